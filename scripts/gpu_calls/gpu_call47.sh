#!/bin/bash
python scripts/profile_cfg3_ops.py amp 2>&1 | tail -34
exit 0
