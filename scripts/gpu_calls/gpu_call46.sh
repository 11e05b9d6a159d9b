#!/bin/bash
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python scripts/profile_cfg3.py cl graphs 2>&1 | head -1
python scripts/profile_cfg3.py amp graphs 2>&1 | head -1
exit 0
