#!/bin/bash
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/append_time.py
exit 0
