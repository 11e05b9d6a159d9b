#!/bin/bash
for v in "" _q8b2 _q4b2 _q4b1; do
  echo "== libevavos_sm100$v"; EVAVOS_LIB=evavos_b200/libevavos_sm100$v.so python scripts/dense_time.py 2>&1 | tail -3
done
exit 0
