#!/bin/bash
# repeat the whole parity file; print every failure line
for i in 1 2 3 4 5 6; do
  timeout 200 python -m pytest tests/test_gpu_parity.py -q 2>&1 | grep -E "^E +AssertionError|passed|failed" | tr '\n' ' '; echo
done
echo "== GAITK_FORK=0"
for i in 1 2 3 4; do
  GAITK_FORK=0 timeout 200 python -m pytest tests/test_gpu_parity.py -q 2>&1 | grep -E "^E +AssertionError|passed|failed" | tr '\n' ' '; echo
done
