#!/bin/bash
# order-dependence of test_fused_step_matches_oracle_random: which earlier test makes it fail?
T=test_fused_step_matches_oracle_random
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -E "AssertionError:|passed|failed" | head -3
names=$(grep -n "^def test_" tests/test_gpu_parity.py | awk -F'[ (]' '{print $2}' | awk "/$T/{exit} {print}")
for n in $names; do
  r=$(timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "$n or $T" 2>&1 | grep -E "AssertionError:|passed|failed" | tr '\n' ' ')
  echo "$n :: $r"
done
