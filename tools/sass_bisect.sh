#!/bin/bash
# Runs the load-time self-check of the re-scheduled SASS one kernel variant per process (a faulting variant takes its
# CUDA context with it): prints, per variant, the mask of variants that differ from nvcc's schedule or the failure.
cd "$(dirname "$0")/.."
for v in $(seq 0 22); do
  printf "variant %2d: " $v
  BQ_B200_SASS_VERIFY_ONLY=$v timeout 120 python -c "
from bosonic_qiskit_b200 import get_engine
e = get_engine(0)
print('bad mask %#x, checked %d' % e.sass_self_check())" 2>&1 | tail -1
done
