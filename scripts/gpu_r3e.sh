#!/bin/bash
mkdir -p gpurun_out
{
echo "=== solver tests"; timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_properties.py tests/test_gpu_fitters.py tests/test_gpu_via_completion.py -m gpu -q -x 2>&1 | tail -8
echo "=== wide vs narrow"; timeout 600 python scripts/time_wide.py 2>&1 | tail -30
echo "=== diag"; timeout 600 python scripts/diag_alternation_solve.py 2>&1 | grep -v "^ \[\|^  \[" | tail -30
} > gpurun_out/round_r3e.log 2>&1
tail -c 5000 gpurun_out/round_r3e.log
