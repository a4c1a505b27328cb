#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_properties.py tests/test_gpu_via_completion.py -m gpu -q -x 2>&1 | tail -4
echo "=== wide vs narrow"; timeout 600 python scripts/time_wide.py 2>&1 | grep -E "B     1:|B    32|B   148|single" 
echo "=== phase clocks"; timeout 300 python scripts/phase_clk.py 2>&1 | tail -6
} > gpurun_out/round_r3j.log 2>&1
tail -c 5000 gpurun_out/round_r3j.log
