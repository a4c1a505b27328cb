#!/bin/bash
mkdir -p gpurun_out
{
echo "=== plan tests"; timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_solver.py -q 2>&1 | tail -5
} > gpurun_out/round_ah.log 2>&1
tail -30 gpurun_out/round_ah.log
