#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fitters"; timeout 900 python -m pytest tests/test_gpu_fitters.py -q -x -s 2>&1 | tail -8
} > gpurun_out/round_s.log 2>&1
tail -30 gpurun_out/round_s.log
