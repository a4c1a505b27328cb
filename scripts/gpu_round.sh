#!/bin/bash
mkdir -p gpurun_out
{
echo "=== keypoint tests"; timeout 600 python -m pytest tests/test_gpu_keypoint.py -q -x 2>&1 | tail -5
echo "=== timing"; timeout 300 python scripts/time_keypoint.py 2>&1 | tail -6
} > gpurun_out/round_m.log 2>&1
tail -30 gpurun_out/round_m.log
