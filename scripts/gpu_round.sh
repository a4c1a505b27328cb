#!/bin/bash
mkdir -p gpurun_out
{
echo "=== halo tests"; timeout 900 python -m pytest tests/test_gpu_halo.py -q 2>&1 | tail -3
echo "=== layer timings"; timeout 300 python scripts/prof_halo_layers.py 8 2 2>&1 | tail -8
echo "=== time"; timeout 300 python scripts/time_scnet.py 1 8 32 2>&1 | tail -4
} > gpurun_out/round_v.log 2>&1
tail -30 gpurun_out/round_v.log
