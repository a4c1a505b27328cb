#!/bin/bash
mkdir -p gpurun_out
{
echo "=== all gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== time scnet"; timeout 300 python scripts/time_scnet.py 1 8 32 2>&1 | tail -3
echo "=== configs"; timeout 600 python scripts/bench_pipeline.py 2>&1 | tail -5
} > gpurun_out/round_ao.log 2>&1
tail -12 gpurun_out/round_ao.log
