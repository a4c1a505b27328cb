#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_pipeline_batch.py tests/test_gpu_warp.py tests/test_gpu_pipeline.py -q 2>&1 | tail -3
echo "=== configs"; timeout 600 python scripts/bench_pipeline.py 2>&1 | tail -4
} > gpurun_out/round_aq.log 2>&1
tail -12 gpurun_out/round_aq.log
