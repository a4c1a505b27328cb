#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== net tests"; timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_plan.py tests/test_gpu_via_completion.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_batch.py -m gpu -q 2>&1 | tail -4
echo "=== scnet timing default"; timeout 600 python scripts/time_scnet.py 1 8 32 2>&1 | tail -3
} > gpurun_out/round_$TAG.log 2>&1
cat gpurun_out/round_$TAG.log
