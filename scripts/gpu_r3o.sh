#!/bin/bash
mkdir -p gpurun_out
{
echo "=== resnet tests"; timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_plan.py tests/test_gpu_pipeline_batch.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|rror|max|final" | tail -12
echo "=== resnet timing"; timeout 300 python scripts/time_resnet.py 2>&1 | tail -1
echo "=== resnet launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_resnet_launches_64img.csv python scripts/prof_resnet.py 64 > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/r2_resnet_launches_64img.csv | head -16
} > gpurun_out/round_r3o.log 2>&1
tail -c 3500 gpurun_out/round_r3o.log
