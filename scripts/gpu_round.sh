#!/bin/bash
mkdir -p gpurun_out
{
echo "=== resnet tests"; timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_plan.py -q 2>&1 | tail -4
echo "=== configs"; PAIRS=32 timeout 600 python scripts/bench_pipeline.py 2>&1 | head -1
echo "=== resnet launches"; RP_SCNET_PLAN=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/resnet_launches3.csv python scripts/prof_resnet.py 64 2>&1 | tail -1
} > gpurun_out/round_ai.log 2>&1
tail -30 gpurun_out/round_ai.log
