#!/bin/bash
mkdir -p gpurun_out
{
echo "=== resnet tests"; timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_scnet.py -q -s 2>&1 | grep "final\|passed\|failed\|rror" | tail -12
echo "=== configs[2]/[3]"; timeout 600 python scripts/bench_pipeline.py 2>&1 | tail -5
echo "=== resnet launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/resnet_launches2.csv python scripts/prof_resnet.py 64 2>&1 | tail -1
} > gpurun_out/round_ad.log 2>&1
tail -20 gpurun_out/round_ad.log
