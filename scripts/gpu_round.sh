#!/bin/bash
mkdir -p gpurun_out
{
echo "=== halo layer tests"; timeout 600 python -m pytest tests/test_gpu_halo.py -q 2>&1 | grep -v "^$" | tail -15
echo "=== scnet/resnet/pipeline tests"; timeout 900 python -m pytest tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_pipeline.py -q -s 2>&1 | grep "final\|passed\|failed\|Error" | tail -20
echo "=== layer timings"; timeout 300 python scripts/prof_halo_layers.py 8 2 2>&1 | tail -8
echo "=== time"; timeout 300 python scripts/time_scnet.py 1 8 2>&1 | tail -3
echo "=== ncu launch list"; RP_SCNET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/scnet_launches_r1g.csv python scripts/prof_scnet.py 8 2>&1 | tail -5
} > gpurun_out/round_i.log 2>&1
tail -50 gpurun_out/round_i.log
