#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_properties.py tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_plan.py tests/test_gpu_pipeline_batch.py -m gpu -q -x 2>&1 | tail -4
echo "=== wide vs narrow"; timeout 600 python scripts/time_wide.py 2>&1 | grep -E "B     1:|B    32|B   148|single" 
echo "=== scnet"; timeout 600 python scripts/time_scnet.py 1 32 2>&1 | tail -2
echo "=== resnet"; timeout 600 python scripts/time_resnet.py 2>&1 | tail -1
echo "=== ncu launch list of one alternation call (no graph)"; RP_SCNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_alternation_launches.csv python scripts/prof_alternation.py > gpurun_out/prof_alt.out 2>&1; tail -3 gpurun_out/prof_alt.out; grep -i "error" gpurun_out/r3_alternation_launches.csv | head -3; python scripts/ncu_launch_table.py gpurun_out/r3_alternation_launches.csv | head -40
} > gpurun_out/round_r3i.log 2>&1
tail -c 7000 gpurun_out/round_r3i.log
