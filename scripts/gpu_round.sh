#!/bin/bash
mkdir -p gpurun_out
{
echo "=== halo layer tests"; timeout 600 python -m pytest tests/test_gpu_halo.py -q -s 2>&1 | grep -v "^$" | tail -45
echo "=== scnet/resnet/pipeline/warp/fd tests"; timeout 900 python -m pytest tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_pipeline.py tests/test_gpu_warp.py tests/test_gpu_fd_objective.py -q -s 2>&1 | grep -v "^conv\|^deconv\|^resnet18\|^\.conv\|^$\|bit-equal" | tail -40
for cfg in "1 1" "1 14"; do set -- $cfg; echo "=== time halo=$1 halo_min=$2 act=bf16"; RP_SCNET_HALO=$1 RP_SCNET_HALO_MIN=$2 timeout 300 python scripts/time_scnet.py 1 8 2>&1 | tail -3; done
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/scnet_launches_r1d.csv python scripts/prof_scnet.py 8 2>&1 | tail -3
} > gpurun_out/round_d.log 2>&1
tail -5 gpurun_out/round_d.log
