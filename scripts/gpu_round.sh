#!/bin/bash
# One GPU-box visit: halo-layer parity, the whole GPU suite, SCNet timing in the four (halo, storage) configurations.
mkdir -p gpurun_out
{
echo "=== halo layer tests"; timeout 600 python -m pytest tests/test_gpu_halo.py -q -x -s 2>&1 | tail -40
echo "=== scnet/resnet tests"; timeout 900 python -m pytest tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_tc.py -q -s 2>&1 | grep -v "^conv\|^deconv\|^resnet18" | tail -40
for cfg in "0 fp32" "1 fp32" "0 bf16" "1 bf16"; do set -- $cfg; echo "=== time halo=$1 act=$2"; RP_SCNET_HALO=$1 RP_SCNET_ACT=$2 timeout 300 python scripts/time_scnet.py 1 8 2>&1 | tail -4; done
} > gpurun_out/round_a.log 2>&1
tail -60 gpurun_out/round_a.log
