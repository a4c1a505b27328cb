#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc3 scnet test"; timeout 900 python -m pytest tests/test_gpu_scnet.py -m gpu -q -x -s -k "split_precision" 2>&1 | grep -v "^$" | tail -70
echo "=== tc3 via completion"; timeout 900 python -m pytest tests/test_gpu_via_completion.py -m gpu -q -x -s -k "tc3" 2>&1 | grep -v "^$" | tail -30
echo "=== timing"; for m in tc tc3 fp32; do RP_SCNET_MODE=$m timeout 600 python scripts/time_scnet.py 1 32 2>&1 | tail -2; done
echo "=== halo tests"; timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_scnet.py tests/test_gpu_resnet.py -m gpu -q 2>&1 | tail -4
} > gpurun_out/round_r3k.log 2>&1
tail -c 9000 gpurun_out/round_r3k.log
