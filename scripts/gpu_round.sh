#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full gpu test suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== time"; timeout 300 python scripts/time_scnet.py 1 8 32 2>&1 | tail -4
echo "=== bench"; timeout 900 python bench.py 2>&1 | tail -2
echo "=== ncu launch list"; RP_SCNET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/scnet_launches_r1h.csv python scripts/prof_scnet.py 8 2>&1 | tail -3
} > gpurun_out/round_k.log 2>&1
tail -30 gpurun_out/round_k.log
