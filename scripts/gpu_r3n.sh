#!/bin/bash
mkdir -p gpurun_out
{
for m in fused passes; do
echo "=== launch list tc3 $m"; RP_SCNET_TC3=$m RP_SCNET_MODE=tc3 RP_SCNET_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/tc3_${m}_launches.csv python scripts/prof_scnet.py 32 > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/tc3_${m}_launches.csv -v | awk '{print $2, $3, $4, $5, $6, $7, $8}' | grep -v bn_finalize | head -150
done
} > gpurun_out/round_r3n.log 2>&1
tail -c 300 gpurun_out/round_r3n.log
