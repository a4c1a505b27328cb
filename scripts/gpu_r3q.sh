#!/bin/bash
mkdir -p gpurun_out
{
for cs in "192 48" "96 48" "64 48" "48 48" "48 24" "48 16" "32 16" "64 24"; do
set -- $cs
echo "=== cap $1 switch $2"; RP_PI_FAST_CAP=$1 RP_PI_SWITCH=$2 timeout 600 python scripts/diag_alternation_solve.py 2>&1 | grep -E "solve " | cut -c1-80
RP_PI_FAST_CAP=$1 RP_PI_SWITCH=$2 timeout 300 python scripts/bench_its.py 2>&1 | tail -1
done
} > gpurun_out/round_r3q.log 2>&1
tail -c 5000 gpurun_out/round_r3q.log
