#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
RP_SCNET_HALO_FLAGS=34 timeout 300 python scripts/prof_halo_layers.py 32 3 2>&1 | tail -16
} > gpurun_out/round_$TAG.log 2>&1
cat gpurun_out/round_$TAG.log
