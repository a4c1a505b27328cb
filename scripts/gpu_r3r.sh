#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/time_chunks.py > gpurun_out/round_r3r.log 2>&1; tail -12 gpurun_out/round_r3r.log
