#!/bin/bash
mkdir -p gpurun_out
{
echo "=== solver tests"; timeout 900 python -m pytest tests/test_gpu_solver.py -q -s -k "topk_clamped or determinism or ragged" 2>&1 | tail -6
echo "=== all gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
} > gpurun_out/round_ab.log 2>&1
tail -12 gpurun_out/round_ab.log
