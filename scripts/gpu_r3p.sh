#!/bin/bash
mkdir -p gpurun_out
{
for cap in 192 128 96 64 48; do
echo "=== cap $cap"; RP_PI_FAST_CAP=$cap timeout 600 python scripts/diag_alternation_solve.py 2>&1 | grep -E "solve |accelerated" | sed 's/columns.*//' | cut -c1-220
done
echo "=== solver tests at default"; timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py -m gpu -q -x 2>&1 | tail -2
} > gpurun_out/round_r3p.log 2>&1
tail -c 4000 gpurun_out/round_r3p.log
