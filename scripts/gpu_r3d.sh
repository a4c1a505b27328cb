#!/bin/bash
mkdir -p gpurun_out
{
echo "=== variant tests"; timeout 900 python -m pytest tests/test_gpu_scnet.py -m gpu -q -x -s -k variants 2>&1 | grep -v "^$" | tail -15
echo "=== diag"; timeout 600 python scripts/diag_alternation_solve.py 2>&1 | tail -90
} > gpurun_out/round_r3d.log 2>&1
tail -c 6000 gpurun_out/round_r3d.log
