#!/bin/bash
mkdir -p gpurun_out
{
echo "=== solver tests"; timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_batch.py tests/test_gpu_fd_objective.py -q -x 2>&1 | tail -5
echo "=== hard graphs"; timeout 300 python scripts/diag_pipeline.py 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "=== configs[2]/[3]"; timeout 600 python scripts/bench_pipeline.py 2>&1 | tail -5
} > gpurun_out/round_q.log 2>&1
tail -30 gpurun_out/round_q.log
