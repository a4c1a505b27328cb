#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== solver tests"; timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_plan.py tests/test_gpu_properties.py tests/test_gpu_via_completion.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_batch.py tests/test_gpu_fd_objective.py -m gpu -q -x 2>&1 | tail -4
echo "=== bench (no extras)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','per_pair_p50_ms')}, d['e2e']['value'], d['e2e_records']['value'], d['parity'])"
} > gpurun_out/round_$TAG.log 2>&1
tail -30 gpurun_out/round_$TAG.log
