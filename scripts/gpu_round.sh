#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== solver tests"; timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_plan.py tests/test_gpu_properties.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_batch.py tests/test_gpu_fd_objective.py -m gpu -q -x 2>&1 | tail -8
echo "=== bench (no extras)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','per_pair_p50_ms')}, d['e2e']['value'], d['e2e_records']['value'], d['parity'])"
echo "=== diag"; timeout 600 python scripts/diag_alternation_solve.py 2>&1 | grep -E "solve|stop after|mean"
echo "=== alternation"; timeout 600 python scripts/time_alternation.py 2>&1 | tail -6
echo "=== scnet launch list P=32"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/scnet_launches_P32_$TAG.csv python scripts/prof_scnet.py 32 > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/scnet_launches_P32_$TAG.csv | head -12
} > gpurun_out/round_$TAG.log 2>&1
tail -60 gpurun_out/round_$TAG.log
