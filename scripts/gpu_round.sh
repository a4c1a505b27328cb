#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== solver tests (default lib)"; timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py tests/test_gpu_properties.py -m gpu -q -x 2>&1 | tail -3
for L in "" relativepose_b200/build/librp_mb5.so; do
echo "=== bench lib='$L'"; RP_B200_LIB=${L:+$PWD/$L} timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','per_pair_p50_ms')}, d['e2e']['value'], d['parity'])"
done
echo "=== solver tests (mb5 lib)"; RP_B200_LIB=$PWD/relativepose_b200/build/librp_mb5.so timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fitters.py -m gpu -q -x 2>&1 | tail -3
} > gpurun_out/round_$TAG.log 2>&1
tail -30 gpurun_out/round_$TAG.log
