#!/bin/bash
mkdir -p gpurun_out
{
echo "=== net tests"; timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_plan.py tests/test_gpu_via_completion.py -m gpu -q -x 2>&1 | tail -4
echo "=== timing"; for m in tc tc3; do RP_SCNET_MODE=$m timeout 600 python scripts/time_scnet.py 1 32 2>&1 | tail -2; done
echo "=== bench"; timeout 1500 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_bench_line_N1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line_N1.json')); print(d['value'], d['e2e']['value'], d['per_pair_p50_ms'], d['clocks']); print({k:(round(v.get('ms',0),2)) for k,v in d['extra'].items() if isinstance(v,dict)})"
echo "=== scnet launch list P=32"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_scnet_launches_P32.csv python scripts/prof_scnet.py 32 > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/r2_scnet_launches_P32.csv | head -14
} > gpurun_out/round_r3l.log 2>&1
tail -c 4000 gpurun_out/round_r3l.log
