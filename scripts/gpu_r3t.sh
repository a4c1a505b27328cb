#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_pipeline_batch.py tests/test_gpu_pipeline.py tests/test_gpu_plan.py tests/test_gpu_warp.py tests/test_gpu_resnet.py -m gpu -q -x 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_bench_line_N1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line_N1.json')); print(d['value'], d['e2e']['value'], d['e2e_records']['value'], d['per_pair_p50_ms'], d['clocks']); print({k:(round(v.get('ms',0),2), round(v.get('ms_slowest_call',0),1)) for k,v in d['extra'].items() if isinstance(v,dict)})"
echo "=== launch list of one alternation call (no graph)"; RP_SCNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_alternation_launches.csv python scripts/prof_alternation.py > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/r2_alternation_launches.csv > gpurun_out/r2_alternation_launches_summary.txt; head -4 gpurun_out/r2_alternation_launches_summary.txt
} > gpurun_out/round_r3t.log 2>&1
tail -12 gpurun_out/round_r3t.log
