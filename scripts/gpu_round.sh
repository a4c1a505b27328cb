#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== scnet timing"; timeout 600 python scripts/time_scnet.py 1 8 32 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line_$TAG.json; python - <<PY
import json
d = json.load(open('gpurun_out/bench_line_$TAG.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'per_pair_p50_ms')}, 'e2e', d['e2e']['value'], 'records', d['e2e_records']['value'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac')}, d['parity'])
for k, v in d['extra'].items():
    print(k, {a: b for a, b in v.items() if a not in ('workload', 'rows', 'ncu', 'peak_source')})
print(d.get('cpu_baseline'))
PY
} > gpurun_out/round_$TAG.log 2>&1
tail -40 gpurun_out/round_$TAG.log
