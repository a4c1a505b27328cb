#!/bin/bash
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r3_bench_line_N1.json 2> gpurun_out/r3_bench_err.log; tail -c 600 gpurun_out/r3_bench_err.log; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3_bench_line_N1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'per_pair_p50_ms')}, 'e2e', d['e2e']['value'], 'rec', d['e2e_records']['value'], d['parity'])
print({k: (round(v.get('pairs_per_s', 0) or 0), round(v.get('ms', 0) or 0, 2)) for k, v in d['extra'].items() if isinstance(v, dict)})
print(d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
echo "=== ncu launch list of one alternation call"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_alternation_launches.csv python scripts/prof_alternation.py > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/r3_alternation_launches.csv | head -30
} > gpurun_out/round_r3h.log 2>&1
tail -c 7000 gpurun_out/round_r3h.log
