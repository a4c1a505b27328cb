#!/bin/bash
mkdir -p gpurun_out
{
echo "=== all gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_line_N1.json; python -c "import json; d=json.load(open('gpurun_out/bench_line_N1.json')); print(d['value'], d['per_pair_p50_ms'], d['e2e']['value'], d['cpu_baseline'], d['clocks'])"
echo "=== time scnet"; timeout 300 python scripts/time_scnet.py 1 8 32 2>&1 | tail -3
} > gpurun_out/round_final3.log 2>&1
tail -12 gpurun_out/round_final3.log
