#!/bin/bash
# Evidence run without the ncu --set full captures (gpurun merges at most 64 MiB back): tests, smoke, bench line, launch lists.
TAG=${1:-r2}
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 1500 python bench.py 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_line_N1.json; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_line_N1.json')); print(d['value'], d['e2e']['value'], d['e2e_records']['value'], d['per_pair_p50_ms'], d['clocks']); print({k:(round(v.get('ms',0),2)) for k,v in d['extra'].items() if isinstance(v,dict)})"
echo "=== launch list of one alternation call (no graph)"; RP_SCNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_alternation_launches.csv python scripts/prof_alternation.py > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/${TAG}_alternation_launches.csv > gpurun_out/${TAG}_alternation_launches_summary.txt; head -8 gpurun_out/${TAG}_alternation_launches_summary.txt
echo "=== dense pairs"; timeout 600 python scripts/diag_alternation_solve.py 2>&1 | grep -E "solve |accelerated" | sed 's/columns.*//' | cut -c1-200
} > gpurun_out/round_final_light_$TAG.log 2>&1
tail -40 gpurun_out/round_final_light_$TAG.log
