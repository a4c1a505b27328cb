#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_bench_line_N1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line_N1.json')); print(d['value'], d['e2e']['value'], d['e2e_records']['value'], d['per_pair_p50_ms'], d['clocks']); print({k:(round(v.get('ms',0),2), round(v.get('ms_slowest_call',0),1)) for k,v in d['extra'].items() if isinstance(v,dict)})"
