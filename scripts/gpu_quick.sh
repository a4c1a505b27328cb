#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_records']['value'], d['per_pair_p50_ms'], d['parity']['max_T_frobenius_err_vs_oracle'])" > gpurun_out/round_quick.log 2>&1; cat gpurun_out/round_quick.log
