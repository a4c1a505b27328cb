#!/bin/bash
# One GPU visit: tests, bench.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "=== via-completion numbers"; python - <<'PY'
import json
try:
    R = json.load(open('gpurun_out/via_completion_parity.json'))
    for k, rows in R.items():
        if isinstance(rows, list):
            for r in rows:
                print(k, {a: (round(b, 6) if isinstance(b, float) else b) for a, b in r.items()})
        else:
            print(k, rows)
except Exception as e:
    print("no report", e)
PY
echo "=== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 2>&1 | tail -3
} > gpurun_out/round_$TAG.log 2>&1
tail -5 gpurun_out/round_$TAG.log
