#!/bin/bash
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench (no extras)"; timeout 600 python bench.py --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_records']['value'], d['per_pair_p50_ms'], d['parity'])"
} > gpurun_out/round_sanity.log 2>&1
tail -12 gpurun_out/round_sanity.log
