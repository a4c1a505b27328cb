#!/bin/bash
mkdir -p gpurun_out
{
for wm in 148 296; do
echo "=== wide_max $wm"; RP_SOLVER_WIDE_MAX=$wm timeout 600 python - <<'PY' 2>&1 | tail -3
import sys, os
sys.argv = ['bench.py']
import torch, bench
dev = torch.device('cuda:0')
r = bench.run_c4(torch, dev, 0, 1, 256, lambda: torch.cuda.synchronize(), lambda x: x)
print('c4 256 pairs:', {k: round(r[k], 1) for k in ('ms', 'pairs_per_s')})
PY
RP_SOLVER_WIDE_MAX=$wm timeout 300 python scripts/time_wide.py 2>&1 | grep -E "B   222|B   296"
done
} > gpurun_out/round_r3s.log 2>&1
tail -c 1500 gpurun_out/round_r3s.log
