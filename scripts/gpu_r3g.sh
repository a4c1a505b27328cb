#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pipeline tests"; timeout 900 python -m pytest tests/test_gpu_pipeline_batch.py tests/test_gpu_pipeline.py tests/test_gpu_warp.py -m gpu -q -x 2>&1 | tail -8
echo "=== h2d / chunks"; timeout 600 python scripts/time_h2d.py 2>&1 | tail -14
echo "=== c4 N=1"; timeout 900 python - <<'PY' 2>&1 | tail -5
import sys, time
sys.argv = ['bench.py']
import torch, bench
dev = torch.device('cuda:0')
r = bench.run_c4(torch, dev, 0, 1, 256, lambda: torch.cuda.synchronize(), lambda x: x)
print({k: r[k] for k in ('ms', 'pairs_per_s')})
r = bench.run_c4(torch, dev, 0, 1, 32, lambda: torch.cuda.synchronize(), lambda x: x)
print('32 pairs:', {k: r[k] for k in ('ms', 'pairs_per_s')})
PY
} > gpurun_out/round_r3g.log 2>&1
tail -c 5000 gpurun_out/round_r3g.log
