#!/bin/bash
# two-GPU check of the torchrun path of bench.py (both arms) + the gloo / sharding tests that need real GPUs
mkdir -p gpurun_out
{
echo "=== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 > gpurun_out/bench_line_N2.txt; python - <<'PY'
import json
txt = open('gpurun_out/bench_line_N2.txt').read().strip().splitlines()
try:
    d = json.loads(txt[-1])
    print(d.get('host_affinity'), {k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'], d['parity'], {k: (v.get('pairs_per_s'), v.get('ms')) for k, v in d['extra'].items()})
except Exception as e:
    print("parse failed", e, txt[-3:])
PY
echo "=== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
echo "=== sharding test on 2 GPUs"; timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k sharded 2>&1 | tail -2
} > gpurun_out/round_N2.log 2>&1
cat gpurun_out/round_N2.log
