#!/bin/bash
mkdir -p gpurun_out
{
echo "=== bench N=2 (torchrun)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line_N2.json; cut -c1-400 gpurun_out/bench_line_N2.json
echo "=== reference arm N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "=== bench N=1"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_line_N1.json; cut -c1-300 gpurun_out/bench_line_N1.json
} > gpurun_out/round_final2.log 2>&1
tail -12 gpurun_out/round_final2.log
