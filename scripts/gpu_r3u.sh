#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scnet.py -m gpu -q -x -k "head_subset" > gpurun_out/round_r3u.log 2>&1; tail -5 gpurun_out/round_r3u.log
