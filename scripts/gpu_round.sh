#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "=== net tests"; timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_scnet.py tests/test_gpu_resnet.py tests/test_gpu_plan.py tests/test_gpu_tc.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_batch.py -m gpu -q 2>&1 | tail -6
echo "=== resnet timing"; timeout 300 python - <<'PY'
import sys, types, torch
sys.path.insert(0, '.')
import bench
from relativepose_b200.model.mymodel import Resnet18_8s
torch.manual_seed(0)
net = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).cuda()
x = torch.randn(64, 7, 160, 640, device='cuda')
print("Resnet18_8s 64 images: %.3f ms" % bench.device_time_ms(torch, lambda: net(x), 10, 5))
PY
} > gpurun_out/round_$TAG.log 2>&1
cat gpurun_out/round_$TAG.log
