#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc3 scnet test (fused)"; timeout 900 python -m pytest tests/test_gpu_scnet.py -m gpu -q -x -s -k "split_precision" 2>&1 | grep -E "tc3 final|passed|failed|Error|error|conv9:act|deconv2f:act" | tail -12
echo "=== tc3 via completion (fused)"; timeout 900 python -m pytest tests/test_gpu_via_completion.py -m gpu -q -x -s -k "tc3" 2>&1 | grep -E "passed|failed|Error|error|assert" | tail -8; python - <<'PY'
import json
d = json.load(open('gpurun_out/via_completion_parity.json'))
for k, rows in d.items():
    if 'tc3' in k:
        print(k, [(round(r['net_f'], 6), round(r['topk_rows_equal'], 4), float('%.3g' % r['dT'])) for r in rows])
PY
echo "=== resnet tc3"; RP_SCNET_MODE=tc3 timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -x 2>&1 | tail -2
echo "=== timing"; for m in fused passes; do RP_SCNET_TC3=$m RP_SCNET_MODE=tc3 timeout 600 python scripts/time_scnet.py 1 32 2>&1 | tail -2; done
RP_SCNET_MODE=tc timeout 600 python scripts/time_scnet.py 32 2>&1 | tail -1
} > gpurun_out/round_r3m.log 2>&1
tail -c 5000 gpurun_out/round_r3m.log
