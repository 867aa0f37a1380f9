#!/bin/bash
# cheapest check after a helper-kernel change: conv-chain tests + ResNet / AlexNet lines
set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_convchain.py -m gpu -q -x 2>&1 | tail -1
for c in resnet18_t2a8 alexnet_w4a4; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 3 2> $O/q_${c}.err | tail -1 > $O/q_${c}.json
  python - <<PY
import json; d=json.load(open("$O/q_${c}.json")); print("$c", d["value"], d["unit"], d["ms_per_step"], d["parity"]["rel_err_vs_oracle"], d["parity"]["graph_replay_equals_eager"])
PY
done
