#!/bin/bash
# conv-chain parity tests + ResNet / AlexNet / VGG lines + forward kernel times of ResNet
set -u
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_convchain.py tests/test_gpu_requant.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -2
for c in alexnet_w4a4 resnet18_t2a8 vgg_w8a8; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 2> $O/q_${c}.err | tail -1 > $O/q_${c}.json
  python - <<PY
import json; d=json.load(open("$O/q_${c}.json")); print("$c", d["value"], d["unit"], d["ms_per_step"], d["parity"]["rel_err_vs_oracle"], d["parity"]["graph_replay_equals_eager"])
PY
done
bash scratch/r2_launches.sh resnet18_t2a8 0 400 | grep -v "weight_\|Fill\|^$" | tail -42 | awk '{print $NF, $0}' | cut -c1-110 | sort -k2 | uniq -c -f1 | sort -k2 -n -r | head -12
