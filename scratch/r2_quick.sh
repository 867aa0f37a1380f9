#!/bin/bash
# quick check after a kernel change: conv-chain parity tests + the three CNN bench lines + headline
set -u
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_convchain.py tests/test_gpu_requant.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -4
for c in alexnet_w4a4 resnet18_t2a8 vgg_w8a8; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 2> $O/q_${c}.err | tail -1 > $O/q_${c}.json
  python - <<PY
import json; d=json.load(open("$O/q_${c}.json")); print("$c", d["value"], d["unit"], d["ms_per_step"], d.get("gathered_logits_equal_single_gpu_run"), d.get("parity"))
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
QTB200_BENCH_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 75 --csv --log-file $O/q_launches_resnet.csv python bench.py --config resnet18_t2a8 --steps 1 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/q_launches_resnet.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:75]:
    print(r[ki][:70], r[vi])
PY
