#!/bin/bash
# end-of-round refresh: smoke, GPU suite, the four 1-GPU lines
set -u
O=gpurun_out
for i in 1 2 3; do timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200; done > $O/r2_smoke.log
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_gpu.log
QTB200_CTA_GROUP=2 timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_gpu_cta_group2.log
timeout 400 python bench.py --steps 20 --warmup 5 2> $O/r2_bench_1gpu.err | tail -1 > $O/r2_bench_1gpu.json
for c in alexnet_w4a4 resnet18_t2a8 vgg_w8a8; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 2> $O/r2_${c}_1gpu.err | tail -1 > $O/r2_${c}_1gpu.json
done
tail -1 $O/r2_smoke.log | cut -c1-80; tail -1 $O/r2_pytest_gpu.log; tail -1 $O/r2_pytest_gpu_cta_group2.log
for f in r2_bench_1gpu r2_alexnet_w4a4_1gpu r2_resnet18_t2a8_1gpu r2_vgg_w8a8_1gpu; do cut -c1-140 $O/$f.json; done
