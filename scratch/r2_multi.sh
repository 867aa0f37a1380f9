#!/bin/bash
# multi-GPU evidence: NG ranks of one node under torchrun; usage: NG=4 bash scratch/r2_multi.sh "bench resnet18_t2a8 vgg_w8a8"
set -u
O=gpurun_out
NG=${NG:-2}
tr() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
for w in $1; do
  if [ $w = bench ]; then
    tr bench.py --gpus $NG --steps 20 --warmup 5 2> $O/r2_bench_${NG}gpu.err | tail -1 > $O/r2_bench_${NG}gpu.json
    cut -c1-160 $O/r2_bench_${NG}gpu.json
  else
    tr bench.py --gpus $NG --config $w --steps 20 --warmup 3 2> $O/r2_${w}_${NG}gpu.err | tail -1 > $O/r2_${w}_${NG}gpu.json
    cut -c1-160 $O/r2_${w}_${NG}gpu.json
  fi
done
