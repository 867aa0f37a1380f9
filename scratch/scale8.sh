#!/bin/bash
# 8-GPU gather variants, quick mode (device-resident step timing only)
run() { # name, env...
  name=$1; shift
  echo "== $name"
  env "$@" QTB200_BENCH_QUICK=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $NG --steps 30 --warmup 5 2>/dev/null | grep quick
}
NG=${NG:-8}
run ce4 QTB200_GATHER=ce
run ce1 QTB200_GATHER=ce QTB200_GATHER_STREAMS=1
run ce7 QTB200_GATHER=ce QTB200_GATHER_STREAMS=7
run nccl QTB200_GATHER=nccl
run sync QTB200_GATHER=sync
