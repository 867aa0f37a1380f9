#!/usr/bin/env python
"""Drop-in use of the package on the BASELINE configs[1] network (needs one B200; `torchrun --nproc-per-node G` shards the batch).

    python examples/xnor_mlp_inference.py [--batch 8192] [--steps 20]

1. a model written against the reference's API (`QuantTorch.layers.LinearXNOR`, `QuantTorch.functions.nnQuantXnor`) -- only
   the import line differs;
2. `fuse_inference` + `code_only_activations()` + `prefetch_operands`: hidden activations stay low-bit operands;
3. `GraphedModule`: the whole forward replayed as one CUDA graph;
4. `save_packed` / `load_packed`: the 2-bit + alpha[k] weights on disk, no fp32 master copy after loading;
5. under torchrun: contiguous batch shards, logits gathered by the copy engines beside the next step.
"""
import argparse
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pytorch_quantize_impls_b200 as QuantTorch          # instead of `import QuantTorch`
from pytorch_quantize_impls_b200 import sharding
from pytorch_quantize_impls_b200.pipeline import GraphedModule


def build(dims=(4096, 4096, 4096, 1000)):
    mods = []
    for i in range(len(dims) - 1):
        mods += [QuantTorch.functions.nnQuantXnor(1), QuantTorch.layers.LinearXNOR(dims[i], dims[i + 1])]
    return torch.nn.Sequential(*mods)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192, help="rows per GPU")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    net = build().to(dev).eval()                          # eval(): weights packed once (2 bit planes + alpha[k])
    x = torch.randn(args.batch, 4096, device=dev)
    with torch.no_grad():
        y_dropin = net(x)                                 # every module returns what the reference's module returns

        # packed checkpoint: 2.1 MB per 4096 x 4096 layer instead of 67 MB
        path = os.path.join(tempfile.gettempdir(), "xnor_mlp_rank%d.qtb" % local)
        QuantTorch.save_packed(net, path)
        served = QuantTorch.load_packed(build().to(dev), path)            # fresh instance, packed-only layers
        served = QuantTorch.prefetch_operands(QuantTorch.fuse_inference(served))

        def forward(t):
            with QuantTorch.code_only_activations():
                return served(t)
        step = GraphedModule(forward, x)
        gather = sharding.PipelinedGather()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = step()
            logits, done = gather.submit(y)               # no-op on one GPU
            if done is not None:
                step.wait_for(done)
        gather.drain()
        e1.record()
        torch.cuda.synchronize()
    if int(os.environ.get("RANK", "0")) == 0:
        ms = e0.elapsed_time(e1) / args.steps
        macs = 4096 * 4096 * 2 + 4096 * 1000
        print("max |fused - drop-in| / max |drop-in| = %.2e" % float((y - y_dropin).abs().max() / y_dropin.abs().max()))
        print("%.3f ms/step, %.0f TOPS, gathered logits %s" % (ms, 2 * args.batch * world * macs / ms / 1e9, tuple(logits.shape)))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
