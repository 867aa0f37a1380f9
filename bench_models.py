#!/usr/bin/env python
"""Images/sec of the BASELINE network configs 3-5 (synthetic inputs, random-init weights, eval mode).

    python bench_models.py --config alexnet_w4a4|resnet18_t2a8|vgg_w8a8 [--batch B] [--steps K]
    python -m torch.distributed.run --nproc-per-node G ... bench_models.py --config resnet18_t2a8 --batch 2048

Under torchrun the global batch is sharded contiguously over the ranks (weights replicated and packed per rank) and the
fp32 logits are all-gathered (the only collective).  Prints one JSON line per run on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (builder, kwargs, input shape, default global batch, GMAC per image of the quantized layers)
    "alexnet_w4a4": ("alexnet_dorefa", dict(bit_width=4), (3, 224, 224), 256, 4.935),
    "resnet18_t2a8": ("resnet18_ternary", dict(act_bits=8), (3, 224, 224), 512, 1.814),
    "vgg_w8a8": ("vgg_dorefa", dict(bit_width=8), (3, 32, 32), 4096, 0.158),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="GLOBAL batch (sharded over ranks)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="fused", choices=["plain", "fused"],
                    help="fused: fusion.fuse_inference (BN+clamp+quantizer in one pass) + code_only_activations()")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import contextlib
    import pytorch_quantize_impls_b200 as Q
    from pytorch_quantize_impls_b200 import _lib, nets, sharding
    builder, kw, shape, dflt, gmac = CONFIGS[args.config]
    B = args.batch or dflt
    torch.manual_seed(1234)
    net = getattr(nets, builder)(**kw).to(dev).eval()
    if args.mode == "fused":
        net = Q.fuse_inference(net)
    mode_ctx = Q.code_only_activations if args.mode == "fused" else contextlib.nullcontext
    g = torch.Generator().manual_seed(1234)
    lo, hi = sharding.shard_bounds(B, rank, world)
    x = torch.rand(hi - lo, *shape, generator=g).to(dev)

    def step():
        with mode_ctx():
            y = net(x)
        return sharding.gather_logits(y, batch=B)

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        print(json.dumps({
            "config": args.config, "mode": args.mode, "n_gpus": world, "global_batch": B, "ms_per_step": round(ms, 3),
            "images_per_sec": round(B / (ms * 1e-3), 1), "quantized_gops": round(2 * gmac * B / (ms * 1e-3), 1),
            "qt_kernel_launches_per_step": _lib.launch_count() // args.steps, "logits_shape": list(y.shape),
            "finite": bool(torch.isfinite(y).all().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
