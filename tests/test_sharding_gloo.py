"""CPU, world_size 2 over gloo: the N>1 host logic (contiguous batch shards + one logits all-gather) reproduces the
single-process result.  The per-rank compute stand-in is the oracle (tests may use it; the product path is CUDA only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import quanttorch_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_quantize_impls_b200 import sharding
    torch.manual_seed(5)
    x = torch.randn(batch, 64)
    w = torch.randn(10, 64) * 0.3
    b = torch.rand(10)

    class Net(torch.nn.Module):
        def forward(self, t):
            return O.linear_xnor(O.xnor_act(t, 1), w, b)
    y = sharding.ShardedInference(Net())(x)
    if batch % world == 0:
        # the pipelined gather (communication stream on CUDA, synchronous on CPU tensors) returns the same rows
        pg = sharding.PipelinedGather(depth=2)
        for _ in range(3):          # rotate through the buffers
            g, ev = pg.submit(Net()(sharding.shard_batch(x)))
        pg.drain()
        assert ev is None and torch.equal(g, y)
    if rank == 0:
        q.put(y)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [16, 17])
def test_sharded_forward_matches_single_process(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    y = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    torch.manual_seed(5)
    x = torch.randn(batch, 64); w = torch.randn(10, 64) * 0.3; b = torch.rand(10)
    ref = O.linear_xnor(O.xnor_act(x, 1), w, b)
    assert torch.equal(y, ref)


def test_shard_bounds_cover_batch():
    from pytorch_quantize_impls_b200.sharding import shard_bounds
    for B in (0, 1, 7, 8, 2048, 2049):
        for G in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
