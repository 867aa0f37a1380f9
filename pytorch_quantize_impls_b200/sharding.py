"""Batch-sharded data-parallel inference.

The path shards naturally: every sample is independent through every layer, so B is cut into G contiguous
shards (one process per GPU, weights replicated and packed per rank) and the ONLY collective is one all-gather
of the fp32 logits.  No data-path collective exists inside the layers.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous [lo, hi) rows of rank `rank`; the first batch % world ranks take one extra row."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world=None):
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(y_local, batch=None):
    """All-gather per-rank logits [B_r, C] into [B, C] on every rank (ragged shards are padded to the largest)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return y_local
    world = dist.get_world_size()
    if batch is None or batch % world == 0:
        out = torch.empty((world * y_local.shape[0],) + tuple(y_local.shape[1:]), dtype=y_local.dtype,
                          device=y_local.device)
        dist.all_gather_into_tensor(out, y_local.contiguous())
        return out
    rows = [shard_bounds(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in rows)
    pad = torch.zeros((mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[:y_local.shape[0]] = y_local
    out = torch.empty((world * mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(rows)], 0)


class ShardedInference(torch.nn.Module):
    """Wrap a replicated network: forward(x_full) runs this rank's shard and returns the gathered logits."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return gather_logits(self.net(shard_batch(x)), batch=x.shape[0])


class PipelinedGather:
    """The logits all-gather of step i on a dedicated stream, overlapped with the kernels of step i+1.

    At 8 GPUs every rank receives 7 x [B/G, classes] fp32 per step (229 MB for the 8192 x 1000 logits of BASELINE
    configs[1]): ~0.26 ms at NVLink-5 line rate, half of the step's compute time.  Steps are independent (batch data
    parallelism), so the collective of one step runs beside the kernels of the next one.

    mode "ce" (default on CUDA): the gather is done by the COPY ENGINES over NVLink through symmetric memory
        (torch.distributed._symmetric_memory).  PUSH form: the gathered buffers themselves are peer-mapped; every rank writes
        its shard straight from the logits tensor into rows [rank*B, (rank+1)*B) of every peer's buffer (posted NVLink
        writes, spread over up to 4 streams), then ONE signal-pad barrier tells everybody that all shards have landed.
    mode "push": the same protocol with the G copies done by ONE small kernel (qt_peer_push: the shard is read once and stored
        to all peers with 16-byte stores from ~32 CTAs that hold no shared memory, so they run beside the tcgen05 CTAs).
    mode "ce_pull": every rank stages its logits in a peer-mapped buffer, barrier, each rank pulls the G-1 remote shards,
        a second barrier releases the stage (one more copy and one more barrier per step than the push form).
        No SM is taken from the compute kernels: the tcgen05 GEMMs are persistent one-CTA-per-SM kernels with ~225 KB of
        shared memory, an NCCL kernel cannot share an SM with them, and running one beside them stalls both (measured:
        2.5 ms/step instead of 0.5 at 2 GPUs).
    mode "nccl": ONE all_gather_into_tensor per step on the communication stream.
    mode "sync": the NCCL collective on the compute stream (no overlap).
    On CPU tensors (gloo tests) it degrades to the synchronous collective."""

    def __init__(self, depth=2, mode="ce", pull_streams=None):
        if mode not in ("ce", "push", "ce_pull", "nccl", "sync"):
            raise ValueError("mode must be 'ce', 'push', 'ce_pull', 'nccl' or 'sync'")
        self.depth, self.mode = depth, mode
        # mode "ce": the G-1 peer pulls of one step are spread over this many streams so that several copy engines
        # (and NVLink ports) work at once; 1 = one pull after the other
        self.pull_streams = pull_streams
        self._pull = []
        self._outs = [None] * depth
        self._done = [None] * depth
        self._stage = [None] * depth       # symmetric staging buffers + their rendezvous handles (mode "ce")
        self._i = 0
        self._comm = None

    def _out(self, slot, y, world):
        shape = (world * y.shape[0],) + tuple(y.shape[1:])
        o = self._outs[slot]
        if o is None or tuple(o.shape) != shape or o.dtype != y.dtype or o.device != y.device:
            o = self._outs[slot] = torch.empty(shape, dtype=y.dtype, device=y.device)
        return o

    def _symmetric_out(self, slot, y, world):
        """Peer-mapped gathered buffer of `slot` (push form) + its rendezvous handle."""
        shape = (world * y.shape[0],) + tuple(y.shape[1:])
        st = self._stage[slot]
        if st is None or tuple(st[0].shape) != shape or st[0].dtype != y.dtype:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(*shape, dtype=y.dtype, device=y.device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            st = self._stage[slot] = (buf, hdl)
            self._outs[slot] = buf
        return st

    def _symmetric_stage(self, slot, y):
        st = self._stage[slot]
        if st is None or tuple(st[0].shape) != tuple(y.shape) or st[0].dtype != y.dtype:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(*y.shape, dtype=y.dtype, device=y.device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            st = self._stage[slot] = (buf, hdl)
        return st

    def submit(self, y_local):
        """Enqueue the gather of `y_local` ([B/G, ...], equal shards).  Returns (gathered tensor, completion event or None).

        Contract for the returned buffer (it is one of `depth` rotating slots, peer-writable in mode "ce"): whatever reads it
        must be enqueued -- on the stream that calls `submit`, or on a stream that stream waits for -- BEFORE the submit that
        re-uses the slot (`depth` submits later).  The push of that later step is ordered after those reads on every rank
        by a cross-rank release barrier."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return y_local, None
        world, rank = dist.get_world_size(), dist.get_rank()
        slot = self._i % self.depth
        self._i += 1
        out = self._out(slot, y_local, world)
        y_local = y_local.contiguous()
        if not y_local.is_cuda or self.mode == "sync":
            dist.all_gather_into_tensor(out, y_local)
            return out, None
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=y_local.device)
        if self.mode in ("ce", "push", "ce_pull"):
            try:
                if self.mode in ("ce", "push"):
                    out, hdl = self._symmetric_out(slot, y_local, world)     # collective on first use of a slot
                else:
                    buf, hdl = self._symmetric_stage(slot, y_local)
            except Exception as err:                                 # no peer mapping on this system: NCCL on the side stream
                import warnings
                warnings.warn("PipelinedGather: symmetric memory unavailable (%s); using NCCL" % (err,))
                self.mode = "nccl"
        comp = torch.cuda.current_stream(y_local.device)
        ready = torch.cuda.Event()
        ready.record(comp)
        self._comm.wait_event(ready)
        with torch.cuda.stream(self._comm):
            if self.mode == "push":
                # same protocol as "ce" (release barrier, pushes, publish barrier), the pushes being ONE kernel: 16-byte stores
                # from a few dozen CTAs straight into every peer's gathered buffer (qt_peer_push)
                from . import _ops as ops
                n = y_local.shape[0]
                full_shape = (world * n,) + tuple(y_local.shape[1:])
                hdl.barrier(channel=self.depth + slot)
                dsts = []
                for step in range(world):
                    dst = (rank + step) % world
                    peer_out = out if dst == rank else hdl.get_buffer(dst, full_shape, y_local.dtype)
                    dsts.append(peer_out[rank * n:(rank + 1) * n])
                ops.peer_push(y_local, dsts, ctas=self.pull_streams or 32)
                hdl.barrier(channel=slot)
            elif self.mode == "ce":
                n = y_local.shape[0]
                # release barrier: a peer may only overwrite my rows of slot `slot` once I am done with what step i - depth left
                # there.  Every rank's communication stream reaches this barrier after its own `ready` event, i.e. after all
                # work its submitting stream had enqueued before this submit -- which is where a consumer of the previous
                # contents of the slot must have been enqueued (see `submit` docstring).  Without it a faster rank could
                # write into a buffer a slower rank is still reading (write-after-read across ranks).
                hdl.barrier(channel=self.depth + slot)
                nps = self.pull_streams if self.pull_streams else min(4, world)
                while len(self._pull) < nps - 1:
                    self._pull.append(torch.cuda.Stream(device=y_local.device))
                start = torch.cuda.Event()
                start.record(self._comm)
                lanes = [self._comm] + self._pull[:nps - 1]
                full_shape = (world * n,) + tuple(y_local.shape[1:])
                for step in range(world):
                    dst = (rank + step) % world                      # my own buffer first, then round the ring
                    peer_out = out if dst == rank else hdl.get_buffer(dst, full_shape, y_local.dtype)
                    lane = lanes[step % len(lanes)]
                    with torch.cuda.stream(lane):
                        if lane is not self._comm and step < len(lanes):
                            lane.wait_event(start)
                        peer_out[rank * n:(rank + 1) * n].copy_(y_local, non_blocking=True)
                        if lane is not self._comm:
                            y_local.record_stream(lane)
                for lane in lanes[1:]:
                    ev = torch.cuda.Event()
                    ev.record(lane)
                    self._comm.wait_event(ev)
                hdl.barrier(channel=slot)                            # every shard of this step has landed in every buffer
            elif self.mode == "ce_pull":
                n = y_local.shape[0]
                buf.copy_(y_local, non_blocking=True)
                hdl.barrier(channel=slot)                            # every rank has staged this step
                nps = self.pull_streams if self.pull_streams else min(4, world - 1)
                while len(self._pull) < nps - 1:
                    self._pull.append(torch.cuda.Stream(device=y_local.device))
                staged = torch.cuda.Event()
                staged.record(self._comm)
                lanes = [self._comm] + self._pull[:nps - 1]
                for step in range(world):
                    src = (rank - step) % world                      # start with my own shard, then walk the ring
                    peer = buf if src == rank else hdl.get_buffer(src, tuple(y_local.shape), y_local.dtype)
                    lane = lanes[step % len(lanes)]
                    with torch.cuda.stream(lane):
                        if lane is not self._comm and step < len(lanes):
                            lane.wait_event(staged)
                        out[src * n:(src + 1) * n].copy_(peer, non_blocking=True)
                for lane in lanes[1:]:
                    ev = torch.cuda.Event()
                    ev.record(lane)
                    self._comm.wait_event(ev)
                hdl.barrier(channel=slot)                            # every rank has read my stage: it may be reused
            else:
                dist.all_gather_into_tensor(out, y_local)
            y_local.record_stream(self._comm)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._done[slot] = done
        return out, done

    def drain(self):
        """The current stream waits for every gather still in flight."""
        for ev in self._done:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        self._done = [None] * self.depth
