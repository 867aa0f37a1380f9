"""Host-to-host inference pipeline: overlaps the H2D copy of batch i+1 and the D2H copy of the logits of batch i-1 with
the kernels of batch i on three CUDA streams (copies and compute on separate streams, PCIe is full duplex).

The reference has nothing comparable (it calls `model(x.to(device))` batch by batch); this is the call a user of this
package makes to push host batches through a quantized network at the rate of the slowest of {H2D, compute, D2H}.
"""
import torch


class HostPipeline:
    def __init__(self, fn, depth=2, graphs=False):
        """fn: callable mapping a device batch to a device result (e.g. an nn.Module in eval mode).
        graphs=True: the forward of each of the `depth` device input buffers is captured once in a CUDA graph
        (GraphedModule) and replayed, one launch per step."""
        self.fn, self.depth, self.graphs = fn, depth, graphs
        self.h2d = torch.cuda.Stream()
        self.d2h = torch.cuda.Stream()
        self._bufs = None
        self._graphed = None

    def _buffers(self, like, dev):
        if self._bufs is None or self._bufs[0].shape != like.shape or self._bufs[0].device != dev:
            self._bufs = [torch.empty(like.shape, dtype=like.dtype, device=dev) for _ in range(self.depth)]
            self._graphed = None
            if self.graphs:
                for b in self._bufs:
                    b.zero_()
                self._graphed = [GraphedModule(self.fn, b) for b in self._bufs]
        return self._bufs

    @torch.no_grad()
    def run(self, host_inputs, host_outputs, device=None):
        """host_inputs / host_outputs: sequences of pinned CPU tensors (outputs are written in place).
        Returns after all device work has been *enqueued*; synchronise (or record an event on the current stream,
        which waits for the last D2H) before reading host_outputs."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        comp = torch.cuda.current_stream()
        bufs = self._buffers(host_inputs[0], dev)
        free = [None] * self.depth          # event: compute finished reading buffer slot
        start = torch.cuda.Event()
        start.record(comp)
        self.h2d.wait_event(start)
        self.d2h.wait_event(start)
        last = None
        for i, (xin, yout) in enumerate(zip(host_inputs, host_outputs)):
            slot = i % self.depth
            with torch.cuda.stream(self.h2d):
                if free[slot] is not None:
                    self.h2d.wait_event(free[slot])
                bufs[slot].copy_(xin, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.h2d)
            comp.wait_event(ready)
            y = self._graphed[slot]() if self._graphed is not None else self.fn(bufs[slot])
            done = torch.cuda.Event()
            done.record(comp)
            free[slot] = done
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                yout.copy_(y, non_blocking=True)
                y.record_stream(self.d2h)
                last = torch.cuda.Event()
                last.record(self.d2h)
            if self._graphed is not None:
                self._graphed[slot].wait_for(last)     # the static output is overwritten by this slot's next replay
        if last is not None:
            comp.wait_event(last)
        return host_outputs


class GraphedModule:
    """fn(x) for ONE fixed input buffer captured in a CUDA graph: a whole quantized forward (quantizer, weight expansion and
    tcgen05 contraction kernels of every layer) is replayed with a single launch, so short steps are not bound by the
    Python / ctypes launch path (7 kernels in 0.5 ms for BASELINE configs[1]).

    The kernels are launched through the C ABI on torch's current stream, which is the capture stream inside
    `torch.cuda.graph`; scratch tensors come from the graph's private memory pool.  `fn` must be free of host
    synchronisation (true of every forward path of this package unless `set_strict(True)`).  The returned tensor is a
    static buffer that the next replay overwrites: `wait_for(event)` makes the next replay wait for a consumer on another
    stream (e.g. the pipelined logits gather)."""

    def __init__(self, fn, static_input, warmup=3, pool=None):
        self.x = static_input
        self._busy = None
        side = torch.cuda.Stream(device=static_input.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # warm up off the capture: attribute setting, tensor-map entry points, caches
            for _ in range(warmup):
                fn(self.x)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.y = fn(self.x)

    def pool(self):
        return self.graph.pool()

    def wait_for(self, event):
        self._busy = event

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if self._busy is not None:
            torch.cuda.current_stream().wait_event(self._busy)
            self._busy = None
        self.graph.replay()
        return self.y
