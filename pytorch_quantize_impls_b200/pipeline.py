"""Host-to-host inference pipeline: overlaps the H2D copy of batch i+1 and the D2H copy of the logits of batch i-1 with
the kernels of batch i on three CUDA streams (copies and compute on separate streams, PCIe is full duplex).

The reference has nothing comparable (it calls `model(x.to(device))` batch by batch); this is the call a user of this
package makes to push host batches through a quantized network at the rate of the slowest of {H2D, compute, D2H}.
"""
import torch


class HostPipeline:
    def __init__(self, fn, depth=2):
        """fn: callable mapping a device batch to a device result (e.g. an nn.Module in eval mode)."""
        self.fn, self.depth = fn, depth
        self.h2d = torch.cuda.Stream()
        self.d2h = torch.cuda.Stream()
        self._bufs = None

    def _buffers(self, like, dev):
        if self._bufs is None or self._bufs[0].shape != like.shape or self._bufs[0].device != dev:
            self._bufs = [torch.empty(like.shape, dtype=like.dtype, device=dev) for _ in range(self.depth)]
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
            y = self.fn(bufs[slot])
            done = torch.cuda.Event()
            done.record(comp)
            free[slot] = done
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                yout.copy_(y, non_blocking=True)
                y.record_stream(self.d2h)
                last = torch.cuda.Event()
                last.record(self.d2h)
        if last is not None:
            comp.wait_event(last)
        return host_outputs
