"""Host-to-host inference pipeline: pinned host batches in, class maps (or logits) out, with the H2D upload of
batch i+1 and the D2H download of result i-1 overlapped with the forward of batch i.

This is the end-to-end path of the evaluator loop (reference models/evaluator.py:156-180 + :89-103:
``batch.to(device)`` -> ``net_G(x1, x2)`` -> ``argmax`` -> ``.cpu()``) written once, with the copies on a side
stream.  Results are identical to calling the module batch by batch.

    pipe = PairPipeline(net, out="argmax_u8")
    for pred in pipe.run(batches):        # batches: iterable of (x1, x2) CPU tensors (pinned => async copies)
        ...                               # pred: (B,H,W) uint8 CPU tensor (valid until the next iteration)

``inputs="u8_hwc"`` takes the decoded images as they come out of the image reader — uint8 (B,H,W,3) — uploads those
bytes (4x less PCIe traffic) and normalises them on the device exactly as the reference loaders do
(dahitra_b200.inputs.normalize_u8; ``kind`` = "levir" or "xbd").
"""
from __future__ import annotations

import torch


class PairPipeline:
    def __init__(self, net, out: str = "argmax_u8", depth: int = 2, inputs: str = "f32_nchw", kind: str = "levir"):
        if out not in ("argmax_u8", "argmax", "logits"):
            raise ValueError(out)
        if inputs not in ("f32_nchw", "u8_hwc"):
            raise ValueError(inputs)
        self.net, self.out, self.depth, self.inputs, self.kind = net, out, max(2, depth), inputs, kind
        self.dev = next(net.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("dahitra_b200: PairPipeline needs the module on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.out_stream = torch.cuda.Stream(self.dev)       # D2H of the results: the next forward does not queue behind it
        self._slots = None
        self._host = None

    def _ensure(self, x1):
        shape = tuple(x1.shape)
        u8 = self.inputs == "u8_hwc"
        if u8:                                   # (B,H,W,3) uint8 on the wire, (B,3,H,W) fp32 for the network
            shape = (shape[0], 3, shape[1], shape[2])
        if self._slots is None or self._slots[0][0].shape != shape:
            raw = lambda: torch.empty(tuple(x1.shape), dtype=torch.uint8, device=self.dev) if u8 else None
            self._slots = [(torch.empty(shape, dtype=torch.float32, device=self.dev),
                            torch.empty(shape, dtype=torch.float32, device=self.dev),
                            torch.cuda.Event(), torch.cuda.Event(), raw(), raw()) for _ in range(self.depth)]
            B, _, H, W = shape
            nc = self.net.output_nc
            hs = {"argmax_u8": ((B, H, W), torch.uint8), "argmax": ((B, H, W), torch.int64),
                  "logits": ((B, nc, H, W), torch.float32)}[self.out]
            self._host = [torch.empty(hs[0], dtype=hs[1]).pin_memory() for _ in range(self.depth)]

    def _upload(self, i, x1, x2):
        d1, d2, ready, free, u1, u2 = self._slots[i % self.depth]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(free)          # the forward that last read this slot has finished
            (u1 if u1 is not None else d1).copy_(x1, non_blocking=True)
            (u2 if u2 is not None else d2).copy_(x2, non_blocking=True)
            ready.record(self.copy_stream)

    @torch.no_grad()
    def run(self, batches):
        it = iter(batches)
        main = torch.cuda.current_stream(self.dev)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._ensure(nxt[0])
        for s in self._slots:
            s[3].record(main)
        self._upload(0, *nxt)
        i = 0
        pending = None                                  # (host buffer, event) of the previous step
        while nxt is not None:
            cur_i = i
            try:
                nxt = next(it)
                self._upload(cur_i + 1, *nxt)           # overlaps with the forward below
            except StopIteration:
                nxt = None
            d1, d2, ready, free, u1, u2 = self._slots[cur_i % self.depth]
            main.wait_event(ready)
            if u1 is not None:                          # normalise on the device, bit-identical to the reference loaders
                from .inputs import normalize_u8
                normalize_u8(u1, self.kind, out=d1)
                normalize_u8(u2, self.kind, out=d2)
            if self.out == "argmax_u8":
                self.net._engine.forward_pair(self.net, d1, d2, want_argmax=True)
                res = self.net._engine.last_argmax
            else:
                # the xBD variant's own forward takes one stacked (B,6,H,W) tensor; the engine takes the two halves
                y = self.net(d1, d2) if getattr(self.net, "VARIANT", "levir") == "levir" else \
                    self.net._engine.forward_pair(self.net, d1, d2)
                res = torch.argmax(y, dim=1) if self.out == "argmax" else y
            free.record(main)
            host = self._host[cur_i % self.depth]
            done = torch.cuda.Event()
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(free)        # recorded right after the forward: the result is complete
                host.copy_(res, non_blocking=True)
                done.record(self.out_stream)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (host, done, res)                 # `res` stays referenced until its copy has completed
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
