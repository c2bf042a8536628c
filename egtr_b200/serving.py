"""Pipelined host-to-host execution of the EGTR forward (throughput serving).

`evaluate_egtr.py`'s loop (`/root/reference/evaluate_egtr.py:26-36`) is synchronous: copy the batch to
the GPU, run the model, read results back — on a B200 the two PCIe copies (21 MB up, 8 MB down per
800x1333 image) cost about as much as a third of the forward.  `PipelinedRunner` keeps the reference's
call shape (pinned host tensors in, host tensors out) but overlaps the three phases of consecutive
batches on three CUDA streams with double-buffered device inputs/outputs:

    copy-in  stream  : H2D of batch i+1
    compute  streams : CUDA-graph replay of the forward for batch i (and, with `concurrency` 2, batch i+1 on a second
                       stream with its own workspace: at batch 1 the decoder and the deep backbone layers are latency-bound
                       chains of small kernels that leave most SMs idle — a second image in flight fills them)
    copy-out stream  : D2H of batch i-1's logits / boxes / pred_rel / pred_connectivity

Every batch still pays its own H2D and D2H; they just no longer serialise with the kernels.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .engine import GraphRunner

RESULT_FIELDS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")


class PipelinedRunner:
    def __init__(self, model, batch: int, height: int, width: int, depth: int = 2, post=None, concurrency: int = 1):
        """`post(outputs) -> dict of device tensors` (optional) runs on the compute stream after each replay — e.g.
        the image-parallel all-gather of per-image records — and its result is what gets copied to the host."""
        self.model = model
        self.post = post
        eng = model.engine()
        self.dev = eng.device
        self.depth = depth
        with torch.cuda.device(self.dev):
            # one captured graph per slot: private static inputs and outputs; slots that share a compute stream replay
            # serially and share a workspace, slots on different compute streams get their own (they overlap in time)
            self.concurrency = max(1, min(concurrency, depth))
            self.slots: List[GraphRunner] = [GraphRunner(eng, batch, height, width, slot=i % self.concurrency, throughput=self.concurrency > 1)
                                              for i in range(depth)]
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            self.s_runs = [torch.cuda.Stream() for _ in range(self.concurrency)]
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
            self.res: List[Dict[str, torch.Tensor]] = []
            for sl in self.slots:
                r = {k: sl.out[k] for k in RESULT_FIELDS}
                self.res.append(post(r) if post is not None else r)
            torch.cuda.synchronize()
            self.host: List[Dict[str, torch.Tensor]] = [
                {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in r.items()} for r in self.res
            ]
            for e in self.ev_run + self.ev_out:
                e.record()  # slots start free
        self.n = 0
        self.h2d_bytes = 0
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.host[0].values())

    def submit(self, pixel_values: torch.Tensor, pixel_mask: Optional[torch.Tensor] = None) -> int:
        """Enqueue one batch given as (ideally pinned) HOST tensors; returns a ticket for `collect`."""
        i = self.n
        s = i % self.depth
        slot = self.slots[s]
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_run[s])  # the previous user of this slot has consumed its inputs
                slot.px.copy_(pixel_values, non_blocking=True)
                if pixel_mask is not None:
                    slot.pm.copy_(pixel_mask, non_blocking=True)
                else:
                    slot.pm.fill_(1)
                self.ev_in[s].record(self.s_in)
            s_run = self.s_runs[s % self.concurrency]
            with torch.cuda.stream(s_run):
                s_run.wait_event(self.ev_in[s])
                s_run.wait_event(self.ev_out[s])  # outputs of this slot have been read back
                slot.graph.replay()
                if self.post is not None:
                    self.res[s] = self.post({k: slot.out[k] for k in RESULT_FIELDS})
                    for v in self.res[s].values():
                        v.record_stream(self.s_out)
                self.ev_run[s].record(s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[s])
                for k, v in self.res[s].items():
                    self.host[s][k].copy_(v, non_blocking=True)
                self.ev_out[s].record(self.s_out)
        self.h2d_bytes = pixel_values.numel() * pixel_values.element_size() + (
            pixel_mask.numel() * pixel_mask.element_size() if pixel_mask is not None else 0)
        self.n += 1
        return i

    def collect(self, ticket: int) -> Dict[str, torch.Tensor]:
        """Block until batch `ticket` is on the host.  The returned pinned tensors are reused `depth` submits later."""
        if not (self.n - self.depth <= ticket < self.n):
            raise ValueError(f"ticket {ticket} is no longer (or not yet) in flight")
        s = ticket % self.depth
        self.ev_out[s].synchronize()
        return self.host[s]

    def run(self, batches):
        """Convenience generator: yields host results for an iterable of (pixel_values, pixel_mask) host batches."""
        pending = []
        for px, pm in batches:
            pending.append(self.submit(px, pm))
            if len(pending) >= self.depth:
                yield {k: v.clone() for k, v in self.collect(pending.pop(0)).items()}
        while pending:
            yield {k: v.clone() for k, v in self.collect(pending.pop(0)).items()}
