"""Pipelined host-to-host execution of the EGTR forward (throughput serving).

`evaluate_egtr.py`'s loop (`/root/reference/evaluate_egtr.py:26-36`) is synchronous: copy the batch to
the GPU, run the model, read results back — on a B200 the two PCIe copies (21 MB up, 8 MB down per
800x1333 image) cost about as much as a third of the forward.  `PipelinedRunner` keeps the reference's
call shape (pinned host tensors in, host tensors out) but overlaps the three phases of consecutive
batches on CUDA streams with per-slot static device inputs/outputs:

    copy-in  stream  : H2D of batch i+1
    compute  streams : CUDA-graph replay of the forward for batch i (and, with `concurrency` > 1, further batches on other
                       streams with their own workspaces: at batch 1 the decoder and the deep backbone layers are
                       latency-bound chains of small kernels that leave most SMs idle — other images in flight fill them)
    copy-out stream  : D2H of batch i-1's results

Two boundary formats on each side (SURVEY.md §8f rows 1 and 2 are part of the captured graph):

    input_format "f32"  : `pixel_values` fp32 [B,3,H,W] + `pixel_mask` int64 — what the reference's collate_fn hands the model
                          (`/root/reference/train_egtr.py:176-186`): 21.3 MB per 800x1333 image over PCIe
                 "u8"   : raw uint8 RGB images [B,h0,w0,3]; resize / normalise / pad / mask run on the device
                          (preprocess.cu, Pillow-exact): 3.2 MB per image
    output "raw"        : logits / pred_boxes / pred_rel / pred_connectivity (8.3 MB per image) — the model's outputs
           "triplets"   : the product of `evaluate_batch` (`/root/reference/train_egtr.py:56-94`): boxes, object scores /
                          classes, top-k (s,o,p) and their scores (6.4 KB per image), extracted on the device (triplets.cu);
                          with `gather=True` under torch.distributed the records of all ranks are all-gathered (ONE
                          collective per batch) and every rank reads back the whole batch's triplets
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .engine import GraphRunner
from .postprocess import TripletRecords
from .preprocess import StaticStager

RESULT_FIELDS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")


class PipelinedRunner:
    def __init__(self, model, batch: int, height: int, width: int, depth: int = 2, post=None, concurrency: int = 1,
                 input_format: str = "f32", output: str = "raw", raw_hw: Optional[Tuple[int, int]] = None,
                 resize: Optional[Tuple[int, int]] = None, topk: int = 100, single: bool = False, gather: bool = False):
        """`post(outputs) -> dict of device tensors` (optional, output "raw" only) runs on the compute stream after each
        replay and its result is what gets copied to the host.  input_format "u8": images arrive at `raw_hw` (default: the
        model size) and are resized on the device to DetrFeatureExtractor's target size for `resize` = (size, max_size), or to
        (height, width) when `resize` is None; the result must fit the (height, width) batch tensor (zero padded)."""
        if input_format not in ("f32", "u8") or output not in ("raw", "triplets"):
            raise ValueError("input_format is 'f32' or 'u8', output is 'raw' or 'triplets'")
        if output == "triplets" and post is not None:
            raise ValueError("`post` applies to raw outputs; triplet records are gathered with gather=True")
        self.model = model
        self.post = post
        self.input_format, self.output = input_format, output
        eng = model.engine()
        cfg = eng.cfg
        self.dev = eng.device
        self.depth = depth
        import torch.distributed as dist
        self.world = dist.get_world_size() if (gather and dist.is_available() and dist.is_initialized()) else 1
        self._dist = dist
        with torch.cuda.device(self.dev):
            # one captured graph per slot: private static inputs and outputs; slots that share a compute stream replay
            # serially and share a workspace, slots on different compute streams get their own (they overlap in time)
            self.concurrency = max(1, min(concurrency, depth))
            self.stagers: List[Optional[StaticStager]] = []
            self.records: List[Optional[TripletRecords]] = []
            self.slots: List[GraphRunner] = []
            for i in range(depth):
                stager = None
                if input_format == "u8":
                    kw = dict(size=resize[0], max_size=resize[1]) if resize is not None else dict(out_hw=(height, width))
                    stager = StaticStager(batch, raw_hw or (height, width), (height, width), self.dev, **kw)
                rec = TripletRecords(batch, cfg.num_queries, cfg.num_labels, cfg.num_rel_labels, cfg.num_labels, self.dev, topk=topk,
                                     single=single) if output == "triplets" else None
                pro = (lambda r, s=stager: s.enqueue(r.px, r.pm)) if stager is not None else None
                epi = (lambda r, out, t=rec: t.enqueue(out)) if rec is not None else None
                self.stagers.append(stager)
                self.records.append(rec)
                self.slots.append(GraphRunner(eng, batch, height, width, slot=i % self.concurrency, throughput=self.concurrency > 1,
                                              prologue=pro, epilogue=epi))
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            # the all-gather of the triplet records gets its own high-priority stream: on the compute streams it would hold them
            # (and the SMs of its spinning kernel) for as long as the slowest peer is late
            self.s_comm = torch.cuda.Stream(priority=-1) if self.world > 1 else None
            self.ev_rep = [torch.cuda.Event() for _ in range(depth)]
            self.s_runs = [torch.cuda.Stream() for _ in range(self.concurrency)]
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
            self.res: List[Dict[str, torch.Tensor]] = []
            for s, sl in enumerate(self.slots):
                if output == "triplets":
                    flat = self.records[s].flat
                    # static gather target per slot: no allocator traffic across the compute streams
                    self.res.append({"triplets": torch.empty(self.world, flat.numel(), dtype=flat.dtype, device=self.dev)
                                     if self.world > 1 else flat})
                else:
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
        """Enqueue one batch given as (ideally pinned) HOST tensors; returns a ticket for `collect`.
        input_format "f32": pixel_values fp32 [B,3,H,W] (+ pixel_mask); "u8": uint8 [B,h0,w0,3] images (no mask)."""
        i = self.n
        s = i % self.depth
        slot = self.slots[s]
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_run[s])  # the previous user of this slot has consumed its inputs
                if self.input_format == "u8":
                    self.stagers[s].u8.copy_(pixel_values, non_blocking=True)
                else:
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
                if self.output != "triplets" and self.post is not None:
                    self.res[s] = self.post({k: slot.out[k] for k in RESULT_FIELDS})
                    for v in self.res[s].values():
                        v.record_stream(self.s_out)
                if self.output == "triplets" and self.world > 1:
                    self.ev_rep[s].record(s_run)
                else:
                    self.ev_run[s].record(s_run)
            if self.output == "triplets" and self.world > 1:  # the path's only collective: per-image triplet records of every rank
                with torch.cuda.stream(self.s_comm):
                    self.s_comm.wait_event(self.ev_rep[s])
                    self._dist.all_gather_into_tensor(self.res[s]["triplets"], self.records[s].flat)
                    self.ev_run[s].record(self.s_comm)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[s])
                for k, v in self.res[s].items():
                    self.host[s][k].copy_(v, non_blocking=True)
                self.ev_out[s].record(self.s_out)
        self.h2d_bytes = pixel_values.numel() * pixel_values.element_size() + (
            pixel_mask.numel() * pixel_mask.element_size() if (pixel_mask is not None and self.input_format == "f32") else 0)
        self.n += 1
        return i

    def collect(self, ticket: int) -> Dict[str, torch.Tensor]:
        """Block until batch `ticket` is on the host.  The returned pinned tensors are reused `depth` submits later.
        output "triplets": the dict of `TripletRecords.decode` (views of the pinned record buffer; with gather, all ranks' images)."""
        if not (self.n - self.depth <= ticket < self.n):
            raise ValueError(f"ticket {ticket} is no longer (or not yet) in flight")
        s = ticket % self.depth
        self.ev_out[s].synchronize()
        if self.output == "triplets":
            return self.records[s].decode(self.host[s]["triplets"])
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
