#!/usr/bin/env python
"""Throughput bench of the EGTR inference hot path on B200 (contract: task prompt §④ / BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--workload B|C|D|E]   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...               # the reference algorithm on the host CPU cores

A step = one pass of the hot path over one batch of synthetic images per GPU: forward (backbone -> encoder -> decoder ->
relation head) + triplet extraction (`evaluate_batch`'s tensor part) and, for N > 1, the ONE all-gather of per-image
triplet records.  Default workload B of BASELINE.json (VG config, 3x800x1333, N_q=200, 150 classes, 50 predicates, batch 1
per GPU — the configuration the metric is quoted on); C / D / E are BASELINE.json configs[2..4] at their per-GPU batch.
Prints ONE JSON line (rank 0).

  value : images/s with the inputs (fp32 pixel_values / pixel_mask) already resident in HBM; every step one CUDA-graph
          replay, several forwards in flight on separate streams with private workspaces; CUDA events around each block
          of K timed steps (barrier + synchronize on both sides), max over ranks; blocks are repeated until >= --min-time
          seconds are timed and the MEDIAN block is reported (min / max beside it).  Inputs rotate over distinct resident
          images (> L2).
  e2e   : images/s through the public serving API (`egtr_b200.serving.PipelinedRunner`) from pinned HOST buffers: per step the
          H2D of the uint8 images (resize / normalise / pad / mask on the device) and the D2H of the triplet records are
          inside the timed region.  `e2e_raw` is the round-1 boundary (fp32 pixel_values + int64 mask in, raw logits /
          pred_rel / pred_connectivity out).
  roofline : the kernel with the largest share of the step — the TMA-fed tcgen05 GEMM — from a CUPTI trace (torch.profiler) of
             the TIMED configuration; `roofline_msda_enc` / `roofline_msda_dec` / `roofline_decoder` / `roofline_relation` report the kernels
             BASELINE.json names (HBM bytes / FLOPs as defined in SURVEY.md §8d).
  reference_gpu : the oracle port (plain torch, the reference's own ops) run eagerly ON THE GPU — the stand-in for the
                  reference's GPU FPS script (`evaluate_egtr.py:26-36`), which cannot travel to the GPU box.
  cpu_baseline : the CPU oracle timed on this box's host cores on a bounded sample (rank 0, N=1).
"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (3x800x1333, N_q=200)"
# BASELINE.json configs[1..4]: (label-space / image-size table of egtr_b200.config.WORKLOADS, per-GPU batch, description)
# L1 wavefront bytes per bilinear sample: fp32 value rows [S, M, D] make every corner its own 128-byte line; the fp16 pair
# records (EGTR_FMT_H16PAIR) put the two x-neighbours in one line
MSDA_L1_BYTES_PER_SAMPLE = {"f32": 4 * 128, "h16": 2 * 128}
BENCH_WORKLOADS = {
    "B": ("B", 1, "VG config, batch 1 per GPU (configs[1])"),
    "C": ("B", 4, "VG config, 32 images image-parallel over 8 GPUs = 4 per GPU (configs[2])"),
    "D": ("D", 4, "Open Images V6 label space, 16 images over 4 GPUs = 4 per GPU (configs[3])"),
    "E": ("E", 8, "stress: N_q=300, 200 predicates, 1024x1024, 64 images over 8 GPUs = 8 per GPU (configs[4])"),
}


def _traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu pass (profiles/r0*_traffic.json, tools/gpu_traffic.sh), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            k = json.load(open(p)).get("kernels", {}).get(kernel)
            if k:
                return k["dram_bytes_per_launch"]
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads():
    """Threads for the CPU legs: all host cores up to 32 — on the 128-core shared GPU hosts torch's CPU
    kernels at these sizes get slower, not faster, beyond that (round-1 measurement: 108 s/forward at 128)."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("EGTR_CPU_THREADS", "32"))))


def build_case(workload, batch):
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    table = BENCH_WORKLOADS[workload][0]
    cfg = workload_config(table)
    H, W = WORKLOADS[table]["image"]
    sd = synth_state_dict(cfg, seed=0)
    px, mask = synth_images(batch, H, W, seed=1)
    return cfg, sd, px, mask, (H, W)


def workload_string(workload, cfg, H, W, batch):
    """Identical in both arms (the driver compares the strings)."""
    return (f"{workload}: {BENCH_WORKLOADS[workload][2]}; 3x{H}x{W}, N_q={cfg.num_queries}, K={cfg.num_labels}, "
            f"P={cfg.num_rel_labels}, {batch} image(s) per GPU per step")


def run_reference(args):
    """The reference algorithm on the host CPU (oracle port; the Python reference cannot travel to the GPU box)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import egtr_oracle as orc
    orc.set_msda_impl("grid_sample")  # the reference's own CPU path for MSDeformAttn (deformable_detr.py:925-960, 1096-1101)
    Bl = args.batch_per_gpu or BENCH_WORKLOADS[args.workload][1]
    cfg, sd, px, mask, (H, W) = build_case(args.workload, 1)
    cores = cpu_threads()
    torch.set_num_threads(cores)
    for _ in range(max(1, args.warmup)):
        orc.forward(sd, cfg, px, mask)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.forward(sd, cfg, px, mask)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} forwards of ONE 3x{H}x{W} image each (a bounded sample of the step's {Bl} image(s)), {max(1, args.warmup)} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, cfg, H, W, Bl), "device": f"host CPU, {cores} threads"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def kernel_trace(fn):
    """Per-kernel GPU durations of `fn()` from a CUPTI activity trace (torch.profiler): [(short name, us, grid CTAs)], or None."""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "trace.json")
            prof.export_chrome_trace(path)
            trace = json.load(open(path))
        out = []
        for ev in trace.get("traceEvents", []):
            if ev.get("cat") != "kernel" or "dur" not in ev:
                continue
            name = ev.get("name", "").replace("void ", "").replace("egtr::<unnamed>::", "").replace("egtr::(anonymous namespace)::", "")
            m = re.match(r"(\w+)(<[^>]*>)?", name)
            short = (m.group(1) + (m.group(2) or "")) if m else name[:40]
            grid = ev.get("args", {}).get("grid", [0, 0, 0])
            ctas = 1
            for g in grid:
                ctas *= max(1, int(g))
            out.append((short, float(ev["dur"]), ctas))
        return out or None
    except Exception as exc:  # noqa: BLE001
        print(f"kernel trace unavailable: {exc!r}", file=sys.stderr)
        return None


def summarize(trace, n_forwards, sms):
    """name -> dict(n per forward, us per forward, sm-weighted us per forward, avg grid)."""
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0])
    for name, us, ctas in trace:
        t = tot[name]
        t[0] += 1
        t[1] += us
        # a persistent one-CTA-per-SM grid of g CTAs holds g / sms of the GPU; other kernels (many small CTAs per SM) are charged in full
        persistent = name.startswith(("gemm_p32_kernel", "gemm_sbf16_kernel", "relhead_kernel", "decoder_kernel"))
        t[2] += us * (min(1.0, ctas / sms) if persistent else 1.0)
        t[3] += ctas
    return {k: dict(n=v[0] / n_forwards, us=v[1] / n_forwards, us_sm_weighted=v[2] / n_forwards, avg_ctas=v[3] / max(1, v[0])) for k, v in tot.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(BENCH_WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="override the workload's per-GPU batch")
    ap.add_argument("--min-time", type=float, default=1.0, help="seconds of timed blocks per leg (each block = --steps steps)")
    ap.add_argument("--cpu-sample", type=int, default=2, help="oracle forwards timed for cpu_baseline (0 = skip)")
    ap.add_argument("--reference-gpu", type=int, default=3, help="eager torch-CUDA forwards of the oracle port timed for reference_gpu (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    args.warmup = max(3, args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from egtr_b200 import _lib
    from egtr_b200.engine import GraphRunner
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.postprocess import TripletRecords
    from egtr_b200.serving import PipelinedRunner

    Bl = args.batch_per_gpu or BENCH_WORKLOADS[args.workload][1]
    cfg, sd, px, mask, (H, W) = build_case(args.workload, Bl)
    N, P, K = cfg.num_queries, cfg.num_rel_labels, cfg.num_labels
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(sd)
    model.cuda().eval()
    eng = model.engine()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_blocks(run, steps):
        """Blocks of `steps` steps, each bracketed by barrier + synchronize and timed with CUDA events, until --min-time
        seconds are timed (at least 3 blocks).  Returns the per-block seconds."""
        blocks, total = [], 0.0
        while len(blocks) < 3 or (total < args.min_time and len(blocks) < 200):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(steps)
            e1.record()
            barrier()
            t = e0.elapsed_time(e1) / 1000.0
            if world > 1:  # the same number of blocks on every rank: decide on the max over ranks
                tt = torch.tensor([t], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt[0])
            blocks.append(t)
            total += t
        return blocks

    # ------------------------------------------------------------ leg 1: device-resident inputs, CUDA-graph replay
    # Throughput mode: `conc` forwards in flight, each a captured CUDA graph (forward + triplet extraction) with its own
    # workspace on its own stream (at batch 1 the decoder and the deep backbone layers are latency-bound chains of small
    # kernels; other images fill the SMs they leave idle).  Inputs rotate over NIMG distinct resident batches (> L2 in total).
    # (batch 1: sixteen images in flight on quarter grids measure +5 % over eight on half grids; at batch 4 - 8 the launches are
    # long enough that two batches on half grids are best: 415 vs 392 images/s at C, 375 vs 283 at E)
    conc = int(os.environ.get("EGTR_PIPE_CONCURRENCY", str(16 if Bl == 1 else max(2, 8 // Bl))))
    depth = int(os.environ.get("EGTR_PIPE_DEPTH", str(2 * conc)))
    NIMG = max(2, 8 // Bl)  # distinct resident input batches the forwards rotate over
    px_d = [torch.roll(px, shifts=17 * i, dims=3).to(dev) for i in range(NIMG)]
    mask_d = mask.to(dev)
    in_bytes = NIMG * (px_d[0].numel() * 4 + mask_d.numel() * 8)
    recs = [TripletRecords(Bl, N, K, P, cfg.num_labels, dev) for _ in range(conc)]
    gathered = [torch.empty(world, r.flat.numel(), dtype=r.flat.dtype, device=dev) for r in recs] if world > 1 else None
    runners = [GraphRunner(eng, Bl, H, W, slot=i, throughput=conc > 1, epilogue=lambda r, out, t=recs[i]: t.enqueue(out)) for i in range(conc)]
    lone = eng.graph_runner(Bl, H, W, slot=0, throughput=False)  # latency configuration of a single forward (split-K on)
    streams = [torch.cuda.Stream() for _ in range(conc)]
    main_s = torch.cuda.current_stream()

    # The path's only collective — the per-image triplet records of every rank — runs on ONE high-priority stream:
    # issued on the sixteen compute streams it kept them (and the SMs its spinning kernels held) waiting on the peers' skew
    # (N = 8: 0.94 of linear; round 2 measured 0.98 with eight streams).
    comm_s = torch.cuda.Stream(priority=-1) if world > 1 else None
    ev_rep = [torch.cuda.Event() for _ in range(conc)]
    ev_comm = [torch.cuda.Event() for _ in range(conc)]

    def run_resident(n):
        for st_ in streams:
            st_.wait_stream(main_s)
        for i in range(n):
            j = i % conc
            with torch.cuda.stream(streams[j]):
                if world > 1:
                    streams[j].wait_event(ev_comm[j])  # the previous records of this slot have been gathered
                runners[j](px_d[i % NIMG], mask_d)
                if world > 1:
                    ev_rep[j].record(streams[j])
            if world > 1:
                with torch.cuda.stream(comm_s):
                    comm_s.wait_event(ev_rep[j])
                    dist.all_gather_into_tensor(gathered[j], recs[j].flat)
                    ev_comm[j].record(comm_s)
        for st_ in streams:
            main_s.wait_stream(st_)
        if world > 1:
            main_s.wait_stream(comm_s)

    run_resident(max(args.warmup, conc))
    barrier()
    # single-forward latency (one batch at a time, L2 flushed before it) for reference
    lat = []
    for i in range(5):
        flush.fill_(1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        lone(px_d[i % NIMG], mask_d)
        s1.record()
        torch.cuda.synchronize()
        lat.append(s0.elapsed_time(s1))
    latency_ms = sorted(lat)[len(lat) // 2]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    blocks_res = timed_blocks(run_resident, args.steps)

    # ------------------------------------------------------------ output check of the timed configuration (untimed)
    # Every forward in flight must reproduce the same image's forward run alone in the latency configuration (which the GPU tests
    # pin against the reference's golden at this size).  A throughput number whose outputs deviate is not a number.
    CHECK_KEYS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")
    chk_n, chk_bad, chk_worst, chk_error, chk_flips = 0, 0, 0.0, None, 0

    def check_err(got, want):
        """Worst max-norm relative error over the four outputs.  A query whose arg-max class differs from the lone forward's (a
        near-tie of its two best logits: `torch.argmax` then selects another frequency-bias row, model/egtr.py:405-413) is counted
        in `class_flips`; it is excused — its pairs left out of pred_rel — only if the lone forward's own margin between the two
        classes is below 2e-3 of max|logits|, otherwise the forward counts as deviating."""
        lg, lw = got["logits"], want["logits"]
        cg, cw = lg.argmax(-1), lw.argmax(-1)
        flip = cg != cw
        nflip, bad = int(flip.sum()), False
        if nflip:
            margin = (lw.gather(-1, cw[..., None]) - lw.gather(-1, cg[..., None])).squeeze(-1)
            bad = bool((margin[flip] > 2e-3 * lw.abs().max()).any())
        keep = ~(flip[:, :, None] | flip[:, None, :])
        e = 0.0
        for k in CHECK_KEYS:
            d = (got[k] - want[k]).abs()
            if k == "pred_rel":
                d = d * keep[..., None]
            e = max(e, float(d.max() / want[k].abs().max()))
        return (1.0 if bad else e), nflip
    try:
        want = []
        for i in range(NIMG):
            o = lone(px_d[i], mask_d)
            want.append({k: o[k].clone() for k in CHECK_KEYS})
        torch.cuda.synchronize()
        for r in range(4):
            for st_ in streams:
                st_.wait_stream(main_s)
            for i in range(conc):
                with torch.cuda.stream(streams[i]):
                    runners[i](px_d[(i + r) % NIMG], mask_d)
            for st_ in streams:
                main_s.wait_stream(st_)
            torch.cuda.synchronize()
            for i in range(conc):
                e, nf = check_err(runners[i].out, want[(i + r) % NIMG])
                chk_n, chk_bad, chk_worst, chk_flips = chk_n + 1, chk_bad + int(e > 1e-3), max(chk_worst, e), chk_flips + nf
        del want
    except Exception as exc:  # noqa: BLE001  (the check must never cost the bench line; it is reported instead)
        chk_error = repr(exc)

    # ------------------------------------------------------------ CUPTI trace of the timed configuration (untimed)
    n_tr = max(conc, min(args.steps, 16))
    tr_timed = kernel_trace(lambda: run_resident(n_tr))
    tr_lone = kernel_trace(lambda: [lone(px_d[i % NIMG], mask_d) for i in range(4)])

    # ------------------------------------------------------------ leg 2: end to end, host buffers in -> host results out
    # public API: egtr_b200.serving.PipelinedRunner — every step pays its own H2D (uint8 images; staging kernels on the device)
    # and D2H (triplet records, all ranks' after the all-gather); copies of neighbouring steps overlap the graph replays.
    g = torch.Generator().manual_seed(2)
    u8_h = torch.randint(0, 256, (Bl, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
    pipe = PipelinedRunner(model, Bl, H, W, depth=depth, concurrency=conc, input_format="u8", output="triplets", gather=world > 1)

    def make_e2e(p, *inputs):
        def run(n):
            pending, last = [], None
            for _ in range(n):
                pending.append(p.submit(*inputs))
                if len(pending) >= depth:
                    last = p.collect(pending.pop(0))
            while pending:
                last = p.collect(pending.pop(0))
            return last
        return run

    run_e2e = make_e2e(pipe, u8_h)
    run_e2e(args.warmup)
    t0 = time.perf_counter()
    blocks_e2e = timed_blocks(run_e2e, args.steps)
    t_e2e_wall = time.perf_counter() - t0
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    del pipe
    # the round-1 boundary for continuity (workload B only): fp32 pixel_values + int64 mask in, raw model outputs out
    px_h, mask_h = px.pin_memory(), mask.pin_memory()
    blocks_raw, raw_h2d, raw_d2h = None, 0, 0
    if args.workload == "B":
        pipe_raw = PipelinedRunner(model, Bl, H, W, depth=depth, concurrency=conc)
        run_raw = make_e2e(pipe_raw, px_h, mask_h)
        run_raw(args.warmup)
        blocks_raw = timed_blocks(run_raw, args.steps)
        raw_h2d, raw_d2h = pipe_raw.h2d_bytes, pipe_raw.d2h_bytes
        del pipe_raw
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------ eager passes: launch count, stage spans (lone forward)
    def step_eager():
        o = model(pixel_values=px_h.to(dev, non_blocking=True), pixel_mask=mask_h.to(dev, non_blocking=True),
                  output_attentions=False, output_attention_states=True, output_hidden_states=True)
        return [o[k].cpu() for k in ("logits", "pred_boxes", "pred_rel", "pred_connectivity")]

    model.use_cuda_graph = False
    eng.probe = None
    # kernels per step in the timed (throughput) configuration = launches of one eager forward with the same settings + triplets
    out_e = eng.forward(px_d[0], mask_d, throughput=conc > 1)
    recs[0].enqueue(out_e)
    torch.cuda.synchronize()
    _lib.call("egtr_launch_count_reset")
    out_e = eng.forward(px_d[0], mask_d, throughput=conc > 1)
    recs[0].enqueue(out_e)
    torch.cuda.synchronize()
    launches = int(_lib.call("egtr_launch_count"))
    eng.probe = {}
    eng.probe_flops = {}
    n_probe = min(args.steps, 10)
    for _ in range(n_probe):
        torch.cuda._sleep(int(2e7))  # ~10 ms head start for the host: the probe events then bracket GPU execution, not launch gaps
        step_eager()
    torch.cuda.synchronize()
    probe, eng.probe = eng.probe, None
    spans = {k: sum(a.elapsed_time(b) for a, b in v) / 1000.0 / n_probe for k, v in probe.items()}  # seconds per forward
    gemm_flops = eng.probe_flops.get("gemm_p32", 0) / n_probe  # algorithmic 2MNK over the gemm_p32 launches of ONE forward

    # max over ranks of the check
    if world > 1:
        t = torch.tensor([chk_worst], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        chk_worst = float(t[0])
        c = torch.tensor([chk_n, chk_bad, chk_flips], device=dev, dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        chk_n, chk_bad, chk_flips = int(c[0]), int(c[1]), int(c[2])

    if rank == 0:
        peaks = _peaks()
        S = sum(h * w for h, w in eng._workspace(Bl, H, W)["shapes"])
        med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
        imgs = world * Bl * args.steps
        t_res, t_e2e, t_raw = med(blocks_res), med(blocks_e2e), (med(blocks_raw) if blocks_raw else None)
        ms_step = 1000 * t_res / args.steps
        k_timed = summarize(tr_timed, n_tr, sms) if tr_timed else {}
        k_lone = summarize(tr_lone, 4, sms) if tr_lone else {}

        def pick(tab, prefix):
            prefixes = prefix if isinstance(prefix, tuple) else (prefix,)
            ks = [k for k in tab if k.startswith(prefixes)]
            if not ks:
                return None
            return dict(n=sum(tab[k]["n"] for k in ks), us=sum(tab[k]["us"] for k in ks), us_w=sum(tab[k]["us_sm_weighted"] for k in ks))

        def hbm_roof(prefix, bytes_per_launch, label):
            """HBM roofline of one kernel class: achieved = algorithmic bytes / average launch duration in the TIMED configuration
            (CUPTI); the same kernel timed with nothing else on the GPU (`lone`) beside it."""
            kt, kl = pick(k_timed, prefix), pick(k_lone, prefix)
            src = "CUPTI trace of the timed configuration"
            if kt is None:  # no trace: CUDA events around the launches of a lone eager forward
                name = label
                if name not in spans or not probe.get(name):
                    return None
                kt = dict(n=len(probe[name]) / n_probe, us=1e6 * spans[name])
                src = "CUDA events around the launches of a lone eager forward (no CUPTI trace)"
            us = kt["us"] / kt["n"]
            ach = bytes_per_launch / us / 1e3
            r = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                 "traffic": _traffic(label), "kernel": label + " (" + prefix[0] + "...>)", "avg_launch_us": us, "launches_per_step": kt["n"],
                 "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peaks["src"] + " (hbm_gbs)", "timing": src}
            if kl is not None:
                us_l = kl["us"] / kl["n"]
                r["lone"] = {"avg_launch_us": us_l, "achieved": bytes_per_launch / us_l / 1e3, "frac": bytes_per_launch / us_l / 1e3 / peaks["hbm"]}
            return r

        # MSDeformAttn, encoder form: SURVEY.md §8d algorithmic bytes = 4*B*[S*C + Lq*M*L*P*3 + Lq*C] = 3584*S per image
        r_msda = hbm_roof(("msda_kernel<true, 32", "msda_kernel<1, 32"), 3584 * S * Bl, "msda_enc")
        r_msda_dec = hbm_roof(("msda_kernel<true, 8", "msda_kernel<1, 8"), (1024 * S + 2560 * N) * Bl, "msda_dec")
        # fused decoder stack (decoder.cu, forwards of at most 512 queries): per layer the MSDeformAttn bytes of SURVEY.md §8d's
        # decoder form plus the layer's weights once (0.95 M parameters as bf16 hi + lo planes = 3.8 MB)
        nl_dec = cfg.decoder_layers
        r_dec = hbm_roof(("decoder_kernel",), nl_dec * ((1024 * S + 2560 * N) * Bl + 3.8e6), "decoder_kernel")
        if r_dec is not None:
            kt_d = pick(k_timed, "decoder_kernel")
            r_dec["avg_launch_us_sm_weighted"] = kt_d["us_w"] / kt_d["n"] if kt_d else None
            r_dec["note"] = ("ONE kernel for the six decoder layers on a cluster of 8 CTAs per image (16 for a lone forward): latency-bound "
                             "by design (phase chain of ~70 cluster barriers), it holds 8 of the 148 SMs; algorithmic bytes as for the "
                             "stand-alone decoder MSDeformAttn launches it replaces + the layer weights")
        if r_msda is not None:
            # the binding roof of the gather is the SM's L1 path, not HBM: one 128-byte wavefront per
            # (query, head, level, point, corner) at 128 B/clk/SM — report the fraction of THAT roof beside the HBM one
            per_sample = MSDA_L1_BYTES_PER_SAMPLE[eng.msda_value]
            l1_bytes = Bl * S * 8 * 16 * per_sample
            floor_us = l1_bytes / (sms * 128.0 * (clocks["sm_mhz"] or 1965.0 if clocks else 1965.0) * 1e6) * 1e6
            r_msda["l1_roof"] = {"bytes_through_l1_per_launch": l1_bytes, "floor_us": floor_us, "frac": floor_us / r_msda["avg_launch_us"],
                                 "note": f"{per_sample} B of L1 wavefronts per bilinear sample (value layout: {eng.msda_value})"}
        # relation head (a13-a16): algorithmic HBM bytes per image: Q/K/h in 13*N*1024, logits 4NK, freq-bias
        # gather min(4N^2P, 4(K+1)^2P), outputs 4N^2(P+1), weights ~5.3 MB once; FLOPs as written in the reference
        rel_bytes = Bl * (13 * N * 1024 + 4 * N * K + min(4 * N * N * P, 4 * (K + 1) ** 2 * P) + 4 * N * N * (P + 1)) + 5.3e6
        rel_flops_ref = Bl * (N * N * (7 * 2 * 512 + 2 * (2 * 512 * 512 + 2 * 512 * 256) + 2 * 256 * P + 2 * 256) + 14 * N * 2 * 256 * 256)
        # executed by the factorised form: 14 per-query GEMMs [N,256]x[256,513], two 256x256 layer-2 GEMMs and layer 3 per pair
        rel_flops_exec = Bl * (14 * N * 2 * 256 * 513 + N * N * (2 * 2 * 256 * 256 + 2 * 256 * P + 2 * 256))
        r_rel = None
        kt = pick(k_timed, "relhead_kernel")
        kl = pick(k_lone, "relhead_kernel")
        if kt is not None:
            us, us_w = kt["us"] / kt["n"], kt["us_w"] / kt["n"]
            pair_flops = Bl * N * N * (2 * 2 * 256 * 256 + 2 * 256 * P + 2 * 256)
            r_rel = {"bound": "tensor", "achieved": pair_flops / us_w / 1e6, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": pair_flops / us_w / 1e6 / peaks["bf16_sustained"], "traffic": _traffic("relhead_kernel"),
                     "kernel": "relhead_kernel (fused pair stage: gating + layer 1 producer, layer 2 / layer 3 tcgen05, sigmoid epilogue)",
                     "avg_launch_us": us, "avg_launch_us_sm_weighted": us_w, "algorithmic_flops_per_launch": pair_flops,
                     "executed_bf16_tflops": 3 * pair_flops / us_w / 1e6, "frac_executed_bf16": 3 * pair_flops / us_w / 1e6 / peaks["bf16_sustained"],
                     "hbm_GBps_on_algorithmic_bytes": rel_bytes / us / 1e3, "hbm_frac": rel_bytes / us / 1e3 / peaks["hbm"],
                     "reference_flops_per_image": rel_flops_ref / Bl, "executed_flops_per_image": rel_flops_exec / Bl,
                     "timing": "CUPTI trace of the timed configuration; sm-weighted = duration x (grid CTAs / SMs)",
                     "note": "2*M*N*K FLOPs of the pair stage as executed (layer 2 of both MLPs + layer 3); bf16x3 products cap frac at 1/3"}
            if kl is not None:
                r_rel["lone"] = {"avg_launch_us": kl["us"] / kl["n"], "frac": pair_flops / (kl["us_w"] / kl["n"]) / 1e6 / peaks["bf16_sustained"]}
        elif "stage_relation" in spans:
            tl = spans["stage_relation"]
            r_rel = {"bound": "tensor", "achieved": rel_flops_ref / tl / 1e12, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": rel_flops_ref / tl / 1e12 / peaks["bf16"], "traffic": None, "kernel": "relation head stage (all its launches)",
                     "stage_us": 1e6 * tl, "hbm_GBps_on_algorithmic_bytes": rel_bytes / tl / 1e9,
                     "hbm_frac": rel_bytes / tl / 1e9 / peaks["hbm"], "note": "FLOPs as written in the reference (SURVEY.md §8d); lone eager forward"}
        if "stage_relation" in spans and r_rel is not None:
            r_rel["stage_us_lone_eager"] = 1e6 * spans["stage_relation"]
        # dominant kernel of the step: the TMA-fed tcgen05 GEMM (every launch of gemm_p32_kernel: backbone convolutions,
        # input_proj, encoder Linears, decoder value projection).  Algorithmic FLOPs = 2*M*N*K summed over its launches; each
        # product is executed as three bf16 MMAs, so the ceiling of `frac` against the bf16 peak is 1/3.
        r_gemm = None
        kt, kl = pick(k_timed, "gemm_p32_kernel"), pick(k_lone, "gemm_p32_kernel")
        total_w = sum(v["us_sm_weighted"] for v in k_timed.values()) if k_timed else None
        if kt is not None and gemm_flops:
            ach = gemm_flops / kt["us_w"] / 1e6
            r_gemm = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                      "traffic": _traffic("gemm_p32_kernel"), "traffic_note": "DRAM bytes per launch, ncu dram__bytes_read+write averaged over the step's launches",
                      "kernel": "gemm_p32_kernel (all launches of a forward)", "launches_per_forward": kt["n"],
                      "avg_launch_us": kt["us"] / kt["n"], "avg_launch_us_sm_weighted": kt["us_w"] / kt["n"],
                      "kernel_ms_per_step_sm_weighted": kt["us_w"] / 1e3, "ms_per_step": ms_step,
                      "algorithmic_flops_per_forward": gemm_flops, "share_of_kernel_sm_time": kt["us_w"] / total_w,
                      "executed_bf16_tflops": 3 * ach, "frac_executed_bf16": 3 * ach / peaks["bf16_sustained"],
                      "peak_source": peaks["src"] + " (bf16_tflops_sustained: kernels timed inside a long step)",
                      "timing": "CUPTI trace (torch.profiler) of the TIMED configuration: graph replays, forwards in flight; a launch's "
                                "time is its duration x (grid CTAs / SMs) — persistent grids hold that share of the GPU while other "
                                "forwards' kernels run beside them, so the sum over a step cannot exceed ms_per_step",
                      "note": "fp32-parity products = 3 bf16 MMAs each (hi*hi + hi*lo + lo*hi): frac is capped at 1/3"}
            if kl is not None:
                r_gemm["lone"] = {"launches_per_forward": kl["n"], "avg_launch_us": kl["us"] / kl["n"], "kernel_ms_per_forward": kl["us"] / 1e3,
                                  "achieved": gemm_flops / kl["us"] / 1e6, "frac": gemm_flops / kl["us"] / 1e6 / peaks["bf16_sustained"],
                                  "what": "the same launches in a lone graph-replayed forward (full grids, split-K on, nothing else on the GPU)"}
        elif "gemm_p32" in spans and gemm_flops:
            tl, n_l = spans["gemm_p32"], len(probe["gemm_p32"]) / n_probe
            ach = gemm_flops / tl / 1e12
            r_gemm = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                      "traffic": _traffic("gemm_p32_kernel"), "kernel": "gemm_p32_kernel (all launches of a forward)", "launches_per_forward": n_l,
                      "avg_launch_us": 1e6 * tl / n_l, "algorithmic_flops_per_forward": gemm_flops,
                      "timing": "CUDA events around every launch of a lone eager forward (no CUPTI trace available)"}
        if os.environ.get("EGTR_BENCH_SHAPES"):  # dev: per-shape GEMM table (warm, in-pipeline timings) on stderr
            cnt = {k: len(v) / n_probe for k, v in probe.items()}
            for k in sorted((k for k in spans if k.startswith("gemm_p32:")), key=lambda k: -spans[k]):
                m_, n_, k_ = [int(v) for v in k.split(":")[1].split("x")]
                us = 1e6 * spans[k] / cnt[k]
                print(f"{k:34s} n={cnt[k]:5.1f} {us:8.1f} us/launch {2 * m_ * n_ * k_ / us / 1e6:7.1f} TFLOP/s  total {1e6 * spans[k]:8.1f} us", file=sys.stderr)
        if os.environ.get("EGTR_BENCH_KERNELS") and k_timed:  # dev: per-kernel table of the timed configuration on stderr
            for k, v in sorted(k_timed.items(), key=lambda kv: -kv[1]["us_sm_weighted"]):
                l_ = k_lone.get(k, {})
                print(f"{k:40s} n={v['n']:6.1f} timed {v['us']:8.1f} us (sm-weighted {v['us_sm_weighted']:8.1f})  lone {l_.get('us', 0):8.1f} us", file=sys.stderr)
        stage = {k[6:]: round(1e3 * v, 3) for k, v in spans.items() if k.startswith("stage_")}
        out = {
            "metric": METRIC, "value": imgs / t_res, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split products with fp32 accumulate on tcgen05 (fp32-equivalent); fp32 elsewhere", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, cfg, H, W, Bl),
                       "parallelism": f"image-parallel x{world}, one all-gather of per-image triplet records per step" if world > 1 else "single GPU",
                       "global_batch": world * Bl, "S": S,
                       "timing": f"blocks of {args.steps} steps, each bracketed by barrier + synchronize and timed with CUDA events (max over ranks); "
                                 f"{len(blocks_res)} blocks, median reported; {conc} forwards in flight per GPU",
                       "l2": f"inputs rotate over {NIMG} distinct resident batches ({in_bytes >> 20} MiB > 126 MB L2); a forward streams > 2 GB through L2",
                       "value_leg": f"CUDA-graph replays (forward + triplet extraction), {conc} graphs with private workspaces on {conc} streams, fp32 inputs resident in HBM",
                       "e2e_leg": f"egtr_b200.serving.PipelinedRunner(input_format='u8', output='triplets'): pinned uint8 images in, triplet records out; "
                                  f"per-step H2D/D2H on copy streams, {conc} compute streams, {depth} slots",
                       "single_forward_latency_ms": latency_ms},
            "timing_spread": {"blocks": len(blocks_res), "block_s_min": min(blocks_res), "block_s_median": t_res, "block_s_max": max(blocks_res),
                              "timed_s_total": sum(blocks_res), "value_best_block": imgs / min(blocks_res)},
            "clocks": clocks,
            "e2e": {"value": imgs / t_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000 * t_e2e / args.steps, "blocks": len(blocks_e2e), "value_best_block": imgs / min(blocks_e2e),
                    "wall_s_all_blocks": t_e2e_wall, "boundary": "uint8 RGB images [B,H,W,3] in (staged on the device), triplet records out"},
            "e2e_raw": None if t_raw is None else {
                "value": imgs / t_raw, "unit": "images/s", "h2d_bytes_per_step": raw_h2d, "d2h_bytes_per_step": raw_d2h,
                "ms_per_step": 1000 * t_raw / args.steps,
                "boundary": "fp32 pixel_values + int64 pixel_mask in, logits / pred_boxes / pred_rel / pred_connectivity out (the round-1 e2e)"},
            "gpu_launches": launches,
            "output_check": {"forwards_checked": chk_n, "deviating": chk_bad, "worst_rel_err": chk_worst, "tolerance": 1e-3, "class_flips": chk_flips,
                             "what": f"{conc} forwards in flight (the timed configuration) vs the same images run alone, max-norm relative error "
                                     "over logits / boxes / pred_rel / pred_connectivity, all ranks", "error": chk_error},
            "roofline": r_gemm if r_gemm is not None else r_msda, "roofline_msda_enc": r_msda, "roofline_msda_dec": r_msda_dec, "roofline_decoder": r_dec, "roofline_relation": r_rel,
            "stage_ms_lone_eager": stage,
        }
        # ---- the reference's GPU path on this box (stand-in: the oracle port, plain torch ops, eager, on the GPU)
        out["reference_gpu"] = None
        if args.reference_gpu > 0 and world == 1:
            try:
                out["reference_gpu"] = reference_gpu(args, dev, imgs / t_e2e)
            except Exception as exc:  # noqa: BLE001
                out["reference_gpu"] = {"error": repr(exc)}
        if args.cpu_sample > 0 and world == 1:
            from oracle import egtr_oracle as orc
            orc.set_msda_impl("grid_sample")
            cfg1, sd1, px1, mask1, _ = build_case(args.workload, 1)
            torch.set_num_threads(cpu_threads())
            orc.forward(sd1, cfg1, px1, mask1)
            t0 = time.perf_counter()
            for _ in range(args.cpu_sample):
                orc.forward(sd1, cfg1, px1, mask1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": args.cpu_sample / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{args.cpu_sample} oracle forwards of one 3x{H}x{W} image after 1 warm-up"}
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def reference_gpu(args, dev, ours_e2e):
    """Stand-in for the reference's GPU FPS script (`evaluate_egtr.py:26-36`: per image H2D of pixel_values / pixel_mask, forward,
    synchronize): the oracle port — the reference's own torch ops (cuDNN convolutions, cuBLAS Linears, `grid_sample` MSDeformAttn as
    shipped, `deformable_detr.py:1096-1101`) — run eagerly on this GPU, fp32.  `/root/reference` does not travel to the GPU box."""
    import torch
    from oracle import egtr_oracle as orc
    orc.set_msda_impl("grid_sample")
    cfg, sd, px, mask, (H, W) = build_case(args.workload, 1)
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    px_h, mask_h = px.pin_memory(), mask.pin_memory()
    res = {}
    for label, tf32 in (("as_shipped", True), ("strict_fp32", False)):
        # PyTorch defaults (what the reference's script runs with): cuDNN convolutions may use TF32, matmuls do not
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.device(dev):
            def one():
                o = orc.forward(sd_d, cfg, px_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True), relation_row_chunk=cfg.num_queries)
                torch.cuda.synchronize()
                return o
            one()
            one()
            t0 = time.perf_counter()
            for _ in range(args.reference_gpu):
                one()
            dt = (time.perf_counter() - t0) / args.reference_gpu
        res[label] = {"images_per_s": 1.0 / dt, "ms_per_image": 1000 * dt, "cudnn_tf32": tf32}
    torch.backends.cudnn.allow_tf32 = True
    return {"kind": "oracle port on torch-CUDA, eager, batch 1, H2D + forward + synchronize per image (evaluate_egtr.py:26-36 loop shape)",
            **res, "unit": "images/s", "speedup_e2e_vs_as_shipped": ours_e2e / res["as_shipped"]["images_per_s"],
            "speedup_e2e_vs_strict_fp32": ours_e2e / res["strict_fp32"]["images_per_s"]}


if __name__ == "__main__":
    main()
