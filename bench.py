#!/usr/bin/env python
"""Throughput bench of the EGTR inference hot path on B200 (contract: task prompt §④ / BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A step = one forward of the hot path over one batch of synthetic images per GPU (workload B of
BASELINE.json: VG config, 3x800x1333, N_q=200, 150 classes, 50 predicates, batch 1 per GPU) plus,
for N > 1, the single all-gather of per-image result records.  Prints ONE JSON line (rank 0).

  value : images/s with the inputs already resident in HBM; every step one CUDA-graph replay of the forward, several
          (default 3) forwards in flight on separate streams with private workspaces; CUDA events around the K timed
          steps, max over ranks; inputs rotate over 8 distinct resident images (> L2).  The latency of a single forward
          (L2 flushed before it) is reported as config.single_forward_latency_ms.
  e2e   : images/s through the public serving API (`egtr_b200.serving.PipelinedRunner`) starting from pinned HOST
          buffers: H2D of the batch and D2H of logits/boxes/pred_rel/pred_connectivity are inside the timed region.
  roofline : the kernel with the largest share of the step — the TMA-fed tcgen05 GEMM — timed with CUDA events around each
             of its launches in an eager pass (tensor bound); `roofline_msda_enc` / `roofline_msda_dec` / `roofline_relation`
             report the kernels BASELINE.json names (HBM bytes / FLOPs as defined in SURVEY.md §8d).
  cpu_baseline : the CPU oracle (a port of the reference forward, oracle/egtr_oracle.py) timed on this
                 box's host cores on a bounded sample (rank 0, N=1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (3x800x1333, N_q=200)"
WORKLOAD = "B"


def _traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu pass (profiles/r01_traffic.json, tools/gpu_traffic.sh), or None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.isfile(p):
        return None
    k = json.load(open(p)).get("kernels", {}).get(kernel)
    return k["dram_bytes_per_launch"] if k else None


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


def build_case(batch):
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config(WORKLOAD)
    H, W = WORKLOADS[WORKLOAD]["image"]
    sd = synth_state_dict(cfg, seed=0)
    px, mask = synth_images(batch, H, W, seed=1)
    return cfg, sd, px, mask, (H, W)


def run_reference(args):
    """The reference algorithm on the host CPU (oracle port; the Python reference cannot travel to the GPU box)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import egtr_oracle as orc
    orc.set_msda_impl("grid_sample")  # the reference's own CPU path for MSDeformAttn (deformable_detr.py:925-960, 1096-1101)
    cfg, sd, px, mask, (H, W) = build_case(1)
    cores = cpu_threads()
    torch.set_num_threads(cores)
    for _ in range(max(1, args.warmup)):
        orc.forward(sd, cfg, px, mask)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.forward(sd, cfg, px, mask)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} forwards of one 3x{H}x{W} image (workload {WORKLOAD}), {max(1, args.warmup)} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: VG config, 1x3x{H}x{W}, N_q={cfg.num_queries}, K={cfg.num_labels}, P={cfg.num_rel_labels}, host CPU"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=2, help="oracle forwards timed for cpu_baseline (0 = skip)")
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
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.parallel import all_gather_records, pack_records, record_layout

    Bl = args.batch_per_gpu
    cfg, sd, px, mask, (H, W) = build_case(Bl)
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(sd)
    model.cuda().eval()
    eng = model.engine()
    layout = record_layout(cfg.num_queries, cfg.num_labels, cfg.num_rel_labels)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------ leg 1: device-resident inputs, CUDA-graph replay
    # Throughput mode: `conc` forwards in flight, each a captured CUDA graph with its own workspace on its own stream
    # (at batch 1 the decoder and the deep backbone layers are latency-bound chains of small kernels; a second/third image
    # fills the SMs they leave idle).  Inputs rotate over NIMG distinct resident images (> L2 in total), so no step finds
    # its input in L2; one forward also streams > 2 GB of activations and weights through the 126 MB L2.
    conc = int(os.environ.get("EGTR_PIPE_CONCURRENCY", "8"))  # forwards in flight on separate compute streams (half-GPU grids: 8)
    depth = int(os.environ.get("EGTR_PIPE_DEPTH", str(2 * conc)))
    NIMG = 8
    px_d = [torch.roll(px, shifts=17 * i, dims=3).to(dev) for i in range(NIMG)]
    mask_d = mask.to(dev)
    in_bytes = NIMG * (px_d[0].numel() * 4 + mask_d.numel() * 8)
    runners = [eng.graph_runner(Bl, H, W, slot=i, throughput=conc > 1) for i in range(conc)]
    lone = eng.graph_runner(Bl, H, W, slot=0, throughput=False)  # latency configuration of a single forward (split-K on)
    streams = [torch.cuda.Stream() for _ in range(conc)]
    main = torch.cuda.current_stream()

    def run_resident(n):
        for st_ in streams:
            st_.wait_stream(main)
        for i in range(n):
            with torch.cuda.stream(streams[i % conc]):
                out = runners[i % conc](px_d[i % NIMG], mask_d)
                if world > 1:
                    all_gather_records(pack_records(out, layout), Bl)
        for st_ in streams:
            main.wait_stream(st_)

    run_resident(max(args.warmup, conc))
    barrier()
    # single-forward latency (one image at a time, L2 flushed before it) for reference
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
    barrier()
    _lib.call("egtr_launch_count_reset")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_resident(args.steps)
    e1.record()
    barrier()
    t_res = e0.elapsed_time(e1) / 1000.0
    # launches inside one replayed graph == launches of one eager forward; count them on an eager pass below

    # ------------------------------------------------------------ output check of the timed configuration (untimed)
    # Every forward in flight must reproduce the same image's forward run alone in the latency configuration (which the GPU tests
    # pin against the reference's golden at this size).  A throughput number whose outputs deviate is not a number.
    CHECK_KEYS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")
    chk_n, chk_bad, chk_worst, chk_error = 0, 0, 0.0, None
    try:
        want = []
        for i in range(NIMG):
            o = lone(px_d[i], mask_d)
            want.append({k: o[k].clone() for k in CHECK_KEYS})
        torch.cuda.synchronize()
        for r in range(4):
            for st_ in streams:
                st_.wait_stream(main)
            for i in range(conc):
                with torch.cuda.stream(streams[i]):
                    runners[i](px_d[(i + r) % NIMG], mask_d)
            for st_ in streams:
                main.wait_stream(st_)
            torch.cuda.synchronize()
            for i in range(conc):
                w_ = want[(i + r) % NIMG]
                e = max(float((runners[i].out[k] - w_[k]).abs().max() / w_[k].abs().max()) for k in CHECK_KEYS)
                chk_n, chk_bad, chk_worst = chk_n + 1, chk_bad + int(e > 1e-3), max(chk_worst, e)
        del want
    except Exception as exc:  # noqa: BLE001  (the check must never cost the bench line; it is reported instead)
        chk_error = repr(exc)

    # ------------------------------------------------------------ leg 2: end to end, host buffers in -> host results out
    # public API: egtr_b200.serving.PipelinedRunner(model, ...) — every step pays its own H2D (pixel_values fp32 +
    # pixel_mask int64) and D2H (logits, boxes, pred_rel, pred_connectivity); copies of neighbouring steps overlap
    # the CUDA-graph replay on separate streams.
    from egtr_b200.serving import PipelinedRunner
    px_h, mask_h = px.pin_memory(), mask.pin_memory()

    _post_bufs = {}

    def post(res):  # N > 1: the single all-gather of per-image records; each rank reads back its own images
        # static per-slot buffers (keyed by the slot's static output): no allocator traffic across the compute streams
        key = res["logits"].data_ptr()
        if key not in _post_bufs:
            rec = layout["_size"][0]
            _post_bufs[key] = (torch.empty(Bl, rec, device=dev), torch.empty(world * Bl, rec, device=dev))
        local, out = _post_bufs[key]
        torch.cat([res[f].reshape(Bl, -1) for f in ("logits", "pred_boxes", "pred_rel", "pred_connectivity")], dim=1, out=local)
        dist.all_gather_into_tensor(out, local)
        return {"records": out[rank * Bl:(rank + 1) * Bl]}

    pipe = PipelinedRunner(model, Bl, H, W, depth=depth, post=post if world > 1 else None, concurrency=conc)

    def run_e2e(n):
        pending = []
        for _ in range(n):
            pending.append(pipe.submit(px_h, mask_h))
            if len(pending) >= depth:
                pipe.collect(pending.pop(0))
        last = None
        while pending:
            last = pipe.collect(pending.pop(0))
        return last

    run_e2e(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    last = run_e2e(args.steps)
    barrier()
    e1.record()
    torch.cuda.synchronize()
    t_e2e_wall = time.perf_counter() - t0
    t_e2e = e0.elapsed_time(e1) / 1000.0
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    def step_e2e():  # eager pass through the plain model API: per-kernel probes and the launch count
        o = model(pixel_values=px_h.to(dev, non_blocking=True), pixel_mask=mask_h.to(dev, non_blocking=True),
                  output_attentions=False, output_attention_states=True, output_hidden_states=True)
        return [o[k].cpu() for k in ("logits", "pred_boxes", "pred_rel", "pred_connectivity")]

    probe, eng.probe = eng.probe, None
    # per-kernel probes and the launch count need eager launches: one extra untimed eager pass per step count
    model.use_cuda_graph = False
    # kernels per step in the timed (throughput) configuration = launches of one eager forward with the same settings
    eng.forward(px_d[0], mask_d, throughput=conc > 1)
    torch.cuda.synchronize()
    _lib.call("egtr_launch_count_reset")
    eng.forward(px_d[0], mask_d, throughput=conc > 1)
    torch.cuda.synchronize()
    launches = int(_lib.call("egtr_launch_count"))
    eng.probe = {}
    eng.probe_flops = {}
    # the probes time ONE forward's launches back to back (nothing else on the GPU): the eager model API runs the
    # single-forward configuration (split-K on), not the throughput one in which few-CTA launches rely on other images
    for _ in range(args.steps):
        torch.cuda._sleep(int(2e7))  # ~10 ms head start for the host: the probe events then bracket GPU execution, not launch gaps
        step_e2e()
    torch.cuda.synchronize()
    probe, eng.probe = eng.probe, None
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    if world > 1:
        t = torch.tensor([t_res, t_e2e, chk_worst], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e, chk_worst = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([chk_n, chk_bad], device=dev, dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        chk_n, chk_bad = int(c[0]), int(c[1])
    spans = {k: sum(a.elapsed_time(b) for a, b in v) / 1000.0 / args.steps for k, v in probe.items()}  # seconds per step
    counts = {k: len(v) // args.steps for k, v in probe.items()}

    if rank == 0:
        peaks = _peaks()
        S = sum(h * w for h, w in eng._workspace(Bl, H, W)["shapes"])
        N, P, K = cfg.num_queries, cfg.num_rel_labels, cfg.num_labels
        img_s = world * Bl * args.steps / t_res
        img_s_e2e = world * Bl * args.steps / t_e2e

        def hbm_roof(name, bytes_per_launch, n):
            if name not in spans or n == 0:
                return None
            t_launch = spans[name] / n
            ach = bytes_per_launch / t_launch / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                    "traffic": _traffic(name), "kernel": name, "avg_launch_us": 1e6 * t_launch, "launches_per_step": n,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peaks["src"] + " (hbm_gbs)"}

        # MSDeformAttn, encoder form: SURVEY.md §8d algorithmic bytes = 4*B*[S*C + Lq*M*L*P*3 + Lq*C] = 3584*S per image
        r_msda = hbm_roof("msda_enc", 3584 * S * Bl, counts.get("msda_enc", 0))
        r_msda_dec = hbm_roof("msda_dec", (1024 * S + 2560 * N) * Bl, counts.get("msda_dec", 0))
        # relation head stage (a13-a16): algorithmic HBM bytes per image: Q/K/h in 13*N*1024,
        # logits 4NK, freq-bias gather min(4N^2P, 4(K+1)^2P), outputs 4N^2(P+1), weights ~5.3 MB once
        rel_bytes = Bl * (13 * N * 1024 + 4 * N * K + min(4 * N * N * P, 4 * (K + 1) ** 2 * P) + 4 * N * N * (P + 1)) + 5.3e6
        rel_flops = Bl * (826880 * N * N + 14 * N * 131072) if P == 50 else None
        r_rel = None
        if "stage_relation" in spans:
            tl = spans["stage_relation"]
            r_rel = {"bound": "tensor", "achieved": (rel_flops or 0) / tl / 1e12, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": (rel_flops or 0) / tl / 1e12 / peaks["bf16"], "traffic": None, "kernel": "relation head stage (all its launches)",
                     "stage_us": 1e6 * tl, "hbm_GBps_on_algorithmic_bytes": rel_bytes / tl / 1e9,
                     "hbm_frac": rel_bytes / tl / 1e9 / peaks["hbm"], "note": "FLOPs as written in the reference (SURVEY.md §8d)"}
        # dominant kernel of the step: the TMA-fed tcgen05 GEMM (every launch of gemm_p32_kernel: backbone convolutions,
        # input_proj, encoder Linears, decoder value projection).  Algorithmic FLOPs = 2*M*N*K summed over its launches; each
        # product is executed as three bf16 MMAs, so the ceiling of `frac` against the bf16 peak is 1/3.
        r_gemm = None
        if "gemm_p32" in spans:
            tl, n_l = spans["gemm_p32"], counts["gemm_p32"]
            fl = eng.probe_flops.get("gemm_p32", 0) / args.steps
            ach = fl / tl / 1e12
            r_gemm = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                      "traffic": _traffic("gemm_p32_kernel"), "traffic_note": "DRAM bytes per launch, ncu dram__bytes_read+write averaged over the step's launches (profiles/r01_traffic.json)",
                      "kernel": "gemm_p32_kernel (all launches of the step)", "launches_per_step": n_l,
                      "avg_launch_us": 1e6 * tl / n_l, "algorithmic_flops_per_step": fl, "share_of_step_kernel_time": tl / sum(v for k, v in spans.items() if k.startswith("stage_")),
                      "executed_bf16_tflops": 3 * ach, "frac_executed_bf16": 3 * ach / peaks["bf16_sustained"],
                      "peak_source": peaks["src"] + " (bf16_tflops_sustained: kernels timed inside a long step)",
                      "note": "fp32-parity products = 3 bf16 MMAs each (hi*hi + hi*lo + lo*hi): frac is capped at 1/3; launches timed one forward at a time in the single-forward configuration (split-K on)"}
        if os.environ.get("EGTR_BENCH_SHAPES"):  # dev: per-shape GEMM table (warm, in-pipeline timings) on stderr
            for k in sorted((k for k in spans if k.startswith("gemm_p32:")), key=lambda k: -spans[k]):
                m_, n_, k_ = [int(v) for v in k.split(":")[1].split("x")]
                us = 1e6 * spans[k] / counts[k]
                print(f"{k:34s} n={counts[k]:3d} {us:8.1f} us/launch {2 * m_ * n_ * k_ / us / 1e6:7.1f} TFLOP/s  total {1e6 * spans[k]:8.1f} us", file=sys.stderr)
        stage = {k[6:]: round(1e3 * v, 3) for k, v in spans.items() if k.startswith("stage_")}
        # dominant kernel class of the step = the stage with the largest share; report the HBM roofline of the
        # encoder gather kernel as `roofline` (BASELINE.json's named kernel) and keep the others beside it
        out = {
            "metric": METRIC, "value": img_s, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split products with fp32 accumulate on tcgen05 (fp32-equivalent); fp32 elsewhere", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: VG config, {Bl}x3x{H}x{W} per GPU, N_q={N}, K={K}, P={P}, S={S}",
                       "parallelism": f"image-parallel x{world}, one all-gather of per-image records" if world > 1 else "single GPU",
                       "global_batch": world * Bl,
                       "timing": f"CUDA events around the {args.steps} timed steps, max over ranks; {conc} forwards in flight per GPU",
                       "l2": f"inputs rotate over {NIMG} distinct resident images ({in_bytes >> 20} MiB > 126 MB L2); a forward streams > 2 GB through L2",
                       "value_leg": f"CUDA-graph replays, {conc} graphs with private workspaces on {conc} streams, inputs resident in HBM",
                       "e2e_leg": f"egtr_b200.serving.PipelinedRunner: pinned host tensors in, host results out; per-step H2D/D2H on copy streams, {conc} compute streams, {depth} slots",
                       "single_forward_latency_ms": latency_ms},
            "clocks": clocks,
            "e2e": {"value": img_s_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000 * t_e2e / args.steps, "wall_ms_per_step": 1000 * t_e2e_wall / args.steps},
            "gpu_launches": launches,
            "output_check": {"forwards_checked": chk_n, "deviating": chk_bad, "worst_rel_err": chk_worst, "tolerance": 1e-3,
                             "what": f"{conc} forwards in flight (the timed configuration) vs the same images run alone, max-norm relative error "
                                     "over logits / boxes / pred_rel / pred_connectivity, all ranks", "error": chk_error},
            "roofline": r_gemm if r_gemm is not None else r_msda, "roofline_msda_enc": r_msda, "roofline_msda_dec": r_msda_dec, "roofline_relation": r_rel,
            "stage_ms": stage,
        }
        if args.cpu_sample > 0 and world == 1:
            from oracle import egtr_oracle as orc
            orc.set_msda_impl("grid_sample")
            cfg1, sd1, px1, mask1, _ = build_case(1)
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


if __name__ == "__main__":
    main()
