#!/usr/bin/env python
"""Benchmark of the DSF model-fitting hot path (contract: see the task's bench.py section).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One step = one pass of the hot path over one batch of synthetic hands: MANO forward ->
depth rasterisation (128x128) -> m2d depth loss -> backward to the 62 MANO/camera parameters.
Workload: BASELINE.json configs[2] ("large-batch render-loss fwd/bwd: batch 4096, 128x128, batch-sharded
across 1/2/4/8 B200"): ONE 4096-hand batch; under torchrun rank r owns the contiguous shard
dist.shard_bounds(4096, N, r) (strong scaling, no data-path collective; one 16-byte all-reduce of the packed
loss record per step).  The weak-scaling figure (4096 hands on every rank) is reported next to it.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hand fits/sec (MANO+raster+loss fwd/bwd)"
UNIT = "fits/s"
CROP = 128
# algorithmic bytes per hand-fit at R=128, 1 view
BYTES_PER_FIT = 2 * CROP * CROP * 4 + 779 * 3 * 4 + 21 * 3 * 4 + 62 * 4 + 62 * 4 + 4 + 60   # 141 232
# the rasteriser stage of that accounting (target in, rendered depth out, vertices + camera in)
RASTER_ALG_BYTES_PER_HAND = 2 * CROP * CROP * 4 + 779 * 3 * 4 + 60                                  # 140 480
# bytes the fused raster launch really moves per hand: verts + view record + sample grids + target in;
# normalised depth + two per-tile vertex-gradient shares + per-tile loss sums / flags out (no pix_to_face plane)
RASTER_MOVED_BYTES_PER_HAND = 779 * 3 * 4 + 20 * 4 + 2 * CROP * 4 + 24 + 2 * CROP * CROP * 4 + 2 * 779 * 3 * 4 + 24


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if "Active" == r[2 + i]
                          or r[2 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (the oracle port of the reference's CPU path)
# --------------------------------------------------------------------------------------------------
def cpu_pipeline(batch: int, seed: int = 0):
    from dsf_b200 import make_synthetic_mano, sample_fit_inputs
    from oracle import mano_oracle as mo
    from oracle import pipeline as pl

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    consts = mo.ManoConstants(make_synthetic_mano(0))
    inp = {k: torch.from_numpy(v) for k, v in sample_fit_inputs(batch, seed=seed).items()}
    from dsf_b200.synthetic import quantise_depth_mm
    with torch.no_grad():
        target, *_ = pl.render(consts, inp["params_target"], inp["center3d"], inp["cube"])
    # same workload definition as the GPU arm: the target is what a depth sensor delivers (integer mm),
    # normalised the way the reference's loader does it on the CPU (data/render_loader.py:738-745)
    mm = quantise_depth_mm(target.detach(), inp["center3d"], inp["cube"])
    target = mo.target_from_u16(mm.to(torch.int32), inp["center3d"], inp["cube"])

    def step():
        return pl.fit_step(consts, inp["params"], inp["center3d"], inp["cube"], target)

    return step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    step, cores = cpu_pipeline(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    sample = (f"{args.steps} steps x {batch} hands (of the {args.batch}-hand workload), 128x128, oracle port: torch MANO "
              f"restatement + C/OpenMP pytorch3d-0.4.0 naive rasteriser + m2d loss + autograd backward")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    per = -(-args.batch // world)
    img_mb = 2 * per * CROP * CROP * 4 / 1e6
    return {
        "workload": f"BASELINE.json configs[2]: large-batch render-loss fwd/bwd, one {args.batch}-hand batch sharded "
                    f"over {world} GPU(s), 128x128 depth, direct crop raster, NYU intrinsics, synthetic hand-shaped "
                    f"MANO (778v/1554f), pytorch3d-0.4.0 rasteriser settings (perspective_correct=False)",
        "hands_per_gpu": per, "global_batch": args.batch, "crop": CROP, "views": 1,
        "parallelism": f"batch-sharded x{world}",
        "target": "rendering of a perturbed parameter set, quantised to integer millimetres like sensor depth",
        "l2_policy": ("inputs larger than L2 (target+rendered images %.0f MB per step vs 126 MB L2)" % img_mb)
                     if img_mb > 126 else
                     ("an L2 flush (write of a 256 MB buffer) between timed steps: the shard's images (%.0f MB) "
                      "would fit the 126 MB L2" % img_mb),
    }


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def time_region(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters     # ms


def stage_times(step, iters=10):
    """CUDA-event time of each stage of the step, launched alone through the C ABI (same stream)."""
    import ctypes as C

    from dsf_b200 import _lib as L

    lib = L.lib()
    B, R, h = step.B, step.R, step.layer._handle
    s = L.stream_ptr()
    dev = step.dev
    g_img = torch.zeros(B, R, R, device=dev)
    g_verts = torch.zeros(B, 779, 3, device=dev)
    p2f = torch.empty(B, R, R, dtype=torch.int32, device=dev)
    vcam = (step.verts * step.cube[:, None] / 2 + step.center3d[:, None]).contiguous()
    prm = step.params
    p = L.DsfManoParams(prm.data_ptr(), 62, 3, prm.data_ptr() + 12, 62, 45, prm.data_ptr() + 192, 62,
                        prm.data_ptr() + 232, 62)
    gp = step.g_params
    g = L.DsfManoGrads(gp.data_ptr(), 62, gp.data_ptr() + 12, 62, gp.data_ptr() + 192, 62, gp.data_ptr() + 232, 62)
    mws = torch.empty(lib.dsf_mano_workspace_floats(B), device=dev)
    rws = torch.empty(lib.dsf_raster_loss_workspace_floats(B, R), device=dev)
    parts, totals = torch.empty(B, 2, device=dev), torch.empty(4, device=dev)
    stages = {
        "mano_forward(3 kernels)": lambda: L.check(lib.dsf_mano_forward(
            h, B, C.byref(p), 8.0, step.verts.data_ptr(), step.joints.data_ptr(), None, mws.data_ptr(), s)),
        # the rasteriser exactly as the fused step launches it: target in; depth, loss sums / totals and the
        # per-tile vertex-gradient shares out (forward + loss + raster backward in one launch)
        "raster_fwd_kernel[fused: +loss +vertex gradient]": lambda: L.check(lib.dsf_raster_loss_grad(
            h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(), step.ys.data_ptr(), R,
            step.target.data_ptr(), 0.99, 0.1, B, step.img.data_ptr(), None, parts.data_ptr(), totals.data_ptr(),
            None, rws.data_ptr(), 0, s)),
        "mano_backward(3 kernels)": lambda: L.check(lib.dsf_mano_backward(
            h, B, C.byref(p), 8.0, step.verts.data_ptr(), step.joints.data_ptr(), g_verts.data_ptr(), None,
            C.byref(g), mws.data_ptr(), s)),
        # the modular (autograd drop-in) path's kernels, not part of the fused step
        "modular: raster_fwd_kernel (image + pix_to_face)": lambda: L.check(lib.dsf_raster_forward(
            h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(), step.ys.data_ptr(), R,
            step.img.data_ptr(), p2f.data_ptr(), None, None, None, None, 0.99, None, 0, s)),
        "modular: depth_loss (2 kernels)": lambda: L.check(lib.dsf_depth_loss(
            0, B, R, step.target.data_ptr(), step.img.data_ptr(), 0.99, 0.1, parts.data_ptr(),
            totals.data_ptr(), g_img.data_ptr(), s)),
        "modular: raster_bwd_kernel": lambda: L.check(lib.dsf_raster_backward(
            h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(), step.ys.data_ptr(), R,
            p2f.data_ptr(), g_img.data_ptr(), g_verts.data_ptr(), 0, s)),
    }
    out = {}
    for name, fn in stages.items():
        fn()
        out[name] = time_region(fn, iters)
    # Fragments-materialising mode (SURVEY 8d accounting F): the rasteriser also writes the barycentrics
    # (3 x fp32 per pixel) next to depth and pix_to_face - the full "z-buffer + pix_to_face + barycentrics"
    # product of north_star piece (2).  Bytes really moved: 20 B/pixel out + the mesh in.
    bary = torch.empty(B, R, R, 3, device=dev)
    frag = lambda: L.check(lib.dsf_raster_forward(
        h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(), step.ys.data_ptr(), R,
        step.img.data_ptr(), p2f.data_ptr(), None, bary.data_ptr(), None, None, 0.99, None, 0, s))
    frag()
    out["modular: raster_fwd_kernel[fragments mode: +barycentrics]"] = time_region(frag, iters)
    return out


FUSED_KEY = "raster_fwd_kernel[fused: +loss +vertex gradient]"
FRAG_KEY = "modular: raster_fwd_kernel[fragments mode: +barycentrics]"
L2_BYTES = 126e6


class L2Flusher:
    """writes a buffer larger than L2 (timing rule: no warm-cache numbers for inputs that would fit L2)"""

    def __init__(self, dev):
        self.buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB

    def __call__(self):
        self.buf.fill_(1.0)


def timed_steps(fn, steps, flush=None):
    """ms per step on the device: back to back when the working set exceeds L2, else one event pair per step
    with an L2 flush (outside the events) in between.  The flush, the event records and the step are enqueued
    back to back (no host synchronisation inside the loop): the flush kernel runs while the host enqueues the
    step behind it, so the event pair brackets the step's device time and not the host's launch latency."""
    if flush is None:
        return time_region(fn, steps)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        flush()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in ev) / steps


def run_ours(args):
    from dsf_b200 import dist as D
    from dsf_b200 import make_synthetic_mano, sample_fit_inputs
    from dsf_b200.fit import FitStep
    from dsf_b200.mano_layer import MANO_SMPL

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: dsf_b200 has no CPU fallback")
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = D.bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa_bind else False
    G = args.batch                                  # the global batch (BASELINE configs[2]: 4096)
    lo, hi = D.shard_bounds(G, rank, world)
    B = hi - lo                                     # this rank's shard
    layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
    inp_all = sample_fit_inputs(G, seed=1000)       # every rank draws the same batch and keeps its shard
    inp = {k: v[lo:hi] for k, v in inp_all.items()}
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in inp.items()}
    # stream slices per step, measured (tools/time_steps.py): 4096 hands 1.124 / 1.084 / 1.091 ms at 1 / 2 / 4 slices,
    # 2048: 0.597 / 0.573 / 0.570, 1024: 0.339 / 0.326 / 0.321, 512: 0.200 / 0.207 (one slice is best below 1024)
    n_chunks = lambda b: args.chunks if args.chunks > 0 else (2 if b >= 4096 else (4 if b >= 1024 else 1))
    mk = lambda b: FitStep(layer, b, CROP, use_graph=not args.no_graph, chunks=n_chunks(b), keep_pix_to_face=False)
    step = mk(B)
    step.set_inputs(host["params"].to(dev), host["center3d"].to(dev), host["cube"].to(dev))
    step.render_target(host["params_target"].to(dev))
    # the target is what a depth sensor delivers: integer millimetres.  It is resident (normalised fp32) for
    # `value`; for `e2e` it travels every step either as the sensor's uint16 (normalised on the device,
    # the default) or as the loader-normalised fp32 image the reference uploads (--target-format f32).
    from dsf_b200.synthetic import quantise_depth_mm
    mm = quantise_depth_mm(step.target, step.center3d, step.cube)
    step.set_inputs(step.params, step.center3d, step.cube, mm)
    torch.cuda.synchronize()
    from dsf_b200.pcl import pack_target_rows
    mm_host = mm.cpu()
    t_pack0 = time.perf_counter()
    packed = pack_target_rows(mm_host, host["center3d"], host["cube"])
    t_pack = time.perf_counter() - t_pack0
    host_targets = {"u16": mm_host.pin_memory(), "f32": step.target.cpu().pin_memory(), "u16rows": packed}
    reducer = D.TotalsReducer(dev, G)
    flush = L2Flusher(dev) if 2 * B * CROP * CROP * 4 <= L2_BYTES else None

    def one_step():
        step.step()
        if world > 1:
            reducer.submit(step.totals)        # one 16-byte all-reduce per step, overlapped with step i+1

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        ms = timed_steps(one_step, args.steps, flush)
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reducer.finish()                       # the last collective is inside the timed region
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1) / args.steps
        if args.steps * ms < 1500:      # keep the sampler alive long enough to see clocks under load
            time_region(step.step, int(1500 / max(ms, 1e-3)) + 1)    # local work only: no collective
    if world > 1:
        torch.distributed.barrier()
    ms = D.max_over_ranks(ms, dev)
    value = G / (ms * 1e-3)
    launches = step.launches_per_step * args.steps

    # weak-scaling companion (round-1 headline): every rank owns a full 4096-hand batch
    weak = None
    if world > 1 and not args.no_weak:
        wstep = mk(G)
        wi = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(G, seed=1000 + rank).items()}
        wstep.set_inputs(wi["params"], wi["center3d"], wi["cube"])
        wstep.render_target(wi["params_target"])
        wred = D.TotalsReducer(dev, G * world)

        def wone():
            wstep.step()
            wred.submit(wstep.totals)

        for _ in range(3):
            wone()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        wms = time_region(wone, args.steps)
        wred.finish()
        torch.cuda.synchronize()
        wms = D.max_over_ranks(wms, dev)
        weak = {"hands_per_gpu": G, "global_batch": G * world, "ms_per_step": wms, "value": G * world / (wms * 1e-3),
                "unit": UNIT, "scaling": "weak"}
        del wstep, wi

    # ---- end to end through the public call with HOST buffers -------------------------------------
    # every step: H2D of that step's inputs (params, centre, cube, target depth) from pinned memory,
    # the fused step, D2H of loss + parameter gradients.  Two FitStep instances ping-pong so the copy
    # of step i+1 (copy stream) overlaps the compute of step i; all copies stay inside the timed region.
    steps_plain = [step, mk(B)]
    # row-run transport: the rasteriser's epilogue decodes the packed crop itself (dsf_fit_step_rows)
    mk_rows = lambda b: FitStep(layer, b, CROP, use_graph=not args.no_graph, chunks=n_chunks(b), keep_pix_to_face=False,
                                fuse_target_rows=True)
    steps_rows = [mk_rows(B), mk_rows(B)]
    h_g = [torch.empty(B, 62).pin_memory() for _ in range(2)]
    h_tot = [torch.empty(4).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()
    computed = [torch.cuda.Event() for _ in range(2)]
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    e2e_iters = max(4, min(args.steps, 20))
    d2h = B * 62 * 4 + 16

    def measure_e2e(fmt):
        host_target = host_targets["u16rows" if fmt == "u16rows_unpack" else fmt]
        steps2 = steps_rows if fmt == "u16rows" else steps_plain

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i])            # buffers of slot i are free again
                steps2[i].set_inputs(host["params"], host["center3d"], host["cube"], host_target)
                ready[i].record(copy_stream)

        for st_ in steps2:                                   # warm both graphs
            st_.set_inputs(host["params"], host["center3d"], host["cube"], host_target)
            st_.step()
        torch.cuda.synchronize()
        for i in range(2):
            done[i].record(main_stream)

        def e2e_run():
            upload(0)
            for it in range(e2e_iters):
                i = it & 1
                if it + 1 < e2e_iters:
                    upload(1 - i)
                main_stream.wait_event(ready[i])
                steps2[i].step()
                computed[i].record(main_stream)
                with torch.cuda.stream(d2h_stream):          # results leave on their own stream, under step i+1
                    d2h_stream.wait_event(computed[i])
                    h_g[i].copy_(steps2[i].g_params, non_blocking=True)
                    h_tot[i].copy_(steps2[i].totals, non_blocking=True)
                    done[i].record(d2h_stream)
            d2h_stream.synchronize()
            main_stream.synchronize()

        e2e_run()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        e2e_run()
        e1.record()
        torch.cuda.synchronize()
        t = D.max_over_ranks(e0.elapsed_time(e1) / e2e_iters, dev)
        h2d = B * (62 + 3 + 3) * 4 + (packed.nbytes if fmt.startswith("u16rows") else B * CROP * CROP * (2 if fmt == "u16" else 4))
        return {"value": G / (t * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "ms_per_step": t, "target_format": fmt,
                "note": "double-buffered: H2D of step i+1 and D2H of step i-1 overlap compute of step i; %.1f MB in per step on this rank; "
                        "target crop travels as %s" % (h2d / 1e6, {
                            "u16": "the sensor's uint16 mm, normalised on the device (dsf_target_from_u16)",
                            "f32": "loader-normalised fp32 (the reference's own hand-off)",
                            "u16rows": "row-run packed uint16 mm (per row only the span between the first and the last "
                                       "non-background pixel; packed by the loader with dsf_pack_u16_rows, decoded + "
                                       "normalised inside the rasteriser's epilogue: dsf_fit_step_rows, no unpack launch, "
                                       "no fp32 target plane)",
                            "u16rows_unpack": "row-run packed uint16 mm, unpacked to an fp32 plane by a separate kernel "
                                              "(dsf_target_from_u16_rows) before the step"}[fmt])}

    e2e = measure_e2e(args.target_format)
    e2e_others = {f: measure_e2e(f) for f in ("u16rows", "u16rows_unpack", "u16", "f32") if f != args.target_format}
    allv = dict(e2e_others)
    allv[args.target_format] = e2e
    # all hand-offs belong next to each other: the reference uploads the loader-normalised fp32 crop
    e2e["note"] += ("; the three hand-offs side by side: reference-style fp32 crop %.3g fits/s, sensor uint16 crop "
                    "%.3g fits/s, row-run packed uint16 crop %.3g fits/s decoded in the rasteriser (%.3g with a separate unpack "
                    "kernel) (packing is loader work, once per sample: %.2f us per hand on one host core here)"
                    % (allv["f32"]["value"], allv["u16"]["value"], allv["u16rows"]["value"],
                       allv["u16rows_unpack"]["value"], 1e6 * t_pack / B))

    if rank != 0:
        return
    # ---- roofline of the dominant kernel, timed live with CUDA events ----
    peak, peak_src = measured_peaks()
    st = stage_times(step)
    frag_ms = st.pop(FRAG_KEY)
    fused_stages = {k: v for k, v in st.items() if not k.startswith("modular:")}
    dom = max(fused_stages, key=fused_stages.get)
    dom_ms = st[FUSED_KEY]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("raster_fwd_kernel", {}).get(str(B))
        except Exception:
            traffic = None
    achieved = RASTER_ALG_BYTES_PER_HAND * B / (dom_ms * 1e-3) / 1e9
    step_gbs = BYTES_PER_FIT * G / (ms * 1e-3) / 1e9 / world
    roofline = {
        "bound": "hbm", "kernel": "raster_fwd_kernel<false> (forward + m2d loss + raster backward in one launch)",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "bytes_per_launch": RASTER_ALG_BYTES_PER_HAND * B, "bytes_per_unit": RASTER_ALG_BYTES_PER_HAND,
        "units_per_launch": B, "kernel_ms": dom_ms,
        "bytes_moved_per_launch": RASTER_MOVED_BYTES_PER_HAND * B,
        "achieved_moved": RASTER_MOVED_BYTES_PER_HAND * B / (dom_ms * 1e-3) / 1e9,
        "stage_ms": st, "slowest_stage": dom,
        "step": {"bytes_per_fit": BYTES_PER_FIT, "achieved": step_gbs, "frac": step_gbs / peak,
                 "note": "per GPU: algorithmic bytes of the whole fit step x this GPU's fits/s over the measured HBM peak"},
    }
    frag_bytes = (779 * 3 * 4 + 20 * 4 + 2 * CROP * 4 + 24 + 5 * CROP * CROP * 4) * B
    roofline["fragments_mode"] = {
        "what": "modular raster kernel also writing barycentrics (3 x fp32 / pixel): depth + pix_to_face + bary out, mesh in; no loss fusion",
        "kernel_ms": frag_ms, "bytes_per_launch": frag_bytes, "achieved": frag_bytes / (frag_ms * 1e-3) / 1e9,
        "frac": frag_bytes / (frag_ms * 1e-3) / 1e9 / peak}
    other = {}
    if world == 1 and not args.no_other_configs:
        for name, b2 in (("C1_batch128", 128), ("batch1024", 1024)):
            s2 = mk(b2)                                 # same slice policy as the headline
            i2 = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(b2, seed=77).items()}
            s2.set_inputs(i2["params"], i2["center3d"], i2["cube"])
            s2.render_target(i2["params_target"])
            for _ in range(5):
                s2.step()
            t2 = time_region(s2.step, 200)
            t2_cold = timed_steps(s2.step, 30, L2Flusher(dev) if flush is None else flush)
            st2 = stage_times(s2, iters=50)
            other[name] = {"hands": b2, "ms_per_step": t2_cold, "fits_per_s": b2 / (t2_cold * 1e-3),
                           "ms_per_step_warm_l2": t2, "fits_per_s_warm_l2": b2 / (t2 * 1e-3),
                           "launches_per_step": s2.launches_per_step,
                           "stage_ms_warm_l2": {k: round(v, 4) for k, v in st2.items() if not k.startswith("modular:")},
                           "step_hbm_frac": BYTES_PER_FIT * b2 / (t2_cold * 1e-3) / 1e9 / peak,
                           "roofline": {"bound": "hbm", "kernel": "raster_fwd_kernel<false>",
                                        "achieved": RASTER_ALG_BYTES_PER_HAND * b2 / (st2[FUSED_KEY] * 1e-3) / 1e9,
                                        "peak": peak, "unit": "GB/s",
                                        "frac": RASTER_ALG_BYTES_PER_HAND * b2 / (st2[FUSED_KEY] * 1e-3) / 1e9 / peak,
                                        "note": "kernel timed back to back (inputs L2-resident at this size)"},
                           "note": "ms_per_step: one event pair per step with an L2 flush in between (inputs would fit "
                                   "L2); *_warm_l2: back-to-back graph replays"}
        # config C4 (BASELINE.json configs[4]): self-penetration + point-to-mesh terms, batch 1024
        from dsf_b200.mesh_loss import _PointFaceDistance
        b4, P = 1024, 2048
        i4 = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(b4, seed=5).items()}
        p4 = i4["params"]
        v4, j4 = layer.get_mano_vertices(p4[:, :3], p4[:, 3:48], p4[:, 48:58], p4[:, 58:], global_scale=1 / 125)
        gen = torch.Generator(device=dev).manual_seed(0)
        idx = torch.randint(0, 778, (b4, P), device=dev, generator=gen)
        pcl = torch.gather(v4, 1, idx[..., None].expand(-1, -1, 3)) + 0.05 * torch.randn(b4, P, 3, device=dev, generator=gen)
        vg = v4.detach().requires_grad_(True)

        def icp():
            d, _ = _PointFaceDistance.apply(pcl, vg, layer.faces_int)
            d.mean().backward()
            vg.grad = None

        jg = j4.detach().requires_grad_(True)

        def coll():
            layer.calculate_coll(jg, v4.detach()).backward()
            jg.grad = None

        for fn in (icp, coll):
            fn()
        t_icp, t_coll = time_region(icp, 5), time_region(coll, 20)

        def icp_fwd():
            _PointFaceDistance.apply(pcl, v4.detach(), layer.faces_int)

        icp_fwd()
        t_icp_fwd = time_region(icp_fwd, 5)
        # FP32 roofline of the point-face scan (compute bound, bytes negligible): work counters of
        # the same launch, flops per pair as documented at dsf_point_face_stats
        from dsf_b200 import _lib as _L
        nF = int(layer.faces_int.shape[0])
        st3 = torch.zeros(4, dtype=torch.int64, device=dev)
        dd = torch.empty(b4, P, device=dev)
        ii = torch.empty(b4, P, dtype=torch.int32, device=dev)
        oo = torch.empty(b4 * (P + nF), dtype=torch.int32, device=dev)
        vv = v4.detach().contiguous()
        _L.check(_L.lib().dsf_point_face_stats(b4, P, vv.shape[1], nF, pcl.data_ptr(), vv.data_ptr(),
                                               layer.faces_int.data_ptr(), dd.data_ptr(), ii.data_ptr(), oo.data_ptr(),
                                               st3.data_ptr(), _L.stream_ptr()))
        n_cull, n_in, n_edge, n_grp = (int(x) for x in st3.tolist())
        pairs = b4 * P * nF
        flops = 16 * (n_cull + n_grp) + 56 * n_in + 120 * n_edge
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        other["C4_icp_batch1024"] = {
            "hands": b4, "points": P, "faces": nF, "ms_fwd_bwd": t_icp, "ms_fwd": t_icp_fwd,
            "point_triangle_tests_per_s": pairs / (t_icp * 1e-3),
            "pairs": {"all": pairs, "group_box_tests": n_grp, "culled_with_their_group": pairs - n_cull - n_in - n_edge,
                      "sphere_culled": n_cull, "evaluated_interior": n_in, "evaluated_edge": n_edge,
                      "evaluated_frac": (n_in + n_edge) / pairs},
            "roofline": {"bound": "fp32", "kernel": "point_face_fwd_all_kernel", "achieved": flops / (t_icp_fwd * 1e-3) / 1e12,
                         "peak": fp32_peak, "unit": "TFLOP/s", "frac": flops / (t_icp_fwd * 1e-3) / 1e12 / fp32_peak,
                         "flops_per_launch": flops,
                         "flops_per_pair": {"sphere_or_box_test": 16, "interior": 56, "edge": 120},
                         "peak_source": "148 SMs x 128 FP32 lanes x 2 (FMA) x 1.965 GHz; ms_fwd includes the point "
                                        "sort launch (< 3 % of it)",
                         "brute_force_equivalent": pairs * 120 / (t_icp_fwd * 1e-3) / 1e12},
            "note": "ICPLoss fwd+bwd; one CTA per hand stages all face records once; box hierarchy (32- and 8-face boxes) + "
                    "per-face bounding-sphere cull over spatially ordered points and faces, surviving pairs queued and "
                    "evaluated with full warps; results identical to the exhaustive scan.  The culls remove work "
                    "instead of speeding it up, so the executed-flop rate drops while the time does: compare "
                    "brute_force_equivalent (the rate an exhaustive scan would need for the same time)"}
        other["C4_coll_batch1024"] = {"hands": b4, "ms_fwd_bwd": t_coll, "hands_per_s": b4 / (t_coll * 1e-3),
                                      "note": "latency bound: one warp per hand, 9.6 KB in per hand = %.0f GB/s"
                                              % (b4 * 9600 / (t_coll * 1e-3) / 1e9)}
        # the reference's own call sequence through the drop-in classes: Render.render -> m2d loss -> backward
        # (one autograd node + the loss op), literal 640^2 -> 480x640 -> 128^2 pixel chain, batch 128 (configs[1])
        from dsf_b200.mano_layer import Render as _Render
        from dsf_b200.render_loss import m2d_loss as _m2d
        for bapi in (128, 1024):
            rapi = _Render(make_synthetic_mano(0), "nyu", (588.03, 587.07, 320.0, 240.0), (640, 480), (CROP, CROP),
                           mode="literal")
            ia = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(bapi, seed=3).items()}
            with torch.no_grad():
                tapi = rapi.render(ia["params_target"], ia["center3d"], ia["cube"])[0].clone()
            papi = ia["params"].clone().requires_grad_(True)

            def api_step():
                img_a = rapi.render(papi, ia["center3d"], ia["cube"])[0]
                _m2d(tapi, img_a).backward()
                papi.grad = None

            for _ in range(5):
                api_step()
            t_api = time_region(api_step, 50)
            other["drop_in_api_literal_batch%d" % bapi] = {
                "hands": bapi, "ms_per_step": t_api, "fits_per_s": bapi / (t_api * 1e-3),
                "note": "Render.render + m2d_loss + backward through torch autograd, eager (no CUDA graph)"}
        # config C3 (BASELINE.json configs[3]): 256x256, 3 camera views per hand, batch 512, depth + silhouette
        # (union-mask) loss through the modular autograd API: MANO once per hand, per view a rigid rotation
        # about center3d (RotationPoints), rasterise, m2d loss, backward to the 62 parameters
        from dsf_b200.mano_layer import Render, RotationPoints
        from dsf_b200.render_loss import m2d_loss
        b3, V3, R3 = 512, 3, 256
        rnd3 = Render(make_synthetic_mano(0), "nyu", (588.03, 587.07, 320.0, 240.0), (640, 480), (R3, R3), mode="direct")
        i3 = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(b3, seed=9).items()}
        rot3 = torch.tensor([[0.0, 0.0, 0.0], [0.0, 2 * np.pi / 3, 0.0], [0.0, -2 * np.pi / 3, 0.0]], device=dev).repeat(b3, 1)
        c3v, cube3v = i3["center3d"].repeat_interleave(V3, 0), i3["cube"].repeat_interleave(V3, 0)

        def c3_images(params):
            v, j = rnd3.mano_layer.get_mano_vertices(params[:, :3], params[:, 3:48], params[:, 48:58], params[:, 58:],
                                                     global_scale=1 / 125)
            vw = (v * i3["cube"][:, None] / 2 + i3["center3d"][:, None]).repeat_interleave(V3, 0)
            jw = (j * i3["cube"][:, None] / 2 + i3["center3d"][:, None]).repeat_interleave(V3, 0)
            vr, _ = RotationPoints(vw, jw, c3v, rot3)
            return rnd3._rasterize(vr, c3v, cube3v)[0]

        with torch.no_grad():
            tgt3 = c3_images(i3["params_target"]).clone()
        pg3 = i3["params"].clone().requires_grad_(True)

        def c3_step():
            m2d_loss(tgt3, c3_images(pg3)).backward()
            pg3.grad = None

        c3_step()
        t_c3 = time_region(c3_step, 10)
        c3_bytes = 6 * R3 * R3 * 4 + 10100 + 180
        other["C3_multiview_256_batch512"] = {
            "hands": b3, "views": V3, "crop": R3, "ms_fwd_bwd": t_c3, "fits_per_s": b3 / (t_c3 * 1e-3),
            "step_hbm_frac": c3_bytes * b3 / (t_c3 * 1e-3) / 1e9 / peak,
            "note": "modular autograd path (image returned, loss as a separate kernel, torch ops for the view rotation)"}
        # the same configuration through the fused multi-view step (dsf_fit_step_views, CUDA-graph replay)
        from dsf_b200.fit import MultiViewFitStep
        mv = MultiViewFitStep(layer, b3, V3, R3, use_graph=not args.no_graph, keep_pix_to_face=False)
        mv.set_inputs(i3["params"], i3["center3d"], i3["cube"], rot3.reshape(b3, V3, 3), tgt3[:, 0])
        for _ in range(3):
            mv.step()
        t_mv = time_region(mv.step, 10)
        other["C3_multiview_256_batch512"]["fused_ms"] = t_mv
        other["C3_multiview_256_batch512"]["fused_fits_per_s"] = b3 / (t_mv * 1e-3)
        other["C3_multiview_256_batch512"]["fused_step_hbm_frac"] = c3_bytes * b3 / (t_mv * 1e-3) / 1e9 / peak
        other["C3_multiview_256_batch512"]["fused_launches_per_step"] = mv.launches_per_step
        del tgt3, rnd3, mv
        # "next" rows: depth crop -> 2048-point cloud (Img2pcl) and the intersection-volume metric (I1)
        from dsf_b200.intersection import PartTopology, intersect_counts
        from dsf_b200.pcl import Img2pcl
        s2.step()
        t_pcl = time_region(lambda: Img2pcl(s2.img, CROP, s2.center3d, s2.M, s2.cube, 2048, seed=1), 50)
        other["img2pcl_batch1024"] = {"hands": 1024, "points": 2048, "ms": t_pcl, "hands_per_s": 1024 / (t_pcl * 1e-3),
                                      "GB_per_s": 1024 * (CROP * CROP * 4 + 2048 * 12) / (t_pcl * 1e-3) / 1e9}
        topo = PartTopology.synthetic_hand()
        v_mm = layer.get_mano_vertices(p4[:256, :3], p4[:256, 3:48] * 3, p4[:256, 48:58], p4[:256, 58:])[0].detach()
        intersect_counts(v_mm, topo, 2.0)
        t_iv = time_region(lambda: intersect_counts(v_mm, topo, 2.0), 5)
        other["I1_intersection_volume_batch256"] = {
            "hands": 256, "pitch_mm": 2.0, "ms": t_iv, "hands_per_s": 256 / (t_iv * 1e-3),
            "note": "15 watertight parts, 91 pairs, curled synthetic hands; float64 ray parity"}
        if not args.no_cpu_baseline:
            # CPU baseline leg for this row: the oracle restatement of the trimesh algorithm on the host cores
            from oracle import intersect_oracle as io
            t0 = time.perf_counter()
            io.intersect_vox(v_mm[:32].cpu().numpy(), topo, 2.0)
            t_iv_cpu = time.perf_counter() - t0
            other["I1_intersection_volume_batch256"]["cpu_baseline"] = {
                "value": 32 / t_iv_cpu, "unit": "hands/s", "cores": os.cpu_count(), "kind": "port",
                "sample": "32 of the 256 hands, oracle/intersect_oracle.c with OpenMP"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cstep, cores = cpu_pipeline(args.ref_batch)
        cstep()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.cpu_seconds:
            cstep()
            n += 1
        dt = time.perf_counter() - t0
        cpu = {"value": args.ref_batch * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} steps x {args.ref_batch} hands of the same workload (128x128, direct raster), "
                         f"oracle port of the reference CPU path, {dt:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches,
        "launches_per_step": step.launches_per_step, "cuda_graph": not args.no_graph, "stream_chunks": step.chunks,
        "roofline": roofline, "cpu_baseline": cpu, "loss": float(step.totals[0]), "other_configs": other,
        **{"e2e_%s_target" % f: v for f, v in e2e_others.items()}, "numa_bound": numa_bound, "weak_scaling": weak,
        "pix_to_face_plane": "not written: the rasteriser's epilogue emits the vertex gradient itself, nothing reads it",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="global batch, sharded over the GPUs")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the weak-scaling companion measurement")
    ap.add_argument("--ref-batch", type=int, default=32, help="hands per CPU reference step (bounded sample)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--chunks", type=int, default=0,
                    help="slices of the shard run on parallel streams (0 = 2 from 2048 hands per GPU, else 1)")
    ap.add_argument("--target-format", choices=["u16rows", "u16", "f32"], default="u16rows",
                    help="how the target depth crop travels host->device in the headline e2e measurement "
                         "(the other two formats are measured as well and reported next to it)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin ranks to their GPU's NUMA node")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    # libraries (NCCL's version banner, torch warnings) may write to fd 1; the contract is ONE JSON
    # line on stdout, so everything else is routed to stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout = old
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    if lines:
        print(lines[-1], flush=True)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
