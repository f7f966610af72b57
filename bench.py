#!/usr/bin/env python3
"""bench.py -- stereo frames/s of the frontend hot path (extract L + extract R + stereo match + grids).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload euroc|kitti|hires|small]
                    [--frames F]

A step = one pass of the hot path over one batch of F synthetic stereo frames per GPU (default workload:
BASELINE.json configs[1], EuRoC MH_01-shaped 752x480 stereo, nFeatures 1000, F = 256 independent frames).
Frames are independent, so N GPUs = N ranks each owning its own F frames (weak scaling, no collective on the
data path; the only collective is the timing reduction).  Prints ONE JSON line on rank 0.

  value     whole-job frames/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public API with pinned HOST images in and keypoints / descriptors / stereo
            arrays back on the host inside the timed region
  roofline  the dominant kernel stage against the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes
            per SURVEY.md section 8(d)
  cpu_baseline / --impl reference: the reference's own FeatureExtractor.cpp / FeatureMatcher.cpp (oracle/_ref, compiled
            unmodified in the build container against a stand-in OpenCV; kind "reference"), or, when that library is
            absent, the C++ restatement oracle/cpu_ref.cpp (kind "port"), on the host cores of the same box, on a
            bounded sample of the same workload
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gtsam_vslam_b200 import synth  # noqa: E402

METRIC = "stereo_frames_per_s"
UNIT = "stereo frames/s"


def level_areas(width, height, n_levels, im_scale=1.2):
    """Level sizes exactly as FeatureExtractor::computePyramid (float32 tables, cvRound)."""
    s = np.float32(1.0)
    out = []
    for l in range(n_levels):
        if l:
            s = np.float32(s * np.float32(im_scale))
        inv = np.float32(1.0) / s
        w = int(np.rint(np.float32(width) * inv))
        h = int(np.rint(np.float32(height) * inv))
        out.append(w * h)
    return out


def algorithmic_bytes(cfg):
    """un-fused stage convention, per image and per stereo frame."""
    a = level_areas(cfg["width"], cfg["height"], cfg["n_levels"])
    A = sum(a)
    nF = cfg["n_features"]
    st = dict(pyramid=(A - a[-1]) + (A - a[0]), fast=A, blur=2 * A, describe=60 * nF)
    b_img = 5 * A - a[0] - a[-1] + 60 * nF
    return st, b_img, 2 * b_img + 13 * nF


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples whose line arrived inside [t0, t1] (the timed region); all samples if that leaves none."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for t, r in self.rows if t0 is not None and t0 <= t <= t1 + 0.03]
        rows = inside if inside else [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_frames(cfg, n, seed0):
    return synth.stereo_batch(cfg["width"], cfg["height"], cfg["disparity"], range(seed0, seed0 + n))


def cpu_arm():
    """(module, kind, description) of the CPU implementation timed as the baseline (oracle/: checker + baseline only)."""
    from oracle import cpu_ref, ref_build
    if os.path.exists(ref_build._SO):
        return ref_build, "reference", ("oracle/_ref/libref_frontend.so = the reference's own src/FeatureExtractor.cpp + "
                                        "src/FeatureMatcher.cpp compiled unmodified (-O3) against a stand-in OpenCV whose resize / FAST / "
                                        "GaussianBlur / sortIdx / fastAtan2 are scalar restatements (no SIMD OpenCV in this image)")
    return cpu_ref, "port", "oracle/cpu_ref.cpp (scalar C++ restatement; oracle/_ref was not built)"


def cpu_reference_rate(cfg, L, R, threads, shape=1):
    """frames/s of the reference's CPU path on this box."""
    mod, _, _ = cpu_arm()
    params = (cfg["n_features"], cfg["n_levels"], 1.2, 19, 31, 20, 7)
    t = time.perf_counter()
    tot, _ = mod.stereo_frames(params, cfg["fx"], cfg["baseline"], L, R, shape, threads)
    dt = time.perf_counter() - t
    return len(L) / dt, dt, tot


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores; rank 0 only."""
    if rank != 0:
        return
    _, kind, what = cpu_arm()
    cores = os.cpu_count() or 1
    sample = max(cores, min(4 * cores, 64))
    L, R = make_frames(cfg, sample, 0)
    for _ in range(args.warmup):
        cpu_reference_rate(cfg, L[:cores], R[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(cfg, L, R, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    st, b_img, b_frame = algorithmic_bytes(cfg)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_step": sample, "host_threads": cores},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} frames per step, frame-parallel over {cores} host threads; {what}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_name(args, cfg):
    return (f"{args.workload}: {cfg['width']}x{cfg['height']} stereo, nFeatures {cfg['n_features']}, {cfg['n_levels']} levels, "
            f"synthetic rectangles texture (SURVEY 8(d)), disparity {cfg['disparity']}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="euroc", choices=["euroc", "kitti", "hires", "small", "mono"])
    ap.add_argument("--frames", type=int, default=0, help="stereo frames per GPU per step (default 256; hires 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="sub-batches pipelined through separate handles in the end-to-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = synth.CONFIGS[args.workload]
    F = args.frames or (32 if args.workload == "hires" else 256)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gtsam_vslam_b200 as gvs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = cfg["width"], cfg["height"]
    L, R = make_frames(cfg, F, rank * F)  # each rank owns its own frames (shard of the global batch)
    hL = torch.from_numpy(L).pin_memory()
    hR = torch.from_numpy(R).pin_memory()
    fe_args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=F, device=local)
    feL, feR = gvs.FeatureExtractor(**fe_args), gvs.FeatureExtractor(**fe_args)
    fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    stream = torch.cuda.Stream()  # a real (non-default) stream: torch events and the library's launches share it
    torch.cuda.set_stream(stream)
    fm.set_stream(stream.cuda_stream)  # the two extractors run concurrently on their own streams, forked/joined here

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        return feL.launch_count() + feR.launch_count() + fm.launch_count()

    # ---- make the batch resident in the handles' pitched level-0 stores (outside the timed region)
    for fe, host in ((feL, hL), (feR, hR)):
        dev, pitch, stride = fe.input_layout()
        rows_per_image = stride // pitch
        store = torch.empty(0, dtype=torch.uint8, device="cuda")
        # view the library's own store through the CUDA array interface
        class _Ext:  # noqa: N801
            __cuda_array_interface__ = {"shape": (F, rows_per_image, pitch), "typestr": "|u1", "data": (dev, False), "version": 2}
        store = torch.as_tensor(_Ext(), device="cuda")
        store[:, :h, :w].copy_(host.cuda(non_blocking=True))
    torch.cuda.synchronize()

    # ---- device-resident throughput
    def step_resident():
        fm.track_stereo_batch_resident(F, w, h)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # (nvidia-smi needs ~0.1 s before its first line: started ahead of the warm-up, filtered by time below)
    for _ in range(args.warmup):
        step_resident()
    fm.sync()
    barrier()
    t_region0 = time.perf_counter()
    l0 = launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    fm.sync()
    t_region1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1) / args.steps
    n_launch = (launches() - l0) // args.steps
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None

    # ---- per-stage device times: same work, same region length, but every kernel in order on ONE stream (timing mode
    # of the library) so that each stage's CUDA-event time is that kernel alone, not its overlap with the other side
    feL.set_stream(stream.cuda_stream); feR.set_stream(stream.cuda_stream)
    feL.enable_timing(True); feR.enable_timing(True)
    stage_acc = {k: 0.0 for k in ("pyramid", "fast", "select", "blur", "describe", "stereo", "grid")}
    prof_steps = max(2, min(args.steps, 5))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_resident(); fm.sync()
    p0.record(stream)
    for _ in range(prof_steps):
        step_resident()
        fm.sync()
        tl, tr = feL.stage_times(), feR.stage_times()
        for k in ("pyramid", "fast", "select", "blur", "describe"):
            stage_acc[k] += (getattr(tl, k + "_ms") + getattr(tr, k + "_ms")) / prof_steps
        stage_acc["stereo"] += tl.stereo_ms / prof_steps
        stage_acc["grid"] += tl.grid_ms / prof_steps
    p1.record(stream)
    torch.cuda.synchronize()
    serial_ms = sum(stage_acc.values())
    feL.enable_timing(False); feR.enable_timing(False)
    feL.set_stream(None); feR.set_stream(None)

    # ---- end to end through the public API: pinned host images in, all results back on the host
    # The public API is used the way a throughput caller would: gvs_pipeline (include/gvs.h; api.StereoPipeline) with `chunks`
    # slots, the batch cut into `chunks` tickets.  All tickets are submitted first and collected in order, so that the
    # host->device copy of ticket k+1 and the device->host copy of ticket k-1 overlap the kernels of ticket k.
    chunks = max(1, min(args.e2e_chunks, F))
    while F % chunks:
        chunks -= 1
    Fc = F // chunks
    pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory().numpy()
    pipe = gvs.StereoPipeline(cfg["fx"], cfg["baseline"], width=w, height=h, frames_per_slot=Fc, n_slots=chunks,
                              nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], device=local)
    cap = Fc * feL.cap_per_image
    bufs = [dict(kl=pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE), kr=pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE),
                 dl=pin(cap * 32, torch.uint8), dr=pin(cap * 32, torch.uint8), ri=pin(cap, torch.int32),
                 li=pin(cap, torch.int32), de=pin(cap, torch.float32), cl=pin(cap, torch.uint8)) for _ in range(chunks)]
    npL, npR = hL.numpy(), hR.numpy()

    def run_e2e(steps):
        """`steps` steps (= steps * chunks tickets) streamed through the pipeline: a ticket is collected when its slot is needed
        again, so the upload of the next tickets (also the next step's) overlaps the kernels and downloads in flight.  Every
        step's images are copied from pinned host memory and every result is read back inside the timed region."""
        pending = []       # (ticket, chunk index)
        nbytes, tl, tr, tm = 0, 0, 0, 0

        def collect():
            nonlocal nbytes, tl, tr, tm
            t, c = pending.pop(0)
            b = bufs[t % chunks]
            offL, offR = pipe.wait(t, Fc)
            nl, nr = int(offL[-1]), int(offR[-1])
            pipe.download_into(t, b["kl"], b["dl"], b["kr"], b["dr"], b["ri"], b["li"], b["de"], b["cl"])
            nbytes += 2 * 4 * (Fc + 1) + (nl + nr) * 60 + nl * 9 + nr * 4
            tl += nl; tr += nr; tm += int((b["ri"][:nl] >= 0).sum())
            pipe.release(t)

        for _ in range(steps):
            for c in range(chunks):
                if len(pending) == chunks:
                    collect()
                pending.append((pipe.submit(npL[c * Fc:(c + 1) * Fc], npR[c * Fc:(c + 1) * Fc]), c))
        while pending:
            collect()
        return nbytes // steps, (tl // steps, tr // steps, tm // steps)

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    d2h, counts = run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    st_bytes, b_img, b_frame = algorithmic_bytes(cfg)
    peak, peak_src = measured_peak()
    dom = max(("pyramid", "fast", "blur"), key=lambda k: stage_acc[k])
    dom_ms = stage_acc[dom]
    achieved = st_bytes[dom] * 2 * F / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    value = world * F / (dev_ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_gpu_per_step": F, "global_frames_per_step": world * F,
                   "parallelism": f"frame-sharded x{world}, no collective on the data path",
                   "l2": f"inputs larger than L2: {2 * F * w * h / 1e6:.0f} MB of level-0 images per GPU per step",
                   "keypoints_per_frame": [counts[0] / F, counts[1] / F], "stereo_matches_per_frame": counts[2] / F},
        "e2e": {"value": world * F / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(2 * F * w * h), "d2h_bytes_per_step": int(d2h),
                "how": f"pinned host images -> gvs_pipeline_submit / wait / download ({chunks} tickets of {Fc} frames per step streamed over "
                       f"{chunks} slots, a ticket is collected when its slot is needed again) -> keypoints, descriptors, stereo arrays back in "
                       f"pinned host memory; wall clock around all {e2e_steps} steps including the final drain"},
        "gpu_launches": int(n_launch),
        "stage_ms": {k: round(v, 4) for k, v in stage_acc.items()},
        "stage_ms_note": f"per step, left+right, each kernel alone on one stream (sum {serial_ms:.3f} ms); the timed region overlaps "
                         "the two sides and the blur on separate streams",
        "roofline": {"bound": "hbm", "kernel": {"pyramid": "pyr_down_kernel (7 launches)", "fast": "fast_cells_kernel",
                                                 "blur": "blur7_kernel"}[dom],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": st_bytes[dom] * F,
                     "launch_ms": dom_ms / 2,
                     "note": "HBM is the roofline SURVEY 8(d) prescribes for this stage; ncu (profiles/README.md) shows the kernel bound by the "
                             "integer ALU pipe instead (ALU 68 %, issue slots 71 %, DRAM 2.9 %, DRAM traffic = algorithmic bytes)",
                     "step": {"achieved": b_frame * F / (dev_ms * 1e-3) / 1e9, "frac": b_frame * F / (dev_ms * 1e-3) / 1e9 / peak,
                              "bytes_per_stereo_frame": b_frame}},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = F  # the whole batch: about 10-30 core-seconds of CPU work
        v_all, dt_all, _ = cpu_reference_rate(cfg, L[:sample], R[:sample], cores, shape=1)
        s2 = min(sample, 6)
        v_ref, dt_ref, _ = cpu_reference_rate(cfg, L[:s2], R[:s2], 2, shape=0)
        _, kind, what = cpu_arm()
        out["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"first {sample} frames of the same batch, frame-parallel over {cores} host threads "
                                         f"({dt_all:.1f} s); {what}",
                               "reference_shaped_2_threads": {"value": v_ref, "sample": f"{s2} frames, 2 extraction threads + serial matcher"}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
