#!/usr/bin/env python3
"""bench.py -- stereo frames/s of the frontend hot path: extract L + extract R + stereo match + key grids + temporal match.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload euroc|kitti|hires|small|mono]
                    [--frames F] [--no-temporal] [--also kitti4096,hires] [--e2e-chunks C] [--e2e-slots S]

A step = one pass of the hot path over one batch of F synthetic stereo frames per GPU (default workload: BASELINE.json
configs[1], EuRoC MH_01-shaped 752x480 stereo, nFeatures 1000, F = 256 consecutive frames of a synthetic sequence).  The
temporal stage tracks every frame against its predecessor in the batch exactly as the reference's tracking thread does after
its first frame: map points from the predecessor's stereo keys (FeatureTracker::initializeMap), removeOutOfFrameMPs with the
predicted pose, FeatureMatcher::matchByProjectionRPred(rad 10) (src/FeatureTracker.cpp:72-112, 910-940, 1176-1205).
Frames only depend on their predecessor, so N GPUs = N ranks each owning its own F-frame run (weak scaling, no collective on the
data path; the only collective is the timing reduction).  Prints ONE JSON line on rank 0.

  value     whole-job frames/s with the batch already resident in HBM (CUDA events, max over ranks), exactly K steps
  sustained the same loop repeated for >= 2 s (clocks sampled over that region)
  e2e       the same through the public API with pinned HOST images in and keypoints / descriptors / stereo arrays /
            temporal matches back on the host inside the timed region (gvs_pipeline: C tickets per step over S slots)
  roofline  the dominant extraction kernel against the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes per
            SURVEY.md section 8(d); alu_roofline: the same kernel against the integer-ALU issue rate that actually bounds it
  cpu_baseline / --impl reference: the reference's own FeatureExtractor.cpp / FeatureMatcher.cpp (oracle/_ref, compiled
            unmodified in the build container against a stand-in OpenCV; kind "reference"), or, when that library is
            absent, the C++ restatement oracle/cpu_ref.cpp (kind "port"), on the host cores of the same box, same frames,
            same temporal step
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
TEMPORAL_RAD = 10.0   # src/FeatureTracker.cpp:1192


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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


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

    def window(self, t0, t1):
        """Summary of the samples whose line arrived inside [t0, t1]; all samples if that leaves none."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.03]
        rows = inside if inside else [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()


def make_sequence(cfg, n, seed):
    """n consecutive stereo frames of one synthetic sequence + the camera motion that explains them (each frame is tracked
    against its predecessor; frame 0 against frame 1, so that every frame of the batch carries the same work)."""
    L, R = synth.stereo_sequence(cfg["width"], cfg["height"], cfg["disparity"], n, seed=seed)
    prev = [1 if n > 1 else -1] + list(range(n - 1))
    return L, R, synth.sequence_motion(cfg, n, prev)


def motion_slice(mo, a, b):
    """The motion of frames [a, b) as a self-contained ticket (predecessors re-indexed; a frame whose predecessor lies outside
    the ticket is tracked against its successor)."""
    out = dict(mo)
    prev = np.array(mo["prev"][a:b]) - a
    pose_prev = mo["pose_prev"][a:b].copy()
    for i in range(b - a):
        if prev[i] < 0 or prev[i] >= b - a:
            prev[i] = i + 1 if i + 1 < b - a else -1
            if prev[i] >= 0:
                pose_prev[i] = mo["pose"][a + prev[i]]
    out.update(prev=prev.astype(np.int32), pose_prev=pose_prev, to_camera=mo["to_camera"][a:b], to_camera_r=mo["to_camera_r"][a:b])
    return out


def cpu_arm():
    """(module, kind, description) of the CPU implementation timed as the baseline (oracle/: checker + baseline only)."""
    from oracle import cpu_ref, ref_build
    if os.path.exists(ref_build._SO):
        return ref_build, "reference", ("oracle/_ref/libref_frontend.so = the reference's own src/FeatureExtractor.cpp + "
                                        "src/FeatureMatcher.cpp compiled unmodified (-O3 -march=x86-64-v3 -ffp-contract=off; the reference's own "
                                        "flags are -O3 -march=native) against a stand-in OpenCV whose resize / FAST / GaussianBlur / sortIdx / "
                                        "fastAtan2 are scalar restatements (no SIMD OpenCV in this image); map-point creation and projection of "
                                        "the temporal step are restated (FeatureTracker.cpp needs GTSAM)")
    return cpu_ref, "port", "oracle/cpu_ref.cpp (scalar C++ restatement; oracle/_ref was not built)"


def cpu_reference_rate(cfg, L, R, threads, shape=1, motion=None):
    """frames/s of the reference's CPU path on this box (same frames, same temporal step when `motion` is given)."""
    mod, _, _ = cpu_arm()
    params = (cfg["n_features"], cfg["n_levels"], 1.2, 19, 31, 20, 7)
    t = time.perf_counter()
    tot, _, nt = mod.track_frames(params, cfg["fx"], cfg["baseline"], L, R, shape, threads, motion, TEMPORAL_RAD)
    dt = time.perf_counter() - t
    return len(L) / dt, dt, tot, int(nt.sum())


def workload_name(args, cfg):
    stages = "extract+stereo+grids" + ("" if args.no_temporal else "+temporal(matchByProjectionRPred rad 10 vs. the previous frame)")
    return (f"{args.workload}: {cfg['width']}x{cfg['height']} stereo, nFeatures {cfg['n_features']}, {cfg['n_levels']} levels, {stages}; "
            f"synthetic rectangles texture (SURVEY 8(d)) as a translating sequence, disparity {cfg['disparity']}")


def run_reference(args, cfg, F, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores; rank 0 only."""
    if rank != 0:
        return
    _, kind, what = cpu_arm()
    cores = os.cpu_count() or 1
    L, R, mo = make_sequence(cfg, F, 1000)
    motion = None if args.no_temporal else mo
    w = min(F, 2 * cores)
    for _ in range(args.warmup):
        cpu_reference_rate(cfg, L[:w], R[:w], cores, motion=None if motion is None else motion_slice(mo, 0, w))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(cfg, L, R, cores, motion=motion)
    dt = (time.perf_counter() - t0) / args.steps
    v = F / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_gpu_per_step": F, "global_frames_per_step": F, "host_threads": cores},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{F} frames per step (the whole batch of our arm), frame-parallel over {cores} host threads; {what}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


class Arm:
    """Our arm on one GPU: handles, resident batch, timed loops."""

    def __init__(self, args, cfg, F, local, seed):
        import torch
        import gtsam_vslam_b200 as gvs
        self.torch, self.gvs, self.args, self.cfg, self.F, self.local = torch, gvs, args, cfg, F, local
        w, h = cfg["width"], cfg["height"]
        self.L, self.R, self.mo = make_sequence(cfg, F, seed)   # each rank owns its own run of frames (shard of the global batch)
        self.hL = torch.from_numpy(self.L).pin_memory()
        self.hR = torch.from_numpy(self.R).pin_memory()
        fe_args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=F, device=local)
        self.feL, self.feR = gvs.FeatureExtractor(**fe_args), gvs.FeatureExtractor(**fe_args)
        self.fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], self.feL, self.feR, h)
        self.stream = torch.cuda.Stream()  # a real (non-default) stream: torch events and the library's launches share it
        torch.cuda.set_stream(self.stream)
        self.fm.set_stream(self.stream.cuda_stream)  # the two extractors run concurrently on their own streams, forked/joined here
        # make the batch resident in the handles' pitched level-0 stores (outside the timed region)
        for fe, host in ((self.feL, self.hL), (self.feR, self.hR)):
            dev, pitch, stride = fe.input_layout()
            rows_per_image = stride // pitch

            class _Ext:  # noqa: N801  (view of the library's own store through the CUDA array interface)
                __cuda_array_interface__ = {"shape": (F, rows_per_image, pitch), "typestr": "|u1", "data": (dev, False), "version": 2}
            store = torch.as_tensor(_Ext(), device="cuda")
            store[:, :h, :w].copy_(host.cuda(non_blocking=True))
        torch.cuda.synchronize()

    def launches(self):
        return self.feL.launch_count() + self.feR.launch_count() + self.fm.launch_count()

    def step_resident(self):
        cfg, mo = self.cfg, self.mo
        self.fm.track_stereo_batch_resident(self.F, cfg["width"], cfg["height"])
        if not self.args.no_temporal:
            self.fm.track_temporal_batch(self.F, mo["prev"], mo["pose_prev"], mo["to_camera"], mo["to_camera_r"], mo["fx"], mo["fy"],
                                         mo["cx"], mo["cy"], cfg["width"], cfg["height"], TEMPORAL_RAD)

    def timed(self, steps):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            self.step_resident()
        e1.record(self.stream)
        self.fm.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def stage_times(self, prof_steps):
        """Per-stage device times: same work, but every kernel in order on ONE stream (timing mode of the library) so that each
        stage's CUDA-event time is that kernel alone, not its overlap with the other side."""
        s = self.stream.cuda_stream
        self.feL.set_stream(s); self.feR.set_stream(s)
        self.feL.enable_timing(True); self.feR.enable_timing(True)
        acc = {k: 0.0 for k in ("pyramid", "fast", "select", "blur", "describe", "stereo", "grid", "temporal")}
        self.step_resident(); self.fm.sync()
        for _ in range(prof_steps):
            self.step_resident()
            self.fm.sync()
            tl, tr = self.feL.stage_times(), self.feR.stage_times()
            for k in ("pyramid", "fast", "select", "blur", "describe"):
                acc[k] += (getattr(tl, k + "_ms") + getattr(tr, k + "_ms")) / prof_steps
            acc["stereo"] += tl.stereo_ms / prof_steps
            acc["grid"] += tl.grid_ms / prof_steps
            if not self.args.no_temporal:
                acc["temporal"] += tl.temporal_ms / prof_steps
        self.feL.enable_timing(False); self.feR.enable_timing(False)
        self.feL.set_stream(None); self.feR.set_stream(None)
        return acc

    def make_e2e(self, chunks, slots=0):
        """End to end through the public API the way a throughput caller uses it: gvs_pipeline (include/gvs.h; api.StereoPipeline)
        with `chunks` slots, the batch cut into `chunks` tickets.  A ticket is collected when its slot is needed again, so the
        upload of the next tickets (also the next step's) overlaps the kernels and downloads in flight."""
        torch, gvs, cfg, F = self.torch, self.gvs, self.cfg, self.F
        w, h = cfg["width"], cfg["height"]
        chunks = max(1, min(chunks, F))
        while F % chunks:
            chunks -= 1
        Fc = F // chunks
        slots = max(slots, chunks) if slots else chunks   # tickets in flight (>= tickets per step: the queue may run into the next step)
        pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory().numpy()
        pipe = gvs.StereoPipeline(cfg["fx"], cfg["baseline"], width=w, height=h, frames_per_slot=Fc, n_slots=slots,
                                  nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], device=self.local)
        cap = Fc * self.feL.cap_per_image
        bufs = [dict(kl=pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE), kr=pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE),
                     dl=pin(cap * 32, torch.uint8), dr=pin(cap * 32, torch.uint8), ri=pin(cap, torch.int32),
                     li=pin(cap, torch.int32), de=pin(cap, torch.float32), cl=pin(cap, torch.uint8),
                     to=pin(Fc + 1, torch.int32), tk=pin(cap, torch.int32), tm=pin(cap * 2, torch.int32), tn=pin(Fc, torch.int32))
                for _ in range(slots)]
        npL, npR = self.hL.numpy(), self.hR.numpy()
        motions = [None if self.args.no_temporal else motion_slice(self.mo, c * Fc, (c + 1) * Fc) for c in range(chunks)]

        def run(steps):
            pending = []       # tickets in flight
            nbytes, tl, tr, ts, tt = 0, 0, 0, 0, 0

            def collect():
                nonlocal nbytes, tl, tr, ts, tt
                t = pending.pop(0)
                b = bufs[t % slots]
                offL, offR = pipe.wait(t, Fc)
                nl, nr = int(offL[-1]), int(offR[-1])
                pipe.download_into(t, b["kl"], b["dl"], b["kr"], b["dr"], b["ri"], b["li"], b["de"], b["cl"])
                nbytes += 2 * 4 * (Fc + 1) + (nl + nr) * 60 + nl * 9 + nr * 4
                if not self.args.no_temporal:
                    pipe.download_temporal_into(t, b["to"], b["tk"], b["tm"], b["tn"])
                    M = int(b["to"][Fc])
                    nbytes += 4 * (Fc + 1) + 4 * Fc + M * 12
                    tt += int(b["tn"].sum())
                tl += nl; tr += nr; ts += int((b["ri"][:nl] >= 0).sum())
                pipe.release(t)

            for _ in range(steps):
                for c in range(chunks):
                    if len(pending) == slots:
                        collect()
                    if motions[c] is None:
                        pending.append(pipe.submit(npL[c * Fc:(c + 1) * Fc], npR[c * Fc:(c + 1) * Fc]))
                    else:
                        pending.append(pipe.submit_temporal(npL[c * Fc:(c + 1) * Fc], npR[c * Fc:(c + 1) * Fc], motions[c], TEMPORAL_RAD))
            while pending:
                collect()
            return nbytes // steps, (tl // steps, tr // steps, ts // steps, tt // steps)

        return run, chunks, Fc


def alu_roofline(cfg, F, fast_launch_ms, sm_mhz, instr_per_px):
    """The FAST kernel is bound by the integer-ALU pipe, not HBM (profiles/): issue rate = one warp instruction per two cycles per
    scheduler (4 per SM, 148 SMs).  instr_per_px = ALU-pipe warp instructions per pixel x 32 lanes from the ncu capture named in
    profiles/README.md; achieved = pixels per second of one launch."""
    A = sum(level_areas(cfg["width"], cfg["height"], cfg["n_levels"]))
    px_per_s = A * F / (fast_launch_ms * 1e-3)
    peak_instr = 148 * 4 * 0.5 * sm_mhz * 1e6                     # ALU-pipe warp instructions per second
    peak_px = peak_instr * 32 / instr_per_px if instr_per_px else None   # pixels per second at a saturated ALU pipe
    return {"bound": "alu_pipe", "achieved": px_per_s / 1e9, "peak": peak_px / 1e9 if peak_px else None, "unit": "Gpixel/s",
            "frac": px_per_s / peak_px if peak_px else None, "alu_thread_instr_per_pixel": instr_per_px,
            "note": "peak = 148 SMs x 4 schedulers x 0.5 ALU-pipe warp instructions per clock x 32 lanes / (ALU-pipe thread instructions per "
                    "pixel of this kernel, from profiles/); a fraction near 1 means the pipe is saturated and only fewer instructions per "
                    "pixel can make the kernel faster"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="euroc", choices=["euroc", "kitti", "hires", "small", "mono"])
    ap.add_argument("--frames", type=int, default=0, help="stereo frames per GPU per step (default 256; hires 32)")
    ap.add_argument("--no-temporal", action="store_true", help="extract + stereo + grids only (round-1 step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=2, help="sub-batches (tickets) per step in the end-to-end leg")
    ap.add_argument("--e2e-slots", type=int, default=6, help="pipeline slots = tickets in flight (at least one per ticket of a step; measured on 256 EuRoC frames, tickets per step / slots: 2/6 48.9 k frames/s, 1/4 48.9 k, 3/6 48.6 k, 2/4 48.1 k, 1/3 47.7 k, 2/8 46.7 k, 4/4 45.4 k, 1/2 40.9 k)")
    ap.add_argument("--also", default="auto", help="extra workloads reported in 'other_workloads' (comma list of kitti4096,hires; auto = both when "
                                                   "--gpus > 1, none otherwise)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = synth.CONFIGS[args.workload]
    F = args.frames or (32 if args.workload == "hires" else 256)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, F, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pin_rank_to_numa_node(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    arm = Arm(args, cfg, F, local, seed=1000 + rank)
    w, h = cfg["width"], cfg["height"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # (nvidia-smi needs ~0.1 s before its first line: started ahead of the warm-up, windows are cut by time below)

    # ---- device-resident throughput: exactly K steps
    for _ in range(args.warmup):
        arm.step_resident()
    arm.fm.sync()
    barrier()
    t_region0 = time.perf_counter()
    l0 = arm.launches()
    dev_ms = arm.timed(args.steps)
    barrier()
    t_region1 = time.perf_counter()
    n_launch = (arm.launches() - l0) // args.steps
    clocks = sampler.window(t_region0, t_region1) if rank == 0 else None

    # ---- sustained: the same loop for >= 2 s
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2200.0 / max(dev_ms, 1e-3)) + 1)
        barrier()
        ts0 = time.perf_counter()
        sus_ms = arm.timed(n_sus)
        barrier()
        ts1 = time.perf_counter()
        sus_ms = max_over_ranks([sus_ms])[0]
        if rank == 0:
            sustained = {"value": world * F / (sus_ms * 1e-3), "unit": UNIT, "ms_per_step": sus_ms, "steps": n_sus,
                         "seconds": n_sus * sus_ms * 1e-3, "clocks": sampler.window(ts0, ts1)}

    # ---- per-stage device times
    stage_acc = arm.stage_times(max(2, min(args.steps, 5)))
    serial_ms = sum(stage_acc.values())

    # ---- end to end through the public API: pinned host images in, all results back on the host
    run_e2e, chunks, Fc = arm.make_e2e(args.e2e_chunks, args.e2e_slots)
    run_e2e(2 + (max(args.e2e_slots, chunks) + chunks - 1) // chunks)   # warm-up: every slot has been used (first use allocates)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(args.steps, 3)
    d2h, counts = run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    feed = feed_ceiling(torch, 2 * F * w * h, d2h, steps=5)   # bare pinned copies of the same bytes, all ranks at the same time
    if rank == 0:
        sampler.stop()

    dev_ms, e2e_ms, feed_ms = max_over_ranks([dev_ms, e2e_ms, feed])

    # ---- the other named configurations (BASELINE.json configs 3 and 5), strong / weak scaled next to the default line
    also = args.also
    if also == "auto":
        also = "kitti4096,hires" if world > 1 else ""
    others = {}
    tma = arm.feL.tma_stages()
    del arm
    torch.cuda.empty_cache()
    for name in [a for a in also.split(",") if a]:
        others[name] = other_workload(args, name, rank, world, local, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    st_bytes, b_img, b_frame = algorithmic_bytes(cfg)
    peak, peak_src, sm_max = measured_peaks()
    dom = max(("pyramid", "fast", "blur"), key=lambda k: stage_acc[k])
    dom_ms = stage_acc[dom]
    achieved = st_bytes[dom] * 2 * F / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    prof = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp)).get(args.workload, {})
        except Exception:
            prof = {}
    traffic = prof.get(dom)
    if isinstance(traffic, dict):   # {"bytes_per_image": ..}: per launch = x F
        traffic = traffic.get("bytes_per_image", 0) * F or None
    value = world * F / (dev_ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_gpu_per_step": F, "global_frames_per_step": world * F,
                   "parallelism": f"frame-sharded x{world}, no collective on the data path",
                   "l2": f"inputs larger than L2: {2 * F * w * h / 1e6:.0f} MB of level-0 images per GPU per step",
                   "keypoints_per_frame": [counts[0] / F, counts[1] / F], "stereo_matches_per_frame": counts[2] / F,
                   "temporal_matches_per_frame": counts[3] / F},
        "sustained": sustained,
        "e2e": {"value": world * F / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": int(2 * F * w * h), "d2h_bytes_per_step": int(d2h),
                "feed_ceiling": {"value": world * F / (feed_ms * 1e-3), "unit": UNIT, "ms_per_step": feed_ms,
                                 "how": "the same H2D + D2H bytes per step as bare pinned cudaMemcpyAsync on two streams, no kernels, all ranks "
                                        "at the same time (max over ranks): what the host link gives at this N"},
                "how": f"pinned host images -> gvs_pipeline_submit_temporal / wait / download ({chunks} tickets of {Fc} frames per step streamed over "
                       f"{max(args.e2e_slots, chunks) if args.e2e_slots else chunks} slots, a ticket is collected when its slot is needed again) -> keypoints, descriptors, stereo arrays and temporal "
                       f"matches back in pinned host memory; wall clock around all {e2e_steps} steps including the final drain"},
        "gpu_launches": int(n_launch),
        "tma_stages": {"pyramid": bool(tma & 1), "fast": bool(tma & 2)},
        "stage_ms": {k: round(v, 4) for k, v in stage_acc.items()},
        "stage_ms_note": f"per step, left+right, each kernel alone on one stream (sum {serial_ms:.3f} ms); the timed region overlaps "
                         "the two sides on separate streams",
        "roofline": {"bound": "hbm", "kernel": {"pyramid": "pyr_fused_kernel", "fast": "fast_cells_kernel", "blur": "blur7_kernel"}[dom],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": st_bytes[dom] * F,
                     "launch_ms": dom_ms / 2,
                     "note": "HBM is the roofline SURVEY 8(d) prescribes for this stage; ncu (profiles/README.md) shows the kernel bound by the "
                             "integer ALU pipe instead: see alu_roofline",
                     "pyramid": {"achieved": st_bytes["pyramid"] * 2 * F / (stage_acc["pyramid"] * 1e-3) / 1e9 if stage_acc["pyramid"] > 0 else None,
                                 "frac": st_bytes["pyramid"] * 2 * F / (stage_acc["pyramid"] * 1e-3) / 1e9 / peak if stage_acc["pyramid"] > 0 else None,
                                 "ms_per_step": stage_acc["pyramid"],
                                 "traffic_per_side": prof.get("pyramid"),   # DRAM bytes of the fused launches of ONE side (ncu, profiles/traffic.json)
                                 "algorithmic_bytes_per_side": int(st_bytes["pyramid"] * F)},
                     "step": {"achieved": b_frame * F / (dev_ms * 1e-3) / 1e9, "frac": b_frame * F / (dev_ms * 1e-3) / 1e9 / peak,
                              "bytes_per_stereo_frame": b_frame}},
        "alu_roofline": alu_roofline(cfg, F, stage_acc["fast"] / 2, (clocks or {}).get("sm_mhz") or sm_max, prof.get("fast_alu_thread_instr_per_pixel")),
        "clocks": clocks,
    }
    if others:
        out["other_workloads"] = others
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        L, R, mo = make_sequence(cfg, F, 1000)
        motion = None if args.no_temporal else mo
        v_all, dt_all, _, nt = cpu_reference_rate(cfg, L, R, cores, shape=1, motion=motion)
        s2 = min(F, 6)
        v_ref, dt_ref, _, _ = cpu_reference_rate(cfg, L[:s2], R[:s2], 2, shape=0, motion=None if motion is None else motion_slice(mo, 0, s2))
        _, kind, what = cpu_arm()
        out["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"the same {F} frames and temporal step, frame-parallel over {cores} host threads "
                                         f"({dt_all:.1f} s, {nt / F:.0f} temporal matches per frame); {what}",
                               "reference_shaped_2_threads": {"value": v_ref, "sample": f"{s2} frames, 2 extraction threads + serial matcher"}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def pin_rank_to_numa_node(local):
    """Keep the rank's host threads (submit / wait / download loop, pinned buffers' first touch) on the CPUs of the GPU's own NUMA
    node, spread over the node's cores by rank so that eight ranks do not share the same few cores."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, n)
        cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            per = max(1, len(cpus) // max(world, 1))
            mine = cpus[(local * per) % len(cpus):][:per] or cpus
            os.sched_setaffinity(0, mine)
    except Exception:
        pass


def feed_ceiling(torch, h2d_bytes, d2h_bytes, steps):
    """ms per step of the bare host<->device feed: the same bytes, pinned memory, H2D and D2H on two streams, no kernels."""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps


def other_workload(args, name, rank, world, local, barrier, max_over_ranks):
    """BASELINE.json configs 3 and 5 next to the default line (same timing rules, shorter):
       kitti4096  a 4096-frame KITTI-shaped batch, STRONG scaled: the 4096 frames are split over the ranks (extraction + stereo)
       hires      2048x1536, 8000 features, 32 frames per GPU, weak scaled (extraction + stereo)"""
    import argparse as _ap
    a2 = _ap.Namespace(**vars(args))
    a2.no_temporal = True
    if name == "kitti4096":
        cfg, total, chunk = synth.CONFIGS["kitti"], 4096, 256
        per_rank = total // world
        a2.workload = "kitti"
        arm = Arm(a2, cfg, chunk, local, seed=2000 + rank)   # the rank's share is processed in chunks of 256 resident frames
        n_chunks = max(1, per_rank // chunk)
        for _ in range(3):
            arm.step_resident()
        arm.fm.sync()
        barrier()
        ms = arm.timed(n_chunks)                               # one pass over the rank's share
        ms_total = max_over_ranks([ms * n_chunks])[0]
        del arm
        return {"config": "KITTI-shaped 4096-frame stereo batch (1241x376, 2000 features), extraction + stereo, strong scaling: "
                          f"{per_rank} frames per GPU in chunks of {chunk} (the same {chunk} resident synthetic frames re-run per chunk)",
                "value": total / (ms_total * 1e-3), "unit": UNIT, "ms_per_batch": ms_total, "scaling": "strong", "n_gpus": world}
    if name == "hires":
        cfg, F = synth.CONFIGS["hires"], 32
        a2.workload = "hires"
        arm = Arm(a2, cfg, F, local, seed=3000 + rank)
        for _ in range(3):
            arm.step_resident()
        arm.fm.sync()
        barrier()
        ms = max_over_ranks([arm.timed(5)])[0]
        del arm
        return {"config": "high-res stress: 2048x1536 stereo, 8 levels, 8000 features, 32 frames per GPU, extraction + stereo, weak scaling",
                "value": world * F / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "scaling": "weak", "n_gpus": world}
    return {"error": f"unknown workload {name}"}


if __name__ == "__main__":
    main()
