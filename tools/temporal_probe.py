#!/usr/bin/env python3
"""BASELINE config 4 (mono temporal): time of the projection matchers on the GPU against the CPU oracle, and the timing of
the real OpenCV primitives (cv2) for one EuRoC-shaped frame next to the scalar restatement (SURVEY 8(d) cross-check).
Run on a GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref

cfg = synth.CONFIGS["mono"]
w, h = cfg["width"], cfg["height"]
seq = synth.shifted_sequence(w, h, 2, seed=4)
fe = gvs.FeatureExtractor(nFeatures=2000, width=w, height=h, max_images=2)
fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], fe, None, h)
fe.extract_batch(seq); fm.assign_grids_batch(2, w, h)
off = fe.offsets(2)
k0, d0 = fe.download(0, 1, int(off[1])); k1, d1 = fe.download(1, 1, int(off[2] - off[1]))
o = cpu_ref.Extractor(2000); o.extractKeysNew(seq[1])
rng = np.random.default_rng(0)
M = len(k0)
mps = dict(desc=d0, predL=np.stack([k0["x"] - 2, k0["y"] - 1], 1).astype(np.float32) + rng.normal(0, 1, (M, 2)).astype(np.float32),
           scaleLevelL=k0["octave"].astype(np.int32), inFrame=np.ones(M, np.uint8))
for rad in (10.0, 1200.0):
    ts = []
    for _ in range(30):
        mL = np.full(len(k1), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        t = time.perf_counter(); n = fm.matchByProjectionMono(mps, 1, mL, mt, rad); ts.append((time.perf_counter() - t) * 1e3)
    tc = []
    for _ in range(3):
        mL2 = np.full(len(k1), -1, np.int32); mt2 = np.full((M, 2), -1, np.int32)
        t = time.perf_counter(); n2 = cpu_ref.match_projection(o, 1, mps, k1, d1, None, None, None, None, w, h, rad, mL2, None, mt2); tc.append((time.perf_counter() - t) * 1e3)
    assert n == n2 and np.array_equal(mL, mL2) and np.array_equal(mt, mt2)
    print(f"matchByProjectionMono M={M} N={len(k1)} rad={rad:g}: {n} matches, GPU host-to-host median {np.median(ts[5:]):.3f} ms, CPU oracle {min(tc):.2f} ms")

# ---- real OpenCV primitives on one EuRoC-shaped frame (native calls only), 1 thread and all threads
try:
    import cv2, math
except Exception as e:
    print("cv2 not importable:", e); sys.exit(0)
c2 = synth.CONFIGS["euroc"]
img = synth.stereo_pair(c2["width"], c2["height"], c2["disparity"], 0)[0]
for nt in (1, os.cpu_count()):
    cv2.setNumThreads(nt)
    d20 = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    best = None
    for rep in range(3):
        t_res = t_fast = t_blur = 0.0; calls = 0
        lvl = img; scale = 1.0
        for l in range(8):
            if l:
                scale *= 1.2
                sz = (int(round(img.shape[1] / np.float32(1.2) ** l)), int(round(img.shape[0] / np.float32(1.2) ** l)))
                t = time.perf_counter(); lvl = cv2.resize(lvl, sz, interpolation=cv2.INTER_LINEAR); t_res += time.perf_counter() - t
            hh, ww = lvl.shape
            minX = 16; maxX = ww - 16; minY = 16; maxY = hh - 16
            nC = int((maxX - minX) / 35); nR = int((maxY - minY) / 35); gW = math.ceil((maxX - minX) / nC); gH = math.ceil((maxY - minY) / nR)
            for iR in range(nR):
                r0 = minY + iR * gH; r1 = min(r0 + gH + 6, maxY)
                if r0 >= maxY - 6: continue
                for iC in range(nC):
                    c0 = minX + iC * gW; c1 = min(c0 + gW + 6, maxX)
                    if c0 >= maxX - 6: continue
                    cell = np.ascontiguousarray(lvl[r0:r1, c0:c1])
                    t = time.perf_counter(); d20.detect(cell); t_fast += time.perf_counter() - t; calls += 1
            t = time.perf_counter(); cv2.GaussianBlur(lvl, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101); t_blur += time.perf_counter() - t
        tot = t_res + t_fast + t_blur
        if best is None or tot < best[0]: best = (tot, t_res, t_fast, t_blur, calls)
    print(f"cv2 {cv2.__version__} threads={nt}: one 752x480 image: resize {best[1]*1e3:.2f} ms, {best[4]} FAST cell calls {best[2]*1e3:.2f} ms (python call overhead included), blur {best[3]*1e3:.2f} ms")
t = time.perf_counter(); ox = cpu_ref.Extractor(1000); ox.extractKeysNew(img); print(f"scalar restatement, whole extractKeysNew of the same image: {(time.perf_counter()-t)*1e3:.1f} ms")
