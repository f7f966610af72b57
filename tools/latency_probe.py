#!/usr/bin/env python3
"""Single-stereo-frame latency through the public API (host image in, TrackedKeys arrays out), the way the reference's
tracking thread calls extractORBAndStereoMatch once per frame.  Run on a GPU box."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth

for name in ("euroc", "kitti"):
    cfg = synth.CONFIGS[name]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_batch(w, h, cfg["disparity"], range(8))
    pl = torch.from_numpy(L).pin_memory().numpy(); pr = torch.from_numpy(R).pin_memory().numpy()
    feL = gvs.FeatureExtractor(cfg["n_features"], cfg["n_levels"], width=w, height=h, max_images=1)
    feR = gvs.FeatureExtractor(cfg["n_features"], cfg["n_levels"], width=w, height=h, max_images=1)
    fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    cap = feL.cap_per_image
    pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory().numpy()
    kl = pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE); kr = pin(cap * 28, torch.uint8).view(gvs.KEYPOINT_DTYPE)
    dl = pin(cap * 32, torch.uint8); dr = pin(cap * 32, torch.uint8)
    ri = pin(cap, torch.int32); li = pin(cap, torch.int32); de = pin(cap, torch.float32); cl = pin(cap, torch.uint8)
    ts = []
    for it in range(60):
        f = it % 8
        t = time.perf_counter()
        fm.track_stereo_batch(pl[f:f + 1], pr[f:f + 1])
        nl = feL.download_into(0, 1, kl, dl); nr = feR.download_into(0, 1, kr, dr)
        fm.stereo_download_into(0, 1, ri, li, de, cl)
        ts.append((time.perf_counter() - t) * 1e3)
    ts = np.array(ts[10:])
    print(f"{name}: single stereo frame host->host latency median {np.median(ts):.3f} ms, p95 {np.percentile(ts, 95):.3f} ms, keys {nl}/{nr}")
