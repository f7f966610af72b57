#!/usr/bin/env python3
"""Debug helper: per-level stage comparison for one custom parameter set (run on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref
w, h, nF, nL, scale, edge, maxF, minF, seed = 512, 384, 4000, 3, 2.0, 19, 12, 3, 55
L, R = synth.stereo_pair(w, h, 10, seed)
fe = gvs.FeatureExtractor(nF, nL, scale, edge, 31, maxF, minF, width=w, height=h, max_images=1)
o = cpu_ref.Extractor(nF, nL, scale, edge, 31, maxF, minF)
kp, d = fe.extractKeysNew(L); print("gpu keys", len(kp))
ko, do = o.extractKeysNew(L)
print("quota", o.featurePerLevel)
for l in range(nL):
    a = fe.level_image(0, l); b = o.level_image(l)
    print("L", l, "pyr", np.array_equal(a, b), a.shape)
    x, y, s = fe.level_candidates(0, l); c = o.level_keys(l, False)
    print("   cand", len(x), len(c["x"]), np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]))
    xk, yk, sk = fe.level_kept(0, l); ck = o.level_keys(l, True)
    print("   kept", len(xk), len(ck["x"]), "passes", o.ssc_passes(l))
    srt, kept = fe.debug_select(l, c["x"], c["y"], c["response"].astype(np.int32))
    s2, k2 = o.ssc(int(o.featurePerLevel[l]), a.shape[1], a.shape[0], c["x"], c["y"], c["response"])
    print("   debug_select kept", len(kept), "oracle", len(k2), "sorted equal", None if srt is None else np.array_equal(srt, s2))
