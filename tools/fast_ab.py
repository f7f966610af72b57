#!/usr/bin/env python3
"""Development A/B of the FAST kernel's tile staging: the same 64 EuRoC images through a handle built with TMA staging and one
built without (GVS_NO_FAST_TMA is read when a handle is created).  Run under ncu with -k regex:fast_cells to compare the two."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs  # noqa: E402
from gtsam_vslam_b200 import synth  # noqa: E402

cfg = synth.CONFIGS["euroc"]
w, h, n = cfg["width"], cfg["height"], 64
imgs = np.ascontiguousarray(synth.stereo_batch(w, h, cfg["disparity"], range(8))[0])
imgs = np.ascontiguousarray(np.tile(imgs, (n // 8, 1, 1)))
args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
out = []
for no_tma in (False, True):
    if no_tma:
        os.environ["GVS_NO_FAST_TMA"] = "1"
    fe = gvs.FeatureExtractor(**args)
    for _ in range(3):
        fe.extract_batch(imgs); fe.sync()
    off = fe.offsets(n)
    kp, desc = fe.download(0, n, int(off[n]))
    out.append((fe.tma_stages(), kp.tobytes() + desc.tobytes()))
print("tma stages", out[0][0], out[1][0], "identical", out[0][1] == out[1][1])
