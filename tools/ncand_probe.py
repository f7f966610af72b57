#!/usr/bin/env python3
"""FAST candidates per pyramid level (before selection) on the bench images: what sizes the selection kernel's shared-memory class.
B200, round 2: EuRoC 4258-4374 at level 0 down to 820-847 at level 7; KITTI 5372-5485 down to 1086-1115."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
for name in ("euroc", "kitti", "mono"):
    cfg = synth.CONFIGS[name]
    w, h = cfg["width"], cfg["height"]
    n = 8
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], n, seed=1000)
    fe = gvs.FeatureExtractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
    fe.extract_batch(np.ascontiguousarray(L)); fe.sync()
    for l in range(cfg["n_levels"]):
        c = [len(fe.level_candidates(i, l)[0]) for i in range(n)]
        print(name, "level", l, "candidates min/max", min(c), max(c))
