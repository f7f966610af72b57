#!/usr/bin/env python3
"""Target of the compute-sanitizer runs (profiles/sanitizer_*.txt): the smoke test plus one small EuRoC-shaped batch through every
stage of the step (extract L/R, stereo, grids, temporal), the ticket pipeline, the projection matchers and the multi-device front
end.  Run on a GPU box:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402
import gtsam_vslam_b200 as gvs  # noqa: E402
from gtsam_vslam_b200 import synth  # noqa: E402

__graft_entry__.smoke()
cfg = synth.CONFIGS["euroc"]
w, h = cfg["width"], cfg["height"]
F = 3
L, R = synth.stereo_sequence(w, h, cfg["disparity"], F, seed=5)
args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=F)
feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
fm.track_stereo_batch(L, R)
prev = [1, 0, 1]
mo = synth.sequence_motion(cfg, F, prev)
fm.track_temporal_batch(F, prev, mo["pose_prev"], mo["to_camera"], mo["to_camera_r"], mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h, 10.0)
t = fm.temporal_download(0, F)
k = gvs.tracked_keys(fm, 1, w, h)
M = len(k.keyPoints)
mps = dict(desc=k.Desc, predL=np.stack([k.keyPoints["x"], k.keyPoints["y"]], 1).astype(np.float32),
           predR=np.stack([k.keyPoints["x"] - cfg["disparity"], k.keyPoints["y"]], 1).astype(np.float32),
           scaleLevelL=k.keyPoints["octave"].astype(np.int32), scaleLevelR=k.keyPoints["octave"].astype(np.int32))
for mode, rad in ((0, 10.0), (1, 1200.0), (2, 120.0)):
    mL = np.full(len(k.keyPoints), -1, np.int32); mR = np.full(len(k.rightKeyPoints), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
    fm._project(1, mode, mps, rad, mL, mR if mode == 0 else None, mt)
pipe = gvs.StereoPipeline(cfg["fx"], cfg["baseline"], width=w, height=h, frames_per_slot=2, n_slots=2, nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
tk = [pipe.submit(L[:2], R[:2]), pipe.submit(L[1:3], R[1:3])]
for x in tk:
    pipe.download(x, 2); pipe.release(x)
mt = gvs.MultiGpuTracker(cfg["fx"], cfg["baseline"], [0, 0], width=w, height=h, frames_per_device=2, nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
mt.track_stereo(L, R)
from gtsam_vslam_b200 import api  # noqa: E402
bad, n = api.check_guards()
print(f"guard tails: {bad} overwritten of {n} device buffers checked (GVS_GUARD={os.environ.get('GVS_GUARD', '')}, GVS_POISON={os.environ.get('GVS_POISON', '')})")
assert bad == 0
print(f"sanitize target ok: {int(t['n_matches'].sum())} temporal matches, {M} keys in frame 1")
