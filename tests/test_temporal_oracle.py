"""CPU: the temporal oracle (oracle/cpu_ref.cpp: ora_temporal_mappoints = initializeMap's point creation + MapPoint::update +
removeOutOfFrameMPs) against an independent restatement in Python (oracle/ref_cv2.py: worldToFrame), and against the geometry of
the synthetic sequence; the knife edge of MapPoint::predictScale."""
import math

import numpy as np

from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_cv2

F32 = np.float32


def _py_mappoints(fe, kl, depth, posePrev, toCam, toCamR, fx, fy, cx, cy, w, h):
    """reference src/FeatureTracker.cpp:89-112 + src/Map.cpp:61-104 + src/FeatureTracker.cpp:910-940, scalar Python doubles"""
    logScale = F32(math.log(float(F32(fe.params[2]))))
    out = dict(key=[], predL=[], predR=[], scaleLevelL=[], scaleLevelR=[])
    P = posePrev
    for i in range(len(kl)):
        if not depth[i] > 0:
            continue
        zp = float(depth[i])
        c = [(float(kl["x"][i]) - cx) * zp / fx, (float(kl["y"][i]) - cy) * zp / fy, zp, 1.0]
        wp = [((float(P[r, 0]) * c[0] + float(P[r, 1]) * c[1]) + float(P[r, 2]) * c[2]) + float(P[r, 3]) * c[3] for r in range(4)]
        e = [wp[0] - float(P[0, 3]), wp[1] - float(P[1, 3]), wp[2] - float(P[2, 3])]
        dist = F32(math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]))
        msd = F32(dist * fe.scalePyramid[int(kl["octave"][i])])
        wa = np.array([wp]); ms = np.array([msd], F32); ls = np.array([logScale], F32); nl = np.array([fe.nLevels], np.int32)
        pl = np.zeros((1, 2), F32); pr = np.zeros((1, 2), F32); ll = np.zeros(1, np.int32); lr = np.zeros(1, np.int32)
        f1 = np.zeros(1, np.uint8); f2 = np.zeros(1, np.uint8)
        c1 = ref_cv2.worldToFrame(wa, toCam, fx, fy, cx, cy, w, h, ms, ls, nl, pl, ll, f1)[0]
        c2 = ref_cv2.worldToFrame(wa, toCamR, fx, fy, cx, cy, w, h, ms, ls, nl, pr, lr, f2)[0]
        if not (c1 and c2):
            continue
        out["key"].append(i); out["predL"].append(pl[0]); out["predR"].append(pr[0]); out["scaleLevelL"].append(ll[0]); out["scaleLevelR"].append(lr[0])
    return {k: np.array(v) for k, v in out.items()}


def test_temporal_mappoints_two_restatements_and_geometry():
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], 3, seed=9)
    fe_args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    frames = []
    for f in range(3):
        oL, oR = cpu_ref.Extractor(**fe_args), cpu_ref.Extractor(**fe_args)
        kl, dl = oL.extractKeysNew(L[f]); kr, dr = oR.extractKeysNew(R[f])
        s = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], h, kl, dl, kr, dr)
        frames.append((oL, kl, dl, kr, dr, s))
    prev = [1, 0, 1]
    mo = synth.sequence_motion(cfg, 3, prev)
    for f in range(3):
        fe, kl, dl, kr, dr, s = frames[prev[f]]
        a = cpu_ref.temporal_mappoints(fe, kl, s["estimatedDepth"], mo["pose_prev"][f], mo["to_camera"][f], mo["to_camera_r"][f],
                                       mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h)
        b = _py_mappoints(fe, kl, s["estimatedDepth"], mo["pose_prev"][f], mo["to_camera"][f], mo["to_camera_r"][f], mo["fx"], mo["fy"],
                          mo["cx"], mo["cy"], w, h)
        assert len(a["key"]) > 20
        for k in ("key", "predL", "predR", "scaleLevelL", "scaleLevelR"):
            assert np.array_equal(a[k], b[k]), (f, k)
        # geometry: the prediction is the key moved by the canvas offset difference (sub-pixel depth noise only)
        ox, oy = synth.sequence_offset(f); px, py = synth.sequence_offset(prev[f])
        exp = np.stack([kl["x"][a["key"]] + (px - ox), kl["y"][a["key"]] + (py - oy)], 1)
        assert np.median(np.abs(a["predL"] - exp)) < 0.5
        assert np.median(np.abs((a["predL"][:, 0] - a["predR"][:, 0]) - cfg["disparity"])) < 0.5
        # the matcher finds most of them again in frame f
        fe2, kl2, dl2, kr2, dr2, s2 = frames[f]
        M = len(a["key"])
        mps = dict(desc=dl[a["key"]], predL=a["predL"], predR=a["predR"], scaleLevelL=a["scaleLevelL"], scaleLevelR=a["scaleLevelR"],
                   inFrame=np.ones(M, np.uint8), inFrameR=np.ones(M, np.uint8))
        mL = np.full(len(kl2), -1, np.int32); mR = np.full(len(kr2), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = cpu_ref.match_projection(fe2, 0, mps, kl2, dl2, kr2, dr2, s2["leftIdxs"], s2["rightIdxs"], w, h, 10.0, mL, mR, mt)
        assert n > 0.4 * M


def test_predict_scale_knife_edge_uses_double_log():
    """MapPoint::predictScale (src/Map.cpp:13-23): `log(dif)` on a float is ::log(double) when only <cmath> is in scope (the
    reference's include chain).  dif == scalePyramid[k] exactly is the case where logf and log disagree."""
    fe = cpu_ref.Extractor(1000, 8)
    logScale = F32(math.log(float(F32(1.2))))
    for k in range(8):
        msd = F32(fe.scalePyramid[k])            # dist == 1 -> dif == scalePyramid[k]
        wp = np.array([[0.0, 0.0, 1.0, 1.0]])
        pred = np.zeros((1, 2), F32); lvl = np.zeros(1, np.int32)
        inF = cpu_ref.world_to_frame(np.eye(4), 400.0, 400.0, 320.0, 240.0, 640, 480, wp, np.array([msd], F32), np.array([logScale], F32),
                                     np.array([8], np.int32), pred, lvl)
        expect = int(math.ceil(math.log(float(msd)) / float(logScale)))
        expect = min(max(expect, 0), 7)
        assert inF[0] == 1 and lvl[0] == expect, (k, lvl[0], expect)
        p2 = np.zeros((1, 2), F32); l2 = np.zeros(1, np.int32); f2 = np.zeros(1, np.uint8)
        ref_cv2.worldToFrame(wp, np.eye(4), 400.0, 400.0, 320.0, 240.0, 640, 480, np.array([msd], F32), np.array([logScale], F32),
                             np.array([8], np.int32), p2, l2, f2)
        assert l2[0] == expect
