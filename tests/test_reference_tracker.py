"""CPU: the restated FeatureTracker / Map functions of the oracle (oracle/cpu_ref.cpp) against the REFERENCE'S OWN CODE: the bodies
of FeatureTracker::assignKeysToGrids, initializeMap, check2dError, findOutliersR / Mono, worldToFrame, removeOutOfFrameMPs / Mono,
PredictMPsPosition compiled from src/FeatureTracker.cpp by oracle/slice_reference.py, and src/Map.cpp compiled whole
(MapPoint::MapPoint / update / predictScale / calcDescriptor).  This pins SURVEY rows a16, (f) rank 1 (calcDescriptor) and rank 3."""
import math

import numpy as np
import pytest

from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_build
from test_oracle_vs_cv2 import _outlier_case, _predict_case, _inv_sigma_factor

pytestmark = pytest.mark.skipif(not ref_build.available(), reason="oracle/_ref needs the reference sources (build container)")

EXT = np.eye(4); EXT[0, 3] = 0.110074


def _frame(cfgname="small", seed=5):
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], seed=seed)
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    st = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], h, kl, dl, kr, dr)
    return cfg, oL, kl, dl, kr, dr, st


def _tracker(cfg, ext=None, fx=None):
    ref_build.use_default_hooks()
    feL = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"]); feR = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"])
    f = cfg["fx"] if fx is None else fx
    return ref_build.RefTracker(feL, feR, f, f, cfg["width"] / 2.0, cfg["height"] / 2.0, cfg["baseline"], cfg["width"], cfg["height"],
                                EXT if ext is None else ext)


@pytest.mark.parametrize("w,h", [(320, 240), (752, 480), (1241, 376), (360, 700), (640, 640)])
def test_assign_keys_to_grids(w, h):
    rng = np.random.default_rng(w + h)
    n = 1500
    kps = np.zeros(n, cpu_ref.KEYPOINT_DTYPE)
    kps["x"] = rng.uniform(-3, w + 3, n).astype(np.float32); kps["y"] = rng.uniform(-3, h + 3, n).astype(np.float32)
    kps["x"][:40] = np.round(kps["x"][:40] * 2) / 2; kps["y"][:40] = np.round(kps["y"][:40] * 2) / 2   # .5 cases of cvRound
    t = _tracker(synth.CONFIGS["small"])
    a = cpu_ref.assign_grid(kps, w, h); b = t.assign_grid(kps, w, h)
    assert a["yGrids"] == b["yGrids"] and a["xMult"] == b["xMult"] and a["yMult"] == b["yMult"]
    assert np.array_equal(a["start"], b["start"]) and np.array_equal(a["items"], b["items"])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_predict_mappoints(mode):
    cfg = synth.CONFIGS["euroc"]
    t = _tracker(cfg)
    cam = (cfg["fx"], cfg["fx"], cfg["width"] / 2.0, cfg["height"] / 2.0, cfg["width"], cfg["height"])
    c = _predict_case(seed=20 + mode, M=1500)
    P = np.linalg.inv(c["toCamera"])                      # predNPose of the reference call
    toC, toCR = t.cameras(P)
    M = len(c["wPos"])

    def state():
        return dict(predL=np.full((M, 2), -7, np.float32), slL=np.full(M, -7, np.int32), inF=np.full(M, 1, np.uint8),
                    predR=np.full((M, 2), -7, np.float32), slR=np.full(M, -7, np.int32), inFR=np.full(M, 1, np.uint8),
                    matches=c["matches"].copy(), mL=c["matchedL"].copy(), mR=c["matchedR"].copy(), out=c["outliers"].copy())
    two = mode != 1
    a = state()
    ka = cpu_ref.predict_mappoints(mode, toC, toCR if two else None, *cam, c["valid"], c["wPos"], c["msd"], c["ls"], c["nlev"], a["predL"], a["slL"],
                                   a["inF"], a["predR"] if two else None, a["slR"] if two else None, a["inFR"] if two else None,
                                   *((a["matches"], a["mL"], a["mR"], a["out"]) if mode == 2 else (None, None, None, None)))
    b = state()
    kb = t.predict_mappoints(mode, np.eye(4), P, c["valid"], c["wPos"], c["msd"], c["ls"], c["nlev"], b["predL"], b["slL"], b["inF"],
                             b["predR"] if two else None, b["slR"] if two else None, b["inFR"] if two else None,
                             *((b["matches"], b["mL"], b["mR"], b["out"]) if mode == 2 else (None, None, None, None)))
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    if mode != 2:
        assert np.array_equal(ka, kb) and 0.02 < kb.mean() < 0.9
    else:
        assert (b["matches"] != c["matches"]).sum() > 50


@pytest.mark.parametrize("mode", [0, 1])
def test_find_outliers(mode):
    cfg, oL, kl, dl, kr, dr, st = _frame()
    w, h = cfg["width"], cfg["height"]
    ext = np.eye(4); ext[0, 3] = cfg["baseline"]
    t = _tracker(cfg, ext)
    c = _outlier_case(kl, kr, st, cfg["fx"], w / 2.0, h / 2.0, cfg["baseline"], seed=31 + mode)
    poseR = t.outlier_cameras(c["pose"])
    cam = (cfg["fx"], cfg["fx"], w / 2.0, h / 2.0)
    invSigma = _inv_sigma_factor(cfg["n_levels"])
    m1 = c["matches"].copy(); o1 = np.full(len(m1), 7, np.uint8)
    a1 = [st["estimatedDepth"].copy(), st["close"].astype(np.uint8).copy(), st["rightIdxs"].copy(), st["leftIdxs"].copy()]
    n1 = cpu_ref.find_outliers(mode, c["pose"], poseR, *cam, c["thres"], c["closeDist"], invSigma, c["valid"], c["wPos"], c["inFrame"], c["inFrameR"],
                               m1, o1, kl, kr, *a1)
    m2 = c["matches"].copy(); o2 = np.full(len(m2), 7, np.uint8)
    a2 = [st["estimatedDepth"].copy(), st["close"].astype(np.uint8).copy(), st["rightIdxs"].copy(), st["leftIdxs"].copy()]
    n2 = t.find_outliers(mode, c["pose"], c["thres"], c["valid"], c["wPos"], c["inFrame"], c["inFrameR"], m2, o2, kl, kr, *a2)
    # (the reference writes vector<bool> MPsOutliers: an untouched 7 comes back as 1)
    assert n1 == n2 and np.array_equal(m1, m2) and np.array_equal(np.minimum(o1, 1), o2)
    for x, y in zip(a1, a2):
        assert np.array_equal(x, y)
    assert n1[0] > 100 and (mode == 1 or n1[1] > 10)


@pytest.mark.parametrize("cfgname,seed", [("small", 3), ("euroc", 7)])
def test_temporal_mappoints(cfgname, seed):
    cfg, oL, kl, dl, kr, dr, st = _frame(cfgname, seed)
    w, h = cfg["width"], cfg["height"]
    ext = np.eye(4); ext[0, 3] = cfg["baseline"]
    t = _tracker(cfg, ext)
    ang = 0.01
    posePrev = np.eye(4); posePrev[:3, 3] = [0.3, -0.2, 0.1]
    posePrev[:3, :3] = [[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]]
    predN = posePrev.copy(); predN[:3, 3] += [0.05, 0.02, 0.3]
    toC, toCR = t.cameras(predN)
    a = cpu_ref.temporal_mappoints(oL, kl, st["estimatedDepth"], posePrev, toC, toCR, cfg["fx"], cfg["fx"], w / 2.0, h / 2.0, w, h)
    b = t.temporal_mappoints(kl, dl, kr, dr, st["rightIdxs"], st["estimatedDepth"], posePrev, predN)
    assert len(b["key"]) > 30 and len(b["key"]) < (st["estimatedDepth"] > 0).sum()   # some points leave the frame
    for k in ("key", "predL", "predR", "scaleLevelL", "scaleLevelR", "wPos", "maxScaleDist"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(b["desc"], dl[b["key"]])       # calcDescriptor keeps the left descriptor of a fresh stereo point


def test_median_descriptor():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 8, 13):
        base = rng.integers(0, 256, 32).astype(np.uint8)
        descs = np.stack([base ^ (rng.random(32) < 0.1 * (1 + i % 3)).astype(np.uint8) * rng.integers(0, 256, 32).astype(np.uint8) for i in range(n)])
        order, chosen = ref_build.median_descriptor(descs)
        assert sorted(order) == list(range(n))
        inorder = descs[order]
        best = cpu_ref.median_descriptor(inorder, np.array([0, n], np.int32))[0]
        assert np.array_equal(inorder[best], chosen), n
