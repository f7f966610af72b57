"""GPU (-m gpu): the CUDA entry points of SURVEY rows a16 and (f) ranks 1 / 3 and the temporal step DIRECTLY against the reference's
own code in oracle/_ref: FeatureTracker::assignKeysToGrids / initializeMap / findOutliersR / findOutliersMono / worldToFrame /
removeOutOfFrameMPs[Mono] / PredictMPsPosition (bodies compiled from src/FeatureTracker.cpp by oracle/slice_reference.py) and
MapPoint (src/Map.cpp compiled whole).  The library is built in the CPU container and travels to the GPU box."""
import math

import numpy as np
import pytest

import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_build
from test_oracle_vs_cv2 import _outlier_case, _predict_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_build.available(), reason="oracle/_ref/libref_frontend.so was not built")]


def _handles(cfg, n):
    w, h = cfg["width"], cfg["height"]
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
    feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
    return feL, feR, gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)


def _tracker(cfg, w=None, h=None):
    ref_build.use_default_hooks()
    feL = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"]); feR = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"])
    w = w or cfg["width"]; h = h or cfg["height"]
    ext = np.eye(4); ext[0, 3] = cfg["baseline"]
    return ref_build.RefTracker(feL, feR, cfg["fx"], cfg["fx"], w / 2.0, h / 2.0, cfg["baseline"], w, h, ext)


@pytest.mark.parametrize("cfgname", ["small", "euroc", "kitti"])
def test_key_grids_against_assignKeysToGrids(cfgname):
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_batch(w, h, cfg["disparity"], [61, 62])
    feL, feR, fm = _handles(cfg, 2)
    fm.track_stereo_batch(L, R)
    t = _tracker(cfg)
    for f in range(2):
        k = gvs.tracked_keys(fm, f, w, h)
        for keys, g in ((k.keyPoints, k.lkeyGrid), (k.rightKeyPoints, k.rkeyGrid)):
            r = t.assign_grid(keys, w, h)
            assert g["yGrids"] == r["yGrids"] and g["xMult"] == r["xMult"] and g["yMult"] == r["yMult"]
            assert np.array_equal(g["start"], r["start"]) and np.array_equal(g["items"], r["items"])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_predict_mappoints_against_the_reference(mode):
    cfg = synth.CONFIGS["euroc"]
    t = _tracker(cfg)
    cam = (cfg["fx"], cfg["fx"], cfg["width"] / 2.0, cfg["height"] / 2.0, cfg["width"], cfg["height"])
    c = _predict_case(seed=40 + mode, M=6000, nl=1100, nr=1050)
    P = np.linalg.inv(c["toCamera"])
    toC, toCR = t.cameras(P)
    M = len(c["wPos"])
    feL, feR, fm = _handles(synth.CONFIGS["small"], 1)

    def state():
        return dict(predL=np.full((M, 2), -7, np.float32), slL=np.full(M, -7, np.int32), inF=np.full(M, 1, np.uint8),
                    predR=np.full((M, 2), -7, np.float32), slR=np.full(M, -7, np.int32), inFR=np.full(M, 1, np.uint8),
                    matches=c["matches"].copy(), mL=c["matchedL"].copy(), mR=c["matchedR"].copy(), out=c["outliers"].copy())
    two = mode != 1
    a = state()
    ka = fm.predictMapPoints(mode, toC, toCR if two else None, *cam, c["valid"], c["wPos"], c["msd"], c["ls"], c["nlev"], a["predL"], a["slL"],
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


@pytest.mark.parametrize("mode", [0, 1])
def test_find_outliers_against_the_reference(mode):
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], seed=15)
    feL, feR, fm = _handles(cfg, 1)
    k = gvs.extractORBAndStereoMatch(fm, L, R)
    t = _tracker(cfg)
    cam = (cfg["fx"], cfg["fx"], w / 2.0, h / 2.0)
    cur = dict(estimatedDepth=k.estimatedDepth.copy(), close=k.close.copy(), rightIdxs=k.rightIdxs.copy(), leftIdxs=k.leftIdxs.copy())
    c = _outlier_case(k.keyPoints, k.rightKeyPoints, cur, cfg["fx"], w / 2.0, h / 2.0, cfg["baseline"], seed=3 + mode, M=2500)
    poseR = t.outlier_cameras(c["pose"])
    m2 = c["matches"].copy(); o2 = np.full(len(m2), 1, np.uint8)
    a2 = [cur["estimatedDepth"].copy(), cur["close"].astype(np.uint8).copy(), cur["rightIdxs"].copy(), cur["leftIdxs"].copy()]
    n2 = t.find_outliers(mode, c["pose"], c["thres"], c["valid"], c["wPos"], c["inFrame"], c["inFrameR"], m2, o2, k.keyPoints, k.rightKeyPoints, *a2)
    m = c["matches"].copy(); out = np.full(len(m), 1, np.uint8)
    if mode == 0:
        cnt = fm.findOutliersR(0, c["pose"], poseR, *cam, c["thres"], c["closeDist"], c["valid"], c["wPos"], c["inFrame"], c["inFrameR"], m, out, keys=k)
    else:
        cnt = (fm.findOutliersMono(0, c["pose"], *cam, c["thres"], c["valid"], c["wPos"], c["inFrame"], m, out), 0)
    assert tuple(cnt) == tuple(n2) and np.array_equal(m, m2) and np.array_equal(out, o2)
    if mode == 0:
        for got, want in zip((k.estimatedDepth, k.close.astype(np.uint8), k.rightIdxs, k.leftIdxs), a2):
            assert np.array_equal(got, want)
        assert n2[1] > 10


@pytest.mark.parametrize("cfgname,F", [("euroc", 4), ("kitti", 3)])
def test_temporal_step_against_the_reference(cfgname, F):
    """gvs_track_temporal_batch == initializeMap + removeOutOfFrameMPs + matchByProjectionRPred, all three run by the reference's
    own source on the CPU."""
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], F, seed=44)
    feL, feR, fm = _handles(cfg, F)
    fm.track_stereo_batch(L, R)
    prev = [1] + list(range(F - 1))
    mo = synth.sequence_motion(cfg, F, prev)
    t = _tracker(cfg)
    # the device gets exactly the inverses the reference computes from predNPose
    toC = np.zeros((F, 4, 4)); toCR = np.zeros((F, 4, 4))
    for f in range(F):
        toC[f], toCR[f] = t.cameras(mo["pose"][f])
    fm.track_temporal_batch(F, prev, mo["pose_prev"], toC, toCR, mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h, 10.0)
    out = fm.temporal_download(0, F, preds=True)
    keys = [gvs.tracked_keys(fm, f, w, h) for f in range(F)]
    offL = feL.offsets(F); offR = feR.offsets(F)
    rfm = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], t.feL, t.feR, h)
    for f in range(F):
        kp = keys[prev[f]]
        r = t.temporal_mappoints(kp.keyPoints, kp.Desc, kp.rightKeyPoints, kp.rightDesc, kp.rightIdxs, kp.estimatedDepth, mo["pose_prev"][f], mo["pose"][f])
        sl = slice(out["mp_offsets"][f], out["mp_offsets"][f + 1])
        assert np.array_equal(out["mp_key"][sl], r["key"]) and np.array_equal(out["predL"][sl], r["predL"]) and np.array_equal(out["predR"][sl], r["predR"])
        M = len(r["key"])
        mps = dict(desc=r["desc"], predL=r["predL"], predR=r["predR"], scaleLevelL=r["scaleLevelL"], scaleLevelR=r["scaleLevelR"],
                   inFrame=np.ones(M, np.uint8), inFrameR=np.ones(M, np.uint8))
        k = keys[f]
        oL = np.full(len(k.keyPoints), -1, np.int32); oR = np.full(len(k.rightKeyPoints), -1, np.int32); ot = np.full((M, 2), -1, np.int32)
        on = rfm.match_projection(0, mps, k.keyPoints, k.Desc, k.rightKeyPoints, k.rightDesc, k.leftIdxs, k.rightIdxs,
                                  t.assign_grid(k.keyPoints, w, h), t.assign_grid(k.rightKeyPoints, w, h), 10.0, oL, oR, ot)
        assert out["n_matches"][f] == on and np.array_equal(out["matches"][sl], ot), f
        assert np.array_equal(out["matchedL"][offL[f]:offL[f + 1]], oL) and np.array_equal(out["matchedR"][offR[f]:offR[f + 1]], oR), f
        assert on > 0.3 * M


def test_median_descriptor_against_calcDescriptor():
    cfg = synth.CONFIGS["small"]
    feL, feR, fm = _handles(cfg, 1)
    rng = np.random.default_rng(8)
    lists, want = [], []
    for n in (1, 2, 3, 4, 7, 12, 20):
        base = rng.integers(0, 256, 32).astype(np.uint8)
        d = np.stack([base ^ ((rng.random(32) < 0.08 * (1 + i % 4)) * rng.integers(0, 256, 32)).astype(np.uint8) for i in range(n)])
        order, chosen = ref_build.median_descriptor(d)
        lists.append(d[order]); want.append(chosen)
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int32)
    best = fm.median_descriptor(np.concatenate(lists), off)
    for l, b, w_ in zip(lists, best, want):
        assert np.array_equal(l[b], w_)
