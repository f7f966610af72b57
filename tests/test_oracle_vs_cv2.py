"""CPU: the scalar C++ oracle (oracle/cpu_ref.cpp) against the real OpenCV primitives (cv2) and against the
committed golden fixtures made by oracle/ref_cv2.py.  This is what pins the oracle (the reference has no tests)."""
import ctypes as C
import hashlib
import math
import os

import cv2
import numpy as np
import pytest

from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_cv2

cv2.setNumThreads(1)
lib = cpu_ref.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("shape", [(376, 1241), (480, 752), (105, 346), (61, 83), (240, 320)])
def test_resize_linear_matches_cv2(shape):
    rng = np.random.default_rng(shape[0])
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    dh, dw = int(round(shape[0] / 1.2)), int(round(shape[1] / 1.2))
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((dh, dw), np.uint8)
    lib.ora_resize_linear(_p(src), shape[1], shape[0], _p(out), dw, dh)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("shape", [(313, 1034), (134, 210), (47, 53), (480, 752)])
def test_gaussian_blur_matches_cv2(shape):
    rng = np.random.default_rng(shape[1])
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.GaussianBlur(src.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    out = np.zeros(shape, np.uint8)
    lib.ora_gaussian_blur7(_p(src), shape[1], shape[0], _p(out))
    assert np.array_equal(out, ref)


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = rng.integers(-1300000, 1300000, 20000); xs = rng.integers(-1300000, 1300000, 20000)
    ys[:4] = [0, 0, 5, -5]; xs[:4] = [0, 7, 0, 0]
    for y, x in zip(ys, xs):
        assert lib.ora_fast_atan2(float(y), float(x)) == np.float32(cv2.fastAtan2(float(y), float(x)))


@pytest.mark.parametrize("n,hi", [(1, 5), (16, 3), (17, 3), (100, 4), (5000, 60), (20000, 12)])
def test_sort_idx_matches_cv2(n, hi):
    rng = np.random.default_rng(n)
    key = rng.integers(7, 7 + hi, n).astype(np.float32)
    ref = cv2.sortIdx(key.reshape(1, -1), cv2.SORT_DESCENDING + cv2.SORT_EVERY_ROW)[0]
    out = np.zeros(n, np.int32)
    lib.ora_sort_idx_desc(_p(key), n, _p(out))
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("t", [20, 7])
def test_fast_detect_matches_cv2(t):
    L, _ = synth.stereo_pair(320, 240, 8, 5)
    det = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for (y0, y1, x0, x1) in [(0, 240, 0, 320), (16, 61, 16, 58), (100, 107, 30, 90), (50, 95, 200, 245)]:
        sub = np.ascontiguousarray(L[y0:y1, x0:x1])
        kps = det.detect(sub)
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
        cap = 100000
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = lib.ora_fast_detect(_p(sub), sub.shape[1], sub.shape[0], t, _p(xs), _p(ys), _p(sc), cap)
        assert n == len(ref)
        assert list(zip(xs[:n].tolist(), ys[:n].tolist(), sc[:n].tolist())) == ref  # same points, scores and order


def test_constructor_tables():
    # the only intrinsic known answers of the reference
    e = cpu_ref.Extractor(2000)
    assert e.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert e.featurePerLevel.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert cpu_ref.Extractor(1000).featurePerLevel.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert cpu_ref.Extractor(8000).featurePerLevel.tolist() == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    r = ref_cv2.FeatureExtractor(2000)
    assert np.array_equal(e.scalePyramid, r.scalePyramid) and np.array_equal(e.scaledPatchSize, r.scaledPatchSize)
    assert np.array_equal(e.sigmaFactor, r.sigmaFactor) and np.array_equal(e.InvSigmaFactor, r.InvSigmaFactor)


def _run_cpp(cfg, L, R):
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    s = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], cfg["height"], kl, dl, kr, dr)
    return oL, oR, kl, dl, kr, dr, s


def _check_against_golden(g, kl, dl, kr, dr, s, w, h):
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], g["kpL"][f]), f
        assert np.array_equal(kr[f], g["kpR"][f]), f
    assert np.array_equal(dl, g["descL"]) and np.array_equal(dr, g["descR"])
    assert np.array_equal(s["rightIdxs"], g["rightIdxs"]) and np.array_equal(s["leftIdxs"], g["leftIdxs"])
    assert np.array_equal(s["estimatedDepth"], g["depth"]) and np.array_equal(s["close"], g["close"])
    gl = cpu_ref.assign_grid(kl, w, h)
    assert gl["yGrids"] == int(g["yGrids"]) and np.array_equal(gl["start"], g["gridL_start"]) and np.array_equal(gl["items"], g["gridL_items"])
    gr = cpu_ref.assign_grid(kr, w, h)
    assert np.array_equal(gr["start"], g["gridR_start"]) and np.array_equal(gr["items"], g["gridR_items"])
    assert np.float32(gl["xMult"]) == g["xMult"] and np.float32(gl["yMult"]) == g["yMult"]


@pytest.mark.parametrize("name", ["small_s0", "small_s1_ramp", "euroc_s0", "kitti_s0"])
def test_cpp_oracle_matches_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = synth.CONFIGS[str(g["cfg"])]
    L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], int(g["seed"]), ramp=bool(g["ramp"]))
    assert hashlib.sha256(L.tobytes()).hexdigest() == str(g["sha_left"])  # the generator is deterministic
    assert hashlib.sha256(R.tobytes()).hexdigest() == str(g["sha_right"])
    if "left" in g:
        assert np.array_equal(L, g["left"]) and np.array_equal(R, g["right"])
    oL, oR, kl, dl, kr, dr, s = _run_cpp(cfg, L, R)
    for l in range(cfg["n_levels"]):
        im = oL.level_image(l)
        assert im.shape[::-1] == tuple(g["level_sizes"][l])
        assert hashlib.sha256(im.tobytes()).hexdigest() == str(g["pyr_sha"][l])
        assert len(oL.level_keys(l, False)["x"]) == int(g["candidates"][l])
    _check_against_golden(g, kl, dl, kr, dr, s, cfg["width"], cfg["height"])


def test_cpp_oracle_matches_cv2_oracle_live():
    """End to end on a fresh seed (not a fixture): both oracles run here and must agree bit for bit."""
    cfg = synth.CONFIGS["small"]
    L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], 11)
    feL = ref_cv2.FeatureExtractor(cfg["n_features"], cfg["n_levels"]); feR = ref_cv2.FeatureExtractor(cfg["n_features"], cfg["n_levels"])
    fm = ref_cv2.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, cfg["height"])
    k = ref_cv2.extractORBAndStereoMatch(feL, feR, fm, L, R, cfg["width"], cfg["height"])
    oL, oR, kl, dl, kr, dr, s = _run_cpp(cfg, L, R)
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], k.keyPoints[f]) and np.array_equal(kr[f], k.rightKeyPoints[f])
    assert np.array_equal(dl, k.Desc) and np.array_equal(dr, k.rightDesc)
    assert np.array_equal(s["rightIdxs"], k.rightIdxs) and np.array_equal(s["leftIdxs"], k.leftIdxs)
    assert np.array_equal(s["estimatedDepth"], k.estimatedDepth) and np.array_equal(s["close"], k.close)
    for l in range(cfg["n_levels"]):
        assert np.array_equal(oL.level_image(l, True), cv2.GaussianBlur(feL.imagePyramid[l].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101))


def test_cpp_projection_matchers_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_s0.npz")); p = np.load(os.path.join(golden_dir, "small_s0_proj.npz"))
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    fe = cpu_ref.Extractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    mps = {k[3:]: p[k] for k in p.files if k.startswith("mp_")}
    nl, nr, M = len(g["kpL"]), len(g["kpR"]), len(mps["desc"])
    for rad in (4, 10, 120):
        mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32); mt = np.full((M, 2), -1, np.int32); mt[::17] = [3, -1]
        n = cpu_ref.match_projection(fe, 0, mps, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"], w, h,
                                     float(rad), mL, mR, mt)
        assert n == int(p[f"rpred_{rad}_n"]) and np.array_equal(mL, p[f"rpred_{rad}_mL"]) and np.array_equal(mR, p[f"rpred_{rad}_mR"])
        assert np.array_equal(mt, p[f"rpred_{rad}_mt"])
        mL = np.full(nl, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = cpu_ref.match_projection(fe, 1, mps, g["kpL"], g["descL"], None, None, None, None, w, h, float(rad), mL, None, mt)
        assert n == int(p[f"mono_{rad}_n"]) and np.array_equal(mL, p[f"mono_{rad}_mL"]) and np.array_equal(mt, p[f"mono_{rad}_mt"])
    last = dict(desc=g["descR"], predL=np.stack([g["kpR"]["x"], g["kpR"]["y"]], 1), scaleLevelL=g["kpR"]["octave"])
    mL = np.full(nl, -1, np.int32); mt = np.full((nr, 2), -1, np.int32)
    n = cpu_ref.match_projection(fe, 2, last, g["kpL"], g["descL"], None, None, None, None, w, h, 120.0, mL, None, mt)
    assert n == int(p["radius_n"]) and np.array_equal(mL, p["radius_mL"]) and np.array_equal(mt[:, 0], p["radius_match"])


def test_analytic_depth():
    """Constant-disparity pair: median depth = fx*b/d."""
    g = synth.CONFIGS["kitti"]
    L, R = synth.stereo_pair(g["width"], g["height"], 24, 2)
    *_, s = _run_cpp(g, L, R)
    m = s["rightIdxs"] >= 0
    assert m.sum() > 500
    assert abs(np.median(s["estimatedDepth"][m]) - g["fx"] * g["baseline"] / 24) < 0.05


def test_cpp_lba_matcher_and_median_descriptor_match_golden(golden_dir):
    """matchByProjectionRPredLBA (incl. the bestLev2 quirk of the left branch) and MapPoint::calcDescriptor."""
    g = np.load(os.path.join(golden_dir, "small_s0.npz")); p = np.load(os.path.join(golden_dir, "small_s0_lba.npz"))
    cfg = synth.CONFIGS["small"]
    fe = cpu_ref.Extractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    pts = {k[3:]: p[k] for k in p.files if k.startswith("pt_")}
    for rad in (4, 15):
        n, out = cpu_ref.match_lba(fe, pts, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"], p["unF"], p["unFR"],
                                   cfg["width"], cfg["height"], float(rad))
        assert n == int(p[f"lba_{rad}_n"]) and np.array_equal(out, p[f"lba_{rad}_out"])
        assert 0 < n < len(pts["desc"])
    assert np.array_equal(cpu_ref.median_descriptor(p["med_descs"], p["med_offsets"]), p["med_best"])


def _distortion_maps(w, h, k=0.15, dx=1.3, dy=-0.7):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / np.float32(w * w)
    return (xx + (xx - w / 2) * np.float32(k) * r2 + np.float32(dx)).astype(np.float32), \
           (yy + (yy - h / 2) * np.float32(k) * r2 + np.float32(dy)).astype(np.float32)


def test_image_preparation_matches_cv2():
    """cv::remap(INTER_LINEAR, CV_32F maps) + cvtColor(BGR2GRAY) restated, against cv2."""
    rng = np.random.default_rng(3)
    for (h, w) in ((240, 320), (97, 131)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mx, my = _distortion_maps(w, h)
        mx[0, :6] = [-5, -0.5, 0.0, w - 1.01, w - 0.5, w + 3]; my[0, :6] = [3, 3, 0.0, h - 1.01, 20, 5]   # border cases
        assert np.array_equal(cpu_ref.prepare_image(img, mx, my), ref_cv2.prepareImage(img, mx, my))
        assert np.array_equal(cpu_ref.prepare_image(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
        g = np.ascontiguousarray(img[..., 1])
        assert np.array_equal(cpu_ref.prepare_image(g, mx, my), cv2.remap(g, mx, my, cv2.INTER_LINEAR))


def _projection_case(seed=0, M=3000):
    rng = np.random.default_rng(seed)
    ang = 0.1
    Rm = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    T = np.eye(4); T[:3, :3] = Rm; T[:3, 3] = [0.3, -0.1, 0.5]
    wPos = np.concatenate([rng.uniform(-8, 8, (M, 2)), rng.uniform(-2, 30, (M, 1)), np.ones((M, 1))], axis=1)
    maxScaleDist = rng.uniform(5, 60, M).astype(np.float32)
    logScale = np.full(M, np.float32(math.log(np.float32(1.2))), np.float32)
    nScaleLev = np.full(M, 8, np.int32)
    return np.linalg.inv(T), wPos, maxScaleDist, logScale, nScaleLev


def test_world_to_frame_cpp_matches_python_restatement():
    """FeatureTracker::worldToFrame + MapPoint::predictScale, C++ oracle vs the Python restatement."""
    Pinv, wPos, msd, ls, nl = _projection_case()
    M = len(wPos)
    cam = (435.2047, 435.2047, 367.4517, 252.2008, 752, 480)
    pa = np.full((M, 2), -7, np.float32); sa = np.full(M, -7, np.int32); fa = np.zeros(M, np.uint8)
    ref_cv2.worldToFrame(wPos, Pinv, *cam, msd, ls, nl, pa, sa, fa)
    pb = np.full((M, 2), -7, np.float32); sb = np.full(M, -7, np.int32)
    fb = cpu_ref.world_to_frame(Pinv, *cam, wPos, msd, ls, nl, pb, sb)
    assert np.array_equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(sa, sb)
    assert 0.1 < fa.mean() < 0.9 and len(np.unique(sa[fa > 0])) >= 5


def _outlier_case(kl, kr, st, fx, cx, cy, baseline, seed=0, M=1500):
    """Inputs for findOutliersR / findOutliersMono on the keys and stereo arrays of one frame: map points back-projected from the
    keys they are matched to (small noise = inliers, large = outliers), stereo partners (true, wrong and missing ones), right-only
    matches, duplicates of a left key, invalid / out-of-frame points and points behind the camera."""
    rng = np.random.default_rng(seed)
    nl, nr = len(kl), len(kr)
    ang = 0.05
    T = np.eye(4)
    T[:3, :3] = [[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]]
    T[:3, 3] = [0.2, -0.05, 0.3]
    Tr = np.eye(4); Tr[0, 3] = -baseline
    TR = Tr @ T                                   # world -> right camera
    Tinv = np.linalg.inv(T)
    matches = np.full((M, 2), -1, np.int32)
    wPos = np.zeros((M, 4))
    for i in range(M):
        u = rng.random()
        first = int(rng.integers(0, nl))
        if i > 0 and rng.random() < 0.1 and matches[i - 1, 0] >= 0:
            first = int(matches[i - 1, 0])        # several map points on the same left key
        z = float(st["estimatedDepth"][first]) if st["estimatedDepth"][first] > 0 else float(rng.uniform(1, 40))
        if rng.random() < 0.05:
            z = float(rng.uniform(0.5, 60))
        pc = np.array([(kl["x"][first] - cx) / fx * z, (kl["y"][first] - cy) / fx * z, z, 1.0])
        pc[:3] += rng.normal(0, 0.002 if rng.random() < 0.7 else 0.2, 3) * z
        if rng.random() < 0.03:
            pc[2] = -abs(pc[2])                   # behind the camera
        wPos[i] = Tinv @ pc
        if u < 0.8:
            matches[i, 0] = first
            r = rng.random()
            if st["rightIdxs"][first] >= 0 and r < 0.7:
                matches[i, 1] = st["rightIdxs"][first]
            elif r < 0.85:
                matches[i, 1] = int(rng.integers(0, nr))
        elif u < 0.92:
            matches[i, 1] = int(rng.integers(0, nr))   # right-only match
            j = matches[i, 1]
            z = float(rng.uniform(1, 40))
            pr = np.array([(kr["x"][j] - cx) / fx * z, (kr["y"][j] - cy) / fx * z, z, 1.0])
            pr[:3] += rng.normal(0, 0.002 if rng.random() < 0.7 else 0.2, 3) * z
            wPos[i] = np.linalg.inv(TR) @ pr
    valid = (rng.random(M) < 0.95).astype(np.uint8)
    inF = (rng.random(M) < 0.95).astype(np.uint8)
    inFR = (rng.random(M) < 0.9).astype(np.uint8)
    return dict(pose=T, poseR=TR, wPos=wPos, matches=matches, valid=valid, inFrame=inF, inFrameR=inFR, thres=7.815,
                closeDist=float(np.float32(baseline) * np.float32(40)))


def _inv_sigma_factor(nLevels, imScale=1.2):
    """FeatureExtractor::InvSigmaFactor (src/FeatureExtractor.cpp:628-646): cumulative float32 scale, squared, inverted."""
    out, sc = [], np.float32(1)
    for l in range(nLevels):
        if l:
            sc = np.float32(sc * np.float32(imScale))
        out.append(np.float32(1) / np.float32(sc * sc))
    return np.array(out, np.float32)


def _run_outliers(fn, mode, c, cam, invSigma, kl, kr, st):
    """fn(mode, pose, poseR, fx, fy, cx, cy, thres, closeDist, invSigma, valid, wPos, inFrame, inFrameR, matches, outliers, kl, kr,
    depth, close, rightIdxs, leftIdxs) on fresh copies of the in-out arrays."""
    m = c["matches"].copy(); out = np.full(len(m), 7, np.uint8)
    arr = [st["estimatedDepth"].copy(), st["close"].astype(np.uint8).copy(), st["rightIdxs"].copy(), st["leftIdxs"].copy()]
    cnt = fn(mode, c["pose"], c["poseR"], *cam, c["thres"], c["closeDist"], invSigma, c["valid"], c["wPos"], c["inFrame"],
             c["inFrameR"], m, out, kl, kr, *arr)
    return cnt, m, out, arr


@pytest.mark.parametrize("mode", [0, 1])
def test_find_outliers_cpp_matches_python_restatement(mode):
    """FeatureTracker::findOutliersR / findOutliersMono (src/FeatureTracker.cpp:582-684), C++ oracle vs the
    Python restatement of the same sequential loop, on a frame's real keys and stereo arrays."""
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], seed=5)
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    st = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], h, kl, dl, kr, dr)
    cam = (cfg["fx"], cfg["fx"], w / 2.0, h / 2.0)
    c = _outlier_case(kl, kr, st, cfg["fx"], w / 2.0, h / 2.0, cfg["baseline"], seed=11 + mode)
    invSigma = _inv_sigma_factor(cfg["n_levels"])
    a = _run_outliers(ref_cv2.findOutliers, mode, c, cam, invSigma, kl, kr, st)
    b = _run_outliers(cpu_ref.find_outliers, mode, c, cam, invSigma, kl, kr, st)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for x, y in zip(a[3], b[3]):
        assert np.array_equal(x, y)
    nIn, nSt = b[0]
    assert nIn > 100 and (b[2] == 1).sum() > 100 and (b[2] == 7).sum() > 10         # inliers, outliers and untouched entries
    if mode == 0:
        assert nSt > 20 and (b[1][:, 1] != c["matches"][:, 1]).sum() > 5           # stereo inliers and rejected partners
        assert (b[3][1] != st["close"].astype(np.uint8)).sum() > 0


def _predict_case(seed=0, M=3000, nl=900, nr=880):
    """Inputs for removeOutOfFrameMPs / PredictMPsPosition: the projection case above for both cameras plus match bookkeeping."""
    rng = np.random.default_rng(seed)
    Pinv, wPos, msd, ls, nlev = _projection_case(seed, M)
    Tr = np.eye(4); Tr[0, 3] = -0.11
    valid = (rng.random(M) < 0.9).astype(np.uint8)
    matches = np.full((M, 2), -1, np.int32)
    pick = rng.random(M) < 0.6
    matches[pick, 0] = rng.permutation(max(nl, M))[:pick.sum()] % nl
    pick = rng.random(M) < 0.5
    matches[pick, 1] = rng.permutation(max(nr, M))[:pick.sum()] % nr
    matchedL = np.full(nl, -1, np.int32); matchedR = np.full(nr, -1, np.int32)
    for i in range(M):
        if matches[i, 0] >= 0: matchedL[matches[i, 0]] = i
        if matches[i, 1] >= 0: matchedR[matches[i, 1]] = i
    outliers = (rng.random(M) < 0.15).astype(np.uint8)
    return dict(toCamera=Pinv, toCameraR=Tr @ Pinv, wPos=wPos, msd=msd, ls=ls, nlev=nlev, valid=valid, matches=matches,
                matchedL=matchedL, matchedR=matchedR, outliers=outliers)


def _run_predict(fn, mode, c, cam):
    M = len(c["wPos"])
    st = dict(predL=np.full((M, 2), -7, np.float32), slL=np.full(M, -7, np.int32), inF=np.full(M, 7, np.uint8),
              predR=np.full((M, 2), -7, np.float32), slR=np.full(M, -7, np.int32), inFR=np.full(M, 7, np.uint8),
              matches=c["matches"].copy(), mL=c["matchedL"].copy(), mR=c["matchedR"].copy(), out=c["outliers"].copy())
    two = mode != 1
    keep = fn(mode, c["toCamera"], c["toCameraR"] if two else None, *cam, c["valid"], c["wPos"], c["msd"], c["ls"], c["nlev"],
              st["predL"], st["slL"], st["inF"], st["predR"] if two else None, st["slR"] if two else None, st["inFR"] if two else None,
              *( (st["matches"], st["mL"], st["mR"], st["out"]) if mode == 2 else (None, None, None, None) ))
    return keep, st


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_predict_mappoints_cpp_matches_python_restatement(mode):
    """removeOutOfFrameMPs / removeOutOfFrameMPsMono / PredictMPsPosition (src/FeatureTracker.cpp:910-1013),
    C++ oracle vs the Python restatement."""
    cam = (435.2047, 435.2047, 367.4517, 252.2008, 752, 480)
    c = _predict_case(seed=4 + mode, M=1200)
    ka, a = _run_predict(ref_cv2.predictMapPoints, mode, c, cam)
    kb, b = _run_predict(cpu_ref.predict_mappoints, mode, c, cam)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    if mode != 2:
        assert np.array_equal(ka, kb) and 0.05 < kb.mean() < 0.9 and (b["inF"] == 7).sum() > 20      # kept, dropped and unvisited points
    else:
        assert (b["matches"] != c["matches"]).sum() > 50 and b["out"].sum() == (c["outliers"] * (1 - c["valid"])).sum()
        assert (b["mL"] != c["matchedL"]).sum() > 20 and (b["mR"] != c["matchedR"]).sum() > 20
