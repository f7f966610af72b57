"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures.

Bars: keypoint coordinates, FAST scores, octaves, descriptors, match indices, grids: bit exact.  Orientation,
sub-pixel disparity -> depth: the kernels use the same un-fused float32 operation order as the oracle, so they
are compared bit exact as well; the stated tolerances of the north star (1e-4 rad, 1e-3 px) are asserted too so
that a future relaxation has a written bound.
"""
import hashlib
import math
import os

import numpy as np
import pytest

import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = math.degrees(1e-4)
DEPTH_REL_TOL = 1e-5


def _handles(cfg, n, **kw):
    w, h = cfg["width"], cfg["height"]
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
    args.update(kw)
    feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
    return feL, feR, gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)


def _oracle_frame(cfg, L, R):
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    s = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], cfg["height"], kl, dl, kr, dr)
    return oL, oR, kl, dl, kr, dr, s


def _assert_keys(gk, gd, ok, od):
    assert len(gk) == len(ok)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(gk[f], ok[f]), f
    assert np.max(np.abs(gk["angle"] - ok["angle"]), initial=0) <= ANGLE_TOL_DEG
    assert np.array_equal(gk["angle"], ok["angle"])
    assert np.array_equal(gd, od)


def _assert_frame(fm, f, cfg, kl, dl, kr, dr, s):
    w, h = cfg["width"], cfg["height"]
    k = gvs.tracked_keys(fm, f, w, h)
    _assert_keys(k.keyPoints, k.Desc, kl, dl)
    _assert_keys(k.rightKeyPoints, k.rightDesc, kr, dr)
    assert np.array_equal(k.rightIdxs, s["rightIdxs"]) and np.array_equal(k.leftIdxs, s["leftIdxs"])
    m = s["rightIdxs"] >= 0
    assert np.all(np.abs(k.estimatedDepth[m] - s["estimatedDepth"][m]) <= DEPTH_REL_TOL * s["estimatedDepth"][m])
    assert np.array_equal(k.estimatedDepth, s["estimatedDepth"]) and np.array_equal(k.close, s["close"])
    for side, kk, gg in ((0, kl, k.lkeyGrid), (1, kr, k.rkeyGrid)):
        o = cpu_ref.assign_grid(kk, w, h)
        assert gg["yGrids"] == o["yGrids"] and gg["xMult"] == o["xMult"] and gg["yMult"] == o["yMult"]
        assert np.array_equal(gg["start"], o["start"]) and np.array_equal(gg["items"], o["items"])
    return k


@pytest.mark.parametrize("name", ["small_s0", "small_s1_ramp", "euroc_s0", "kitti_s0"])
def test_golden_fixture(name, golden_dir):
    """CUDA path vs the committed fixtures (made with cv2 under the reference's control flow)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = synth.CONFIGS[str(g["cfg"])]
    L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], int(g["seed"]), ramp=bool(g["ramp"]))
    assert hashlib.sha256(L.tobytes()).hexdigest() == str(g["sha_left"])
    feL, feR, fm = _handles(cfg, 1)
    k = gvs.extractORBAndStereoMatch(fm, L, R)
    _assert_keys(k.keyPoints, k.Desc, g["kpL"], g["descL"])
    _assert_keys(k.rightKeyPoints, k.rightDesc, g["kpR"], g["descR"])
    assert np.array_equal(k.rightIdxs, g["rightIdxs"]) and np.array_equal(k.leftIdxs, g["leftIdxs"])
    assert np.array_equal(k.estimatedDepth, g["depth"]) and np.array_equal(k.close, g["close"])
    assert np.array_equal(k.lkeyGrid["start"], g["gridL_start"]) and np.array_equal(k.lkeyGrid["items"], g["gridL_items"])
    assert np.array_equal(k.rkeyGrid["start"], g["gridR_start"]) and np.array_equal(k.rkeyGrid["items"], g["gridR_items"])
    for l in range(cfg["n_levels"]):
        assert hashlib.sha256(feL.level_image(0, l).tobytes()).hexdigest() == str(g["pyr_sha"][l])
    assert np.array_equal(feL.scalePyramid, g["scalePyramid"]) and np.array_equal(feL.featurePerLevel, g["featurePerLevel"])


@pytest.mark.parametrize("cfgname,seeds", [("small", [20, 21, 22, 23]), ("euroc", [5, 6]), ("kitti", [7])])
def test_stages_against_oracle(cfgname, seeds):
    """Stage by stage on fresh seeds: pyramid, blur, ordered FAST candidates, SSC selection, then everything."""
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    Ls, Rs = synth.stereo_batch(w, h, cfg["disparity"], seeds)
    feL, feR, fm = _handles(cfg, len(seeds))
    fm.track_stereo_batch(Ls, Rs)
    for f in range(len(seeds)):
        oL, oR, kl, dl, kr, dr, s = _oracle_frame(cfg, Ls[f], Rs[f])
        for fe, o in ((feL, oL), (feR, oR)):
            for l in range(cfg["n_levels"]):
                assert np.array_equal(fe.level_image(f, l), o.level_image(l)), ("pyramid", l)
                assert np.array_equal(fe.level_image(f, l, True), o.level_image(l, True)), ("blur", l)
                x, y, sc = fe.level_candidates(f, l); c = o.level_keys(l, False)
                assert np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(sc, c["response"].astype(np.int32)), ("cand", l)
                x, y, sc = fe.level_kept(f, l); c = o.level_keys(l, True)
                assert np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(sc, c["response"].astype(np.int32)), ("kept", l)
        _assert_frame(fm, f, cfg, kl, dl, kr, dr, s)


def test_single_image_api_and_tables():
    """gvs_extract == FeatureExtractor::extractKeysNew; public tables == the reference constructor's."""
    cfg = synth.CONFIGS["euroc"]
    L, _ = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], 9)
    fe = gvs.FeatureExtractor(nFeatures=2000, width=cfg["width"], height=cfg["height"])  # mono default ctor (System.cpp:31)
    o = cpu_ref.Extractor(2000)
    kp, d = fe.extractKeysNew(L)
    ok, od = o.extractKeysNew(L)
    _assert_keys(kp, d, ok, od)
    for a in ("scalePyramid", "scaleInvPyramid", "sigmaFactor", "InvSigmaFactor", "scaledPatchSize", "featurePerLevel"):
        assert np.array_equal(getattr(fe, a), getattr(o, a)), a
    assert fe.featurePerLevel.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]


def test_edge_cases_blank_and_sparse_images():
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    fe = gvs.FeatureExtractor(nFeatures=300, nLevels=4, width=w, height=h)
    o = cpu_ref.Extractor(300, 4)
    # blank image: no corner anywhere -> zero keypoints, outputs untouched (src/FeatureExtractor.cpp:498-499)
    kp, d = fe.extractKeysNew(np.full((h, w), 77, np.uint8))
    assert len(kp) == 0 and len(d) == 0
    # sparse image: fewer candidates than the quota on every level -> no ssc, emission order is kept (:610)
    img = np.full((h, w), 40, np.uint8)
    rng = np.random.default_rng(1)
    for _ in range(12):
        x, y = int(rng.integers(30, w - 60)), int(rng.integers(30, h - 60))
        img[y:y + 17, x:x + 23] = 200
    kp, d = fe.extractKeysNew(img)
    ok, od = o.extractKeysNew(img)
    assert 0 < len(ok) < 300
    _assert_keys(kp, d, ok, od)
    # weak corners only: every cell falls back to the min threshold (cv::FAST(.., 7), :592-593)
    img2 = (40 + 12 * ((np.add.outer(np.arange(h) // 9, np.arange(w) // 9) % 2))).astype(np.uint8)
    kp, d = fe.extractKeysNew(img2)
    ok, od = o.extractKeysNew(img2)
    assert len(ok) > 0 and ok["response"].max() < 20
    _assert_keys(kp, d, ok, od)
    # saturated noise: maximum corner density
    img3 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    kp, d = fe.extractKeysNew(img3)
    ok, od = o.extractKeysNew(img3)
    _assert_keys(kp, d, ok, od)


def _killer(n):
    k = n // 2
    a = [0] * n
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i; a[i] = k + i
        a[k + i - 1] = 2 * i
    return np.array(a)


@pytest.mark.parametrize("case", ["ties2", "ties8", "wide", "sorted", "reversed", "equal", "organ", "killer", "tiny", "big"])
def test_sort_and_ssc_emulation(case):
    """The device emulation of cv::sortIdx (libstdc++ introsort incl. median-of-3, depth limit, heapsort) and of ssc
    on crafted candidate lists, against std::sort / the oracle's ssc."""
    cfg = synth.CONFIGS["kitti"]
    fe = gvs.FeatureExtractor(nFeatures=2000, width=cfg["width"], height=cfg["height"])
    o = cpu_ref.Extractor(2000)
    lw, lh = fe.level_size(0)
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    n = {"tiny": 500, "big": 20000, "killer": 6000}.get(case, 5000)
    if case == "ties2": s = rng.integers(20, 22, n)
    elif case == "ties8": s = rng.integers(7, 15, n)
    elif case == "wide": s = rng.integers(7, 200, n)
    elif case == "sorted": s = np.sort(rng.integers(7, 120, n))
    elif case == "reversed": s = np.sort(rng.integers(7, 120, n))[::-1].copy()
    elif case == "equal": s = np.full(n, 33)
    elif case == "organ": s = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]) % 200 + 7
    elif case == "killer": s = _killer(n) * 240 // n + 7      # median-of-3 killer: exercises the heapsort fallback
    else: s = rng.integers(7, 60, n)
    s = np.clip(s, 1, 254).astype(np.int32)
    n = len(s)
    pos = rng.permutation((lw - 40) * (lh - 40))[:n]
    x = (pos % (lw - 40) + 20).astype(np.int32); y = (pos // (lw - 40) + 20).astype(np.int32)
    sidx, kept = fe.debug_select(0, x, y, s)
    osrt, okept = o.ssc(int(fe.featurePerLevel[0]), lw, lh, x, y, s.astype(np.float32))
    if n > fe.featurePerLevel[0]:
        assert np.array_equal(sidx, osrt)
        assert np.array_equal(kept, okept)
    else:
        assert np.array_equal(kept, np.arange(n))


def test_ssc_small_widths_use_hash_grid():
    """Few candidates above the quota force the binary search down to tiny widths (cell grid larger than shared memory)."""
    cfg = synth.CONFIGS["kitti"]
    fe = gvs.FeatureExtractor(nFeatures=2000, width=cfg["width"], height=cfg["height"])
    o = cpu_ref.Extractor(2000)
    lw, lh = fe.level_size(0)
    rng = np.random.default_rng(3)
    for n in (440, 480, 700):
        pos = rng.permutation((lw - 40) * (lh - 40))[:n]
        x = (pos % (lw - 40) + 20).astype(np.int32); y = (pos // (lw - 40) + 20).astype(np.int32)
        s = rng.integers(7, 90, n).astype(np.int32)
        sidx, kept = fe.debug_select(0, x, y, s)
        osrt, okept = o.ssc(int(fe.featurePerLevel[0]), lw, lh, x, y, s.astype(np.float32))
        assert np.array_equal(sidx, osrt) and np.array_equal(kept, okept), n
    # pairs of points 2 px apart: width 2 merges each pair (count < Kmin), so the search ends at width 1 (cell = half a pixel)
    for pairs in (225, 235):      # 450 keys: inside [Kmin, Kmax] at width 1; 470: above Kmax -> bracket exhausted, last pass returned
        pos = rng.permutation(((lw - 60) // 8) * ((lh - 60) // 8))[:pairs]
        px = (pos % ((lw - 60) // 8) * 8 + 24).astype(np.int32); py = (pos // ((lw - 60) // 8) * 8 + 24).astype(np.int32)
        x = np.concatenate([px, px + 2]); y = np.concatenate([py, py])
        s = rng.integers(7, 60, 2 * pairs).astype(np.int32)
        sidx, kept = fe.debug_select(0, x, y, s)
        osrt, okept = o.ssc(int(fe.featurePerLevel[0]), lw, lh, x, y, s.astype(np.float32))
        assert len(okept) == 2 * pairs
        assert np.array_equal(sidx, osrt) and np.array_equal(kept, okept), pairs


def test_projection_matchers_against_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_s0.npz")); p = np.load(os.path.join(golden_dir, "small_s0_proj.npz"))
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    feL, feR, fm = _handles(cfg, 1)
    gvs.extractORBAndStereoMatch(fm, g["left"], g["right"])
    mps = {k[3:]: p[k] for k in p.files if k.startswith("mp_")}
    nl, nr, M = len(g["kpL"]), len(g["kpR"]), len(mps["desc"])
    for rad in (4, 10, 120):
        mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32); mt = np.full((M, 2), -1, np.int32); mt[::17] = [3, -1]
        n = fm.matchByProjectionRPred(mps, 0, mL, mR, mt, float(rad))
        assert n == int(p[f"rpred_{rad}_n"]) and np.array_equal(mL, p[f"rpred_{rad}_mL"]) and np.array_equal(mR, p[f"rpred_{rad}_mR"])
        assert np.array_equal(mt, p[f"rpred_{rad}_mt"])
        mL = np.full(nl, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = fm.matchByProjectionMono(mps, 0, mL, mt, float(rad))
        assert n == int(p[f"mono_{rad}_n"]) and np.array_equal(mL, p[f"mono_{rad}_mL"]) and np.array_equal(mt, p[f"mono_{rad}_mt"])
    last = dict(desc=g["descR"], predL=np.stack([g["kpR"]["x"], g["kpR"]["y"]], 1), scaleLevelL=g["kpR"]["octave"])
    mL = np.full(nl, -1, np.int32); mt = np.full((nr, 2), -1, np.int32)
    n = fm.matchByRadius(last, 0, mL, mt, 120.0)
    assert n == int(p["radius_n"]) and np.array_equal(mL, p["radius_mL"]) and np.array_equal(mt[:, 0], p["radius_match"])


def test_mono_temporal_config4():
    """BASELINE config 4: 752x480 mono, default ctor (2000 features), M = previous frame's keys, rad 1200 and 10."""
    cfg = synth.CONFIGS["mono"]
    w, h = cfg["width"], cfg["height"]
    seq = synth.shifted_sequence(w, h, 2, seed=4)
    fe = gvs.FeatureExtractor(nFeatures=2000, width=w, height=h, max_images=2)
    fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], fe, None, h)
    fe.extract_batch(seq)
    fm.assign_grids_batch(2, w, h)
    off = fe.offsets(2)
    k0, d0 = fe.download(0, 1, int(off[1]))
    k1, d1 = fe.download(1, 1, int(off[2] - off[1]))
    o = cpu_ref.Extractor(2000)
    ok1, od1 = o.extractKeysNew(seq[1])
    _assert_keys(k1, d1, ok1, od1)
    rng = np.random.default_rng(0)
    M = len(k0)
    mps = dict(desc=d0, predL=np.stack([k0["x"] - 2, k0["y"] - 1], 1).astype(np.float32) + rng.normal(0, 1, (M, 2)).astype(np.float32),
               scaleLevelL=k0["octave"].astype(np.int32), inFrame=np.ones(M, np.uint8))
    for rad in (10.0, 1200.0):
        mL = np.full(len(k1), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = fm.matchByProjectionMono(mps, 1, mL, mt, rad)
        oL = np.full(len(k1), -1, np.int32); ot = np.full((M, 2), -1, np.int32)
        on = cpu_ref.match_projection(o, 1, mps, ok1, od1, None, None, None, None, w, h, rad, oL, None, ot)
        assert n == on and np.array_equal(mL, oL) and np.array_equal(mt, ot), rad
        assert n > 0.5 * M


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (1000, 32), dtype=np.uint8); b = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    b[:10] = a[:10]
    cfg = synth.CONFIGS["small"]
    feL, feR, fm = _handles(cfg, 1)
    d = fm.DescriptorDistance(a, b)
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(axis=1))


def test_batch_invariance_and_idempotence_full_size():
    """BASELINE configs[1] at its full batch size: a 256-frame batch gives, frame by frame, exactly what single-frame
    calls give (frames are independent), twice in a row (idempotence), with sorted responses inside each level."""
    cfg = synth.CONFIGS["euroc"]
    w, h = cfg["width"], cfg["height"]
    F = 256
    base = synth.stereo_batch(w, h, cfg["disparity"], range(100, 108))
    Ls = np.ascontiguousarray(np.tile(base[0], (F // 8, 1, 1))); Rs = np.ascontiguousarray(np.tile(base[1], (F // 8, 1, 1)))
    feL, feR, fm = _handles(cfg, F)
    digests = []
    for _ in range(2):
        fm.track_stereo_batch(Ls, Rs)
        offL, offR = feL.offsets(F), feR.offsets(F)
        kl, dl = feL.download(0, F, int(offL[F])); kr, dr = feR.download(0, F, int(offR[F]))
        ridx, lidx, dep, cl = fm.stereo_download(0, F)
        digests.append(hashlib.sha256(b"".join(a.tobytes() for a in (offL, offR, kl, dl, kr, dr, ridx, lidx, dep, cl))).hexdigest())
    assert digests[0] == digests[1]
    # frames f and f+8 are the same image: identical slices
    for f in range(8):
        a = slice(offL[f], offL[f + 1])
        for rep in range(1, F // 8):
            b = slice(offL[f + 8 * rep], offL[f + 8 * rep + 1])
            assert kl[a].tobytes() == kl[b].tobytes() and dl[a].tobytes() == dl[b].tobytes()
            assert ridx[a].tobytes() == ridx[b].tobytes() and dep[a].tobytes() == dep[b].tobytes()
    # against the oracle on the 8 distinct frames
    for f in range(8):
        oL, oR, okl, odl, okr, odr, s = _oracle_frame(cfg, base[0][f], base[1][f])
        _assert_keys(kl[offL[f]:offL[f + 1]], dl[offL[f]:offL[f + 1]], okl, odl)
        assert np.array_equal(ridx[offL[f]:offL[f + 1]], s["rightIdxs"]) and np.array_equal(lidx[offR[f]:offR[f + 1]], s["leftIdxs"])
    # size-independent properties: per level, responses of SSC-selected keys are non-increasing; octaves ascend
    for f in range(0, F, 37):
        k = kl[offL[f]:offL[f + 1]]
        assert np.all(np.diff(k["octave"]) >= 0)
        for l in range(cfg["n_levels"]):
            r = k["response"][k["octave"] == l]
            assert np.all(np.diff(r) <= 0)
        m = ridx[offL[f]:offL[f + 1]]
        li = lidx[offR[f]:offR[f + 1]]
        # leftIdxs / rightIdxs are mutually consistent where both are set
        j = np.nonzero(li >= 0)[0]
        assert np.all(m[li[j]] == j)


def test_errors_are_loud():
    cfg = synth.CONFIGS["small"]
    feL, feR, fm = _handles(cfg, 1)
    with pytest.raises(gvs.GvsError):
        fm.stereo_match_batch(1)  # nothing extracted yet
    with pytest.raises(gvs.GvsError):
        feL.extractKeysNew(np.zeros((10, 10), np.uint8))  # wrong size
    with pytest.raises(gvs.GvsError):
        gvs.FeatureExtractor(width=64, height=64)  # pyramid level smaller than the border


def test_lba_matcher_and_median_descriptor(golden_dir):
    """SURVEY 8(f) rank 1 (the backend's matcher): gvs_match_projection_lba and gvs_median_descriptor."""
    g = np.load(os.path.join(golden_dir, "small_s0.npz")); p = np.load(os.path.join(golden_dir, "small_s0_lba.npz"))
    cfg = synth.CONFIGS["small"]
    feL, feR, fm = _handles(cfg, 1)
    pts = {k[3:]: p[k] for k in p.files if k.startswith("pt_")}
    for rad in (4, 15):
        n, out = fm.matchByProjectionRPredLBA(pts, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"], p["unF"],
                                             p["unFR"], cfg["width"], cfg["height"], float(rad))
        assert n == int(p[f"lba_{rad}_n"]) and np.array_equal(out, p[f"lba_{rad}_out"])
    assert np.array_equal(fm.median_descriptor(p["med_descs"], p["med_offsets"]), p["med_best"])
    # larger random case against the C++ oracle
    rng = np.random.default_rng(2)
    sizes = rng.integers(1, 120, 40)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    descs = rng.integers(0, 256, (int(offs[-1]), 32), dtype=np.uint8)
    descs[:, 8:] = descs[0, 8:]   # make distances small enough to tie often
    assert np.array_equal(fm.median_descriptor(descs, offs), cpu_ref.median_descriptor(descs, offs))


def test_hires_stress_config5():
    """BASELINE config 5: 2048x1536, 8 levels, 8000 features -- the largest sizes: more than 6144 candidates per level
    (global-memory class of the selection kernel, 32-bit indices) and more keys per image than the shared-memory sort holds."""
    cfg = synth.CONFIGS["hires"]
    L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], 1)
    feL, feR, fm = _handles(cfg, 1)
    fm.track_stereo_batch(L[None].copy(), R[None].copy())
    oL, oR, kl, dl, kr, dr, s = _oracle_frame(cfg, L, R)
    assert len(kl) > 7000 and max(len(oL.level_keys(l, False)["x"]) for l in range(8)) > 6144
    _assert_frame(fm, 0, cfg, kl, dl, kr, dr, s)


def test_raw_frames_remap_and_gray_on_device():
    """BGR frames + rectification maps in, same keys as remap + cvtColor + extractKeysNew on the CPU."""
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], 41)
    rng = np.random.default_rng(0)
    bgr = np.stack([L, R, np.clip(L.astype(int) + rng.integers(-9, 10, L.shape), 0, 255).astype(np.uint8)], axis=-1)  # B, G, R planes
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / np.float32(w * w)
    mx = (xx + (xx - w / 2) * np.float32(0.12) * r2 + np.float32(0.8)).astype(np.float32)
    my = (yy + (yy - h / 2) * np.float32(0.12) * r2 - np.float32(0.4)).astype(np.float32)
    fe = gvs.FeatureExtractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=2)
    o = cpu_ref.Extractor(cfg["n_features"], cfg["n_levels"])
    batch = np.ascontiguousarray(np.stack([bgr, bgr[:, ::-1].copy()]))
    for maps in ((mx, my), (None, None)):
        fe.set_rectify_maps(*maps)
        fe.extract_batch_raw(batch)
        off = fe.offsets(2)
        for i in range(2):
            gray = cpu_ref.prepare_image(batch[i], *maps)
            assert np.array_equal(fe.level_image(i, 0), gray)
            ok, od = o.extractKeysNew(gray)
            kp, d = fe.download(i, 1, int(off[i + 1] - off[i]))
            _assert_keys(kp, d, ok, od)
    # single-channel raw frames through the same entry point
    fe.set_rectify_maps(mx, my)
    fe.extract_batch_raw(np.ascontiguousarray(np.stack([L, R])))
    assert np.array_equal(fe.level_image(1, 0), cpu_ref.prepare_image(R, mx, my))


def test_strided_host_images_and_resident_input():
    """Host images with a row stride (2-D copy path) and the device-resident input path give the same keys as the tight path."""
    import torch
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    Ls, _ = synth.stereo_batch(w, h, cfg["disparity"], [50, 51, 52])
    fe = gvs.FeatureExtractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=3)
    fe.extract_batch(Ls)
    off = fe.offsets(3)
    ref = fe.download(0, 3, int(off[3]))
    ref = (ref[0].copy(), ref[1].copy())
    padded = np.zeros((3, h + 2, w + 24), np.uint8)
    padded[:, :h, :w] = Ls
    lib = gvs.load_library()
    assert lib.gvs_extract_batch_host(fe.h, padded.ctypes.data, padded.strides[0], padded.strides[1], 3) == 0
    off2 = fe.offsets(3)
    got = fe.download(0, 3, int(off2[3]))
    assert np.array_equal(off, off2) and got[0].tobytes() == ref[0].tobytes() and np.array_equal(got[1], ref[1])
    dev, pitch, stride = fe.input_layout()

    class _Ext:
        __cuda_array_interface__ = {"shape": (3, stride // pitch, pitch), "typestr": "|u1", "data": (dev, False), "version": 2}
    store = torch.as_tensor(_Ext(), device="cuda")
    store.zero_()
    store[:, :h, :w].copy_(torch.from_numpy(Ls).cuda())
    torch.cuda.synchronize()
    fe.extract_batch_resident(3)
    off3 = fe.offsets(3)
    got = fe.download(0, 3, int(off3[3]))
    assert np.array_equal(off, off3) and got[0].tobytes() == ref[0].tobytes() and np.array_equal(got[1], ref[1])


def test_world_to_frame_projection():
    """batched FeatureTracker::worldToFrame + MapPoint::predictScale (double precision) against the oracle."""
    from test_oracle_vs_cv2 import _projection_case
    Pinv, wPos, msd, ls, nl = _projection_case(seed=3, M=20000)
    M = len(wPos)
    cam = (435.2047, 435.2047, 367.4517, 252.2008, 752, 480)
    cfg = synth.CONFIGS["small"]
    feL, feR, fm = _handles(cfg, 1)
    pa = np.full((M, 2), -7, np.float32); sa = np.full(M, -7, np.int32)
    fa = fm.worldToFrame(Pinv, *cam, wPos, msd, ls, nl, pa, sa)
    pb = np.full((M, 2), -7, np.float32); sb = np.full(M, -7, np.int32)
    fb = cpu_ref.world_to_frame(Pinv, *cam, wPos, msd, ls, nl, pb, sb)
    assert np.array_equal(fa, fb) and np.array_equal(sa, sb)
    assert np.max(np.abs(pa - pb), initial=0) <= 1e-3 and np.array_equal(pa, pb)   # tolerance of the north star, and bit exact


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_predict_mappoints(mode):
    """FeatureTracker::removeOutOfFrameMPs / removeOutOfFrameMPsMono / PredictMPsPosition on the device against
    the oracle's sequential loops (double precision projection bit exact, match bookkeeping exact)."""
    from test_oracle_vs_cv2 import _predict_case, _run_predict
    cam = (435.2047, 435.2047, 367.4517, 252.2008, 752, 480)
    c = _predict_case(seed=9 + mode, M=20000, nl=2100, nr=2050)
    feL, feR, fm = _handles(synth.CONFIGS["small"], 1)
    ka, a = _run_predict(fm.predictMapPoints, mode, c, cam)
    kb, b = _run_predict(cpu_ref.predict_mappoints, mode, c, cam)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    if mode != 2:
        assert np.array_equal(ka, kb) and 0.05 < kb.mean() < 0.9
    else:
        assert (b["matches"] != c["matches"]).sum() > 500
    with pytest.raises(gvs.GvsError):
        bad = c["matches"].copy(); bad[0, 0] = 10 ** 6
        _run_predict(fm.predictMapPoints, 2, dict(c, matches=bad), cam)


@pytest.mark.parametrize("mode", [0, 1])
def test_find_outliers(mode):
    """FeatureTracker::findOutliersR / findOutliersMono on the device-resident frame against the oracle's
    sequential loop: outlier flags, match pairs, counts and the cleared stereo arrays bit exact; the device-resident stereo
    arrays follow (a second download sees the cleared pairs)."""
    from test_oracle_vs_cv2 import _outlier_case, _run_outliers, _inv_sigma_factor
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], seed=5)
    feL, feR, fm = _handles(cfg, 1)
    k = gvs.extractORBAndStereoMatch(fm, L, R)
    oL, oR, kl, dl, kr, dr, st = _oracle_frame(cfg, L, R)
    assert np.array_equal(k.rightIdxs, st["rightIdxs"]) and np.array_equal(feL.InvSigmaFactor, _inv_sigma_factor(cfg["n_levels"]))
    cam = (cfg["fx"], cfg["fx"], w / 2.0, h / 2.0)
    for seed in (1, 2):
        # the oracle starts from the stereo arrays the device currently holds (they change from call to call)
        cur = dict(estimatedDepth=k.estimatedDepth.copy(), close=k.close.copy(), rightIdxs=k.rightIdxs.copy(), leftIdxs=k.leftIdxs.copy())
        c = _outlier_case(kl, kr, cur, cfg["fx"], w / 2.0, h / 2.0, cfg["baseline"], seed=seed, M=2500)
        ref = _run_outliers(cpu_ref.find_outliers, mode, c, cam, feL.InvSigmaFactor, kl, kr, cur)
        m = c["matches"].copy(); out = np.full(len(m), 7, np.uint8)
        if mode == 0:
            cnt = fm.findOutliersR(0, c["pose"], c["poseR"], *cam, c["thres"], c["closeDist"], c["valid"], c["wPos"], c["inFrame"],
                                   c["inFrameR"], m, out, keys=k)
        else:
            cnt = (fm.findOutliersMono(0, c["pose"], *cam, c["thres"], c["valid"], c["wPos"], c["inFrame"], m, out), 0)
        assert tuple(cnt) == tuple(ref[0]), (cnt, ref[0])
        assert np.array_equal(m, ref[1]) and np.array_equal(out, ref[2])
        if mode == 0:
            for got, want in zip((k.estimatedDepth, k.close, k.rightIdxs, k.leftIdxs), ref[3]):
                assert np.array_equal(got, want)
            r2, l2, d2, c2 = fm.stereo_download(0, 1)     # the device state itself
            assert np.array_equal(r2, ref[3][2]) and np.array_equal(l2, ref[3][3]) and np.array_equal(d2, ref[3][0]) and np.array_equal(c2, ref[3][1])
            assert (ref[3][2] != cur["rightIdxs"]).sum() > 0
    with pytest.raises(gvs.GvsError):
        bad = np.array([[10 ** 6, -1]], np.int32)
        fm.findOutliersMono(0, np.eye(4), *cam, 7.815, None, np.ones((1, 4)), None, bad, np.zeros(1, np.uint8))


def test_cuda_against_the_compiled_reference():
    """The CUDA path against the reference's OWN source (oracle/_ref/libref_frontend.so: FeatureExtractor.cpp and
    FeatureMatcher.cpp compiled unmodified; built in the CPU container, the .so travels), on fresh seeds; with the real
    OpenCV primitives (cv2 hooks) when cv2 is importable on this box."""
    from oracle import ref_build
    if not ref_build.available():
        pytest.skip("oracle/_ref/libref_frontend.so was not built (needs the reference sources in the build container)")
    try:
        import cv2  # noqa: F401
        ref_build.use_cv2_hooks()
    except Exception:
        ref_build.use_default_hooks()
    try:
        for cfgname, seeds in (("small", [40, 41, 42]), ("euroc", [43, 44]), ("kitti", [45])):
            cfg = synth.CONFIGS[cfgname]
            w, h = cfg["width"], cfg["height"]
            L, R = synth.stereo_batch(w, h, cfg["disparity"], seeds)
            feL, feR, fm = _handles(cfg, len(seeds))
            fm.track_stereo_batch(L, R)
            for f in range(len(seeds)):
                rL = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"]); rR = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"])
                kl, dl = rL.extractKeysNew(L[f]); kr, dr = rR.extractKeysNew(R[f])
                s = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], rL, rR, h).findStereoMatchesORB2R(kl, dl, kr, dr)
                assert (s["rightIdxs"] >= 0).sum() > 0.3 * len(kl)
                _assert_frame(fm, f, cfg, kl, dl, kr, dr, s)
    finally:
        ref_build.use_default_hooks()


@pytest.mark.parametrize("w,h,nF,nL,scale,edge,maxF,minF,seed", [
    (640, 480, 500, 6, 1.5, 19, 20, 7, 50),      # coarse pyramid
    (401, 303, 800, 5, 1.3, 19, 30, 10, 51),     # odd sizes, other thresholds
    (1280, 720, 1500, 8, 1.2, 19, 20, 7, 52),    # 720p
    (752, 480, 1000, 12, 1.1, 19, 20, 7, 53),    # 12 fine levels
    (320, 240, 300, 4, 1.2, 25, 40, 5, 54),      # wider edge, far-apart thresholds
    (512, 384, 4000, 3, 2.0, 19, 12, 3, 55),     # many features, factor 2, low thresholds
    (333, 257, 150, 2, 1.2, 19, 20, 20, 56),     # max == min threshold
    (1191, 360, 1000, 7, 1.1, 19, 19, 13, 57),   # last cell column 1 px wide and word aligned (found by tools/fuzz_parity.py)
    (1054, 174, 2000, 4, 1.3, 22, 17, 10, 58),   # same, other alignment
    (360, 700, 800, 5, 1.2, 19, 20, 7, 59),      # portrait: more than 64 key-grid rows
])
def test_parameter_sweep(w, h, nF, nL, scale, edge, maxF, minF, seed):
    """Non-default constructor parameters and image sizes: CUDA == oracle (== compiled reference when it is present)."""
    from oracle import ref_build
    L, R = synth.stereo_pair(w, h, 10, seed)
    feL = gvs.FeatureExtractor(nF, nL, scale, edge, 31, maxF, minF, width=w, height=h, max_images=1)
    feR = gvs.FeatureExtractor(nF, nL, scale, edge, 31, maxF, minF, width=w, height=h, max_images=1)
    fm = gvs.FeatureMatcher(300.0, 0.2, feL, feR, h)
    fm.track_stereo_batch(L[None], R[None])
    cfg = dict(width=w, height=h)
    oL, oR = cpu_ref.Extractor(nF, nL, scale, edge, 31, maxF, minF), cpu_ref.Extractor(nF, nL, scale, edge, 31, maxF, minF)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    assert len(kl) > 0.5 * min(nF, 1500)
    s = cpu_ref.stereo_match(oL, oR, 300.0, 0.2, h, kl, dl, kr, dr)
    _assert_frame(fm, 0, cfg, kl, dl, kr, dr, s)
    if ref_build.available():
        rL, rR = ref_build.RefExtractor(nF, nL, scale, edge, 31, maxF, minF), ref_build.RefExtractor(nF, nL, scale, edge, 31, maxF, minF)
        kl2, dl2 = rL.extractKeysNew(L); kr2, dr2 = rR.extractKeysNew(R)
        s2 = ref_build.RefMatcher(300.0, 0.2, rL, rR, h).findStereoMatchesORB2R(kl2, dl2, kr2, dr2)
        _assert_frame(fm, 0, cfg, kl2, dl2, kr2, dr2, s2)


def test_every_width_alignment_of_the_last_cell():
    """Widths 300..371 at one level: the last cell column takes every width from 1 px up and every alignment against the
    4-px words of the FAST kernel (a 1-px last cell that is one word wide once broke an index division)."""
    rng = np.random.default_rng(9)
    for w in range(300, 372):
        h = 110 + (w % 7)
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img[:, : w // 2] = synth.stereo_pair(w, h, 0, w)[0][:, : w // 2]
        fe = gvs.FeatureExtractor(600, 1, 1.2, 19, 31, 20, 7, width=w, height=h, max_images=1)
        o = cpu_ref.Extractor(600, 1, 1.2, 19, 31, 20, 7)
        gk, gd = fe.extractKeysNew(img)
        ok, od = o.extractKeysNew(img)
        x, y, s = fe.level_candidates(0, 0); c = o.level_keys(0, False)
        assert np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(s, c["response"].astype(np.int32)), w
        _assert_keys(gk, gd, ok, od)


def test_frame_pipeline_tickets_and_slots():
    """more tickets than slots through gvs_pipeline; every ticket equals the oracle; slot discipline errors."""
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    pipe = gvs.StereoPipeline(cfg["fx"], cfg["baseline"], width=w, height=h, frames_per_slot=2, n_slots=3,
                              nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    batches = [synth.stereo_batch(w, h, cfg["disparity"], [60 + 2 * t, 61 + 2 * t][: 1 + t % 2]) for t in range(7)]
    tickets = []

    def collect(t):
        L, R = batches[t]
        n = len(L)
        r = pipe.download(t, n)
        for f in range(n):
            _, _, kl, dl, kr, dr, s = _oracle_frame(cfg, L[f], R[f])
            a, b = r["offL"][f], r["offL"][f + 1]; c, d = r["offR"][f], r["offR"][f + 1]
            _assert_keys(r["kpL"][a:b], r["descL"][a:b], kl, dl)
            _assert_keys(r["kpR"][c:d], r["descR"][c:d], kr, dr)
            assert np.array_equal(r["rightIdxs"][a:b], s["rightIdxs"]) and np.array_equal(r["leftIdxs"][c:d], s["leftIdxs"])
            assert np.array_equal(r["estimatedDepth"][a:b], s["estimatedDepth"]) and np.array_equal(r["close"][a:b], s["close"])
        pipe.release(t)

    for t in range(len(batches)):
        if t >= 3:
            collect(t - 3)                      # free the slot this ticket needs
        assert pipe.submit(*batches[t]) == t
        tickets.append(t)
    with pytest.raises(gvs.GvsError):           # all three slots busy
        pipe.submit(*batches[0])
    for t in range(len(batches) - 3, len(batches)):
        collect(t)
    with pytest.raises(gvs.GvsError):           # released ticket
        pipe.wait(0, 1)
    with pytest.raises(gvs.GvsError):           # more frames than a slot holds
        pipe.submit(*synth.stereo_batch(w, h, cfg["disparity"], [1, 2, 3]))
    pipe.close()


@pytest.mark.parametrize("cfgname", ["euroc", "kitti", "hires", "small"])
def test_tma_staging_is_active_and_changes_nothing(cfgname, monkeypatch):
    """The fused pyramid and the FAST kernel stage their tiles with TMA box loads on the BASELINE shapes; the same handles
    built with the ordinary-load kernels (GVS_NO_FUSED_PYR / GVS_NO_FAST_TMA, read when a handle is created) give
    byte-identical keys and descriptors."""
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    n = 3
    imgs = np.ascontiguousarray(synth.stereo_batch(w, h, cfg["disparity"], range(40, 40 + n))[0])
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)

    def run():
        fe = gvs.FeatureExtractor(**args)
        fe.extract_batch(imgs); fe.sync()
        off = fe.offsets(n)
        kp, desc = fe.download(0, n, int(off[n]))
        return fe.tma_stages(), off.tobytes() + kp.tobytes() + desc.tobytes()

    stages, with_tma = run()
    assert stages == 3, f"TMA staging not active on {cfgname}: {stages}"
    monkeypatch.setenv("GVS_NO_FUSED_PYR", "1"); monkeypatch.setenv("GVS_NO_FAST_TMA", "1")
    stages, without = run()
    assert stages == 0 and without == with_tma


@pytest.mark.parametrize("cfgname", ["euroc", "small"])
def test_fast_several_strips_per_cta(cfgname, monkeypatch):
    """GVS_FAST_ITEMS > 1: a CTA of the FAST kernel walks several strips and requests the next strip's tile (TMA) while it writes the
    current strip's records.  Not the default (it measures slower), but it must give the same bytes."""
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    n = 9   # (more than 8 images: smaller batches replay a captured graph, which would keep the first setting)
    imgs = np.ascontiguousarray(synth.stereo_batch(w, h, cfg["disparity"], range(70, 70 + n))[0])
    fe = gvs.FeatureExtractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)

    def run():
        fe.extract_batch(imgs); fe.sync()
        off = fe.offsets(n)
        kp, desc = fe.download(0, n, int(off[n]))
        return off.tobytes() + kp.tobytes() + desc.tobytes()

    ref = run()
    for k in ("2", "3", "7"):
        monkeypatch.setenv("GVS_FAST_ITEMS", k)
        assert run() == ref, k
