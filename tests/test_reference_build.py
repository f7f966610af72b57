"""The reference's OWN FeatureExtractor.cpp / FeatureMatcher.cpp (oracle/_ref/libref_frontend.so: compiled unmodified from
/root/reference against the OpenCV / Eigen stand-ins of oracle/shim) against the oracle restatement and the golden vectors.

Two hook tables: "cv2" routes resize / FAST / GaussianBlur / sortIdx / fastAtan2 to the real OpenCV (python cv2), i.e.
reference source + real OpenCV primitives; "default" uses the oracle's restated primitives (no cv2 needed).
Skipped when neither the prebuilt library nor the reference sources are present.
"""
import os

import numpy as np
import pytest

from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_build

pytestmark = pytest.mark.skipif(not ref_build.available(), reason="oracle/_ref not built and /root/reference absent")

try:
    import cv2  # noqa: F401
    HOOKS = ["default", "cv2"]
except Exception:  # pragma: no cover
    HOOKS = ["default"]


@pytest.fixture(params=HOOKS)
def hooks(request):
    (ref_build.use_cv2_hooks if request.param == "cv2" else ref_build.use_default_hooks)()
    yield request.param
    ref_build.use_default_hooks()


def _same_keys(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def _pair(cfg):
    args = (cfg["n_features"], cfg["n_levels"])
    return ref_build.RefExtractor(*args), ref_build.RefExtractor(*args), cpu_ref.Extractor(*args), cpu_ref.Extractor(*args)


def test_constructor_tables_match_reference_object():
    for nF, nL, s in ((2000, 8, 1.2), (1000, 8, 1.2), (300, 4, 1.2), (500, 6, 1.5), (1200, 12, 1.1)):
        r = ref_build.RefExtractor(nF, nL, s); o = cpu_ref.Extractor(nF, nL, s)
        for t in ("scalePyramid", "scaleInvPyramid", "sigmaFactor", "InvSigmaFactor", "scaledPatchSize", "featurePerLevel"):
            assert np.array_equal(getattr(r, t), getattr(o, t)), t


@pytest.mark.parametrize("name,seed", [("small", 0), ("small", 1), ("euroc", 0), ("kitti", 0), ("kitti", 3)])
def test_extract_and_stereo_match_oracle(hooks, name, seed):
    """extractKeysNew x2 + findStereoMatchesORB2R of the compiled reference == oracle, bit for bit."""
    cfg = synth.CONFIGS[name]
    L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], seed)
    rL, rR, oL, oR = _pair(cfg)
    kl, dl = rL.extractKeysNew(L); kr, dr = rR.extractKeysNew(R)
    kl2, dl2 = oL.extractKeysNew(L); kr2, dr2 = oR.extractKeysNew(R)
    assert len(kl) > 0.9 * cfg["n_features"]
    assert _same_keys(kl, kl2) and _same_keys(kr, kr2) and np.array_equal(dl, dl2) and np.array_equal(dr, dr2)
    for l in range(cfg["n_levels"]):
        assert np.array_equal(rL.level_image(l), oL.level_image(l))
    fm = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], rL, rR, cfg["height"])
    s = fm.findStereoMatchesORB2R(kl, dl, kr, dr)
    s2 = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], cfg["height"], kl, dl, kr, dr)
    assert (s["rightIdxs"] >= 0).sum() > 0.3 * len(kl)
    for k in ("rightIdxs", "leftIdxs", "estimatedDepth", "close"):
        assert np.array_equal(s[k], s2[k]), k


def test_ramp_image_fallback_threshold(hooks):
    """Low-contrast ramp: cells fall back to minFastThreshold (src/FeatureExtractor.cpp:592-593)."""
    cfg = synth.CONFIGS["small"]
    L, _ = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], 1, ramp=True)
    rL, _, oL, _ = _pair(cfg)
    k1, d1 = rL.extractKeysNew(L); k2, d2 = oL.extractKeysNew(L)
    assert _same_keys(k1, k2) and np.array_equal(d1, d2)


def test_degenerate_images(hooks):
    cfg = synth.CONFIGS["small"]
    rL, _, oL, _ = _pair(cfg)
    flat = np.full((cfg["height"], cfg["width"]), 77, np.uint8)
    k1, _ = rL.extractKeysNew(flat); k2, _ = oL.extractKeysNew(flat)
    assert len(k1) == 0 and len(k2) == 0
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (cfg["height"], cfg["width"]), dtype=np.uint8)      # far more candidates than the quota
    k1, d1 = rL.extractKeysNew(noise); k2, d2 = oL.extractKeysNew(noise)
    assert len(k1) > 0 and _same_keys(k1, k2) and np.array_equal(d1, d2)


def test_ssc_member_on_adversarial_points(hooks):
    """FeatureExtractor::ssc (public member) on tie-heavy responses: the kept set depends on cv::sortIdx's tie order."""
    rng = np.random.default_rng(11)
    r = ref_build.RefExtractor(1000, 8); o = cpu_ref.Extractor(1000, 8)
    for n, hi, want in ((5000, 6, 217), (3000, 250, 150), (900, 2, 100), (20000, 40, 217)):
        x = rng.integers(16, 736, n).astype(np.int32); y = rng.integers(16, 464, n).astype(np.int32)
        resp = rng.integers(7, 7 + hi, n).astype(np.float32)
        kept_ref = r.ssc(want, 752, 480, x, y, resp)
        _, kept_ora = o.ssc(want, 752, 480, x, y, resp)
        assert np.array_equal(kept_ref, kept_ora)


def test_reference_matches_committed_golden(hooks, golden_dir):
    """The fixtures of tests/golden (made by oracle/ref_cv2.py) are reproduced by the compiled reference."""
    for name in ("small_s0", "euroc_s0", "kitti_s0"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        cfg = synth.CONFIGS[str(g["cfg"])]
        L, R = synth.stereo_pair(cfg["width"], cfg["height"], cfg["disparity"], int(g["seed"]), ramp=bool(g["ramp"]))
        rL, rR, _, _ = _pair(cfg)
        kl, dl = rL.extractKeysNew(L); kr, dr = rR.extractKeysNew(R)
        for f, gf in (("x", "x"), ("y", "y"), ("angle", "angle"), ("response", "response"), ("octave", "octave"), ("size", "size")):
            assert np.array_equal(kl[f], g["kpL"][gf]) and np.array_equal(kr[f], g["kpR"][gf]), (name, f)
        assert np.array_equal(dl, g["descL"]) and np.array_equal(dr, g["descR"])
        s = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], rL, rR, cfg["height"]).findStereoMatchesORB2R(kl, dl, kr, dr)
        assert np.array_equal(s["rightIdxs"], g["rightIdxs"]) and np.array_equal(s["leftIdxs"], g["leftIdxs"])
        assert np.array_equal(s["estimatedDepth"], g["depth"]) and np.array_equal(s["close"], g["close"])


def _small_frame(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_s0.npz"))
    cfg = synth.CONFIGS["small"]
    rL, rR, oL, _ = _pair(cfg)
    fm = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], rL, rR, cfg["height"])
    gl = cpu_ref.assign_grid(g["kpL"], cfg["width"], cfg["height"]); gr = cpu_ref.assign_grid(g["kpR"], cfg["width"], cfg["height"])
    # the CSR grids equal the golden ones (restated FeatureTracker::assignKeysToGrids over real data)
    assert np.array_equal(gl["start"], g["gridL_start"]) and np.array_equal(gl["items"], g["gridL_items"])
    assert np.array_equal(gr["start"], g["gridR_start"]) and np.array_equal(gr["items"], g["gridR_items"])
    return g, cfg, fm, oL, gl, gr


def test_projection_matchers_of_the_reference(golden_dir):
    """matchByProjectionRPred / Mono / matchByRadius of the compiled reference == golden == oracle."""
    g, cfg, fm, oL, gl, gr = _small_frame(golden_dir)
    p = np.load(os.path.join(golden_dir, "small_s0_proj.npz"))
    w, h = cfg["width"], cfg["height"]
    mps = {k[3:]: p[k] for k in p.files if k.startswith("mp_")}
    nl, nr, M = len(g["kpL"]), len(g["kpR"]), len(mps["desc"])
    for rad in (4, 10, 120):
        mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32); mt = np.full((M, 2), -1, np.int32); mt[::17] = [3, -1]
        n = fm.match_projection(0, mps, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"], gl, gr, float(rad), mL, mR, mt)
        assert n == int(p[f"rpred_{rad}_n"]) and np.array_equal(mL, p[f"rpred_{rad}_mL"]) and np.array_equal(mR, p[f"rpred_{rad}_mR"])
        assert np.array_equal(mt, p[f"rpred_{rad}_mt"])
        mL = np.full(nl, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = fm.match_projection(1, mps, g["kpL"], g["descL"], None, None, None, None, gl, None, float(rad), mL, None, mt)
        assert n == int(p[f"mono_{rad}_n"]) and np.array_equal(mL, p[f"mono_{rad}_mL"]) and np.array_equal(mt, p[f"mono_{rad}_mt"])
    last = dict(desc=g["descR"], predL=np.stack([g["kpR"]["x"], g["kpR"]["y"]], 1), scaleLevelL=g["kpR"]["octave"])
    mL = np.full(nl, -1, np.int32); mt = np.full((nr, 2), -1, np.int32)
    n = fm.match_projection(2, last, g["kpL"], g["descL"], None, None, None, None, gl, None, 120.0, mL, None, mt)
    assert n == int(p["radius_n"]) and np.array_equal(mL, p["radius_mL"]) and np.array_equal(mt[:, 0], p["radius_match"])


def test_projection_matchers_random_cases_vs_oracle(golden_dir):
    """Fresh random map points (not the fixture's): reference == oracle for all three matchers."""
    g, cfg, fm, oL, gl, gr = _small_frame(golden_dir)
    w, h = cfg["width"], cfg["height"]
    nl, nr = len(g["kpL"]), len(g["kpR"])
    for seed in range(4):
        rng = np.random.default_rng(100 + seed)
        M = 400
        src = rng.integers(0, nl, M)
        desc = g["descL"][src].copy()
        flip = rng.integers(0, 256, (M, 32)) < 12
        desc ^= (flip * rng.integers(1, 256, (M, 32))).astype(np.uint8)
        predL = np.stack([g["kpL"]["x"][src], g["kpL"]["y"][src]], 1) + rng.normal(0, 3, (M, 2)).astype(np.float32)
        predR = predL - np.array([cfg["disparity"], 0], np.float32)
        mps = dict(desc=desc, predL=predL.astype(np.float32), predR=predR.astype(np.float32),
                   scaleLevelL=np.clip(g["kpL"]["octave"][src] + rng.integers(-1, 2, M), 0, cfg["n_levels"] - 1).astype(np.int32),
                   scaleLevelR=np.clip(g["kpL"]["octave"][src] + rng.integers(-1, 2, M), 0, cfg["n_levels"] - 1).astype(np.int32),
                   inFrame=(rng.random(M) < 0.9).astype(np.uint8), inFrameR=(rng.random(M) < 0.8).astype(np.uint8))
        for mode in (0, 1):
            for rad in (3.0, 12.0):
                a = [np.full(nl, -1, np.int32), np.full(nr, -1, np.int32), np.full((M, 2), -1, np.int32)]
                b = [x.copy() for x in a]
                kr, dr, li, ri = (g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"]) if mode == 0 else (None, None, None, None)
                n1 = fm.match_projection(mode, mps, g["kpL"], g["descL"], kr, dr, li, ri, gl, gr if mode == 0 else None, rad,
                                         a[0], a[1] if mode == 0 else None, a[2])
                n2 = cpu_ref.match_projection(oL, mode, mps, g["kpL"], g["descL"], kr, dr, li, ri, w, h, rad,
                                              b[0], b[1] if mode == 0 else None, b[2])
                assert n1 == n2 and n1 > 0 and all(np.array_equal(x, y) for x, y in zip(a, b)), (seed, mode, rad)


def test_lba_matcher_of_the_reference(golden_dir):
    """matchByProjectionRPredLBA of the compiled reference == golden (incl. the bestLev2 quirk of the left branch)."""
    g, cfg, fm, oL, gl, gr = _small_frame(golden_dir)
    p = np.load(os.path.join(golden_dir, "small_s0_lba.npz"))
    pts = {k[3:]: p[k] for k in p.files if k.startswith("pt_")}
    for rad in (4, 15):
        n, out = fm.match_lba(pts, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"], p["unF"], p["unFR"],
                              gl, gr, float(rad))
        assert n == int(p[f"lba_{rad}_n"]) and np.array_equal(out, p[f"lba_{rad}_out"])


def test_descriptor_distance_of_the_reference():
    rng = np.random.default_rng(2)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        want = int(np.unpackbits(a ^ b).sum())
        assert ref_build.descriptor_distance(a, b) == want == cpu_ref.lib().ora_descriptor_distance(cpu_ref._p(a), cpu_ref._p(b))
