"""GPU (-m gpu): the batched temporal matchers (gvs_match_projection_batch, gvs_track_temporal_batch) through the C ABI against
the reference's own matchers (oracle/_ref: FeatureMatcher.cpp compiled unmodified) and the restated map-point creation /
projection (oracle/cpu_ref.cpp), at EuRoC and KITTI size.  All comparisons are bit exact (indices, counts, float32 predictions).
"""
import numpy as np
import pytest

import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref, ref_build

pytestmark = pytest.mark.gpu


def _handles(cfg, n):
    w, h = cfg["width"], cfg["height"]
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
    feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
    return feL, feR, gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)


class _CpuMatcher:
    """matchByProjection* on the CPU: the compiled reference when oracle/_ref was built, else the restatement."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ref = ref_build.available()
        if self.ref:
            ref_build.use_default_hooks()
            self.feL = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"]); self.feR = ref_build.RefExtractor(cfg["n_features"], cfg["n_levels"])
            self.fm = ref_build.RefMatcher(cfg["fx"], cfg["baseline"], self.feL, self.feR, cfg["height"])
        else:
            self.fe = cpu_ref.Extractor(cfg["n_features"], cfg["n_levels"])

    def match(self, mode, mps, k, rad, mL, mR, mt):
        w, h = self.cfg["width"], self.cfg["height"]
        two = mode == 0
        if self.ref:
            gl = cpu_ref.assign_grid(k.keyPoints, w, h)
            gr = cpu_ref.assign_grid(k.rightKeyPoints, w, h) if two else None
            return self.fm.match_projection(mode, mps, k.keyPoints, k.Desc, k.rightKeyPoints if two else None, k.rightDesc if two else None,
                                            k.leftIdxs if two else None, k.rightIdxs if two else None, gl, gr, rad, mL, mR, mt)
        return cpu_ref.match_projection(self.fe, mode, mps, k.keyPoints, k.Desc, k.rightKeyPoints if two else None,
                                        k.rightDesc if two else None, k.leftIdxs if two else None, k.rightIdxs if two else None, w, h, rad,
                                        mL, mR, mt)


def _synthetic_mappoints(rng, kprev, shift, noise, with_right_from=None):
    """map points = a previous frame's keys, prediction = key position + motion + N(0, noise)."""
    M = len(kprev.keyPoints)
    pl = np.stack([kprev.keyPoints["x"] + shift[0], kprev.keyPoints["y"] + shift[1]], 1).astype(np.float32)
    pl += rng.normal(0, noise, (M, 2)).astype(np.float32)
    mps = dict(desc=kprev.Desc.copy(), predL=pl, scaleLevelL=kprev.keyPoints["octave"].astype(np.int32),
               inFrame=(rng.random(M) > 0.05).astype(np.uint8))
    if with_right_from is not None:
        d = float(with_right_from)
        mps["predR"] = (pl - np.array([d, 0], np.float32)).astype(np.float32)
        lr = mps["scaleLevelL"] + rng.integers(-1, 2, M)
        mps["scaleLevelR"] = np.clip(lr, 0, 7).astype(np.int32)
        mps["inFrameR"] = (rng.random(M) > 0.05).astype(np.uint8)
    return mps


@pytest.mark.parametrize("path", ["auto", "windows", "lists"])
@pytest.mark.parametrize("cfgname", ["euroc", "kitti"])
def test_match_projection_batch_against_reference(cfgname, path, monkeypatch):
    """RPred (rad 10 / 40 / 120), Mono (rad 1200), Radius (rad 120) for a batch of frames in ONE call each, with pre-filled in-out
    arrays, against the reference matcher run frame by frame.  The call has two device paths -- a thread per search window with the
    four best keys kept (tracking radii) and stored candidate lists (radii that cover much of the image); `path` forces one of them
    (GVS_PROJ_WIDE, read per call) or leaves the choice to the library: all three must agree with the reference."""
    if path != "auto":
        monkeypatch.setenv("GVS_PROJ_WIDE", "1" if path == "lists" else "0")
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    F = 4
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], F, seed=21)
    feL, feR, fm = _handles(cfg, F)
    fm.track_stereo_batch(L, R)
    keys = [gvs.tracked_keys(fm, f, w, h) for f in range(F)]
    cpu = _CpuMatcher(cfg)
    rng = np.random.default_rng(5)
    prev = [1, 0, 1, 2]
    per = []
    for f in range(F):
        p = prev[f]
        ox, oy = synth.sequence_offset(f); px, py = synth.sequence_offset(p)
        per.append(_synthetic_mappoints(rng, keys[p], (px - ox, py - oy), 1.0, with_right_from=cfg["disparity"]))
    off = np.concatenate([[0], np.cumsum([len(m["desc"]) for m in per])]).astype(np.int32)
    cat = {k: np.concatenate([m[k] for m in per]) for k in per[0]}
    M = int(off[-1])
    offL = feL.offsets(F); offR = feR.offsets(F)

    def fresh():
        mL = np.full(int(offL[-1]), -1, np.int32); mR = np.full(int(offR[-1]), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        # some map points and keys are already taken when the call starts (the second call of a frame, src/FeatureTracker.cpp:1239)
        mt[::13] = [0, -1]
        mL[::29] = 7; mR[::31] = 9
        return mL, mR, mt

    for mode, rads in ((gvs.GVS_MATCH_STEREO_PRED, (10.0, 40.0, 120.0)), (gvs.GVS_MATCH_MONO, (10.0, 1200.0)), (gvs.GVS_MATCH_RADIUS, (120.0,))):
        for rad in rads:
            mL, mR, mt = fresh()
            n = fm.match_projection_batch(mode, 0, F, cat, off, rad, mL, mR if mode == 0 else None, mt)
            oL, oR, ot = fresh()
            total = 0
            for f in range(F):
                sl = slice(off[f], off[f + 1])
                mps = {k: v[sl] for k, v in cat.items()}
                fl = oL[offL[f]:offL[f + 1]]; fr = oR[offR[f]:offR[f + 1]]; ft = ot[sl]
                on = cpu.match(mode, mps, keys[f], rad, fl, fr if mode == 0 else None, ft)
                assert n[f] == on, (mode, rad, f, n[f], on)
                total += on
            assert np.array_equal(mt, ot), (mode, rad)
            assert np.array_equal(mL, oL), (mode, rad)
            if mode == 0:
                assert np.array_equal(mR, oR), (mode, rad)
            assert total > 0.3 * M, (mode, rad, total, M)
    # a sub-range of the batch gives the same answer as the whole batch
    mL, mR, mt = fresh()
    n_all = fm.match_projection_batch(0, 0, F, cat, off, 10.0, mL, mR, mt)
    sl = slice(off[1], off[3])
    sub = {k: v[sl] for k, v in cat.items()}
    sL = np.full(int(offL[3] - offL[1]), -1, np.int32); sR = np.full(int(offR[3] - offR[1]), -1, np.int32); st = np.full((int(off[3] - off[1]), 2), -1, np.int32)
    fL, fR, ft = fresh()
    sL[:] = fL[offL[1]:offL[3]]; sR[:] = fR[offR[1]:offR[3]]; st[:] = ft[sl]
    n_sub = fm.match_projection_batch(0, 1, 2, sub, off[1:4] - off[1], 10.0, sL, sR, st)
    assert np.array_equal(n_sub, n_all[1:3]) and np.array_equal(st, mt[sl]) and np.array_equal(sL, mL[offL[1]:offL[3]])


@pytest.mark.parametrize("cfgname,F", [("euroc", 6), ("kitti", 4), ("small", 9), ("hires", 2)])   # hires: > 14 k keys per side, the resolve kernel's fallback
def test_track_temporal_batch(cfgname, F):
    """extract + stereo + temporal match, device resident: map points from the predecessor's stereo keys (initializeMap, update,
    removeOutOfFrameMPs restated in oracle/cpu_ref.cpp) and matchByProjectionRPred (compiled reference)."""
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], F, seed=33)
    feL, feR, fm = _handles(cfg, F)
    fm.track_stereo_batch(L, R)
    prev = [1] + list(range(F - 1))
    if F > 5:
        prev[5] = -1          # a frame without predecessor
    mo = synth.sequence_motion(cfg, F, prev)
    cpu = _CpuMatcher(cfg)
    ofe = cpu_ref.Extractor(cfg["n_features"], cfg["n_levels"])
    keys = [gvs.tracked_keys(fm, f, w, h) for f in range(F)]
    for rad in (10.0, 120.0):
        fm.track_temporal_batch(F, prev, mo["pose_prev"], mo["to_camera"], mo["to_camera_r"], mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h, rad)
        t = fm.temporal_download(0, F, preds=True)
        offL = feL.offsets(F); offR = feR.offsets(F)
        matched_total = 0
        for f in range(F):
            sl = slice(t["mp_offsets"][f], t["mp_offsets"][f + 1])
            if prev[f] < 0:
                assert sl.stop == sl.start and t["n_matches"][f] == 0
                continue
            kp = keys[prev[f]]
            o = cpu_ref.temporal_mappoints(ofe, kp.keyPoints, kp.estimatedDepth, mo["pose_prev"][f], mo["to_camera"][f], mo["to_camera_r"][f],
                                           mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h)
            assert np.array_equal(t["mp_key"][sl], o["key"]), (f, rad)
            assert np.array_equal(t["predL"][sl], o["predL"]) and np.array_equal(t["predR"][sl], o["predR"]), (f, rad)
            M = len(o["key"])
            assert M > 0.5 * (kp.estimatedDepth > 0).sum()
            mps = dict(desc=kp.Desc[o["key"]], predL=o["predL"], predR=o["predR"], scaleLevelL=o["scaleLevelL"], scaleLevelR=o["scaleLevelR"],
                       inFrame=np.ones(M, np.uint8), inFrameR=np.ones(M, np.uint8))
            oL = np.full(len(keys[f].keyPoints), -1, np.int32); oR = np.full(len(keys[f].rightKeyPoints), -1, np.int32)
            ot = np.full((M, 2), -1, np.int32)
            on = cpu.match(0, mps, keys[f], rad, oL, oR, ot)
            assert t["n_matches"][f] == on, (f, rad, t["n_matches"][f], on)
            assert np.array_equal(t["matches"][sl], ot), (f, rad)
            assert np.array_equal(t["matchedL"][offL[f]:offL[f + 1]], oL) and np.array_equal(t["matchedR"][offR[f]:offR[f + 1]], oR), (f, rad)
            matched_total += on
        assert matched_total > 0.3 * t["mp_offsets"][-1], (rad, matched_total)
    # a partial download equals the slice of the full one
    cnt = min(2, F - 1)
    part = fm.temporal_download(1, cnt)
    o1, o3 = t["mp_offsets"][1], t["mp_offsets"][1 + cnt]
    assert np.array_equal(part["matches"], t["matches"][o1:o3]) and np.array_equal(part["mp_key"], t["mp_key"][o1:o3])


def test_stale_batch_is_rejected():
    """grids / stereo arrays of an older batch must not be combined with newly extracted keys."""
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], 3, seed=2)
    feL, feR, fm = _handles(cfg, 3)
    fm.track_stereo_batch(L, R)
    k = gvs.tracked_keys(fm, 0, w, h)
    feL.extract_batch(L[:1])          # re-extract the left side only: the matcher's grids and stereo arrays are stale now
    feL.sync()
    M = 4
    mps = dict(desc=k.Desc[:M], predL=np.zeros((M, 2), np.float32), predR=np.zeros((M, 2), np.float32), scaleLevelL=np.zeros(M, np.int32),
               scaleLevelR=np.zeros(M, np.int32))
    mL = np.full(len(k.keyPoints), -1, np.int32); mR = np.full(len(k.rightKeyPoints), -1, np.int32); mt = np.full((M, 2), -1, np.int32)
    with pytest.raises(gvs.GvsError):
        fm.matchByProjectionRPred(mps, 0, mL, mR, mt, 10.0)
    with pytest.raises(gvs.GvsError):
        fm.stereo_download(0, 1)
    with pytest.raises(gvs.GvsError):
        fm.grid_download(0, 0)
    mo = synth.sequence_motion(cfg, 1, [-1])
    with pytest.raises(gvs.GvsError):
        fm.track_temporal_batch(1, [-1], mo["pose_prev"], mo["to_camera"], mo["to_camera_r"], mo["fx"], mo["fy"], mo["cx"], mo["cy"], w, h, 10.0)
    # a scale level outside the pyramid is an error, not an out-of-bounds read
    fm.track_stereo_batch(L, R)
    mps["scaleLevelL"] = np.full(M, 99, np.int32)
    with pytest.raises(gvs.GvsError):
        fm.matchByProjectionRPred(mps, 0, mL, mR, mt, 10.0)


def test_kitti_4096_frame_batch_chunked():
    """BASELINE config 3: a 4096-frame KITTI-shaped batch (extraction + stereo + temporal), in chunks of 512 through one handle
    set.  Size-independent properties: every copy of an image gives the same result in every chunk and slot (batch invariance),
    and the distinct images match the oracle."""
    cfg = synth.CONFIGS["kitti"]
    w, h = cfg["width"], cfg["height"]
    D, chunk, total = 16, 512, 4096
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], D, seed=77)
    feL, feR, fm = _handles(cfg, chunk)
    rng = np.random.default_rng(1)
    ref = {}
    seen = 0
    for c in range(total // chunk):
        pick = rng.integers(0, D, chunk)
        pick[:D] = np.arange(D)
        fm.track_stereo_batch(np.ascontiguousarray(L[pick]), np.ascontiguousarray(R[pick]))
        offL = feL.offsets(chunk); offR = feR.offsets(chunk)
        kl, dl = feL.download(0, chunk, int(offL[-1])); kr, dr = feR.download(0, chunk, int(offR[-1]))
        ri, li, de, cl = fm.stereo_download(0, chunk)
        for f in range(chunk):
            sl = slice(offL[f], offL[f + 1]); sr = slice(offR[f], offR[f + 1])
            sig = (kl[sl].tobytes(), dl[sl].tobytes(), kr[sr].tobytes(), dr[sr].tobytes(), ri[sl].tobytes(), li[sr].tobytes(), de[sl].tobytes(),
                   cl[sl].tobytes())
            if pick[f] not in ref:
                ref[int(pick[f])] = sig
            assert sig == ref[int(pick[f])], (c, f, int(pick[f]))
            seen += 1
    assert seen == total and len(ref) == D
    # the distinct images against the oracle
    o = cpu_ref.Extractor(cfg["n_features"], cfg["n_levels"]); o2 = cpu_ref.Extractor(cfg["n_features"], cfg["n_levels"])
    for i in (0, D - 1):
        okl, odl = o.extractKeysNew(L[i]); okr, odr = o2.extractKeysNew(R[i])
        s = cpu_ref.stereo_match(o, o2, cfg["fx"], cfg["baseline"], h, okl, odl, okr, odr)
        assert ref[i][0] == okl.tobytes() and ref[i][1] == odl.tobytes() and ref[i][2] == okr.tobytes() and ref[i][4] == s["rightIdxs"].tobytes()
        assert ref[i][6] == s["estimatedDepth"].tobytes()


def test_predict_scale_knife_edge_on_the_device():
    """dif == scalePyramid[k] exactly, for every level k: kernel == cpu_ref == ::log(double) arithmetic."""
    import math
    cfg = synth.CONFIGS["small"]
    fe = gvs.FeatureExtractor(nFeatures=1000, nLevels=8, width=640, height=480, max_images=1)
    fm = gvs.FeatureMatcher(400.0, 0.1, fe, None, 480)
    logScale = np.float32(math.log(float(np.float32(1.2))))
    msd = fe.scalePyramid.astype(np.float32).copy()
    M = len(msd)
    wp = np.tile(np.array([0.0, 0.0, 1.0, 1.0]), (M, 1))
    pred = np.zeros((M, 2), np.float32); lvl = np.zeros(M, np.int32)
    inF = fm.worldToFrame(np.eye(4), 400.0, 400.0, 320.0, 240.0, 640, 480, wp, msd, np.full(M, logScale, np.float32), np.full(M, 8, np.int32), pred, lvl)
    p2 = np.zeros((M, 2), np.float32); l2 = np.zeros(M, np.int32)
    cpu_ref.world_to_frame(np.eye(4), 400.0, 400.0, 320.0, 240.0, 640, 480, wp, msd, np.full(M, logScale, np.float32), np.full(M, 8, np.int32), p2, l2)
    expect = [min(max(int(math.ceil(math.log(float(v)) / float(logScale))), 0), 7) for v in msd]
    assert inF.all() and list(lvl) == expect and np.array_equal(lvl, l2)
