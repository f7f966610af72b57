"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/cpu_ref.cpp (liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ora_extractor_create.restype = C.c_void_p
        _lib.ora_extractor_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.ora_extractor_destroy.argtypes = [C.c_void_p]
        _lib.ora_fast_atan2.restype = C.c_float
        _lib.ora_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.ora_stereo_frames.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Extractor:
    """Mirror of FeatureExtractor (reference include/FeatureExtractor.h:52-97)."""

    def __init__(self, nFeatures=2000, nLevels=8, imScale=1.2, edgeThreshold=19, patchSize=31,
                 maxFastThreshold=20, minFastThreshold=7):
        self.params = (int(nFeatures), int(nLevels), float(imScale), int(edgeThreshold), int(patchSize),
                       int(maxFastThreshold), int(minFastThreshold))
        self.nFeatures, self.nLevels = int(nFeatures), int(nLevels)
        self.h = C.c_void_p(lib().ora_extractor_create(*self.params))
        L = self.nLevels
        self.scalePyramid = np.zeros(L, np.float32)
        self.scaleInvPyramid = np.zeros(L, np.float32)
        self.sigmaFactor = np.zeros(L, np.float32)
        self.InvSigmaFactor = np.zeros(L, np.float32)
        self.scaledPatchSize = np.zeros(L, np.int32)
        self.featurePerLevel = np.zeros(L, np.int32)
        self.umax = np.zeros(16, np.int32)
        lib().ora_extractor_tables(self.h, _p(self.scalePyramid), _p(self.scaleInvPyramid), _p(self.sigmaFactor),
                                   _p(self.InvSigmaFactor), _p(self.scaledPatchSize), _p(self.featurePerLevel),
                                   _p(self.umax))

    def __del__(self):
        try:
            lib().ora_extractor_destroy(self.h)
        except Exception:
            pass

    def extractKeysNew(self, image: np.ndarray):
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        cap = max(4 * self.nFeatures, 1024)
        kp = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().ora_extract(self.h, _p(image), w, h, w, _p(kp), _p(desc), cap)
        assert n >= 0
        return kp[:n].copy(), desc[:n].copy()

    def level_image(self, level: int, blurred: bool = False) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        lib().ora_level_size(self.h, level, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        lib().ora_level_image(self.h, level, int(blurred), _p(out))
        return out

    def level_keys(self, level: int, kept: bool):
        n = lib().ora_level_count(self.h, level, int(kept))
        x = np.zeros(n, np.int32); y = np.zeros(n, np.int32); r = np.zeros(n, np.float32)
        a = np.zeros(n, np.float32); m10 = np.zeros(n, np.int32); m01 = np.zeros(n, np.int32)
        lib().ora_level_keys(self.h, level, int(kept), _p(x), _p(y), _p(r), _p(a), _p(m10), _p(m01))
        return dict(x=x, y=y, response=r, angle=a, m10=m10, m01=m01)

    def ssc(self, numRetPoints: int, cols: int, rows: int, x, y, resp):
        """FeatureExtractor::ssc (reference src/FeatureExtractor.cpp:368-468): (sorted permutation, kept input indices)."""
        x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32); r = np.ascontiguousarray(resp, np.float32)
        n = len(x)
        out = np.zeros(max(n, 1), np.int32); srt = np.zeros(max(n, 1), np.int32)
        k = lib().ora_ssc(self.h, numRetPoints, cols, rows, n, _p(x), _p(y), _p(r), _p(out), _p(srt))
        return srt[:n].copy(), out[:k].copy()

    def ssc_passes(self, level: int) -> int:
        return lib().ora_level_ssc_passes(self.h, level)


def stereo_match(feL: Extractor, feR: Extractor, fx: float, baseline: float, imageHeight: int, kl, dl, kr, dr):
    """findStereoMatchesORB2R (reference src/FeatureMatcher.cpp:528-708) on the pyramids of the last extract."""
    NL, NR = len(kl), len(kr)
    ridx = np.zeros(NL, np.int32); lidx = np.zeros(NR, np.int32)
    dep = np.zeros(NL, np.float32); close = np.zeros(NL, np.uint8)
    stats = np.zeros(2, np.int32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl); dr = np.ascontiguousarray(dr)
    lib().ora_stereo_match(feL.h, feR.h, C.c_double(fx), C.c_float(baseline), imageHeight, _p(kl), _p(dl), NL,
                           _p(kr), _p(dr), NR, _p(ridx), _p(lidx), _p(dep), _p(close), _p(stats))
    return dict(rightIdxs=ridx, leftIdxs=lidx, estimatedDepth=dep, close=close.astype(bool),
                hamming_tests=int(stats[0]), accepted=int(stats[1]))


def assign_grid(kps, width: int, height: int):
    """assignKeysToGrids (reference src/FeatureTracker.cpp:28-54) in CSR form."""
    kps = np.ascontiguousarray(kps)
    n = len(kps)
    start = np.zeros(64 * 513 + 1, np.int32)   # yGrids = ceil(64 * height / width): more than 64 rows for portrait images
    items = np.zeros(max(n, 1), np.int32)
    xm, ym = C.c_float(), C.c_float()
    yg = lib().ora_assign_grid(_p(kps), n, width, height, _p(start), _p(items), C.byref(xm), C.byref(ym))
    return dict(xGrids=64, yGrids=yg, xMult=xm.value, yMult=ym.value, start=start[:yg * 64 + 1].copy(), items=items[:n].copy())


def match_projection(fe: Extractor, mode: int, mps: dict, kl, dl, kr, dr, leftIdxs, rightIdxs, width, height, rad,
                     matchedL, matchedR, matches):
    """mode 0/1/2 = matchByProjectionRPred / matchByProjectionMono / matchByRadius
    (reference src/FeatureMatcher.cpp:254-389 / 391-456 / 458-526).  Updates matchedL/R, matches in place."""
    M = len(mps["desc"])
    f = lambda k, dt: np.ascontiguousarray(mps[k], dt) if k in mps and mps[k] is not None else None
    desc = f("desc", np.uint8); pL = f("predL", np.float32); pR = f("predR", np.float32)
    lL = f("scaleLevelL", np.int32); lR = f("scaleLevelR", np.int32)
    iF = f("inFrame", np.uint8); iR = f("inFrameR", np.uint8)
    kl = np.ascontiguousarray(kl); dl = np.ascontiguousarray(dl)
    NR = 0 if kr is None else len(kr)
    if kr is not None:
        kr = np.ascontiguousarray(kr); dr = np.ascontiguousarray(dr)
    li = None if leftIdxs is None else np.ascontiguousarray(leftIdxs, np.int32)
    ri = None if rightIdxs is None else np.ascontiguousarray(rightIdxs, np.int32)
    assert matchedL.dtype == np.int32 and matches.dtype == np.int32 and matches.shape == (M, 2)
    return lib().ora_match_projection(fe.h, mode, M, _p(desc), _p(pL), _p(pR), _p(lL), _p(lR), _p(iF), _p(iR),
                                      _p(kl), _p(dl), len(kl), _p(kr), _p(dr), NR, _p(li), _p(ri), width, height,
                                      C.c_float(rad), _p(matchedL), _p(matchedR), _p(matches))


def match_lba(fe: Extractor, pts: dict, kl, dl, kr, dr, leftIdxs, rightIdxs, unMatchedF, unMatchedFR, width, height, rad):
    """matchByProjectionRPredLBA (reference src/FeatureMatcher.cpp:66-252), flattened inputs: (nMatches, out (M,2))."""
    c = lambda a, dt: np.ascontiguousarray(a, dt)
    M = len(pts["desc"])
    out = np.zeros((M, 2), np.int32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    args = [c(pts["desc"], np.uint8), c(pts["predL"], np.float32), c(pts["predR"], np.float32), c(pts["predScale"], np.int32),
            c(pts["lastL"], np.float32), c(pts["lastR"], np.float32)]
    n = lib().ora_match_lba(fe.h, M, *[_p(a) for a in args], _p(kl), _p(c(dl, np.uint8)), len(kl), _p(kr), _p(c(dr, np.uint8)), len(kr),
                            _p(c(leftIdxs, np.int32)), _p(c(rightIdxs, np.int32)), _p(c(unMatchedF, np.int32)),
                            _p(c(unMatchedFR, np.int32)), width, height, C.c_float(rad), _p(out))
    return n, out


def median_descriptor(descs: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """MapPoint::calcDescriptor (reference src/Map.cpp:145-210) for len(offsets)-1 map points."""
    descs = np.ascontiguousarray(descs, np.uint8); offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros(len(offsets) - 1, np.int32)
    lib().ora_median_descriptor(_p(descs), _p(offsets), len(out), _p(out))
    return out


def prepare_image(image: np.ndarray, mapx=None, mapy=None) -> np.ndarray:
    """cv::remap(INTER_LINEAR) + cvtColor(BGR2GRAY) restated (reference src/VIOSlam.cpp:282-301, src/FeatureTracker.cpp:1131-1135)."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    ch = 1 if image.ndim == 2 else image.shape[2]
    out = np.zeros((h, w), np.uint8)
    mx = None if mapx is None else np.ascontiguousarray(mapx, np.float32)
    my = None if mapy is None else np.ascontiguousarray(mapy, np.float32)
    lib().ora_prepare_image(_p(image), w, h, ch, _p(mx), _p(my), _p(out))
    return out


def world_to_frame(poseInv, fx, fy, cx, cy, width, height, wPos, maxScaleDist, logScale, nScaleLev, pred, scaleLevel):
    """worldToFrame + predictScale (reference src/FeatureTracker.cpp:685-741, src/Map.cpp:13-23); pred/scaleLevel in-out."""
    c = lambda a, dt: np.ascontiguousarray(a, dt)
    M = len(wPos)
    inF = np.zeros(M, np.uint8)
    assert pred.dtype == np.float32 and scaleLevel.dtype == np.int32
    lib().ora_world_to_frame(_p(c(poseInv, np.float64)), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), width, height, M,
                             _p(c(wPos, np.float64)), _p(c(maxScaleDist, np.float32)), _p(c(logScale, np.float32)),
                             _p(c(nScaleLev, np.int32)), _p(pred), _p(scaleLevel), _p(inF))
    return inF


def predict_mappoints(mode, toCamera, toCameraR, fx, fy, cx, cy, width, height, valid, wPos, maxScaleDist, logScale, nScaleLev,
                      predL, scaleLevelL, inFrame, predR=None, scaleLevelR=None, inFrameR=None, matches=None, matchedL=None,
                      matchedR=None, outliers=None):
    """removeOutOfFrameMPs (mode 0) / ...Mono (1) / PredictMPsPosition (2), reference src/FeatureTracker.cpp:910-1013; the in-out
    arrays are modified in place; returns keep (modes 0/1) or None."""
    c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
    M = len(wPos)
    keep = np.zeros(M, np.uint8)
    k = [c(toCamera, np.float64), c(toCameraR, np.float64), c(valid, np.uint8), c(wPos, np.float64), c(maxScaleDist, np.float32),
         c(logScale, np.float32), c(nScaleLev, np.int32)]
    lib().ora_predict_mappoints(mode, _p(k[0]), _p(k[1]), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), width, height, M,
                                _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]), _p(predL), _p(scaleLevelL), _p(inFrame), _p(predR),
                                _p(scaleLevelR), _p(inFrameR), _p(keep), _p(matches), _p(matchedL), _p(matchedR), _p(outliers))
    return keep if mode != 2 else None


def find_outliers(mode, pose, poseR, fx, fy, cx, cy, thres, closeDist, invSigma, valid, wPos, inFrame, inFrameR, matches, outliers,
                  kl, kr, depth, close, rightIdxs, leftIdxs):
    """findOutliersR (mode 0) / findOutliersMono (mode 1), reference src/FeatureTracker.cpp:582-684; matches, outliers and the
    four stereo arrays are modified in place; returns (nInliers increments, nStereo)."""
    c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
    for a, dt in ((matches, np.int32), (outliers, np.uint8)):
        assert a.dtype == dt and a.flags.c_contiguous
    cnt = np.zeros(2, np.int32)
    keep = [c(pose, np.float64), c(poseR, np.float64), c(invSigma, np.float32), c(valid, np.uint8), c(wPos, np.float64),
            c(inFrame, np.uint8), c(inFrameR, np.uint8), np.ascontiguousarray(kl), None if kr is None else np.ascontiguousarray(kr)]
    lib().ora_find_outliers(mode, _p(keep[0]), _p(keep[1]), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                            C.c_double(thres), C.c_float(closeDist), _p(keep[2]), len(matches), _p(keep[3]), _p(keep[4]), _p(keep[5]),
                            _p(keep[6]), _p(matches), _p(outliers), _p(keep[7]), _p(keep[8]), _p(depth), _p(close), _p(rightIdxs),
                            _p(leftIdxs), _p(cnt))
    return int(cnt[0]), int(cnt[1])


def temporal_mappoints(fe: Extractor, kl, depth, posePrev, toCamera, toCameraR, fx, fy, cx, cy, width, height):
    """Map points of a frame made from the stereo keys of its predecessor: initializeMap's point creation + MapPoint::update +
    removeOutOfFrameMPs (reference src/FeatureTracker.cpp:89-112, src/Map.cpp:61-104, src/FeatureTracker.cpp:910-940)."""
    c = lambda a, dt: np.ascontiguousarray(a, dt)
    kl = np.ascontiguousarray(kl)
    n = len(kl)
    key = np.zeros(max(n, 1), np.int32); pL = np.zeros((max(n, 1), 2), np.float32); pR = np.zeros((max(n, 1), 2), np.float32)
    lL = np.zeros(max(n, 1), np.int32); lR = np.zeros(max(n, 1), np.int32)
    wp = np.zeros((max(n, 1), 4), np.float64); msd = np.zeros(max(n, 1), np.float32)
    imScale = np.float32(fe.params[2])
    logScale = np.float32(np.log(np.float64(imScale)))
    lib().ora_temporal_mappoints.restype = C.c_int
    M = lib().ora_temporal_mappoints(_p(kl), n, _p(c(depth, np.float32)), _p(c(posePrev, np.float64)), _p(c(toCamera, np.float64)),
                                     _p(c(toCameraR, np.float64)), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                     width, height, _p(fe.scalePyramid), fe.nLevels, C.c_float(logScale), _p(key), _p(pL), _p(pR),
                                     _p(lL), _p(lR), _p(wp), _p(msd))
    return dict(key=key[:M].copy(), predL=pL[:M].copy(), predR=pR[:M].copy(), scaleLevelL=lL[:M].copy(), scaleLevelR=lR[:M].copy(),
                wPos=wp[:M].copy(), maxScaleDist=msd[:M].copy(), logScale=logScale)


def stereo_frames(params, fx, baseline, L: np.ndarray, R: np.ndarray, shape: int, threads: int):
    """CPU baseline: run F stereo frames; returns (total matches, per-frame left key counts)."""
    L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8)
    F, h, w = L.shape
    nk = np.zeros(F, np.int32)
    nF, nL, s, e, p, mx, mn = params
    tot = lib().ora_stereo_frames(nF, nL, C.c_float(s), e, p, mx, mn, C.c_double(fx), C.c_float(baseline),
                                  _p(L), _p(R), F, w, h, shape, threads, _p(nk))
    return tot, nk


def track_frames(params, fx, baseline, L: np.ndarray, R: np.ndarray, shape: int, threads: int, motion: dict | None = None, rad: float = 10.0):
    """CPU baseline (restatement): extract + stereo (+ temporal matching against frame motion["prev"][f] when `motion` =
    gtsam_vslam_b200.synth.sequence_motion(...) is given).  Returns (total matches, left key counts, temporal match counts)."""
    L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8)
    F, h, w = L.shape
    nk = np.zeros(F, np.int32); nt = np.zeros(F, np.int32)
    nF, nL, s, e, p, mx, mn = params
    lib().ora_track_frames.restype = C.c_long
    c = lambda a, dt: np.ascontiguousarray(a, dt)
    if motion is None:
        extra = [None, None, None, None, C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0), C.c_float(0), None]
        keep = []
    else:
        keep = [c(motion["prev"], np.int32), c(motion["pose_prev"], np.float64), c(motion["to_camera"], np.float64), c(motion["to_camera_r"], np.float64)]
        extra = [_p(a) for a in keep] + [C.c_double(motion["fx"]), C.c_double(motion["fy"]), C.c_double(motion["cx"]), C.c_double(motion["cy"]),
                                         C.c_float(rad), _p(nt)]
    tot = lib().ora_track_frames(nF, nL, C.c_float(s), e, p, mx, mn, C.c_double(fx), C.c_float(baseline),
                                 _p(L), _p(R), F, w, h, shape, threads, _p(nk), *extra)
    return tot, nk, nt
