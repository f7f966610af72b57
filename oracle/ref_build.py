"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libref_frontend.so.

That library is the reference's OWN src/FeatureExtractor.cpp and src/FeatureMatcher.cpp, compiled unmodified from where
they lie under /root/reference (recipe: oracle/Makefile target `ref`) against the OpenCV / Eigen stand-ins of
oracle/shim.  The five OpenCV image primitives the path calls go through a hook table (oracle/ref_hooks.h):

* `use_cv2_hooks()` routes them to the real OpenCV (python cv2) -> "reference source + real OpenCV", used to pin the
  restatement and to generate the golden vectors of tests/golden/ref_*.npz (tests/golden/make_golden_ref.py);
* the default table uses the restatements of liboracle.so, so the library also runs where cv2 or /root/reference is absent
  (the built .so travels to the GPU box; the sources do not).

Only tests/ and tests/golden/make_golden_ref.py import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .cpu_ref import KEYPOINT_DTYPE, _p, build as build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_frontend.so")
REFERENCE_ROOT = os.environ.get("GVS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    """True if the library exists or can be built here (reference sources present)."""
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def build(force: bool = False) -> str | None:
    """Compile oracle/_ref from the reference sources when they are present; otherwise use the prebuilt file."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        build_oracle()
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []) + ["ref", f"REFERENCE={REFERENCE_ROOT}"],
                              stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


class _Hooks(C.Structure):
    _fields_ = [
        ("resize_linear", C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int)),
        ("fast_nms", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)),
        ("gaussian_blur7", C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_void_p)),
        ("sort_idx_desc", C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)),
        ("fast_atan2", C.CFUNCTYPE(C.c_float, C.c_float, C.c_float)),
    ]


_lib = None
_keepalive = []


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libref_frontend.so is absent and the reference sources are not here to build it")
        C.CDLL(os.path.join(_HERE, "liboracle.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(so)
        _lib.ref_extractor_create.restype = C.c_void_p
        _lib.ref_extractor_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.ref_extractor_destroy.argtypes = [C.c_void_p]
        _lib.gvs_ref_set_hooks.argtypes = [C.c_void_p]
    return _lib


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def use_cv2_hooks() -> None:
    """Route resize / FAST / GaussianBlur / sortIdx / fastAtan2 of the compiled reference to the real OpenCV."""
    import cv2

    def resize_linear(src, sw, sh, dst, dw, dh):
        _view(dst, (dh, dw), np.uint8)[...] = cv2.resize(_view(src, (sh, sw), np.uint8), (dw, dh), interpolation=cv2.INTER_LINEAR)

    det = {}

    def fast_nms(img, w, h, threshold, xs, ys, sc, cap):
        d = det.get(threshold)
        if d is None:
            d = det[threshold] = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=True)
        kps = d.detect(_view(img, (h, w), np.uint8).copy())
        n = len(kps)
        assert n <= cap
        if n:
            _view(xs, (n,), np.int32)[...] = [int(k.pt[0]) for k in kps]
            _view(ys, (n,), np.int32)[...] = [int(k.pt[1]) for k in kps]
            _view(sc, (n,), np.int32)[...] = [int(k.response) for k in kps]
        return n

    def gaussian_blur7(src, w, h, dst):
        _view(dst, (h, w), np.uint8)[...] = cv2.GaussianBlur(_view(src, (h, w), np.uint8), (7, 7), 2, sigmaY=2,
                                                             borderType=cv2.BORDER_REFLECT_101)

    def sort_idx_desc(key, n, idx):
        k = _view(key, (1, n), np.float32)
        _view(idx, (n,), np.int32)[...] = cv2.sortIdx(k, cv2.SORT_DESCENDING | cv2.SORT_EVERY_ROW).reshape(-1)

    def fast_atan2(y, x):
        return float(cv2.fastAtan2(float(y), float(x)))

    h = _Hooks(_Hooks._fields_[0][1](resize_linear), _Hooks._fields_[1][1](fast_nms), _Hooks._fields_[2][1](gaussian_blur7),
               _Hooks._fields_[3][1](sort_idx_desc), _Hooks._fields_[4][1](fast_atan2))
    _keepalive.append(h)
    lib().gvs_ref_set_hooks(C.byref(h))


def use_default_hooks() -> None:
    lib().gvs_ref_set_hooks(None)


class RefExtractor:
    """The reference's FeatureExtractor object (include/FeatureExtractor.h:52-97) behind C entry points."""

    def __init__(self, nFeatures=2000, nLevels=8, imScale=1.2, edgeThreshold=19, patchSize=31,
                 maxFastThreshold=20, minFastThreshold=7):
        self.nFeatures, self.nLevels = int(nFeatures), int(nLevels)
        self.h = C.c_void_p(lib().ref_extractor_create(int(nFeatures), int(nLevels), float(imScale), int(edgeThreshold),
                                                       int(patchSize), int(maxFastThreshold), int(minFastThreshold)))
        L = self.nLevels
        self.scalePyramid = np.zeros(L, np.float32); self.scaleInvPyramid = np.zeros(L, np.float32)
        self.sigmaFactor = np.zeros(L, np.float32); self.InvSigmaFactor = np.zeros(L, np.float32)
        self.scaledPatchSize = np.zeros(L, np.int32); self.featurePerLevel = np.zeros(L, np.int32)
        lib().ref_extractor_tables(self.h, _p(self.scalePyramid), _p(self.scaleInvPyramid), _p(self.sigmaFactor),
                                   _p(self.InvSigmaFactor), _p(self.scaledPatchSize), _p(self.featurePerLevel))

    def __del__(self):
        try:
            lib().ref_extractor_destroy(self.h)
        except Exception:
            pass

    def extractKeysNew(self, image: np.ndarray):
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        cap = max(4 * self.nFeatures, 1024)
        kp = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = lib().ref_extract(self.h, _p(image), w, h, _p(kp), _p(desc), cap)
        assert n >= 0
        return kp[:n].copy(), desc[:n].copy()

    def level_image(self, level: int) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        lib().ref_level_size(self.h, level, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        lib().ref_level_image(self.h, level, _p(out))
        return out

    def ssc(self, numRetPoints: int, cols: int, rows: int, x, y, resp):
        x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32); r = np.ascontiguousarray(resp, np.float32)
        out = np.zeros(max(len(x), 1), np.int32)
        k = lib().ref_ssc(self.h, numRetPoints, cols, rows, len(x), _p(x), _p(y), _p(r), _p(out))
        return out[:k].copy()


class RefMatcher:
    """The reference's FeatureMatcher object (include/FeatureMatcher.h:21-60)."""

    def __init__(self, fx: float, baseline: float, feL: RefExtractor, feR: RefExtractor, imageHeight: int):
        self.feL, self.feR = feL, feR
        lib().ref_matcher_create.restype = C.c_void_p
        lib().ref_matcher_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p(lib().ref_matcher_create(feL.h, feR.h, C.c_double(fx), C.c_float(baseline), int(imageHeight)))

    def __del__(self):
        try:
            lib().ref_matcher_destroy(self.h)
        except Exception:
            pass

    def findStereoMatchesORB2R(self, kl, dl, kr, dr):
        """On the pyramids of the last extractKeysNew of feL / feR."""
        kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
        dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
        NL, NR = len(kl), len(kr)
        ridx = np.zeros(NL, np.int32); lidx = np.zeros(NR, np.int32); dep = np.zeros(NL, np.float32); close = np.zeros(NL, np.uint8)
        lib().ref_stereo_match(self.h, _p(kl), _p(dl), NL, _p(kr), _p(dr), NR, _p(ridx), _p(lidx), _p(dep), _p(close))
        return dict(rightIdxs=ridx, leftIdxs=lidx, estimatedDepth=dep, close=close.astype(bool))

    @staticmethod
    def _grid_args(gl, gr):
        a = [np.ascontiguousarray(gl["start"], np.int32), np.ascontiguousarray(gl["items"], np.int32)]
        a += [None, None] if gr is None else [np.ascontiguousarray(gr["start"], np.int32), np.ascontiguousarray(gr["items"], np.int32)]
        return a, [_p(x) for x in a] + [int(gl["yGrids"]), int(gl["xGrids"]), C.c_float(gl["xMult"]), C.c_float(gl["yMult"])]

    def match_projection(self, mode: int, mps: dict, kl, dl, kr, dr, leftIdxs, rightIdxs, gridL, gridR, rad,
                         matchedL, matchedR, matches):
        """Same contract as oracle.cpu_ref.match_projection, with the key grids (oracle.cpu_ref.assign_grid) passed in."""
        M = len(mps["desc"])
        f = lambda k, dt: np.ascontiguousarray(mps[k], dt) if k in mps and mps[k] is not None else None
        desc = f("desc", np.uint8); pL = f("predL", np.float32); pR = f("predR", np.float32)
        lL = f("scaleLevelL", np.int32); lR = f("scaleLevelR", np.int32); iF = f("inFrame", np.uint8); iR = f("inFrameR", np.uint8)
        kl = np.ascontiguousarray(kl); dl = np.ascontiguousarray(dl)
        NR = 0 if kr is None else len(kr)
        if kr is not None:
            kr = np.ascontiguousarray(kr); dr = np.ascontiguousarray(dr)
        li = None if leftIdxs is None else np.ascontiguousarray(leftIdxs, np.int32)
        ri = None if rightIdxs is None else np.ascontiguousarray(rightIdxs, np.int32)
        keep, gargs = self._grid_args(gridL, gridR)
        return lib().ref_match_projection(self.h, mode, M, _p(desc), _p(pL), _p(pR), _p(lL), _p(lR), _p(iF), _p(iR), _p(kl), _p(dl),
                                          len(kl), _p(kr), _p(dr), NR, _p(li), _p(ri), *gargs, C.c_float(rad),
                                          _p(matchedL), _p(matchedR), _p(matches))

    def match_lba(self, pts: dict, kl, dl, kr, dr, leftIdxs, rightIdxs, unMatchedF, unMatchedFR, gridL, gridR, rad):
        c = lambda a, dt: np.ascontiguousarray(a, dt)
        M = len(pts["desc"])
        out = np.zeros((M, 2), np.int32)
        kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
        args = [c(pts["desc"], np.uint8), c(pts["predL"], np.float32), c(pts["predR"], np.float32), c(pts["predScale"], np.int32),
                c(pts["lastL"], np.float32), c(pts["lastR"], np.float32)]
        more = [c(dl, np.uint8), c(dr, np.uint8), c(leftIdxs, np.int32), c(rightIdxs, np.int32), c(unMatchedF, np.int32),
                c(unMatchedFR, np.int32)]
        keep, gargs = self._grid_args(gridL, gridR)
        n = lib().ref_match_lba(self.h, M, *[_p(a) for a in args], _p(kl), _p(more[0]), len(kl), _p(kr), _p(more[1]), len(kr),
                                _p(more[2]), _p(more[3]), _p(more[4]), _p(more[5]), *gargs, C.c_float(rad), _p(out))
        return n, out


def descriptor_distance(a, b) -> int:
    return lib().ref_descriptor_distance(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))


def stereo_frames(params, fx, baseline, L: np.ndarray, R: np.ndarray, shape: int, threads: int):
    """CPU baseline on the compiled reference: same contract as oracle.cpu_ref.stereo_frames."""
    L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8)
    F, h, w = L.shape
    nk = np.zeros(F, np.int32)
    nF, nL, s, e, p, mx, mn = params
    lib().ref_stereo_frames.restype = C.c_long
    tot = lib().ref_stereo_frames(nF, nL, C.c_float(s), e, p, mx, mn, C.c_double(fx), C.c_float(baseline),
                                  _p(L), _p(R), F, w, h, shape, threads, _p(nk))
    return tot, nk


def track_frames(params, fx, baseline, L: np.ndarray, R: np.ndarray, shape: int, threads: int, motion: dict | None = None, rad: float = 10.0):
    """CPU baseline on the compiled reference: extract + stereo (+ temporal matching against frame motion["prev"][f] when `motion`
    = gtsam_vslam_b200.synth.sequence_motion(...) is given).  Returns (total matches, left key counts, temporal match counts)."""
    L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8)
    F, h, w = L.shape
    nk = np.zeros(F, np.int32); nt = np.zeros(F, np.int32)
    nF, nL, s, e, p, mx, mn = params
    lib().ref_track_frames.restype = C.c_long
    c = lambda a, dt: np.ascontiguousarray(a, dt)
    if motion is None:
        extra = [None, None, None, None, C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0), C.c_float(0), None]
        keep = []
    else:
        keep = [c(motion["prev"], np.int32), c(motion["pose_prev"], np.float64), c(motion["to_camera"], np.float64), c(motion["to_camera_r"], np.float64)]
        extra = [_p(a) for a in keep] + [C.c_double(motion["fx"]), C.c_double(motion["fy"]), C.c_double(motion["cx"]), C.c_double(motion["cy"]),
                                         C.c_float(rad), _p(nt)]
    tot = lib().ref_track_frames(nF, nL, C.c_float(s), e, p, mx, mn, C.c_double(fx), C.c_float(baseline),
                                 _p(L), _p(R), F, w, h, shape, threads, _p(nk), *extra)
    return tot, nk, nt


class RefTracker:
    """The reference's FeatureTracker object (src/FeatureTracker.cpp:23) for the functions of it that oracle/slice_reference.py
    compiles from the reference source: assignKeysToGrids, initializeMap, check2dError, findOutliersR / Mono, worldToFrame,
    removeOutOfFrameMPs / Mono, PredictMPsPosition -- and, through them, MapPoint (src/Map.cpp, compiled whole)."""

    def __init__(self, feL: RefExtractor, feR: RefExtractor, fx, fy, cx, cy, baseline, width, height, extrinsics=None):
        self.feL, self.feR = feL, feR
        ext = np.ascontiguousarray(np.eye(4) if extrinsics is None else extrinsics, np.float64)
        lib().ref_tracker_create.restype = C.c_void_p
        lib().ref_tracker_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p(lib().ref_tracker_create(feL.h, feR.h, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                                     C.c_float(baseline), int(width), int(height), _p(ext)))

    def __del__(self):
        try:
            lib().ref_tracker_destroy(self.h)
        except Exception:
            pass

    def cameras(self, predNPose):
        """(toCamera, toCameraR) = (predNPose^-1, (predNPose * extrinsics)^-1) with this build's arithmetic (:913-914)."""
        a = np.zeros((4, 4)); b = np.zeros((4, 4))
        lib().ref_tracker_cameras(self.h, _p(np.ascontiguousarray(predNPose, np.float64)), _p(a), _p(b))
        return a, b

    def outlier_cameras(self, estimPose):
        """toCameraR = (estimPose^-1 * extrinsics)^-1 (:584-585)."""
        a = np.zeros((4, 4))
        lib().ref_outlier_cameras(self.h, _p(np.ascontiguousarray(estimPose, np.float64)), _p(a))
        return a

    def assign_grid(self, kps, width: int, height: int):
        """FeatureTracker::assignKeysToGrids (:28-54) in CSR form (same dict as oracle.cpu_ref.assign_grid)."""
        kps = np.ascontiguousarray(kps)
        n = len(kps)
        start = np.zeros(64 * 513 + 1, np.int32); items = np.zeros(max(n, 1), np.int32)
        xm, ym = C.c_float(), C.c_float()
        yg = lib().ref_assign_grid(self.h, _p(kps), n, width, height, _p(start), _p(items), C.byref(xm), C.byref(ym))
        return dict(xGrids=64, yGrids=yg, xMult=xm.value, yMult=ym.value, start=start[:yg * 64 + 1].copy(), items=items[:n].copy())

    def predict_mappoints(self, mode, currCamPose, predNPose, valid, wPos, maxScaleDist, logScale, nScaleLev, predL, scaleLevelL, inFrame,
                          predR=None, scaleLevelR=None, inFrameR=None, matches=None, matchedL=None, matchedR=None, outliers=None):
        """removeOutOfFrameMPs (0) / ...Mono (1) / PredictMPsPosition (2) of the reference; in-out arrays in place; keep or None."""
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        M = len(wPos)
        keep = np.zeros(M, np.uint8)
        k = [c(currCamPose, np.float64), c(predNPose, np.float64), c(valid, np.uint8), c(wPos, np.float64), c(maxScaleDist, np.float32),
             c(logScale, np.float32), c(nScaleLev, np.int32)]
        lib().ref_predict_mappoints(self.h, mode, _p(k[0]), _p(k[1]), M, _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]), _p(predL),
                                    _p(scaleLevelL), _p(inFrame), _p(predR), _p(scaleLevelR), _p(inFrameR), _p(keep), _p(matches),
                                    _p(matchedL), 0 if matchedL is None else len(matchedL), _p(matchedR),
                                    0 if matchedR is None else len(matchedR), _p(outliers))
        return keep if mode != 2 else None

    def find_outliers(self, mode, estimPose, thres, valid, wPos, inFrame, inFrameR, matches, outliers, kl, kr, depth, close, rightIdxs,
                      leftIdxs):
        """findOutliersR (0) / findOutliersMono (1) of the reference; returns (nInliers increments, nStereo)."""
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        cnt = np.zeros(2, np.int32)
        keep = [c(estimPose, np.float64), c(valid, np.uint8), c(wPos, np.float64), c(inFrame, np.uint8), c(inFrameR, np.uint8),
                np.ascontiguousarray(kl), None if kr is None else np.ascontiguousarray(kr)]
        lib().ref_find_outliers(self.h, mode, _p(keep[0]), C.c_double(thres), len(matches), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]),
                                _p(matches), _p(outliers), _p(keep[5]), len(kl), _p(keep[6]), 0 if kr is None else len(kr), _p(depth),
                                _p(close), _p(rightIdxs), _p(leftIdxs), _p(cnt))
        return int(cnt[0]), int(cnt[1])

    def temporal_mappoints(self, kl, dl, kr, dr, rightIdxs, depth, posePrev, predNPose):
        """initializeMap (:72-123, MapPoint::update / calcDescriptor inside) on a frame's stereo keys, then removeOutOfFrameMPs."""
        c = lambda a, dt: np.ascontiguousarray(a, dt)
        kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
        n = max(len(kl), 1)
        key = np.zeros(n, np.int32); pL = np.zeros((n, 2), np.float32); pR = np.zeros((n, 2), np.float32)
        lL = np.zeros(n, np.int32); lR = np.zeros(n, np.int32); wp = np.zeros((n, 4)); msd = np.zeros(n, np.float32)
        desc = np.zeros((n, 32), np.uint8)
        keep = [c(dl, np.uint8), c(dr, np.uint8), c(rightIdxs, np.int32), c(depth, np.float32), c(posePrev, np.float64), c(predNPose, np.float64)]
        M = lib().ref_temporal_mappoints(self.h, _p(kl), _p(keep[0]), len(kl), _p(kr), _p(keep[1]), len(kr), _p(keep[2]), _p(keep[3]),
                                         _p(keep[4]), _p(keep[5]), _p(key), _p(pL), _p(pR), _p(lL), _p(lR), _p(wp), _p(msd), _p(desc))
        return dict(key=key[:M].copy(), predL=pL[:M].copy(), predR=pR[:M].copy(), scaleLevelL=lL[:M].copy(), scaleLevelR=lR[:M].copy(),
                    wPos=wp[:M].copy(), maxScaleDist=msd[:M].copy(), desc=desc[:M].copy())


def median_descriptor(descs: np.ndarray):
    """MapPoint::calcDescriptor (src/Map.cpp:145-210) on one observation list: (order in which the reference visits the
    observations, the 32 bytes it keeps)."""
    descs = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32)
    order = np.zeros(max(len(descs), 1), np.int32); chosen = np.zeros(32, np.uint8)
    lib().ref_median_descriptor(_p(descs), len(descs), _p(order), _p(chosen))
    return order[:len(descs)].copy(), chosen
