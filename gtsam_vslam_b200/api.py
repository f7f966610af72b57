"""Host-side mirror of the reference's feature layer on top of the C ABI (include/gvs.h, libgvs.so).

Class and method names follow the reference (include/FeatureExtractor.h:52-97, include/FeatureMatcher.h:22-65,
src/FeatureTracker.cpp:28-70) so that parity tests read like the reference's call sites.  Everything here is
ctypes plumbing: all computation happens in the CUDA kernels behind the C ABI, and every call fails loudly
(GvsError) when the extension or a CUDA device is missing -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgvs.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KEYPOINT_DTYPE.itemsize == 28

GVS_MATCH_STEREO_PRED, GVS_MATCH_MONO, GVS_MATCH_RADIUS = 0, 1, 2

# every symbol include/gvs.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "gvs_last_error", "gvs_version", "gvs_default_params", "gvs_device_count",
    "gvs_extractor_create", "gvs_extractor_destroy", "gvs_extractor_tables", "gvs_tables_from_params", "gvs_extractor_level_size",
    "gvs_extract", "gvs_extract_batch_host", "gvs_extractor_input", "gvs_extract_batch_resident",
    "gvs_extractor_sync", "gvs_extractor_set_rectify_maps", "gvs_extract_batch_host_raw", "gvs_extract_offsets", "gvs_extract_download", "gvs_extractor_set_stream",
    "gvs_debug_level_image", "gvs_debug_level_candidates", "gvs_debug_level_kept", "gvs_debug_select", "gvs_debug_check_guards", "gvs_extractor_tma_stages",
    "gvs_matcher_create", "gvs_matcher_destroy", "gvs_stereo_match_batch", "gvs_stereo_download",
    "gvs_assign_grids_batch", "gvs_grid_download", "gvs_track_stereo_batch_host",
    "gvs_track_stereo_batch_resident", "gvs_matcher_sync", "gvs_matcher_set_stream", "gvs_match_projection", "gvs_match_projection_batch",
    "gvs_track_temporal_batch", "gvs_temporal_download", "gvs_descriptor_distance",
    "gvs_match_projection_lba", "gvs_median_descriptor", "gvs_project_mappoints", "gvs_find_outliers", "gvs_predict_mappoints", "gvs_extractor_enable_timing", "gvs_extractor_stage_times", "gvs_matcher_launch_count",
    "gvs_extractor_launch_count",
    "gvs_pipeline_create", "gvs_pipeline_destroy", "gvs_pipeline_submit", "gvs_pipeline_wait", "gvs_pipeline_download",
    "gvs_pipeline_handles", "gvs_pipeline_release", "gvs_pipeline_submit_temporal", "gvs_pipeline_download_temporal",
    "gvs_multi_create", "gvs_multi_destroy", "gvs_multi_device_count", "gvs_multi_track_stereo",
]


class GvsError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("n_levels", C.c_int32), ("im_scale", C.c_float),
                ("edge_threshold", C.c_int32), ("patch_size", C.c_int32), ("max_fast_threshold", C.c_int32),
                ("min_fast_threshold", C.c_int32)]


class _MapPoints(C.Structure):
    _fields_ = [("count", C.c_int32), ("desc", C.c_void_p), ("pred_l", C.c_void_p), ("pred_r", C.c_void_p),
                ("scale_level_l", C.c_void_p), ("scale_level_r", C.c_void_p), ("in_frame", C.c_void_p),
                ("in_frame_r", C.c_void_p)]


class _TemporalParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("rad", C.c_float), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32), ("prev", C.c_void_p), ("pose_prev", C.c_void_p), ("to_camera", C.c_void_p),
                ("to_camera_r", C.c_void_p)]


class _LbaFrame(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("kp_left", C.c_void_p), ("desc_left", C.c_void_p),
                ("kp_right", C.c_void_p), ("desc_right", C.c_void_p), ("left_idxs", C.c_void_p), ("right_idxs", C.c_void_p),
                ("unmatched_f", C.c_void_p), ("unmatched_fr", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class _LbaPoints(C.Structure):
    _fields_ = [("count", C.c_int32), ("desc", C.c_void_p), ("pred_l", C.c_void_p), ("pred_r", C.c_void_p),
                ("pred_scale", C.c_void_p), ("last_l", C.c_void_p), ("last_r", C.c_void_p)]


class StageTimes(C.Structure):
    _fields_ = [("pyramid_ms", C.c_float), ("fast_ms", C.c_float), ("select_ms", C.c_float), ("blur_ms", C.c_float),
                ("describe_ms", C.c_float), ("stereo_ms", C.c_float), ("grid_ms", C.c_float), ("temporal_ms", C.c_float),
                ("launches", C.c_int32)]


_lib = None


def load_library() -> C.CDLL:
    """Load libgvs.so.  Raises GvsError if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GvsError(f"{LIB_PATH} is missing: build the CUDA extension first (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gvs_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError if the header and the library disagree
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.gvs_extractor_create.argtypes = [C.POINTER(vp), C.POINTER(_Params), i32, i32, i32, i32]
    lib.gvs_extractor_destroy.argtypes = [vp]
    lib.gvs_extractor_tma_stages.argtypes = [vp]
    lib.gvs_extractor_tables.argtypes = [vp] + [vp] * 6
    lib.gvs_tables_from_params.argtypes = [C.POINTER(_Params)] + [vp] * 6
    lib.gvs_extractor_level_size.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.gvs_extract.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, C.POINTER(i32)]
    lib.gvs_extract_batch_host.argtypes = [vp, vp, C.c_size_t, i32, i32]
    lib.gvs_extractor_input.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(C.c_size_t)]
    lib.gvs_extract_batch_resident.argtypes = [vp, i32]
    lib.gvs_extractor_sync.argtypes = [vp]
    lib.gvs_extractor_set_rectify_maps.argtypes = [vp, vp, vp]
    lib.gvs_extract_batch_host_raw.argtypes = [vp, vp, i32, i32]
    lib.gvs_extract_offsets.argtypes = [vp, vp]
    lib.gvs_extract_download.argtypes = [vp, i32, i32, vp, vp, i32]
    lib.gvs_extractor_set_stream.argtypes = [vp, vp]
    lib.gvs_debug_level_image.argtypes = [vp, i32, i32, i32, vp]
    lib.gvs_debug_level_candidates.argtypes = [vp, i32, i32, vp, vp, vp, i32, C.POINTER(i32)]
    lib.gvs_debug_level_kept.argtypes = [vp, i32, i32, vp, vp, vp, i32, C.POINTER(i32)]
    lib.gvs_debug_select.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, C.POINTER(i32)]
    lib.gvs_matcher_create.argtypes = [C.POINTER(vp), vp, vp, C.c_double, f32, i32]
    lib.gvs_matcher_destroy.argtypes = [vp]
    lib.gvs_stereo_match_batch.argtypes = [vp, i32]
    lib.gvs_stereo_download.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.gvs_assign_grids_batch.argtypes = [vp, i32, i32, i32]
    lib.gvs_grid_download.argtypes = [vp, i32, i32, vp, vp, C.POINTER(i32), C.POINTER(f32), C.POINTER(f32)]
    lib.gvs_track_stereo_batch_host.argtypes = [vp, vp, vp, C.c_size_t, i32, i32, i32, i32]
    lib.gvs_track_stereo_batch_resident.argtypes = [vp, i32, i32, i32]
    lib.gvs_matcher_sync.argtypes = [vp]
    lib.gvs_matcher_set_stream.argtypes = [vp, vp]
    lib.gvs_match_projection.argtypes = [vp, i32, i32, C.POINTER(_MapPoints), f32, vp, vp, vp, C.POINTER(i32)]
    lib.gvs_match_projection_batch.argtypes = [vp, i32, i32, i32, C.POINTER(_MapPoints), vp, f32, vp, vp, vp, vp]
    lib.gvs_track_temporal_batch.argtypes = [vp, i32, C.POINTER(_TemporalParams)]
    lib.gvs_temporal_download.argtypes = [vp, i32, i32] + [vp] * 8
    lib.gvs_descriptor_distance.argtypes = [vp, vp, vp, i32, vp]
    lib.gvs_match_projection_lba.argtypes = [vp, C.POINTER(_LbaFrame), C.POINTER(_LbaPoints), f32, vp, C.POINTER(i32)]
    lib.gvs_median_descriptor.argtypes = [vp, vp, vp, i32, vp]
    dbl = C.c_double
    lib.gvs_project_mappoints.argtypes = [vp, vp, dbl, dbl, dbl, dbl, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.gvs_predict_mappoints.argtypes = [vp, i32, vp, vp, dbl, dbl, dbl, dbl, i32, i32, i32] + [vp] * 14 + [i32, vp, i32, vp]
    lib.gvs_find_outliers.argtypes = [vp, i32, i32, vp, vp, dbl, dbl, dbl, dbl, dbl, f32, i32] + [vp] * 10 + [C.POINTER(i32), C.POINTER(i32)]
    lib.gvs_extractor_enable_timing.argtypes = [vp, i32]
    lib.gvs_extractor_stage_times.argtypes = [vp, C.POINTER(StageTimes)]
    lib.gvs_matcher_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.gvs_extractor_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    i64 = C.c_int64
    lib.gvs_pipeline_create.argtypes = [C.POINTER(vp), C.POINTER(_Params), i32, i32, i32, i32, C.c_double, f32, i32]
    lib.gvs_pipeline_destroy.argtypes = [vp]
    lib.gvs_pipeline_submit.argtypes = [vp, vp, vp, C.c_size_t, i32, i32, C.POINTER(i64)]
    lib.gvs_pipeline_submit_temporal.argtypes = [vp, vp, vp, C.c_size_t, i32, i32, C.POINTER(_TemporalParams), C.POINTER(i64)]
    lib.gvs_pipeline_download_temporal.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.gvs_pipeline_wait.argtypes = [vp, i64, vp, vp]
    lib.gvs_pipeline_download.argtypes = [vp, i64, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp]
    lib.gvs_pipeline_handles.argtypes = [vp, i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.gvs_pipeline_release.argtypes = [vp, i64]
    lib.gvs_multi_create.argtypes = [C.POINTER(vp), C.POINTER(_Params), i32, i32, i32, C.c_double, f32, vp, i32]
    lib.gvs_multi_destroy.argtypes = [vp]
    lib.gvs_multi_device_count.argtypes = [vp]
    lib.gvs_multi_track_stereo.argtypes = [vp, vp, vp, C.c_size_t, i32, i32, C.POINTER(_TemporalParams), vp, vp, vp, vp, i32, vp, vp, i32,
                                           vp, vp, vp, vp, vp, vp, vp, i32, vp]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise GvsError(f"gvs error {rc}: {load_library().gvs_last_error().decode()}")


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def check_guards():
    """(overwritten guard tails, buffers checked) -- GVS_GUARD=1 debugging aid, see include/gvs.h."""
    n = C.c_int(0)
    bad = load_library().gvs_debug_check_guards(C.byref(n))
    return bad, n.value


def device_count() -> int:
    return load_library().gvs_device_count()


@dataclass
class TrackedKeys:
    """include/FeatureExtractor.h:18-50.  Grids are CSR (cell_start, items) instead of vector<vector<vector<int>>>."""
    keyPoints: np.ndarray = None
    rightKeyPoints: np.ndarray = None
    Desc: np.ndarray = None
    rightDesc: np.ndarray = None
    lkeyGrid: dict = field(default_factory=dict)
    rkeyGrid: dict = field(default_factory=dict)
    rightIdxs: np.ndarray = None
    leftIdxs: np.ndarray = None
    estimatedDepth: np.ndarray = None
    close: np.ndarray = None
    xMult: float = 0.0
    yMult: float = 0.0
    xGrids: int = 0
    yGrids: int = 0


class FeatureExtractor:
    """include/FeatureExtractor.h:52-97 with `max_images` image slots (batch of independent images)."""

    def __init__(self, nFeatures=2000, nLevels=8, imScale=1.2, edgeThreshold=19, patchSize=31, maxFastThreshold=20,
                 minFastThreshold=7, *, width: int, height: int, max_images: int = 1, device: int = 0):
        lib = load_library()
        self._lib = lib
        self.nFeatures, self.nLevels, self.imScale = int(nFeatures), int(nLevels), float(imScale)
        self.edgeThreshold, self.patchSize = int(edgeThreshold), int(patchSize)
        self.maxFastThreshold, self.minFastThreshold = int(maxFastThreshold), int(minFastThreshold)
        self.width, self.height, self.max_images, self.device = int(width), int(height), int(max_images), int(device)
        p = _Params(self.nFeatures, self.nLevels, self.imScale, self.edgeThreshold, self.patchSize,
                    self.maxFastThreshold, self.minFastThreshold)
        self.h = C.c_void_p()
        _check(lib.gvs_extractor_create(C.byref(self.h), C.byref(p), self.width, self.height, self.max_images, self.device))
        L = self.nLevels
        self.scalePyramid = np.zeros(L, np.float32)
        self.scaleInvPyramid = np.zeros(L, np.float32)
        self.sigmaFactor = np.zeros(L, np.float32)
        self.InvSigmaFactor = np.zeros(L, np.float32)
        self.scaledPatchSize = np.zeros(L, np.int32)
        self.featurePerLevel = np.zeros(L, np.int32)
        _check(lib.gvs_extractor_tables(self.h, _p(self.scalePyramid), _p(self.scaleInvPyramid), _p(self.sigmaFactor),
                                        _p(self.InvSigmaFactor), _p(self.scaledPatchSize), _p(self.featurePerLevel)))
        self.cap_per_image = int(2 * self.featurePerLevel.sum() + 64 * L)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.gvs_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- FeatureExtractor::extractKeysNew (src/FeatureExtractor.cpp:481-533)
    def extractKeysNew(self, image: np.ndarray):
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        cap = self.cap_per_image
        kp = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(self._lib.gvs_extract(self.h, _p(image), w, h, w, _p(kp), _p(desc), cap, C.byref(n)))
        return kp[:n.value].copy(), desc[:n.value].copy()

    # -- batched
    def extract_batch(self, images: np.ndarray) -> None:
        """images: (n, H, W) uint8 host array (pinned or pageable).  Asynchronous."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.flags.c_contiguous
        n, h, w = images.shape
        assert (h, w) == (self.height, self.width)
        self._keep = images
        _check(self._lib.gvs_extract_batch_host(self.h, _p(images), h * w, w, n))

    def set_rectify_maps(self, mapx: np.ndarray | None, mapy: np.ndarray | None) -> None:
        """cv::initUndistortRectifyMap outputs (CV_32FC1), or None/None for already rectified cameras."""
        if mapx is None:
            _check(self._lib.gvs_extractor_set_rectify_maps(self.h, None, None))
            return
        mx = np.ascontiguousarray(mapx, np.float32); my = np.ascontiguousarray(mapy, np.float32)
        assert mx.shape == (self.height, self.width) and my.shape == mx.shape
        _check(self._lib.gvs_extractor_set_rectify_maps(self.h, _p(mx), _p(my)))

    def extract_batch_raw(self, images: np.ndarray) -> None:
        """Raw camera frames (n, H, W, 3) BGR or (n, H, W) gray: remap + BGR2GRAY on the device, then extraction."""
        assert images.dtype == np.uint8 and images.flags.c_contiguous and images.shape[1:3] == (self.height, self.width)
        ch = 1 if images.ndim == 3 else images.shape[3]
        self._keep = images
        _check(self._lib.gvs_extract_batch_host_raw(self.h, _p(images), ch, images.shape[0]))

    def input_layout(self):
        dev, pitch, stride = C.c_void_p(), C.c_int(), C.c_size_t()
        _check(self._lib.gvs_extractor_input(self.h, C.byref(dev), C.byref(pitch), C.byref(stride)))
        return dev.value, pitch.value, stride.value

    def extract_batch_resident(self, n: int) -> None:
        _check(self._lib.gvs_extract_batch_resident(self.h, n))

    def sync(self) -> None:
        _check(self._lib.gvs_extractor_sync(self.h))

    def tma_stages(self) -> int:
        """bit 0: the fused pyramid kernel, bit 1: the FAST kernel stage their tiles with TMA box loads (include/gvs.h)."""
        return int(self._lib.gvs_extractor_tma_stages(self.h))

    def offsets(self, n: int) -> np.ndarray:
        """Dense key offsets of the first n images of the last batch (n + 1 entries)."""
        off = np.zeros(self.max_images + 1, np.int32)   # the C call writes one entry per image of the last batch, plus one
        _check(self._lib.gvs_extract_offsets(self.h, _p(off)))
        return off[:n + 1].copy()

    def download(self, first: int, count: int, total: int):
        kp = np.zeros(max(total, 1), KEYPOINT_DTYPE)
        desc = np.zeros((max(total, 1), 32), np.uint8)
        _check(self._lib.gvs_extract_download(self.h, first, count, _p(kp), _p(desc), total))
        return kp[:total], desc[:total]

    def download_into(self, first: int, count: int, kp: np.ndarray, desc: np.ndarray) -> int:
        """Like download() but into caller arrays (use pinned memory for full PCIe speed); returns the key count."""
        off = self.offsets(first + count)
        total = int(off[first + count] - off[first])
        assert kp.dtype == KEYPOINT_DTYPE and len(kp) >= total and desc.dtype == np.uint8 and desc.size >= total * 32
        _check(self._lib.gvs_extract_download(self.h, first, count, _p(kp), _p(desc), len(kp)))
        return total

    def set_stream(self, cuda_stream: int | None) -> None:
        _check(self._lib.gvs_extractor_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def level_size(self, level: int):
        w, h = C.c_int(), C.c_int()
        _check(self._lib.gvs_extractor_level_size(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # -- parity taps
    def level_image(self, image: int, level: int, blurred: bool = False) -> np.ndarray:
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        _check(self._lib.gvs_debug_level_image(self.h, image, level, int(blurred), _p(out)))
        return out

    def _level_keys(self, fn, image: int, level: int):
        cap = 1 << 20
        x = np.zeros(cap, np.int32); y = np.zeros(cap, np.int32); s = np.zeros(cap, np.int32)
        n = C.c_int(0)
        _check(fn(self.h, image, level, _p(x), _p(y), _p(s), cap, C.byref(n)))
        return x[:n.value].copy(), y[:n.value].copy(), s[:n.value].copy()

    def level_candidates(self, image: int, level: int):
        return self._level_keys(self._lib.gvs_debug_level_candidates, image, level)

    def level_kept(self, image: int, level: int):
        return self._level_keys(self._lib.gvs_debug_level_kept, image, level)

    def debug_select(self, level: int, x, y, score):
        """Selection stage alone on a caller-made candidate list: (sorted permutation or None, kept input indices)."""
        x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32); s = np.ascontiguousarray(score, np.int32)
        n = len(x)
        sidx = np.zeros(max(n, 1), np.int32); kept = np.zeros(max(n, 1), np.int32)
        nk = C.c_int(0)
        _check(self._lib.gvs_debug_select(self.h, level, n, _p(x), _p(y), _p(s), _p(sidx), _p(kept), max(n, 1), C.byref(nk)))
        return (sidx[:n] if n and sidx[0] >= 0 else None), kept[:nk.value].copy()

    def enable_timing(self, on: bool = True) -> None:
        _check(self._lib.gvs_extractor_enable_timing(self.h, int(on)))

    def stage_times(self) -> StageTimes:
        t = StageTimes()
        _check(self._lib.gvs_extractor_stage_times(self.h, C.byref(t)))
        return t

    def launch_count(self) -> int:
        v = C.c_int64(0)
        _check(self._lib.gvs_extractor_launch_count(self.h, C.byref(v)))
        return v.value


class FeatureMatcher:
    """include/FeatureMatcher.h:22-65.  zedptr is reduced to the three values the path reads: fx, baseline, height."""

    def __init__(self, fx: float, baseline: float, feLeft: FeatureExtractor, feRight: FeatureExtractor | None,
                 imageHeight: int = 360):
        self._lib = load_library()
        self.feLeft, self.feRight = feLeft, feRight if feRight is not None else feLeft
        self.fx, self.baseline, self.imageHeight = float(fx), float(baseline), int(imageHeight)
        self.closeNumber = 40
        self.h = C.c_void_p()
        _check(self._lib.gvs_matcher_create(C.byref(self.h), feLeft.h, self.feRight.h, self.fx, self.baseline, self.imageHeight))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.gvs_matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _dense(fe: FeatureExtractor, first: int, count: int):
        off = fe.offsets(first + count)
        return off, int(off[first]), int(off[first + count] - off[first])

    # -- findStereoMatchesORB2R (src/FeatureMatcher.cpp:528-708) over the last extracted batch
    def stereo_match_batch(self, n_frames: int) -> None:
        _check(self._lib.gvs_stereo_match_batch(self.h, n_frames))

    def stereo_download(self, first: int, count: int):
        _, _, nl = self._dense(self.feLeft, first, count)
        _, _, nr = self._dense(self.feRight, first, count)
        ridx = np.zeros(max(nl, 1), np.int32); dep = np.zeros(max(nl, 1), np.float32); cl = np.zeros(max(nl, 1), np.uint8)
        lidx = np.zeros(max(nr, 1), np.int32)
        _check(self._lib.gvs_stereo_download(self.h, first, count, _p(ridx), _p(lidx), _p(dep), _p(cl)))
        return ridx[:nl], lidx[:nr], dep[:nl], cl[:nl].astype(bool)

    def stereo_download_into(self, first: int, count: int, ridx, lidx, dep, cl) -> None:
        _check(self._lib.gvs_stereo_download(self.h, first, count, _p(ridx), _p(lidx), _p(dep), _p(cl)))

    def assign_grids_batch(self, n_frames: int, width: int, height: int) -> None:
        _check(self._lib.gvs_assign_grids_batch(self.h, n_frames, width, height))

    def grid_download(self, frame: int, side: int):
        fe = self.feRight if side else self.feLeft
        off = fe.offsets(frame + 1)
        n = int(off[frame + 1] - off[frame])
        yg, xm, ym = C.c_int(), C.c_float(), C.c_float()
        _check(self._lib.gvs_grid_download(self.h, frame, side, None, None, C.byref(yg), C.byref(xm), C.byref(ym)))  # sizes first
        start = np.zeros(64 * yg.value + 1, np.int32); items = np.zeros(max(n, 1), np.int32)
        _check(self._lib.gvs_grid_download(self.h, frame, side, _p(start), _p(items), C.byref(yg), C.byref(xm), C.byref(ym)))
        return dict(xGrids=64, yGrids=yg.value, xMult=xm.value, yMult=ym.value, start=start[:64 * yg.value + 1].copy(),
                    items=items[:n].copy())

    def track_stereo_batch(self, left: np.ndarray, right: np.ndarray) -> None:
        """FeatureTracker::extractORBAndStereoMatch (src/FeatureTracker.cpp:56-70) for a batch of host frames."""
        assert left.shape == right.shape and left.dtype == np.uint8 and left.flags.c_contiguous and right.flags.c_contiguous
        n, h, w = left.shape
        self._keep = (left, right)
        _check(self._lib.gvs_track_stereo_batch_host(self.h, _p(left), _p(right), h * w, w, n, w, h))

    def track_stereo_batch_resident(self, n_frames: int, width: int, height: int) -> None:
        _check(self._lib.gvs_track_stereo_batch_resident(self.h, n_frames, width, height))

    def sync(self) -> None:
        _check(self._lib.gvs_matcher_sync(self.h))

    def set_stream(self, cuda_stream: int | None) -> None:
        _check(self._lib.gvs_matcher_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def launch_count(self) -> int:
        v = C.c_int64(0)
        _check(self._lib.gvs_matcher_launch_count(self.h, C.byref(v)))
        return v.value

    # -- matchByProjectionRPred / matchByProjectionMono / matchByRadius (src/FeatureMatcher.cpp:254-526)
    def _project(self, frame, mode, mps, rad, matchedL, matchedR, matches):
        M = len(mps["desc"])
        keep = []

        def arr(key, dt):
            v = mps.get(key)
            if v is None:
                return None
            a = np.ascontiguousarray(v, dt)
            keep.append(a)
            return a.ctypes.data

        s = _MapPoints(M, arr("desc", np.uint8), arr("predL", np.float32), arr("predR", np.float32),
                       arr("scaleLevelL", np.int32), arr("scaleLevelR", np.int32), arr("inFrame", np.uint8),
                       arr("inFrameR", np.uint8))
        assert matchedL.dtype == np.int32 and matches.dtype == np.int32 and matches.shape == (M, 2) and matches.flags.c_contiguous
        n = C.c_int(0)
        _check(self._lib.gvs_match_projection(self.h, frame, mode, C.byref(s), float(rad), _p(matchedL),
                                              _p(matchedR) if matchedR is not None else None, _p(matches), C.byref(n)))
        return n.value

    def matchByProjectionRPred(self, mps, frame, matchedIdxsL, matchedIdxsR, matchesIdxs, rad):
        return self._project(frame, GVS_MATCH_STEREO_PRED, mps, rad, matchedIdxsL, matchedIdxsR, matchesIdxs)

    def matchByProjectionMono(self, mps, frame, matchedIdxsL, matchesIdxs, rad):
        return self._project(frame, GVS_MATCH_MONO, mps, rad, matchedIdxsL, None, matchesIdxs)

    def matchByRadius(self, mps, frame, matchedIdxsL, matchesIdxs, rad):
        return self._project(frame, GVS_MATCH_RADIUS, mps, rad, matchedIdxsL, None, matchesIdxs)

    def match_projection_batch(self, mode, first_frame, n_frames, mps, mp_offsets, rad, matchedL, matchedR, matches):
        """matchByProjectionRPred / Mono / matchByRadius for frames [first_frame, first_frame + n_frames) in one call: frame f is
        matched against map points [mp_offsets[f], mp_offsets[f+1]) of `mps`; matchedL / matchedR are dense over the frames' keys,
        matches is (M, 2); all three in-out.  Returns the per-frame match counts."""
        M = len(mps["desc"])
        keep = []

        def arr(key, dt):
            v = mps.get(key)
            if v is None:
                return None
            x = np.ascontiguousarray(v, dt)
            keep.append(x)
            return x.ctypes.data

        s = _MapPoints(M, arr("desc", np.uint8), arr("predL", np.float32), arr("predR", np.float32),
                       arr("scaleLevelL", np.int32), arr("scaleLevelR", np.int32), arr("inFrame", np.uint8),
                       arr("inFrameR", np.uint8))
        off = np.ascontiguousarray(mp_offsets, np.int32)
        assert len(off) == n_frames + 1
        assert matchedL.dtype == np.int32 and matches.dtype == np.int32 and matches.shape == (M, 2) and matches.flags.c_contiguous
        n = np.zeros(n_frames, np.int32)
        _check(self._lib.gvs_match_projection_batch(self.h, first_frame, n_frames, mode, C.byref(s), _p(off), float(rad), _p(matchedL),
                                                    _p(matchedR) if matchedR is not None else None, _p(matches), _p(n)))
        return n

    def track_temporal_batch(self, n_frames, prev, posePrev, toCamera, toCameraR, fx, fy, cx, cy, width, height, rad, mode=0) -> None:
        """Device-resident temporal tracking of the last stereo batch (include/gvs.h: gvs_track_temporal_batch): map points of
        frame f from the stereo keys of frame prev[f] (initializeMap), removeOutOfFrameMPs, matchByProjectionRPred.  Asynchronous."""
        c = lambda x, dt: np.ascontiguousarray(x, dt)
        k = [c(prev, np.int32), c(posePrev, np.float64), c(toCamera, np.float64), c(toCameraR, np.float64)]
        assert len(k[0]) == n_frames and k[1].size == 16 * n_frames and k[2].size == 16 * n_frames and k[3].size == 16 * n_frames
        self._tkeep = k
        p = _TemporalParams(int(mode), float(rad), float(fx), float(fy), float(cx), float(cy), int(width), int(height),
                            k[0].ctypes.data, k[1].ctypes.data, k[2].ctypes.data, k[3].ctypes.data)
        _check(self._lib.gvs_track_temporal_batch(self.h, n_frames, C.byref(p)))

    def temporal_download(self, first: int, count: int, preds: bool = False):
        """Results of track_temporal_batch for frames [first, first + count)."""
        off = np.zeros(count + 1, np.int32); n = np.zeros(count, np.int32)
        _check(self._lib.gvs_temporal_download(self.h, first, count, _p(off), None, None, _p(n), None, None, None, None))
        M = int(off[-1])
        _, _, nl = self._dense(self.feLeft, first, count)
        _, _, nr = self._dense(self.feRight, first, count)
        key = np.zeros(max(M, 1), np.int32); mt = np.zeros((max(M, 1), 2), np.int32)
        mL = np.zeros(max(nl, 1), np.int32); mR = np.zeros(max(nr, 1), np.int32)
        pL = np.zeros((max(M, 1), 2), np.float32) if preds else None
        pR = np.zeros((max(M, 1), 2), np.float32) if preds else None
        _check(self._lib.gvs_temporal_download(self.h, first, count, None, _p(key), _p(mt), None, _p(mL), _p(mR), _p(pL), _p(pR)))
        out = dict(mp_offsets=off, mp_key=key[:M], matches=mt[:M], n_matches=n, matchedL=mL[:nl], matchedR=mR[:nr])
        if preds:
            out["predL"], out["predR"] = pL[:M], pR[:M]
        return out

    def temporal_download_into(self, first: int, count: int, off, key, mt, n) -> int:
        """Throughput form: offsets / source keys / matches / counts into caller (pinned) arrays; returns the map point count."""
        _check(self._lib.gvs_temporal_download(self.h, first, count, _p(off), _p(key), _p(mt), _p(n), None, None, None, None))
        return int(off[count])

    # -- matchByProjectionRPredLBA (src/FeatureMatcher.cpp:66-252); flattened inputs, see include/gvs.h
    def matchByProjectionRPredLBA(self, pts: dict, kl, dl, kr, dr, leftIdxs, rightIdxs, unMatchedF, unMatchedFR, width, height, rad):
        c = lambda a, dt: np.ascontiguousarray(a, dt)
        kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
        keep = [kl, c(dl, np.uint8), kr, c(dr, np.uint8), c(leftIdxs, np.int32), c(rightIdxs, np.int32), c(unMatchedF, np.int32),
                c(unMatchedFR, np.int32)]
        f = _LbaFrame(len(kl), len(kr), *[a.ctypes.data for a in keep], width, height)
        M = len(pts["desc"])
        pk = [c(pts["desc"], np.uint8), c(pts["predL"], np.float32), c(pts["predR"], np.float32), c(pts["predScale"], np.int32),
              c(pts["lastL"], np.float32), c(pts["lastR"], np.float32)]
        p = _LbaPoints(M, *[a.ctypes.data for a in pk])
        out = np.zeros((M, 2), np.int32)
        n = C.c_int(0)
        _check(self._lib.gvs_match_projection_lba(self.h, C.byref(f), C.byref(p), float(rad), _p(out), C.byref(n)))
        return n.value, out

    # -- MapPoint::calcDescriptor (src/Map.cpp:145-210), batched over map points
    def median_descriptor(self, descs: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        descs = np.ascontiguousarray(descs, np.uint8); offsets = np.ascontiguousarray(offsets, np.int32)
        out = np.zeros(len(offsets) - 1, np.int32)
        _check(self._lib.gvs_median_descriptor(self.h, _p(descs), _p(offsets), len(out), _p(out)))
        return out

    # -- FeatureTracker::worldToFrame + MapPoint::predictScale (src/FeatureTracker.cpp:685-741, src/Map.cpp:13-23), batched
    def worldToFrame(self, poseInv, fx, fy, cx, cy, width, height, wPos, maxScaleDist, logScale, nScaleLev, pred, scaleLevel):
        c = lambda a, dt: np.ascontiguousarray(a, dt)
        M = len(wPos)
        inF = np.zeros(max(M, 1), np.uint8)
        assert pred.dtype == np.float32 and scaleLevel.dtype == np.int32 and pred.flags.c_contiguous
        keep = [c(poseInv, np.float64), c(wPos, np.float64), c(maxScaleDist, np.float32), c(logScale, np.float32), c(nScaleLev, np.int32)]
        _check(self._lib.gvs_project_mappoints(self.h, _p(keep[0]), fx, fy, cx, cy, width, height, M, _p(keep[1]), _p(keep[2]), _p(keep[3]),
                                               _p(keep[4]), _p(pred), _p(scaleLevel), _p(inF)))
        return inF[:M]

    # -- FeatureTracker::removeOutOfFrameMPs / removeOutOfFrameMPsMono / PredictMPsPosition (src/FeatureTracker.cpp:910-1013), batched
    def predictMapPoints(self, mode, toCamera, toCameraR, fx, fy, cx, cy, width, height, valid, wPos, maxScaleDist, logScale, nScaleLev,
                         predL, scaleLevelL, inFrame, predR=None, scaleLevelR=None, inFrameR=None, matches=None, matchedL=None,
                         matchedR=None, outliers=None):
        """mode 0 removeOutOfFrameMPs, 1 removeOutOfFrameMPsMono (returns keep), 2 PredictMPsPosition; in-out arrays updated in place."""
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        M = len(wPos)
        keep = np.zeros(max(M, 1), np.uint8)
        for a, dt in ((predL, np.float32), (scaleLevelL, np.int32), (inFrame, np.uint8), (predR, np.float32), (scaleLevelR, np.int32),
                      (inFrameR, np.uint8), (matches, np.int32), (matchedL, np.int32), (matchedR, np.int32), (outliers, np.uint8)):
            assert a is None or (a.dtype == dt and a.flags.c_contiguous)
        k = [c(toCamera, np.float64), c(toCameraR, np.float64), c(valid, np.uint8), c(wPos, np.float64), c(maxScaleDist, np.float32),
             c(logScale, np.float32), c(nScaleLev, np.int32)]
        _check(self._lib.gvs_predict_mappoints(self.h, mode, _p(k[0]), _p(k[1]), fx, fy, cx, cy, width, height, M, _p(k[2]), _p(k[3]), _p(k[4]),
                                               _p(k[5]), _p(k[6]), _p(predL), _p(scaleLevelL), _p(inFrame), _p(predR), _p(scaleLevelR),
                                               _p(inFrameR), _p(keep), _p(matches), _p(matchedL), 0 if matchedL is None else len(matchedL),
                                               _p(matchedR), 0 if matchedR is None else len(matchedR), _p(outliers)))
        return keep[:M] if mode != 2 else None

    # -- FeatureTracker::findOutliersR / findOutliersMono (src/FeatureTracker.cpp:582-684) against frame `frame` of the last batch
    def _find_outliers(self, mode, frame, estimPose, toCameraR, fx, fy, cx, cy, thres, closeDist, valid, wPos, inFrame, inFrameR,
                       matches, outliers, keys=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        assert matches.dtype == np.int32 and matches.flags.c_contiguous and outliers.dtype == np.uint8
        keep = [c(estimPose, np.float64), c(toCameraR, np.float64), c(valid, np.uint8), c(wPos, np.float64), c(inFrame, np.uint8),
                c(inFrameR, np.uint8)]
        st = [None] * 4
        if keys is not None:  # the caller's TrackedKeys copies of the stereo arrays follow the device state
            st = [keys.estimatedDepth, keys.close, keys.rightIdxs, keys.leftIdxs]
            assert st[0].dtype == np.float32 and st[1].dtype in (np.uint8, np.bool_) and st[2].dtype == np.int32 and st[3].dtype == np.int32
        nIn, nSt = C.c_int(0), C.c_int(0)
        _check(self._lib.gvs_find_outliers(self.h, frame, mode, _p(keep[0]), _p(keep[1]), fx, fy, cx, cy, thres, closeDist, len(matches),
                                           _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]), _p(matches), _p(outliers), _p(st[0]),
                                           _p(st[1]), _p(st[2]), _p(st[3]), C.byref(nIn), C.byref(nSt)))
        return nIn.value, nSt.value

    def findOutliersR(self, frame, estimPose, toCameraR, fx, fy, cx, cy, thres, closeDist, valid, wPos, inFrame, inFrameR, matches,
                      outliers, keys=None):
        """Returns (number of nInliers increments, nStereo); matches / outliers (and keys' stereo arrays) are updated in place."""
        return self._find_outliers(0, frame, estimPose, toCameraR, fx, fy, cx, cy, thres, closeDist, valid, wPos, inFrame, inFrameR,
                                   matches, outliers, keys)

    def findOutliersMono(self, frame, estimPose, fx, fy, cx, cy, thres, valid, wPos, inFrame, matches, outliers):
        return self._find_outliers(1, frame, estimPose, None, fx, fy, cx, cy, thres, 0.0, valid, wPos, inFrame, None, matches,
                                   outliers)[0]

    def DescriptorDistance(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        _check(self._lib.gvs_descriptor_distance(self.h, _p(a), _p(b), len(a), _p(out)))
        return out


class StereoPipeline:
    """gvs_pipeline (include/gvs.h): `n_slots` handle sets in a ring; `submit` is asynchronous, `wait` /
    `download` collect a ticket, `release` frees its slot.  Ticket t uses slot t % n_slots."""

    def __init__(self, fx: float, baseline: float, *, width: int, height: int, frames_per_slot: int = 1, n_slots: int = 2,
                 nFeatures=2000, nLevels=8, imScale=1.2, edgeThreshold=19, patchSize=31, maxFastThreshold=20,
                 minFastThreshold=7, device: int = 0):
        self._lib = load_library()
        self.width, self.height, self.frames_per_slot, self.n_slots = int(width), int(height), int(frames_per_slot), int(n_slots)
        p = _Params(int(nFeatures), int(nLevels), float(imScale), int(edgeThreshold), int(patchSize), int(maxFastThreshold),
                    int(minFastThreshold))
        self.h = C.c_void_p()
        _check(self._lib.gvs_pipeline_create(C.byref(self.h), C.byref(p), self.width, self.height, self.frames_per_slot,
                                             self.n_slots, float(fx), float(baseline), int(device)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.gvs_pipeline_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, left: np.ndarray, right: np.ndarray) -> int:
        """left/right: (n, H, W) uint8, C-contiguous, ideally pinned; must stay alive until wait() returns."""
        assert left.shape == right.shape and left.dtype == np.uint8 and left.flags.c_contiguous and right.flags.c_contiguous
        n, h, w = left.shape
        assert (h, w) == (self.height, self.width)
        t = C.c_int64(-1)
        _check(self._lib.gvs_pipeline_submit(self.h, _p(left), _p(right), w * h, w, n, C.byref(t)))
        return t.value

    def submit_temporal(self, left: np.ndarray, right: np.ndarray, motion: dict, rad: float) -> int:
        """submit() followed by the temporal tracking of the ticket's frames; motion = synth.sequence_motion-like dict with prev,
        pose_prev, to_camera, to_camera_r (per frame of the ticket) and fx, fy, cx, cy."""
        n, h, w = left.shape
        c = lambda x, dt: np.ascontiguousarray(x, dt)
        k = [c(motion["prev"], np.int32), c(motion["pose_prev"], np.float64), c(motion["to_camera"], np.float64), c(motion["to_camera_r"], np.float64)]
        assert len(k[0]) == n
        p = _TemporalParams(0, float(rad), float(motion["fx"]), float(motion["fy"]), float(motion["cx"]), float(motion["cy"]), w, h,
                            k[0].ctypes.data, k[1].ctypes.data, k[2].ctypes.data, k[3].ctypes.data)
        t = C.c_int64(-1)
        _check(self._lib.gvs_pipeline_submit_temporal(self.h, _p(left), _p(right), w * h, w, n, C.byref(p), C.byref(t)))
        return t.value

    def download_temporal_into(self, ticket: int, off, key, mt, n) -> None:
        _check(self._lib.gvs_pipeline_download_temporal(self.h, ticket, _p(off), _p(key), _p(mt), _p(n)))

    def wait(self, ticket: int, n_frames: int):
        offL = np.zeros(n_frames + 1, np.int32); offR = np.zeros(n_frames + 1, np.int32)
        _check(self._lib.gvs_pipeline_wait(self.h, ticket, _p(offL), _p(offR)))
        return offL, offR

    def download_into(self, ticket: int, kl, dl, kr, dr, ridx, lidx, dep, cl) -> None:
        """Dense results of the ticket into caller arrays (pinned for speed); capacities are taken from the arrays."""
        _check(self._lib.gvs_pipeline_download(self.h, ticket, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                               _p(ridx), _p(lidx), _p(dep), _p(cl)))

    def download(self, ticket: int, n_frames: int):
        offL, offR = self.wait(ticket, n_frames)
        nl, nr = int(offL[-1]), int(offR[-1])
        kl = np.zeros(max(nl, 1), KEYPOINT_DTYPE); dl = np.zeros((max(nl, 1), 32), np.uint8)
        kr = np.zeros(max(nr, 1), KEYPOINT_DTYPE); dr = np.zeros((max(nr, 1), 32), np.uint8)
        ridx = np.zeros(max(nl, 1), np.int32); lidx = np.zeros(max(nr, 1), np.int32)
        dep = np.zeros(max(nl, 1), np.float32); cl = np.zeros(max(nl, 1), np.uint8)
        self.download_into(ticket, kl, dl, kr, dr, ridx, lidx, dep, cl)
        return dict(offL=offL, offR=offR, kpL=kl[:nl], descL=dl[:nl], kpR=kr[:nr], descR=dr[:nr], rightIdxs=ridx[:nl],
                    leftIdxs=lidx[:nr], estimatedDepth=dep[:nl], close=cl[:nl].astype(bool))

    def release(self, ticket: int) -> None:
        _check(self._lib.gvs_pipeline_release(self.h, ticket))

    def slot_launch_count(self, ticket: int) -> int:
        fl, fr, fm = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self._lib.gvs_pipeline_handles(self.h, ticket, C.byref(fl), C.byref(fr), C.byref(fm)))
        tot = 0
        for fn, hnd in ((self._lib.gvs_extractor_launch_count, fl), (self._lib.gvs_extractor_launch_count, fr),
                        (self._lib.gvs_matcher_launch_count, fm)):
            v = C.c_int64(0)
            _check(fn(hnd, C.byref(v)))
            tot += v.value
        return tot


class MultiGpuTracker:
    """gvs_multi (include/gvs.h): one process, one handle set and one host thread per device, frames split into
    contiguous ranges over the devices, results gathered on the host into dense arrays."""

    def __init__(self, fx: float, baseline: float, devices, *, width: int, height: int, frames_per_device: int, nFeatures=2000, nLevels=8,
                 imScale=1.2, edgeThreshold=19, patchSize=31, maxFastThreshold=20, minFastThreshold=7):
        self._lib = load_library()
        self.width, self.height, self.frames_per_device = int(width), int(height), int(frames_per_device)
        self.devices = np.ascontiguousarray(devices, np.int32)
        p = _Params(int(nFeatures), int(nLevels), float(imScale), int(edgeThreshold), int(patchSize), int(maxFastThreshold), int(minFastThreshold))
        self.cap_per_image = 2 * int(nFeatures) + 64 * int(nLevels)
        self.h = C.c_void_p()
        _check(self._lib.gvs_multi_create(C.byref(self.h), C.byref(p), self.width, self.height, self.frames_per_device, float(fx),
                                          float(baseline), _p(self.devices), len(self.devices)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.gvs_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def track_stereo(self, left: np.ndarray, right: np.ndarray, motion: dict | None = None, rad: float = 10.0):
        """All frames over all devices; returns a dict of dense arrays with global offsets (and the temporal results when `motion`,
        a synth.sequence_motion-like dict, is given)."""
        assert left.shape == right.shape and left.dtype == np.uint8 and left.flags.c_contiguous and right.flags.c_contiguous
        n, h, w = left.shape
        cap = n * self.cap_per_image
        offL = np.zeros(n + 1, np.int32); offR = np.zeros(n + 1, np.int32)
        kl = np.zeros(cap, KEYPOINT_DTYPE); kr = np.zeros(cap, KEYPOINT_DTYPE)
        dl = np.zeros((cap, 32), np.uint8); dr = np.zeros((cap, 32), np.uint8)
        ri = np.zeros(cap, np.int32); li = np.zeros(cap, np.int32); de = np.zeros(cap, np.float32); cl = np.zeros(cap, np.uint8)
        tp = None
        mo = mk = mt = nm = None
        keep = []
        if motion is not None:
            c = lambda x, dt: np.ascontiguousarray(x, dt)
            keep = [c(motion["prev"], np.int32), c(motion["pose_prev"], np.float64), c(motion["to_camera"], np.float64), c(motion["to_camera_r"], np.float64)]
            tp = C.byref(_TemporalParams(0, float(rad), float(motion["fx"]), float(motion["fy"]), float(motion["cx"]), float(motion["cy"]), w, h,
                                         keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data))
            mo = np.zeros(n + 1, np.int32); mk = np.zeros(cap, np.int32); mt = np.zeros((cap, 2), np.int32); nm = np.zeros(n, np.int32)
        _check(self._lib.gvs_multi_track_stereo(self.h, _p(left), _p(right), h * w, w, n, tp, _p(offL), _p(offR), _p(kl), _p(dl), cap, _p(kr), _p(dr),
                                                cap, _p(ri), _p(li), _p(de), _p(cl), _p(mo), _p(mk), _p(mt), cap, _p(nm)))
        nl, nr = int(offL[-1]), int(offR[-1])
        out = dict(offL=offL, offR=offR, kpL=kl[:nl], descL=dl[:nl], kpR=kr[:nr], descR=dr[:nr], rightIdxs=ri[:nl], leftIdxs=li[:nr],
                   estimatedDepth=de[:nl], close=cl[:nl].astype(bool))
        if motion is not None:
            M = int(mo[-1])
            out.update(mp_offsets=mo, mp_key=mk[:M], matches=mt[:M], n_matches=nm)
        return out


def extractORBAndStereoMatch(fm: FeatureMatcher, leftIm: np.ndarray, rightIm: np.ndarray) -> TrackedKeys:
    """FeatureTracker::extractORBAndStereoMatch (src/FeatureTracker.cpp:56-70) for one stereo pair."""
    h, w = leftIm.shape
    fm.track_stereo_batch(np.ascontiguousarray(leftIm[None]), np.ascontiguousarray(rightIm[None]))
    return tracked_keys(fm, 0, w, h)


def tracked_keys(fm: FeatureMatcher, frame: int, width: int, height: int) -> TrackedKeys:
    """Assemble the TrackedKeys of one frame of the last batch from the device-resident results."""
    k = TrackedKeys()
    offL = fm.feLeft.offsets(frame + 1); offR = fm.feRight.offsets(frame + 1)
    nl = int(offL[frame + 1] - offL[frame]); nr = int(offR[frame + 1] - offR[frame])
    k.keyPoints, k.Desc = fm.feLeft.download(frame, 1, nl)
    k.rightKeyPoints, k.rightDesc = fm.feRight.download(frame, 1, nr)
    k.rightIdxs, k.leftIdxs, k.estimatedDepth, k.close = fm.stereo_download(frame, 1)
    k.lkeyGrid = fm.grid_download(frame, 0)
    k.rkeyGrid = fm.grid_download(frame, 1)
    k.xGrids, k.yGrids = 64, k.lkeyGrid["yGrids"]
    k.xMult, k.yMult = k.lkeyGrid["xMult"], k.lkeyGrid["yMult"]
    return k
