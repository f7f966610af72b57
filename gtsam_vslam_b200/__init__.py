"""gtsam_vslam_b200 -- B200-native frontend hot path of gtsam-vSLAM (feature extraction + stereo / temporal matching).

The product is the CUDA library csrc/ -> libgvs.so behind the C ABI of include/gvs.h; this package is the thin
host-side mirror of the reference's FeatureExtractor / FeatureMatcher / FeatureTracker interfaces.
"""
from .api import (FeatureExtractor, FeatureMatcher, GVS_MATCH_MONO, GVS_MATCH_RADIUS, GVS_MATCH_STEREO_PRED, GvsError, KEYPOINT_DTYPE, LIB_PATH, MultiGpuTracker, SYMBOLS, StageTimes, StereoPipeline, TrackedKeys,
                  device_count, extractORBAndStereoMatch, load_library, tracked_keys)

__all__ = ["FeatureExtractor", "FeatureMatcher", "GVS_MATCH_MONO", "GVS_MATCH_RADIUS", "GVS_MATCH_STEREO_PRED", "GvsError", "KEYPOINT_DTYPE", "LIB_PATH", "MultiGpuTracker", "SYMBOLS", "StageTimes",
           "StereoPipeline", "TrackedKeys", "device_count", "extractORBAndStereoMatch", "load_library", "tracked_keys"]
