/*
 * gvs.h -- C ABI of the B200-native frontend hot path of gtsam-vSLAM.
 *
 * The reference (christoskokas/gtsam-vSLAM) has no FFI / plugin layer: its boundary for this path is the
 * public C++ surface of FeatureExtractor, FeatureMatcher and two FeatureTracker helpers.  Every entry point
 * below names the reference interface it replaces (file:line under the reference tree).  The C++ adapter
 * classes in include/gvs_adapter.hpp keep the reference's class and member names on top of this ABI.
 *
 * Conventions
 *  - plain pointers and sizes only; all functions return 0 (GVS_OK) or a negative gvs_status;
 *    gvs_last_error() gives the text for the calling thread.  The reference itself has no error channel
 *    (void returns, unchecked indexing); conditions that are UB there are GVS_ERR_* here.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with GVS_ERR_CUDA.
 *  - a gvs_extractor is one FeatureExtractor instance with `max_images` image slots, so that a batch of
 *    independent frames runs as one set of launches.  Handles own their CUDA stream.
 *  - threads: different handles may be driven from different host threads concurrently (the reference runs its two
 *    extractors on two std::threads, src/FeatureTracker.cpp:58-61; gvs_multi drives one handle set per thread).  One
 *    handle is not re-entrant, and a gvs_matcher shares the streams and the resident results of the two extractors it
 *    was created on: calls on a matcher and on its extractors must come from one thread at a time.  A second matcher
 *    over the same extractors (the reference wires FeatureTracker::fm and System::mFeatureMatcher that way) is only
 *    safe for the calls that do not touch the resident frame: gvs_match_projection_lba, gvs_median_descriptor,
 *    gvs_descriptor_distance, gvs_project_mappoints, gvs_predict_mappoints -- and only while no other thread is inside
 *    a call on the same extractors; give a thread of its own (LocalMapper) its own extractor pair instead.
 *  - results of the last extract stay resident on the device (pyramids, keypoints, descriptors) and are what
 *    the matcher calls read, exactly like FeatureMatcher reads feLeft->imagePyramid (src/FeatureMatcher.cpp:606).
 */
#ifndef GVS_H_
#define GVS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVS_MAX_LEVELS 12
#define GVS_GRID_X 64 /* FeatureTracker::assignKeysToGrids xGrids, src/FeatureTracker.cpp:31 */

typedef enum gvs_status {
    GVS_OK = 0,
    GVS_ERR_INVALID = -1,  /* bad argument */
    GVS_ERR_CUDA = -2,     /* CUDA runtime / driver error, or no device */
    GVS_ERR_CAPACITY = -3, /* an internal or caller buffer was too small; nothing was silently dropped */
    GVS_ERR_STATE = -4     /* call order violated (e.g. match before extract) */
} gvs_status;

/* Same memory layout as cv::KeyPoint (28 bytes): the adapter memcpy's into std::vector<cv::KeyPoint>. */
typedef struct gvs_keypoint {
    float x, y;     /* pt, level coordinates times scalePyramid[octave] (src/FeatureExtractor.cpp:523-524) */
    float size;     /* (float)scaledPatchSize[octave] (:601) */
    float angle;    /* degrees in [0,360), cv::fastAtan2 of the intensity centroid (:338) */
    float response; /* FAST score (integer valued) */
    int32_t octave; /* pyramid level */
    int32_t class_id; /* -1 */
} gvs_keypoint;

/* FeatureExtractor constructor arguments, include/FeatureExtractor.h:80 (defaults in gvs_default_params). */
typedef struct gvs_extractor_params {
    int32_t n_features;
    int32_t n_levels;
    float im_scale;
    int32_t edge_threshold;
    int32_t patch_size;
    int32_t max_fast_threshold;
    int32_t min_fast_threshold;
} gvs_extractor_params;

typedef struct gvs_extractor gvs_extractor; /* replaces class FeatureExtractor, include/FeatureExtractor.h:52-97 */
typedef struct gvs_matcher gvs_matcher;     /* replaces class FeatureMatcher,  include/FeatureMatcher.h:22-65 */

const char* gvs_last_error(void);
int gvs_version(void);
void gvs_default_params(gvs_extractor_params* p); /* 2000, 8, 1.2f, 19, 31, 20, 7 */
int gvs_device_count(void);

/* ------------------------------------------------------------------ FeatureExtractor ---------------- */

/* FeatureExtractor::FeatureExtractor, src/FeatureExtractor.cpp:620-682.  Image size is fixed per handle
 * (the reference re-derives level sizes per call from image.cols/rows, :347; cameras do not change size). */
int gvs_extractor_create(gvs_extractor** out, const gvs_extractor_params* params, int width, int height,
                         int max_images, int device);
int gvs_extractor_destroy(gvs_extractor* fe);

/* Public tables read by KeyFrame / pose optimisation (include/FeatureExtractor.h:71-77;
 * src/FeatureTracker.cpp:75-79).  Each array has n_levels entries; any pointer may be NULL. */
int gvs_extractor_tables(const gvs_extractor* fe, float* scale_pyramid, float* scale_inv_pyramid, float* sigma_factor,
                         float* inv_sigma_factor, int32_t* scaled_patch_size, int32_t* feature_per_level);
/* The same tables from the constructor arguments alone (no handle, no device: they do not depend on the image size), for
 * callers that must publish them before the first image arrives (KeyFrame copies them at once, src/FeatureTracker.cpp:75-79). */
int gvs_tables_from_params(const gvs_extractor_params* params, float* scale_pyramid, float* scale_inv_pyramid, float* sigma_factor,
                           float* inv_sigma_factor, int32_t* scaled_patch_size, int32_t* feature_per_level);
/* imagePyramid[level].cols / .rows (computePyramid, src/FeatureExtractor.cpp:346-347) */
int gvs_extractor_level_size(const gvs_extractor* fe, int level, int* width, int* height);

/* FeatureExtractor::extractKeysNew, src/FeatureExtractor.cpp:481-533, one image (slot 0).
 * gray: host pointer, u8, `stride` bytes per row.  out/desc: caller arrays of `cap` keypoints / cap*32 bytes.
 * Zero keypoints leaves out/desc untouched and sets *n = 0 (:498-499). */
int gvs_extract(gvs_extractor* fe, const uint8_t* gray, int width, int height, int stride, gvs_keypoint* out,
                uint8_t* desc, int cap, int* n);

/* Batched form of extractKeysNew over n_images independent images (throughput path).
 * Host images: image i starts at images + i*image_stride, rows are row_stride bytes apart.  Asynchronous on
 * the handle's stream; gvs_extractor_sync / the download calls wait. */
int gvs_extract_batch_host(gvs_extractor* fe, const uint8_t* images, size_t image_stride, int row_stride, int n_images);
/* Device-resident input: the caller fills the handle's own pitched level-0 store (device pointer *dev,
 * rows *pitch bytes apart, images *image_stride bytes apart), then runs n_images slots. */
int gvs_extractor_input(gvs_extractor* fe, uint8_t** dev, int* pitch, size_t* image_stride);
int gvs_extract_batch_resident(gvs_extractor* fe, int n_images);
/* Upstream image preparation on the device, so that raw camera frames go straight to the GPU:
 * cv::remap(INTER_LINEAR) with the CV_32FC1 maps of cv::initUndistortRectifyMap (src/VIOSlam.cpp:282-301) followed by
 * cvtColor(BGR2GRAY) (src/FeatureTracker.cpp:1131-1135).  Maps: width*height floats each, NULL/NULL = images are already
 * rectified (Camera.rectified).  Raw images: tightly packed, `channels` = 3 (BGR, what cv::imread(IMREAD_COLOR) returns) or 1. */
int gvs_extractor_set_rectify_maps(gvs_extractor* fe, const float* map_x, const float* map_y);
int gvs_extract_batch_host_raw(gvs_extractor* fe, const uint8_t* images, int channels, int n_images);
int gvs_extractor_sync(gvs_extractor* fe);

/* Per-image results of the last batch.  offsets has n_images+1 entries: image i owns [offsets[i], offsets[i+1])
 * of the dense keypoint / descriptor arrays.  Waits for the batch. */
int gvs_extract_offsets(gvs_extractor* fe, int32_t* offsets);
/* Copy keypoints / descriptors of images [first, first+count) (dense, in image order) to the host. */
int gvs_extract_download(gvs_extractor* fe, int first, int count, gvs_keypoint* kps, uint8_t* desc, int cap);

/* Use an externally owned CUDA stream (cudaStream_t as void*) for all work of this handle; NULL restores the
 * handle's own stream. */
int gvs_extractor_set_stream(gvs_extractor* fe, void* cuda_stream);

/* Which stages of this handle stage their tiles with TMA box loads (cp.async.bulk.tensor): bit 0 = the fused pyramid kernel,
 * bit 1 = the FAST kernel.  0 when the driver offers no tensor-map encoder or the geometry does not fit a TMA box (the kernels
 * then stage with ordinary loads; same results).  Negative on a null handle. */
int gvs_extractor_tma_stages(const gvs_extractor* fe);

/* Parity taps (tests only; they stall the stream).  imagePyramid[level] / its blurred clone
 * (src/FeatureExtractor.cpp:512-515), tightly packed w*h bytes. */
int gvs_debug_level_image(gvs_extractor* fe, int image, int level, int blurred, uint8_t* out);
/* Ordered FAST candidates of a level before ssc (computeKeypointsORBNew, :563-608): returns count in *n. */
int gvs_debug_level_candidates(gvs_extractor* fe, int image, int level, int32_t* x, int32_t* y, int32_t* score, int cap,
                               int* n);
/* Keys of a level after ssc (:609-613) in output order, with the integer moments of computeOrientation. */
int gvs_debug_level_kept(gvs_extractor* fe, int image, int level, int32_t* x, int32_t* y, int32_t* score, int cap, int* n);

/* Debugging aid (compute-sanitizer by other means): with GVS_GUARD=1 in the environment every device buffer of the library carries
 * a guard tail; returns how many tails have been overwritten (0 = no write past the end of any buffer) and the number checked.
 * Call it before destroying the handles of interest.  Without GVS_GUARD it checks nothing (*n_checked = 0). */
int gvs_debug_check_guards(int* n_checked);

/* Run the selection stage alone (ordering + cv::sortIdx emulation + ssc, src/FeatureExtractor.cpp:368-468) on a
 * caller-made candidate list of pyramid level `level` (n candidates in emission order, x/y inside the level,
 * score in [1,254]).  Outputs: sorted_idx (n entries, may be NULL) = the permutation cv::sortIdx(SORT_DESCENDING)
 * returns for these scores; kept_idx (cap entries) = indices into the input list of the keys ssc keeps, in
 * output order (all n in input order when n <= featurePerLevel[level]).  Overwrites the handle's last batch. */
int gvs_debug_select(gvs_extractor* fe, int level, int n, const int32_t* x, const int32_t* y, const int32_t* score,
                     int32_t* sorted_idx, int32_t* kept_idx, int cap, int* n_kept);

/* ------------------------------------------------------------------ FeatureMatcher ------------------ */

/* FeatureMatcher::FeatureMatcher, src/FeatureMatcher.cpp:8.  fx = zedptr->mCameraLeft->fx (double,
 * include/Camera.h:72), baseline = zedptr->mBaseline (float, :86), image_height sizes the row buckets (:732). */
int gvs_matcher_create(gvs_matcher** out, gvs_extractor* fe_left, gvs_extractor* fe_right, double fx, float baseline,
                       int image_height);
int gvs_matcher_destroy(gvs_matcher* fm);

/* FeatureMatcher::findStereoMatchesORB2R, src/FeatureMatcher.cpp:528-708, for frames [0,n_frames): frame f
 * matches image slot f of fe_left against slot f of fe_right (device resident, last batch).  Asynchronous. */
int gvs_stereo_match_batch(gvs_matcher* fm, int n_frames);
/* rightIdxs / estimatedDepth / close are indexed like the left keypoints of frames [first,first+count)
 * (dense, left offsets); leftIdxs like the right keypoints (right offsets).  Any pointer may be NULL. */
int gvs_stereo_download(gvs_matcher* fm, int first, int count, int32_t* right_idxs, int32_t* left_idxs,
                        float* estimated_depth, uint8_t* close);

/* FeatureTracker::assignKeysToGrids, src/FeatureTracker.cpp:28-54, for both sides of frames [0,n_frames)
 * (device resident CSR: the grids feed gvs_match_projection).  width/height = zedPtr->mWidth/mHeight. */
int gvs_assign_grids_batch(gvs_matcher* fm, int n_frames, int width, int height);
/* CSR download for one frame and side (0 left, 1 right): cell_start has y_grids*64+1 entries, items n.
 * y_grids = ceil(64 * height / width) (64 for landscape images up to square; at most 512): call with cell_start = items =
 * NULL first to learn it.  Any pointer may be NULL. */
int gvs_grid_download(gvs_matcher* fm, int frame, int side, int32_t* cell_start, int32_t* items, int* y_grids,
                      float* x_mult, float* y_mult);

/* FeatureTracker::extractORBAndStereoMatch, src/FeatureTracker.cpp:56-70, for n_frames host frames:
 * extract left and right (two streams, like the two std::threads), stereo match, assign both grids. */
int gvs_track_stereo_batch_host(gvs_matcher* fm, const uint8_t* left, const uint8_t* right, size_t image_stride,
                                int row_stride, int n_frames, int width, int height);
int gvs_track_stereo_batch_resident(gvs_matcher* fm, int n_frames, int width, int height);
int gvs_matcher_sync(gvs_matcher* fm);
/* Fork/join the two track calls above from/into an externally owned CUDA stream (cudaStream_t as void*): the
 * extractors keep running concurrently on their own streams (like the reference's two extraction threads), but
 * everything is ordered after the work already queued on `cuda_stream`, and work queued on it afterwards waits
 * for the batch.  NULL detaches. */
int gvs_matcher_set_stream(gvs_matcher* fm, void* cuda_stream);

/* Map points as a structure of arrays: the fields of MapPoint that the projection matchers read
 * (include/Map.h; src/FeatureMatcher.cpp:264-267,307-310). */
typedef struct gvs_mappoints {
    int32_t count;
    const uint8_t* desc;          /* count x 32, MapPoint::desc */
    const float* pred_l;          /* count x 2, MapPoint::predL */
    const float* pred_r;          /* count x 2, MapPoint::predR (NULL for mono / radius) */
    const int32_t* scale_level_l; /* MapPoint::scaleLevelL */
    const int32_t* scale_level_r; /* MapPoint::scaleLevelR (NULL for mono / radius) */
    const uint8_t* in_frame;      /* MapPoint::inFrame  (NULL = all true) */
    const uint8_t* in_frame_r;    /* MapPoint::inFrameR (NULL = all true) */
} gvs_mappoints;

typedef enum gvs_match_mode {
    GVS_MATCH_STEREO_PRED = 0, /* FeatureMatcher::matchByProjectionRPred, src/FeatureMatcher.cpp:254-389 */
    GVS_MATCH_MONO = 1,        /* FeatureMatcher::matchByProjectionMono,  :391-456 */
    GVS_MATCH_RADIUS = 2       /* FeatureMatcher::matchByRadius,          :458-526 (map points = last keyframe's
                                  keys: pred_l = pt, scale_level_l = octave, desc = Desc rows) */
} gvs_match_mode;

/* One call of matchByProjection* / matchByRadius against frame `frame` of the last batch (host arrays in,
 * host arrays out).  matched_l (n_left) / matched_r (n_right) / matches (count x 2, pair<int,int>) are
 * in-out exactly as in the reference (callers pre-fill them with -1 or with earlier matches).  For
 * GVS_MATCH_RADIUS matches[i][0] receives the matched key index (the reference appends (KeyFrame*, idx)
 * to keyframeIdxMatchs[i], :522).  *n_matches = the reference's return value.  Requires
 * gvs_assign_grids_batch (and gvs_stereo_match_batch for mode 0) on the extractors' current batch: GVS_ERR_STATE otherwise.
 * A scale level outside [0, n_levels) is GVS_ERR_INVALID (the reference indexes scalePyramid unchecked). */
int gvs_match_projection(gvs_matcher* fm, int frame, int mode, const gvs_mappoints* mps, float rad, int32_t* matched_l,
                         int32_t* matched_r, int32_t* matches, int* n_matches);

/* The same call for frames [first_frame, first_frame + n_frames) at once (two launches for the whole batch, no host loop):
 * frame first_frame + f is matched against the map points [mp_offsets[f], mp_offsets[f + 1]) of `mps` (mp_offsets[0] = 0,
 * mp_offsets[n_frames] = mps->count).  matched_l / matched_r are dense over the frames' keys (like gvs_extract_download of the
 * same range), matches is mps->count x 2, n_matches has n_frames entries (the reference's return value per frame).
 * Reference call sites: src/FeatureTracker.cpp:1205,1239 (stereo), :1400-1416 (mono), :1518 (radius). */
int gvs_match_projection_batch(gvs_matcher* fm, int first_frame, int n_frames, int mode, const gvs_mappoints* mps,
                               const int32_t* mp_offsets, float rad, int32_t* matched_l, int32_t* matched_r, int32_t* matches,
                               int32_t* n_matches);

/* Temporal tracking of consecutive stereo frames, device resident end to end (the timed step of bench.py: extract + stereo +
 * temporal match).  For every frame f of the last batch with prev[f] >= 0 the call replays, on the device,
 *   FeatureTracker::initializeMap's point creation (src/FeatureTracker.cpp:89-112) with MapPoint::update (src/Map.cpp:61-104):
 *     one map point per left key of frame prev[f] with estimatedDepth > 0, wp = pose_prev[f] * ((x-cx)z/fx, (y-cy)z/fy, z, 1),
 *     desc = its descriptor row, maxScaleDist = |wp - t(pose_prev[f])| * scalePyramid[octave];
 *   FeatureTracker::removeOutOfFrameMPs (:910-940) -> worldToFrame (:685-741) -> MapPoint::predictScale (src/Map.cpp:13-23):
 *     both projections with to_camera[f] / to_camera_r[f]; points that leave either camera are dropped, order kept;
 *   FeatureMatcher::matchByProjectionRPred (src/FeatureMatcher.cpp:254-389) of the surviving points against frame f with radius
 *     `rad`, on fresh matchedIdxsL / matchedIdxsR / matchesIdxs (:1176-1178).
 * Frames are independent of each other (a frame only reads the stereo results of prev[f]), so the whole batch is a handful of
 * launches.  Asynchronous (the parameter arrays are free again when the call returns); requires gvs_track_stereo_batch_* (or
 * stereo + grids) on the current batch. */
typedef struct gvs_temporal_params {
    int32_t mode;              /* GVS_MATCH_STEREO_PRED */
    float rad;                 /* 10 after the first frame, 120 at curFrame == 1 (src/FeatureTracker.cpp:1189-1192) */
    double fx, fy, cx, cy;     /* FeatureTracker::fx .. cy = zedPtr->mCameraLeft intrinsics (src/FeatureTracker.cpp:23) */
    int32_t width, height;     /* zedPtr->mWidth / mHeight (worldToFrame's bounds test) */
    const int32_t* prev;       /* n_frames: index of the frame whose stereo keys become frame f's map points, -1 = none */
    const double* pose_prev;   /* n_frames x 16, row-major: zedPtr->mCameraPose.pose when frame prev[f] was tracked */
    const double* to_camera;   /* n_frames x 16: predNPose^-1 of frame f (:914) */
    const double* to_camera_r; /* n_frames x 16: (predNPose * extrinsics)^-1 (:913) */
} gvs_temporal_params;
int gvs_track_temporal_batch(gvs_matcher* fm, int n_frames, const gvs_temporal_params* params);
/* Results for frames [first, first + count) (waits for the batch).  mp_offsets (count + 1, relative to the first frame of the
 * range): frame f's map points; mp_key: the left key of frame prev[f] each map point was created from (= activeMapPoints after
 * removeOutOfFrameMPs, in order); matches (x 2) / n_matches (count) / matched_l / matched_r as in gvs_match_projection_batch;
 * pred_l / pred_r (x 2 floats): MapPoint::predL / predR (parity taps).  Any pointer may be NULL. */
int gvs_temporal_download(gvs_matcher* fm, int first, int count, int32_t* mp_offsets, int32_t* mp_key, int32_t* matches,
                          int32_t* n_matches, int32_t* matched_l, int32_t* matched_r, float* pred_l, float* pred_r);

/* FeatureMatcher::matchByProjectionRPredLBA, src/FeatureMatcher.cpp:66-252 (LocalMapper::findNewPoints,
 * src/OptimizationBA.cpp:367).  newKF is any stored key frame, so its keys come from the host: */
typedef struct gvs_lba_frame {
    int32_t n_left, n_right;
    const gvs_keypoint* kp_left;   /* newKF->keys.keyPoints */
    const uint8_t* desc_left;      /* newKF->keys.Desc, n_left x 32 */
    const gvs_keypoint* kp_right;  /* newKF->keys.rightKeyPoints */
    const uint8_t* desc_right;     /* newKF->keys.rightDesc */
    const int32_t* left_idxs;      /* newKF->keys.leftIdxs, n_right */
    const int32_t* right_idxs;     /* newKF->keys.rightIdxs, n_left */
    const int32_t* unmatched_f;    /* newKF->unMatchedF, n_left (>= 0: the key already belongs to a map point) */
    const int32_t* unmatched_fr;   /* newKF->unMatchedFR, n_right */
    int32_t width, height;         /* image size the key grids were built with (src/FeatureTracker.cpp:67-68) */
} gvs_lba_frame;
/* The candidate points p4d[i] (:72-110), flattened by the caller exactly as the reference reads them: */
typedef struct gvs_lba_points {
    int32_t count;
    const uint8_t* desc;        /* mp->desc, or lastKeys.Desc / rightDesc row (:76-86) */
    const float* pred_l;        /* predPoints[i].first  (searched only if x > 0 && y > 0, :104) */
    const float* pred_r;        /* predPoints[i].second (:153) */
    const int32_t* pred_scale;  /* clamp(cvCeil(log(maxDistsScale[i]/dist)/logScale), 0, nScaleLev-1) (:88-95), host double math */
    const float* last_l;        /* lastKFKeyPoint used by the left branch  (:231-235) */
    const float* last_r;        /* lastKFKeyPoint used by the right branch (:211-215) */
} gvs_lba_points;
/* out_pairs: count x 2; (lIdx, rIdx) appended to matchedIdxs[i] (:226,:246) or (-2,-2) if nothing is appended.
 * *n_matches = the reference's return value. */
int gvs_match_projection_lba(gvs_matcher* fm, const gvs_lba_frame* new_kf, const gvs_lba_points* pts, float rad,
                             int32_t* out_pairs, int* n_matches);

/* FeatureTracker::worldToFrame (src/FeatureTracker.cpp:685-741) with MapPoint::predictScale (src/Map.cpp:13-23) for `count`
 * map points and one camera (it produces predL/predR, scaleLevelL/R, inFrame/inFrameR, the inputs of
 * gvs_match_projection).  pose_inv: predPoseInv, 16 doubles row-major; world_pos: count x 4 doubles (MapPoint::wp);
 * max_scale_dist: MapPoint::maxScaleDist; log_scale / n_scale_lev: lastObsKF->logScale / nScaleLev per point.
 * pred (count x 2) and scale_level are in-out: a point that is not in the frame only gets in_frame = 0 (:697-718). */
int gvs_project_mappoints(gvs_matcher* fm, const double* pose_inv, double fx, double fy, double cx, double cy, int width, int height,
                          int count, const double* world_pos, const float* max_scale_dist, const float* log_scale,
                          const int32_t* n_scale_lev, float* pred, int32_t* scale_level, uint8_t* in_frame);

/* The worldToFrame loops of the tracking thread, for `count` active map points at once:
 *   mode 0  FeatureTracker::removeOutOfFrameMPs     (src/FeatureTracker.cpp:910-940): both cameras, keep[i] = c1 && c2
 *   mode 1  FeatureTracker::removeOutOfFrameMPsMono (:942-967): left camera only, keep[i] = c1
 *   mode 2  FeatureTracker::PredictMPsPosition      (:969-1013): both cameras; a point that leaves a camera, or that is flagged in
 *           outliers[], loses its match on that side (matched_l / matched_r slot and matches[i] set to -1), outliers[i] is reset
 * to_camera = predNPose^-1, to_camera_r = (predNPose * extrinsics)^-1 (16 doubles row-major; the inverses stay on the host, :913-914).
 * valid[i] = the reference does not `continue` for point i (mp != nullptr, and for modes 0/1 !mp->GetIsOutlier()); NULL = all.
 * pred_l / scale_level_l / in_frame (and the _r arrays; NULL in mode 1) are MapPoint::predL / scaleLevelL / inFrame, in-out: only
 * visited points are written, pred and scale level only when the point is in the frame (worldToFrame, :685-741).
 * keep (modes 0/1, out): 1 = the point stays in activeMapPoints (the caller compacts the vector and calls setActive, :926-937).
 * matches (count x 2), matched_l (n_left), matched_r (n_right), outliers (count): mode 2 only, in-out. */
int gvs_predict_mappoints(gvs_matcher* fm, int mode, const double* to_camera, const double* to_camera_r, double fx, double fy,
                          double cx, double cy, int width, int height, int count, const uint8_t* valid, const double* world_pos,
                          const float* max_scale_dist, const float* log_scale, const int32_t* n_scale_lev, float* pred_l,
                          int32_t* scale_level_l, uint8_t* in_frame, float* pred_r, int32_t* scale_level_r, uint8_t* in_frame_r,
                          uint8_t* keep, int32_t* matches, int32_t* matched_l, int n_left, int32_t* matched_r, int n_right,
                          uint8_t* outliers);

/* FeatureTracker::findOutliersR (src/FeatureTracker.cpp:582-649; mode 0) and findOutliersMono (:651-684; mode 1) with
 * check2dError (:147-164), for the `count` active map points against frame `frame` of the last batch (the
 * outlier scoring that follows every pose estimate).  The frame's keys and stereo arrays are the device-resident ones, and the
 * stereo arrays are UPDATED on the device when a stereo partner is rejected (:636-641), so the next gvs_match_projection call
 * sees what the reference's next matchByProjectionRPred sees in keysLeft.
 *   pose      estimPose, 16 doubles row-major;  pose_r  toCameraR = (estimPose^-1 * extrinsics)^-1 (:584-585; mode 0 only)
 *   thres     chi-square threshold; close_dist = zedPtr->mBaseline * fm.closeNumber (:623)
 *   valid     activeMapPoints[i] != nullptr (NULL = all valid); world_pos  count x 4 doubles (MapPoint::getWordPose4d)
 *   in_frame / in_frame_r  MapPoint::inFrame / inFrameR (NULL = all true)
 *   matches   count x 2 (pair<int,int>), in-out: .second is set to -1 where the right observation is rejected (:641);
 *             indices must lie inside the frame's key arrays (the reference indexes unchecked)
 *   outliers  count, in-out: written exactly where the reference writes MPsOutliers[i]
 *   est_depth, close (n_left), right_idxs (n_left), left_idxs (n_right): optional host copies of the frame's stereo arrays
 *             (TrackedKeys::estimatedDepth / close / rightIdxs / leftIdxs), refreshed from the device after the call
 *   *n_inliers = number of nInliers++ executed (the reference increments its int& argument); *n_stereo = findOutliersR's
 *             return value (0 in mode 1, where the reference returns nInliers). */
int gvs_find_outliers(gvs_matcher* fm, int frame, int mode, const double* pose, const double* pose_r, double fx, double fy,
                      double cx, double cy, double thres, float close_dist, int count, const uint8_t* valid,
                      const double* world_pos, const uint8_t* in_frame, const uint8_t* in_frame_r, int32_t* matches,
                      uint8_t* outliers, float* est_depth, uint8_t* close, int32_t* right_idxs, int32_t* left_idxs,
                      int* n_inliers, int* n_stereo);

/* MapPoint::calcDescriptor, src/Map.cpp:145-210, for n_points map points at once: descs = the observation descriptors
 * of all points back to back (32 bytes each), offsets[n_points + 1]; best_idx[p] = index inside point p's own list of
 * the descriptor with the least median Hamming distance to the others (-1 for an empty list). */
int gvs_median_descriptor(gvs_matcher* fm, const uint8_t* descs, const int32_t* offsets, int n_points, int32_t* best_idx);

/* FeatureMatcher::DescriptorDistance, src/FeatureMatcher.cpp:710-726, for n pairs of 32-byte descriptors
 * (host arrays; computed on the device). */
int gvs_descriptor_distance(gvs_matcher* fm, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* ---- frame pipeline ---------------------------------------------------------------------------
 * The reference tracks one frame at a time, synchronously (System::trackNewImage -> FeatureTracker::TrackImageT ->
 * extractORBAndStereoMatch, src/FeatureTracker.cpp:1120-1145 and :56-70).  A gvs_pipeline is a ring of n_slots independent
 * handle sets (two extractors + one matcher, own CUDA streams), each taking up to frames_per_slot stereo frames per ticket:
 * while the host collects ticket t, tickets t+1 ... upload and compute.  Ticket t uses slot t % n_slots; a slot is free again
 * after gvs_pipeline_release.  Not thread safe (one submitting thread), like the reference's tracking thread. */
typedef struct gvs_pipeline gvs_pipeline;
int gvs_pipeline_create(gvs_pipeline** out, const gvs_extractor_params* params, int width, int height, int frames_per_slot,
                        int n_slots, double fx, float baseline, int device);
int gvs_pipeline_destroy(gvs_pipeline* p);
/* Asynchronous gvs_track_stereo_batch_host on the next slot (host images should be pinned and must stay valid until
 * gvs_pipeline_wait returns).  GVS_ERR_STATE if that slot still holds an unreleased ticket. */
int gvs_pipeline_submit(gvs_pipeline* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride,
                        int n_frames, int64_t* ticket);
/* The same followed by gvs_track_temporal_batch(temporal) on the ticket's frames (prev[] indexes frames of the ticket);
 * temporal == NULL is gvs_pipeline_submit.  The parameter arrays are free again when the call returns. */
int gvs_pipeline_submit_temporal(gvs_pipeline* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride,
                                 int n_frames, const gvs_temporal_params* temporal, int64_t* ticket);
/* Blocks until the ticket is computed; offsets (n_frames + 1 entries each, may be NULL) as gvs_extract_offsets.  Device-side
 * failures of the ticket (GVS_ERR_CAPACITY ...) are returned here. */
int gvs_pipeline_wait(gvs_pipeline* p, int64_t ticket, int32_t* left_offsets, int32_t* right_offsets);
/* Results of all frames of the ticket, dense (as gvs_extract_download / gvs_stereo_download).  Any pointer may be NULL. */
int gvs_pipeline_download(gvs_pipeline* p, int64_t ticket, gvs_keypoint* kp_left, uint8_t* desc_left, int cap_left,
                          gvs_keypoint* kp_right, uint8_t* desc_right, int cap_right, int32_t* right_idxs, int32_t* left_idxs,
                          float* estimated_depth, uint8_t* close);
/* Temporal results of all frames of the ticket (as gvs_temporal_download over the whole ticket). */
int gvs_pipeline_download_temporal(gvs_pipeline* p, int64_t ticket, int32_t* mp_offsets, int32_t* mp_key, int32_t* matches,
                                   int32_t* n_matches);
/* The slot's handles, e.g. for gvs_match_projection / gvs_grid_download against the ticket's frames (valid until release). */
int gvs_pipeline_handles(gvs_pipeline* p, int64_t ticket, gvs_extractor** fe_left, gvs_extractor** fe_right, gvs_matcher** fm);
int gvs_pipeline_release(gvs_pipeline* p, int64_t ticket);

/* ---- several GPUs in one process ----------------------------------------------------------------------
 * Stereo frames are independent: a batch of n_frames is cut into contiguous ranges [g*n/G, (g+1)*n/G) over the G devices of the
 * handle; every device has its own handle set and is driven by its own host thread; results are gathered on the host into the
 * caller's dense arrays (offsets are global over the batch).  No collective and no peer access on the data path.  A device may
 * be named more than once (two handle sets on one GPU). */
typedef struct gvs_multi gvs_multi;
int gvs_multi_create(gvs_multi** out, const gvs_extractor_params* params, int width, int height, int frames_per_device, double fx,
                     float baseline, const int* devices, int n_devices);
int gvs_multi_destroy(gvs_multi* p);
int gvs_multi_device_count(const gvs_multi* p);
/* FeatureTracker::extractORBAndStereoMatch (src/FeatureTracker.cpp:56-70) for n_frames host frames over all devices, followed by
 * gvs_track_temporal_batch when `temporal` is given (prev[] indexes frames of the whole batch; a frame and its predecessor must
 * fall on the same device: GVS_ERR_INVALID otherwise).  left_offsets / right_offsets (n_frames + 1) and, with temporal,
 * mp_offsets (n_frames + 1) / n_matches (n_frames) are required; any other output may be NULL.  cap_* are the capacities of the
 * caller's key / map point arrays.  Blocks until everything is on the host. */
int gvs_multi_track_stereo(gvs_multi* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride, int n_frames,
                           const gvs_temporal_params* temporal, int32_t* left_offsets, int32_t* right_offsets, gvs_keypoint* kp_left,
                           uint8_t* desc_left, int cap_left, gvs_keypoint* kp_right, uint8_t* desc_right, int cap_right,
                           int32_t* right_idxs, int32_t* left_idxs, float* estimated_depth, uint8_t* close, int32_t* mp_offsets,
                           int32_t* mp_key, int32_t* matches, int cap_mp, int32_t* n_matches);

/* ------------------------------------------------------------------ measurement --------------------- */

/* Launch counters and device timings of the last batch, for bench.py (gpu_launches, per-stage ms). */
typedef struct gvs_stage_times {
    float pyramid_ms, fast_ms, select_ms, blur_ms, describe_ms, stereo_ms, grid_ms, temporal_ms;
    int32_t launches;
} gvs_stage_times;
int gvs_extractor_enable_timing(gvs_extractor* fe, int on);
int gvs_extractor_stage_times(gvs_extractor* fe, gvs_stage_times* t);
int gvs_matcher_launch_count(gvs_matcher* fm, int64_t* launches); /* kernels launched since creation */
int gvs_extractor_launch_count(gvs_extractor* fe, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GVS_H_ */
