// Drop-in C++ adapter: the reference's FeatureExtractor / FeatureMatcher / TrackedKeys (same namespace, class,
// member and method names, same argument meaning) implemented on the C ABI of include/gvs.h.
//
// A maintainer of gtsam-vSLAM replaces include/FeatureExtractor.h + src/FeatureExtractor.cpp and
// include/FeatureMatcher.h + src/FeatureMatcher.cpp by this header and links libgvs.so;
// System, FeatureTracker, KeyFrame, Map and LocalMapper compile against it unchanged apart from the two
// constructor calls.  The constructors keep the reference's argument lists (include/FeatureExtractor.h:80,
// include/FeatureMatcher.h:40): the device handles are created lazily from the size of the first image.
//
//   reference                                                       here
//   FeatureExtractor::FeatureExtractor   src/FeatureExtractor.cpp:620   FeatureExtractor(nfeatures = 2000, ..., minFastThreshold = 7)
//   FeatureExtractor::extractKeysNew     src/FeatureExtractor.cpp:481   extractKeysNew(image, keypoints, descriptors)
//   public tables                        include/FeatureExtractor.h:71  scalePyramid, scaleInvPyramid, sigmaFactor, ...
//   FeatureMatcher::findStereoMatchesORB2R src/FeatureMatcher.cpp:528   findStereoMatchesORB2R(lImage, rImage, rightDesc, rightKeys, keysLeft)
//   FeatureMatcher::matchByProjectionRPred src/FeatureMatcher.cpp:254   matchByProjectionRPred(activeMapPoints, keysLeft, ...)
//   FeatureMatcher::matchByProjectionMono  src/FeatureMatcher.cpp:391   matchByProjectionMono(...)
//   FeatureMatcher::matchByRadius          src/FeatureMatcher.cpp:458   matchByRadius(...)
//   FeatureMatcher::matchByProjectionRPredLBA src/FeatureMatcher.cpp:66 matchByProjectionRPredLBA(lastKF, newKF, ...) (template)
//   MapPoint::calcDescriptor               src/Map.cpp:145              FeatureMatcher::medianDescriptor(vDescriptors)
//   FeatureMatcher::DescriptorDistance     src/FeatureMatcher.cpp:710   DescriptorDistance(a, b)
//   FeatureTracker::assignKeysToGrids      src/FeatureTracker.cpp:28    FeatureMatcher::assignKeysToGrids(keysLeft, width, height)
//   FeatureTracker::extractORBAndStereoMatch src/FeatureTracker.cpp:56  FeatureMatcher::extractORBAndStereoMatch(leftIm, rightIm, keysLeft)
//
// Error behaviour: the reference returns void and never reports errors; here a failing C-ABI call throws
// std::runtime_error with gvs_last_error() (there is no CPU fallback to fall through to).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifdef GVS_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include "gvs_cv_shim.hpp"
#endif
#include "gvs.h"

namespace GTSAM_VIOSLAM {

#ifdef GVS_HAVE_OPENCV
#define GVS_CV_8U CV_8U
#else
#define GVS_CV_8U cv::CV_8U
#endif

static_assert(sizeof(cv::KeyPoint) == sizeof(gvs_keypoint), "cv::KeyPoint must be the 28-byte POD the ABI copies");

inline void gvsCheck(int rc) {
    if (rc != GVS_OK) throw std::runtime_error(std::string("gvs: ") + gvs_last_error());
}

struct TrackedKeys {  // include/FeatureExtractor.h:18-50
    std::vector<cv::KeyPoint> keyPoints, rightKeyPoints;
    std::vector<std::vector<std::vector<int>>> lkeyGrid, rkeyGrid;
    std::vector<int> rightIdxs, leftIdxs;
    std::vector<float> estimatedDepth;
    std::vector<bool> close;
    cv::Mat Desc, rightDesc;
    float xMult = 0, yMult = 0;
    int xGrids = 0, yGrids = 0;
    void getKeys(const TrackedKeys& k) {
        keyPoints = k.keyPoints; rightKeyPoints = k.rightKeyPoints; lkeyGrid = k.lkeyGrid; rkeyGrid = k.rkeyGrid;
        rightIdxs = k.rightIdxs; leftIdxs = k.leftIdxs; Desc = k.Desc.clone(); rightDesc = k.rightDesc.clone();
        close = k.close; estimatedDepth = k.estimatedDepth; xMult = k.xMult; yMult = k.yMult; xGrids = k.xGrids; yGrids = k.yGrids;
    }
};

class FeatureExtractor {
  public:
    const int nFeatures;
    const size_t nLevels;
    const float imScale;
    const int edgeThreshold, patchSize;
    const int halfPatchSize{15};
    const int maxFastThreshold, minFastThreshold;
    // include/FeatureExtractor.h:71.  The pyramid lives on the device (only findStereoMatchesORB2R reads it, src/FeatureMatcher.cpp:
    // 606-623, and that runs on the device too); set mirrorImagePyramid to have extractKeysNew copy the levels back into this member.
    std::vector<cv::Mat> imagePyramid;
    bool mirrorImagePyramid = false;
    std::vector<int> scaledPatchSize, featurePerLevel;
    std::vector<float> scalePyramid, scaleInvPyramid, sigmaFactor, InvSigmaFactor;

    // include/FeatureExtractor.h:80: the reference's seven arguments and defaults (src/System.cpp:31,56-57).  The device handle is
    // created by the first call that sees an image (the reference also learns the image size per call, src/FeatureExtractor.cpp:347).
    FeatureExtractor(const int _nfeatures = 2000, const int _nLevels = 8, const float _imScale = 1.2f, const int _edgeThreshold = 19,
                     const int _patchSize = 31, const int _maxFastThreshold = 20, const int _minFastThreshold = 7)
        : nFeatures(_nfeatures), nLevels(_nLevels), imScale(_imScale), edgeThreshold(_edgeThreshold), patchSize(_patchSize),
          maxFastThreshold(_maxFastThreshold), minFastThreshold(_minFastThreshold) {
        const gvs_extractor_params p = params();
        scalePyramid.resize(nLevels); scaleInvPyramid.resize(nLevels); sigmaFactor.resize(nLevels); InvSigmaFactor.resize(nLevels);
        scaledPatchSize.resize(nLevels); featurePerLevel.resize(nLevels);
        gvsCheck(gvs_tables_from_params(&p, scalePyramid.data(), scaleInvPyramid.data(), sigmaFactor.data(), InvSigmaFactor.data(),
                                        scaledPatchSize.data(), featurePerLevel.data()));
    }
    ~FeatureExtractor() { gvs_extractor_destroy(h_); }
    FeatureExtractor(const FeatureExtractor&) = delete;
    FeatureExtractor& operator=(const FeatureExtractor&) = delete;

    // src/FeatureExtractor.cpp:481-533.  Zero keypoints leaves both outputs untouched (:498-499).
    void extractKeysNew(cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {
        ensure(image.cols, image.rows);
        const int cap = 2 * nFeatures + 64 * (int)nLevels;
        std::vector<cv::KeyPoint> k(cap);
        cv::Mat d(cap, 32, GVS_CV_8U);
        int n = 0;
        gvsCheck(gvs_extract(h_, image.data, image.cols, image.rows, (int)image.step, reinterpret_cast<gvs_keypoint*>(k.data()),
                             d.data, cap, &n));
        if (mirrorImagePyramid) fetchImagePyramid();
        if (n <= 0) return;
        k.resize(n);
        keypoints = std::move(k);
        descriptors = cv::Mat(n, 32, GVS_CV_8U);
        std::memcpy(descriptors.data, d.data, (size_t)n * 32);
    }
    // FeatureExtractor::computePyramid (src/FeatureExtractor.cpp:342-366): the levels of `image` in imagePyramid (the device runs
    // its whole per-image pipeline; the levels are what is copied back)
    void computePyramid(const cv::Mat& image) {
        ensure(image.cols, image.rows);
        gvsCheck(gvs_extract_batch_host(h_, image.data, (size_t)image.step * image.rows, (int)image.step, 1));
        fetchImagePyramid();
    }
    // imagePyramid[level] of the last image, copied from the device (a parity tap: it waits for the stream)
    void fetchImagePyramid() {
        imagePyramid.resize(nLevels);
        for (size_t l = 0; l < nLevels; l++) {
            int w = 0, h = 0;
            gvsCheck(gvs_extractor_level_size(h_, (int)l, &w, &h));
            imagePyramid[l] = cv::Mat(h, w, GVS_CV_8U);
            gvsCheck(gvs_debug_level_image(h_, 0, (int)l, 0, imagePyramid[l].data));
        }
    }
    // FeatureExtractor::ssc (src/FeatureExtractor.cpp:368-468) as the extractor itself calls it (:612): tolerance 0.1, (cols, rows)
    // = a pyramid level's size and numRetPoints = that level's featurePerLevel.  The selection stage of the device pipeline is run
    // on the given points (pt = level coordinates, response = FAST score); other argument combinations are not on the path and
    // throw std::invalid_argument.
    std::vector<cv::KeyPoint> ssc(std::vector<cv::KeyPoint> keyPoints, int numRetPoints, float tolerance, int cols, int rows) {
        if (!h_) throw std::logic_error("FeatureExtractor::ssc: no image has been seen yet (the level sizes are unknown)");
        int level = -1;
        for (size_t l = 0; l < nLevels && level < 0; l++) {
            int w = 0, h = 0;
            gvsCheck(gvs_extractor_level_size(h_, (int)l, &w, &h));
            if (w == cols && h == rows && featurePerLevel[l] == numRetPoints) level = (int)l;
        }
        if (level < 0 || tolerance != 0.1f) throw std::invalid_argument("FeatureExtractor::ssc: only the extractor's own per-level calls run on the device");
        const int n = (int)keyPoints.size();
        std::vector<int32_t> x(n), y(n), sc(n), kept(n + 1);
        for (int i = 0; i < n; i++) { x[i] = (int)keyPoints[i].pt.x; y[i] = (int)keyPoints[i].pt.y; sc[i] = (int)keyPoints[i].response; }
        int nk = 0;
        gvsCheck(gvs_debug_select(h_, level, n, x.data(), y.data(), sc.data(), nullptr, kept.data(), n + 1, &nk));
        std::vector<cv::KeyPoint> out(nk);
        for (int i = 0; i < nk; i++) out[i] = keyPoints[kept[i]];
        return out;
    }

    // device handle for images of this size (created on first use; the size of a camera does not change)
    gvs_extractor* ensure(int width, int height) {
        if (!h_) {
            const gvs_extractor_params p = params();
            gvsCheck(gvs_extractor_create(&h_, &p, width, height, 1, device));
            width_ = width; height_ = height;
        } else if (width != width_ || height != height_) {
            throw std::invalid_argument("FeatureExtractor: image size differs from the first image's");
        }
        return h_;
    }
    gvs_extractor* handle() const { return h_; }
    int width() const { return width_; }
    int height() const { return height_; }
    int device = 0;   // CUDA device of the lazily created handle

  private:
    gvs_extractor_params params() const {
        return gvs_extractor_params{nFeatures, (int32_t)nLevels, imScale, edgeThreshold, patchSize, maxFastThreshold, minFastThreshold};
    }
    gvs_extractor* h_ = nullptr;
    int width_ = 0, height_ = 0;
};

class FeatureMatcher {
  public:
    const int closeNumber{40};
    std::shared_ptr<FeatureExtractor> feLeft, feRight;
    const int imageHeight;

    // include/FeatureMatcher.h:40 (src/System.cpp:32,58, src/FeatureTracker.cpp:23): StereoCameraT is the reference's StereoCamera;
    // the path reads zedptr->mCameraLeft->fx (double, include/Camera.h:72) and zedptr->mBaseline (float, :86).
    template <class StereoCameraT>
    FeatureMatcher(std::shared_ptr<StereoCameraT> _zed, std::shared_ptr<FeatureExtractor> _feLeft, std::shared_ptr<FeatureExtractor> _feRight,
                   const int _imageHeight = 360)
        : feLeft(std::move(_feLeft)), feRight(std::move(_feRight)), imageHeight(_imageHeight), fx_(_zed->mCameraLeft->fx),
          baseline_(_zed->mBaseline) {}
    // the same with the two camera values spelled out
    FeatureMatcher(double fx, float baseline, std::shared_ptr<FeatureExtractor> _feLeft, std::shared_ptr<FeatureExtractor> _feRight,
                   int _imageHeight = 360)
        : feLeft(std::move(_feLeft)), feRight(std::move(_feRight)), imageHeight(_imageHeight), fx_(fx), baseline_(baseline) {}
    ~FeatureMatcher() { gvs_matcher_destroy(h_); }
    FeatureMatcher(const FeatureMatcher&) = delete;
    FeatureMatcher& operator=(const FeatureMatcher&) = delete;

    // src/FeatureMatcher.cpp:528-708.  The images and rightDesc/rightKeys arguments are kept for signature parity: the
    // device already holds the pyramids, keys and descriptors of the two extractKeysNew calls that precede this call.
    void findStereoMatchesORB2R(const cv::Mat&, const cv::Mat&, const cv::Mat&, std::vector<cv::KeyPoint>& rightKeys, TrackedKeys& keysLeft) {
        gvsCheck(gvs_stereo_match_batch(ensureHandle(), 1));
        // (sized from the device's own key counts: the caller's vectors may be stale when an extraction returned no keys)
        int offL[2] = {0, 0}, offR[2] = {0, 0};
        gvsCheck(gvs_extract_offsets(feLeft->handle(), offL));
        gvsCheck(gvs_extract_offsets(feRight->handle(), offR));
        const size_t nl = (size_t)offL[1], nr = (size_t)offR[1];
        if (nl != keysLeft.keyPoints.size() || nr != rightKeys.size())
            throw std::logic_error("findStereoMatchesORB2R: keysLeft / rightKeys are not the keys of the last extractKeysNew calls");
        keysLeft.rightIdxs.assign(nl, -1); keysLeft.leftIdxs.assign(nr, -1); keysLeft.estimatedDepth.assign(nl, -1.0f);
        std::vector<uint8_t> cl(nl + 1);
        gvsCheck(gvs_stereo_download(h_, 0, 1, keysLeft.rightIdxs.data(), keysLeft.leftIdxs.data(), keysLeft.estimatedDepth.data(), cl.data()));
        keysLeft.close.assign(nl, false);
        for (size_t i = 0; i < nl; i++) keysLeft.close[i] = cl[i] != 0;
    }

    // FeatureTracker::assignKeysToGrids for both sides (src/FeatureTracker.cpp:28-54, :67-68)
    void assignKeysToGrids(TrackedKeys& keysLeft, int width, int height) {
        gvsCheck(gvs_assign_grids_batch(ensureHandle(), 1, width, height));
        for (int side = 0; side < 2; side++) {
            const size_t n = side ? keysLeft.rightKeyPoints.size() : keysLeft.keyPoints.size();
            int yg = 0;
            gvsCheck(gvs_grid_download(h_, 0, side, nullptr, nullptr, &yg, &keysLeft.xMult, &keysLeft.yMult));  // rows first
            std::vector<int> start((size_t)GVS_GRID_X * yg + 1), items(n + 1);
            gvsCheck(gvs_grid_download(h_, 0, side, start.data(), items.data(), &yg, &keysLeft.xMult, &keysLeft.yMult));
            keysLeft.xGrids = GVS_GRID_X; keysLeft.yGrids = yg;
            auto& grid = side ? keysLeft.rkeyGrid : keysLeft.lkeyGrid;
            grid.assign(yg, std::vector<std::vector<int>>(GVS_GRID_X));
            for (int r = 0; r < yg; r++)
                for (int c = 0; c < GVS_GRID_X; c++)
                    grid[r][c].assign(items.begin() + start[r * GVS_GRID_X + c], items.begin() + start[r * GVS_GRID_X + c + 1]);
        }
    }

    // FeatureTracker::extractORBAndStereoMatch (src/FeatureTracker.cpp:56-70): the two extractions run concurrently on the
    // two handles' streams (the reference uses two std::threads), then stereo matching and both grids.
    void extractORBAndStereoMatch(cv::Mat& leftIm, cv::Mat& rightIm, TrackedKeys& keysLeft) {
        gvsCheck(gvs_extract_batch_host(feLeft->ensure(leftIm.cols, leftIm.rows), leftIm.data, leftIm.step * leftIm.rows, (int)leftIm.step, 1));
        gvsCheck(gvs_extract_batch_host(feRight->ensure(rightIm.cols, rightIm.rows), rightIm.data, rightIm.step * rightIm.rows, (int)rightIm.step, 1));
        keysLeft.keyPoints.clear(); keysLeft.rightKeyPoints.clear();   // (a frame without keys must not keep the previous frame's)
        fetch(*feLeft, keysLeft.keyPoints, keysLeft.Desc);
        fetch(*feRight, keysLeft.rightKeyPoints, keysLeft.rightDesc);
        findStereoMatchesORB2R(leftIm, rightIm, keysLeft.rightDesc, keysLeft.rightKeyPoints, keysLeft);
        assignKeysToGrids(keysLeft, leftIm.cols, leftIm.rows);
    }

    // src/FeatureMatcher.cpp:254-389 / :391-456.  MapPointT needs: cv::Mat desc; cv::Point2f predL, predR; int scaleLevelL,
    // scaleLevelR; bool inFrame, inFrameR (include/Map.h).  Runs against the frame of the last extractORBAndStereoMatch.
    template <class MapPointT>
    int matchByProjectionRPred(std::vector<MapPointT*>& activeMapPoints, TrackedKeys& keysLeft, std::vector<int>& matchedIdxsL,
                               std::vector<int>& matchedIdxsR, std::vector<std::pair<int, int>>& matchesIdxs, const float rad) {
        return project(GVS_MATCH_STEREO_PRED, activeMapPoints, keysLeft, matchedIdxsL, &matchedIdxsR, matchesIdxs, rad);
    }
    template <class MapPointT>
    int matchByProjectionMono(std::vector<MapPointT*>& activeMapPoints, TrackedKeys& keysLeft, std::vector<int>& matchedIdxsL,
                              std::vector<std::pair<int, int>>& matchesIdxs, const float rad) {
        return project(GVS_MATCH_MONO, activeMapPoints, keysLeft, matchedIdxsL, nullptr, matchesIdxs, rad);
    }
    // src/FeatureMatcher.cpp:458-526 with the reference's seven arguments (call site src/FeatureTracker.cpp:1518): matchedIdxsL is
    // in-out, (actKeyFrame, idx) is appended to keyframeIdxMatchs[i] for every matched key i of lastKeys (:522) and matchesIdxs is
    // not touched (the reference never reads or writes it in this function).  actKeys must be the last extracted frame.
    template <class KeyFrameT>
    int matchByRadius(TrackedKeys& lastKeys, TrackedKeys& actKeys, std::vector<int>& matchedIdxsL, std::vector<std::pair<int, int>>& matchesIdxs,
                      const float rad, std::vector<std::vector<std::pair<KeyFrameT*, int>>>& keyframeIdxMatchs, KeyFrameT* actKeyFrame) {
        const int M = (int)lastKeys.keyPoints.size();
        std::vector<float> pred(2 * (size_t)M);
        std::vector<int32_t> lvl(M);
        std::vector<uint8_t> desc((size_t)M * 32);
        for (int i = 0; i < M; i++) {
            pred[2 * i] = lastKeys.keyPoints[i].pt.x; pred[2 * i + 1] = lastKeys.keyPoints[i].pt.y; lvl[i] = lastKeys.keyPoints[i].octave;
            std::memcpy(&desc[(size_t)i * 32], lastKeys.Desc.ptr(i), 32);
        }
        gvs_mappoints mp{M, desc.data(), pred.data(), nullptr, lvl.data(), nullptr, nullptr, nullptr};
        std::vector<std::pair<int, int>> found(M, std::make_pair(-1, -1));
        int n = 0;
        (void)actKeys; (void)matchesIdxs;
        gvsCheck(gvs_match_projection(ensureHandle(), 0, GVS_MATCH_RADIUS, &mp, rad, matchedIdxsL.data(), nullptr,
                                      reinterpret_cast<int32_t*>(found.data()), &n));
        for (int i = 0; i < M; i++)
            if (found[i].first >= 0) keyframeIdxMatchs[i].emplace_back(std::make_pair(actKeyFrame, found[i].first));
        return n;
    }

    // FeatureMatcher::getMatchIdxs (src/FeatureMatcher.cpp:13-64) and destributeRightKeys (:728-752): public helpers of the
    // reference class that nothing outside FeatureMatcher.cpp calls.  They only index the CALLER'S host containers (the key
    // grids / key vectors of a TrackedKeys); the device matchers never use them (they enumerate the same windows / buckets in
    // their kernels).  Kept so that code written against the reference header still compiles.
    void getMatchIdxs(const cv::Point2f& predP, std::vector<int>& idxs, const TrackedKeys& keysLeft, const int predictedScale, const float radius,
                      bool right) const {
        const float trackX = predP.x, trackY = predP.y;
        const int minX = std::max(0, (int)std::floor((trackX - radius) * keysLeft.xMult));
        const int maxX = std::min(keysLeft.xGrids - 1, (int)std::ceil((trackX + radius) * keysLeft.xMult));
        const int minY = std::max(0, (int)std::floor((trackY - radius) * keysLeft.yMult));
        const int maxY = std::min(keysLeft.yGrids - 1, (int)std::ceil((trackY + radius) * keysLeft.yMult));
        if (minX >= keysLeft.xGrids || minY >= keysLeft.yGrids || maxX < 0 || maxY < 0) return;
        for (int row = minY; row <= maxY; row++)
            for (int col = minX; col <= maxX; col++)
                for (int idx : (right ? keysLeft.rkeyGrid : keysLeft.lkeyGrid)[row][col]) {
                    const cv::KeyPoint& k = right ? keysLeft.rightKeyPoints[idx] : keysLeft.keyPoints[idx];
                    if (k.octave > predictedScale + 1 || k.octave < predictedScale - 1) continue;
                    if (std::fabs(k.pt.x - trackX) < radius && std::fabs(k.pt.y - trackY) < radius) idxs.push_back(idx);
                }
    }
    void destributeRightKeys(const std::vector<cv::KeyPoint>& rightKeys, std::vector<std::vector<int>>& indexes) const {
        indexes.assign(imageHeight, std::vector<int>());
        for (size_t i = 0; i < rightKeys.size(); i++) {
            const float y = rightKeys[i].pt.y, r = 2.0f * feLeft->scalePyramid[rightKeys[i].octave];
            const int mx = (int)std::ceil(y + r), mn = (int)std::floor(y - r);
            for (int p = mn; p <= mx; p++)
                if (p >= 0 && p < imageHeight) indexes[p].push_back((int)i);
        }
    }

    // src/FeatureMatcher.cpp:66-252 (LocalMapper::findNewPoints, src/OptimizationBA.cpp:367).  KeyFrameT / the Eigen types are
    // the reference's own (include/KeyFrame.h); this template is only instantiated inside the reference tree.  The scale
    // prediction (:88-95) stays on the host in double precision exactly as the reference computes it.
    template <class KeyFrameT, class Vec4>
    int matchByProjectionRPredLBA(const KeyFrameT* lastKF, KeyFrameT* newKF,
                                  std::vector<std::vector<std::pair<KeyFrameT*, std::pair<int, int>>>>& matchedIdxs, const float rad,
                                  const std::vector<std::pair<cv::Point2f, cv::Point2f>>& predPoints,
                                  const std::vector<float>& maxDistsScale, std::vector<std::pair<Vec4, std::pair<int, int>>>& p4d) {
        const TrackedKeys& lastKeys = lastKF->keys;
        const TrackedKeys& keys = newKF->keys;
        const auto toFindPose = newKF->pose.pose;
        const int M = (int)p4d.size();
        std::vector<uint8_t> desc((size_t)M * 32);
        std::vector<float> pl(2 * (size_t)M), pr(2 * (size_t)M), ll(2 * (size_t)M), lr(2 * (size_t)M);
        std::vector<int32_t> ps(M);
        for (int i = 0; i < M; i++) {
            const auto& wPos = p4d[i].first;
            const std::pair<int, int>& kp = p4d[i].second;
            const cv::Mat* d;
            if (kp.first >= 0) {
                const auto* mp = lastKF->localMapPoints[kp.first];
                d = mp ? &mp->desc : nullptr;
                std::memcpy(&desc[(size_t)i * 32], d ? d->ptr(0) : lastKeys.Desc.ptr(kp.first), 32);
            } else std::memcpy(&desc[(size_t)i * 32], lastKeys.rightDesc.ptr(kp.second), 32);
            const double px = wPos(0) - toFindPose(0, 3), py = wPos(1) - toFindPose(1, 3), pz = wPos(2) - toFindPose(2, 3);
            const float dist = (float)std::sqrt(px * px + py * py + pz * pz);
            const float dif = maxDistsScale[i] / dist;
            // the reference's unqualified log(dif) on a float (src/FeatureMatcher.cpp:91) is ::log(double) under <cmath>: double arithmetic, as in gvs.h
            int predScale = (int)std::ceil(std::log((double)dif) / (double)lastKF->logScale);
            if (predScale < 0) predScale = 0; else if (predScale >= lastKF->nScaleLev) predScale = lastKF->nScaleLev - 1;
            ps[i] = predScale;
            pl[2 * i] = predPoints[i].first.x; pl[2 * i + 1] = predPoints[i].first.y;
            pr[2 * i] = predPoints[i].second.x; pr[2 * i + 1] = predPoints[i].second.y;
            const cv::Point2f kl = kp.first >= 0 ? lastKeys.keyPoints[kp.first].pt : lastKeys.rightKeyPoints[kp.second].pt;   // :231-235
            const cv::Point2f kr = kp.second >= 0 ? lastKeys.rightKeyPoints[kp.second].pt : lastKeys.keyPoints[kp.first].pt;  // :211-215
            ll[2 * i] = kl.x; ll[2 * i + 1] = kl.y; lr[2 * i] = kr.x; lr[2 * i + 1] = kr.y;
        }
        gvs_lba_frame f{(int32_t)keys.keyPoints.size(), (int32_t)keys.rightKeyPoints.size(),
                        reinterpret_cast<const gvs_keypoint*>(keys.keyPoints.data()), keys.Desc.data,
                        reinterpret_cast<const gvs_keypoint*>(keys.rightKeyPoints.data()), keys.rightDesc.data, keys.leftIdxs.data(),
                        keys.rightIdxs.data(), newKF->unMatchedF.data(), newKF->unMatchedFR.data(), feLeft->width(), feLeft->height()};
        gvs_lba_points p{M, desc.data(), pl.data(), pr.data(), ps.data(), ll.data(), lr.data()};
        std::vector<int32_t> out(2 * (size_t)M);
        int n = 0;
        gvsCheck(gvs_match_projection_lba(ensureHandle(), &f, &p, rad, out.data(), &n));
        for (int i = 0; i < M; i++)
            if (out[2 * i] != -2) matchedIdxs[i].emplace_back(std::make_pair(newKF, std::make_pair(out[2 * i], out[2 * i + 1])));
        return n;
    }

    // FeatureTracker::findOutliersR (src/FeatureTracker.cpp:582-649) with the reference's arguments (plus the camera members the
    // tracker holds: fx, fy, cx, cy, zedPtr->extrinsics).  Mat4 / MapPointT are the reference's own Eigen::Matrix4d / MapPoint;
    // the template is only instantiated inside the reference tree.  The two inverses (:584-585) stay on the host.
    template <class Mat4, class MapPointT>
    int findOutliersR(const Mat4& estimPose, const Mat4& extrinsics, double fx, double fy, double cx, double cy, float baseline,
                      std::vector<MapPointT*>& activeMapPoints, TrackedKeys& keysLeft, std::vector<std::pair<int, int>>& matchesIdxs,
                      const double thres, std::vector<bool>& MPsOutliers, int& nInliers) {
        const Mat4 toCameraR = (estimPose.inverse() * extrinsics).inverse();
        return findOutliersImpl(0, estimPose, &toCameraR, fx, fy, cx, cy, baseline * (float)closeNumber, activeMapPoints, &keysLeft,
                                matchesIdxs, thres, MPsOutliers, nInliers);
    }
    // FeatureTracker::findOutliersMono (src/FeatureTracker.cpp:651-684); returns nInliers like the reference
    template <class Mat4, class MapPointT>
    int findOutliersMono(const Mat4& estimPose, double fx, double fy, double cx, double cy, std::vector<MapPointT*>& activeMapPoints,
                         std::vector<std::pair<int, int>>& matchesIdxs, const double thres, std::vector<bool>& MPsOutliers, int& nInliers) {
        findOutliersImpl(1, estimPose, (const Mat4*)nullptr, fx, fy, cx, cy, 0.f, activeMapPoints, (TrackedKeys*)nullptr, matchesIdxs, thres,
                         MPsOutliers, nInliers);
        return nInliers;
    }

    // FeatureTracker::removeOutOfFrameMPs (src/FeatureTracker.cpp:910-940) with the reference's arguments (plus the tracker's camera
    // members).  Mat4 / MapPointT are the reference's Eigen::Matrix4d / MapPoint (fields wp via getWordPose4d(), maxScaleDist,
    // lastObsKF->logScale / nScaleLev, predL/R, scaleLevelL/R, inFrame/R, GetIsOutlier(), setActive()).
    template <class Mat4, class MapPointT>
    void removeOutOfFrameMPs(const Mat4& predNPose, const Mat4& extrinsics, double fx, double fy, double cx, double cy, int width, int height,
                             std::vector<MapPointT*>& activeMapPoints) {
        const Mat4 toRCamera = (predNPose * extrinsics).inverse(), toCamera = predNPose.inverse();
        std::vector<uint8_t> keep;
        predictImpl(0, toCamera, &toRCamera, fx, fy, cx, cy, width, height, activeMapPoints, keep, nullptr, nullptr, nullptr, nullptr);
        size_t j = 0;
        for (size_t i = 0; i < activeMapPoints.size(); i++) {
            MapPointT* mp = activeMapPoints[i];
            if (!mp || mp->GetIsOutlier()) continue;
            mp->setActive(keep[i] != 0);
            if (keep[i]) activeMapPoints[j++] = mp;
        }
        activeMapPoints.resize(j);
    }
    // FeatureTracker::PredictMPsPosition (src/FeatureTracker.cpp:969-1013); the caller holds map->mapMutex like the reference
    template <class Mat4, class MapPointT>
    void PredictMPsPosition(const Mat4& predNPose, const Mat4& extrinsics, double fx, double fy, double cx, double cy, int width, int height,
                            std::vector<MapPointT*>& activeMapPoints, std::vector<int>& matchedIdxsL, std::vector<int>& matchedIdxsR,
                            std::vector<std::pair<int, int>>& matchesIdxs, std::vector<bool>& MPsOutliers) {
        const Mat4 toRCamera = (predNPose * extrinsics).inverse(), toCamera = predNPose.inverse();
        std::vector<uint8_t> keep;
        predictImpl(2, toCamera, &toRCamera, fx, fy, cx, cy, width, height, activeMapPoints, keep, &matchesIdxs, &matchedIdxsL, &matchedIdxsR,
                    &MPsOutliers);
    }

    // MapPoint::calcDescriptor (src/Map.cpp:145-210) for one observation list: index of the chosen descriptor
    int medianDescriptor(const std::vector<cv::Mat>& vDescriptors) {
        const int N = (int)vDescriptors.size();
        if (N == 0) return -1;
        std::vector<uint8_t> d((size_t)N * 32);
        for (int i = 0; i < N; i++) std::memcpy(&d[(size_t)i * 32], vDescriptors[i].ptr(0), 32);
        const int32_t off[2] = {0, N};
        int32_t best = -1;
        gvsCheck(gvs_median_descriptor(ensureHandle(), d.data(), off, 1, &best));
        return best;
    }

    // src/FeatureMatcher.cpp:710-726
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(a.ptr(0));
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(b.ptr(0));
        int d = 0;
        for (int i = 0; i < 8; i++) d += __builtin_popcount(pa[i] ^ pb[i]);  // one 32-byte pair: not worth a launch
        return d;
    }
    // the device handle (created when both extractors have seen an image: the matcher reads their resident results)
    gvs_matcher* ensureHandle() {
        if (!h_) {
            if (!feLeft || !feLeft->handle() || (feRight && !feRight->handle()))
                throw std::logic_error("FeatureMatcher: extract an image with both extractors first");
            gvsCheck(gvs_matcher_create(&h_, feLeft->handle(), feRight ? feRight->handle() : nullptr, fx_, baseline_, imageHeight));
        }
        return h_;
    }
    gvs_matcher* handle() const { return h_; }

  private:
    template <class Mat4, class MapPointT>
    void predictImpl(int mode, const Mat4& toCamera, const Mat4* toRCamera, double fx, double fy, double cx, double cy, int width, int height,
                     std::vector<MapPointT*>& mps, std::vector<uint8_t>& keep, std::vector<std::pair<int, int>>* matchesIdxs,
                     std::vector<int>* matchedIdxsL, std::vector<int>* matchedIdxsR, std::vector<bool>* MPsOutliers) {
        const int M = (int)mps.size();
        keep.assign(M, 0);
        if (M == 0) return;
        double tc[16], tr[16];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) { tc[4 * r + c] = toCamera(r, c); tr[4 * r + c] = toRCamera ? (*toRCamera)(r, c) : 0.0; }
        std::vector<uint8_t> valid(M), inF(M), inFR(M), out(M);
        std::vector<double> wp(4 * (size_t)M, 0.0);
        std::vector<float> msd(M, 1.f), ls(M, 1.f), pl(2 * (size_t)M), pr(2 * (size_t)M);
        std::vector<int32_t> nsl(M, 1), sl(M), sr(M);
        for (int i = 0; i < M; i++) {
            const MapPointT* mp = mps[i];
            valid[i] = mp != nullptr && (mode == 2 || !mp->GetIsOutlier());
            if (MPsOutliers) out[i] = (*MPsOutliers)[i];
            if (!mp) continue;
            const auto w = mp->getWordPose4d();
            for (int c = 0; c < 4; c++) wp[4 * (size_t)i + c] = w(c);
            msd[i] = mp->maxScaleDist; ls[i] = mp->lastObsKF->logScale; nsl[i] = mp->lastObsKF->nScaleLev;
            pl[2 * i] = mp->predL.x; pl[2 * i + 1] = mp->predL.y; pr[2 * i] = mp->predR.x; pr[2 * i + 1] = mp->predR.y;
            sl[i] = mp->scaleLevelL; sr[i] = mp->scaleLevelR; inF[i] = mp->inFrame; inFR[i] = mp->inFrameR;
        }
        const bool two = mode != 1;
        gvsCheck(gvs_predict_mappoints(ensureHandle(), mode, tc, two ? tr : nullptr, fx, fy, cx, cy, width, height, M, valid.data(), wp.data(), msd.data(),
                                       ls.data(), nsl.data(), pl.data(), sl.data(), inF.data(), two ? pr.data() : nullptr,
                                       two ? sr.data() : nullptr, two ? inFR.data() : nullptr, keep.data(),
                                       matchesIdxs ? reinterpret_cast<int32_t*>(matchesIdxs->data()) : nullptr,
                                       matchedIdxsL ? matchedIdxsL->data() : nullptr, matchedIdxsL ? (int)matchedIdxsL->size() : 0,
                                       matchedIdxsR ? matchedIdxsR->data() : nullptr, matchedIdxsR ? (int)matchedIdxsR->size() : 0,
                                       MPsOutliers ? out.data() : nullptr));
        for (int i = 0; i < M; i++) {
            MapPointT* mp = mps[i];
            if (!valid[i]) continue;
            mp->predL = cv::Point2f(pl[2 * i], pl[2 * i + 1]); mp->scaleLevelL = sl[i]; mp->inFrame = inF[i] != 0;
            if (two) { mp->predR = cv::Point2f(pr[2 * i], pr[2 * i + 1]); mp->scaleLevelR = sr[i]; mp->inFrameR = inFR[i] != 0; }
        }
        if (MPsOutliers) for (int i = 0; i < M; i++) (*MPsOutliers)[i] = out[i] != 0;
    }
    template <class Mat4, class MapPointT>
    int findOutliersImpl(int mode, const Mat4& estimPose, const Mat4* toCameraR, double fx, double fy, double cx, double cy, float closeDist,
                         std::vector<MapPointT*>& activeMapPoints, TrackedKeys* keys, std::vector<std::pair<int, int>>& matchesIdxs,
                         const double thres, std::vector<bool>& MPsOutliers, int& nInliers) {
        const int M = (int)matchesIdxs.size();
        if (M == 0) return 0;
        double pose[16], poseR[16];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) { pose[4 * r + c] = estimPose(r, c); poseR[4 * r + c] = toCameraR ? (*toCameraR)(r, c) : 0.0; }
        std::vector<uint8_t> valid(M), inF(M), inFR(M), out(M);
        std::vector<double> wp(4 * (size_t)M, 0.0);
        for (int i = 0; i < M; i++) {
            const MapPointT* mp = activeMapPoints[i];
            valid[i] = mp != nullptr; out[i] = MPsOutliers[i];
            if (!mp) continue;
            inF[i] = mp->inFrame; inFR[i] = mp->inFrameR;
            const auto w = mp->getWordPose4d();
            for (int c = 0; c < 4; c++) wp[4 * (size_t)i + c] = w(c);
        }
        std::vector<uint8_t> close8;
        if (keys) close8.assign(keys->close.begin(), keys->close.end());
        int nIn = 0, nStereo = 0;
        gvsCheck(gvs_find_outliers(ensureHandle(), 0, mode, pose, toCameraR ? poseR : nullptr, fx, fy, cx, cy, thres, closeDist, M, valid.data(), wp.data(),
                                   inF.data(), mode == 0 ? inFR.data() : nullptr, reinterpret_cast<int32_t*>(matchesIdxs.data()), out.data(),
                                   keys ? keys->estimatedDepth.data() : nullptr, keys ? close8.data() : nullptr,
                                   keys ? keys->rightIdxs.data() : nullptr, keys ? keys->leftIdxs.data() : nullptr, &nIn, &nStereo));
        for (int i = 0; i < M; i++) MPsOutliers[i] = out[i] != 0;
        if (keys) for (size_t i = 0; i < close8.size(); i++) keys->close[i] = close8[i] != 0;
        nInliers += nIn;
        return nStereo;
    }
    static void fetch(FeatureExtractor& fe, std::vector<cv::KeyPoint>& k, cv::Mat& d) {
        int off[2] = {0, 0};
        gvsCheck(gvs_extract_offsets(fe.handle(), off));
        const int n = off[1];
        if (n <= 0) return;
        k.assign(n, cv::KeyPoint());
        d = cv::Mat(n, 32, GVS_CV_8U);
        gvsCheck(gvs_extract_download(fe.handle(), 0, 1, reinterpret_cast<gvs_keypoint*>(k.data()), d.data, n));
    }
    template <class MapPointT>
    int project(int mode, std::vector<MapPointT*>& mps, TrackedKeys&, std::vector<int>& mL, std::vector<int>* mR,
                std::vector<std::pair<int, int>>& matches, float rad) {
        const int M = (int)mps.size();
        std::vector<float> pl(2 * (size_t)M), pr(2 * (size_t)M);
        std::vector<int32_t> ll(M), lr(M);
        std::vector<uint8_t> inl(M), inr(M), desc((size_t)M * 32);
        for (int i = 0; i < M; i++) {
            const MapPointT* mp = mps[i];
            // the reference skips an already matched point before it touches mp (:259-263), so a null entry there is harmless:
            // its slot stays zero / out of frame (an unmatched null entry is a null dereference in the reference as well)
            if (!mp && (matches[i].first >= 0 || matches[i].second >= 0)) continue;
            pl[2 * i] = mp->predL.x; pl[2 * i + 1] = mp->predL.y; pr[2 * i] = mp->predR.x; pr[2 * i + 1] = mp->predR.y;
            ll[i] = mp->scaleLevelL; lr[i] = mp->scaleLevelR; inl[i] = mp->inFrame; inr[i] = mp->inFrameR;
            std::memcpy(&desc[(size_t)i * 32], mp->desc.ptr(0), 32);
        }
        gvs_mappoints s{M, desc.data(), pl.data(), pr.data(), ll.data(), lr.data(), inl.data(), inr.data()};
        static_assert(sizeof(std::pair<int, int>) == 8, "pair<int,int> must be two packed ints");
        int n = 0;
        gvsCheck(gvs_match_projection(ensureHandle(), 0, mode, &s, rad, mL.data(), mR ? mR->data() : nullptr,
                                      reinterpret_cast<int32_t*>(matches.data()), &n));
        return n;
    }
    gvs_matcher* h_ = nullptr;
    double fx_ = 0;
    float baseline_ = 0;
};

}  // namespace GTSAM_VIOSLAM
