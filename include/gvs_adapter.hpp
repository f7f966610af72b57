// Drop-in C++ adapter: the reference's FeatureExtractor / FeatureMatcher / TrackedKeys (same namespace, class,
// member and method names, same argument meaning) implemented on the C ABI of include/gvs.h.
//
// A maintainer of gtsam-vSLAM replaces include/FeatureExtractor.h + src/FeatureExtractor.cpp and
// include/FeatureMatcher.h + src/FeatureMatcher.cpp by this header and links libgvs.so;
// System, FeatureTracker, KeyFrame, Map and LocalMapper compile against it unchanged apart from the two
// constructor calls that now also pass the image size.
//
//   reference                                                       here
//   FeatureExtractor::FeatureExtractor   src/FeatureExtractor.cpp:620   FeatureExtractor(..., width, height)
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
    std::vector<int> scaledPatchSize, featurePerLevel;
    std::vector<float> scalePyramid, scaleInvPyramid, sigmaFactor, InvSigmaFactor;
    // imagePyramid stays on the device: only findStereoMatchesORB2R reads it (src/FeatureMatcher.cpp:606-623)

    FeatureExtractor(int _nfeatures, int _nLevels, float _imScale, int _edgeThreshold, int _patchSize, int _maxFastThreshold,
                     int _minFastThreshold, int width, int height, int device = 0)
        : nFeatures(_nfeatures), nLevels(_nLevels), imScale(_imScale), edgeThreshold(_edgeThreshold), patchSize(_patchSize),
          maxFastThreshold(_maxFastThreshold), minFastThreshold(_minFastThreshold), width_(width), height_(height) {
        gvs_extractor_params p{_nfeatures, _nLevels, _imScale, _edgeThreshold, _patchSize, _maxFastThreshold, _minFastThreshold};
        gvsCheck(gvs_extractor_create(&h_, &p, width, height, 1, device));
        scalePyramid.resize(nLevels); scaleInvPyramid.resize(nLevels); sigmaFactor.resize(nLevels); InvSigmaFactor.resize(nLevels);
        scaledPatchSize.resize(nLevels); featurePerLevel.resize(nLevels);
        gvsCheck(gvs_extractor_tables(h_, scalePyramid.data(), scaleInvPyramid.data(), sigmaFactor.data(), InvSigmaFactor.data(),
                                      scaledPatchSize.data(), featurePerLevel.data()));
    }
    FeatureExtractor(int width, int height) : FeatureExtractor(2000, 8, 1.2f, 19, 31, 20, 7, width, height) {}
    ~FeatureExtractor() { gvs_extractor_destroy(h_); }
    FeatureExtractor(const FeatureExtractor&) = delete;
    FeatureExtractor& operator=(const FeatureExtractor&) = delete;

    // src/FeatureExtractor.cpp:481-533.  Zero keypoints leaves both outputs untouched (:498-499).
    void extractKeysNew(cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {
        const int cap = 2 * nFeatures + 64 * (int)nLevels;
        std::vector<cv::KeyPoint> k(cap);
        cv::Mat d(cap, 32, cv::CV_8U);
        int n = 0;
        gvsCheck(gvs_extract(h_, image.data, image.cols, image.rows, (int)image.step, reinterpret_cast<gvs_keypoint*>(k.data()),
                             d.data, cap, &n));
        if (n <= 0) return;
        k.resize(n);
        keypoints = std::move(k);
        descriptors = cv::Mat(n, 32, cv::CV_8U);
        std::memcpy(descriptors.data, d.data, (size_t)n * 32);
    }
    gvs_extractor* handle() const { return h_; }
    int width() const { return width_; }
    int height() const { return height_; }

  private:
    gvs_extractor* h_ = nullptr;
    int width_, height_;
};

class FeatureMatcher {
  public:
    const int closeNumber{40};
    std::shared_ptr<FeatureExtractor> feLeft, feRight;
    const int imageHeight;

    // zedptr is reduced to the values the path reads: mCameraLeft->fx (double), mBaseline (float) (include/Camera.h:72,86)
    FeatureMatcher(double fx, float baseline, std::shared_ptr<FeatureExtractor> _feLeft, std::shared_ptr<FeatureExtractor> _feRight,
                   int _imageHeight = 360)
        : feLeft(std::move(_feLeft)), feRight(std::move(_feRight)), imageHeight(_imageHeight) {
        gvsCheck(gvs_matcher_create(&h_, feLeft->handle(), feRight ? feRight->handle() : nullptr, fx, baseline, _imageHeight));
    }
    ~FeatureMatcher() { gvs_matcher_destroy(h_); }
    FeatureMatcher(const FeatureMatcher&) = delete;
    FeatureMatcher& operator=(const FeatureMatcher&) = delete;

    // src/FeatureMatcher.cpp:528-708.  The images and rightDesc/rightKeys arguments are kept for signature parity: the
    // device already holds the pyramids, keys and descriptors of the two extractKeysNew calls that precede this call.
    void findStereoMatchesORB2R(const cv::Mat&, const cv::Mat&, const cv::Mat&, std::vector<cv::KeyPoint>& rightKeys, TrackedKeys& keysLeft) {
        gvsCheck(gvs_stereo_match_batch(h_, 1));
        const size_t nl = keysLeft.keyPoints.size(), nr = rightKeys.size();
        keysLeft.rightIdxs.assign(nl, -1); keysLeft.leftIdxs.assign(nr, -1); keysLeft.estimatedDepth.assign(nl, -1.0f);
        std::vector<uint8_t> cl(nl + 1);
        gvsCheck(gvs_stereo_download(h_, 0, 1, keysLeft.rightIdxs.data(), keysLeft.leftIdxs.data(), keysLeft.estimatedDepth.data(), cl.data()));
        keysLeft.close.assign(nl, false);
        for (size_t i = 0; i < nl; i++) keysLeft.close[i] = cl[i] != 0;
    }

    // FeatureTracker::assignKeysToGrids for both sides (src/FeatureTracker.cpp:28-54, :67-68)
    void assignKeysToGrids(TrackedKeys& keysLeft, int width, int height) {
        gvsCheck(gvs_assign_grids_batch(h_, 1, width, height));
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
        gvsCheck(gvs_extract_batch_host(feLeft->handle(), leftIm.data, leftIm.step * leftIm.rows, (int)leftIm.step, 1));
        gvsCheck(gvs_extract_batch_host(feRight->handle(), rightIm.data, rightIm.step * rightIm.rows, (int)rightIm.step, 1));
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
    // src/FeatureMatcher.cpp:458-526: matches[i].first receives the index matched by lastKeys.keyPoints[i] (or stays -1); the
    // caller appends (actKeyFrame, idx) to keyframeIdxMatchs[i] as :522 does.  actKeys must be the last extracted frame.
    int matchByRadius(TrackedKeys& lastKeys, TrackedKeys& actKeys, std::vector<int>& matchedIdxsL, std::vector<std::pair<int, int>>& matches,
                      const float rad) {
        const int M = (int)lastKeys.keyPoints.size();
        std::vector<float> pred(2 * (size_t)M);
        std::vector<int32_t> lvl(M);
        std::vector<uint8_t> desc((size_t)M * 32);
        for (int i = 0; i < M; i++) {
            pred[2 * i] = lastKeys.keyPoints[i].pt.x; pred[2 * i + 1] = lastKeys.keyPoints[i].pt.y; lvl[i] = lastKeys.keyPoints[i].octave;
            std::memcpy(&desc[(size_t)i * 32], lastKeys.Desc.ptr(i), 32);
        }
        gvs_mappoints mp{M, desc.data(), pred.data(), nullptr, lvl.data(), nullptr, nullptr, nullptr};
        matches.assign(M, std::make_pair(-1, -1));
        int n = 0;
        (void)actKeys;
        gvsCheck(gvs_match_projection(h_, 0, GVS_MATCH_RADIUS, &mp, rad, matchedIdxsL.data(), nullptr,
                                      reinterpret_cast<int32_t*>(matches.data()), &n));
        return n;
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
        gvsCheck(gvs_match_projection_lba(h_, &f, &p, rad, out.data(), &n));
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
        gvsCheck(gvs_median_descriptor(h_, d.data(), off, 1, &best));
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
        gvsCheck(gvs_predict_mappoints(h_, mode, tc, two ? tr : nullptr, fx, fy, cx, cy, width, height, M, valid.data(), wp.data(), msd.data(),
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
        gvsCheck(gvs_find_outliers(h_, 0, mode, pose, toCameraR ? poseR : nullptr, fx, fy, cx, cy, thres, closeDist, M, valid.data(), wp.data(),
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
        d = cv::Mat(n, 32, cv::CV_8U);
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
            pl[2 * i] = mp->predL.x; pl[2 * i + 1] = mp->predL.y; pr[2 * i] = mp->predR.x; pr[2 * i + 1] = mp->predR.y;
            ll[i] = mp->scaleLevelL; lr[i] = mp->scaleLevelR; inl[i] = mp->inFrame; inr[i] = mp->inFrameR;
            std::memcpy(&desc[(size_t)i * 32], mp->desc.ptr(0), 32);
        }
        gvs_mappoints s{M, desc.data(), pl.data(), pr.data(), ll.data(), lr.data(), inl.data(), inr.data()};
        static_assert(sizeof(std::pair<int, int>) == 8, "pair<int,int> must be two packed ints");
        int n = 0;
        gvsCheck(gvs_match_projection(h_, 0, mode, &s, rad, mL.data(), mR ? mR->data() : nullptr,
                                      reinterpret_cast<int32_t*>(matches.data()), &n));
        return n;
    }
    gvs_matcher* h_ = nullptr;
};

}  // namespace GTSAM_VIOSLAM
