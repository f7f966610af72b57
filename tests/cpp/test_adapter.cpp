// Exercises include/gvs_adapter.hpp (the C++ drop-in with the reference's class names) end to end on one stereo pair read
// from two raw u8 files; prints counts and FNV-1a checksums that tests/test_gpu_adapter.py compares with the ctypes path.
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "../../include/gvs_adapter.hpp"

using namespace GTSAM_VIOSLAM;

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
// stand-ins for the reference's Eigen::Vector4d / Eigen::Matrix4d (pure translations are enough here) and MapPoint
struct Vec4 { double v[4]; double operator()(int i) const { return v[i]; } };
struct Mat4 {
    double t[3] = {0, 0, 0};
    double operator()(int r, int c) const { return r == c ? 1.0 : (c == 3 && r < 3 ? t[r] : 0.0); }
    Mat4 inverse() const { Mat4 m; for (int i = 0; i < 3; i++) m.t[i] = -t[i]; return m; }
    Mat4 operator*(const Mat4& o) const { Mat4 m; for (int i = 0; i < 3; i++) m.t[i] = t[i] + o.t[i]; return m; }
};
struct KeyFrameStub { float logScale; int nScaleLev; };
// stand-ins for the reference's StereoCamera / Camera (include/Camera.h:72,86): the members the matcher constructor reads
struct CameraStub { double fx = 0; };
struct StereoCameraStub { std::shared_ptr<CameraStub> mCameraLeft; float mBaseline = 0; };
struct MapPoint {
    cv::Mat desc; cv::Point2f predL, predR; int scaleLevelL = 0, scaleLevelR = 0; bool inFrame = true, inFrameR = true;
    Vec4 wp{{0, 0, 0, 1}};
    float maxScaleDist = 30.f; const KeyFrameStub* lastObsKF = nullptr; bool isOutlier = false, isActive = true;
    Vec4 getWordPose4d() const { return wp; }
    bool GetIsOutlier() const { return isOutlier; }
    void setActive(bool a) { isActive = a; }
};

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: left.raw right.raw w h nFeatures nLevels\n"); return 2; }
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), nF = std::atoi(argv[5]), nL = std::atoi(argv[6]);
    cv::Mat L(h, w, cv::CV_8U), R(h, w, cv::CV_8U);
    FILE* f = std::fopen(argv[1], "rb"); if (!f || std::fread(L.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3; std::fclose(f);
    f = std::fopen(argv[2], "rb"); if (!f || std::fread(R.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3; std::fclose(f);
    try {
        // exactly the reference's argument lists: include/FeatureExtractor.h:80 (src/System.cpp:56-57), include/FeatureMatcher.h:40
        // (src/FeatureTracker.cpp:23); the default constructor of src/System.cpp:31 is exercised below
        auto feL = std::make_shared<FeatureExtractor>(nF, nL, 1.2f, 19, 31, 20, 7);
        auto feR = std::make_shared<FeatureExtractor>(nF, nL, 1.2f, 19, 31, 20, 7);
        auto zed = std::make_shared<StereoCameraStub>();
        zed->mCameraLeft = std::make_shared<CameraStub>(); zed->mCameraLeft->fx = 200.0; zed->mBaseline = 0.1f;
        if ((int)feL->scalePyramid.size() != nL || feL->featurePerLevel.empty()) return 4;   // tables are public before the first image
        FeatureMatcher fm(zed, feL, feR, h);
        TrackedKeys k;
        fm.extractORBAndStereoMatch(L, R, k);
        int matches = 0;
        for (int v : k.rightIdxs) matches += v >= 0;
        size_t gridItems = 0;
        for (auto& row : k.lkeyGrid) for (auto& c : row) gridItems += c.size();
        std::printf("nl %zu nr %zu matches %d grid %zu yGrids %d\n", k.keyPoints.size(), k.rightKeyPoints.size(), matches, gridItems, k.yGrids);
        std::printf("kp %016llx desc %016llx ridx %016llx depth %016llx\n",
                    (unsigned long long)fnv(k.keyPoints.data(), k.keyPoints.size() * sizeof(cv::KeyPoint)),
                    (unsigned long long)fnv(k.Desc.data, (size_t)k.Desc.rows * 32),
                    (unsigned long long)fnv(k.rightIdxs.data(), k.rightIdxs.size() * 4),
                    (unsigned long long)fnv(k.estimatedDepth.data(), k.estimatedDepth.size() * 4));
        // the single-image class interface
        std::vector<cv::KeyPoint> kp; cv::Mat d;
        feL->extractKeysNew(L, kp, d);
        std::printf("single %zu %016llx\n", kp.size(), (unsigned long long)fnv(kp.data(), kp.size() * sizeof(cv::KeyPoint)));
        // temporal matcher through the class interface: map points = this frame's left keys
        fm.extractORBAndStereoMatch(L, R, k);
        std::vector<MapPoint> store(k.keyPoints.size());
        std::vector<MapPoint*> mps;
        for (size_t i = 0; i < store.size(); i++) {
            store[i].desc = k.Desc.row((int)i).clone(); store[i].predL = k.keyPoints[i].pt; store[i].predR = k.keyPoints[i].pt;
            store[i].scaleLevelL = store[i].scaleLevelR = k.keyPoints[i].octave; store[i].inFrameR = false;
            mps.push_back(&store[i]);
        }
        std::vector<int> mL(k.keyPoints.size(), -1), mR(k.rightKeyPoints.size(), -1);
        std::vector<std::pair<int, int>> pairs(mps.size(), {-1, -1});
        int n = fm.matchByProjectionRPred(mps, k, mL, mR, pairs, 4.0f);
        std::printf("proj %d of %zu\n", n, mps.size());
        // FeatureTracker::findOutliersR through the adapter: map point i = left key i back-projected with its stereo depth (5 m
        // without one), matched to (i, rightIdxs[i]); every 5th stereo partner is replaced by its neighbour and must be rejected
        const double fx = 200.0, cx = w / 2.0, cy = h / 2.0;
        Mat4 estimPose, extrinsics; extrinsics.t[0] = 0.1f;
        for (size_t i = 0; i < store.size(); i++) {
            const double z = k.estimatedDepth[i] > 0 ? (double)k.estimatedDepth[i] : 5.0;
            store[i].wp = Vec4{{((double)k.keyPoints[i].pt.x - cx) / fx * z, ((double)k.keyPoints[i].pt.y - cy) / fx * z, z, 1.0}};
            store[i].inFrameR = true;
            int second = k.rightIdxs[i];
            if (second >= 0 && i % 5 == 0) second = (second + 1) % (int)k.rightKeyPoints.size();
            pairs[i] = {(int)i, second};
        }
        std::vector<bool> outl(mps.size(), false);
        int nIn = 0;
        const int nStereo = fm.findOutliersR(estimPose, extrinsics, fx, fx, cx, cy, 0.1f, mps, k, pairs, 7.815, outl, nIn);
        int nOut = 0, nCleared = 0;
        for (size_t i = 0; i < outl.size(); i++) { nOut += outl[i]; nCleared += pairs[i].second < 0; }
        std::printf("outliers in %d stereo %d out %d ridx %016llx close %d\n", nIn, nStereo, nOut,
                    (unsigned long long)fnv(k.rightIdxs.data(), k.rightIdxs.size() * 4), (int)std::count(k.close.begin(), k.close.end(), true));
        // FeatureTracker::PredictMPsPosition and removeOutOfFrameMPs through the adapter: every 7th point is moved far out of the
        // frame, every 11th is flagged as an outlier (PredictMPsPosition) / is an outlier map point (removeOutOfFrameMPs)
        const KeyFrameStub kf{(float)std::log((double)1.2f), nL};   // KeyFrame::logScale
        for (size_t i = 0; i < store.size(); i++) {
            store[i].lastObsKF = &kf;
            if (i % 7 == 0) store[i].wp.v[0] += 1000.0;
            pairs[i] = {(int)i, k.rightIdxs[i]};
            outl[i] = i % 11 == 0;
        }
        std::fill(mL.begin(), mL.end(), -1); std::fill(mR.begin(), mR.end(), -1);
        for (size_t i = 0; i < pairs.size(); i++) { mL[pairs[i].first] = (int)i; if (pairs[i].second >= 0) mR[pairs[i].second] = (int)i; }
        fm.PredictMPsPosition(estimPose, extrinsics, fx, fx, cx, cy, w, h, mps, mL, mR, pairs, outl);
        int inL = 0, inR = 0, scaleSum = 0;
        for (auto& mp : store) { inL += mp.inFrame; inR += mp.inFrameR; scaleSum += mp.scaleLevelL + 3 * mp.scaleLevelR; }
        std::printf("predict inL %d inR %d scale %d pairs %016llx mL %016llx mR %016llx\n", inL, inR, scaleSum,
                    (unsigned long long)fnv(pairs.data(), pairs.size() * 8), (unsigned long long)fnv(mL.data(), mL.size() * 4),
                    (unsigned long long)fnv(mR.data(), mR.size() * 4));
        for (size_t i = 0; i < store.size(); i++) store[i].isOutlier = i % 11 == 0;
        std::vector<MapPoint*> act = mps;
        fm.removeOutOfFrameMPs(estimPose, extrinsics, fx, fx, cx, cy, w, h, act);
        int active = 0;
        for (auto& mp : store) active += mp.isActive;
        std::printf("remove kept %zu active %d\n", act.size(), active);
        // matchByRadius with the reference's seven arguments (include/FeatureMatcher.h:63, call site src/FeatureTracker.cpp:1518):
        // the right keys of this frame against its left keys; matchedIdxsL is in-out, matchesIdxs stays untouched
        fm.extractORBAndStereoMatch(L, R, k);
        TrackedKeys last;
        last.keyPoints = k.rightKeyPoints; last.Desc = k.rightDesc.clone();
        std::vector<int> mRad(k.keyPoints.size(), -1);
        std::vector<std::pair<int, int>> untouched(last.keyPoints.size(), {-7, -7});
        std::vector<std::vector<std::pair<KeyFrameStub*, int>>> kfMatches(last.keyPoints.size());
        KeyFrameStub actKF{0.f, nL};
        const int nRad = fm.matchByRadius(last, k, mRad, untouched, 120.0f, kfMatches, &actKF);
        int appended = 0, same = 1;
        std::vector<int> radIdx(last.keyPoints.size(), -1);
        for (size_t i = 0; i < kfMatches.size(); i++) { appended += (int)kfMatches[i].size(); if (!kfMatches[i].empty()) { radIdx[i] = kfMatches[i].back().second; same &= kfMatches[i].back().first == &actKF; } }
        for (auto& p : untouched) same &= p.first == -7 && p.second == -7;
        std::printf("radius %d appended %d ok %d idx %016llx mL %016llx\n", nRad, appended, same, (unsigned long long)fnv(radIdx.data(), radIdx.size() * 4),
                    (unsigned long long)fnv(mRad.data(), mRad.size() * 4));
        // the default constructor (src/System.cpp:31: mono, 2000 features, 8 levels), imagePyramid mirror, the public ssc helper
        auto feM = std::make_shared<FeatureExtractor>();
        feM->mirrorImagePyramid = true;
        std::vector<cv::KeyPoint> km; cv::Mat dm;
        feM->extractKeysNew(L, km, dm);
        uint64_t ph = 1469598103934665603ull;
        for (auto& m : feM->imagePyramid) ph = fnv(m.data, (size_t)m.rows * m.cols, ph);
        std::printf("default %zu levels %zu pyr %016llx l1 %dx%d\n", km.size(), feM->imagePyramid.size(), (unsigned long long)ph,
                    feM->imagePyramid[1].cols, feM->imagePyramid[1].rows);
        FeatureMatcher fmMono(zed, feM, feM);   // src/System.cpp:32
        std::vector<cv::KeyPoint> grid;
        for (int y = 20; y < h - 20; y += 5) for (int x = 20; x < w - 20; x += 5) { cv::KeyPoint q; q.pt = cv::Point2f((float)x, (float)y); q.response = (float)(1 + (x * 7 + y * 13) % 200); grid.push_back(q); }
        std::vector<cv::KeyPoint> keptS = feM->ssc(grid, feM->featurePerLevel[0], 0.1f, w, h);
        std::printf("ssc %zu of %zu quota %d %016llx\n", keptS.size(), grid.size(), feM->featurePerLevel[0], (unsigned long long)fnv(keptS.data(), keptS.size() * sizeof(cv::KeyPoint)));
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
