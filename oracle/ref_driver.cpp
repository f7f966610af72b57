// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's OWN FeatureExtractor and FeatureMatcher (compiled unmodified from
// /root/reference/src/FeatureExtractor.cpp and src/FeatureMatcher.cpp against oracle/shim, see oracle/Makefile target
// `ref`), so that python tests can run them next to the restatement (oracle/cpu_ref.cpp) and generate golden vectors.
// Only objects and members the reference declares public are touched.  Also compiled into the library: the reference's
// src/Map.cpp (unmodified, whole) and the functions of src/FeatureTracker.cpp listed in oracle/slice_reference.py (bodies copied by
// line range into a generated file at build time; the whole file needs GTSAM).  KeyFrame.cpp / Camera.cpp are not on the path
// and not compiled; the three constructors of their classes that this driver needs in order to hand objects to the reference
// code are defined at the bottom of this file as plain member initialisations.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "FeatureTracker.h"   // reference header (with oracle/shim/gtsam/navigation/ImuBias.h standing in for GTSAM): pulls
                              // FeatureMatcher.h, Settings.h, Camera.h, Map.h, KeyFrame.h, FeatureExtractor.h, Conversions.h
#include "ref_hooks.h"

// default primitives: the cv2-pinned restatements exported by liboracle.so (oracle/cpu_ref.cpp)
extern "C" {
void ora_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void ora_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst);
float ora_fast_atan2(float y, float x);
void ora_sort_idx_desc(const float* key, int n, int* idx);
int ora_fast_detect(const uint8_t* img, int w, int h, int threshold, int* xs, int* ys, int* scores, int cap);
int ora_assign_grid(const void* kps, int n, int width, int height, int* start, int* items, float* xMult, float* yMult);
int ora_temporal_mappoints(const void* kl, int NL, const float* depth, const double* posePrev, const double* toCamera,
                           const double* toCameraR, double fx, double fy, double cx, double cy, int width, int height,
                           const float* scalePyramid, int nLevels, float logScale, int* mpKey, float* predL, float* predR, int* lvlL,
                           int* lvlR, double* wPos, float* maxScaleDist);
}

namespace {
const gvs_ref_hooks kDefaults = {ora_resize_linear, ora_fast_detect, ora_gaussian_blur7, ora_sort_idx_desc, ora_fast_atan2};
gvs_ref_hooks gHooks = kDefaults;
using namespace GTSAM_VIOSLAM;
typedef std::shared_ptr<FeatureExtractor> ExtractorPtr;
inline FeatureExtractor* FE(void* h) { return ((ExtractorPtr*)h)->get(); }

struct Matcher {
    std::shared_ptr<StereoCamera> zed;
    std::shared_ptr<FeatureMatcher> fm;
};

std::vector<cv::KeyPoint> keyVector(const void* kps, int n) {
    const cv::KeyPoint* k = (const cv::KeyPoint*)kps;
    return std::vector<cv::KeyPoint>(k, k + n);
}
cv::Mat descMat(const uint8_t* d, int n) {
    cv::Mat m(n, 32, CV_8U);
    if (n) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}
// CSR (start[yGrids*xGrids+1], items) -> TrackedKeys grid (FeatureTracker::assignKeysToGrids' output layout)
std::vector<std::vector<std::vector<int>>> gridFromCsr(const int* start, const int* items, int yGrids, int xGrids) {
    std::vector<std::vector<std::vector<int>>> g(yGrids, std::vector<std::vector<int>>(xGrids));
    for (int r = 0; r < yGrids; ++r)
        for (int c = 0; c < xGrids; ++c) g[r][c].assign(items + start[r * xGrids + c], items + start[r * xGrids + c + 1]);
    return g;
}
// TrackedKeys of one frame from flat arrays (+ CSR grids)
void fillKeys(TrackedKeys& k, const void* kpL, const uint8_t* descL, int NL, const void* kpR, const uint8_t* descR, int NR,
              const int* leftIdxs, const int* rightIdxs, const int* gridLStart, const int* gridLItems, const int* gridRStart,
              const int* gridRItems, int yGrids, int xGrids, float xMult, float yMult) {
    k.keyPoints = keyVector(kpL, NL); k.Desc = descMat(descL, NL);
    if (kpR) { k.rightKeyPoints = keyVector(kpR, NR); k.rightDesc = descMat(descR, NR); }
    if (leftIdxs) k.leftIdxs.assign(leftIdxs, leftIdxs + NR);
    if (rightIdxs) k.rightIdxs.assign(rightIdxs, rightIdxs + NL);
    k.xGrids = xGrids; k.yGrids = yGrids; k.xMult = xMult; k.yMult = yMult;
    k.lkeyGrid = gridFromCsr(gridLStart, gridLItems, yGrids, xGrids);
    if (gridRStart) k.rkeyGrid = gridFromCsr(gridRStart, gridRItems, yGrids, xGrids);
}
}  // namespace

extern "C" {

const gvs_ref_hooks* gvs_ref_get_hooks(void) { return &gHooks; }

void gvs_ref_set_hooks(const gvs_ref_hooks* h) {
    gHooks = kDefaults;
    if (!h) return;
    if (h->resize_linear) gHooks.resize_linear = h->resize_linear;
    if (h->fast_nms) gHooks.fast_nms = h->fast_nms;
    if (h->gaussian_blur7) gHooks.gaussian_blur7 = h->gaussian_blur7;
    if (h->sort_idx_desc) gHooks.sort_idx_desc = h->sort_idx_desc;
    if (h->fast_atan2) gHooks.fast_atan2 = h->fast_atan2;
}

void* ref_extractor_create(int nFeatures, int nLevels, float imScale, int edgeThreshold, int patchSize, int maxFast,
                           int minFast) {
    return new ExtractorPtr(std::make_shared<FeatureExtractor>(nFeatures, nLevels, imScale, edgeThreshold, patchSize, maxFast, minFast));
}
void ref_extractor_destroy(void* h) { delete (ExtractorPtr*)h; }

// public tables (include/FeatureExtractor.h:71-77)
void ref_extractor_tables(void* h, float* scale, float* invScale, float* sigma, float* invSigma, int* patchSz, int* quota) {
    FeatureExtractor* fe = FE(h);
    for (size_t i = 0; i < fe->nLevels; ++i) {
        scale[i] = fe->scalePyramid[i]; invScale[i] = fe->scaleInvPyramid[i]; sigma[i] = fe->sigmaFactor[i];
        invSigma[i] = fe->InvSigmaFactor[i]; patchSz[i] = fe->scaledPatchSize[i]; quota[i] = fe->featurePerLevel[i];
    }
}

// FeatureExtractor::extractKeysNew; kps = cv::KeyPoint[cap] (28 bytes each), desc = cap x 32 bytes.  Returns n or -1.
int ref_extract(void* h, const uint8_t* img, int w, int hgt, void* kps, uint8_t* desc, int cap) {
    FeatureExtractor* fe = FE(h);
    cv::Mat image(hgt, w, CV_8UC1, (void*)img);
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    fe->extractKeysNew(image, keys, d);
    if ((int)keys.size() > cap) return -1;
    if (!keys.empty()) std::memcpy(kps, keys.data(), keys.size() * sizeof(cv::KeyPoint));
    for (size_t i = 0; i < keys.size(); ++i) std::memcpy(desc + i * 32, d.ptr((int)i), 32);
    return (int)keys.size();
}

// pyramid level of the last extract (FeatureExtractor::imagePyramid, public)
void ref_level_size(void* h, int level, int* w, int* hgt) {
    FeatureExtractor* fe = FE(h);
    *w = fe->imagePyramid[level].cols; *hgt = fe->imagePyramid[level].rows;
}
void ref_level_image(void* h, int level, uint8_t* out) {
    const cv::Mat& m = FE(h)->imagePyramid[level];
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
}

// FeatureExtractor::ssc on explicit points (public member): returns the kept count, out = indices into the input
int ref_ssc(void* h, int numRetPoints, int cols, int rows, int n, const int* x, const int* y, const float* resp, int* out) {
    FeatureExtractor* fe = FE(h);
    std::vector<cv::KeyPoint> in(n);
    for (int i = 0; i < n; ++i) { in[i] = cv::KeyPoint((float)x[i], (float)y[i], 7.f, -1.f, resp[i]); in[i].class_id = i; }
    std::vector<cv::KeyPoint> kept = fe->ssc(in, numRetPoints, 0.1f, cols, rows);
    for (size_t i = 0; i < kept.size(); ++i) out[i] = kept[i].class_id;
    return (int)kept.size();
}

// ---- FeatureMatcher (reference include/FeatureMatcher.h:21-60) ------------------------------------------------------

void* ref_matcher_create(void* feL, void* feR, double fx, float baseline, int imageHeight) {
    Matcher* m = new Matcher;
    m->zed = std::make_shared<StereoCamera>();
    m->zed->mBaseline = baseline;
    m->zed->mCameraLeft = std::make_shared<Camera>(nullptr, std::string());
    m->zed->mCameraLeft->fx = fx;
    m->fm = std::make_shared<FeatureMatcher>(m->zed, *(ExtractorPtr*)feL, *(ExtractorPtr*)feR, imageHeight);
    return m;
}
void ref_matcher_destroy(void* h) { delete (Matcher*)h; }

// FeatureMatcher::findStereoMatchesORB2R on the pyramids of the last extractKeysNew of the two extractors
void ref_stereo_match(void* h, const void* kl, const uint8_t* dl, int NL, const void* kr, const uint8_t* dr, int NR,
                      int* rightIdxs, int* leftIdxs, float* depth, uint8_t* close) {
    Matcher* m = (Matcher*)h;
    TrackedKeys keys;
    keys.keyPoints = keyVector(kl, NL); keys.Desc = descMat(dl, NL);
    std::vector<cv::KeyPoint> rightKeys = keyVector(kr, NR);
    cv::Mat rightDesc = descMat(dr, NR), none;
    m->fm->findStereoMatchesORB2R(none, none, rightDesc, rightKeys, keys);
    for (int i = 0; i < NL; ++i) { rightIdxs[i] = keys.rightIdxs[i]; depth[i] = keys.estimatedDepth[i]; close[i] = keys.close[i]; }
    for (int i = 0; i < NR; ++i) leftIdxs[i] = keys.leftIdxs[i];
}

// mode 0/1/2 = matchByProjectionRPred / matchByProjectionMono / matchByRadius; same flattened arguments as
// ora_match_projection (oracle/cpu_ref.cpp) except that the key grids are passed in (FeatureTracker.cpp is not compiled).
int ref_match_projection(void* h, int mode, int M, const uint8_t* mpDesc, const float* predL, const float* predR,
                         const int* lvlL, const int* lvlR, const uint8_t* inFrame, const uint8_t* inFrameR, const void* kpL,
                         const uint8_t* descL, int NL, const void* kpR, const uint8_t* descR, int NR, const int* leftIdxs,
                         const int* rightIdxs, const int* gridLStart, const int* gridLItems, const int* gridRStart,
                         const int* gridRItems, int yGrids, int xGrids, float xMult, float yMult, float rad,
                         int* matchedIdxsL, int* matchedIdxsR, int* matchesIdxs /* M x 2 */) {
    Matcher* m = (Matcher*)h;
    TrackedKeys keys;
    fillKeys(keys, kpL, descL, NL, kpR, descR, NR, leftIdxs, rightIdxs, gridLStart, gridLItems, gridRStart, gridRItems, yGrids,
             xGrids, xMult, yMult);
    std::vector<int> mL(matchedIdxsL, matchedIdxsL + NL), mR;
    if (matchedIdxsR) mR.assign(matchedIdxsR, matchedIdxsR + NR);
    std::vector<std::pair<int, int>> pairs(M);
    for (int i = 0; i < M; ++i) pairs[i] = {matchesIdxs[2 * i], matchesIdxs[2 * i + 1]};
    int n = 0;
    if (mode == 2) {
        // "map points" are the previous frame's keys: predL = pt, lvlL = octave, mpDesc = Desc rows
        TrackedKeys last;
        last.keyPoints.resize(M);
        for (int i = 0; i < M; ++i) { last.keyPoints[i].pt = cv::Point2f(predL[2 * i], predL[2 * i + 1]); last.keyPoints[i].octave = lvlL[i]; }
        last.Desc = descMat(mpDesc, M);
        std::vector<std::vector<std::pair<KeyFrame*, int>>> kfMatches(M);
        n = m->fm->matchByRadius(last, keys, mL, pairs, rad, kfMatches, nullptr);
        for (int i = 0; i < M; ++i)
            if (!kfMatches[i].empty()) pairs[i].first = kfMatches[i].back().second;
    } else {
        std::vector<MapPoint*> mps(M);
        cv::KeyPoint obs;
        for (int i = 0; i < M; ++i) {
            MapPoint* mp = new MapPoint(Eigen::Vector4d(0, 0, 0, 1), descMat(mpDesc + (size_t)i * 32, 1), obs, 0, (unsigned long)i);
            mp->predL = cv::Point2f(predL[2 * i], predL[2 * i + 1]);
            if (predR) mp->predR = cv::Point2f(predR[2 * i], predR[2 * i + 1]);
            mp->scaleLevelL = lvlL[i];
            if (lvlR) mp->scaleLevelR = lvlR[i];
            if (inFrame) mp->inFrame = inFrame[i] != 0;
            if (inFrameR) mp->inFrameR = inFrameR[i] != 0;
            mps[i] = mp;
        }
        n = mode == 0 ? m->fm->matchByProjectionRPred(mps, keys, mL, mR, pairs, rad)
                      : m->fm->matchByProjectionMono(mps, keys, mL, pairs, rad);
        for (MapPoint* mp : mps) delete mp;
    }
    std::memcpy(matchedIdxsL, mL.data(), (size_t)NL * sizeof(int));
    if (matchedIdxsR) std::memcpy(matchedIdxsR, mR.data(), (size_t)NR * sizeof(int));
    for (int i = 0; i < M; ++i) { matchesIdxs[2 * i] = pairs[i].first; matchesIdxs[2 * i + 1] = pairs[i].second; }
    return n;
}

// FeatureMatcher::matchByProjectionRPredLBA with the flattened inputs of ora_match_lba.  The KeyFrame inputs are rebuilt
// so that the reference code sees exactly these values: map point i is key pair (i, i) of a synthetic last key frame whose
// left / right key i sits at lastL[i] / lastR[i] with descriptor mpDesc[i]; the new key frame's pose is the identity, every
// world point is at distance 1 and maxDistsScale[i] = 1.2^(predScale[i] - 0.5), which makes
// ceil(log(maxDist / dist) / logScale) = predScale[i] (src/FeatureMatcher.cpp:88-95).
int ref_match_lba(void* h, int M, const uint8_t* mpDesc, const float* predL, const float* predR, const int* predScale,
                  const float* lastL, const float* lastR, const void* kpL, const uint8_t* descL, int NL, const void* kpR,
                  const uint8_t* descR, int NR, const int* leftIdxs, const int* rightIdxs, const int* unMatchedF,
                  const int* unMatchedFR, const int* gridLStart, const int* gridLItems, const int* gridRStart,
                  const int* gridRItems, int yGrids, int xGrids, float xMult, float yMult, float rad, int* out /* M x 2 */) {
    Matcher* m = (Matcher*)h;
    cv::Mat noIm;
    KeyFrame lastKF(Eigen::Matrix4d::Identity(), noIm, noIm, 0, 0), newKF(Eigen::Matrix4d::Identity(), noIm, noIm, 1, 1);
    const int nLev = (int)m->fm->feLeft->nLevels;
    lastKF.logScale = std::log(m->fm->feLeft->imScale);
    lastKF.nScaleLev = nLev;
    lastKF.keys.keyPoints.resize(M); lastKF.keys.rightKeyPoints.resize(M);
    for (int i = 0; i < M; ++i) {
        lastKF.keys.keyPoints[i].pt = cv::Point2f(lastL[2 * i], lastL[2 * i + 1]);
        lastKF.keys.rightKeyPoints[i].pt = cv::Point2f(lastR[2 * i], lastR[2 * i + 1]);
    }
    lastKF.keys.Desc = descMat(mpDesc, M);
    lastKF.keys.rightDesc = descMat(mpDesc, M);
    lastKF.localMapPoints.assign(M, nullptr);
    fillKeys(newKF.keys, kpL, descL, NL, kpR, descR, NR, leftIdxs, rightIdxs, gridLStart, gridLItems, gridRStart, gridRItems,
             yGrids, xGrids, xMult, yMult);
    newKF.unMatchedF.assign(unMatchedF, unMatchedF + NL);
    newKF.unMatchedFR.assign(unMatchedFR, unMatchedFR + NR);
    std::vector<std::vector<std::pair<KeyFrame*, std::pair<int, int>>>> matched(M);
    std::vector<std::pair<cv::Point2f, cv::Point2f>> pred(M);
    std::vector<float> maxDists(M);
    std::vector<std::pair<Eigen::Vector4d, std::pair<int, int>>> p4d(M);
    for (int i = 0; i < M; ++i) {
        pred[i] = {cv::Point2f(predL[2 * i], predL[2 * i + 1]), cv::Point2f(predR[2 * i], predR[2 * i + 1])};
        maxDists[i] = (float)std::exp((double)lastKF.logScale * ((double)predScale[i] - 0.5));
        p4d[i] = {Eigen::Vector4d(0, 0, 1, 1), {i, i}};
    }
    int n = m->fm->matchByProjectionRPredLBA(&lastKF, &newKF, matched, rad, pred, maxDists, p4d);
    for (int i = 0; i < M; ++i) {
        out[2 * i] = out[2 * i + 1] = -2;
        if (!matched[i].empty()) { out[2 * i] = matched[i].back().second.first; out[2 * i + 1] = matched[i].back().second.second; }
    }
    return n;
}

// CPU baseline on the reference's own code.  Per frame: extractKeysNew x2 + findStereoMatchesORB2R + the two key grids (through
// the restated assignKeysToGrids of liboracle.so: FeatureTracker.cpp is not compiled), i.e. FeatureTracker::extractORBAndStereoMatch
// (src/FeatureTracker.cpp:56-70); then, when `prev` is given, the temporal step of bench.py: the map points of frame f from the
// stereo keys of frame prev[f] (ora_temporal_mappoints: initializeMap / MapPoint::update / removeOutOfFrameMPs restated) matched
// with the reference's own matchByProjectionRPred(rad) on fresh index arrays (src/FeatureTracker.cpp:1176-1205).
// shape 0 = reference-shaped (frames in sequence, two extraction threads per frame, serial matcher), shape 1 = frame-parallel
// over `threads` workers.  Uses whatever hook table is installed (default: restated scalar primitives, thread safe).
// Returns stereo matches + temporal matches (so the work cannot be optimised away).
long ref_track_frames(int nF, int nL, float s, int edge, int patch, int maxF, int minF, double fx, float baseline,
                      const uint8_t* L, const uint8_t* R, int frames, int w, int h, int shape, int threads, int* nKeysOut,
                      const int* prev, const double* posePrev, const double* toCamera, const double* toCameraR, double cfx, double cfy,
                      double ccx, double ccy, float rad, int* nTemporalOut) {
    std::atomic<long> total{0};
    std::vector<TrackedKeys> all(prev ? frames : 0);
    auto one = [&](int f, bool twoThreads) {
        ExtractorPtr feL = std::make_shared<FeatureExtractor>(nF, nL, s, edge, patch, maxF, minF);
        ExtractorPtr feR = std::make_shared<FeatureExtractor>(nF, nL, s, edge, patch, maxF, minF);
        Matcher* m = (Matcher*)ref_matcher_create(&feL, &feR, fx, baseline, h);
        cv::Mat li(h, w, CV_8UC1, (void*)(L + (size_t)f * w * h)), ri(h, w, CV_8UC1, (void*)(R + (size_t)f * w * h));
        TrackedKeys local;
        TrackedKeys& keys = prev ? all[f] : local;
        if (twoThreads) {
            std::thread a([&] { feL->extractKeysNew(li, keys.keyPoints, keys.Desc); });
            std::thread b([&] { feR->extractKeysNew(ri, keys.rightKeyPoints, keys.rightDesc); });
            a.join(); b.join();
        } else {
            feL->extractKeysNew(li, keys.keyPoints, keys.Desc);
            feR->extractKeysNew(ri, keys.rightKeyPoints, keys.rightDesc);
        }
        m->fm->findStereoMatchesORB2R(li, ri, keys.rightDesc, keys.rightKeyPoints, keys);
        const int nl = (int)keys.keyPoints.size(), nr = (int)keys.rightKeyPoints.size();
        const int yGrids = (int)std::ceil(64.f / ((float)w / (float)h));
        std::vector<int> start(64 * (yGrids + 1) + 1), items(std::max(1, std::max(nl, nr)));
        float xm, ym;
        ora_assign_grid(keys.keyPoints.data(), nl, w, h, start.data(), items.data(), &xm, &ym);
        if (prev) keys.lkeyGrid = gridFromCsr(start.data(), items.data(), yGrids, 64);
        ora_assign_grid(keys.rightKeyPoints.data(), nr, w, h, start.data(), items.data(), &xm, &ym);
        if (prev) { keys.rkeyGrid = gridFromCsr(start.data(), items.data(), yGrids, 64); keys.xGrids = 64; keys.yGrids = yGrids; keys.xMult = xm; keys.yMult = ym; }
        long matched = 0;
        for (int v : keys.rightIdxs) matched += v >= 0;
        total += matched;
        if (nKeysOut) nKeysOut[f] = nl;
        ref_matcher_destroy(m);
    };
    auto temporal = [&](int f) {
        if (nTemporalOut) nTemporalOut[f] = 0;
        const int p = prev[f];
        if (p < 0) return;
        ExtractorPtr feL = std::make_shared<FeatureExtractor>(nF, nL, s, edge, patch, maxF, minF);
        Matcher* m = (Matcher*)ref_matcher_create(&feL, &feL, fx, baseline, h);
        const TrackedKeys& pk = all[p];
        TrackedKeys& keys = all[f];
        const int np = (int)pk.keyPoints.size();
        std::vector<int> key(std::max(np, 1)), ll(std::max(np, 1)), lr(std::max(np, 1));
        std::vector<float> pl(2 * std::max(np, 1)), pr(2 * std::max(np, 1));
        const float logScale = (float)std::log((double)s);
        const int M = ora_temporal_mappoints(pk.keyPoints.data(), np, pk.estimatedDepth.data(), posePrev + 16 * f, toCamera + 16 * f,
                                             toCameraR + 16 * f, cfx, cfy, ccx, ccy, w, h, feL->scalePyramid.data(), nL, logScale, key.data(),
                                             pl.data(), pr.data(), ll.data(), lr.data(), nullptr, nullptr);
        std::vector<MapPoint*> mps(M);
        cv::KeyPoint obs;
        for (int i = 0; i < M; ++i) {
            MapPoint* mp = new MapPoint(Eigen::Vector4d(0, 0, 0, 1), pk.Desc.row(key[i]), obs, 0, (unsigned long)i);
            mp->predL = cv::Point2f(pl[2 * i], pl[2 * i + 1]); mp->predR = cv::Point2f(pr[2 * i], pr[2 * i + 1]);
            mp->scaleLevelL = ll[i]; mp->scaleLevelR = lr[i]; mp->inFrame = true; mp->inFrameR = true;
            mps[i] = mp;
        }
        std::vector<int> mL(keys.keyPoints.size(), -1), mR(keys.rightKeyPoints.size(), -1);
        std::vector<std::pair<int, int>> pairs(M, std::make_pair(-1, -1));
        const int n = m->fm->matchByProjectionRPred(mps, keys, mL, mR, pairs, rad);
        for (MapPoint* mp : mps) delete mp;
        total += n;
        if (nTemporalOut) nTemporalOut[f] = n;
        ref_matcher_destroy(m);
    };
    auto parallel = [&](auto&& fn) {
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, threads); t++)
            pool.emplace_back([&] { for (int f; (f = next++) < frames;) fn(f); });
        for (auto& t : pool) t.join();
    };
    if (shape == 0) {
        for (int f = 0; f < frames; f++) one(f, true);
        if (prev) for (int f = 0; f < frames; f++) temporal(f);
    } else {
        parallel([&](int f) { one(f, false); });
        if (prev) parallel(temporal);
    }
    return total.load();
}

long ref_stereo_frames(int nF, int nL, float s, int edge, int patch, int maxF, int minF, double fx, float baseline,
                       const uint8_t* L, const uint8_t* R, int frames, int w, int h, int shape, int threads, int* nKeysOut) {
    return ref_track_frames(nF, nL, s, edge, patch, maxF, minF, fx, baseline, L, R, frames, w, h, shape, threads, nKeysOut, nullptr, nullptr,
                            nullptr, nullptr, 0, 0, 0, 0, 0.f, nullptr);
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return FeatureMatcher::DescriptorDistance(descMat(a, 1), descMat(b, 1));
}

}  // extern "C"


// ---- FeatureTracker functions next to the path (reference src/FeatureTracker.cpp, bodies compiled from the reference source by
// oracle/slice_reference.py) and MapPoint (src/Map.cpp, compiled whole) ------------------------------------------------------
namespace {
Eigen::Matrix4d mat4(const double* rowMajor) {
    Eigen::Matrix4d m;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = rowMajor[4 * r + c];
    return m;
}
void mat4Out(const Eigen::Matrix4d& m, double* rowMajor) {
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) rowMajor[4 * r + c] = m(r, c);
}
struct Tracker {
    std::shared_ptr<StereoCamera> zed;
    ExtractorPtr feL, feR;
    std::shared_ptr<Map> map;
    std::unique_ptr<FeatureTracker> ft;
};
// a key frame that carries what MapPoint::predictScale reads (lastObsKF->logScale / nScaleLev)
KeyFrame* scaleKF(float logScale, int nScaleLev) {
    cv::Mat none;
    KeyFrame* kf = new KeyFrame(Eigen::Matrix4d::Identity(), none, none, 0, 0);
    kf->logScale = logScale; kf->nScaleLev = nScaleLev;
    return kf;
}
}  // namespace

extern "C" {

// FeatureTracker::FeatureTracker (src/FeatureTracker.cpp:23) on a camera described by the values the functions below read
void* ref_tracker_create(void* feL, void* feR, double fx, double fy, double cx, double cy, float baseline, int width, int height,
                         const double* extrinsics) {
    Tracker* t = new Tracker;
    t->zed = std::make_shared<StereoCamera>();
    t->zed->mBaseline = baseline; t->zed->mWidth = width; t->zed->mHeight = height;
    t->zed->mCameraLeft = std::make_shared<Camera>(nullptr, std::string());
    t->zed->mCameraLeft->fx = fx; t->zed->mCameraLeft->fy = fy; t->zed->mCameraLeft->cx = cx; t->zed->mCameraLeft->cy = cy;
    if (extrinsics) t->zed->extrinsics = mat4(extrinsics);
    t->feL = *(ExtractorPtr*)feL; t->feR = *(ExtractorPtr*)feR;
    t->map = std::make_shared<Map>();
    t->ft.reset(new FeatureTracker(t->zed, t->feL, t->feR, t->map));
    return t;
}
void ref_tracker_destroy(void* h) { delete (Tracker*)h; }

// the matrices the reference derives from its pose arguments, with the arithmetic of THIS build (mini-Eigen inverse): tests give
// exactly these to the CUDA entry points, whose callers compute them on the host as well (include/gvs.h)
void ref_tracker_cameras(void* h, const double* predNPose, double* toCamera, double* toCameraR) {   // :913-914, :972-973
    Tracker* t = (Tracker*)h;
    const Eigen::Matrix4d P = mat4(predNPose);
    mat4Out(P.inverse(), toCamera);
    mat4Out((P * t->zed->extrinsics).inverse(), toCameraR);
}
void ref_outlier_cameras(void* h, const double* estimPose, double* toCameraR) {   // :584-585
    Tracker* t = (Tracker*)h;
    const Eigen::Matrix4d E = mat4(estimPose);
    mat4Out((E.inverse() * t->zed->extrinsics).inverse(), toCameraR);
}

// FeatureTracker::assignKeysToGrids (:28-54), CSR out (start: yGrids * 64 + 1 entries); returns yGrids
int ref_assign_grid(void* h, const void* kps, int n, int width, int height, int* start, int* items, float* xMult, float* yMult) {
    Tracker* t = (Tracker*)h;
    TrackedKeys keys;
    std::vector<cv::KeyPoint> k = keyVector(kps, n);
    std::vector<std::vector<std::vector<int>>> grid;
    t->ft->assignKeysToGrids(keys, k, grid, width, height);
    int run = 0;
    for (int r = 0; r < keys.yGrids; ++r)
        for (int c = 0; c < keys.xGrids; ++c) {
            start[r * keys.xGrids + c] = run;
            for (int v : grid[r][c]) items[run++] = v;
        }
    start[keys.yGrids * keys.xGrids] = run;
    *xMult = keys.xMult; *yMult = keys.yMult;
    return keys.yGrids;
}

// removeOutOfFrameMPs (:910-940; mode 0), removeOutOfFrameMPsMono (:942-967; mode 1), PredictMPsPosition (:969-1013; mode 2) on map
// points rebuilt from flat arrays (same contract as ora_predict_mappoints, but the POSE predNPose goes in: the reference inverts
// it itself).  valid[i] == 0 stands for a null entry (mode 2) / an outlier map point (modes 0/1).
void ref_predict_mappoints(void* h, int mode, const double* currCamPose, const double* predNPose, int M, const uint8_t* valid,
                           const double* wPos, const float* maxScaleDist, const float* logScale, const int* nScaleLev, float* predL,
                           int* scaleLevelL, uint8_t* inFrame, float* predR, int* scaleLevelR, uint8_t* inFrameR, uint8_t* keep, int* matches,
                           int* matchedL, int nL, int* matchedR, int nR, uint8_t* outliers) {
    Tracker* t = (Tracker*)h;
    std::vector<MapPoint*> mps(M, nullptr), all;
    std::vector<KeyFrame*> kfs;
    cv::KeyPoint obs;
    cv::Mat d(1, 32, CV_8U);
    for (int i = 0; i < M; ++i) {
        const bool ok = !valid || valid[i];
        if (mode == 2 && !ok) continue;   // null entry
        MapPoint* mp = new MapPoint(Eigen::Vector4d(wPos[4 * i], wPos[4 * i + 1], wPos[4 * i + 2], wPos[4 * i + 3]), d, obs, 0, (unsigned long)i);
        KeyFrame* kf = scaleKF(logScale[i], nScaleLev[i]);
        kfs.push_back(kf);
        mp->lastObsKF = kf; mp->maxScaleDist = maxScaleDist[i];
        mp->predL = cv::Point2f(predL[2 * i], predL[2 * i + 1]); mp->scaleLevelL = scaleLevelL[i]; mp->inFrame = inFrame[i] != 0;
        if (mode != 1) { mp->predR = cv::Point2f(predR[2 * i], predR[2 * i + 1]); mp->scaleLevelR = scaleLevelR[i]; mp->inFrameR = inFrameR[i] != 0; }
        if (!ok) mp->SetIsOutlier(true);
        mps[i] = mp; all.push_back(mp);
    }
    const Eigen::Matrix4d C = mat4(currCamPose), P = mat4(predNPose);
    if (mode == 2) {
        std::vector<int> mL(matchedL, matchedL + nL), mR(matchedR, matchedR + nR);
        std::vector<std::pair<int, int>> pairs(M);
        std::vector<bool> out(M);
        for (int i = 0; i < M; ++i) { pairs[i] = {matches[2 * i], matches[2 * i + 1]}; out[i] = outliers[i] != 0; }
        t->ft->PredictMPsPosition(C, P, mps, mL, mR, pairs, out);
        std::memcpy(matchedL, mL.data(), (size_t)nL * sizeof(int)); std::memcpy(matchedR, mR.data(), (size_t)nR * sizeof(int));
        for (int i = 0; i < M; ++i) { matches[2 * i] = pairs[i].first; matches[2 * i + 1] = pairs[i].second; outliers[i] = out[i]; }
    } else {
        std::vector<MapPoint*> act = mps;
        if (mode == 0) t->ft->removeOutOfFrameMPs(C, P, act); else t->ft->removeOutOfFrameMPsMono(C, P, act);
        for (int i = 0; i < M; ++i) keep[i] = 0;
        for (MapPoint* mp : act) keep[mp->idx] = 1;
    }
    for (int i = 0; i < M; ++i) {
        MapPoint* mp = mps[i];
        if (!mp) continue;
        predL[2 * i] = mp->predL.x; predL[2 * i + 1] = mp->predL.y; scaleLevelL[i] = mp->scaleLevelL; inFrame[i] = mp->inFrame;
        if (mode != 1) { predR[2 * i] = mp->predR.x; predR[2 * i + 1] = mp->predR.y; scaleLevelR[i] = mp->scaleLevelR; inFrameR[i] = mp->inFrameR; }
    }
    for (MapPoint* mp : all) delete mp;
    for (KeyFrame* kf : kfs) delete kf;
}

// findOutliersR (:582-649; mode 0) / findOutliersMono (:651-684; mode 1) with check2dError (:147-164); same contract as
// ora_find_outliers except that the reference computes toCameraR from estimPose and the extrinsics itself
void ref_find_outliers(void* h, int mode, const double* estimPose, double thres, int M, const uint8_t* valid, const double* wPos,
                       const uint8_t* inFrame, const uint8_t* inFrameR, int* matches, uint8_t* outliers, const void* kl, int NL,
                       const void* kr, int NR, float* depth, uint8_t* close, int* rightIdxs, int* leftIdxs, int* counts) {
    Tracker* t = (Tracker*)h;
    TrackedKeys keys;
    keys.keyPoints = keyVector(kl, NL);
    if (kr) keys.rightKeyPoints = keyVector(kr, NR);
    if (mode == 0) {
        keys.estimatedDepth.assign(depth, depth + NL); keys.rightIdxs.assign(rightIdxs, rightIdxs + NL); keys.leftIdxs.assign(leftIdxs, leftIdxs + NR);
        keys.close.resize(NL);
        for (int i = 0; i < NL; ++i) keys.close[i] = close[i] != 0;
    }
    std::vector<MapPoint*> mps(M, nullptr);
    cv::KeyPoint obs;
    cv::Mat d(1, 32, CV_8U);
    for (int i = 0; i < M; ++i) {
        if (valid && !valid[i]) continue;
        MapPoint* mp = new MapPoint(Eigen::Vector4d(wPos[4 * i], wPos[4 * i + 1], wPos[4 * i + 2], wPos[4 * i + 3]), d, obs, 0, (unsigned long)i);
        mp->inFrame = inFrame ? inFrame[i] != 0 : true; mp->inFrameR = inFrameR ? inFrameR[i] != 0 : true;
        mps[i] = mp;
    }
    std::vector<std::pair<int, int>> pairs(M);
    std::vector<bool> out(M);
    for (int i = 0; i < M; ++i) { pairs[i] = {matches[2 * i], matches[2 * i + 1]}; out[i] = outliers[i] != 0; }
    std::vector<float> weights;
    int nInliers = 0, nStereo = 0;
    const Eigen::Matrix4d E = mat4(estimPose);
    if (mode == 0) nStereo = t->ft->findOutliersR(E, mps, keys, pairs, thres, out, weights, nInliers);
    else t->ft->findOutliersMono(E, mps, keys, pairs, thres, out, weights, nInliers);
    for (int i = 0; i < M; ++i) { matches[2 * i] = pairs[i].first; matches[2 * i + 1] = pairs[i].second; outliers[i] = out[i]; }
    if (mode == 0) {
        for (int i = 0; i < NL; ++i) { depth[i] = keys.estimatedDepth[i]; close[i] = keys.close[i]; rightIdxs[i] = keys.rightIdxs[i]; }
        for (int i = 0; i < NR; ++i) leftIdxs[i] = keys.leftIdxs[i];
    }
    counts[0] = nInliers; counts[1] = nStereo;
    for (MapPoint* mp : mps) delete mp;
}

// The map points of the temporal step, built by the reference itself: FeatureTracker::initializeMap (:72-123) on the stereo keys of
// a frame tracked at camera pose posePrev (MapPoint::MapPoint / MapPoint::update / MapPoint::calcDescriptor of src/Map.cpp run inside),
// then FeatureTracker::removeOutOfFrameMPs (:910-940) with predNPose.  Outputs per surviving point, in order: source key, predL/R,
// scaleLevelL/R, world position, maxScaleDist, descriptor.  Returns their number.
int ref_temporal_mappoints(void* h, const void* kl, const uint8_t* dl, int NL, const void* kr, const uint8_t* dr, int NR,
                           const int* rightIdxs, const float* depth, const double* posePrev, const double* predNPose, int* mpKey,
                           float* predL, float* predR, int* lvlL, int* lvlR, double* wPos, float* maxScaleDist, uint8_t* desc) {
    Tracker* t = (Tracker*)h;
    // a fresh tracker state per call: initializeMap appends to the map
    Tracker fresh;
    fresh.zed = std::make_shared<StereoCamera>(*t->zed);
    fresh.zed->mCameraPose.pose = mat4(posePrev);
    fresh.feL = t->feL; fresh.feR = t->feR; fresh.map = std::make_shared<Map>();
    fresh.ft.reset(new FeatureTracker(fresh.zed, fresh.feL, fresh.feR, fresh.map));
    TrackedKeys keys;
    keys.keyPoints = keyVector(kl, NL); keys.Desc = descMat(dl, NL);
    keys.rightKeyPoints = keyVector(kr, NR); keys.rightDesc = descMat(dr, NR);
    keys.rightIdxs.assign(rightIdxs, rightIdxs + NL); keys.estimatedDepth.assign(depth, depth + NL);
    keys.leftIdxs.assign(NR, -1); keys.close.assign(NL, false);
    fresh.ft->initializeMap(keys);
    std::vector<MapPoint*> all = fresh.map->activeMapPoints;
    KeyFrame* kf = all.empty() ? nullptr : all[0]->lastObsKF;
    fresh.ft->removeOutOfFrameMPs(mat4(posePrev), mat4(predNPose), fresh.map->activeMapPoints);
    int M = 0;
    for (MapPoint* mp : fresh.map->activeMapPoints) {
        mpKey[M] = mp->kFMatches[kf].first;
        predL[2 * M] = mp->predL.x; predL[2 * M + 1] = mp->predL.y; predR[2 * M] = mp->predR.x; predR[2 * M + 1] = mp->predR.y;
        lvlL[M] = mp->scaleLevelL; lvlR[M] = mp->scaleLevelR;
        if (wPos) for (int c = 0; c < 4; ++c) wPos[4 * M + c] = mp->wp(c);
        if (maxScaleDist) maxScaleDist[M] = mp->maxScaleDist;
        if (desc) std::memcpy(desc + (size_t)M * 32, mp->desc.ptr(0), 32);
        ++M;
    }
    for (MapPoint* mp : all) delete mp;
    for (auto& kv : fresh.map->keyFrames) delete kv.second;
    return M;
}

// MapPoint::calcDescriptor (src/Map.cpp:145-210) for one map point observed by n key frames (one left descriptor each).  The
// reference walks an unordered_map<KeyFrame*, ...>: order[] receives the order in which it visits the observations (ties are
// resolved by that order), chosen[] the 32 bytes it picks.
void ref_median_descriptor(const uint8_t* descs, int n, int* order, uint8_t* chosen) {
    cv::KeyPoint obs;
    MapPoint mp(Eigen::Vector4d(0, 0, 0, 1), descMat(descs, n ? 1 : 0), obs, 0, 0);
    std::vector<KeyFrame*> kfs(n);
    cv::Mat none;
    for (int i = 0; i < n; ++i) {
        kfs[i] = new KeyFrame(Eigen::Matrix4d::Identity(), none, none, i, i);
        kfs[i]->keys.Desc = descMat(descs + (size_t)i * 32, 1);
        mp.kFMatches.insert({kfs[i], {0, -1}});
    }
    {
        std::unordered_map<KeyFrame*, std::pair<int, int>> observations = mp.kFMatches;   // the copy the reference iterates (:150)
        int k = 0;
        for (auto& kv : observations) order[k++] = (int)kv.first->numb;
    }
    mp.calcDescriptor();
    std::memcpy(chosen, mp.desc.ptr(0), 32);
    for (KeyFrame* kf : kfs) delete kf;
}

}  // extern "C"

// ---- link stubs -----------------------------------------------------------------------------------------------------
// Constructors that the reference defines in files which are not on the path (src/KeyFrame.cpp, src/Camera.cpp) and which this driver needs to create the objects the matcher reads.  They only initialise members; no
// behaviour of the reference is restated here.
namespace GTSAM_VIOSLAM {
CameraPose::CameraPose(Eigen::Matrix4d _pose, std::chrono::time_point<std::chrono::high_resolution_clock> _timestamp)
    : pose(_pose), timestamp(_timestamp) {}
Camera::Camera(std::shared_ptr<ConfigFile> configFile, const std::string&) : mConfigFile(configFile) {}
// src/Camera.cpp:22-25; referenced by MapPoint::updatePos (src/Map.cpp:224), which nothing here calls
Eigen::Matrix4d CameraPose::getInvPose() const { return poseInverse; }
KeyFrame::KeyFrame(Eigen::Matrix4d _pose, cv::Mat& _leftIm, cv::Mat& rLIm, const int _numb, const int _frameIdx)
    : pose(_pose), leftIm(_leftIm), rLeftIm(rLIm), numb(_numb), frameIdx(_frameIdx) {}
}  // namespace GTSAM_VIOSLAM
