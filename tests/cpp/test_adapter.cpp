// Exercises include/gvs_adapter.hpp (the C++ drop-in with the reference's class names) end to end on one stereo pair read
// from two raw u8 files; prints counts and FNV-1a checksums that tests/test_gpu_adapter.py compares with the ctypes path.
#include <cstdio>
#include <cstdlib>
#include "../../include/gvs_adapter.hpp"

using namespace GTSAM_VIOSLAM;

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
struct MapPoint { cv::Mat desc; cv::Point2f predL, predR; int scaleLevelL = 0, scaleLevelR = 0; bool inFrame = true, inFrameR = true; };

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: left.raw right.raw w h nFeatures nLevels\n"); return 2; }
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), nF = std::atoi(argv[5]), nL = std::atoi(argv[6]);
    cv::Mat L(h, w, cv::CV_8U), R(h, w, cv::CV_8U);
    FILE* f = std::fopen(argv[1], "rb"); if (!f || std::fread(L.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3; std::fclose(f);
    f = std::fopen(argv[2], "rb"); if (!f || std::fread(R.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3; std::fclose(f);
    try {
        auto feL = std::make_shared<FeatureExtractor>(nF, nL, 1.2f, 19, 31, 20, 7, w, h);
        auto feR = std::make_shared<FeatureExtractor>(nF, nL, 1.2f, 19, 31, 20, 7, w, h);
        FeatureMatcher fm(200.0, 0.1f, feL, feR, h);
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
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
