// Handle structures behind the opaque types of include/gvs.h, shared by the host-side translation units (context.cu: extractor
// and per-frame matcher entry points; temporal.cu: batched temporal matching; multi.cu: multi-device front end).  Internal.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <map>

#include "common.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX: ranges show up in nsys / ncu timelines, a no-op without a tool attached

namespace gvs {

// Entry points run on the handle's device and leave the caller's current device as they found it (a caller such as PyTorch
// resolves "cuda" to the current device).
// NVTX range for the duration of a C ABI call (tracing)
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope&) = delete;
    NvtxScope& operator=(const NvtxScope&) = delete;
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != device) prev = cur;
        cudaSetDevice(device);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Debugging aids for device memory (compute-sanitizer is not available on the GPU pool this was developed on):
//   GVS_POISON=1  every fresh allocation is filled with 0xA5, so that a kernel relying on cudaMalloc returning zeros shows up as a
//                 parity failure in the tests / tools/fuzz_parity.py (initcheck by other means);
//   GVS_GUARD=1   every allocation gets a 256-byte tail filled with 0xC3 and is registered; gvs_debug_check_guards() (gvs.h)
//                 verifies all tails: an out-of-bounds write past the end of a buffer is reported with the buffer's size.
constexpr size_t kGuardBytes = 256;
void guardRegister(void* p, size_t bytes);
void guardForget(void* p);
inline cudaError_t devFree(void* p) { if (p) guardForget(p); return cudaFree(p); }
template <typename T>
static cudaError_t devAlloc(T** p, size_t count) {
    *p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    static const bool poison = getenv("GVS_POISON") != nullptr, guard = getenv("GVS_GUARD") != nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes + (guard ? kGuardBytes : 0));
    if (e == cudaSuccess && poison) e = cudaMemset(*p, 0xA5, bytes);
    if (e == cudaSuccess && guard) { e = cudaMemset(reinterpret_cast<uint8_t*>(*p) + bytes, 0xC3, kGuardBytes); guardRegister(*p, bytes); }
    if (e == cudaSuccess && (poison || guard)) cudaStreamSynchronize(cudaStreamLegacy);   // (handles use non-blocking streams)
    return e;
}

}  // namespace gvs

using namespace gvs;  // (internal header: the handle structs are global, their members are gvs:: types)

struct gvs_extractor {
    int device = 0;
    gvs_extractor_params params{};
    int width = 0, height = 0, maxImages = 0;
    Geometry g{};
    std::vector<float> scale, invScale, sigma, invSigma;
    std::vector<int> patch, quota;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    cudaStream_t auxStream = nullptr;  // blur runs here, concurrently with FAST + selection
    cudaEvent_t doneEvent = nullptr, evPyr = nullptr, evBlur = nullptr;
    // device
    uint8_t* slab = nullptr;
    uint8_t* blurSlab = nullptr;
    uint8_t* tight = nullptr;   // staging for tightly packed host batches (one contiguous DMA, then repack)
    size_t tightBytes = 0;
    float* mapX = nullptr;      // rectification maps of cv::remap (optional, gvs_extractor_set_rectify_maps)
    float* mapY = nullptr;
    ResizeTap* taps = nullptr;
    uint4* xWeights = nullptr;        // per group of 4 destination columns: a0 | a1 << 16 of each column
    uint4* xSelectors = nullptr;      // ... PRMT selector (| 1 << 16 when the pair sits in the upper 8 bytes of the window)
    unsigned short* xBase = nullptr;  // ... source offset of the window: sx of the first column (narrow levels) or the 4-byte aligned
                                      //     start of the 12-byte window (other levels)
    PyrPlan pyrPlan{};                // fused pyramid (nGroups == 0: legacy per-level kernels)
    PyrSpan* pyrSpans = nullptr;
    alignas(64) CUtensorMap pyrMaps[kPyrMaxGroups];   // TMA tile maps of the groups' base levels
    unsigned pyrNarrowMask = 0;       // bit l: every column group of level l spans at most 8 source bytes (pyr_down_kernel<true>)
    alignas(64) CUtensorMap fastMaps[kMaxLevels];   // TMA tile maps of the levels for the FAST kernel
    bool fastTma = false;
    FastItem* fastItems = nullptr;
    int nFastItems = 0;
    // CUDA graphs of the extraction pipeline for small batches (latency mode): one instantiated graph per batch size
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int> graphLaunches;
    BlurItem* blurItems = nullptr;
    int nBlurItems = 0;
    CandRec* recs = nullptr;
    int* nCand = nullptr;
    unsigned short* cellCnt = nullptr;
    KeptKey* kept = nullptr;
    int* keptCnt = nullptr;
    int* imgOff = nullptr;
    unsigned* orderedDbg = nullptr;
    unsigned* sortedDbg = nullptr;
    int* momentsDbg = nullptr;
    uint8_t* selScratch = nullptr;
    size_t selScratchPerCta = 0;
    int selCtas = 0;
    int* workCounter = nullptr;
    int* status = nullptr;
    gvs_keypoint* kps = nullptr;
    uint8_t* desc = nullptr;
    uint2* auxKey = nullptr;   // per key: stereo bucket row range + pt.y (compact copy for the matcher's scan)
    int totalCap = 0;
    // host (pinned)
    int* hOff = nullptr;
    int* hStatus = nullptr;
    int lastN = 0;
    uint64_t generation = 0;  // bumped by every batch that replaces the resident results (runExtract, gvs_debug_select)
    bool pending = false;
    bool blurValid = false;   // the blurred slab is only materialised on demand (parity tap)
    int launches = 0;
    int64_t launchesTotal = 0;
    bool timing = false;
    cudaEvent_t ev[8] = {};
    gvs_stage_times times{};
};

struct gvs_matcher {
    gvs_extractor* L = nullptr;
    gvs_extractor* R = nullptr;
    int device = 0;
    double fx = 0;
    float baseline = 0;
    int imageHeight = 0;
    int* rightIdxs = nullptr;
    int* leftIdxs = nullptr;
    int* sadBest = nullptr;
    void* sortScratch = nullptr;
    float* depth = nullptr;
    uint8_t* close = nullptr;
    // grids
    int* cellStartL = nullptr;
    int* cellStartR = nullptr;
    int* itemsL = nullptr;
    int* itemsR = nullptr;
    int maxCells = GVS_GRID_X * GVS_GRID_X;   // set at creation from the image shape
    static constexpr int kMaxGridRows = 512;   // 64 x 512 cells: the counting sort of grid_assign_kernel keeps them in shared memory
    int xGrids = 0, yGrids = 0;
    float xMult = 0, yMult = 0;
    bool gridsValid = false, stereoValid = false;
    int lastFrames = 0, gridFrames = 0;
    // the extractors' batch generations the stereo arrays / key grids were computed from: a call that combines them with the
    // resident keys checks that nothing was re-extracted in between (GVS_ERR_STATE otherwise)
    uint64_t stereoGenL = 0, stereoGenR = 0, gridGenL = 0, gridGenR = 0;
    // batched temporal tracking (temporal.cu), allocated on first use
    uint8_t* tFlag = nullptr;
    float* tPredL = nullptr; float* tPredR = nullptr;
    int* tLvlL = nullptr; int* tLvlR = nullptr; int* tDescRow = nullptr;
    double* tWPos = nullptr; float* tMaxScaleDist = nullptr;
    int* tMpCnt = nullptr; int* tMpOff = nullptr; int* tMpSlot = nullptr; int* tMpKey = nullptr;
    int* tMatchedL = nullptr; int* tMatchedR = nullptr; int* tMatches = nullptr; int* tNMatches = nullptr;
    uint8_t* tRecs = nullptr;
    double* tPoses = nullptr; int* tPrev = nullptr; int* tStatus = nullptr;
    std::vector<uint8_t> tStage;   // pageable staging of the per-frame poses / predecessor list
    int* hMpOff = nullptr; int* hNMatches = nullptr; int* hTStatus = nullptr;   // pinned
    int temporalFrames = 0;
    uint64_t temporalGen = 0;
    cudaEvent_t evT[2] = {};
    float temporalMs = 0;
    cudaEvent_t evR = nullptr, evFork = nullptr, evJoin = nullptr, evHold = nullptr, evLast = nullptr;
    cudaStream_t callerStream = nullptr;  // optional: work is forked from / joined into this stream
    // projection scratch (grown on demand)
    uint8_t* pj = nullptr;
    size_t pjBytes = 0;
    // the calls that do not read the resident frame (LBA matcher, calcDescriptor, DescriptorDistance, worldToFrame loops) run on the
    // matcher's own stream with their own scratch: a second matcher over the same extractors may serve another host thread
    // (System::mFeatureMatcher in LocalMapper) without touching the tracker's streams
    cudaStream_t ownStream = nullptr;
    uint8_t* pjOwn = nullptr;
    size_t pjOwnBytes = 0;
    unsigned* pool = nullptr;
    int* rowStart = nullptr;               // stereo row buckets (destributeRightKeys), see stereo.cu
    unsigned short* rowItems = nullptr;
    int rowItemsStride = 0;
    int poolCap = 0;
    int* pjStatus = nullptr;
    int launches = 0;
    int64_t launchesTotal = 0;
    cudaEvent_t ev[4] = {};
    float stereoMs = 0, gridMs = 0;
};

namespace gvs {
// context.cu
int syncExtract(gvs_extractor* fe);                 // wait for the handle's last batch, surface device-side status bits
int joinRight(gvs_matcher* m);                      // L's stream waits for R's last batch
int holdRight(gvs_matcher* m);                      // R's stream waits for the matcher work queued on L's stream
int growScratch(gvs_matcher* m, size_t bytes);      // m->pj of at least `bytes` (synchronises when it has to reallocate)
// temporal.cu
int checkResident(gvs_matcher* m, bool needStereo, int frame0, int nFrames);   // grids / stereo arrays belong to the current batch
}  // namespace gvs
