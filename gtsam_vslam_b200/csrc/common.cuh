// Shared declarations of the sm_100a frontend kernels (internal; the public surface is include/gvs.h).
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/gvs.h"

namespace gvs {

constexpr int kMaxLevels = GVS_MAX_LEVELS;
constexpr int kHalfPatch = 15;  // FeatureExtractor::halfPatchSize, include/FeatureExtractor.h:67
constexpr int kFastEdge = 3;    // computeKeypointsORBNew, src/FeatureExtractor.cpp:538
constexpr int kCellW = 35;      // W, :542

// Geometry and quotas of one pyramid level.  Lives in kernel parameter space (__grid_constant__).
struct LevelGeom {
    int w, h, pitch;         // pitch: bytes per row in the slab (multiple of 64)
    unsigned off;            // byte offset of the level inside one image's slab
    int nCols, nRows;        // FAST cell grid (src/FeatureExtractor.cpp:555-556)
    int gridW, gridH;        // cell pitch (:558-559)
    int cellBase;            // first entry of this level in the per-image cell-count array
    int quota, kmin, kmax;   // featurePerLevel, SSC tolerance band (:405-406)
    int sscHigh;             // closed-form upper bracket (:382-395)
    int candOff, candCap;    // candidate records of this level inside one image's record buffer
    int keptOff, keptCap;    // kept keys of this level inside one image's staging buffer
    int tabX, tabY;          // offsets of the resize tables of this level (dst level), in entries
    int xmax;                // first dst column whose source column is clamped (cv::resize)
    float scale, invScale;   // scalePyramid, scaleInvPyramid
    int scaledPatch;         // scaledPatchSize
    int fastItemBase;        // first FAST work item of this level
};

struct Geometry {
    int nLevels;
    int width, height;
    unsigned slabBytes;      // bytes per image in the pyramid slab
    int cellsPerImage;
    int candPerImage;        // records per image
    int keptPerImage;        // staging keys per image
    int tMin, tMax;          // minFastThreshold, maxFastThreshold
    int edge;                // edgeThreshold: keypoints lie in [edge, w-edge) x [edge, h-edge)
    int umax[kHalfPatch + 1];
    LevelGeom lv[kMaxLevels];
};

// One FAST work item = a strip of up to kFastCellsPerItem consecutive cells of one cell row.
#ifndef GVS_FAST_CELLS
#define GVS_FAST_CELLS 6
#endif
constexpr int kFastCellsPerItem = GVS_FAST_CELLS;
struct FastItem {
    short level, iR, iC0, nC;
};

// Resize tables entry for one destination column / row (cv::resize INTER_LINEAR fixed point, 11 bits).
struct alignas(8) ResizeTap {   // (row taps are fetched with one 8-byte load: {ofs | a0 << 16, a1 | pad << 16})
    short ofs, a0, a1, pad;
};

// 8-byte candidate record written by the FAST kernel in arbitrary order.
//   w0 = x | y << 12 | score << 24 ; w1 = cell | rank << 16   (rank = position inside the cell's emission order)
struct CandRec {
    unsigned w0, w1;
};

// Kept key after selection (staging, per level region): x | y << 12 | score << 24
typedef unsigned KeptKey;

#define GVS_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return gvs::fail(GVS_ERR_CUDA, #expr, cudaGetErrorString(_e)); \
    } while (0)

int fail(int code, const char* what, const char* detail = nullptr);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember the largest request per (kernel, device).  Handles
// are driven from several host threads (two extractors, gvs_multi workers), so the cache is guarded and only ever grows.
std::mutex& smemAttrMutex();
template <typename K>
inline void ensureDynamicSmem(K kernel, size_t bytes, size_t* cache /* [64] zero-initialised, one per call site */) {
    // (opt in from 40 KB: a kernel's static shared memory counts against the 48 KB default as well)
    if (bytes <= 40 * 1024) return;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::lock_guard<std::mutex> lock(smemAttrMutex());
    if (cache[dev] < bytes) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        cache[dev] = bytes;
    }
}

struct StageTimer;

// ---- fused pyramid (pyramid.cu): a CTA owns a tile of the coarsest level of a group of consecutive levels and produces that
// tile and the regions of the finer levels of the group it depends on from ONE TMA load of the group's base level.
constexpr int kPyrMaxFuse = 4;      // output levels per fused group
constexpr int kPyrMaxGroups = 4;    // (3 + 4 + 4 covers the 11 output levels of a 12-level pyramid)
struct PyrSpan {                    // one tile column (or tile row) of a group; index 0 = base level, k = base + k
    short lo[kPyrMaxFuse + 1];      // first computed (base level: loaded) column / row
    short hi[kPyrMaxFuse + 1];      // one past the last computed
    short ownLo[kPyrMaxFuse + 1];   // the part this tile writes to global memory: [ownLo, ownHi)
    short ownHi[kPyrMaxFuse + 1];
    unsigned magic[kPyrMaxFuse + 1]; // column spans: 2^32 / (number of 4-column groups of the level's region) + 1
};
struct PyrGroup {
    int base, nOut;                 // levels base+1 .. base+nOut are produced from level base
    int nTx, nTy;                   // tiles
    int spanX, spanY;               // first entries of the group in the span table
    int boxW, boxH;                 // TMA box of the base level (bytes x rows)
    int bufOff[kPyrMaxFuse];        // shared-memory offsets of the level buffers (0 = base level box)
    int bufPitch[kPyrMaxFuse];
    int smemBytes;
};
struct PyrPlan {
    int nGroups;                    // 0 = the fused path is not available for this geometry (legacy kernels run)
    PyrGroup grp[kPyrMaxGroups];
};

// ---- kernel launchers (one translation unit each) -------------------------------------------------
void launchPyramid(const Geometry& g, uint8_t* slab, const ResizeTap* taps, const uint4* xw, const uint4* xs,
                   const unsigned short* xbase, unsigned narrowMask, int nImages, cudaStream_t s, int* launches);
void launchPyramidFused(const Geometry& g, uint8_t* slab, const ResizeTap* taps, const uint4* xw, const uint4* xs,
                        const unsigned short* xbase, const PyrPlan& plan, const PyrSpan* spans, const CUtensorMap* baseMaps /* per group, host */,
                        int nImages, cudaStream_t s, int* launches);
void launchPrepare(const Geometry& g, const uint8_t* src, int channels, const float* mapx, const float* mapy, uint8_t* slab,
                   int nImages, cudaStream_t s, int* launches);
void launchRepack(const Geometry& g, const uint8_t* tight, uint8_t* slab, int nImages, cudaStream_t s, int* launches);
// levelMaps: TMA tile maps of the pyramid levels with the box of fastTmaBox (HOST array of nLevels entries, passed to the kernel as a parameter), or nullptr
void launchFast(const Geometry& g, const uint8_t* slab, const FastItem* items, int nItems, int nImages, CandRec* recs,
                int* nCand, unsigned short* cellCnt, int* status, const CUtensorMap* levelMaps, cudaStream_t s, int* launches);
bool fastTmaBox(const Geometry& g, int* boxWords, int* boxH);
void launchSelect(const Geometry& g, int nImages, const CandRec* recs, const int* nCand, const unsigned short* cellCnt,
                  KeptKey* kept, int* keptCnt, unsigned* orderedDbg, unsigned* sortedDbg, uint8_t* scratch,
                  size_t scratchPerCta, int nCtas, int* workCounter, int* status, cudaStream_t s, int* launches);
size_t selectScratchBytes(const Geometry& g);
int selectMaxCtas(int device);
void launchOffsets(const Geometry& g, int nImages, const int* keptCnt, int* imgOff, int totalCap, int* status,
                   cudaStream_t s, int* launches);
struct BlurItem { short level, tx, ty, pad; };  // one 64x32 output tile
constexpr int kBlurTW = 64, kBlurTH = 32;
void launchBlur(const Geometry& g, const uint8_t* slab, uint8_t* blurSlab, const BlurItem* items, int nItems, int nImages,
                cudaStream_t s, int* launches);
void launchDescribeFused(const Geometry& g, const uint8_t* slab, int nImages, const KeptKey* kept, const int* keptCnt,
                         const int* imgOff, gvs_keypoint* kps, uint8_t* desc, uint2* auxKey, int* momentsDbg, int totalCap,
                         cudaStream_t s, int* launches);

struct StereoArgs {
    const uint8_t* slabL;
    const uint8_t* slabR;
    const gvs_keypoint* kpL;
    const gvs_keypoint* kpR;
    const uint8_t* descL;
    const uint8_t* descR;
    const int* offL;  // nFrames+1
    const int* offR;
    const int* keptCntL;  // per image per level counts (nImages x kMaxLevels)
    const int* keptCntR;
    const uint2* auxR; // per right key: bucket row range (mn | mx << 16) and pt.y, written by the descriptor kernel
    int* rowStart;             // per frame (imageHeight + 1): CSR over image rows of the right keys whose bucket range holds the row
    unsigned short* rowItems;  // per frame rowItemsStride entries: right key indices (inside the frame), row by row
    int rowItemsStride;
    int bucketRows;            // rows that have a bucket: min(imageHeight of the matcher, image height)
    int* rightIdxs;   // dense, left offsets
    int* leftIdxs;    // dense, right offsets
    float* depth;
    uint8_t* close;
    int* sadBest;     // dense, left offsets (scratch)
    void* sortScratch; // nFrames x pow2ceil(keys per image) x 8 bytes, only when the sort does not fit shared memory
    double fx;
    float baseline;
    int imageHeight;
    int closeNumber;
    int thDist;
};
void launchStereo(const Geometry& g, const StereoArgs& a, int nFrames, int maxKeysPerImage, cudaStream_t s, int* launches);

struct GridArgs {
    const gvs_keypoint* kps;
    const int* off;     // nImages+1
    int* cellStart;     // nImages x (maxCells+1)
    int* items;         // dense, same offsets as kps
    int xGrids, yGrids;
    float xMult, yMult;
    int maxCells;
};
void launchGrid(const GridArgs& a, int nImages, cudaStream_t s, int* launches);
void launchGrid2(const GridArgs& a, const GridArgs* b, int nImages, cudaStream_t s, int* launches);

struct ProjArgs {
    int mode;           // gvs_match_mode
    int M;
    const uint8_t* mpDesc;
    const float* predL;
    const float* predR;
    const int* lvlL;
    const int* lvlR;
    const uint8_t* inFrame;
    const uint8_t* inFrameR;
    const gvs_keypoint* kpL;
    const gvs_keypoint* kpR;
    const uint8_t* descL;
    const uint8_t* descR;
    int NL, NR;
    const int* cellStartL;
    const int* itemsL;
    const int* cellStartR;
    const int* itemsR;
    const int* leftIdxs;   // stereo partner arrays of the frame (mode 0)
    const int* rightIdxs;
    int xGrids, yGrids;
    float xMult, yMult;
    float rad;
    float scalePyramid[kMaxLevels];
    // candidate lists (phase 1 output / phase 2 input)
    void* recs;            // 2*M ProjRec (project.cu)
    int* poolTop;          // [2] list pool counters (left, right)
    unsigned* candL;       // packed idx | dist << 16 | level << 25, see project.cu
    unsigned* candR;
    int candCap;           // entries per pool
    // in/out
    int* matchedL;
    int* matchedR;
    int* matches;          // M x 2
    int* nMatches;
    int* status;
};
// stored candidate lists (wide search windows): one argument block per frame, in device memory and on the host
void launchProjectWide(const ProjArgs* frames, const ProjArgs* hostFrames, int nFrames, cudaStream_t s, int* launches);
void launchProjectLBA(const ProjArgs& a, const float* lastL, const float* lastR, int* out, cudaStream_t s, int* launches);
void launchMedianDescriptor(const uint8_t* descs, const int* offsets, int P, int* bestIdx, cudaStream_t s, int* launches);
void launchWorldToFrame(int M, const double* poseInv, double fx, double fy, double cx, double cy, int width, int height,
                        const double* wPos, const float* maxScaleDist, const float* logScale, const int* nScaleLev, float* pred,
                        int* scaleLevel, uint8_t* inFrame, cudaStream_t s, int* launches);
// FeatureTracker::removeOutOfFrameMPs / removeOutOfFrameMPsMono / PredictMPsPosition
struct PredictArgs {
    int mode, M;               // 0 removeOutOfFrameMPs, 1 ...Mono, 2 PredictMPsPosition
    const double* toCamera;    // predNPose^-1, row-major
    const double* toCameraR;   // (predNPose * extrinsics)^-1
    double fx, fy, cx, cy;
    int width, height;
    const uint8_t* valid;
    const double* wPos;
    const float* maxScaleDist;
    const float* logScale;
    const int* nScaleLev;
    float* predL; int* scaleLevelL; uint8_t* inFrame;      // in-out
    float* predR; int* scaleLevelR; uint8_t* inFrameR;
    uint8_t* keep;             // modes 0/1
    int* matches; int* matchedL; int* matchedR; uint8_t* outliers;   // mode 2, in-out
};
void launchPredictMapPoints(const PredictArgs& a, cudaStream_t s, int* launches);

// FeatureTracker::findOutliersR / findOutliersMono
struct OutlierArgs {
    int mode;                  // 0 findOutliersR, 1 findOutliersMono
    int M, NL, NR;
    const double* pose;        // estimPose, row-major
    const double* poseR;       // toCameraR (mode 0)
    double fx, fy, cx, cy, thres;
    float closeDist;           // zedPtr->mBaseline * fm.closeNumber
    const float* invSigma;     // feLeft->InvSigmaFactor
    const double* wPos;        // M x 4
    const uint8_t* valid;      // mp != nullptr (NULL = all)
    const uint8_t* inFrame;
    const uint8_t* inFrameR;
    int* matches;              // M x 2, in-out
    uint8_t* outliers;         // M, in-out
    const gvs_keypoint* kpL;   // the frame's keys
    const gvs_keypoint* kpR;
    float* depth;              // the frame's stereo arrays, in-out
    uint8_t* close;
    int* rightIdxs;
    int* leftIdxs;
    int* act;                  // M scratch
    int* firstClear;           // NL scratch, preset to INT_MAX-like
    int* counters;             // [0] inliers, [1] stereo
};
void launchFindOutliers(const OutlierArgs& a, cudaStream_t s, int* launches);
void launchHamming(const uint8_t* a, const uint8_t* b, int n, int* out, cudaStream_t s, int* launches);

// status bits written by kernels into a device int (checked on sync)
enum : int { kStatCandOverflow = 1, kStatKeptOverflow = 2, kStatOutputOverflow = 4, kStatProjOverflow = 8 };

// ---- small device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* levelPtr(const uint8_t* slab, const Geometry& g, int img, int level) {
    return slab + (size_t)img * g.slabBytes + g.lv[level].off;
}
__device__ __forceinline__ uint8_t* levelPtr(uint8_t* slab, const Geometry& g, int img, int level) {
    return slab + (size_t)img * g.slabBytes + g.lv[level].off;
}

}  // namespace gvs
