// Host side of the C ABI (include/gvs.h): handle lifetime, level geometry and tables (the FeatureExtractor
// constructor, reference src/FeatureExtractor.cpp:620-682), launch sequencing, transfers.  No CPU compute path:
// every entry point either launches the CUDA kernels or fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"
#include "handles.cuh"

namespace gvs {

static thread_local std::string g_lastError;

int fail(int code, const char* what, const char* detail) {
    g_lastError = what ? what : "";
    if (detail) { g_lastError += ": "; g_lastError += detail; }
    return code;
}

std::mutex& smemAttrMutex() {
    static std::mutex m;
    return m;
}

static std::mutex g_guardMutex;
static std::vector<std::pair<void*, size_t>> g_guards;   // GVS_GUARD=1: (allocation, payload bytes) of the live buffers
void guardRegister(void* p, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_guardMutex);
    g_guards.emplace_back(p, bytes);
}

void guardForget(void* p) {
    static const bool guard = getenv("GVS_GUARD") != nullptr;
    if (!guard) return;
    std::lock_guard<std::mutex> lock(g_guardMutex);
    for (auto& g : g_guards)
        if (g.first == p) { g = g_guards.back(); g_guards.pop_back(); break; }
}

static inline int cvRoundF(float v) { return (int)lrintf(v); }

}  // namespace gvs

using namespace gvs;

namespace {

// FeatureExtractor::FeatureExtractor, the tables that depend on the constructor arguments only (src/FeatureExtractor.cpp:622-659)
void buildTables(const gvs_extractor_params& p, std::vector<float>& scale, std::vector<float>& invScale, std::vector<float>& sigma,
                 std::vector<float>& invSigma, std::vector<int>& patch, std::vector<int>& quota) {
    const int L = p.n_levels;
    scale.assign(L, 1.f); invScale.assign(L, 1.f); sigma.assign(L, 1.f); invSigma.assign(L, 1.f);
    patch.assign(L, p.patch_size); quota.assign(L, 0);
    for (int i = 1; i < L; i++) {
        scale[i] = scale[i - 1] * p.im_scale;
        patch[i] = (int)(p.patch_size * scale[i]);
        sigma[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < L; i++) { invScale[i] = 1.0f / scale[i]; invSigma[i] = 1.0f / sigma[i]; }
    const float factor = 1.0f / p.im_scale;
    float nd = p.n_features * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; l++) { quota[l] = cvRoundF(nd); sum += quota[l]; nd *= factor; }
    quota[L - 1] = std::max(p.n_features - sum, 0);
}

int buildGeometry(gvs_extractor* fe) {
    const gvs_extractor_params& p = fe->params;
    const int L = p.n_levels;
    Geometry& g = fe->g;
    std::memset(&g, 0, sizeof(g));
    g.nLevels = L;
    g.width = fe->width;
    g.height = fe->height;
    g.tMin = p.min_fast_threshold;
    g.tMax = p.max_fast_threshold;
    g.edge = p.edge_threshold;
    // ---- tables, float32 cumulative (src/FeatureExtractor.cpp:622-644) and per-level quota (:648-659)
    buildTables(p, fe->scale, fe->invScale, fe->sigma, fe->invSigma, fe->patch, fe->quota);
    // ---- umax (:666-680)
    {
        const int hp = kHalfPatch;
        int v, v0, vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
        const double hp2 = hp * hp;
        for (v = 0; v <= vmax; ++v) g.umax[v] = (int)lrint(std::sqrt(hp2 - v * v));
        for (v = hp, v0 = 0; v >= vmin; --v) {
            while (g.umax[v0] == g.umax[v0 + 1]) ++v0;
            g.umax[v] = v0;
            ++v0;
        }
    }
    // ---- level geometry
    unsigned off = 0;
    int cellBase = 0, candOff = 0, keptOff = 0, tab = 0;
    for (int l = 0; l < L; l++) {
        LevelGeom& lv = g.lv[l];
        lv.w = cvRoundF((float)fe->width * fe->invScale[l]);    // computePyramid :347
        lv.h = cvRoundF((float)fe->height * fe->invScale[l]);
        if (lv.w > 4095 || lv.h > 4095) return fail(GVS_ERR_INVALID, "image larger than 4095 px is not supported");
        if (lv.w < 2 * p.edge_threshold + 8 || lv.h < 2 * p.edge_threshold + 8)
            return fail(GVS_ERR_INVALID, "pyramid level smaller than the border: reduce n_levels");
        lv.pitch = (lv.w + 63) & ~63;
        lv.off = off;
        off += (unsigned)lv.pitch * lv.h;
        off = (off + 255u) & ~255u;
        // FAST cells (computeKeypointsORBNew :544-559)
        const int minB = p.edge_threshold - kFastEdge;
        const int maxX = lv.w - p.edge_threshold + kFastEdge, maxY = lv.h - p.edge_threshold + kFastEdge;
        const int wid = maxX - minB, hig = maxY - minB;
        const float W = (float)kCellW;
        lv.nCols = (int)(wid / W);
        lv.nRows = (int)(hig / W);
        lv.gridW = lv.nCols > 0 ? (int)std::ceil((float)wid / lv.nCols) : 0;
        lv.gridH = lv.nRows > 0 ? (int)std::ceil((float)hig / lv.nRows) : 0;
        if (lv.gridW > 255 || lv.gridH > 255) return fail(GVS_ERR_INVALID, "FAST cell larger than 255 px");
        lv.cellBase = cellBase;
        cellBase += lv.nCols * lv.nRows;
        if (lv.nCols * lv.nRows > 65535) return fail(GVS_ERR_INVALID, "too many FAST cells in a level");
        // quota and SSC constants (ssc :382-406, called with tolerance 0.1f, cols, rows :612)
        lv.quota = fe->quota[l];
        {
            const int K = lv.quota, rows = lv.h, cols = lv.w;
            const int exp1 = rows + cols + 2 * K;
            const long long exp2 = ((long long)4 * cols + (long long)4 * K + (long long)4 * rows * K + (long long)rows * rows +
                                    (long long)cols * cols - (long long)2 * rows * cols + (long long)4 * rows * cols * K);
            const double exp3 = std::sqrt((double)exp2), exp4 = K - 1;
            const double sol1 = -std::round((exp1 + exp3) / exp4), sol2 = -std::round((exp1 - exp3) / exp4);
            lv.sscHigh = (int)((sol1 > sol2) ? sol1 : sol2);
            const float tol = 0.1f;
            const unsigned uK = (unsigned)K;
            lv.kmin = (int)(unsigned)std::round(uK - (uK * tol));
            lv.kmax = (int)(unsigned)std::round(uK + (uK * tol));
        }
        lv.candOff = candOff;
        lv.candCap = std::min(((lv.w * lv.h) / 8 + 256 + 3) & ~3, 65535 * 16);  // iid noise at a low threshold keeps ~1 px in 9 after the 3x3 NMS
        candOff += lv.candCap;
        lv.keptOff = keptOff;
        lv.keptCap = 2 * lv.quota + 64;
        keptOff += lv.keptCap;
        lv.scale = fe->scale[l];
        lv.invScale = fe->invScale[l];
        lv.scaledPatch = fe->patch[l];
        if (l > 0) {
            lv.tabX = tab; tab += (lv.w + 3) & ~3;
            lv.tabY = tab; tab += (lv.h + 3) & ~3;
        }
    }
    // whole rows of level 0 per image, so that a batch of host images is one strided 3-D copy
    const unsigned unit = (unsigned)g.lv[0].pitch * 4u;
    g.slabBytes = (off + unit - 1) / unit * unit;
    g.cellsPerImage = cellBase;
    g.candPerImage = candOff;
    g.keptPerImage = keptOff;
    if (g.keptPerImage > 65535) return fail(GVS_ERR_INVALID, "n_features too large (more than 65535 keys per image)");
    return tab;
}

// cv::resize INTER_LINEAR coefficient tables, float32 exactly as OpenCV computes them
void buildTaps(const gvs_extractor* fe, std::vector<ResizeTap>& taps, int nTaps) {
    taps.assign(nTaps, ResizeTap{0, 0, 0, 0});
    const Geometry& g = fe->g;
    for (int l = 1; l < g.nLevels; l++) {
        const LevelGeom& S = g.lv[l - 1];
        const LevelGeom& D = g.lv[l];
        const double scale_x = 1.0 / ((double)D.w / S.w), scale_y = 1.0 / ((double)D.h / S.h);
        for (int dx = 0; dx < D.w; dx++) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)std::floor(fx);
            fx -= sx;
            bool clampR = false;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; clampR = true; }
            ResizeTap t;
            t.ofs = (short)sx;
            if (clampR) { t.a0 = 2048; t.a1 = 0; t.pad = (short)sx; }
            else { t.a0 = (short)cvRoundF((1.f - fx) * 2048); t.a1 = (short)cvRoundF(fx * 2048); t.pad = (short)(sx + 1); }
            taps[D.tabX + dx] = t;
        }
        for (int dy = 0; dy < D.h; dy++) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            int sy = (int)std::floor(fy);
            fy -= sy;
            ResizeTap t;
            t.ofs = (short)std::min(std::max(sy, 0), S.h - 1);
            t.pad = (short)std::min(std::max(sy + 1, 0), S.h - 1);
            t.a0 = (short)cvRoundF((1.f - fy) * 2048);
            t.a1 = (short)cvRoundF(fy * 2048);
            taps[D.tabY + dy] = t;
        }
    }
}

// cuTensorMapEncodeTiled through the runtime (libcuda is not linked): nullptr when the driver does not offer it
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensorMapEncoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// tile map over pyramid level `level` of every image slot: (x, y rows, image); x counts bytes, or 32-bit words with `words`
// (a box dimension is limited to 256 elements)
bool encodeLevelMap(const gvs_extractor* fe, int level, int boxW, int boxH, CUtensorMap* out, bool words = false) {
    EncodeTiledFn enc = tensorMapEncoder();
    if (!enc) return false;
    const LevelGeom& lv = fe->g.lv[level];
    const cuuint64_t dims[3] = {(cuuint64_t)(words ? lv.pitch / 4 : lv.pitch), (cuuint64_t)lv.h, (cuuint64_t)fe->maxImages};
    const cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)fe->g.slabBytes};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u}, es[3] = {1u, 1u, 1u};
    return enc(out, words ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, fe->slab + lv.off, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Plan of the fused pyramid (pyramid.cu: pyr_fused_kernel).  Groups of consecutive levels; per group the tiles of its coarsest
// level and, walking the resize tables backwards, the regions of the finer levels every tile depends on.  Returns false (legacy
// per-level kernels) when the geometry does not fit: a level that is not "narrow" (scale factor above ~1.75), a halo that does
// not cover the neighbour's first column, or a region larger than a TMA box.
bool buildPyrPlan(gvs_extractor* fe, const std::vector<ResizeTap>& taps, std::vector<PyrSpan>& spans) {
    const Geometry& g = fe->g;
    PyrPlan& P = fe->pyrPlan;
    P.nGroups = 0;
    spans.clear();
    const int L = g.nLevels;
    if (L < 2 || getenv("GVS_NO_FUSED_PYR")) return false;
    for (int l = 1; l < L; l++)
        if (!((fe->pyrNarrowMask >> l) & 1u)) return false;
    auto up = [](int v, int a) { return (v + a - 1) / a * a; };
    int base = 0;
    while (base < L - 1) {
        if (P.nGroups == kPyrMaxGroups) { P.nGroups = 0; return false; }
        const bool first = base == 0;
        const int nOut = std::min(first ? 3 : kPyrMaxFuse, L - 1 - base), top = base + nOut;
        PyrGroup G{};
        G.base = base; G.nOut = nOut;
        int TW = first ? 128 : 64, TH = first ? 64 : 32;   // tile of the group's coarsest level (measured: larger tiles amortise the TMA wait)
        if (const char* e = getenv("GVS_PYR_TILE")) {        // tuning knob: "TWa,THa,TWb,THb"
            int v[4] = {TW, TH, 64, 32};
            if (sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) >= 2) { TW = first ? v[0] : v[2]; TH = first ? v[1] : v[3]; }
            TW = std::max(TW & ~3, 4); TH = std::max(TH, 2);
        }
        for (;; TW = std::max(TW / 2, 4), TH = std::max(TH / 2, 2)) {
            auto build = [&](bool isX, int T, std::vector<PyrSpan>& out) -> bool {
                out.clear();
                const int dimTop = isX ? g.lv[top].w : g.lv[top].h, n = (dimTop + T - 1) / T;
                for (int i = 0; i < n; i++) {
                    PyrSpan sp{};
                    sp.lo[nOut] = (short)(i * T); sp.hi[nOut] = (short)std::min(i * T + T, dimTop);
                    for (int k = nOut - 1; k >= 0; k--) {   // level base + k feeds level base + k + 1
                        const LevelGeom& D = g.lv[base + k + 1];
                        const int tab = isX ? D.tabX : D.tabY, dimSrc = isX ? g.lv[base + k].w : g.lv[base + k].h;
                        int lo = taps[tab + sp.lo[k + 1]].ofs, hi = taps[tab + sp.hi[k + 1] - 1].pad + 1;
                        if (isX) lo &= k == 0 ? ~15 : ~3;
                        if (i == n - 1 && k > 0) hi = dimSrc;   // the last tile of a produced level runs to its end
                        sp.lo[k] = (short)lo; sp.hi[k] = (short)std::min(hi, dimSrc);
                    }
                    out.push_back(sp);
                }
                for (int k = 1; k <= nOut; k++) {
                    const int dim = isX ? g.lv[base + k].w : g.lv[base + k].h;
                    for (int i = 0; i < n; i++) {
                        PyrSpan& sp = out[i];
                        const unsigned nG = (unsigned)((sp.hi[k] - sp.lo[k] + 3) >> 2);
                        sp.magic[k] = nG > 1 ? 0xffffffffu / nG + 1u : 0u;
                        sp.ownLo[k] = (short)(i == 0 ? 0 : sp.lo[k]);
                        sp.ownHi[k] = (short)(i == n - 1 ? dim : out[i + 1].lo[k]);
                        if (i == 0 && sp.lo[k] != 0) return false;
                        if (sp.ownHi[k] > sp.hi[k] || sp.ownHi[k] < sp.ownLo[k]) return false;           // the halo must reach the neighbour's first column
                        if (isX && ((sp.ownLo[k] & 3) || (i < n - 1 && (sp.ownHi[k] & 3)))) return false;   // whole 4-px store groups
                        if (i == n - 1 && sp.hi[k] != dim) return false;
                    }
                }
                return true;
            };
            std::vector<PyrSpan> sx, sy;
            if (!build(true, TW, sx) || !build(false, TH, sy)) return false;
            int wMax[kPyrMaxFuse + 1] = {}, hMax[kPyrMaxFuse + 1] = {};
            for (const PyrSpan& sp : sx) for (int k = 0; k <= nOut; k++) wMax[k] = std::max(wMax[k], sp.hi[k] - sp.lo[k]);
            for (const PyrSpan& sp : sy) for (int k = 0; k <= nOut; k++) hMax[k] = std::max(hMax[k], sp.hi[k] - sp.lo[k]);
            G.boxW = up(wMax[0] + 16, 16); G.boxH = hMax[0];
            if ((G.boxW > 256 || G.boxH > 256) && (TW > 4 || TH > 2)) continue;   // smaller tiles
            if (G.boxW > 256 || G.boxH > 256) return false;
            // two shared-memory regions used in turn: level k is read while level k + 1 is written, level k - 1 is dead by then
            G.bufPitch[0] = G.boxW;
            int size[2] = {up(G.boxW * G.boxH + 16, 128), 0};
            for (int k = 1; k < nOut; k++) {
                G.bufPitch[k] = up(wMax[k] + 16, 4);
                size[k & 1] = std::max(size[k & 1], up(G.bufPitch[k] * hMax[k] + 16, 128));
            }
            for (int k = 0; k < nOut; k++) G.bufOff[k] = (k & 1) ? size[0] : 0;
            const int off = size[0] + size[1];
            G.smemBytes = off + 16;
            if (G.smemBytes > 96 * 1024) return false;
            G.nTx = (int)sx.size(); G.nTy = (int)sy.size();
            G.spanX = (int)spans.size(); spans.insert(spans.end(), sx.begin(), sx.end());
            G.spanY = (int)spans.size(); spans.insert(spans.end(), sy.begin(), sy.end());
            break;
        }
        if (!encodeLevelMap(fe, base, G.boxW, G.boxH, &fe->pyrMaps[P.nGroups])) { P.nGroups = 0; return false; }
        P.grp[P.nGroups++] = G;
        base = top;
    }
    return true;
}

void destroyExtractor(gvs_extractor* fe) {
    if (!fe) return;
    DeviceGuard deviceGuard(fe->device);
    if (fe->ownStream) cudaStreamSynchronize(fe->ownStream);
    devFree(fe->slab); devFree(fe->blurSlab); devFree(fe->tight); devFree(fe->mapX); devFree(fe->mapY); devFree(fe->xWeights); devFree(fe->xSelectors); devFree(fe->xBase); devFree(fe->pyrSpans); devFree(fe->taps); devFree(fe->fastItems); devFree(fe->blurItems);
    devFree(fe->recs); devFree(fe->nCand); devFree(fe->cellCnt); devFree(fe->kept); devFree(fe->keptCnt);
    devFree(fe->imgOff); devFree(fe->orderedDbg); devFree(fe->sortedDbg); devFree(fe->momentsDbg); devFree(fe->selScratch);
    devFree(fe->workCounter); devFree(fe->status); devFree(fe->kps); devFree(fe->desc); devFree(fe->auxKey);
    if (fe->hOff) cudaFreeHost(fe->hOff);
    if (fe->hStatus) cudaFreeHost(fe->hStatus);
    for (auto& kv : fe->graphs) cudaGraphExecDestroy(kv.second);
    if (fe->doneEvent) cudaEventDestroy(fe->doneEvent);
    if (fe->evPyr) cudaEventDestroy(fe->evPyr);
    if (fe->evBlur) cudaEventDestroy(fe->evBlur);
    if (fe->auxStream) cudaStreamDestroy(fe->auxStream);
    for (auto& e : fe->ev) if (e) cudaEventDestroy(e);
    if (fe->ownStream) cudaStreamDestroy(fe->ownStream);
    delete fe;
}

// the whole extraction pipeline for image slots [0, n) (src/FeatureExtractor.cpp:481-533 for each image): stream work only
static int enqueueExtract(gvs_extractor* fe, int n, cudaStream_t s) {
    const Geometry& g = fe->g;
    fe->launches = 0;
    const bool T = fe->timing;
    GVS_CUDA_OK(cudaMemsetAsync(fe->status, 0, sizeof(int), s));
    if (T) cudaEventRecord(fe->ev[0], s);
    if (fe->pyrPlan.nGroups > 0)
        launchPyramidFused(g, fe->slab, fe->taps, fe->xWeights, fe->xSelectors, fe->xBase, fe->pyrPlan, fe->pyrSpans, fe->pyrMaps, n, s, &fe->launches);
    else
        launchPyramid(g, fe->slab, fe->taps, fe->xWeights, fe->xSelectors, fe->xBase, fe->pyrNarrowMask, n, s, &fe->launches);
    if (T) cudaEventRecord(fe->ev[1], s);
    launchFast(g, fe->slab, fe->fastItems, fe->nFastItems, n, fe->recs, fe->nCand, fe->cellCnt, fe->status, fe->fastTma ? fe->fastMaps : nullptr, s,
               &fe->launches);
    if (T) cudaEventRecord(fe->ev[2], s);
    launchSelect(g, n, fe->recs, fe->nCand, fe->cellCnt, fe->kept, fe->keptCnt, fe->orderedDbg, nullptr, fe->selScratch,
                 fe->selScratchPerCta, fe->selCtas, fe->workCounter, fe->status, s, &fe->launches);
    launchOffsets(g, n, fe->keptCnt, fe->imgOff, fe->totalCap, fe->status, s, &fe->launches);
    // the 7x7 blur is fused into the descriptor kernel (evaluated only at the sample positions of the kept keys)
    if (T) { cudaEventRecord(fe->ev[3], s); cudaEventRecord(fe->ev[4], s); }
    launchDescribeFused(g, fe->slab, n, fe->kept, fe->keptCnt, fe->imgOff, fe->kps, fe->desc, fe->auxKey, fe->momentsDbg,
                        fe->totalCap, s, &fe->launches);
    if (T) cudaEventRecord(fe->ev[5], s);
    GVS_CUDA_OK(cudaMemcpyAsync(fe->hOff, fe->imgOff, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(fe->hStatus, fe->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    return GVS_OK;
}

// Captured graphs freeze the kernel arguments: drop them when a lazily allocated debug tap changes a pointer.
static void invalidateGraphs(gvs_extractor* fe) {
    for (auto& kv : fe->graphs) cudaGraphExecDestroy(kv.second);
    fe->graphs.clear();
    fe->graphLaunches.clear();
}

// Small batches are launch bound (13 launches per side for a frame that needs ~0.2 ms of kernels): the pipeline of a batch
// size is captured once into a CUDA graph and replayed.  All kernel arguments are fixed per handle and batch size.
constexpr int kGraphMaxImages = 8;

// rerun = the same resident input is processed again (a parity tap was switched on): the results are the same, so the batch
// generation, which the matcher compares, stays
int runExtract(gvs_extractor* fe, int n, bool rerun = false) {
    NvtxScope nvtx("gvs:extract");
    cudaStream_t s = fe->stream;
    static const bool noGraph = getenv("GVS_NO_GRAPH") != nullptr;
    // (a caller's stream that is itself being captured takes the plain launches: they become part of the caller's graph)
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess) { cudaGetLastError(); capturing = cudaStreamCaptureStatusNone; }
    if (!fe->timing && !noGraph && n <= kGraphMaxImages && capturing == cudaStreamCaptureStatusNone) {
        auto it = fe->graphs.find(n);
        if (it == fe->graphs.end()) {
            cudaGraph_t graph = nullptr;
            GVS_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueueExtract(fe, n, s);
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            GVS_CUDA_OK(e);
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            GVS_CUDA_OK(ei);
            it = fe->graphs.emplace(n, exec).first;
            fe->graphLaunches[n] = fe->launches;
        }
        GVS_CUDA_OK(cudaGraphLaunch(it->second, s));
        fe->launches = fe->graphLaunches[n];
    } else {
        const int rc = enqueueExtract(fe, n, s);
        if (rc) return rc;
    }
    fe->blurValid = false;
    GVS_CUDA_OK(cudaEventRecord(fe->doneEvent, s));
    GVS_CUDA_OK(cudaGetLastError());
    fe->lastN = n;
    if (!rerun) ++fe->generation;
    fe->pending = true;
    fe->launchesTotal += fe->launches;
    return GVS_OK;
}

}  // namespace

namespace gvs {
int syncExtract(gvs_extractor* fe) {
    DeviceGuard deviceGuard(fe->device);
    GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
    if (fe->pending) {
        fe->pending = false;
        if (fe->timing) {
            float* t[5] = {&fe->times.pyramid_ms, &fe->times.fast_ms, &fe->times.select_ms, &fe->times.blur_ms,
                           &fe->times.describe_ms};
            for (int i = 0; i < 5; i++) cudaEventElapsedTime(t[i], fe->ev[i], fe->ev[i + 1]);
            fe->times.launches = fe->launches;
        }
    }
    const int st = *fe->hStatus;
    if (st & kStatCandOverflow) return fail(GVS_ERR_CAPACITY, "FAST candidate buffer overflow (image denser than 1 corner per 8 px)");
    if (st & kStatKeptOverflow) return fail(GVS_ERR_CAPACITY, "kept-key staging overflow (ssc returned more than 2*quota+64 keys)");
    if (st & kStatOutputOverflow) return fail(GVS_ERR_CAPACITY, "dense output overflow");
    return GVS_OK;
}
}  // namespace gvs

namespace {

int checkBatch(gvs_extractor* fe, int n) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    if (n < 1 || n > fe->maxImages) return fail(GVS_ERR_INVALID, "n_images out of range for this handle");
    return GVS_OK;
}

}  // namespace

namespace gvs {
int joinRight(gvs_matcher* m) {  // make L's stream wait for R's last batch
    if (m->R != m->L && m->R->stream != m->L->stream) {
        GVS_CUDA_OK(cudaEventRecord(m->evR, m->R->stream));
        GVS_CUDA_OK(cudaStreamWaitEvent(m->L->stream, m->evR, 0));
    }
    return GVS_OK;
}

// The matcher's kernels run on L's stream and read R's keys, descriptors and pyramid: whatever is queued on R's stream afterwards
// (the next batch's upload and extraction) must wait for them.
int holdRight(gvs_matcher* m) {
    {   // remember the end of the matcher's own work for gvs_matcher_destroy
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(m->L->stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
        if (cap == cudaStreamCaptureStatusNone) GVS_CUDA_OK(cudaEventRecord(m->evLast, m->L->stream));
    }
    if (m->R != m->L && m->R->stream != m->L->stream) {
        // (not while the caller captures L's stream into a graph: a fork that is never joined would invalidate the capture, and the
        // caller's graph already orders whole batches)
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(m->L->stream, &capturing) != cudaSuccess) { cudaGetLastError(); capturing = cudaStreamCaptureStatusNone; }
        if (capturing != cudaStreamCaptureStatusNone) return GVS_OK;
        GVS_CUDA_OK(cudaEventRecord(m->evHold, m->L->stream));
        GVS_CUDA_OK(cudaStreamWaitEvent(m->R->stream, m->evHold, 0));
    }
    return GVS_OK;
}

int growScratch(gvs_matcher* m, size_t bytes) {
    if (bytes <= m->pjBytes) return GVS_OK;
    GVS_CUDA_OK(cudaStreamSynchronize(m->L->stream));
    devFree(m->pj); m->pj = nullptr; m->pjBytes = 0;
    GVS_CUDA_OK(devAlloc(&m->pj, bytes));
    m->pjBytes = bytes;
    return GVS_OK;
}
}  // namespace gvs

extern "C" {

const char* gvs_last_error(void) { return g_lastError.c_str(); }
int gvs_version(void) { return 200; }
void gvs_default_params(gvs_extractor_params* p) {
    if (!p) return;
    p->n_features = 2000; p->n_levels = 8; p->im_scale = 1.2f; p->edge_threshold = 19; p->patch_size = 31;
    p->max_fast_threshold = 20; p->min_fast_threshold = 7;
}
int gvs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gvs_extractor_tma_stages(const gvs_extractor* fe) {
    if (!fe) return fail(GVS_ERR_INVALID, "null handle");
    return (fe->pyrPlan.nGroups > 0 ? 1 : 0) | (fe->fastTma ? 2 : 0);
}

int gvs_debug_check_guards(int* n_checked) {
    std::lock_guard<std::mutex> lock(g_guardMutex);
    if (n_checked) *n_checked = (int)g_guards.size();
    std::vector<uint8_t> tail(kGuardBytes);
    int bad = 0;
    for (auto& g : g_guards) {
        if (cudaMemcpy(tail.data(), static_cast<uint8_t*>(g.first) + g.second, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        for (size_t i = 0; i < kGuardBytes; i++)
            if (tail[i] != 0xC3) {
                ++bad;
                std::fprintf(stderr, "gvs guard: %zu-byte device buffer %p overwritten at +%zu past its end\n", g.second, g.first, i);
                break;
            }
    }
    return bad;
}

int gvs_extractor_create(gvs_extractor** out, const gvs_extractor_params* params, int width, int height, int max_images,
                         int device) {
    if (!out || !params) return fail(GVS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (params->n_levels < 1 || params->n_levels > kMaxLevels) return fail(GVS_ERR_INVALID, "n_levels must be in [1, 12]");
    if (params->n_features < 2 || params->im_scale <= 1.0f) return fail(GVS_ERR_INVALID, "bad n_features / im_scale");
    if (params->patch_size != 31) return fail(GVS_ERR_INVALID, "patch_size must be 31 (halfPatchSize is fixed to 15 in the reference)");
    if (params->edge_threshold < 19) return fail(GVS_ERR_INVALID, "edge_threshold must be >= 19 (descriptor reach)");
    if (params->min_fast_threshold < 1 || params->max_fast_threshold < params->min_fast_threshold || params->max_fast_threshold > 254)
        return fail(GVS_ERR_INVALID, "bad FAST thresholds");
    if (max_images < 1 || max_images > 65535) return fail(GVS_ERR_INVALID, "max_images must be in [1, 65535]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GVS_ERR_CUDA, "no CUDA device: this library has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(GVS_ERR_INVALID, "bad device index");
    DeviceGuard deviceGuard(device);
    gvs_extractor* fe = new gvs_extractor();
    fe->device = device; fe->params = *params; fe->width = width; fe->height = height; fe->maxImages = max_images;
    const int nTaps = buildGeometry(fe);
    if (nTaps < 0) { delete fe; return nTaps; }
    const Geometry& g = fe->g;
    const size_t B = (size_t)max_images;
#define GVS_TRY(expr)                                                                \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            int rc = fail(GVS_ERR_CUDA, #expr, cudaGetErrorString(_e));              \
            destroyExtractor(fe);                                                    \
            return rc;                                                               \
        }                                                                            \
    } while (0)
    GVS_TRY(cudaStreamCreateWithFlags(&fe->ownStream, cudaStreamNonBlocking));
    fe->stream = fe->ownStream;
    GVS_TRY(cudaStreamCreateWithFlags(&fe->auxStream, cudaStreamNonBlocking));
    GVS_TRY(cudaEventCreateWithFlags(&fe->doneEvent, cudaEventDisableTiming));
    GVS_TRY(cudaEventCreateWithFlags(&fe->evPyr, cudaEventDisableTiming));
    GVS_TRY(cudaEventCreateWithFlags(&fe->evBlur, cudaEventDisableTiming));
    for (auto& e : fe->ev) GVS_TRY(cudaEventCreate(&e));
    GVS_TRY(devAlloc(&fe->slab, B * g.slabBytes + 256));
    GVS_TRY(cudaMemset(fe->slab, 0, B * g.slabBytes + 256));
    {
        std::vector<ResizeTap> taps;
        buildTaps(fe, taps, std::max(nTaps, 4));
        GVS_TRY(devAlloc(&fe->taps, taps.size()));
        GVS_TRY(cudaMemcpy(fe->taps, taps.data(), taps.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
        // horizontal taps regrouped for pyr_down_kernel: 4 destination columns share one 12-byte source window
        const size_t nGroups = taps.size() / 4 + 1;
        std::vector<unsigned> w(4 * nGroups, 0), sl(4 * nGroups, 0);
        std::vector<unsigned short> base(nGroups, 0);
        for (int l = 1; l < g.nLevels; l++) {
            const LevelGeom& D = g.lv[l];
            // narrow level: the pairs (sx, sx+1) of the 4 columns of every group lie within 8 bytes of the first column's sx
            bool narrow = true;
            for (int dx0 = 0; dx0 < D.w; dx0 += 4) {
                const int last = std::min(dx0 + 3, D.w - 1);
                if (taps[D.tabX + last].pad - taps[D.tabX + dx0].ofs > 7) narrow = false;
            }
            if (narrow) fe->pyrNarrowMask |= 1u << l;
            for (int dx0 = 0; dx0 < D.w; dx0 += 4) {
                const size_t gi = (size_t)(D.tabX + dx0) / 4;
                const int b = narrow ? taps[D.tabX + dx0].ofs : (taps[D.tabX + dx0].ofs & ~3);
                base[gi] = (unsigned short)b;
                for (int i = 0; i < 4 && dx0 + i < D.w; i++) {
                    const ResizeTap& t = taps[D.tabX + dx0 + i];
                    const int o = t.ofs - b, o1 = t.pad - b;
                    if (o < 0 || o1 > 11 || o1 < o) { int rc = fail(GVS_ERR_INVALID, "im_scale too large for the pyramid kernel's 12-byte source window"); destroyExtractor(fe); return rc; }
                    const int hi = narrow ? 0 : o >= 4;
                    w[4 * gi + i] = (unsigned)(unsigned short)t.a0 | ((unsigned)(unsigned short)t.a1 << 16);
                    sl[4 * gi + i] = (unsigned)(o - 4 * hi) | ((unsigned)(o1 - 4 * hi) << 4) | ((unsigned)hi << 16);
                }
            }
        }
        GVS_TRY(devAlloc(&fe->xWeights, nGroups));
        GVS_TRY(devAlloc(&fe->xSelectors, nGroups));
        GVS_TRY(devAlloc(&fe->xBase, nGroups));
        GVS_TRY(cudaMemcpy(fe->xWeights, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        GVS_TRY(cudaMemcpy(fe->xSelectors, sl.data(), sl.size() * 4, cudaMemcpyHostToDevice));
        GVS_TRY(cudaMemcpy(fe->xBase, base.data(), base.size() * 2, cudaMemcpyHostToDevice));
        std::vector<PyrSpan> spans;
        if (buildPyrPlan(fe, taps, spans)) {
            GVS_TRY(devAlloc(&fe->pyrSpans, spans.size()));
            GVS_TRY(cudaMemcpy(fe->pyrSpans, spans.data(), spans.size() * sizeof(PyrSpan), cudaMemcpyHostToDevice));
        }
    }
    {
        std::vector<FastItem> items;
        for (int l = 0; l < g.nLevels; l++) {
            fe->g.lv[l].fastItemBase = (int)items.size();
            // a cell row is cut into the fewest strips of at most kFastCellsPerItem cells, of equal length +- 1
            const int nCols = g.lv[l].nCols, nStrips = (nCols + kFastCellsPerItem - 1) / kFastCellsPerItem;
            for (int iR = 0; iR < g.lv[l].nRows; iR++)
                for (int k = 0; k < nStrips; k++) {
                    const int c0 = (int)((long long)k * nCols / nStrips), c1 = (int)((long long)(k + 1) * nCols / nStrips);
                    items.push_back(FastItem{(short)l, (short)iR, (short)c0, (short)(c1 - c0)});
                }
        }
        fe->nFastItems = (int)items.size();
        {   // TMA tile maps of every level for the FAST kernel's tile staging (box = the largest strip tile with its 3-px halo)
            int bw = 0, bh = 0;
            fe->fastTma = !getenv("GVS_NO_FAST_TMA") && fastTmaBox(fe->g, &bw, &bh);
            for (int l = 0; l < g.nLevels && fe->fastTma; l++) fe->fastTma = encodeLevelMap(fe, l, bw, bh, &fe->fastMaps[l], true);
        }
        GVS_TRY(devAlloc(&fe->fastItems, items.size()));
        if (!items.empty()) GVS_TRY(cudaMemcpy(fe->fastItems, items.data(), items.size() * sizeof(FastItem), cudaMemcpyHostToDevice));
        std::vector<BlurItem> bi;
        for (int l = 0; l < g.nLevels; l++)
            for (int ty = 0; ty * kBlurTH < g.lv[l].h; ty++)
                for (int tx = 0; tx * kBlurTW < g.lv[l].w; tx++) bi.push_back(BlurItem{(short)l, (short)tx, (short)ty, 0});
        fe->nBlurItems = (int)bi.size();
        GVS_TRY(devAlloc(&fe->blurItems, bi.size()));
        GVS_TRY(cudaMemcpy(fe->blurItems, bi.data(), bi.size() * sizeof(BlurItem), cudaMemcpyHostToDevice));
    }
    GVS_TRY(devAlloc(&fe->recs, B * g.candPerImage));
    GVS_TRY(devAlloc(&fe->nCand, B * kMaxLevels));
    GVS_TRY(devAlloc(&fe->cellCnt, B * std::max(g.cellsPerImage, 1)));
    GVS_TRY(devAlloc(&fe->kept, B * g.keptPerImage));
    GVS_TRY(devAlloc(&fe->keptCnt, B * kMaxLevels));
    GVS_TRY(cudaMemset(fe->keptCnt, 0, B * kMaxLevels * sizeof(int)));
    GVS_TRY(devAlloc(&fe->imgOff, B + 1));
    GVS_TRY(devAlloc(&fe->workCounter, 2));
    GVS_TRY(devAlloc(&fe->status, 1));
    fe->selScratchPerCta = selectScratchBytes(g);
    fe->selCtas = std::min(selectMaxCtas(device), (int)B * g.nLevels);
    GVS_TRY(devAlloc(&fe->selScratch, fe->selScratchPerCta * fe->selCtas));
    fe->totalCap = (int)std::min<size_t>(B * g.keptPerImage, (size_t)1 << 30);
    GVS_TRY(devAlloc(&fe->kps, (size_t)fe->totalCap));
    GVS_TRY(devAlloc(&fe->desc, (size_t)fe->totalCap * 32));
    GVS_TRY(devAlloc(&fe->auxKey, (size_t)fe->totalCap));
    GVS_TRY(cudaMallocHost(reinterpret_cast<void**>(&fe->hOff), sizeof(int) * (B + 1)));
    GVS_TRY(cudaMallocHost(reinterpret_cast<void**>(&fe->hStatus), sizeof(int)));
    *fe->hStatus = 0;
    std::memset(fe->hOff, 0, sizeof(int) * (B + 1));
    // the initialising memsets above ran on the legacy default stream, which the handle's non-blocking streams do not wait for
    // (only the legacy stream, not the device: another host thread may be capturing its handle's stream into a CUDA graph, and a
    // device-wide synchronisation is not permitted then -- found by tests/test_gpu_multi.py)
    GVS_TRY(cudaStreamSynchronize(cudaStreamLegacy));
#undef GVS_TRY
    *out = fe;
    return GVS_OK;
}

int gvs_extractor_destroy(gvs_extractor* fe) {
    destroyExtractor(fe);
    return GVS_OK;
}

int gvs_extractor_tables(const gvs_extractor* fe, float* scale, float* inv_scale, float* sigma, float* inv_sigma,
                         int32_t* scaled_patch, int32_t* feature_per_level) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    for (int i = 0; i < fe->g.nLevels; i++) {
        if (scale) scale[i] = fe->scale[i];
        if (inv_scale) inv_scale[i] = fe->invScale[i];
        if (sigma) sigma[i] = fe->sigma[i];
        if (inv_sigma) inv_sigma[i] = fe->invSigma[i];
        if (scaled_patch) scaled_patch[i] = fe->patch[i];
        if (feature_per_level) feature_per_level[i] = fe->quota[i];
    }
    return GVS_OK;
}

int gvs_tables_from_params(const gvs_extractor_params* params, float* scale, float* inv_scale, float* sigma, float* inv_sigma,
                           int32_t* scaled_patch, int32_t* feature_per_level) {
    if (!params) return fail(GVS_ERR_INVALID, "null argument");
    if (params->n_levels < 1 || params->n_levels > kMaxLevels) return fail(GVS_ERR_INVALID, "n_levels must be in [1, 12]");
    if (params->n_features < 2 || params->im_scale <= 1.0f) return fail(GVS_ERR_INVALID, "bad n_features / im_scale");
    std::vector<float> sc, isc, sg, isg;
    std::vector<int> pt, qt;
    buildTables(*params, sc, isc, sg, isg, pt, qt);
    for (int i = 0; i < params->n_levels; i++) {
        if (scale) scale[i] = sc[i];
        if (inv_scale) inv_scale[i] = isc[i];
        if (sigma) sigma[i] = sg[i];
        if (inv_sigma) inv_sigma[i] = isg[i];
        if (scaled_patch) scaled_patch[i] = pt[i];
        if (feature_per_level) feature_per_level[i] = qt[i];
    }
    return GVS_OK;
}

int gvs_extractor_level_size(const gvs_extractor* fe, int level, int* width, int* height) {
    if (!fe || level < 0 || level >= fe->g.nLevels) return fail(GVS_ERR_INVALID, "bad level");
    if (width) *width = fe->g.lv[level].w;
    if (height) *height = fe->g.lv[level].h;
    return GVS_OK;
}

int gvs_extractor_set_stream(gvs_extractor* fe, void* cuda_stream) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    DeviceGuard deviceGuard(fe->device);
    GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
    fe->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : fe->ownStream;
    return GVS_OK;
}

int gvs_extractor_input(gvs_extractor* fe, uint8_t** dev, int* pitch, size_t* image_stride) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    if (dev) *dev = fe->slab + fe->g.lv[0].off;
    if (pitch) *pitch = fe->g.lv[0].pitch;
    if (image_stride) *image_stride = fe->g.slabBytes;
    return GVS_OK;
}

int gvs_extract_batch_resident(gvs_extractor* fe, int n_images) {
    int rc = checkBatch(fe, n_images);
    if (rc) return rc;
    DeviceGuard deviceGuard(fe->device);
    return runExtract(fe, n_images);
}

int gvs_extract_batch_host(gvs_extractor* fe, const uint8_t* images, size_t image_stride, int row_stride, int n_images) {
    int rc = checkBatch(fe, n_images);
    if (rc) return rc;
    if (!images || row_stride < fe->width) return fail(GVS_ERR_INVALID, "bad image pointer / stride");
    DeviceGuard deviceGuard(fe->device);
    const LevelGeom& l0 = fe->g.lv[0];
    if (row_stride == fe->width && image_stride == (size_t)fe->width * fe->height) {
        // tightly packed batch: one contiguous DMA into a staging buffer, then a repack kernel into the pitched store
        const size_t need = (size_t)fe->maxImages * fe->width * fe->height + 16;
        if (fe->tightBytes < need) {
            GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
            devFree(fe->tight); fe->tight = nullptr; fe->tightBytes = 0;
            GVS_CUDA_OK(devAlloc(&fe->tight, need));
            fe->tightBytes = need;
        }
        GVS_CUDA_OK(cudaMemcpyAsync(fe->tight, images, (size_t)n_images * image_stride, cudaMemcpyHostToDevice, fe->stream));
        int launches = 0;
        launchRepack(fe->g, fe->tight, fe->slab, n_images, fe->stream, &launches);
        fe->launchesTotal += launches;
        return runExtract(fe, n_images);
    }
    for (int i = 0; i < n_images; i++)
        GVS_CUDA_OK(cudaMemcpy2DAsync(fe->slab + (size_t)i * fe->g.slabBytes + l0.off, l0.pitch, images + (size_t)i * image_stride,
                                      row_stride, fe->width, fe->height, cudaMemcpyHostToDevice, fe->stream));
    return runExtract(fe, n_images);
}

int gvs_extractor_set_rectify_maps(gvs_extractor* fe, const float* map_x, const float* map_y) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    DeviceGuard deviceGuard(fe->device);
    GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
    if (!map_x || !map_y) {
        devFree(fe->mapX); devFree(fe->mapY);
        fe->mapX = fe->mapY = nullptr;
        return GVS_OK;
    }
    const size_t n = (size_t)fe->width * fe->height;
    if (!fe->mapX) GVS_CUDA_OK(devAlloc(&fe->mapX, n));
    if (!fe->mapY) GVS_CUDA_OK(devAlloc(&fe->mapY, n));
    GVS_CUDA_OK(cudaMemcpy(fe->mapX, map_x, n * sizeof(float), cudaMemcpyHostToDevice));
    GVS_CUDA_OK(cudaMemcpy(fe->mapY, map_y, n * sizeof(float), cudaMemcpyHostToDevice));
    return GVS_OK;
}

int gvs_extract_batch_host_raw(gvs_extractor* fe, const uint8_t* images, int channels, int n_images) {
    int rc = checkBatch(fe, n_images);
    if (rc) return rc;
    if (!images || (channels != 1 && channels != 3)) return fail(GVS_ERR_INVALID, "raw images must be tightly packed, 1 or 3 channels");
    DeviceGuard deviceGuard(fe->device);
    const size_t per = (size_t)fe->width * fe->height * channels;
    const size_t need = (size_t)fe->maxImages * fe->width * fe->height * 3 + 16;
    if (fe->tightBytes < need) {
        GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
        devFree(fe->tight); fe->tight = nullptr; fe->tightBytes = 0;
        GVS_CUDA_OK(devAlloc(&fe->tight, need));
        fe->tightBytes = need;
    }
    GVS_CUDA_OK(cudaMemcpyAsync(fe->tight, images, per * n_images, cudaMemcpyHostToDevice, fe->stream));
    int launches = 0;
    launchPrepare(fe->g, fe->tight, channels, fe->mapX, fe->mapY, fe->slab, n_images, fe->stream, &launches);
    fe->launchesTotal += launches;
    return runExtract(fe, n_images);
}

int gvs_extractor_sync(gvs_extractor* fe) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    return syncExtract(fe);
}

int gvs_extract_offsets(gvs_extractor* fe, int32_t* offsets) {
    if (!fe || !offsets) return fail(GVS_ERR_INVALID, "null argument");
    if (fe->lastN <= 0) return fail(GVS_ERR_STATE, "no batch has been extracted");
    int rc = syncExtract(fe);
    if (rc) return rc;
    std::memcpy(offsets, fe->hOff, sizeof(int) * (fe->lastN + 1));
    return GVS_OK;
}

int gvs_extract_download(gvs_extractor* fe, int first, int count, gvs_keypoint* kps, uint8_t* desc, int cap) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    if (fe->lastN <= 0) return fail(GVS_ERR_STATE, "no batch has been extracted");
    if (first < 0 || count < 0 || first + count > fe->lastN) return fail(GVS_ERR_INVALID, "image range out of the last batch");
    int rc = syncExtract(fe);
    if (rc) return rc;
    const int o0 = fe->hOff[first], n = fe->hOff[first + count] - o0;
    if (n > cap) return fail(GVS_ERR_CAPACITY, "caller buffer too small for the keypoints of this range");
    if (n > 0) {
        if (kps) GVS_CUDA_OK(cudaMemcpyAsync(kps, fe->kps + o0, (size_t)n * sizeof(gvs_keypoint), cudaMemcpyDeviceToHost, fe->stream));
        if (desc) GVS_CUDA_OK(cudaMemcpyAsync(desc, fe->desc + (size_t)o0 * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, fe->stream));
        GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
    }
    return GVS_OK;
}

int gvs_extract(gvs_extractor* fe, const uint8_t* gray, int width, int height, int stride, gvs_keypoint* out, uint8_t* desc,
                int cap, int* n) {
    if (!fe || !gray || !n) return fail(GVS_ERR_INVALID, "null argument");
    if (width != fe->width || height != fe->height) return fail(GVS_ERR_INVALID, "image size differs from the handle's");
    int rc = gvs_extract_batch_host(fe, gray, (size_t)stride * height, stride, 1);
    if (rc) return rc;
    int off[2];
    rc = gvs_extract_offsets(fe, off);
    if (rc) return rc;
    *n = off[1];
    if (*n == 0) return GVS_OK;  // outputs untouched (src/FeatureExtractor.cpp:498-499)
    return gvs_extract_download(fe, 0, 1, out, desc, cap);
}

int gvs_extractor_enable_timing(gvs_extractor* fe, int on) {
    if (!fe) return fail(GVS_ERR_INVALID, "null extractor");
    fe->timing = on != 0;
    return GVS_OK;
}
int gvs_extractor_stage_times(gvs_extractor* fe, gvs_stage_times* t) {
    if (!fe || !t) return fail(GVS_ERR_INVALID, "null argument");
    *t = fe->times;
    return GVS_OK;
}
int gvs_extractor_launch_count(gvs_extractor* fe, int64_t* launches) {
    if (!fe || !launches) return fail(GVS_ERR_INVALID, "null argument");
    *launches = fe->launchesTotal;
    return GVS_OK;
}

// ---- parity taps
int gvs_debug_level_image(gvs_extractor* fe, int image, int level, int blurred, uint8_t* out) {
    if (!fe || !out || level < 0 || level >= fe->g.nLevels || image < 0 || image >= fe->maxImages)
        return fail(GVS_ERR_INVALID, "bad argument");
    DeviceGuard deviceGuard(fe->device);
    const LevelGeom& lv = fe->g.lv[level];
    if (blurred && !fe->blurValid && fe->lastN > 0) {  // cv::GaussianBlur of every level, dense (src/FeatureExtractor.cpp:512-515)
        if (!fe->blurSlab) GVS_CUDA_OK(devAlloc(&fe->blurSlab, (size_t)fe->maxImages * fe->g.slabBytes + 256));
        int launches = 0;
        launchBlur(fe->g, fe->slab, fe->blurSlab, fe->blurItems, fe->nBlurItems, fe->lastN, fe->stream, &launches);
        fe->launchesTotal += launches;
        fe->blurValid = true;
    }
    if (blurred && !fe->blurSlab) return fail(GVS_ERR_STATE, "no batch has been extracted");
    const uint8_t* src = (blurred ? fe->blurSlab : fe->slab) + (size_t)image * fe->g.slabBytes + lv.off;
    GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
    GVS_CUDA_OK(cudaMemcpy2D(out, lv.w, src, lv.pitch, lv.w, lv.h, cudaMemcpyDeviceToHost));
    return GVS_OK;
}

static int unpackKeys(const std::vector<unsigned>& v, int n, int32_t* x, int32_t* y, int32_t* score) {
    for (int i = 0; i < n; i++) {
        if (x) x[i] = v[i] & 0xfff;
        if (y) y[i] = (v[i] >> 12) & 0xfff;
        if (score) score[i] = v[i] >> 24;
    }
    return GVS_OK;
}

int gvs_debug_level_candidates(gvs_extractor* fe, int image, int level, int32_t* x, int32_t* y, int32_t* score, int cap, int* n) {
    if (!fe || !n || level < 0 || level >= fe->g.nLevels || image < 0 || image >= fe->lastN) return fail(GVS_ERR_INVALID, "bad argument");
    DeviceGuard deviceGuard(fe->device);
    if (!fe->orderedDbg) {
        // first use: allocate the tap and re-run the last batch so that it gets filled
        GVS_CUDA_OK(cudaStreamSynchronize(fe->stream));
        GVS_CUDA_OK(devAlloc(&fe->orderedDbg, (size_t)fe->maxImages * fe->g.candPerImage));
        invalidateGraphs(fe);
        int rc = runExtract(fe, fe->lastN, true);
        if (rc) return rc;
    }
    int rc = syncExtract(fe);
    if (rc) return rc;
    const LevelGeom& lv = fe->g.lv[level];
    int cnt = 0;
    GVS_CUDA_OK(cudaMemcpy(&cnt, fe->nCand + image * kMaxLevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (cnt > cap) return fail(GVS_ERR_CAPACITY, "caller buffer too small");
    std::vector<unsigned> v(std::max(cnt, 1));
    GVS_CUDA_OK(cudaMemcpy(v.data(), fe->orderedDbg + (size_t)image * fe->g.candPerImage + lv.candOff, sizeof(unsigned) * cnt,
                           cudaMemcpyDeviceToHost));
    return unpackKeys(v, cnt, x, y, score);
}

int gvs_debug_level_kept(gvs_extractor* fe, int image, int level, int32_t* x, int32_t* y, int32_t* score, int cap, int* n) {
    if (!fe || !n || level < 0 || level >= fe->g.nLevels || image < 0 || image >= fe->lastN) return fail(GVS_ERR_INVALID, "bad argument");
    DeviceGuard deviceGuard(fe->device);
    int rc = syncExtract(fe);
    if (rc) return rc;
    const LevelGeom& lv = fe->g.lv[level];
    int cnt = 0;
    GVS_CUDA_OK(cudaMemcpy(&cnt, fe->keptCnt + image * kMaxLevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (cnt > cap) return fail(GVS_ERR_CAPACITY, "caller buffer too small");
    std::vector<unsigned> v(std::max(cnt, 1));
    GVS_CUDA_OK(cudaMemcpy(v.data(), fe->kept + (size_t)image * fe->g.keptPerImage + lv.keptOff, sizeof(unsigned) * cnt,
                           cudaMemcpyDeviceToHost));
    return unpackKeys(v, cnt, x, y, score);
}

int gvs_debug_select(gvs_extractor* fe, int level, int n, const int32_t* x, const int32_t* y, const int32_t* score,
                     int32_t* sorted_idx, int32_t* kept_idx, int cap, int* n_kept) {
    if (!fe || !x || !y || !score || !kept_idx || !n_kept || level < 0 || level >= fe->g.nLevels) return fail(GVS_ERR_INVALID, "bad argument");
    const LevelGeom& lv = fe->g.lv[level];
    if (n < 0 || n > lv.candCap || n > 65535) return fail(GVS_ERR_CAPACITY, "more candidates than the level's record buffer (or 65535)");
    DeviceGuard deviceGuard(fe->device);
    cudaStream_t s = fe->stream;
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    const Geometry& g = fe->g;
    if (!fe->sortedDbg) GVS_CUDA_OK(devAlloc(&fe->sortedDbg, (size_t)fe->maxImages * g.candPerImage));
    if (!fe->orderedDbg) {
        GVS_CUDA_OK(devAlloc(&fe->orderedDbg, (size_t)fe->maxImages * g.candPerImage));
        invalidateGraphs(fe);
    }
    // all candidates in cell 0 with rank = input position, every other cell / level empty
    std::vector<CandRec> recs(std::max(n, 1));
    for (int i = 0; i < n; i++) {
        if (x[i] < 0 || x[i] >= lv.w || y[i] < 0 || y[i] >= lv.h || score[i] < 1 || score[i] > 254) return fail(GVS_ERR_INVALID, "candidate out of range");
        recs[i].w0 = (unsigned)x[i] | ((unsigned)y[i] << 12) | ((unsigned)score[i] << 24);
        recs[i].w1 = (unsigned)i << 16;
    }
    std::vector<unsigned short> cells(std::max(g.cellsPerImage, 1), 0);
    if (lv.nCols * lv.nRows < 1) return fail(GVS_ERR_INVALID, "level has no FAST cells");
    cells[lv.cellBase] = (unsigned short)n;
    std::vector<int> nc(kMaxLevels, 0);
    nc[level] = n;
    GVS_CUDA_OK(cudaMemcpyAsync(fe->recs + lv.candOff, recs.data(), sizeof(CandRec) * n, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(fe->cellCnt, cells.data(), sizeof(unsigned short) * cells.size(), cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(fe->nCand, nc.data(), sizeof(int) * kMaxLevels, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemsetAsync(fe->status, 0, sizeof(int), s));
    int launches = 0;
    launchSelect(g, 1, fe->recs, fe->nCand, fe->cellCnt, fe->kept, fe->keptCnt, fe->orderedDbg, fe->sortedDbg, fe->selScratch,
                 fe->selScratchPerCta, fe->selCtas, fe->workCounter, fe->status, s, &launches);
    fe->launchesTotal += launches;
    int cnt = 0, st = 0;
    GVS_CUDA_OK(cudaMemcpyAsync(&cnt, fe->keptCnt + level, sizeof(int), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(&st, fe->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    GVS_CUDA_OK(cudaGetLastError());
    fe->lastN = 0;  // the batch state of the handle is gone
    ++fe->generation;
    if (st) return fail(GVS_ERR_CAPACITY, "selection overflow");
    *n_kept = cnt;
    if (cnt > cap) return fail(GVS_ERR_CAPACITY, "caller buffer too small");
    std::vector<unsigned> sorted(std::max(n, 1)), kept(std::max(cnt, 1));
    GVS_CUDA_OK(cudaMemcpy(kept.data(), fe->kept + lv.keptOff, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost));
    if (n > lv.quota) {
        GVS_CUDA_OK(cudaMemcpy(sorted.data(), fe->sortedDbg + lv.candOff, sizeof(unsigned) * n, cudaMemcpyDeviceToHost));
        if (sorted_idx) for (int i = 0; i < n; i++) sorted_idx[i] = (int)sorted[i];
    } else if (sorted_idx) {
        for (int i = 0; i < n; i++) sorted_idx[i] = -1;  // the reference does not sort when n <= quota
    }
    // map kept keys back to input indices (x, y are unique per candidate in FAST output; fall back to first match)
    std::vector<int> lut((size_t)lv.w * lv.h, -1);
    for (int i = n - 1; i >= 0; i--) lut[(size_t)y[i] * lv.w + x[i]] = i;
    for (int i = 0; i < cnt; i++) kept_idx[i] = lut[(size_t)((kept[i] >> 12) & 0xfff) * lv.w + (kept[i] & 0xfff)];
    return GVS_OK;
}

// ------------------------------------------------------------------ matcher
int gvs_matcher_create(gvs_matcher** out, gvs_extractor* feL, gvs_extractor* feR, double fx, float baseline, int image_height) {
    if (!out || !feL) return fail(GVS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!feR) feR = feL;  // mono: System::InitializeMonocular passes the left extractor twice (src/System.cpp:32)
    if (feL->device != feR->device) return fail(GVS_ERR_INVALID, "extractors live on different devices");
    // the stereo and projection kernels index both pyramids and both key sets with the LEFT handle's geometry and scale tables
    if (feL->g.nLevels != feR->g.nLevels || feL->width != feR->width || feL->height != feR->height ||
        feL->params.im_scale != feR->params.im_scale || feL->params.edge_threshold != feR->params.edge_threshold ||
        feL->params.patch_size != feR->params.patch_size || feL->g.slabBytes != feR->g.slabBytes)
        return fail(GVS_ERR_INVALID, "left and right extractors differ in geometry (size, levels, scale factor, edge threshold)");
    for (int l = 0; l < feL->g.nLevels; l++)
        if (feL->g.lv[l].w != feR->g.lv[l].w || feL->g.lv[l].h != feR->g.lv[l].h || feL->g.lv[l].pitch != feR->g.lv[l].pitch ||
            feL->g.lv[l].off != feR->g.lv[l].off)
            return fail(GVS_ERR_INVALID, "left and right extractors differ in pyramid layout");
    if (image_height < 1) return fail(GVS_ERR_INVALID, "bad image_height");
    DeviceGuard deviceGuard(feL->device);
    gvs_matcher* m = new gvs_matcher();
    m->L = feL; m->R = feR; m->device = feL->device; m->fx = fx; m->baseline = baseline; m->imageHeight = image_height;
    {   // key grid: 64 columns, ceil(64 / (w/h)) rows (src/FeatureTracker.cpp:30-34); sized for this handle's image shape
        const int yG = (int)std::ceil((float)GVS_GRID_X / ((float)feL->width / (float)feL->height));
        if (yG > gvs_matcher::kMaxGridRows) { delete m; return fail(GVS_ERR_INVALID, "image taller than 8x its width is not supported"); }
        m->maxCells = GVS_GRID_X * std::max(GVS_GRID_X, yG);
    }
    const size_t B = (size_t)std::min(feL->maxImages, feR->maxImages);
    bool ok = devAlloc(&m->rightIdxs, (size_t)feL->totalCap) == cudaSuccess && devAlloc(&m->depth, (size_t)feL->totalCap) == cudaSuccess &&
              devAlloc(&m->close, (size_t)feL->totalCap) == cudaSuccess && devAlloc(&m->sadBest, (size_t)feL->totalCap) == cudaSuccess &&
              devAlloc(&m->leftIdxs, (size_t)feR->totalCap) == cudaSuccess &&
              devAlloc(&m->cellStartL, B * (m->maxCells + 1)) == cudaSuccess && devAlloc(&m->cellStartR, B * (m->maxCells + 1)) == cudaSuccess &&
              devAlloc(&m->itemsL, (size_t)feL->totalCap) == cudaSuccess && devAlloc(&m->itemsR, (size_t)feR->totalCap) == cudaSuccess &&
              devAlloc(&m->pjStatus, 2) == cudaSuccess && cudaEventCreateWithFlags(&m->evR, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evFork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evJoin, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evHold, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evLast, cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithFlags(&m->ownStream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : m->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    for (auto& e : m->evT) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    if (!ok) {
        int rc = fail(GVS_ERR_CUDA, "matcher allocation", cudaGetErrorString(cudaGetLastError()));
        gvs_matcher_destroy(m);
        return rc;
    }
    *out = m;
    return GVS_OK;
}

int gvs_matcher_destroy(gvs_matcher* m) {
    if (!m) return GVS_OK;
    // (the extractors may be gone already when handles are torn down in arbitrary order, e.g. at interpreter exit: only the
    // matcher's own state is touched here)
    DeviceGuard deviceGuard(m->device);
    if (m->evLast) cudaEventSynchronize(m->evLast);
    if (m->ownStream) { cudaStreamSynchronize(m->ownStream); cudaStreamDestroy(m->ownStream); }
    devFree(m->pjOwn);   // the matcher's last enqueued work (no device-wide synchronisation: see gvs_extractor_create)
    devFree(m->rightIdxs); devFree(m->leftIdxs); devFree(m->sadBest); devFree(m->sortScratch); devFree(m->depth); devFree(m->close);
    devFree(m->cellStartL); devFree(m->cellStartR); devFree(m->itemsL); devFree(m->itemsR); devFree(m->pj);
    devFree(m->rowStart); devFree(m->rowItems);
    devFree(m->pool); devFree(m->pjStatus);
    if (m->evR) cudaEventDestroy(m->evR);
    if (m->evFork) cudaEventDestroy(m->evFork);
    if (m->evJoin) cudaEventDestroy(m->evJoin);
    if (m->evHold) cudaEventDestroy(m->evHold);
    if (m->evLast) cudaEventDestroy(m->evLast);
    for (auto& e : m->ev) if (e) cudaEventDestroy(e);
    for (auto& e : m->evT) if (e) cudaEventDestroy(e);
    devFree(m->tFlag); devFree(m->tPredL); devFree(m->tPredR); devFree(m->tLvlL); devFree(m->tLvlR); devFree(m->tDescRow);
    devFree(m->tWPos); devFree(m->tMaxScaleDist); devFree(m->tMpCnt); devFree(m->tMpOff); devFree(m->tMpSlot); devFree(m->tMpKey);
    devFree(m->tMatchedL); devFree(m->tMatchedR); devFree(m->tMatches); devFree(m->tNMatches); devFree(m->tRecs);
    devFree(m->tPoses); devFree(m->tPrev); devFree(m->tStatus);
    if (m->hMpOff) cudaFreeHost(m->hMpOff);
    if (m->hNMatches) cudaFreeHost(m->hNMatches);
    if (m->hTStatus) cudaFreeHost(m->hTStatus);
    delete m;
    return GVS_OK;
}

static int forkFromCaller(gvs_matcher* m) {
    if (!m->callerStream) return GVS_OK;
    GVS_CUDA_OK(cudaEventRecord(m->evFork, m->callerStream));
    if (m->L->stream != m->callerStream) GVS_CUDA_OK(cudaStreamWaitEvent(m->L->stream, m->evFork, 0));
    if (m->R != m->L && m->R->stream != m->callerStream) GVS_CUDA_OK(cudaStreamWaitEvent(m->R->stream, m->evFork, 0));
    return GVS_OK;
}
static int joinIntoCaller(gvs_matcher* m) {
    if (!m->callerStream || m->L->stream == m->callerStream) return GVS_OK;
    GVS_CUDA_OK(cudaEventRecord(m->evJoin, m->L->stream));
    GVS_CUDA_OK(cudaStreamWaitEvent(m->callerStream, m->evJoin, 0));
    return GVS_OK;
}

int gvs_matcher_set_stream(gvs_matcher* m, void* cuda_stream) {
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    m->callerStream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return GVS_OK;
}

int gvs_stereo_match_batch(gvs_matcher* m, int n_frames) {
    NvtxScope nvtx("gvs:stereo_match");
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (n_frames < 1 || n_frames > m->L->lastN || n_frames > m->R->lastN) return fail(GVS_ERR_STATE, "extract both sides first");
    DeviceGuard deviceGuard(m->L->device);
    int rc = joinRight(m);
    if (rc) return rc;
    cudaStream_t s = m->L->stream;
    StereoArgs a{};
    a.slabL = m->L->slab; a.slabR = m->R->slab; a.kpL = m->L->kps; a.kpR = m->R->kps; a.descL = m->L->desc; a.descR = m->R->desc;
    a.offL = m->L->imgOff; a.offR = m->R->imgOff; a.keptCntL = m->L->keptCnt; a.keptCntR = m->R->keptCnt;
    a.auxR = m->R->auxKey;
    if (!m->rowStart) {   // first stereo match of this matcher: the row buckets of the right keys
        // a key lies in the buckets of rows floor(y - r) .. ceil(y + r), r = 2 * scalePyramid[octave] (src/FeatureMatcher.cpp:741-744)
        const int dup = 2 * (int)std::ceil(2.0 * (double)m->R->scale[m->R->g.nLevels - 1]) + 3;
        const size_t B = (size_t)std::min(m->L->maxImages, m->R->maxImages);
        m->rowItemsStride = m->R->g.keptPerImage * dup;
        GVS_CUDA_OK(devAlloc(&m->rowStart, B * (size_t)(std::min(m->imageHeight, m->L->height) + 1)));
        GVS_CUDA_OK(devAlloc(&m->rowItems, B * (size_t)m->rowItemsStride));
    }
    a.rowStart = m->rowStart; a.rowItems = m->rowItems; a.rowItemsStride = m->rowItemsStride;
    a.bucketRows = std::min(m->imageHeight, m->L->height);   // (a left key's row is below both)
    a.rightIdxs = m->rightIdxs; a.leftIdxs = m->leftIdxs; a.depth = m->depth; a.close = m->close; a.sadBest = m->sadBest;
    a.fx = m->fx; a.baseline = m->baseline; a.imageHeight = m->imageHeight; a.closeNumber = 40; a.thDist = 75;
    {
        size_t n2 = 1;
        while (n2 < (size_t)m->L->g.keptPerImage) n2 <<= 1;
        if (n2 * 8 > 160 * 1024 && !m->sortScratch) {
            uint8_t* p = nullptr;
            GVS_CUDA_OK(devAlloc(&p, (size_t)std::min(m->L->maxImages, m->R->maxImages) * n2 * 8));
            m->sortScratch = p;
        }
        a.sortScratch = m->sortScratch;
    }
    m->launches = 0;
    if (m->L->timing) cudaEventRecord(m->ev[0], s);
    GVS_CUDA_OK(cudaMemsetAsync(m->leftIdxs, 0xff, sizeof(int) * (size_t)std::min<size_t>((size_t)n_frames * m->R->g.keptPerImage, m->R->totalCap), s));
    launchStereo(m->L->g, a, n_frames, m->L->g.keptPerImage, s, &m->launches);
    if (m->L->timing) cudaEventRecord(m->ev[1], s);
    GVS_CUDA_OK(cudaGetLastError());
    m->launchesTotal += m->launches;
    m->stereoValid = true;
    m->lastFrames = n_frames;
    m->stereoGenL = m->L->generation; m->stereoGenR = m->R->generation;
    return holdRight(m);
}

int gvs_matcher_sync(gvs_matcher* m) {
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    int rc = syncExtract(m->L);
    if (rc) return rc;
    if (m->R != m->L) rc = syncExtract(m->R);
    if (rc) return rc;
    if (m->L->timing && m->stereoValid) {
        cudaEventElapsedTime(&m->stereoMs, m->ev[0], m->ev[1]);
        if (m->gridsValid) cudaEventElapsedTime(&m->gridMs, m->ev[2], m->ev[3]);
        if (m->temporalFrames > 0 && m->temporalGen == m->L->generation) cudaEventElapsedTime(&m->temporalMs, m->evT[0], m->evT[1]);
        m->L->times.stereo_ms = m->stereoMs;
        m->L->times.grid_ms = m->gridMs;
        m->L->times.temporal_ms = m->temporalMs;
    }
    return GVS_OK;
}

int gvs_stereo_download(gvs_matcher* m, int first, int count, int32_t* right_idxs, int32_t* left_idxs, float* depth, uint8_t* close) {
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (!m->stereoValid || m->stereoGenL != m->L->generation || m->stereoGenR != m->R->generation)
        return fail(GVS_ERR_STATE, "no stereo batch has been matched for the current frames");
    if (first < 0 || count < 0 || first + count > m->lastFrames) return fail(GVS_ERR_INVALID, "frame range out of the last batch");
    int rc = gvs_matcher_sync(m);
    if (rc) return rc;
    cudaStream_t s = m->L->stream;
    const int l0 = m->L->hOff[first], nl = m->L->hOff[first + count] - l0;
    const int r0 = m->R->hOff[first], nr = m->R->hOff[first + count] - r0;
    if (nl > 0) {
        if (right_idxs) GVS_CUDA_OK(cudaMemcpyAsync(right_idxs, m->rightIdxs + l0, sizeof(int) * nl, cudaMemcpyDeviceToHost, s));
        if (depth) GVS_CUDA_OK(cudaMemcpyAsync(depth, m->depth + l0, sizeof(float) * nl, cudaMemcpyDeviceToHost, s));
        if (close) GVS_CUDA_OK(cudaMemcpyAsync(close, m->close + l0, (size_t)nl, cudaMemcpyDeviceToHost, s));
    }
    if (nr > 0 && left_idxs) GVS_CUDA_OK(cudaMemcpyAsync(left_idxs, m->leftIdxs + r0, sizeof(int) * nr, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    return GVS_OK;
}

int gvs_assign_grids_batch(gvs_matcher* m, int n_frames, int width, int height) {
    NvtxScope nvtx("gvs:assign_grids");
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (n_frames < 1 || n_frames > m->L->lastN) return fail(GVS_ERR_STATE, "extract first");
    if (width < 1 || height < 1) return fail(GVS_ERR_INVALID, "bad size");
    DeviceGuard deviceGuard(m->L->device);
    // src/FeatureTracker.cpp:30-34 (float32)
    const float imageRatio = (float)width / (float)height;
    m->xGrids = GVS_GRID_X;
    m->yGrids = (int)std::ceil((float)m->xGrids / imageRatio);
    if (m->xGrids * m->yGrids > m->maxCells) return fail(GVS_ERR_INVALID, "key grid larger than the one this matcher was created for (image shape differs from the extractors')");
    m->xMult = (float)m->xGrids / (float)width;
    m->yMult = (float)m->yGrids / (float)height;
    int rc = joinRight(m);
    if (rc) return rc;
    cudaStream_t s = m->L->stream;
    int launches = 0;
    if (m->L->timing) cudaEventRecord(m->ev[2], s);
    GridArgs a{};
    a.xGrids = m->xGrids; a.yGrids = m->yGrids; a.xMult = m->xMult; a.yMult = m->yMult; a.maxCells = m->maxCells;
    a.kps = m->L->kps; a.off = m->L->imgOff; a.cellStart = m->cellStartL; a.items = m->itemsL;
    if (m->R != m->L && n_frames <= m->R->lastN) {
        GridArgs r = a;
        r.kps = m->R->kps; r.off = m->R->imgOff; r.cellStart = m->cellStartR; r.items = m->itemsR;
        launchGrid2(a, &r, n_frames, s, &launches);
    } else {
        launchGrid(a, n_frames, s, &launches);
    }
    if (m->L->timing) cudaEventRecord(m->ev[3], s);
    GVS_CUDA_OK(cudaGetLastError());
    m->launchesTotal += launches;
    m->launches += launches;
    m->gridsValid = true;
    m->gridFrames = n_frames;
    m->gridGenL = m->L->generation; m->gridGenR = m->R->generation;
    if (m->lastFrames < n_frames) m->lastFrames = n_frames;
    return holdRight(m);
}

int gvs_grid_download(gvs_matcher* m, int frame, int side, int32_t* cell_start, int32_t* items, int* y_grids, float* x_mult, float* y_mult) {
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (!m->gridsValid || m->gridGenL != m->L->generation || (m->R != m->L && m->gridGenR != m->R->generation))
        return fail(GVS_ERR_STATE, "grids have not been assigned for the current frames");
    gvs_extractor* fe = side ? m->R : m->L;
    if (frame < 0 || frame >= fe->lastN) return fail(GVS_ERR_INVALID, "bad frame");
    int rc = gvs_matcher_sync(m);
    if (rc) return rc;
    const int nc = m->xGrids * m->yGrids;
    const int o0 = fe->hOff[frame], n = fe->hOff[frame + 1] - o0;
    if (cell_start) GVS_CUDA_OK(cudaMemcpy(cell_start, (side ? m->cellStartR : m->cellStartL) + (size_t)frame * (m->maxCells + 1), sizeof(int) * (nc + 1), cudaMemcpyDeviceToHost));
    if (items && n > 0) GVS_CUDA_OK(cudaMemcpy(items, (side ? m->itemsR : m->itemsL) + o0, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (y_grids) *y_grids = m->yGrids;
    if (x_mult) *x_mult = m->xMult;
    if (y_mult) *y_mult = m->yMult;
    return GVS_OK;
}

int gvs_track_stereo_batch_resident(gvs_matcher* m, int n_frames, int width, int height) {
    NvtxScope nvtx("gvs:track_stereo_batch");
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (m->R == m->L) return fail(GVS_ERR_STATE, "stereo tracking needs two extractors");
    DeviceGuard deviceGuard(m->L->device);
    int rc = forkFromCaller(m);
    if (rc) return rc;
    rc = gvs_extract_batch_resident(m->L, n_frames);
    if (rc) return rc;
    rc = gvs_extract_batch_resident(m->R, n_frames);
    if (rc) return rc;
    rc = gvs_stereo_match_batch(m, n_frames);
    if (rc) return rc;
    rc = gvs_assign_grids_batch(m, n_frames, width, height);
    if (rc) return rc;
    return joinIntoCaller(m);
}

int gvs_track_stereo_batch_host(gvs_matcher* m, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride,
                                int n_frames, int width, int height) {
    NvtxScope nvtx("gvs:track_stereo_batch_host");
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (m->R == m->L) return fail(GVS_ERR_STATE, "stereo tracking needs two extractors");
    DeviceGuard deviceGuard(m->L->device);
    int rc = forkFromCaller(m);
    if (rc) return rc;
    rc = gvs_extract_batch_host(m->L, left, image_stride, row_stride, n_frames);
    if (rc) return rc;
    rc = gvs_extract_batch_host(m->R, right, image_stride, row_stride, n_frames);
    if (rc) return rc;
    rc = gvs_stereo_match_batch(m, n_frames);
    if (rc) return rc;
    rc = gvs_assign_grids_batch(m, n_frames, width, height);
    if (rc) return rc;
    return joinIntoCaller(m);
}

int gvs_matcher_launch_count(gvs_matcher* m, int64_t* launches) {
    if (!m || !launches) return fail(GVS_ERR_INVALID, "null argument");
    *launches = m->launchesTotal;
    return GVS_OK;
}

int gvs_match_projection_lba(gvs_matcher* m, const gvs_lba_frame* kf, const gvs_lba_points* pts, float rad, int32_t* out_pairs,
                             int* n_matches) {
    if (!m || !kf || !pts || !out_pairs || !n_matches) return fail(GVS_ERR_INVALID, "null argument");
    const int M = pts->count, NL = kf->n_left, NR = kf->n_right;
    *n_matches = 0;
    if (M <= 0) return GVS_OK;
    if (NL < 0 || NR < 0 || NL > 65535 || NR > 65535 || kf->width < 1 || kf->height < 1) return fail(GVS_ERR_INVALID, "bad key frame");
    if (!pts->desc || !pts->pred_l || !pts->pred_r || !pts->pred_scale || !pts->last_l || !pts->last_r)
        return fail(GVS_ERR_INVALID, "candidate point arrays missing");
    if ((NL && (!kf->kp_left || !kf->desc_left || !kf->right_idxs || !kf->unmatched_f)) ||
        (NR && (!kf->kp_right || !kf->desc_right || !kf->left_idxs || !kf->unmatched_fr)))
        return fail(GVS_ERR_INVALID, "key frame arrays missing");
    for (int i = 0; i < M; i++)
        if (pts->pred_scale[i] < 0 || pts->pred_scale[i] >= m->L->g.nLevels) return fail(GVS_ERR_INVALID, "pred_scale out of range");
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->ownStream;   // (does not read the resident frame: the matcher's own stream and scratch)
    const float imageRatio = (float)kf->width / (float)kf->height;
    const int xGrids = GVS_GRID_X, yGrids = (int)std::ceil((float)xGrids / imageRatio);
    if (xGrids * yGrids > m->maxCells) return fail(GVS_ERR_INVALID, "key grid larger than the one this matcher was created for (image shape differs from the extractors')");
    const float xMult = (float)xGrids / (float)kf->width, yMult = (float)yGrids / (float)kf->height;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o += al(std::max<size_t>(bytes, 4)); return r; };
    const size_t oDesc = take((size_t)M * 32), oPL = take((size_t)M * 8), oPR = take((size_t)M * 8), oLv = take((size_t)M * 4),
                 oLL = take((size_t)M * 8), oLR = take((size_t)M * 8), oIF = take(M), oIR = take(M), oOut = take((size_t)M * 8),
                 oRec = take((size_t)M * 2 * 20), oCnt = take(16), oKL = take((size_t)NL * 28), oKR = take((size_t)NR * 28),
                 oDL = take((size_t)NL * 32), oDR = take((size_t)NR * 32), oLI = take((size_t)NR * 4), oRI = take((size_t)NL * 4),
                 oUF = take((size_t)NL * 4), oUR = take((size_t)NR * 4), oCSL = take((size_t)(m->maxCells + 1) * 4),
                 oCSR = take((size_t)(m->maxCells + 1) * 4), oItL = take((size_t)NL * 4), oItR = take((size_t)NR * 4), oOff = take(32);
    if (o > m->pjOwnBytes) {
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        devFree(m->pjOwn); m->pjOwn = nullptr; m->pjOwnBytes = 0;
        GVS_CUDA_OK(devAlloc(&m->pjOwn, o));
        m->pjOwnBytes = o;
    }
    const long long need = std::min<long long>((long long)M * std::max(std::max(NL, NR), 1), (long long)1 << 28);
    if (need > m->poolCap) {
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        devFree(m->pool); m->pool = nullptr; m->poolCap = 0;
        GVS_CUDA_OK(devAlloc(&m->pool, (size_t)need * 2));
        m->poolCap = (int)need;
    }
    uint8_t* b = m->pjOwn;
    std::vector<uint8_t> inL(M), inR(M);
    for (int i = 0; i < M; i++) {
        inL[i] = pts->pred_l[2 * i] > 0 && pts->pred_l[2 * i + 1] > 0;   // :104
        inR[i] = pts->pred_r[2 * i] > 0 && pts->pred_r[2 * i + 1] > 0;   // :153
    }
    const int offs[4] = {0, NL, 0, NR};
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    GVS_CUDA_OK(up(oDesc, pts->desc, (size_t)M * 32)); GVS_CUDA_OK(up(oPL, pts->pred_l, (size_t)M * 8));
    GVS_CUDA_OK(up(oPR, pts->pred_r, (size_t)M * 8)); GVS_CUDA_OK(up(oLv, pts->pred_scale, (size_t)M * 4));
    GVS_CUDA_OK(up(oLL, pts->last_l, (size_t)M * 8)); GVS_CUDA_OK(up(oLR, pts->last_r, (size_t)M * 8));
    GVS_CUDA_OK(up(oIF, inL.data(), M)); GVS_CUDA_OK(up(oIR, inR.data(), M));
    GVS_CUDA_OK(up(oKL, kf->kp_left, (size_t)NL * 28)); GVS_CUDA_OK(up(oKR, kf->kp_right, (size_t)NR * 28));
    GVS_CUDA_OK(up(oDL, kf->desc_left, (size_t)NL * 32)); GVS_CUDA_OK(up(oDR, kf->desc_right, (size_t)NR * 32));
    GVS_CUDA_OK(up(oLI, kf->left_idxs, (size_t)NR * 4)); GVS_CUDA_OK(up(oRI, kf->right_idxs, (size_t)NL * 4));
    GVS_CUDA_OK(up(oUF, kf->unmatched_f, (size_t)NL * 4)); GVS_CUDA_OK(up(oUR, kf->unmatched_fr, (size_t)NR * 4));
    GVS_CUDA_OK(up(oOff, offs, sizeof(offs)));
    GVS_CUDA_OK(cudaMemsetAsync(m->pjStatus, 0, 2 * sizeof(int), s));
    int launches = 0;
    GridArgs ga{};
    ga.xGrids = xGrids; ga.yGrids = yGrids; ga.xMult = xMult; ga.yMult = yMult; ga.maxCells = m->maxCells;
    ga.kps = reinterpret_cast<gvs_keypoint*>(b + oKL); ga.off = reinterpret_cast<int*>(b + oOff); ga.cellStart = reinterpret_cast<int*>(b + oCSL);
    ga.items = reinterpret_cast<int*>(b + oItL);
    launchGrid(ga, 1, s, &launches);
    ga.kps = reinterpret_cast<gvs_keypoint*>(b + oKR); ga.off = reinterpret_cast<int*>(b + oOff) + 2; ga.cellStart = reinterpret_cast<int*>(b + oCSR);
    ga.items = reinterpret_cast<int*>(b + oItR);
    launchGrid(ga, 1, s, &launches);
    ProjArgs a{};
    a.mode = 3; a.M = M;
    a.mpDesc = b + oDesc; a.predL = reinterpret_cast<float*>(b + oPL); a.predR = reinterpret_cast<float*>(b + oPR);
    a.lvlL = reinterpret_cast<int*>(b + oLv); a.lvlR = reinterpret_cast<int*>(b + oLv);
    a.inFrame = b + oIF; a.inFrameR = b + oIR;
    a.kpL = reinterpret_cast<gvs_keypoint*>(b + oKL); a.descL = b + oDL; a.NL = NL;
    a.kpR = reinterpret_cast<gvs_keypoint*>(b + oKR); a.descR = b + oDR; a.NR = NR;
    a.cellStartL = reinterpret_cast<int*>(b + oCSL); a.itemsL = reinterpret_cast<int*>(b + oItL);
    a.cellStartR = reinterpret_cast<int*>(b + oCSR); a.itemsR = reinterpret_cast<int*>(b + oItR);
    a.leftIdxs = reinterpret_cast<int*>(b + oLI); a.rightIdxs = reinterpret_cast<int*>(b + oRI);
    a.xGrids = xGrids; a.yGrids = yGrids; a.xMult = xMult; a.yMult = yMult; a.rad = rad;
    for (int i = 0; i < kMaxLevels; i++) a.scalePyramid[i] = i < m->L->g.nLevels ? m->L->scale[i] : 1.f;
    a.recs = b + oRec; a.poolTop = reinterpret_cast<int*>(b + oCnt);
    a.candL = m->pool; a.candR = m->pool + m->poolCap; a.candCap = m->poolCap;
    a.matchedL = reinterpret_cast<int*>(b + oUF); a.matchedR = reinterpret_cast<int*>(b + oUR);  // static exclusion arrays
    a.matches = nullptr; a.nMatches = m->pjStatus + 1; a.status = m->pjStatus;
    launchProjectLBA(a, reinterpret_cast<float*>(b + oLL), reinterpret_cast<float*>(b + oLR), reinterpret_cast<int*>(b + oOut), s, &launches);
    m->launchesTotal += launches;
    int st[2] = {0, 0};
    GVS_CUDA_OK(cudaMemcpyAsync(out_pairs, b + oOut, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(st, m->pjStatus, sizeof(st), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    GVS_CUDA_OK(cudaGetLastError());
    if (st[0] & kStatProjOverflow) return fail(GVS_ERR_CAPACITY, "projection candidate pool overflow");
    *n_matches = st[1];
    return GVS_OK;
}

int gvs_project_mappoints(gvs_matcher* m, const double* pose_inv, double fx, double fy, double cx, double cy, int width, int height,
                          int count, const double* world_pos, const float* max_scale_dist, const float* log_scale,
                          const int32_t* n_scale_lev, float* pred, int32_t* scale_level, uint8_t* in_frame) {
    if (!m || !pose_inv || count < 0) return fail(GVS_ERR_INVALID, "bad argument");
    if (count == 0) return GVS_OK;
    if (!world_pos || !max_scale_dist || !log_scale || !n_scale_lev || !pred || !scale_level || !in_frame) return fail(GVS_ERR_INVALID, "null array");
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->ownStream;   // (does not read the resident frame: the matcher's own stream and scratch)
    const size_t M = (size_t)count;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t oPose = 0, oW = al(128), oMs = oW + al(M * 32), oLs = oMs + al(M * 4), oNl = oLs + al(M * 4), oPr = oNl + al(M * 4),
                 oSl = oPr + al(M * 8), oIf = oSl + al(M * 4), total = oIf + al(M);
    if (total > m->pjOwnBytes) {
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        devFree(m->pjOwn); m->pjOwn = nullptr; m->pjOwnBytes = 0;
        GVS_CUDA_OK(devAlloc(&m->pjOwn, total));
        m->pjOwnBytes = total;
    }
    uint8_t* b = m->pjOwn;
    GVS_CUDA_OK(cudaMemcpyAsync(b + oPose, pose_inv, 128, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(b + oW, world_pos, M * 32, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(b + oMs, max_scale_dist, M * 4, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(b + oLs, log_scale, M * 4, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(b + oNl, n_scale_lev, M * 4, cudaMemcpyHostToDevice, s));
    GVS_CUDA_OK(cudaMemcpyAsync(b + oPr, pred, M * 8, cudaMemcpyHostToDevice, s));        // in-out: untouched when not in frame
    GVS_CUDA_OK(cudaMemcpyAsync(b + oSl, scale_level, M * 4, cudaMemcpyHostToDevice, s));
    int launches = 0;
    launchWorldToFrame(count, reinterpret_cast<double*>(b + oPose), fx, fy, cx, cy, width, height, reinterpret_cast<double*>(b + oW),
                       reinterpret_cast<float*>(b + oMs), reinterpret_cast<float*>(b + oLs), reinterpret_cast<int*>(b + oNl),
                       reinterpret_cast<float*>(b + oPr), reinterpret_cast<int*>(b + oSl), b + oIf, s, &launches);
    m->launchesTotal += launches;
    GVS_CUDA_OK(cudaMemcpyAsync(pred, b + oPr, M * 8, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(scale_level, b + oSl, M * 4, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(in_frame, b + oIf, M, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    GVS_CUDA_OK(cudaGetLastError());
    return GVS_OK;
}

int gvs_predict_mappoints(gvs_matcher* m, int mode, const double* to_camera, const double* to_camera_r, double fx, double fy,
                          double cx, double cy, int width, int height, int count, const uint8_t* valid, const double* world_pos,
                          const float* max_scale_dist, const float* log_scale, const int32_t* n_scale_lev, float* pred_l,
                          int32_t* scale_level_l, uint8_t* in_frame, float* pred_r, int32_t* scale_level_r, uint8_t* in_frame_r,
                          uint8_t* keep, int32_t* matches, int32_t* matched_l, int n_left, int32_t* matched_r, int n_right,
                          uint8_t* outliers) {
    if (!m || !to_camera || count < 0 || mode < 0 || mode > 2) return fail(GVS_ERR_INVALID, "bad argument");
    if (count == 0) return GVS_OK;
    const bool two = mode != 1;
    if (!world_pos || !max_scale_dist || !log_scale || !n_scale_lev || !pred_l || !scale_level_l || !in_frame) return fail(GVS_ERR_INVALID, "null array");
    if (two && (!to_camera_r || !pred_r || !scale_level_r || !in_frame_r)) return fail(GVS_ERR_INVALID, "right-camera arrays missing");
    if (mode != 2 && !keep) return fail(GVS_ERR_INVALID, "keep is required");
    if (mode == 2) {
        if (!matches || !matched_l || !matched_r || !outliers || n_left < 0 || n_right < 0) return fail(GVS_ERR_INVALID, "match arrays missing");
        for (int i = 0; i < count; i++)
            if (matches[2 * i] >= n_left || matches[2 * i + 1] >= n_right) return fail(GVS_ERR_INVALID, "match index outside matched_l / matched_r");
    }
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->ownStream;   // (does not read the resident frame: the matcher's own stream and scratch)
    const size_t M = (size_t)count, NL = (size_t)std::max(n_left, 1), NR = (size_t)std::max(n_right, 1);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += al(bytes); return at; };
    const size_t oPose = take(256), oVa = take(M), oW = take(M * 32), oMs = take(M * 4), oLs = take(M * 4), oNl = take(M * 4),
                 oPl = take(M * 8), oSl = take(M * 4), oIf = take(M), oPr = take(M * 8), oSr = take(M * 4), oIr = take(M), oKeep = take(M),
                 oMt = take(M * 8), oML = take(NL * 4), oMR = take(NR * 4), oOut = take(M);
    if (o > m->pjOwnBytes) {
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        devFree(m->pjOwn); m->pjOwn = nullptr; m->pjOwnBytes = 0;
        GVS_CUDA_OK(devAlloc(&m->pjOwn, o));
        m->pjOwnBytes = o;
    }
    uint8_t* b = m->pjOwn;
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        return src ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    GVS_CUDA_OK(up(oPose, to_camera, 128));
    GVS_CUDA_OK(up(oPose + 128, two ? to_camera_r : nullptr, 128));
    GVS_CUDA_OK(up(oVa, valid, M));
    GVS_CUDA_OK(up(oW, world_pos, M * 32));
    GVS_CUDA_OK(up(oMs, max_scale_dist, M * 4));
    GVS_CUDA_OK(up(oLs, log_scale, M * 4));
    GVS_CUDA_OK(up(oNl, n_scale_lev, M * 4));
    GVS_CUDA_OK(up(oPl, pred_l, M * 8));
    GVS_CUDA_OK(up(oSl, scale_level_l, M * 4));
    GVS_CUDA_OK(up(oIf, in_frame, M));
    if (two) {
        GVS_CUDA_OK(up(oPr, pred_r, M * 8));
        GVS_CUDA_OK(up(oSr, scale_level_r, M * 4));
        GVS_CUDA_OK(up(oIr, in_frame_r, M));
    }
    if (mode == 2) {
        GVS_CUDA_OK(up(oMt, matches, M * 8));
        GVS_CUDA_OK(up(oML, matched_l, (size_t)n_left * 4));
        GVS_CUDA_OK(up(oMR, matched_r, (size_t)n_right * 4));
        GVS_CUDA_OK(up(oOut, outliers, M));
    }
    PredictArgs a{};
    a.mode = mode; a.M = count;
    a.toCamera = reinterpret_cast<double*>(b + oPose); a.toCameraR = reinterpret_cast<double*>(b + oPose + 128);
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.width = width; a.height = height;
    a.valid = valid ? b + oVa : nullptr; a.wPos = reinterpret_cast<double*>(b + oW);
    a.maxScaleDist = reinterpret_cast<float*>(b + oMs); a.logScale = reinterpret_cast<float*>(b + oLs); a.nScaleLev = reinterpret_cast<int*>(b + oNl);
    a.predL = reinterpret_cast<float*>(b + oPl); a.scaleLevelL = reinterpret_cast<int*>(b + oSl); a.inFrame = b + oIf;
    a.predR = reinterpret_cast<float*>(b + oPr); a.scaleLevelR = reinterpret_cast<int*>(b + oSr); a.inFrameR = b + oIr;
    a.keep = b + oKeep;
    a.matches = reinterpret_cast<int*>(b + oMt); a.matchedL = reinterpret_cast<int*>(b + oML); a.matchedR = reinterpret_cast<int*>(b + oMR);
    a.outliers = b + oOut;
    int launches = 0;
    launchPredictMapPoints(a, s, &launches);
    m->launchesTotal += launches;
    auto down = [&](void* dst, size_t off, size_t bytes) -> cudaError_t {
        return dst && bytes ? cudaMemcpyAsync(dst, b + off, bytes, cudaMemcpyDeviceToHost, s) : cudaSuccess;
    };
    GVS_CUDA_OK(down(pred_l, oPl, M * 8));
    GVS_CUDA_OK(down(scale_level_l, oSl, M * 4));
    GVS_CUDA_OK(down(in_frame, oIf, M));
    if (two) {
        GVS_CUDA_OK(down(pred_r, oPr, M * 8));
        GVS_CUDA_OK(down(scale_level_r, oSr, M * 4));
        GVS_CUDA_OK(down(in_frame_r, oIr, M));
    }
    if (mode != 2) GVS_CUDA_OK(down(keep, oKeep, M));
    else {
        GVS_CUDA_OK(down(matches, oMt, M * 8));
        GVS_CUDA_OK(down(matched_l, oML, (size_t)n_left * 4));
        GVS_CUDA_OK(down(matched_r, oMR, (size_t)n_right * 4));
        GVS_CUDA_OK(down(outliers, oOut, M));
    }
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    GVS_CUDA_OK(cudaGetLastError());
    return GVS_OK;
}

int gvs_find_outliers(gvs_matcher* m, int frame, int mode, const double* pose, const double* pose_r, double fx, double fy,
                      double cx, double cy, double thres, float close_dist, int count, const uint8_t* valid,
                      const double* world_pos, const uint8_t* in_frame, const uint8_t* in_frame_r, int32_t* matches,
                      uint8_t* outliers, float* est_depth, uint8_t* close, int32_t* right_idxs, int32_t* left_idxs,
                      int* n_inliers, int* n_stereo) {
    if (!m || !pose || !n_inliers || !n_stereo || count < 0) return fail(GVS_ERR_INVALID, "bad argument");
    if (mode != 0 && mode != 1) return fail(GVS_ERR_INVALID, "bad mode");
    if (mode == 0 && (!pose_r || !m->stereoValid || m->stereoGenL != m->L->generation || m->stereoGenR != m->R->generation))
        return fail(GVS_ERR_STATE, "findOutliersR needs toCameraR and a stereo batch of the current frames");
    if (frame < 0 || frame >= m->L->lastN || (mode == 0 && frame >= m->lastFrames)) return fail(GVS_ERR_INVALID, "bad frame");
    *n_inliers = 0; *n_stereo = 0;
    int rc = gvs_matcher_sync(m);
    if (rc) return rc;
    const int oL = m->L->hOff[frame], NL = m->L->hOff[frame + 1] - oL;
    const bool two = mode == 0;
    const int oR = two ? m->R->hOff[frame] : 0, NR = two ? m->R->hOff[frame + 1] - oR : 0;
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->L->stream;
    if (count > 0) {
        if (!world_pos || !matches || !outliers) return fail(GVS_ERR_INVALID, "null array");
        for (int i = 0; i < count; i++)
            if (matches[2 * i] >= NL || (two && matches[2 * i + 1] >= NR))
                return fail(GVS_ERR_INVALID, "match index outside the frame's keys");
        const size_t M = (size_t)count;
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t oPose = 0, oSig = al(256), oW = oSig + al(kMaxLevels * 4), oVa = oW + al(M * 32), oIf = oVa + al(M), oIr = oIf + al(M),
                     oMt = oIr + al(M), oOut = oMt + al(M * 8), oAct = oOut + al(M), oFc = oAct + al(M * 4),
                     oCnt = oFc + al((size_t)std::max(NL, 1) * 4), total = oCnt + al(16);
        if (total > m->pjBytes) {
            devFree(m->pj); m->pj = nullptr; m->pjBytes = 0;
            GVS_CUDA_OK(devAlloc(&m->pj, total));
            m->pjBytes = total;
        }
        uint8_t* b = m->pj;
        auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
            return src ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
        };
        GVS_CUDA_OK(up(oPose, pose, 128));
        GVS_CUDA_OK(up(oPose + 128, pose_r, 128));
        GVS_CUDA_OK(up(oSig, &m->L->invSigma[0], (size_t)m->L->g.nLevels * 4));
        GVS_CUDA_OK(up(oW, world_pos, M * 32));
        GVS_CUDA_OK(up(oVa, valid, M));
        GVS_CUDA_OK(up(oIf, in_frame, M));
        GVS_CUDA_OK(up(oIr, in_frame_r, M));
        GVS_CUDA_OK(up(oMt, matches, M * 8));
        GVS_CUDA_OK(up(oOut, outliers, M));
        GVS_CUDA_OK(cudaMemsetAsync(b + oFc, 0x7f, (size_t)std::max(NL, 1) * 4, s));
        GVS_CUDA_OK(cudaMemsetAsync(b + oCnt, 0, 16, s));
        OutlierArgs a{};
        a.mode = mode; a.M = count; a.NL = NL; a.NR = NR;
        a.pose = reinterpret_cast<double*>(b + oPose); a.poseR = reinterpret_cast<double*>(b + oPose + 128);
        a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.thres = thres; a.closeDist = close_dist;
        a.invSigma = reinterpret_cast<float*>(b + oSig);
        a.wPos = reinterpret_cast<double*>(b + oW);
        a.valid = valid ? b + oVa : nullptr; a.inFrame = in_frame ? b + oIf : nullptr; a.inFrameR = in_frame_r ? b + oIr : nullptr;
        a.matches = reinterpret_cast<int*>(b + oMt); a.outliers = b + oOut;
        a.kpL = m->L->kps + oL; a.kpR = two ? m->R->kps + oR : nullptr;
        a.depth = two ? m->depth + oL : nullptr; a.close = two ? m->close + oL : nullptr;
        a.rightIdxs = two ? m->rightIdxs + oL : nullptr; a.leftIdxs = two ? m->leftIdxs + oR : nullptr;
        a.act = reinterpret_cast<int*>(b + oAct); a.firstClear = reinterpret_cast<int*>(b + oFc);
        a.counters = reinterpret_cast<int*>(b + oCnt);
        int launches = 0;
        launchFindOutliers(a, s, &launches);
        m->launchesTotal += launches;
        int cnt[2] = {0, 0};
        GVS_CUDA_OK(cudaMemcpyAsync(matches, b + oMt, M * 8, cudaMemcpyDeviceToHost, s));
        GVS_CUDA_OK(cudaMemcpyAsync(outliers, b + oOut, M, cudaMemcpyDeviceToHost, s));
        GVS_CUDA_OK(cudaMemcpyAsync(cnt, b + oCnt, 8, cudaMemcpyDeviceToHost, s));
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        *n_inliers = cnt[0]; *n_stereo = cnt[1];
    }
    if (two) {
        if (est_depth) GVS_CUDA_OK(cudaMemcpyAsync(est_depth, m->depth + oL, (size_t)NL * 4, cudaMemcpyDeviceToHost, s));
        if (close) GVS_CUDA_OK(cudaMemcpyAsync(close, m->close + oL, (size_t)NL, cudaMemcpyDeviceToHost, s));
        if (right_idxs) GVS_CUDA_OK(cudaMemcpyAsync(right_idxs, m->rightIdxs + oL, (size_t)NL * 4, cudaMemcpyDeviceToHost, s));
        if (left_idxs) GVS_CUDA_OK(cudaMemcpyAsync(left_idxs, m->leftIdxs + oR, (size_t)NR * 4, cudaMemcpyDeviceToHost, s));
        GVS_CUDA_OK(cudaStreamSynchronize(s));
    }
    GVS_CUDA_OK(cudaGetLastError());
    return GVS_OK;
}

int gvs_median_descriptor(gvs_matcher* m, const uint8_t* descs, const int32_t* offsets, int n_points, int32_t* best_idx) {
    if (!m || !offsets || !best_idx || n_points < 0) return fail(GVS_ERR_INVALID, "bad argument");
    if (n_points == 0) return GVS_OK;
    const int total = offsets[n_points];
    if (total < 0 || (total > 0 && !descs)) return fail(GVS_ERR_INVALID, "bad descriptor list");
    for (int p = 0; p < n_points; p++)
        if (offsets[p + 1] < offsets[p] || offsets[p + 1] - offsets[p] > 65535) return fail(GVS_ERR_INVALID, "bad offsets");
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->ownStream;   // (does not read the resident frame: the matcher's own stream and scratch)
    uint8_t* dd = nullptr; int* doff = nullptr; int* dbest = nullptr;
    cudaError_t e = devAlloc(&dd, (size_t)std::max(total, 1) * 32);
    if (e == cudaSuccess) e = devAlloc(&doff, (size_t)n_points + 1);
    if (e == cudaSuccess) e = devAlloc(&dbest, (size_t)n_points);
    if (e == cudaSuccess && total) e = cudaMemcpyAsync(dd, descs, (size_t)total * 32, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(doff, offsets, sizeof(int) * (n_points + 1), cudaMemcpyHostToDevice, s);
    int launches = 0;
    if (e == cudaSuccess) { launchMedianDescriptor(dd, doff, n_points, dbest, s, &launches); e = cudaMemcpyAsync(best_idx, dbest, sizeof(int) * n_points, cudaMemcpyDeviceToHost, s); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    devFree(dd); devFree(doff); devFree(dbest);
    m->launchesTotal += launches;
    if (e != cudaSuccess) return fail(GVS_ERR_CUDA, "gvs_median_descriptor", cudaGetErrorString(e));
    return GVS_OK;
}

int gvs_descriptor_distance(gvs_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
    if (!m || !a || !b || !dist || n < 0) return fail(GVS_ERR_INVALID, "bad argument");
    if (n == 0) return GVS_OK;
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->ownStream;   // (does not read the resident frame: the matcher's own stream and scratch)
    uint8_t *da = nullptr, *db = nullptr;
    int* dd = nullptr;
    cudaError_t e = devAlloc(&da, (size_t)n * 32);
    if (e == cudaSuccess) e = devAlloc(&db, (size_t)n * 32);
    if (e == cudaSuccess) e = devAlloc(&dd, (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, s);
    int launches = 0;
    if (e == cudaSuccess) { launchHamming(da, db, n, dd, s, &launches); e = cudaMemcpyAsync(dist, dd, sizeof(int) * n, cudaMemcpyDeviceToHost, s); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    devFree(da); devFree(db); devFree(dd);
    m->launchesTotal += launches;
    if (e != cudaSuccess) return fail(GVS_ERR_CUDA, "gvs_descriptor_distance", cudaGetErrorString(e));
    return GVS_OK;
}

}  // extern "C"
