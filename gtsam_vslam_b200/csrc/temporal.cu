// Batched temporal matching: FeatureMatcher::matchByProjectionRPred / matchByProjectionMono / matchByRadius (reference
// src/FeatureMatcher.cpp:254-526, with getMatchIdxs :13-64) for ALL frames of a batch in two launches, everything device
// resident, no host loop and no synchronisation; plus the device-side producer of their inputs for consecutive stereo frames:
// map points created from the previous frame's stereo keys (FeatureTracker::initializeMap, src/FeatureTracker.cpp:89-112, with
// MapPoint::update, src/Map.cpp:61-104) and projected into the current frame (FeatureTracker::removeOutOfFrameMPs :910-940 ->
// worldToFrame :685-741 -> MapPoint::predictScale src/Map.cpp:13-23).
//
// The reference matcher is greedy over map points in index order.  Two phases keep that order:
//   phase 1 (projb_candidates_kernel, one THREAD per (frame, map point, side), the frame's key grid and keys staged in shared
//     memory per CTA): walks the key-grid window in the reference's order (grid row, then one contiguous CSR range per row) and
//     keeps the four smallest keys (dist << 20 | position in the candidate list).  The best / second-best rule
//     (src/FeatureMatcher.cpp:281-302, strict <) only ever depends on the TWO smallest keys of the list: the best is the first
//     occurrence of the minimum, and the second best is the smallest key among the others (a prefix element with the second's
//     distance would have the smaller key, so "prefix minimum unless the suffix is strictly smaller" of SURVEY App. A.9 is the
//     same rule).  Four are kept so that phase 2 can drop candidates that get matched in the meantime.  Nothing is stored per
//     candidate: no candidate pool, no capacity to overflow.  Keys that are matched before the call (in-out matchedIdxs) are
//     skipped here already: entries only change from -1 to >= 0 during a call.
//   phase 2 (projb_resolve_kernel, one warp per frame, lane = map point): walks the frame's map points in order, 32 at a time.  A map
//     point decides from the first two of its stored candidates that are still free; if fewer than two are left of a list that
//     was longer than four, its lane walks the window again with the skip test.  The longest prefix of the chunk in which
//     no map point takes a deciding candidate of a later one is committed; the rest is re-evaluated.  matchedIdxsL/R of the
//     frame live in shared memory during the walk.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "handles.cuh"
#include "mappoint_math.cuh"

namespace gvs {

// candidate entry: idx (16 bits) | dist (9 bits) << 16 | level (4 bits) << 25 ; 0xffffffff = none
__device__ __forceinline__ unsigned tPack(int idx, int dist, int lev) { return (unsigned)idx | ((unsigned)dist << 16) | ((unsigned)lev << 25); }
__device__ __forceinline__ int tIdx(unsigned e) { return (int)(e & 0xffff); }
__device__ __forceinline__ int tDist(unsigned e) { return (int)((e >> 16) & 0x1ff); }
__device__ __forceinline__ int tLev(unsigned e) { return (int)((e >> 25) & 0xf); }
constexpr unsigned kNoCand = 0xffffffffu;

constexpr int kTopK = 4;
struct TRec { unsigned e[kTopK]; int n; };   // per (map point, side): the kTopK smallest keys' entries in key order, n = list length

struct ProjBatchArgs {
    int mode;                  // gvs_match_mode
    int nFrames, frame0;       // frames frame0 .. frame0 + nFrames - 1 of the extractors' last batch
    int nLevels;
    // map points: frame f owns [mpOff[f], mpOff[f + 1]); their data sits at index mpSlot[i] (or i) of the arrays below
    const int* mpOff;
    const int* mpSlot;         // optional
    const uint8_t* mpDesc;     // 32-byte rows
    const int* mpDescRow;      // optional: row of map point data index j inside mpDesc (else j)
    const float* predL;        // 2 floats per data index
    const float* predR;
    const int* lvlL;
    const int* lvlR;
    const uint8_t* inFrame;    // optional (all true)
    const uint8_t* inFrameR;
    // the frames' keys (dense arrays of the extractors, offsets on the device)
    const gvs_keypoint* kpL; const gvs_keypoint* kpR;
    const uint8_t* descL; const uint8_t* descR;
    const int* offL; const int* offR;
    const int* cellStartL; const int* cellStartR;   // per frame (maxCells + 1) entries
    const int* itemsL; const int* itemsR;           // dense, key offsets
    const int* leftIdxs; const int* rightIdxs;      // stereo partner arrays (mode 0), dense
    int maxCells, xGrids, yGrids;
    float xMult, yMult, rad;
    float scalePyramid[kMaxLevels];
    // scratch + in-out
    TRec* recs;                // 2 per map point
    int* matchedL; int* matchedR;   // dense: entry of key k of frame f at (off[f] - off[frame0]) + k
    int* matches;              // 2 per map point
    int* nMatches;             // per frame
    int* status;
    int capL, capR;            // shared-memory capacity of the resolve kernel (keys per frame and side)
};

struct TWindow {
    int minX, maxX, minY, maxY, lvl;
    float trackX, trackY, radius;
    bool ok;
};

// getMatchIdxs window (src/FeatureMatcher.cpp:18-30)
__device__ __forceinline__ TWindow tWindow(const ProjBatchArgs& a, float trackX, float trackY, int lvl, bool inFrame) {
    TWindow w;
    w.trackX = trackX; w.trackY = trackY; w.lvl = lvl;
    w.radius = __fmul_rn(a.scalePyramid[min(max(lvl, 0), kMaxLevels - 1)], a.rad);
    w.minX = max(0, (int)floorf(__fmul_rn(__fsub_rn(trackX, w.radius), a.xMult)));
    w.maxX = min(a.xGrids - 1, (int)ceilf(__fmul_rn(__fadd_rn(trackX, w.radius), a.xMult)));
    w.minY = max(0, (int)floorf(__fmul_rn(__fsub_rn(trackY, w.radius), a.yMult)));
    w.maxY = min(a.yGrids - 1, (int)ceilf(__fmul_rn(__fadd_rn(trackY, w.radius), a.yMult)));
    w.ok = inFrame && !(w.minX >= a.xGrids || w.minY >= a.yGrids || w.maxX < 0 || w.maxY < 0) && w.minX <= w.maxX;
    return w;
}

struct TSide {
    const char* kp;            // keys of the frame: x, y as floats at +0 / +4, the octave at +octOfs, kpStride bytes apart
    int kpStride, octOfs;      // (28, 20: gvs_keypoint in global memory; 12, 8: the staged copy in shared memory)
    const uint8_t* descs;
    const int* cellStart;
    const int* items;
};
struct TKey3 { float x, y; int octave; };   // staged key

// One THREAD: the ordered candidate list of a window (grid rows top to bottom, one contiguous CSR range per row) reduced to its
// kTopK smallest keys (entries in key order) and its length.  `skip` (optional) = matchedIdxs of the side, indexed by key: entries
// >= 0 are not candidates (src/FeatureMatcher.cpp:283,323,418,485).  A window holds a few dozen keys at the tracking radii, so a
// thread per (map point, side) keeps 32 independent walks of a warp in flight where a warp per walk would keep one.
__device__ __forceinline__ void tScanLane(const TSide& sd, const TWindow& w, int xGrids, const uint4& q0, const uint4& q1, const int* skip,
                                          bool radiusMode, TRec& rec) {
    unsigned k[kTopK];
#pragma unroll
    for (int i = 0; i < kTopK; i++) { k[i] = kNoCand; rec.e[i] = kNoCand; }
    rec.n = 0;
    if (!w.ok) return;
    int run = 0;
    for (int row = w.minY; row <= w.maxY; row++) {
        const int p0 = sd.cellStart[row * xGrids + w.minX], p1 = sd.cellStart[row * xGrids + w.maxX + 1];
        for (int p = p0; p < p1; p++) {
            const int idx = sd.items[p];
            const char* kp = sd.kp + (size_t)idx * sd.kpStride;
            const int lev = *reinterpret_cast<const int*>(kp + sd.octOfs);
            if (lev > w.lvl + 1 || lev < w.lvl - 1) continue;
            if (skip && skip[idx] >= 0) continue;
            const float kx = *reinterpret_cast<const float*>(kp), ky = *reinterpret_cast<const float*>(kp + 4);
            const float dx = __fsub_rn(kx, w.trackX), dy = __fsub_rn(ky, w.trackY);
            if (!(fabsf(dx) < w.radius && fabsf(dy) < w.radius)) continue;
            if (radiusMode) {  // Converter::checkPixelParallax (include/Conversions.h:140-144)
                const double ex = (double)w.trackX - (double)kx, ey = (double)w.trackY - (double)ky;
                if (!(sqrt(ex * ex + ey * ey) > 10.0)) continue;
            }
            const uint4* dk = reinterpret_cast<const uint4*>(sd.descs + (size_t)idx * 32);
            const uint4 k0 = __ldg(dk), k1 = __ldg(dk + 1);
            const int d = __popc(q0.x ^ k0.x) + __popc(q0.y ^ k0.y) + __popc(q0.z ^ k0.z) + __popc(q0.w ^ k0.w) +
                          __popc(q1.x ^ k1.x) + __popc(q1.y ^ k1.y) + __popc(q1.z ^ k1.z) + __popc(q1.w ^ k1.w);
            unsigned key = ((unsigned)d << 20) | (unsigned)run, ent = tPack(idx, d, lev);
            ++run;
            if (key < k[kTopK - 1]) {   // insert into the ascending list
#pragma unroll
                for (int i = 0; i < kTopK; i++) {
                    if (key < k[i]) { const unsigned tk = k[i], te = rec.e[i]; k[i] = key; rec.e[i] = ent; key = tk; ent = te; }
                }
            }
        }
    }
    rec.n = run;
}

// everything a thread needs to scan one (map point, side) of frame `frame`
struct TTask {
    TSide sd;
    TWindow w;
    uint4 q0, q1;
};

__device__ __forceinline__ TTask tMakeTask(const ProjBatchArgs& a, int frame, int mpi, int side, int* status) {
    TTask t;
    const int j = a.mpSlot ? a.mpSlot[mpi] : mpi;
    const float* pred = (side ? a.predR : a.predL) + 2 * (size_t)j;
    int lvl = side ? a.lvlR[j] : a.lvlL[j];
    if (lvl < 0 || lvl >= a.nLevels) {   // the reference indexes scalePyramid unchecked: reported as GVS_ERR_INVALID
        atomicOr(status, kStatProjOverflow);
        lvl = min(max(lvl, 0), a.nLevels - 1);
    }
    const bool inFrame = a.mode == GVS_MATCH_RADIUS ? true
                                                    : (side ? (a.inFrameR ? a.inFrameR[j] != 0 : true) : (a.inFrame ? a.inFrame[j] != 0 : true));
    t.w = tWindow(a, pred[0], pred[1], lvl, inFrame);
    const int o = side ? a.offR[frame] : a.offL[frame];
    t.sd.kp = reinterpret_cast<const char*>((side ? a.kpR : a.kpL) + o);
    t.sd.kpStride = (int)sizeof(gvs_keypoint); t.sd.octOfs = 20;
    t.sd.descs = (side ? a.descR : a.descL) + (size_t)o * 32;
    t.sd.cellStart = (side ? a.cellStartR : a.cellStartL) + (size_t)frame * (a.maxCells + 1);
    t.sd.items = (side ? a.itemsR : a.itemsL) + o;
    const int row = a.mpDescRow ? a.mpDescRow[j] : j;
    const uint4* dq = reinterpret_cast<const uint4*>(a.mpDesc + (size_t)row * 32);
    t.q0 = __ldg(dq); t.q1 = __ldg(dq + 1);
    return t;
}

// phase 1.  CTA = (slice of 256 map points of a frame, frame, side), thread = one window walk.  Every walk of the CTA reads the same
// key grid, keys and matched flags, so they are staged in shared memory once (cellStart, items, {x, y, octave}, matchedIdxs: 20 bytes
// per key + 4 per grid cell); a walk then only goes to global memory for its map point and for the descriptors of the candidates
// that pass the window test.  STAGED = false (frames with more keys than the shared memory holds): the same walk on global memory.
constexpr int kCandThreads = 256;
template <bool STAGED>
__global__ void __launch_bounds__(kCandThreads)
projb_candidates_kernel(const ProjBatchArgs a, int capKeys) {
    extern __shared__ __align__(16) uint8_t tsm[];
    const int f = blockIdx.y, frame = a.frame0 + f, side = blockIdx.z;
    const int tid = threadIdx.x;
    const int mp0 = a.mpOff[f], M = a.mpOff[f + 1] - mp0;
    const int i0 = blockIdx.x * kCandThreads;
    if (i0 >= M) return;
    const int o = side ? a.offR[frame] : a.offL[frame], N = (side ? a.offR[frame + 1] : a.offL[frame + 1]) - o;
    const int* gSkip = (side ? a.matchedR + (a.offR[frame] - a.offR[a.frame0]) : a.matchedL + (a.offL[frame] - a.offL[a.frame0]));
    TSide sd;
    sd.descs = (side ? a.descR : a.descL) + (size_t)o * 32;
    const int* gCell = (side ? a.cellStartR : a.cellStartL) + (size_t)frame * (a.maxCells + 1);
    const int* gItems = (side ? a.itemsR : a.itemsL) + o;
    const gvs_keypoint* gKp = (side ? a.kpR : a.kpL) + o;
    const int* skip = gSkip;
    if (STAGED && N <= capKeys) {
        const int nCells = a.xGrids * a.yGrids + 1;
        int* sCell = reinterpret_cast<int*>(tsm);
        int* sItems = sCell + ((nCells + 3) & ~3);
        int* sSkip = sItems + capKeys;
        TKey3* sKp = reinterpret_cast<TKey3*>(sSkip + capKeys);
        for (int i = tid; i < nCells; i += kCandThreads) sCell[i] = gCell[i];
        for (int i = tid; i < N; i += kCandThreads) {
            sItems[i] = gItems[i];
            sSkip[i] = gSkip[i];
            const gvs_keypoint k = gKp[i];
            sKp[i] = TKey3{k.x, k.y, k.octave};
        }
        __syncthreads();
        sd.kp = reinterpret_cast<const char*>(sKp); sd.kpStride = (int)sizeof(TKey3); sd.octOfs = 8;
        sd.cellStart = sCell; sd.items = sItems; skip = sSkip;
    } else {
        sd.kp = reinterpret_cast<const char*>(gKp); sd.kpStride = (int)sizeof(gvs_keypoint); sd.octOfs = 20;
        sd.cellStart = gCell; sd.items = gItems;
    }
    const bool radiusMode = a.mode == GVS_MATCH_RADIUS;
    // the CTA's slices of 256 map points: blockIdx.x, blockIdx.x + gridDim.x, ... (large batches run one CTA per frame and side, so
    // that the frame is staged once; small ones spread a frame over several CTAs)
    for (int i = i0 + tid; i < M; i += gridDim.x * kCandThreads) {
    const int mpi = mp0 + i;
    TRec rec;
#pragma unroll
    for (int q = 0; q < kTopK; q++) rec.e[q] = kNoCand;
    rec.n = 0;
    // a map point that is matched already is skipped by the walk (:262,:399); matchByRadius has no such test
    const bool skipMP = !radiusMode && (a.matches[2 * (size_t)mpi] >= 0 || a.matches[2 * (size_t)mpi + 1] >= 0);
    if (!skipMP) {
        const int j = a.mpSlot ? a.mpSlot[mpi] : mpi;
        const float* pred = (side ? a.predR : a.predL) + 2 * (size_t)j;
        int lvl = side ? a.lvlR[j] : a.lvlL[j];
        if (lvl < 0 || lvl >= a.nLevels) {   // the reference indexes scalePyramid unchecked: reported as GVS_ERR_INVALID
            atomicOr(a.status, kStatProjOverflow);
            lvl = min(max(lvl, 0), a.nLevels - 1);
        }
        const bool inFrame = radiusMode ? true : (side ? (a.inFrameR ? a.inFrameR[j] != 0 : true) : (a.inFrame ? a.inFrame[j] != 0 : true));
        const TWindow w = tWindow(a, pred[0], pred[1], lvl, inFrame);
        const int row = a.mpDescRow ? a.mpDescRow[j] : j;
        const uint4* dq = reinterpret_cast<const uint4*>(a.mpDesc + (size_t)row * 32);
        const uint4 q0 = __ldg(dq), q1 = __ldg(dq + 1);
        tScanLane(sd, w, a.xGrids, q0, q1, skip, radiusMode, rec);
    }
    a.recs[2 * (size_t)mpi + side] = rec;
    }
}

struct TDecision { int bestDist, bestIdx, bestLev, bestLev2, secDist; };
// best = smallest remaining key, second best = second smallest (see the header of this file)
__device__ __forceinline__ TDecision tDecide(unsigned e1, unsigned e2) {
    TDecision d{256, -1, -1, -1, 256};
    if (e1 == kNoCand) return d;
    d.bestDist = tDist(e1); d.bestIdx = tIdx(e1); d.bestLev = tLev(e1);
    if (e2 != kNoCand) { d.secDist = tDist(e2); d.bestLev2 = tLev(e2); }
    return d;
}
// the first two stored candidates that are still free; false = the stored list cannot answer (it was truncated and fewer than
// two of its entries are left): enumerate the window again
__device__ __forceinline__ bool tFirstTwo(const TRec& r, const int* matched, unsigned& e1, unsigned& e2) {
    e1 = e2 = kNoCand;
#pragma unroll
    for (int i = 0; i < kTopK; i++) {
        const unsigned e = r.e[i];
        if (e == kNoCand || matched[tIdx(e)] >= 0) continue;
        if (e1 == kNoCand) e1 = e; else if (e2 == kNoCand) e2 = e;
    }
    return e2 != kNoCand || r.n <= kTopK;
}

// phase 2: ONE WARP per frame, lane = map point of the current chunk of 32.  Deciding from the stored candidates is a few
// shared-memory look-ups per map point, so a thread per map point is enough; only the rare re-enumeration of a window is done
// by the whole warp.  The walk is serial per frame (that is the reference's semantics), frames run in parallel.
// MARKS: the round's tentative marks live in shared memory next to the matched flags (frames of up to ~14 k keys per side); without
// them (very large key counts: the flags alone fill the shared memory) conflicts are found by comparing lanes pairwise.
template <bool MARKS>
__global__ void __launch_bounds__(32)
projb_resolve_kernel(const ProjBatchArgs a) {
    extern __shared__ int sMatched[];  // 2 x (capL + capR)
    const int f = blockIdx.x, frame = a.frame0 + f;
    const int lane = threadIdx.x;
    const bool stereo = a.mode == GVS_MATCH_STEREO_PRED;
    const int mp0 = a.mpOff[f], M = a.mpOff[f + 1] - mp0;
    const int oL = a.offL[frame], NL = a.offL[frame + 1] - oL;
    const int oR = stereo ? a.offR[frame] : 0, NR = stereo ? a.offR[frame + 1] - oR : 0;
    if (M <= 0) { if (lane == 0) a.nMatches[f] = 0; return; }
    if (NL > a.capL || NR > a.capR) { if (lane == 0) { atomicOr(a.status, kStatOutputOverflow); a.nMatches[f] = 0; } return; }
    int* mL = sMatched;
    int* mR = sMatched + a.capL;
    int* tL = sMatched + a.capL + a.capR;   // per key: the first lane of the current round that wants to mark it (kNoLane = nobody)
    int* tR = tL + a.capL;
    constexpr int kNoLane = 0x7fffffff;
    int* gML = a.matchedL + (oL - a.offL[a.frame0]);
    int* gMR = stereo ? a.matchedR + (oR - a.offR[a.frame0]) : nullptr;
    for (int i = lane; i < NL; i += 32) { mL[i] = gML[i]; if (MARKS) tL[i] = kNoLane; }
    for (int i = lane; i < NR; i += 32) { mR[i] = gMR[i]; if (MARKS) tR[i] = kNoLane; }
    __syncwarp();
    const int* leftIdxs = stereo ? a.leftIdxs + oR : nullptr;
    const int* rightIdxs = stereo ? a.rightIdxs + oL : nullptr;
    const bool radiusMode = a.mode == GVS_MATCH_RADIUS;
    auto keyOf = [&](unsigned e) { return e == kNoCand ? -1 : tIdx(e); };
    int total = 0;
    // the records of a chunk are requested one chunk ahead: the walk is a serial chain per frame and a chunk is only a few hundred
    // cycles of work, less than one trip to L2
    auto loadChunk = [&](int base, TRec& rl, TRec& rr, bool& skipMP) {
#pragma unroll
        for (int i = 0; i < kTopK; i++) { rl.e[i] = kNoCand; rr.e[i] = kNoCand; }
        rl.n = 0; rr.n = 0;
        skipMP = true;
        if (base + lane < M) {
            const int mpi = mp0 + base + lane;
            // a map point that is matched already is skipped by the walk (:262,:399); matchByRadius has no such test
            skipMP = !radiusMode && (a.matches[2 * (size_t)mpi] >= 0 || a.matches[2 * (size_t)mpi + 1] >= 0);
            rl = a.recs[2 * (size_t)mpi];
            if (stereo) rr = a.recs[2 * (size_t)mpi + 1];
        }
    };
    TRec nrl, nrr;
    bool nskip;
    loadChunk(0, nrl, nrr, nskip);
    for (int base = 0; base < M; base += 32) {
        const int chunk = min(32, M - base);
        const int mpi = mp0 + base + lane;
        TRec rl = nrl, rr = nrr;
        const bool skipMP = nskip;
        loadChunk(base + 32, nrl, nrr, nskip);   // (matches[] of a later chunk is not written by this one: every lane writes its own map point)
        int done = 0;
        while (done < chunk) {
            const bool pend = lane >= done && lane < chunk && !skipMP;
            // ---- (A) tentative decision of every pending map point against the current matched arrays
            unsigned l1 = kNoCand, l2 = kNoCand, r1 = kNoCand, r2 = kNoCand;
            bool needL = false, needR = false;
            if (pend) {
                needL = !tFirstTwo(rl, mL, l1, l2);
                if (stereo) needR = !tFirstTwo(rr, mR, r1, r2);
            }
            if (needL) {   // rare: the stored list ran dry
                const TTask t = tMakeTask(a, frame, mpi, 0, a.status);
                tScanLane(t.sd, t.w, a.xGrids, t.q0, t.q1, mL, radiusMode, rl);
                l1 = rl.e[0]; l2 = rl.e[1];
            }
            if (needR) {
                const TTask t = tMakeTask(a, frame, mpi, 1, a.status);
                tScanLane(t.sd, t.w, a.xGrids, t.q0, t.q1, mR, false, rr);
                r1 = rr.e[0]; r2 = rr.e[1];
            }
            int accept = 0, takeL = -1, takeR = -1, out0 = -1, out1 = -1;
            if (pend) {
                TDecision d = tDecide(l1, l2);
                bool right = false;
                int bestIdxR = -1;
                if (stereo) {
                    const TDecision dr = tDecide(r1, r2);
                    bestIdxR = dr.bestIdx;
                    if (d.bestDist > dr.bestDist) {   // :353-360
                        d.bestDist = dr.bestDist; d.secDist = dr.secDist; d.bestLev = dr.bestLev; d.bestLev2 = dr.bestLev2;
                        right = true;
                    }
                }
                bool ok;
                if (a.mode == GVS_MATCH_MONO) {
                    // (ratioProj + 0.1) * secDist is evaluated in double (float + double), threshold 100 + 50 (:442-447)
                    const double thr = ((double)0.8f + 0.1) * (double)d.secDist;
                    ok = d.bestDist <= 150 && !(d.bestLev == d.bestLev2 && (double)d.bestDist >= thr) &&
                         (d.bestLev != d.bestLev2 || (double)d.bestDist < thr);
                } else {
                    const float thr = __fmul_rn(0.8f, (float)d.secDist);
                    ok = d.bestDist <= 100 && !(d.bestLev == d.bestLev2 && (float)d.bestDist >= thr) &&
                         (d.bestLev != d.bestLev2 || (float)d.bestDist < thr);
                }
                if (ok) {
                    accept = 1;
                    if (!stereo) { takeL = d.bestIdx; out0 = d.bestIdx; }
                    else if (right) {
                        takeR = bestIdxR; out1 = bestIdxR;
                        const int li = leftIdxs[bestIdxR];
                        if (li >= 0) { out0 = li; takeL = li; }
                    } else {
                        takeL = d.bestIdx; out0 = d.bestIdx;
                        const int ri = rightIdxs[d.bestIdx];
                        if (ri >= 0) { out1 = ri; takeR = ri; }
                    }
                }
            }
            // ---- (B) longest prefix of the pending lanes in which nobody takes a deciding candidate of a later lane: every accepting
            // lane leaves its lane number on the keys it would mark (minimum wins), every pending lane looks at its deciding keys
            const int dl0 = keyOf(l1), dl1 = keyOf(l2), dr0 = keyOf(r1), dr1 = keyOf(r2);
            bool unsafe = false;
            if (MARKS) {
                if (accept) {
                    if (takeL >= 0) atomicMin(&tL[takeL], lane);
                    if (takeR >= 0) atomicMin(&tR[takeR], lane);
                }
                __syncwarp();
                if (pend) {
                    if (dl0 >= 0 && tL[dl0] < lane) unsafe = true;
                    if (dl1 >= 0 && tL[dl1] < lane) unsafe = true;
                    if (dr0 >= 0 && tR[dr0] < lane) unsafe = true;
                    if (dr1 >= 0 && tR[dr1] < lane) unsafe = true;
                }
                __syncwarp();
                if (accept) {
                    if (takeL >= 0) tL[takeL] = kNoLane;
                    if (takeR >= 0) tR[takeR] = kNoLane;
                }
            } else {
                unsigned accMask = __ballot_sync(0xffffffffu, accept != 0);
                accMask &= ~0u << done;
                for (unsigned am = accMask; am; am &= am - 1) {
                    const int j = __ffs(am) - 1;
                    const int tl = __shfl_sync(0xffffffffu, takeL, j), tr = __shfl_sync(0xffffffffu, takeR, j);
                    if (lane > j && pend) {
                        if (tl >= 0 && (tl == dl0 || tl == dl1)) unsafe = true;
                        if (tr >= 0 && (tr == dr0 || tr == dr1)) unsafe = true;
                    }
                }
            }
            const unsigned um = __ballot_sync(0xffffffffu, unsafe);
            const int upto = um ? __ffs(um) - 1 : chunk;     // lanes [done, upto) are final
            const bool mine = lane >= done && lane < upto && accept;
            const int i = base + lane;                       // the reference's map point index inside the frame
            // several committed map points may mark the same key (a stereo partner): the last one in index order wins
            const int tl = mine ? takeL : -1, tr = mine ? takeR : -1;
            const unsigned pl = __match_any_sync(0xffffffffu, tl >= 0 ? tl : -1 - lane);
            const unsigned pr = __match_any_sync(0xffffffffu, tr >= 0 ? tr : -1 - lane);
            if (tl >= 0 && (31 - __clz(pl)) == lane) mL[tl] = i;
            if (tr >= 0 && (31 - __clz(pr)) == lane) mR[tr] = i;
            if (mine) {
                int* mt = a.matches + 2 * (size_t)mpi;
                if (!stereo) mt[0] = out0;
                else {
                    if (out0 >= 0) mt[0] = out0;
                    if (out1 >= 0) mt[1] = out1;
                }
            }
            total += __popc(__ballot_sync(0xffffffffu, mine));
            done = upto;
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = lane; i < NL; i += 32) gML[i] = mL[i];
    for (int i = lane; i < NR; i += 32) gMR[i] = mR[i];
    if (lane == 0) a.nMatches[f] = total;
}

static void launchProjectBatch(const ProjBatchArgs& a, int maxMpPerFrame, cudaStream_t s, int* launches) {
    if (a.nFrames <= 0) return;
    const int nSides = a.mode == GVS_MATCH_STEREO_PRED ? 2 : 1;
    // CTAs per (frame, side): enough of them to fill the GPU (148 SMs x ~4 CTAs), at most one per 256 map points
    const int maxSlices = std::max((maxMpPerFrame + kCandThreads - 1) / kCandThreads, 1);
    const int nSlices = std::min(maxSlices, std::max(1, (592 + a.nFrames * nSides - 1) / (a.nFrames * nSides)));
    const int capKeys = std::max(a.capL, a.capR), nCells = a.xGrids * a.yGrids + 1;
    const size_t stage = (size_t)((nCells + 3) & ~3) * 4 + (size_t)capKeys * (4 + 4 + sizeof(TKey3)) + 16;
    static size_t cfgA[64], cfgB[64];
    if (stage <= 160 * 1024) {
        ensureDynamicSmem(projb_candidates_kernel<true>, stage, cfgA);
        projb_candidates_kernel<true><<<dim3(nSlices, a.nFrames, nSides), kCandThreads, stage, s>>>(a, capKeys);
    } else {
        projb_candidates_kernel<false><<<dim3(nSlices, a.nFrames, nSides), kCandThreads, 0, s>>>(a, capKeys);
    }
    ++*launches;
    const size_t flags = (size_t)(a.capL + a.capR) * sizeof(int);   // matched flags (+ as much again for the round's tentative marks)
    static size_t cfgC[64];
    if (2 * flags <= 200 * 1024) {
        ensureDynamicSmem(projb_resolve_kernel<true>, 2 * flags, cfgB);
        projb_resolve_kernel<true><<<a.nFrames, 32, 2 * flags, s>>>(a);
    } else {
        ensureDynamicSmem(projb_resolve_kernel<false>, flags, cfgC);
        projb_resolve_kernel<false><<<a.nFrames, 32, flags, s>>>(a);
    }
    ++*launches;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Map points of frame f = the stereo keys of frame prev[f] (initializeMap), projected into frame f's two cameras
// (removeOutOfFrameMPs).  Data is written per slot = f * slotStride + key index; the ordered list of surviving slots
// (activeMapPoints after the compaction at src/FeatureTracker.cpp:926-939) is built by tmp_compact_kernel.
struct TemporalArgs {
    int nFrames, slotStride, nLevels;
    const int* prev;            // nFrames
    const double* posePrev;     // nFrames x 16: zedPtr->mCameraPose.pose when frame prev[f] created its points
    const double* toCamera;     // nFrames x 16: predNPose^-1
    const double* toCameraR;    // nFrames x 16: (predNPose * extrinsics)^-1
    double fx, fy, cx, cy;
    int width, height;
    float logScale;             // kF->logScale = log(feLeft->imScale) (src/FeatureTracker.cpp:79)
    float scalePyramid[kMaxLevels];
    const gvs_keypoint* kpL; const int* offL;
    const float* depth;         // TrackedKeys::estimatedDepth, dense left offsets
    // per slot
    uint8_t* flag;
    float* predL; float* predR; int* lvlL; int* lvlR; int* descRow;
    double* wPos; float* maxScaleDist;
    // per frame
    int* mpCnt; int* mpOff;
    int* mpSlot; int* mpKey;    // dense map point lists
    int cap;                    // capacity of the dense lists
    int* status;
};

__global__ void __launch_bounds__(128)
tmp_project_kernel(const TemporalArgs a) {
    const int f = blockIdx.y, p = a.prev[f];
    if (p < 0) return;
    const int o = a.offL[p], n = a.offL[p + 1] - o;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n && i < a.slotStride) {
        const size_t slot = (size_t)f * a.slotStride + i;
        const float depth = a.depth[o + i];
        if (depth > 0.f) {   // src/FeatureTracker.cpp:93
            const gvs_keypoint k = a.kpL[o + i];
            // :95-100 (double): p = pose * ((x - cx) * zp / fx, (y - cy) * zp / fy, zp, 1)
            const double zp = (double)depth;
            double c[4];
            c[0] = __ddiv_rn(__dmul_rn(__dsub_rn((double)k.x, a.cx), zp), a.fx);
            c[1] = __ddiv_rn(__dmul_rn(__dsub_rn((double)k.y, a.cy), zp), a.fy);
            c[2] = zp; c[3] = 1.0;
            const double* P = a.posePrev + 16 * (size_t)f;
            double w3[3];
            mulPose(P, c, w3);
            // Eigen evaluates all four rows; the last row of a rigid pose is (0, 0, 0, 1)
            double w[4] = {w3[0], w3[1], w3[2],
                           __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[12], c[0]), __dmul_rn(P[13], c[1])), __dmul_rn(P[14], c[2])), __dmul_rn(P[15], c[3]))};
            // MapPoint::update (src/Map.cpp:61-104): dist = |wp3d - kF translation| (float), maxScaleDist = dist * scaleFactor[level]
            const double ex = __dsub_rn(w[0], P[3]), ey = __dsub_rn(w[1], P[7]), ez = __dsub_rn(w[2], P[11]);
            const float dist = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
            const float maxScaleDist = __fmul_rn(dist, a.scalePyramid[min(max(k.octave, 0), kMaxLevels - 1)]);
            float pl[2], pr[2];
            int ll = 0, lr = 0;
            const bool c1 = worldToFrameOne(a.toCamera + 16 * (size_t)f, w, a.fx, a.fy, a.cx, a.cy, a.width, a.height, maxScaleDist, a.logScale,
                                            a.nLevels, pl, &ll);
            const bool c2 = worldToFrameOne(a.toCameraR + 16 * (size_t)f, w, a.fx, a.fy, a.cx, a.cy, a.width, a.height, maxScaleDist, a.logScale,
                                            a.nLevels, pr, &lr);
            keep = c1 && c2;   // :920-930
            if (keep) {
                a.predL[2 * slot] = pl[0]; a.predL[2 * slot + 1] = pl[1];
                a.predR[2 * slot] = pr[0]; a.predR[2 * slot + 1] = pr[1];
                a.lvlL[slot] = ll; a.lvlR[slot] = lr;
                a.descRow[slot] = o + i;
                a.wPos[4 * slot] = w[0]; a.wPos[4 * slot + 1] = w[1]; a.wPos[4 * slot + 2] = w[2]; a.wPos[4 * slot + 3] = w[3];
                a.maxScaleDist[slot] = maxScaleDist;
            }
        }
        a.flag[slot] = keep ? 1 : 0;
    }
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0 && cnt) atomicAdd(a.mpCnt + f, cnt);
}

// exclusive scan of the per-frame counts (one CTA; nFrames <= 65535)
__global__ void __launch_bounds__(1024)
tmp_scan_kernel(const TemporalArgs a) {
    __shared__ int sWarp[32];
    __shared__ int sRun;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sRun = 0;
    __syncthreads();
    for (int base = 0; base < a.nFrames; base += 1024) {
        const int f = base + tid;
        const int v = f < a.nFrames ? (a.prev[f] >= 0 ? a.mpCnt[f] : 0) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = sWarp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            sWarp[lane] = wi - w;
        }
        __syncthreads();
        const int excl = sRun + sWarp[warp] + incl - v;
        if (f < a.nFrames) a.mpOff[f] = excl;
        __syncthreads();
        if (tid == 1023) sRun = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        a.mpOff[a.nFrames] = sRun;
        if (sRun > a.cap) atomicOr(a.status, kStatOutputOverflow);
    }
}

// ordered compaction of the kept slots of one frame
__global__ void __launch_bounds__(256)
tmp_compact_kernel(const TemporalArgs a) {
    __shared__ int sWarp[8];
    __shared__ int sRun;
    const int f = blockIdx.x, p = a.prev[f];
    if (p < 0) return;
    const int n = min(a.offL[p + 1] - a.offL[p], a.slotStride);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = a.mpOff[f];
    if (a.mpOff[a.nFrames] > a.cap) return;
    if (tid == 0) sRun = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + tid;
        const bool k = i < n && a.flag[(size_t)f * a.slotStride + i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int before = sRun;
        for (int w = 0; w < warp; w++) before += sWarp[w];
        if (k) {
            const int r = o + before + __popc(m & ((1u << lane) - 1u));
            a.mpSlot[r] = f * a.slotStride + i;
            a.mpKey[r] = i;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += sWarp[w]; sRun += t; }
        __syncthreads();
    }
}

}  // namespace gvs

using namespace gvs;

namespace {

size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// the grid-related part of ProjBatchArgs shared by all entry points
void fillFrameArgs(gvs_matcher* m, ProjBatchArgs& a, int mode, float rad) {
    const bool two = mode == GVS_MATCH_STEREO_PRED;
    a.mode = mode; a.rad = rad; a.nLevels = m->L->g.nLevels;
    a.kpL = m->L->kps; a.descL = m->L->desc; a.offL = m->L->imgOff;
    a.kpR = two ? m->R->kps : nullptr; a.descR = two ? m->R->desc : nullptr; a.offR = two ? m->R->imgOff : m->L->imgOff;
    a.cellStartL = m->cellStartL; a.itemsL = m->itemsL;
    a.cellStartR = two ? m->cellStartR : nullptr; a.itemsR = two ? m->itemsR : nullptr;
    a.leftIdxs = two ? m->leftIdxs : nullptr; a.rightIdxs = two ? m->rightIdxs : nullptr;
    a.maxCells = m->maxCells; a.xGrids = m->xGrids; a.yGrids = m->yGrids; a.xMult = m->xMult; a.yMult = m->yMult;
    for (int i = 0; i < kMaxLevels; i++) a.scalePyramid[i] = i < m->L->g.nLevels ? m->L->scale[i] : 1.f;
    a.capL = m->L->g.keptPerImage;
    a.capR = two ? m->R->g.keptPerImage : 0;
}

}  // namespace

namespace gvs {

// The state checks every call that reads the resident frame shares (the grids / stereo arrays must belong to the
// extractors' current batch).
int checkResident(gvs_matcher* m, bool needStereo, int frame0, int nFrames) {
    if (!m->gridsValid || m->gridGenL != m->L->generation || (m->R != m->L && m->gridGenR != m->R->generation))
        return fail(GVS_ERR_STATE, "call gvs_assign_grids_batch for the current batch first");
    if (needStereo && (!m->stereoValid || m->stereoGenL != m->L->generation || m->stereoGenR != m->R->generation))
        return fail(GVS_ERR_STATE, "call gvs_stereo_match_batch for the current batch first");
    if (frame0 < 0 || nFrames < 1 || frame0 + nFrames > m->gridFrames || frame0 + nFrames > m->L->lastN ||
        (needStereo && (frame0 + nFrames > m->lastFrames || frame0 + nFrames > m->R->lastN)))
        return fail(GVS_ERR_INVALID, "frame range outside the last batch");
    return GVS_OK;
}

}  // namespace gvs

extern "C" {

int gvs_match_projection_batch(gvs_matcher* m, int first_frame, int n_frames, int mode, const gvs_mappoints* mps, const int32_t* mp_offsets,
                               float rad, int32_t* matched_l, int32_t* matched_r, int32_t* matches, int32_t* n_matches) {
    NvtxScope nvtx("gvs:match_projection_batch");
    if (!m || !mps || !mp_offsets || !matched_l || !matches || !n_matches) return fail(GVS_ERR_INVALID, "null argument");
    if (mode < 0 || mode > 2) return fail(GVS_ERR_INVALID, "bad mode");
    const bool two = mode == GVS_MATCH_STEREO_PRED;
    if (two && (!matched_r || !mps->pred_r || !mps->scale_level_r))
        return fail(GVS_ERR_STATE, "stereo-pred matching needs matched_r, pred_r and scale_level_r");
    int rc = checkResident(m, two, first_frame, n_frames);
    if (rc) return rc;
    const int M = mps->count;
    if (M < 0 || mp_offsets[0] != 0 || mp_offsets[n_frames] != M) return fail(GVS_ERR_INVALID, "mp_offsets must run from 0 to mps->count");
    int maxMp = 0;
    for (int f = 0; f < n_frames; f++) {
        if (mp_offsets[f + 1] < mp_offsets[f]) return fail(GVS_ERR_INVALID, "mp_offsets must be non-decreasing");
        maxMp = std::max(maxMp, mp_offsets[f + 1] - mp_offsets[f]);
    }
    for (int f = 0; f < n_frames; f++) n_matches[f] = 0;
    if (M == 0) return GVS_OK;
    if (!mps->desc || !mps->pred_l || !mps->scale_level_l) return fail(GVS_ERR_INVALID, "map point arrays missing");
    rc = gvs_matcher_sync(m);   // the host offsets (hOff) size the in-out arrays
    if (rc) return rc;
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->L->stream;
    const int oL = m->L->hOff[first_frame], NL = m->L->hOff[first_frame + n_frames] - oL;
    const int oR = two ? m->R->hOff[first_frame] : 0, NR = two ? m->R->hOff[first_frame + n_frames] - oR : 0;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += al256(std::max<size_t>(bytes, 4)); return at; };
    const size_t oDesc = take((size_t)M * 32), oPL = take((size_t)M * 8), oPR = take((size_t)M * 8), oLL = take((size_t)M * 4),
                 oLR = take((size_t)M * 4), oIF = take(M), oIR = take(M), oML = take((size_t)NL * 4), oMR = take((size_t)NR * 4),
                 oMt = take((size_t)M * 8), oRec = take((size_t)M * 2 * sizeof(TRec)), oOff = take((size_t)(n_frames + 1) * 4),
                 oCnt = take((size_t)n_frames * 4), oSt = take(16);
    // A search radius that covers a large part of the image (matchByProjectionMono with rad 1200, BASELINE config 4: hundreds of
    // candidates per map point and most keys contested) goes to the stored-list kernels of project.cu: a thread per window and a
    // re-enumeration per conflict (temporal.cu) are built for the tracking radii, where a window holds a few dozen keys.
    bool wide = false;
    size_t poolNeed = 0;
    {
        const float r = rad * m->L->scale[std::min(2, m->L->g.nLevels - 1)];
        const float wx = std::min((float)m->xGrids, 2.f * r * m->xMult + 2.f), wy = std::min((float)m->yGrids, 2.f * r * m->yMult + 2.f);
        const float keysPerFrame = (float)NL / (float)n_frames;
        wide = keysPerFrame * 0.5f * (wx * wy) / ((float)m->xGrids * (float)m->yGrids) > 128.f;
        // a handful of frames (the per-frame call of the tracking thread): the stored-list resolve spends 32 warps on a frame where
        // the other spends one -- measured on one EuRoC frame, 1937 map points, rad 10: 0.53 ms against 0.91 ms host to host
        if (n_frames <= 4) wide = true;
        if (const char* e = getenv("GVS_PROJ_WIDE")) wide = atoi(e) != 0;   // (tests drive both paths)
        if (wide)
            for (int f = 0; f < n_frames; f++) {
                const int frame = first_frame + f;
                const int nl = m->L->hOff[frame + 1] - m->L->hOff[frame], nr = two ? m->R->hOff[frame + 1] - m->R->hOff[frame] : 0;
                poolNeed += 2 * (size_t)(mp_offsets[f + 1] - mp_offsets[f]) * (size_t)std::max(std::max(nl, nr), 1);
            }
        if (poolNeed > ((size_t)1 << 30)) { wide = false; poolNeed = 0; }   // (4 GB of lists: take the other path)
    }
    const size_t oFr = take(wide ? (size_t)n_frames * sizeof(ProjArgs) : 4), oPT = take(wide ? (size_t)n_frames * 8 : 4);
    rc = growScratch(m, o);
    if (rc) return rc;
    if (wide) {
        for (int i = 0; i < M; i++) {
            const int l = mps->scale_level_l[i];
            if (l < 0 || l >= m->L->g.nLevels || (two && (mps->scale_level_r[i] < 0 || mps->scale_level_r[i] >= m->L->g.nLevels)))
                return fail(GVS_ERR_INVALID, "map point scale level out of range");
        }
        if (poolNeed > (size_t)m->poolCap * 2) {
            GVS_CUDA_OK(cudaStreamSynchronize(s));
            devFree(m->pool); m->pool = nullptr; m->poolCap = 0;
            GVS_CUDA_OK(devAlloc(&m->pool, poolNeed));
            m->poolCap = (int)(poolNeed / 2);
        }
    }
    uint8_t* b = m->pj;
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        return (src && bytes) ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    GVS_CUDA_OK(up(oDesc, mps->desc, (size_t)M * 32));
    GVS_CUDA_OK(up(oPL, mps->pred_l, (size_t)M * 8));
    GVS_CUDA_OK(up(oLL, mps->scale_level_l, (size_t)M * 4));
    GVS_CUDA_OK(up(oIF, mps->in_frame, (size_t)M));
    if (two) {
        GVS_CUDA_OK(up(oPR, mps->pred_r, (size_t)M * 8));
        GVS_CUDA_OK(up(oLR, mps->scale_level_r, (size_t)M * 4));
        GVS_CUDA_OK(up(oIR, mps->in_frame_r, (size_t)M));
        GVS_CUDA_OK(up(oMR, matched_r, (size_t)NR * 4));
    }
    GVS_CUDA_OK(up(oML, matched_l, (size_t)NL * 4));
    GVS_CUDA_OK(up(oMt, matches, (size_t)M * 8));
    GVS_CUDA_OK(up(oOff, mp_offsets, (size_t)(n_frames + 1) * 4));
    GVS_CUDA_OK(cudaMemsetAsync(b + oSt, 0, 16, s));
    ProjBatchArgs a{};
    fillFrameArgs(m, a, mode, rad);
    a.nFrames = n_frames; a.frame0 = first_frame;
    a.mpOff = reinterpret_cast<int*>(b + oOff); a.mpSlot = nullptr; a.mpDesc = b + oDesc; a.mpDescRow = nullptr;
    a.predL = reinterpret_cast<float*>(b + oPL); a.predR = reinterpret_cast<float*>(b + oPR);
    a.lvlL = reinterpret_cast<int*>(b + oLL); a.lvlR = reinterpret_cast<int*>(b + oLR);
    a.inFrame = mps->in_frame ? b + oIF : nullptr; a.inFrameR = (two && mps->in_frame_r) ? b + oIR : nullptr;
    a.recs = reinterpret_cast<TRec*>(b + oRec);
    a.matchedL = reinterpret_cast<int*>(b + oML); a.matchedR = reinterpret_cast<int*>(b + oMR);
    a.matches = reinterpret_cast<int*>(b + oMt); a.nMatches = reinterpret_cast<int*>(b + oCnt); a.status = reinterpret_cast<int*>(b + oSt);
    int launches = 0;
    if (wide) {
        // stored candidate lists (project.cu): one argument block per frame; pool regions of M_f x max(NL_f, NR_f) entries per side
        static_assert(sizeof(TRec) == 20, "the stored-list records of project.cu (ProjRec, 20 bytes) reuse the TRec scratch region");
        std::vector<ProjArgs> fr((size_t)n_frames);
        size_t poolAt = 0;
        for (int f = 0; f < n_frames; f++) {
            const int frame = first_frame + f, mp0 = mp_offsets[f], Mf = mp_offsets[f + 1] - mp0;
            const int fl = m->L->hOff[frame], nl = m->L->hOff[frame + 1] - fl;
            const int fr0 = two ? m->R->hOff[frame] : 0, nr = two ? m->R->hOff[frame + 1] - fr0 : 0;
            ProjArgs& q = fr[(size_t)f];
            q = ProjArgs{};
            q.mode = mode; q.M = Mf;
            q.mpDesc = b + oDesc + (size_t)mp0 * 32; q.predL = a.predL + 2 * (size_t)mp0; q.predR = a.predR + 2 * (size_t)mp0;
            q.lvlL = a.lvlL + mp0; q.lvlR = a.lvlR + mp0;
            q.inFrame = a.inFrame ? a.inFrame + mp0 : nullptr; q.inFrameR = a.inFrameR ? a.inFrameR + mp0 : nullptr;
            q.kpL = m->L->kps + fl; q.descL = m->L->desc + (size_t)fl * 32; q.NL = nl;
            q.kpR = two ? m->R->kps + fr0 : nullptr; q.descR = two ? m->R->desc + (size_t)fr0 * 32 : nullptr; q.NR = nr;
            q.cellStartL = m->cellStartL + (size_t)frame * (m->maxCells + 1); q.itemsL = m->itemsL + fl;
            q.cellStartR = two ? m->cellStartR + (size_t)frame * (m->maxCells + 1) : nullptr; q.itemsR = two ? m->itemsR + fr0 : nullptr;
            q.leftIdxs = two ? m->leftIdxs + fr0 : nullptr; q.rightIdxs = two ? m->rightIdxs + fl : nullptr;
            q.xGrids = a.xGrids; q.yGrids = a.yGrids; q.xMult = a.xMult; q.yMult = a.yMult; q.rad = rad;
            for (int i = 0; i < kMaxLevels; i++) q.scalePyramid[i] = a.scalePyramid[i];
            q.recs = b + oRec + (size_t)mp0 * 2 * sizeof(TRec);
            q.poolTop = reinterpret_cast<int*>(b + oPT) + 2 * f;
            const size_t cap = (size_t)Mf * (size_t)std::max(std::max(nl, nr), 1);
            q.candL = m->pool + poolAt; q.candR = m->pool + poolAt + cap; q.candCap = (int)cap;
            poolAt += 2 * cap;
            q.matchedL = a.matchedL + (fl - oL); q.matchedR = two ? a.matchedR + (fr0 - oR) : nullptr;
            q.matches = a.matches + 2 * (size_t)mp0; q.nMatches = a.nMatches + f; q.status = a.status + 2;
        }
        GVS_CUDA_OK(cudaMemcpyAsync(b + oFr, fr.data(), fr.size() * sizeof(ProjArgs), cudaMemcpyHostToDevice, s));   // (pageable: staged before the call returns)
        GVS_CUDA_OK(cudaMemsetAsync(b + oPT, 0, (size_t)n_frames * 8, s));
        launchProjectWide(reinterpret_cast<const ProjArgs*>(b + oFr), fr.data(), n_frames, s, &launches);
    } else {
        launchProjectBatch(a, maxMp, s, &launches);
    }
    m->launchesTotal += launches;
    int st = 0;
    GVS_CUDA_OK(cudaMemcpyAsync(matched_l, b + oML, (size_t)NL * 4, cudaMemcpyDeviceToHost, s));
    if (two) GVS_CUDA_OK(cudaMemcpyAsync(matched_r, b + oMR, (size_t)NR * 4, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(matches, b + oMt, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(n_matches, b + oCnt, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(&st, b + oSt, sizeof(int), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    GVS_CUDA_OK(cudaGetLastError());
    if (st & kStatProjOverflow) return fail(GVS_ERR_INVALID, "map point scale level out of range");
    if (st & kStatOutputOverflow) return fail(GVS_ERR_CAPACITY, "a frame has more keys than the handle's per-image capacity");
    return GVS_OK;
}

int gvs_match_projection(gvs_matcher* m, int frame, int mode, const gvs_mappoints* mps, float rad, int32_t* matched_l,
                         int32_t* matched_r, int32_t* matches, int* n_matches) {
    if (!mps || !n_matches) return fail(GVS_ERR_INVALID, "null argument");
    const int32_t off[2] = {0, mps->count};
    int32_t n = 0;
    const int rc = gvs_match_projection_batch(m, frame, 1, mode, mps, off, rad, matched_l, matched_r, matches, &n);
    *n_matches = n;
    return rc;
}

int gvs_track_temporal_batch(gvs_matcher* m, int n_frames, const gvs_temporal_params* tp) {
    NvtxScope nvtx("gvs:track_temporal_batch");
    if (!m || !tp || !tp->prev || !tp->pose_prev || !tp->to_camera || !tp->to_camera_r) return fail(GVS_ERR_INVALID, "null argument");
    if (m->R == m->L) return fail(GVS_ERR_STATE, "stereo temporal tracking needs two extractors");
    int rc = checkResident(m, true, 0, n_frames);
    if (rc) return rc;
    if (tp->width < 1 || tp->height < 1) return fail(GVS_ERR_INVALID, "bad image size");
    for (int f = 0; f < n_frames; f++)
        if (tp->prev[f] >= n_frames || tp->prev[f] < -1) return fail(GVS_ERR_INVALID, "prev[] must name a frame of the batch or -1");
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->L->stream;
    gvs_extractor* L = m->L;
    const int stride = L->g.keptPerImage;
    const size_t B = (size_t)std::min(L->maxImages, m->R->maxImages);
    const size_t cap = B * (size_t)stride;
    if (!m->tFlag) {   // first use: per-slot and per-map-point state, sized for the handle's batch capacity
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        bool ok = devAlloc(&m->tFlag, cap) == cudaSuccess && devAlloc(&m->tPredL, cap * 2) == cudaSuccess && devAlloc(&m->tPredR, cap * 2) == cudaSuccess &&
                  devAlloc(&m->tLvlL, cap) == cudaSuccess && devAlloc(&m->tLvlR, cap) == cudaSuccess && devAlloc(&m->tDescRow, cap) == cudaSuccess &&
                  devAlloc(&m->tWPos, cap * 4) == cudaSuccess && devAlloc(&m->tMaxScaleDist, cap) == cudaSuccess &&
                  devAlloc(&m->tMpCnt, B + 1) == cudaSuccess && devAlloc(&m->tMpOff, B + 2) == cudaSuccess && devAlloc(&m->tMpSlot, cap) == cudaSuccess &&
                  devAlloc(&m->tMpKey, cap) == cudaSuccess && devAlloc(&m->tMatchedL, cap) == cudaSuccess &&
                  devAlloc(&m->tMatchedR, B * (size_t)m->R->g.keptPerImage) == cudaSuccess && devAlloc(&m->tMatches, cap * 2) == cudaSuccess &&
                  devAlloc(&m->tRecs, cap * 2 * sizeof(TRec)) == cudaSuccess && devAlloc(&m->tNMatches, B + 1) == cudaSuccess &&
                  devAlloc(&m->tPoses, B * 48 + 16) == cudaSuccess && devAlloc(&m->tPrev, B + 1) == cudaSuccess && devAlloc(&m->tStatus, 4) == cudaSuccess &&
                  cudaMallocHost(reinterpret_cast<void**>(&m->hMpOff), sizeof(int) * (B + 2)) == cudaSuccess &&
                  cudaMallocHost(reinterpret_cast<void**>(&m->hNMatches), sizeof(int) * (B + 1)) == cudaSuccess &&
                  cudaMallocHost(reinterpret_cast<void**>(&m->hTStatus), sizeof(int) * 4) == cudaSuccess;
        if (!ok) return fail(GVS_ERR_CUDA, "temporal state allocation", cudaGetErrorString(cudaGetLastError()));
        *m->hTStatus = 0;
    }
    const size_t F = (size_t)n_frames;
    // The caller's arrays are free again when this returns, whatever memory they live in: they are gathered into a pageable
    // buffer of the handle, and an asynchronous copy FROM PAGEABLE memory is staged by the driver before the call returns.
    m->tStage.resize(F * (48 * sizeof(double)) + F * sizeof(int));
    {
        uint8_t* h = m->tStage.data();
        std::memcpy(h, tp->pose_prev, F * 128);
        std::memcpy(h + F * 128, tp->to_camera, F * 128);
        std::memcpy(h + F * 256, tp->to_camera_r, F * 128);
        std::memcpy(h + F * 384, tp->prev, F * 4);
        GVS_CUDA_OK(cudaMemcpyAsync(m->tPoses, h, F * 128, cudaMemcpyHostToDevice, s));
        GVS_CUDA_OK(cudaMemcpyAsync(m->tPoses + B * 16, h + F * 128, F * 128, cudaMemcpyHostToDevice, s));
        GVS_CUDA_OK(cudaMemcpyAsync(m->tPoses + B * 32, h + F * 256, F * 128, cudaMemcpyHostToDevice, s));
        GVS_CUDA_OK(cudaMemcpyAsync(m->tPrev, h + F * 384, F * 4, cudaMemcpyHostToDevice, s));
    }
    GVS_CUDA_OK(cudaMemsetAsync(m->tMpCnt, 0, F * 4, s));
    GVS_CUDA_OK(cudaMemsetAsync(m->tStatus, 0, 16, s));
    GVS_CUDA_OK(cudaMemsetAsync(m->tMatchedL, 0xff, F * (size_t)stride * 4, s));   // fresh matchedIdxs / matchesIdxs (:1176-1178)
    GVS_CUDA_OK(cudaMemsetAsync(m->tMatchedR, 0xff, F * (size_t)m->R->g.keptPerImage * 4, s));
    GVS_CUDA_OK(cudaMemsetAsync(m->tMatches, 0xff, F * (size_t)stride * 8, s));
    TemporalArgs t{};
    t.nFrames = n_frames; t.slotStride = stride; t.nLevels = L->g.nLevels;
    t.prev = m->tPrev; t.posePrev = m->tPoses; t.toCamera = m->tPoses + B * 16; t.toCameraR = m->tPoses + B * 32;
    t.fx = tp->fx; t.fy = tp->fy; t.cx = tp->cx; t.cy = tp->cy; t.width = tp->width; t.height = tp->height;
    t.logScale = (float)std::log((double)L->params.im_scale);   // unqualified log(float) -> ::log(double) under <cmath>
    for (int i = 0; i < kMaxLevels; i++) t.scalePyramid[i] = i < L->g.nLevels ? L->scale[i] : 1.f;
    t.kpL = L->kps; t.offL = L->imgOff; t.depth = m->depth;
    t.flag = m->tFlag; t.predL = m->tPredL; t.predR = m->tPredR; t.lvlL = m->tLvlL; t.lvlR = m->tLvlR; t.descRow = m->tDescRow;
    t.wPos = m->tWPos; t.maxScaleDist = m->tMaxScaleDist;
    t.mpCnt = m->tMpCnt; t.mpOff = m->tMpOff; t.mpSlot = m->tMpSlot; t.mpKey = m->tMpKey; t.cap = (int)std::min<size_t>(cap, 0x7fffffff);
    t.status = m->tStatus;
    int launches = 0;
    if (L->timing) cudaEventRecord(m->evT[0], s);
    tmp_project_kernel<<<dim3((stride + 127) / 128, n_frames), 128, 0, s>>>(t);
    tmp_scan_kernel<<<1, 1024, 0, s>>>(t);
    tmp_compact_kernel<<<n_frames, 256, 0, s>>>(t);
    launches += 3;
    ProjBatchArgs a{};
    fillFrameArgs(m, a, tp->mode == GVS_MATCH_MONO ? GVS_MATCH_MONO : GVS_MATCH_STEREO_PRED, tp->rad);
    a.nFrames = n_frames; a.frame0 = 0;
    a.mpOff = m->tMpOff; a.mpSlot = m->tMpSlot; a.mpDesc = L->desc; a.mpDescRow = m->tDescRow;
    a.predL = m->tPredL; a.predR = m->tPredR; a.lvlL = m->tLvlL; a.lvlR = m->tLvlR; a.inFrame = nullptr; a.inFrameR = nullptr;
    a.recs = reinterpret_cast<TRec*>(m->tRecs);
    a.matchedL = m->tMatchedL; a.matchedR = m->tMatchedR; a.matches = m->tMatches; a.nMatches = m->tNMatches; a.status = m->tStatus;
    launchProjectBatch(a, stride, s, &launches);
    if (L->timing) cudaEventRecord(m->evT[1], s);
    GVS_CUDA_OK(cudaMemcpyAsync(m->hMpOff, m->tMpOff, (F + 1) * 4, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(m->hNMatches, m->tNMatches, F * 4, cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaMemcpyAsync(m->hTStatus, m->tStatus, sizeof(int), cudaMemcpyDeviceToHost, s));
    GVS_CUDA_OK(cudaGetLastError());
    m->launchesTotal += launches;
    m->launches += launches;
    m->temporalFrames = n_frames;
    m->temporalGen = L->generation;
    return holdRight(m);
}

int gvs_temporal_download(gvs_matcher* m, int first, int count, int32_t* mp_offsets, int32_t* mp_key, int32_t* matches, int32_t* n_matches,
                          int32_t* matched_l, int32_t* matched_r, float* pred_l, float* pred_r) {
    if (!m) return fail(GVS_ERR_INVALID, "null matcher");
    if (m->temporalFrames <= 0 || m->temporalGen != m->L->generation) return fail(GVS_ERR_STATE, "no temporal batch for the current frames");
    if (first < 0 || count < 1 || first + count > m->temporalFrames) return fail(GVS_ERR_INVALID, "frame range out of the last temporal batch");
    int rc = gvs_matcher_sync(m);
    if (rc) return rc;
    if (*m->hTStatus & kStatOutputOverflow) return fail(GVS_ERR_CAPACITY, "temporal map point lists overflow");
    DeviceGuard deviceGuard(m->L->device);
    cudaStream_t s = m->L->stream;
    const int o0 = m->hMpOff[first], M = m->hMpOff[first + count] - o0;
    if (mp_offsets) for (int f = 0; f <= count; f++) mp_offsets[f] = m->hMpOff[first + f] - o0;
    if (n_matches) std::memcpy(n_matches, m->hNMatches + first, sizeof(int) * (size_t)count);
    if (M > 0) {
        if (mp_key) GVS_CUDA_OK(cudaMemcpyAsync(mp_key, m->tMpKey + o0, (size_t)M * 4, cudaMemcpyDeviceToHost, s));
        if (matches) GVS_CUDA_OK(cudaMemcpyAsync(matches, m->tMatches + 2 * (size_t)o0, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
    }
    const int l0 = m->L->hOff[first], nl = m->L->hOff[first + count] - l0;
    const int r0 = m->R->hOff[first], nr = m->R->hOff[first + count] - r0;
    if (matched_l && nl > 0) GVS_CUDA_OK(cudaMemcpyAsync(matched_l, m->tMatchedL + l0, (size_t)nl * 4, cudaMemcpyDeviceToHost, s));
    if (matched_r && nr > 0) GVS_CUDA_OK(cudaMemcpyAsync(matched_r, m->tMatchedR + r0, (size_t)nr * 4, cudaMemcpyDeviceToHost, s));
    if ((pred_l || pred_r) && M > 0) {
        // predictions live per slot (frame * stride + source key): fetch the frames' slot ranges and gather on the host
        // (parity tap, not a throughput path)
        const size_t stride = (size_t)m->L->g.keptPerImage, s0 = (size_t)first * stride, ns = (size_t)count * stride;
        std::vector<int> slot((size_t)M);
        std::vector<float> pl(pred_l ? ns * 2 : 0), pr(pred_r ? ns * 2 : 0);
        GVS_CUDA_OK(cudaMemcpyAsync(slot.data(), m->tMpSlot + o0, (size_t)M * 4, cudaMemcpyDeviceToHost, s));
        if (pred_l) GVS_CUDA_OK(cudaMemcpyAsync(pl.data(), m->tPredL + 2 * s0, ns * 8, cudaMemcpyDeviceToHost, s));
        if (pred_r) GVS_CUDA_OK(cudaMemcpyAsync(pr.data(), m->tPredR + 2 * s0, ns * 8, cudaMemcpyDeviceToHost, s));
        GVS_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < M; i++) {
            const size_t j = (size_t)slot[i] - s0;
            if (pred_l) { pred_l[2 * (size_t)i] = pl[2 * j]; pred_l[2 * (size_t)i + 1] = pl[2 * j + 1]; }
            if (pred_r) { pred_r[2 * (size_t)i] = pr[2 * j]; pred_r[2 * (size_t)i + 1] = pr[2 * j + 1]; }
        }
    }
    GVS_CUDA_OK(cudaStreamSynchronize(s));
    return GVS_OK;
}

}  // extern "C"
