// Map-point side kernels of the matcher (everything here works on caller-provided arrays of ONE call):
//   * FeatureMatcher::matchByProjectionRPredLBA (reference src/FeatureMatcher.cpp:66-252): proj_candidates_kernel (ordered
//     candidate lists of a stored key frame, reduced to the three deciding candidates of SURVEY App. A.9) + lba_resolve_kernel;
//   * MapPoint::calcDescriptor (src/Map.cpp:145-210), FeatureTracker::worldToFrame / removeOutOfFrameMPs / PredictMPsPosition /
//     findOutliersR / findOutliersMono (src/FeatureTracker.cpp:582-1013), FeatureMatcher::DescriptorDistance (:710-726).
// The batched projection matchers against the device-resident frames (matchByProjectionRPred / Mono / matchByRadius) live in
// temporal.cu.
#include <algorithm>

#include "common.cuh"
#include "mappoint_math.cuh"

namespace gvs {

// list entry: idx (16 bits) | dist (9 bits) << 16 | level (4 bits) << 25 ; 0xffffffff = none
__device__ __forceinline__ unsigned packCand(int idx, int dist, int lev) { return (unsigned)idx | ((unsigned)dist << 16) | ((unsigned)lev << 25); }
__device__ __forceinline__ int candIdx(unsigned e) { return (int)(e & 0xffff); }
__device__ __forceinline__ int candDist(unsigned e) { return (int)((e >> 16) & 0x1ff); }
__device__ __forceinline__ int candLev(unsigned e) { return (int)((e >> 25) & 0xf); }
constexpr unsigned kNone = 0xffffffffu;

constexpr int kModeLBA = 3;  // internal: FeatureMatcher::matchByProjectionRPredLBA

struct ProjRec {  // per (map point, side)
    int start, len;
    unsigned eBest, ePre, eSuf;
};

__device__ __forceinline__ unsigned warpMin(unsigned v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce an ordered candidate list (optionally skipping matched keys) to (eBest, ePre, eSuf).  Warp-wide.
__device__ void reduceList(const unsigned* list, int len, const int* matched, int lane, unsigned& eBest, unsigned& ePre,
                           unsigned& eSuf) {
    // key = dist << 20 | position  (len < 2^20)
    unsigned kb = kNone;
    for (int p = lane; p < len; p += 32) {
        const unsigned e = list[p];
        if (matched && matched[candIdx(e)] >= 0) continue;
        kb = min(kb, ((unsigned)candDist(e) << 20) | (unsigned)p);
    }
    kb = warpMin(kb);
    if (kb == kNone) { eBest = ePre = eSuf = kNone; return; }
    const int pBest = (int)(kb & 0xfffff);
    unsigned kp = kNone, ks = kNone;
    for (int p = lane; p < len; p += 32) {
        if (p == pBest) continue;
        const unsigned e = list[p];
        if (matched && matched[candIdx(e)] >= 0) continue;
        const unsigned k = ((unsigned)candDist(e) << 20) | (unsigned)p;
        if (p < pBest) kp = min(kp, k); else ks = min(ks, k);
    }
    kp = warpMin(kp);
    ks = warpMin(ks);
    eBest = list[pBest];
    ePre = kp == kNone ? kNone : list[kp & 0xfffff];
    eSuf = ks == kNone ? kNone : list[ks & 0xfffff];
}

__device__ __forceinline__ void projCandidatesBody(const ProjArgs& a, ProjRec* __restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const int nSides = (a.mode == GVS_MATCH_STEREO_PRED || a.mode == kModeLBA) ? 2 : 1;
    const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < a.M * nSides; w += warpsPerGrid) {
        const int mp = w / nSides, side = w - mp * nSides;
        ProjRec rec;
        rec.start = 0; rec.len = 0; rec.eBest = rec.ePre = rec.eSuf = kNone;
        const float* pred = side ? a.predR : a.predL;
        const int lvl = side ? a.lvlR[mp] : a.lvlL[mp];
        const bool inFrame = a.mode == GVS_MATCH_RADIUS ? true : (side ? (a.inFrameR ? a.inFrameR[mp] != 0 : true)
                                                                       : (a.inFrame ? a.inFrame[mp] != 0 : true));
        const gvs_keypoint* kps = side ? a.kpR : a.kpL;
        const uint8_t* descs = side ? a.descR : a.descL;
        const int* cellStart = side ? a.cellStartR : a.cellStartL;
        const int* items = side ? a.itemsR : a.itemsL;
        // LBA: keys already owned by a map point are excluded up front (newKF->unMatchedF/FR, src/FeatureMatcher.cpp:118,167);
        // the arrays are not modified by the call, so map points are independent
        const int* excl = a.mode == kModeLBA ? (side ? a.matchedR : a.matchedL) : nullptr;
        const float trackX = pred[2 * mp], trackY = pred[2 * mp + 1];
        const float radius = __fmul_rn(a.scalePyramid[min(max(lvl, 0), kMaxLevels - 1)], a.rad);
        // getMatchIdxs window (src/FeatureMatcher.cpp:18-30)
        const int minX = max(0, (int)floorf(__fmul_rn(__fsub_rn(trackX, radius), a.xMult)));
        const int maxX = min(a.xGrids - 1, (int)ceilf(__fmul_rn(__fadd_rn(trackX, radius), a.xMult)));
        const int minY = max(0, (int)floorf(__fmul_rn(__fsub_rn(trackY, radius), a.yMult)));
        const int maxY = min(a.yGrids - 1, (int)ceilf(__fmul_rn(__fadd_rn(trackY, radius), a.yMult)));
        const bool window = inFrame && !(minX >= a.xGrids || minY >= a.yGrids || maxX < 0 || maxY < 0) && minX <= maxX;
        if (window) {
            const uint4* dq = reinterpret_cast<const uint4*>(a.mpDesc + (size_t)mp * 32);
            const uint4 q0 = __ldg(dq), q1 = __ldg(dq + 1);
            auto passes = [&](int idx) {
                const gvs_keypoint k = kps[idx];
                if (k.octave > lvl + 1 || k.octave < lvl - 1) return -1;
                if (excl && excl[idx] >= 0) return -1;
                const float dx = __fsub_rn(k.x, trackX), dy = __fsub_rn(k.y, trackY);
                if (!(fabsf(dx) < radius && fabsf(dy) < radius)) return -1;
                if (a.mode == GVS_MATCH_RADIUS) {  // Converter::checkPixelParallax (include/Conversions.h:140-144)
                    const double ex = (double)trackX - (double)k.x, ey = (double)trackY - (double)k.y;
                    if (!(sqrt(ex * ex + ey * ey) > 10.0)) return -1;
                }
                return k.octave;
            };
            // pass 1: count
            int cnt = 0;
            for (int row = minY; row <= maxY; row++) {
                const int p0 = cellStart[row * a.xGrids + minX], p1 = cellStart[row * a.xGrids + maxX + 1];
                for (int p = p0 + lane; p < p1; p += 32) cnt += passes(items[p]) >= 0;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            int start = 0;
            if (lane == 0 && cnt > 0) start = atomicAdd(a.poolTop + side, cnt);
            start = __shfl_sync(0xffffffffu, start, 0);
            if (cnt > 0 && start + cnt > a.candCap) {
                if (lane == 0) atomicOr(a.status, kStatProjOverflow);
                cnt = 0;
            }
            // pass 2: distances, ordered write
            if (cnt > 0) {
                unsigned* list = (side ? a.candR : a.candL) + start;
                int run = 0;
                for (int row = minY; row <= maxY; row++) {
                    const int p0 = cellStart[row * a.xGrids + minX], p1 = cellStart[row * a.xGrids + maxX + 1];
                    for (int base = p0; base < p1; base += 32) {
                        const int p = base + lane;
                        int idx = -1, lev = -1;
                        if (p < p1) { idx = items[p]; lev = passes(idx); }
                        const unsigned m = __ballot_sync(0xffffffffu, lev >= 0);
                        if (lev >= 0) {
                            const uint4* dk = reinterpret_cast<const uint4*>(descs + (size_t)idx * 32);
                            const uint4 k0 = __ldg(dk), k1 = __ldg(dk + 1);
                            const int dist = __popc(q0.x ^ k0.x) + __popc(q0.y ^ k0.y) + __popc(q0.z ^ k0.z) + __popc(q0.w ^ k0.w) +
                                             __popc(q1.x ^ k1.x) + __popc(q1.y ^ k1.y) + __popc(q1.z ^ k1.z) + __popc(q1.w ^ k1.w);
                            list[run + __popc(m & lt)] = packCand(idx, dist, lev);
                        }
                        run += __popc(m);
                    }
                }
                __syncwarp();
                rec.start = start; rec.len = cnt;
                reduceList(list, cnt, nullptr, lane, rec.eBest, rec.ePre, rec.eSuf);
            }
        }
        if (lane == 0) recs[w] = rec;
    }
}

__global__ void __launch_bounds__(256)
proj_candidates_kernel(const ProjArgs a, ProjRec* __restrict__ recs) { projCandidatesBody(a, recs); }
// one grid row per frame, the frames' arguments in device memory
__global__ void __launch_bounds__(256)
proj_candidates_frames_kernel(const ProjArgs* __restrict__ frames) {
    const ProjArgs& a = frames[blockIdx.y];
    if ((int)blockIdx.x * 8 >= 2 * a.M) return;
    projCandidatesBody(a, reinterpret_cast<ProjRec*>(a.recs));
}

struct Decision { int bestDist, bestIdx, bestLev, bestLev2, secDist; };
__device__ __forceinline__ Decision decide(unsigned eBest, unsigned ePre, unsigned eSuf) {
    Decision d{256, -1, -1, -1, 256};
    if (eBest == kNone) return d;
    d.bestDist = candDist(eBest); d.bestIdx = candIdx(eBest); d.bestLev = candLev(eBest);
    if (ePre != kNone) { d.secDist = candDist(ePre); d.bestLev2 = candLev(ePre); }
    if (eSuf != kNone && candDist(eSuf) < d.secDist) { d.secDist = candDist(eSuf); d.bestLev2 = candLev(eSuf); }
    return d;
}

// Wide windows (gvs_match_projection_batch with a search radius that covers a large part of the image: hundreds of candidates per
// map point): the candidate lists are stored (proj_candidates_kernel above) and phase 2 works on the stored lists.
// phase 2.  The reference walks the map points in index order; map point i only depends on an earlier one j if j takes a key
// that is one of i's (at most six) deciding candidates.  A CTA of 32 warps handles 32 map points at a time: every warp
// computes the tentative decision of its map point against the current matched arrays (rescanning its stored list when a
// deciding candidate is already taken), then warp 0 finds the longest prefix of the chunk in which no map point is affected by
// an earlier one of the same chunk, commits it, and the rest is re-evaluated.  The first pending map point is always safe,
// so every round makes progress; without conflicts a chunk costs one round.
constexpr int kResolveWarps = 32;

__global__ void __launch_bounds__(32 * kResolveWarps)
proj_resolve_frames_kernel(const ProjArgs* __restrict__ frames) {
    const ProjArgs& a = frames[blockIdx.x];
    const ProjRec* __restrict__ recs = reinterpret_cast<const ProjRec*>(a.recs);
    extern __shared__ int sMatched[];  // NL + NR
    __shared__ int sDec[kResolveWarps][6];                     // key ids of the deciding candidates (left 0..2, right 3..5), -1 = none
    __shared__ int sTakeL[kResolveWarps], sTakeR[kResolveWarps];   // keys the map point marks if accepted (-1 = none)
    __shared__ int sOut0[kResolveWarps], sOut1[kResolveWarps];     // its matchesIdxs pair
    __shared__ int sAccept[kResolveWarps];
    __shared__ int sDone, sTotal;
    int* mL = sMatched;
    int* mR = sMatched + a.NL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool stereo = a.mode == GVS_MATCH_STEREO_PRED;
    for (int i = tid; i < a.NL; i += blockDim.x) mL[i] = a.matchedL[i];
    if (stereo) for (int i = tid; i < a.NR; i += blockDim.x) mR[i] = a.matchedR[i];
    if (tid == 0) sTotal = 0;
    const int nSides = stereo ? 2 : 1;
    auto taken = [&](unsigned e, const int* m) { return e != kNone && m[candIdx(e)] >= 0; };
    auto keyOf = [&](unsigned e) { return e == kNone ? -1 : candIdx(e); };
    for (int base = 0; base < a.M; base += kResolveWarps) {
        const int chunk = min(kResolveWarps, a.M - base);
        __syncthreads();
        if (tid == 0) sDone = 0;
        bool first = true;
        while (true) {
            __syncthreads();
            const int done = sDone;
            if (done >= chunk) break;
            // ---- (A) tentative decisions of the pending map points whose inputs changed
            if (warp >= done && warp < chunk) {
                bool redo = first;
                if (!first) {
#pragma unroll
                    for (int q = 0; q < 6; q++) {
                        const int kq = sDec[warp][q];
                        if (kq >= 0 && (q < 3 ? mL[kq] : mR[kq]) >= 0) redo = true;
                    }
                }
                if (redo) {
                    const int mp = base + warp;
                    const bool skipMP = a.mode != GVS_MATCH_RADIUS && (a.matches[2 * mp] >= 0 || a.matches[2 * mp + 1] >= 0);
                    int accept = 0, takeL = -1, takeR = -1, out0 = -1, out1 = -1;
                    unsigned eb = kNone, ep = kNone, es = kNone, rb = kNone, rp = kNone, rs = kNone;
                    if (!skipMP) {
                        const ProjRec rl = recs[mp * nSides];
                        eb = rl.eBest; ep = rl.ePre; es = rl.eSuf;
                        if (taken(eb, mL) || taken(ep, mL) || taken(es, mL)) reduceList(a.candL + rl.start, rl.len, mL, lane, eb, ep, es);
                        Decision d = decide(eb, ep, es);
                        bool right = false;
                        int bestIdxR = -1;
                        if (stereo) {
                            const ProjRec rr = recs[mp * nSides + 1];
                            rb = rr.eBest; rp = rr.ePre; rs = rr.eSuf;
                            if (taken(rb, mR) || taken(rp, mR) || taken(rs, mR)) reduceList(a.candR + rr.start, rr.len, mR, lane, rb, rp, rs);
                            const Decision dr = decide(rb, rp, rs);
                            bestIdxR = dr.bestIdx;
                            if (d.bestDist > dr.bestDist) {
                                d.bestDist = dr.bestDist; d.secDist = dr.secDist; d.bestLev = dr.bestLev; d.bestLev2 = dr.bestLev2;
                                right = true;
                            }
                        }
                        bool ok;
                        if (a.mode == GVS_MATCH_MONO) {
                            // (ratioProj + 0.1) * secDist is evaluated in double (float + double), threshold 100 + 50
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
                                const int li = a.leftIdxs[bestIdxR];
                                if (li >= 0) { out0 = li; takeL = li; }
                            } else {
                                takeL = d.bestIdx; out0 = d.bestIdx;
                                const int ri = a.rightIdxs[d.bestIdx];
                                if (ri >= 0) { out1 = ri; takeR = ri; }
                            }
                        }
                    }
                    if (lane == 0) {
                        sDec[warp][0] = keyOf(eb); sDec[warp][1] = keyOf(ep); sDec[warp][2] = keyOf(es);
                        sDec[warp][3] = keyOf(rb); sDec[warp][4] = keyOf(rp); sDec[warp][5] = keyOf(rs);
                        sAccept[warp] = accept; sTakeL[warp] = takeL; sTakeR[warp] = takeR; sOut0[warp] = out0; sOut1[warp] = out1;
                    }
                }
            }
            first = false;
            __syncthreads();
            // ---- (B) warp 0, lane = map point of the chunk: longest safe prefix, commit
            if (warp == 0) {
                bool unsafe = false;
                int dl0 = -1, dl1 = -1, dl2 = -1, dr0 = -1, dr1 = -1, dr2 = -1;
                if (lane < chunk) { dl0 = sDec[lane][0]; dl1 = sDec[lane][1]; dl2 = sDec[lane][2]; dr0 = sDec[lane][3]; dr1 = sDec[lane][4]; dr2 = sDec[lane][5]; }
                for (int j = done; j < chunk - 1; j++) {
                    if (!sAccept[j]) continue;   // (uniform)
                    const int tl = sTakeL[j], tr = sTakeR[j];
                    if (lane > j && lane < chunk) {
                        if (tl >= 0 && (tl == dl0 || tl == dl1 || tl == dl2)) unsafe = true;
                        if (tr >= 0 && (tr == dr0 || tr == dr1 || tr == dr2)) unsafe = true;
                    }
                }
                const unsigned um = __ballot_sync(0xffffffffu, unsafe);
                const int upto = um ? __ffs(um) - 1 : chunk;     // lanes [done, upto) are final
                const bool mine = lane >= done && lane < upto && sAccept[lane];
                const int i = base + lane;
                // several committed map points may mark the same key (a stereo partner): the last one in index order wins
                const int tl = mine ? sTakeL[lane] : -1, tr = mine ? sTakeR[lane] : -1;
                const unsigned pl = __match_any_sync(0xffffffffu, tl >= 0 ? tl : -1 - lane);
                const unsigned pr = __match_any_sync(0xffffffffu, tr >= 0 ? tr : -1 - lane);
                if (tl >= 0 && (31 - __clz(pl)) == lane) mL[tl] = i;
                if (tr >= 0 && (31 - __clz(pr)) == lane) mR[tr] = i;
                if (mine) {
                    if (a.mode != GVS_MATCH_STEREO_PRED) a.matches[2 * i] = sOut0[lane];
                    else {
                        if (sOut0[lane] >= 0) a.matches[2 * i] = sOut0[lane];
                        if (sOut1[lane] >= 0) a.matches[2 * i + 1] = sOut1[lane];
                    }
                }
                const int nAcc = __popc(__ballot_sync(0xffffffffu, mine));
                if (lane == 0) { sTotal += nAcc; sDone = upto; }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < a.NL; i += blockDim.x) a.matchedL[i] = mL[i];
    if (stereo) for (int i = tid; i < a.NR; i += blockDim.x) a.matchedR[i] = mR[i];
    if (tid == 0) *a.nMatches = sTotal;
}

// frames: nFrames argument blocks in device memory (host copy in hostFrames for the launch geometry); every frame has its own
// pool counters (poolTop), pool regions (candL / candR / candCap), recs and nMatches
void launchProjectWide(const ProjArgs* frames, const ProjArgs* hostFrames, int nFrames, cudaStream_t s, int* launches) {
    int maxM = 0, maxKeys = 0;
    for (int f = 0; f < nFrames; f++) { maxM = std::max(maxM, hostFrames[f].M); maxKeys = std::max(maxKeys, hostFrames[f].NL + hostFrames[f].NR); }
    if (maxM <= 0) return;
    proj_candidates_frames_kernel<<<dim3((2 * maxM + 7) / 8, nFrames), 256, 0, s>>>(frames);
    ++*launches;
    const size_t sm = (size_t)maxKeys * sizeof(int);
    static size_t cfg[64];
    ensureDynamicSmem(proj_resolve_frames_kernel, sm + 4096, cfg);   // (+ the kernel's static shared memory counts against the 48 KB default)
    proj_resolve_frames_kernel<<<nFrames, 32 * kResolveWarps, sm, s>>>(frames);
    ++*launches;
}

// FeatureMatcher::matchByProjectionRPredLBA decision (src/FeatureMatcher.cpp:191-249), one thread per candidate point.
// Left branch quirk (:132-136): a later second-best takes the *best* key's level.
__global__ void lba_resolve_kernel(const ProjArgs a, const ProjRec* __restrict__ recs, const float* __restrict__ lastL,
                                   const float* __restrict__ lastR, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.M;
    bool accept = false;
    int o0 = -2, o1 = -2;
    if (valid) {
        const ProjRec rl = recs[2 * i], rr = recs[2 * i + 1];
        Decision d{256, -1, -1, -1, 256};
        if (rl.eBest != kNone) {
            d.bestDist = candDist(rl.eBest); d.bestIdx = candIdx(rl.eBest); d.bestLev = candLev(rl.eBest);
            if (rl.ePre != kNone) { d.secDist = candDist(rl.ePre); d.bestLev2 = candLev(rl.ePre); }
            if (rl.eSuf != kNone && candDist(rl.eSuf) < d.secDist) { d.secDist = candDist(rl.eSuf); d.bestLev2 = d.bestLev; }
        }
        const Decision dr = decide(rr.eBest, rr.ePre, rr.eSuf);
        bool right = false;
        if (d.bestDist > dr.bestDist) { d.bestDist = dr.bestDist; d.secDist = dr.secDist; d.bestLev = dr.bestLev; d.bestLev2 = dr.bestLev2; right = true; }
        const float thr = __fmul_rn(0.6f, (float)d.secDist);
        bool ok = d.bestDist <= 50 && !(d.bestLev == d.bestLev2 && (float)d.bestDist >= thr) &&
                  (d.bestLev != d.bestLev2 || (float)d.bestDist < thr);
        if (ok) {
            const float* last = right ? lastR + 2 * i : lastL + 2 * i;
            const float* pred = right ? a.predR + 2 * i : a.predL + 2 * i;
            const double dx = (double)pred[0] - (double)last[0], dy = (double)pred[1] - (double)last[1];
            ok = sqrt(dx * dx + dy * dy) > 10.0;  // Converter::checkPixelParallax (include/Conversions.h:140-144)
        }
        if (ok) {
            accept = true;
            if (right) { const int li = a.leftIdxs[dr.bestIdx]; o0 = li >= 0 ? li : -1; o1 = dr.bestIdx; }
            else { const int ri = a.rightIdxs[d.bestIdx]; o0 = d.bestIdx; o1 = ri >= 0 ? ri : -1; }
        }
        out[2 * i] = o0; out[2 * i + 1] = o1;
    }
    const int n = __syncthreads_count(accept);
    if (threadIdx.x == 0 && n) atomicAdd(a.nMatches, n);
}

void launchProjectLBA(const ProjArgs& a, const float* lastL, const float* lastR, int* out, cudaStream_t s, int* launches) {
    ProjRec* recs = reinterpret_cast<ProjRec*>(a.recs);
    cudaMemsetAsync(a.poolTop, 0, 2 * sizeof(int), s);
    cudaMemsetAsync(a.nMatches, 0, sizeof(int), s);
    if (a.M <= 0) return;
    proj_candidates_kernel<<<(2 * a.M + 7) / 8, 256, 0, s>>>(a, recs);
    ++*launches;
    lba_resolve_kernel<<<(a.M + 127) / 128, 128, 0, s>>>(a, recs, lastL, lastR, out);
    ++*launches;
}

// MapPoint::calcDescriptor (src/Map.cpp:145-210): one CTA per map point, one warp per observation row; the median of the
// row's distances (element int(0.5*(N-1)) of the sorted row, self distance 0 included) comes from a 257-bin histogram.
__global__ void __launch_bounds__(128)
median_descriptor_kernel(const uint8_t* __restrict__ descs, const int* __restrict__ offsets, int* __restrict__ bestIdx) {
    __shared__ int hist[4][264];
    __shared__ unsigned best[4];
    const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = offsets[p], N = offsets[p + 1] - o;
    if (N <= 0) { if (threadIdx.x == 0) bestIdx[p] = -1; return; }
    const int k = (int)(0.5 * (N - 1));
    unsigned mine = 0xffffffffu;
    for (int i = warp; i < N; i += 4) {
        for (int b = lane; b < 264; b += 32) hist[warp][b] = 0;
        __syncwarp();
        const uint4* di = reinterpret_cast<const uint4*>(descs + (size_t)(o + i) * 32);
        const uint4 a0 = __ldg(di), a1 = __ldg(di + 1);
        for (int j = lane; j < N; j += 32) {
            const uint4* dj = reinterpret_cast<const uint4*>(descs + (size_t)(o + j) * 32);
            const uint4 b0 = __ldg(dj), b1 = __ldg(dj + 1);
            const int d = i == j ? 0 : __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                       __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            atomicAdd(&hist[warp][d], 1);
        }
        __syncwarp();
        // k-th smallest: first bin whose inclusive prefix count exceeds k (lane owns 9 consecutive bins)
        int c[9], tot = 0;
#pragma unroll
        for (int t = 0; t < 9; t++) { const int b = lane * 9 + t; c[t] = b < 257 ? hist[warp][b] : 0; tot += c[t]; }
        int incl = tot;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, s2); if (lane >= s2) incl += v; }
        int run = incl - tot, med = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < 9; t++) { run += c[t]; if (run > k && med == 0x7fffffff) med = lane * 9 + t; }
#pragma unroll
        for (int s2 = 16; s2; s2 >>= 1) med = min(med, __shfl_xor_sync(0xffffffffu, med, s2));
        mine = min(mine, ((unsigned)med << 16) | (unsigned)i);  // least median, first index on ties (strict < in the reference)
        __syncwarp();
    }
    if (lane == 0) best[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) bestIdx[p] = (int)(min(min(best[0], best[1]), min(best[2], best[3])) & 0xffff);
}

void launchMedianDescriptor(const uint8_t* descs, const int* offsets, int P, int* bestIdx, cudaStream_t s, int* launches) {
    if (P <= 0) return;
    median_descriptor_kernel<<<P, 128, 0, s>>>(descs, offsets, bestIdx);
    ++*launches;
}

__global__ void world_to_frame_kernel(int M, const double* __restrict__ poseInv, double fx, double fy, double cx, double cy, int width,
                                      int height, const double* __restrict__ wPos, const float* __restrict__ maxScaleDist,
                                      const float* __restrict__ logScale, const int* __restrict__ nScaleLev, float* __restrict__ pred,
                                      int* __restrict__ scaleLevel, uint8_t* __restrict__ inFrame) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    inFrame[i] = worldToFrameOne(poseInv, wPos + 4 * (size_t)i, fx, fy, cx, cy, width, height, maxScaleDist[i], logScale[i], nScaleLev[i],
                                 pred + 2 * (size_t)i, scaleLevel + i) ? 1 : 0;
}

// FeatureTracker::removeOutOfFrameMPs (src/FeatureTracker.cpp:910-940; mode 0), removeOutOfFrameMPsMono (:942-967; mode 1) and
// PredictMPsPosition (:969-1013; mode 2), thread per active map point.  Every iteration of the reference's loops only touches its
// own map point and, in mode 2, the matched-index slots of its own match (distinct matches own distinct slots; equal ones would
// all write -1), so the loops are order independent.
__global__ void predict_mappoints_kernel(const PredictArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    if (a.mode != 2) a.keep[i] = 0;
    if (a.valid && !a.valid[i]) return;        // if ( !mp ) continue;  (modes 0/1 also: if ( mp->GetIsOutlier() ) continue;)
    const double* w = a.wPos + 4 * (size_t)i;
    const bool c1 = worldToFrameOne(a.toCamera, w, a.fx, a.fy, a.cx, a.cy, a.width, a.height, a.maxScaleDist[i], a.logScale[i],
                                    a.nScaleLev[i], a.predL + 2 * (size_t)i, a.scaleLevelL + i);
    a.inFrame[i] = c1 ? 1 : 0;
    bool c2 = true;
    if (a.mode != 1) {
        c2 = worldToFrameOne(a.toCameraR, w, a.fx, a.fy, a.cx, a.cy, a.width, a.height, a.maxScaleDist[i], a.logScale[i], a.nScaleLev[i],
                             a.predR + 2 * (size_t)i, a.scaleLevelR + i);
        a.inFrameR[i] = c2 ? 1 : 0;
    }
    if (a.mode != 2) { a.keep[i] = (c1 && c2) ? 1 : 0; return; }
    const bool out = a.outliers[i] != 0;
    if (out) a.outliers[i] = 0;
    const int first = a.matches[2 * i], second = a.matches[2 * i + 1];
    if ((!c1 || out) && first >= 0) { a.matchedL[first] = -1; a.matches[2 * i] = -1; }
    if ((!c2 || out) && second >= 0) { a.matchedR[second] = -1; a.matches[2 * i + 1] = -1; }
}

void launchPredictMapPoints(const PredictArgs& a, cudaStream_t s, int* launches) {
    if (a.M <= 0) return;
    predict_mappoints_kernel<<<(a.M + 127) / 128, 128, 0, s>>>(a);
    ++*launches;
}

void launchWorldToFrame(int M, const double* poseInv, double fx, double fy, double cx, double cy, int width, int height,
                        const double* wPos, const float* maxScaleDist, const float* logScale, const int* nScaleLev, float* pred,
                        int* scaleLevel, uint8_t* inFrame, cudaStream_t s, int* launches) {
    if (M <= 0) return;
    world_to_frame_kernel<<<(M + 127) / 128, 128, 0, s>>>(M, poseInv, fx, fy, cx, cy, width, height, wPos, maxScaleDist, logScale,
                                                          nScaleLev, pred, scaleLevel, inFrame);
    ++*launches;
}

// FeatureTracker::findOutliersR / findOutliersMono (src/FeatureTracker.cpp:582-649, :651-684) with check2dError (:147-164).
// The reference walks the matches in index order; everything a match computes is independent of the others except the
// stereo-partner check, which reads keysLeft.close[nIdx] and clears it (with the depth and the stereo pair) when the right
// observation is an outlier.  If several matches point at the same left key, only those up to and including the first one that
// clears see close == true.  Hence two passes: pass 1 (thread per match) evaluates the reprojection tests against the
// original `close` and records, per left key, the smallest match index that would clear it (atomicMin); pass 2 counts a
// stereo inlier only before that index and applies the clearing at that index.  Double precision, explicit round-to-nearest
// operations in the reference's left-to-right order (no FMA), like world_to_frame_kernel.
// FeatureTracker::check2dError, :147-164
__device__ __forceinline__ bool check2dError(const double (&p)[3], float ox, float oy, double fx, double fy, double cx, double cy,
                                             double thres, double weight) {
    if (p[2] <= 0) return true;
    const double invZ = __ddiv_rn(1.0, p[2]);
    const double u = __dadd_rn(__dmul_rn(__dmul_rn(fx, p[0]), invZ), cx), v = __dadd_rn(__dmul_rn(__dmul_rn(fy, p[1]), invZ), cy);
    const double eU = __dsub_rn((double)ox, u), eV = __dsub_rn((double)oy, v);
    const double error = __dmul_rn(__dadd_rn(__dmul_rn(eU, eU), __dmul_rn(eV, eV)), weight);
    return error > thres;
}

__global__ void find_outliers_kernel(const OutlierArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    a.act[i] = 0;
    if (a.valid && !a.valid[i]) return;                     // if ( !mp ) continue;
    const int first = a.matches[2 * i], second = a.matches[2 * i + 1];
    const double* w = a.wPos + 4 * (size_t)i;
    double p[3];
    int nIdx;
    bool right = false;
    if (first >= 0) {
        if (a.inFrame && !a.inFrame[i]) return;
        mulPose(a.pose, w, p);
        nIdx = first;
    } else if (a.mode == 0 && second >= 0) {
        if (a.inFrameR && !a.inFrameR[i]) return;
        right = true;
        mulPose(a.poseR, w, p);
        nIdx = second;
    } else return;
    const gvs_keypoint k = right ? a.kpR[nIdx] : a.kpL[nIdx];
    const bool outlier = check2dError(p, k.x, k.y, a.fx, a.fy, a.cx, a.cy, a.thres, (double)a.invSigma[k.octave]);
    a.outliers[i] = outlier ? 1 : 0;
    if (outlier) return;
    atomicAdd(&a.counters[0], 1);                           // nInliers++
    if (a.mode != 0 || right || !(p[2] < (double)a.closeDist) || !a.close[nIdx] || second < 0) return;
    double pr[3];
    mulPose(a.poseR, w, pr);
    const gvs_keypoint kr = a.kpR[second];
    if (!check2dError(pr, kr.x, kr.y, a.fx, a.fy, a.cx, a.cy, a.thres, (double)a.invSigma[kr.octave])) a.act[i] = 1;
    else { a.act[i] = 2; atomicMin(&a.firstClear[nIdx], i); }
}

__global__ void find_outliers_apply_kernel(const OutlierArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    const int act = a.act[i];
    if (act == 0) return;
    const int nIdx = a.matches[2 * i], fc = a.firstClear[nIdx];
    if (act == 1) { if (i < fc) atomicAdd(&a.counters[1], 1); return; }   // nStereo++ while close[nIdx] is still set
    if (i != fc) return;
    a.depth[nIdx] = -1.f;
    a.close[nIdx] = 0;
    const int rIdx = a.rightIdxs[nIdx];
    a.rightIdxs[nIdx] = -1;
    if (rIdx >= 0 && rIdx < a.NR) a.leftIdxs[rIdx] = -1;   // close[nIdx] implies a stereo partner; the reference indexes unchecked
    a.matches[2 * i + 1] = -1;
}

void launchFindOutliers(const OutlierArgs& a, cudaStream_t s, int* launches) {
    if (a.M <= 0) return;
    find_outliers_kernel<<<(a.M + 127) / 128, 128, 0, s>>>(a);
    ++*launches;
    if (a.mode == 0) {
        find_outliers_apply_kernel<<<(a.M + 127) / 128, 128, 0, s>>>(a);
        ++*launches;
    }
}

__global__ void hamming_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* pa = reinterpret_cast<const uint4*>(a + (size_t)i * 32);
    const uint4* pb = reinterpret_cast<const uint4*>(b + (size_t)i * 32);
    const uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    out[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
             __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

void launchHamming(const uint8_t* a, const uint8_t* b, int n, int* out, cudaStream_t s, int* launches) {
    if (n <= 0) return;
    hamming_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, n, out);
    ++*launches;
}

}  // namespace gvs
