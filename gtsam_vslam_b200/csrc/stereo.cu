// K8 stereo_match + K8b stereo_finalize + K9 grid_assign.
// Replaces FeatureMatcher::findStereoMatchesORB2R and destributeRightKeys (reference src/FeatureMatcher.cpp:528-708,
// :728-752) and FeatureTracker::assignKeysToGrids (src/FeatureTracker.cpp:28-54) for a batch of frames.
//
// stereo_match: one warp per left key.
//   * the row buckets of destributeRightKeys are built per frame by row_bucket_kernel (CSR over image rows): right key j is in
//     bucket[yKeyL] iff floor(yKeyR - r) <= yKeyL <= ceil(yKeyR + r), r = 2*scalePyramid[octR] (float32); the reference's buckets
//     hold indices in ascending order, so its "first minimum in bucket order" is the minimum of (distance, j), whatever the order
//     here.  Right keys are stored level by level, so the level test octL-1..octL+1 is a range test on j.
//   * the reference's "u" test uses pt.y (sic): yL - fx <= yR <= yL.
//   * Hamming distance: 8 x __popc on 32-bit words, warp-shuffle argmin on (dist << 20 | j).
//   * SAD refinement on the unblurred level octL of both pyramids: 22 lanes = 11 shifts x 2 row halves, a patch row = 3 words
//     compared four bytes at a time (VABSDIFF4 with accumulate);
//     skipped shifts keep 0.0 (App. B.3); float32 parabola and depth with explicit round-to-nearest ops.
//   * leftIdxs[j] is last-writer-wins in left-index order (App. B.4) = atomicMax over the accepted left indices.
// stereo_finalize: one CTA per frame; bitonic sorts of (depth, idx) and (SAD, idx) give the lowest floor(1%)
//   depths and the median SAD; both outlier passes are then order independent.
// Bound: integer pipe / L2 latency; the HBM traffic of this stage is negligible (descriptors + keypoints).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace gvs {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// the 11 bytes at p as three words (the last one holds 3 bytes): 4 aligned word loads + funnel shifts.  The rows are padded
// (pitch a multiple of 64, 256 spare bytes behind the last slab), so the aligned words around the patch are always readable.
__device__ __forceinline__ void patchRow(const uint8_t* p, unsigned& w0, unsigned& w1, unsigned& w2) {
    const unsigned sh = ((unsigned)(size_t)p & 3u) * 8u;
    const unsigned* q = reinterpret_cast<const unsigned*>(p - ((size_t)p & 3u));
    const unsigned a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2), a3 = __ldg(q + 3);
    w0 = __funnelshift_r(a0, a1, sh); w1 = __funnelshift_r(a1, a2, sh); w2 = __funnelshift_r(a2, a3, sh) & 0x00ffffffu;
}

// destributeRightKeys (src/FeatureMatcher.cpp:728-752) for a batch: per frame a CSR over image rows, row r listing the right keys
// whose bucket range [mn, mx] (written by the descriptor kernel) holds r.  The order inside a row is arbitrary: the match takes the
// minimum of (distance, key index), which does not depend on it.  One CTA per frame; counters, scan and cursors in shared memory.
__global__ void __launch_bounds__(256)
row_bucket_kernel(const StereoArgs a) {
    extern __shared__ int sRow[];   // bucketRows + 1
    __shared__ int sWarp[8];
    const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int offR = a.offR[frame], nR = a.offR[frame + 1] - offR;
    const int H = a.bucketRows;
    for (int r = tid; r <= H; r += 256) sRow[r] = 0;
    __syncthreads();
    for (int j = tid; j < nR; j += 256) {
        const unsigned rg = a.auxR[offR + j].x;
        const int mn = (int)(rg & 0xffffu), mx = min((int)(rg >> 16), H - 1);
        for (int r = mn; r <= mx; r++) atomicAdd(&sRow[r], 1);
    }
    __syncthreads();
    // exclusive scan over the rows: thread t owns rows [t * per, (t + 1) * per)
    const int per = (H + 255) / 256;
    const int r0 = min(tid * per, H), r1 = min(r0 + per, H);
    int sum = 0;
    for (int r = r0; r < r1; r++) sum += sRow[r];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w = 0; w < warp; w++) base += sWarp[w];
    int total = 0;
    for (int w = 0; w < 8; w++) total += sWarp[w];
    for (int r = r0; r < r1; r++) { const int c = sRow[r]; sRow[r] = base; base += c; }
    if (tid == 0) sRow[H] = total;
    __syncthreads();
    int* rs = a.rowStart + (size_t)frame * (H + 1);
    for (int r = tid; r <= H; r += 256) rs[r] = sRow[r];
    __syncthreads();
    unsigned short* items = a.rowItems + (size_t)frame * a.rowItemsStride;
    for (int j = tid; j < nR; j += 256) {
        const unsigned rg = a.auxR[offR + j].x;
        const int mn = (int)(rg & 0xffffu), mx = min((int)(rg >> 16), H - 1);
        for (int r = mn; r <= mx; r++) {
            const int p = atomicAdd(&sRow[r], 1);
            if (p < a.rowItemsStride) items[p] = (unsigned short)j;
        }
    }
}

__global__ void __launch_bounds__(256)
stereo_match_kernel(const __grid_constant__ Geometry g, const StereoArgs a) {
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int offL = a.offL[frame], nL = a.offL[frame + 1] - offL;
    const int offR = a.offR[frame], nR = a.offR[frame + 1] - offR;
    __shared__ int sLvlStart[kMaxLevels + 1];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int l = 0; l < kMaxLevels; l++) { sLvlStart[l] = run; run += l < g.nLevels ? a.keptCntR[frame * kMaxLevels + l] : 0; }
        sLvlStart[kMaxLevels] = run;
    }
    __syncthreads();
    const float maxD = (float)a.fx;
    const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nL; i += warpsPerGrid) {
        const gvs_keypoint kl = a.kpL[offL + i];
        int rIdx = -1, sad = 0;
        float depth = -1.f;
        bool close = false;
        const int yKey = __float2int_rn(kl.y);
        const float uL = kl.y;
        const int octL = kl.octave;
        const float minU = __fsub_rn(uL, maxD), maxU = uL;
        if (!(maxU < 0.f) && yKey >= 0 && yKey < a.imageHeight && yKey < a.bucketRows) {   // (bucketRows = min(imageHeight, image rows))
            const uint4* dl = reinterpret_cast<const uint4*>(a.descL + (size_t)(offL + i) * 32);
            const uint4 l0 = __ldg(dl), l1 = __ldg(dl + 1);
            const int j0 = sLvlStart[max(octL - 1, 0)], j1 = sLvlStart[min(octL + 2, kMaxLevels)];
            unsigned best = 0xffffffffu;
            // the row bucket of yKey: right keys j with mn_j <= yKey <= mx_j; of these the levels octL-1 .. octL+1 (key ranges j0 .. j1)
            const int* rs = a.rowStart + (size_t)frame * (a.bucketRows + 1);
            const unsigned short* items = a.rowItems + (size_t)frame * a.rowItemsStride;
            const int p1 = min(rs[yKey + 1], a.rowItemsStride);
            for (int p = rs[yKey] + lane; p < p1; p += 32) {
                const int j = items[p];
                if (j < j0 || j >= j1) continue;
                const float yR = __uint_as_float(__ldg(&a.auxR[offR + j].y));
                if (!(yR >= minU && yR <= maxU)) continue;
                const uint4* dr = reinterpret_cast<const uint4*>(a.descR + (size_t)(offR + j) * 32);
                const int dist = hamming256(l0, l1, __ldg(dr), __ldg(dr + 1));
                best = min(best, ((unsigned)dist << 20) | (unsigned)j);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            const int bestDist = best == 0xffffffffu ? 256 : (int)(best >> 20);
            if (bestDist <= a.thDist) {
                const int bestIdx = (int)(best & 0xfffff);
                const float kRx = a.kpR[offR + bestIdx].x;
                const LevelGeom& lv = g.lv[octL];
                const float scale = lv.invScale;
                const int scuL = (int)roundf(__fmul_rn(kl.x, scale)), scvL = (int)roundf(__fmul_rn(kl.y, scale));
                const int scuR = (int)roundf(__fmul_rn(kRx, scale));
                const uint8_t* imL = levelPtr(a.slabL, g, frame, octL);
                const uint8_t* imR = levelPtr(a.slabR, g, frame, octL);
                // 22 lanes: shift s = lane % 11 (xMov = s - 5), rows half*6 .. half*6+5 (second half has 5 rows)
                int part = 0;
                const int s = lane % 11, half = lane / 11;
                const int startW = scuR + (s - 5) - 5, endW = scuR + (s - 5) + 6;
                const bool skip = startW < 0 || endW >= lv.w;
                if (lane < 22 && !skip) {
                    const int rEnd = half ? 11 : 6;
                    // 11 pixels of a row = 3 words cut out of 4 aligned ones; |l - r| summed four bytes at a time (VABSDIFF4)
                    for (int r = half * 6; r < rEnd; r++) {
                        unsigned l0, l1, l2, r0, r1, r2;
                        patchRow(imL + (size_t)(scvL - 5 + r) * lv.pitch + (scuL - 5), l0, l1, l2);
                        patchRow(imR + (size_t)(scvL - 5 + r) * lv.pitch + startW, r0, r1, r2);
                        part += (int)__vsadu4(l0, r0) + (int)__vsadu4(l1, r1) + (int)__vsadu4(l2, r2);
                    }
                }
                part += __shfl_down_sync(0xffffffffu, part, 11);  // lanes 0..10 now hold the full SAD of shift s
                const unsigned skipMask = __ballot_sync(0xffffffffu, skip);
                int bestDistW = 0x7fffffff, bestX = 0;
                float d[11];
#pragma unroll
                for (int t = 0; t < 11; t++) {
                    const int v = __shfl_sync(0xffffffffu, part, t);
                    const bool sk = (skipMask >> t) & 1;
                    d[t] = sk ? 0.0f : (float)v;
                    if (!sk && (float)bestDistW > (float)v) { bestX = t - 5; bestDistW = v; }
                }
                if (bestX != -5 && bestX != 5) {
                    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
                    for (int t = 1; t < 10; t++)
                        if (t == bestX + 5) { dist1 = d[t - 1]; dist2 = d[t]; dist3 = d[t + 1]; }
                    const float delta = __fdiv_rn(__fsub_rn(dist1, dist3),
                                                  __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
                    if (!(delta > 1.f || delta < -1.f)) {
                        const float newuR = __fmul_rn(lv.scale, __fadd_rn(__fadd_rn((float)scuR, (float)bestX), delta));
                        const float disparity = __fsub_rn(kl.x, newuR);
                        if (disparity > 0.0f && (double)disparity < a.fx) {
                            depth = __fdiv_rn(__fmul_rn((float)a.fx, a.baseline), disparity);
                            rIdx = bestIdx;
                            sad = bestDistW;
                            close = depth < __fmul_rn(a.baseline, (float)a.closeNumber);
                        }
                    }
                }
            }
        }
        if (lane == 0) {
            a.rightIdxs[offL + i] = rIdx;
            a.depth[offL + i] = depth;
            a.close[offL + i] = close ? 1 : 0;
            a.sadBest[offL + i] = sad;
            if (rIdx >= 0) atomicMax(&a.leftIdxs[offR + rIdx], i);
        }
    }
}

// ascending bitonic sort of n 64-bit keys (n a power of two) in shared memory
__device__ void bitonicSort(unsigned long long* k, int n) {
    for (int size = 2; size <= n; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = k[lo], y = k[hi];
                if ((x > y) == up) { k[lo] = y; k[hi] = x; }
            }
        }
    __syncthreads();
}

// The two outlier passes of findStereoMatchesORB2R (src/FeatureMatcher.cpp:655-706) need the floor(1 %) smallest (depth, index)
// pairs of the frame's matches and the median SAD, not sorted arrays: the first is a handful of block-wide arg-min rounds (a bitonic
// sort when the frame has more than kMinRounds * 100 matches), the second a two-level radix select on the 15-bit SAD.
constexpr int kMinRounds = 16;

__device__ __forceinline__ unsigned long long warpMin64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

__global__ void __launch_bounds__(256)
stereo_finalize_kernel(const StereoArgs a, int sortCap, unsigned long long* __restrict__ gKeys) {
    extern __shared__ unsigned long long smKeys[];
    // sort keys live in shared memory when the per-image key capacity fits, in global scratch otherwise (very large nFeatures)
    unsigned long long* sKeys = gKeys ? gKeys + (size_t)blockIdx.x * sortCap : smKeys;
    __shared__ int sCount, sHist[256], sSel[2];
    __shared__ unsigned long long sWarpMin[8];
    __shared__ float sMed;
    const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int offL = a.offL[frame], nL = a.offL[frame + 1] - offL;
    const int offR = a.offR[frame];
    int mine = 0;
    for (int i = tid; i < nL; i += blockDim.x) mine += a.rightIdxs[offL + i] >= 0;
    if (tid == 0) sCount = 0;
    __syncthreads();
    if (mine) atomicAdd(&sCount, mine);
    __syncthreads();
    const int V = sCount;   // matched left keys of the frame
    if (V <= 0) return;
    // ---- pass A: the floor(1 %) smallest (depth, idx) are dropped.  They are marked through the close[] byte (2 = "drop"; close is
    // rewritten below) and applied after pass B has read the SADs: the two passes commute.
    const int endDe = (int)floor((double)V * 0.01);
    if (endDe <= kMinRounds) {
        unsigned long long last = 0;
        for (int r = 0; r < endDe; r++) {
            unsigned long long best = ~0ull;
            for (int i = tid; i < nL; i += blockDim.x) {
                if (a.rightIdxs[offL + i] < 0) continue;
                const unsigned long long k = ((unsigned long long)__float_as_uint(a.depth[offL + i]) << 32) | (unsigned)i;
                if ((r == 0 || k > last) && k < best) best = k;   // (keys are distinct: the index is part of them)
            }
            best = warpMin64(best);
            if (lane == 0) sWarpMin[warp] = best;
            __syncthreads();
            best = sWarpMin[0];
#pragma unroll
            for (int w = 1; w < 8; w++) best = sWarpMin[w] < best ? sWarpMin[w] : best;
            last = best;
            if (tid == 0) a.close[offL + (int)(last & 0xffffffffu)] = 2;
            __syncthreads();
        }
    } else {
        if (tid == 0) sCount = 0;
        __syncthreads();
        for (int i = tid; i < nL; i += blockDim.x)
            if (a.rightIdxs[offL + i] >= 0) {
                const int s = atomicAdd(&sCount, 1);
                if (s < sortCap) sKeys[s] = ((unsigned long long)__float_as_uint(a.depth[offL + i]) << 32) | (unsigned)i;
            }
        __syncthreads();
        const int Vs = min(V, sortCap);
        int n2 = 1;
        while (n2 < Vs) n2 <<= 1;
        for (int i = Vs + tid; i < n2; i += blockDim.x) sKeys[i] = ~0ull;
        bitonicSort(sKeys, n2);
        for (int i = tid; i < min(endDe, Vs); i += blockDim.x) a.close[offL + (int)(sKeys[i] & 0xffffffffu)] = 2;
    }
    // ---- pass B: the median SAD = element V / 2 of the ascending (SAD, idx) order; only its SAD is used.  SAD <= 11 * 11 * 255 < 2^15:
    // bits 14..7 pick one of 256 bins, bits 6..0 one of 128 inside it.
    int target = V / 2, prefix = 0;
    for (int level = 0; level < 2; level++) {
        __syncthreads();
        sHist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < nL; i += blockDim.x) {
            if (a.rightIdxs[offL + i] < 0) continue;
            const int v = a.sadBest[offL + i];
            if (level == 0) atomicAdd(&sHist[min(v >> 7, 255)], 1);
            else if (min(v >> 7, 255) == prefix) atomicAdd(&sHist[v & 127], 1);
        }
        __syncthreads();
        if (warp == 0) {   // the bin that holds rank `target`: 8 bins per lane, warp scan
            int c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { c[k] = sHist[8 * lane + k]; sum += c[k]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int run = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (target >= run && target < run + c[k]) { sSel[0] = 8 * lane + k; sSel[1] = target - run; }
                run += c[k];
            }
        }
        __syncthreads();
        if (level == 0) { prefix = sSel[0]; target = sSel[1]; }
    }
    if (tid == 0) {
        const int medianD = (prefix << 7) | sSel[0];
        sMed = __fmul_rn((float)medianD, __fmul_rn(1.5f, 1.4f));
    }
    __syncthreads();
    const float medDistD = sMed;
    for (int i = tid; i < nL; i += blockDim.x) {
        const int r = a.rightIdxs[offL + i];
        if (r < 0) continue;
        const bool drop = a.close[offL + i] == 2 || !((float)a.sadBest[offL + i] < medDistD);
        if (drop) {
            a.leftIdxs[offR + r] = -1;
            a.rightIdxs[offL + i] = -1;
            a.depth[offL + i] = -1.f;
            a.close[offL + i] = 0;
        }
    }
}

void launchStereo(const Geometry& g, const StereoArgs& a, int nFrames, int maxKeysPerImage, cudaStream_t s, int* launches) {
    if (nFrames <= 0) return;
    const int warpsPerCta = 8;
    static const int keysPerWarp = [] { const char* e = getenv("GVS_STEREO_KEYS"); return e ? std::max(atoi(e), 1) : 8; }();
    int gx = (maxKeysPerImage / 2 + keysPerWarp * warpsPerCta - 1) / (keysPerWarp * warpsPerCta);  // left keys per warp at typical fill
    if (gx < 1) gx = 1;
    // the row buckets of the right keys first (one CTA per frame)
    static size_t cfgRow[64];
    const size_t rowSm = ((size_t)a.bucketRows + 1) * sizeof(int);
    ensureDynamicSmem(row_bucket_kernel, rowSm, cfgRow);
    row_bucket_kernel<<<nFrames, 256, rowSm, s>>>(a);
    ++*launches;
    stereo_match_kernel<<<dim3(gx, nFrames), 32 * warpsPerCta, 0, s>>>(g, a);
    ++*launches;
    int n2 = 1;
    while (n2 < maxKeysPerImage) n2 <<= 1;
    size_t sm = (size_t)n2 * 8;
    unsigned long long* gKeys = nullptr;
    if (sm > 160 * 1024) { gKeys = reinterpret_cast<unsigned long long*>(a.sortScratch); sm = 0; }
    static size_t cfg[64];
    ensureDynamicSmem(stereo_finalize_kernel, sm, cfg);
    stereo_finalize_kernel<<<nFrames, 256, sm, s>>>(a, n2, gKeys);
    ++*launches;
}

// ------------------------------------------------------------------ K9 grid_assign
// One warp per image: bin = cvRound(pt * mult) clamped (src/FeatureTracker.cpp:40-49); CSR with insertion
// (= index) order inside a cell: counting pass, exclusive scan, stable fill in chunks of 32 keys.
__global__ void __launch_bounds__(32)
grid_assign_kernel(const GridArgs a0, const GridArgs a1) {   // blockIdx.y = side (the grids of both sides in one launch)
    extern __shared__ int sCell[];  // xGrids * yGrids + 1
    const GridArgs& a = blockIdx.y ? a1 : a0;
    const int img = blockIdx.x, lane = threadIdx.x;
    const int off = a.off[img], n = a.off[img + 1] - off;
    const int nCells = a.xGrids * a.yGrids;
    for (int c = lane; c <= nCells; c += 32) sCell[c] = 0;
    __syncwarp();
    auto cellOf = [&](int i) {
        const gvs_keypoint k = a.kps[off + i];
        int xPos = __float2int_rn(__fmul_rn(k.x, a.xMult)), yPos = __float2int_rn(__fmul_rn(k.y, a.yMult));
        xPos = min(max(xPos, 0), a.xGrids - 1);
        yPos = min(max(yPos, 0), a.yGrids - 1);
        return yPos * a.xGrids + xPos;
    };
    for (int i = lane; i < n; i += 32) atomicAdd(&sCell[cellOf(i)], 1);
    __syncwarp();
    // exclusive scan over cells
    int run = 0;
    for (int base = 0; base < nCells; base += 32) {
        const int c = base + lane;
        const int v = c < nCells ? sCell[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (c < nCells) sCell[c] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) sCell[nCells] = run;
    __syncwarp();
    int* cs = a.cellStart + (size_t)img * (a.maxCells + 1);
    for (int c = lane; c <= nCells; c += 32) cs[c] = sCell[c];
    __syncwarp();
    // stable fill
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const int c = i < n ? cellOf(i) : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (i < n) a.items[off + sCell[c] + __popc(peers & lt)] = i;
        __syncwarp();
        if (i < n && (peers & lt) == 0) sCell[c] += __popc(peers);
        __syncwarp();
    }
}

void launchGrid(const GridArgs& a, int nImages, cudaStream_t s, int* launches) { launchGrid2(a, nullptr, nImages, s, launches); }

// b != nullptr: the grids of a second set of keys (the right images) in the same launch
void launchGrid2(const GridArgs& a, const GridArgs* b, int nImages, cudaStream_t s, int* launches) {
    if (nImages <= 0) return;
    const size_t sm = ((size_t)a.xGrids * a.yGrids + 1) * sizeof(int);
    static size_t cfg[64];
    ensureDynamicSmem(grid_assign_kernel, sm, cfg);
    grid_assign_kernel<<<dim3(nImages, b ? 2 : 1), 32, sm, s>>>(a, b ? *b : a);
    ++*launches;
}

}  // namespace gvs
