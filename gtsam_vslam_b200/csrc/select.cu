// K3+K4 select: per (image, level): restore the reference's candidate order, sort by response exactly like
// cv::sortIdx (libstdc++ introsort, unstable: the tie order is data dependent), and run the SSC binary search
// (Suppression via Square Covering).  Replaces the ordered emplace_back of computeKeypointsORBNew
// (reference src/FeatureExtractor.cpp:596-603) and FeatureExtractor::ssc (:368-468).
//
// One persistent CTA per resident slot, problems (level-major, biggest first) taken from an atomic counter.
//
// (a) order:   exclusive scan of the per-cell counts (cell-row-major) + rank inside the cell gives each
//              record's position in the reference's emission order; keys and coordinates are scattered
//              into shared memory in that order.
// (b) sort:    std::sort(idx, by key[a] < key[b]) followed by array reversal is emulated
//              exactly but in parallel:
//              * __unguarded_partition is a two-pointer sweep that swaps the k-th element >= pivot from
//                the left with the k-th element <= pivot from the right until the pointers cross; the set
//                of swaps and the cut are closed-form in prefix/suffix ranks, so a partition is three
//                data-parallel passes (count, rank+assign, swap).  Segments >= 1024 are partitioned by the
//                whole CTA, smaller ones breadth-first by one warp each; median-of-3, the depth limit
//                2*floor(log2 n) and the heapsort fallback (one lane, rare) follow libstdc++ literally.
//              * __final_insertion_sort moves each element left past strictly greater ones, i.e. it is a
//                stable sort of the partitioned array, done here as a stable counting sort on the 8-bit
//                score; the descending order is the reversed array.
//              Validated against cv2.sortIdx / std::sort in tests (including median-of-3 killer inputs).
// (c) ssc:     a pass keeps a point iff its cell (side width/2) is not covered by an earlier kept point, and
//              a kept point covers the 5x5 cells around it.  Hence at most the first point of a cell can be
//              kept and a cell "fires" iff no neighbouring cell (+-2) with a smaller first rank fired: an
//              exact greedy independent set on the cell grid, resolved in parallel rounds (a decision is
//              final once made).  Kept points are compacted in sorted order by a block scan.
//              floor(y / c) with c = width / 2.0 equals (2*y) / width in integers (exact).
#include <cstdlib>

#include "common.cuh"

namespace gvs {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSmallMax = 6144;    // problems with n <= kSmallMax run entirely in shared memory
constexpr int kBigKeySmem = 64 * 1024;  // global-memory class: the 8-bit keys of a problem of up to this many candidates stay in shared memory
constexpr int kBigSeg = 1024;      // segments at least this long are partitioned by the whole CTA
constexpr int kSscWarps = 7;       // widths evaluated speculatively per super-round

template <typename IdxT>
struct SelArrays {
    uint8_t* key;    // [n]   score, emission order
    unsigned* xy;    // [n]   x | y << 12 | score << 24, emission order
    unsigned* sxy;   // [n]   the same in sorted order (global-memory class only, nullptr otherwise)
    IdxT* idx;       // [n]   permutation
    IdxT* tmp;       // [n]   scratch / sorted order
    int* cellOff;    // [cells + 1]
    int* seg;        // small-segment frontiers 2 x segCap x 3
    int segCap;
    int* hist;       // kSelWarps x 256 counting-sort counters
    unsigned* bitmaps;      // kSscWarps x kSscBitmapWords coverage bitmaps (shared)
    void* lists;            // kSscWarps x listCap kept positions (shared)
    int listCap;
    unsigned* globalBitmap; // fallback bitmap for tiny widths (global scratch)
};

struct SelShared {
    int prob;
    int bigTop;
    int bigSeg[64][3];
    int nSeg[2];
    int wA[kSelWarps], wB[kSelWarps], wM[kSelWarps], wMin[kSelWarps];
    int cut, pv, m, any;
    int scan[kSelWarps];
    int keptTotal;
};

#ifdef GVS_SEL_PROFILE
// phase clocks (thread 0 of every CTA): 0 order, 1 sort phase 1 (CTA-wide partitions), 2 sort phase 2 (warp partitions),
// 3 final counting sort, 4 ssc, 5 output, 6 problems
__device__ unsigned long long gSelProf[8];
#define GVS_PROF_DECL long long profT = clock64();
#define GVS_PROF_DECL2 profT = clock64();
#define GVS_PROF(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&gSelProf[i], (unsigned long long)(t_ - profT)); profT = t_; } } while (0)
#else
#define GVS_PROF_DECL
#define GVS_PROF_DECL2
#define GVS_PROF(i) do { } while (0)
#endif

__device__ __forceinline__ unsigned lanemaskLt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ unsigned lanemaskLe() { unsigned m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

// ---- libstdc++ pieces executed by a single lane -----------------------------------------------------
// The 8-bit keys are permuted together with idx (key[p] is the key of the element at position p), so no pass needs the
// double indirection key[idx[p]] and key loads are contiguous.
template <typename IdxT>
__device__ __forceinline__ void swapPos(uint8_t* key, IdxT* idx, int i, int j) {
    const IdxT t = idx[i]; idx[i] = idx[j]; idx[j] = t;
    const uint8_t k = key[i]; key[i] = key[j]; key[j] = k;
}

template <typename IdxT>
__device__ void moveMedianToFirst(uint8_t* key, IdxT* idx, int result, int a, int b, int c) {
    const int ka = key[a], kb = key[b], kc = key[c];
    int sel;
    if (ka < kb) sel = kb < kc ? b : (ka < kc ? c : a);
    else sel = ka < kc ? a : (kb < kc ? c : b);
    swapPos(key, idx, result, sel);
}

template <typename IdxT>
__device__ void adjustHeap(uint8_t* key, IdxT* idx, int hole, int len, IdxT value, uint8_t vkey) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (key[second] < key[second - 1]) second--;
        idx[hole] = idx[second]; key[hole] = key[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        idx[hole] = idx[second - 1]; key[hole] = key[second - 1];
        hole = second - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && key[parent] < vkey) {
        idx[hole] = idx[parent]; key[hole] = key[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    idx[hole] = value; key[hole] = vkey;
}

// std::__partial_sort(first, last, last) on positions [0, len) of (idx, key)
template <typename IdxT>
__device__ void heapSort(uint8_t* key, IdxT* idx, int len) {
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            adjustHeap(key, idx, parent, len, idx[parent], key[parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        const IdxT v = idx[last]; const uint8_t vk = key[last];
        idx[last] = idx[0]; key[last] = key[0];
        adjustHeap(key, idx, 0, last, v, vk);
    }
}

// ---- warp-level pieces of the parallel partition ----------------------------------------------------
// count of (key >= pv) and (key <= pv) over positions [a, b)
__device__ __forceinline__ void countAB(const uint8_t* key, int a, int b, int pv, int lane, int& cA, int& cB) {
    int na = 0, nb = 0;
    for (int p = a + lane; p < b; p += 32) {
        const int k = key[p];
        na += k >= pv;
        nb += k <= pv;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { na += __shfl_xor_sync(0xffffffffu, na, o); nb += __shfl_xor_sync(0xffffffffu, nb, o); }
    cA = na; cB = nb;
}

// rank + assign over [a, b): runA/runB = counts before a; totalB = count of B over the whole partition range.
// Writes posA[k] / posB[k] for swapped elements; returns number swapped (A side) and the smallest unswapped A position.
template <typename IdxT>
__device__ __forceinline__ void assignAB(const uint8_t* key, int a, int b, int pv, int lane, int runA,
                                         int runB, int totalB, IdxT* posA, IdxT* posB, int& mOut, int& minUnsw) {
    int m = 0, mn = 0x7fffffff;
    const unsigned lt = lanemaskLt(), le = lanemaskLe();
    for (int base4 = a; base4 < b; base4 += 128) {
        int kk[4];  // the keys of four chunks are fetched before the (store-carrying) rank code: four loads in flight
#pragma unroll
        for (int j = 0; j < 4; j++) { const int p = base4 + 32 * j + lane; kk[j] = p < b ? (int)key[p] : -1; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int base = base4 + 32 * j;
            if (base >= b) break;
            const int p = base + lane;
            const int k = kk[j];
            const bool fa = k >= pv, fb = k >= 0 && k <= pv;   // k = -1 marks a position past the end
            const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb);
            const int ra = runA + __popc(ba & lt);
            const int rbx = totalB - (runB + __popc(bb & le));
            const bool swA = fa && rbx > ra, swB = fb && ra > rbx;
            if (swA) posA[ra] = (IdxT)p;
            if (swB) posB[rbx] = (IdxT)p;
            if (fa && !swA) mn = min(mn, p);
            m += __popc(__ballot_sync(0xffffffffu, swA));
            runA += __popc(ba);
            runB += __popc(bb);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mOut = m; minUnsw = mn;
}

// Segments of at most 32 elements entirely in registers (lane = position): median-of-3, __unguarded_partition and the
// descent into the one child that can still exceed 16 elements, with shuffles and ballots only -- one load and one store of
// idx per element instead of several shared-memory round trips per partition.  Same swaps as partitionWarp: the swapped
// "A" elements (key >= pivot, scanning up) are a prefix of the A lanes, the swapped "B" elements (key <= pivot, scanning
// down) a suffix of the B lanes, the r-th of one is exchanged with the r-th of the other.
template <typename IdxT>
__device__ void sortSmallSegmentWarp(uint8_t* key, IdxT* idx, int first, int last, int depth, int lane) {
    const int m0 = last - first;
    IdxT e = lane < m0 ? idx[first + lane] : (IdxT)0;
    int k = lane < m0 ? (int)key[first + lane] : 0;
    int f = 0, l = m0;  // current sub-segment, lane relative
    const unsigned ltm = lanemaskLt(), lem = lanemaskLe();
    while (l - f > 16) {
        if (depth == 0) {  // introsort's heapsort fallback: back to shared memory, one lane
            if (lane < m0) { idx[first + lane] = e; key[first + lane] = (uint8_t)k; }
            __syncwarp();
            if (lane == 0) heapSort(key + first + f, idx + first + f, l - f);
            __syncwarp();
            return;
        }
        --depth;
        // __move_median_to_first(f, f + 1, mid, l - 1)
        const int ia = f + 1, ib = f + (l - f) / 2, ic = l - 1;
        const int ka = __shfl_sync(0xffffffffu, k, ia), kb = __shfl_sync(0xffffffffu, k, ib), kc = __shfl_sync(0xffffffffu, k, ic);
        int sel;
        if (ka < kb) sel = kb < kc ? ib : (ka < kc ? ic : ia);
        else sel = ka < kc ? ia : (kb < kc ? ic : ib);
        {
            const int src = lane == f ? sel : (lane == sel ? f : lane);
            e = (IdxT)__shfl_sync(0xffffffffu, (int)e, src);
            k = __shfl_sync(0xffffffffu, k, src);
        }
        const int pv = __shfl_sync(0xffffffffu, k, f);
        // __unguarded_partition(f + 1, l, pivot at f)
        const bool in = lane > f && lane < l;
        const bool fa = in && k >= pv, fb = in && k <= pv;
        const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb);
        const int ra = __popc(ba & ltm), rbx = __popc(bb) - __popc(bb & lem);
        const bool swA = fa && rbx > ra, swB = fb && ra > rbx;
        int src = lane;
        if (swA) src = __fns(bb, 31, -(ra + 1));       // the (ra+1)-th B lane from the top
        else if (swB) src = __fns(ba, 0, rbx + 1);     // the (rbx+1)-th A lane from the bottom
        const unsigned unswA = __ballot_sync(0xffffffffu, fa && !swA), swBm = __ballot_sync(0xffffffffu, swB);
        e = (IdxT)__shfl_sync(0xffffffffu, (int)e, src);
        k = __shfl_sync(0xffffffffu, k, src);
        const int cut = min(unswA ? __ffs(unswA) - 1 : 0x7fffffff, swBm ? __ffs(swBm) - 1 : 0x7fffffff);
        // at most one child can still be larger than 16 (the two sizes add up to at most 32)
        if (l - cut > 16) f = cut;
        else l = cut;
    }
    if (lane < m0) { idx[first + lane] = e; key[first + lane] = (uint8_t)k; }
    __syncwarp();
}

// Partition [first, last) with one warp; returns cut.
template <typename IdxT>
__device__ int partitionWarp(uint8_t* key, IdxT* idx, IdxT* tmp, int first, int last, int lane) {
    const int len = last - first, lo = first + 1;
    // __move_median_to_first(first, first + 1, mid, last - 1): three lanes fetch, every lane decides
    const int ia = first + 1, ib = first + len / 2, ic = last - 1;
    const int kq = key[lane == 0 ? ia : (lane == 1 ? ib : ic)];
    const int ka = __shfl_sync(0xffffffffu, kq, 0), kb = __shfl_sync(0xffffffffu, kq, 1), kc = __shfl_sync(0xffffffffu, kq, 2);
    int sel, pv;
    if (ka < kb) { if (kb < kc) { sel = ib; pv = kb; } else if (ka < kc) { sel = ic; pv = kc; } else { sel = ia; pv = ka; } }
    else if (ka < kc) { sel = ia; pv = ka; }
    else if (kb < kc) { sel = ic; pv = kc; }
    else { sel = ib; pv = kb; }
    if (lane == 0) swapPos(key, idx, first, sel);
    __syncwarp();
    int cA, cB;
    countAB(key, lo, last, pv, lane, cA, cB);
    IdxT* posA = tmp + first;
    IdxT* posB = tmp + first + (len + 1) / 2;
    int m, mn;
    assignAB(key, lo, last, pv, lane, 0, 0, cB, posA, posB, m, mn);
    __syncwarp();
    const int bm1 = m > 0 ? (int)posB[m - 1] : 0x7fffffff;
    for (int k = lane; k < m; k += 32) swapPos(key, idx, (int)posA[k], (int)posB[k]);
    __syncwarp();
    return min(mn, bm1);
}

template <typename IdxT>
__device__ void sortDescending(const SelArrays<IdxT>& A, SelShared& S, int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* key = A.key;
    IdxT* idx = A.idx;
    IdxT* tmp = A.tmp;
    GVS_PROF_DECL
    for (int i = tid; i < n; i += kSelThreads) idx[i] = (IdxT)i;
    if (tid == 0) {
        S.bigTop = 0; S.nSeg[0] = 0; S.nSeg[1] = 0;
        if (n > 16) {
            const int depth = 2 * (31 - __clz(n));
            if (n >= kBigSeg) { S.bigSeg[0][0] = 0; S.bigSeg[0][1] = n; S.bigSeg[0][2] = depth; S.bigTop = 1; }
            else { A.seg[0] = 0; A.seg[1] = n; A.seg[2] = depth; S.nSeg[0] = 1; }
        }
    }
    __syncthreads();

    // ---- phase 1: big segments, whole CTA per partition
    while (S.bigTop > 0) {
        const int top = S.bigTop - 1;
        const int first = S.bigSeg[top][0], last = S.bigSeg[top][1], depth = S.bigSeg[top][2];
        __syncthreads();
        if (depth == 0) {
            if (tid == 0) { heapSort(key + first, idx + first, last - first); S.bigTop = top; }
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            moveMedianToFirst(key, idx, first, first + 1, first + (last - first) / 2, last - 1);
            S.pv = key[first];
        }
        __syncthreads();
        const int pv = S.pv, lo = first + 1, len = last - first;
        const int chunk = (((last - lo) + kSelWarps - 1) / kSelWarps + 31) & ~31;
        const int a = min(lo + warp * chunk, last), b = min(a + chunk, last);
        int cA, cB;
        countAB(key, a, b, pv, lane, cA, cB);
        if (lane == 0) { S.wA[warp] = cA; S.wB[warp] = cB; }
        __syncthreads();
        int runA = 0, runB = 0, totalB = 0;
#pragma unroll
        for (int w = 0; w < kSelWarps; w++) {
            if (w < warp) { runA += S.wA[w]; runB += S.wB[w]; }
            totalB += S.wB[w];
        }
        IdxT* posA = tmp + first;
        IdxT* posB = tmp + first + (len + 1) / 2;
        int m, mn;
        assignAB(key, a, b, pv, lane, runA, runB, totalB, posA, posB, m, mn);
        if (lane == 0) { S.wM[warp] = m; S.wMin[warp] = mn; }
        __syncthreads();
        int mTot = 0, mnAll = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < kSelWarps; w++) { mTot += S.wM[w]; mnAll = min(mnAll, S.wMin[w]); }
        const int bm1 = mTot > 0 ? (int)posB[mTot - 1] : 0x7fffffff;
        const int cut = min(mnAll, bm1);
        __syncthreads();  // everyone has read posB[mTot-1] before tmp is reused by children
        for (int k = tid; k < mTot; k += kSelThreads) swapPos(key, idx, (int)posA[k], (int)posB[k]);
        if (tid == 0) {
            int t2 = top;
            const int d = depth - 1;
            auto push = [&](int f, int l) {
                if (l - f >= kBigSeg) { S.bigSeg[t2][0] = f; S.bigSeg[t2][1] = l; S.bigSeg[t2][2] = d; t2++; }
                else if (l - f > 16) { int s = S.nSeg[0]++; A.seg[3 * s] = f; A.seg[3 * s + 1] = l; A.seg[3 * s + 2] = d; }
            };
            push(cut, last);
            push(first, cut);
            S.bigTop = t2;
        }
        __syncthreads();
    }

    GVS_PROF(1);
    // ---- phase 2: small segments, breadth first, one warp per segment
    int cur = 0;
    while (true) {
        const int ns = S.nSeg[cur];
        if (ns == 0) break;
        int* segIn = A.seg + cur * 3 * A.segCap;
        int* segOut = A.seg + (cur ^ 1) * 3 * A.segCap;
        for (int s = warp; s < ns; s += kSelWarps) {
            const int first = segIn[3 * s], last = segIn[3 * s + 1], depth = segIn[3 * s + 2];
            if (depth == 0) {
                if (lane == 0) heapSort(key + first, idx + first, last - first);
                continue;
            }
            if (last - first <= 32) { sortSmallSegmentWarp(key, idx, first, last, depth, lane); continue; }
            const int cut = partitionWarp(key, idx, tmp, first, last, lane);
            // children of at most 32 elements are finished right away in registers, larger ones go to the next level
            if (last - cut > 32) {
                if (lane == 0) { int o = atomicAdd(&S.nSeg[cur ^ 1], 1); segOut[3 * o] = cut; segOut[3 * o + 1] = last; segOut[3 * o + 2] = depth - 1; }
            } else if (last - cut > 16) sortSmallSegmentWarp(key, idx, cut, last, depth - 1, lane);
            if (cut - first > 32) {
                if (lane == 0) { int o = atomicAdd(&S.nSeg[cur ^ 1], 1); segOut[3 * o] = first; segOut[3 * o + 1] = cut; segOut[3 * o + 2] = depth - 1; }
            } else if (cut - first > 16) sortSmallSegmentWarp(key, idx, first, cut, depth - 1, lane);
        }
        __syncthreads();
        if (tid == 0) S.nSeg[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }

    GVS_PROF(2);
    // ---- final insertion sort == stable sort by key; reversed -> tmp holds the descending order
    int* hist = A.hist;
    for (int i = tid; i < kSelWarps * 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    const int chunk = ((n + kSelWarps - 1) / kSelWarps + 31) & ~31;
    const int a = min(warp * chunk, n), b = min(a + chunk, n);
    for (int base = a; base < b; base += 32) {
        const int p = base + lane;
        const int k = p < b ? (int)key[p] : 256 + lane;  // inactive lanes get unique keys
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (p < b && (peers & lanemaskLt()) == 0) hist[warp * 256 + k] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive scan in (key, warp) order; thread k owns key k
        int tot = 0;
        for (int w = 0; w < kSelWarps; w++) tot += hist[w * 256 + tid];
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) S.scan[warp] = incl;
        __syncthreads();
        int base = incl - tot;
        for (int w = 0; w < warp; w++) base += S.scan[w];
        for (int w = 0; w < kSelWarps; w++) { const int c = hist[w * 256 + tid]; hist[w * 256 + tid] = base; base += c; }
    }
    __syncthreads();
    for (int base = a; base < b; base += 32) {
        const int p = base + lane;
        const IdxT v = p < b ? idx[p] : (IdxT)0;
        const int k = p < b ? (int)key[p] : 256 + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (p < b) {
            const int pos = hist[warp * 256 + k] + __popc(peers & lanemaskLt());
            tmp[n - 1 - pos] = v;
        }
        __syncwarp();
        if (p < b && (peers & lanemaskLt()) == 0) hist[warp * 256 + k] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    GVS_PROF(3);
}

// ---- SSC (src/FeatureExtractor.cpp:410-461) --------------------------------------------------------------
// One pass at `width`, executed by ONE warp over all n points in sorted order with a coverage bitmap
// (1 bit per cell of side width/2).  32 points are tested against the bitmap at once; the uncovered ones are
// resolved in order inside the chunk with ballots (the first uncovered lane is kept, covers its 5x5 cells and
// knocks out the later lanes that fall inside), so the result is exactly the sequential greedy cover.
// floor(y / c), c = width / 2.0  ==  (2*y) / width in integers (exact: a non-integer quotient is at least
// 1/width away from an integer, far more than the double rounding error; an integer quotient is exact).
// `limit`: a speculative pass may stop as soon as it has kept more than `limit` points (the search only needs to know
// "more than Kmax" to choose its direction); it then returns limit + 1 and its list is incomplete.
// The bitmap carries a border of two cells on every side (row pitch ncc + 4), so that the 5 x 5 block of a kept point never
// needs clamping: lane r < 5 sets the five bits starting at bit(point) + (r - 2) * pitch - 2.
__host__ __device__ inline int sscBitmapCells(int cols, int rows, int width) {
    return ((2 * cols) / width + 5) * ((2 * rows) / width + 5);
}

template <typename IdxT>
__device__ int sscPassWarp(const unsigned* xy, const IdxT* sorted, int n, int width, int ncc, int ncr, unsigned* bitmap,
                           IdxT* list, int listCap, int lane, int limit) {
    const int pitch = ncc + 4;
    const int bmWords = (pitch * (ncr + 4) + 31) >> 5;
    for (int i = lane; i < bmWords; i += 32) bitmap[i] = 0;
    __syncwarp();
    // floor(v / width) = umulhi(v + 1, floor((2^32 - 1) / width)) for every width >= 1 while (v + 1) * width < 2^32
    // (round-down magic number; unlike the round-up form it also exists for width 1)
    const unsigned magic = 0xffffffffu / (unsigned)width;
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        int row = 0, col = 0, bit = 0;
        bool active = false;
        if (i < n) {
            const unsigned q = xy[sorted[i]];
            col = (int)__umulhi(2 * (q & 0xfff) + 1, magic);
            row = (int)__umulhi(2 * ((q >> 12) & 0xfff) + 1, magic);
            bit = (row + 2) * pitch + col + 2;
            active = !((bitmap[bit >> 5] >> (bit & 31)) & 1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, active);
        const int rc = (row << 16) | col;
        // Two uncovered lanes per step: the first one (f1) is always kept; the second one (f2) is kept in the same step unless
        // f1 covers it, which every lane can tell from the two shuffled cells.  Lanes 0..4 / 8..12 set the two 5 x 5 blocks.
        const int coverOff2 = ((lane & 7) - 2) * pitch - 2;
        while (mask) {
            const int f1 = __ffs(mask) - 1;
            const unsigned rest = mask & (mask - 1);
            const int f2 = rest ? __ffs(rest) - 1 : f1;
            const int c1 = __shfl_sync(0xffffffffu, rc, f1), c2 = __shfl_sync(0xffffffffu, rc, f2);
            const int r1 = c1 >> 16, q1 = c1 & 0xffff, r2 = c2 >> 16, q2 = c2 & 0xffff;
            const bool both = rest != 0 && !((unsigned)(r2 - r1 + 2) <= 4u && (unsigned)(q2 - q1 + 2) <= 4u);
            const bool second = lane >= 8;
            if ((lane & 7) < 5 && lane < 16 && (!second || both)) {
                const int sbit = ((second ? r2 : r1) + 2) * pitch + (second ? q2 : q1) + 2 + coverOff2, sh = sbit & 31;
                atomicOr(&bitmap[sbit >> 5], 0x1fu << sh);
                if (sh > 27) atomicOr(&bitmap[(sbit >> 5) + 1], 0x1fu >> (32 - sh));
            } else if (lane == 5) { if (cnt < listCap) list[cnt] = (IdxT)(base + f1); }
            else if (lane == 13 && both) { if (cnt + 1 < listCap) list[cnt + 1] = (IdxT)(base + f2); }
            cnt += both ? 2 : 1;
            if ((unsigned)(row - r1 + 2) <= 4u && (unsigned)(col - q1 + 2) <= 4u) active = false;
            if (both && (unsigned)(row - r2 + 2) <= 4u && (unsigned)(col - q2 + 2) <= 4u) active = false;
            mask = __ballot_sync(0xffffffffu, active);
        }
        __syncwarp();
        if (cnt > limit) return limit + 1;
    }
    return cnt;
}

// sorted == nullptr: xy is already in sorted order (the global-memory class gathers it once after the sort, so that a pass streams
// it instead of chasing sorted -> xy through L2).  kU chunks of 32 points are loaded per outer step (kU = 4 for the global class:
// four independent loads in flight; 1 for the shared-memory class, where look-ahead only costs registers).
template <typename IdxT, int kU>
__device__ int sscPassStream(const unsigned* xy, const IdxT* sorted, int n, int width, int ncc, int ncr, unsigned* bitmap,
                           IdxT* list, int listCap, int lane, int limit) {
    const int pitch = ncc + 4;
    const int bmWords = (pitch * (ncr + 4) + 31) >> 5;
    for (int i = lane; i < bmWords; i += 32) bitmap[i] = 0;
    __syncwarp();
    // floor(v / width) = umulhi(v + 1, floor((2^32 - 1) / width)) for every width >= 1 while (v + 1) * width < 2^32
    // (round-down magic number; unlike the round-up form it also exists for width 1)
    const unsigned magic = 0xffffffffu / (unsigned)width;
    int cnt = 0;
    const int coverOff2 = ((lane & 7) - 2) * pitch - 2;
    for (int base0 = 0; base0 < n; base0 += 32 * kU) {
        unsigned q4[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const int i = base0 + 32 * u + lane;
            q4[u] = i < n ? (sorted ? xy[sorted[i]] : xy[i]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const int base = base0 + 32 * u;
            if (base >= n) break;
            const int i = base + lane;
            int row = 0, col = 0, bit = 0;
            bool active = false;
            if (i < n) {
                const unsigned q = q4[u];
                col = (int)__umulhi(2 * (q & 0xfff) + 1, magic);
                row = (int)__umulhi(2 * ((q >> 12) & 0xfff) + 1, magic);
                bit = (row + 2) * pitch + col + 2;
                active = !((bitmap[bit >> 5] >> (bit & 31)) & 1);
            }
            unsigned mask = __ballot_sync(0xffffffffu, active);
            const int rc = (row << 16) | col;
            // Two uncovered lanes per step: the first one (f1) is always kept; the second one (f2) is kept in the same step unless
            // f1 covers it, which every lane can tell from the two shuffled cells.  Lanes 0..4 / 8..12 set the two 5 x 5 blocks.
            while (mask) {
                const int f1 = __ffs(mask) - 1;
                const unsigned rest = mask & (mask - 1);
                const int f2 = rest ? __ffs(rest) - 1 : f1;
                const int c1 = __shfl_sync(0xffffffffu, rc, f1), c2 = __shfl_sync(0xffffffffu, rc, f2);
                const int r1 = c1 >> 16, q1 = c1 & 0xffff, r2 = c2 >> 16, q2 = c2 & 0xffff;
                const bool both = rest != 0 && !((unsigned)(r2 - r1 + 2) <= 4u && (unsigned)(q2 - q1 + 2) <= 4u);
                const bool second = lane >= 8;
                if ((lane & 7) < 5 && lane < 16 && (!second || both)) {
                    const int sbit = ((second ? r2 : r1) + 2) * pitch + (second ? q2 : q1) + 2 + coverOff2, sh = sbit & 31;
                    atomicOr(&bitmap[sbit >> 5], 0x1fu << sh);
                    if (sh > 27) atomicOr(&bitmap[(sbit >> 5) + 1], 0x1fu >> (32 - sh));
                } else if (lane == 5) { if (cnt < listCap) list[cnt] = (IdxT)(base + f1); }
                else if (lane == 13 && both) { if (cnt + 1 < listCap) list[cnt + 1] = (IdxT)(base + f2); }
                cnt += both ? 2 : 1;
                if ((unsigned)(row - r1 + 2) <= 4u && (unsigned)(col - q1 + 2) <= 4u) active = false;
                if (both && (unsigned)(row - r2 + 2) <= 4u && (unsigned)(col - q2 + 2) <= 4u) active = false;
                mask = __ballot_sync(0xffffffffu, active);
            }
            __syncwarp();
            if (cnt > limit) return limit + 1;
        }
    }
    return cnt;
}

// the pass of the shared-memory class (kU == 1) or of the global-memory class
template <typename IdxT, int kU>
__device__ __forceinline__ int sscPass(const unsigned* xy, const IdxT* sorted, int n, int width, int ncc, int ncr, unsigned* bitmap,
                                       IdxT* list, int listCap, int lane, int limit) {
    if (kU == 1) return sscPassWarp<IdxT>(xy, sorted, n, width, ncc, ncr, bitmap, list, listCap, lane, limit);
    return sscPassStream<IdxT, kU>(xy, sorted, n, width, ncc, ncr, bitmap, list, listCap, lane, limit);
}

// The reference's binary search over the width.  The widths it can visit form a binary tree; three levels of
// the tree (7 nodes) are evaluated speculatively by 7 warps at once, then the path is walked with the counts.
// Returns the kept count; `outList` (shared, listCap entries) receives the kept sorted-positions.
constexpr int kSscBitmapWords = 384;  // per-warp shared bitmap: 12288 cells (keeps the CTA at 3 per SM)
constexpr size_t kSscListSmemBudget = 16 * 1024;  // 7 kept lists of u16: in shared memory up to this size (quota <= ~550 per level)
__host__ __device__ inline size_t sscListScratchBytes(int listCap) { return ((size_t)kSscWarps * listCap * 4 + 255) & ~(size_t)255; }
template <typename IdxT, int kU>
__device__ int sscSearch(const SelArrays<IdxT>& A, SelShared& S, const unsigned* xy, const IdxT* sorted, int n, const LevelGeom& lv,
                         unsigned* bitmaps, IdxT* lists, int listCap, unsigned* globalBitmap, const IdxT** outList) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = lv.quota, cols = lv.w, rows = lv.h;
    int high = lv.sscHigh;
    int low = max(1, (int)floor(sqrt((double)n / (double)K)));
    int finalNode = -1, finalCount = 0;
    while (true) {
#ifdef GVS_SEL_PROFILE
        if (tid == 0) atomicAdd(&gSelProf[7], 1ull);
#endif
        if (low > high) break;  // (width == prevWidth cannot occur once low <= high: the bracket always shrinks)
        // node k of the 3-level tree: bracket [nl, nh], width nw; children 2k+1 (count < Kmin) and 2k+2 (count > Kmax)
        int nl[7], nh[7], nw[7];
        nl[0] = low; nh[0] = high;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            nw[k] = nl[k] + (nh[k] - nl[k]) / 2;
            if (k < 3) { nl[2 * k + 1] = nl[k]; nh[2 * k + 1] = nw[k] - 1; nl[2 * k + 2] = nw[k] + 1; nh[2 * k + 2] = nh[k]; }
        }
        auto fits = [&](int k) { return nl[k] <= nh[k] && sscBitmapCells(cols, rows, nw[k]) <= kSscBitmapWords * 32; };
        const bool rootFits = fits(0);
        if (warp < 7) {
            bool valid = true;  // all ancestors must be valid brackets as well
            for (int k = warp; valid; k = (k - 1) / 2) { valid = nl[k] <= nh[k]; if (k == 0) break; }
            int c = -1;
            if (rootFits) {
                if (valid && fits(warp))
                    c = sscPass<IdxT, kU>(xy, sorted, n, nw[warp], (2 * cols) / nw[warp] + 1, (2 * rows) / nw[warp] + 1,
                                    bitmaps + warp * kSscBitmapWords, lists + warp * listCap, listCap, lane, lv.kmax);
            } else if (warp == 0) {  // tiny width: one pass, bitmap in global scratch
                c = sscPass<IdxT, kU>(xy, sorted, n, nw[0], (2 * cols) / nw[0] + 1, (2 * rows) / nw[0] + 1, globalBitmap, lists,
                                listCap, lane, lv.kmax);
            }
            if (lane == 0) S.wA[warp] = c;
        }
        __syncthreads();
        // walk (every thread does the same walk on the shared counts)
        int k = 0;
        bool done = false;
        while (true) {
            const int c = S.wA[k];
            finalNode = k; finalCount = c;
            if (c >= lv.kmin && c <= lv.kmax) { done = true; break; }
            const int child = c < lv.kmin ? 2 * k + 1 : 2 * k + 2;
            const int cl = c < lv.kmin ? nl[k] : nw[k] + 1, ch = c < lv.kmin ? nw[k] - 1 : nh[k];
            if (cl > ch) { done = true; break; }          // the reference stops and returns the last computed pass
            if (child >= 7 || S.wA[child] < 0) { low = cl; high = ch; break; }  // not evaluated: next super-round
            k = child;
        }
        __syncthreads();
        if (done) {
            if (finalCount > lv.kmax) {
                // bracket exhausted on a truncated pass: the reference returns that pass in full (:412-416) -> redo it
                const int w = nw[finalNode];
                const int ncc = (2 * cols) / w + 1, ncr = (2 * rows) / w + 1;
                const bool fitsSmem = sscBitmapCells(cols, rows, w) <= kSscBitmapWords * 32;
                if (warp == 0) {
                    const int c = sscPass<IdxT, kU>(xy, sorted, n, w, ncc, ncr, fitsSmem ? bitmaps : globalBitmap, lists, listCap, lane,
                                              0x7fffffff);
                    if (lane == 0) S.wA[0] = c;
                }
                __syncthreads();
                finalNode = 0; finalCount = S.wA[0];
                __syncthreads();
            }
            break;
        }
    }
    *outList = finalNode >= 0 ? lists + finalNode * listCap : lists;
    return finalNode >= 0 ? finalCount : 0;
}

template <typename IdxT, bool kSmall>
__global__ void __launch_bounds__(kSelThreads)
select_kernel(const __grid_constant__ Geometry g, int nImages, const CandRec* __restrict__ recs,
              const int* __restrict__ nCand, const unsigned short* __restrict__ cellCnt, KeptKey* __restrict__ kept,
              int* __restrict__ keptCnt, unsigned* __restrict__ orderedDbg, unsigned* __restrict__ sortedDbg,
              uint8_t* __restrict__ scratch,
              size_t scratchPerCta, int* __restrict__ workCounter, int* __restrict__ status, int maxCells, int segCap,
              int nMaxGlobal, int listCap, int listsGlobal) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ SelShared S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    SelArrays<IdxT> A;
    size_t bigKeyOff = 0;
    uint8_t* myScratch = scratch + (size_t)blockIdx.x * scratchPerCta;
    {
        // shared layout: cellOff | union(sort: seg + hist ; ssc: bitmaps + lists) | [small only] xy | idx | tmp | key
        uint8_t* p = smem;
        A.cellOff = reinterpret_cast<int*>(p); p += ((maxCells + 1) * 4 + 15) & ~15;
        const size_t segBytes = (size_t)2 * 3 * segCap * 4, histBytes = kSelWarps * 256 * 4;
        // the 7 speculative kept lists live in shared memory unless the per-level quota is large (listsGlobal)
        const size_t bmBytes = (size_t)kSscWarps * kSscBitmapWords * 4;
        const size_t listBytes = listsGlobal ? 0 : (size_t)kSscWarps * listCap * sizeof(IdxT);
        A.seg = reinterpret_cast<int*>(p);
        A.hist = reinterpret_cast<int*>(p + segBytes);
        A.segCap = segCap;
        A.bitmaps = reinterpret_cast<unsigned*>(p);
        A.lists = listsGlobal ? myScratch + scratchPerCta - sscListScratchBytes(listCap) : p + bmBytes;
        A.listCap = listCap;
        const size_t unionBytes = segBytes + histBytes > bmBytes + listBytes ? segBytes + histBytes : bmBytes + listBytes;
        p += (unionBytes + 15) & ~15;
        bigKeyOff = (size_t)(p - smem);
        if (kSmall) {
            A.xy = reinterpret_cast<unsigned*>(p); p += (size_t)kSmallMax * 4;
            A.idx = reinterpret_cast<IdxT*>(p); p += (size_t)kSmallMax * sizeof(IdxT);
            A.tmp = reinterpret_cast<IdxT*>(p); p += (size_t)kSmallMax * sizeof(IdxT);
            A.key = p;
            A.sxy = nullptr;
            A.globalBitmap = reinterpret_cast<unsigned*>(myScratch);
        } else {
            // global class: all per-candidate arrays and the segment frontiers live in global scratch
            uint8_t* q = myScratch;
            A.xy = reinterpret_cast<unsigned*>(q); q += (size_t)nMaxGlobal * 4;
            A.idx = reinterpret_cast<IdxT*>(q); q += (size_t)nMaxGlobal * sizeof(IdxT);
            A.tmp = reinterpret_cast<IdxT*>(q); q += (size_t)nMaxGlobal * sizeof(IdxT);
            A.seg = reinterpret_cast<int*>(q); A.segCap = nMaxGlobal / 17 + 2; q += (size_t)2 * 3 * A.segCap * 4;
            A.key = q; q += ((size_t)nMaxGlobal + 15) & ~(size_t)15;   // (problems above kBigKeySmem candidates; smaller ones: shared, below)
            A.sxy = reinterpret_cast<unsigned*>(q); q += (size_t)nMaxGlobal * 4;
            A.globalBitmap = reinterpret_cast<unsigned*>(q);
            A.hist = reinterpret_cast<int*>(A.bitmaps);  // the union holds only hist during the sort
        }
    }

    uint8_t* const keyGlobal = A.key;
    uint8_t* const keyShared = kSmall ? A.key : smem + bigKeyOff;   // (global class: behind cellOff and the union)
    const int nProb = nImages * g.nLevels;
    while (true) {
        __syncthreads();
        if (tid == 0) S.prob = atomicAdd(workCounter, 1);
        __syncthreads();
        const int prob = S.prob;
        if (prob >= nProb) break;
        const int level = prob / nImages, img = prob - level * nImages;
        const LevelGeom& lv = g.lv[level];
        int n = nCand[img * kMaxLevels + level];
        if (n > lv.candCap) n = lv.candCap;  // overflow already flagged by the FAST kernel
        if ((n <= kSmallMax) != kSmall) continue;
        if (!kSmall) A.key = n <= kBigKeySmem ? keyShared : keyGlobal;   // the partition passes mostly read keys
        int* keptCntOut = keptCnt + img * kMaxLevels + level;
        KeptKey* keptOut = kept + (size_t)img * g.keptPerImage + lv.keptOff;
        if (n == 0) { if (tid == 0) *keptCntOut = 0; continue; }
        GVS_PROF_DECL

        // ---- (a) order: scan cell counts, scatter records
        const int nCells = lv.nCols * lv.nRows;
        const unsigned short* cc = cellCnt + (size_t)img * g.cellsPerImage + lv.cellBase;
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nCells; base += 32) {
                const int i = base + lane;
                const int c = i < nCells ? (int)cc[i] : 0;
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (i < nCells) A.cellOff[i] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        const CandRec* rin = recs + (size_t)img * g.candPerImage + lv.candOff;
        for (int i = tid; i < n; i += kSelThreads) {
            const CandRec r = rin[i];
            const int pos = A.cellOff[r.w1 & 0xffff] + (int)(r.w1 >> 16);
            if (pos < n) { A.xy[pos] = r.w0; A.key[pos] = (uint8_t)(r.w0 >> 24); }
        }
        __syncthreads();
        if (orderedDbg) {
            unsigned* o = orderedDbg + (size_t)img * g.candPerImage + lv.candOff;
            for (int i = tid; i < n; i += kSelThreads) o[i] = A.xy[i];
        }
        if (n <= lv.quota) {  // no ssc (src/FeatureExtractor.cpp:610): keys stay in emission order
            for (int i = tid; i < n; i += kSelThreads) keptOut[i] = A.xy[i];
            if (tid == 0) *keptCntOut = n;
            continue;
        }

        // ---- (b) sort
        GVS_PROF(0);
        sortDescending(A, S, n);
        const IdxT* sorted = A.tmp;
        if (sortedDbg) {  // parity tap: the permutation itself and, after ssc, the kept positions
            unsigned* o = sortedDbg + (size_t)img * g.candPerImage + lv.candOff;
            for (int i = tid; i < n; i += kSelThreads) o[i] = (unsigned)sorted[i];
        }

        // ---- (c) ssc binary search (src/FeatureExtractor.cpp:396-461)
        const IdxT* res = nullptr;
        GVS_PROF_DECL2
        if (!kSmall) {   // coordinates in sorted order: the passes stream them (one gather here instead of one per pass and point)
            for (int i = tid; i < n; i += kSelThreads) A.sxy[i] = A.xy[sorted[i]];
            __syncthreads();
        }
        int count = sscSearch<IdxT, kSmall ? 1 : 4>(A, S, kSmall ? A.xy : A.sxy, kSmall ? sorted : nullptr, n, lv, A.bitmaps,
                                                    reinterpret_cast<IdxT*>(A.lists), A.listCap, A.globalBitmap, &res);
        GVS_PROF(4);
        if (count > lv.keptCap || count > A.listCap) { if (tid == 0) atomicOr(status, kStatKeptOverflow); count = min(lv.keptCap, A.listCap); }
        for (int i = tid; i < count; i += kSelThreads) keptOut[i] = kSmall ? A.xy[sorted[res[i]]] : A.sxy[res[i]];
        if (tid == 0) *keptCntOut = count;
        GVS_PROF(5);
#ifdef GVS_SEL_PROFILE
        if (tid == 0) atomicAdd(&gSelProf[6], 1ull);
#endif
    }
}

static size_t selectSmemBytes(bool small, int maxCells, int segCap, int listCap /* 0: lists in global scratch */) {
    size_t b = ((size_t)(maxCells + 1) * 4 + 15) & ~(size_t)15;
    const size_t segBytes = (size_t)2 * 3 * segCap * 4, histBytes = kSelWarps * 256 * 4;
    size_t u = std::max(segBytes + histBytes, (size_t)kSscWarps * kSscBitmapWords * 4 + (size_t)kSscWarps * listCap * (small ? 2 : 4));
    b += (u + 15) & ~(size_t)15;
    if (small) b += (size_t)kSmallMax * (4 + 2 + 2 + 1);
    else b += kBigKeySmem;
    return b;
}

static int maxListCap(const Geometry& g) {
    int m = 0;
    for (int l = 0; l < g.nLevels; l++) m = std::max(m, g.lv[l].keptCap);
    return m;
}

static int maxCandCap(const Geometry& g) {
    int m = 0;
    for (int l = 0; l < g.nLevels; l++) m = std::max(m, g.lv[l].candCap);
    return m;
}

size_t selectScratchBytes(const Geometry& g) {
    // global class arrays (xy, idx, tmp, seg frontiers, key) + the fallback SSC bitmap for width 1:
    // (2h+1) x (2w+1) cells, one bit each
    const size_t nMax = (size_t)std::max(maxCandCap(g), kSmallMax);
    size_t b = nMax * (4 + 4 + 4 + 4) + (size_t)2 * 3 * (nMax / 17 + 2) * 4 + ((nMax + 15) & ~(size_t)15);   // xy, idx, tmp, sxy, seg, key
    b += ((size_t)sscBitmapCells(g.lv[0].w, g.lv[0].h, 1) + 31) / 32 * 4 + 16;
    b = (b + 255) & ~(size_t)255;
    return b + sscListScratchBytes(maxListCap(g));   // tail: the SSC kept lists when they do not fit shared memory
}

int selectMaxCtas(int device) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int per = 3;
    if (const char* e = getenv("GVS_SELECT_CTAS_PER_SM")) per = std::max(1, atoi(e));  // tuning knob (default 3 = shared-memory limit)
    return (sms > 0 ? sms : 148) * per;
}

void launchSelect(const Geometry& g, int nImages, const CandRec* recs, const int* nCand, const unsigned short* cellCnt,
                  KeptKey* kept, int* keptCnt, unsigned* orderedDbg, unsigned* sortedDbg, uint8_t* scratch,
                  size_t scratchPerCta, int nCtas, int* workCounter, int* status, cudaStream_t s, int* launches) {
    int maxCells = 0;
    for (int l = 0; l < g.nLevels; l++) maxCells = std::max(maxCells, g.lv[l].nCols * g.lv[l].nRows);
    const int segCapSmall = kSmallMax / 17 + 2;
    const int listCap = maxListCap(g);
    const int listsGlobal = (size_t)kSscWarps * listCap * 2 > kSscListSmemBudget;
    const size_t smSmall = selectSmemBytes(true, maxCells, segCapSmall, listsGlobal ? 0 : listCap);
    const size_t smBig = selectSmemBytes(false, maxCells, 1, listsGlobal ? 0 : listCap);
    static size_t cfgS[64], cfgB[64];
    ensureDynamicSmem(select_kernel<unsigned short, true>, smSmall, cfgS);
    ensureDynamicSmem(select_kernel<unsigned, false>, smBig, cfgB);
    const int nProb = nImages * g.nLevels;
    const int grid = std::min(nCtas, nProb);
    cudaMemsetAsync(workCounter, 0, 2 * sizeof(int), s);
    select_kernel<unsigned short, true><<<grid, kSelThreads, smSmall, s>>>(
        g, nImages, recs, nCand, cellCnt, kept, keptCnt, orderedDbg, sortedDbg, scratch, scratchPerCta, workCounter, status,
        maxCells, segCapSmall, maxCandCap(g), listCap, listsGlobal);
    ++*launches;
#ifdef GVS_SEL_PROFILE
    {
        static int calls = 0;
        if (++calls % 40 == 0) {
            cudaStreamSynchronize(s);
            unsigned long long h[8];
            cudaMemcpyFromSymbol(h, gSelProf, sizeof(h));
            const double tot = (double)(h[0] + h[1] + h[2] + h[3] + h[4] + h[5]);
            fprintf(stderr, "[select profile] problems %llu  order %.1f%%  sort-cta %.1f%%  sort-warp %.1f%%  sort-final %.1f%%  ssc %.1f%%  out %.1f%%  cycles/problem %.0f  ssc rounds/problem %.2f  ssc cycles/round %.0f\n",
                    h[6], 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, 100 * h[5] / tot, tot / (double)h[6],
                    h[7] / (double)h[6], (double)h[4] / (double)h[7]);
        }
    }
#endif
    bool needBig = false;
    for (int l = 0; l < g.nLevels; l++) needBig |= g.lv[l].candCap > kSmallMax;
    if (needBig) {
        select_kernel<unsigned, false><<<grid, kSelThreads, smBig, s>>>(
            g, nImages, recs, nCand, cellCnt, kept, keptCnt, orderedDbg, sortedDbg, scratch, scratchPerCta, workCounter + 1,
            status, maxCells, 1, maxCandCap(g), listCap, listsGlobal);
        ++*launches;
    }
}

// exclusive scan of per-image totals -> dense output offsets (one CTA; nImages is small)
__global__ void offsets_kernel(int nLevels, int nImages, const int* __restrict__ keptCnt, int* __restrict__ imgOff,
                               int totalCap, int* __restrict__ status) {
    __shared__ int sRun;
    __shared__ int sW[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sRun = 0;
    __syncthreads();
    for (int base = 0; base < nImages; base += blockDim.x) {
        const int i = base + tid;
        int c = 0;
        if (i < nImages)
            for (int l = 0; l < nLevels; l++) c += keptCnt[i * kMaxLevels + l];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) sW[warp] = incl;
        __syncthreads();
        int add = sRun;
        for (int w = 0; w < warp; w++) add += sW[w];
        if (i < nImages) imgOff[i] = add + incl - c;
        __syncthreads();
        if (tid == blockDim.x - 1) sRun = add + incl;
        __syncthreads();
    }
    if (tid == 0) {
        imgOff[nImages] = sRun;
        if (sRun > totalCap) atomicOr(status, kStatOutputOverflow);
    }
}

void launchOffsets(const Geometry& g, int nImages, const int* keptCnt, int* imgOff, int totalCap, int* status,
                   cudaStream_t s, int* launches) {
    offsets_kernel<<<1, 256, 0, s>>>(g.nLevels, nImages, keptCnt, imgOff, totalCap, status);
    ++*launches;
}

}  // namespace gvs
