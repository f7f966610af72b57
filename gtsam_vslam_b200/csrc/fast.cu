// K2 fast_cells: FAST-9/16 corner score, cell-adaptive threshold and cell-local non-max suppression for every
// pyramid level of every image in one launch.  Replaces the ~900 per-cell cv::FAST calls (+ threshold-7
// fallback) of FeatureExtractor::computeKeypointsORBNew (reference src/FeatureExtractor.cpp:547-608).
//
// Dense reformulation, bit and order identical to the per-cell calls:
//   best(p)  = max over the 16 arcs of 9 consecutive ring pixels of max(min(ring) - I(p), I(p) - max(ring))
//   corner at threshold t  <=>  best > t ;  score = best - 1 (independent of t)
//   cell (iR,iC) owns x in [19+iC*gW, min(19+(iC+1)*gW, w-19)), y likewise; neighbours outside the cell's own
//   range count as 0 in the 3x3 non-max test (each cell was a separate sub-image in the reference).
//   A pixel survives NMS at threshold T iff score >= T and it is a strict local maximum of the raw scores
//   inside the cell (a neighbour below T is below the pixel anyway), so one local-max pass serves both
//   thresholds: T_cell = maxFast if any local max has score >= maxFast, else minFast  (the reference falls
//   back when the first call returns no keypoint *after* NMS, :592-593).
//
// Work item = up to kFastCellsPerItem consecutive cells of one cell row (about 150 x 40 px); CTA = 256 threads.
//   1. the (rows+6) x (cols+6) u8 tile is staged in shared memory with 4-byte loads;
//   Every warp then owns a contiguous share of the tile's 4-pixel unit words and private queue regions for it (no shared
//   counters, no atomics, no CTA barrier between 2a and 2b):
//   2a. a pre-test on 4 pixels per instruction (VABSDIFF4 on the up/down and left/right ring pairs) discards pixels that
//      cannot be corners at the pass threshold; surviving pixel pairs are compacted into the warp's queues with ballots;
//   2b. exact scores of the surviving pairs in u16x2 lanes: 21 word loads, ring pairs extracted with PRMT
//      (lane = 257 * pixel), 9-wide circular min/max as min3 of min3 with the native VIMNMX3.U16x2
//      (78 instructions per pair for both polarities);
//   3. the unit words with a non-zero score are compacted the same way; strict local maxima, four pixels at a time (vcmpgtu4
//      against the 8 shifted neighbour words, masked at cell borders), are appended to a per-cell list in shared memory, ranked
//      in row-major order inside the cell and written as 8-byte records {x, y, score, cell, rank} to a per-(image, level)
//      buffer; K3/K4 (select.cu) restores the reference's emission order from (cell, rank).
//   Pass 0 runs 2a-3 at maxFastThreshold on every cell; pass 1 redoes the cells that listed no key at minFastThreshold.
// Bound: integer ALU, not HBM: see DESIGN.md.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace gvs {

// ring offsets, OpenCV order
#define GVS_RING_DX {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}
#define GVS_RING_DY {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3}

constexpr int kFastWarps = 8;  // warps per CTA (256 threads)

struct FastSmemLayout {
    int tileWords, tileRows;    // image tile: words per row, rows
    int scoreWords, scoreRows;  // score tile: words per row (1 zero word each side), rows (1 zero row each side)
    int listCap;                // entries per cell list
    int tileOff, scoreOff, listOff, maskOff, queueOff, queueCap, total;
};

// tma: the tile is filled by TMA box loads (row pitch a multiple of 16 bytes, buffer 128-byte aligned).  ONE buffer: a second one
// costs a CTA per SM (measured: 2.11 ms against 1.91 with one), so the next strip's tile is requested when the current strip has
// read its tile for the last time -- it arrives while the strip finds its local maxima and writes its records
// maxSW: width in pixels of the widest strip (cells of a strip x cell pitch of its level), maxGW / maxGH: largest cell pitch
__host__ __device__ inline FastSmemLayout fastLayout(int maxSW, int maxGW, int maxGH, bool tma = false) {
    FastSmemLayout L;
    const int maxUnits = (maxSW + 3 + 3) / 4 + 1;  // unit words covering [X0, X1)
    // (a TMA box starts at a 16-byte aligned x -- the copy faults otherwise -- so up to 3 words of lead-in are fetched with the tile)
    // and the row pitch is an odd multiple of 16 bytes: corners line up vertically, and a pitch of 0 mod 32 words would put the same
    // column of every row into one bank
    L.tileWords = tma ? ((maxUnits + 2 + 3 + 3) & ~3) | 4 : maxUnits + 2;
    L.tileRows = maxGH + 6;
    L.scoreWords = maxUnits + 2;
    L.scoreRows = maxGH + 2;
    L.listCap = ((maxGW + 1) / 2) * ((maxGH + 1) / 2);
    L.tileOff = 0;
    const int tileBytes = tma ? (L.tileWords * L.tileRows * 4 + 127) & ~127 : L.tileWords * L.tileRows * 4;
    L.scoreOff = L.tileOff + tileBytes;
    L.maskOff = L.scoreOff + L.scoreWords * L.scoreRows * 4;
    L.listOff = ((L.maskOff + maxUnits * 12 + 15) / 16) * 16;
    L.queueOff = L.listOff + kFastCellsPerItem * L.listCap * 4;
    // per pair parity; every warp owns a private region of `chunk` entries (chunk = its share of the unit words rounded up to 32)
    L.queueCap = maxUnits * maxGH + kFastWarps * 32 + kFastWarps;
    L.total = L.queueOff + 2 * L.queueCap * 2;
    return L;
}

// index of the cell (inside the strip) that owns strip column lx
__device__ __forceinline__ int cellOfStrip(int lx, int gW) {
    int c = 0;
#pragma unroll
    for (int k = 1; k < kFastCellsPerItem; k++) c += lx >= k * gW;
    return c;
}

// u16x2 lanes = 257 * (byte q, byte q+1) of the 12-byte window {w0, w1, w2}; q is a compile-time constant
template <int Q>
__device__ __forceinline__ unsigned pairAt(unsigned w0, unsigned w1, unsigned w2) {
    if (Q + 1 <= 7) return __byte_perm(w0, w1, Q | (Q << 4) | ((Q + 1) << 8) | ((Q + 1) << 12));
    return __byte_perm(w1, w2, (Q - 4) | ((Q - 4) << 4) | ((Q - 3) << 8) | ((Q - 3) << 12));
}

// B = max over the 16 arcs of min of 9 consecutive, D = min over arcs of max of 9 consecutive (u16x2 lanes)
__device__ __forceinline__ void arcMinMax(const unsigned (&r)[16], unsigned& B, unsigned& D) {
    unsigned mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mx3[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    unsigned mn9[16], mx9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn9[k] = __vimin3_u16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
        mx9[k] = __vimax3_u16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
    }
    unsigned b5[6], d5[6];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        b5[k] = __vimax3_u16x2(mn9[3 * k], mn9[3 * k + 1], mn9[3 * k + 2]);
        d5[k] = __vimin3_u16x2(mx9[3 * k], mx9[3 * k + 1], mx9[3 * k + 2]);
    }
    b5[5] = mn9[15]; d5[5] = mx9[15];
    B = __vmaxu2(__vimax3_u16x2(b5[0], b5[1], b5[2]), __vimax3_u16x2(b5[3], b5[4], b5[5]));
    D = __vminu2(__vimin3_u16x2(d5[0], d5[1], d5[2]), __vimin3_u16x2(d5[3], d5[4], d5[5]));
}

template <int P>  // pair P of the unit word: pixels 2P, 2P+1
__device__ __forceinline__ unsigned scorePair(const unsigned (&W)[7][3], int tMin) {
    constexpr int dx[16] = GVS_RING_DX, dy[16] = GVS_RING_DY;
    unsigned r[16];
#define GVS_R(k) r[k] = pairAt<4 + 2 * P + dx[k]>(W[3 + dy[k]][0], W[3 + dy[k]][1], W[3 + dy[k]][2]);
    GVS_R(0) GVS_R(1) GVS_R(2) GVS_R(3) GVS_R(4) GVS_R(5) GVS_R(6) GVS_R(7)
    GVS_R(8) GVS_R(9) GVS_R(10) GVS_R(11) GVS_R(12) GVS_R(13) GVS_R(14) GVS_R(15)
#undef GVS_R
    unsigned B, D;
    arcMinMax(r, B, D);
    const unsigned c = W[3][1];
    const int v0 = (c >> (16 * P)) & 0xff, v1 = (c >> (16 * P + 8)) & 0xff;
    const int best0 = max((int)((B >> 8) & 0xff) - v0, v0 - (int)((D >> 8) & 0xff));
    const int best1 = max((int)(B >> 24) - v1, v1 - (int)(D >> 24));
    const unsigned s0 = best0 > tMin ? (unsigned)(best0 - 1) : 0u, s1 = best1 > tMin ? (unsigned)(best1 - 1) : 0u;
    return s0 | (s1 << 8);
}

#ifndef GVS_FAST_CTAS
#define GVS_FAST_CTAS 3
#endif

// geometry of one work item (a strip of cells of one cell row)
struct FastStrip {
    int level, y0, y1, X0, X1, xu, nW, rows, cols, gW;
};
__device__ __forceinline__ FastStrip fastStrip(const Geometry& g, const FastItem& it) {
    FastStrip s;
    const LevelGeom& lv = g.lv[it.level];
    const int edge = g.edge;  // (edgeThreshold - fastEdge) + 3: first tested pixel of the first cell
    s.level = it.level;
    s.y0 = edge + it.iR * lv.gridH; s.y1 = min(s.y0 + lv.gridH, lv.h - edge);
    s.X0 = edge + it.iC0 * lv.gridW; s.X1 = min(s.X0 + it.nC * lv.gridW, lv.w - edge);
    s.rows = s.y1 - s.y0; s.cols = s.X1 - s.X0;
    s.gW = lv.gridW;
    // unit words: 4-px aligned words covering [X0, X1); tile word 0 is the word left of unit 0
    s.xu = s.X0 & ~3;                 // x of unit 0
    // number of unit words; at least 2: i / 1 has no 32-bit magic number, and a narrow last cell can leave a single word (the
    // extra word lies inside the row -- X1 <= w - edge -- and is fully masked out)
    s.nW = max((s.X1 - s.xu + 3) >> 2, 2);
    return s;
}

// USE_TMA: a CTA takes itemsPerCta consecutive strips of one image; the (rows+6) x (cols+8) tile of strip k+1 is fetched by one TMA
// box load (cp.async.bulk.tensor) into the second tile buffer while strip k is processed.  Otherwise (no tensor-map encoder, or a
// strip wider than a TMA box) every thread stages its share of the tile with 4-byte loads.
struct FastMaps { CUtensorMap m[kMaxLevels]; };

template <bool USE_TMA>
__global__ void __launch_bounds__(256, GVS_FAST_CTAS)
fast_cells_kernel(const __grid_constant__ Geometry g, const __grid_constant__ FastMaps maps /* TMA tile map per level */, const uint8_t* __restrict__ slab,
                  const FastItem* __restrict__ items, int nItems, int itemsPerCta, int maxSW, int maxGW, int maxGH, CandRec* __restrict__ recs,
                  int* __restrict__ nCand, unsigned short* __restrict__ cellCnt, int* __restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int sCnt[kFastCellsPerItem], sBase[kFastCellsPerItem];
    __shared__ __align__(8) unsigned long long sBar;

    const int img = blockIdx.y;
    const FastSmemLayout L = fastLayout(maxSW, maxGW, maxGH, USE_TMA);
    unsigned* score = reinterpret_cast<unsigned*>(smem + L.scoreOff);
    unsigned* masks = reinterpret_cast<unsigned*>(smem + L.maskOff);  // per unit: in-range, has-left, has-right byte masks
    unsigned* lists = reinterpret_cast<unsigned*>(smem + L.listOff);
    const int tid = threadIdx.x;
    const int first = blockIdx.x * itemsPerCta, last = min(first + itemsPerCta, nItems);
    // one TMA box load of the tile of strip `idx` (x from xu - 4, y from y0 - 3); issued by thread 0
    auto issue = [&](int idx) {
        const FastStrip n = fastStrip(g, items[idx]);
        if (n.rows <= 0 || n.cols <= 0) return;
        const unsigned bar = smemAddr(&sBar);
        mbarExpectTx(bar, (unsigned)(L.tileWords * 4 * L.tileRows));
        // the maps have 32-bit elements (a box may be 256 of them wide); x must be a multiple of 16 bytes (tools/tma_probe.cu), so
        // the box starts up to 3 words left of the strip's tile
        tmaLoadTile3d(smemAddr(smem + L.tileOff), &maps.m[n.level], ((n.xu - 4) & ~15) >> 2, n.y0 - 3, img, bar);
    };
    if (USE_TMA) {
        if (tid == 0) {
            mbarInit(smemAddr(&sBar), 1);
            if (first < last) issue(first);
        }
        __syncthreads();   // (the barrier is initialised before anybody polls it)
    }
    unsigned used = 0;   // tile loads consumed so far (the mbarrier's phase)
    for (int idx = first; idx < last; idx++) {
    const FastItem it = items[idx];
    const LevelGeom& lv = g.lv[it.level];
    const FastStrip S = fastStrip(g, it);
    const int y0 = S.y0, X0 = S.X0, rows = S.rows, cols = S.cols, gW = S.gW, xu = S.xu, nW = S.nW;
    if (tid < kFastCellsPerItem) sCnt[tid] = 0;
    if (rows <= 0 || cols <= 0) {  // empty strip: still publish zero counts for its cells
        if (tid < it.nC) cellCnt[(size_t)img * g.cellsPerImage + lv.cellBase + it.iR * lv.nCols + it.iC0 + tid] = 0;
        if (USE_TMA && tid == 0 && idx + 1 < last) issue(idx + 1);   // (nothing of this strip is in the tile buffer)
        __syncthreads();
        continue;
    }
    unsigned* tile = reinterpret_cast<unsigned*>(smem + L.tileOff) + (USE_TMA ? ((xu - 4) & 15) >> 2 : 0);
    const int tw = nW + 2, th = rows + 6;   // tile words per row / rows
    // i / d = umulhi(i, 2^32 / d + 1) for the small i used here (d >= 2)
    const unsigned magicT = 0xffffffffu / (unsigned)tw + 1u, magicU = 0xffffffffu / (unsigned)nW + 1u;

    // ---- 1. stage the image tile (x from xu-4, y from y0-3), zero the score tile, build the per-unit byte masks
    if (!USE_TMA) {
        const uint8_t* src = levelPtr(slab, g, img, it.level) + (size_t)(y0 - 3) * lv.pitch + (xu - 4);
        for (int i = tid; i < tw * th; i += 256) {
            const int r = (int)__umulhi((unsigned)i, magicT), c = i - r * tw;
            tile[r * L.tileWords + c] = __ldg(reinterpret_cast<const unsigned*>(src + (size_t)r * lv.pitch) + c);
        }
    }
    for (int i = tid; i < L.scoreWords * (rows + 2); i += 256) score[i] = 0;
    for (int j = tid; j < nW; j += 256) {
        unsigned inr = 0, hasL = 0, hasR = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int lx = xu + 4 * j + b - X0;  // position inside the strip
            if (lx < 0 || lx >= cols) continue;
            const int cl = cellOfStrip(lx, gW);
            const int cx0 = cl * gW, cx1 = min(cx0 + gW, cols);
            inr |= 0xffu << (8 * b);
            if (lx > cx0) hasL |= 0xffu << (8 * b);
            if (lx + 1 < cx1) hasR |= 0xffu << (8 * b);
        }
        masks[3 * j] = inr; masks[3 * j + 1] = hasL; masks[3 * j + 2] = hasR;
    }
    if (USE_TMA) {   // one thread polls the tile's mbarrier (a polling warp takes issue slots from the CTAs that compute); the CTA
        if (tid == 0) mbarWait(smemAddr(&sBar), used & 1u);   // barrier below passes the arrival on
        used++;
    }
    __syncthreads();

    // The reference runs cv::FAST on a cell at maxFastThreshold and, only if that returns nothing, again at minFastThreshold
    // (src/FeatureExtractor.cpp:591-593).  Same here: pass 0 works at tMax on every cell of the strip; pass 1 (rare on textured
    // images) redoes the cells that produced no key at tMin.  A pass = 2a pre-test + compaction, 2b exact scores, 3 local maxima.
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem + L.queueOff);
    unsigned short* score16 = reinterpret_cast<unsigned short*>(score);
    const int tMin = g.tMin, tMax = g.tMax;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const int nUnits = nW * rows;
    // Every warp owns a share of the unit words (at most `chunk`) and private queue regions for it, so that compaction needs no
    // shared counters or atomics and phase 2b follows 2a without a CTA barrier.
    const int chunk = (((nUnits + kFastWarps - 1) / kFastWarps) + 31) & ~31;
    // Ownership is interleaved -- warp w takes the 32-unit groups w, w + 8, w + 16, ... -- so that corner-dense parts of the strip are
    // spread over all warps (the barriers between the phases wait for the slowest warp).
    const int u0 = warp * 32, u1 = nUnits;
    constexpr int uStep = kFastWarps * 32;
    unsigned short* q0 = queue + warp * chunk;                 // pair 0 of a unit / phase 3: units with a non-zero score
    unsigned short* q1 = queue + L.queueCap + warp * chunk;    // pair 1 of a unit
    unsigned redo = 0;  // pass 1: bit c = cell c of the strip is redone
    // bytes of unit word j that lie in a cell that is being redone
    auto redoBytes = [&](int j) {
        unsigned m = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int lx = xu + 4 * j + b - X0;
            if (lx >= 0 && lx < cols && ((redo >> cellOfStrip(lx, gW)) & 1u)) m |= 0xffu << (8 * b);
        }
        return m;
    };
    for (int pass = 0; pass < 2; pass++) {
        const int T = pass == 0 ? tMax : tMin;

        // ---- 2a. pre-test on 4 pixels at once (native VABSDIFF4): a 9-arc contains ring pixel k or k+8 for every k, so a
        // corner at threshold T needs |I(ring_k) - I(p)| > T for k in {0 or 8} and for k in {4 or 12}.  Per byte,
        // d > T  <=>  bit 7 of ((d & 0x7f) + (127 - T)) | d   (T <= 127; above that the test degrades to d >= 128, still
        // necessary).  Surviving pixel pairs are compacted into the warp's two queues (one per pair position: 2b has no
        // divergent code).
        const unsigned kT = (unsigned)max(127 - T, 0) * 0x01010101u;
        auto gt = [&](unsigned d) { return ((d & 0x7f7f7f7fu) + kT) | d; };
        int n0 = 0, n1 = 0;
        for (int base = u0; base < u1; base += uStep) {
            const int i = base + lane;
            unsigned f = 0;  // bit 7 of byte b: pixel b of the unit passes the pre-test
            if (i < u1) {
                const int ry = (int)__umulhi((unsigned)i, magicU), j = i - ry * nW;
                const unsigned* p = tile + (ry + 3) * L.tileWords + j;
                const unsigned c = p[1];
                const unsigned up = p[1 - 3 * L.tileWords], dn = p[1 + 3 * L.tileWords];
                const unsigned lf = __funnelshift_r(p[0], c, 8), rt = __funnelshift_r(c, p[2], 24);
                const unsigned v = gt(__vabsdiffu4(up, c)) | gt(__vabsdiffu4(dn, c));
                const unsigned h = gt(__vabsdiffu4(lf, c)) | gt(__vabsdiffu4(rt, c));
                f = v & h & 0x80808080u & masks[3 * j];
                if (pass) f &= redoBytes(j);
            }
            const bool p0 = (f & 0x0000ffffu) != 0, p1 = (f & 0xffff0000u) != 0;
            const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
            if (p0) q0[n0 + __popc(b0 & ltMask)] = (unsigned short)i;
            if (p1) q1[n1 + __popc(b1 & ltMask)] = (unsigned short)i;
            n0 += __popc(b0); n1 += __popc(b1);
        }
        __syncwarp();

        // ---- 2b. exact scores of the warp's surviving pairs: 21 word loads, ring pairs by PRMT, min3-of-min3 arcs; a score
        // below the pass threshold is stored as 0, exactly like the score buffer of cv::FAST at that threshold
        for (int q = lane; q < n0 + n1; q += 32) {
            const bool second = q >= n0;
            const int i = second ? q1[q - n0] : q0[q];
            const int ry = (int)__umulhi((unsigned)i, magicU), j = i - ry * nW;
            unsigned W[7][3];
#pragma unroll
            for (int r = 0; r < 7; r++) {
                const unsigned* p = tile + (ry + r) * L.tileWords + j;
                W[r][0] = p[0]; W[r][1] = p[1]; W[r][2] = p[2];
            }
            const unsigned m = masks[3 * j];
            if (!second) score16[2 * ((ry + 1) * L.scoreWords + j + 1)] = (unsigned short)(scorePair<0>(W, T) & m);
            else score16[2 * ((ry + 1) * L.scoreWords + j + 1) + 1] = (unsigned short)(scorePair<1>(W, T) & (m >> 16));
        }
        __syncthreads();

        // ---- 3. strict local maxima inside each cell, four pixels at a time -> per-cell lists.  The warp first compacts the
        // unit words of its share that hold a non-zero score (about a third on textured images) and tests only those.
        int nz = 0;
        for (int base = u0; base < u1; base += uStep) {
            const int i = base + lane;
            bool any = false;
            if (i < u1) {
                const int ry = (int)__umulhi((unsigned)i, magicU), j = i - ry * nW;
                unsigned c = score[(ry + 1) * L.scoreWords + j + 1];
                if (pass) c &= redoBytes(j);   // the other cells were listed by pass 0
                any = c != 0;
            }
            const unsigned b = __ballot_sync(0xffffffffu, any);
            if (any) q0[nz + __popc(b & ltMask)] = (unsigned short)i;
            nz += __popc(b);
        }
        __syncwarp();
        for (int q = lane; q < nz; q += 32) {
            const int i = q0[q];
            const int ry = (int)__umulhi((unsigned)i, magicU), j = i - ry * nW;
            const unsigned* sp = score + (ry + 1) * L.scoreWords + j + 1;
            const unsigned c = sp[0];
            const unsigned hasL = masks[3 * j + 1], hasR = masks[3 * j + 2];
            const int P = L.scoreWords;
            const unsigned up = sp[-P], dn = sp[P];
            unsigned lm = __vcmpgtu4(c, up) & __vcmpgtu4(c, dn);
            const unsigned l0 = __funnelshift_r(sp[-1], c, 24) & hasL, l1 = __funnelshift_r(sp[-P - 1], up, 24) & hasL,
                           l2 = __funnelshift_r(sp[P - 1], dn, 24) & hasL;
            lm &= __vcmpgtu4(c, l0) & __vcmpgtu4(c, l1) & __vcmpgtu4(c, l2);
            const unsigned r0 = __funnelshift_r(c, sp[1], 8) & hasR, r1 = __funnelshift_r(up, sp[-P + 1], 8) & hasR,
                           r2 = __funnelshift_r(dn, sp[P + 1], 8) & hasR;
            lm &= __vcmpgtu4(c, r0) & __vcmpgtu4(c, r1) & __vcmpgtu4(c, r2);
            if (pass) lm &= redoBytes(j);   // the other cells were listed by pass 0
            while (lm) {
                const int b = (__ffs(lm) - 1) >> 3;
                lm &= ~(0xffu << (8 * b));
                const int lx = xu + 4 * j + b - X0, s = (c >> (8 * b)) & 0xff;
                const int cl = cellOfStrip(lx, gW);
                const int slot = atomicAdd(&sCnt[cl], 1);
                lists[cl * L.listCap + slot] = ((unsigned)ry << 16) | ((unsigned)(lx - cl * gW) << 8) | (unsigned)s;
            }
        }
        __syncthreads();
        if (pass || tMin >= tMax) break;
        for (int c = 0; c < it.nC; c++) redo |= (sCnt[c] == 0 ? 1u : 0u) << c;   // same value in every thread
        if (!redo) break;
    }
    // the tile has been read for the last time (every warp is past its exact scores): request the next strip's tile into the same
    // buffer; it arrives while this strip writes its records
    if (USE_TMA && tid == 0 && idx + 1 < last) issue(idx + 1);

    // ---- 4. reserve, rank inside the cell (row-major), write: every listed key is emitted (the lists already hold exactly the
    // keys of the threshold that applies to their cell)
    int total = 0;
    for (int c = 0; c < it.nC; c++) total += sCnt[c];
    if (tid == 0) {
        int tot = 0;
        for (int c = 0; c < it.nC; c++) { sBase[c] = tot; tot += sCnt[c]; }
        const int base = tot ? atomicAdd(&nCand[img * kMaxLevels + it.level], tot) : 0;
        if (base + tot > lv.candCap) atomicOr(status, kStatCandOverflow);
        for (int c = 0; c < it.nC; c++) sBase[c] += base;
    }
    if (tid < it.nC)
        cellCnt[(size_t)img * g.cellsPerImage + lv.cellBase + it.iR * lv.nCols + it.iC0 + tid] = (unsigned short)sCnt[tid];
    __syncthreads();
    CandRec* out = recs + (size_t)img * g.candPerImage + lv.candOff;
    for (int e = tid; e < total; e += 256) {
        int c = 0, off = e;
        while (off >= sCnt[c]) { off -= sCnt[c]; c++; }
        const unsigned* lc = lists + c * L.listCap;
        const unsigned key = lc[off];
        int rank = 0;
        const int n = sCnt[c];
        for (int j = 0; j < n; j++) rank += (lc[j] >> 8) < (key >> 8);
        const int slot = sBase[c] + rank;
        if (slot >= lv.candCap) continue;
        const int x = X0 + c * gW + (int)((key >> 8) & 0xff), y = y0 + (int)(key >> 16);
        CandRec rec;
        rec.w0 = (unsigned)x | ((unsigned)y << 12) | ((key & 0xff) << 24);
        rec.w1 = (unsigned)(it.iR * lv.nCols + it.iC0 + c) | ((unsigned)rank << 16);
        out[slot] = rec;
    }
    if (idx + 1 < last) __syncthreads();   // the next strip reuses the score tile, the masks, the lists and the counters
    }   // strips of this CTA
}

// box of the TMA tile loads (32-bit words x rows), or false when a strip does not fit a TMA box (256 x 256 elements)
struct FastDims { int maxSW, maxGW, maxGH; };
static FastDims fastDims(const Geometry& g) {
    FastDims d{0, 0, 0};
    for (int l = 0; l < g.nLevels; l++) {
        const LevelGeom& lv = g.lv[l];
        d.maxGW = std::max(d.maxGW, lv.gridW); d.maxGH = std::max(d.maxGH, lv.gridH);
        if (lv.nCols <= 0) continue;
        const int nStrips = (lv.nCols + kFastCellsPerItem - 1) / kFastCellsPerItem;   // the cut of gvs_extractor_create
        d.maxSW = std::max(d.maxSW, ((lv.nCols + nStrips - 1) / nStrips) * lv.gridW);
    }
    return d;
}

bool fastTmaBox(const Geometry& g, int* boxWords, int* boxH) {
    const FastDims d = fastDims(g);
    const FastSmemLayout L = fastLayout(d.maxSW, d.maxGW, d.maxGH, true);
    *boxWords = L.tileWords; *boxH = L.tileRows;
    return *boxWords <= 256 && *boxH <= 256 && L.total <= 200 * 1024;
}

void launchFast(const Geometry& g, const uint8_t* slab, const FastItem* items, int nItems, int nImages, CandRec* recs,
                int* nCand, unsigned short* cellCnt, int* status, const CUtensorMap* levelMaps, cudaStream_t s, int* launches) {
    const FastDims d = fastDims(g);
    cudaMemsetAsync(nCand, 0, sizeof(int) * (size_t)nImages * kMaxLevels, s);
    if (nItems <= 0) return;
    static size_t cfgA[64], cfgB[64];
    // strips per CTA (a CTA that takes several requests the next strip's tile while it writes the current strip's records).  Measured
    // on 256 EuRoC frames: 1 strip 1.89 ms, 2: 1.90, 3: 1.92, 4: 1.92, 8: 1.98 -- many short CTAs schedule better than few long
    // ones and four CTAs per SM already hide the tile latency, so the default is 1
    const int perCta = [] { const char* e = getenv("GVS_FAST_ITEMS"); const int v = e ? atoi(e) : 1; return std::min(std::max(v, 1), 64); }();   // (read per launch: tests switch it)
    FastMaps pm{};
    if (levelMaps) {
        const FastSmemLayout L = fastLayout(d.maxSW, d.maxGW, d.maxGH, true);
        memcpy(pm.m, levelMaps, sizeof(CUtensorMap) * (size_t)g.nLevels);
        ensureDynamicSmem(fast_cells_kernel<true>, (size_t)L.total, cfgA);
        fast_cells_kernel<true><<<dim3((nItems + perCta - 1) / perCta, nImages), 256, L.total, s>>>(g, pm, slab, items, nItems, perCta, d.maxSW, d.maxGW, d.maxGH, recs,
                                                                                                 nCand, cellCnt, status);
    } else {
        const FastSmemLayout L = fastLayout(d.maxSW, d.maxGW, d.maxGH, false);
        ensureDynamicSmem(fast_cells_kernel<false>, (size_t)L.total, cfgB);
        fast_cells_kernel<false><<<dim3(nItems, nImages), 256, L.total, s>>>(g, pm, slab, items, nItems, 1, d.maxSW, d.maxGW, d.maxGH, recs, nCand, cellCnt, status);
    }
    ++*launches;
}

}  // namespace gvs
