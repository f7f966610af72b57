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
// Work item = up to 4 consecutive cells of one cell row (about 150 x 40 px); CTA = 256 threads.
//   1. the (rows+6) x (cols+6) u8 tile is staged in shared memory with 4-byte loads;
//   2. scores are computed two pixels at a time in u16x2 lanes with the native VIMNMX3.U16x2 min/max
//      (9-wide circular min/max as min3 of min3: 32 + 8 instructions per polarity per pixel pair);
//   3. local maxima are appended to a per-cell list in shared memory, filtered by T_cell, ranked in
//      row-major order inside the cell (k^2/2 compares on ~20 entries) and written as 8-byte records
//      {x, y, score, cell, rank} to a per-(image, level) buffer; K3/K4 (select.cu) restores the reference's
//      emission order from (cell, rank).
// Bound: integer ALU (about 50 lane-ops per pixel), not HBM: see DESIGN.md.
#include "common.cuh"

namespace gvs {

// ring offsets, OpenCV order
__constant__ int8_t kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int8_t kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

struct FastSmemLayout {
    int tilePitch, tileRows;    // image tile (bytes per row, rows)
    int scorePitch, scoreRows;  // score tile including a 1-px zero frame
    int listCap;                // entries per cell list
    int tileOff, scoreOff, listOff, total;
};

__host__ __device__ inline FastSmemLayout fastLayout(int maxGW, int maxGH) {
    FastSmemLayout L;
    L.tilePitch = ((kFastCellsPerItem * maxGW + 6 + 3 + 3) / 4) * 4 + 4;
    L.tileRows = maxGH + 6;
    L.scorePitch = ((kFastCellsPerItem * maxGW + 2 + 3) / 4) * 4;
    L.scoreRows = maxGH + 2;
    L.listCap = ((maxGW + 1) / 2) * ((maxGH + 1) / 2);
    L.tileOff = 0;
    L.scoreOff = L.tileOff + L.tilePitch * L.tileRows;
    L.listOff = ((L.scoreOff + L.scorePitch * L.scoreRows + 15) / 16) * 16;
    L.total = L.listOff + kFastCellsPerItem * L.listCap * 4;
    return L;
}

__device__ __forceinline__ unsigned pair16(const uint8_t* p) { return (unsigned)p[0] | ((unsigned)p[1] << 16); }

__global__ void __launch_bounds__(256)
fast_cells_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ slab,
                  const FastItem* __restrict__ items, int maxGW, int maxGH, CandRec* __restrict__ recs,
                  int* __restrict__ nCand, unsigned short* __restrict__ cellCnt, int* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int sCnt[kFastCellsPerItem], sMax[kFastCellsPerItem], sEmit[kFastCellsPerItem], sBase[kFastCellsPerItem];

    const FastItem it = items[blockIdx.x];
    const int img = blockIdx.y;
    const LevelGeom& lv = g.lv[it.level];
    const FastSmemLayout L = fastLayout(maxGW, maxGH);
    uint8_t* tile = smem + L.tileOff;
    uint8_t* score = smem + L.scoreOff;
    unsigned* lists = reinterpret_cast<unsigned*>(smem + L.listOff);

    const int edge = g.edge;  // (edgeThreshold - fastEdge) + 3: first tested pixel of the first cell
    const int y0 = edge + it.iR * lv.gridH, y1 = min(y0 + lv.gridH, lv.h - edge);
    const int X0 = edge + it.iC0 * lv.gridW, X1 = min(X0 + it.nC * lv.gridW, lv.w - edge);
    const int tid = threadIdx.x;
    if (tid < kFastCellsPerItem) { sCnt[tid] = 0; sMax[tid] = 0; sEmit[tid] = 0; }
    const int rows = y1 - y0, cols = X1 - X0;
    if (rows <= 0 || cols <= 0) {  // empty strip: still publish zero counts for its cells
        if (tid < it.nC) cellCnt[(size_t)img * g.cellsPerImage + lv.cellBase + it.iR * lv.nCols + it.iC0 + tid] = 0;
        return;
    }

    // ---- 1. stage the image tile: x in [xa, xa + 4*tw), y in [y0-3, y1+3)
    const uint8_t* src = levelPtr(slab, g, img, it.level);
    const int xa = (X0 - 3) & ~3;
    const int tw = (X1 + 3 - xa + 3) >> 2;  // words per row
    const int th = rows + 6;
    for (int i = tid; i < tw * th; i += 256) {
        const int r = i / tw, c = i - r * tw;
        const unsigned v = __ldg(reinterpret_cast<const unsigned*>(src + (size_t)(y0 - 3 + r) * lv.pitch + xa) + c);
        *reinterpret_cast<unsigned*>(tile + r * L.tilePitch + 4 * c) = v;
    }
    // zero the score tile (frame included)
    for (int i = tid; i < (L.scorePitch * (rows + 2)) >> 2; i += 256) reinterpret_cast<unsigned*>(score)[i] = 0;
    __syncthreads();

    // ---- 2. scores, two horizontally adjacent pixels per step
    const int pairsPerRow = (cols + 1) >> 1;
    const int tMin = g.tMin;
    for (int i = tid; i < pairsPerRow * rows; i += 256) {
        const int ry = i / pairsPerRow, px = (i - ry * pairsPerRow) * 2;
        const uint8_t* c = tile + (ry + 3) * L.tilePitch + (X0 - xa) + px;
        unsigned r[16];
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = pair16(c + kRingDy[k] * L.tilePitch + kRingDx[k]);
        unsigned mn3[16], mx3[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            mn3[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
            mx3[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        }
        unsigned B = 0, D = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const unsigned mn9 = __vimin3_u16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
            const unsigned mx9 = __vimax3_u16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
            B = __vmaxu2(B, mn9);
            D = __vminu2(D, mx9);
        }
        const int v0 = c[0], v1 = c[1];
        const int best0 = max((int)(B & 0xffff) - v0, v0 - (int)(D & 0xffff));
        const int best1 = max((int)(B >> 16) - v1, v1 - (int)(D >> 16));
        uint8_t* sp = score + (ry + 1) * L.scorePitch + 1 + px;
        if (best0 > tMin) sp[0] = (uint8_t)(best0 - 1);
        if (best1 > tMin && px + 1 < cols) sp[1] = (uint8_t)(best1 - 1);
    }
    __syncthreads();

    // ---- 3. strict local maxima inside each cell -> per-cell lists
    const int gW = lv.gridW;
    for (int i = tid; i < cols * rows; i += 256) {
        const int ry = i / cols, lx = i - ry * cols;
        const uint8_t* sp = score + (ry + 1) * L.scorePitch + 1 + lx;
        const int s = sp[0];
        if (s == 0) continue;
        const int cl = (lx >= gW) + (lx >= 2 * gW) + (lx >= 3 * gW);
        const int cx0 = cl * gW, cx1 = min(cx0 + gW, cols);
        const bool hasL = lx > cx0, hasR = lx + 1 < cx1;
        const int P = L.scorePitch;
        bool ok = s > sp[-P] && s > sp[P];
        if (hasL) ok = ok && s > sp[-1] && s > sp[-P - 1] && s > sp[P - 1];
        if (hasR) ok = ok && s > sp[1] && s > sp[-P + 1] && s > sp[P + 1];
        if (!ok) continue;
        const int slot = atomicAdd(&sCnt[cl], 1);
        lists[cl * L.listCap + slot] = ((unsigned)ry << 16) | ((unsigned)(lx - cx0) << 8) | (unsigned)s;
        atomicMax(&sMax[cl], s);
    }
    __syncthreads();

    // ---- 4. threshold per cell, count emitted, reserve, rank inside the cell (row-major), write
    const int tMax = g.tMax;
    int total = 0;
    for (int c = 0; c < it.nC; c++) total += sCnt[c];
    for (int e = tid; e < total; e += 256) {
        int c = 0, off = e;
        while (off >= sCnt[c]) { off -= sCnt[c]; c++; }
        const int T = sMax[c] >= tMax ? tMax : tMin;
        if ((int)(lists[c * L.listCap + off] & 0xff) >= T) atomicAdd(&sEmit[c], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int c = 0; c < it.nC; c++) { sBase[c] = tot; tot += sEmit[c]; }
        const int base = tot ? atomicAdd(&nCand[img * kMaxLevels + it.level], tot) : 0;
        if (base + tot > lv.candCap) atomicOr(status, kStatCandOverflow);
        for (int c = 0; c < it.nC; c++) sBase[c] += base;
    }
    if (tid < it.nC)
        cellCnt[(size_t)img * g.cellsPerImage + lv.cellBase + it.iR * lv.nCols + it.iC0 + tid] = (unsigned short)sEmit[tid];
    __syncthreads();
    CandRec* out = recs + (size_t)img * g.candPerImage + lv.candOff;
    for (int e = tid; e < total; e += 256) {
        int c = 0, off = e;
        while (off >= sCnt[c]) { off -= sCnt[c]; c++; }
        const unsigned* lc = lists + c * L.listCap;
        const unsigned key = lc[off];
        const int T = sMax[c] >= tMax ? tMax : tMin;
        if ((int)(key & 0xff) < T) continue;
        int rank = 0;
        const int n = sCnt[c];
        for (int j = 0; j < n; j++) {
            const unsigned o = lc[j];
            rank += ((int)(o & 0xff) >= T) && ((o >> 8) < (key >> 8));
        }
        const int slot = sBase[c] + rank;
        if (slot >= lv.candCap) continue;
        const int x = X0 + c * gW + (int)((key >> 8) & 0xff), y = y0 + (int)(key >> 16);
        CandRec rec;
        rec.w0 = (unsigned)x | ((unsigned)y << 12) | ((key & 0xff) << 24);
        rec.w1 = (unsigned)(it.iR * lv.nCols + it.iC0 + c) | ((unsigned)rank << 16);
        out[slot] = rec;
    }
}

void launchFast(const Geometry& g, const uint8_t* slab, const FastItem* items, int nItems, int nImages, CandRec* recs,
                int* nCand, unsigned short* cellCnt, int* status, cudaStream_t s, int* launches) {
    int maxGW = 0, maxGH = 0;
    for (int l = 0; l < g.nLevels; l++) { maxGW = std::max(maxGW, g.lv[l].gridW); maxGH = std::max(maxGH, g.lv[l].gridH); }
    const FastSmemLayout L = fastLayout(maxGW, maxGH);
    static int configured = 0;
    if (L.total > 48 * 1024 && configured < L.total) {
        cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
        configured = L.total;
    }
    cudaMemsetAsync(nCand, 0, sizeof(int) * (size_t)nImages * kMaxLevels, s);
    if (nItems > 0) {
        fast_cells_kernel<<<dim3(nItems, nImages), 256, L.total, s>>>(g, slab, items, maxGW, maxGH, recs, nCand, cellCnt, status);
        ++*launches;
    }
}

}  // namespace gvs
