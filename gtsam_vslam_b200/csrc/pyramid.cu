// K1 pyr_down: level l = cv::resize(level l-1, INTER_LINEAR) for every image of the batch.
// Replaces FeatureExtractor::computePyramid (reference src/FeatureExtractor.cpp:342-366); the 19-px
// copyMakeBorder frame is dead data (never read on the path) and is not materialised.
//
// Arithmetic (bit exact with OpenCV's 8-bit fixed-point bilinear):
//   H[dx]  = S[sx]*a0 + S[sx+1]*a1                       (a0+a1 = 2048, int32)
//   out    = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// Coefficients are precomputed on the host exactly as cv::resize does (float32 rounding) and read from a
// small table that stays in L1/L2.  HBM-bound: reads level l-1 once (neighbouring threads share rows
// through L1), writes level l once with 4-byte coalesced stores.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace gvs {

// Thread = 4 consecutive destination pixels of two consecutive rows.  Their source bytes lie in a short window whose position
// is the same for every row, so the host stores, per group of 4 columns, the window offset and per column
// {a0 | a1 << 16, PRMT selector of (S[sx], S[sx+1]) inside the window}; both tables are shared by the two rows.
// NARROW (every group of the level spans at most 8 source bytes: scale factors up to about 1.75): the window starts at the
// first column's sx; two funnel shifts align the three loaded words to it and one PRMT per pixel picks the pair.
// Otherwise the window is 12 bytes from the aligned offset and the selector carries which 8-byte half holds the pair.
// Per pixel and source row: 1 PRMT + 1 IDP2A (a0*S[sx] + a1*S[sx+1]); the horizontal interpolation of a source row that
// both destination rows need (sy(dy+1) = sy(dy) + 1, the common case at scale 1.2) is computed once.  The vertical blend
// ((b*(H>>4))>>16 as a high multiply by b << 16) and the 4-byte store follow.  3 aligned word loads per source row.
template <bool NARROW>
__global__ void __launch_bounds__(256)
pyr_down_kernel(const __grid_constant__ Geometry g, uint8_t* __restrict__ slab, const ResizeTap* __restrict__ taps,
                const uint4* __restrict__ xw, const uint4* __restrict__ xs, const unsigned short* __restrict__ xbase, int level) {
    const LevelGeom& S = g.lv[level - 1];
    const LevelGeom& D = g.lv[level];
    const int grp = blockIdx.x * 32 + threadIdx.x;  // group of 4 destination columns
    const int dx0 = grp * 4;
    const int dy = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (dx0 >= D.w || dy >= D.h) return;
    const bool two = dy + 1 < D.h;
    const int img = blockIdx.z;
    const uint8_t* src = levelPtr((const uint8_t*)slab, g, img, level - 1);
    uint8_t* dst = levelPtr(slab, g, img, level);

    const int gi = D.tabX / 4 + grp;
    const ResizeTap t0 = taps[D.tabY + dy], t1 = taps[D.tabY + dy + (two ? 1 : 0)];
    const unsigned base = xbase[gi];
    const unsigned al = NARROW ? (base & ~3u) : base, sh = (base & 3u) * 8u;
    const uint4 wv = __ldg(xw + gi), sv = __ldg(xs + gi);
    const unsigned wts[4] = {wv.x, wv.y, wv.z, wv.w}, sel[4] = {sv.x, sv.y, sv.z, sv.w};
    // horizontal interpolation of one source row for the 4 columns, already shifted right by 4
    auto hrow = [&](int row, unsigned (&h)[4]) {
        const unsigned* r = reinterpret_cast<const unsigned*>(src + (size_t)row * S.pitch + al);
        const unsigned p0 = __ldg(r), p1 = __ldg(r + 1), p2 = __ldg(r + 2);
        if (NARROW) {
            const unsigned lo = __funnelshift_r(p0, p1, sh), hi = __funnelshift_r(p1, p2, sh);
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = __dp2a_lo(wts[i], __byte_perm(lo, hi, sel[i]), 0u) >> 4;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const bool up = (sel[i] >> 16) != 0;
                h[i] = __dp2a_lo(wts[i], __byte_perm(up ? p1 : p0, up ? p2 : p1, sel[i]), 0u) >> 4;
            }
        }
    };
    auto blend = [&](const unsigned (&h0)[4], const unsigned (&h1)[4], const ResizeTap& t, int y) {
        const unsigned b0 = (unsigned)t.a0 << 16, b1 = (unsigned)t.a1 << 16;   // (b*(h))>>16 == umulhi(b << 16, h), b <= 2048
        unsigned out = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned v = (__umulhi(b0, h0[i]) + __umulhi(b1, h1[i]) + 2) >> 2;
            out |= (v & 0xff) << (8 * i);
        }
        *reinterpret_cast<unsigned*>(dst + (size_t)y * D.pitch + dx0) = out;
    };
    unsigned hA0[4], hA1[4], hB0[4], hB1[4];
    hrow(t0.ofs, hA0);
    if (t0.pad != t0.ofs) hrow(t0.pad, hA1);
    else {
#pragma unroll
        for (int i = 0; i < 4; i++) hA1[i] = hA0[i];
    }
    blend(hA0, hA1, t0, dy);
    if (!two) return;
    // the rows of dy + 1 (uniform across the warp: all its threads share dy)
    if (t1.ofs == t0.pad) {
#pragma unroll
        for (int i = 0; i < 4; i++) hB0[i] = hA1[i];
    } else if (t1.ofs == t0.ofs) {
#pragma unroll
        for (int i = 0; i < 4; i++) hB0[i] = hA0[i];
    } else hrow(t1.ofs, hB0);
    if (t1.pad == t1.ofs) {
#pragma unroll
        for (int i = 0; i < 4; i++) hB1[i] = hB0[i];
    } else if (t1.pad == t0.pad) {
#pragma unroll
        for (int i = 0; i < 4; i++) hB1[i] = hA1[i];
    } else hrow(t1.pad, hB1);
    blend(hB0, hB1, t1, dy + 1);
}

// ---- fused pyramid -----------------------------------------------------------------------------------------------------
// The level chain is sequential (level l is a resize of level l-1), but it is local: a tile of level l+3 depends on a slightly
// larger tile of level l+2, ... of level l.  A CTA owns a tile of the coarsest level of a group (host plan: PyrGroup / PyrSpan),
// fetches the region of the group's base level it depends on with ONE TMA tile load (cp.async.bulk.tensor, halo included,
// out-of-image parts zero filled and never read) and produces the regions of the group's levels one after the other in shared
// memory with the arithmetic of pyr_down_kernel<true>; every level's pixels are written to global memory by exactly one tile
// (the "own" ranges partition the level).  HBM traffic per image: the base levels once + every produced level once (the
// un-fused chain re-reads every intermediate level), 2-3 launches instead of nLevels - 1, and no exposed global-load latency
// in the inner loop.
constexpr int kPyrMaxThreads = 512;
__global__ void __launch_bounds__(kPyrMaxThreads)
pyr_fused_kernel(const __grid_constant__ Geometry g, const __grid_constant__ CUtensorMap baseMap, const __grid_constant__ PyrGroup G,
                 uint8_t* __restrict__ slab, const ResizeTap* __restrict__ taps, const uint4* __restrict__ xw,
                 const uint4* __restrict__ xs, const unsigned short* __restrict__ xbase, const PyrSpan* __restrict__ spans) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long barrier;
    const int tid = threadIdx.x, img = blockIdx.y;
    const int ty = blockIdx.x / G.nTx, tx = blockIdx.x - ty * G.nTx;
    __shared__ PyrSpan SX, SY;   // (indexed by the level inside the group: shared memory, not registers)
    if (tid < (int)(sizeof(PyrSpan) / sizeof(short))) {
        reinterpret_cast<short*>(&SX)[tid] = reinterpret_cast<const short*>(spans + G.spanX + tx)[tid];
        reinterpret_cast<short*>(&SY)[tid] = reinterpret_cast<const short*>(spans + G.spanY + ty)[tid];
    }
    static_assert(sizeof(PyrSpan) / sizeof(short) <= 256, "one thread per 16-bit word of a span");
    const unsigned bar = smemAddr(&barrier);
    if (tid == 0) {
        const short* gx = reinterpret_cast<const short*>(spans + G.spanX + tx);
        const short* gy = reinterpret_cast<const short*>(spans + G.spanY + ty);
        mbarInit(bar, 1);
        mbarExpectTx(bar, (unsigned)(G.boxW * G.boxH));
        tmaLoadTile3d(smemAddr(sm), &baseMap, gx[0], gy[0], img, bar);   // PyrSpan::lo[0]
    }
    if (tid == 0) mbarWait(bar, 0);   // one thread polls (a polling warp takes issue slots from the CTAs that compute) ...
    __syncthreads();                  // ... and the CTA barrier passes the arrival on
    for (int k = 1; k <= G.nOut; k++) {
        const int level = G.base + k;
        const LevelGeom& D = g.lv[level];
        const uint8_t* sbuf = sm + G.bufOff[k - 1];
        const int sPitch = G.bufPitch[k - 1], sX0 = SX.lo[k - 1], sY0 = SY.lo[k - 1];
        uint8_t* dbuf = k < G.nOut ? sm + G.bufOff[k] : nullptr;
        const int dPitch = k < G.nOut ? G.bufPitch[k] : 0;
        const int cLo = SX.lo[k], cHi = SX.hi[k], rLo = SY.lo[k], rHi = SY.hi[k];
        const int oxLo = SX.ownLo[k], oxHi = SX.ownHi[k], oyLo = SY.ownLo[k], oyHi = SY.ownHi[k];
        uint8_t* dst = levelPtr(slab, g, img, level);
        // thread = one group of 4 destination columns (its column tables stay in registers) and every nStripes-th row of the
        // region.  Both source rows of a destination row are interpolated on the spot: no dependence between the rows of a
        // thread, no reuse bookkeeping -- the inner loop is straight-line code on shared memory.
        const int nGx = (cHi - cLo + 3) >> 2, nRows = rHi - rLo;
        const int nStripes = max((int)blockDim.x / nGx, 1);   // stripes of row pairs
        const unsigned magic = SX.magic[k];   // stripe = tid / nGx = umulhi(tid, magic) (host: 2^32 / nGx + 1)
        const int stripe = nGx > 1 ? (int)__umulhi((unsigned)tid, magic) : tid, gx = tid - stripe * nGx;
        if (stripe < nStripes) {
            const int dx0 = cLo + 4 * gx;
            const int gi = (D.tabX + dx0) >> 2;
            const unsigned base = xbase[gi];
            const unsigned sh = (base & 3u) * 8u;
            const uint4 wv = __ldg(xw + gi), sv = __ldg(xs + gi);
            const uint8_t* s0 = sbuf + ((int)(base & ~3u) - sX0) - sY0 * sPitch;
            uint8_t* d0 = dbuf ? dbuf + (dx0 - cLo) - rLo * dPitch : nullptr;
            uint8_t* g0 = dst + dx0;
            const bool ownX = dx0 >= oxLo && dx0 < oxHi;
            const int dp = D.pitch;
            const uint2* ty = reinterpret_cast<const uint2*>(taps + D.tabY);   // {ofs | a0 << 16, a1 | pad << 16}: one 8-byte load
            auto hrow = [&](int row, unsigned& h0, unsigned& h1, unsigned& h2, unsigned& h3) {
                const unsigned* r = reinterpret_cast<const unsigned*>(s0 + row * sPitch);
                const unsigned p0 = r[0], p1 = r[1], p2 = r[2];
                const unsigned lo = __funnelshift_r(p0, p1, sh), hi = __funnelshift_r(p1, p2, sh);
                h0 = __dp2a_lo(wv.x, __byte_perm(lo, hi, sv.x), 0u) >> 4;
                h1 = __dp2a_lo(wv.y, __byte_perm(lo, hi, sv.y), 0u) >> 4;
                h2 = __dp2a_lo(wv.z, __byte_perm(lo, hi, sv.z), 0u) >> 4;
                h3 = __dp2a_lo(wv.w, __byte_perm(lo, hi, sv.w), 0u) >> 4;
            };
            // vertical blend of 4 pixels: (hi(b0*u) + hi(b1*v) + 2) >> 2 per pixel; the four 10-bit sums are packed into 16-bit
            // lanes first so that two shifts and one PRMT (bytes 0 and 2 of each word) finish all four
            auto blend = [&](unsigned b0, unsigned b1, unsigned u0, unsigned u1, unsigned u2, unsigned u3, unsigned v0, unsigned v1,
                             unsigned v2, unsigned v3) {
                const unsigned s0 = __umulhi(b0, u0) + 2u + __umulhi(b1, v0), s1 = __umulhi(b0, u1) + 2u + __umulhi(b1, v1);
                const unsigned s2 = __umulhi(b0, u2) + 2u + __umulhi(b1, v2), s3 = __umulhi(b0, u3) + 2u + __umulhi(b1, v3);
                const unsigned p01 = __byte_perm(s0, s1, 0x5410) >> 2, p23 = __byte_perm(s2, s3, 0x5410) >> 2;
                return __byte_perm(p01, p23, 0x6420);
            };
            // thread rows: pairs (y, y + 1) of consecutive destination rows, every nStripes-th pair: the second row usually shares
            // its upper source row with the first one's lower row (uniform across the warps of a stripe)
            const int y0 = rLo + 2 * stripe;
            uint8_t* dptr = d0 ? d0 + y0 * dPitch : nullptr;        // running pointers: one add per pair
            uint8_t* gptr = g0 + (size_t)y0 * dp;
            const int dStep = 2 * nStripes * dPitch, gStep = 2 * nStripes * dp;
            for (int y = y0; y < rHi; y += 2 * nStripes, dptr += dStep, gptr += gStep) {
                const uint2 t = __ldg(ty + y);
                const int rA = (int)(t.x & 0xffffu), rB = (int)(t.y >> 16);
                unsigned a0, a1, a2, a3, c0, c1, c2, c3;
                hrow(rA, a0, a1, a2, a3);
                hrow(rB, c0, c1, c2, c3);
                const unsigned w0 = blend(t.x & 0xffff0000u, t.y << 16, a0, a1, a2, a3, c0, c1, c2, c3);
                if (d0) *reinterpret_cast<unsigned*>(dptr) = w0;
                if (ownX && y >= oyLo && y < oyHi) *reinterpret_cast<unsigned*>(gptr) = w0;
                if (y + 1 >= rHi) break;
                const uint2 t2 = __ldg(ty + y + 1);
                const int rC = (int)(t2.x & 0xffffu), rD = (int)(t2.y >> 16);
                unsigned e0, e1, e2, e3, f0, f1, f2, f3;
                if (rC == rB) { e0 = c0; e1 = c1; e2 = c2; e3 = c3; }
                else if (rC == rA) { e0 = a0; e1 = a1; e2 = a2; e3 = a3; }
                else hrow(rC, e0, e1, e2, e3);
                if (rD == rC) { f0 = e0; f1 = e1; f2 = e2; f3 = e3; }
                else if (rD == rB) { f0 = c0; f1 = c1; f2 = c2; f3 = c3; }
                else hrow(rD, f0, f1, f2, f3);
                const unsigned w1 = blend(t2.x & 0xffff0000u, t2.y << 16, e0, e1, e2, e3, f0, f1, f2, f3);
                if (d0) *reinterpret_cast<unsigned*>(dptr + dPitch) = w1;
                if (ownX && y + 1 >= oyLo && y + 1 < oyHi) *reinterpret_cast<unsigned*>(gptr + dp) = w1;
            }
        }
        __syncthreads();
    }
}

void launchPyramidFused(const Geometry& g, uint8_t* slab, const ResizeTap* taps, const uint4* xw, const uint4* xs,
                        const unsigned short* xbase, const PyrPlan& plan, const PyrSpan* spans, const CUtensorMap* baseMaps,
                        int nImages, cudaStream_t s, int* launches) {
    static size_t cfg[64];
    for (int i = 0; i < plan.nGroups; i++) {
        const PyrGroup& G = plan.grp[i];
        ensureDynamicSmem(pyr_fused_kernel, (size_t)G.smemBytes, cfg);
        static const int threads = [] { const char* e = getenv("GVS_PYR_THREADS"); const int v = e ? atoi(e) : 256; return std::min(std::max(v & ~31, 64), kPyrMaxThreads); }();
        pyr_fused_kernel<<<dim3(G.nTx * G.nTy, nImages), threads, G.smemBytes, s>>>(g, baseMaps[i], G, slab, taps, xw, xs, xbase, spans);
        ++*launches;
    }
}

// Tightly packed host batches arrive with one contiguous DMA (about 55 GB/s over PCIe Gen5, against about 8 GB/s for a
// pitched copy with 752-byte rows); this kernel spreads them into the pitched level-0 store (16-byte stores).
__global__ void __launch_bounds__(256)
repack_level0_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ tight, uint8_t* __restrict__ slab) {
    const LevelGeom& lv = g.lv[0];
    const int img = blockIdx.z, y = blockIdx.y * 4 + threadIdx.y;
    const int x = (blockIdx.x * 64 + threadIdx.x) * 16;
    if (x >= lv.w || y >= lv.h) return;
    const uint8_t* src = tight + ((size_t)img * lv.h + y) * lv.w + x;
    uint8_t* dst = slab + (size_t)img * g.slabBytes + lv.off + (size_t)y * lv.pitch + x;
    uint4 v;
    if (x + 16 <= lv.w && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        v = __ldg(reinterpret_cast<const uint4*>(src));
    } else {
        unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (x + i < lv.w) w[i >> 2] |= (unsigned)src[i] << (8 * (i & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(dst) = v;
}

void launchRepack(const Geometry& g, const uint8_t* tight, uint8_t* slab, int nImages, cudaStream_t s, int* launches) {
    dim3 grid((g.lv[0].w + 1023) / 1024, (g.lv[0].h + 3) / 4, nImages);
    repack_level0_kernel<<<grid, dim3(64, 4), 0, s>>>(g, tight, slab);
    ++*launches;
}

// Upstream image preparation: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with the CV_32FC1 rectification maps
// (reference src/VIOSlam.cpp:282-301) fused with cvtColor BGR2GRAY (src/FeatureTracker.cpp:1131-1135), written straight into the
// pitched level-0 store.  OpenCV arithmetic: sx = cvRound(mapx * 32) (float32), 5 fractional bits, weights (32-fx)(32-fy)*32 etc.
// (the first one saturated to short), channel = (sum + 2^14) >> 15; gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15.
// Thread = 4 destination pixels of one row (one 4-byte store).  HBM/L2 gather bound: 8 map bytes + 4 taps per pixel.
template <int CH>
__global__ void __launch_bounds__(256)
prepare_level0_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ srcImgs, const float* __restrict__ mapx,
                      const float* __restrict__ mapy, uint8_t* __restrict__ slab) {
    const LevelGeom& lv = g.lv[0];
    const int img = blockIdx.z, y = blockIdx.y * 4 + threadIdx.y;
    const int x0 = (blockIdx.x * 64 + threadIdx.x) * 4;
    if (x0 >= lv.w || y >= lv.h) return;
    const uint8_t* src = srcImgs + (size_t)img * lv.w * lv.h * CH;
    unsigned out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        int v[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) v[c] = 0;
        if (x < lv.w) {
            if (mapx) {
                const size_t m = (size_t)y * lv.w + x;
                const int sx = __float2int_rn(__fmul_rn(__ldg(mapx + m), 32.f)), sy = __float2int_rn(__fmul_rn(__ldg(mapy + m), 32.f));
                const int fx = sx & 31, fy = sy & 31;
                const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
                const int wt[4] = {min((32 - fx) * (32 - fy) * 32, 32767), fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int px = ix + (t & 1), py = iy + (t >> 1);
                    if (px >= 0 && px < lv.w && py >= 0 && py < lv.h) {
                        const uint8_t* p = src + ((size_t)py * lv.w + px) * CH;
#pragma unroll
                        for (int c = 0; c < CH; c++) v[c] += (int)p[c] * wt[t];
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; c++) v[c] = min(max((v[c] + 16384) >> 15, 0), 255);
            } else {
                const uint8_t* p = src + ((size_t)y * lv.w + x) * CH;
#pragma unroll
                for (int c = 0; c < CH; c++) v[c] = p[c];
            }
        }
        const int gray = CH == 3 ? (v[0] * 3735 + v[CH > 1 ? 1 : 0] * 19235 + v[CH > 2 ? 2 : 0] * 9798 + 16384) >> 15 : v[0];
        out |= (unsigned)(gray & 0xff) << (8 * i);
    }
    *reinterpret_cast<unsigned*>(slab + (size_t)img * g.slabBytes + lv.off + (size_t)y * lv.pitch + x0) = out;
}

void launchPrepare(const Geometry& g, const uint8_t* src, int channels, const float* mapx, const float* mapy, uint8_t* slab,
                   int nImages, cudaStream_t s, int* launches) {
    dim3 grid((g.lv[0].w + 255) / 256, (g.lv[0].h + 3) / 4, nImages);
    if (channels == 3) prepare_level0_kernel<3><<<grid, dim3(64, 4), 0, s>>>(g, src, mapx, mapy, slab);
    else prepare_level0_kernel<1><<<grid, dim3(64, 4), 0, s>>>(g, src, mapx, mapy, slab);
    ++*launches;
}

void launchPyramid(const Geometry& g, uint8_t* slab, const ResizeTap* taps, const uint4* xw, const uint4* xs,
                   const unsigned short* xbase, unsigned narrowMask, int nImages, cudaStream_t s, int* launches) {
    for (int l = 1; l < g.nLevels; l++) {
        dim3 block(32, 8);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 15) / 16, nImages);
        if ((narrowMask >> l) & 1u) pyr_down_kernel<true><<<grid, block, 0, s>>>(g, slab, taps, xw, xs, xbase, l);
        else pyr_down_kernel<false><<<grid, block, 0, s>>>(g, slab, taps, xw, xs, xbase, l);
        ++*launches;
    }
}

}  // namespace gvs
