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
#include "common.cuh"

namespace gvs {

// Thread = 4 consecutive destination pixels of one row.  Their source bytes lie in a 12-byte window that starts at a
// 4-byte aligned offset which is the same for every row, so the host stores, per group of 4 columns, the window offset
// and per column {a0 | a1 << 16, PRMT selector of (S[sx], S[sx+1]) inside the window}.  Per pixel: 2 PRMT + 2 IDP2A for
// the two horizontal interpolations, then the fixed-point vertical blend.  3 aligned word loads per source row.
__global__ void __launch_bounds__(256)
pyr_down_kernel(const __grid_constant__ Geometry g, uint8_t* __restrict__ slab, const ResizeTap* __restrict__ taps,
                const uint4* __restrict__ xw, const uint4* __restrict__ xs, const unsigned short* __restrict__ xbase, int level) {
    const LevelGeom& S = g.lv[level - 1];
    const LevelGeom& D = g.lv[level];
    const int grp = blockIdx.x * 32 + threadIdx.x;  // group of 4 destination columns
    const int dx0 = grp * 4;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    if (dx0 >= D.w || dy >= D.h) return;
    const int img = blockIdx.z;
    const uint8_t* src = levelPtr((const uint8_t*)slab, g, img, level - 1);
    uint8_t* dst = levelPtr(slab, g, img, level);

    const ResizeTap ty = taps[D.tabY + dy];
    const int gi = D.tabX / 4 + grp;
    const unsigned base = xbase[gi];
    const unsigned* r0 = reinterpret_cast<const unsigned*>(src + (size_t)ty.ofs * S.pitch + base);
    const unsigned* r1 = reinterpret_cast<const unsigned*>(src + (size_t)ty.pad * S.pitch + base);
    const unsigned p0 = __ldg(r0), p1 = __ldg(r0 + 1), p2 = __ldg(r0 + 2);
    const unsigned q0 = __ldg(r1), q1 = __ldg(r1 + 1), q2 = __ldg(r1 + 2);
    const uint4 wv = __ldg(xw + gi), sv = __ldg(xs + gi);
    const unsigned wts[4] = {wv.x, wv.y, wv.z, wv.w}, sel[4] = {sv.x, sv.y, sv.z, sv.w};
    const unsigned b0 = (unsigned)ty.a0, b1 = (unsigned)ty.a1;
    unsigned out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool hi = (sel[i] >> 16) != 0;
        const unsigned s0 = __byte_perm(hi ? p1 : p0, hi ? p2 : p1, sel[i]);
        const unsigned s1 = __byte_perm(hi ? q1 : q0, hi ? q2 : q1, sel[i]);
        const unsigned h0 = __dp2a_lo(wts[i], s0, 0u), h1 = __dp2a_lo(wts[i], s1, 0u);
        const unsigned v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (v & 0xff) << (8 * i);
    }
    *reinterpret_cast<unsigned*>(dst + (size_t)dy * D.pitch + dx0) = out;
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

void launchPyramid(const Geometry& g, uint8_t* slab, const ResizeTap* taps, const uint4* xw, const uint4* xs,
                   const unsigned short* xbase, int nImages, cudaStream_t s, int* launches) {
    for (int l = 1; l < g.nLevels; l++) {
        dim3 block(32, 8);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 7) / 8, nImages);
        pyr_down_kernel<<<grid, block, 0, s>>>(g, slab, taps, xw, xs, xbase, l);
        ++*launches;
    }
}

}  // namespace gvs
