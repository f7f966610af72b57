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

__global__ void __launch_bounds__(256)
pyr_down_kernel(const __grid_constant__ Geometry g, uint8_t* __restrict__ slab, const ResizeTap* __restrict__ taps,
                int level) {
    const LevelGeom& S = g.lv[level - 1];
    const LevelGeom& D = g.lv[level];
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    if (dx0 >= D.w || dy >= D.h) return;
    const int img = blockIdx.z;
    const uint8_t* src = levelPtr((const uint8_t*)slab, g, img, level - 1);
    uint8_t* dst = levelPtr(slab, g, img, level);

    const ResizeTap ty = taps[D.tabY + dy];
    const uint8_t* r0 = src + (size_t)ty.ofs * S.pitch;
    const uint8_t* r1 = src + (size_t)ty.pad * S.pitch;
    const int b0 = ty.a0, b1 = ty.a1;

    // four column taps = 32 bytes, 16-byte aligned (tabX is a multiple of 4)
    const uint4* tp = reinterpret_cast<const uint4*>(taps + D.tabX + dx0);
    const uint4 t01 = __ldg(tp), t23 = __ldg(tp + 1);
    const unsigned tw[8] = {t01.x, t01.y, t01.z, t01.w, t23.x, t23.y, t23.z, t23.w};
    unsigned out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int sx = (int)(tw[2 * i] & 0xffff), a0 = (int)(short)(tw[2 * i] >> 16);
        const int a1 = (int)(short)(tw[2 * i + 1] & 0xffff), sx1 = (int)(tw[2 * i + 1] >> 16);
        const int h0 = r0[sx] * a0 + r0[sx1] * a1;
        const int h1 = r1[sx] * a0 + r1[sx1] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (unsigned)(v & 0xff) << (8 * i);
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

void launchPyramid(const Geometry& g, uint8_t* slab, const ResizeTap* taps, int nImages, cudaStream_t s, int* launches) {
    for (int l = 1; l < g.nLevels; l++) {
        dim3 block(32, 8);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 7) / 8, nImages);
        pyr_down_kernel<<<grid, block, 0, s>>>(g, slab, taps, l);
        ++*launches;
    }
}

}  // namespace gvs
