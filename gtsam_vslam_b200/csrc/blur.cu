// K6 blur7: cv::GaussianBlur(7x7, sigma=2, BORDER_REFLECT_101) of every pyramid level of every image
// (reference src/FeatureExtractor.cpp:512-515; the clone is blurred only to sample descriptors from it).
//
// Arithmetic (bit exact with OpenCV's 8-bit fixed-point path): taps {18,34,48,56,48,34,18}/256,
//   h = sum t[k]*src[x+k-3]            (<= 65280, u16)
//   out = (sum t[k]*h[y+k-3] + 2^15) >> 16
// One CTA = one 64x32 output tile of one (image, level): the (32+6) x (64+8) source tile is staged in shared
// memory (reflect-101 at the image border), the horizontal pass uses two dp4a per pixel on funnel-shifted
// words, the vertical pass keeps u16 rows in shared memory and writes 4-byte coalesced words.
// Bound: HBM for large batches (reads A, writes A per image).
#include "common.cuh"

namespace gvs {

__device__ __forceinline__ int reflect101(int p, int n) {
    // valid for -n < p < 2n-1 (halo of 4 on levels at least 8 wide); general loop for tiny levels
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
blur7_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ slab, uint8_t* __restrict__ blurSlab,
             const BlurItem* __restrict__ items) {
    __shared__ __align__(16) uint8_t sSrc[(kBlurTH + 6) * (kBlurTW + 8)];
    __shared__ __align__(16) unsigned short sH[(kBlurTH + 6) * kBlurTW];
    const BlurItem it = items[blockIdx.x];
    const int img = blockIdx.y;
    const LevelGeom& lv = g.lv[it.level];
    const uint8_t* src = levelPtr(slab, g, img, it.level);
    uint8_t* dst = levelPtr(blurSlab, g, img, it.level);
    const int x0 = it.tx * kBlurTW, y0 = it.ty * kBlurTH;
    const int tid = threadIdx.x;
    constexpr int SP = kBlurTW + 8;  // source tile pitch (bytes): image x = x0 - 4 + i

    // ---- stage source
    for (int i = tid; i < (kBlurTH + 6) * (SP / 4); i += 256) {
        const int r = i / (SP / 4), c = i - r * (SP / 4);
        const int iy = reflect101(y0 - 3 + r, lv.h);
        const int ix = x0 - 4 + 4 * c;
        const uint8_t* row = src + (size_t)iy * lv.pitch;
        unsigned v;
        if (ix >= 0 && ix + 3 < lv.w) v = __ldg(reinterpret_cast<const unsigned*>(row + ix));
        else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int xx = ix + b;
                // columns far right of the image (tile padding) are never used by a stored pixel: clamp cheaply
                const int rx = xx < 2 * lv.w - 1 ? reflect101(xx, lv.w) : lv.w - 1;
                v |= (unsigned)row[rx] << (8 * b);
            }
        }
        *reinterpret_cast<unsigned*>(sSrc + r * SP + 4 * c) = v;
    }
    __syncthreads();

    // ---- horizontal pass: unit = (row, word of 4 px)
    const unsigned tLo = 0x38302212u, tHi = 0x00122230u;  // {18,34,48,56}, {48,34,18,0}
    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW / 4); i += 256) {
        const int r = i / (kBlurTW / 4), c = i - r * (kBlurTW / 4);
        const unsigned* w = reinterpret_cast<const unsigned*>(sSrc + r * SP + 4 * c);
        const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
        unsigned h[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned a = j == 3 ? w1 : __funnelshift_r(w0, w1, 8 * (1 + j));
            const unsigned b = j == 3 ? w2 : __funnelshift_r(w1, w2, 8 * (1 + j));
            h[j] = __dp4a(a, tLo, __dp4a(b, tHi, 0u));
        }
        *reinterpret_cast<uint2*>(sH + r * kBlurTW + 4 * c) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
    }
    __syncthreads();

    // ---- vertical pass: thread = (word column, pair of rows)
    {
        const int c = tid & 15, rp = tid >> 4;  // 16 word columns x 16 row pairs
        const int ry = 2 * rp;
        unsigned acc0[4] = {32768u, 32768u, 32768u, 32768u}, acc1[4] = {32768u, 32768u, 32768u, 32768u};
        const unsigned T[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint2 q = *reinterpret_cast<const uint2*>(sH + (ry + k) * kBlurTW + 4 * c);
            const unsigned v[4] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (k < 7) acc0[j] += T[k] * v[j];
                if (k > 0) acc1[j] += T[k - 1] * v[j];
            }
        }
        const int ox = x0 + 4 * c, oy = y0 + ry;
        if (ox < lv.w) {
            const unsigned o0 = (acc0[0] >> 16) | ((acc0[1] >> 16) << 8) | ((acc0[2] >> 16) << 16) | ((acc0[3] >> 16) << 24);
            const unsigned o1 = (acc1[0] >> 16) | ((acc1[1] >> 16) << 8) | ((acc1[2] >> 16) << 16) | ((acc1[3] >> 16) << 24);
            if (oy < lv.h) *reinterpret_cast<unsigned*>(dst + (size_t)oy * lv.pitch + ox) = o0;
            if (oy + 1 < lv.h) *reinterpret_cast<unsigned*>(dst + (size_t)(oy + 1) * lv.pitch + ox) = o1;
        }
    }
}

void launchBlur(const Geometry& g, const uint8_t* slab, uint8_t* blurSlab, const BlurItem* items, int nItems, int nImages,
                cudaStream_t s, int* launches) {
    if (nItems <= 0) return;
    blur7_kernel<<<dim3(nItems, nImages), 256, 0, s>>>(g, slab, blurSlab, items);
    ++*launches;
}

}  // namespace gvs
