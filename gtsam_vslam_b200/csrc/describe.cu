// K5+K6+K7 describe: intensity-centroid orientation on the unblurred level, rotated BRIEF-256 on the 7x7-blurred level
// (the blur is evaluated inside the kernel, only where a key samples it), and assembly of the final cv::KeyPoint records at
// their dense output position.  One warp per kept key.
// Replaces FeatureExtractor::computeOrientation / computeAllOrientations (reference src/FeatureExtractor.cpp:315-340,
// :471-479), computeOrbDescriptor / computeDescriptors (:268-313) and the scale-up + concatenation loop of
// extractKeysNew (:518-530).
//
// Orientation: lane v+15 sums row v of the radius-15 disc (umax table, :666-680): m10 = sum u*I, m01 = sum v*I in
// int32 (exact, so the summation order is free), then cv::fastAtan2's float32 polynomial with explicit
// round-to-nearest ops (no FMA contraction).
// Descriptor: lane i builds byte i from 8 tests = 16 samples I[cvRound(x*b + y*a)][cvRound(x*a - y*b)],
// a = (float)cos((double)angle*pi/180), b = (float)sin(...), float32 products and sums rounded separately,
// cvRound = round-half-even (__float2int_rn).  The 32 pattern bytes of a lane live in registers.
// Bound: L2 gather (709 + 512 byte reads per key), not HBM.
#include "common.cuh"

namespace gvs {

__device__ __align__(16) const int8_t kPatternDev[1024] = {
#include "../../include/gvs_orb_pattern.inc"
};

__device__ __forceinline__ float fastAtan2Deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
                p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// The 7x7 sigma-2 blur (same Q8 taps and rounding as blur7_kernel) is evaluated
// only where a key needs it.  Per key the warp stages the 43 x 52-byte source patch (reach 18 + 3 taps, reflect-101 at
// the image border) in shared memory, runs the horizontal pass for the 43 rows (two dp4a per pixel, u16 results in shared
// memory) and evaluates the vertical pass at the 512 sample positions only.  The blurred slab is neither written nor read.
constexpr int kPR = 21, kPRows = 2 * kPR + 1, kPWords = 13, kHCols = 44;
constexpr int kPBytes = (kPRows * kPWords * 4 + 15) & ~15;                       // 2240

__device__ __forceinline__ int reflect101i(int p, int n) {
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

constexpr int kHtRows = 44;  // H is stored column-major (column = image x, 44 u16 per column incl. 1 pad) so that the 7 vertical
                             // taps of a sample are contiguous: 4 word loads instead of 7 scattered u16 loads
constexpr int kDescWarpBytes = 2 * kPBytes + ((kHCols * kHtRows * 2 + 15) & ~15);  // two patch buffers + H^T = 8352

__device__ __forceinline__ void cpAsync4(void* smemDst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}

struct KeyRef { int level, x, y, score, xs, ys; };

__global__ void __launch_bounds__(256, 3)
describe_fused_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ slab, const KeptKey* __restrict__ kept,
                      const int* __restrict__ keptCnt, const int* __restrict__ imgOff, gvs_keypoint* __restrict__ kps,
                      uint8_t* __restrict__ desc, uint2* __restrict__ auxKey, int* __restrict__ momentsDbg, int totalCap) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = blockIdx.y;
    uint8_t* wbase = dsm + warp * kDescWarpBytes;
    unsigned short* H = reinterpret_cast<unsigned short*>(wbase + 2 * kPBytes);
    unsigned pat[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(kPatternDev + lane * 32);
        const uint4 a = p[0], b = p[1];
        pat[0] = a.x; pat[1] = a.y; pat[2] = a.z; pat[3] = a.w; pat[4] = b.x; pat[5] = b.y; pat[6] = b.z; pat[7] = b.w;
    }
    __shared__ int pre[kMaxLevels + 1];  // keys of this image before level l (dense output order)
    __shared__ unsigned oriW[8 * 32], oriO[8 * 32];  // orientation weights per (word, lane = disc row), see below
    for (int t = threadIdx.x; t < 8 * 32; t += blockDim.x) {
        const int i = t >> 5, v = (t & 31) - kHalfPatch;
        const int d = (t & 31) <= 2 * kHalfPatch ? g.umax[v < 0 ? -v : v] : -1;
        unsigned w = 0, o = 0;
        for (int b = 0; b < 4; b++) {
            const int u = 4 * i + b - kHalfPatch;
            if (u >= -d && u <= d) { w |= (unsigned)(u + kHalfPatch) << (8 * b); o |= 1u << (8 * b); }
        }
        oriW[t] = w; oriO[t] = o;
    }
    if (threadIdx.x == 0) {
        int run = 0;
        for (int l = 0; l < kMaxLevels; l++) { pre[l] = run; run += l < g.nLevels ? keptCnt[img * kMaxLevels + l] : 0; }
        pre[kMaxLevels] = run;
    }
    __syncthreads();
    const int outBase = imgOff[img], nKeys = pre[kMaxLevels];
    const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);

    auto keyOf = [&](int k) {
        KeyRef r;
        int level = 0;
#pragma unroll
        for (int l = 1; l < kMaxLevels; l++) level += k >= pre[l] && l < g.nLevels;
        const unsigned q = kept[(size_t)img * g.keptPerImage + g.lv[level].keptOff + (k - pre[level])];
        r.level = level; r.x = q & 0xfff; r.y = (q >> 12) & 0xfff; r.score = q >> 24;
        r.xs = ((r.x - kPR) & ~3) - 4; r.ys = r.y - kPR;
        return r;
    };
    // stage the 43 x 13-word patch of a key into `P` asynchronously with aligned 4-byte cp.async.  Lane = (row parity, word
    // column): 26 lanes copy two rows per step with nothing but pointer increments in the loop.  Keys whose patch touches the
    // image border (they sit >= 19 px inside, the patch reaches 21 + alignment) reflect the row index (101) and gather the
    // words that cross the left / right border byte by byte.
    const int stSub = lane >= kPWords ? 1 : 0, stCol = lane - stSub * kPWords;   // lanes 26..31 idle
    auto stage = [&](const KeyRef& kr, unsigned* P) {
        const LevelGeom& lv = g.lv[kr.level];
        const uint8_t* src = levelPtr(slab, g, img, kr.level);
        if (lane < 2 * kPWords) {
            unsigned sp = (unsigned)__cvta_generic_to_shared(P + stSub * kPWords + stCol);
            const int gx = kr.xs + 4 * stCol;
            if (kr.xs >= 0 && kr.xs + 4 * kPWords <= lv.w && kr.ys >= 0 && kr.ys + kPRows <= lv.h) {
                const uint8_t* gp = src + (size_t)(kr.ys + stSub) * lv.pitch + gx;
                const size_t step = 2 * (size_t)lv.pitch;
#pragma unroll 11
                for (int r = stSub; r < kPRows; r += 2) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sp), "l"(gp));
                    gp += step; sp += 2 * kPWords * 4;
                }
            } else {
                const bool wordIn = gx >= 0 && gx + 3 < lv.w;
                for (int r = stSub; r < kPRows; r += 2, sp += 2 * kPWords * 4) {
                    int iy = kr.ys + r;
                    iy = iy < 0 ? -iy : (iy >= lv.h ? 2 * lv.h - 2 - iy : iy);
                    const uint8_t* row = src + (size_t)iy * lv.pitch;
                    if (wordIn) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sp), "l"(row + gx));
                    } else {
                        unsigned v = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            int xx = gx + b;
                            xx = xx < 0 ? -xx : (xx >= lv.w ? 2 * lv.w - 2 - xx : xx);
                            v |= (unsigned)row[xx] << (8 * b);
                        }
                        P[r * kPWords + stCol] = v;
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    int k = blockIdx.x * (blockDim.x >> 5) + warp;
    int buf = 0;
    KeyRef cur{};
    if (k < nKeys) { cur = keyOf(k); stage(cur, reinterpret_cast<unsigned*>(wbase)); }
    for (; k < nKeys; k += warpsPerGrid, buf ^= 1) {
        unsigned* P = reinterpret_cast<unsigned*>(wbase + buf * kPBytes);
        const int level = cur.level, x = cur.x, y = cur.y, score = cur.score, ox = cur.x - cur.xs;  // ox in [25, 28]
        const LevelGeom& lv = g.lv[level];
        const int o = outBase + k;
        // prefetch the next key's patch into the other buffer, then wait for this key's
        KeyRef nxt{};
        const bool more = k + warpsPerGrid < nKeys;
        if (more) { nxt = keyOf(k + warpsPerGrid); stage(nxt, reinterpret_cast<unsigned*>(wbase + (buf ^ 1) * kPBytes)); }
        if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
        __syncwarp();
        if (o >= totalCap) { cur = nxt; continue; }

        // ---- orientation: lane = row of the radius-15 disc
        int m10 = 0, m01 = 0;
        if (lane < 2 * kHalfPatch + 1) {
            const int v = lane - kHalfPatch;
            // row v of the disc: bytes u = -umax[|v|] .. umax[|v|] around the key column.  The 31 bytes u = -15 .. 15 are
            // 8 words after a funnel shift; the disc extent is folded into per-lane weight tables (oriW = u + 15, oriO = 1,
            // 0 outside the disc), so the two sums are 16 dp4a:  m10 row sum = sum (u + 15) I - 15 sum I.
            const unsigned* rw = P + (kPR + v) * kPWords + ((ox - kHalfPatch) >> 2);
            const unsigned sh = ((ox - kHalfPatch) & 3) * 8;
            unsigned s0 = 0, s1 = 0, lo = rw[0];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const unsigned hi = rw[i + 1];
                const unsigned xw = __funnelshift_r(lo, hi, sh);
                s0 = __dp4a(xw, oriO[i * 32 + lane], s0);
                s1 = __dp4a(xw, oriW[i * 32 + lane], s1);
                lo = hi;
            }
            m10 = (int)s1 - kHalfPatch * (int)s0;
            m01 = v * (int)s0;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, off); m01 += __shfl_xor_sync(0xffffffffu, m01, off); }
        const float angle = fastAtan2Deg((float)m01, (float)m10);

        // ---- horizontal blur pass on the patch: unit = (word m in 1..11, row), consecutive lanes = consecutive rows
        // (conflict-free: patch rows are 13 words apart, H^T columns hold consecutive rows) -> 4 u16 results.  The 7-byte window
        // of output 4m + j lies in the words (m-1, m, m+1); the taps {18,34,48,56,48,34,18} are laid out per word and per j as
        // dp4a constants, so no byte shifts are needed (10 dp4a per unit; exact u32 sums, the order is irrelevant).
        for (int t = lane; t < kPRows * 11; t += 32) {
            const int m = 1 + t / kPRows, r = t - (m - 1) * kPRows;
            const unsigned* w = P + r * kPWords + m;
            const unsigned w0 = w[-1], w1 = w[0], w2 = w[1];
            unsigned short* h = H + (4 * (m - 1)) * kHtRows + r;
            h[0 * kHtRows] = (unsigned short)__dp4a(w0, 0x30221200u, __dp4a(w1, 0x12223038u, 0u));
            h[1 * kHtRows] = (unsigned short)__dp4a(w0, 0x22120000u, __dp4a(w1, 0x22303830u, __dp4a(w2, 0x00000012u, 0u)));
            h[2 * kHtRows] = (unsigned short)__dp4a(w0, 0x12000000u, __dp4a(w1, 0x30383022u, __dp4a(w2, 0x00001222u, 0u)));
            h[3 * kHtRows] = (unsigned short)__dp4a(w1, 0x38302212u, __dp4a(w2, 0x00122230u, 0u));
        }
        __syncwarp();

        // ---- rotated BRIEF on vertically blurred samples
        const float factorPI = (float)(3.14159265358979323846 / 180.0);
        const float ang = __fmul_rn(angle, factorPI);
        double sd, cd;
        sincos((double)ang, &sd, &cd);  // one shared range reduction; same results as separate sin() / cos() (parity-tested)
        const float a = (float)cd, b = (float)sd;
        // H^T as words: column c starts at word c * 22; the taps of a sample at (ix, iy) are rows 18 + iy .. 24 + iy of column ox - 4 + ix
        const unsigned* Hw = reinterpret_cast<const unsigned*>(H);
        auto sample = [&](float px, float py) {
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            const int r0 = kPR - 3 + iy;
            const unsigned* c = Hw + (ox - 4 + ix) * (kHtRows / 2) + (r0 >> 1);
            const unsigned sh = (r0 & 1) * 16;
            const unsigned w0 = c[0], w1 = c[1], w2 = c[2], w3 = c[3];
            const unsigned x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = __funnelshift_r(w2, w3, sh),
                           x3 = (w3 >> sh) & 0xffffu;
            // taps {18,34 | 48,56 | 48,34 | 18} on (v0,v1 | v2,v3 | v4,v5 | v6): exact u32 sums, order is irrelevant
            const unsigned acc = __dp2a_lo(x0, 0x2212u, __dp2a_lo(x1, 0x3830u, __dp2a_lo(x2, 0x2230u, 32768u + 18u * x3)));
            return (int)(acc >> 16);
        };
        unsigned byte = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned w = pat[j];
            const int t0 = sample((float)(int8_t)(w & 0xff), (float)(int8_t)((w >> 8) & 0xff));
            const int t1 = sample((float)(int8_t)((w >> 16) & 0xff), (float)(int8_t)(w >> 24));
            byte |= (unsigned)(t0 < t1) << j;
        }
        desc[(size_t)o * 32 + lane] = (uint8_t)byte;

        if (lane < 7) {
            float fx = (float)x, fy = (float)y;
            if (level != 0) { fx = __fmul_rn(fx, lv.scale); fy = __fmul_rn(fy, lv.scale); }
            unsigned wv;
            switch (lane) {
                case 0: wv = __float_as_uint(fx); break;
                case 1: wv = __float_as_uint(fy); break;
                case 2: wv = __float_as_uint((float)lv.scaledPatch); break;
                case 3: wv = __float_as_uint(angle); break;
                case 4: wv = __float_as_uint((float)score); break;
                case 5: wv = (unsigned)level; break;
                default: wv = 0xffffffffu; break;
            }
            reinterpret_cast<unsigned*>(kps + o)[lane] = wv;
        }
        if (lane == 7) {
            // compact record for the stereo matcher: the row bucket range of destributeRightKeys
            // (src/FeatureMatcher.cpp:741-744: floor(yKey - r) .. ceil(yKey + r), r = 2*scalePyramid[octave]) and pt.y
            float fy = (float)y;
            if (level != 0) fy = __fmul_rn(fy, lv.scale);
            const int yKey = __float2int_rn(fy);
            const float r = __fmul_rn(2.0f, lv.scale);
            const int mn = (int)floorf(__fsub_rn((float)yKey, r)), mx = (int)ceilf(__fadd_rn((float)yKey, r));
            auxKey[o] = make_uint2(((unsigned)mn & 0xffffu) | ((unsigned)mx << 16), __float_as_uint(fy));
        }
        if (momentsDbg && lane == 0) { momentsDbg[2 * o] = m10; momentsDbg[2 * o + 1] = m01; }
        cur = nxt;
        __syncwarp();  // everyone is done with H and this patch buffer before they are refilled
    }
}

void launchDescribeFused(const Geometry& g, const uint8_t* slab, int nImages, const KeptKey* kept, const int* keptCnt,
                         const int* imgOff, gvs_keypoint* kps, uint8_t* desc, uint2* auxKey, int* momentsDbg, int totalCap,
                         cudaStream_t s, int* launches) {
    const int warpsPerCta = 8;
    const int smem = warpsPerCta * kDescWarpBytes;
    static size_t cfg[64];
    ensureDynamicSmem(describe_fused_kernel, (size_t)smem, cfg);
    // Keys per image are only known on the device (about keptPerImage / 2); warps take keys in a grid-stride loop.  A CTA has a
    // fixed set-up cost (pattern and orientation tables), so large batches give every warp about 12 keys (measured on 256-frame
    // batches: 3 keys 1.33 ms, 6: 1.22, 12: 1.17, 24: 1.18, 48: 1.25); small batches keep 3 so that a single frame still spreads
    // over the whole GPU (148 SMs x 3 CTAs x 8 warps).
    const long estKeys = (long)nImages * (g.keptPerImage / 2);
    const int keysPerWarp = (int)std::min<long>(12, std::max<long>(3, estKeys / (148 * 3 * warpsPerCta)));
    int gx = (g.keptPerImage / 2 + keysPerWarp * warpsPerCta - 1) / (keysPerWarp * warpsPerCta);
    if (gx < 1) gx = 1;
    describe_fused_kernel<<<dim3(gx, nImages), 32 * warpsPerCta, smem, s>>>(g, slab, kept, keptCnt, imgOff, kps, desc, auxKey, momentsDbg, totalCap);
    ++*launches;
}

}  // namespace gvs
