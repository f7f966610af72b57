// Probe of the TMA box-load parameters the FAST kernel uses (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_probe tools/tma_probe.cu && gpurun_out/tma_probe
// For every case: encode a u8 3-D map (pitch x h x images), load one box at (x, y, z) into shared memory, compare with the
// zero-filled expectation.  Cases vary the coordinate alignment, the box against the tensor size, where the descriptor lives
// (kernel parameter / global memory with and without the tensormap proxy fence) and the shared-memory offset of the destination.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../gtsam_vslam_b200/csrc/tma.cuh"

using namespace gvs;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap m[4]; };

template <int MODE>
__global__ void probe(const __grid_constant__ Maps pm, const CUtensorMap* gm, int which, int x, int y, int z, int boxW, int boxH, int dstOff,
                      uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int pad[12];
    __shared__ __align__(8) unsigned long long sBar[2];
    pad[threadIdx.x % 12] = 0;
    if (threadIdx.x == 0) {
        mbarInit(smemAddr(&sBar[0]), 1);
        mbarInit(smemAddr(&sBar[1]), 1);
        const unsigned bar = smemAddr(&sBar[1]);
        mbarExpectTx(bar, (unsigned)(boxW * boxH));
        const unsigned dst = smemAddr(smem + dstOff);
        if (MODE == 0) {
            switch (which) {
                case 0: tmaLoadTile3d(dst, &pm.m[0], x, y, z, bar); break;
                case 1: tmaLoadTile3d(dst, &pm.m[1], x, y, z, bar); break;
                case 2: tmaLoadTile3d(dst, &pm.m[2], x, y, z, bar); break;
                default: tmaLoadTile3d(dst, &pm.m[3], x, y, z, bar); break;
            }
        } else if (MODE == 1) {
            tmaLoadTile3d(dst, gm + which, x, y, z, bar);
        } else if (MODE == 2) {
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(gm + which) : "memory");
            tmaLoadTile3d(dst, gm + which, x, y, z, bar);
        } else {
            tmaLoadTile3d(dst, &pm.m[which], x, y, z, bar);   // dynamic index into the parameter
        }
    }
    __syncthreads();
    mbarWait(smemAddr(&sBar[1]), 0);
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[i] = smem[dstOff + i];
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;   // one case per process: a faulting case takes the context with it
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encoder\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int pitch = 768, h = 480, imgs = 3;
    const size_t slab = (size_t)pitch * h + 4096;
    std::vector<uint8_t> host(slab * imgs);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d, *out;
    cudaMalloc(&d, host.size()); cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 256 * 256);
    struct Case { const char* name; int mode, pitch, w, h, boxW, boxH, x, y, z, dstOff; };
    const Case cases[] = {
        {"param x=16 y=8", 0, 768, 768, 480, 256, 57, 16, 8, 1, 0},
        {"param x=-16 y=-3", 0, 768, 768, 480, 256, 57, -16, -3, 1, 0},
        {"param x=-4 (not 16-aligned)", 0, 768, 768, 480, 256, 57, -4, -3, 1, 0},
        {"param x=36 (not 16-aligned)", 0, 768, 768, 480, 256, 57, 36, 11, 2, 0},
        {"param dst second buffer", 0, 768, 768, 480, 256, 57, 48, 11, 2, 14592},
        {"param box wider than tensor", 0, 64, 64, 40, 256, 57, -16, -3, 0, 0},
        {"param box past right/bottom", 0, 768, 768, 480, 256, 57, 640, 460, 2, 0},
        {"param dynamic index", 3, 768, 768, 480, 256, 57, 32, 11, 2, 0},
        {"global no fence", 1, 768, 768, 480, 256, 57, 32, 11, 2, 0},
        {"global + fence", 2, 768, 768, 480, 256, 57, 32, 11, 2, 0},
    };
    int ci = -1;
    for (const Case& c : cases) {
        if (++ci != only && only >= 0) continue;
        Maps pm;
        const cuuint64_t dims[3] = {(cuuint64_t)c.pitch, (cuuint64_t)c.h, (cuuint64_t)imgs};
        const cuuint64_t strides[2] = {(cuuint64_t)c.pitch, (cuuint64_t)slab};
        const cuuint32_t box[3] = {(cuuint32_t)c.boxW, (cuuint32_t)c.boxH, 1u}, es[3] = {1u, 1u, 1u};
        CUresult r = enc(&pm.m[2], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-32s encode failed %d\n", c.name, (int)r); continue; }
        pm.m[0] = pm.m[1] = pm.m[3] = pm.m[2];
        CUtensorMap* gm;
        cudaMalloc(&gm, sizeof(pm)); cudaMemcpy(gm, &pm, sizeof(pm), cudaMemcpyHostToDevice);
        const int smemBytes = c.dstOff + c.boxW * c.boxH;
        cudaMemset(out, 0xEE, 256 * 256);
#define LAUNCH(M) { cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes); \
                    probe<M><<<1, 256, smemBytes>>>(pm, gm, 2, c.x, c.y, c.z, c.boxW, c.boxH, c.dstOff, out); }
        if (c.mode == 0) LAUNCH(0) else if (c.mode == 1) LAUNCH(1) else if (c.mode == 2) LAUNCH(2) else LAUNCH(3)
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-32s CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 2; }   // the context is gone
        std::vector<uint8_t> got((size_t)c.boxW * c.boxH);
        cudaMemcpy(got.data(), out, got.size(), cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int r2 = 0; r2 < c.boxH; r2++)
            for (int k = 0; k < c.boxW; k++) {
                const int gx = c.x + k, gy = c.y + r2;
                const uint8_t want = (gx < 0 || gx >= c.pitch || gy < 0 || gy >= c.h) ? 0 : host[(size_t)c.z * slab + (size_t)gy * c.pitch + gx];
                bad += got[(size_t)r2 * c.boxW + k] != want;
            }
        printf("%-32s %s (%zu mismatches)\n", c.name, bad ? "WRONG" : "ok", bad);
        cudaFree(gm);
    }
    return 0;
}
