// Double-precision map-point arithmetic shared by project.cu and temporal.cu: explicit round-to-nearest operations in the
// reference's left-to-right order (no FMA contraction), so that the results equal the un-fused host arithmetic bit for bit.
#pragma once
#include "common.cuh"

namespace gvs {

// FeatureTracker::worldToFrame + MapPoint::predictScale (src/FeatureTracker.cpp:685-741, src/Map.cpp:13-23), thread per map
// point.  Double precision, explicit round-to-nearest operations in the reference's left-to-right order (no FMA).
// one map point into one camera; returns worldToFrame's bool and updates pred / scaleLevel only when the point is in the frame
__device__ __forceinline__ bool worldToFrameOne(const double* __restrict__ poseInv, const double* __restrict__ w, double fx, double fy,
                                                double cx, double cy, int width, int height, float maxScaleDist, float logScale,
                                                int nScaleLev, float* pred, int* scaleLevel) {
    double pt[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
        pt[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(poseInv[4 * r], w[0]), __dmul_rn(poseInv[4 * r + 1], w[1])), __dmul_rn(poseInv[4 * r + 2], w[2])),
                          __dmul_rn(poseInv[4 * r + 3], w[3]));
    if (pt[2] <= 0.0) return false;
    const double invZ = __ddiv_rn(1.0, pt[2]);
    const double u = __dadd_rn(__dmul_rn(__dmul_rn(fx, pt[0]), invZ), cx), v = __dadd_rn(__dmul_rn(__dmul_rn(fy, pt[1]), invZ), cy);
    if (u < 0 || v < 0 || u >= width || v >= height) return false;
    const float dist = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(pt[0], pt[0]), __dmul_rn(pt[1], pt[1])), __dmul_rn(pt[2], pt[2])));
    const float dif = __fdiv_rn(maxScaleDist, dist);
    int scale = (int)ceil(__ddiv_rn(log((double)dif), (double)logScale));
    if (scale < 0) scale = 0; else if (scale >= nScaleLev) scale = nScaleLev - 1;
    *scaleLevel = scale; pred[0] = (float)u; pred[1] = (float)v;
    return true;
}

// 4x4 (row-major) times (x, y, z, w), first three rows: Eigen's `pose * vec4` accumulated column by column
__device__ __forceinline__ void mulPose(const double* __restrict__ P, const double* __restrict__ w, double (&pt)[3]) {
#pragma unroll
    for (int r = 0; r < 3; r++)
        pt[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[4 * r], w[0]), __dmul_rn(P[4 * r + 1], w[1])), __dmul_rn(P[4 * r + 2], w[2])),
                          __dmul_rn(P[4 * r + 3], w[3]));
}

}  // namespace gvs
