// Pinhole geometry shared by the depth EM and the P3P instance collector.
// Behavioural source: reference gpu-kernels/optimize_depth.cu:54-82 and collect_p3p_instances.cu:36-55
// (identical helper code in both translation units).  Rounding points follow the reference build.
#pragma once
#include "common.cuh"
#include "residual_model.cuh"

namespace vb {

__device__ __forceinline__ void backproject(const CamBlock& C, float px, float py, float depth, float& ox, float& oy,
                                            float& oz) {
    ox = f_mul(f_fma(C.K4inv[0], px, C.K4inv[1]), depth);
    oy = f_mul(f_fma(C.K4inv[2], py, C.K4inv[3]), depth);
    oz = depth;
}

__device__ __forceinline__ void project(const CamBlock& C, float ox, float oy, float oz, float& px, float& py) {
    px = f_div(f_fma(ox, C.K4[0], f_mul(oz, C.K4[1])), oz);
    py = f_div(f_fma(oy, C.K4[2], f_mul(oz, C.K4[3])), oz);
}

__device__ __forceinline__ void rigid_move(const float* R, const float* t, float& ox, float& oy, float& oz) {
    const float nx = f_fma(oz, R[2], f_fma(ox, R[0], f_mul(oy, R[1])));
    const float ny = f_fma(oz, R[5], f_fma(ox, R[3], f_mul(oy, R[4])));
    const float nz = f_fma(oz, R[8], f_fma(ox, R[6], f_mul(oy, R[7])));
    ox = f_add(nx, t[0]);
    oy = f_add(ny, t[1]);
    oz = f_add(nz, t[2]);
}

// stacked-texture fetch: layer d lives at rows [d*h, (d+1)*h) (reference: gmat.h:175-179)
template <typename T>
__device__ __forceinline__ T fetch_stack(cudaTextureObject_t tex, float x, float y, int d, int h) {
    return tex2D<T>(tex, f_add(x, 0.5f), f_add(f_add(y, (float)((size_t)d * (size_t)h)), 0.5f));
}

}  // namespace vb
