// voldor_b200 — shared device/host helpers for the sm_100a kernels.
//
// Storage types used by every kernel family of the EM hot path.  They replace the reference's GMat<T>
// (reference: gpu-kernels/gmat.h:4-204) with two purpose-built containers:
//   * TexStack<T>  – pitched stack of `layers` images bound to ONE pitch2D texture with linear filtering,
//                    clamp addressing and unnormalised coordinates.  The stacking (layer d starts at row
//                    d*h) and the "grow only, keep the larger layer count" policy are observable through
//                    bilinear fetches near the bottom row of a layer (gmat.h:19-22,39-66,175-179), so they
//                    are kept bit-for-bit.
//   * Plane<T>     – dense layered array with a 128-byte aligned row pitch for coalesced / vectorised
//                    access by the non-interpolated streams (rigidness, depth, cost, RNG words).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define VB_DIV_CEIL(x, y) (((x) + (y)-1) / (y))

// Error convention of the reference ABI (gpu-kernels/utils.h:21-26): print "GPUassert", return the code.
#define VB_RETURN_IF_CUDA_ERROR()                                                                        \
    do {                                                                                                 \
        cudaError_t vb_code_ = cudaGetLastError();                                                       \
        if (vb_code_ != cudaSuccess) {                                                                   \
            printf("GPUassert : %s\n%s at line %d\n", cudaGetErrorString(vb_code_), __FILE__, __LINE__); \
            return (int)vb_code_;                                                                        \
        }                                                                                                \
    } while (0)

#define VB_CUDA(call)                                                                                    \
    do {                                                                                                 \
        cudaError_t vb_code_ = (call);                                                                   \
        if (vb_code_ != cudaSuccess) {                                                                   \
            printf("GPUassert : %s\n%s at line %d\n", cudaGetErrorString(vb_code_), __FILE__, __LINE__); \
            return (int)vb_code_;                                                                        \
        }                                                                                                \
    } while (0)

namespace vb {

constexpr int kMaxFrames = 16;      // reference: optimize_depth.cu:20, collect_p3p_instances.cu:10
constexpr int kMaxPriorFrames = 16; // reference: optimize_depth.cu:21

// ---------------------------------------------------------------------------------------------------
// TexStack<T>: layered pitched image + stacked pitch2D texture (bilinear, clamp, unnormalised).
// ---------------------------------------------------------------------------------------------------
template <typename T>
struct TexStack {
    T* ptr = nullptr;
    size_t pitch = 0;  // bytes
    int w = 0, h = 0, layers = 0;
    cudaTextureObject_t tex = 0;

    // Returns 1 when (re)allocated — contents and texture are then fresh — else 0.
    // `lazy_layers` reproduces GMat::create(..., lazy_depth=true): an allocation with at least the
    // requested layer count is kept (gmat.h:19-22).
    int ensure(int w_, int h_, int layers_, bool lazy_layers) {
        if ((w_ == w && h_ == h && layers_ == layers) || (lazy_layers && w_ == w && h_ == h && layers_ <= layers))
            return 0;
        release();
        if (w_ <= 0 || h_ <= 0 || layers_ <= 0) return 1;
        if (cudaMallocPitch((void**)&ptr, &pitch, (size_t)w_ * sizeof(T), (size_t)h_ * layers_) != cudaSuccess) {
            ptr = nullptr;
            return 1;
        }
        w = w_, h = h_, layers = layers_;
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<T>();
        rd.res.pitch2D.devPtr = ptr;
        rd.res.pitch2D.width = (size_t)w;
        rd.res.pitch2D.height = (size_t)h * layers;
        rd.res.pitch2D.pitchInBytes = pitch;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        return 1;
    }
    void release() {
        if (tex) cudaDestroyTextureObject(tex);
        if (ptr) cudaFree(ptr);
        tex = 0, ptr = nullptr, pitch = 0, w = h = layers = 0;
    }
    T* layer(int d) const { return (T*)((char*)ptr + (size_t)d * h * pitch); }
    cudaError_t upload_layer(const T* host, int d, cudaStream_t s) {
        return cudaMemcpy2DAsync(layer(d), pitch, host, (size_t)w * sizeof(T), (size_t)w * sizeof(T), h,
                                 cudaMemcpyDefault, s);
    }
    cudaError_t download_layer(T* host, int d, cudaStream_t s) const {
        return cudaMemcpy2DAsync(host, (size_t)w * sizeof(T), layer(d), pitch, (size_t)w * sizeof(T), h,
                                 cudaMemcpyDefault, s);
    }
};

// ---------------------------------------------------------------------------------------------------
// Plane<T>: dense layered array, row pitch (in elements) rounded up to 128 bytes.
// ---------------------------------------------------------------------------------------------------
template <typename T>
struct Plane {
    T* ptr = nullptr;
    int w = 0, h = 0, layers = 0;
    int pitch = 0;  // elements

    int ensure(int w_, int h_, int layers_, bool lazy_layers) {
        if ((w_ == w && h_ == h && layers_ == layers) || (lazy_layers && w_ == w && h_ == h && layers_ <= layers))
            return 0;
        release();
        if (w_ <= 0 || h_ <= 0 || layers_ <= 0) return 1;
        const int per128 = 128 / (int)sizeof(T) > 0 ? 128 / (int)sizeof(T) : 1;
        int p = VB_DIV_CEIL(w_, per128) * per128;
        if (cudaMalloc((void**)&ptr, (size_t)p * h_ * layers_ * sizeof(T)) != cudaSuccess) {
            ptr = nullptr;
            return 1;
        }
        w = w_, h = h_, layers = layers_, pitch = p;
        return 1;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, w = h = layers = pitch = 0;
    }
    size_t layer_elems() const { return (size_t)pitch * h; }
    T* layer(int d) const { return ptr + (size_t)d * layer_elems(); }
    cudaError_t upload_layer(const T* host, int d, cudaStream_t s) {
        return cudaMemcpy2DAsync(layer(d), (size_t)pitch * sizeof(T), host, (size_t)w * sizeof(T),
                                 (size_t)w * sizeof(T), h, cudaMemcpyDefault, s);
    }
    cudaError_t download_layer(T* host, int d, cudaStream_t s) const {
        return cudaMemcpy2DAsync(host, (size_t)w * sizeof(T), layer(d), (size_t)pitch * sizeof(T),
                                 (size_t)w * sizeof(T), h, cudaMemcpyDefault, s);
    }
};

// Camera block handed to kernels by value (kernel parameter space = constant bank, dynamically indexable).
// Layout follows what the reference keeps in __constant__ memory (optimize_depth.cu:24-35).
struct CamBlock {
    float K4[4];      // fx, cx, fy, cy
    float K4inv[4];   // 1/fx, -cx/fx, 1/fy, -cy/fy
    float R[kMaxFrames][9];
    float t[kMaxFrames][3];
};
struct PriorCamBlock {
    float R[kMaxPriorFrames][9];
    float t[kMaxPriorFrames][3];
};

inline void fill_K(CamBlock& cb, const float* h_K) {
    // reference: optimize_depth.cu:345-350 — K4 = {fx,cx,fy,cy}, K4_inv = {1/fx, -cx/fx, 1/fy, -cy/fy}
    cb.K4[0] = h_K[0], cb.K4[1] = h_K[2], cb.K4[2] = h_K[4], cb.K4[3] = h_K[5];
    cb.K4inv[0] = 1.f / h_K[0], cb.K4inv[1] = -h_K[2] / h_K[0], cb.K4inv[2] = 1.f / h_K[4],
    cb.K4inv[3] = -h_K[5] / h_K[4];
}

}  // namespace vb
