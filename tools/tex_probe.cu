// tex_probe — does a software bilinear lerp reproduce tex2D bit for bit?
//
// The depth EM and the collector fetch flows / priors at data-dependent fractional positions through ONE pitch2D
// texture over all layers stacked in y (csrc/common.cuh TexStack, reference gpu-kernels/gmat.h:39-66,175-179): linear
// filter, clamp addressing, unnormalised coordinates.  The texture unit filters with 1.8 fixed-point weights, and the
// stacking makes rows h-1..h of layer d blend with row 0 of layer d+1.  Staging those streams through TMA + shared
// memory instead (BASELINE.json north_star) is only parity-safe if a software lerp gives identical bits, so this
// probe compares tex2D against candidate emulations over >= 1e8 coordinates (uniform over the stack plus dense
// samples around the layer seams and the outer borders) and prints the mismatch count of every variant.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/tex_probe tools/tex_probe.cu
//   tools/_build/tex_probe [samples_log2=27]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

constexpr int kVariants = 8;
static const char* kNames[kVariants] = {
    "q=rint(c*256)-128, exact 4-term sum rounded once (RN)",
    "q=rint(c*256)-128, exact 4-term sum truncated (RZ)",
    "q=rint(c*256)-128, fp32 fma chain T00,T10,T01,T11",
    "q=rint(c*256)-128, two-stage fp32 lerp (x then y) with fma",
    "q=floor(c*256)-128, exact 4-term sum rounded once (RN)",
    "q=rint((c-0.5)*256), exact 4-term sum rounded once (RN)",
    "q=rint(c*256)-128, two-stage lerp in double, rounded once",
    "q=rint(c*256)-128, fp32 (1-a)*(1-b)*T.. products summed left to right",
};

struct Probe {
    const float2* data;
    size_t pitch;  // bytes
    int w, H;      // H = h * layers (stack height)
    cudaTextureObject_t tex;
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float2 texel(const Probe& P, int x, int y) {
    x = min(max(x, 0), P.w - 1);
    y = min(max(y, 0), P.H - 1);
    return *(const float2*)((const char*)P.data + (size_t)y * P.pitch + (size_t)x * sizeof(float2));
}

__device__ float emulate(int variant, const Probe& P, float cx, float cy, int comp) {
    long qx, qy;
    if (variant == 4) {
        qx = (long)floorf(cx * 256.f) - 128, qy = (long)floorf(cy * 256.f) - 128;
    } else if (variant == 5) {
        qx = (long)rintf((cx - 0.5f) * 256.f), qy = (long)rintf((cy - 0.5f) * 256.f);
    } else {
        qx = (long)rintf(cx * 256.f) - 128, qy = (long)rintf(cy * 256.f) - 128;
    }
    const int ix = (int)(qx >> 8), iy = (int)(qy >> 8);
    const float a = (float)(qx & 255) * (1.f / 256.f), b = (float)(qy & 255) * (1.f / 256.f);
    const float2 t00 = texel(P, ix, iy), t10 = texel(P, ix + 1, iy), t01 = texel(P, ix, iy + 1), t11 = texel(P, ix + 1, iy + 1);
    const float v00 = comp ? t00.y : t00.x, v10 = comp ? t10.y : t10.x, v01 = comp ? t01.y : t01.x, v11 = comp ? t11.y : t11.x;
    const float w00 = (1.f - a) * (1.f - b), w10 = a * (1.f - b), w01 = (1.f - a) * b, w11 = a * b;  // exact in fp32
    switch (variant) {
        case 0: case 4: case 5: {
            const double s = (double)w00 * v00 + (double)w10 * v10 + (double)w01 * v01 + (double)w11 * v11;
            return __double2float_rn(s);
        }
        case 1: {
            const double s = (double)w00 * v00 + (double)w10 * v10 + (double)w01 * v01 + (double)w11 * v11;
            return __double2float_rz(s);
        }
        case 2: return __fmaf_rn(w11, v11, __fmaf_rn(w01, v01, __fmaf_rn(w10, v10, __fmul_rn(w00, v00))));
        case 3: {
            const float top = __fmaf_rn(a, __fsub_rn(v10, v00), v00), bot = __fmaf_rn(a, __fsub_rn(v11, v01), v01);
            return __fmaf_rn(b, __fsub_rn(bot, top), top);
        }
        case 6: {
            const double top = (double)v00 + (double)a * ((double)v10 - (double)v00);
            const double bot = (double)v01 + (double)a * ((double)v11 - (double)v01);
            return __double2float_rn(top + (double)b * (bot - top));
        }
        default:
            return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, v00), __fmul_rn(w10, v10)), __fmul_rn(w01, v01)), __fmul_rn(w11, v11));
    }
}

__global__ void k_probe(Probe P, int h, unsigned long long n, unsigned long long* mismatches, float* examples) {
    unsigned long long local[kVariants] = {0};
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t r0 = hash32((uint32_t)i * 2654435761u + 1u), r1 = hash32(r0 ^ (uint32_t)(i >> 32) ^ 0x9e3779b9u), r2 = hash32(r1 + 77u);
        float x = -2.f + (float)(r0 >> 8) * (1.f / 16777216.f) * (float)(P.w + 4);
        float y;
        const uint32_t mode = r2 & 7u;
        if (mode < 5) {
            y = -2.f + (float)(r1 >> 8) * (1.f / 16777216.f) * (float)(P.H + 4);               // anywhere in the stack
        } else if (mode < 7) {
            const int layer = 1 + (int)((r2 >> 3) % (uint32_t)(P.H / h - 1));
            y = (float)(layer * h) - 1.5f + (float)(r1 >> 8) * (1.f / 16777216.f) * 2.f;         // around a layer seam
        } else {
            y = ((r2 >> 3) & 1u) ? -1.f + (float)(r1 >> 8) * (1.f / 16777216.f) * 2.f           // top border
                                 : (float)P.H - 1.5f + (float)(r1 >> 8) * (1.f / 16777216.f) * 2.f;  // bottom border
            if ((r2 >> 4) & 1u) x = ((r2 >> 5) & 1u) ? -1.f + (float)(r0 >> 8) * (1.f / 16777216.f) * 2.f
                                                     : (float)P.w - 1.5f + (float)(r0 >> 8) * (1.f / 16777216.f) * 2.f;
        }
        // the kernels pass (x + 0.5, d*h + y + 0.5): emulate the same rounded sums
        const float cx = __fadd_rn(x, 0.5f), cy = __fadd_rn(y, 0.5f);
        const float2 t = tex2D<float2>(P.tex, cx, cy);
#pragma unroll
        for (int v = 0; v < kVariants; v++) {
            const float ex = emulate(v, P, cx, cy, 0), ey = emulate(v, P, cx, cy, 1);
            const bool bad = __float_as_uint(ex) != __float_as_uint(t.x) || __float_as_uint(ey) != __float_as_uint(t.y);
            if (bad) {
                local[v]++;
                if (v == 0) {
                    const unsigned long long slot = atomicAdd(&mismatches[kVariants], 1ull);
                    if (slot < 8) {
                        float* e = examples + slot * 6;
                        e[0] = cx, e[1] = cy, e[2] = t.x, e[3] = ex, e[4] = t.y, e[5] = ey;
                    }
                }
            }
        }
    }
    for (int v = 0; v < kVariants; v++)
        if (local[v]) atomicAdd(&mismatches[v], local[v]);
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 27;
    const unsigned long long n = 1ull << lg;
    const int w = 640, h = 480, layers = 8, H = h * layers;
    Probe P;
    P.w = w, P.H = H;
    float2* d = nullptr;
    CK(cudaMallocPitch((void**)&d, &P.pitch, (size_t)w * sizeof(float2), (size_t)H));
    std::vector<float2> host((size_t)w * H);
    uint32_t s = 12345u;
    for (auto& v : host) {
        s = s * 1664525u + 1013904223u; v.x = ((int)(s >> 8) - (1 << 23)) * (1.f / (1 << 18));   // flows of +-32 px
        s = s * 1664525u + 1013904223u; v.y = ((int)(s >> 8) - (1 << 23)) * (1.f / (1 << 20));
    }
    CK(cudaMemcpy2D(d, P.pitch, host.data(), (size_t)w * sizeof(float2), (size_t)w * sizeof(float2), H, cudaMemcpyHostToDevice));
    P.data = d;
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float2>();
    rd.res.pitch2D.devPtr = d, rd.res.pitch2D.width = w, rd.res.pitch2D.height = H, rd.res.pitch2D.pitchInBytes = P.pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear, td.readMode = cudaReadModeElementType, td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&P.tex, &rd, &td, nullptr));
    unsigned long long* dm = nullptr;
    float* de = nullptr;
    CK(cudaMalloc((void**)&dm, (kVariants + 1) * sizeof(unsigned long long)));
    CK(cudaMemset(dm, 0, (kVariants + 1) * sizeof(unsigned long long)));
    CK(cudaMalloc((void**)&de, 8 * 6 * sizeof(float)));
    CK(cudaMemset(de, 0, 8 * 6 * sizeof(float)));
    k_probe<<<148 * 8, 256>>>(P, h, n, dm, de);
    CK(cudaDeviceSynchronize());
    unsigned long long hm[kVariants + 1];
    float he[48];
    CK(cudaMemcpy(hm, dm, sizeof(hm), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(he, de, sizeof(he), cudaMemcpyDeviceToHost));
    printf("{\"samples\": %llu, \"texture\": \"pitch2D float2 %dx%d, %d layers stacked in y, linear, clamp, unnormalised\", \"variants\": [\n", n, w, h, layers);
    for (int v = 0; v < kVariants; v++)
        printf("  {\"emulation\": \"%s\", \"mismatching_fetches\": %llu, \"fraction\": %.3e}%s\n", kNames[v], hm[v], (double)hm[v] / (double)n,
               v + 1 < kVariants ? "," : "");
    printf("], \"counter_examples_variant0\": [");
    const int ne = (int)(hm[kVariants] < 8 ? hm[kVariants] : 8);
    for (int k = 0; k < ne; k++)
        printf("%s{\"cx\": %.9g, \"cy\": %.9g, \"tex_x\": %.9g, \"emu_x\": %.9g, \"tex_y\": %.9g, \"emu_y\": %.9g}", k ? ", " : "", he[k * 6], he[k * 6 + 1],
               he[k * 6 + 2], he[k * 6 + 3], he[k * 6 + 4], he[k * 6 + 5]);
    printf("]}\n");
    return 0;
}
