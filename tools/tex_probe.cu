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
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

constexpr int kVariants = 16;
static const char* kNames[kVariants] = {
    "8 fraction bits (the documented 1.8 fixed point), exact 4-term sum rounded once",
    "8 fraction bits, two-stage fp32 lerp (x then y) with fma",
    "9 fraction bits, exact sum rounded once",
    "10 fraction bits, exact sum rounded once",
    "12 fraction bits, exact sum rounded once",
    "16 fraction bits, exact sum rounded once",
    "full fp32 fractions a=c-0.5-floor(c-0.5), exact sum rounded once",
    "full fp32 fractions, two-stage fp32 lerp with fma: t=fma(a,T10-T00,T00)",
    "full fp32 fractions, two-stage fp32 lerp: t=fma(a,T10,fma(-a,T00,T00))",
    "full fp32 fractions, fp32 weights (1-a)(1-b).. fma chain",
    "9 fraction bits, two-stage fp32 lerp with fma",
    "10 fraction bits, two-stage fp32 lerp with fma",
    "12 fraction bits, two-stage fp32 lerp with fma",
    "16 fraction bits, two-stage fp32 lerp with fma",
    "23 fraction bits (rint(c*2^23)), exact sum rounded once",
    "8 fraction bits truncated (floor), two-stage fp32 lerp with fma",
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
    static const int kBits[kVariants] = {8, 8, 9, 10, 12, 16, 0, 0, 0, 0, 9, 10, 12, 16, 23, -8};
    static const int kForm[kVariants] = {0, 1, 0, 0, 0, 0, 0, 1, 2, 3, 1, 1, 1, 1, 0, 1};
    const int bits = kBits[variant], form = kForm[variant];
    int ix, iy;
    float a, b;
    if (bits == 0) {
        const float bx = cx - 0.5f, by = cy - 0.5f;
        const float fxl = floorf(bx), fyl = floorf(by);
        ix = (int)fxl, iy = (int)fyl;
        a = bx - fxl, b = by - fyl;
    } else {
        const int nb = bits < 0 ? -bits : bits;
        const double sc = (double)(1ll << nb);
        const long long qx = (bits < 0 ? (long long)floor((double)cx * sc) : (long long)rint((double)cx * sc)) - (1ll << (nb - 1));
        const long long qy = (bits < 0 ? (long long)floor((double)cy * sc) : (long long)rint((double)cy * sc)) - (1ll << (nb - 1));
        ix = (int)(qx >> nb), iy = (int)(qy >> nb);
        a = (float)((double)(qx & ((1ll << nb) - 1)) / sc), b = (float)((double)(qy & ((1ll << nb) - 1)) / sc);
    }
    const float2 t00 = texel(P, ix, iy), t10 = texel(P, ix + 1, iy), t01 = texel(P, ix, iy + 1), t11 = texel(P, ix + 1, iy + 1);
    const float v00 = comp ? t00.y : t00.x, v10 = comp ? t10.y : t10.x, v01 = comp ? t01.y : t01.x, v11 = comp ? t11.y : t11.x;
    if (form == 0) {
        const double da = a, db = b;
        const double s = (1 - da) * (1 - db) * v00 + da * (1 - db) * v10 + (1 - da) * db * v01 + da * db * v11;
        return __double2float_rn(s);
    } else if (form == 1) {
        const float top = __fmaf_rn(a, __fsub_rn(v10, v00), v00), bot = __fmaf_rn(a, __fsub_rn(v11, v01), v01);
        return __fmaf_rn(b, __fsub_rn(bot, top), top);
    } else if (form == 2) {
        const float top = __fmaf_rn(a, v10, __fmaf_rn(-a, v00, v00)), bot = __fmaf_rn(a, v11, __fmaf_rn(-a, v01, v01));
        return __fmaf_rn(b, bot, __fmaf_rn(-b, top, top));
    } else {
        const float w00 = __fmul_rn(1.f - a, 1.f - b), w10 = __fmul_rn(a, 1.f - b), w01 = __fmul_rn(1.f - a, b), w11 = __fmul_rn(a, b);
        return __fmaf_rn(w11, v11, __fmaf_rn(w01, v01, __fmaf_rn(w10, v10, __fmul_rn(w00, v00))));
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

// ---- weight sweep: read the filter weight the hardware applies as a function of the coordinate -------------------
// texels are 0 everywhere except one column (x sweep) / one row (y sweep) of ones, so a fetch half a texel around it
// returns the weight itself.
__global__ void k_weights(cudaTextureObject_t tex, float x0, float y0, int n, float* out_x, float* out_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = (float)i / (float)n;  // [0, 1)
    out_x[i] = tex2D<float2>(tex, x0 + d, y0).x;   // alpha(x)
    out_y[i] = tex2D<float2>(tex, x0, y0 + d).y;   // beta(y)
}
static int weight_sweep(const char* path) {
    const int w = 256, H = 256, n = 1 << 16;
    float2* d = nullptr;
    size_t pitch = 0;
    CK(cudaMallocPitch((void**)&d, &pitch, (size_t)w * sizeof(float2), (size_t)H));
    std::vector<float2> host((size_t)w * H, make_float2(0.f, 0.f));
    for (int y = 0; y < H; y++) host[(size_t)y * w + 101].x = 1.f;  // column of ones in .x
    for (int x = 0; x < w; x++) host[(size_t)51 * w + x].y = 1.f;   // row of ones in .y
    CK(cudaMemcpy2D(d, pitch, host.data(), (size_t)w * sizeof(float2), (size_t)w * sizeof(float2), H, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float2>();
    rd.res.pitch2D.devPtr = d, rd.res.pitch2D.width = w, rd.res.pitch2D.height = H, rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear, td.readMode = cudaReadModeElementType, td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    float *ox, *oy;
    CK(cudaMalloc((void**)&ox, n * sizeof(float)));
    CK(cudaMalloc((void**)&oy, n * sizeof(float)));
    // texel 100 has its centre at 100.5: sweeping the coordinate over [100.5, 101.5) moves the weight of texel 101 over [0, 1)
    k_weights<<<(n + 255) / 256, 256>>>(tex, 100.5f, 50.5f, n, ox, oy);
    CK(cudaDeviceSynchronize());
    std::vector<float> hx(n), hy(n);
    CK(cudaMemcpy(hx.data(), ox, n * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hy.data(), oy, n * sizeof(float), cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "wb");
    if (!f) return 3;
    fwrite(hx.data(), sizeof(float), n, f);
    fwrite(hy.data(), sizeof(float), n, f);
    fclose(f);
    int steps = 0;
    for (int i = 1; i < n; i++) steps += hx[i] != hx[i - 1];
    printf("{\"weight_sweep\": \"%s\", \"samples\": %d, \"distinct_steps_x\": %d, \"alpha_at_0\": %.9g, \"alpha_at_quarter\": %.9g, \"alpha_last\": %.9g}\n",
           path, n, steps, hx[0], hx[n / 4], hx[n - 1]);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 2 && !strcmp(argv[1], "weights")) return weight_sweep(argv[2]);
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
