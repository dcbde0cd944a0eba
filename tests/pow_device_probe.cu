// TEST INFRASTRUCTURE.  Exhaustive check of the lean power function of the Fisk pdf (csrc/residual_model.cuh:
// log2_parts / exp2_scaled = the main path of the CUDA math library's powf without its special-case code) against the
// library's powf: EVERY normal positive float as base, for each exponent the residual model can produce.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tests/_build/libpow_probe.so tests/pow_device_probe.cu
#include <cuda_runtime.h>
#include "../voldor_b200/csrc/residual_model.cuh"

__global__ void k_pow_probe(const float* ys, int n_y, unsigned long long* mismatches, unsigned* first_bad) {
    // bases: all bit patterns 0x00800000 .. 0x7f7fffff
    const unsigned long long total = 0x7f000000ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)(0x00800000ull + i));
        const vb::Log2Parts L = vb::log2_parts(x);
        for (int k = 0; k < n_y; k++) {
            const float y = ys[k];
            const float mine = vb::exp2_scaled(L, y);
            const float lib = powf(x, y);
            if (__float_as_uint(mine) != __float_as_uint(lib)) {
                if (atomicAdd(&mismatches[k], 1ull) == 0) first_bad[k] = __float_as_uint(x);
            }
        }
    }
}

// also the composed pdf against a copy that calls the library everywhere (random residuals / shapes)
__device__ float fisk_pdf_library(float residual, vb::FiskShape k) {
    const float x = fmaxf(vb::f_mul(residual, 0.5f), FLT_EPSILON);
    const float q = vb::f_div(vb::f_mul(x, x), k.s);
    const float a = powf(q, vb::f_sub(-1.f, k.c)), t = powf(q, -k.c);
    const float b = powf(vb::f_add(t, 1.0f), -2.f);
    return vb::f_div(vb::f_mul(vb::f_mul(k.c, a), b), k.s);
}
__global__ void k_pdf_probe(unsigned long long n, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + (unsigned)(i >> 32) * 40503u + 12345u;
        h ^= h >> 15, h *= 0x2c1b3c6du, h ^= h >> 12, h *= 0x297a2d39u, h ^= h >> 15;
        unsigned g = h * 0x9e3779b9u + 7u;
        g ^= g >> 16, g *= 0x85ebca6bu, g ^= g >> 13;
        // residual: any float bit pattern (incl. NaN/inf/denormals/negatives); magnitude for the shape in [0, 300]
        const float residual = __uint_as_float(h);
        const float mag = (float)(g >> 8) * (300.f / 16777216.f);
        const vb::FiskShape k = vb::fisk_shape_scale(mag);
        if (__float_as_uint(vb::fisk_pdf(residual, k)) != __float_as_uint(fisk_pdf_library(residual, k))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" int pow_probe_run(const float* h_ys, int n_y, unsigned long long* h_mismatches, unsigned* h_first_bad,
                             unsigned long long pdf_samples, unsigned long long* h_pdf_mismatches) {
    float* ys;
    unsigned long long* mm;
    unsigned* fb;
    cudaMalloc((void**)&ys, n_y * sizeof(float));
    cudaMalloc((void**)&mm, (n_y + 1) * sizeof(unsigned long long));
    cudaMalloc((void**)&fb, n_y * sizeof(unsigned));
    cudaMemcpy(ys, h_ys, n_y * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(mm, 0, (n_y + 1) * sizeof(unsigned long long));
    cudaMemset(fb, 0, n_y * sizeof(unsigned));
    k_pow_probe<<<148 * 16, 256>>>(ys, n_y, mm, fb);
    k_pdf_probe<<<148 * 16, 256>>>(pdf_samples, mm + n_y);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_mismatches, mm, n_y * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_pdf_mismatches, mm + n_y, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_first_bad, fb, n_y * sizeof(unsigned), cudaMemcpyDeviceToHost);
    cudaFree(ys), cudaFree(mm), cudaFree(fb);
    return (int)e;
}
