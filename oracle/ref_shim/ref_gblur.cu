// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// extern "C" trampoline onto the reference's gblur_gpu(GMatf, GMatf&, sigma, ksize) (reference: gpu-kernels/gblur.cu:47-72,
// gblur.h:6).  That function has C++ linkage, takes the reference's own GMat<float> by value and is not declared in
// gpu_kernels.h, so the harness cannot reach it through ctypes: this shim builds the GMatf's from plain [depth][h][w]
// host buffers with the reference's own copy helpers (gmat.h:133-169).  gblur.h / gmat.h are taken from the reference tree
// by -I at build time (oracle/Makefile); nothing from the reference is copied into this repository.
#include "gblur.h"

extern "C" int ref_gblur_gpu(const float* h_src, float* h_dst, int w, int h, int depth, float sigma, int ksize) {
    GMatf src, dst;
    src.create(w, h, depth);
    src.copy_from_host(h_src, make_cudaPos(0, 0, 0), w, h, depth);
    const int rc = gblur_gpu(src, dst, sigma, ksize);
    if (rc == cudaSuccess) {
        cudaDeviceSynchronize();
        dst.copy_to_host(h_dst, make_cudaPos(0, 0, 0), w, h, depth);
    }
    src.free();
    dst.free();
    const cudaError_t e = cudaGetLastError();
    return rc != cudaSuccess ? rc : (int)e;
}
