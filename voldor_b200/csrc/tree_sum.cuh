// Warp-level restatement of the reference's block-tree vector sum.
//
// Behavioural source: reference gpu-kernels/reduce_vector_sum.h:3-57.  The reference reduces an array in
// passes of 512-element blocks (256 threads): s[t] = x[t] (+ x[t+256]); then s[t] += s[t+128], s[t+64]
// (shared memory, two barriers), then the 32-lane "warp_reduce" s[t] += s[t+32], +16, +8, +4, +2, +1; block
// results are appended and reduced again until one value remains.  Floating-point addition is not
// associative, so mean-shift / robust-fit iterates only match the reference if this exact tree is used
// (SURVEY §9 Q13).
//
// Here ONE WARP reduces one 512-element block with no shared memory and no barrier: lane t holds the 16
// elements t+32j; tree levels that pair indices 256/128/64/32 apart are lane-local register adds, the last
// five levels are shuffles.  The result (valid in lane 0) is bit-identical to the reference tree.
#pragma once
#include <cuda_runtime.h>

namespace vb {

// get(i): element i of this 512-block, called only for i < count (count in [1,512]).
template <class Get>
__device__ __forceinline__ float tree_sum_512(Get get, int count, int lane) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int t = lane + 32 * j;
        float v = 0.f;
        if (t < count) {
            v = get(t);
            if (t + 256 < count) v = __fadd_rn(v, get(t + 256));
        }
        a[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) a[j] = __fadd_rn(a[j], a[j + 4]);  // stride 128
    a[0] = __fadd_rn(a[0], a[2]);                                  // stride 64
    a[1] = __fadd_rn(a[1], a[3]);
    float v = __fadd_rn(a[0], a[1]);                               // stride 32
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

inline int tree_levels_ok(int n) { return n <= 512 * 512; }

}  // namespace vb
