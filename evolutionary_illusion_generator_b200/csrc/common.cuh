// Shared definitions for the EIGen B200 fitness engine kernels.
#pragma once
#ifndef EIG_EMU
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#define EIG_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define EIG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace eig {

// launch counter: bench.py reports how many of OUR kernels ran inside the timed region
struct LaunchCounter { long long n; };
inline LaunchCounter& launch_counter() { static LaunchCounter c{0}; return c; }
#define EIG_COUNT_LAUNCH() (++::eig::launch_counter().n)

// View of an NHWC fp32 activation tensor living inside a wider "concat" buffer.
// When `lo` is non-null the value is stored split for the 3xTF32 tensor-core path:
// hi = tf32-rounded value, lo = value - hi (exact), so hi + lo reconstructs the fp32 value bit-exactly.
struct View {
    float* hi;
    float* lo;
    int pitch;  // floats per pixel of the underlying buffer
    int coff;   // first channel of this view
    int C;      // channels in this view
};

__device__ __forceinline__ float tf32_round(float v) {
#ifdef EIG_EMU
    unsigned u = __float_as_uint(v);
    u = (u + 0x1000u) & ~0x1fffu;
    return __uint_as_float(u);
#else
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
#endif
}

__device__ __forceinline__ void view_store(const View& v, long long pix, int c, float val) {
    long long idx = pix * v.pitch + v.coff + c;
    if (v.lo) {
        float h = tf32_round(val);
        v.hi[idx] = h;
        v.lo[idx] = __fsub_rn(val, h);
    } else {
        v.hi[idx] = val;
    }
}

__device__ __forceinline__ float view_load(const float* hi, const float* lo, long long idx) {
    float v = hi[idx];
    if (lo) v = __fadd_rn(v, lo[idx]);
    return v;
}

}  // namespace eig
