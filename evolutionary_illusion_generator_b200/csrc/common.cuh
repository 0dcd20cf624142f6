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
// Programmatic dependent launch (sm_90+): a kernel launched with EIG_LAUNCH_PDL may start while its predecessor in the
// stream is still draining; it must execute EIG_PDL_WAIT() before it touches anything the predecessor reads or writes
// (weights are constants and may be staged earlier), and it lets its own successor start early with EIG_PDL_TRIGGER().
#define EIG_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define EIG_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
template <class... KArgs, class... Args>
inline cudaError_t eig_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define EIG_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) eig_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#else
#define EIG_PDL_TRIGGER() do { } while (0)
#define EIG_PDL_WAIT() do { } while (0)
#define EIG_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) EIG_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#endif

namespace eig {

// launch counter: bench.py reports how many of OUR kernels ran inside the timed region
struct LaunchCounter { long long n; };
inline LaunchCounter& launch_counter() { static LaunchCounter c{0}; return c; }
#define EIG_COUNT_LAUNCH() (++::eig::launch_counter().n)

// View of an NHWC activation tensor living inside a wider "concat" buffer.
//   lo == nullptr : plain fp32 at `hi` (exact-fp32 SIMT mode).
//   lo != nullptr : SPLIT-FP16 storage for the tensor-core path: `hi` and `lo` point to two fp16 planes of the same
//                   geometry holding hi = fp16(16 v) and lo = fp16(16 v - hi), i.e. 22 significant bits of v in the same
//                   4 bytes per element.  The planes are exactly the MMA operands (conv_tc.cuh loads them with TMA
//                   straight into swizzled shared memory); every other reader uses view_load (value = (hi + lo) / 16).
struct View {
    float* hi;
    float* lo;
    int pitch;  // elements per pixel of the underlying buffer
    int coff;   // first channel of this view
    int C;      // channels in this view
};

typedef unsigned short h16;   // raw IEEE binary16 bits
#define EIG_ACT_SCALE 16.0f
#define EIG_ACT_INV 0.0625f

#ifdef EIG_EMU
// software binary16 conversions for the g++ kernel-source emulator (round to nearest even, subnormals, inf)
inline h16 f2h(float f) {
    unsigned x;
    memcpy(&x, &f, 4);
    const unsigned sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (h16)(sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u));
    if (x >= 0x477ff000u) return (h16)(sign | 0x7c00u);   // >= 65520 rounds to infinity
    if (x < 0x38800000u) {                                // below 2^-14: subnormal half (or zero)
        float af;
        memcpy(&af, &x, 4);
        return (h16)(sign | (unsigned)lrintf(af * 16777216.0f));
    }
    const unsigned mant = x & 0x7fffffu, ex = (x >> 23) - 112u;
    unsigned h = (ex << 10) | (mant >> 13);
    const unsigned rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (h16)(sign | h);
}
inline float h2f(h16 h) {
    const unsigned sign = ((unsigned)h & 0x8000u) << 16, ex = (h >> 10) & 0x1fu, mant = h & 0x3ffu;
    float r;
    if (ex == 0) r = ldexpf((float)mant, -24);
    else if (ex == 31) r = mant ? NAN : INFINITY;
    else r = ldexpf((float)(mant | 0x400u), (int)ex - 25);
    unsigned u;
    memcpy(&u, &r, 4);
    u |= sign;
    memcpy(&r, &u, 4);
    return r;
}
#else
}  // namespace eig
#include <cuda_fp16.h>
namespace eig {
__host__ __device__ __forceinline__ h16 f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
__host__ __device__ __forceinline__ float h2f(h16 h) { return __half2float(__ushort_as_half(h)); }
#endif

// Split-fp16 storage covers |v| < 65504 / 16 = 4094.  PredNet activations are O(1); a weight file that drives them out of
// range (or to NaN) raises this flag instead of silently producing inf - inf: the scoring kernel hands it to the host
// with the fitness vector and eig_eval_host fails with EIG_E_RANGE (use conv_mode SIMT for such a model).
#ifdef EIG_EMU
static int g_eig_range_flag = 0;
#define EIG_NOTE_RANGE(x) do { if (!(fabsf(x) <= 65504.f)) g_eig_range_flag = 1; } while (0)
#else
__device__ int g_eig_range_flag = 0;
#ifdef __CUDA_ARCH__
#define EIG_NOTE_RANGE(x) do { if (!(fabsf(x) <= 65504.f)) g_eig_range_flag = 1; } while (0)
#else
#define EIG_NOTE_RANGE(x) do { } while (0)
#endif
#endif

__host__ __device__ __forceinline__ void split16(float v, h16* h, h16* l) {
    const float x = v * EIG_ACT_SCALE;        // exact (power of two)
    EIG_NOTE_RANGE(x);
    const h16 hh = f2h(x);
    *h = hh;
    *l = f2h(x - h2f(hh));                    // the difference is exact in fp32
}
__host__ __device__ __forceinline__ float join16(h16 h, h16 l) { return (h2f(h) + h2f(l)) * EIG_ACT_INV; }   // exact sum

__device__ __forceinline__ void view_store(const View& v, long long pix, int c, float val) {
    const long long idx = pix * v.pitch + v.coff + c;
    if (v.lo) split16(val, reinterpret_cast<h16*>(v.hi) + idx, reinterpret_cast<h16*>(v.lo) + idx);
    else v.hi[idx] = val;
}

// four consecutive channels (c a multiple of 4); falls back to scalar stores when the view is not vector aligned
__device__ __forceinline__ void view_store_vec4(const View& v, long long pix, int c, const float* val) {
    const long long idx = pix * v.pitch + v.coff + c;
    if ((v.pitch | v.coff | c) & 3) {
        for (int i = 0; i < 4; ++i) view_store(v, pix, c + i, val[i]);
    } else if (v.lo) {
        h16 h[4], l[4];
        for (int i = 0; i < 4; ++i) split16(val[i], &h[i], &l[i]);
        uint2 uh, ul;
        uh.x = (unsigned)h[0] | ((unsigned)h[1] << 16); uh.y = (unsigned)h[2] | ((unsigned)h[3] << 16);
        ul.x = (unsigned)l[0] | ((unsigned)l[1] << 16); ul.y = (unsigned)l[2] | ((unsigned)l[3] << 16);
        *reinterpret_cast<uint2*>(reinterpret_cast<h16*>(v.hi) + idx) = uh;
        *reinterpret_cast<uint2*>(reinterpret_cast<h16*>(v.lo) + idx) = ul;
    } else {
        *reinterpret_cast<float4*>(v.hi + idx) = make_float4(val[0], val[1], val[2], val[3]);
    }
}

__device__ __forceinline__ float view_load(const float* hi, const float* lo, long long idx) {
    if (lo) return join16(reinterpret_cast<const h16*>(hi)[idx], reinterpret_cast<const h16*>(lo)[idx]);
    return hi[idx];
}

}  // namespace eig
