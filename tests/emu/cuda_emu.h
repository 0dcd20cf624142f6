// TEST-ONLY kernel-source emulator.  NOT part of the product, never linked into libeig.so.
//
// The build container has no GPU.  To exercise the *actual* SIMT kernel sources (csrc/*.cuh: indexing,
// shared-memory staging, barriers, warp shuffles) under `pytest -m "not gpu"`, tests/emu/build_emu.sh
// compiles them with g++ and this shim instead of nvcc.  Every CUDA thread of a block runs as a ucontext
// fiber; __syncthreads / __syncwarp / __shfl_*_sync / __ballot_sync yield to a scheduler that releases a
// barrier once every live participant has arrived.  Blocks run one after the other, so `__shared__`
// variables are plain statics and atomics are plain read-modify-writes.  tcgen05/TMA kernels are not
// emulated (they are compiled out under EIG_EMU) and the numbers produced here are never reported.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define EIG_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct double2 { double x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

typedef int cudaError_t;
typedef int cudaStream_t;
typedef struct { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace emu {
enum { RUN = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
    ucontext_t ctx;
    int state;
    unsigned wait_mask;
    uint3 tid;
    int linear;
    uint64_t xchg;
    int pred;
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<char*> stacks;
    ucontext_t sched;
    Fiber* cur = nullptr;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    unsigned char* dyn_smem = nullptr;
    size_t dyn_cap = 0;
    std::function<void()> body;
};
inline State& st() { static State s; return s; }
static const size_t kStack = 256 * 1024;

inline void trampoline() {
    State& s = st();
    s.body();
    s.cur->state = DONE;
    swapcontext(&s.cur->ctx, &s.sched);
}
inline void yield_to_sched() {
    State& s = st();
    swapcontext(&s.cur->ctx, &s.sched);
}
inline void block_barrier() {
    st().cur->state = WAIT_BLOCK;
    yield_to_sched();
}
inline void warp_barrier(unsigned mask) {
    State& s = st();
    s.cur->state = WAIT_WARP;
    s.cur->wait_mask = mask;
    yield_to_sched();
}
inline void run_block(int nthreads) {
    State& s = st();
    if ((int)s.fibers.size() < nthreads) {
        s.fibers.resize(nthreads);
        while ((int)s.stacks.size() < nthreads) s.stacks.push_back((char*)malloc(kStack));
    }
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = s.fibers[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = s.stacks[t];
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
        f.state = RUN;
        f.linear = t;
        f.tid.x = t % s.bdim.x;
        f.tid.y = (t / s.bdim.x) % s.bdim.y;
        f.tid.z = t / (s.bdim.x * s.bdim.y);
    }
    int live = nthreads;
    while (live > 0) {
        bool progressed = false;
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            if (f.state != RUN) continue;
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
            progressed = true;
            if (f.state == DONE) --live;
        }
        // release warp barriers
        int nwarps = (nthreads + 31) / 32;
        for (int w = 0; w < nwarps; ++w) {
            int lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
            unsigned mask = 0;
            bool any_wait = false;
            for (int t = lo; t < hi; ++t)
                if (s.fibers[t].state == WAIT_WARP) { any_wait = true; mask = s.fibers[t].wait_mask; break; }
            if (!any_wait) continue;
            bool all = true;
            for (int t = lo; t < hi; ++t) {
                if (!((mask >> (t - lo)) & 1u)) continue;
                int stt = s.fibers[t].state;
                if (stt == DONE) continue;
                if (stt != WAIT_WARP) { all = false; break; }
            }
            if (all) {
                for (int t = lo; t < hi; ++t)
                    if (s.fibers[t].state == WAIT_WARP) s.fibers[t].state = RUN;
                progressed = true;
            }
        }
        // release the block barrier
        bool all_block = live > 0;
        for (int t = 0; t < nthreads && all_block; ++t) {
            int stt = s.fibers[t].state;
            if (stt != DONE && stt != WAIT_BLOCK) all_block = false;
        }
        if (all_block) {
            for (int t = 0; t < nthreads; ++t)
                if (s.fibers[t].state == WAIT_BLOCK) s.fibers[t].state = RUN;
            progressed = true;
        }
        if (!progressed && live > 0) {
            fprintf(stderr, "cuda_emu: deadlock (divergent barrier?) in block (%u,%u,%u)\n", s.bid.x, s.bid.y, s.bid.z);
            abort();
        }
    }
}
template <class K, class... A>
inline void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    State& s = st();
    s.bdim = block;
    s.gdim = grid;
    if (smem > s.dyn_cap) {
        free(s.dyn_smem);
        s.dyn_smem = (unsigned char*)malloc(smem);
        s.dyn_cap = smem;
    }
    s.body = [=]() { kernel(args...); };
    int nthreads = block.x * block.y * block.z;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                s.bid = uint3{x, y, z};
                run_block(nthreads);
            }
}
inline Fiber& lane_fiber(int lane) {
    State& s = st();
    int base = (s.cur->linear / 32) * 32;
    return s.fibers[base + lane];
}
inline int lane_count() {
    State& s = st();
    int n = s.bdim.x * s.bdim.y * s.bdim.z;
    int base = (s.cur->linear / 32) * 32;
    return n - base < 32 ? n - base : 32;
}
template <class T>
inline T shfl_from(unsigned mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    State& s = st();
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    s.cur->xchg = bits;
    warp_barrier(mask);
    T out = v;
    if (src >= 0 && src < lane_count() && ((mask >> src) & 1u)) {
        uint64_t b = lane_fiber(src).xchg;
        memcpy(&out, &b, sizeof(T));
    }
    warp_barrier(mask);
    return out;
}
}  // namespace emu

#define threadIdx (emu::st().cur->tid)
#define blockIdx (emu::st().bid)
#define blockDim (emu::st().bdim)
#define gridDim (emu::st().gdim)
#define warpSize 32

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
static inline int emu_lane() { return emu::st().cur->linear % 32; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
    int lane = emu_lane();
    return emu::shfl_from(m, v, (lane / width) * width + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
    (void)width;
    return emu::shfl_from(m, v, emu_lane() ^ x);
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int src = lane + (int)d;
    if (src / width != lane / width) src = lane;
    return emu::shfl_from(m, v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int src = lane - (int)d;
    if (src < 0 || src / width != lane / width) src = lane;
    return emu::shfl_from(m, v, src);
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
    emu::State& s = emu::st();
    s.cur->pred = pred ? 1 : 0;
    emu::warp_barrier(m);
    unsigned out = 0;
    int n = emu::lane_count();
    for (int l = 0; l < n; ++l)
        if (((m >> l) & 1u) && emu::lane_fiber(l).state != emu::DONE && emu::lane_fiber(l).pred) out |= 1u << l;
    emu::warp_barrier(m);
    return out;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    unsigned b = __ballot_sync(m, p);
    unsigned live = 0;
    int n = emu::lane_count();
    for (int l = 0; l < n; ++l) if ((m >> l) & 1u) live |= 1u << l;
    return b == live;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
static inline int __double2int_rz(double a) { return (int)a; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __ll2float_rn(long long a) { return (float)a; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }

#define EIG_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define EIG_DYN_SMEM(name) unsigned char* name = emu::st().dyn_smem
