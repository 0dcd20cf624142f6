// tcgen05 tensor-core 3x3 convolution for PredNet layers 1..3 (sm_100a): TMA-staged implicit GEMM, 3xTF32
// split-precision MMAs accumulating in TMEM, the three PredNet epilogues fused behind tcgen05.ld.
//
// Reference semantics are those of conv_simt.cuh (Chainer `L.Convolution2D(cin, cout, 3, pad=1)` +
// ConvLSTM / error-unit / pooling epilogues, /root/reference/chainer_prednet/PredNet/net.py:46-62,94-126,
// 187-209); the SIMT kernel is the exact-fp32 twin this one is checked against on the GPU (tests/gpu/tc_check.cu).
//
// GEMM view: D[m][n] = sum_{tap, c} A[m + shift(tap)][c] * Wt[tap][c][n]
//   m   = "flat padded" pixel index inside a CTA region: m = h * P + w, P = TW + 2.  The CTA loads ONE halo box
//         (KBT channels x P columns x NT*TH+2 rows) per channel block with a single 4-D TMA (negative / out of
//         range coordinates are zero-filled by the TMA unit = the conv's zero padding) and the nine taps are nine
//         row-shifted views of that box: tap (ky,kx) of MMA tile t starts (t*TH + ky) * P + kx rows into the box.
//         Rows with w >= TW are computed and dropped (2/P waste); the halo is fetched once instead of 9 times.
//   K   = KBT channels per block (16 -> 64-byte swizzle rows, 32 -> 128-byte), UMMA_K = 8
//   N   = output channels of this CTA (<= 256; the 4 gates of an LSTM cell are adjacent columns)
// 3xTF32: activations live in HBM as plain fp32; the converter warps split each staged box into hi = tf32(v) and
//   lo = v - hi in shared memory, weights are pre-split on the host; every k-step issues lo*hi + hi*lo + hi*hi into
//   the same fp32 TMEM accumulator (the lo*lo term, 2^-22 relative, is dropped).
#pragma once
#include <cuda.h>
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "common.cuh"
#include "conv_simt.cuh"

namespace eig {

enum { EPI_RAW = 3 };  // test only: out = acc + bias, no activation (conv3x3_tc_kernel only)
enum { TC_THREADS = 512, TC_SMEM_LIMIT = 227 * 1024 };

struct TcWeights {
    float* d = nullptr;  // [2 planes][9 taps][KBn][N][KBT] fp32 (hi plane, then lo plane)
    int cin = 0, N = 0, KBT = 16, KBn = 0, Ncta = 0, gz = 1;
    CUtensorMap map[3];  // weight-tile boxes of Ncta, Ncta/2, Ncta/4 rows (cluster size 1, 2, 4)
    int max_csize = 1;
    bool ok = false;
};

struct TcParams {
    int B, H, W;
    int TW, TH, P, NT;
    int tiles_x, tiles_y, regions;  // regions = spatial CTA regions (tiles_x * tiles_y * B)
    int csize, groups_per_nz, groups;  // cluster size, groups (= csize regions sharing one weight slice) per N slice, total
    int KBn, Ncta, N;
    int a_plane_bytes, b_plane_bytes, a_box_bytes;
    int SA, SB;
    int tmem_cols;
    int staging_bytes, stage_ld;  // ConvA pooling tile: 128 rows x stage_ld floats
    long long* dbg;               // optional [grid][16] cycle counters per role (tests/gpu/tc_check timing mode), else null
    ConvArgs ca;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// one elected lane of a converged warp; the surrounding code stays warp-uniform so that descriptors, barrier
// addresses and loop counters live in uniform registers (an `if (lane == 0)` around the issue loop makes ptxas wrap
// every tcgen05.mma in an R2UR "waterfall" of ~85 cycles - measured, see profiles/)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float lstm_cell_v(float gi, float gf, float gc, float go, const float4 b, const float4 pe,
                                             float c_old, float* c_new) {
    const float i = chainer_sigmoid(__fadd_rn(__fadd_rn(gi, b.x), __fmul_rn(c_old, pe.x)));
    const float f = chainer_sigmoid(__fadd_rn(__fadd_rn(gf, b.y), __fmul_rn(c_old, pe.y)));
    const float cn = __fadd_rn(__fmul_rn(tanhf(__fadd_rn(gc, b.z)), i), __fmul_rn(f, c_old));
    const float o = chainer_sigmoid(__fadd_rn(__fadd_rn(go, b.w), __fmul_rn(c_old, pe.z)));
    *c_new = cn;
    return __fmul_rn(o, tanhf(cn));
}

__device__ __forceinline__ void view_store4(const View& v, long long pix, int c, const float* val) {
    const long long idx = pix * v.pitch + v.coff + c;
    if ((v.pitch | v.coff) & 3) {  // layer-0 concat buffer (pitch 2*C0 + R1 + C0): not 16-byte aligned
#pragma unroll
        for (int i = 0; i < 4; ++i) view_store(v, pix, c + i, val[i]);
        return;
    }
    if (v.lo) {
        float4 h, l;
        h.x = tf32_round(val[0]); h.y = tf32_round(val[1]); h.z = tf32_round(val[2]); h.w = tf32_round(val[3]);
        l.x = __fsub_rn(val[0], h.x); l.y = __fsub_rn(val[1], h.y); l.z = __fsub_rn(val[2], h.z); l.w = __fsub_rn(val[3], h.w);
        *reinterpret_cast<float4*>(v.hi + idx) = h;
        *reinterpret_cast<float4*>(v.lo + idx) = l;
    } else {
        *reinterpret_cast<float4*>(v.hi + idx) = make_float4(val[0], val[1], val[2], val[3]);
    }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int KBT> struct TcSwz;
template <> struct TcSwz<32> { static constexpr uint64_t layout = 2, sbo = 1024; };  // SWIZZLE_128B
template <> struct TcSwz<16> { static constexpr uint64_t layout = 4, sbo = 512; };   // SWIZZLE_64B

// K-major swizzled shared-memory matrix descriptor (rows of KBT*4 bytes, 8-row atoms `sbo` bytes apart).  Measured on
// B200 (tests/gpu/tc_check): the swizzle XOR is a function of the absolute smem address, so a start address shifted by
// whole rows is legal with base offset 0 - that is what makes the nine taps nine views of one halo box.
template <int KBT>
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(TcSwz<KBT>::sbo >> 4) << 32;  // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= TcSwz<KBT>::layout << 61;
    return d;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Work decomposition: a *group* is `csize` spatially consecutive CTA regions that use the same weight slice n0; the
// CTAs of one cluster walk the groups in lockstep (cluster k takes groups k, k + #clusters, ...), CTA rank r of the
// cluster computes region r of the group, and every weight tile is fetched once per cluster (each CTA loads 1/csize
// of it and multicasts).  A rank without a region (tail group) still loads its slice and releases the stages.
struct TcRegion { int x0, y0, b, n0; bool active; };
__device__ __forceinline__ TcRegion tc_region(const TcParams& p, int group, int rank) {
    TcRegion r;
    const int nz = group / p.groups_per_nz;
    const int sp = (group - nz * p.groups_per_nz) * p.csize + rank;
    r.active = sp < p.regions;
    const int tiles = p.tiles_x * p.tiles_y;
    const int tile = sp % tiles;
    r.b = r.active ? sp / tiles : 0;
    r.n0 = nz * p.Ncta;
    r.x0 = (tile % p.tiles_x) * p.TW;
    r.y0 = (tile / p.tiles_x) * (p.NT * p.TH);
    return r;
}

// Persistent, warp-specialised: grid = min(regions, #SM) CTAs of 512 threads, each looping over CTA regions
// (NT stacked MMA tiles of one genome x Ncta output channels).
//   warps 0-3   converter: raw fp32 halo box (TMA) -> hi = tf32(v) in place, lo = v - hi in the second plane
//   warp  4     weight producer (one thread): per-tap weight tiles, hi and lo planes (pre-split on the host)
//   warp  5     MMA issuer (one thread) + TMEM allocator; accumulators double-buffered in TMEM
//   warp  6     activation producer (one thread): one halo box per channel block, runs SA stages ahead
//   warp  7     idle
//   warps 8-15  epilogue (TMEM lane quarter = warp & 3, column half = (warp - 8) >> 2): drains accumulator set i
//               while set i^1 is being computed
template <int KBT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, const TcParams p) {
    constexpr int RB = KBT * 4;       // bytes per operand row
    constexpr int KSTEPS = KBT / 8;   // UMMA_K = 8 for tf32
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sbase = raw_addr + pad;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int crank = p.csize > 1 ? (int)cluster_rank() : 0;
    const int cluster_id = blockIdx.x / p.csize, n_clusters = gridDim.x / p.csize;
    const uint16_t cmask = (uint16_t)((1u << p.csize) - 1u);
    const int a_stage_bytes = 2 * p.a_plane_bytes, b_stage_bytes = 2 * p.b_plane_bytes;
    const uint32_t sA = sbase, sB = sbase + p.SA * a_stage_bytes;
    const uint32_t pipe_bytes = p.SA * a_stage_bytes + p.SB * b_stage_bytes;
    float* stage = reinterpret_cast<float*>(smem + pipe_bytes);
    const uint32_t sBar = sbase + pipe_bytes + p.staging_bytes;
    // barriers: fullA[SA] convA[SA] emptyA[SA] fullB[SB] emptyB[SB] accFull[2] accEmpty[2]
    const uint32_t fullA = sBar, convA = fullA + 8 * p.SA, emptyA = convA + 8 * p.SA;
    const uint32_t fullB = emptyA + 8 * p.SA, emptyB = fullB + 8 * p.SB;
    const uint32_t accFull = emptyB + 8 * p.SB, accEmpty = accFull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pipe_bytes + p.staging_bytes + 8 * (3 * p.SA + 2 * p.SB + 4));

    if (warp == 4 && lane == 0) {
        for (int i = 0; i < p.SA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(convA + 8 * i, 4); mbar_init(emptyA + 8 * i, 1); }
        for (int i = 0; i < p.SB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, p.csize); }
        for (int i = 0; i < 2; ++i) { mbar_init(accFull + 8 * i, 1); mbar_init(accEmpty + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (p.csize > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int acc_stride = p.NT * p.Ncta;

    if (warp < 4) {
        // ===== converter =====
        const int chunks = p.a_box_bytes >> 4;
        int ia = 0;
        long long c_wait = 0, c0 = clock64(), cq;
        for (int grp = cluster_id; grp < p.groups; grp += n_clusters) {
            if (!tc_region(p, grp, crank).active) continue;
            for (int kb = 0; kb < p.KBn; ++kb, ++ia) {
                const int s = ia % p.SA;
                cq = clock64();
                mbar_wait(fullA + 8 * s, (ia / p.SA) & 1);
                c_wait += clock64() - cq;
                float4* p0 = reinterpret_cast<float4*>(smem + s * a_stage_bytes);
                float4* p1 = reinterpret_cast<float4*>(smem + s * a_stage_bytes + p.a_plane_bytes);
#pragma unroll 2
                for (int i = threadIdx.x; i < chunks; i += 128) {
                    const float4 v = p0[i];
                    float4 h, l;
                    h.x = tf32_round(v.x); h.y = tf32_round(v.y); h.z = tf32_round(v.z); h.w = tf32_round(v.w);
                    l.x = __fsub_rn(v.x, h.x); l.y = __fsub_rn(v.y, h.y); l.z = __fsub_rn(v.z, h.z); l.w = __fsub_rn(v.w, h.w);
                    p0[i] = h;
                    p1[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(convA + 8 * s);
            }
        }
        if (p.dbg && threadIdx.x == 0) {
            long long* d = p.dbg + (long long)blockIdx.x * 16;
            d[7] = clock64() - c0; d[8] = c_wait;
        }
    } else if (warp == 6) {
        // ===== activation (A) producer =====
        if (lane == 0) {
            int ia = 0;
            for (int grp = cluster_id; grp < p.groups; grp += n_clusters) {
                const TcRegion r = tc_region(p, grp, crank);
                if (!r.active) continue;
                for (int kb = 0; kb < p.KBn; ++kb, ++ia) {
                    const int s = ia % p.SA;
                    mbar_wait(emptyA + 8 * s, ((ia / p.SA) & 1) ^ 1);
                    mbar_expect_tx(fullA + 8 * s, p.a_box_bytes);
                    tma_load_4d(sA + s * a_stage_bytes, &mA, fullA + 8 * s, kb * KBT, r.x0 - 1, r.y0 - 1, r.b);
                }
            }
        }
    } else if (warp == 4) {
        // ===== weight (B) producer: this CTA's 1/csize slice of every tile, multicast to the whole cluster =====
        if (lane == 0) {
            int ib = 0;
            long long b_wait = 0, b0 = clock64(), bq;
            const int slice_rows = p.Ncta / p.csize;
            const uint32_t slice_off = (uint32_t)(crank * slice_rows * RB);
            for (int grp = cluster_id; grp < p.groups; grp += n_clusters) {
                const int n0 = (grp / p.groups_per_nz) * p.Ncta + crank * slice_rows;
                for (int kb = 0; kb < p.KBn; ++kb) {
                    for (int tap = 0; tap < 9; ++tap, ++ib) {
                        const int s = ib % p.SB;
                        bq = clock64();
                        mbar_wait(emptyB + 8 * s, ((ib / p.SB) & 1) ^ 1);
                        b_wait += clock64() - bq;
                        mbar_expect_tx(fullB + 8 * s, 2 * p.b_plane_bytes);
                        const uint32_t dst = sB + s * b_stage_bytes + slice_off;
                        const int row_hi = (tap * p.KBn + kb) * p.N + n0, row_lo = ((9 + tap) * p.KBn + kb) * p.N + n0;
                        if (p.csize > 1) {
                            tma_load_2d_mc(dst, &mB, fullB + 8 * s, 0, row_hi, cmask);
                            tma_load_2d_mc(dst + p.b_plane_bytes, &mB, fullB + 8 * s, 0, row_lo, cmask);
                        } else {
                            tma_load_2d(dst, &mB, fullB + 8 * s, 0, row_hi);
                            tma_load_2d(dst + p.b_plane_bytes, &mB, fullB + 8 * s, 0, row_lo);
                        }
                    }
                }
            }
            if (p.dbg) {
                long long* d = p.dbg + (long long)blockIdx.x * 16;
                d[9] = clock64() - b0; d[10] = b_wait;
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer: the warp walks the loops together, one elected lane issues =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Ncta >> 3) << 17) | ((128u >> 4) << 24);
        // descriptor = {hi word: constant, lo word: (address >> 4) | LBO}; offsets are plain adds on the lo word
        const uint64_t desc_hi = tc_smem_desc<KBT>(0) & 0xffffffff00000000ull;
        const uint32_t lo_flag = 1u << 16;
        const uint32_t plane_a16 = (uint32_t)p.a_plane_bytes >> 4, plane_b16 = (uint32_t)p.b_plane_bytes >> 4;
        const uint32_t tile16 = (uint32_t)(p.TH * p.P * RB) >> 4;
        int ia = 0, ib = 0, it = 0;
        long long t_acc = 0, t_a = 0, t_b = 0, t0 = clock64(), tq;
        for (int grp = cluster_id; grp < p.groups; grp += n_clusters) {
            if (!tc_region(p, grp, crank).active) {
                // no region for this rank in the tail group: keep the weight ring moving for the cluster
                for (int k = 0; k < 9 * p.KBn; ++k, ++ib) {
                    const int sb = ib % p.SB;
                    mbar_wait(fullB + 8 * sb, (ib / p.SB) & 1);
                    if (elect_one()) tc_commit_mc(emptyB + 8 * sb, cmask);
                    __syncwarp();
                }
                continue;
            }
            const int set = it & 1;
            tq = clock64();
            mbar_wait(accEmpty + 8 * set, ((it >> 1) & 1) ^ 1);
            t_acc += clock64() - tq;
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(set * acc_stride);
            for (int kb = 0; kb < p.KBn; ++kb, ++ia) {
                const int sa = ia % p.SA;
                tq = clock64();
                mbar_wait(convA + 8 * sa, (ia / p.SA) & 1);
                t_a += clock64() - tq;
                const uint32_t a16 = (((sA + sa * a_stage_bytes) & 0x3FFFFu) >> 4) | lo_flag;
                for (int tap = 0; tap < 9; ++tap, ++ib) {
                    const int sb = ib % p.SB;
                    tq = clock64();
                    mbar_wait(fullB + 8 * sb, (ib / p.SB) & 1);
                    t_b += clock64() - tq;
                    tc_fence_after();
                    const uint32_t tap16 = (uint32_t)(((tap / 3) * p.P + (tap % 3)) * RB) >> 4;
                    const uint32_t b16 = (((sB + sb * b_stage_bytes) & 0x3FFFFu) >> 4) | lo_flag;
                    const uint32_t acc0 = (kb | tap) ? 1u : 0u;
                    if (elect_one()) {
                        for (int t = 0; t < p.NT; ++t) {
                            const uint32_t at16 = a16 + tap16 + (uint32_t)t * tile16;
                            const uint32_t d = d0 + (uint32_t)(t * p.Ncta);
#pragma unroll
                            for (int ks = 0; ks < KSTEPS; ++ks) {
                                const uint64_t dah = desc_hi | (at16 + 2 * ks), dal = desc_hi | (at16 + plane_a16 + 2 * ks);
                                const uint64_t dbh = desc_hi | (b16 + 2 * ks), dbl = desc_hi | (b16 + plane_b16 + 2 * ks);
                                tc_mma_tf32(d, dal, dbh, idesc, ks ? 1u : acc0);
                                tc_mma_tf32(d, dah, dbl, idesc, 1u);
                                tc_mma_tf32(d, dah, dbh, idesc, 1u);
                            }
                        }
                        if (p.csize > 1) tc_commit_mc(emptyB + 8 * sb, cmask);
                        else tc_commit(emptyB + 8 * sb);
                        if (tap == 8) tc_commit(emptyA + 8 * sa);
                        if (tap == 8 && kb == p.KBn - 1) tc_commit(accFull + 8 * set);
                    }
                    __syncwarp();
                }
            }
            ++it;
        }
        if (p.dbg && lane == 0) {
            long long* d = p.dbg + (long long)blockIdx.x * 16;
            d[0] = clock64() - t0; d[1] = t_acc; d[2] = t_a; d[3] = t_b; d[4] = it;
        }
    } else if (warp >= 8) {
        // ===== epilogue warps 8..15 =====
        const ConvArgs& a = p.ca;
        const int q4 = warp & 3, half = (warp - 8) >> 2;
        const int m = q4 * 32 + lane;
        const int etid = (warp - 8) * 32 + lane;
        const int hh = m / p.P, ww = m - hh * p.P;
        int it = 0;
        long long e_wait = 0, e0 = clock64(), eq;
        for (int grp = cluster_id; grp < p.groups; grp += n_clusters) {
            const TcRegion r = tc_region(p, grp, crank);
            if (!r.active) continue;
            const int set = it & 1, b = r.b, n0 = r.n0;
            eq = clock64();
            mbar_wait(accFull + 8 * set, (it >> 1) & 1);
            e_wait += clock64() - eq;
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(set * acc_stride);
            for (int t = 0; t < p.NT; ++t) {
                const int y = r.y0 + t * p.TH + hh, x = r.x0 + ww;
                const bool valid = hh < p.TH && ww < p.TW && y < p.H && x < p.W;
                const long long pix = valid ? ((long long)b * p.H + y) * p.W + x : 0;
                const uint32_t tcol = lane_addr + (uint32_t)(t * p.Ncta);
                if (a.epi == EPI_LSTM) {
                    const int R = a.N >> 2;
                    const long long ppix = valid ? (long long)y * p.W + x : 0;
                    // software pipeline: the state / peephole loads of chunk c+32 are in flight while chunk c is computed
                    float4 cold, pq[4];
                    int c0 = half * 16;
                    if (c0 < p.Ncta) {
                        const int r0 = (n0 + c0) >> 2;
                        cold = *reinterpret_cast<const float4*>(a.cstate + pix * R + r0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) pq[q] = *reinterpret_cast<const float4*>(a.peep + (ppix * R + r0 + q) * 4);
                    }
                    for (; c0 < p.Ncta; c0 += 32) {
                        float v[16];
                        tmem_ld16(tcol + c0, v);
                        const int r0 = (n0 + c0) >> 2;
                        const float4 ccur = cold;
                        float4 pcur[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) pcur[q] = pq[q];
                        if (c0 + 32 < p.Ncta) {
                            const int r1 = (n0 + c0 + 32) >> 2;
                            cold = *reinterpret_cast<const float4*>(a.cstate + pix * R + r1);
#pragma unroll
                            for (int q = 0; q < 4; ++q) pq[q] = *reinterpret_cast<const float4*>(a.peep + (ppix * R + r1 + q) * 4);
                        }
                        if (!valid) continue;
                        const float co[4] = {ccur.x, ccur.y, ccur.z, ccur.w};
                        float cn[4], hn[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bq = *reinterpret_cast<const float4*>(a.bias + n0 + c0 + q * 4);
                            hn[q] = lstm_cell_v(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3], bq, pcur[q], co[q], &cn[q]);
                        }
                        *reinterpret_cast<float4*>(a.cstate + pix * R + r0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                        view_store4(a.dstH, pix, r0, hn);
                        if (a.dstUp.hi) {
                            const int W2 = p.W * 2;
                            const long long ub = ((long long)b * p.H * 2 + y * 2) * W2 + x * 2;
                            view_store4(a.dstUp, ub, r0, hn);
                            view_store4(a.dstUp, ub + 1, r0, hn);
                            view_store4(a.dstUp, ub + W2, r0, hn);
                            view_store4(a.dstUp, ub + W2 + 1, r0, hn);
                        }
                    }
                } else if (a.epi == EPI_CONVP || a.epi == EPI_RAW) {
                    for (int c0 = half * 16; c0 < p.Ncta; c0 += 32) {
                        float v[16];
                        tmem_ld16(tcol + c0, v);
                        if (!valid) continue;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bq = *reinterpret_cast<const float4*>(a.bias + n0 + c0 + q * 4);
                            float o[4] = {__fadd_rn(v[q * 4], bq.x), __fadd_rn(v[q * 4 + 1], bq.y), __fadd_rn(v[q * 4 + 2], bq.z),
                                          __fadd_rn(v[q * 4 + 3], bq.w)};
                            if (a.epi == EPI_CONVP) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    o[i] = o[i] > 0.f ? o[i] : 0.f;
                                    if (a.clip && o[i] > 1.f) o[i] = 1.f;
                                }
                            }
                            *reinterpret_cast<float4*>(a.outP + pix * a.N + n0 + c0 + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    }
                } else {  // EPI_CONVA: relu -> staging tile -> 2x2 max-pool -> error units at half resolution
                    for (int c0 = half * 16; c0 < p.Ncta; c0 += 32) {
                        float v[16];
                        tmem_ld16(tcol + c0, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float o = __fadd_rn(v[i], a.bias[n0 + c0 + i]);
                            stage[m * p.stage_ld + c0 + i] = o > 0.f ? o : 0.f;
                        }
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    const int Hp = p.H >> 1, Wp = p.W >> 1, tw2 = p.TW >> 1, th2 = p.TH >> 1;
                    const int items = th2 * tw2 * p.Ncta;
                    for (int base = 0; base < items; base += 4 * 256) {
                        float mx[4], pv[4];
                        long long ppos[4];
                        int nn[4];
                        bool okk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {   // loads first: four independent global reads in flight
                            const int idx = base + u * 256 + etid;
                            okk[u] = idx < items;
                            const int n = idx % p.Ncta, pp = idx / p.Ncta;
                            const int ph = pp / tw2, pw = pp - ph * tw2;
                            const int py = ((r.y0 + t * p.TH) >> 1) + ph, px = (r.x0 >> 1) + pw;
                            okk[u] = okk[u] && py < Hp && px < Wp;
                            nn[u] = n0 + n;
                            ppos[u] = okk[u] ? ((long long)b * Hp + py) * Wp + px : 0;
                            pv[u] = okk[u] ? a.P[ppos[u] * a.N + nn[u]] : 0.f;
                            const int m00 = okk[u] ? (2 * ph) * p.P + 2 * pw : 0;
                            const float* s0 = stage + m00 * p.stage_ld + n;
                            mx[u] = fmaxf(fmaxf(s0[0], s0[p.stage_ld]), fmaxf(s0[p.P * p.stage_ld], s0[(p.P + 1) * p.stage_ld]));
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!okk[u]) continue;
                            const float ep = __fsub_rn(mx[u], pv[u]), en = __fsub_rn(pv[u], mx[u]);
                            view_store(a.dstE, ppos[u], nn[u], ep > 0.f ? ep : 0.f);
                            view_store(a.dstE, ppos[u], a.N + nn[u], en > 0.f ? en : 0.f);
                        }
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty + 8 * set);
            ++it;
        }
        if (p.dbg && warp == 8 && lane == 0) {
            long long* d = p.dbg + (long long)blockIdx.x * 16;
            d[5] = clock64() - e0; d[6] = e_wait;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.csize > 1) cluster_sync_all();   // nobody exits while a peer may still multicast into / arrive on this CTA
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EigEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
    EigEncodeTiledFn encode = nullptr;
    bool probed = false, available = false;
    std::string reason, last_error;
    int kbt = 16;      // channels per K block (EIG_TC_KB = 16 | 32)
    int force_nt = 0;  // EIG_TC_NT: cap on MMA tiles per CTA region
    long long* dbg = nullptr;  // device buffer for the per-role cycle counters (tests only)
    int last_grid = 0, last_csize = 0, last_nt = 0, last_sa = 0, last_sb = 0;
    int max_cluster = 4;  // EIG_TC_CLUSTER: cap on the multicast cluster size (1, 2 or 4)
    std::map<std::tuple<int, int, size_t>, int> max_clusters;  // (KBT, csize, smem) -> co-resident clusters
    int n_sm = 148;
    std::map<std::tuple<const void*, int, int, int, int, int, int, int, int>, CUtensorMap> amaps;
};
inline TcState& tc_state() { static TcState s; return s; }

inline bool tc_available() {
    TcState& s = tc_state();
    if (s.probed) return s.available;
    s.probed = true;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { s.reason = "no CUDA device"; return false; }
    if (prop.major != 10) { s.reason = "tcgen05 needs an sm_100-class GPU (found sm_" + std::to_string(prop.major * 10 + prop.minor) + ")"; return false; }
    s.n_sm = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        s.reason = "cuTensorMapEncodeTiled not exported by the driver";
        return false;
    }
    s.encode = (EigEncodeTiledFn)fn;
    if (cudaFuncSetAttribute(conv3x3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess ||
        cudaFuncSetAttribute(conv3x3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess) {
        s.reason = "cannot raise the dynamic shared memory limit";
        cudaGetLastError();
        return false;
    }
    if (const char* e = getenv("EIG_TC_KB")) s.kbt = atoi(e) == 32 ? 32 : 16;
    if (const char* e = getenv("EIG_TC_NT")) s.force_nt = atoi(e);
    if (const char* e = getenv("EIG_TC_CLUSTER")) s.max_cluster = atoi(e) >= 4 ? 4 : (atoi(e) >= 2 ? 2 : 1);
    s.available = true;
    return true;
}
inline std::string tc_unavailable_reason() { return tc_state().reason; }
inline std::string tc_last_error() { return tc_state().last_error; }
inline void tc_set_kb(int kbt) { tc_state().kbt = kbt == 32 ? 32 : 16; }
inline void tc_set_max_nt(int nt) { tc_state().force_nt = nt; }
inline void tc_set_max_cluster(int c) { tc_state().max_cluster = c >= 4 ? 4 : (c >= 2 ? 2 : 1); }

inline float tc_host_tf32(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;
    u = (u + 0x1000u) & ~0x1fffu;  // round to nearest, ties away (cvt.rna.tf32.f32)
    float r;
    memcpy(&r, &u, 4);
    return r;
}

inline void tc_free(TcWeights& w) {
    if (w.d) cudaFree(w.d);
    w.d = nullptr;
    w.ok = false;
}

// wv: [9][cin][npad] fp32 (the SIMT layout), N valid columns; max_ncta caps the output channels of one CTA
inline int tc_pack(TcWeights& w, const float* wv, int cin, int N, int npad, int max_ncta = 256) {
    if (!tc_available()) return 0;  // no tensor-core path on this device: nothing to pack
    TcState& s = tc_state();
    tc_free(w);
    if (N % 16 || cin % 4) return 0;  // not a tensor-core shape: w.ok stays false, the caller keeps the SIMT kernel
    const int KBT = s.kbt;
    w.cin = cin; w.N = N; w.KBT = KBT; w.KBn = (cin + KBT - 1) / KBT;
    w.gz = (N + max_ncta - 1) / max_ncta;
    while (N % w.gz || (N / w.gz) % 16) ++w.gz;
    w.Ncta = N / w.gz;
    const size_t plane = (size_t)9 * w.KBn * N * KBT;
    std::vector<float> pk(2 * plane, 0.f);
    for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < cin; ++c)
            for (int n = 0; n < N; ++n) {
                const float v = wv[((size_t)tap * cin + c) * npad + n];
                const float hi = tc_host_tf32(v);
                const size_t o = (((size_t)tap * w.KBn + c / KBT) * N + n) * KBT + c % KBT;
                pk[o] = hi;
                pk[plane + o] = v - hi;
            }
    if (cudaMalloc((void**)&w.d, pk.size() * sizeof(float)) != cudaSuccess) { s.last_error = "tc_pack: cudaMalloc failed"; return -1; }
    if (cudaMemcpy(w.d, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { s.last_error = "tc_pack: upload failed"; return -1; }
    const cuuint64_t gdim[2] = {(cuuint64_t)KBT, (cuuint64_t)2 * 9 * w.KBn * N};
    const cuuint64_t gstr[1] = {(cuuint64_t)KBT * sizeof(float)};
    const cuuint32_t est[2] = {1, 1};
    w.max_csize = 1;
    for (int ci = 0; ci < 3; ++ci) {
        const int cs = 1 << ci;
        if (w.Ncta % (8 * cs)) break;   // every CTA's slice must be whole 8-row swizzle atoms
        const cuuint32_t box[2] = {(cuuint32_t)KBT, (cuuint32_t)(w.Ncta / cs)};
        const CUresult r = s.encode(&w.map[ci], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w.d, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    KBT == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { s.last_error = "tc_pack: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -1; }
        w.max_csize = cs;
    }
    w.ok = true;
    return 0;
}

struct TcGeom { int TW, TH, P, NT, SA, SB, a_plane, b_plane, tmem_cols, tiles_x, tiles_y, regions, staging, stage_ld; size_t smem; };

inline int tc_round_up(int v, int m) { return (v + m - 1) / m * m; }

// Picks the flat-padded tile (TW x TH, P = TW + 2, (TH-1)*P + TW <= 128) with the best MMA-row efficiency, then the
// number of stacked tiles per CTA region (weight-tile reuse) that still load-balances over the SMs and fits shared memory.
inline bool tc_geometry(int B, int H, int W, int KBT, int Ncta, int gz, bool pooled, int force_nt, int n_sm, int csize, TcGeom& g) {
    double best = -1.0;
    int bTW = 0, bTH = 0;
    const int step = pooled ? 2 : 1;
    for (int TW = step; TW <= W && TW <= 126; TW += step) {
        const int P = TW + 2;
        int TH = (128 - TW) / P + 1;
        if (TH > H) TH = H;
        if (pooled) TH &= ~1;
        if (TH < 1) continue;
        const long long tiles = (long long)((W + TW - 1) / TW) * ((H + TH - 1) / TH);
        const double eff = (double)W * H / (tiles * 128.0) - 0.02 * (2.0 / TH) - 0.02 * (2.0 / TW);
        if (eff > best) { best = eff; bTW = TW; bTH = TH; }
    }
    if (best < 0) return false;
    g.TW = bTW; g.TH = bTH; g.P = bTW + 2;
    g.tiles_x = (W + g.TW - 1) / g.TW;
    const int row_tiles = (H + g.TH - 1) / g.TH;
    const int RB = KBT * 4;
    g.b_plane = Ncta * RB;
    g.stage_ld = Ncta + 1;
    g.staging = pooled ? tc_round_up(128 * g.stage_ld * 4, 1024) : 0;
    int nt_cap = 256 / Ncta;  // two accumulator sets of NT * Ncta columns in the 512 TMEM columns
    if (nt_cap > row_tiles) nt_cap = row_tiles;
    if (nt_cap < 1) nt_cap = 1;
    if (force_nt > 0 && nt_cap > force_nt) nt_cap = force_nt;
    int pick = 0;
    double pick_eff = -1.0;
    TcGeom cand[9];
    for (int NT = nt_cap; NT >= 1; --NT) {
        const int rows = std::max((NT * g.TH + 2) * g.P, (NT - 1) * g.TH * g.P + 2 * g.P + 2 + 128);
        const int a_plane = tc_round_up(rows * RB, 1024);
        int SA = 3, SB = 0;
        for (; SA >= 2; --SA) {   // three activation stages when the weight ring still gets >= 4
            const long long left = (long long)TC_SMEM_LIMIT - 4096 - g.staging - (long long)SA * 2 * a_plane;
            SB = left > 0 ? (int)(left / (2 * g.b_plane)) : 0;
            if (SB > 8) SB = 8;
            if (SB >= (SA == 3 ? 4 : 2)) break;
        }
        if (SA < 2) continue;
        TcGeom c = g;
        c.NT = NT; c.SA = SA; c.SB = SB; c.a_plane = a_plane;
        c.tiles_y = (row_tiles + NT - 1) / NT;
        c.regions = c.tiles_x * c.tiles_y * B;
        int cols = 32;
        while (cols < 2 * NT * Ncta) cols <<= 1;
        c.tmem_cols = cols;
        c.smem = (size_t)SA * 2 * a_plane + (size_t)SB * 2 * g.b_plane + g.staging + 8 * (3 * SA + 2 * SB + 4) + 16 + 1024;
        const int slots = n_sm / csize;                                    // clusters that run side by side
        const int groups = ((c.regions + csize - 1) / csize) * gz;
        const int rounds = (groups + slots - 1) / slots;
        const double eff = (double)c.regions * gz / ((double)rounds * slots * csize);
        cand[NT] = c;
        if (eff >= 0.85) { pick = NT; break; }       // largest NT that still fills the machine evenly
        if (eff > pick_eff) { pick_eff = eff; pick = NT; }
    }
    if (!pick) return false;
    g = cand[pick];
    return true;
}

inline int tc_conv(const TcWeights& w, const ConvArgs& a, cudaStream_t stream) {
    TcState& s = tc_state();
    if (!tc_available()) { s.last_error = s.reason; return -1; }
    if (!w.ok) { s.last_error = "tc_conv: weights not packed"; return -1; }
    if (a.in_lo) { s.last_error = "tc_conv: split activation planes are not used any more (pass in_lo = nullptr)"; return -1; }
    if (a.Cin != w.cin || a.N != w.N) { s.last_error = "tc_conv: shape mismatch with packed weights"; return -1; }
    if ((a.in_coff & 3) || (a.in_pitch & 3)) { s.last_error = "tc_conv: view not 16-byte aligned"; return -1; }
    const bool pooled = a.epi == EPI_CONVA;
    if (pooled && ((a.H | a.W) & 1)) { s.last_error = "tc_conv: pooled conv needs even H, W"; return -1; }
    if (pooled && w.Ncta > 128) { s.last_error = "tc_conv: pooled conv needs <= 128 channels per CTA"; return -1; }
    int csize = std::min(w.max_csize, s.max_cluster);
    TcGeom g;
    if (!tc_geometry(a.B, a.H, a.W, w.KBT, w.Ncta, w.gz, pooled, s.force_nt, s.n_sm, csize, g)) { s.last_error = "tc_conv: no tile geometry fits"; return -1; }
    while (csize > 1 && g.regions < csize) csize >>= 1;
    const int box_rows = g.NT * g.TH + 2;
    auto key = std::make_tuple((const void*)(a.in_hi + a.in_coff), a.Cin, a.W, a.H, a.B, a.in_pitch, g.P, box_rows, w.KBT);
    auto it = s.amaps.find(key);
    if (it == s.amaps.end()) {
        CUtensorMap map;
        const cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        const cuuint64_t gstr[3] = {(cuuint64_t)a.in_pitch * 4, (cuuint64_t)a.W * a.in_pitch * 4, (cuuint64_t)a.H * a.W * a.in_pitch * 4};
        const cuuint32_t box[4] = {(cuuint32_t)w.KBT, (cuuint32_t)g.P, (cuuint32_t)box_rows, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        const CUresult r = s.encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(a.in_hi + a.in_coff), gdim, gstr, box, est,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, w.KBT == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { s.last_error = "tc_conv: cuTensorMapEncodeTiled(A) failed (" + std::to_string((int)r) + ")"; return -1; }
        it = s.amaps.emplace(key, map).first;
    }
    TcParams p;
    memset(&p, 0, sizeof p);
    p.B = a.B; p.H = a.H; p.W = a.W;
    p.TW = g.TW; p.TH = g.TH; p.P = g.P; p.NT = g.NT; p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y; p.regions = g.regions;
    p.KBn = w.KBn; p.Ncta = w.Ncta; p.N = w.N;
    p.a_plane_bytes = g.a_plane; p.b_plane_bytes = g.b_plane; p.a_box_bytes = w.KBT * 4 * g.P * box_rows;
    p.SA = g.SA; p.SB = g.SB; p.tmem_cols = g.tmem_cols;
    p.staging_bytes = g.staging; p.stage_ld = g.stage_ld;
    p.ca = a;
    if (g.smem > TC_SMEM_LIMIT) { s.last_error = "tc_conv: shared memory budget exceeded"; return -1; }
    p.csize = csize;
    p.dbg = s.dbg;
    p.groups_per_nz = (g.regions + csize - 1) / csize;
    p.groups = p.groups_per_nz * w.gz;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = g.smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    void (*kern)(const CUtensorMap, const CUtensorMap, const TcParams) = w.KBT == 32 ? conv3x3_tc_kernel<32> : conv3x3_tc_kernel<16>;
    int slots = s.n_sm / csize;
    if (csize > 1) {   // how many clusters can be resident at once (GPC boundaries cost a few SMs)
        auto ck = std::make_tuple(w.KBT, csize, g.smem);
        auto ci = s.max_clusters.find(ck);
        if (ci == s.max_clusters.end()) {
            int n = 0;
            cfg.gridDim = dim3(s.n_sm / csize * csize);
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = s.n_sm / csize / 2; }
            ci = s.max_clusters.emplace(ck, n).first;
        }
        slots = std::min(slots, ci->second);
    }
    const int n_clusters = std::min(p.groups, slots);
    cfg.gridDim = dim3(n_clusters * csize);
    s.last_grid = n_clusters * csize; s.last_csize = csize; s.last_nt = g.NT; s.last_sa = g.SA; s.last_sb = g.SB;
    const int ci_map = csize == 4 ? 2 : (csize == 2 ? 1 : 0);
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, it->second, w.map[ci_map], p);
    if (le != cudaSuccess) { s.last_error = std::string("tc_conv launch: ") + cudaGetErrorString(le); cudaGetLastError(); return -1; }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s.last_error = std::string("tc_conv launch: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

}  // namespace eig
