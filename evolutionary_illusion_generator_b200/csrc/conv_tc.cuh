// tcgen05 tensor-core 3x3 convolution for PredNet layers 1..3 (sm_100a): TMA-staged implicit GEMM, 3xTF32
// split-precision MMAs accumulating in TMEM, the three PredNet epilogues fused behind tcgen05.ld.
//
// Reference semantics are those of conv_simt.cuh (Chainer `L.Convolution2D(cin, cout, 3, pad=1)` +
// ConvLSTM / error-unit / pooling epilogues, /root/reference/chainer_prednet/PredNet/net.py:46-62,94-126,
// 187-209); the SIMT kernel is the exact-fp32 twin this one is checked against on the GPU.
//
// GEMM view: D[m][n] = sum_{tap, c} A[m + shift(tap)][c] * Wt[tap][c][n]
//   m   = "flat padded" pixel index inside a CTA region: m = h * P + w, P = TW + 2.  The CTA loads ONE halo box
//         (32 channels x P columns x NT*TH+2 rows) per 32-channel block with a single 4-D TMA (negative / out of
//         range coordinates are zero-filled by the TMA unit = the conv's zero padding) and the nine taps are nine
//         row-shifted views of that box: tap (ky,kx) of MMA tile t starts (t*TH + ky) * P + kx rows into the box.
//         Rows with w >= TW are computed and dropped (2/P waste); the halo is fetched once instead of 9 times.
//   K   = 32 channels per block (one 128-byte swizzle row), UMMA_K = 8 -> 4 k-steps per (block, tap)
//   N   = output channels of this CTA (<= 256; the 4 gates of an LSTM cell are adjacent columns)
// 3xTF32: activations and weights are stored as hi = tf32(v), lo = v - hi; every k-step issues
//   lo*hi + hi*lo + hi*hi into the same fp32 TMEM accumulator (the lo*lo term, 2^-22 relative, is dropped).
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer, warp 5 MMA
// issuer + TMEM allocator.  Two mbarrier rings: A (halo boxes, SA stages) and B (per-tap weight tiles, SB stages).
#pragma once
#include <cuda.h>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "common.cuh"
#include "conv_simt.cuh"

namespace eig {

enum { EPI_RAW = 3 };  // test only: out = acc + bias, no activation (conv3x3_tc_kernel only)
enum { TC_KB = 32, TC_THREADS = 192, TC_SMEM_LIMIT = 227 * 1024 };

struct TcWeights {
    float* d = nullptr;  // [2 planes][9 taps][KBn][N][32] fp32 (hi plane, then lo plane)
    int cin = 0, N = 0, KBn = 0, Ncta = 0, gz = 1;
    CUtensorMap map;
    bool ok = false;
};

struct TcParams {
    int B, H, W;
    int TW, TH, P, NT;
    int tiles_x, tiles_y;
    int KBn, Ncta, N;
    int a_plane_bytes, b_plane_bytes, a_box_bytes;
    int SA, SB;
    int a_mode;  // 1 (product): halo box, row-shifted descriptors with base offset 0; 2: one box per tap (9x the loads,
                 // same numbers - kept as the cross-check); 0: base offset = row phase (WRONG on B200, kept for the probe)
    int tmem_cols;
    int stage_ld;  // floats per row of the ConvA staging tile
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
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
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
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
                  const __grid_constant__ CUtensorMap mB, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sbase = raw_addr + pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_stage_bytes = 2 * p.a_plane_bytes, b_stage_bytes = 2 * p.b_plane_bytes;
    const uint32_t sA = sbase, sB = sbase + p.SA * a_stage_bytes;
    const uint32_t pipe_bytes = p.SA * a_stage_bytes + p.SB * b_stage_bytes;
    const uint32_t sBar = sbase + pipe_bytes;  // fullA[SA], emptyA[SA], fullB[SB], emptyB[SB], accFull
    const uint32_t fullA = sBar, emptyA = sBar + 8 * p.SA, fullB = sBar + 16 * p.SA, emptyB = fullB + 8 * p.SB;
    const uint32_t accFull = emptyB + 8 * p.SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pipe_bytes + 8 * (2 * p.SA + 2 * p.SB + 1));

    const int tile = blockIdx.x;
    const int x0 = (tile % p.tiles_x) * p.TW, y0 = (tile / p.tiles_x) * (p.NT * p.TH);
    const int b = blockIdx.y;
    const int n0 = blockIdx.z * p.Ncta;

    if (warp == 4 && lane == 0) {
        for (int i = 0; i < p.SA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
        for (int i = 0; i < p.SB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
        mbar_init(accFull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            int ia = 0, ib = 0;
            for (int kb = 0; kb < p.KBn; ++kb) {
                for (int tap = 0; tap < 9; ++tap) {
                    if (p.a_mode == 2 || tap == 0) {
                        const int s = ia % p.SA;
                        mbar_wait(emptyA + 8 * s, ((ia / p.SA) & 1) ^ 1);
                        mbar_expect_tx(fullA + 8 * s, 2 * p.a_box_bytes);
                        const int sx = p.a_mode == 2 ? tap % 3 : 0, sy = p.a_mode == 2 ? tap / 3 : 0;
                        const uint32_t dst = sA + s * a_stage_bytes;
                        tma_load_4d(dst, &mA_hi, fullA + 8 * s, kb * TC_KB, x0 - 1 + sx, y0 - 1 + sy, b);
                        tma_load_4d(dst + p.a_plane_bytes, &mA_lo, fullA + 8 * s, kb * TC_KB, x0 - 1 + sx, y0 - 1 + sy, b);
                        ++ia;
                    }
                    const int s = ib % p.SB;
                    mbar_wait(emptyB + 8 * s, ((ib / p.SB) & 1) ^ 1);
                    mbar_expect_tx(fullB + 8 * s, 2 * p.b_plane_bytes);
                    const uint32_t dst = sB + s * b_stage_bytes;
                    const int row_hi = (tap * p.KBn + kb) * p.N + n0;
                    const int row_lo = ((9 + tap) * p.KBn + kb) * p.N + n0;
                    tma_load_2d(dst, &mB, fullB + 8 * s, 0, row_hi);
                    tma_load_2d(dst + p.b_plane_bytes, &mB, fullB + 8 * s, 0, row_lo);
                    ++ib;
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Ncta >> 3) << 17) | ((128u >> 4) << 24);
            int ia = 0, ib = 0, cur_sa = 0;
            for (int kb = 0; kb < p.KBn; ++kb) {
                for (int tap = 0; tap < 9; ++tap) {
                    if (p.a_mode == 2 || tap == 0) {
                        cur_sa = ia % p.SA;
                        mbar_wait(fullA + 8 * cur_sa, (ia / p.SA) & 1);
                        ++ia;
                    }
                    const int sb = ib % p.SB;
                    mbar_wait(fullB + 8 * sb, (ib / p.SB) & 1);
                    tc_fence_after();
                    const int tap_rows = p.a_mode == 2 ? 0 : (tap / 3) * p.P + (tap % 3);
                    const uint32_t b_hi = sB + sb * b_stage_bytes, b_lo = b_hi + p.b_plane_bytes;
                    for (int t = 0; t < p.NT; ++t) {
                        const uint32_t a_hi = sA + cur_sa * a_stage_bytes + (uint32_t)(t * p.TH * p.P + tap_rows) * 128u;
                        const uint32_t a_lo = a_hi + p.a_plane_bytes;
                        const uint32_t boff = p.a_mode == 0 ? ((a_hi >> 7) & 7u) : 0u;
                        const uint32_t d = tmem_base + (uint32_t)(t * p.Ncta);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t dah = tc_smem_desc(a_hi + ks * 32, boff), dal = tc_smem_desc(a_lo + ks * 32, boff);
                            const uint64_t dbh = tc_smem_desc(b_hi + ks * 32, 0), dbl = tc_smem_desc(b_lo + ks * 32, 0);
                            const uint32_t first = (kb == 0 && tap == 0 && ks == 0) ? 0u : 1u;
                            tc_mma_tf32(d, dal, dbh, idesc, first);
                            tc_mma_tf32(d, dah, dbl, idesc, 1u);
                            tc_mma_tf32(d, dah, dbh, idesc, 1u);
                        }
                    }
                    tc_commit(emptyB + 8 * sb);
                    ++ib;
                    if (p.a_mode == 2 || tap == 8) tc_commit(emptyA + 8 * cur_sa);
                }
            }
            tc_commit(accFull);
        }
    } else {
        // ===== epilogue warps 0..3: TMEM lane quarter = warp =====
        mbar_wait(accFull, 0);
        tc_fence_after();
        const ConvArgs& a = p.ca;
        const int m = warp * 32 + lane;
        const int hh = m / p.P, ww = m - hh * p.P;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* stage = reinterpret_cast<float*>(smem);
        for (int t = 0; t < p.NT; ++t) {
            const int y = y0 + t * p.TH + hh, x = x0 + ww;
            const bool valid = hh < p.TH && ww < p.TW && y < p.H && x < p.W;
            const long long pix = ((long long)b * p.H + y) * p.W + x;
            const uint32_t tcol = lane_addr + (uint32_t)(t * p.Ncta);
            if (a.epi == EPI_LSTM) {
                const int R = a.N >> 2;
                const long long ppix = (long long)y * p.W + x;
                for (int c0 = 0; c0 < p.Ncta; c0 += 16) {
                    float v[16];
                    tmem_ld16(tcol + c0, v);
                    if (!valid) continue;
                    const int r0 = (n0 + c0) >> 2;
                    const float4 cold = *reinterpret_cast<const float4*>(a.cstate + pix * R + r0);
                    const float co[4] = {cold.x, cold.y, cold.z, cold.w};
                    float cn[4], hn[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bq = *reinterpret_cast<const float4*>(a.bias + n0 + c0 + q * 4);
                        const float4 pq = *reinterpret_cast<const float4*>(a.peep + (ppix * R + r0 + q) * 4);
                        hn[q] = lstm_cell_v(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3], bq, pq, co[q], &cn[q]);
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
                for (int c0 = 0; c0 < p.Ncta; c0 += 16) {
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
                for (int c0 = 0; c0 < p.Ncta; c0 += 16) {
                    float v[16];
                    tmem_ld16(tcol + c0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float o = __fadd_rn(v[i], a.bias[n0 + c0 + i]);
                        stage[m * p.stage_ld + c0 + i] = o > 0.f ? o : 0.f;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int Hp = p.H >> 1, Wp = p.W >> 1, tw2 = p.TW >> 1, th2 = p.TH >> 1;
                const int items = th2 * tw2 * p.Ncta;
                for (int idx = m; idx < items; idx += 128) {
                    const int n = idx % p.Ncta, pp = idx / p.Ncta;
                    const int ph = pp / tw2, pw = pp - ph * tw2;
                    const int py = ((y0 + t * p.TH) >> 1) + ph, px = (x0 >> 1) + pw;
                    if (py >= Hp || px >= Wp) continue;
                    const int m00 = (2 * ph) * p.P + 2 * pw;
                    const float* s0 = stage + m00 * p.stage_ld + n;
                    const float mx = fmaxf(fmaxf(s0[0], s0[p.stage_ld]), fmaxf(s0[p.P * p.stage_ld], s0[(p.P + 1) * p.stage_ld]));
                    const long long ppos = ((long long)b * Hp + py) * Wp + px;
                    const float pv = a.P[ppos * a.N + n0 + n];
                    const float ep = __fsub_rn(mx, pv), en = __fsub_rn(pv, mx);
                    view_store(a.dstE, ppos, n0 + n, ep > 0.f ? ep : 0.f);
                    view_store(a.dstE, ppos, a.N + n0 + n, en > 0.f ? en : 0.f);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        tc_fence_before();
    }
    __syncthreads();
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
    int a_mode = 1;  // measured on B200: the 128B swizzle is a function of the absolute smem address, base offset stays 0
    int force_nt = 0;
    std::map<std::tuple<const void*, const void*, int, int, int, int, int, int, int>, std::pair<CUtensorMap, CUtensorMap>> amaps;
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
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        s.reason = "cuTensorMapEncodeTiled not exported by the driver";
        return false;
    }
    s.encode = (EigEncodeTiledFn)fn;
    if (cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess) {
        s.reason = "cannot raise the dynamic shared memory limit";
        cudaGetLastError();
        return false;
    }
    if (const char* e = getenv("EIG_TC_AMODE")) s.a_mode = atoi(e);
    if (const char* e = getenv("EIG_TC_NT")) s.force_nt = atoi(e);
    s.available = true;
    return true;
}
inline std::string tc_unavailable_reason() { return tc_state().reason; }
inline std::string tc_last_error() { return tc_state().last_error; }
inline void tc_set_a_mode(int m) { tc_state().a_mode = m; }

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

// wv: [9][cin][npad] fp32 (the SIMT layout), N valid columns
inline int tc_pack(TcWeights& w, const float* wv, int cin, int N, int npad) {
    if (!tc_available()) return 0;  // no tensor-core path on this device: nothing to pack
    TcState& s = tc_state();
    tc_free(w);
    if (N % 16 || cin % 4) return 0;  // not a tensor-core shape: w.ok stays false, the caller keeps the SIMT kernel
    w.cin = cin; w.N = N; w.KBn = (cin + TC_KB - 1) / TC_KB;
    w.gz = (N + 255) / 256;
    while (N % w.gz || (N / w.gz) % 16) ++w.gz;
    w.Ncta = N / w.gz;
    const size_t plane = (size_t)9 * w.KBn * N * TC_KB;
    std::vector<float> pk(2 * plane, 0.f);
    for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < cin; ++c)
            for (int n = 0; n < N; ++n) {
                const float v = wv[((size_t)tap * cin + c) * npad + n];
                const float hi = tc_host_tf32(v);
                const size_t o = (((size_t)tap * w.KBn + c / TC_KB) * N + n) * TC_KB + c % TC_KB;
                pk[o] = hi;
                pk[plane + o] = v - hi;
            }
    if (cudaMalloc((void**)&w.d, pk.size() * sizeof(float)) != cudaSuccess) { s.last_error = "tc_pack: cudaMalloc failed"; return -1; }
    if (cudaMemcpy(w.d, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { s.last_error = "tc_pack: upload failed"; return -1; }
    const cuuint64_t gdim[2] = {TC_KB, (cuuint64_t)2 * 9 * w.KBn * N};
    const cuuint64_t gstr[1] = {TC_KB * sizeof(float)};
    const cuuint32_t box[2] = {TC_KB, (cuuint32_t)w.Ncta};
    const cuuint32_t est[2] = {1, 1};
    const CUresult r = s.encode(&w.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w.d, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { s.last_error = "tc_pack: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -1; }
    w.ok = true;
    return 0;
}

struct TcGeom { int TW, TH, P, NT, SA, SB, a_plane, b_plane, tmem_cols, tiles_x, tiles_y; size_t smem; };

inline int tc_round_up(int v, int m) { return (v + m - 1) / m * m; }

// picks the flat-padded tile (TW x TH, P = TW + 2, (TH-1)*P + TW <= 128) with the best MMA-row efficiency
inline bool tc_geometry(int B, int H, int W, int Ncta, int gz, bool pooled, int force_nt, TcGeom& g) {
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
    g.b_plane = Ncta * 128;
    int nt_max = 512 / Ncta;
    if (nt_max > row_tiles) nt_max = row_tiles;
    for (int NT = nt_max; NT >= 1; --NT) {
        const int rows = std::max((NT * g.TH + 2) * g.P, (NT - 1) * g.TH * g.P + 2 * g.P + 2 + 128);
        const int a_plane = tc_round_up(rows * 128, 1024);
        const int SA = 2;
        const long long left = (long long)TC_SMEM_LIMIT - 2048 - (long long)SA * 2 * a_plane;
        int SB = (int)(left / (2 * g.b_plane));
        if (SB > 8) SB = 8;
        const long long ctas = (long long)g.tiles_x * ((row_tiles + NT - 1) / NT) * B * gz;
        const bool fits = SB >= 2;
        const bool enough = ctas >= 2 * 148 || NT == 1;
        if (force_nt > 0 ? (NT <= force_nt && fits) : (fits && enough)) {
            g.NT = NT; g.SA = SA; g.SB = SB; g.a_plane = a_plane;
            g.tiles_y = (row_tiles + NT - 1) / NT;
            int cols = 32;
            while (cols < NT * Ncta) cols <<= 1;
            g.tmem_cols = cols;
            g.smem = (size_t)SA * 2 * a_plane + (size_t)SB * 2 * g.b_plane + 8 * (2 * SA + 2 * SB + 1) + 16 + 1024;
            return true;
        }
    }
    return false;
}

inline int tc_conv(const TcWeights& w, const ConvArgs& a, cudaStream_t stream) {
    TcState& s = tc_state();
    if (!tc_available()) { s.last_error = s.reason; return -1; }
    if (!w.ok) { s.last_error = "tc_conv: weights not packed"; return -1; }
    if (!a.in_lo) { s.last_error = "tc_conv: input view has no lo plane"; return -1; }
    if (a.Cin != w.cin || a.N != w.N) { s.last_error = "tc_conv: shape mismatch with packed weights"; return -1; }
    if ((a.in_coff & 3) || (a.in_pitch & 3)) { s.last_error = "tc_conv: view not 16-byte aligned"; return -1; }
    const bool pooled = a.epi == EPI_CONVA;
    if (pooled && ((a.H | a.W) & 1)) { s.last_error = "tc_conv: pooled conv needs even H, W"; return -1; }
    TcGeom g;
    if (!tc_geometry(a.B, a.H, a.W, w.Ncta, w.gz, pooled, s.force_nt, g)) { s.last_error = "tc_conv: no tile geometry fits"; return -1; }
    const int box_rows = g.NT * g.TH + 2;
    auto key = std::make_tuple((const void*)(a.in_hi + a.in_coff), (const void*)(a.in_lo + a.in_coff), a.Cin, a.W, a.H, a.B, a.in_pitch, g.P, box_rows);
    auto it = s.amaps.find(key);
    if (it == s.amaps.end()) {
        std::pair<CUtensorMap, CUtensorMap> maps;
        const cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        const cuuint64_t gstr[3] = {(cuuint64_t)a.in_pitch * 4, (cuuint64_t)a.W * a.in_pitch * 4, (cuuint64_t)a.H * a.W * a.in_pitch * 4};
        const cuuint32_t box[4] = {TC_KB, (cuuint32_t)g.P, (cuuint32_t)box_rows, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        for (int k = 0; k < 2; ++k) {
            const float* base = (k == 0 ? a.in_hi : a.in_lo) + a.in_coff;
            const CUresult r = s.encode(k == 0 ? &maps.first : &maps.second, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstr, box,
                                        est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { s.last_error = "tc_conv: cuTensorMapEncodeTiled(A) failed (" + std::to_string((int)r) + ")"; return -1; }
        }
        it = s.amaps.emplace(key, maps).first;
    }
    TcParams p;
    memset(&p, 0, sizeof p);
    p.B = a.B; p.H = a.H; p.W = a.W;
    p.TW = g.TW; p.TH = g.TH; p.P = g.P; p.NT = g.NT; p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y;
    p.KBn = w.KBn; p.Ncta = w.Ncta; p.N = w.N;
    p.a_plane_bytes = g.a_plane; p.b_plane_bytes = g.b_plane; p.a_box_bytes = TC_KB * 4 * g.P * box_rows;
    p.SA = g.SA; p.SB = g.SB; p.a_mode = s.a_mode; p.tmem_cols = g.tmem_cols;
    p.stage_ld = w.Ncta + 1;
    p.ca = a;
    size_t smem = g.smem;
    if (pooled) {
        const size_t need = (size_t)128 * p.stage_ld * 4 + 2048;
        if (need > smem) smem = need;
        if ((size_t)128 * p.stage_ld * 4 > (size_t)g.SA * 2 * g.a_plane + (size_t)g.SB * 2 * g.b_plane) {
            s.last_error = "tc_conv: staging tile would overlap the barriers";
            return -1;
        }
    }
    if (smem > TC_SMEM_LIMIT) { s.last_error = "tc_conv: shared memory budget exceeded"; return -1; }
    conv3x3_tc_kernel<<<dim3(g.tiles_x * g.tiles_y, a.B, w.gz), dim3(TC_THREADS), smem, stream>>>(it->second.first, it->second.second, w.map, p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s.last_error = std::string("tc_conv launch: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

}  // namespace eig
