// PredNet layer-0 kernels (full image resolution, C0 = 1 or 3 channels): exact-fp32 SIMT, HBM/L2-bound.
// These stages have no tensor-core shape (K <= 9*57, N <= 48) - they are fused so that every pixel-sized tensor is
// read and written once per time step:
//   l0_conva1_kernel : E0 = [relu(x-P0), relu(P0-x)] on the fly -> ConvA1 (2*C0 -> C1) -> relu -> 2x2 max-pool
//                      -> E1 = [relu(A1-P1), relu(P1-A1)] into the layer-1 concat buffer        (net.py:187-194)
//   l0_lstm_kernel   : ConvLSTM0 on [E0 | up2x(R1) | h0]: E0 recomputed from x and P0; the up2x(R1) taps arrive as
//                      half-resolution partial sums Z computed on the tensor cores next to ConvP1 (no up-sampled copy
//                      in HBM, 6x fewer CUDA-core MACs); 4 gates + cell update fused          (net.py:94-126,202-203)
//   l0_convp_kernel  : P0 = min(relu(ConvP0(h0)), 1)                                              (net.py:207)
// Accumulation order: input channel, then ky, then kx, with fused multiply-add (same as conv_simt.cuh).
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"

namespace eig {

struct L0Args {
    int B, H, W, C0, C1;
    const float* x;        // [B,H,W,C0] input frame of this step (== P0 on the self-fed extension steps)
    const float* P0;       // [B,H,W,C0] prediction of the previous step
    // ConvA1
    const float* wA;       // [9][2*C0][C1pad]
    const float* bA;       // [C1]
    int C1pad;
    const float* P1;       // [B,H/2,W/2,C1]
    View dstE1;            // layer-1 concat buffer, channels [0, 2*C1)
    // ConvLSTM0
    const float* wL;       // [9][ctot0][4*C0], ctot0 = 2*C0 + C1 + C0
    const float* bL;       // [4*C0] gate-interleaved
    const float* peep;     // [H,W,C0,4]
    const float* Z;        // [B,H/2,W/2,4*4*C0] half-resolution partial sums of the up-sampled-R1 taps (see l0_lstm_kernel)
    const float* h_prev;   // [B,H,W,C0]
    float* h_next;         // [B,H,W,C0]
    float* cstate;         // [B,H,W,C0]
    // ConvP0
    const float* wP;       // [9][C0][C0pad]
    const float* bP;       // [C0]
    int C0pad;
    float* P0_out;         // [B,H,W,C0]
};

enum { L0_TW = 32, L0_TH = 8 };

// ---------------------------------------------------------------------------------------------- E0 for the tensor cores
// E0 = [relu(x - P0), relu(P0 - x)] as an 8-channel split-fp16 tensor (channels 2*C0 .. 7 zero): the input of ConvA1 when
// it runs on the tcgen05 kernel (wide first layers, C1 >= 32).  hi / lo: [B*H*W][8] fp16 planes.
__global__ void __launch_bounds__(256) l0_e0_kernel(const float* x, const float* P0, h16* hi, h16* lo, long long npix, int C0) {
    EIG_PDL_WAIT();   // (no early trigger: multi-wave grid - a successor holding whole SMs would starve our later CTAs)
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    h16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = 0; l[i] = 0; }
    for (int c = 0; c < C0; ++c) {
        const float xv = x[p * C0 + c], pv = P0[p * C0 + c];
        const float ep = __fsub_rn(xv, pv), en = __fsub_rn(pv, xv);
        split16(ep > 0.f ? ep : 0.f, &h[c], &l[c]);
        split16(en > 0.f ? en : 0.f, &h[C0 + c], &l[C0 + c]);
    }
    uint4 uh, ul;
    uh.x = h[0] | ((unsigned)h[1] << 16); uh.y = h[2] | ((unsigned)h[3] << 16); uh.z = h[4] | ((unsigned)h[5] << 16); uh.w = h[6] | ((unsigned)h[7] << 16);
    ul.x = l[0] | ((unsigned)l[1] << 16); ul.y = l[2] | ((unsigned)l[3] << 16); ul.z = l[4] | ((unsigned)l[5] << 16); ul.w = l[6] | ((unsigned)l[7] << 16);
    reinterpret_cast<uint4*>(hi)[p] = uh;
    reinterpret_cast<uint4*>(lo)[p] = ul;
}

// ---------------------------------------------------------------------------------------------- ConvA1
// Persistent: each CTA keeps the ConvA1 weights in shared memory and walks work items = (32x8 pixel tile, genome),
// item = blockIdx.x, blockIdx.x + gridDim.x, ...  The x / P0 halo of the NEXT item is fetched into registers while the
// current one is computed (double-buffered E0 tile).  thread = one pooled pixel x CPT output channels.
enum { L0_PF = 4 };   // prefetch registers per thread per tensor: 340 halo positions x C0 channels / blockDim <= 4
template <int CPT>
__global__ void __launch_bounds__(256) l0_conva1_kernel(L0Args a, int n_items) {
    constexpr int SW = L0_TW + 2, SH = L0_TH + 2;
    EIG_DYN_SMEM(smem);
    const int cin = 2 * a.C0;
    const int e_sz = cin * SH * SW;
    float* sE = reinterpret_cast<float*>(smem);                 // [2 buffers][2*C0][SH][SW]
    float* sW = sE + 2 * e_sz;                                  // [9][2*C0][C1pad]
    float* sOut = sW + 9 * cin * a.C1pad;                       // [64 pooled pixels][C1pad + 1]
    const int ldo = a.C1pad + 1;   // odd pitch: the 64 pooled pixels of a warp pair hit different banks
    const int tiles_x = (a.W + L0_TW - 1) / L0_TW;
    const int tiles = tiles_x * ((a.H + L0_TH - 1) / L0_TH);
    const int n_halo = SH * SW * a.C0;                          // (position, channel) pairs of one halo tile
    EIG_PDL_TRIGGER();
    for (int i = threadIdx.x; i < 9 * cin * a.C1pad; i += blockDim.x) sW[i] = a.wA[i];   // constants: before the wait
    EIG_PDL_WAIT();

    // the halo slots a thread prefetches are the same for every item: decode them once
    float pf_x[L0_PF], pf_p[L0_PF];
    int pf_cy[L0_PF], pf_cx[L0_PF], pf_dst[L0_PF], pf_src[L0_PF];   // halo row / column, smem offset, global offset
#pragma unroll
    for (int k = 0; k < L0_PF; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        const int c = i % a.C0, pos = i / a.C0;
        pf_cy[k] = i < n_halo ? pos / SW : -100000;          // an impossible row: never inside the image
        pf_cx[k] = pos % SW;
        pf_dst[k] = c * SH * SW + pos;
        pf_src[k] = ((pf_cy[k] - 1) * a.W + (pf_cx[k] - 1)) * a.C0 + c;
    }
    auto fetch = [&](int item) {   // halo of `item` -> registers (zeros outside the image)
        const int b = item / tiles, tile = item - b * tiles;
        const int x0 = (tile % tiles_x) * L0_TW, y0 = (tile / tiles_x) * L0_TH;
        const long long base = (((long long)b * a.H + y0) * a.W + x0) * a.C0;
#pragma unroll
        for (int k = 0; k < L0_PF; ++k) {
            const int gy = y0 + pf_cy[k] - 1, gx = x0 + pf_cx[k] - 1;
            const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
            pf_x[k] = in ? a.x[base + pf_src[k]] : 0.f;
            pf_p[k] = in ? a.P0[base + pf_src[k]] : 0.f;
        }
    };
    auto stash = [&](float* dst) {  // registers -> E0 = [relu(x - P0), relu(P0 - x)] tile in shared memory
#pragma unroll
        for (int k = 0; k < L0_PF; ++k) {
            if (pf_cy[k] >= 0) {
                float ep = __fsub_rn(pf_x[k], pf_p[k]), en = __fsub_rn(pf_p[k], pf_x[k]);
                ep = ep > 0.f ? ep : 0.f; en = en > 0.f ? en : 0.f;
                dst[pf_dst[k]] = ep;
                dst[a.C0 * SH * SW + pf_dst[k]] = en;
            }
        }
    };
    // small test networks run fewer threads than the halo needs registers for: they stage without the prefetch
    const bool pf_ok = n_halo <= L0_PF * (int)blockDim.x;
    auto stage_direct = [&](int it2, float* dst) {
        const int b = it2 / tiles, tile = it2 - b * tiles;
        const int x0 = (tile % tiles_x) * L0_TW, y0 = (tile / tiles_x) * L0_TH;
        const long long img = (long long)b * a.H * a.W;
        for (int i = threadIdx.x; i < n_halo; i += blockDim.x) {
            const int c = i % a.C0, pos = i / a.C0;
            const int cy = pos / SW, cx = pos - cy * SW;
            const int gy = y0 + cy - 1, gx = x0 + cx - 1;
            float ep = 0.f, en = 0.f;
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                const long long idx = (img + (long long)gy * a.W + gx) * a.C0 + c;
                const float xv = a.x[idx], pv = a.P0[idx];
                ep = __fsub_rn(xv, pv); en = __fsub_rn(pv, xv);
                ep = ep > 0.f ? ep : 0.f; en = en > 0.f ? en : 0.f;
            }
            dst[c * SH * SW + pos] = ep;
            dst[(a.C0 + c) * SH * SW + pos] = en;
        }
    };
    int item = blockIdx.x, cur = 0;
    if (item < n_items) {
        if (pf_ok) { fetch(item); stash(sE); }
        else stage_direct(item, sE);
    }
    __syncthreads();

    const int pp = threadIdx.x & 63, grp = threadIdx.x >> 6;    // pooled pixel in the tile, channel group
    const int px = pp & 15, py = pp >> 4;
    const int n0 = grp * CPT;
    const int Hp = a.H >> 1, Wp = a.W >> 1;
    const int c2 = 2 * a.C1;
    for (; item < n_items; item += gridDim.x) {
        const int nxt = item + gridDim.x;
        if (pf_ok && nxt < n_items) fetch(nxt);                 // global loads in flight during the FMAs below
        const float* sEc = sE + cur * e_sz;
        float acc[4][CPT];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int n = 0; n < CPT; ++n) acc[j][n] = 0.f;
        for (int c = 0; c < cin; ++c) {
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) in[r][q] = sEc[(c * SH + 2 * py + r) * SW + 2 * px + q];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wrow = sW + ((ky * 3 + kx) * cin + c) * a.C1pad + n0;
                    float wv[CPT];
#pragma unroll
                    for (int n = 0; n < CPT; n += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(wrow + n);
                        wv[n] = q.x; wv[n + 1] = q.y; wv[n + 2] = q.z; wv[n + 3] = q.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int n = 0; n < CPT; ++n)
                            acc[j][n] = __fmaf_rn(in[(j >> 1) + ky][(j & 1) + kx], wv[n], acc[j][n]);
                }
        }
        // pooled, biased, rectified A1 -> shared memory, then the error units are written with consecutive threads on
        // consecutive channels of one pixel (coalesced rows of the layer-1 concat buffer)
#pragma unroll
        for (int n = 0; n < CPT; ++n) {
            if (n0 + n >= a.C1) continue;
            const float bn = a.bA[n0 + n];
            float m = fmaxf(fmaxf(__fadd_rn(acc[0][n], bn), __fadd_rn(acc[1][n], bn)),
                            fmaxf(__fadd_rn(acc[2][n], bn), __fadd_rn(acc[3][n], bn)));
            sOut[pp * ldo + n0 + n] = fmaxf(m, 0.f);  // relu commutes with max
        }
        __syncthreads();
        {
            const int b = item / tiles, tile = item - b * tiles;
            const int x0 = (tile % tiles_x) * L0_TW, y0 = (tile / tiles_x) * L0_TH;
            const long long prow = ((long long)b * Hp + (y0 >> 1)) * Wp + (x0 >> 1);
            if ((a.C1 & 3) == 0) {
                // item = (pooled pixel q, group of 4 channels): one float4 P1 load, two vector stores (E+ and E-); the
                // walk advances (q, g) by a fixed step, so it needs no division; 4 items per batch keep the loads in flight
                const int g4 = a.C1 >> 2;
                const int dq = (int)blockDim.x / g4, dg = (int)blockDim.x - dq * g4;
                int q = (int)threadIdx.x / g4, g = (int)threadIdx.x - q * g4;
                while (q < 64) {
                    float4 pv[4];
                    float m[4][4];
                    long long ppos[4];
                    int ch0[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        ppos[u] = -1;
                        if (q < 64) {
                            const int gpy = (y0 >> 1) + (q >> 4), gpx = (x0 >> 1) + (q & 15);
                            if (gpy < Hp && gpx < Wp) {
                                ppos[u] = prow + (long long)(q >> 4) * Wp + (q & 15);
                                ch0[u] = 4 * g;
#pragma unroll
                                for (int k = 0; k < 4; ++k) m[u][k] = sOut[q * ldo + 4 * g + k];
                                pv[u] = *reinterpret_cast<const float4*>(a.P1 + ppos[u] * a.C1 + 4 * g);
                            }
                        }
                        q += dq; g += dg;
                        if (g >= g4) { g -= g4; ++q; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (ppos[u] < 0) continue;
                        const float p4[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
                        float ep[4], en[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float d1 = __fsub_rn(m[u][k], p4[k]), d2 = __fsub_rn(p4[k], m[u][k]);
                            ep[k] = d1 > 0.f ? d1 : 0.f; en[k] = d2 > 0.f ? d2 : 0.f;
                        }
                        view_store_vec4(a.dstE1, ppos[u], ch0[u], ep);
                        view_store_vec4(a.dstE1, ppos[u], a.C1 + ch0[u], en);
                    }
                }
            } else {
                // scalar walk for channel counts that are not a multiple of 4 (small test networks)
                const int dq = (int)blockDim.x / c2, dch = (int)blockDim.x - dq * c2;
                int q = (int)threadIdx.x / c2, ch = (int)threadIdx.x - q * c2;
                while (q < 64) {
                    const int gpy = (y0 >> 1) + (q >> 4), gpx = (x0 >> 1) + (q & 15);
                    if (gpy < Hp && gpx < Wp) {
                        const long long ppos = prow + (long long)(q >> 4) * Wp + (q & 15);
                        const int n = ch < a.C1 ? ch : ch - a.C1;
                        const float m = sOut[q * ldo + n], pv = a.P1[ppos * a.C1 + n];
                        const float e = ch < a.C1 ? __fsub_rn(m, pv) : __fsub_rn(pv, m);
                        view_store(a.dstE1, ppos, ch, e > 0.f ? e : 0.f);
                    }
                    q += dq; ch += dch;
                    if (ch >= c2) { ch -= c2; ++q; }
                }
            }
        }
        if (nxt < n_items) {
            if (pf_ok) stash(sE + (cur ^ 1) * e_sz);
            else stage_direct(nxt, sE + (cur ^ 1) * e_sz);
        }
        __syncthreads();
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------- ConvA1, narrow first layers
// One CTA per (32x8 pixel tile, genome), no persistence and no register prefetch: for C1 <= 16 (the gray BASELINE network,
// C0 = 1, C1 = 16) the persistent kernel above needs 161 registers per thread = 3 CTAs of 128 threads per SM, and a
// ConvA1 that is 10 % of a C2 PredNet step then runs on 12 warps per SM (measured 26-29 us for 177 MFMA).  This variant
// stays under 80 registers (6+ CTAs per SM) and lets the block scheduler hide the global-memory latency instead.
// Same arithmetic, same accumulation order (input channel, ky, kx; fused multiply-add) as l0_conva1_kernel.
template <int CPT, int C0>
__global__ void __launch_bounds__(128, 6) l0_conva1_tile_kernel(L0Args a) {
    constexpr int SW = L0_TW + 2, SH = L0_TH + 2, NPOS = SH * SW;
    EIG_DYN_SMEM(smem);
    constexpr int cin = 2 * C0;
    float* sE = reinterpret_cast<float*>(smem);                 // [2*C0][SH][SW]
    float* sW = sE + cin * SH * SW;                             // [9][2*C0][C1pad]
    float* sOut = sW + 9 * cin * a.C1pad;                       // [64 pooled pixels][C1pad + 1]
    const int ldo = a.C1pad + 1;
    const int tiles_x = (a.W + L0_TW - 1) / L0_TW;
    const int x0 = (blockIdx.x % tiles_x) * L0_TW, y0 = (blockIdx.x / tiles_x) * L0_TH;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 9 * cin * a.C1pad; i += blockDim.x) sW[i] = a.wA[i];   // constants: before the wait
    EIG_PDL_WAIT();
    {   // halo staging, one position (all C0 channels) per thread and round; every global load of a thread is issued before
        // the first one is used (the round count is a compile-time constant for the thread counts this kernel is launched with)
        const long long img = (long long)b * a.H * a.W;
        constexpr int ROUNDS = (NPOS + 63) / 64;                // blockDim >= 64
        float xv[ROUNDS][C0], pv[ROUNDS][C0];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pos = (int)threadIdx.x + r * (int)blockDim.x;
            const int cy = pos / SW, cx = pos - cy * SW;
            const int gy = y0 + cy - 1, gx = x0 + cx - 1;
            const bool in = pos < NPOS && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
            const long long idx = (img + (long long)gy * a.W + gx) * C0;
#pragma unroll
            for (int c = 0; c < C0; ++c) { xv[r][c] = in ? a.x[idx + c] : 0.f; pv[r][c] = in ? a.P0[idx + c] : 0.f; }
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pos = (int)threadIdx.x + r * (int)blockDim.x;
            if (pos < NPOS) {
#pragma unroll
                for (int c = 0; c < C0; ++c) {
                    float ep = __fsub_rn(xv[r][c], pv[r][c]), en = __fsub_rn(pv[r][c], xv[r][c]);
                    ep = ep > 0.f ? ep : 0.f; en = en > 0.f ? en : 0.f;
                    sE[c * NPOS + pos] = ep;
                    sE[(C0 + c) * NPOS + pos] = en;
                }
            }
        }
    }
    __syncthreads();
    const int pp = threadIdx.x & 63, grp = threadIdx.x >> 6;    // pooled pixel in the tile, channel group
    const int px = pp & 15, py = pp >> 4;
    const int n0 = grp * CPT;
    float acc[4][CPT];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < CPT; ++n) acc[j][n] = 0.f;
    if (n0 < a.C1pad) {
        for (int c = 0; c < cin; ++c) {
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) in[r][q] = sE[(c * SH + 2 * py + r) * SW + 2 * px + q];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wrow = sW + ((ky * 3 + kx) * cin + c) * a.C1pad + n0;
                    float wv[CPT];
#pragma unroll
                    for (int n = 0; n < CPT; n += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(wrow + n);
                        wv[n] = q.x; wv[n + 1] = q.y; wv[n + 2] = q.z; wv[n + 3] = q.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int n = 0; n < CPT; ++n)
                            acc[j][n] = __fmaf_rn(in[(j >> 1) + ky][(j & 1) + kx], wv[n], acc[j][n]);
                }
        }
#pragma unroll
        for (int n = 0; n < CPT; ++n) {
            if (n0 + n >= a.C1) continue;
            const float bn = a.bA[n0 + n];
            const float m = fmaxf(fmaxf(__fadd_rn(acc[0][n], bn), __fadd_rn(acc[1][n], bn)),
                                  fmaxf(__fadd_rn(acc[2][n], bn), __fadd_rn(acc[3][n], bn)));
            sOut[pp * ldo + n0 + n] = fmaxf(m, 0.f);  // relu commutes with max
        }
    }
    __syncthreads();
    // error units: consecutive threads on consecutive channels of one pooled pixel (coalesced rows of the concat buffer)
    const int Hp = a.H >> 1, Wp = a.W >> 1;
    const long long prow = ((long long)b * Hp + (y0 >> 1)) * Wp + (x0 >> 1);
    if ((a.C1 & 3) == 0 && (int)blockDim.x % (a.C1 >> 2) == 0) {
        const int g4 = a.C1 >> 2;
        const int g = (int)threadIdx.x % g4, dq = (int)blockDim.x / g4;   // a thread keeps its channel group and strides over pixels
        for (int q = (int)threadIdx.x / g4; q < 64; q += dq) {
            const int gpy = (y0 >> 1) + (q >> 4), gpx = (x0 >> 1) + (q & 15);
            if (gpy >= Hp || gpx >= Wp) continue;
            const long long ppos = prow + (long long)(q >> 4) * Wp + (q & 15);
            const float4 pv = *reinterpret_cast<const float4*>(a.P1 + ppos * a.C1 + 4 * g);
            const float p4[4] = {pv.x, pv.y, pv.z, pv.w};
            float ep[4], en[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float m = sOut[q * ldo + 4 * g + k];
                const float d1 = __fsub_rn(m, p4[k]), d2 = __fsub_rn(p4[k], m);
                ep[k] = d1 > 0.f ? d1 : 0.f; en[k] = d2 > 0.f ? d2 : 0.f;
            }
            view_store_vec4(a.dstE1, ppos, 4 * g, ep);
            view_store_vec4(a.dstE1, ppos, a.C1 + 4 * g, en);
        }
    } else {
        const int c2 = 2 * a.C1;
        for (int i = threadIdx.x; i < 64 * c2; i += blockDim.x) {
            const int q = i / c2, ch = i - q * c2;
            const int gpy = (y0 >> 1) + (q >> 4), gpx = (x0 >> 1) + (q & 15);
            if (gpy >= Hp || gpx >= Wp) continue;
            const long long ppos = prow + (long long)(q >> 4) * Wp + (q & 15);
            const int n = ch < a.C1 ? ch : ch - a.C1;
            const float m = sOut[q * ldo + n], pv = a.P1[ppos * a.C1 + n];
            const float e = ch < a.C1 ? __fsub_rn(m, pv) : __fsub_rn(pv, m);
            view_store(a.dstE1, ppos, ch, e > 0.f ? e : 0.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------- ConvLSTM0
// The 3x3 taps over the nearest-neighbour up-sampled R1 are NOT evaluated here: for the four pixel parities they
// collapse to 2x2 taps over R1 at half resolution, i.e. to one 3x3 convolution of R1 with 4 * NG output columns
// (column = parity * NG + gate column; weights pre-summed at load time, eig_api.cu:build_z_weights).  That convolution
// rides along with ConvP1 on the layer-1 conv kernel (ConvArgs::outZ) and this kernel starts its accumulators from Z,
// leaving only the E0 and h0 taps (3*C0 input channels) for the CUDA cores.
// CTA = 32x8 pixels of one genome, 128 threads, thread = 1x2 pixel strip x all NG = 4*C0 gate columns.
template <int C0>
__global__ void __launch_bounds__(128) l0_lstm_kernel(L0Args a) {
    constexpr int NG = 4 * C0, CIN = 3 * C0, SW = L0_TW + 2, SH = L0_TH + 2;
    __shared__ float sIn[CIN][SH][SW];
    __shared__ __align__(16) float sWt[9 * CIN * NG];
    const int tiles_x = (a.W + L0_TW - 1) / L0_TW;
    const int x0 = (blockIdx.x % tiles_x) * L0_TW, y0 = (blockIdx.x / tiles_x) * L0_TH;
    const int b = blockIdx.y;
    const int ctot = 2 * C0 + a.C1 + C0;
    const long long img = (long long)b * a.H * a.W;
    // (no early trigger here: the grid has more CTAs than fit at once, and a tcgen05 successor holding whole SMs would
    // starve the later ones - measured slower)
    for (int i = threadIdx.x; i < 9 * CIN * NG; i += blockDim.x) {   // weights are constants: staged before the wait
        const int n = i % NG, r = i / NG;
        const int k = r % CIN, tap = r / CIN;
        const int c = k < 2 * C0 ? k : k + a.C1;   // skip the R1 channels of the full weight tensor
        sWt[i] = a.wL[((long long)tap * ctot + c) * NG + n];
    }
    EIG_PDL_WAIT();
    for (int i = threadIdx.x; i < SH * SW * C0; i += blockDim.x) {
        const int c = i % C0, pp = i / C0;
        const int cy = pp / SW, cx = pp - cy * SW;
        const int gy = y0 + cy - 1, gx = x0 + cx - 1;
        float ep = 0.f, en = 0.f, hv = 0.f;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
            const long long idx = (img + (long long)gy * a.W + gx) * C0 + c;
            const float xv = a.x[idx], pv = a.P0[idx];
            ep = __fsub_rn(xv, pv); en = __fsub_rn(pv, xv);
            ep = ep > 0.f ? ep : 0.f; en = en > 0.f ? en : 0.f;
            hv = a.h_prev[idx];
        }
        sIn[c][cy][cx] = ep;
        sIn[C0 + c][cy][cx] = en;
        sIn[2 * C0 + c][cy][cx] = hv;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // strip: columns 2*tx, 2*tx+1 of row ty
    const int gy = y0 + ty, gx0 = x0 + 2 * tx;
    if (gy >= a.H || gx0 >= a.W) return;
    const int H1 = a.H >> 1, W1 = a.W >> 1;
    float acc[2][NG];
    {   // both pixels of the strip share the half-resolution pixel; parity = (y & 1) * 2 + (x & 1)
        const float* z = a.Z + (((long long)b * H1 + (gy >> 1)) * W1 + (gx0 >> 1)) * (4 * NG) + (gy & 1) * 2 * NG;
#pragma unroll
        for (int n = 0; n < NG; n += 4) {
            const float4 z0 = *reinterpret_cast<const float4*>(z + n), z1 = *reinterpret_cast<const float4*>(z + NG + n);
            acc[0][n] = z0.x; acc[0][n + 1] = z0.y; acc[0][n + 2] = z0.z; acc[0][n + 3] = z0.w;
            acc[1][n] = z1.x; acc[1][n + 1] = z1.y; acc[1][n + 2] = z1.z; acc[1][n + 3] = z1.w;
        }
    }
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float in[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) in[q] = sIn[k][ty + ky][2 * tx + q];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* wrow = sWt + ((ky * 3 + kx) * CIN + k) * NG;
#pragma unroll
                for (int n = 0; n < NG; n += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(wrow + n);
                    acc[0][n] = __fmaf_rn(in[kx], q.x, acc[0][n]); acc[0][n + 1] = __fmaf_rn(in[kx], q.y, acc[0][n + 1]);
                    acc[0][n + 2] = __fmaf_rn(in[kx], q.z, acc[0][n + 2]); acc[0][n + 3] = __fmaf_rn(in[kx], q.w, acc[0][n + 3]);
                    acc[1][n] = __fmaf_rn(in[kx + 1], q.x, acc[1][n]); acc[1][n + 1] = __fmaf_rn(in[kx + 1], q.y, acc[1][n + 1]);
                    acc[1][n + 2] = __fmaf_rn(in[kx + 1], q.z, acc[1][n + 2]); acc[1][n + 3] = __fmaf_rn(in[kx + 1], q.w, acc[1][n + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int gx = gx0 + j;
        if (gx >= a.W) continue;
        const long long pix = img + (long long)gy * a.W + gx;
        const long long ppix = (long long)gy * a.W + gx;
#pragma unroll
        for (int r = 0; r < C0; ++r) {
            const float hnew = lstm_cell(acc[j][r * 4], acc[j][r * 4 + 1], acc[j][r * 4 + 2], acc[j][r * 4 + 3], a.bL + r * 4,
                                         a.peep + (ppix * C0 + r) * 4, a.cstate + pix * C0 + r);
            a.h_next[pix * C0 + r] = hnew;
        }
    }
}

// ---------------------------------------------------------------------------------------------- ConvP0
template <int C0>
__global__ void __launch_bounds__(256) l0_convp_kernel(L0Args a) {
    __shared__ float sW[9 * C0 * C0];
    __shared__ float sB[C0];
    for (int i = threadIdx.x; i < 9 * C0 * C0; i += blockDim.x) {
        const int n = i % C0, r = i / C0;
        sW[i] = a.wP[(long long)r * a.C0pad + n];
    }
    if (threadIdx.x < C0) sB[threadIdx.x] = a.bP[threadIdx.x];
    EIG_PDL_WAIT();
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)a.B * a.H * a.W;
    if (p >= npix) return;
    const int gx = (int)(p % a.W), gy = (int)((p / a.W) % a.H);
    float acc[C0];
#pragma unroll
    for (int n = 0; n < C0; ++n) acc[n] = 0.f;
#pragma unroll
    for (int c = 0; c < C0; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yy = gy + ky - 1, xx = gx + kx - 1;
                float v = 0.f;
                if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) v = a.h_next[(p + (long long)(ky - 1) * a.W + (kx - 1)) * C0 + c];
#pragma unroll
                for (int n = 0; n < C0; ++n) acc[n] = __fmaf_rn(v, sW[((ky * 3 + kx) * C0 + c) * C0 + n], acc[n]);
            }
#pragma unroll
    for (int n = 0; n < C0; ++n) {
        float v = __fadd_rn(acc[n], sB[n]);
        v = v > 0.f ? v : 0.f;
        a.P0_out[p * C0 + n] = v > 1.f ? 1.f : v;
    }
}

}  // namespace eig
