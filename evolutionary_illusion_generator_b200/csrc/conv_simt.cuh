// fp32 SIMT 3x3 convolution (pad 1, stride 1, cross-correlation) on NHWC activations with the three fused
// PredNet epilogues.  This is the exact-fp32 path: it serves the tiny-channel layer-0 stages permanently
// (ConvA1, ConvLSTM0, ConvP0: K <= 513, N <= 48, HBM/L2-bound - no tensor-core shape) and every layer when
// the engine runs in `conv_mode = SIMT` (parity debugging, and the baseline the tcgen05 kernel is checked
// against on the GPU).
//
// Reference semantics: `L.Convolution2D(cin, cout, 3, pad=1)` (/root/reference/chainer_prednet/PredNet/
// net.py:46-62,145-147), the ConvLSTM gate equations (net.py:94-126), the error units, 2x2 max-pool and
// nearest x2 up-sampling of `PredNet.__call__` (net.py:187-209).
//
// Tiling: one CTA = a 16x8 pixel tile of one genome x (NW * TN) output channels.  A warp covers the whole
// pixel tile (lane -> column tx = lane&15, rows (lane>>4)*4 .. +3) and owns TN consecutive output channels,
// so weight reads are warp-wide broadcasts and each lane keeps a 4 x TN accumulator block in registers.
// Input channels are streamed through shared memory CK = 8 at a time together with their 9 x CK x (NW*TN)
// weight slab.
#pragma once
#include "common.cuh"

namespace eig {

enum { EPI_CONVP = 0, EPI_CONVA = 1, EPI_LSTM = 2 };

struct ConvArgs {
    // input view (B, H, W, Cin) inside a buffer with `in_pitch` floats per pixel
    const float* in_hi;
    const float* in_lo;  // nullable: non-null = split-fp16 storage, in_hi / in_lo are the two fp16 planes (common.cuh View)
    int in_pitch, in_coff, Cin;
    int B, H, W;
    const float* wgt;   // [9][Cin][Npad], Npad = N rounded up to a multiple of 4
    const float* bias;  // [N] (gate-interleaved for LSTM)
    int N, Npad;
    int epi;
    // EPI_CONVP: out = relu(acc + b) (clipped to 1 when clip != 0) -> outP [B,H,W,nP] plain fp32.
    // When nP != 0 only the first nP columns are ConvP outputs; columns [nP, N) are written raw (no bias, no relu) to
    // outZ [B,H,W,N-nP]: the half-resolution partial sums of ConvLSTM0's up-sampled-R1 taps (conv_l0.cuh).
    float* outP;
    int clip;
    int nP;
    float* outZ;
    // EPI_CONVA: A = maxpool2x2(relu(acc + b)); E = [relu(A-P), relu(P-A)] -> dstE view at (H/2, W/2)
    const float* P;  // [B, H/2, W/2, N]
    View dstE;
    // EPI_LSTM (N = 4R, column n = r*4 + gate, gates i,f,c,o): state c [B,H,W,R] in place,
    // peephole [H,W,R,4] (i,f,o,unused); h -> dstH (same res) and, if dstUp.hi, 2x2 replicated into dstUp
    float* cstate;
    const float* peep;
    View dstH;
    View dstUp;
    // tcgen05 kernel only: half-resolution partial sums [B][H/2][W/2][4 parities][N] of the taps over the up-sampled
    // R_{n+1}, added to the gate pre-activations (nullptr = those taps are part of this convolution's K range)
    const float* Zin;
};

__device__ __forceinline__ float chainer_sigmoid(float v) {
    // Chainer's forward: tanh(x * 0.5) * 0.5 + 0.5
    return __fadd_rn(__fmul_rn(tanhf(__fmul_rn(v, 0.5f)), 0.5f), 0.5f);
}

// One ConvLSTM cell update (net.py:94-126).  Returns h', updates c in place.
__device__ __forceinline__ float lstm_cell(float gi, float gf, float gc, float go, const float* bias4,
                                           const float* peep4, float* cptr) {
    const float c_old = *cptr;
    const float i = chainer_sigmoid(__fadd_rn(__fadd_rn(gi, bias4[0]), __fmul_rn(c_old, peep4[0])));
    const float f = chainer_sigmoid(__fadd_rn(__fadd_rn(gf, bias4[1]), __fmul_rn(c_old, peep4[1])));
    const float cn = __fadd_rn(__fmul_rn(tanhf(__fadd_rn(gc, bias4[2])), i), __fmul_rn(f, c_old));
    const float o = chainer_sigmoid(__fadd_rn(__fadd_rn(go, bias4[3]), __fmul_rn(c_old, peep4[2])));
    *cptr = cn;
    return __fmul_rn(o, tanhf(cn));
}

// REV = true walks the nine taps of every input channel backwards: same products, another fp32 summation order - the
// yardstick for "what a different but equally exact fp32 implementation does to the frames"
// (profiles/experiments/pass_ablation.py; eig_set_option "simt_reverse_taps").
template <int TN, bool REV = false>
__global__ void __launch_bounds__(256) conv3x3_simt_kernel(ConvArgs a) {
    constexpr int CK = 8, TW = 16, TH = 8, SROW = 20;  // SROW: padded smem row (bank-conflict free half-warps)
    EIG_DYN_SMEM(smem);
    const int nw = blockDim.x >> 5;
    const int ncta = nw * TN;  // output channels per CTA
    float* sIn = reinterpret_cast<float*>(smem);               // [CK][TH+2][SROW]
    float* sW = sIn + CK * (TH + 2) * SROW;                     // [9][CK][ncta]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (a.W + TW - 1) / TW;
    const int tile = blockIdx.x;
    const int x0 = (tile % tiles_x) * TW, y0 = (tile / tiles_x) * TH;
    const int b = blockIdx.y;
    const int n_cta0 = blockIdx.z * ncta;
    const int tx = lane & 15, ty = lane >> 4;
    const int n0 = n_cta0 + warp * TN;

    float acc[4][TN];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[j][n] = 0.f;

    const long long img_base = (long long)b * a.H * a.W;
    for (int c0 = 0; c0 < a.Cin; c0 += CK) {
        // stage the input halo tile: (TH+2) x (TW+2) pixels x CK channels
        for (int i = threadIdx.x; i < (TH + 2) * (TW + 2) * CK; i += blockDim.x) {
            const int ck = i % CK;
            const int pix = i / CK;
            const int cx = pix % (TW + 2), cy = pix / (TW + 2);
            const int gx = x0 + cx - 1, gy = y0 + cy - 1;
            float v = 0.f;
            if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H && c0 + ck < a.Cin) {
                const long long idx = (img_base + (long long)gy * a.W + gx) * a.in_pitch + a.in_coff + c0 + ck;
                v = view_load(a.in_hi, a.in_lo, idx);
            }
            sIn[(ck * (TH + 2) + cy) * SROW + cx] = v;
        }
        // stage the weight slab [9][CK][ncta]
        for (int i = threadIdx.x; i < 9 * CK * ncta; i += blockDim.x) {
            const int n = i % ncta;
            const int r = i / ncta;
            const int ck = r % CK, tap = r / CK;
            float v = 0.f;
            if (c0 + ck < a.Cin && n_cta0 + n < a.Npad)
                v = a.wgt[((long long)tap * a.Cin + c0 + ck) * a.Npad + n_cta0 + n];
            sW[i] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ck = 0; ck < CK; ++ck) {
            float in[6][3];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) in[r][c] = sIn[(ck * (TH + 2) + ty * 4 + r) * SROW + tx + c];
#pragma unroll
            for (int ky0 = 0; ky0 < 3; ++ky0)
#pragma unroll
                for (int kx0 = 0; kx0 < 3; ++kx0) {
                    const int ky = REV ? 2 - ky0 : ky0, kx = REV ? 2 - kx0 : kx0;
                    const float* wrow = sW + ((ky * 3 + kx) * CK + ck) * ncta + warp * TN;
                    float wv[TN];
#pragma unroll
                    for (int n = 0; n < TN; n += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(wrow + n);
                        wv[n] = q.x; wv[n + 1] = q.y; wv[n + 2] = q.z; wv[n + 3] = q.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int n = 0; n < TN; ++n) acc[j][n] = __fmaf_rn(in[j + ky][kx], wv[n], acc[j][n]);
                }
        }
        __syncthreads();
    }

    const int gx = x0 + tx;
    if (a.epi == EPI_CONVP) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gy = y0 + ty * 4 + j;
            if (gx >= a.W || gy >= a.H) continue;
            const long long pix = img_base + (long long)gy * a.W + gx;
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                if (n0 + n >= a.N) continue;
                const int nP = a.nP ? a.nP : a.N;
                if (n0 + n >= nP) { a.outZ[pix * (a.N - nP) + n0 + n - nP] = acc[j][n]; continue; }
                float v = __fadd_rn(acc[j][n], a.bias[n0 + n]);
                v = v > 0.f ? v : 0.f;
                if (a.clip && v > 1.f) v = 1.f;
                a.outP[pix * nP + n0 + n] = v;
            }
        }
    } else if (a.epi == EPI_CONVA) {
        const int Hp = a.H >> 1, Wp = a.W >> 1;
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
            const int gy = y0 + ty * 4 + jp * 2;
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                const float bn = (n0 + n < a.N) ? a.bias[n0 + n] : 0.f;
                float v0 = __fadd_rn(acc[jp * 2][n], bn), v1 = __fadd_rn(acc[jp * 2 + 1][n], bn);
                float m = fmaxf(fmaxf(v0, v1), 0.f);  // relu commutes with max
                const float mo = __shfl_xor_sync(0xffffffffu, m, 1);
                m = fmaxf(m, mo);
                if ((tx & 1) == 0 && gx < a.W && gy < a.H && n0 + n < a.N) {
                    const long long pp = ((long long)b * Hp + (gy >> 1)) * Wp + (gx >> 1);
                    const float pv = a.P[pp * a.N + n0 + n];
                    const float ep = __fsub_rn(m, pv), en = __fsub_rn(pv, m);
                    view_store(a.dstE, pp, n0 + n, ep > 0.f ? ep : 0.f);
                    view_store(a.dstE, pp, a.N + n0 + n, en > 0.f ? en : 0.f);
                }
            }
        }
    } else {  // EPI_LSTM
        const int R = a.N >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gy = y0 + ty * 4 + j;
            if (gx >= a.W || gy >= a.H) continue;
            const long long pix = img_base + (long long)gy * a.W + gx;
            const long long ppix = (long long)gy * a.W + gx;
#pragma unroll
            for (int q = 0; q < TN / 4; ++q) {
                const int n = n0 + q * 4;
                if (n >= a.N) continue;
                const int r = n >> 2;
                const float hnew = lstm_cell(acc[j][q * 4], acc[j][q * 4 + 1], acc[j][q * 4 + 2], acc[j][q * 4 + 3],
                                             a.bias + n, a.peep + (ppix * R + r) * 4, a.cstate + pix * R + r);
                view_store(a.dstH, pix, r, hnew);
                if (a.dstUp.hi) {
                    const int W2 = a.W * 2;
                    const long long ub = ((long long)b * a.H * 2 + gy * 2) * W2 + gx * 2;
                    view_store(a.dstUp, ub, r, hnew);
                    view_store(a.dstUp, ub + 1, r, hnew);
                    view_store(a.dstUp, ub + W2, r, hnew);
                    view_store(a.dstUp, ub + W2 + 1, r, hnew);
                }
            }
        }
    }
}

// write_image (call_prednet.py:51-61): u8 = trunc(P0 * 255) in fp32; then cv2 gray (optical_flow.py:62-65):
// (B*3735 + G*19235 + R*9798 + 16384) >> 15 on the RGB triple, identity for one channel.
__global__ void __launch_bounds__(256) quantize_gray_kernel(const float* P0, unsigned char* img, unsigned char* gray,
                                                            long long npix, int C0) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    int ch[3];
    for (int c = 0; c < C0; ++c) {
        const float v = __fmul_rn(P0[p * C0 + c], 255.0f);
        ch[c] = (int)v & 0xff;  // values are in [0,255]: astype(uint8) truncates
        if (img) img[p * C0 + c] = (unsigned char)ch[c];
    }
    if (gray) gray[p] = (C0 == 3) ? (unsigned char)((ch[2] * 3735 + ch[1] * 19235 + ch[0] * 9798 + 16384) >> 15)
                                  : (unsigned char)ch[0];
}

// reset_state (net.py:159-164, called per genome in call_prednet.py:203): every state region that step 0 reads before it
// is written, zeroed by one launch (blockIdx.y = region) instead of one memset node per region.
#define RESET_MAX_REGIONS 16
struct ResetArgs {
    void* ptr[RESET_MAX_REGIONS];
    unsigned long long bytes[RESET_MAX_REGIONS];
};
__global__ void __launch_bounds__(256) reset_state_kernel(ResetArgs a) {
    unsigned char* p = static_cast<unsigned char*>(a.ptr[blockIdx.y]);
    const unsigned long long n = a.bytes[blockIdx.y];
    unsigned long long head = (16 - (reinterpret_cast<unsigned long long>(p) & 15)) & 15;
    if (head > n) head = n;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = tid; i < head; i += stride) p[i] = 0;
    uint4* q = reinterpret_cast<uint4*>(p + head);
    const unsigned long long nq = (n - head) >> 4;
    uint4 z; z.x = 0; z.y = 0; z.z = 0; z.w = 0;
    for (unsigned long long i = tid; i < nq; i += stride) q[i] = z;
    for (unsigned long long i = head + (nq << 4) + tid; i < n; i += stride) p[i] = 0;
}

// gray from an interleaved u8 image (used for the rendered input image in the single-image pairing)
__global__ void __launch_bounds__(256) gray_u8_kernel(const unsigned char* img, unsigned char* gray, long long npix, int C0) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    if (C0 == 3) {
        const int r = img[p * 3], g = img[p * 3 + 1], b = img[p * 3 + 2];
        gray[p] = (unsigned char)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
    } else {
        gray[p] = img[p];
    }
}

}  // namespace eig
