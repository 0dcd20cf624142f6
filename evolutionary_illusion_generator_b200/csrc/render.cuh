// K1 cppn_render: evaluate every genome's CPPN at every pixel in one launch.
//
// Replaces, per genome, `get_image_from_cppn` (/root/reference/generate_illusion.py:372-460) and the
// `Node.__call__` recursion (/root/reference/pytorch_neat/pytorch_neat/cppn.py:75-108).  The genome is a
// flat program (see genome.py:flatten_genome) staged in shared memory; node values of the block's pixels
// live in shared memory as [slot][thread] fp64; all pixel-dependent arithmetic is fp64 with separately
// rounded multiply and add (torch evaluates `w*x`, `sum`, `response*pre+bias` as separate ops).
// Epilogue fuses the background override, `*255` and numpy's float->uint8 cast (truncate, wrap mod 256),
// and also emits the PredNet input x = float32(float64(u8)/255) (call_prednet.py:29-49).
#pragma once
#include "common.cuh"

namespace eig {

struct RenderArgs {
    const unsigned char* blob;   // concatenated genome programs
    const long long* offsets;    // [n+1] byte offsets into blob
    const double* xmat;          // [npix]
    const double* ymat;          // [npix]
    int npix;
    int c_dim;                   // 1 or 3
    int mode;                    // 0: gradient (gray or colour), 1: gray + np.round, 2: colour palette
    double bg;                   // 1 = white, 0 = black
    unsigned char* img;          // [n][npix][c_dim]
    float* x;                    // [n][npix][c_dim]  (may be null)
    int max_blob_bytes;          // smem reserved for the program (multiple of 16)
};

struct NodeRec { int act, agg, t0, nt; double bias, resp; };
struct TermRec { double w; int src, pad; };

__device__ __forceinline__ double cppn_act(int act, double v) {
    switch (act) {
        case 0: return __ddiv_rn(1.0, __dadd_rn(1.0, exp(__dsub_rn(0.0, __dmul_rn(5.0, v)))));  // torch.sigmoid(5x)
        case 1: return tanh(__dmul_rn(2.5, v));
        case 2: return fabs(v);
        case 3: return exp(__dmul_rn(-5.0, __dmul_rn(v, v)));
        case 4: return v;
        case 5: return sin(v);
        default: return v > 0.0 ? v : (v != v ? v : 0.0);  // relu, NaN passes through
    }
}

// numpy's float64 -> uint8 cast on x86-64: truncate to int32 (cvttsd2si), keep the low byte;
// NaN / out-of-int32 give the "integer indefinite" 0x80000000 -> 0.
__device__ __forceinline__ unsigned char numpy_u8(double v) {
    if (!(fabs(v) < 2147483648.0)) return 0;
    int i = __double2int_rz(v);
    return (unsigned char)(i & 0xff);
}

__global__ void __launch_bounds__(128) cppn_render_kernel(RenderArgs a) {
    EIG_DYN_SMEM(smem);
    const int g = blockIdx.y;
    const int nt = blockDim.x;
    const int t = threadIdx.x;
    const long long o0 = a.offsets[g], o1 = a.offsets[g + 1];
    const int nwords = (int)((o1 - o0) >> 3);
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.blob + o0);
    for (int i = t; i < nwords; i += nt) prog[i] = src[i];
    __syncthreads();
    const int* hdr = reinterpret_cast<const int*>(prog);
    const int n_nodes = hdr[1], n_out = hdr[3];
    const int* outs = hdr + 4;
    const int outs_words = (n_out + 1) >> 1;
    const NodeRec* nodes = reinterpret_cast<const NodeRec*>(prog + 2 + outs_words);
    const TermRec* terms = reinterpret_cast<const TermRec*>(nodes + n_nodes);
    double* vals = reinterpret_cast<double*>(smem + a.max_blob_bytes);

    const int p = blockIdx.x * nt + t;
    const bool live = p < a.npix;
    const double xv = live ? a.xmat[p] : 0.0;
    const double yv = live ? a.ymat[p] : 0.0;
    vals[0 * nt + t] = xv;
    vals[1 * nt + t] = yv;
    vals[2 * nt + t] = 1.0;
    for (int i = 0; i < n_nodes; ++i) {
        const NodeRec nd = nodes[i];
        double acc = 0.0;
        for (int k = 0; k < nd.nt; ++k) {
            const TermRec tr = terms[nd.t0 + k];
            const double term = __dmul_rn(tr.w, vals[tr.src * nt + t]);
            acc = (k == 0) ? term : (nd.agg == 0 ? __dadd_rn(acc, term) : __dmul_rn(acc, term));
        }
        vals[(3 + i) * nt + t] = cppn_act(nd.act, __dadd_rn(__dmul_rn(nd.resp, acc), nd.bias));
    }
    if (!live) return;
    const bool is_bg = (xv == -1.0);
    const long long base = ((long long)g * a.npix + p) * a.c_dim;
    unsigned char px[3];
    if (a.mode == 2) {
        const int o = outs[0];
        double v = vals[(o & 0x3fffffff) * nt + t];
        unsigned char idx = numpy_u8(__dmul_rn(v, 4.0));
        px[0] = (idx == 0 || idx == 1) ? 255 : 0;
        px[1] = (idx == 0 || idx == 2) ? 255 : 0;
        px[2] = (idx == 0 || idx == 3) ? 255 : 0;
        if (is_bg) px[0] = px[1] = px[2] = numpy_u8(__dmul_rn(a.bg, 255.0));
    } else {
        for (int c = 0; c < a.c_dim; ++c) {
            if (c >= n_out) { px[c] = 0; continue; }   // np.zeros image_array for missing outputs
            const int o = outs[c];
            double v = vals[(o & 0x3fffffff) * nt + t];
            const bool f32_const = (o >> 30) & 1;      // constant plane: the reference keeps it float32
            if (a.c_dim == 1 && f32_const) {
                float f = (float)v;
                if (is_bg) f = (float)a.bg;
                if (a.mode == 1) f = rintf(f);
                float s = __fmul_rn(f, 255.0f);
                // numpy float32 -> uint8 takes the same truncating path
                px[c] = numpy_u8((double)s);
            } else {
                if (is_bg) v = a.bg;
                if (a.mode == 1) v = rint(v);
                px[c] = numpy_u8(__dmul_rn(v, 255.0));
            }
        }
    }
    for (int c = 0; c < a.c_dim; ++c) {
        a.img[base + c] = px[c];
        if (a.x) a.x[base + c] = (float)__ddiv_rn((double)px[c], 255.0);
    }
}

}  // namespace eig
