// Sparse optical flow on the GPU: Shi-Tomasi corners + pyramidal Lucas-Kanade, following OpenCV's
// arithmetic so that results match the cv2 calls of /root/reference/optical_flow/optical_flow.py:51-82
// (`goodFeaturesToTrack(maxCorners=100, qualityLevel=0.3, minDistance=7, blockSize=7)`,
//  `calcOpticalFlowPyrLK(winSize=(50,50), maxLevel=2, criteria=(EPS|COUNT, 10, 0.03))`).
// The specification is oracle/flow.py (numpy restatement pinned against the cv2 4.13 binary).
//   min_eig_kernel     : Sobel 3x3 (fp32, OpenCV's FMA order) -> products -> 7x7 box sums in fp64 -> min eigen
//                        value map + per-image max (ordered-int atomicMax)
//   corner_select_kernel: threshold 0.3*max, 3x3 non-max suppression, bitonic sort (value desc, address desc),
//                        greedy 7-px spacing, at most 100 corners           (one CTA per image)
//   pyr_down_kernel    : 5x5 [1 4 6 4 1] integer pyramid level               (both frames)
//   scharr_kernel      : int16 Scharr derivatives of frame 1
//   lk_track_kernel    : one 128-thread CTA per corner; the 50x50 window sums (A11,A12,A22,b1,b2) are exact int64
//                        sums reduced with warp shuffles (+ one shared-memory exchange); <= 10 Newton steps per level
//   collect_vectors_kernel: rows [x, y, dx, dy] of the tracked corners, in corner order
#pragma once
#include "common.cuh"

namespace eig {

enum { FLOW_MAX_CORNERS = 100, FLOW_WIN = 50, FLOW_BLOCK = 7, FLOW_MAX_LEVELS = 3, FLOW_MAX_ITERS = 10 };

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}
__device__ __forceinline__ int float_order_key(float f) {
    // monotonic float -> int map so that atomicMax(int) orders floats (negative values included)
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float float_from_order_key(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// ---------------------------------------------------------------------------------------------- min eigen map
struct EigArgs {
    const unsigned char* gray;  // [B][H][W]
    float* eig;                 // [B][H][W]
    int* eig_max_key;           // [B], initialised to INT_MIN
    int H, W;
};

__device__ __forceinline__ void sobel_products(const unsigned char* g, int H, int W, int y, int x, float s, float s2,
                                               float* out3) {
    // derivative at in-image pixel (y,x); neighbours use BORDER_REFLECT_101
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const float a00 = g[ym * W + xm], a01 = g[ym * W + x], a02 = g[ym * W + xp];
    const float a10 = g[y * W + xm], a11 = g[y * W + x], a12 = g[y * W + xp];
    const float a20 = g[yp * W + xm], a21 = g[yp * W + x], a22 = g[yp * W + xp];
    // dx: row difference, then column smoothing [s, 2s, s] as fma(up+down, s, mid*2s)
    const float rdU = a02 - a00, rdM = a12 - a10, rdD = a22 - a20;
    const float dx = __fmaf_rn(__fadd_rn(rdD, rdU), s, __fmul_rn(rdM, s2));
    // dy: row smoothing [s, 2s, s] (FMA chain in OpenCV's 32-pixel SIMD body, plain in its scalar tail), then
    // column difference
    float rU, rD;
    if (x < (W & ~31)) {
        rU = __fmaf_rn(a02, s, __fmaf_rn(a01, s2, __fmul_rn(a00, s)));
        rD = __fmaf_rn(a22, s, __fmaf_rn(a21, s2, __fmul_rn(a20, s)));
    } else {
        rU = __fadd_rn(__fadd_rn(__fmul_rn(a00, s), __fmul_rn(a01, s2)), __fmul_rn(a02, s));
        rD = __fadd_rn(__fadd_rn(__fmul_rn(a20, s), __fmul_rn(a21, s2)), __fmul_rn(a22, s));
    }
    const float dy = __fsub_rn(rD, rU);
    (void)a11;
    out3[0] = __fmul_rn(dx, dx);
    out3[1] = __fmul_rn(dx, dy);
    out3[2] = __fmul_rn(dy, dy);
}

__global__ void __launch_bounds__(256) min_eig_kernel(EigArgs a) {
    constexpr int TW = 32, TH = 16, R = FLOW_BLOCK / 2, CW = TW + 2 * R, CH = TH + 2 * R;
    __shared__ float sCov[CH][CW][3];
    __shared__ double sRow[CH][TW][3];
    __shared__ int sMax;
    const int b = blockIdx.y;
    const int tiles_x = (a.W + TW - 1) / TW;
    const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
    const unsigned char* g = a.gray + (long long)b * a.H * a.W;
    const double scale = 1.0 / (4 * FLOW_BLOCK * 255.0);
    const float s = (float)scale, s2 = (float)(2.0 * scale);
    if (threadIdx.x == 0) sMax = (int)0x80000000;
    for (int i = threadIdx.x; i < CH * CW; i += blockDim.x) {
        const int cy = i / CW, cx = i % CW;
        // the box filter reflects the covariance maps (BORDER_REFLECT_101), i.e. it re-reads in-image pixels
        const int y = reflect101(y0 + cy - R, a.H), x = reflect101(x0 + cx - R, a.W);
        float o[3] = {0.f, 0.f, 0.f};
        if (y0 + cy - R < a.H + R && x0 + cx - R < a.W + R && y >= 0 && y < a.H && x >= 0 && x < a.W)
            sobel_products(g, a.H, a.W, y, x, s, s2, o);
        sCov[cy][cx][0] = o[0]; sCov[cy][cx][1] = o[1]; sCov[cy][cx][2] = o[2];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CH * TW; i += blockDim.x) {
        const int cy = i / TW, cx = i % TW;
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int k = 0; k < FLOW_BLOCK; ++k) {
            r0 = __dadd_rn(r0, (double)sCov[cy][cx + k][0]);
            r1 = __dadd_rn(r1, (double)sCov[cy][cx + k][1]);
            r2 = __dadd_rn(r2, (double)sCov[cy][cx + k][2]);
        }
        sRow[cy][cx][0] = r0; sRow[cy][cx][1] = r1; sRow[cy][cx][2] = r2;
    }
    __syncthreads();
    int best = (int)0x80000000;
    for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
        const int ty = i / TW, tx = i % TW;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= a.H || x >= a.W) continue;
        double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
        for (int k = 0; k < FLOW_BLOCK; ++k) {
            c0 = __dadd_rn(c0, sRow[ty + k][tx][0]);
            c1 = __dadd_rn(c1, sRow[ty + k][tx][1]);
            c2 = __dadd_rn(c2, sRow[ty + k][tx][2]);
        }
        const float A = __fmul_rn((float)c0, 0.5f), Bv = (float)c1, C = __fmul_rn((float)c2, 0.5f);
        const float d = __fsub_rn(A, C);
        const float e = __fsub_rn(__fadd_rn(A, C), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(Bv, Bv))));
        a.eig[((long long)b * a.H + y) * a.W + x] = e;
        const int key = float_order_key(e);
        best = key > best ? key : best;
    }
    atomicMax(&sMax, best);
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(a.eig_max_key + b, sMax);
}

// ---------------------------------------------------------------------------------------------- corner selection
struct CornerArgs {
    const float* eig;              // [B][H][W]
    const int* eig_max_key;        // [B]
    unsigned long long* cand;      // [B][H*W] scratch
    float* corners;                // [B][100][2]
    int* ncorners;                 // [B]
    int H, W;
};

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = keys[i], y = keys[ixj];
                    const bool up = (i & k) == 0;  // descending overall
                    if (up ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024) corner_select_kernel(CornerArgs a) {
    constexpr int SMEM_KEYS = 2048;
    __shared__ unsigned long long sKeys[SMEM_KEYS];
    __shared__ int sCount;
    const int b = blockIdx.x;
    const float* eig = a.eig + (long long)b * a.H * a.W;
    unsigned long long* cand = a.cand + (long long)b * a.H * a.W;
    const float maxv = float_from_order_key(a.eig_max_key[b]);
    const float thr = (float)((double)maxv * 0.3);
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    const int iw = a.W - 2, ih = a.H - 2;
    for (int i = threadIdx.x; i < iw * ih; i += blockDim.x) {
        const int y = 1 + i / iw, x = 1 + i % iw;
        const float v = eig[y * a.W + x];
        if (!(v > thr) || v == 0.f) continue;
        bool is_max = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const float nv = eig[(y + dy) * a.W + x + dx];
                // dilate works on the thresholded map: neighbours <= thr count as 0 (< v)
                if (nv > thr && nv > v) is_max = false;
            }
        if (!is_max) continue;
        const int slot = atomicAdd(&sCount, 1);
        const unsigned key = (unsigned)float_order_key(v) ^ 0x80000000u;  // unsigned-monotonic
        cand[slot] = ((unsigned long long)key << 32) | (unsigned)(y * a.W + x);
    }
    __syncthreads();
    const int count = sCount;
    int n_pow2 = 1;
    while (n_pow2 < count) n_pow2 <<= 1;
    unsigned long long* keys;
    if (n_pow2 <= SMEM_KEYS) {
        keys = sKeys;
        for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) keys[i] = i < count ? cand[i] : 0ull;
    } else {
        keys = cand;  // H*W is not necessarily a power of two: pad virtually by clamping the sort size
        // fall back: sort the first power-of-two prefix that fits in the scratch; H*W >= count always, and
        // n_pow2 can exceed H*W only when count > H*W/2, which a 3x3 non-max suppression cannot produce
        // without plateaus.  Clamp defensively.
        while (n_pow2 > a.H * a.W) n_pow2 >>= 1;
        for (int i = count + threadIdx.x; i < n_pow2; i += blockDim.x) keys[i] = 0ull;
    }
    __syncthreads();
    if (count > 1) bitonic_sort_desc(keys, n_pow2);
    __syncthreads();
    const int total = count < n_pow2 ? count : n_pow2;
    // Greedy spacing (accept a candidate when no accepted corner lies within squared distance 49), serial over the sorted
    // list.  Warp 0 takes 32 candidates at a time: a shared bitmap holds every pixel closer than 7 px to an accepted
    // corner, so "is this candidate blocked" is one bit test; inside a batch the surviving lanes are accepted in order,
    // each acceptance marking its disc before the later lanes look again.
    EIG_DYN_SMEM(smem_blocked);
    unsigned* blocked = reinterpret_cast<unsigned*>(smem_blocked);
    const int n_words = (a.H * a.W + 31) >> 5;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) blocked[i] = 0u;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int nacc = 0;
        for (int base = 0; base < total && nacc < FLOW_MAX_CORNERS; base += 32) {
            const int i = base + lane;
            bool live = i < total;
            unsigned addr = 0;
            if (live) {
                addr = (unsigned)(keys[i] & 0xffffffffu);
                live = !((blocked[addr >> 5] >> (addr & 31)) & 1u);
            }
            unsigned m = __ballot_sync(0xffffffffu, live);
            while (m && nacc < FLOW_MAX_CORNERS) {
                const int l = __ffs((int)m) - 1;
                const unsigned acc_addr = __shfl_sync(0xffffffffu, addr, l);
                const int ay = (int)(acc_addr / (unsigned)a.W), ax = (int)(acc_addr % (unsigned)a.W);
                if (lane == 0) {
                    a.corners[((long long)b * FLOW_MAX_CORNERS + nacc) * 2] = (float)ax;
                    a.corners[((long long)b * FLOW_MAX_CORNERS + nacc) * 2 + 1] = (float)ay;
                }
                ++nacc;
                for (int t = lane; t < 13 * 13; t += 32) {        // dx, dy in [-6, 6]: dx*dx + dy*dy < 49
                    const int dy = t / 13 - 6, dx = t % 13 - 6;
                    const int y = ay + dy, x = ax + dx;
                    if (dx * dx + dy * dy < 49 && y >= 0 && y < a.H && x >= 0 && x < a.W) {
                        const unsigned p = (unsigned)(y * a.W + x);
                        atomicOr(&blocked[p >> 5], 1u << (p & 31));
                    }
                }
                __syncwarp();
                if (lane <= l) live = false;
                else if (live) live = !((blocked[addr >> 5] >> (addr & 31)) & 1u);
                m = __ballot_sync(0xffffffffu, live);
            }
        }
        if (lane == 0) a.ncorners[b] = nacc;
    }
}

// ---------------------------------------------------------------------------------------------- pyramid + Scharr
// dst level (oh, ow) from src level (h, w); images laid out [n_img][h][w]
__global__ void __launch_bounds__(256) pyr_down_kernel(const unsigned char* src, unsigned char* dst, int h, int w,
                                                       int oh, int ow, int n_img) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_img * oh * ow) return;
    const int x = (int)(i % ow), y = (int)((i / ow) % oh), img = (int)(i / ((long long)ow * oh));
    const unsigned char* s = src + (long long)img * h * w;
    int acc = 0;
    const int wt[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
        const int yy = reflect101(2 * y + dy - 2, h);
        int row = 0;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) row += wt[dx] * s[yy * w + reflect101(2 * x + dx - 2, w)];
        acc += wt[dy] * row;
    }
    dst[i] = (unsigned char)((acc + 128) >> 8);
}

// calcScharrDeriv: (dx, dy) int16 interleaved, reflect101 at the image edge
__global__ void __launch_bounds__(256) scharr_kernel(const unsigned char* src, short* deriv, int h, int w, int n_img) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_img * h * w) return;
    const int x = (int)(i % w), y = (int)((i / w) % h), img = (int)(i / ((long long)w * h));
    const unsigned char* s = src + (long long)img * h * w;
    const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const int t0m = (s[ym * w + xm] + s[yp * w + xm]) * 3 + s[y * w + xm] * 10;
    const int t0p = (s[ym * w + xp] + s[yp * w + xp]) * 3 + s[y * w + xp] * 10;
    const int t1m = s[yp * w + xm] - s[ym * w + xm];
    const int t1c = s[yp * w + x] - s[ym * w + x];
    const int t1p = s[yp * w + xp] - s[ym * w + xp];
    deriv[i * 2] = (short)(t0p - t0m);
    deriv[i * 2 + 1] = (short)((t1p + t1m) * 3 + t1c * 10);
}

// ---------------------------------------------------------------------------------------------- LK tracking
struct LkArgs {
    const unsigned char* img1[FLOW_MAX_LEVELS];  // per level [B][h][w]
    const unsigned char* img2[FLOW_MAX_LEVELS];
    const short* deriv1[FLOW_MAX_LEVELS];        // per level [B][h][w][2]
    int lh[FLOW_MAX_LEVELS], lw[FLOW_MAX_LEVELS];
    int n_levels;
    const float* corners;   // [B][100][2]
    const int* ncorners;    // [B]
    float* next_pts;        // [B][100][2]
    unsigned char* status;  // [B][100]
    int B;
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int* w) {
    const float sc = 16384.f;
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    w[0] = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), sc));
    w[1] = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), sc));
    w[2] = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), sc));
    w[3] = 16384 - w[0] - w[1] - w[2];
}

// One CTA of LK_THREADS threads per corner: the 50x50 window is spread over the CTA (exact int64 partial sums, so the
// result does not depend on the partition), reduced with warp shuffles and one shared-memory exchange per sum.
#define LK_THREADS 128
__device__ __forceinline__ void lk_block_sum3(long long& a, long long& b, long long& c, long long (*sRed)[3]) {
    a = warp_sum_ll(a); b = warp_sum_ll(b); c = warp_sum_ll(c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                       // the previous use of sRed is over
    if (lane == 0) { sRed[warp][0] = a; sRed[warp][1] = b; sRed[warp][2] = c; }
    __syncthreads();
    a = 0; b = 0; c = 0;
#pragma unroll
    for (int w = 0; w < LK_THREADS / 32; ++w) { a += sRed[w][0]; b += sRed[w][1]; c += sRed[w][2]; }
}

__global__ void __launch_bounds__(LK_THREADS) lk_track_kernel(LkArgs a) {
    constexpr int WIN = FLOW_WIN, NPX = WIN * WIN;
    __shared__ short I[NPX];
    __shared__ short Ix[NPX];
    __shared__ short Iy[NPX];
    __shared__ long long sRed[LK_THREADS / 32][3];
    const int tid = threadIdx.x;
    const int pt = blockIdx.x;
    const int b = pt / FLOW_MAX_CORNERS, k = pt % FLOW_MAX_CORNERS;
    if (b >= a.B) return;
    if (k >= a.ncorners[b]) return;  // block-uniform
    const float px0 = a.corners[((long long)b * FLOW_MAX_CORNERS + k) * 2];
    const float py0 = a.corners[((long long)b * FLOW_MAX_CORNERS + k) * 2 + 1];
    const float half = 24.5f;
    const float flt_scale = 1.f / (1 << 20);
    float nx = 0.f, ny = 0.f;
    bool status = true;
    for (int level = a.n_levels - 1; level >= 0; --level) {
        const int rows = a.lh[level], cols = a.lw[level];
        const unsigned char* im1 = a.img1[level] + (long long)b * rows * cols;
        const unsigned char* im2 = a.img2[level] + (long long)b * rows * cols;
        const short* dv = a.deriv1[level] + (long long)b * rows * cols * 2;
        const float inv = 1.f / (float)(1 << level);
        const float prx = __fmul_rn(px0, inv), pry = __fmul_rn(py0, inv);
        if (level == a.n_levels - 1) { nx = prx; ny = pry; }
        else { nx = __fmul_rn(nx, 2.f); ny = __fmul_rn(ny, 2.f); }
        const float pxw = __fsub_rn(prx, half), pyw = __fsub_rn(pry, half);
        const int ipx = __float2int_rd(pxw), ipy = __float2int_rd(pyw);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) status = false;
            continue;
        }
        int w[4];
        lk_weights(__fsub_rn(pxw, (float)ipx), __fsub_rn(pyw, (float)ipy), w);
        long long s11 = 0, s12 = 0, s22 = 0;
        __syncthreads();   // the window buffers of the previous level are no longer read
        for (int i = tid; i < NPX; i += LK_THREADS) {
            const int wy = i / WIN, wx = i % WIN;
            const int gy = ipy + wy, gx = ipx + wx;
            // image: 50-px BORDER_REFLECT_101 frame around the level; derivatives: zero outside
            const int y0r = reflect101(gy, rows), y1r = reflect101(gy + 1, rows);
            const int x0r = reflect101(gx, cols), x1r = reflect101(gx + 1, cols);
            const int iv = (im1[y0r * cols + x0r] * w[0] + im1[y0r * cols + x1r] * w[1] + im1[y1r * cols + x0r] * w[2] +
                            im1[y1r * cols + x1r] * w[3] + (1 << 8)) >> 9;
            int dx00 = 0, dy00 = 0, dx01 = 0, dy01 = 0, dx10 = 0, dy10 = 0, dx11 = 0, dy11 = 0;
            const bool yin0 = gy >= 0 && gy < rows, yin1 = gy + 1 >= 0 && gy + 1 < rows;
            const bool xin0 = gx >= 0 && gx < cols, xin1 = gx + 1 >= 0 && gx + 1 < cols;
            if (yin0 && xin0) { dx00 = dv[(gy * cols + gx) * 2]; dy00 = dv[(gy * cols + gx) * 2 + 1]; }
            if (yin0 && xin1) { dx01 = dv[(gy * cols + gx + 1) * 2]; dy01 = dv[(gy * cols + gx + 1) * 2 + 1]; }
            if (yin1 && xin0) { dx10 = dv[((gy + 1) * cols + gx) * 2]; dy10 = dv[((gy + 1) * cols + gx) * 2 + 1]; }
            if (yin1 && xin1) { dx11 = dv[((gy + 1) * cols + gx + 1) * 2]; dy11 = dv[((gy + 1) * cols + gx + 1) * 2 + 1]; }
            const int ixv = (dx00 * w[0] + dx01 * w[1] + dx10 * w[2] + dx11 * w[3] + (1 << 13)) >> 14;
            const int iyv = (dy00 * w[0] + dy01 * w[1] + dy10 * w[2] + dy11 * w[3] + (1 << 13)) >> 14;
            I[i] = (short)iv; Ix[i] = (short)ixv; Iy[i] = (short)iyv;
            s11 += (long long)ixv * ixv; s12 += (long long)ixv * iyv; s22 += (long long)iyv * iyv;
        }
        lk_block_sum3(s11, s12, s22, sRed);   // also makes the window visible to the whole CTA
        const float A11 = __fmul_rn(__ll2float_rn(s11), flt_scale);
        const float A12 = __fmul_rn(__ll2float_rn(s12), flt_scale);
        const float A22 = __fmul_rn(__ll2float_rn(s22), flt_scale);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float min_eig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * WIN * WIN));
        if (min_eig < 1e-4f || D < 1.1920928955078125e-07f) {
            if (level == 0) status = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        float cx = __fsub_rn(nx, half), cy = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < FLOW_MAX_ITERS; ++j) {
            const int icx = __float2int_rd(cx), icy = __float2int_rd(cy);
            if (icx < -WIN || icx >= cols || icy < -WIN || icy >= rows) {
                if (level == 0) status = false;
                break;
            }
            lk_weights(__fsub_rn(cx, (float)icx), __fsub_rn(cy, (float)icy), w);
            long long sb1 = 0, sb2 = 0, unused = 0;
            for (int i = tid; i < NPX; i += LK_THREADS) {
                const int wy = i / WIN, wx = i % WIN;
                const int y0r = reflect101(icy + wy, rows), y1r = reflect101(icy + wy + 1, rows);
                const int x0r = reflect101(icx + wx, cols), x1r = reflect101(icx + wx + 1, cols);
                const int jv = (im2[y0r * cols + x0r] * w[0] + im2[y0r * cols + x1r] * w[1] +
                                im2[y1r * cols + x0r] * w[2] + im2[y1r * cols + x1r] * w[3] + (1 << 8)) >> 9;
                const int diff = jv - I[i];
                sb1 += (long long)diff * Ix[i];
                sb2 += (long long)diff * Iy[i];
            }
            lk_block_sum3(sb1, sb2, unused, sRed);
            const float b1 = __fmul_rn(__ll2float_rn(sb1), flt_scale);
            const float b2 = __fmul_rn(__ll2float_rn(sb2), flt_scale);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            cx = __fadd_rn(cx, ddx); cy = __fadd_rn(cy, ddy);
            nx = __fadd_rn(cx, half); ny = __fadd_rn(cy, half);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= 0.03 * 0.03) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(ddx, 0.5f));
                ny = __fsub_rn(ny, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        if (level == 0 && status) {
            const int ifx = __float2int_rd(__fsub_rn(nx, half)), ify = __float2int_rd(__fsub_rn(ny, half));
            if (ifx < -WIN || ifx >= cols || ify < -WIN || ify >= rows) status = false;
        }
    }
    if (tid == 0) {
        a.next_pts[((long long)b * FLOW_MAX_CORNERS + k) * 2] = nx;
        a.next_pts[((long long)b * FLOW_MAX_CORNERS + k) * 2 + 1] = ny;
        a.status[(long long)b * FLOW_MAX_CORNERS + k] = status ? 1 : 0;
    }
}

// rows [x0, y0, x1-x0, y1-y0] (fp32) for corners with status 1, in corner order (optical_flow.py:73-82).
// One warp per genome: ballot + popcount compaction keeps the corner order.
__global__ void collect_vectors_kernel(const float* corners, const int* ncorners, const float* next_pts,
                                       const unsigned char* status, float* vectors, int* nvec, int B) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int nc = ncorners[b];
    int n = 0;
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        const long long i = (long long)b * FLOW_MAX_CORNERS + k;
        const bool keep = k < nc && status[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            float* v = vectors + ((long long)b * FLOW_MAX_CORNERS + n + __popc(m & ((1u << lane) - 1u))) * 4;
            v[0] = corners[i * 2]; v[1] = corners[i * 2 + 1];
            v[2] = __fsub_rn(next_pts[i * 2], corners[i * 2]);
            v[3] = __fsub_rn(next_pts[i * 2 + 1], corners[i * 2 + 1]);
        }
        n += __popc(m);
    }
    if (lane == 0) nvec[b] = n;
}

}  // namespace eig
