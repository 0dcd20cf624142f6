// K9 score_vectors: one warp per genome turns its <= 100 flow vectors into the fitness scalar.
//
// Restates the live scoring branches of the reference with the same mixed fp32 / fp64 typing numpy 2.x
// gives them (rows are float32, python-float literals are weak scalars):
//   plausibility_ratio        /root/reference/fitness_calculator.py:18-27
//   strength_number           fitness_calculator.py:32-41     (mean |dx| only - `my` is unused)
//   horizontal_symmetry_score fitness_calculator.py:81-120    (line 101 broadcasts x into both columns)
//   swarm_score               fitness_calculator.py:124-159   (`% 2 * math.pi` precedence, fp32 throughout)
//   rotation_symmetry_score   fitness_calculator.py:166-215
//   branch logic              /root/reference/generate_illusion.py:557-616
// Zero-length vectors give 0/0 = NaN exactly like the reference; no guard.
// All sums are warp-shuffle reductions in fp64 (deterministic: fixed lane order).
#pragma once
#include "common.cuh"
#include "flow.cuh"

namespace eig {

enum { STRUCT_BANDS = 0, STRUCT_CIRCLES = 1, STRUCT_FREE = 2, STRUCT_CIRCLES_FREE = 3 };

struct ScoreArgs {
    const float* vectors;  // [B][100][4]
    const int* nvec;       // [B]
    double* fitness;       // [B]
    double* status;        // nullable: 1 sticky slot, set when the split-fp16 range flag was raised during this evaluation
    int B, structure, w, h;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// population variance and mean of n values held as (lane-strided) arrays in shared memory
__device__ __forceinline__ void warp_mean_var(const double* v, int n, int lane, double* mean, double* var) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += v[i];
    s = warp_sum_d(s);
    const double m = s / n;
    double q = 0.0;
    for (int i = lane; i < n; i += 32) { const double d = v[i] - m; q += d * d; }
    q = warp_sum_d(q);
    *mean = m;
    *var = q / n;
}

__global__ void __launch_bounds__(32) score_kernel(ScoreArgs a) {
    constexpr int MAXV = FLOW_MAX_CORNERS;
    __shared__ float sx[MAXV], sy[MAXV], sdx[MAXV], sdy[MAXV];
    __shared__ double t0[MAXV], t1[MAXV];
    __shared__ float sang[MAXV];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= a.B) return;
    if (b == 0 && lane == 0 && a.status && g_eig_range_flag) { *a.status = 1.0; g_eig_range_flag = 0; }
    const int nraw = a.nvec[b];
    const float limit = a.structure == STRUCT_BANDS ? 0.15f : (a.structure == STRUCT_FREE ? 0.4f : 0.3f);
    // plausibility filter, order preserving compaction
    int n = 0;
    for (int base = 0; base < nraw; base += 32) {
        const int i = base + lane;
        bool keep = false;
        float vx = 0.f, vy = 0.f, dx = 0.f, dy = 0.f;
        if (i < nraw) {
            const float* v = a.vectors + ((long long)b * MAXV + i) * 4;
            vx = v[0]; vy = v[1]; dx = v[2]; dy = v[3];
            const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            keep = !(norm > limit);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            sx[pos] = vx; sy[pos] = vy; sdx[pos] = dx; sdy[pos] = dy;
        }
        n += __popc(m);
    }
    __syncwarp();
    double score = 0.0;
    const float fpi = 3.14159265358979323846f;

    // strength_number(good, limit): fp32 result
    float strength = 0.f;
    if (n > 0 && a.structure != STRUCT_BANDS) {
        double sabs = 0.0;
        for (int i = lane; i < n; i += 32) {
            sabs += (double)fabsf(sdx[i]);
            t0[i] = (double)__fsqrt_rn(__fadd_rn(__fmul_rn(sdx[i], sdx[i]), __fmul_rn(sdy[i], sdy[i])));
        }
        __syncwarp();
        sabs = warp_sum_d(sabs);
        double mean, var;
        warp_mean_var(t0, n, lane, &mean, &var);
        const float mx = (float)(sabs / n);
        const float vf = (float)var;
        strength = __fmul_rn(__fdiv_rn(mx, limit), __fsub_rn(1.f, vf < 1.f ? vf : 1.f));
        __syncwarp();
    }

    if (a.structure == STRUCT_CIRCLES || a.structure == STRUCT_CIRCLES_FREE) {
        if (n > 24) {
            const float cxf = (float)(a.w / 2.0), cyf = (float)(a.h / 2.0);
            const float r1 = (float)(a.h / 2.0);
            // compact the rows inside the radius limits
            int m = 0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                bool keep = false;
                double rx = 0.0, ry = 0.0;
                if (i < n) {
                    const float pxf = __fsub_rn(sx[i], cxf), pyf = __fsub_rn(sy[i], cyf);
                    const float df = __fsqrt_rn(__fadd_rn(__fmul_rn(pxf, pxf), __fmul_rn(pyf, pyf)));
                    keep = !(df < 0.f || df > r1 || df == 0.f);
                    if (keep) {
                        const double px = pxf, py = pyf, dist = df;
                        double ux = sdx[i], uy = sdy[i];
                        const double nrm = sqrt(ux * ux + uy * uy);
                        ux = ux / nrm; uy = uy / nrm;
                        const double ex = px + ux, ey = py + uy;
                        rx = (ex * px + ey * py) / dist - dist;
                        ry = (-ex * py + ey * px) / dist;
                    }
                }
                const unsigned bm = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int pos = m + __popc(bm & ((1u << lane) - 1u));
                    t0[pos] = rx; t1[pos] = ry;
                }
                m += __popc(bm);
            }
            __syncwarp();
            double rot = 0.0;
            if (m >= 2) {
                double mean, vx, vy;
                warp_mean_var(t0, m, lane, &mean, &vx);
                warp_mean_var(t1, m, lane, &mean, &vy);
                rot = ((1 - vx) * (1 - vx) + (1 - vy) * (1 - vy)) / 2;
            }
            score = 0.7 * rot + 0.3 * (double)strength;
        }
    } else if (a.structure == STRUCT_BANDS) {
        if (n > 0) {
            const float lim1 = (float)((a.h / 4.0) * 2);
            const int middle = (int)(((a.h / 4.0) * 2) / 2);
            int m = 0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                bool keep = false;
                double c0 = 0.0, c1 = 0.0;
                if (i < n) {
                    const float y = sy[i];
                    keep = !(y < 0.f || y > lim1);
                    if (keep) {
                        const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(sdx[i], sdx[i]), __fmul_rn(sdy[i], sdy[i])));
                        const float nx = __fdiv_rn(sdx[i], nrm), ny = __fdiv_rn(sdy[i], nrm);
                        if (y < (float)middle) { c0 = nx; c1 = nx; }
                        else { c0 = -nx; c1 = ny; }
                    }
                }
                const unsigned bm = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int pos = m + __popc(bm & ((1u << lane) - 1u));
                    t0[pos] = c0; t1[pos] = c1;
                }
                m += __popc(bm);
            }
            __syncwarp();
            if (m > 0) {
                double mean0, var0, mean1, var1;
                warp_mean_var(t0, m, lane, &mean0, &var0);
                warp_mean_var(t1, m, lane, &mean1, &var1);
                score = ((1 - var0) + fabs(mean0) + (1 - fabs(mean1))) / 3;
            }
        }
    } else if (a.structure == STRUCT_FREE) {
        if (n > 0) {
            // swarm_score: everything is float32 in the reference
            for (int i = lane; i < n; i += 32) {
                const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(sdx[i], sdx[i]), __fmul_rn(sdy[i], sdy[i])));
                const float nx = __fdiv_rn(sdx[i], nrm);
                sang[i] = acosf(nx);
                t0[i] = (double)nx;
            }
            __syncwarp();
            float total = 0.f;
            for (int aidx = 0; aidx < n; ++aidx) {
                const float ax = sx[aidx], ay = sy[aidx];
                const float v_angle = (float)acos(t0[aidx]);
                double loss = 0.0;
                for (int j = lane; j < n; j += 32) {
                    const float ddx = __fsub_rn(sx[j], ax), ddy = __fsub_rn(sy[j], ay);
                    float f = __fdiv_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), 10000.f);
                    f = f > 1.f ? 1.f : f;
                    const float close = f < 1.f ? 1.f : 0.f;
                    const float opt = __fmul_rn(fmodf(__fadd_rn(v_angle, __fmul_rn(f, fpi)), 2.f), fpi);
                    loss += (double)__fmul_rn(close, fabsf(__fsub_rn(sang[j], opt)));
                }
                loss = warp_sum_d(loss);
                const float temp = __fsub_rn(fpi, __fdiv_rn((float)loss, (float)n));
                total = __fadd_rn(total, __fdiv_rn(temp, fpi));
            }
            const float swarm = __fdiv_rn(total, (float)n);
            const double number = (double)(n < 15 ? n : 15) / 15;
            score = 0.5 * (double)swarm + 0.1 * (double)strength + 0.4 * number;
        }
    }
    if (lane == 0) a.fitness[b] = score;
}

}  // namespace eig
