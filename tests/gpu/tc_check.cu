// GPU self-check of the tcgen05 convolution (conv_tc.cuh) - TEST ONLY, run by tests/test_gpu_parity.py on the B200.
//   part 1: raw accumulators (EPI_RAW) against a float64 CPU convolution
//   part 2: the three fused epilogues against the exact-fp32 SIMT kernel on identical inputs
// with the tiles-per-region cap at 1 (no weight-tile reuse) and unlimited; inputs are in split-fp16 storage, one output
// view is written in split-fp16 storage as the product does.
// usage: tc_check            exit code 0 = every shape / epilogue passes;  tc_check time = per-role cycle counters.
// Tolerances: the tensor core's fp32 accumulator truncates, so over K = 9*192 the result sits ~1e-5 (relative to the
// output scale) from float64 and from the SIMT kernel's round-to-nearest fp32 sums; the bar is 4e-5 + 1e-5 * K/1000.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../evolutionary_illusion_generator_b200/csrc/common.cuh"
#include "../../evolutionary_illusion_generator_b200/csrc/conv_simt.cuh"
#include "../../evolutionary_illusion_generator_b200/csrc/conv_tc.cuh"

using namespace eig;
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static void launch_simt(const ConvArgs& a) {
    const int tiles = ((a.W + 15) / 16) * ((a.H + 7) / 8);
    if (a.N <= 16) {
        const int nw = (a.N + 3) / 4;
        const size_t smem = (8 * 10 * 20 + 9 * 8 * nw * 4) * sizeof(float);
        conv3x3_simt_kernel<4><<<dim3(tiles, a.B, 1), dim3(32 * nw), smem>>>(a);
    } else {
        int nw = (a.N + 15) / 16;
        if (nw > 4) nw = 4;
        const int gz = (a.N + nw * 16 - 1) / (nw * 16);
        const size_t smem = (8 * 10 * 20 + 9 * 8 * nw * 16) * sizeof(float);
        conv3x3_simt_kernel<16><<<dim3(tiles, a.B, gz), dim3(32 * nw), smem>>>(a);
    }
    CHECK(cudaGetLastError());
}

template <class T> static T* dev(const std::vector<T>& h) {
    T* d; CHECK(cudaMalloc(&d, h.size() * sizeof(T) + 64)); CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d;
}
template <class T> static std::vector<T> host(const T* d, size_t n) {
    std::vector<T> h(n); CHECK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost)); return h;
}
static double g_tol = 6e-5;
static double maxdiff(const std::vector<float>& a, const std::vector<float>& b, double* scale) {
    double m = 0, s = 0;
    for (size_t i = 0; i < a.size(); ++i) { m = std::max(m, (double)fabsf(a[i] - b[i])); s = std::max(s, (double)fabsf(b[i])); if (a[i] != a[i]) m = 1e30; }
    *scale = s; return m;
}

struct Problem {
    int B, H, W, pitch, coff, Cin, N;
    std::vector<float> hi, wv, bias;
    float *d_hi, *d_lo, *d_w, *d_b;
    TcWeights tw;
};

static void make_problem(Problem& p, int B, int H, int W, int pitch, int coff, int Cin, int N, unsigned seed, int max_ncta = 0) {
    p.B = B; p.H = H; p.W = W; p.pitch = pitch; p.coff = coff; p.Cin = Cin; p.N = N;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    const size_t px = (size_t)B * H * W;
    p.hi.assign(px * pitch, 0.f);
    // the activation buffer is in split-fp16 storage (two fp16 planes: hi, then lo); the references use the value the
    // planes encode, (hi + lo) / 16
    std::vector<h16> planes(2 * px * pitch);
    for (size_t i = 0; i < px * pitch; ++i) {
        float v = u(rng);
        if (i % 7 == 3) v *= 1e-3f;     // small values: lo becomes a subnormal half
        if (i % 11 == 5) v *= 4.f;      // larger activations
        split16(v, &planes[i], &planes[px * pitch + i]);
        p.hi[i] = join16(planes[i], planes[px * pitch + i]);
    }
    p.wv.resize((size_t)9 * Cin * N);
    const float sc = 1.f / sqrtf(9.f * Cin);
    for (auto& v : p.wv) v = u(rng) * sc * 1.7f;
    p.bias.resize(N);
    for (auto& v : p.bias) v = u(rng) * 0.1f;
    p.d_hi = reinterpret_cast<float*>(dev(planes)); p.d_w = dev(p.wv); p.d_b = dev(p.bias);
    p.d_lo = reinterpret_cast<float*>(reinterpret_cast<h16*>(p.d_hi) + px * pitch);
    if (tc_pack(p.tw, p.wv.data(), Cin, N, N, max_ncta ? max_ncta : (N <= 128 ? 128 : 256)) || !p.tw.ok) { printf("tc_pack failed: %s\n", tc_last_error().c_str()); exit(2); }
}

static ConvArgs base_args(const Problem& p) {
    ConvArgs a; memset(&a, 0, sizeof a);
    a.in_hi = p.d_hi; a.in_lo = p.d_lo; a.in_pitch = p.pitch; a.in_coff = p.coff; a.Cin = p.Cin;
    a.B = p.B; a.H = p.H; a.W = p.W; a.wgt = p.d_w; a.bias = p.d_b; a.N = p.N; a.Npad = p.N;
    return a;
}

static bool check_raw(Problem& p, int mode) {
    ConvArgs a = base_args(p);
    a.epi = EPI_RAW;
    const size_t px = (size_t)p.B * p.H * p.W;
    float* d_out; CHECK(cudaMalloc(&d_out, px * p.N * 4)); CHECK(cudaMemset(d_out, 0xff, px * p.N * 4));
    a.outP = d_out;
    if (tc_conv(p.tw, a, 0)) { printf("  tc_conv failed: %s\n", tc_last_error().c_str()); return false; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  mode %d: kernel error %s\n", mode, cudaGetErrorString(e)); exit(3); }
    std::vector<float> got = host(d_out, px * p.N);
    // CPU reference on a sample of pixels (all channels)
    std::mt19937 rng(7);
    double worst = 0, scale = 0;
    const int samples = 600;
    for (int sidx = 0; sidx < samples; ++sidx) {
        int b = rng() % p.B, y = rng() % p.H, x = rng() % p.W;
        if (sidx < 8) { y = (sidx & 1) ? p.H - 1 : 0; x = (sidx & 2) ? p.W - 1 : 0; b = (sidx & 4) ? p.B - 1 : 0; }
        for (int n = 0; n < p.N; ++n) {
            double acc = 0;
            for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
                const int yy = y + ky - 1, xx = x + kx - 1;
                if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
                const size_t base = (((size_t)b * p.H + yy) * p.W + xx) * p.pitch + p.coff;
                for (int c = 0; c < p.Cin; ++c)
                    acc += (double)p.hi[base + c] * (double)p.wv[((size_t)(ky * 3 + kx) * p.Cin + c) * p.N + n];
            }
            acc += p.bias[n];
            const double g = got[(((size_t)b * p.H + y) * p.W + x) * p.N + n];
            const double d = fabs(g - acc);
            worst = std::max(worst, g == g ? d : 1e30); scale = std::max(scale, fabs(acc));
        }
    }
    cudaFree(d_out);
    const bool ok = worst <= g_tol * std::max(scale, 1.0);
    printf("  raw  B%d %dx%d Cin%d(coff %d) N%d max_nt %d: worst abs err %.3e (scale %.2f) %s\n", p.B, p.W, p.H, p.Cin, p.coff, p.N, mode, worst, scale, ok ? "ok" : "FAIL");
    return ok;
}

static View mkview(float* hi, float* lo, int pitch, int coff, int C) { View v; v.hi = hi; v.lo = lo; v.pitch = pitch; v.coff = coff; v.C = C; return v; }

static bool cmp(const char* what, const float* d_a, const float* d_b, size_t n, double tol) {
    std::vector<float> a = host(d_a, n), b = host(d_b, n);
    double sc; const double m = maxdiff(a, b, &sc);
    const bool ok = m <= tol * std::max(sc, 1.0);
    printf("    %-10s max |tc - simt| = %.3e (scale %.2f) %s\n", what, m, sc, ok ? "ok" : "FAIL");
    return ok;
}

static bool check_epilogues(Problem& p, int mode) {
    bool ok = true;
    const size_t px = (size_t)p.B * p.H * p.W;
    std::mt19937 rng(99);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    printf("  epilogues B%d %dx%d Cin%d N%d max_nt %d\n", p.B, p.W, p.H, p.Cin, p.N, mode);
    {   // ConvP (relu + clip)
        float *o1, *o2; CHECK(cudaMalloc(&o1, px * p.N * 4)); CHECK(cudaMalloc(&o2, px * p.N * 4));
        ConvArgs a = base_args(p); a.epi = EPI_CONVP; a.clip = 1;
        a.outP = o1; if (tc_conv(p.tw, a, 0)) { printf("tc_conv: %s\n", tc_last_error().c_str()); return false; }
        a.outP = o2; launch_simt(a);
        CHECK(cudaDeviceSynchronize());
        ok &= cmp("ConvP", o1, o2, px * p.N, g_tol);
        cudaFree(o1); cudaFree(o2);
    }
    if (!(p.H & 1) && !(p.W & 1) && p.tw.Ncta <= 128) {   // ConvA (pool + error units, hi/lo split)
        const size_t pp = px / 4;
        std::vector<float> P(pp * p.N); for (auto& v : P) v = fabsf(u(rng));
        float* dP = dev(P);
        float* e[2];
        for (auto& q : e) { CHECK(cudaMalloc(&q, pp * 2 * p.N * 4)); CHECK(cudaMemset(q, 0, pp * 2 * p.N * 4)); }
        ConvArgs a = base_args(p); a.epi = EPI_CONVA; a.P = dP;
        a.dstE = mkview(e[0], nullptr, 2 * p.N, 0, 2 * p.N); if (tc_conv(p.tw, a, 0)) { printf("tc_conv: %s\n", tc_last_error().c_str()); return false; }
        a.dstE = mkview(e[1], nullptr, 2 * p.N, 0, 2 * p.N); launch_simt(a);
        CHECK(cudaDeviceSynchronize());
        ok &= cmp("ConvA", e[0], e[1], pp * 2 * p.N, g_tol);
        for (auto& q : e) cudaFree(q); cudaFree(dP);
    }
    if (p.N % 16 == 0) {   // LSTM
        const int R = p.N / 4;
        std::vector<float> c0(px * R), peep((size_t)p.H * p.W * R * 4);
        for (auto& v : c0) v = u(rng); for (auto& v : peep) v = u(rng) * 0.1f;
        float* dpe = dev(peep);
        float *cs[2], *hh[2], *up[2];
        for (auto& q : cs) q = dev(c0);
        for (auto& q : hh) { CHECK(cudaMalloc(&q, px * (R + 8) * 4)); CHECK(cudaMemset(q, 0, px * (R + 8) * 4)); }
        for (auto& q : up) { CHECK(cudaMalloc(&q, px * 4 * (R + 4) * 4)); CHECK(cudaMemset(q, 0, px * 4 * (R + 4) * 4)); }
        ConvArgs a = base_args(p); a.epi = EPI_LSTM; a.peep = dpe;
        // h goes to a plain fp32 view, the up-sampled copy to a split-fp16 view (the product layout of the concat buffers)
        const size_t n_up = px * 4 * (R + 4);
        auto lo_of = [&](float* b) { return reinterpret_cast<float*>(reinterpret_cast<h16*>(b) + n_up); };
        a.cstate = cs[0]; a.dstH = mkview(hh[0], nullptr, R + 8, 4, R); a.dstUp = mkview(up[0], lo_of(up[0]), R + 4, 4, R);
        if (tc_conv(p.tw, a, 0)) { printf("tc_conv: %s\n", tc_last_error().c_str()); return false; }
        a.cstate = cs[1]; a.dstH = mkview(hh[1], nullptr, R + 8, 4, R); a.dstUp = mkview(up[1], lo_of(up[1]), R + 4, 4, R);
        launch_simt(a);
        CHECK(cudaDeviceSynchronize());
        ok &= cmp("LSTM c", cs[0], cs[1], px * R, g_tol);
        ok &= cmp("LSTM h", hh[0], hh[1], px * (R + 8), g_tol);
        {
            std::vector<h16> ua = host(reinterpret_cast<const h16*>(up[0]), 2 * n_up), ub = host(reinterpret_cast<const h16*>(up[1]), 2 * n_up);
            std::vector<float> fa(n_up), fb(n_up);
            for (size_t i = 0; i < n_up; ++i) { fa[i] = join16(ua[i], ua[n_up + i]); fb[i] = join16(ub[i], ub[n_up + i]); }
            double sc; const double md = maxdiff(fa, fb, &sc);
            const bool o2 = md <= g_tol * std::max(sc, 1.0);
            printf("    %-10s max |tc - simt| = %.3e (scale %.2f) %s\n", "LSTM up16", md, sc, o2 ? "ok" : "FAIL");
            ok &= o2;
        }
        for (auto& q : cs) cudaFree(q); for (auto& q : hh) cudaFree(q); for (auto& q : up) cudaFree(q); cudaFree(dpe);
    }
    return ok;
}

static int g_passes = 7;   // MMA products per k-step in timing mode (TcParams::passes)
// timing mode: `tc_check time` runs the PredNet layer shapes at population 32 with the per-role cycle counters on
static void time_shape(const char* name, int B, int H, int W, int pitch, int coff, int Cin, int N, int epi, int max_nt) {
    tc_set_max_nt(max_nt);
    Problem p; make_problem(p, B, H, W, pitch, coff, Cin, N, 5, epi == EPI_CONVA ? 128 : 0);   // the product packs ConvA with <= 128 channels per CTA pair (pooling tile)
    const size_t px = (size_t)B * H * W;
    ConvArgs a = base_args(p);
    a.epi = epi;
    float *out = nullptr, *cst = nullptr, *peep = nullptr, *dh = nullptr, *Pp = nullptr, *dE = nullptr;
    if (epi == EPI_CONVP) { CHECK(cudaMalloc(&out, px * N * 4)); a.outP = out; }
    if (epi == EPI_LSTM) {
        const int R = N / 4;
        CHECK(cudaMalloc(&cst, px * R * 4)); CHECK(cudaMemset(cst, 0, px * R * 4));
        CHECK(cudaMalloc(&peep, (size_t)H * W * R * 16)); CHECK(cudaMemset(peep, 0, (size_t)H * W * R * 16));
        CHECK(cudaMalloc(&dh, px * R * 4));
        a.cstate = cst; a.peep = peep; a.dstH = mkview(dh, nullptr, R, 0, R);
    }
    if (epi == EPI_CONVA) {
        CHECK(cudaMalloc(&Pp, px / 4 * N * 4)); CHECK(cudaMemset(Pp, 0, px / 4 * N * 4));
        CHECK(cudaMalloc(&dE, px / 4 * 2 * N * 4));
        a.P = Pp; a.dstE = mkview(dE, nullptr, 2 * N, 0, 2 * N);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_product = 0;
    {   // the product instantiation (no cycle counters)
        for (int i = 0; i < 3; ++i) tc_conv(p.tw, a, 0, g_passes);
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) tc_conv(p.tw, a, 0, g_passes);
        cudaEventRecord(e1);
        CHECK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_product, e0, e1);
    }
    long long* dbg; CHECK(cudaMalloc(&dbg, 160 * 16 * 8)); CHECK(cudaMemset(dbg, 0, 160 * 16 * 8));
    tc_state().dbg = dbg;
    for (int i = 0; i < 3; ++i) tc_conv(p.tw, a, 0, g_passes);
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) if (tc_conv(p.tw, a, 0, g_passes)) { printf("tc_conv: %s\n", tc_last_error().c_str()); exit(2); }
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> d = host(dbg, 160 * 16);
    const int grid = tc_state().last_grid;
    double s[16] = {0};
    for (int c = 0; c < grid; ++c) for (int k = 0; k < 16; ++k) s[k] += (double)d[c * 16 + k] / grid;
    for (int k = 0; k < 5; ++k) s[k] *= 2;
    s[14] *= 2; s[15] *= 2;   // the MMA warp only runs in the leader CTA of each pair
    const double flop = 2.0 * px * 9.0 * Cin * N;
    printf("%-8s B%d %dx%d Cin%d N%d cluster %d NT %d SA %d SB %d grid %d: %7.1f us %6.1f TFLOP/s(fp32-equiv) [instrumented %.1f us] | MMA thr: total %6.0f waitAcc %5.0f waitA %5.0f waitB %5.0f issue %6.0f commit %6.0f regions %.1f | epi: total %6.0f wait %6.0f | conv: total %6.0f wait %6.0f | Bprod: wait %6.0f | kernel: prologue %5.0f body %6.0f teardown %5.0f = %.1f us @1.965GHz (cycles, avg per CTA)\n",
           name, B, W, H, Cin, N, 2, tc_state().last_nt, tc_state().last_sa, tc_state().last_sb, grid, 1e2 * ms_product,
           flop / (1e-4 * ms_product) / 1e12, 1e3 * ms / reps, s[0], s[1], s[2], s[3], s[14], s[15], s[4], s[5], s[6], s[7], s[8], s[10], s[11], s[12], s[13], (s[11] + s[12] + s[13]) / 1965.0);
    tc_state().dbg = nullptr;
    cudaFree(dbg); cudaFree(out); cudaFree(cst); cudaFree(peep); cudaFree(dh); cudaFree(Pp); cudaFree(dE);
    cudaFree(p.d_hi); cudaFree(p.d_w); cudaFree(p.d_b); tc_free(p.tw);
}

// `tc_check time c3 [passes]`: the eight convolutions of one PredNet step of BASELINE configs[2] (pop 128 colour)
static int timing_c3(int passes) {
    g_passes = passes;
    if (const char* e = getenv("EIG_TC_DBGFLAGS")) tc_state().dbg_flags = atoi(e);
    printf("--- workload C3 shapes (B 128, channels 3,48,96,192), MMA products mask %d, debug flags %d (1 = no operand loads, 2 = no epilogue work)\n", passes, tc_state().dbg_flags);
    time_shape("A2", 128, 60, 80, 240, 0, 96, 96, EPI_CONVA, 0);
    time_shape("A3", 128, 30, 40, 480, 0, 192, 192, EPI_CONVA, 0);
    time_shape("LSTM3", 128, 15, 20, 576, 0, 576, 768, EPI_LSTM, 0);
    time_shape("LSTM2", 128, 30, 40, 480, 0, 480, 384, EPI_LSTM, 0);
    time_shape("LSTM1", 128, 60, 80, 240, 0, 240, 192, EPI_LSTM, 0);
    time_shape("P1+Z", 128, 60, 80, 240, 192, 48, 96, EPI_CONVP, 0);
    time_shape("P2", 128, 30, 40, 480, 384, 96, 96, EPI_CONVP, 0);
    time_shape("P3", 128, 15, 20, 576, 384, 192, 192, EPI_CONVP, 0);
    g_passes = 7;
    return 0;
}

static int timing_main() {
    if (const char* e = getenv("EIG_TC_DBGFLAGS")) tc_state().dbg_flags = atoi(e);
    if (const char* e = getenv("EIG_TC_PASSES")) g_passes = atoi(e);
    printf("--- workload C2 shapes, MMA products mask %d, debug flags %d\n", g_passes, tc_state().dbg_flags);
    {
        for (int max_nt = 0; max_nt <= 0; ++max_nt) {
            printf("--- product kernel: CTA pairs, lo*hi, hi*lo (keep A), hi*hi (re-use A)\n");
            time_shape("LSTM1", 32, 60, 80, 80, 0, 80, 64, EPI_LSTM, max_nt);
            time_shape("LSTM2", 32, 30, 40, 160, 0, 160, 128, EPI_LSTM, max_nt);
            time_shape("LSTM3", 32, 15, 20, 192, 0, 192, 256, EPI_LSTM, max_nt);
            time_shape("ConvA2", 32, 60, 80, 80, 0, 32, 32, EPI_CONVA, max_nt);
            time_shape("ConvA3", 32, 30, 40, 160, 0, 64, 64, EPI_CONVA, max_nt);
            time_shape("ConvP1", 32, 60, 80, 80, 64, 16, 16, EPI_CONVP, max_nt);
            time_shape("ConvP2", 32, 30, 40, 160, 128, 32, 32, EPI_CONVP, max_nt);
            time_shape("ConvP3", 32, 15, 20, 192, 128, 64, 64, EPI_CONVP, max_nt);
            time_shape("LSTM1x4", 128, 60, 80, 80, 0, 80, 64, EPI_LSTM, max_nt);
            time_shape("A1gray", 32, 120, 160, 8, 0, 8, 16, EPI_CONVA, max_nt);
            time_shape("A1col", 128, 120, 160, 8, 0, 8, 48, EPI_CONVA, max_nt);
        }
    }
    return 0;
}

int main(int argc, char** argv) {
    if (!tc_available()) { printf("tensor-core path unavailable: %s\n", tc_unavailable_reason().c_str()); return 4; }
    if (argc > 2 && !strcmp(argv[1], "time") && !strcmp(argv[2], "c3")) return timing_c3(argc > 3 ? atoi(argv[3]) : 7);
    if (argc > 1 && !strcmp(argv[1], "time")) return timing_main();
    struct Shape { int B, H, W, pitch, coff, Cin, N; };
    const Shape shapes[] = {
        {2, 15, 20, 192, 0, 192, 256},   // gray ConvLSTM3
        {2, 30, 40, 160, 0, 160, 128},   // gray ConvLSTM2
        {5, 60, 80, 80, 0, 80, 64},      // gray ConvLSTM1 (partial last 32-channel block)
        {2, 60, 80, 80, 64, 16, 16},     // gray ConvP1 (view at a channel offset)
        {2, 30, 40, 160, 0, 64, 32},     // gray ConvA-like
        {1, 16, 24, 40, 8, 32, 48},      // odd sizes
        {40, 30, 40, 96, 0, 96, 96},     // more regions than SMs: the persistent loop + accumulator double buffering
        {3, 30, 40, 480, 0, 480, 384},   // colour ConvLSTM2: N split over two CTA columns
    };
    bool all_ok = true;
    struct Cfg { int max_nt; };
    const Cfg cfgs[] = {{0}, {1}};
    for (const Cfg& c : cfgs) {
        printf("== tiles per region %s ==\n", c.max_nt ? "capped at 1" : "auto");
        tc_set_max_nt(c.max_nt);
        unsigned seed = 1;
        bool all = true;
        for (const Shape& s : shapes) {
            Problem p; make_problem(p, s.B, s.H, s.W, s.pitch, s.coff, s.Cin, s.N, seed++);
            g_tol = 4e-5 + 1e-5 * (9.0 * s.Cin / 1000.0);
            bool ok = check_raw(p, c.max_nt);
            if (ok) ok &= check_epilogues(p, c.max_nt);
            all &= ok;
            cudaFree(p.d_hi); cudaFree(p.d_w); cudaFree(p.d_b); tc_free(p.tw);
        }
        printf("== %s ==\n", all ? "PASS" : "FAIL");
        all_ok &= all;
    }
    printf("RESULT %s\n", all_ok ? "PASS" : "FAIL");
    return all_ok ? 0 : 1;
}
