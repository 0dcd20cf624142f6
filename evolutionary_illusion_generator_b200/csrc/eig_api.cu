// libeig.so: context, weight repacking, stage orchestration and the C ABI (include/eig.h).
// Host-side C++ only orchestrates; every number on the hot path is produced by the CUDA kernels in
// render.cuh / conv_simt.cuh / conv_tc.cuh / flow.cuh / score.cuh.  There is no CPU fallback.
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#ifdef EIG_EMU
#include "cuda_emu.h"
#endif
#include "common.cuh"
#include "render.cuh"
#include "conv_simt.cuh"
#include "conv_l0.cuh"
#include "flow.cuh"
#include "score.cuh"
#ifndef EIG_EMU
#include "conv_tc.cuh"
#endif
#include "../../include/eig.h"


using namespace eig;
#define TO_STREAM(p) ((cudaStream_t)(intptr_t)(p))

// Optional per-kernel-class device timing (bench.py's roofline pass): every launch is bracketed by a pair of
// CUDA events on the launching stream; eig_profile_end sums them per class.
enum { CLS_RENDER = 0, CLS_CONV_SIMT = 1, CLS_CONV_TC = 2, CLS_ELEMENTWISE = 3, CLS_FLOW = 4, CLS_SCORE = 5, CLS_L0 = 6, CLS_COUNT = 8 };
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
};

// Error text and launch instrumentation belong to the context an entry point was called with: every extern "C" function
// that takes a context opens a Scope, and fail() / the launch macros reach the context through it.  g_err keeps the
// most recent message of the calling thread for eig_last_error() (failures before a context exists: eig_create).
struct CtxCommon { std::string err; Profiler prof; };
static thread_local std::string g_err;
static thread_local CtxCommon* g_cur = nullptr;
struct Scope {
    CtxCommon* prev;
    explicit Scope(CtxCommon* c) : prev(g_cur) { g_cur = c; }
    ~Scope() { g_cur = prev; }
};
static int fail(int code, const std::string& msg) {
    g_err = msg;
    if (g_cur) g_cur->err = msg;
    return code;
}
#define CK(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(EIG_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));              \
    } while (0)
#define CKL()                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = cudaGetLastError();                                                          \
        if (e_ != cudaSuccess) return fail(EIG_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

static bool prof_on() { return g_cur && g_cur->prof.on; }
static void prof_pre(int cls, cudaStream_t s) {
    if (!prof_on()) return;
    Profiler& pr = g_cur->prof;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    pr.ev.push_back(a); pr.ev.push_back(b); pr.cls.push_back(cls);
}
static void prof_post(cudaStream_t s) {
    if (!prof_on()) return;
    cudaEventRecord(g_cur->prof.ev.back(), s);
}
#define LAUNCH_K(cls, kernel, grid, block, smem, s, ...)              \
    do {                                                              \
        prof_pre(cls, s);                                             \
        EIG_LAUNCH(kernel, grid, block, smem, s, __VA_ARGS__);        \
        prof_post(s);                                                 \
        EIG_COUNT_LAUNCH();                                           \
    } while (0)

#define LAUNCH_K_PDL(cls, kernel, grid, block, smem, s, ...)          \
    do {                                                              \
        prof_pre(cls, s);                                             \
        EIG_LAUNCH_PDL(kernel, grid, block, smem, s, __VA_ARGS__);    \
        prof_post(s);                                                 \
        EIG_COUNT_LAUNCH();                                           \
    } while (0)

struct LayerW {            // repacked weights of one PredNet layer (device)
    float* convA = nullptr; float* convA_b = nullptr;   // [9][2C_{n-1}][Npad]
    float* convP = nullptr; float* convP_b = nullptr;   // [9][R_n][Npad]
    float* lstm = nullptr;  float* lstm_b = nullptr;    // [9][Ctot][4R], bias [4R] gate-interleaved
    float* peep = nullptr;                              // [H][W][R][4]
#ifndef EIG_EMU
    TcWeights tcA, tcP, tcL;                            // tensor-core layouts (conv_tc.cuh)
    TcWeights tcZ;                                      // folded up-sampled-R taps of THIS layer's ConvLSTM: R_{n+1} -> Z_n (layers 1, 2)
#endif
};

struct eig_ctx : CtxCommon {
    int device = 0, w = 0, h = 0, c_dim = 0, ch[4] = {0, 0, 0, 0}, cap = 0;
    int H[4], W[4], ctot[4];
    int conv_mode = EIG_CONV_SIMT;
    bool have_w = false, have_grid = false;
    bool render_only = false;        // eig_create_render: only the CPPN stage (grid planes, image, input frame) exists
    double *xmat = nullptr, *ymat = nullptr;
    LayerW lw[4];
    // activations
    float* X[4][2] = {{nullptr}};    // concat buffers [B][H][W][ctot] = [E_n | up(R_{n+1}) | h_n], double-buffered over time steps
    float* cst[4] = {nullptr};       // cell state [B][H][W][R]
    float* h0[2] = {nullptr, nullptr};  // layer-0 hidden state [B][h][w][C0], double-buffered over time steps
    float* P[4] = {nullptr};         // predictions [B][H][W][C]
    h16* E0s = nullptr;              // [2 planes][B][h][w][8] split-fp16 E0, input of ConvA1 on the tensor cores (conv_l0.cuh)
    bool conva1_tc = false;          // ConvA1 runs on the tcgen05 kernel (set at weight load: C1 >= 32)
    float* Z = nullptr;              // [B][H/2][W/2][16*C0] partial sums of ConvLSTM0's up-sampled-R1 taps (conv_l0.cuh)
    // Folding for layers 1 and 2 (tensor-core mode): ConvLSTM_n's taps over the nearest-neighbour up-sampled R_{n+1}
    // collapse, per pixel parity, to 2x2 taps at half resolution; Zf[n] = [B][H_{n+1}][W_{n+1}][4 parities][4*C_n] holds
    // those partial sums (one tap-masked convolution of R_{n+1}, launched right after ConvLSTM_{n+1}), ConvLSTM_n skips
    // the K blocks of its up(R) slice and adds Zf[n] in its epilogue, and ConvLSTM_{n+1} no longer writes the 2x2-replicated
    // copy of R_{n+1}.  -22 % of the MMAs of ConvLSTM1/2.
    float* Zf[3] = {nullptr, nullptr, nullptr};
    int fold = -1;                   // eig_set_option "fold": 0 off, 1 on wherever the shapes allow, -1 auto (by population size)
    bool skip_zero_state = true;     // eig_set_option "skip_zero_state": step 0 skips the K blocks that only hold the zero state
    int npz = 0;                     // columns of the layer-1 ConvP+Z convolution: C1 + 16*C0
    float* x_in = nullptr;           // [B][h][w][c]
    unsigned char* img = nullptr;    // rendered [B][h][w][c]
    unsigned char* frames = nullptr; // [3][B][h][w][c]
    // flow
    int n_levels = 1, lh[FLOW_MAX_LEVELS], lwid[FLOW_MAX_LEVELS];
    unsigned char* gray[FLOW_MAX_LEVELS] = {nullptr};  // [2B][lh][lw]  (frame 1 images first, then frame 2)
    short* deriv[FLOW_MAX_LEVELS] = {nullptr};         // [B][lh][lw][2]
    float* eigmap = nullptr; int* eigmax = nullptr; unsigned long long* cand = nullptr;
    float* corners = nullptr; int* ncorners = nullptr; float* next_pts = nullptr; unsigned char* status = nullptr;
    float* vectors = nullptr; int* nvec = nullptr;
    double* fitness = nullptr;
    // host staging for eig_eval_host
    void* d_blob = nullptr; size_t d_blob_cap = 0; long long* d_off = nullptr;
    size_t corner_smem = 0;   // bitmap of the greedy corner spacing: one bit per pixel
    int seq_t = 0, seq_n = -1;   // stateful stepping (eig_prednet_reset / eig_prednet_forward)
    void* h_pin = nullptr; size_t h_pin_cap = 0;
    cudaStream_t stream = 0;
    // ConvP_2 / ConvP_3 are off the critical path of a PredNet step: they run on a side stream, ordered by events
    cudaStream_t side = 0;
    cudaEvent_t ev_lstm[4] = {nullptr, nullptr, nullptr, nullptr}, ev_p[4] = {nullptr, nullptr, nullptr, nullptr};
    bool p_pending[4] = {false, false, false, false};
    bool overlap = true;
    // CUDA graphs of everything after the render (PredNet sequence + flow + score), keyed by what the launches depend on;
    // a key is run once un-captured (lazy initialisations), captured on its second use and replayed afterwards
    struct GraphEntry { int seen = 0; long long launches = 0; void* exec = nullptr; };
    std::map<std::tuple<int, int, int, int>, GraphEntry> graphs;
    bool use_graphs = true;
    // tensor-core numerics: MMA products per k-step for each convolution (kind 0 ConvA, 1 ConvP, 2 ConvLSTM) x layer, as a
    // 3-bit mask (conv_tc.cuh TcParams::passes; 7 = a_lo*w_hi + a_hi*w_lo + a_hi*w_hi).  Steps t < early_until use
    // (mask & early_mask) instead.  Set through eig_set_option; the defaults are what the parity tests pin.
    int passes[3][4] = {{7, 7, 7, 7}, {7, 7, 7, 7}, {7, 7, 7, 7}};
    int early_until = 0, early_mask = 7;
    int precision = 0;
    int simt_reverse_taps = 0;   // diagnostic: the exact-fp32 kernel sums the taps in reverse order (ConvArgs::rev_taps)
#ifndef EIG_EMU
    TcMapCache amaps;   // activation tensor maps of this context's buffers
#endif
    std::vector<void*> allocs;
};
enum { KIND_A = 0, KIND_P = 1, KIND_L = 2 };
static const long long FOLD_AUTO_MIN_WORK = 200000000LL;   // "fold" auto: B * H_n * W_n * C_n * C_{n+1} from which the folded form is used (measured: +10 % at 16 colour genomes = 3.5e8, -1.5 % at 32 gray genomes = 7.9e7; profiles/r2/bench_c3_pop16_fold{0,1}_f.json, bench_c3_*_f.json "also")
// Precision profiles of the tensor-core path (eig_set_option "precision"; measured in profiles/r2/pass_ablation_*.md):
//   0 exact    : three products (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi) in every convolution - fp32-grade, 2^-22 per product
//   1 balanced : single fp16 product in the convolutions of layers 2 and 3 (ConvA2/3, ConvP2/3, ConvLSTM2/3), whose
//                rounding never reaches the uint8 frames (frame bytes that differ from the exact-fp32 path: unchanged),
//                three products in layer 1 (ConvA1, ConvLSTM1, ConvP1 + the ConvLSTM0 partial sums), which drive P0 directly
//   2 fast     : single product everywhere (frames still within 1 LSB; ~15x more bytes differ than in profile 0)
static void apply_precision(eig_ctx* c, int profile) {
    for (int kd = 0; kd < 3; ++kd)
        for (int n = 1; n < 4; ++n) c->passes[kd][n] = profile == 0 ? 7 : profile == 2 ? 4 : (n == 1 ? 7 : 4);
    c->precision = profile;
}
static int passes_for(const eig_ctx* c, int kind, int layer, int t) {
    int m = c->passes[kind][layer];
    if (t < c->early_until && (m & c->early_mask)) m &= c->early_mask;
    return m;
}

template <class T>
static cudaError_t dalloc(eig_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e == cudaSuccess) { c->allocs.push_back(q); *p = (T*)q; }
    return e;
}

// captured graphs hold raw pointers to the weight buffers: a new weight file invalidates them
static void drop_graphs(eig_ctx* c) {
#ifndef EIG_EMU
    for (auto& kv : c->graphs) if (kv.second.exec) cudaGraphExecDestroy((cudaGraphExec_t)kv.second.exec);
#endif
    c->graphs.clear();
}

extern "C" const char* eig_last_error(void) { return g_err.c_str(); }
extern "C" const char* eig_error(const eig_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
extern "C" int eig_version(void) { return 100; }
extern "C" int64_t eig_launch_count(void) { return launch_counter().n; }

static int create_buffers(eig_ctx* c, int w, int h, int c_dim, const int channels[4], int max_genomes);
// the folded form needs the K blocks of [E_n | up(R_{n+1}) | h_n] to start on 32-channel boundaries and one parity
// (4*C_n gate columns) to be a whole number of N slices
static bool fold_shape_ok(const eig_ctx* c, int n) {
    return n >= 1 && n <= 2 && (2 * c->ch[n]) % 32 == 0 && c->ch[n + 1] % 32 == 0 && (4 * c->ch[n]) % 32 == 0;
}

extern "C" int eig_create(eig_ctx** out, int device, int w, int h, int c_dim, const int channels[4], int max_genomes) {
    if (!out || !channels || max_genomes <= 0) return fail(EIG_E_INVALID, "eig_create: null/invalid argument");
    if (w % 8 || h % 8 || w <= 0 || h <= 0) return fail(EIG_E_INVALID, "eig_create: w and h must be positive multiples of 8");
    if (c_dim != channels[0] || (c_dim != 1 && c_dim != 3)) return fail(EIG_E_INVALID, "eig_create: c_dim must equal channels[0] and be 1 or 3");
    if (channels[1] > 64) return fail(EIG_E_INVALID, "eig_create: channels[1] > 64 is not supported by the fused layer-0 kernel");
    if (w <= FLOW_WIN || h <= FLOW_WIN) return fail(EIG_E_INVALID, "eig_create: image must be larger than the 50-px LK window");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(EIG_E_NODEVICE, "eig_create: no CUDA device visible; this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EIG_E_INVALID, "eig_create: bad device index");
    CK(cudaSetDevice(device));
    eig_ctx* c = new eig_ctx();
    Scope sc(c);
    c->device = device;
    const int rc = create_buffers(c, w, h, c_dim, channels, max_genomes);
    if (rc != EIG_OK) { g_cur = sc.prev; eig_destroy(c); return rc; }   // nothing of a half-built context survives (e.g. out of memory)
    *out = c;
    return EIG_OK;
}

// A context for the CPPN stage alone (`get_image_from_cppn` / the 800x800 `enhanced.png` mosaic, generate_illusion.py:
// 372-460, 664-671): no PredNet, no flow, hence none of their size rules (any w, h > 0).
extern "C" int eig_create_render(eig_ctx** out, int device, int w, int h, int c_dim, int max_genomes) {
    if (!out || max_genomes <= 0 || w <= 0 || h <= 0) return fail(EIG_E_INVALID, "eig_create_render: null/invalid argument");
    if (c_dim != 1 && c_dim != 3) return fail(EIG_E_INVALID, "eig_create_render: c_dim must be 1 or 3");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(EIG_E_NODEVICE, "eig_create_render: no CUDA device visible; this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EIG_E_INVALID, "eig_create_render: bad device index");
    CK(cudaSetDevice(device));
    eig_ctx* c = new eig_ctx();
    Scope sc(c);
    c->device = device; c->render_only = true;
    c->w = w; c->h = h; c->c_dim = c_dim; c->cap = max_genomes; c->ch[0] = c_dim;
    const size_t npx = (size_t)max_genomes * w * h;
    cudaError_t e = dalloc(c, &c->x_in, npx * c_dim);
    if (e == cudaSuccess) e = dalloc(c, &c->img, npx * c_dim);
    if (e == cudaSuccess) e = dalloc(c, &c->xmat, (size_t)w * h);
    if (e == cudaSuccess) e = dalloc(c, &c->ymat, (size_t)w * h);
    if (e == cudaSuccess) e = dalloc(c, &c->d_off, (size_t)max_genomes + 1);
    if (e != cudaSuccess) { g_cur = sc.prev; eig_destroy(c); return fail(EIG_E_CUDA, std::string("eig_create_render: ") + cudaGetErrorString(e)); }
    c->use_graphs = false; c->overlap = false;
    *out = c;
    return EIG_OK;
}

static int create_buffers(eig_ctx* c, int w, int h, int c_dim, const int channels[4], int max_genomes) {
    c->w = w; c->h = h; c->c_dim = c_dim; c->cap = max_genomes;
    for (int n = 0; n < 4; ++n) { c->ch[n] = channels[n]; c->H[n] = h >> n; c->W[n] = w >> n; }
    for (int n = 0; n < 4; ++n) c->ctot[n] = 2 * c->ch[n] + (n < 3 ? c->ch[n + 1] : 0) + c->ch[n];
    const size_t B = max_genomes;
    for (int n = 0; n < 4; ++n) {
        const size_t px = B * c->H[n] * c->W[n];
        for (int k = 0; k < 2; ++k) {
            if (n >= 1) CK(dalloc(c, &c->X[n][k], px * c->ctot[n]));   // layer 0 keeps no concat buffer (conv_l0.cuh)
            else CK(dalloc(c, &c->h0[k], px * c->ch[0]));
        }
        CK(dalloc(c, &c->cst[n], px * c->ch[n]));
        CK(dalloc(c, &c->P[n], px * c->ch[n]));
    }
    const size_t npx = B * w * h;
    c->npz = c->ch[1] + 16 * c->ch[0];
    CK(dalloc(c, &c->Z, B * c->H[1] * c->W[1] * 16 * c->ch[0]));
    for (int n = 1; n <= 2; ++n)
        if (fold_shape_ok(c, n)) CK(dalloc(c, &c->Zf[n], B * c->H[n + 1] * c->W[n + 1] * 16 * c->ch[n]));
    CK(dalloc(c, &c->E0s, 2 * npx * 8));
    CK(dalloc(c, &c->x_in, npx * c_dim));
    CK(dalloc(c, &c->img, npx * c_dim));
    CK(dalloc(c, &c->frames, 3 * npx * c_dim));
    // pyramid geometry (buildOpticalFlowPyramid: stop when the next level is <= the window)
    c->n_levels = 1; c->lh[0] = h; c->lwid[0] = w;
    for (int l = 1; l < FLOW_MAX_LEVELS; ++l) {
        const int nw2 = (c->lwid[l - 1] + 1) / 2, nh2 = (c->lh[l - 1] + 1) / 2;
        if (nw2 <= FLOW_WIN || nh2 <= FLOW_WIN) break;
        c->lwid[l] = nw2; c->lh[l] = nh2; c->n_levels = l + 1;
    }
    for (int l = 0; l < c->n_levels; ++l) {
        CK(dalloc(c, &c->gray[l], 2 * B * c->lh[l] * c->lwid[l]));
        CK(dalloc(c, &c->deriv[l], 2 * B * c->lh[l] * c->lwid[l]));
    }
    CK(dalloc(c, &c->eigmap, npx)); CK(dalloc(c, &c->eigmax, B)); CK(dalloc(c, &c->cand, npx));
    CK(dalloc(c, &c->corners, B * FLOW_MAX_CORNERS * 2)); CK(dalloc(c, &c->ncorners, B));
    CK(dalloc(c, &c->next_pts, B * FLOW_MAX_CORNERS * 2)); CK(dalloc(c, &c->status, B * FLOW_MAX_CORNERS));
    CK(dalloc(c, &c->vectors, B * FLOW_MAX_CORNERS * 4)); CK(dalloc(c, &c->nvec, B));
    CK(dalloc(c, &c->fitness, B + 1)); CK(dalloc(c, &c->d_off, B + 1));   // fitness[B] = status slot (sticky range flag)
    CK(cudaMemset(c->fitness + B, 0, sizeof(double)));
    CK(dalloc(c, &c->xmat, (size_t)w * h)); CK(dalloc(c, &c->ymat, (size_t)w * h));
    c->corner_smem = (((size_t)w * h + 31) / 32) * sizeof(unsigned);
    if (c->corner_smem > 200 * 1024) return fail(EIG_E_CAPACITY, "eig_create: image too large for the corner-selection bitmap (w*h <= 1.6 Mpx)");
#ifndef EIG_EMU
    if (c->corner_smem + 17 * 1024 > 48 * 1024)
        CK(cudaFuncSetAttribute(corner_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->corner_smem));
    {   // the host entry point runs on its own high-priority stream; the side stream (ConvP_2/3, off the critical path)
        // gets the lowest priority so that its CTAs only take SMs the critical-path kernels leave idle
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest));
        CK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, least));
    }
    for (int n = 2; n < 4; ++n) {
        CK(cudaEventCreateWithFlags(&c->ev_lstm[n], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_p[n], cudaEventDisableTiming));
    }
    if (const char* e = getenv("EIG_PRECISION")) { const int v = atoi(e); if (v >= 0 && v <= 2) apply_precision(c, v); }
    if (const char* e = getenv("EIG_FOLD")) { const int v = atoi(e); c->fold = v < 0 ? -1 : (v != 0); }
    if (const char* e = getenv("EIG_NO_OVERLAP")) c->overlap = atoi(e) == 0;
    if (const char* e = getenv("EIG_NO_GRAPH")) c->use_graphs = atoi(e) == 0;
#else
    c->overlap = false;
    c->use_graphs = false;
#endif
    return EIG_OK;
}

extern "C" void eig_destroy(eig_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
#ifndef EIG_EMU
    for (int n = 0; n < 4; ++n) { tc_free(c->lw[n].tcA); tc_free(c->lw[n].tcP); tc_free(c->lw[n].tcL); tc_free(c->lw[n].tcZ); }
#endif
#ifndef EIG_EMU
    drop_graphs(c);
    for (int n = 2; n < 4; ++n) { if (c->ev_lstm[n]) cudaEventDestroy(c->ev_lstm[n]); if (c->ev_p[n]) cudaEventDestroy(c->ev_p[n]); }
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
#endif
    for (cudaEvent_t e : c->prof.ev) cudaEventDestroy(e);
    for (void* p : c->allocs) cudaFree(p);
    if (c->d_blob) cudaFree(c->d_blob);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    delete c;
}

extern "C" int eig_set_conv_mode(eig_ctx* c, int mode) {
    Scope sc(c);
    if (!c) return fail(EIG_E_INVALID, "null ctx");
    if (mode != EIG_CONV_SIMT && mode != EIG_CONV_TC) return fail(EIG_E_INVALID, "unknown conv mode");
#ifdef EIG_EMU
    if (mode == EIG_CONV_TC) return fail(EIG_E_INVALID, "tensor-core path is not available in the emulator build");
#else
    if (mode == EIG_CONV_TC && !tc_available()) return fail(EIG_E_INVALID, "tensor-core path unavailable: " + tc_unavailable_reason());
#endif
    c->conv_mode = mode;
    return EIG_OK;
}

extern "C" int eig_set_option(eig_ctx* c, const char* key, int value) {
    if (!c || !key) return fail(EIG_E_INVALID, "eig_set_option: null argument");
    Scope sc(c);
    const std::string k = key;
    bool ok = false;
    if (k == "early_until") { c->early_until = value; ok = true; }
    else if (k == "early_mask") { if (!(value & 7)) return fail(EIG_E_INVALID, "eig_set_option: empty pass mask"); c->early_mask = value & 7; ok = true; }
    else if (k == "precision") {
        if (value < 0 || value > 2) return fail(EIG_E_INVALID, "eig_set_option: precision must be 0 (exact), 1 (balanced) or 2 (fast)");
        apply_precision(c, value); ok = true;
    }
    else if (k == "skip_zero_state") { c->skip_zero_state = value != 0; ok = true; }
    else if (k == "fold") { c->fold = value < 0 ? -1 : (value != 0); ok = true; }
    else if (k == "simt_reverse_taps") { c->simt_reverse_taps = value != 0; ok = true; }
    else if (k == "graphs") { c->use_graphs = value != 0; ok = true; }
    else if (k == "overlap") { c->overlap = value != 0; ok = true; }
    else if (k.compare(0, 7, "passes.") == 0) {
        if (!(value & 7)) return fail(EIG_E_INVALID, "eig_set_option: empty pass mask");
        const std::string t = k.substr(7);
        const char kinds[3] = {'A', 'P', 'L'};
        for (int kd = 0; kd < 3; ++kd)
            for (int n = 1; n < 4; ++n)
                if (t == "all" || (t.size() == 1 && t[0] == kinds[kd]) || (t.size() == 2 && t[0] == kinds[kd] && t[1] == '0' + n)) { c->passes[kd][n] = value & 7; ok = true; }
    }
    if (!ok) return fail(EIG_E_INVALID, "eig_set_option: unknown key " + k);
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());
    drop_graphs(c);   // captured graphs bake the launch parameters in
    return EIG_OK;
}

extern "C" int eig_set_grid(eig_ctx* c, const double* hx, const double* hy) {
    Scope sc(c);
    if (!c || !hx || !hy) return fail(EIG_E_INVALID, "eig_set_grid: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(c->xmat, hx, sizeof(double) * c->w * c->h, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->ymat, hy, sizeof(double) * c->w * c->h, cudaMemcpyHostToDevice));
    c->have_grid = true;
    return EIG_OK;
}

// ------------------------------------------------------------------------------------------------ weights
namespace {
struct HostT { const float* p; int64_t s[4]; };
typedef std::map<std::string, HostT> WMap;

int need(const WMap& m, const std::string& k, int64_t s0, int64_t s1, int64_t s2, int64_t s3, const HostT** out) {
    auto it = m.find(k);
    if (it == m.end()) return fail(EIG_E_INVALID, "eig_load_weights: missing tensor " + k);
    const int64_t* s = it->second.s;
    if (s[0] != s0 || s[1] != s1 || s[2] != s2 || s[3] != s3) {
        char buf[256];
        snprintf(buf, sizeof buf, "eig_load_weights: %s has shape (%lld,%lld,%lld,%lld), expected (%lld,%lld,%lld,%lld)",
                 k.c_str(), (long long)s[0], (long long)s[1], (long long)s[2], (long long)s[3], (long long)s0,
                 (long long)s1, (long long)s2, (long long)s3);
        return fail(EIG_E_INVALID, buf);
    }
    *out = &it->second;
    return EIG_OK;
}
template <class T>
cudaError_t upload(eig_ctx* c, T** dst, const std::vector<T>& v) {
    cudaError_t e = dalloc(c, dst, v.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}
// Chainer conv weight (Cout, Cin, 3, 3) -> dst[tap][cin_off + cin][Npad] at column col(n)
void scatter_conv(std::vector<float>& dst, int cin_total, int npad, int cin_off, const HostT* t, int cout, int cin,
                  int col_mul, int col_add) {
    for (int n = 0; n < cout; ++n)
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < 9; ++tap)
                dst[((size_t)tap * cin_total + cin_off + ci) * npad + n * col_mul + col_add] =
                    t->p[((size_t)n * cin + ci) * 9 + tap];
}
// A ConvLSTM's 3x3 taps over the nearest-neighbour up-sampled R_{n+1}, folded to half resolution: output pixel (2Y+py, 2X+px)
// reads full-resolution row 2Y+py+ky-1 = low-resolution row Y + floor((py+ky-1)/2), so per parity the taps {0,1,2}
// collapse onto low-resolution offsets {-1,0,0} (p = 0) or {0,0,+1} (p = 1).  Zero padding agrees on both grids.
// dst: [9][Cup][npad]; the folded columns start at col0: column col0 + (py*2+px)*NG + n.
// w: the ConvLSTM's weights [9][ctot][NG]; the R_{n+1} channels are [coff, coff + Cup).
static const int kLowOff[2][3] = {{-1, 0, 0}, {0, 0, 1}};
void build_fold_weights(std::vector<float>& dst, int npad, int col0, int Cup, const std::vector<float>& w, int NG, int ctot, int coff) {
    std::vector<double> acc((size_t)9 * Cup * 4 * NG, 0.0);
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    const int tap_lr = (kLowOff[py][ky] + 1) * 3 + (kLowOff[px][kx] + 1);
                    for (int ch = 0; ch < Cup; ++ch)
                        for (int n = 0; n < NG; ++n)
                            acc[((size_t)tap_lr * Cup + ch) * 4 * NG + (py * 2 + px) * NG + n] +=
                                (double)w[((size_t)(ky * 3 + kx) * ctot + coff + ch) * NG + n];
                }
    for (int tap = 0; tap < 9; ++tap)
        for (int ch = 0; ch < Cup; ++ch)
            for (int j = 0; j < 4 * NG; ++j)
                dst[((size_t)tap * Cup + ch) * npad + col0 + j] = (float)acc[((size_t)tap * Cup + ch) * 4 * NG + j];
}
// the low-resolution taps parity (py, px) reads: bit = (dy+1)*3 + (dx+1)
unsigned short fold_tap_mask(int parity) {
    unsigned m = 0;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) m |= 1u << ((kLowOff[parity >> 1][ky] + 1) * 3 + (kLowOff[parity & 1][kx] + 1));
    return (unsigned short)m;
}
// layer 0: the Z columns ride behind the C1 ConvP1 columns (conv_l0.cuh)
void build_z_weights(std::vector<float>& dst, int npad, int C1, const std::vector<float>& w0, int C0, int ctot0) {
    build_fold_weights(dst, npad, C1, C1, w0, 4 * C0, ctot0, 2 * C0);
}
}  // namespace

extern "C" int eig_load_weights(eig_ctx* c, int nt, const char* const* names, const float* const* ptrs, const int64_t* shapes) {
    Scope sc(c);
    if (!c || !names || !ptrs || !shapes || nt <= 0) return fail(EIG_E_INVALID, "eig_load_weights: null argument");
    if (c->render_only) return fail(EIG_E_STATE, "eig_load_weights: render-only context");
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());   // no evaluation may still be reading the old weights
    drop_graphs(c);
    WMap m;
    for (int i = 0; i < nt; ++i) {
        std::string k = names[i];
        const std::string pre = "predictor/";
        if (k.compare(0, pre.size(), pre) == 0) k = k.substr(pre.size());
        HostT t; t.p = ptrs[i];
        for (int d = 0; d < 4; ++d) t.s[d] = shapes[i * 4 + d];
        m[k] = t;
    }
    const char gates[4] = {'i', 'f', 'c', 'o'};
    std::vector<float> lstm0_host;   // [9][ctot0][4*C0], needed again for the layer-1 ConvP+Z weights
    for (int n = 0; n < 4; ++n) {
        LayerW& L = c->lw[n];
        const int C = c->ch[n], R = C, Hn = c->H[n], Wn = c->W[n];
        const HostT* t = nullptr;
        int rc;
        char nm[96];
        if (n >= 1) {
            const int cin = 2 * c->ch[n - 1], npad = (C + 3) & ~3;
            std::vector<float> wv((size_t)9 * cin * npad, 0.f), bv(C);
            snprintf(nm, sizeof nm, "ConvA%d/W", n);
            if ((rc = need(m, nm, C, cin, 3, 3, &t))) return rc;
            scatter_conv(wv, cin, npad, 0, t, C, cin, 1, 0);
            snprintf(nm, sizeof nm, "ConvA%d/b", n);
            if ((rc = need(m, nm, C, 1, 1, 1, &t))) return rc;
            for (int i = 0; i < C; ++i) bv[i] = t->p[i];
            CK(upload(c, &L.convA, wv)); CK(upload(c, &L.convA_b, bv));
#ifndef EIG_EMU
            if (n >= 2 && (rc = tc_pack(L.tcA, wv.data(), cin, C, npad, 128))) return fail(EIG_E_CUDA, "tc_pack ConvA: " + tc_last_error());
            if (n == 1) {   // ConvA1 over the 8-channel split-fp16 E0 tensor (channels >= 2*C0 are zero)
                std::vector<float> w8((size_t)9 * 8 * npad, 0.f);
                for (int tap = 0; tap < 9; ++tap)
                    for (int ci = 0; ci < cin; ++ci)
                        for (int k = 0; k < npad; ++k) w8[((size_t)tap * 8 + ci) * npad + k] = wv[((size_t)tap * cin + ci) * npad + k];
                if ((rc = tc_pack(L.tcA, w8.data(), 8, C, npad, 128))) return fail(EIG_E_CUDA, "tc_pack ConvA1: " + tc_last_error());
                c->conva1_tc = L.tcA.ok && C >= 32;
                if (const char* e = getenv("EIG_CONVA1_TC")) c->conva1_tc = L.tcA.ok && atoi(e) != 0;
            }
#endif
        }
        {
            // layer 1 carries the Z columns of ConvLSTM0 next to its ConvP columns (see build_z_weights)
            const int cin = R, ncol = n == 1 ? c->npz : C, npad = (ncol + 3) & ~3;
            std::vector<float> wv((size_t)9 * cin * npad, 0.f), bv(npad, 0.f);
            snprintf(nm, sizeof nm, "ConvP%d/W", n);
            if ((rc = need(m, nm, C, cin, 3, 3, &t))) return rc;
            scatter_conv(wv, cin, npad, 0, t, C, cin, 1, 0);
            snprintf(nm, sizeof nm, "ConvP%d/b", n);
            if ((rc = need(m, nm, C, 1, 1, 1, &t))) return rc;
            for (int i = 0; i < C; ++i) bv[i] = t->p[i];
            if (n == 1) build_z_weights(wv, npad, C, lstm0_host, c->ch[0], c->ctot[0]);
            CK(upload(c, &L.convP, wv)); CK(upload(c, &L.convP_b, bv));
#ifndef EIG_EMU
            if (n >= 1 && (rc = tc_pack(L.tcP, wv.data(), cin, ncol, npad))) return fail(EIG_E_CUDA, "tc_pack ConvP: " + tc_last_error());
#endif
        }
        {
            const int ctot = c->ctot[n], N = 4 * R;
            const int rup = n < 3 ? c->ch[n + 1] : 0;
            std::vector<float> wv((size_t)9 * ctot * N, 0.f), bv(N), pv((size_t)Hn * Wn * R * 4, 0.f);
            for (int g = 0; g < 4; ++g) {
                snprintf(nm, sizeof nm, "ConvLSTM%d/x_%c0/W", n, gates[g]);
                if ((rc = need(m, nm, R, 2 * C, 3, 3, &t))) return rc;
                scatter_conv(wv, ctot, N, 0, t, R, 2 * C, 4, g);
                if (n < 3) {
                    snprintf(nm, sizeof nm, "ConvLSTM%d/x_%c1/W", n, gates[g]);
                    if ((rc = need(m, nm, R, rup, 3, 3, &t))) return rc;
                    scatter_conv(wv, ctot, N, 2 * C, t, R, rup, 4, g);
                }
                snprintf(nm, sizeof nm, "ConvLSTM%d/h_%c/W", n, gates[g]);
                if ((rc = need(m, nm, R, R, 3, 3, &t))) return rc;
                scatter_conv(wv, ctot, N, 2 * C + rup, t, R, R, 4, g);
                snprintf(nm, sizeof nm, "ConvLSTM%d/h_%c/b", n, gates[g]);
                if ((rc = need(m, nm, R, 1, 1, 1, &t))) return rc;
                for (int r = 0; r < R; ++r) bv[r * 4 + g] = t->p[r];
            }
            const char pg[3] = {'i', 'f', 'o'};
            for (int g = 0; g < 3; ++g) {
                snprintf(nm, sizeof nm, "ConvLSTM%d/c_%c/W", n, pg[g]);
                if ((rc = need(m, nm, 1, R, Hn, Wn, &t))) return rc;
                for (int r = 0; r < R; ++r)
                    for (int y = 0; y < Hn; ++y)
                        for (int x = 0; x < Wn; ++x)
                            pv[(((size_t)y * Wn + x) * R + r) * 4 + g] = t->p[((size_t)r * Hn + y) * Wn + x];
            }
            CK(upload(c, &L.lstm, wv)); CK(upload(c, &L.lstm_b, bv)); CK(upload(c, &L.peep, pv));
            if (n == 0) lstm0_host = wv;
#ifndef EIG_EMU
            if (n >= 1 && (rc = tc_pack(L.tcL, wv.data(), ctot, N, N))) return fail(EIG_E_CUDA, "tc_pack ConvLSTM: " + tc_last_error());
            if (fold_shape_ok(c, n)) {   // Z_n = folded up(R_{n+1}) taps: [9][C_{n+1}][4 parities x N], one parity per group of N slices
                std::vector<float> zw((size_t)9 * rup * 4 * N, 0.f);
                build_fold_weights(zw, 4 * N, 0, rup, wv, N, ctot, 2 * C);
                int ncta = 0;
                for (int d = 32; d <= 256 && d <= N; d += 32) if (N % d == 0) ncta = d;   // largest slice that divides one parity
                if (ncta && (rc = tc_pack(L.tcZ, zw.data(), rup, 4 * N, 4 * N, ncta))) return fail(EIG_E_CUDA, "tc_pack fold: " + tc_last_error());
                if (L.tcZ.ok && (L.tcZ.Ncta != ncta || L.tcZ.gz > 16)) tc_free(L.tcZ);
                if (L.tcZ.ok) for (int z = 0; z < L.tcZ.gz; ++z) L.tcZ.tap_mask[z] = fold_tap_mask(z * ncta / N);
            }
#endif
        }
    }
    c->have_w = true;
    return EIG_OK;
}

// ------------------------------------------------------------------------------------------------ launches
static int launch_conv(eig_ctx* c, const ConvArgs& a, cudaStream_t s) {
    const int tiles = ((a.W + 15) / 16) * ((a.H + 7) / 8);
    if (a.N <= 16) {
        const int nw = (a.N + 3) / 4;
        const size_t smem = (8 * 10 * 20 + 9 * 8 * nw * 4) * sizeof(float);
        auto k = c->simt_reverse_taps ? conv3x3_simt_kernel<4, true> : conv3x3_simt_kernel<4, false>;
        LAUNCH_K(CLS_CONV_SIMT, k, dim3(tiles, a.B, 1), dim3(32 * nw), smem, s, a);
    } else {
        int nw = (a.N + 15) / 16;
        if (nw > 4) nw = 4;
        const int gz = (a.N + nw * 16 - 1) / (nw * 16);
        const size_t smem = (8 * 10 * 20 + 9 * 8 * nw * 16) * sizeof(float);
        auto k = c->simt_reverse_taps ? conv3x3_simt_kernel<16, true> : conv3x3_simt_kernel<16, false>;
        LAUNCH_K(CLS_CONV_SIMT, k, dim3(tiles, a.B, gz), dim3(32 * nw), smem, s, a);
    }
    CKL();
    return EIG_OK;
}

static View mkview(float* hi, float* lo, int pitch, int coff, int C) { View v; v.hi = hi; v.lo = lo; v.pitch = pitch; v.coff = coff; v.C = C; return v; }
// lo plane of a layer's concat buffer: in tensor-core mode the buffer holds two fp16 planes (split-fp16 storage, common.cuh
// View) in the bytes the fp32 tensor would take; the lo plane starts after cap * H * W * ctot halves.  nullptr = plain fp32.
static float* lo_plane(eig_ctx* c, int n, float* base) {
    if (c->conv_mode != EIG_CONV_TC) return nullptr;
    return reinterpret_cast<float*>(reinterpret_cast<h16*>(base) + (size_t)c->cap * c->H[n] * c->W[n] * c->ctot[n]);
}

static L0Args l0_args(eig_ctx* c, const float* x, int B, int cur, int nxt) {
    L0Args a;
    memset(&a, 0, sizeof a);
    a.B = B; a.H = c->h; a.W = c->w; a.C0 = c->ch[0]; a.C1 = c->ch[1];
    a.x = x; a.P0 = c->P[0];
    a.wA = c->lw[1].convA; a.bA = c->lw[1].convA_b; a.C1pad = (c->ch[1] + 3) & ~3;
    a.P1 = c->P[1];
    a.dstE1 = mkview(c->X[1][cur], lo_plane(c, 1, c->X[1][cur]), c->ctot[1], 0, 2 * c->ch[1]);
    a.wL = c->lw[0].lstm; a.bL = c->lw[0].lstm_b; a.peep = c->lw[0].peep;
    a.Z = c->Z;
    a.h_prev = c->h0[cur]; a.h_next = c->h0[nxt]; a.cstate = c->cst[0];
    a.wP = c->lw[0].convP; a.bP = c->lw[0].convP_b; a.C0pad = (c->ch[0] + 3) & ~3;
    a.P0_out = c->P[0];
    return a;
}

// One PredNet time step (net.py:175-211) for B genomes.  x: [B][h][w][c] input frame, t: step index.
// Layer 0 runs on the fused full-resolution kernels of conv_l0.cuh; layers 1..3 on the tcgen05 kernel (conv_mode TC,
// shapes with N % 16 == 0) or on the exact-fp32 SIMT kernel.
// `last`: no step follows, so the predictions P_2 / P_3 (only read by the next step's error units) are not computed
static int prednet_step(eig_ctx* c, const float* x, int B, int t, bool last, cudaStream_t s) {
    const int cur = t & 1, nxt = cur ^ 1;
    const bool first_step = t == 0 && c->skip_zero_state;   // every caller resets the state right before step 0
    const bool tc = c->conv_mode == EIG_CONV_TC;
    const bool side_ok = c->overlap && !prof_on();   // the per-class profiler times launches on one stream
    int rc;
    const L0Args l0 = l0_args(c, x, B, cur, nxt);
    const int l0_tiles = ((c->w + L0_TW - 1) / L0_TW) * ((c->h + L0_TH - 1) / L0_TH);
#ifndef EIG_EMU
    if (tc && c->conva1_tc) {   // E0 -> split-fp16 tensor, then ConvA1 + pool + E1 on the tcgen05 kernel
        const long long npix = (long long)B * c->h * c->w;
        h16* e_lo = c->E0s + (size_t)c->cap * c->h * c->w * 8;
        LAUNCH_K_PDL(CLS_L0, l0_e0_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, x, (const float*)c->P[0], c->E0s, e_lo, npix, c->ch[0]);
        CKL();
        ConvArgs a;
        memset(&a, 0, sizeof a);
        a.in_hi = reinterpret_cast<const float*>(c->E0s); a.in_lo = reinterpret_cast<const float*>(e_lo);
        a.in_pitch = 8; a.in_coff = 0; a.Cin = 8;
        a.B = B; a.H = c->h; a.W = c->w;
        a.wgt = nullptr; a.bias = c->lw[1].convA_b; a.N = c->ch[1]; a.Npad = (c->ch[1] + 3) & ~3;
        a.epi = EPI_CONVA; a.P = c->P[1];
        a.dstE = mkview(c->X[1][cur], lo_plane(c, 1, c->X[1][cur]), c->ctot[1], 0, 2 * c->ch[1]);
        prof_pre(CLS_CONV_TC, s); rc = tc_conv(c->lw[1].tcA, a, s, passes_for(c, KIND_A, 1, t), &c->amaps); prof_post(s); EIG_COUNT_LAUNCH();
        if (rc) return fail(EIG_E_CUDA, "tc_conv ConvA1: " + tc_last_error());
    } else
#endif
    {   // E0 -> ConvA1 -> pool -> E1 (layer-1 concat buffer)
        const int c1pad = l0.C1pad;
        const size_t smem = ((size_t)2 * 2 * l0.C0 * (L0_TH + 2) * (L0_TW + 2) + (size_t)9 * 2 * l0.C0 * c1pad + (size_t)64 * (c1pad + 1) + 16) * sizeof(float);
        const int n_items = l0_tiles * B;
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
        // persistent grid: a few CTAs per SM (the register file allows 2 of the 256-thread variant), at most one per item
        // persistent grid: as many CTAs as can be resident (occupancy query per variant), at most one per item
        auto grid_of = [&](const void* fn, int threads) {
            int per_sm = 1;
#ifndef EIG_EMU
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
#endif
            return dim3((unsigned)std::min(n_items, nsm * per_sm));
        };
        // narrow first layers (C1 <= 16, the gray network): one CTA per tile and genome, high occupancy instead of persistence
        const size_t smem_tile = ((size_t)2 * l0.C0 * (L0_TH + 2) * (L0_TW + 2) + (size_t)9 * 2 * l0.C0 * c1pad + (size_t)64 * (c1pad + 1) + 16) * sizeof(float);
        if (c1pad <= 4 && smem_tile <= 48 * 1024) {
            if (l0.C0 == 1) { auto k = l0_conva1_tile_kernel<4, 1>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(64), smem_tile, s, l0); }
            else { auto k = l0_conva1_tile_kernel<4, 3>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(64), smem_tile, s, l0); }
        } else if (c1pad <= 16 && smem_tile <= 48 * 1024) {
            const int th = 64 * ((c1pad + 7) / 8);
            if (l0.C0 == 1) { auto k = l0_conva1_tile_kernel<8, 1>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(th), smem_tile, s, l0); }
            else { auto k = l0_conva1_tile_kernel<8, 3>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(th), smem_tile, s, l0); }
        }
        else if (c1pad <= 4) { auto k = l0_conva1_kernel<4>; LAUNCH_K_PDL(CLS_L0, k, grid_of((const void*)k, 64), dim3(64), smem, s, l0, n_items); }
        else if (c1pad <= 16) { const int th = 64 * ((c1pad + 7) / 8); auto k = l0_conva1_kernel<8>; LAUNCH_K_PDL(CLS_L0, k, grid_of((const void*)k, th), dim3(th), smem, s, l0, n_items); }
        else if (c1pad <= 48) { const int th = 64 * ((c1pad + 11) / 12); auto k = l0_conva1_kernel<12>; LAUNCH_K_PDL(CLS_L0, k, grid_of((const void*)k, th), dim3(th), smem, s, l0, n_items); }
        else { const int th = 64 * ((c1pad + 15) / 16); auto k = l0_conva1_kernel<16>; LAUNCH_K_PDL(CLS_L0, k, grid_of((const void*)k, th), dim3(th), smem, s, l0, n_items); }
        CKL();
    }
    for (int n = 2; n < 4; ++n) {  // ConvA_n: E_{n-1} (res n-1) -> pool -> E_n
#ifndef EIG_EMU
        if (c->p_pending[n]) { CK(cudaStreamWaitEvent(s, c->ev_p[n], 0)); c->p_pending[n] = false; }   // its epilogue reads P_n
#endif
        ConvArgs a;
        memset(&a, 0, sizeof a);
        a.in_hi = c->X[n - 1][cur]; a.in_lo = lo_plane(c, n - 1, c->X[n - 1][cur]);
        a.in_pitch = c->ctot[n - 1]; a.in_coff = 0; a.Cin = 2 * c->ch[n - 1];
        a.B = B; a.H = c->H[n - 1]; a.W = c->W[n - 1];
        a.wgt = c->lw[n].convA; a.bias = c->lw[n].convA_b; a.N = c->ch[n]; a.Npad = (c->ch[n] + 3) & ~3;
        a.epi = EPI_CONVA; a.P = c->P[n];
        a.dstE = mkview(c->X[n][cur], lo_plane(c, n, c->X[n][cur]), c->ctot[n], 0, 2 * c->ch[n]);
#ifndef EIG_EMU
        if (tc && c->lw[n].tcA.ok && tc_view_ok(a)) {
            // step 0 of a sequence: P_{n-1} is still the zero state, so E-_{n-1} = relu(P - A) is exactly zero (A >= 0):
            // the K blocks that only hold E- contribute exact zeros and are skipped (bit-identical result)
            int skip_lo = 0, skip_hi = 0;
            if (first_step) { skip_lo = (c->ch[n - 1] + TC_KB - 1) / TC_KB; skip_hi = c->lw[n].tcA.KBn; }
            prof_pre(CLS_CONV_TC, s); rc = tc_conv(c->lw[n].tcA, a, s, passes_for(c, KIND_A, n, t), &c->amaps, skip_lo, skip_hi); prof_post(s); EIG_COUNT_LAUNCH();
            if (rc) return fail(EIG_E_CUDA, "tc_conv ConvA: " + tc_last_error());
            continue;
        }
#endif
        if ((rc = launch_conv(c, a, s))) return rc;
    }
    auto conv_p = [&](int n, cudaStream_t s) -> int {   // ConvP_n: R_n -> P_n (layer 1 also emits Z for ConvLSTM0)
        ConvArgs a;
        memset(&a, 0, sizeof a);
        const int hoff = 2 * c->ch[n] + (n < 3 ? c->ch[n + 1] : 0);
        a.in_hi = c->X[n][nxt]; a.in_lo = lo_plane(c, n, c->X[n][nxt]);
        a.in_pitch = c->ctot[n]; a.in_coff = hoff; a.Cin = c->ch[n];
        a.B = B; a.H = c->H[n]; a.W = c->W[n];
        a.wgt = c->lw[n].convP; a.bias = c->lw[n].convP_b;
        a.N = n == 1 ? c->npz : c->ch[n]; a.Npad = (a.N + 3) & ~3;
        a.epi = EPI_CONVP; a.outP = c->P[n]; a.clip = 0;
        if (n == 1) { a.nP = c->ch[1]; a.outZ = c->Z; }
#ifndef EIG_EMU
        if (tc && c->lw[n].tcP.ok && tc_view_ok(a)) { prof_pre(CLS_CONV_TC, s); const int r2 = tc_conv(c->lw[n].tcP, a, s, passes_for(c, KIND_P, n, t), &c->amaps); prof_post(s); EIG_COUNT_LAUNCH(); if (r2) return fail(EIG_E_CUDA, "tc_conv ConvP: " + tc_last_error()); return EIG_OK; }
#endif
        return launch_conv(c, a, s);
    };
#ifndef EIG_EMU
    // folded up-sampled-R taps (see eig_ctx::Zf): per consumer layer n = 1, 2
    auto fold_on = [&](int n) {
        if (!tc || n < 1 || n > 2 || !c->Zf[n] || !c->lw[n].tcZ.ok || !c->lw[n].tcL.ok || c->fold == 0) return false;
        // the producer of Z_n is the tcgen05 launch of ConvLSTM_{n+1}: that layer must be on the tensor-core path too
        const int hoff_up = 2 * c->ch[n + 1] + (n + 1 < 3 ? c->ch[n + 2] : 0);
        if (!c->lw[n + 1].tcL.ok || (c->ctot[n + 1] & 7) || (c->ctot[n] & 7) || (hoff_up & 7)) return false;
        if (c->fold == 1) return true;
        // small problems are launch-bound: one more launch costs more than the MMAs it saves.  Work saved ~ pixels x C_n x C_{n+1}
        return (long long)B * c->H[n] * c->W[n] * c->ch[n] * c->ch[n + 1] >= FOLD_AUTO_MIN_WORK;
    };
#else
    auto fold_on = [&](int) { return false; };
#endif
    for (int n = 3; n >= 1; --n) {  // ConvLSTM_n
        ConvArgs a;
        memset(&a, 0, sizeof a);
        a.in_hi = c->X[n][cur]; a.in_lo = lo_plane(c, n, c->X[n][cur]);
        a.in_pitch = c->ctot[n]; a.in_coff = 0; a.Cin = c->ctot[n];
        a.B = B; a.H = c->H[n]; a.W = c->W[n];
        a.wgt = c->lw[n].lstm; a.bias = c->lw[n].lstm_b; a.N = 4 * c->ch[n]; a.Npad = a.N;
        a.epi = EPI_LSTM; a.cstate = c->cst[n]; a.peep = c->lw[n].peep;
        const int hoff = 2 * c->ch[n] + (n < 3 ? c->ch[n + 1] : 0);
        a.dstH = mkview(c->X[n][nxt], lo_plane(c, n, c->X[n][nxt]), c->ctot[n], hoff, c->ch[n]);
        // R_n up-sampled x2 into the concat buffer of layer n-1 (layer 0 gets R_1 through Z instead)
        const bool fold_here = fold_on(n), fold_below = n >= 2 && fold_on(n - 1);
        if (n >= 2 && !fold_below) a.dstUp = mkview(c->X[n - 1][cur], lo_plane(c, n - 1, c->X[n - 1][cur]), c->ctot[n - 1], 2 * c->ch[n - 1], c->ch[n]);
#ifndef EIG_EMU
        if (tc && c->lw[n].tcL.ok && tc_view_ok(a)) {
            int skip_lo = 0, skip_hi = 0;
            if (fold_here) { a.Zin = c->Zf[n]; skip_lo = 2 * c->ch[n] / TC_KB; skip_hi = (2 * c->ch[n] + c->ch[n + 1]) / TC_KB; }
            if (first_step) {
                // step 0: E-_n (P_n is the zero state) and h_n (the zero state) are exactly zero.  With the up-sampled slice
                // folded away (or absent: layer 3) everything behind E+_n goes; otherwise only the h tail can (one range).
                const int kbn = c->lw[n].tcL.KBn;
                if (fold_here || n == 3) { skip_lo = (c->ch[n] + TC_KB - 1) / TC_KB; skip_hi = kbn; }
                else if ((2 * c->ch[n] + c->ch[n + 1]) % TC_KB == 0) { skip_lo = (2 * c->ch[n] + c->ch[n + 1]) / TC_KB; skip_hi = kbn; }
            }
            prof_pre(CLS_CONV_TC, s); rc = tc_conv(c->lw[n].tcL, a, s, passes_for(c, KIND_L, n, t), &c->amaps, skip_lo, skip_hi); prof_post(s); EIG_COUNT_LAUNCH();
            if (rc) return fail(EIG_E_CUDA, "tc_conv ConvLSTM: " + tc_last_error());
            if (fold_below) {   // Z_{n-1}: the tap-masked half-resolution convolution of the h_n just written, raw fp32 partial sums
                ConvArgs z;
                memset(&z, 0, sizeof z);
                z.in_hi = c->X[n][nxt]; z.in_lo = lo_plane(c, n, c->X[n][nxt]);
                z.in_pitch = c->ctot[n]; z.in_coff = hoff; z.Cin = c->ch[n];
                z.B = B; z.H = c->H[n]; z.W = c->W[n];
                z.N = 16 * c->ch[n - 1]; z.Npad = z.N; z.epi = EPI_RAW; z.outP = c->Zf[n - 1];
                prof_pre(CLS_CONV_TC, s); rc = tc_conv(c->lw[n - 1].tcZ, z, s, passes_for(c, KIND_L, n - 1, t), &c->amaps); prof_post(s); EIG_COUNT_LAUNCH();
                if (rc) return fail(EIG_E_CUDA, "tc_conv fold: " + tc_last_error());
            }
        }
        else
#endif
        if ((rc = launch_conv(c, a, s))) return rc;
#ifndef EIG_EMU
        if (side_ok && n >= 2 && !last) {   // ConvP_n only needs h_n: it overlaps ConvLSTM_{n-1} .. ConvP_0 of this step
            CK(cudaEventRecord(c->ev_lstm[n], s));
            CK(cudaStreamWaitEvent(c->side, c->ev_lstm[n], 0));
            if ((rc = conv_p(n, c->side))) return rc;
            CK(cudaEventRecord(c->ev_p[n], c->side));
            c->p_pending[n] = true;
        }
#endif
    }
    if ((rc = conv_p(1, s))) return rc;   // P_1 and Z (half-resolution partial sums of ConvLSTM0's R1 taps)
    {   // ConvLSTM_0 on [E0 | up(R1) | h0] (R1 taps via Z), then ConvP_0 -> P0 (this step's prediction)
        if (c->ch[0] == 1) { auto k = l0_lstm_kernel<1>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(128), 0, s, l0); }
        else { auto k = l0_lstm_kernel<3>; LAUNCH_K_PDL(CLS_L0, k, dim3(l0_tiles, B), dim3(128), 0, s, l0); }
        CKL();
        const long long npix = (long long)B * c->h * c->w;
        if (c->ch[0] == 1) { auto k = l0_convp_kernel<1>; LAUNCH_K_PDL(CLS_L0, k, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, l0); }
        else { auto k = l0_convp_kernel<3>; LAUNCH_K_PDL(CLS_L0, k, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, l0); }
        CKL();
    }
    if (!side_ok && !last)
        for (int n = 2; n < 4; ++n)
            if ((rc = conv_p(n, s))) return rc;
    return EIG_OK;
}

// the main stream waits for the side-stream ConvP launches still in flight (before anything overwrites / reads P_2, P_3)
static int join_side(eig_ctx* c, cudaStream_t s) {
#ifndef EIG_EMU
    for (int n = 2; n < 4; ++n)
        if (c->p_pending[n]) { CK(cudaStreamWaitEvent(s, c->ev_p[n], 0)); c->p_pending[n] = false; }
#endif
    (void)c; (void)s;
    return EIG_OK;
}

// Step 0 reads, before anything writes them: h_prev inside the even concat buffers (X[n][0], h0[0]), the cell states and the
// predictions P_n.  The odd buffers are fully written before they are read (h by step 0, E / up(R) by step 1).
static int prednet_reset(eig_ctx* c, int B, cudaStream_t s) {
    ResetArgs ra;
    memset(&ra, 0, sizeof ra);
    int r = 0;
    unsigned long long most = 0;
    auto add = [&](void* p, size_t bytes) { ra.ptr[r] = p; ra.bytes[r] = bytes; if (bytes > most) most = bytes; ++r; };
    for (int n = 0; n < 4; ++n) {
        const size_t px = (size_t)B * c->H[n] * c->W[n];
        if (n == 0) add(c->h0[0], px * c->ch[0] * sizeof(float));
        else if (float* lo = lo_plane(c, n, c->X[n][0])) {
            // split-fp16 storage: the lo plane starts after the hi plane of all `cap` genomes, so each plane has its own
            // prefix of B genomes to clear
            add(c->X[n][0], px * c->ctot[n] * sizeof(h16));
            add(lo, px * c->ctot[n] * sizeof(h16));
        } else add(c->X[n][0], px * c->ctot[n] * sizeof(float));
        add(c->cst[n], px * c->ch[n] * sizeof(float));
        add(c->P[n], px * c->ch[n] * sizeof(float));
    }
    static_assert(RESET_MAX_REGIONS >= 15, "regions");
    unsigned bx = (unsigned)((most / 16 + 255) / 256);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    LAUNCH_K(CLS_ELEMENTWISE, reset_state_kernel, dim3(bx, r), dim3(256), 0, s, ra);
    CKL();
    return EIG_OK;
}

// runs the frame protocol; frame k (0 = prediction #n_in, 1.. = extensions) is quantised to frames_out
// (nullable) and its gray version to gray_dst[k] (nullable entries)
static int prednet_sequence(eig_ctx* c, const float* d_x, int B, int n_in, int n_ext, unsigned char* frames_out,
                            unsigned char* const* gray_dst, cudaStream_t s) {
    int rc;
    if ((rc = join_side(c, s))) return rc;   // a previous sequence may still have ConvP launches on the side stream
    c->seq_n = -1;                           // the stepping state (eig_prednet_forward) does not survive a whole-sequence run
    if ((rc = prednet_reset(c, B, s))) return rc;
    const long long npix = (long long)B * c->h * c->w;
    for (int t = 0; t < n_in + n_ext; ++t) {
        // extension steps are fed the previous, unquantised prediction (call_prednet.py:185,200)
        const float* x = t < n_in ? d_x : c->P[0];
        if ((rc = prednet_step(c, x, B, t, t == n_in + n_ext - 1, s))) return rc;
        if (t >= n_in - 1) {
            const int k = t - (n_in - 1);
            unsigned char* fo = frames_out ? frames_out + (size_t)k * npix * c->c_dim : nullptr;
            unsigned char* go = gray_dst ? gray_dst[k] : nullptr;
            if (fo || go) {
                LAUNCH_K(CLS_ELEMENTWISE, quantize_gray_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s,
                           (const float*)c->P[0], fo, go, npix, c->c_dim);
                CKL();
            }
        }
    }
    return join_side(c, s);
}

static int check_ready(eig_ctx* c, int n, bool need_w, bool need_grid, bool render_call = false) {
    if (!c) return fail(EIG_E_INVALID, "null ctx");
    if (c->render_only && !render_call) return fail(EIG_E_STATE, "this context was made by eig_create_render: it only renders");
    if (n <= 0) return fail(EIG_E_INVALID, "n must be positive");
    if (n > c->cap) return fail(EIG_E_CAPACITY, "population larger than max_genomes given to eig_create");
    if (need_w && !c->have_w) return fail(EIG_E_STATE, "PredNet weights not loaded (eig_load_weights)");
    if (need_grid && !c->have_grid) return fail(EIG_E_STATE, "grid planes not set (eig_set_grid)");
    return EIG_OK;
}

extern "C" int eig_cppn_render(eig_ctx* c, const void* d_blob, const int64_t* d_offsets, int n, int max_slots,
                               int max_blob_bytes, int mode, double bg, uint8_t* d_img, float* d_x, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, false, true, true))) return rc;
    if (!d_blob || !d_offsets || !d_img) return fail(EIG_E_INVALID, "eig_cppn_render: null pointer");
    if (mode < 0 || mode > 2) return fail(EIG_E_INVALID, "eig_cppn_render: mode must be 0, 1 or 2");
    CK(cudaSetDevice(c->device));
    RenderArgs a;
    a.blob = (const unsigned char*)d_blob; a.offsets = (const long long*)d_offsets;
    a.xmat = c->xmat; a.ymat = c->ymat; a.npix = c->w * c->h; a.c_dim = c->c_dim; a.mode = mode; a.bg = bg;
    a.img = d_img; a.x = d_x;
    a.max_blob_bytes = (max_blob_bytes + 15) & ~15;
    int nt = 128;
    size_t smem = (size_t)a.max_blob_bytes + (size_t)max_slots * nt * sizeof(double);
    while (smem > 200 * 1024 && nt > 32) { nt >>= 1; smem = (size_t)a.max_blob_bytes + (size_t)max_slots * nt * sizeof(double); }
    if (smem > 200 * 1024) return fail(EIG_E_CAPACITY, "eig_cppn_render: genome program too large for shared memory");
    auto k = cppn_render_kernel;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH_K(CLS_RENDER, k, dim3((a.npix + nt - 1) / nt, n), dim3(nt), smem, TO_STREAM(stream), a);
    CKL();
    return EIG_OK;
}

extern "C" int eig_prednet_run(eig_ctx* c, const float* d_x, int n, int n_in, int n_ext, uint8_t* d_frames, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, true, false))) return rc;
    if (!d_x || !d_frames || n_in < 1 || n_ext < 0) return fail(EIG_E_INVALID, "eig_prednet_run: bad argument");
    CK(cudaSetDevice(c->device));
    return prednet_sequence(c, d_x, n, n_in, n_ext, d_frames, nullptr, TO_STREAM(stream));
}

// stateful single steps: sequences of distinct frames, every prediction readable (test_image_list, call_prednet.py:129-205)
extern "C" int eig_prednet_reset(eig_ctx* c, int n, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, true, false))) return rc;
    CK(cudaSetDevice(c->device));
    cudaStream_t s = TO_STREAM(stream);
    if ((rc = join_side(c, s))) return rc;
    c->seq_t = 0;
    c->seq_n = n;
    return prednet_reset(c, n, s);
}

extern "C" int eig_prednet_forward(eig_ctx* c, const float* d_x, int n, float* d_pred, uint8_t* d_frame, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, true, false))) return rc;
    if (!d_x) return fail(EIG_E_INVALID, "eig_prednet_forward: null input");
    if (c->seq_n != n) return fail(EIG_E_STATE, "eig_prednet_forward: call eig_prednet_reset with the same n first");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = TO_STREAM(stream);
    if ((rc = prednet_step(c, d_x, n, c->seq_t, false, s))) return rc;
    ++c->seq_t;
    const long long npix = (long long)n * c->h * c->w;
    if (d_pred) CK(cudaMemcpyAsync(d_pred, c->P[0], sizeof(float) * npix * c->c_dim, cudaMemcpyDeviceToDevice, s));
    if (d_frame) {
        LAUNCH_K(CLS_ELEMENTWISE, quantize_gray_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s,
                   (const float*)c->P[0], d_frame, (unsigned char*)nullptr, npix, c->c_dim);
        CKL();
    }
    return EIG_OK;
}

// corners + LK + vector rows from the gray level-0 images already in c->gray[0] ([0,B) frame 1, [B,2B) frame 2)
static int flow_from_gray(eig_ctx* c, int B, cudaStream_t s) {
    const int H = c->h, W = c->w;
    for (int l = 1; l < c->n_levels; ++l) {
        // frame-2 images of every level sit right after the B frame-1 images actually in use
        const long long tot = 2LL * B * c->lh[l] * c->lwid[l];
        LAUNCH_K(CLS_FLOW, pyr_down_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, s, (const unsigned char*)c->gray[l - 1],
                   c->gray[l], c->lh[l - 1], c->lwid[l - 1], c->lh[l], c->lwid[l], 2 * B);
        CKL();
    }
    for (int l = 0; l < c->n_levels; ++l) {
        const long long tot = (long long)B * c->lh[l] * c->lwid[l];
        LAUNCH_K(CLS_FLOW, scharr_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, s, (const unsigned char*)c->gray[l],
                   c->deriv[l], c->lh[l], c->lwid[l], B);
        CKL();
    }
    CK(cudaMemsetAsync(c->eigmax, 0x80, sizeof(int) * B, s));  // 0x80808080: below every real key
    EigArgs ea; ea.gray = c->gray[0]; ea.eig = c->eigmap; ea.eig_max_key = c->eigmax; ea.H = H; ea.W = W;
    LAUNCH_K(CLS_FLOW, min_eig_kernel, dim3(((W + 31) / 32) * ((H + 15) / 16), B), dim3(256), 0, s, ea);
    CKL();
    CornerArgs ca; ca.eig = c->eigmap; ca.eig_max_key = c->eigmax; ca.cand = c->cand; ca.corners = c->corners;
    ca.ncorners = c->ncorners; ca.H = H; ca.W = W;
    LAUNCH_K(CLS_FLOW, corner_select_kernel, dim3(B), dim3(1024), c->corner_smem, s, ca);
    CKL();
    LkArgs la;
    memset(&la, 0, sizeof la);
    for (int l = 0; l < c->n_levels; ++l) {
        la.img1[l] = c->gray[l];
        la.img2[l] = c->gray[l] + (size_t)B * c->lh[l] * c->lwid[l];
        la.deriv1[l] = c->deriv[l];
        la.lh[l] = c->lh[l]; la.lw[l] = c->lwid[l];
    }
    la.n_levels = c->n_levels; la.corners = c->corners; la.ncorners = c->ncorners; la.next_pts = c->next_pts;
    la.status = c->status; la.B = B;
    LAUNCH_K(CLS_FLOW, lk_track_kernel, dim3(B * FLOW_MAX_CORNERS), dim3(LK_THREADS), 0, s, la);
    CKL();
    LAUNCH_K(CLS_FLOW, collect_vectors_kernel, dim3((B + 3) / 4), dim3(128), 0, s, (const float*)c->corners, (const int*)c->ncorners,
               (const float*)c->next_pts, (const unsigned char*)c->status, c->vectors, c->nvec, B);
    CKL();
    return EIG_OK;
}

extern "C" int eig_flow(eig_ctx* c, const uint8_t* d_img1, const uint8_t* d_img2, int n, float* d_corners, int* d_ncorners,
                        float* d_vectors, int* d_nvec, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, false, false))) return rc;
    if (!d_img1 || !d_img2 || !d_vectors || !d_nvec) return fail(EIG_E_INVALID, "eig_flow: null pointer");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = TO_STREAM(stream);
    const long long npix = (long long)n * c->h * c->w;
    LAUNCH_K(CLS_ELEMENTWISE, gray_u8_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, d_img1, c->gray[0], npix, c->c_dim);
    LAUNCH_K(CLS_ELEMENTWISE, gray_u8_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, d_img2, c->gray[0] + npix, npix, c->c_dim);
    CKL();
    if ((rc = flow_from_gray(c, n, s))) return rc;
    CK(cudaMemcpyAsync(d_vectors, c->vectors, sizeof(float) * n * FLOW_MAX_CORNERS * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(d_nvec, c->nvec, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    if (d_corners) CK(cudaMemcpyAsync(d_corners, c->corners, sizeof(float) * n * FLOW_MAX_CORNERS * 2, cudaMemcpyDeviceToDevice, s));
    if (d_ncorners) CK(cudaMemcpyAsync(d_ncorners, c->ncorners, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    return EIG_OK;
}

static int score_launch(eig_ctx* c, const float* vec, const int* nvec, int n, int structure, double* fit, cudaStream_t s) {
    if (structure < 0 || structure > 3) return fail(EIG_E_INVALID, "unknown structure id");
    ScoreArgs sa; sa.vectors = vec; sa.nvec = nvec; sa.fitness = fit; sa.B = n; sa.structure = structure; sa.w = c->w; sa.h = c->h;
    sa.status = c->fitness + c->cap;   // sticky range flag of this context, read and cleared by eig_eval_host / eig_range_check
    LAUNCH_K(CLS_SCORE, score_kernel, dim3(n), dim3(32), 0, s, sa);
    CKL();
    return EIG_OK;
}

extern "C" int eig_score(eig_ctx* c, const float* d_vectors, const int* d_nvec, int n, int structure, double* d_fitness, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, false, false))) return rc;
    if (!d_vectors || !d_nvec || !d_fitness) return fail(EIG_E_INVALID, "eig_score: null pointer");
    CK(cudaSetDevice(c->device));
    return score_launch(c, d_vectors, d_nvec, n, structure, d_fitness, TO_STREAM(stream));
}

// everything after the render: frame protocol, flow, score (all launch arguments depend only on n, the modes and ctx buffers)
static int eval_after_render(eig_ctx* c, int n, int structure, int pair_mode, double* d_fitness, cudaStream_t s) {
    int rc;
    const long long npix = (long long)n * c->h * c->w;
    unsigned char* g1 = c->gray[0];
    unsigned char* g2 = c->gray[0] + npix;
    unsigned char* gd[3];
    int n_ext;
    if (pair_mode == EIG_PAIR_POPULATION) {
        gd[0] = g1; gd[1] = g2; gd[2] = nullptr; n_ext = 1;   // the reference's 22nd forward is dead work
    } else {
        gd[0] = nullptr; gd[1] = nullptr; gd[2] = g2; n_ext = 2;
        LAUNCH_K(CLS_ELEMENTWISE, gray_u8_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, (const unsigned char*)c->img, g1, npix, c->c_dim);
        CKL();
    }
    if ((rc = prednet_sequence(c, c->x_in, n, 20, n_ext, c->frames, gd, s))) return rc;
    if ((rc = flow_from_gray(c, n, s))) return rc;
    return score_launch(c, c->vectors, c->nvec, n, structure, d_fitness, s);
}

// the evaluation writes the context's own fitness vector (so a captured graph does not depend on the caller's pointer)
static int copy_fitness(eig_ctx* c, int n, double* d_fitness, cudaStream_t s) {
    if (d_fitness != c->fitness)
        CK(cudaMemcpyAsync(d_fitness, c->fitness, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    return EIG_OK;
}

extern "C" int eig_eval(eig_ctx* c, const void* d_blob, const int64_t* d_offsets, int n, int max_slots, int max_blob_bytes,
                        int structure, int render_mode, int pair_mode, double* d_fitness, void* stream) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, true, true))) return rc;
    if (!d_fitness) return fail(EIG_E_INVALID, "eig_eval: null fitness pointer");
    if (pair_mode != EIG_PAIR_POPULATION && pair_mode != EIG_PAIR_SINGLE_IMAGE) return fail(EIG_E_INVALID, "unknown pair mode");
    c->seq_n = -1;   // the evaluation (possibly a graph replay) overwrites the recurrent state of eig_prednet_forward
    cudaStream_t s = TO_STREAM(stream);
    if ((rc = eig_cppn_render(c, d_blob, d_offsets, n, max_slots, max_blob_bytes, render_mode, 1.0, c->img, c->x_in, stream))) return rc;
#ifndef EIG_EMU
    if (c->use_graphs && !prof_on()) {
        auto key = std::make_tuple(n, structure, pair_mode, c->conv_mode);
        eig_ctx::GraphEntry& ge = c->graphs[key];
        if (ge.exec) {
            CK(cudaGraphLaunch((cudaGraphExec_t)ge.exec, s));
            launch_counter().n += ge.launches;
            return copy_fitness(c, n, d_fitness, s);
        }
        if (ge.seen++ >= 1) {   // second use of this key: capture, instantiate, replay from now on
            const long long before = launch_counter().n;
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                rc = eval_after_render(c, n, structure, pair_mode, c->fitness, s);
                cudaGraph_t g = nullptr;
                const cudaError_t ee = cudaStreamEndCapture(s, &g);
                cudaGraphExec_t exec = nullptr;
                if (rc == EIG_OK && ee == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) {
                    cudaGraphDestroy(g);
                    ge.exec = exec;
                    ge.launches = launch_counter().n - before;
                    CK(cudaGraphLaunch(exec, s));
                    return copy_fitness(c, n, d_fitness, s);
                }
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                launch_counter().n = before;
            }
            cudaGetLastError();      // a refused capture (legacy stream, caller already capturing) must not surface as a launch error below
            c->use_graphs = false;   // capture is not possible here: run directly
        }
    }
#endif
    if ((rc = eval_after_render(c, n, structure, pair_mode, c->fitness, s))) return rc;
    return copy_fitness(c, n, d_fitness, s);
}

extern "C" int eig_range_check(eig_ctx* c, void* stream) {
    Scope sc(c);
    if (!c) return fail(EIG_E_INVALID, "null ctx");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = TO_STREAM(stream);
    double status = 0.0;
    CK(cudaMemcpyAsync(&status, c->fitness + c->cap, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemsetAsync(c->fitness + c->cap, 0, sizeof(double), s));
    CK(cudaStreamSynchronize(s));
    if (status != 0.0)
        return fail(EIG_E_RANGE, "an activation left the split-fp16 range (|v| >= 4094) or became NaN in tensor-core mode; "
                                 "this weight file needs conv_mode EIG_CONV_SIMT");
    return EIG_OK;
}

extern "C" int eig_eval_host(eig_ctx* c, const void* h_blob, const int64_t* h_offsets, int n, int max_slots, int structure,
                             int render_mode, int pair_mode, double* h_fitness) {
    Scope sc(c);
    int rc;
    if ((rc = check_ready(c, n, true, true))) return rc;
    if (!h_blob || !h_offsets || !h_fitness) return fail(EIG_E_INVALID, "eig_eval_host: null pointer");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)h_offsets[n];
    int max_blob = 0;
    for (int i = 0; i < n; ++i) { const int b = (int)(h_offsets[i + 1] - h_offsets[i]); if (b > max_blob) max_blob = b; }
    if (bytes > c->d_blob_cap) {
        if (c->d_blob) CK(cudaFree(c->d_blob));
        c->d_blob = nullptr;
        c->d_blob_cap = bytes * 2 + 4096;
        CK(cudaMalloc(&c->d_blob, c->d_blob_cap));
    }
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->d_blob, h_blob, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->d_off, h_offsets, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    if ((rc = eig_eval(c, c->d_blob, (const int64_t*)c->d_off, n, max_slots, max_blob, structure, render_mode, pair_mode,
                       c->fitness, (void*)(intptr_t)s)))
        return rc;
    CK(cudaMemcpyAsync(h_fitness, c->fitness, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    return eig_range_check(c, (void*)(intptr_t)s);
}

extern "C" int eig_debug_buffers(eig_ctx* c, uint8_t** d_img, uint8_t** d_frames, float** d_vectors, int** d_nvec,
                                 float** d_corners, int** d_ncorners) {
    Scope sc(c);
    if (!c) return fail(EIG_E_INVALID, "null ctx");
    if (d_img) *d_img = c->img;
    if (d_frames) *d_frames = c->frames;
    if (d_vectors) *d_vectors = c->vectors;
    if (d_nvec) *d_nvec = c->nvec;
    if (d_corners) *d_corners = c->corners;
    if (d_ncorners) *d_ncorners = c->ncorners;
    return EIG_OK;
}

extern "C" int eig_memcpy_d2h(void* h_dst, const void* d_src, int64_t bytes) {
    if (!h_dst || !d_src || bytes < 0) return fail(EIG_E_INVALID, "eig_memcpy_d2h: bad argument");
    CK(cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return EIG_OK;
}

extern "C" int eig_profile_begin(eig_ctx* c) {
    if (!c) return fail(EIG_E_INVALID, "null ctx");
    Scope sc(c);
    Profiler& pr = c->prof;
    for (cudaEvent_t e : pr.ev) cudaEventDestroy(e);
    pr.ev.clear(); pr.cls.clear();
    pr.on = true;
    return EIG_OK;
}

extern "C" int eig_profile_end(eig_ctx* c, double* ms_per_class, int64_t* launches_per_class) {
    if (!c || !ms_per_class || !launches_per_class) return fail(EIG_E_INVALID, "eig_profile_end: null pointer");
    Scope sc(c);
    Profiler& pr = c->prof;
    pr.on = false;
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < CLS_COUNT; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
    for (size_t i = 0; i < pr.cls.size(); ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pr.ev[2 * i], pr.ev[2 * i + 1]));
        ms_per_class[pr.cls[i]] += ms;
        launches_per_class[pr.cls[i]] += 1;
    }
    for (cudaEvent_t e : pr.ev) cudaEventDestroy(e);
    pr.ev.clear(); pr.cls.clear();
    return EIG_OK;
}
