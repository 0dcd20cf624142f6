/* libeig.so - C ABI of the B200-native EIGen fitness engine.
 *
 * The reference (LanaSina/evolutionary_illusion_generator) is pure Python and has no FFI; its hot path is the
 * body of `get_fitnesses_neat` (/root/reference/generate_illusion.py:478-673).  This header is the seam a
 * maintainer binds with ctypes: every entry point names the reference function it
 * replaces.  Conventions: plain pointers and sizes only; return 0 on success, a negative EIG_E_* code on
 * failure (never throws, never exits); `eig_last_error` gives the message.  One context per process per GPU,
 * not thread-safe.  Pointers named d_* are device pointers, h_* host pointers.  All work is enqueued on the
 * `stream` argument (a cudaStream_t passed as void*; NULL = default stream) and is asynchronous unless the
 * function name ends in _host.
 */
#ifndef EIG_H
#define EIG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eig_ctx eig_ctx;

enum {
    EIG_OK = 0,
    EIG_E_INVALID = -1,   /* bad argument */
    EIG_E_CUDA = -2,      /* CUDA runtime / driver error */
    EIG_E_STATE = -3,     /* weights or grid not loaded yet */
    EIG_E_CAPACITY = -4,  /* population larger than max_genomes, genome too large for shared memory */
    EIG_E_NODEVICE = -5,  /* no CUDA device: there is no CPU fallback */
    EIG_E_RANGE = -6      /* tensor-core mode: an activation left the split-fp16 range (|v| >= 4094) or became NaN */
};

/* structure ids = StructureType, generate_illusion.py:25-29 */
enum { EIG_BANDS = 0, EIG_CIRCLES = 1, EIG_FREE = 2, EIG_CIRCLES_FREE = 3 };
/* frame pairing: population path (generate_illusion.py:543-546: prediction #20 vs extension #1) or the
 * single-image path (fitness_calculator.py:493-498: input image vs extension #2) */
enum { EIG_PAIR_POPULATION = 0, EIG_PAIR_SINGLE_IMAGE = 1 };
/* convolution engine for PredNet layers 1..3: exact-fp32 SIMT kernels, or tcgen05 tensor cores (CTA pairs, 3-pass split fp16 with fp32 accumulation) */
enum { EIG_CONV_SIMT = 0, EIG_CONV_TC = 1 };

/* message of the most recent failure on the calling thread (any context; also failures of eig_create) */
const char* eig_last_error(void);
/* message of the most recent failure of THIS context (two contexts in one process do not overwrite each other) */
const char* eig_error(const eig_ctx* ctx);
int eig_version(void);
/* number of CUDA kernels this library has launched so far (bench.py's `gpu_launches`) */
int64_t eig_launch_count(void);

/* Replaces `net.PredNet(w, h, channels)` + per-call buffer setup (call_prednet.py:209-231).
 * channels[4] = PredNet channels per layer, c_dim = channels[0] (1 or 3); w, h divisible by 8. */
int eig_create(eig_ctx** out, int device, int w, int h, int c_dim, const int channels[4], int max_genomes);
/* A context for `get_image_from_cppn` alone (generate_illusion.py:372-460; the 800x800 `enhanced.png` mosaic, 664-671):
 * only eig_set_grid and eig_cppn_render work on it (everything else returns EIG_E_STATE); any w, h > 0. */
int eig_create_render(eig_ctx** out, int device, int w, int h, int c_dim, int max_genomes);
void eig_destroy(eig_ctx* ctx);

/* conv_mode: EIG_CONV_SIMT / EIG_CONV_TC.  Returns EIG_E_INVALID if the mode is not compiled in. */
int eig_set_conv_mode(eig_ctx* ctx, int conv_mode);

/* Tuning / diagnostic knobs (no reference counterpart).  Keys:
 *   "passes.all" | "passes.A" | "passes.P" | "passes.L" | "passes.<A|P|L><1..3>" : which of the three split-fp16 MMA products
 *        a convolution issues per k-step (bit 0 a_lo*w_hi, bit 1 a_hi*w_lo, bit 2 a_hi*w_hi; default 7 = all, the setting
 *        the parity tests pin; profiles/r2/pass_ablation.md has the measured cost of every cheaper mix);
 *   "precision" 0/1/2 : preset of those masks (0 exact, the default; 1 single product in layers 2+3; 2 single product everywhere);
 *   "early_until" = T, "early_mask" = m : PredNet steps t < T use (mask & m);
 *   "fold" -1/0/1 : folded form of the ConvLSTM taps over up-sampled states (auto by problem size / off / on);
 *   "skip_zero_state" 0/1 : step 0 skips the K blocks that only hold the zero state (bit-identical, default on);
 *   "simt_reverse_taps" 0/1 : exact-fp32 kernel sums the taps in reverse order (the ablation's yardstick);
 *   "graphs" 0/1 : CUDA-graph replay of everything after the render; "overlap" 0/1 : ConvP2/3 on the side stream.
 * Synchronises the device and drops captured graphs. */
int eig_set_option(eig_ctx* ctx, const char* key, int value);

/* Replaces `serializers.load_npz(initmodel, model)` (call_prednet.py:231).  names[i] use the Chainer npz keys
 * ("predictor/ConvLSTM2/x_i0/W", ...); host_ptrs[i] is contiguous fp32; shapes is n_tensors x 4 (unused
 * trailing dims = 1).  Repacks into the kernels' layouts once. */
int eig_load_weights(eig_ctx* ctx, int n_tensors, const char* const* names, const float* const* host_ptrs,
                     const int64_t* shapes);

/* Replaces the `create_grid` hand-off (generate_illusion.py:501): h*w fp64 planes, x_mat == -1 = background. */
int eig_set_grid(eig_ctx* ctx, const double* h_x_mat, const double* h_y_mat);

/* Replaces `get_image_from_cppn` for n genomes (generate_illusion.py:372-460).  d_blob/d_offsets: flattened
 * programs (genome.py) and n+1 byte offsets; max_slots = largest program slot count.
 * mode: 0 gradient, 1 gray without gradient (np.round), 2 colour palette (gradient=0).
 * d_img: [n][h][w][c_dim] uint8.  d_x (nullable): [n][h][w][c_dim] fp32 = read_image() of that picture. */
int eig_cppn_render(eig_ctx* ctx, const void* d_blob, const int64_t* d_offsets, int n, int max_slots,
                    int max_blob_bytes, int mode, double bg, uint8_t* d_img, float* d_x, void* stream);

/* Replaces `test_prednet(... extension_start=20, extension_duration=2)` for n independent sequences
 * (call_prednet.py:129-205): reset, n_input_steps forwards on d_x, then n_ext self-fed forwards.
 * d_frames: [(n_ext+1)][n][h][w][c_dim] uint8 = prediction #n_input_steps, extension #1.. (write_image). */
int eig_prednet_run(eig_ctx* ctx, const float* d_x, int n, int n_input_steps, int n_ext, uint8_t* d_frames,
                    void* stream);

/* The same network stepped frame by frame, for sequences of DISTINCT frames and when every prediction is wanted
 * (`test_image_list`, call_prednet.py:129-205): eig_prednet_reset = `prednet.reset_state()` (net.py:159-164) for n
 * parallel sequences; eig_prednet_forward = one `model(x, y)` call (call_prednet.py:155): d_x [n][h][w][c] fp32 in,
 * d_pred (nullable) the unquantised prediction `model.y.data` that the extension steps feed back (line 185/200),
 * d_frame (nullable) what `write_image` stores (uint8 truncation of P0*255). */
int eig_prednet_reset(eig_ctx* ctx, int n, void* stream);
int eig_prednet_forward(eig_ctx* ctx, const float* d_x, int n, float* d_pred, uint8_t* d_frame, void* stream);

/* Replaces `lucas_kanade(file1, file2)` for n image pairs (optical_flow/optical_flow.py:40-89).
 * d_img1/d_img2: [n][h][w][c_dim] uint8 (RGB order).  Outputs (all nullable except d_vectors/d_nvec):
 * d_corners [n][100][2] fp32, d_ncorners [n], d_vectors [n][100][4] fp32 rows (x, y, dx, dy), d_nvec [n]. */
int eig_flow(eig_ctx* ctx, const uint8_t* d_img1, const uint8_t* d_img2, int n, float* d_corners,
             int* d_ncorners, float* d_vectors, int* d_nvec, void* stream);

/* Replaces the scoring branches (generate_illusion.py:557-616 / fitness_calculator.calculate_fitness). */
int eig_score(eig_ctx* ctx, const float* d_vectors, const int* d_nvec, int n, int structure, double* d_fitness,
              void* stream);

/* The whole hot path for n genomes: render -> PredNet (20 + n_ext) -> flow -> score. d_fitness: [n] fp64. */
int eig_eval(eig_ctx* ctx, const void* d_blob, const int64_t* d_offsets, int n, int max_slots,
             int max_blob_bytes, int structure, int render_mode, int pair_mode, double* d_fitness, void* stream);

/* Same, from HOST buffers (pinned or pageable): H2D of the genome blob, eval, D2H of the fitness vector,
 * stream synchronised on return.  This is what `get_fitnesses_neat` calls. */
int eig_eval_host(eig_ctx* ctx, const void* h_blob, const int64_t* h_offsets, int n, int max_slots,
                  int structure, int render_mode, int pair_mode, double* h_fitness);

/* Resident-path companion of eig_eval_host's range report: synchronises `stream`, returns EIG_E_RANGE if any eig_eval
 * since the last check (or the last eig_eval_host) saw an activation leave the split-fp16 range, and clears the flag.
 * The reference has no counterpart (its Chainer convolutions are plain fp32, net.py:45-62). */
int eig_range_check(eig_ctx* ctx, void* stream);

/* Debug / test taps: device pointers to the context's internal buffers of the last eig_eval
 * (rendered images [n][h][w][c] u8, frames [3][n][h][w][c] u8, vectors, nvec, corners, ncorners). */
int eig_debug_buffers(eig_ctx* ctx, uint8_t** d_img, uint8_t** d_frames, float** d_vectors, int** d_nvec,
                      float** d_corners, int** d_ncorners);

/* Synchronous device -> host copy of `bytes` bytes (tests read the debug taps with it; no torch type needed). */
int eig_memcpy_d2h(void* h_dst, const void* d_src, int64_t bytes);

/* Per-kernel-class device timing for bench.py's roofline pass: between begin and end every kernel launch is
 * bracketed by CUDA events on its stream.  Classes: 0 render, 1 conv SIMT (layers 1-3), 2 conv tcgen05, 3 element-wise,
 * 4 flow, 5 score, 6 fused layer-0 kernels (arrays of 8).  eig_profile_end synchronises the device. */
int eig_profile_begin(eig_ctx* ctx);
int eig_profile_end(eig_ctx* ctx, double* ms_per_class, int64_t* launches_per_class);

#ifdef __cplusplus
}
#endif
#endif /* EIG_H */
