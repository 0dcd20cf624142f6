"""Engine: one libeig context per process per GPU, with PyTorch tensors as the device buffers.

Host code only prepares inputs (flattened genomes, grid planes, weight arrays) and launches; all arithmetic of
the hot path runs in the CUDA kernels behind the C ABI (include/eig.h).  Without a CUDA device and the built
`libeig.so` construction fails - there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, genome as G, grid as grid_mod, weights as weights_mod

MAX_CORNERS = 100
RENDER_GRADIENT, RENDER_GRAY_ROUND, RENDER_PALETTE = 0, 1, 2


def render_mode_for(c_dim, gradient):
    """generate_illusion.py:391-458: colour takes the gradient branch only for `gradient == 1` and the palette branch
    otherwise (line 391/405); gray rounds only for `gradient == 0` (line 450) and is a plain gradient for any other value."""
    if c_dim > 1:
        return RENDER_GRADIENT if gradient == 1 else RENDER_PALETTE
    return RENDER_GRAY_ROUND if gradient == 0 else RENDER_GRADIENT


class Engine:
    def __init__(self, w, h, channels, max_genomes, device=None, lib=None, tensor_device=None, render_only=False):
        """render_only: a context for the CPPN stage alone (`eig_create_render`): `channels` only supplies c_dim, any
        image size works, and only set_grid / render are available."""
        self.w, self.h = int(w), int(h)
        self.channels = [int(c) for c in channels]
        self.c_dim = self.channels[0]
        self.max_genomes = int(max_genomes)
        self.render_only = bool(render_only)
        if lib is None:
            lib = _lib.get_library()
            if not torch.cuda.is_available():
                raise RuntimeError("evolutionary_illusion_generator_b200 needs a CUDA device (B200); "
                                   "there is no CPU fallback")
            dev_index = torch.cuda.current_device() if device is None else int(device)
            self.tdev = torch.device("cuda", dev_index)
        else:  # explicit library: the test-only emulator binds host memory
            dev_index = 0
            self.tdev = torch.device(tensor_device or "cpu")
        self.lib = lib
        self.dev_index = dev_index
        ctx = C.c_void_p()
        ch = (C.c_int * 4)(*self.channels)
        if self.render_only:
            lib.check(lib.eig_create_render(C.byref(ctx), dev_index, self.w, self.h, self.c_dim, self.max_genomes))
        else:
            lib.check(lib.eig_create(C.byref(ctx), dev_index, self.w, self.h, self.c_dim, ch, self.max_genomes))
        self.ctx = ctx
        self._grid_key = None
        self.structure = None
        # whole-path evaluations run on a high-priority stream: the library puts ConvP_2/3 on a lowest-priority side
        # stream, and the block scheduler then gives the critical-path kernels the SMs first
        self._hp_stream = torch.cuda.Stream(device=self.tdev, priority=-1) if self.tdev.type == "cuda" else None
        self._up_stream = torch.cuda.Stream(device=self.tdev) if self.tdev.type == "cuda" else None
        self._fit_buf = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.eig_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    def _stream(self):
        if self.tdev.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)
        return C.c_void_p(0)

    def sync(self):
        if self.tdev.type == "cuda":
            torch.cuda.synchronize(self.tdev)

    def set_conv_mode(self, mode):
        self.lib.check(self.lib.eig_set_conv_mode(self.ctx, int(mode)))

    def set_option(self, key, value):
        """Tuning / diagnostic knob of the library (include/eig.h: eig_set_option)."""
        self.lib.check(self.lib.eig_set_option(self.ctx, key.encode(), int(value)))

    def load_weights(self, weights):
        """weights: path to a Chainer npz (`serializers.save_npz` layout) or a dict name -> ndarray."""
        if isinstance(weights, str):
            weights = weights_mod.load_npz(weights)
        weights_mod.check_weights(weights, self.w, self.h, self.channels)
        names = sorted(weights)
        arrs = [np.ascontiguousarray(weights[k], dtype=np.float32) for k in names]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        shp = np.ones((n, 4), dtype=np.int64)
        for i, a in enumerate(arrs):
            shp[i, :a.ndim] = a.shape
        self.lib.check(self.lib.eig_load_weights(self.ctx, n, c_names, c_ptrs,
                                                 shp.ctypes.data_as(C.POINTER(C.c_int64))))

    def set_grid(self, structure=None, grid=None, scaling=10):
        """Upload the CPPN input planes of `structure` (or explicit planes)."""
        if grid is None:
            key = (int(structure), scaling)
            if key == self._grid_key:
                return
            grid = grid_mod.create_grid(structure, self.w, self.h, scaling)
            self._grid_key = key
        else:
            self._grid_key = None
        x = np.ascontiguousarray(np.asarray(grid["x_mat"], dtype=np.float64).reshape(self.h, self.w))
        y = np.ascontiguousarray(np.asarray(grid["y_mat"], dtype=np.float64).reshape(self.h, self.w))
        self.lib.check(self.lib.eig_set_grid(self.ctx, x.ctypes.data, y.ctypes.data))
        self.structure = structure

    # ------------------------------------------------------------------ helpers
    def _to_dev(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.tdev, non_blocking=False) if self.tdev.type == "cuda" else t.clone()

    def upload_programs(self, programs):
        """[FlatProgram] -> (d_blob, d_offsets, n, max_slots, max_blob_bytes) resident on the device."""
        blob, offsets, max_slots = G.pack_population(programs)
        max_blob = int(np.diff(offsets).max()) if len(programs) else 0
        return (self._to_dev(blob), self._to_dev(offsets), len(programs), max_slots, max_blob)

    # ------------------------------------------------------------------ stages
    def render(self, programs, mode=RENDER_GRADIENT, bg=1.0):
        d_blob, d_off, n, max_slots, max_blob = self.upload_programs(programs)
        img = torch.empty((n, self.h, self.w, self.c_dim), dtype=torch.uint8, device=self.tdev)
        x = torch.empty((n, self.h, self.w, self.c_dim), dtype=torch.float32, device=self.tdev)
        self.lib.check(self.lib.eig_cppn_render(self.ctx, d_blob.data_ptr(), d_off.data_ptr(), n, max_slots, max_blob,
                                                int(mode), float(bg), img.data_ptr(), x.data_ptr(), self._stream()))
        self.sync()
        return img, x

    def prednet(self, x, n_input_steps=20, n_ext=2):
        """x: (n,h,w,c) float32 device tensor -> uint8 frames (n_ext+1, n, h, w, c)."""
        n = x.shape[0]
        x = x.contiguous()
        frames = torch.empty((n_ext + 1, n, self.h, self.w, self.c_dim), dtype=torch.uint8, device=self.tdev)
        self.lib.check(self.lib.eig_prednet_run(self.ctx, x.data_ptr(), n, n_input_steps, n_ext, frames.data_ptr(),
                                                self._stream()))
        self.sync()
        return frames

    def prednet_reset(self, n=1):
        """`prednet.reset_state()` for n parallel sequences (stateful stepping, see prednet_forward)."""
        self._seq_n = int(n)
        self.lib.check(self.lib.eig_prednet_reset(self.ctx, self._seq_n, self._stream()))

    def prednet_forward(self, x):
        """One forward step on x (n,h,w,c) float32 with the state carried over from the previous call:
        -> (unquantised prediction float32 (n,h,w,c), uint8 frame as `write_image` stores it)."""
        x = x.contiguous()
        n = x.shape[0]
        pred = torch.empty((n, self.h, self.w, self.c_dim), dtype=torch.float32, device=self.tdev)
        frame = torch.empty((n, self.h, self.w, self.c_dim), dtype=torch.uint8, device=self.tdev)
        self.lib.check(self.lib.eig_prednet_forward(self.ctx, x.data_ptr(), n, pred.data_ptr(), frame.data_ptr(),
                                                    self._stream()))
        self.sync()
        return pred, frame

    def flow(self, img1, img2):
        """img1/img2: (n,h,w,c) uint8 device tensors -> (corners, ncorners, vectors, nvec) device tensors."""
        n = img1.shape[0]
        img1, img2 = img1.contiguous(), img2.contiguous()
        corners = torch.zeros((n, MAX_CORNERS, 2), dtype=torch.float32, device=self.tdev)
        ncorners = torch.zeros((n,), dtype=torch.int32, device=self.tdev)
        vectors = torch.zeros((n, MAX_CORNERS, 4), dtype=torch.float32, device=self.tdev)
        nvec = torch.zeros((n,), dtype=torch.int32, device=self.tdev)
        self.lib.check(self.lib.eig_flow(self.ctx, img1.data_ptr(), img2.data_ptr(), n, corners.data_ptr(),
                                         ncorners.data_ptr(), vectors.data_ptr(), nvec.data_ptr(), self._stream()))
        self.sync()
        return corners, ncorners, vectors, nvec

    def score(self, vectors, nvec, structure):
        n = vectors.shape[0]
        fit = torch.zeros((n,), dtype=torch.float64, device=self.tdev)
        self.lib.check(self.lib.eig_score(self.ctx, vectors.contiguous().data_ptr(), nvec.contiguous().data_ptr(), n,
                                          int(structure), fit.data_ptr(), self._stream()))
        self.sync()
        return fit

    # ------------------------------------------------------------------ whole path
    def evaluate_resident(self, resident, structure, render_mode=RENDER_GRADIENT, pair_mode=_lib.PAIR_POPULATION,
                          out=None):
        """Genome programs already on the device (upload_programs) -> device fp64 fitness tensor.  Asynchronous."""
        d_blob, d_off, n, max_slots, max_blob = resident
        if out is None:
            out = torch.empty((n,), dtype=torch.float64, device=self.tdev)
        if self._hp_stream is None:
            self.lib.check(self.lib.eig_eval(self.ctx, d_blob.data_ptr(), d_off.data_ptr(), n, max_slots, max_blob,
                                             int(structure), int(render_mode), int(pair_mode), out.data_ptr(),
                                             self._stream()))
            return out
        cur = torch.cuda.current_stream(self.tdev)
        self._hp_stream.wait_stream(cur)                 # inputs / `out` may have been produced on the caller's stream
        self.lib.check(self.lib.eig_eval(self.ctx, d_blob.data_ptr(), d_off.data_ptr(), n, max_slots, max_blob,
                                         int(structure), int(render_mode), int(pair_mode), out.data_ptr(),
                                         C.c_void_p(self._hp_stream.cuda_stream)))
        cur.wait_stream(self._hp_stream)                 # the caller's stream sees the fitness vector in order
        return out

    def stream_chunk(self, n):
        """Genomes per chunk of `evaluate_streamed`: large enough that a chunk still fills the GPU (a chunk below
        ~4e8 layer-0-pixel x channel units is latency-bound: splitting C2's 32 gray genomes would only lengthen
        the evaluation), at most 8 chunks."""
        units = self.w * self.h * sum(self.channels)
        chunk = -(-int(4e8) // units)
        chunk = max(-(-chunk // 8) * 8, -(-n // 8))
        return min(max(chunk, 1), max(n, 1))

    def _staging(self, k, blob_bytes, n_offsets):
        """Pinned host + device buffers of chunk index k, grown geometrically."""
        if not hasattr(self, "_stage"):
            self._stage = []
        while len(self._stage) <= k:
            self._stage.append(None)
        st = self._stage[k]
        if st is None or st[0].numel() < blob_bytes or st[1].numel() < n_offsets:
            cb = max(2 * blob_bytes, 1 << 16)
            co = max(2 * n_offsets, 1 << 10)
            st = (torch.empty((cb,), dtype=torch.uint8).pin_memory(), torch.empty((co,), dtype=torch.int64).pin_memory(),
                  torch.empty((cb,), dtype=torch.uint8, device=self.tdev), torch.empty((co,), dtype=torch.int64, device=self.tdev))
            self._stage[k] = st
        return st

    def evaluate_streamed(self, items, flatten, structure, render_mode=RENDER_GRADIENT,
                          pair_mode=_lib.PAIR_POPULATION, chunk=None):
        """[(genome_id, genome)] -> device fp64 fitness tensor (a view of an engine-owned buffer), with the host-side
        flattening of chunk k+1 overlapped with the GPU evaluation of chunk k (SURVEY.md §8 f row 4: flattening is
        0.15-0.3 ms of Python per genome, a third of the GPU time of a generation when done up front).
        `flatten(genome_id, genome) -> FlatProgram`.  Chunks are queued on the engine's evaluation stream in population
        order; uploads go through pinned memory on a copy stream.  One synchronisation, at the end (range check)."""
        n = len(items)
        if n > self.max_genomes:
            raise ValueError("population of %d exceeds the engine capacity %d" % (n, self.max_genomes))
        if self._fit_buf is None:
            self._fit_buf = torch.empty((self.max_genomes,), dtype=torch.float64, device=self.tdev)
        out = self._fit_buf
        if n == 0:
            return out[:0]
        chunk = self.stream_chunk(n) if chunk is None else max(1, int(chunk))
        cuda = self.tdev.type == "cuda"
        ev = self._hp_stream if cuda else None
        if cuda:
            ev.wait_stream(torch.cuda.current_stream(self.tdev))
        keep = []
        for c0 in range(0, n, chunk):
            progs = [flatten(gid, g) for gid, g in items[c0:c0 + chunk]]
            blob, offsets, max_slots = G.pack_population(progs)
            max_blob = int(np.diff(offsets).max())
            if cuda:
                # pinned staging and device copies are kept per chunk index and re-used by later calls (the previous call
                # ended with a synchronisation): cudaHostAlloc per generation cost more than the flattening itself
                hb, ho, db, do = self._staging(len(keep), blob.nbytes, offsets.size)
                hb[:blob.nbytes].copy_(torch.from_numpy(blob))
                ho[:offsets.size].copy_(torch.from_numpy(offsets))
                with torch.cuda.stream(self._up_stream):
                    db[:blob.nbytes].copy_(hb[:blob.nbytes], non_blocking=True)
                    do[:offsets.size].copy_(ho[:offsets.size], non_blocking=True)
                ev.wait_stream(self._up_stream)
                keep.append((hb, ho, db, do))
                stream = C.c_void_p(ev.cuda_stream)
            else:
                db, do = torch.from_numpy(blob), torch.from_numpy(offsets)
                keep.append((db, do))
                stream = C.c_void_p(0)
            self.lib.check(self.lib.eig_eval(self.ctx, db.data_ptr(), do.data_ptr(), len(progs), max_slots, max_blob,
                                             int(structure), int(render_mode), int(pair_mode),
                                             out[c0:].data_ptr(), stream))
        self.lib.check(self.lib.eig_range_check(self.ctx, C.c_void_p(ev.cuda_stream) if cuda else C.c_void_p(0)))
        if cuda:
            torch.cuda.current_stream(self.tdev).wait_stream(ev)
        del keep
        return out[:n]

    def evaluate_host(self, blob, offsets, max_slots, structure, render_mode=RENDER_GRADIENT,
                      pair_mode=_lib.PAIR_POPULATION, out=None):
        """Host buffers in, host fitness out (H2D + kernels + D2H inside; synchronous)."""
        n = len(offsets) - 1
        if out is None:
            out = np.empty((n,), dtype=np.float64)
        self.lib.check(self.lib.eig_eval_host(self.ctx, blob.ctypes.data, offsets.ctypes.data, n, int(max_slots),
                                              int(structure), int(render_mode), int(pair_mode), out.ctypes.data))
        return out

    def evaluate(self, programs, structure, render_mode=RENDER_GRADIENT, pair_mode=_lib.PAIR_POPULATION):
        """[FlatProgram] -> numpy fitness vector, through the host entry point."""
        blob, offsets, max_slots = G.pack_population(programs)
        return self.evaluate_host(blob, offsets, max_slots, structure, render_mode, pair_mode)

    def debug_buffers(self, n):
        """Views (copied to numpy) of the context's internal buffers after the last evaluate*."""
        ptrs = [C.c_void_p() for _ in range(6)]
        self.lib.check(self.lib.eig_debug_buffers(self.ctx, *[C.byref(p) for p in ptrs]))
        self.sync()

        def grab(ptr, shape, dtype):
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            host = np.empty(shape, dtype=dtype)
            self.lib.check(self.lib.eig_memcpy_d2h(host.ctypes.data, ptr.value, nbytes))
            return host

        h, w, c = self.h, self.w, self.c_dim
        return dict(image=grab(ptrs[0], (n, h, w, c), np.uint8),
                    frames=grab(ptrs[1], (3, n, h, w, c), np.uint8),
                    vectors=grab(ptrs[2], (n, MAX_CORNERS, 4), np.float32),
                    nvec=grab(ptrs[3], (n,), np.int32),
                    corners=grab(ptrs[4], (n, MAX_CORNERS, 2), np.float32),
                    ncorners=grab(ptrs[5], (n,), np.int32))
