"""Oracle: PredNet inference in torch-CPU fp32.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PINNED TO THE REFERENCE'S OWN CODE UNDER A CHAINER STAND-IN: Chainer (third-party, v5-v7) cannot be installed offline
and the reference holds no test / golden vector for this stage.  `tests/golden/chainer_shim` restates the few Chainer
primitives net.py uses; with it the reference's net.py, call_prednet.py and the whole get_fitnesses_neat run unmodified
(tests/golden/ref_harness.py), and tests/golden/reference_pipeline.npz holds their frames and fitness values.  This file
reproduces them (frames within 1 LSB, fitness within 4e-4 relative; tests/test_oracle_golden.py).  What stays
unpinned is the arithmetic inside Chainer's primitives, restated from its documentation.  This file restates
  * `PredNet.__call__`   /root/reference/chainer_prednet/PredNet/net.py:175-211
  * `ConvLSTM.__call__`  net.py:84-126,  `EltFilter.__call__` net.py:30-34
  * frame protocol       /root/reference/chainer_prednet/PredNet/call_prednet.py:129-205 (`test_image_list`),
                         `read_image` 29-49, `write_image` 51-61
with the Chainer semantics listed in SURVEY.md §8(c): Convolution2D = cross-correlation, stride 1, pad 1,
W:(out,in,3,3); max_pooling_2d(2, stride=2); unpooling_2d(2, stride=2, cover_all=False) = nearest x2;
clipped_relu(x, 1) = min(max(x, 0), 1); sigmoid(x) = tanh(x*0.5)*0.5+0.5 (Chainer's CPU forward);
gate pre-activation summed as x0-conv + x1-conv + (h-conv + bias) + peephole, all fp32.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def chainer_sigmoid(x):
    return torch.tanh(x * 0.5) * 0.5 + 0.5


class PredNetOracle:
    def __init__(self, weights, channels, w, h, prefix="predictor/"):
        """weights: dict name -> ndarray in the Chainer npz layout."""
        self.ch = list(channels)
        self.L = len(self.ch)
        self.w, self.h = w, h
        self.p = {k[len(prefix):] if k.startswith(prefix) else k: _t(v) for k, v in weights.items()}
        self.reset_state()

    def reset_state(self):
        self.P = [None] * self.L
        self.hs = [None] * self.L
        self.cs = [None] * self.L

    def _conv(self, name, x, bias=True):
        b = self.p.get(name + "/b") if bias else None
        return F.conv2d(x, self.p[name + "/W"], b, padding=1)

    def _lstm(self, n, xs):
        pre = "ConvLSTM%d/" % n
        B = xs[0].shape[0]
        hh, ww = xs[0].shape[2], xs[0].shape[3]
        if self.hs[n] is None:
            self.hs[n] = torch.zeros(B, self.ch[n], hh, ww)
        if self.cs[n] is None:
            self.cs[n] = torch.zeros(B, self.ch[n], hh, ww)
        h_old, c_old = self.hs[n], self.cs[n]

        def gate(g, peep):
            acc = self._conv(pre + "x_%s0" % g, xs[0], bias=False)
            for k in range(1, len(xs)):
                acc = acc + self._conv(pre + "x_%s%d" % (g, k), xs[k], bias=False)
            acc = acc + self._conv(pre + "h_%s" % g, h_old)
            if peep:
                acc = acc + c_old * self.p[pre + "c_%s/W" % g]
            return acc

        i = chainer_sigmoid(gate("i", True))
        f = chainer_sigmoid(gate("f", True))
        c_new = torch.tanh(gate("c", False)) * i + f * c_old
        o = chainer_sigmoid(gate("o", True))  # peephole on the OLD cell state (net.py:116-124)
        h_new = o * torch.tanh(c_new)
        self.cs[n], self.hs[n] = c_new, h_new
        return h_new

    def step(self, x):
        """x: (B, C0, h, w) float32 -> P0 (B, C0, h, w)."""
        B = x.shape[0]
        hh, ww = self.h, self.w
        for n in range(self.L):
            if self.P[n] is None:
                self.P[n] = torch.zeros(B, self.ch[n], hh, ww)
            hh, ww = hh // 2, ww // 2
        E = [None] * self.L
        E[0] = torch.cat((F.relu(x - self.P[0]), F.relu(self.P[0] - x)), dim=1)
        for n in range(1, self.L):
            A = F.max_pool2d(F.relu(self._conv("ConvA%d" % n, E[n - 1])), 2, stride=2)
            E[n] = torch.cat((F.relu(A - self.P[n]), F.relu(self.P[n] - A)), dim=1)
        R = [None] * self.L
        for n in reversed(range(self.L)):
            if n == self.L - 1:
                R[n] = self._lstm(n, (E[n],))
            else:
                up = F.interpolate(R[n + 1], scale_factor=2, mode="nearest")
                R[n] = self._lstm(n, (E[n], up))
            pn = self._conv("ConvP%d" % n, R[n])
            self.P[n] = torch.clamp(pn, 0.0, 1.0) if n == 0 else F.relu(pn)
        return self.P[0]


def image_to_input(img_u8):
    """read_image (call_prednet.py:29-49): (h,w[,3]) uint8 -> (C,h,w) float32 = float32(float64(u8)/255)."""
    a = np.asarray(img_u8)
    a = a.reshape(1, a.shape[0], a.shape[1]) if a.ndim == 2 else a.transpose(2, 0, 1)
    return (a / 255).astype(np.float32)


def prediction_to_image(p0):
    """write_image (call_prednet.py:51-61): (C,h,w) float32 -> uint8 by fp32 multiply and truncation."""
    a = np.array(p0, dtype=np.float32, copy=True)
    a *= 255
    a = a.transpose(1, 2, 0).astype(np.uint8)
    return a[:, :, 0] if a.shape[2] == 1 else a


def run_genome_frames(net, img_u8, repeat=20, extension=2):
    """The test_image_list protocol for one genome: reset, `repeat` forwards on the same image, then
    `extension` forwards fed with the previous *unquantised* prediction.  Returns the uint8 frames
    [pred #repeat, ext #1, ..., ext #extension] (files {20i+19}.png, {20i+20}_extended.png, ...)."""
    net.reset_state()
    x = torch.from_numpy(image_to_input(img_u8))[None]
    with torch.no_grad():
        for _ in range(repeat):
            p = net.step(x)
        frames = [prediction_to_image(p[0].numpy())]
        for _ in range(extension):
            p = net.step(p.clone())
            frames.append(prediction_to_image(p[0].numpy()))
    net.reset_state()
    return frames
