"""chainer.functions used by net.py: float32 numpy forwards with Chainer's documented semantics."""
import numpy

from ..variable import Variable, _raw
from . import loss  # noqa: F401


def _v(a):
    return Variable(numpy.ascontiguousarray(a))


def relu(x):
    a = _raw(x)
    return _v(numpy.maximum(a, 0, dtype=a.dtype))


def clipped_relu(x, z=20.0):
    a = _raw(x)
    return _v(numpy.minimum(numpy.maximum(0, a), z).astype(a.dtype, copy=False))


def sigmoid(x):
    a = _raw(x)
    half = a.dtype.type(0.5)
    return _v(numpy.tanh(a * half) * half + half)        # the form of Chainer's CPU forward


def tanh(x):
    return _v(numpy.tanh(_raw(x)))


def concat(xs, axis=1):
    return _v(numpy.concatenate([_raw(x) for x in xs], axis=axis))


def im2col(x, kh, kw, sy, sx, ph, pw, pval=0.0, cover_all=False):
    """(n, c, h, w) -> (n, c, kh, kw, out_h, out_w), the layout Chainer's conv / pooling forwards use."""
    n, c, h, w = x.shape

    def out_size(size, k, s, p):
        return (size + p * 2 - k + s - 1) // s + 1 if cover_all else (size + p * 2 - k) // s + 1

    out_h, out_w = out_size(h, kh, sy, ph), out_size(w, kw, sx, pw)
    img = numpy.pad(x, ((0, 0), (0, 0), (ph, ph + sy - 1), (pw, pw + sx - 1)), mode="constant", constant_values=(pval,))
    col = numpy.ndarray((n, c, kh, kw, out_h, out_w), dtype=x.dtype)
    for j in range(kh):
        for i in range(kw):
            col[:, :, j, i, :, :] = img[:, :, j:j + sy * out_h:sy, i:i + sx * out_w:sx]
    return col


def convolution_2d(x, W, b=None, stride=1, pad=0):
    """Cross-correlation: y[n, o, y, x] = sum_{c, j, i} W[o, c, j, i] * xpad[n, c, y*s + j, x*s + i] (+ b[o])."""
    a, w = _raw(x), _raw(W)
    kh, kw = w.shape[2:]
    col = im2col(a, kh, kw, stride, stride, pad, pad)
    y = numpy.tensordot(col, w, ((1, 2, 3), (1, 2, 3))).astype(a.dtype, copy=False)      # (n, out_h, out_w, out_c)
    if b is not None:
        y += _raw(b)
    return _v(numpy.rollaxis(y, 3, 1))


def max_pooling_2d(x, ksize, stride=None, pad=0, cover_all=True):
    a = _raw(x)
    stride = ksize if stride is None else stride
    col = im2col(a, ksize, ksize, stride, stride, pad, pad, pval=-float("inf"), cover_all=cover_all)
    n, c, kh, kw, out_h, out_w = col.shape
    return _v(col.reshape(n, c, kh * kw, out_h, out_w).max(axis=2))


def unpooling_2d(x, ksize, stride=None, pad=0, outsize=None, cover_all=True):
    """Every input pixel is spread over its ksize x ksize window (overlaps add; none for ksize == stride)."""
    a = _raw(x)
    stride = ksize if stride is None else stride
    n, c, h, w = a.shape
    if outsize is None:
        def size(s):
            return stride * (s - 1) + ksize - 2 * pad - (stride - 1 if cover_all else 0)
        outsize = (size(h), size(w))
    out_h, out_w = outsize
    canvas = numpy.zeros((n, c, out_h + 2 * pad + stride - 1, out_w + 2 * pad + stride - 1), dtype=a.dtype)
    for j in range(ksize):
        for i in range(ksize):
            canvas[:, :, j:j + stride * h:stride, i:i + stride * w:stride] += a
    return _v(canvas[:, :, pad:pad + out_h, pad:pad + out_w])
