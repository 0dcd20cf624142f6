"""PredNet weight files: Chainer `save_npz` key layout, synthetic generator, npz I/O.

Key layout = what `serializers.load_npz(initmodel, L.Classifier(PredNet(w, h, channels)))` reads
(/root/reference/chainer_prednet/PredNet/call_prednet.py:215-231; link names from net.py:45-62,143-154):

  predictor/ConvA{n}/{W,b}                n=1..3   W:(C_n, 2*C_{n-1}, 3, 3)
  predictor/ConvP{n}/{W,b}                n=0..3   W:(C_n, R_n, 3, 3)
  predictor/ConvLSTM{n}/h_{i,f,c,o}/{W,b}          W:(R_n, R_n, 3, 3)
  predictor/ConvLSTM{n}/c_{i,f,o}/W                (1, R_n, H_n, W_n)   peephole maps, tied to the resolution
  predictor/ConvLSTM{n}/x_{i,f,c,o}0/W             (R_n, 2*C_n, 3, 3)
  predictor/ConvLSTM{n}/x_{i,f,c,o}1/W             (R_n, R_{n+1}, 3, 3) n<3
"""
import numpy as np

PREFIX = "predictor/"


def layer_sizes(w, h, channels):
    out = []
    for c in channels:
        out.append((c, h, w))
        w, h = w // 2, h // 2
    return out


def expected_shapes(w, h, channels):
    L = len(channels)
    sz = layer_sizes(w, h, channels)
    shapes = {}
    for n in range(L):
        C, H, W = sz[n]
        if n > 0:
            shapes["ConvA%d/W" % n] = (C, 2 * channels[n - 1], 3, 3)
            shapes["ConvA%d/b" % n] = (C,)
        shapes["ConvP%d/W" % n] = (C, C, 3, 3)
        shapes["ConvP%d/b" % n] = (C,)
        pre = "ConvLSTM%d/" % n
        for g in "ifco":
            shapes[pre + "h_%s/W" % g] = (C, C, 3, 3)
            shapes[pre + "h_%s/b" % g] = (C,)
            shapes[pre + "x_%s0/W" % g] = (C, 2 * C, 3, 3)
            if n < L - 1:
                shapes[pre + "x_%s1/W" % g] = (C, channels[n + 1], 3, 3)
        for g in "ifo":
            shapes[pre + "c_%s/W" % g] = (1, C, H, W)
    return {PREFIX + k: v for k, v in shapes.items()}


def synthetic_weights(w, h, channels, seed=0, bias_std=0.0):
    """Seeded random weights with Chainer's default initialisers: LeCun-normal convolutions
    N(0, 1/fan_in) with zero bias, peephole maps N(0, 1/(W*H*C)) (net.py:18-19)."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shp in sorted(expected_shapes(w, h, channels).items()):
        if name.endswith("/b"):
            out[name] = (rng.normal(0, bias_std, shp) if bias_std else np.zeros(shp)).astype(np.float32)
        elif "/c_" in name:
            std = np.sqrt(1.0 / (shp[1] * shp[2] * shp[3]))
            out[name] = rng.normal(0, std, shp).astype(np.float32)
        else:
            fan_in = shp[1] * 9
            out[name] = rng.normal(0, np.sqrt(1.0 / fan_in), shp).astype(np.float32)
    return out


def save_npz(path, weights):
    np.savez(path, **weights)


def load_npz(path):
    with np.load(path) as f:
        return {k: np.ascontiguousarray(f[k], dtype=np.float32) for k in f.files}


def check_weights(weights, w, h, channels):
    want = expected_shapes(w, h, channels)
    missing = [k for k in want if k not in weights]
    if missing:
        raise KeyError("weight file lacks %d tensors, e.g. %s" % (len(missing), missing[:3]))
    for k, shp in want.items():
        if tuple(weights[k].shape) != tuple(shp):
            raise ValueError("%s has shape %s, expected %s for %dx%d channels %s (peephole maps tie a weight "
                             "file to one resolution, net.py:12)" % (k, weights[k].shape, shp, w, h, channels))
