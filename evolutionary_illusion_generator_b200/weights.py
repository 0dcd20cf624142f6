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


def synthetic_predictor_weights(w, h, channels, seed=0, k=1.0, g=1.5, rand0=0.15, b_i=3.0, b_f=5.0, b_o=3.0):
    """`synthetic_weights` re-shaped so that the random network behaves like a trained predictor (P0 tracks the
    input, tiny frame-to-frame flow of 0.01-0.05 px, as the published weights give): plain LeCun-normal
    weights make P0 unrelated to the frame, every flow vector exceeds the plausibility limits of
    generate_illusion.py:569-597 and every genome scores 0, which would make fitness parity vacuous.
    Layer 0 becomes an error integrator: the random layer-0 convolutions are scaled by `rand0`, the cell
    candidate gets +k / -k centre taps on E0 = [relu(x-P0), relu(P0-x)], the gates are biased open (input,
    output) and to remember (forget), and ConvP0 gets an identity centre tap of gain g.  Layers 1-3 keep the
    plain initialisation; they modulate layer 0 through the x_*1 convolutions.  Same shapes, same arithmetic
    cost as any other weight file."""
    wt = synthetic_weights(w, h, channels, seed=seed)
    c0 = channels[0]
    pre = PREFIX + "ConvLSTM0/"
    for name in wt:
        if (name.startswith(pre) or name.startswith(PREFIX + "ConvP0/")) and name.endswith("/W") and "/c_" not in name:
            wt[name] *= np.float32(rand0)
    for r in range(c0):
        wt[pre + "x_c0/W"][r, r, 1, 1] += k
        wt[pre + "x_c0/W"][r, c0 + r, 1, 1] -= k
        wt[PREFIX + "ConvP0/W"][r, r, 1, 1] += g
    wt[pre + "h_i/b"][:] = b_i
    wt[pre + "h_f/b"][:] = b_f
    wt[pre + "h_o/b"][:] = b_o
    return wt


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
