"""Numerical prototype (CPU, float64) of the next algorithmic lever named in DESIGN.md: the taps of a 3x3 convolution over a
nearest-neighbour x2 up-sampled tensor, folded to half resolution.

ConvLSTM_n reads [E_n | up2(R_{n+1}) | h_n].  For an output pixel of parity (py, px) the nine taps over up2(R) touch only
a 2x2 neighbourhood of R, so the up(R) part of the convolution is, per parity, a 2x2 convolution of R with pre-summed
weights: 16 instead of 36 multiply-accumulates per half-resolution pixel, channel and output column (-56 % of that part,
-22 % of ConvLSTM_1/2).  As one half-resolution 3x3 pass with 4N output columns (parity-major) each tap feeds only the
parities that use it - corner taps one, edge taps two, the centre tap four - which is the column slicing a tensor-core
kernel needs (`build_z_weights` in csrc/eig_api.cu is the layer-0 instance of the same table, without the slicing).
Run: python profiles/experiments/parity_fold_proto.py
"""
import numpy as np

# half-resolution row offset (-1, 0, +1 -> index 0, 1, 2) read by full-resolution tap ky for output row parity py
#   py = 0: ky 0 -> row y2-1, ky 1, 2 -> row y2        py = 1: ky 0, 1 -> row y2, ky 2 -> row y2+1
HALF_TAP = {0: (0, 1, 1), 1: (1, 1, 2)}


def direct(R, W):
    """conv3x3(pad 1) over up2(R).  R: (C, H2, W2), W: (N, C, 3, 3) -> (N, 2*H2, 2*W2)."""
    up = np.repeat(np.repeat(R, 2, axis=1), 2, axis=2)
    C, H, Wd = up.shape
    pad = np.zeros((C, H + 2, Wd + 2))
    pad[:, 1:-1, 1:-1] = up
    out = np.zeros((W.shape[0], H, Wd))
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("nc,chw->nhw", W[:, :, ky, kx], pad[:, ky:ky + H, kx:kx + Wd])
    return out


def fold_weights(W):
    """(N, C, 3, 3) -> (2, 2, N, C, 3, 3): per output parity a half-resolution 3x3 kernel with at most 2x2 non-zero taps."""
    N, C = W.shape[:2]
    F = np.zeros((2, 2, N, C, 3, 3))
    for py in range(2):
        for px in range(2):
            for ky in range(3):
                for kx in range(3):
                    F[py, px, :, :, HALF_TAP[py][ky], HALF_TAP[px][kx]] += W[:, :, ky, kx]
    return F


def folded(R, W):
    F = fold_weights(W)
    C, H2, W2 = R.shape
    pad = np.zeros((C, H2 + 2, W2 + 2))
    pad[:, 1:-1, 1:-1] = R
    out = np.zeros((W.shape[0], 2 * H2, 2 * W2))
    macs = 0
    for ty in range(3):
        for tx in range(3):
            users = [(py, px) for py in range(2) for px in range(2) if np.any(F[py, px, :, :, ty, tx])]
            for py, px in users:        # one MMA per (tap, parity group): D[:, columns of these parities] += A_tap x B
                out[:, py::2, px::2] += np.einsum("nc,chw->nhw", F[py, px, :, :, ty, tx], pad[:, ty:ty + H2, tx:tx + W2])
            macs += len(users)
    return out, macs, F


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    R, W = rng.randn(6, 5, 7), rng.randn(8, 6, 3, 3)
    want = direct(R, W)
    got, macs, F = folded(R, W)
    print("max abs difference folded vs direct: %.2e" % np.abs(got - want).max())
    print("multiply-accumulates per half-resolution pixel, channel and output column: %d folded, %d direct" % (macs, 36))
    print("parities (py, px) fed by each half-resolution tap (ty, tx):")
    for ty in range(3):
        print("   ", ["".join("%d%d " % (py, px) for py in range(2) for px in range(2) if np.any(F[py, px, :, :, ty, tx])).strip()
                      for tx in range(3)])
    assert np.abs(got - want).max() < 1e-12 and macs == 16
