"""Test helper: interpret a FlatProgram with torch-CPU fp64 exactly the way csrc/render.cuh does
(separately rounded multiply and add, fp64 activations) so the flattener can be checked without a GPU."""
import numpy as np
import torch

from evolutionary_illusion_generator_b200 import genome as G


def _act(a, t):
    if a == 0:
        return torch.sigmoid(5 * t)
    if a == 1:
        return torch.tanh(2.5 * t)
    if a == 2:
        return torch.abs(t)
    if a == 3:
        return torch.exp(-5.0 * t ** 2)
    if a == 4:
        return t
    if a == 5:
        return torch.sin(t)
    if a == 6:
        return torch.nn.functional.relu(t)
    raise ValueError(a)


def run_program(prog, x, y):
    """x, y: float64 1-D tensors -> list of float64 tensors, one per output slot."""
    slots = [x, y, torch.ones_like(x)]
    for act, agg, t0, nt, bias, resp in prog.nodes:
        acc = None
        for wgt, s in prog.terms[t0:t0 + nt]:
            term = wgt * slots[s]
            if acc is None:
                acc = term
            elif agg == 0:
                acc = acc + term
            else:
                acc = acc * term
        slots.append(_act(act, resp * acc + bias))
    return [slots[s & 0x3fffffff] for s in prog.out_slots]


def render_flat(prog, grid, c_dim, w, h, bg=1, gradient=1):
    x_dat = np.asarray(grid["x_mat"], np.float64).reshape(h, w)
    y_dat = np.asarray(grid["y_mat"], np.float64).reshape(h, w)
    outs = run_program(prog, torch.tensor(x_dat.flatten()), torch.tensor(y_dat.flatten()))
    is_bg = x_dat == -1
    if c_dim > 1 and gradient == 1:
        arr = np.zeros((h, w, c_dim))
        for c in range(min(c_dim, len(outs))):
            arr[:, :, c] = outs[c].numpy().reshape(h, w)
            arr[:, :, c][is_bg] = bg
        return np.array(arr * 255.0, dtype=np.uint8)
    if c_dim > 1:
        idx = np.array(outs[0].numpy().reshape(h, w) * 4.0, dtype=np.uint8)
        img = np.zeros((h, w, 3), np.uint8)
        img[idx == 0] = 255
        for c in range(3):
            img[:, :, c][idx == c + 1] = 255
        img[is_bg] = bg * 255
        return img
    arr = outs[0].numpy().reshape(h, w).copy()
    arr[is_bg] = bg
    if gradient == 0:
        arr = np.round(arr)
    return np.array(arr * 255.0, dtype=np.uint8)
