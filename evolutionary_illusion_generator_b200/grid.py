"""CPPN input planes (`create_grid`, /root/reference/generate_illusion.py:196-317; `fill_circle` 38-117).

Genome-independent, computed once per (structure, w, h) on the host in fp64 and uploaded with
`eig_set_grid`; cached.  Vectorised numpy, bit-identical to the reference's scalar loops (checked in
tests/test_grid.py against the oracle restatement and, where the reference tree exists, the reference).
Returned planes have shape (h, w); `x_mat == -1` marks background.
"""
import math
from enum import IntEnum

import numpy as np


class StructureType(IntEnum):
    Bands = 0
    Circles = 1
    Free = 2
    CirclesFree = 3


_cache = {}


def _ring_edges(n=10):
    e = np.zeros(n)
    e[n - 1] = 1
    for i in range(2, n + 1):
        e[n - i] = e[n - i + 1] * 1.5
    return e / e[0]


def _theta(x, y):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(x == 0, math.pi / 2.0, np.arctan(y * 1.0 / np.where(x == 0, 1.0, x)))
    return np.where(x < 0, t + math.pi, t)


def polar_cells(x, y, max_radius, direction=1, structure=StructureType.Circles):
    """Vectorised `fill_circle`: x, y arrays of centre-relative coordinates -> (r, theta) arrays."""
    edges = _ring_edges()
    n = len(edges)
    r_total = np.sqrt(x * x + y * y)
    inside = r_total <= max_radius / 2
    radius = np.minimum(1, r_total / (max_radius / 2))
    r = np.full(x.shape, -1.0)
    ring = np.zeros(x.shape, dtype=np.int64)
    assigned = np.zeros(x.shape, dtype=bool)
    for i in range(1, n - 1):
        hit = (~assigned) & (radius > edges[i])
        rv = (radius - edges[i]) / (edges[i - 1] - edges[i])
        if direction < 0:
            rv = 1 - rv
        r = np.where(hit, rv, r)
        ring = np.where(hit, n - i - 1, ring)
        assigned |= hit
    if structure in (StructureType.Circles, StructureType.CirclesFree):
        theta = _theta(x, y)
        theta = np.where(ring % 2 == 1, theta + math.pi / 4.0, theta)
        if structure == StructureType.Circles:
            theta = theta % (math.pi / 6.0)
        if direction < 0:
            theta = (math.pi / 6.0) - theta
    else:   # fill_circle leaves theta at 0 for Bands / Free (generate_illusion.py:66-104)
        theta = np.zeros(x.shape)
    white = (r > 0.9) | (r < 0.1)
    r_out = np.where(white, -1.0, r / 0.8)
    theta = np.where(white, 0.0, theta)
    return np.where(inside, r_out, -1.0), np.where(inside, theta, 0.0)


def create_grid(structure, x_res=32, y_res=32, scaling=1.0):
    key = (int(structure), x_res, y_res, float(scaling))
    if key in _cache:
        g = _cache[key]
        return {"x_mat": g[0].copy(), "y_mat": g[1].copy()}
    w, h = x_res, y_res
    st = StructureType(int(structure))
    if st == StructureType.Bands:
        y_rep, padding = 4, 10
        y_len = int(h / y_rep)
        sc = scaling / y_rep
        seg = np.concatenate((np.linspace(-1 * sc, sc, num=y_len - padding), np.zeros(padding)))
        y_range = np.tile(seg, y_rep)
        x_rep = 10
        x_len = int(w / x_rep)
        sc = scaling / x_rep
        x_range = np.tile(np.linspace(-1 * sc, sc, num=x_len), x_rep)
        flip = np.ones((h, 1))
        start = y_len
        while start < h:
            flip[max(0, start - padding):start] = 0
            stop = min(h, start + y_len)
            flip[max(stop - padding, 0):stop] = 0
            flip[start:stop] = -flip[start:stop]
            start += 2 * y_len
        x_mat = np.matmul(flip, x_range.reshape((1, w))).reshape(h, w)
        y_mat = np.matmul(y_range.reshape((h, 1)), np.ones((1, w))).reshape(h, w)
    else:
        x_range = np.linspace(-1 * scaling, scaling, num=w)
        y_range = np.linspace(-1 * scaling, scaling, num=h)
        y_mat = np.matmul(y_range.reshape((h, 1)), np.ones((1, w)))
        x_mat = np.matmul(np.ones((h, 1)), x_range.reshape((1, w)))
        if st != StructureType.Free:
            xs = (np.arange(w) - (w / 2))[None, :] * np.ones((h, 1))
            ys = (np.arange(h) - (h / 2))[:, None] * np.ones((1, w))
            if st == StructureType.Circles:
                x_mat, y_mat = polar_cells(xs, ys, h, 1, StructureType.Circles)
            else:
                r_len = int(h / 6)
                r_total = np.sqrt(xs * xs + ys * ys)
                x_mat = (np.minimum(r_total, h / 2) % r_len) / r_len
                th = _theta(xs, ys)
                th = np.where((r_total / r_len).astype(np.int64) % 2 == 1, th + math.pi / 4.0, th)
                y_mat = np.where(r_total < h / 2, th, 0.0)
    _cache[key] = (np.ascontiguousarray(x_mat, np.float64), np.ascontiguousarray(y_mat, np.float64))
    return create_grid(structure, x_res, y_res, scaling)


_enh_cache = {}


def enhanced_image_grid(x_res, y_res, structure):
    """The 3x3 + 2x2 circle mosaic of the per-generation `enhanced.png` export
    (/root/reference/generate_illusion.py:121-193), vectorised and cached (it is genome-independent; the reference
    spends ~10 s per generation in its scalar loops).  Circles alternate direction with their index."""
    key = (int(x_res), int(y_res), int(structure))
    if key not in _enh_cache:
        st = StructureType(int(structure))
        c_rows = c_cols = 3
        y_step, x_step = int(y_res / c_cols), int(x_res / c_cols)
        x_mat = np.ones((y_res, x_res)) * -1
        y_mat = np.ones((y_res, x_res)) * -1
        yy, xx = np.meshgrid(np.arange(y_step), np.arange(x_step), indexing="ij")
        for row in range(c_rows):
            for col in range(c_cols):
                index = row * c_cols + col
                direction = -1 if index % 2 == 0 else 1
                real_x, real_y = col * x_step + xx, row * y_step + yy
                x = real_x - (x_step * col + x_step / 2)
                y = real_y - (y_step * row + y_step / 2)
                r, theta = polar_cells(x, y, y_step, direction, st)
                x_mat[real_y, real_x] = r
                y_mat[real_y, real_x] = theta
        sub = c_rows - 1
        for row in range(sub):
            for col in range(sub):
                index = c_rows * c_cols + row * sub + col
                direction = -1 if index % 2 == 0 else 1
                real_x = col * x_step + xx + int(x_step / 2)
                real_y = row * y_step + yy + int(y_step / 2)
                x = real_x - (x_step * col + x_step)
                y = real_y - (y_step * row + x_step)      # sic: the reference uses x_step for the y centre (line 147)
                hit = np.sqrt(x * x + y * y) < x_step / 2
                r, theta = polar_cells(x, y, y_step, direction, st)
                x_mat[real_y[hit], real_x[hit]] = r[hit]
                y_mat[real_y[hit], real_x[hit]] = theta[hit]
        _enh_cache[key] = (x_mat, y_mat)
    g = _enh_cache[key]
    return {"x_mat": g[0].copy(), "y_mat": g[1].copy()}
