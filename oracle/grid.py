"""Oracle: CPPN input planes.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Scalar restatement of `create_grid` (/root/reference/generate_illusion.py:196-317) and `fill_circle`
(generate_illusion.py:38-117).  Returns fp64 planes of shape (h, w); x_mat == -1 marks background.
Bands is reshaped to (h, w) (the reference returns (1, w*h, 1), generate_illusion.py:236-237).
"""
import math

import numpy as np

BANDS, CIRCLES, FREE, CIRCLES_FREE = 0, 1, 2, 3


def ring_edges(n=10):
    # generate_illusion.py:41-48 : geometric ring radii, ratio 1.5, normalised by the largest
    e = np.zeros(n)
    e[n - 1] = 1
    for i in range(2, n + 1):
        e[n - i] = e[n - i + 1] * 1.5
    return e / e[0]


def polar_cell(x, y, max_radius, direction, structure=CIRCLES):
    """(r, theta) of one pixel relative to the circle centre; generate_illusion.py:38-117."""
    edges = ring_edges()
    n = len(edges)
    r_total = np.sqrt(x * x + y * y)
    theta = 0
    r = -1
    if r_total <= max_radius / 2:
        radius = min(1, r_total / (max_radius / 2))
        ring = 0
        for i in range(1, n - 1):
            if radius > edges[i]:
                r = (radius - edges[i]) / (edges[i - 1] - edges[i])
                if direction < 0:
                    r = 1 - r
                ring = n - i - 1
                break
        if structure in (CIRCLES, CIRCLES_FREE):
            theta = math.pi / 2.0 if x == 0 else np.arctan(y * 1.0 / x)
            if x < 0:
                theta = theta + math.pi
            if ring % 2 == 1:
                theta = theta + math.pi / 4.0
            if structure == CIRCLES:
                theta = theta % (math.pi / 6.0)
            if direction < 0:
                theta = (math.pi / 6.0) - theta
        if (r > 0.9) or (r < 0.1):
            r = -1
            theta = 0
        else:
            r = r / 0.8
    return r, theta


def create_grid(structure, w, h, scaling=10.0):
    if structure == BANDS:
        y_rep, padding = 4, 10
        y_len = int(h / y_rep)
        sc = scaling / y_rep
        seg = np.concatenate((np.linspace(-sc, sc, num=y_len - padding), np.zeros(padding)))
        y_range = np.tile(seg, y_rep)
        x_rep = 10
        x_len = int(w / x_rep)
        sc = scaling / x_rep
        x_range = np.tile(np.linspace(-sc, sc, num=x_len), x_rep)
        flip = np.ones((h, 1))
        start = y_len
        while start < h:
            m0 = max(0, start - padding)
            flip[m0:start] = 0
            stop = min(h, start + y_len)
            m0 = max(stop - padding, 0)
            flip[m0:stop] = 0
            flip[start:stop] = -flip[start:stop]
            start += 2 * y_len
        x_mat = np.matmul(flip, x_range.reshape((1, w)))
        y_mat = np.matmul(y_range.reshape((h, 1)), np.ones((1, w)))
        return {"x_mat": x_mat.reshape(h, w), "y_mat": y_mat.reshape(h, w)}
    x_range = np.linspace(-scaling, scaling, num=w)
    y_range = np.linspace(-scaling, scaling, num=h)
    y_mat = np.matmul(y_range.reshape((h, 1)), np.ones((1, w)))
    x_mat = np.matmul(np.ones((h, 1)), x_range.reshape((1, w)))
    if structure == FREE:
        return {"x_mat": x_mat, "y_mat": y_mat}
    if structure == CIRCLES:
        for xx in range(w):
            x = xx - (w / 2)
            for yy in range(h):
                y = yy - (h / 2)
                x_mat[yy, xx], y_mat[yy, xx] = polar_cell(x, y, h, 1)
        return {"x_mat": x_mat, "y_mat": y_mat}
    if structure == CIRCLES_FREE:
        r_len = int(h / 6)
        for xx in range(w):
            x = xx - (w / 2)
            for yy in range(h):
                y = yy - (h / 2)
                r_total = np.sqrt(x * x + y * y)
                r = (min(r_total, h / 2) % r_len) / r_len
                theta = 0
                if r_total < h / 2:
                    theta = math.pi / 2.0 if x == 0 else np.arctan(y * 1.0 / x)
                    if x < 0:
                        theta = theta + math.pi
                    if int(r_total / r_len) % 2 == 1:
                        theta = theta + math.pi / 4.0
                x_mat[yy, xx] = r
                y_mat[yy, xx] = theta
        return {"x_mat": x_mat, "y_mat": y_mat}
    raise ValueError("unknown structure %r" % (structure,))


def enhanced_image_grid(x_res, y_res, structure):
    """Scalar restatement of `enhanced_image_grid` (/root/reference/generate_illusion.py:121-193): 3x3 circles plus a
    2x2 overlay, directions alternating with the circle index."""
    c_rows = c_cols = 3
    y_step, x_step = int(y_res / c_cols), int(x_res / c_cols)
    sub_rows = sub_cols = c_rows - 1
    centers = [None] * (c_rows * c_cols + sub_rows * sub_cols)
    for y in range(c_rows):
        for x in range(c_cols):
            centers[y * c_cols + x] = [x_step * x + x_step / 2, y_step * y + y_step / 2]
    for y in range(sub_rows):
        for x in range(sub_cols):
            centers[c_rows * c_cols + y * sub_cols + x] = [x_step * x + x_step, y_step * y + x_step]
    y_mat = np.ones((y_res, x_res)) * -1
    x_mat = np.ones((y_res, x_res)) * -1
    for row in range(c_rows):
        for col in range(c_cols):
            index = row * c_cols + col
            direction = -1 if index % 2 == 0 else 1
            for xx in range(x_step):
                real_x = col * x_step + xx
                x = real_x - centers[index][0]
                for yy in range(y_step):
                    real_y = row * y_step + yy
                    y = real_y - centers[index][1]
                    r, theta = polar_cell(x, y, y_step, direction, structure)
                    x_mat[real_y, real_x] = r
                    y_mat[real_y, real_x] = theta
    for row in range(sub_rows):
        for col in range(sub_cols):
            index = c_rows * c_cols + row * sub_rows + col
            direction = -1 if index % 2 == 0 else 1
            for xx in range(x_step):
                real_x = (col * x_step + xx) + int(x_step / 2)
                x = real_x - centers[index][0]
                for yy in range(y_step):
                    real_y = (row * y_step + yy) + int(y_step / 2)
                    y = real_y - centers[index][1]
                    if np.sqrt(x * x + y * y) < x_step / 2:
                        r, theta = polar_cell(x, y, y_step, direction, structure)
                        x_mat[real_y, real_x] = r
                        y_mat[real_y, real_x] = theta
    return {"x_mat": x_mat, "y_mat": y_mat}
