"""Oracle: motion-vector scoring.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Vectorised numpy restatement of the live scoring branches of the reference:
  plausibility_ratio            /root/reference/fitness_calculator.py:18-27
  strength_number               fitness_calculator.py:32-41
  horizontal_symmetry_score     fitness_calculator.py:81-120   (incl. the line-101 slice broadcast)
  swarm_score                   fitness_calculator.py:124-159  (incl. the `% 2 * math.pi` precedence)
  rotation_symmetry_score       fitness_calculator.py:166-215
  branch logic                  /root/reference/generate_illusion.py:557-616 (= fitness_calculator.py:505-548,
                                with score_d initialised to 0 as generate_illusion.py:566 does)
Pinned against the reference functions themselves by tests/test_oracle_vs_reference.py.
Input rows are (x, y, dx, dy) float32, as optical_flow.py:73-82 produces them.
"""
import math

import numpy as np

BANDS, CIRCLES, FREE, CIRCLES_FREE = 0, 1, 2, 3
NO_VECTOR_SENTINEL = np.array([[0, 0, -1000, 0]], dtype=np.float64)  # generate_illusion.py:554


def plausible(v, limit):
    v = np.asarray(v)
    norm = np.sqrt(v[:, 2] * v[:, 2] + v[:, 3] * v[:, 3])
    return v[~(norm > limit)]


def strength_number(v, max_norm):
    v = np.asarray(v)
    mx = np.mean(np.abs(v[:, 2]))
    norms = np.sqrt(v[:, 2] * v[:, 2] + v[:, 3] * v[:, 3])
    var = np.var(norms)
    return (mx / max_norm) * (1 - min(var, 1))


def horizontal_symmetry_score(v, limits):
    v = np.asarray(v)
    middle = int(limits[1] / 2)
    sel = v[~((v[:, 1] < limits[0]) | (v[:, 1] > limits[1]))]
    if len(sel) == 0:
        return 0
    nrm = np.sqrt(sel[:, 2] * sel[:, 2] + sel[:, 3] * sel[:, 3])
    nx, ny = sel[:, 2] / nrm, sel[:, 3] / nrm
    upper = sel[:, 1] < middle
    m = np.zeros((len(sel), 2))
    m[:, 0] = np.where(upper, nx, -nx)
    m[:, 1] = np.where(upper, nx, ny)  # fitness_calculator.py:101 broadcasts x into both columns
    return ((1 - np.var(m[:, 0])) + abs(np.mean(m[:, 0])) + (1 - abs(np.mean(m[:, 1])))) / 3


def swarm_score(v):
    nv = np.array(v)
    n = len(nv)
    norms = np.sqrt(nv[:, 2] * nv[:, 2] + nv[:, 3] * nv[:, 3])
    nv[:, 2] = nv[:, 2] / norms
    nv[:, 3] = nv[:, 3] / norms
    angles = np.arccos(nv[:, 2])
    score = 0
    for a in nv:
        x = nv[:, 0] - a[0]
        y = nv[:, 1] - a[1]
        f = (x * x + y * y) / (100 * 100)
        f = np.where(f > 1, 1, f)
        close = 1 - np.where(f < 1, 0, f)
        optimal = (math.acos(a[2]) + f * math.pi) % 2 * math.pi
        loss = close * abs(angles - optimal)
        score = score + (math.pi - (sum(loss) / n)) / math.pi
    return score / n


def rotation_symmetry_score(v, w, h, limits):
    v = np.asarray(v)
    cx, cy = w / 2, h / 2
    px = v[:, 0] - cx
    py = v[:, 1] - cy
    dist = np.sqrt(px * px + py * py)
    keep = ~((dist < limits[0]) | (dist > limits[1]) | (dist == 0))
    if keep.sum() < 2:
        return 0
    px, py, dist = px[keep].astype(np.float64), py[keep].astype(np.float64), dist[keep].astype(np.float64)
    dx, dy = v[keep, 2].astype(np.float64), v[keep, 3].astype(np.float64)
    nrm = np.sqrt(dx * dx + dy * dy)
    dx, dy = dx / nrm, dy / nrm
    ex, ey = px + dx, py + dy
    rx = (ex * px + ey * py) / dist - dist
    ry = (-ex * py + ey * px) / dist
    vx, vy = np.var(rx), np.var(ry)
    return ((1 - vx) * (1 - vx) + (1 - vy) * (1 - vy)) / 2


def fitness_from_vectors(structure, vectors, w, h):
    """One genome's fitness from its flow vectors (None / empty -> the no-vector sentinel)."""
    if vectors is None or len(vectors) == 0:
        vectors = NO_VECTOR_SENTINEL
    score = 0
    if structure == BANDS:
        good = plausible(vectors, 0.15)
        if len(good) > 0:
            score = horizontal_symmetry_score(good, [0, (h / 4) * 2])
    elif structure in (CIRCLES, CIRCLES_FREE):
        good = plausible(vectors, 0.3)
        if len(good) > 24:
            score = 0.7 * rotation_symmetry_score(good, w, h, [0, h / 2]) + 0.3 * strength_number(good, 0.3)
    elif structure == FREE:
        good = plausible(vectors, 0.4)
        if len(good) > 0:
            score = (0.5 * swarm_score(good) + 0.1 * strength_number(good, 0.4)
                     + 0.4 * (min(len(good), 15) / 15))
    else:
        raise ValueError("structure %r has no live scoring branch in the reference" % (structure,))
    return float(score)
