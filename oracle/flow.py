"""Oracle: sparse optical flow (Shi-Tomasi corners + pyramidal Lucas-Kanade).  TEST INFRASTRUCTURE ONLY.

The reference calls OpenCV for this stage (/root/reference/optical_flow/optical_flow.py:40-89:
`cv2.cvtColor(BGR2GRAY)`, `cv2.goodFeaturesToTrack(maxCorners=100, qualityLevel=0.3, minDistance=7,
blockSize=7)`, `cv2.calcOpticalFlowPyrLK(winSize=(50,50), maxLevel=2, criteria=(EPS|COUNT, 10, 0.03))`).
OpenCV is third-party and unpinned by the reference; this image has opencv-python-headless 4.13.0.

Two checkers live here:
  * `lucas_kanade_cv2`  - the reference's call sequence on in-memory images with the cv2 binary itself;
  * `lucas_kanade_np`   - a from-scratch numpy restatement of OpenCV's published algorithm (imgproc
    color/featureselect/corner/deriv/pyramids, video/lkpyramid), which is the specification the CUDA kernels
    follow.  It is pinned against the cv2 binary by tests/test_oracle_flow.py: gray / eigen map / corner list
    bit-identical, LK end points within 1e-4 px (the only difference: OpenCV accumulates the 2500-pixel
    window sums in fp32 SIMD lanes, here and on the GPU they are exact integer sums).
"""
import numpy as np

MAX_CORNERS = 100
QUALITY = 0.3
MIN_DIST = 7
BLOCK = 7
WIN = 50
MAX_LEVEL = 2
MAX_ITERS = 10
EPS = 0.03
MIN_EIG_THRESHOLD = 1e-4


# ----------------------------------------------------------------------------- gray
def to_gray(img_u8):
    """(h,w,3) RGB uint8 -> gray as cv2.imread(png)+cvtColor(BGR2GRAY) gives (optical_flow.py:62-65);
    (h,w) passes through (an 'L' png is read as three equal channels and the formula is the identity)."""
    a = np.asarray(img_u8)
    if a.ndim == 2:
        return a.copy()
    r, g, b = (a[:, :, k].astype(np.int32) for k in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------- Shi-Tomasi
def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def sobel_scaled(gray):
    """cv2.Sobel(8U -> 32F, ksize 3, scale = 1/(4*7*255)), BORDER_REFLECT_101, with the float operation
    order of the AVX2/FMA3 build (verified bit-exact against cv2 4.13, tests/test_oracle_flow.py)."""
    p = np.pad(gray.astype(np.float32), 1, mode="reflect")
    scale = 1.0 / (4 * BLOCK * 255.0)
    s = np.float32(scale)
    s2 = np.float32(2.0 * scale)
    left, mid, right = p[:, :-2], p[:, 1:-1], p[:, 2:]
    rd = right - left
    S = np.full(rd[1:-1].shape, s, np.float32)
    dx = _fma32(rd[2:] + rd[:-2], S, rd[1:-1] * s2)
    Sr = np.full(mid.shape, s, np.float32)
    S2r = np.full(mid.shape, s2, np.float32)
    row = _fma32(right, Sr, _fma32(mid, S2r, left * s))
    # OpenCV's row filter handles 32 pixels per SIMD step with FMA; the last (w mod 32) columns go through
    # its scalar loop, which is the same sum without FMA contraction.
    tail = (gray.shape[1] // 32) * 32
    row[:, tail:] = ((left * s + mid * s2) + right * s)[:, tail:]
    dy = row[2:] - row[:-2]
    return dx, dy


def _box_sum(a):
    h, w = a.shape
    r = BLOCK // 2
    p = np.pad(a.astype(np.float64), r, mode="reflect")
    rows = np.zeros((h + 2 * r, w))
    for k in range(BLOCK):
        rows += p[:, k:k + w]
    out = np.zeros((h, w))
    for k in range(BLOCK):
        out += rows[k:k + h]
    return out.astype(np.float32)


def min_eig_map(gray):
    """cv2.cornerMinEigenVal(gray, 7, ksize=3): products in fp32, 7x7 box sums accumulated in fp64."""
    dx, dy = sobel_scaled(gray)
    a = _box_sum(dx * dx) * np.float32(0.5)
    b = _box_sum(dx * dy)
    c = _box_sum(dy * dy) * np.float32(0.5)
    return (a + c) - np.sqrt((a - c) * (a - c) + b * b)


def good_features(gray):
    """cv2.goodFeaturesToTrack -> float32 (n,2) array of (x,y), in OpenCV's order."""
    eig = min_eig_map(gray)
    h, w = eig.shape
    thr = np.float32(float(eig.max()) * QUALITY)
    eig = np.where(eig > thr, eig, np.float32(0))
    pad = np.pad(eig, 1, mode="constant", constant_values=-np.inf)
    dil = np.full_like(eig, -np.inf)
    for dy in range(3):
        for dx in range(3):
            dil = np.maximum(dil, pad[dy:dy + h, dx:dx + w])
    keep = (eig != 0) & (eig == dil)
    keep[0, :] = keep[-1, :] = False
    keep[:, 0] = keep[:, -1] = False
    ys, xs = np.nonzero(keep)
    vals = eig[ys, xs]
    addr = ys.astype(np.int64) * w + xs
    order = np.lexsort((-addr, -vals.astype(np.float64)))  # value desc, ties -> higher address first
    out = []
    for k in order:
        x, y = int(xs[k]), int(ys[k])
        ok = True
        for (px, py) in out:
            if (x - px) * (x - px) + (y - py) * (y - py) < MIN_DIST * MIN_DIST:
                ok = False
                break
        if ok:
            out.append((x, y))
            if len(out) == MAX_CORNERS:
                break
    return np.array(out, dtype=np.float32).reshape(-1, 2)


# ----------------------------------------------------------------------------- pyramid + derivatives
def pyr_down(img):
    """cv2.pyrDown for uint8: separable [1 4 6 4 1], integer, (sum + 128) >> 8, BORDER_REFLECT_101."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    p = np.pad(img.astype(np.int32), 2, mode="reflect")
    cols = 2 * np.arange(ow)
    rows = p[:, cols + 2] * 6 + (p[:, cols + 1] + p[:, cols + 3]) * 4 + p[:, cols] + p[:, cols + 4]
    r = 2 * np.arange(oh)
    out = rows[r + 2] * 6 + (rows[r + 1] + rows[r + 3]) * 4 + rows[r] + rows[r + 4]
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(gray):
    """buildOpticalFlowPyramid: levels stop when the NEXT level would be <= the window in either dimension."""
    levels = [gray]
    h, w = gray.shape
    for _ in range(MAX_LEVEL):
        w2, h2 = (w + 1) // 2, (h + 1) // 2
        if w2 <= WIN or h2 <= WIN:
            break
        levels.append(pyr_down(levels[-1]))
        h, w = h2, w2
    return levels


def scharr(img):
    """calcScharrDeriv: int16 (dx, dy), reflect101 at the image edge."""
    p = np.pad(img.astype(np.int32), 1, mode="reflect")
    t0 = (p[:-2] + p[2:]) * 3 + p[1:-1] * 10  # vertical smoothing
    t1 = p[2:] - p[:-2]  # vertical difference
    dx = t0[:, 2:] - t0[:, :-2]
    dy = (t1[:, 2:] + t1[:, :-2]) * 3 + t1[:, 1:-1] * 10
    return dx.astype(np.int16), dy.astype(np.int16)


def _cv_round(v):
    return int(np.rint(np.float32(v)))


def _weights(a, b):
    one = np.float32(1.0)
    sc = np.float32(1 << 14)
    a, b = np.float32(a), np.float32(b)
    w00 = _cv_round((one - a) * (one - b) * sc)
    w01 = _cv_round(a * (one - b) * sc)
    w10 = _cv_round((one - a) * b * sc)
    return w00, w01, w10, (1 << 14) - w00 - w01 - w10


def _bilinear(img, ix, iy, wts, shift):
    w00, w01, w10, w11 = wts
    a = img[iy:iy + WIN, ix:ix + WIN]
    b = img[iy:iy + WIN, ix + 1:ix + WIN + 1]
    c = img[iy + 1:iy + WIN + 1, ix:ix + WIN]
    d = img[iy + 1:iy + WIN + 1, ix + 1:ix + WIN + 1]
    return (a * w00 + b * w01 + c * w10 + d * w11 + (1 << (shift - 1))) >> shift


def pyr_lk(gray1, gray2, pts):
    """calcOpticalFlowPyrLK(gray1, gray2, pts, winSize=(50,50), maxLevel=2, (EPS|COUNT,10,0.03)).
    Returns (next_pts float32 (n,2), status uint8 (n,))."""
    f32 = np.float32
    pyr1, pyr2 = build_pyramid(gray1), build_pyramid(gray2)
    n = len(pts)
    nxt = np.zeros((n, 2), f32)
    status = np.ones(n, np.uint8)
    half = f32((WIN - 1) * 0.5)
    flt_scale = f32(1.0 / (1 << 20))
    eps2 = min(max(EPS, 0.0), 10.0)
    eps2 = eps2 * eps2  # criteria.epsilon *= criteria.epsilon, in double
    top = len(pyr1) - 1
    for level in range(top, -1, -1):
        I = np.pad(pyr1[level].astype(np.int32), WIN, mode="reflect")
        J = np.pad(pyr2[level].astype(np.int32), WIN, mode="reflect")
        dx, dy = scharr(pyr1[level])
        DX = np.pad(dx.astype(np.int32), WIN, mode="constant")
        DY = np.pad(dy.astype(np.int32), WIN, mode="constant")
        rows, cols = pyr1[level].shape
        inv = f32(1.0 / (1 << level))
        for k in range(n):
            prev = (f32(pts[k][0]) * inv, f32(pts[k][1]) * inv)
            if level == top:
                cur = prev
            else:
                cur = (nxt[k, 0] * f32(2), nxt[k, 1] * f32(2))
            nxt[k] = cur
            px, py = f32(prev[0] - half), f32(prev[1] - half)
            ipx, ipy = int(np.floor(px)), int(np.floor(py))
            if ipx < -WIN or ipx >= cols or ipy < -WIN or ipy >= rows:
                if level == 0:
                    status[k] = 0
                continue
            wts = _weights(f32(px - f32(ipx)), f32(py - f32(ipy)))
            Iw = _bilinear(I, ipx + WIN, ipy + WIN, wts, 9)
            Ix = _bilinear(DX, ipx + WIN, ipy + WIN, wts, 14)
            Iy = _bilinear(DY, ipx + WIN, ipy + WIN, wts, 14)
            A11 = f32(f32(int((Ix * Ix).sum())) * flt_scale)
            A12 = f32(f32(int((Ix * Iy).sum())) * flt_scale)
            A22 = f32(f32(int((Iy * Iy).sum())) * flt_scale)
            D = f32(f32(A11 * A22) - f32(A12 * A12))
            dd = f32(A11 - A22)
            min_eig = f32(f32(f32(A22 + A11) - np.sqrt(f32(f32(dd * dd) + f32(f32(f32(4) * A12) * A12))))
                          / f32(2 * WIN * WIN))
            if min_eig < f32(MIN_EIG_THRESHOLD) or D < np.finfo(np.float32).eps:
                if level == 0:
                    status[k] = 0
                continue
            D = f32(f32(1) / D)
            cx, cy = f32(cur[0] - half), f32(cur[1] - half)
            pdx = pdy = f32(0)
            for j in range(MAX_ITERS):
                icx, icy = int(np.floor(cx)), int(np.floor(cy))
                if icx < -WIN or icx >= cols or icy < -WIN or icy >= rows:
                    if level == 0:
                        status[k] = 0
                    break
                wj = _weights(f32(cx - f32(icx)), f32(cy - f32(icy)))
                diff = _bilinear(J, icx + WIN, icy + WIN, wj, 9) - Iw
                b1 = f32(f32(int((diff * Ix).sum())) * flt_scale)
                b2 = f32(f32(int((diff * Iy).sum())) * flt_scale)
                ddx = f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D)
                ddy = f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)
                cx, cy = f32(cx + ddx), f32(cy + ddy)
                nxt[k] = (f32(cx + half), f32(cy + half))
                if float(ddx) * float(ddx) + float(ddy) * float(ddy) <= eps2:
                    break
                if j > 0 and abs(float(f32(ddx + pdx))) < 0.01 and abs(float(f32(ddy + pdy))) < 0.01:
                    nxt[k, 0] = f32(nxt[k, 0] - f32(ddx * f32(0.5)))
                    nxt[k, 1] = f32(nxt[k, 1] - f32(ddy * f32(0.5)))
                    break
                pdx, pdy = ddx, ddy
            if level == 0 and status[k]:
                fx, fy = f32(nxt[k, 0] - half), f32(nxt[k, 1] - half)
                ifx, ify = int(np.floor(fx)), int(np.floor(fy))
                if ifx < -WIN or ifx >= cols or ify < -WIN or ify >= rows:
                    status[k] = 0
    return nxt, status


def vectors_from_tracks(p0, p1, status):
    """optical_flow.py:73-82: rows [x0, y0, x1-x0, y1-y0] in float32 for points with status 1."""
    rows = []
    for (a, b), (c, d), s in zip(p1, p0, status):
        if s == 1:
            rows.append([c, d, np.float32(a) - np.float32(c), np.float32(b) - np.float32(d)])
    return np.array(rows, dtype=np.float32).reshape(-1, 4)


def lucas_kanade_np(img1_u8, img2_u8):
    g1, g2 = to_gray(img1_u8), to_gray(img2_u8)
    p0 = good_features(g1)
    if len(p0) == 0:
        return np.zeros((0, 4), np.float32)
    p1, st = pyr_lk(g1, g2, p0)
    return vectors_from_tracks(p0, p1, st)


def lucas_kanade_cv2(img1_u8, img2_u8):
    """The reference's own call sequence on in-memory images (file round trip through PNG is lossless)."""
    import cv2

    def gray(a):
        a = np.asarray(a)
        bgr = cv2.cvtColor(a, cv2.COLOR_GRAY2BGR) if a.ndim == 2 else cv2.cvtColor(a, cv2.COLOR_RGB2BGR)
        return cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)

    g1, g2 = gray(img1_u8), gray(img2_u8)
    p0 = cv2.goodFeaturesToTrack(g1, mask=None, maxCorners=MAX_CORNERS, qualityLevel=QUALITY,
                                 minDistance=MIN_DIST, blockSize=BLOCK)
    if p0 is None:
        return np.zeros((0, 4), np.float32)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(g1, g2, p0, None, winSize=(WIN, WIN), maxLevel=MAX_LEVEL,
                                         criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, MAX_ITERS, EPS))
    return vectors_from_tracks(p0.reshape(-1, 2), p1.reshape(-1, 2), st.reshape(-1))
