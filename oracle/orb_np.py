"""ORACLE -- test infrastructure, never imported by the product path.

CPU restatement (numpy) of the ORB stages that ``cv2.ORB_create().detectAndCompute`` performs for the
reference's per-frame call ``VisualOdometry.compute_current_image_elements``
(/root/reference/scripts/visual_odometry_v3.py:370-379, detector created at :96).

The arithmetic itself lives in OpenCV (third-party, not vendored by the reference, not pinned by it; pinned for
this build to cv2 4.13.0 -- SURVEY.md §8c).  Every function below restates one stage of OpenCV's published ORB
algorithm as verified bit-for-bit against cv2 4.13.0; ``tests/test_oracle_golden.py`` and
the fixtures under ``tests/golden`` (made by ``tests/golden/make_golden.py`` from cv2 itself) pin it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SCALE_FACTOR = float(np.float32(1.2))  # cv2 stores 1.2f and widens it to double (A.0)
EDGE_THRESHOLD = 31
PATCH_SIZE = 31
HALF_PATCH = 15
FAST_THRESHOLD = 20
HARRIS_BLOCK = 7
HARRIS_K = np.float32(0.04)
UMAX = (15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3)
FAST_CIRCLE = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
               (-3, 0), (-3, 1), (-2, 2), (-1, 3))  # (dx, dy)

BRIEF_PATTERN = np.load(os.path.join(_HERE, "brief_pattern.npy")).astype(np.int32)  # (256, 4): x0 y0 x1 y1


# ---------------------------------------------------------------------------------------------- A.0 geometry
def layer_scales(nlevels: int = 8) -> np.ndarray:
    return np.array([np.float32(math.pow(SCALE_FACTOR, L)) for L in range(nlevels)], dtype=np.float32)


def level_sizes(width: int, height: int, nlevels: int = 8):
    out = []
    for s in layer_scales(nlevels):
        # cvRound(width / scale): float width / float scale in float32?  cv2: Size(cvRound(w*1/scale)...) uses
        # scale = 1/layerScale (float); verified equal to round-half-even of the double quotient on all sizes used.
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(width) * inv)), int(np.rint(np.float32(height) * inv))))
    return out


def features_per_level(nfeatures: int, nlevels: int = 8):
    factor = np.float32(1.0 / SCALE_FACTOR)
    nd = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(math.pow(float(factor), nlevels)))
    nd = np.float32(nd)
    quota, total = [], 0
    for _ in range(nlevels - 1):
        q = int(np.rint(nd))
        quota.append(q)
        total += q
        nd = np.float32(nd * factor)
    quota.append(max(nfeatures - total, 0))
    return quota


# ---------------------------------------------------------------------------------------------- A.1 pyramid
def _axis_coeffs(dst: int, src: int):
    scale = 1.0 / (dst / src)
    ofs = np.empty(dst, dtype=np.int64)
    c1 = np.empty(dst, dtype=np.int64)
    for d in range(dst):
        f = scale * (d + 0.5) - 0.5
        i = math.floor(f)
        if i < 0:
            ofs[d], c1[d] = 0, 0
        elif i >= src - 1:
            ofs[d], c1[d] = src - 1, 0
        else:
            ofs[d] = i
            c1[d] = int(np.rint((f - i) * 256.0))
    return ofs, c1


def resize_linear_exact(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=INTER_LINEAR_EXACT) for u8 single channel."""
    src_h, src_w = img.shape
    ox, cx1 = _axis_coeffs(dst_w, src_w)
    oy, cy1 = _axis_coeffs(dst_h, src_h)
    cx0, cy0 = 256 - cx1, 256 - cy1
    s = img.astype(np.int64)
    ox1 = np.minimum(ox + 1, src_w - 1)
    oy1 = np.minimum(oy + 1, src_h - 1)
    hrow = s[:, ox] * cx0[None, :] + s[:, ox1] * cx1[None, :]          # (src_h, dst_w), 8 fractional bits
    out = (hrow[oy, :] * cy0[:, None] + hrow[oy1, :] * cy1[:, None] + (1 << 15)) >> 16
    return out.astype(np.uint8)


def build_pyramid(img: np.ndarray, nlevels: int = 8):
    sizes = level_sizes(img.shape[1], img.shape[0], nlevels)
    levels = [np.ascontiguousarray(img)]
    for L in range(1, nlevels):
        levels.append(resize_linear_exact(levels[-1], sizes[L][0], sizes[L][1]))
    return levels


# ---------------------------------------------------------------------------------------------- A.2 FAST
def fast_score_map(img: np.ndarray, threshold: int = FAST_THRESHOLD) -> np.ndarray:
    """u8 map: FAST-9/16 corner score (m-1) where the pixel is a corner (m > threshold), else 0."""
    h, w = img.shape
    out = np.zeros((h, w), dtype=np.uint8)
    if h < 7 or w < 7:
        return out
    v = img[3:h - 3, 3:w - 3].astype(np.int16)
    d = np.stack([v - img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int16) for dx, dy in FAST_CIRCLE])
    dd = np.concatenate([d, d[:8]])           # cyclic
    best = np.full(v.shape, -32768, dtype=np.int16)
    for k in range(16):
        arc = dd[k:k + 9]
        best = np.maximum(best, np.maximum(arc.min(axis=0), (-arc).min(axis=0)))
    corner = best > threshold
    out[3:h - 3, 3:w - 3] = np.where(corner, best - 1, 0).astype(np.uint8)
    return out


def fast_nms(score: np.ndarray) -> np.ndarray:
    """Keep iff strictly greater than all 8 neighbours (non-corners are 0)."""
    h, w = score.shape
    p = np.zeros((h + 2, w + 2), dtype=np.uint8)
    p[1:-1, 1:-1] = score
    keep = score > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= score > p[1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx]
    return keep


def fast_detect(img: np.ndarray, threshold: int = FAST_THRESHOLD, border: int = 0):
    """Raster-ordered (x, y, score) int32 array, == cv2.FastFeatureDetector_create(threshold, True).detect,
    optionally culled to border <= x < W-border, border <= y < H-border (KeyPointsFilter::runByImageBorder)."""
    score = fast_score_map(img, threshold)
    keep = fast_nms(score)
    if border > 0:
        h, w = img.shape
        m = np.zeros_like(keep)
        if h > 2 * border and w > 2 * border:
            m[border:h - border, border:w - border] = True
        keep &= m
    ys, xs = np.nonzero(keep)    # row-major == raster order
    return np.stack([xs, ys, score[ys, xs].astype(np.int64)], axis=1).astype(np.int32)


# ---------------------------------------------------------------------------------------------- A.3 Harris
def harris_responses(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """float32 Harris response of the 7x7 block centred on each (x, y); float32 ops, no FMA."""
    I = img.astype(np.int32)
    r = HARRIS_BLOCK // 2
    a = np.zeros(len(xs), dtype=np.int64)
    b = np.zeros(len(xs), dtype=np.int64)
    c = np.zeros(len(xs), dtype=np.int64)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            y = ys + dy
            x = xs + dx
            ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    f = np.float32
    scale = f(1.0) / (f(4 * HARRIS_BLOCK) * f(255.0))
    s4 = f(f(f(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(np.float32), b.astype(np.float32), c.astype(np.float32)
    t1 = af * bf
    t2 = cf * cf
    sm = af + bf
    t3 = (HARRIS_K * sm) * sm
    return (((t1 - t2) - t3) * s4).astype(np.float32)


# ---------------------------------------------------------------------------------------------- A.4 selection
_retain_lib = None


def _lib():
    global _retain_lib
    if _retain_lib is None:
        path = os.path.join(_HERE, "_build", "liboracle_retain.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _retain_lib = ctypes.CDLL(path)
        _retain_lib.oracle_retain_best.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _retain_lib.oracle_retain_best.restype = ctypes.c_int
    return _retain_lib


def retain_best(response: np.ndarray, n_points: int) -> np.ndarray:
    """Indices (into ``response``) that cv::KeyPointsFilter::retainBest keeps, in the order it leaves them."""
    resp = np.ascontiguousarray(response, dtype=np.float32)
    out = np.empty(max(len(resp), 1), dtype=np.int32)
    k = _lib().oracle_retain_best(resp.ctypes.data, len(resp), int(n_points), out.ctypes.data)
    return out[:k].copy()


def retain_best_set(response: np.ndarray, n_points: int) -> np.ndarray:
    """Same SET as retain_best, by definition (k-th largest threshold, ties kept); order = input order."""
    response = np.asarray(response)
    if len(response) <= n_points:
        return np.arange(len(response))
    if n_points == 0:
        return np.arange(0)
    thr = np.sort(response)[::-1][n_points - 1]
    return np.nonzero(response >= thr)[0]


# ---------------------------------------------------------------------------------------------- A.5 IC angle
def fast_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """cv::fastAtan2 (degrees) on float32 arrays, float32 arithmetic, no FMA."""
    f = np.float32
    s = f(57.29577951308232)
    p1 = f(f(0.9997878412794807) * s)
    p3 = f(f(-0.3258083974640975) * s)
    p5 = f(f(0.1555786518463281) * s)
    p7 = f(f(-0.04432655554792128) * s)
    eps = f(2.220446049250313e-16)
    y = np.asarray(y, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    ax, ay = np.abs(x), np.abs(y)
    swap = ax < ay
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(swap, ax / (ay + eps), ay / (ax + eps)).astype(np.float32)
    c2 = (c * c).astype(np.float32)
    a = ((((p7 * c2).astype(np.float32) + p5) * c2 + p3).astype(np.float32) * c2 + p1).astype(np.float32) * c
    a = a.astype(np.float32)
    a = np.where(swap, f(90.0) - a, a).astype(np.float32)
    a = np.where(x < 0, f(180.0) - a, a).astype(np.float32)
    a = np.where(y < 0, f(360.0) - a, a).astype(np.float32)
    return a


def ic_angles(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    I = img.astype(np.int64)
    m10 = np.zeros(len(xs), dtype=np.int64)
    m01 = np.zeros(len(xs), dtype=np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * I[ys, xs + u]
    for v in range(1, HALF_PATCH + 1):
        d = UMAX[v]
        vsum = np.zeros(len(xs), dtype=np.int64)
        for u in range(-d, d + 1):
            vp = I[ys + v, xs + u]
            vm = I[ys - v, xs + u]
            vsum += vp - vm
            m10 += u * (vp + vm)
        m01 += v * vsum
    return fast_atan2(m01.astype(np.float32), m10.astype(np.float32))


# ---------------------------------------------------------------------------------------------- A.6 blur
def gaussian_kernel_7_2() -> np.ndarray:
    """cv2.getGaussianKernel(7, 2, CV_32F): computed in double, normalised, stored as float32."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    k = k / k.sum()
    return k.astype(np.float32)


def _fma32(a, b, c):
    """float32 fused multiply-add of arrays.  Here a is a 24-bit kernel tap, b an 8-bit pixel or a 24-bit row
    value and c a float32 of comparable magnitude, so a*b (<= 48 significant bits) and a*b + c (<= 53) are both
    EXACT in float64; the final cast is then the single rounding a hardware FMA performs."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def gaussian_blur_7x7(img: np.ndarray) -> np.ndarray:
    """The blur ORB applies before descriptors: separable 7-tap sigma=2 float32 filter with FMA accumulation,
    BORDER_REFLECT_101, result rounded half-to-even to u8 (== cv2.sepFilter2D(img,-1,k32,k32,REFLECT_101))."""
    k = gaussian_kernel_7_2()
    h, w = img.shape
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    # row pass: s = k0*p[-3]; s = fma(k_i, p_i, s) for i = 1..6
    rows = (k[0] * p[:, 0:w]).astype(np.float32)
    for i in range(1, 7):
        rows = _fma32(np.full_like(rows, k[i]), p[:, i:i + w], rows)
    # column pass on float rows: s = k3*c; s = fma(k_{3+d}, (r[+d] + r[-d]), s)
    s = (k[3] * rows[3:3 + h]).astype(np.float32)
    for d in (1, 2, 3):
        pair = (rows[3 + d:3 + d + h] + rows[3 - d:3 - d + h]).astype(np.float32)
        s = _fma32(np.full_like(s, k[3 + d]), pair, s)
    return np.clip(np.rint(s), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------- A.7 rBRIEF
def brief_descriptors(blurred: np.ndarray, xs: np.ndarray, ys: np.ndarray, angles_deg: np.ndarray) -> np.ndarray:
    """(N, 32) u8 steered-BRIEF descriptors for integer level coordinates (xs, ys) on the BLURRED level."""
    f = np.float32
    ang = angles_deg.astype(np.float32) * f(np.pi / f(180.0))
    a = np.cos(ang.astype(np.float64)).astype(np.float32)
    b = np.sin(ang.astype(np.float64)).astype(np.float32)
    n = len(xs)
    desc = np.zeros((n, 32), dtype=np.uint8)
    pat = BRIEF_PATTERN.astype(np.float32)

    def value(px, py):
        x = ((px[None, :] * a[:, None]).astype(np.float32) - (py[None, :] * b[:, None]).astype(np.float32)).astype(np.float32)
        y = ((px[None, :] * b[:, None]).astype(np.float32) + (py[None, :] * a[:, None]).astype(np.float32)).astype(np.float32)
        ix = np.rint(x).astype(np.int64)
        iy = np.rint(y).astype(np.int64)
        return blurred[ys[:, None] + iy, xs[:, None] + ix]

    t0 = value(pat[:, 0], pat[:, 1])
    t1 = value(pat[:, 2], pat[:, 3])
    bits = (t0 < t1).astype(np.uint8).reshape(n, 32, 8)
    for j in range(8):
        desc |= (bits[:, :, j] << j).astype(np.uint8)
    return desc


# ---------------------------------------------------------------------------------------------- whole ORB
def orb_detect_and_compute(img: np.ndarray, nfeatures: int = 500, nlevels: int = 8, order_exact: bool = True):
    """Restatement of cv2.ORB_create(nfeatures).detectAndCompute(img, None).

    Returns dict with: pt (N,2) f32, size (N,) f32, angle (N,) f32, response (N,) f32, octave (N,) i32,
    lvl_xy (N,2) i32 integer level coordinates, desc (N,32) u8, and per-level intermediates under 'levels'.
    """
    levels = build_pyramid(img, nlevels)
    scales = layer_scales(nlevels)
    quota = features_per_level(nfeatures, nlevels)
    keep_fn = retain_best if order_exact else retain_best_set
    out = {k: [] for k in ("pt", "size", "angle", "response", "octave", "lvl_xy")}
    dbg = []
    for L, lev in enumerate(levels):
        cand = fast_detect(lev, FAST_THRESHOLD, EDGE_THRESHOLD)
        idx1 = keep_fn(cand[:, 2].astype(np.float32), 2 * quota[L])
        c1 = cand[idx1]
        resp = harris_responses(lev, c1[:, 0].astype(np.int64), c1[:, 1].astype(np.int64)) if len(c1) else np.zeros(0, np.float32)
        idx2 = keep_fn(resp, quota[L])
        c2 = c1[idx2]
        r2 = resp[idx2]
        xs, ys = c2[:, 0].astype(np.int64), c2[:, 1].astype(np.int64)
        ang = ic_angles(lev, xs, ys) if len(c2) else np.zeros(0, np.float32)
        s = scales[L]
        out["pt"].append(np.stack([xs.astype(np.float32) * s, ys.astype(np.float32) * s], 1).astype(np.float32).reshape(-1, 2))
        out["size"].append(np.full(len(c2), np.float32(PATCH_SIZE) * s, dtype=np.float32))
        out["angle"].append(ang)
        out["response"].append(r2)
        out["octave"].append(np.full(len(c2), L, dtype=np.int32))
        out["lvl_xy"].append(c2[:, :2].astype(np.int32).reshape(-1, 2))
        dbg.append({"candidates": cand, "after_fast_cut": c1, "harris": resp})
    res = {k: np.concatenate(v) for k, v in out.items()}
    descs = []
    for L, lev in enumerate(levels):
        m = res["octave"] == L
        if not m.any():
            continue
        blurred = gaussian_blur_7x7(lev)
        inv = np.float32(1.0) / scales[L]
        cx = np.rint(res["pt"][m, 0] * inv).astype(np.int64)
        cy = np.rint(res["pt"][m, 1] * inv).astype(np.int64)
        descs.append(brief_descriptors(blurred, cx, cy, res["angle"][m]))
    res["desc"] = np.concatenate(descs) if descs else np.zeros((0, 32), np.uint8)
    res["levels"] = dbg
    res["pyramid"] = levels
    return res
