"""ORACLE -- test infrastructure, never imported by the product path.

CPU restatement (numpy, float64) of the pair-level stages the reference runs through OpenCV:

* ``cv.BFMatcher(NORM_HAMMING, crossCheck=True).match`` + ``sorted(key=distance)``
  (/root/reference/scripts/visual_odometry_v3.py:75, :219, :221) and the kNN(k=2) + ratio-0.75 variant (:203, :227);
* ``cv.findEssentialMat(p_prev, p_cur, K, RANSAC, prob=0.999, threshold=1.0)`` (:297-300);
* ``cv.recoverPose(E, p_prev, p_cur, K)`` (:303-306).

OpenCV (third-party, unpinned by the reference, pinned here to cv2 4.13.0) holds the arithmetic.  What is restated is
its published algorithm and its observable control flow: RNG stream, 5-point sampling
with single-index redraw, Nister 5-point minimal solver, Sampson error cast to float32, strict '>' update, adaptive
iteration count, SVD decomposition, linear triangulation and the '>=' cheirality cascade.  Pinned against cv2 itself by
``tests/test_oracle_golden.py`` and the fixtures under ``tests/golden``.
"""
from __future__ import annotations

import math

import numpy as np

DBL_MIN = 2.2250738585072014e-308


# ------------------------------------------------------------------------------------------------ A.8 matching
_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint16)


def hamming_matrix(d1: np.ndarray, d2: np.ndarray) -> np.ndarray:
    """(N1, N2) int32 Hamming distances between rows of two (N, 32) u8 descriptor arrays."""
    out = np.empty((len(d1), len(d2)), dtype=np.int32)
    step = max(1, (1 << 24) // max(1, len(d2) * 32))
    for s in range(0, len(d1), step):
        x = d1[s:s + step, None, :] ^ d2[None, :, :]
        out[s:s + step] = _POP8[x].sum(axis=2)
    return out


def bf_match_crosscheck(d1: np.ndarray, d2: np.ndarray):
    """cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(d1, d2): (queryIdx, trainIdx, distance) ordered by
    queryIdx; mutual nearest neighbours with lowest-index tie-breaking on both sides."""
    if len(d1) == 0 or len(d2) == 0:
        return np.zeros((0, 3), dtype=np.int32)
    D = hamming_matrix(d1, d2)
    fwd = D.argmin(axis=1)          # first minimum = lowest index
    bwd = D.argmin(axis=0)
    q = np.arange(len(d1))
    ok = bwd[fwd] == q
    return np.stack([q[ok], fwd[ok], D[q[ok], fwd[ok]]], axis=1).astype(np.int32)


def bf_knn2(d1: np.ndarray, d2: np.ndarray):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(d1, d2, k=2): per query the two smallest by (distance, trainIdx).
    Returns (idx (N1,2) int32, dist (N1,2) int32)."""
    D = hamming_matrix(d1, d2)
    key = D.astype(np.int64) * (D.shape[1] + 1) + np.arange(D.shape[1])[None, :]
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    return order.astype(np.int32), np.take_along_axis(D, order, axis=1).astype(np.int32)


def ratio_and_reverse_check(d1: np.ndarray, d2: np.ndarray, ratio: float = 0.75):
    """Config-4 matcher: knn k=2 + Lowe ratio (m.distance < 0.75 * n.distance, float32 distances as cv2 stores them)
    + reverse 1-NN check.  (queryIdx, trainIdx, distance) ordered by queryIdx."""
    idx, dist = bf_knn2(d1, d2)
    D = hamming_matrix(d1, d2)
    bwd = D.argmin(axis=0)
    q = np.arange(len(d1))
    # cv2 DMatch.distance is float32; python compares m.distance < 0.75*n.distance in double
    ok = dist[:, 0].astype(np.float64) < ratio * dist[:, 1].astype(np.float64)
    ok &= bwd[idx[:, 0]] == q
    return np.stack([q[ok], idx[ok, 0], dist[ok, 0]], axis=1).astype(np.int32)


def sort_matches(m: np.ndarray) -> np.ndarray:
    """Python's stable sorted(matches, key=distance): order (distance, queryIdx)."""
    return m[np.argsort(m[:, 2], kind="stable")]


# ------------------------------------------------------------------------------------------------ 5-point solver
# monomial order for cubic polynomials in (x, y, z) -- Nister's: the first ten are eliminated
_MONO = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]
_MONO_IDX = {m: i for i, m in enumerate(_MONO)}


class _Poly:
    """Sparse polynomial in x, y, z: dict exponent-tuple -> coeff."""
    __slots__ = ("c",)

    def __init__(self, c=None):
        self.c = c or {}

    def __add__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) + v
        return _Poly(r)

    def __sub__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) - v
        return _Poly(r)

    def __mul__(self, o):
        if not isinstance(o, _Poly):
            return _Poly({k: v * o for k, v in self.c.items()})
        r = {}
        for k1, v1 in self.c.items():
            for k2, v2 in o.c.items():
                k = (k1[0] + k2[0], k1[1] + k2[1], k1[2] + k2[2])
                r[k] = r.get(k, 0.0) + v1 * v2
        return _Poly(r)

    def vec20(self):
        out = np.zeros(20)
        for k, v in self.c.items():
            out[_MONO_IDX[k]] = v
        return out


def null_space_5x9(Q: np.ndarray) -> np.ndarray:
    """Orthonormal basis (4, 9) of the right null space of the 5x9 epipolar constraint matrix (rows of Vt)."""
    _, _, vt = np.linalg.svd(Q, full_matrices=True)
    return vt[5:9]


def five_point_constraints(EE: np.ndarray) -> np.ndarray:
    """10x20 coefficient matrix of det(E)=0 and 2 E E^T E - tr(E E^T) E = 0 for E = x X + y Y + z Z + W."""
    X, Y, Z, W = (EE[i].reshape(3, 3) for i in range(4))
    E = [[_Poly({(1, 0, 0): X[i, j], (0, 1, 0): Y[i, j], (0, 0, 1): Z[i, j], (0, 0, 0): W[i, j]}) for j in range(3)]
         for i in range(3)]
    EEt = [[E[i][0] * E[j][0] + E[i][1] * E[j][1] + E[i][2] * E[j][2] for j in range(3)] for i in range(3)]
    half_tr = (EEt[0][0] + EEt[1][1] + EEt[2][2]) * 0.5
    L = [[EEt[i][j] - half_tr if i == j else EEt[i][j] for j in range(3)] for i in range(3)]
    rows = []
    for i in range(3):
        for j in range(3):
            rows.append((L[i][0] * E[0][j] + L[i][1] * E[1][j] + L[i][2] * E[2][j]).vec20())
    det = (E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1])
           - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0])
           + E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]))
    rows.append(det.vec20())
    return np.array(rows)


_MONO_ARR = np.array(_MONO, dtype=np.float64)


def _polish_xyz(A, x, y, z, iters=3):
    """Gauss-Newton on the ten cubic constraints A . monomials(x, y, z) = 0.  The degree-10 hidden-variable
    polynomial loses digits when roots cluster; the original system is well conditioned, and a few steps bring every
    solution to the ~1e-13 agreement with cv2's solver that the RANSAC replay relies on."""
    v = np.array([x, y, z], dtype=np.float64)
    e = _MONO_ARR
    for _ in range(iters):
        with np.errstate(all="ignore"):
            mon = v[0] ** e[:, 0] * v[1] ** e[:, 1] * v[2] ** e[:, 2]
            J = np.empty((20, 3))
            for d in range(3):
                ed = e.copy()
                ed[:, d] = np.maximum(ed[:, d] - 1, 0)
                J[:, d] = e[:, d] * (v[0] ** ed[:, 0] * v[1] ** ed[:, 1] * v[2] ** ed[:, 2])
        r = A @ mon
        Jr = A @ J
        try:
            step = np.linalg.solve(Jr.T @ Jr, Jr.T @ r)
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(step)):
            break
        v = v - step
    return v[0], v[1], v[2]


def five_point(x1: np.ndarray, x2: np.ndarray, polish: bool = False):
    """Nister 5-point minimal solver on normalised coordinates: all real E (3x3, ||E||_F = 1) with x2^T E x1 = 0."""
    Q = np.empty((5, 9))
    for i in range(5):
        a, b = x1[i]
        c, d = x2[i]
        Q[i] = [c * a, c * b, c, d * a, d * b, d, a, b, 1.0]
    EE = null_space_5x9(Q)
    A = five_point_constraints(EE)
    try:
        Bm = np.linalg.solve(A[:, :10], A[:, 10:])     # reduced rows: monomial_i + Bm[i] . [xz2 xz x yz2 yz y z3 z2 z 1]
    except np.linalg.LinAlgError:
        return []
    # rows for x^2 z (4), x^2 (5), y^2 z (6), y^2 (7), xyz (8), xy (9):  k = e - z f etc.
    def row_minus_z(e, f):
        # polynomials in z multiplying x, y, 1:  e: [xz2 xz x | yz2 yz y | z3 z2 z 1]
        re, rf = Bm[e], Bm[f]
        px = np.array([0.0, re[0], re[1], re[2]]) - np.array([rf[0], rf[1], rf[2], 0.0])          # z^3..z^0
        py = np.array([0.0, re[3], re[4], re[5]]) - np.array([rf[3], rf[4], rf[5], 0.0])
        p1 = np.array([0.0, re[6], re[7], re[8], re[9]]) - np.array([rf[6], rf[7], rf[8], rf[9], 0.0])  # z^4..z^0
        return px, py, p1
    B = [row_minus_z(4, 5), row_minus_z(6, 7), row_minus_z(8, 9)]
    pm = np.polymul
    det = (pm(pm(B[0][0], B[1][1]) - pm(B[0][1], B[1][0]), B[2][2])
           + pm(pm(B[0][1], B[1][2]), B[2][0]) - pm(pm(B[0][2], B[1][1]), B[2][0])
           + pm(pm(B[0][2], B[1][0]), B[2][1]) - pm(pm(B[0][0], B[1][2]), B[2][1]))
    det = np.atleast_1d(det)
    if len(det) < 11:
        det = np.concatenate([np.zeros(11 - len(det)), det])
    if not np.all(np.isfinite(det)) or det[0] == 0:
        return []
    roots = np.roots(det)
    ddet = np.polyder(det)
    sols = []
    for r in roots:
        if abs(r.imag) > 1e-10:
            continue
        z = r.real
        for _ in range(3):      # Newton polish on the degree-10 polynomial (np.roots is a companion-matrix eigensolve)
            dv = np.polyval(ddet, z)
            if dv == 0:
                break
            z = z - np.polyval(det, z) / dv
        Bz = np.array([[np.polyval(B[i][0], z), np.polyval(B[i][1], z), np.polyval(B[i][2], z)] for i in range(3)])
        _, _, vt = np.linalg.svd(Bz)
        xy1 = vt[2]
        if abs(xy1[2]) < 1e-10:
            continue
        x, y = xy1[0] / xy1[2], xy1[1] / xy1[2]
        if polish:
            x, y, z = _polish_xyz(A, x, y, z)
        Ev = x * EE[0] + y * EE[1] + z * EE[2] + EE[3]
        Ev = Ev / np.linalg.norm(Ev)
        sols.append((z, Ev.reshape(3, 3)))
    sols.sort(key=lambda s: s[0])
    return [s[1] for s in sols]


# ------------------------------------------------------------------------------------------------ A.9 RANSAC
class CvRNG:
    """cv::RNG: multiply-with-carry, state seeded to 2^64-1 by RANSACPointSetRegistrator."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, n):
        return self.next() % n


def sample_stream(count: int, n_iters: int, model_points: int = 5) -> np.ndarray:
    """(n_iters, 5) positions cv2's RANSAC would draw for iterations 0..n_iters-1 with ``count`` correspondences."""
    rng = CvRNG()
    out = np.empty((n_iters, model_points), dtype=np.int32)
    for it in range(n_iters):
        idx = []
        while len(idx) < model_points:
            v = rng.uniform(count)
            if v in idx:
                continue
            idx.append(v)
        out[it] = idx
    return out


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - math.pow(1.0 - ep, model_points)
    if denom < DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def _fma(a: np.ndarray, b: float, c: float) -> np.ndarray:
    """Correctly rounded a * b + c in float64 (math.fma needs Python 3.13): the exact product a * b = hi + lo by Dekker's split,
    then a compensated sum -- exact to the last bit for the magnitudes met here (coordinates of a few thousand at most)."""
    a = np.asarray(a, dtype=np.float64)
    hi = a * b
    split = 134217729.0
    a1 = a * split
    a1 = a1 - (a1 - a)
    a2 = a - a1
    b1 = b * split
    b1 = b1 - (b1 - b)
    b2 = b - b1
    lo = ((a1 * b1 - hi) + a1 * b2 + a2 * b1) + a2 * b2
    s = hi + c
    bb = s - hi
    err = (hi - (s - bb)) + (c - bb)
    return s + (err + lo)


def normalize_points(p: np.ndarray, K: np.ndarray) -> np.ndarray:
    """cv.findEssentialMat's K-normalisation bit for bit: OpenCV evaluates `(col - cx) / fx` as alpha * p + beta with
    alpha = 1 / fx, beta = -cx * alpha, and its AVX2/FMA3 convertTo kernel fuses the multiply-add (checked against cv2 4.13.0:
    five-point calls on coordinates normalised this way are bit-identical to calls with K, tests/test_oracle_golden.py)."""
    p = np.asarray(p, dtype=np.float64)
    ax, ay = 1.0 / K[0, 0], 1.0 / K[1, 1]
    return np.stack([_fma(p[:, 0], ax, -(K[0, 2] * ax)), _fma(p[:, 1], ay, -(K[1, 2] * ay))], axis=1)


def sampson_errors(E: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """float32 error vector exactly as cv2's EMEstimatorCallback::computeError (float64 math, cast at the end)."""
    X1 = np.concatenate([x1, np.ones((len(x1), 1))], axis=1)
    X2 = np.concatenate([x2, np.ones((len(x2), 1))], axis=1)
    Ex1 = X1 @ E.T
    Etx2 = X2 @ E
    x2tEx1 = np.sum(X2 * Ex1, axis=1)
    den = Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return (x2tEx1 * x2tEx1 / den).astype(np.float32)


def find_essential_mat(p1, p2, K, prob=0.999, threshold=1.0, max_iters=1000, solver=five_point, return_trace=False,
                       exhaustive=False):
    """cv2.findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters) -> (E 3x3 or None, mask (N,) u8).
    ``exhaustive``: never shrink the iteration count (all max_iters hypotheses are scored, first best wins) -- the
    "all hypotheses scored" mode of BASELINE configs[4]; cv2 itself cannot be made to do this (it asserts prob < 1)."""
    x1 = normalize_points(p1, K)
    x2 = normalize_points(p2, K)
    n = len(x1)
    thr = threshold / ((K[0, 0] + K[1, 1]) / 2.0)
    t32 = np.float32(thr * thr)
    if n < 5:
        return (None, np.zeros(n, np.uint8)) + (({},) if return_trace else ())
    rng = CvRNG()
    niters = max(max_iters, 1)
    max_good = 0
    best_E, best_mask = None, np.zeros(n, np.uint8)
    it = 0
    trace = {"iters_run": 0, "models_scored": 0, "best_iter": -1, "best_model": -1}
    while it < niters:
        idx = []
        while len(idx) < 5:
            v = rng.uniform(n)
            if v in idx:
                continue
            idx.append(v)
        models = solver(x1[idx], x2[idx])
        for mi, E in enumerate(models):
            err = sampson_errors(E, x1, x2)
            mask = err <= t32
            good = int(mask.sum())
            trace["models_scored"] += 1
            if good > max(max_good, 4):
                best_E, best_mask, max_good = E, mask.astype(np.uint8), good
                if not exhaustive:
                    niters = ransac_update_num_iters(prob, (n - good) / n, 5, niters)
                trace["best_iter"], trace["best_model"] = it, mi
        it += 1
    trace["iters_run"] = it
    return (best_E, best_mask) + ((trace,) if return_trace else ())


# ------------------------------------------------------------------------------------------------ A.10 recoverPose
def decompose_essential(E: np.ndarray):
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


def triangulate_dlt(P0: np.ndarray, P1: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """cv2.triangulatePoints: per point the right singular vector of the 4x4 DLT matrix for the smallest s.v."""
    n = len(x1)
    A = np.empty((n, 4, 4))
    A[:, 0] = x1[:, 0:1] * P0[2] - P0[0]
    A[:, 1] = x1[:, 1:2] * P0[2] - P0[1]
    A[:, 2] = x2[:, 0:1] * P1[2] - P1[0]
    A[:, 3] = x2[:, 1:2] * P1[2] - P1[1]
    _, _, vt = np.linalg.svd(A)
    return vt[:, 3, :].T        # (4, n)


def recover_pose(E, p1, p2, K, distance_thresh=50.0):
    """cv2.recoverPose(E, p1, p2, K) -> (good, R, t (3,1), mask (N,) u8 0/255, candidate index 0..3)."""
    x1 = normalize_points(p1, K)
    x2 = normalize_points(p2, K)
    R1, R2, t = decompose_essential(np.asarray(E, dtype=np.float64))
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    masks, goods = [], []
    for R, tt in cands:
        P = np.hstack([R, tt.reshape(3, 1)])
        Q = triangulate_dlt(P0, P, x1, x2)
        with np.errstate(divide="ignore", invalid="ignore"):
            m = (Q[2] * Q[3]) > 0
            Qn = Q / Q[3]
            m &= Qn[2] < distance_thresh
            Q2 = P @ Qn
            m &= (Q2[2] > 0) & (Q2[2] < distance_thresh)
        masks.append(m)
        goods.append(int(m.sum()))
    g = goods
    if g[0] >= g[1] and g[0] >= g[2] and g[0] >= g[3]:
        k = 0
    elif g[1] >= g[0] and g[1] >= g[2] and g[1] >= g[3]:
        k = 1
    elif g[2] >= g[0] and g[2] >= g[1] and g[2] >= g[3]:
        k = 2
    else:
        k = 3
    R, tt = cands[k]
    return g[k], R, tt.reshape(3, 1), (masks[k].astype(np.uint8) * 255), k
