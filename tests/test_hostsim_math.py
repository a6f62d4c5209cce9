"""CPU: the __host__ build of the per-element code the CUDA kernels execute (csrc/mathcore.cuh, csrc/select.cuh)
against the oracle -- catches arithmetic mistakes before any GPU time is spent."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ptr
from oracle import orb_np as O, pose_np as P
from droplet_visual_odometry_b200 import synth


def test_retain_best_replay_matches_libstdcpp(hostsim):
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(1, 4000))
        k = int(rng.integers(0, n + 5))
        if trial % 3 == 0:
            resp = rng.integers(20, 60, n).astype(np.float32)          # heavy ties (FAST scores)
        elif trial % 3 == 1:
            resp = rng.random(n).astype(np.float32)                    # Harris-like
        else:
            resp = np.sort(rng.integers(20, 255, n)).astype(np.float32)[::(-1 if trial % 2 else 1)].copy()
        ref = O.retain_best(resp, k)
        out = np.empty(n + 1, np.int32)
        m = hostsim.hs_retain_best(ptr(resp), n, k, ptr(out))
        assert m == len(ref) and np.array_equal(out[:m], ref)


def test_retain_best_paired_formulation_matches_libstdcpp(hostsim):
    """the data-parallel (stopper-list) form of nth_element + partition that k_select runs, vs the real std:: calls"""
    rng = np.random.default_rng(7)
    for trial in range(200):
        n = int(rng.integers(1, 6000))
        k = int(rng.integers(0, n + 5))
        kind = trial % 4
        if kind == 0:
            resp = rng.integers(20, 60, n).astype(np.float32)
        elif kind == 1:
            resp = rng.random(n).astype(np.float32)
        elif kind == 2:
            resp = np.sort(rng.integers(20, 255, n)).astype(np.float32)[::(-1 if trial % 8 < 4 else 1)].copy()
        else:
            resp = np.full(n, 33.0, np.float32)
            resp[rng.integers(0, n, max(1, n // 50))] = 90.0
        ref = O.retain_best(resp, k)
        out = np.empty(n + 1, np.int32)
        for tail in (0, 3, 64, 100000):
            m = hostsim.hs_retain_best_paired(ptr(resp), n, k, tail, ptr(out))
            assert m == len(ref) and np.array_equal(out[:m], ref), (trial, n, k, tail)


def test_heap_select_fallback(hostsim):
    rng = np.random.default_rng(1)
    for trial in range(100):
        n = int(rng.integers(4, 2000))
        nth = int(rng.integers(0, n))
        resp = rng.integers(20, 80, n).astype(np.float32) if trial % 2 else rng.random(n).astype(np.float32)
        assert hostsim.hs_heap_select_check(ptr(resp), n, nth) == 0


def test_fast_score_map(hostsim, golden):
    img = np.ascontiguousarray(golden["frame0"][:200, :320])
    out = np.zeros_like(img)
    hostsim.hs_fast_score_map(ptr(img), 320, 200, 320, 20, ptr(out))
    assert np.array_equal(out, O.fast_score_map(img))


def test_fast_packed_prefilter_never_rejects_a_corner(hostsim, golden):
    import ctypes
    rng = np.random.default_rng(5)
    imgs = [np.ascontiguousarray(golden["frame0"][:240, :320]), rng.integers(0, 256, (120, 160)).astype(np.uint8),
            (rng.integers(0, 2, (120, 160)) * 41 + 100).astype(np.uint8)]
    for img in imgs:
        for t in (1, 20, 21, 60):
            passed = ctypes.c_int()
            h, w = img.shape
            assert hostsim.hs_fast_prefilter_check(ptr(img), w, h, w, t, ctypes.byref(passed)) == 0
            assert passed.value < img.size


def test_fast_atan2_and_harris(hostsim):
    rng = np.random.default_rng(2)
    ys = rng.integers(-60000, 60000, 5000).astype(np.float32)
    xs = rng.integers(-60000, 60000, 5000).astype(np.float32)
    ref = O.fast_atan2(ys, xs)
    mine = np.array([hostsim.hs_fast_atan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    assert np.array_equal(ref, mine)
    a = rng.integers(0, 50_000_000, 2000); b = rng.integers(0, 50_000_000, 2000); c = rng.integers(-30_000_000, 30_000_000, 2000)
    f = np.float32
    scale = f(1.0) / (f(28) * f(255.0))
    s4 = f(f(f(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(f), b.astype(f), c.astype(f)
    ref = (((af * bf) - (cf * cf)) - ((f(0.04) * (af + bf)) * (af + bf))) * s4
    mine = np.array([hostsim.hs_harris(int(x), int(y), int(z)) for x, y, z in zip(a, b, c)], f)
    assert np.array_equal(ref, mine)


def test_rng_and_iteration_rule(hostsim):
    r = np.empty(200, np.uint32)
    hostsim.hs_rng_stream(ptr(r), 200)
    g = P.CvRNG()
    assert all(int(r[i]) == g.next() for i in range(200))
    for n_, g_ in [(1500, 896), (400, 300), (3000, 1124), (100, 5), (100, 100), (2000, 7), (50, 49)]:
        for mx in (1000, 4096, 50):
            assert hostsim.hs_update_iters(0.999, (n_ - g_) / n_, 5, mx) == P.ransac_update_num_iters(0.999, (n_ - g_) / n_, 5, mx)


def test_five_point_solver(hostsim):
    rng = np.random.default_rng(3)
    worst = []
    for trial in range(120):
        p1, p2, K, R, t, _ = synth.synthetic_correspondences(5, 0.0, 0.3, seed=trial)
        if trial % 2:
            p2 = (p2 + rng.normal(0, 30, p2.shape)).astype(np.float32)
        x1 = np.ascontiguousarray(P.normalize_points(p1, K)); x2 = np.ascontiguousarray(P.normalize_points(p2, K))
        ref = P.five_point(x1, x2)
        models = np.zeros(90)
        n = hostsim.hs_five_point(ptr(x1), ptr(x2), ptr(models))
        assert n == len(ref)
        X1 = np.c_[x1, np.ones(5)]; X2 = np.c_[x2, np.ones(5)]
        for k in range(n):
            e = models[9 * k:9 * k + 9].reshape(3, 3)
            assert abs(np.linalg.norm(e) - 1) < 1e-12
            assert np.abs(np.sum(X2 * (X1 @ e.T), 1)).max() < 1e-10        # epipolar constraint on the 5 samples
            worst.append(min(np.abs(e - r).max() for r in ref + [-r for r in ref]))
    # agreement with the LAPACK-based oracle is limited by the conditioning of the hidden-variable polynomial
    assert np.median(worst) < 1e-9 and np.mean(np.array(worst) < 1e-5) > 0.9


def test_normalisation_is_cv2s_fused_form():
    """cv.findEssentialMat normalises with fma(p, 1/f, -(c/f)) (MatExpr alpha*A + beta through the FMA3 convertTo kernel).  Probe:
    a five-point call with K must be bit-identical to the call on pose_np.normalize_points' output with K = I; (p - c) / f is not."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    hits, plain_hits, n = 0, 0, 0
    for K in (synth.camera_matrix(1280, 1024), np.array([[1173.854081, 0, 747.788206], [0, 1170.565083, 574.700374], [0, 0, 1.0]])):
        for _ in range(30):
            p1 = rng.uniform(0, 1280, (5, 2)).astype(np.float32).astype(np.float64)
            p2 = (p1 + rng.normal(size=(5, 2)) * 8).astype(np.float32).astype(np.float64)
            E0 = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.999, threshold=1.0)[0]
            if E0 is None:
                continue
            n += 1
            E1 = cv2.findEssentialMat(P.normalize_points(p1, K), P.normalize_points(p2, K), np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=1.0)[0]
            hits += int(E1 is not None and E1.shape == E0.shape and np.array_equal(E0, E1))
            a = np.stack([(p1[:, 0] - K[0, 2]) / K[0, 0], (p1[:, 1] - K[1, 2]) / K[1, 1]], 1)
            b = np.stack([(p2[:, 0] - K[0, 2]) / K[0, 0], (p2[:, 1] - K[1, 2]) / K[1, 1]], 1)
            E2 = cv2.findEssentialMat(a, b, np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=1.0)[0]
            plain_hits += int(E2 is not None and E2.shape == E0.shape and np.array_equal(E0, E2))
    if hits != n:
        pytest.skip("this host's cv2 does not take the FMA3 convertTo path (%d of %d calls bit-identical)" % (hits, n))
    assert plain_hits < n // 4


def test_kernel_solver_works_in_cv2s_null_space_basis(hostsim):
    """The hidden-variable coordinates (x, y, z) of a solution are coordinates in the null-space basis, so the basis is observable:
    cv2 emits its solutions in the order cv::solvePoly finds the roots z of a polynomial that depends on the basis.  The numpy
    prototype of cv2's basis (Gram-Schmidt of cv::SVD's fixed +-1/9 fill-in vectors) + Durand-Kerner reproduces cv2's ORDER; the
    kernel solver's basis must span the same four vectors: every model the prototype finds, the kernel solver finds, with the
    same z (models are emitted by ascending z)."""
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "analysis"))
    import cv2_five_point_prototype as proto
    rng = np.random.default_rng(4)
    ordered = 0
    for trial in range(25):
        X = rng.uniform(-1, 1, (5, 3)) + np.array([0, 0, 4.0])
        R, _ = cv2.Rodrigues(rng.normal(size=3) * 0.1)
        t = rng.normal(size=3) * 0.3
        x1 = np.ascontiguousarray(X[:, :2] / X[:, 2:])
        Xc = (R @ X.T).T + t
        x2 = np.ascontiguousarray(Xc[:, :2] / Xc[:, 2:])
        Ecv = cv2.findEssentialMat(x1, x2, np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=1.0)[0]
        cvs = [Ecv[3 * i:3 * i + 3] for i in range(len(Ecv) // 3)]
        ms = proto.five_point_cvlike(x1, x2, "rowmajor", False)
        ordered += int(len(ms) == len(cvs) and all(min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-6 for a, b in zip(ms, cvs)))
        # the basis itself: the kernel's four vectors against the prototype's, up to rounding
        Q = np.array([[c * a, c * b, c, d * a, d * b, d, a, b, 1.0] for (a, b), (c, d) in zip(x1, x2)])
        want = proto.basis_cv(Q)
        got = np.zeros((4, 9))
        hostsim.hs_null_space(ptr(np.ascontiguousarray(Q)), ptr(got))
        assert np.abs(got - want).max() < 1e-9, trial
    # with any other basis the order agrees on ~5 % of the samples (2 of 40 for the transposed conventions); ill-conditioned
    # samples, where cv2's own root set moves with the last bit, keep this below 100 %
    assert ordered >= 20, "prototype reproduced cv2's solution order on %d of 25 samples" % ordered


def test_ransac_replay_with_kernel_solver_matches_golden(hostsim, golden):
    def solver(x1, x2):
        models = np.zeros(90)
        n = hostsim.hs_five_point(ptr(np.ascontiguousarray(x1)), ptr(np.ascontiguousarray(x2)), ptr(models))
        return [models[9 * k:9 * k + 9].reshape(3, 3).copy() for k in range(n)]
    E, mask = P.find_essential_mat(golden["c5_p1"], golden["c5_p2"], golden["c5_K"], solver=solver)
    assert np.array_equal(mask, golden["c5_mask"])
    assert min(np.abs(E - golden["c5_E"]).max(), np.abs(E + golden["c5_E"]).max()) < 1e-4


def test_pose_decomposition_and_cheirality(hostsim, golden):
    E = np.ascontiguousarray(golden["c5_E"])
    R1 = np.zeros(9); R2 = np.zeros(9); t = np.zeros(3)
    hostsim.hs_decompose(ptr(E), ptr(R1), ptr(R2), ptr(t))
    r1, r2, tt = P.decompose_essential(E)
    cands_ref = [(r1, tt), (r2, tt), (r1, -tt), (r2, -tt)]
    for R in (R1.reshape(3, 3), R2.reshape(3, 3)):
        assert abs(np.linalg.det(R) - 1) < 1e-12
        assert min(np.abs(R - c[0]).max() for c in cands_ref) < 1e-9
    assert min(np.abs(t - tt).max(), np.abs(t + tt).max()) < 1e-9
    # cheirality votes for the candidate cv2 chose reproduce cv2's mask
    Rg, tg = np.ascontiguousarray(golden["c5_R"]), np.ascontiguousarray(golden["c5_t"].ravel())
    x1 = P.normalize_points(golden["c5_p1"], golden["c5_K"]); x2 = P.normalize_points(golden["c5_p2"], golden["c5_K"])
    mine = np.array([hostsim.hs_cheirality(ptr(Rg), ptr(tg), float(a[0]), float(a[1]), float(b[0]), float(b[1]), 50.0)
                     for a, b in zip(x1, x2)], np.uint8) * 255
    assert (mine != golden["c5_pose_mask"]).sum() <= 1
    # Sampson error in the kernel's arithmetic == oracle's float32 vector
    err = P.sampson_errors(E, x1, x2)
    mine = np.array([hostsim.hs_sampson(ptr(E), float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(x1, x2)], np.float32)
    assert np.mean(mine == err) > 0.98 and np.allclose(mine, err, rtol=1e-6)   # numpy matmul sums in another order: 1-ulp float32 differences


def test_sampson_inlier_shortcut_equals_exact_decision(hostsim, golden):
    """the division-free fast path of sampson_inlier decides exactly like (float)(num/den) <= t32, also at the boundary"""
    hostsim.hs_sampson_inlier.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_float]
    E = np.ascontiguousarray(golden["c5_E"])
    x1 = P.normalize_points(golden["c5_p1"], golden["c5_K"]); x2 = P.normalize_points(golden["c5_p2"], golden["c5_K"])
    errs = np.array([hostsim.hs_sampson(ptr(E), float(a[0]), float(a[1]), float(b[0]), float(b[1])) for a, b in zip(x1, x2)], np.float32)
    # thresholds: the usual one, plus each sample's own error and its float neighbours (boundary cases)
    K = golden["c5_K"]
    t_usual = np.float32((1.0 / ((K[0, 0] + K[1, 1]) / 2)) ** 2)
    idx = np.random.default_rng(0).choice(len(x1), 300, replace=False)
    for i in idx:
        a, b = x1[i], x2[i]
        for t in (t_usual, errs[i], np.nextafter(errs[i], np.float32(0)), np.nextafter(errs[i], np.float32(1))):
            if not t > 0:
                continue
            got = hostsim.hs_sampson_inlier(ptr(E), float(a[0]), float(a[1]), float(b[0]), float(b[1]), float(t))
            assert got == int(errs[i] <= t)


def test_cheirality_pair_equals_two_separate_triangulations(hostsim, golden):
    """one triangulation serves (R, t) and (R, -t): same decisions as triangulating both"""
    hostsim.hs_cheirality_pair.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_double] * 5
    Rg, tg = np.ascontiguousarray(golden["c5_R"]), np.ascontiguousarray(golden["c5_t"].ravel())
    tn = np.ascontiguousarray(-tg)
    x1 = P.normalize_points(golden["c5_p1"], golden["c5_K"]); x2 = P.normalize_points(golden["c5_p2"], golden["c5_K"])
    bad = 0
    for a, b in zip(x1[:1500], x2[:1500]):
        args = (float(a[0]), float(a[1]), float(b[0]), float(b[1]), 50.0)
        both = hostsim.hs_cheirality_pair(ptr(Rg), ptr(tg), *args)
        sep = hostsim.hs_cheirality(ptr(Rg), ptr(tg), *args) | (hostsim.hs_cheirality(ptr(Rg), ptr(tn), *args) << 1)
        bad += int(both != sep)
    assert bad == 0


def test_fast_one_imad_ring_comparison_is_exact_for_every_pixel_centre_and_threshold(hostsim):
    """The device build of the exact FAST test folds `p > v + t` and `p < v - t` into one multiply-add per ring pixel;
    all 2^24 (v, t, p) combinations agree with the plain comparisons."""
    assert hostsim.hs_fast_ring_flags_check() == 0


def test_ring_has9_equals_cyclic_run_length_on_all_65536_masks(hostsim):
    assert hostsim.hs_ring_has9_check() == 0
