"""CPU: the __host__ build of the per-element code the CUDA kernels execute (csrc/mathcore.cuh, csrc/select.cuh)
against the oracle -- catches arithmetic mistakes before any GPU time is spent."""
import ctypes

import numpy as np

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
