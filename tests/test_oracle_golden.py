"""CPU: the numpy oracle against the golden vectors cv2 4.13.0 produced (tests/golden/make_golden.py), and against
live cv2 where importable.  This is what pins the oracle (the reference itself ships no fixtures)."""
import numpy as np
import pytest

from oracle import orb_np as O, pose_np as P, chain_np as N

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_geometry_tables():
    assert O.features_per_level(500) == [109, 90, 75, 63, 52, 44, 36, 31]
    assert O.features_per_level(2000) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert O.features_per_level(10000) == [2172, 1810, 1508, 1257, 1047, 873, 727, 606]
    assert O.level_sizes(1280, 1024) == [(1280, 1024), (1067, 853), (889, 711), (741, 593), (617, 494), (514, 412), (429, 343), (357, 286)]
    assert O.level_sizes(2448, 2048)[-1] == (683, 572)


def test_pyramid_fast_blur_vs_golden(golden):
    a = golden["frame0"]
    assert np.array_equal(O.resize_linear_exact(a, 400, 300), golden["level1"])
    assert np.array_equal(O.fast_detect(a), golden["fast0"])
    assert np.array_equal(O.gaussian_blur_7x7(a), golden["blur0"])


def test_orb_vs_golden_order_exact(golden):
    for i, name in enumerate(("frame0", "frame1")):
        r = O.orb_detect_and_compute(golden[name], int(golden["nfeatures"]))
        for k in ("pt", "size", "angle", "response", "octave", "desc"):
            assert np.array_equal(r[k], golden["f%d_%s" % (i, k)]), (name, k)


def test_matching_vs_golden(golden):
    d0, d1 = golden["f0_desc"], golden["f1_desc"]
    assert np.array_equal(P.sort_matches(P.bf_match_crosscheck(d0, d1)), golden["matches"])
    idx, dist = P.bf_knn2(d0, d1)
    assert np.array_equal(dist, golden["knn_dist"])
    assert np.array_equal(idx, golden["knn_idx"])


def test_ransac_and_pose_vs_golden(golden):
    m = golden["matches"]
    p1, p2 = golden["f0_pt"][m[:, 0]], golden["f1_pt"][m[:, 1]]
    E, mask = P.find_essential_mat(p1, p2, golden["K"])
    assert min(np.abs(E - golden["E"]).max(), np.abs(E + golden["E"]).max()) < 1e-6
    assert np.array_equal(mask, golden["ransac_mask"])
    good, R, t, pmask, _ = P.recover_pose(golden["E"], p1, p2, golden["K"])
    assert good == int(golden["good"]) and np.array_equal(pmask, golden["pose_mask"])
    assert np.abs(R - golden["R"]).max() < 1e-9 and np.abs(t - golden["t"]).max() < 1e-9


def test_ransac_heavy_vs_golden(golden):
    E, mask = P.find_essential_mat(golden["c5_p1"], golden["c5_p2"], golden["c5_K"])
    assert min(np.abs(E - golden["c5_E"]).max(), np.abs(E + golden["c5_E"]).max()) < 1e-6   # minimal-solver conditioning, see DESIGN.md
    assert np.array_equal(mask, golden["c5_mask"])
    good, R, t, pmask, _ = P.recover_pose(golden["c5_E"], golden["c5_p1"], golden["c5_p2"], golden["c5_K"])
    assert good == int(golden["c5_good"]) and np.array_equal(pmask, golden["c5_pose_mask"])
    assert np.abs(R - golden["c5_R"]).max() < 1e-9


def test_five_point_vs_golden(golden):
    x1 = P.normalize_points(golden["five_p1"], golden["five_K"])
    x2 = P.normalize_points(golden["five_p2"], golden["five_K"])
    mine = P.five_point(x1, x2)
    ref = golden["five_E"].reshape(-1, 3, 3)
    assert len(mine) == len(ref)
    for e in ref:
        assert min(min(np.abs(e - m).max(), np.abs(e + m).max()) for m in mine) < 1e-7


def test_edge_cases():
    # empty / tiny inputs
    assert P.bf_match_crosscheck(np.zeros((0, 32), np.uint8), np.zeros((3, 32), np.uint8)).shape == (0, 3)
    E, mask = P.find_essential_mat(np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32), np.eye(3))
    assert E is None and mask.shape == (3,)
    flat = np.full((200, 300), 127, np.uint8)     # no texture: no keypoints at all
    r = O.orb_detect_and_compute(flat, 100)
    assert len(r["pt"]) == 0 and r["desc"].shape == (0, 32)
    # retainBest: count <= n keeps order untouched; n == 0 clears; ties at the boundary are all kept
    assert list(O.retain_best(np.array([3, 1, 2], np.float32), 5)) == [0, 1, 2]
    assert len(O.retain_best(np.array([3, 1, 2], np.float32), 0)) == 0
    assert sorted(O.retain_best(np.array([5, 7, 5, 5, 1], np.float32), 2)) == [0, 1, 2, 3]


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_cv2_whole_chain():
    from droplet_visual_odometry_b200 import synth
    from oracle import cv2_chain as C
    frames, _, K = synth.render_sequence(2, width=640, height=480)
    a, b = frames[0].numpy(), frames[1].numpy()
    rc = C.frame_pair(a, b, K, 400)
    rn = N.frame_pair(a, b, K, 400)
    for k in ("pt", "angle", "response", "octave", "desc"):
        assert np.array_equal(rc["feats_prev"][k], rn["feats_prev"][k]), k
    assert np.array_equal(rc["matches"], rn["matches"])
    assert np.array_equal(rc["ransac_mask"], rn["ransac_mask"])
    assert min(np.abs(rc["E"] - rn["E"]).max(), np.abs(rc["E"] + rn["E"]).max()) < 1e-6
    assert np.abs(rc["R"] - rn["R"]).max() < 1e-8 and np.array_equal(rc["pose_mask"], rn["pose_mask"])
    # config-4 matcher composition
    assert np.array_equal(C.match_knn_ratio(rc["feats_prev"]["desc"], rc["feats_cur"]["desc"]),
                          P.sort_matches(P.ratio_and_reverse_check(rc["feats_prev"]["desc"], rc["feats_cur"]["desc"])))


def test_ingest_restatement_equals_cv2_gray_and_undistort():
    """oracle/ingest_np.py (what k_ingest is checked against on the GPU) == cv2.cvtColor + cv2.undistort, bit for bit, with
    the reference's two calibrations (/root/reference/Parameters/*.yaml values) and random images"""
    cv2 = pytest.importorskip("cv2")
    from oracle import ingest_np as I
    rng = np.random.default_rng(11)
    cases = [((1440, 1080), [1173.854081, 0, 747.788206, 0, 1170.565083, 574.700374, 0, 0, 1], [-0.296079, 0.099771, 0.000222, 0.000109, 0.0]),
             ((640, 480), [612.0, 0, 322.5, 0, 611.2, 238.1, 0, 0, 1], [0.11, -0.23, 0.001, -0.0007, 0.05]),
             ((333, 257), [300.0, 0, 160.0, 0, 305.0, 130.0, 0, 0, 1], [-0.35, 0.15, 0.0, 0.0, -0.03, 0.01, 0.002, 0.0005])]
    for (w, h), Kl, D in cases:
        K = np.array(Kl, np.float64).reshape(3, 3)
        D = np.array(D, np.float64)
        newK, _ = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 1, (w, h))
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        grey = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(I.bgr_to_gray(bgr), grey)
        ref = cv2.undistort(grey, K, D, None, newK)
        assert np.array_equal(I.ingest(bgr, K, D, newK), ref), (w, h)
        assert np.array_equal(I.ingest(grey, K, D, newK), ref)
