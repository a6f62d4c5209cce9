"""CPU: host-side mirror of the reference interface -- pose algebra, text formats, partitioning, chaining."""
import math
import os

import numpy as np
import pytest

from droplet_visual_odometry_b200 import pose_estimation_module as PEM, transformations_lite as T, sequence as S
from droplet_visual_odometry_b200._native import POSE_DTYPE


def rx(a): return np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
def ry(a): return np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
def rz(a): return np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])


def test_euler_conventions():
    a, b, c = 0.3, -0.2, 0.5
    assert np.allclose(T.euler_matrix(a, b, c, "sxyz")[:3, :3], rz(c) @ ry(b) @ rx(a))          # visual_odometry_v3.py:140
    assert np.allclose(T.euler_matrix(a, b, c, "rxyz")[:3, :3], rx(a) @ ry(b) @ rz(c))
    M = np.eye(4); M[:3, :3] = rx(a) @ ry(b) @ rz(c)
    assert np.allclose(T.euler_from_matrix(M, "rxyz"), (a, b, c))                              # :334
    for axes in ("sxyz", "rxyz", "szyx", "rzxz", "syxz"):
        M = T.euler_matrix(a, b, c, axes)      # compare matrices: repeated-axis conventions return the equivalent triple
        assert np.allclose(T.euler_matrix(*T.euler_from_matrix(M, axes), axes=axes), M)


def test_quaternion_roundtrip_all_shepperd_branches():
    for angles in ((0.1, 0.2, 0.3), (3.0, 0.1, 0.2), (0.1, 3.0, 0.2), (0.2, 0.1, 3.0)):
        R = T.euler_matrix(*angles)[:3, :3]
        q = PEM.rotation_matrix_to_quaternion(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.allclose(T.quaternion_matrix(q)[:3, :3], R)
        M = PEM.transformation_from_translation_quaternion([1, 2, 3], q)
        assert np.allclose(M[:3, :3], R) and PEM.translation_from_transformation_matrix(M) == [1, 2, 3]
        assert np.allclose(T.euler_from_quaternion(q), T.euler_from_matrix(M))


def test_stamped_line_format_and_writer(tmp_path):
    line = PEM.format_stamped_line(12.5, [1.0, 2.0, 3.5], [0.0, 0.0, 0.0, 1.0])
    assert line == "12.5 1.0 2.0 3.5 0.0 0.0 0.0 1.0 \n"          # trailing space before the newline (reference :80-86)
    p = tmp_path / "t.txt"
    PEM.write_to_output_file(str(p), 1.0, [0, 0, 0], [0, 0, 0, 1])
    PEM.write_to_output_file(str(p), 2.0, [1, 0, 0], [0, 0, 0, 1])
    rows = np.genfromtxt(str(p))
    assert rows.shape == (2, 8) and rows[1, 1] == 1.0
    PEM.clear_txt_file_contents(str(p))
    assert p.read_text() == ""


def test_velocity_and_frame_algebra():
    rel = np.eye(4); rel[:3, 3] = [2.0, 4.0, 6.0]
    v = PEM.get_velocity_between_timestamps(rel, 1.0, 3.0)
    assert np.allclose(v[:3, 3], [1, 2, 3]) and np.allclose(v[:3, :3], np.eye(3) / 2)
    A = T.euler_matrix(0.1, 0.2, 0.3); A[:3, 3] = [1, 2, 3]
    B = T.euler_matrix(-0.2, 0.1, 0.4); B[:3, 3] = [0, 1, 0]
    assert np.allclose(PEM.get_marker_to_marker_transformation(A, B), np.linalg.inv(A) @ B)
    assert np.allclose(PEM.get_camera_to_camera_transformation(A, B), A @ np.linalg.inv(B))


def test_shard_pairs_covers_everything_once():
    for n in (0, 1, 7, 8, 999, 1000):
        for g in (1, 2, 3, 4, 8):
            seen = []
            for r in range(g):
                s, e = S.shard_pairs(n, g, r)
                assert 0 <= s <= e <= n
                seen += list(range(s, e))
            assert seen == list(range(n))


def _fake_poses(n, seed=0):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=POSE_DTYPE)
    for i in range(n):
        p[i]["R"] = T.euler_matrix(*(rng.normal(0, 0.01, 3)))[:3, :3].reshape(9)
        t = rng.normal(size=3); p[i]["t"] = t / np.linalg.norm(t)
        p[i]["status"] = 0 if i % 7 else (2 if i else 0)
    return p


def test_chain_matches_reference_construction_and_writes_files(tmp_path):
    poses = _fake_poses(10)
    rel = S.poses_to_relatives(poses)
    assert np.array_equal(rel[7], np.eye(4))                                       # failed pair -> identity
    # the reference's quirk: angles extracted as rotating-xyz, rebuilt as static-xyz (first-order equal to R)
    R = poses[1]["R"].reshape(3, 3)
    assert np.abs(rel[1][:3, :3] - R).max() < 1e-3 and not np.array_equal(rel[1][:3, :3], R)
    assert np.allclose(S.relative_transform(R, poses[1]["t"], exact_rotation=True)[:3, :3], R)
    absolute = S.chain(rel)
    assert len(absolute) == 11 and np.allclose(absolute[3], rel[0] @ rel[1] @ rel[2])
    ts = [0.1 * i for i in range(11)]
    paths = S.write_trajectory(str(tmp_path), ts, poses)
    a = np.genfromtxt(paths["absolute"]); r = np.genfromtxt(paths["relative"]); v = np.genfromtxt(paths["velocity"])
    assert a.shape == (11, 8) and r.shape == (10, 8) and v.shape == (10, 8)
    assert np.allclose(a[3, 1:4], absolute[3][:3, 3]) and np.allclose(r[:, 0], ts[1:])
    assert open(paths["absolute"]).readline().endswith(" \n")
    assert os.path.basename(paths["legacy"]) == "stamped_traj_estimate.txt"


def test_run_sharded_single_process_is_identity():
    poses = _fake_poses(9)
    out = S.run_sharded(10, lambda a, b: poses[a:b])
    assert np.array_equal(out, poses)
    with pytest.raises(ValueError):
        S.run_sharded(10, lambda a, b: poses[:1])


def test_dropin_shims_resolve_the_reference_module_names():
    """`from visual_odometry_v3 import VisualOdometry` / `import pose_estimation_module as PEM` (the reference's own import
    lines, trajectory_evaluation_dual_process.py:21,23) work with dropin/ on sys.path; no GPU needed to import."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "dropin"))
    try:
        for name in ("visual_odometry_v3", "pose_estimation_module"):
            sys.modules.pop(name, None)
        vo3 = importlib.import_module("visual_odometry_v3")
        pem = importlib.import_module("pose_estimation_module")
        assert hasattr(vo3, "VisualOdometry") and callable(vo3.VisualOdometry.visual_odometry_calculations)
        for fn in ("rotation_matrix_to_quaternion", "write_to_output_file", "get_velocity_between_timestamps", "clear_txt_file_contents"):
            assert hasattr(pem, fn), fn
    finally:
        sys.path.pop(0)
        for name in ("visual_odometry_v3", "pose_estimation_module"):
            sys.modules.pop(name, None)


def test_frame_folder_reads_in_name_order_and_by_slice(tmp_path):
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (5, 48, 64), dtype=np.uint8)
    for i, stamp in enumerate((10.5, 2.25, 7.0, 30.0, 4.5)):
        np.save(tmp_path / ("%012.6f.npy" % stamp), imgs[i])
    ff = S.FrameFolder(str(tmp_path))
    order = np.argsort([10.5, 2.25, 7.0, 30.0, 4.5])
    assert len(ff) == 5 and ff.timestamps == sorted([10.5, 2.25, 7.0, 30.0, 4.5])
    assert np.array_equal(ff[1:4], imgs[order[1:4]]) and np.array_equal(ff[0], imgs[order[0]])
    try:
        import cv2
    except ImportError:
        return
    d2 = tmp_path / "png"
    d2.mkdir()
    for i in range(3):
        cv2.imwrite(str(d2 / ("frame_%03d.png" % i)), imgs[i])
    fp = S.FrameFolder(str(d2))
    assert fp.timestamps == [0.0, 1.0, 2.0] and np.array_equal(fp[0:3], imgs[:3])


def test_bench_host_helpers_degrade_without_a_gpu():
    """bench.py's measurement helpers are best effort: without NVML / nvidia-smi they return empty results, never raise,
    and never change the process's CPU affinity."""
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    bound = bench._bind_to_gpu_numa_node(0)
    assert bound is None or (isinstance(bound, int) and bound > 0)
    if bound is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
    clocks = bench.ClockSampler(0).stop()
    assert set(clocks) >= {"sm_mhz", "sm_max_mhz", "reasons"}
