"""ORACLE -- test infrastructure, never imported by the product path.

Loads the REFERENCE's own module, unmodified, from where it lies (/root/reference/scripts/visual_odometry_v3.py) so that
its methods can be executed in this container and their outputs recorded / compared:

  VisualOdometry.__init__ + parse_camera_intrinsics         visual_odometry_v3.py:29-87, :145-166
  compute_current_image_elements                            :370-379
  bf (cv.BFMatcher(NORM_HAMMING, crossCheck=True)) + sorted :75, :219-221
  previous_current_matching                                 :349-368   (KeyPoint_convert, then ...)
  get_transformation_between_two_frames                     :293-345   (findEssentialMat, recoverPose, P = K[R|t], ...)
  get_scaling_factor_from_triangulation                     :263-291   (cv.triangulatePoints on fiducial corners)

The reference imports three packages that are not installed here and are unrelated to the arithmetic of the path
(SURVEY.md 8c): ``transformations`` (Gohlke), ``tf`` (ROS) and ``matplotlib``.  They are replaced by stub modules.  The
``transformations`` stub is written on scipy.spatial.transform.Rotation -- deliberately NOT on the product's
transformations_lite -- so that the euler round trip of :334-339 is checked against an independent implementation:
  euler_matrix(ai, aj, ak, 'sxyz')  = static x, y, z   = Rotation.from_euler('xyz', ...)   (extrinsic, lower case)
  euler_from_matrix(M, 'rxyz')      = rotating x, y, z = Rotation.as_euler('XYZ')          (intrinsic, upper case)

/root/reference exists only in the build container: ``available()`` is False on the GPU box, where the tests use the
fixture recorded here (tests/golden/golden_reference_module.npz, written by tests/golden/make_reference_golden.py).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF_SCRIPTS = "/root/reference/scripts"
_module = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SCRIPTS, "visual_odometry_v3.py"))


def _transformations_stub():
    from scipy.spatial.transform import Rotation

    m = types.ModuleType("transformations")

    def _seq(axes):
        frame, letters = axes[0], axes[1:]
        if frame not in "sr" or len(letters) != 3:
            raise ValueError("axes must look like 'sxyz' / 'rxyz'")
        return letters.lower() if frame == "s" else letters.upper()

    def euler_matrix(ai, aj, ak, axes="sxyz"):
        M = np.eye(4)
        M[:3, :3] = Rotation.from_euler(_seq(axes), [ai, aj, ak]).as_matrix()
        return M

    def euler_from_matrix(matrix, axes="sxyz"):
        a = Rotation.from_matrix(np.asarray(matrix, dtype=np.float64)[:3, :3]).as_euler(_seq(axes))
        return float(a[0]), float(a[1]), float(a[2])

    def translation_matrix(direction):
        M = np.eye(4)
        M[:3, 3] = np.asarray(direction, dtype=np.float64)[:3]
        return M

    def quaternion_matrix(q):      # [x, y, z, w], as tf.transformations
        M = np.eye(4)
        M[:3, :3] = Rotation.from_quat(np.asarray(q, dtype=np.float64)).as_matrix()
        return M

    def euler_from_quaternion(q, axes="sxyz"):
        return tuple(float(v) for v in Rotation.from_quat(np.asarray(q, dtype=np.float64)).as_euler(_seq(axes)))

    m.euler_matrix, m.euler_from_matrix, m.translation_matrix = euler_matrix, euler_from_matrix, translation_matrix
    m.quaternion_matrix, m.euler_from_quaternion = quaternion_matrix, euler_from_quaternion
    return m


def _install_stubs():
    tr = _transformations_stub()
    stubs = {"transformations": tr}
    tf = types.ModuleType("tf")
    tf.transformations = tr
    tf.quaternion_matrix, tf.euler_from_quaternion, tf.euler_from_matrix = tr.quaternion_matrix, tr.euler_from_quaternion, tr.euler_from_matrix
    stubs["tf"] = tf
    stubs["tf.transformations"] = tr
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    stubs["matplotlib"], stubs["matplotlib.pyplot"] = mpl, plt
    tk = types.ModuleType("mpl_toolkits")
    m3 = types.ModuleType("mpl_toolkits.mplot3d")
    m3.axes3d = m3.Axes3D = None
    tk.mplot3d = m3
    stubs["mpl_toolkits"], stubs["mpl_toolkits.mplot3d"] = tk, m3
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    return saved


def load():
    """The reference's ``visual_odometry_v3`` module object (imported once under a private name, its banner prints muted)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise RuntimeError("the reference checkout is not present at " + REF_SCRIPTS)
    import importlib.util
    saved = _install_stubs()
    saved_names = {k: sys.modules.get(k) for k in ("pose_estimation_module",)}
    sys.path.insert(0, REF_SCRIPTS)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            spec = importlib.util.spec_from_file_location("_reference_visual_odometry_v3", os.path.join(REF_SCRIPTS, "visual_odometry_v3.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)       # imports the reference's own pose_estimation_module too (:14)
    finally:
        sys.path.remove(REF_SCRIPTS)
        for k, v in list(saved.items()) + list(saved_names.items()):     # leave no stub (or reference module) behind
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _module = mod
    return mod


def write_controlled_calibration(path, K, dist=(0.0, 0.0, 0.0, 0.0, 0.0)):
    """A calibration file in the schema the reference reads with controlled=True (:156-161)."""
    import yaml
    with open(path, "w") as f:
        yaml.safe_dump({"camera_matrix": {"rows": 3, "cols": 3, "data": [float(v) for v in np.asarray(K).ravel()]},
                        "distortion_coefficients": {"rows": 1, "cols": 5, "data": [float(v) for v in dist]}}, f)


def make_vo(calibration_file_path, real_marker_length, nfeatures=None, mode="orb"):
    """Construct the reference's VisualOdometry(controlled=True) -- the only configuration in which its projection matrix is
    initialised (:164-166).  ``nfeatures`` swaps the detector for cv.ORB_create(nfeatures=...) (the literal is ORB_create())."""
    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):
        vo = ref.VisualOdometry(mode=mode, calibration_file_path=calibration_file_path, controlled=True,
                                real_marker_length=real_marker_length)
    if nfeatures is not None:
        import cv2
        vo.feature_detector = cv2.ORB_create(nfeatures=int(nfeatures))
    return vo


def pair_through_reference(vo, previous_image, current_image, robot_previous_position_transformation, previous_marker_corners,
                           current_marker_corners):
    """visual_odometry_calculations (:384-408) executed with the reference's own methods, except for the three lines of
    get_matches_between_two_frames that index a DMatch (:234-238, TypeError in ORB mode): those are restated by their
    evident intent.  Everything from previous_current_matching (:349) on is the reference's code."""
    with contextlib.redirect_stdout(io.StringIO()):
        kp_prev, d_prev, _ = vo.compute_current_image_elements(previous_image)
        kp_cur, d_cur, _ = vo.compute_current_image_elements(current_image)
        matches = vo.bf.match(d_prev, d_cur)                                 # :219
        matches = sorted(matches, key=lambda x: x.distance)                  # :221
        top_prev = [kp_prev[m.queryIdx] for m in matches]                    # :237 (intent)
        top_cur = [kp_cur[m.trainIdx] for m in matches]                      # :238 (intent)
        cur_pose, rel = vo.previous_current_matching(top_prev, top_cur, robot_previous_position_transformation,
                                                     previous_marker_corners, current_marker_corners)
    feats = lambda kps, d: {"pt": np.array([k.pt for k in kps], np.float32).reshape(-1, 2),      # noqa: E731
                            "angle": np.array([k.angle for k in kps], np.float32), "response": np.array([k.response for k in kps], np.float32),
                            "octave": np.array([k.octave for k in kps], np.int32), "size": np.array([k.size for k in kps], np.float32), "desc": d}
    return {"feats_prev": feats(kp_prev, d_prev), "feats_cur": feats(kp_cur, d_cur),
            "matches": np.array([(m.queryIdx, m.trainIdx, int(m.distance)) for m in matches], np.int32).reshape(-1, 3),
            "E": np.array(vo.essential_matrix), "cur_pose": np.array(cur_pose), "rel": np.array(rel),
            "projection": np.array(vo.previous_projection_matrix)}
