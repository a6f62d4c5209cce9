"""Drop-in for the reference's ``visual_odometry_v3`` (/root/reference/scripts/visual_odometry_v3.py): class
``VisualOdometry`` with the reference's constructor, public attributes and method names, where every cv2 call on the
per-frame-pair hot path (:373 detectAndCompute, :219 bf.match, :221 sorted, :355/:358 KeyPoint_convert,
:297-300 findEssentialMat, :303-306 recoverPose) is replaced by libdvo's sm_100a kernels through the C ABI.

Deliberate, documented deviations from the shipped reference:
  * ORB mode in the reference raises TypeError at :234 (``passed_ratio_test[i][0]`` on a DMatch).  The evident intent
    -- every cross-checked match, sorted by distance, goes to findEssentialMat -- is what runs here.
  * ``compute_current_image_elements`` returns ``None`` for the drawn image (the reference draws keypoints with
    cv.drawKeypoints and no caller uses the result, :375).
  * marker corners may be ``None``: then no metric scale is applied and the translation stays unit-norm (the reference
    needs fiducial corners and ``controlled=True`` to have a projection matrix, :164-166, :263-291).
  * SIFT/SURF/FLANN modes (:99-106) are float-descriptor paths outside the ORB contract and are not provided.
There is no CPU fallback: without libdvo.so or without a CUDA device the constructor raises.
"""
from __future__ import annotations

import math

import numpy as np
import yaml
from yaml.loader import SafeLoader

from . import _native
from . import pose_estimation_module as PEM  # noqa: F401  (the reference imports it under this name, :14)
from . import transformations_lite as transf

VERBOSE = False
DEFAULT_STARTING_ROBOT_TRANSLATION = [0, 0, 0]
DEFAULT_STARTING_ROBOT_EULER = [0, 0, 0]
DEFAULT_NFEATURES = 500   # cv.ORB_create() default, the reference's literal (:96)


class KeyPoint:
    """Minimal stand-in for cv2.KeyPoint (pt, size, angle, response, octave, class_id)."""
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y, size, angle=-1.0, response=0.0, octave=0, class_id=-1):
        self.pt = (float(x), float(y))
        self.size, self.angle, self.response, self.octave, self.class_id = float(size), float(angle), float(response), int(octave), int(class_id)


class DMatch:
    """Minimal stand-in for cv2.DMatch."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx, trainIdx, distance, imgIdx=0):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = int(queryIdx), int(trainIdx), int(imgIdx), float(distance)


def keypoints_from_features(f):
    return [KeyPoint(p[0], p[1], s, a, r, o) for p, s, a, r, o in zip(f["pt"], f["size"], f["angle"], f["response"], f["octave"])]


def keypoints_to_array(kps):
    """cv.KeyPoint_convert: list of keypoints -> (n, 2) float32."""
    return np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)


class PairEngine:
    """Thin, allocation-once wrapper of one libdvo context for two-frame work (the reference's unit of work)."""

    def __init__(self, width, height, nfeatures=DEFAULT_NFEATURES, device=0, matcher=_native.DVO_MATCH_CROSSCHECK, max_frames=2, **kw):
        self.ctx = _native.Context(width, height, nfeatures=nfeatures, max_frames=max_frames, matcher=matcher, device=device, **kw)

    def features(self, img, slot=0):
        self.ctx.load_frames(img, slot)
        self.ctx.orb(slot, 1)
        return self.ctx.features(slot)

    def _collect(self, K, fa=None, fb=None):
        pose = self.ctx.poses(0, 1)[0]
        if pose["frame_flags"]:
            raise _native.DvoError("DVO_E_CAPACITY: a frame of this pair holds a truncated keypoint set (%s)"
                                   % _native.describe_frame_flags(int(pose["frame_flags"])))
        arr = self.ctx.pair_arrays(0, pose["n_matches"])
        out = {"status": int(pose["status"]), "E": pose["E"].reshape(3, 3).copy(), "R": pose["R"].reshape(3, 3).copy(),
               "t": pose["t"].reshape(3, 1).copy(), "good": int(pose["n_good"]), "n_inliers": int(pose["n_inliers"]),
               "candidate": int(pose["candidate"]), "ransac_iters": int(pose["ransac_iters"]), "best_iter": int(pose["best_iter"]),
               "matches": arr["matches"], "p_prev": arr["p_prev"], "p_cur": arr["p_cur"], "ransac_mask": arr["ransac_mask"],
               "pose_mask": arr["pose_mask"], "feats_prev": fa, "feats_cur": fb}
        return out

    def frame_pair(self, img_prev, img_cur, K, want_features=True):
        """ORB x2 -> match -> E-RANSAC -> recoverPose for one pair; returns host copies of everything."""
        c = self.ctx
        c.load_frames(img_prev, 0)
        c.load_frames(img_cur, 1)
        c.orb(0, 2)
        c.pairs(0, 0, 1, K)
        fa = c.features(0) if want_features else None
        fb = c.features(1) if want_features else None
        return self._collect(K, fa, fb)

    def match_and_pose(self, pt_prev, desc_prev, pt_cur, desc_cur, K):
        c = self.ctx
        c.set_features(0, pt_prev, desc_prev)
        c.set_features(1, pt_cur, desc_cur)
        c.pairs(0, 0, 1, K)
        return self._collect(K)

    def match_only(self, pt_prev, desc_prev, pt_cur, desc_cur, K):
        """bf.match + sorted + KeyPoint_convert (dvo_match): no RANSAC, no pose.  The correspondences stay in pair slot 0
        so that pose_of_last_match() can run findEssentialMat + recoverPose on them without another upload."""
        c = self.ctx
        c.set_features(0, pt_prev, desc_prev)
        c.set_features(1, pt_cur, desc_cur)
        c.match(0, 0, 1, K)
        n = c.match_count(0)
        arr = c.pair_arrays(0, n)
        return {"matches": arr["matches"], "p_prev": arr["p_prev"], "p_cur": arr["p_cur"]}

    def pose_of_last_match(self, K):
        self.ctx.pose(0, 0, 1, K)
        return self._collect(K)

    def pose_from_points(self, p_prev, p_cur, K):
        self.ctx.pose_points(p_prev, p_cur, K, 0)
        return self._collect(K)


NORM_HAMMING = 6     # cv.NORM_HAMMING


class OrbFeatureDetector:
    """What ``VisualOdometry.feature_detector`` is in the reference: the object ``cv.ORB_create()`` returns
    (visual_odometry_v3.py:96), reduced to the one call the reference makes on it (:373) plus the getters of the parameters
    this build fixes.  The work runs in libdvo."""

    def __init__(self, owner):
        self._owner = owner

    def detectAndCompute(self, image, mask=None):
        if mask is not None:
            raise NotImplementedError("detectAndCompute with a mask is not on the reference's path (it passes None, :373)")
        kps, desc, _ = self._owner.compute_current_image_elements(image)
        return kps, desc

    def getMaxFeatures(self):
        return self._owner.nfeatures

    def getScaleFactor(self):
        return 1.2000000476837158

    def getNLevels(self):
        return 8

    def getEdgeThreshold(self):
        return 31

    def getFirstLevel(self):
        return 0

    def getWTA_K(self):
        return 2

    def getScoreType(self):
        return 0      # cv.ORB_HARRIS_SCORE

    def getPatchSize(self):
        return 31

    def getFastThreshold(self):
        return 20

    def descriptorSize(self):
        return 32

    def defaultNorm(self):
        return NORM_HAMMING


class BruteForceMatcher:
    """What ``VisualOdometry.bf`` is in the reference: ``cv.BFMatcher(normType=NORM_HAMMING, crossCheck=True)`` (:75), with the
    one method the ORB path calls on it (``match``, :219).  Matches come back in cv2's order (by queryIdx)."""

    def __init__(self, owner, norm_type, cross_check):
        self._owner, self.normType, self.crossCheck = owner, norm_type, cross_check

    def match(self, queryDescriptors, trainDescriptors):
        q = np.ascontiguousarray(queryDescriptors, dtype=np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(trainDescriptors, dtype=np.uint8).reshape(-1, 32)
        eng = self._owner._engine_for_descriptors(max(len(q), len(t)))
        res = eng.match_only(np.zeros((len(q), 2), np.float32), q, np.zeros((len(t), 2), np.float32), t, np.eye(3))
        m = res["matches"]
        m = m[np.argsort(m[:, 0], kind="stable")]
        return [DMatch(a, b, d) for a, b, d in m]

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        raise NotImplementedError("knnMatch belongs to the float-descriptor modes (:203-215); the k=2 ratio matcher of BASELINE "
                                  "configs[3] is the context option matcher=DVO_MATCH_KNN_RATIO")


class VisualOdometry:
    def __init__(self, starting_translation=None, starting_euler=None, to_sort=False, mode="ORB",
                 calibration_file_path="", controlled=False, real_marker_length=0.0,
                 nfeatures=DEFAULT_NFEATURES, device=0, camera_matrix=None):
        if starting_euler is None:
            starting_euler = DEFAULT_STARTING_ROBOT_EULER
        if starting_translation is None:
            starting_translation = DEFAULT_STARTING_ROBOT_TRANSLATION
        self.controlled = controlled
        if not controlled:                      # reference :39-44
            self.frame_height, self.frame_width = 1080, 1400
        else:
            self.frame_height, self.frame_width = 480, 640
        self.starting_translation = starting_translation
        self.starting_euler = starting_euler
        self.robot_current_translation = None
        self.essential_matrix = None
        self.calibration_file_path = calibration_file_path
        self.distortion_coefficient_matrix = None
        self.intrinsic_coefficient_matrix = None
        self.previous_projection_matrix = None
        if camera_matrix is not None:           # synthetic runs: K given directly instead of a YAML file
            self.intrinsic_coefficient_matrix = np.asarray(camera_matrix, dtype=np.float64).reshape(3, 3)
            self.distortion_coefficient_matrix = np.zeros(5)
            if controlled:
                self.previous_projection_matrix = self.intrinsic_coefficient_matrix @ np.hstack((np.eye(3), np.zeros((3, 1))))
        else:
            self.parse_camera_intrinsics()      # raises on a bad path, like the reference (:62, :147)
        self.to_sort = to_sort
        self.mode = mode
        self.real_marker_length = real_marker_length
        if mode.lower() != "orb":
            raise NotImplementedError("only mode='orb' is on the accelerated hot path (SIFT/SURF/FLANN: visual_odometry_v3.py:99-106)")
        self.nfeatures = int(nfeatures)
        self.device = int(device)
        # reference :70, :75 -- the detector / matcher objects and their parameters are public attributes
        self.feature_detector, self.norm_type, self.cross_check = self.return_feature_matching_parameters(mode)
        self.bf = BruteForceMatcher(self, self.norm_type, self.cross_check)
        self._engine = None                     # created on first image (frame size known then)
        self._engine_size = None
        _native.load_library()                  # fail now, loudly, if the CUDA library is missing
        _native._torch()                        # ... or if no CUDA device is visible
        self.robot_position_list = []
        self.ground_truth_list = []
        self.frame_translations = []
        self.matches_dictionary = []
        self.projection_matrix_list = []
        self.plot_4D_counter = 1
        self.last_pair = None                   # raw (R, t, E, masks, matched points) of the most recent pair
        self.robot_curr_position = self.make_transform_mat(translation=self.starting_translation, euler=self.starting_euler)

    # ------------------------------------------------------------------ utilities (reference :93-166)
    def return_feature_matching_parameters(self, mode):       # reference :93-107
        if mode.lower() == "orb":
            return OrbFeatureDetector(self), NORM_HAMMING, True
        raise NotImplementedError("only mode='orb' is on the accelerated hot path (SIFT/SURF/FLANN: visual_odometry_v3.py:99-106)")

    def _engine_for_descriptors(self, n):
        """Engine for descriptor-only work (bf.match before any image was seen, or on more rows than the image engine holds)."""
        if self._engine is not None and n <= self._engine.ctx.max_keypoints:
            return self._engine
        eng = getattr(self, "_desc_engine", None)
        if eng is None or n > eng.ctx.max_keypoints:
            eng = PairEngine(256, 256, nfeatures=max(int(n), 500), device=self.device)
            self._desc_engine = eng
        return eng

    def _engine_for(self, height, width):
        if self._engine is None or self._engine_size != (height, width):
            self._engine = PairEngine(width, height, nfeatures=self.nfeatures, device=self.device)
            self._engine_size = (height, width)
        return self._engine

    def make_transform_mat(self, translation, euler):          # reference :138-142
        rx, ry, rz = euler
        rotation = transf.euler_matrix(rx, ry, rz, axes="sxyz")
        return transf.translation_matrix(translation).dot(rotation)

    def parse_camera_intrinsics(self):                          # reference :145-166
        with open(self.calibration_file_path) as camera_calibration:
            data = yaml.load(camera_calibration, Loader=SafeLoader)
        if not self.controlled:
            self.distortion_coefficient_matrix = np.array(data["distortion_coeffs"][0])
            self.intrinsic_coefficient_matrix = np.array(data["intrinsic_coeffs"][0]).reshape((3, 3))
        else:
            self.intrinsic_coefficient_matrix = np.array(data["camera_matrix"]["data"]).reshape((3, 3))
            self.distortion_coefficient_matrix = np.array(data["distortion_coefficients"]["data"]).reshape((1, 5))
            self.previous_projection_matrix = np.matmul(self.intrinsic_coefficient_matrix, np.hstack((np.eye(3), np.zeros((3, 1)))))

    def undistort_image(self, distorted_image, new_camera_matrix):     # reference :110-113
        """cv.undistort(src, cameraMatrix, distCoeffs, newCameraMatrix) on the GPU (k_ingest), bit-exact with cv2; accepts the
        grey image the reference passes, or a BGR image (then cv.cvtColor(BGR2GRAY) is applied first, as :132 does)."""
        img = np.ascontiguousarray(distorted_image, dtype=np.uint8)
        eng = self._engine_for(img.shape[0], img.shape[1])
        c = eng.ctx
        c.set_undistort(self.intrinsic_coefficient_matrix, np.asarray(self.distortion_coefficient_matrix, dtype=np.float64).ravel(),
                        new_camera_matrix, channels=3 if img.ndim == 3 else 1)
        try:
            c.load_frames(img, 0)
            out = c.tap_image(0, 0, 0)
        finally:
            c.set_undistort(None)
        return out

    def ros_img_msg_to_opencv_image(self, image_message, msg_type):    # reference :115-135
        """decode on the host (cv.imdecode is entropy decoding, outside the hot path), grey + undistort on the GPU."""
        import cv2 as cv
        new_camera_matrix, _ = cv.getOptimalNewCameraMatrix(self.intrinsic_coefficient_matrix, self.distortion_coefficient_matrix,
                                                            (self.frame_width, self.frame_height), 1, (self.frame_width, self.frame_height))
        if msg_type == "compressed":
            image_np = cv.imdecode(np.frombuffer(image_message.data, np.uint8), cv.IMREAD_COLOR)
        elif msg_type == "usb_raw":
            image_np = np.frombuffer(image_message.data, dtype=np.uint8).reshape((image_message.height, image_message.width, -1))
        else:
            raise ValueError("unknown msg_type " + str(msg_type))
        if image_np.ndim == 3 and image_np.shape[2] == 1:
            image_np = image_np[:, :, 0]
        return self.undistort_image(image_np, new_camera_matrix)

    # ------------------------------------------------------------------ hot path
    def compute_current_image_elements(self, input_image):     # reference :370-379
        img = np.ascontiguousarray(input_image, dtype=np.uint8)
        f = self._engine_for(*img.shape).features(img, slot=0)
        self._last_features = f
        return keypoints_from_features(f), f["desc"], None

    def get_matches_between_two_frames(self, previous_key_points, previous_descriptors, current_key_points, current_descriptors):
        """reference :191-239 (ORB branch, intended semantics)."""
        if self._engine is None:
            raise _native.DvoError("no frame has been processed yet: frame size unknown")
        res = self._engine.match_only(keypoints_to_array(previous_key_points), previous_descriptors,
                                      keypoints_to_array(current_key_points), current_descriptors, self.intrinsic_coefficient_matrix)
        m = res["matches"]
        matches = [DMatch(q, t, d) for q, t, d in m]
        top_prev = [previous_key_points[q] for q in m[:, 0]]
        top_cur = [current_key_points[t] for t in m[:, 1]]
        self._pending_match = res   # the same correspondences will be asked for a pose next: they are already on the device
        return matches, top_prev, top_cur

    def get_scaling_factor_from_triangulation(self, current_projection_matrix, previous_marker_corners, current_marker_corners):
        """reference :263-291: cv.triangulatePoints on the fiducial corners of the two frames, then the distance between the
        first two triangulated corners measured on the RAW homogeneous vectors (no division by w, :272-279).  The vectors
        come from libdvo's restatement of OpenCV's DLT + Jacobi SVD (dvo_triangulate_points_host), which reproduces cv2's
        null-vector sign: with the opposite sign for one of the two points the 'distance' becomes |X0 + X1|."""
        self.projection_matrix_list.append(self.previous_projection_matrix)
        a = np.asarray(previous_marker_corners)
        b = np.asarray(current_marker_corners)
        marker_corners_4D = _native.triangulate_points(self.previous_projection_matrix, current_projection_matrix,
                                                       a.reshape(-1, 2), b.reshape(-1, 2))
        if a.dtype == np.float32:       # cv2 returns points4D in the type of projPoints1
            marker_corners_4D = marker_corners_4D.astype(np.float32)
        marker_Xs, marker_Ys, marker_Zs = marker_corners_4D[0, :], marker_corners_4D[1, :], marker_corners_4D[2, :]
        return math.sqrt((marker_Xs[0] - marker_Xs[1]) ** 2 + (marker_Ys[0] - marker_Ys[1]) ** 2 + (marker_Zs[0] - marker_Zs[1]) ** 2)

    def get_transformation_between_two_frames(self, array_previous_key_points, array_current_key_points,
                                              previous_marker_corners=None, current_marker_corners=None):
        """reference :293-345."""
        pend = getattr(self, "_pending_match", None)
        self._pending_match = None
        if pend is not None and len(pend["p_prev"]) == len(array_previous_key_points) and \
                np.array_equal(pend["p_prev"], np.asarray(array_previous_key_points, np.float32).reshape(-1, 2)) and \
                np.array_equal(pend["p_cur"], np.asarray(array_current_key_points, np.float32).reshape(-1, 2)):
            res = self._engine.pose_of_last_match(self.intrinsic_coefficient_matrix)
        else:
            h, w = self._engine_size if self._engine_size else (self.frame_height, self.frame_width)
            res = self._engine_for(h, w).pose_from_points(array_previous_key_points, array_current_key_points,
                                                          self.intrinsic_coefficient_matrix)
        return self._finish_pair(res, previous_marker_corners, current_marker_corners)

    def _finish_pair(self, res, previous_marker_corners, current_marker_corners):
        self.last_pair = res
        if res["status"] != _native.PAIR_OK:
            # cv.findEssentialMat returns None for < 5 points and cv.recoverPose then throws in the reference
            raise _native.DvoError("pose estimation failed for this pair: status %d (%d matches)" % (res["status"], len(res["matches"])))
        self.essential_matrix = res["E"]
        relative_rotation, translation = res["R"], res["t"]
        current_projection_matrix = self.intrinsic_coefficient_matrix.dot(np.hstack((relative_rotation, translation.reshape(-1, 1))))
        translation = translation.transpose()[0]
        if previous_marker_corners is not None and current_marker_corners is not None and self.previous_projection_matrix is not None:
            d = self.get_scaling_factor_from_triangulation(current_projection_matrix, previous_marker_corners, current_marker_corners)
            translation = translation * (self.real_marker_length / d)          # reference :321-325
        new_rotation_mat = np.vstack((np.hstack((np.array(relative_rotation), np.zeros((3, 1)))), [0, 0, 0, 1]))
        euler = np.array(transf.euler_from_matrix(new_rotation_mat, "rxyz"))   # reference :334 (extract rotating-xyz ...
        prev_to_curr = self.make_transform_mat(translation=translation, euler=euler)   # ... rebuild static-xyz, :339; quirk kept)
        self.frame_translations.append(prev_to_curr)
        self.previous_projection_matrix = current_projection_matrix
        return prev_to_curr

    def previous_current_matching(self, top_previous_key_points, top_current_key_points, robot_previous_position_transformation,
                                  previous_marker_corners=None, current_marker_corners=None):      # reference :349-368
        a = keypoints_to_array(top_previous_key_points)
        b = keypoints_to_array(top_current_key_points)
        rel = self.get_transformation_between_two_frames(a, b, previous_marker_corners, current_marker_corners)
        return robot_previous_position_transformation.dot(rel), rel

    def visual_odometry_calculations(self, previous_image, current_image, robot_previous_position_transformation,
                                     previous_marker_corners=None, current_marker_corners=None):   # reference :384-408
        prev = np.ascontiguousarray(previous_image, dtype=np.uint8)
        cur = np.ascontiguousarray(current_image, dtype=np.uint8)
        res = self._engine_for(*prev.shape).frame_pair(prev, cur, self.intrinsic_coefficient_matrix, want_features=False)
        rel = self._finish_pair(res, previous_marker_corners, current_marker_corners)
        return robot_previous_position_transformation.dot(rel), rel

    def relative_pose(self, previous_image, current_image):
        """What BASELINE.json's north star calls 'the per-frame-pair call returning R, t and matched keypoints'."""
        prev = np.ascontiguousarray(previous_image, dtype=np.uint8)
        cur = np.ascontiguousarray(current_image, dtype=np.uint8)
        res = self._engine_for(*prev.shape).frame_pair(prev, cur, self.intrinsic_coefficient_matrix, want_features=False)
        self.last_pair = res
        return res["R"], res["t"], res["p_prev"], res["p_cur"], res


if __name__ == "__main__":
    pass
