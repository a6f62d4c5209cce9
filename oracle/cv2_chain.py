"""ORACLE / CPU BASELINE -- test and measurement infrastructure, never imported by the product path.

The reference's own hot path, issued call for call against the cv2 wheel in this image (cv2 4.13.0), with the
reference's literal parameters:

  cv.ORB_create()                               /root/reference/scripts/visual_odometry_v3.py:96   (nfeatures per config)
  detectAndCompute(img, None)                   :373   (twice per pair, :387 and :391)
  cv.BFMatcher(NORM_HAMMING, crossCheck=True)   :75 with :97-98
  bf.match(prev_desc, cur_desc)                 :219
  sorted(matches, key=distance)                 :221
  kp_prev[m.queryIdx], kp_cur[m.trainIdx]       :237-238 (the shipped code indexes [i][0] at :234-235 and raises
                                                TypeError in ORB mode; the evident intent is restated here)
  cv.KeyPoint_convert                           :355, :358
  cv.findEssentialMat(p_prev, p_cur, K, RANSAC, prob=0.999, threshold=1.0)   :297-300
  cv.recoverPose(E, p_prev, p_cur, K)           :303-306

This module is what ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time (kind "reference"), and
what the parity tests compare the CUDA path with when cv2 is importable.  cv2 is a dependency of the oracle only.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 lives in the image; the numpy restatement (orb_np / pose_np) stands in if it is ever absent
    import cv2 as cv
except Exception:  # pragma: no cover
    cv = None

RATIO = 0.75  # /root/reference/scripts/visual_odometry_v3.py:227


def available() -> bool:
    return cv is not None


def orb_features(img: np.ndarray, nfeatures: int = 500, nlevels: int = 8):
    """detectAndCompute -> dict of arrays (pt, size, angle, response, octave, desc)."""
    det = cv.ORB_create(nfeatures=nfeatures, nlevels=nlevels)
    kps, desc = det.detectAndCompute(img, None)
    n = len(kps)
    out = {
        "pt": np.array([k.pt for k in kps], dtype=np.float32).reshape(n, 2),
        "size": np.array([k.size for k in kps], dtype=np.float32),
        "angle": np.array([k.angle for k in kps], dtype=np.float32),
        "response": np.array([k.response for k in kps], dtype=np.float32),
        "octave": np.array([k.octave for k in kps], dtype=np.int32),
        "desc": desc if desc is not None else np.zeros((0, 32), np.uint8),
    }
    return out


def match_crosscheck(d_prev: np.ndarray, d_cur: np.ndarray) -> np.ndarray:
    """bf.match + stable sort by distance -> (M, 3) int32 rows (queryIdx, trainIdx, distance)."""
    bf = cv.BFMatcher(cv.NORM_HAMMING, crossCheck=True)
    ms = sorted(bf.match(d_prev, d_cur), key=lambda m: m.distance)
    return np.array([(m.queryIdx, m.trainIdx, int(m.distance)) for m in ms], dtype=np.int32).reshape(-1, 3)


def match_knn_ratio(d_prev: np.ndarray, d_cur: np.ndarray) -> np.ndarray:
    """Config 4: knnMatch(k=2) on a crossCheck=False matcher + ratio 0.75 + reverse 1-NN check, sorted by distance."""
    bf = cv.BFMatcher(cv.NORM_HAMMING, crossCheck=False)
    knn = bf.knnMatch(d_prev, d_cur, k=2)
    rev = bf.match(d_cur, d_prev)
    back = {m.queryIdx: m.trainIdx for m in rev}
    good = []
    for pair in knn:
        if len(pair) < 2:
            continue
        m, n = pair
        if m.distance < RATIO * n.distance and back.get(m.trainIdx, -1) == m.queryIdx:
            good.append(m)
    good = sorted(good, key=lambda m: m.distance)
    return np.array([(m.queryIdx, m.trainIdx, int(m.distance)) for m in good], dtype=np.int32).reshape(-1, 3)


def pose_from_points(p_prev: np.ndarray, p_cur: np.ndarray, K: np.ndarray, max_iters: int = 1000):
    """findEssentialMat + recoverPose exactly as the reference calls them.  Returns dict or status -1."""
    out = {"status": 0, "E": None, "ransac_mask": None, "R": None, "t": None, "pose_mask": None, "good": 0}
    if len(p_prev) < 5:
        out["status"] = -1
        return out
    E, mask = cv.findEssentialMat(points1=p_prev, points2=p_cur, cameraMatrix=K, method=cv.RANSAC, prob=0.999,
                                  threshold=1.0, maxIters=max_iters)
    if E is None or E.shape != (3, 3):
        out["status"] = -2
        return out
    good, R, t, pmask = cv.recoverPose(E=E, points1=p_prev, points2=p_cur, cameraMatrix=K)
    out.update(E=E, ransac_mask=mask[:, 0].astype(np.uint8), R=R, t=t, pose_mask=pmask[:, 0].astype(np.uint8), good=int(good))
    return out


def frame_pair(img_prev: np.ndarray, img_cur: np.ndarray, K: np.ndarray, nfeatures: int = 500, matcher: str = "crosscheck",
               feats_prev=None, feats_cur=None):
    """The whole per-pair chain.  ``feats_*`` let a sequence runner reuse a frame's features; the reference itself
    recomputes both frames per pair (visual_odometry_v3.py:387-392) -- pass None for that behaviour."""
    fa = feats_prev if feats_prev is not None else orb_features(img_prev, nfeatures)
    fb = feats_cur if feats_cur is not None else orb_features(img_cur, nfeatures)
    if len(fa["desc"]) == 0 or len(fb["desc"]) == 0:
        m = np.zeros((0, 3), np.int32)
    elif matcher == "crosscheck":
        m = match_crosscheck(fa["desc"], fb["desc"])
    else:
        m = match_knn_ratio(fa["desc"], fb["desc"])
    p_prev = fa["pt"][m[:, 0]].astype(np.float32).reshape(-1, 2)
    p_cur = fb["pt"][m[:, 1]].astype(np.float32).reshape(-1, 2)
    res = pose_from_points(p_prev, p_cur, np.asarray(K, dtype=np.float64))
    res.update(matches=m, p_prev=p_prev, p_cur=p_cur, feats_prev=fa, feats_cur=fb)
    return res


def marker_scaled_transform(R, t, K, prev_projection, prev_corners, cur_corners, real_marker_length):
    """The tail of get_transformation_between_two_frames, /root/reference/scripts/visual_odometry_v3.py:309-345, call for call:
    P = K [R|t] (:309), cv.triangulatePoints on the fiducial corners (:265), distance of the first two RAW homogeneous
    points (:272-279), t *= real_marker_length / distance (:321-325), euler_from_matrix(R, 'rxyz') (:334) fed to
    euler_matrix(..., 'sxyz') (:140) behind translation_matrix(t) (:141-142).  The two Gohlke conventions are taken from
    scipy.spatial.transform.Rotation (rotating xyz = intrinsic 'XYZ', static xyz = extrinsic 'xyz').
    Returns (4x4 prev_to_curr, current projection matrix, measured distance)."""
    from scipy.spatial.transform import Rotation
    K = np.asarray(K, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(3, 1)
    P = K.dot(np.hstack((R, t)))
    X = cv.triangulatePoints(projMatr1=np.asarray(prev_projection, dtype=np.float64), projMatr2=P,
                             projPoints1=np.asarray(prev_corners).T, projPoints2=np.asarray(cur_corners).T)
    d = float(np.sqrt((X[0, 0] - X[0, 1]) ** 2 + (X[1, 0] - X[1, 1]) ** 2 + (X[2, 0] - X[2, 1]) ** 2))
    ts = t[:, 0] * (real_marker_length / d)
    e = Rotation.from_matrix(R).as_euler("XYZ")
    M = np.eye(4)
    M[:3, :3] = Rotation.from_euler("xyz", e).as_matrix()
    T = np.eye(4)
    T[:3, 3] = ts
    return T.dot(M), P, d
