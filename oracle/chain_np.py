"""ORACLE -- the per-pair chain assembled from the numpy restatements (orb_np + pose_np); mirrors
``oracle.cv2_chain.frame_pair`` (reference call sequence: /root/reference/scripts/visual_odometry_v3.py:384-408)."""
from __future__ import annotations

import numpy as np

from . import orb_np, pose_np


def orb_features(img, nfeatures=500, nlevels=8):
    r = orb_np.orb_detect_and_compute(img, nfeatures, nlevels)
    return {k: r[k] for k in ("pt", "size", "angle", "response", "octave", "desc", "lvl_xy")}


def pose_from_points(p_prev, p_cur, K, max_iters=1000, exhaustive=False):
    out = {"status": 0, "E": None, "ransac_mask": None, "R": None, "t": None, "pose_mask": None, "good": 0}
    if len(p_prev) < 5:
        out["status"] = -1
        return out
    E, mask = pose_np.find_essential_mat(p_prev, p_cur, K, 0.999, 1.0, max_iters, exhaustive=exhaustive)
    if E is None:
        out["status"] = -2
        return out
    good, R, t, pmask, cand = pose_np.recover_pose(E, p_prev, p_cur, K)
    out.update(E=E, ransac_mask=mask, R=R, t=t, pose_mask=pmask, good=good, cand=cand)
    return out


def frame_pair(img_prev, img_cur, K, nfeatures=500, matcher="crosscheck", feats_prev=None, feats_cur=None):
    fa = feats_prev if feats_prev is not None else orb_features(img_prev, nfeatures)
    fb = feats_cur if feats_cur is not None else orb_features(img_cur, nfeatures)
    if matcher == "crosscheck":
        m = pose_np.sort_matches(pose_np.bf_match_crosscheck(fa["desc"], fb["desc"]))
    else:
        m = pose_np.sort_matches(pose_np.ratio_and_reverse_check(fa["desc"], fb["desc"]))
    p_prev = fa["pt"][m[:, 0]].astype(np.float32).reshape(-1, 2)
    p_cur = fb["pt"][m[:, 1]].astype(np.float32).reshape(-1, 2)
    res = pose_from_points(p_prev, p_cur, np.asarray(K, dtype=np.float64))
    res.update(matches=m, p_prev=p_prev, p_cur=p_cur, feats_prev=fa, feats_cur=fb)
    return res
