"""Host-side restatement of the few functions the reference takes from Christoph Gohlke's ``transformations`` module
and from ROS ``tf.transformations`` (neither is installed here, neither is vendored by the reference):

  transformations.euler_matrix / translation_matrix    /root/reference/scripts/visual_odometry_v3.py:140-141
  transformations.euler_from_matrix(M, 'rxyz')         /root/reference/scripts/visual_odometry_v3.py:334
  tf.quaternion_matrix, tf.transformations.euler_from_quaternion   /root/reference/scripts/pose_estimation_module.py:17,119

Conventions follow the published module: axes strings 'sxyz' ... 'rzyz'; quaternions are ROS order [x, y, z, w].
Microseconds of float64 work per pair -- host code, not a kernel.
"""
from __future__ import annotations

import math

import numpy as np

_EPS = np.finfo(float).eps * 4.0
_NEXT_AXIS = [1, 2, 0, 1]
_AXES2TUPLE = {
    "sxyz": (0, 0, 0, 0), "sxyx": (0, 0, 1, 0), "sxzy": (0, 1, 0, 0), "sxzx": (0, 1, 1, 0), "syzx": (1, 0, 0, 0),
    "syzy": (1, 0, 1, 0), "syxz": (1, 1, 0, 0), "syxy": (1, 1, 1, 0), "szxy": (2, 0, 0, 0), "szxz": (2, 0, 1, 0),
    "szyx": (2, 1, 0, 0), "szyz": (2, 1, 1, 0), "rzyx": (0, 0, 0, 1), "rxyx": (0, 0, 1, 1), "ryzx": (0, 1, 0, 1),
    "rxzx": (0, 1, 1, 1), "rxzy": (1, 0, 0, 1), "ryzy": (1, 0, 1, 1), "rzxy": (1, 1, 0, 1), "ryxy": (1, 1, 1, 1),
    "ryxz": (2, 0, 0, 1), "rzxz": (2, 0, 1, 1), "rxyz": (2, 1, 0, 1), "rzyz": (2, 1, 1, 1)}


def _axes(axes):
    firstaxis, parity, repetition, frame = _AXES2TUPLE[axes.lower()] if isinstance(axes, str) else axes
    i = firstaxis
    j = _NEXT_AXIS[i + parity]
    k = _NEXT_AXIS[i - parity + 1]
    return i, j, k, parity, repetition, frame


def translation_matrix(direction):
    M = np.identity(4)
    M[:3, 3] = np.asarray(direction, dtype=np.float64)[:3]
    return M


def euler_matrix(ai, aj, ak, axes="sxyz"):
    i, j, k, parity, repetition, frame = _axes(axes)
    if frame:
        ai, ak = ak, ai
    if parity:
        ai, aj, ak = -ai, -aj, -ak
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    M = np.identity(4)
    if repetition:
        M[i, i] = cj; M[i, j] = sj * si; M[i, k] = sj * ci
        M[j, i] = sj * sk; M[j, j] = -cj * ss + cc; M[j, k] = -cj * cs - sc
        M[k, i] = -sj * ck; M[k, j] = cj * sc + cs; M[k, k] = cj * cc - ss
    else:
        M[i, i] = cj * ck; M[i, j] = sj * sc - cs; M[i, k] = sj * cc + ss
        M[j, i] = cj * sk; M[j, j] = sj * ss + cc; M[j, k] = sj * cs - sc
        M[k, i] = -sj; M[k, j] = cj * si; M[k, k] = cj * ci
    return M


def euler_from_matrix(matrix, axes="sxyz"):
    i, j, k, parity, repetition, frame = _axes(axes)
    M = np.array(matrix, dtype=np.float64, copy=False)[:3, :3]
    if repetition:
        sy = math.sqrt(M[i, j] * M[i, j] + M[i, k] * M[i, k])
        if sy > _EPS:
            ax = math.atan2(M[i, j], M[i, k]); ay = math.atan2(sy, M[i, i]); az = math.atan2(M[j, i], -M[k, i])
        else:
            ax = math.atan2(-M[j, k], M[j, j]); ay = math.atan2(sy, M[i, i]); az = 0.0
    else:
        cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
        if cy > _EPS:
            ax = math.atan2(M[k, j], M[k, k]); ay = math.atan2(-M[k, i], cy); az = math.atan2(M[j, i], M[i, i])
        else:
            ax = math.atan2(-M[j, k], M[j, j]); ay = math.atan2(-M[k, i], cy); az = 0.0
    if parity:
        ax, ay, az = -ax, -ay, -az
    if frame:
        ax, az = az, ax
    return ax, ay, az


def quaternion_matrix(quaternion):
    """ROS tf.transformations convention: quaternion = [x, y, z, w]."""
    q = np.array(quaternion[:4], dtype=np.float64, copy=True)
    nq = float(np.dot(q, q))
    if nq < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
        (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)


def euler_from_quaternion(quaternion, axes="sxyz"):
    return euler_from_matrix(quaternion_matrix(quaternion), axes)
