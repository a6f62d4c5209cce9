"""Deterministic synthetic scenes for the VO hot path.

Input generator shared by the parity tests and ``bench.py``: a textured piecewise-planar room
(floor + back wall + left wall) seen by a pinhole camera that moves on a smooth closed path, so
that every consecutive frame pair has a known ground-truth relative motion ``(R, t_hat)`` in the
convention of ``cv2.recoverPose`` (``x_cur = R x_prev + t``).

The camera matrix is the reference's calibration
(``/root/reference/Parameters/camera_calibration.yaml:29``: fx 1173.854081, fy 1170.565083) with the
principal point moved to the centre of the rendered size, no distortion.

Rendering is plain ray casting against three planes with bilinear texture lookup, written with torch
tensor ops so that the same code renders a handful of frames on the CPU for tests and a
1000-frame sequence on the GPU for the bench.  Nothing here is on the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

REF_FX = 1173.854081
REF_FY = 1170.565083
TEXTURE_SEED = 20261018
TEXTURE_SIZE = 4096
TEXEL_METRES = 0.005  # one texel is 5 mm on every plane

_texture_cache: dict = {}


def camera_matrix(width: int, height: int) -> np.ndarray:
    """K of the reference calibration with (cx, cy) at the image centre."""
    return np.array([[REF_FX, 0.0, width / 2.0], [0.0, REF_FY, height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def make_texture(seed: int = TEXTURE_SEED, size: int = TEXTURE_SIZE) -> np.ndarray:
    """u8 (size, size): four octaves of smooth value noise + ~12000 constant-grey rectangles."""
    key = (seed, size)
    if key in _texture_cache:
        return _texture_cache[key]
    rng = np.random.default_rng(seed)
    acc = torch.zeros(1, 1, size, size, dtype=torch.float32)
    for cells, weight in ((size // 256, 1.0), (size // 64, 0.6), (size // 16, 0.4), (size // 4, 0.25)):
        grid = torch.from_numpy(rng.random((1, 1, cells, cells), dtype=np.float32))
        acc += weight * torch.nn.functional.interpolate(grid, size=(size, size), mode="bilinear", align_corners=False)
    acc -= acc.min()
    acc /= acc.max()
    tex = (acc[0, 0] * 255.0).round().to(torch.uint8).numpy().copy()
    n_rect = 12000 * (size * size) // (4096 * 4096) + 16
    xs = rng.integers(0, size - 48, n_rect)
    ys = rng.integers(0, size - 48, n_rect)
    ws = rng.integers(8, 41, n_rect)
    hs = rng.integers(8, 41, n_rect)
    gs = rng.integers(0, 256, n_rect)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        tex[y:y + h, x:x + w] = g
    _texture_cache[key] = tex
    return tex


def _rot_xyz(ax: float, ay: float, az: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return rz @ ry @ rx


@dataclass
class CameraPose:
    """World -> camera: ``x_cam = R (x_world - C)``."""
    R: np.ndarray
    C: np.ndarray


def trajectory(n_frames: int, period: int = 48, start_index: int = 0) -> list:
    """Smooth path: a 1.5 m circle in the x-z plane (one lap per ``period`` frames) with small yaw/pitch/roll
    oscillation, plus slow incommensurate drifts so that no two frames of a long sequence are identical.

    Per step: translation ~0.2 m (2-6 % of the 3.4-10 m scene depth, i.e. depth/baseline < 50 = cv2's recoverPose
    distance threshold for most points), rotation < 0.5 deg.
    """
    poses = []
    for i in range(start_index, start_index + n_frames):
        th = 2.0 * math.pi * i / period
        C = np.array([1.5 * math.sin(th) + 0.25 * math.sin(2.0 * math.pi * i / 1013.0),
                      0.15 * math.sin(2.0 * th) + 0.1 * math.sin(2.0 * math.pi * i / 733.0),
                      1.5 * (1.0 - math.cos(th))], dtype=np.float64)
        yaw = math.radians(3.0) * math.sin(th + 0.7) + math.radians(1.0) * math.sin(2.0 * math.pi * i / 911.0)
        pitch = math.radians(1.5) * math.sin(2.0 * th + 0.2)
        roll = math.radians(2.0) * math.cos(th)
        poses.append(CameraPose(R=_rot_xyz(pitch, yaw, roll), C=C))
    return poses


def relative_motion(prev: CameraPose, cur: CameraPose):
    """Ground truth in cv2.recoverPose's convention: x_cur = R x_prev + t (t returned unit-norm)."""
    R = cur.R @ prev.R.T
    t = cur.R @ (prev.C - cur.C)
    n = np.linalg.norm(t)
    return R, (t / n if n > 0 else t)


# planes: (normal n, offset d) with n.x = d, and the two in-plane axes used for texture coordinates
_PLANES = (
    (np.array([0.0, 1.0, 0.0]), 1.5, np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]), (37.0, 11.0)),      # floor y = 1.5
    (np.array([0.0, 0.0, 1.0]), 10.0, np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]), (5.0, 1301.0)),    # back wall z = 10
    (np.array([-1.0, 0.0, 0.0]), 4.0, np.array([0.0, 0.0, 1.0]), np.array([0.0, 1.0, 0.0]), (2203.0, 517.0)),  # left wall x = -4
)


def render_frame(pose: CameraPose, width: int, height: int, texture: torch.Tensor, device=None) -> torch.Tensor:
    """Ray-cast one u8 (height, width) frame.  ``texture`` is a float32 (S, S) tensor on ``device``."""
    device = texture.device if device is None else device
    K = camera_matrix(width, height)
    S = texture.shape[0]
    u = torch.arange(width, device=device, dtype=torch.float32)
    v = torch.arange(height, device=device, dtype=torch.float32)
    xn = ((u - float(K[0, 2])) / float(K[0, 0]))[None, :].expand(height, width)
    yn = ((v - float(K[1, 2])) / float(K[1, 1]))[:, None].expand(height, width)
    Rt = pose.R.T  # camera -> world
    dirs = [Rt[r, 0] * xn + Rt[r, 1] * yn + Rt[r, 2] for r in range(3)]
    best_t = torch.full((height, width), float("inf"), device=device, dtype=torch.float32)
    tex_u = torch.zeros((height, width), device=device, dtype=torch.float32)
    tex_v = torch.zeros((height, width), device=device, dtype=torch.float32)
    for n, d, a0, a1, (ou, ov) in _PLANES:
        denom = float(n[0]) * dirs[0] + float(n[1]) * dirs[1] + float(n[2]) * dirs[2]
        num = d - float(n @ pose.C)
        t = num / denom
        ok = (t > 1e-3) & (t < best_t)
        pu = sum((float(pose.C[r]) + t * dirs[r]) * float(a0[r]) for r in range(3)) / TEXEL_METRES + ou
        pv = sum((float(pose.C[r]) + t * dirs[r]) * float(a1[r]) for r in range(3)) / TEXEL_METRES + ov
        best_t = torch.where(ok, t, best_t)
        tex_u = torch.where(ok, pu, tex_u)
        tex_v = torch.where(ok, pv, tex_v)
    tex_u = torch.remainder(tex_u, float(S))
    tex_v = torch.remainder(tex_v, float(S))
    u0 = torch.floor(tex_u)
    v0 = torch.floor(tex_v)
    fu = tex_u - u0
    fv = tex_v - v0
    u0 = u0.long() % S
    v0 = v0.long() % S
    u1 = (u0 + 1) % S
    v1 = (v0 + 1) % S
    flat = texture.reshape(-1)
    p00 = flat[v0 * S + u0]
    p01 = flat[v0 * S + u1]
    p10 = flat[v1 * S + u0]
    p11 = flat[v1 * S + u1]
    val = (p00 * (1 - fu) + p01 * fu) * (1 - fv) + (p10 * (1 - fu) + p11 * fu) * fv
    return val.round().clamp_(0, 255).to(torch.uint8)


def marker_corners(pose: CameraPose, width: int, height: int, marker_length: float = 0.4, centre=(0.3, -0.2, 10.0)) -> np.ndarray:
    """Pixel coordinates (4, 2) float64 of the corners of a square fiducial of side ``marker_length`` on the back wall, in the
    order a STag detector reports them (clockwise from top-left): the input get_scaling_factor_from_triangulation expects
    (/root/reference/scripts/traj_eval_ground_truth.py:303-311 hands such (n, 2) arrays to visual_odometry_calculations)."""
    K = camera_matrix(width, height)
    h = marker_length / 2.0
    cx, cy, cz = centre
    world = np.array([[cx - h, cy - h, cz], [cx + h, cy - h, cz], [cx + h, cy + h, cz], [cx - h, cy + h, cz]], dtype=np.float64)
    cam = (pose.R @ (world - pose.C).T).T
    uv = (K @ cam.T).T
    return uv[:, :2] / uv[:, 2:3]


def render_sequence(n_frames: int, width: int = 1280, height: int = 1024, device="cpu", period: int = 48,
                    seed: int = TEXTURE_SEED, texture_size: int = TEXTURE_SIZE, start_index: int = 0):
    """Returns (frames u8 (n, H, W) tensor on ``device``, poses, K)."""
    tex = torch.from_numpy(make_texture(seed, texture_size)).to(device=device, dtype=torch.float32)
    poses = trajectory(n_frames, period, start_index)
    frames = torch.empty((n_frames, height, width), dtype=torch.uint8, device=device)
    for i, p in enumerate(poses):
        frames[i] = render_frame(p, width, height, tex)
    return frames, poses, camera_matrix(width, height)


def synthetic_correspondences(n: int, outlier_frac: float = 0.4, noise_px: float = 0.3, seed: int = 7,
                              width: int = 1280, height: int = 1024):
    """Config-5 input: n 3-D points at depth 4-12 seen from two poses, Gaussian pixel noise,
    the first ``outlier_frac`` rows replaced by uniform image points, rows shuffled with the seeded rng.

    Returns (p_prev f32 (n,2), p_cur f32 (n,2), K, R_gt, t_gt_unit, inlier_truth bool (n,)).
    """
    rng = np.random.default_rng(seed)
    K = camera_matrix(width, height)
    R = _rot_xyz(math.radians(0.4), math.radians(-0.3), math.radians(0.25))
    t = np.array([0.12, -0.03, 0.05])
    p1 = np.empty((n, 2))
    p2 = np.empty((n, 2))
    filled = 0
    while filled < n:
        m = (n - filled) * 2 + 16
        uv = np.stack([rng.uniform(0, width, m), rng.uniform(0, height, m)], 1)
        z = rng.uniform(4.0, 12.0, m)
        X = np.stack([(uv[:, 0] - K[0, 2]) / K[0, 0] * z, (uv[:, 1] - K[1, 2]) / K[1, 1] * z, z], 1)
        X2 = X @ R.T + t
        uv2 = np.stack([X2[:, 0] / X2[:, 2] * K[0, 0] + K[0, 2], X2[:, 1] / X2[:, 2] * K[1, 1] + K[1, 2]], 1)
        ok = (uv2[:, 0] >= 0) & (uv2[:, 0] < width) & (uv2[:, 1] >= 0) & (uv2[:, 1] < height) & (X2[:, 2] > 0)
        k = min(int(ok.sum()), n - filled)
        p1[filled:filled + k] = uv[ok][:k]
        p2[filled:filled + k] = uv2[ok][:k]
        filled += k
    p1 += rng.normal(0, noise_px, p1.shape)
    p2 += rng.normal(0, noise_px, p2.shape)
    n_out = int(round(outlier_frac * n))
    p2[:n_out] = np.stack([rng.uniform(0, width, n_out), rng.uniform(0, height, n_out)], 1)
    truth = np.ones(n, dtype=bool)
    truth[:n_out] = False
    perm = rng.permutation(n)
    return (p1[perm].astype(np.float32), p2[perm].astype(np.float32), K, R, t / np.linalg.norm(t), truth[perm])
