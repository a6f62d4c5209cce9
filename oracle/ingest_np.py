"""TEST INFRASTRUCTURE -- CPU restatement of the reference's image ingest after decode
(/root/reference/scripts/visual_odometry_v3.py:115-135): cv.cvtColor(BGR2GRAY) then cv.undistort(gray, K, dist, newK).
cv2 4.13.0 behaviour restated (each step verified bit-exact against cv2 in this container, tests/test_oracle_golden.py):

  BGR2GRAY (8u)      gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15
  cv.undistort       == remap(src, initUndistortRectifyMap(K, dist, I, newK, size, CV_16SC2), INTER_LINEAR, BORDER_CONSTANT 0)
  map (float64)      [x y w]^T = inv(newK) [j i 1]^T ; x/=w, y/=w ; r2 = x^2+y^2 ;
                     kr = (1 + ((k3 r2 + k2) r2 + k1) r2) / (1 + ((k6 r2 + k5) r2 + k4) r2) ;
                     xd = x kr + p1 2xy + p2 (r2 + 2x^2) ; yd = y kr + p1 (r2 + 2y^2) + p2 2xy ;
                     u = fx xd + cx, v = fy yd + cy ; iu = rint(32 u), iv = rint(32 v)
                     integer part (iu >> 5, iv >> 5) saturated to int16, fraction (iu & 31, iv & 31)
  remap (8u, linear) four taps (0 outside the image), int16 weights w[fy][fx][4] = rint(32768 (1-ay|ay)(1-ax|ax)) in float32
                     with the largest (or smallest) adjusted so they sum to 32768 ; out = (sum w p + 2^14) >> 15
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
from __future__ import annotations

import numpy as np


def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def undistort_map(K, dist, newK, width: int, height: int):
    """(iu, iv) int64 arrays: 32 x the source coordinates, rounded to nearest even."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    d = np.zeros(8)
    dd = np.asarray(dist, np.float64).ravel()
    d[:min(8, dd.size)] = dd[:8]
    k1, k2, p1, p2, k3, k4, k5, k6 = d
    ir = np.linalg.inv(np.asarray(newK, np.float64).reshape(3, 3))
    ii = np.arange(height, dtype=np.float64)[:, None]
    jj = np.arange(width, dtype=np.float64)[None, :]
    X = ii * ir[0, 1] + ir[0, 2] + jj * ir[0, 0]
    Y = ii * ir[1, 1] + ir[1, 2] + jj * ir[1, 0]
    W = ii * ir[2, 1] + ir[2, 2] + jj * ir[2, 0]
    w = 1.0 / W
    x, y = X * w, Y * w
    x2, y2 = x * x, y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    u = K[0, 0] * xd + K[0, 2]
    v = K[1, 1] * yd + K[1, 2]
    lim = float(2 ** 31 - 1)
    iu = np.rint(np.clip(u * 32, -lim - 1, lim)).astype(np.int64)
    iv = np.rint(np.clip(v * 32, -lim - 1, lim)).astype(np.int64)
    return iu, iv


def bilinear_weight_table() -> np.ndarray:
    """(32, 32, 4) int32: cv2's BilinearTab_i for INTER_LINEAR (weights of taps 00, 01, 10, 11 for fraction fy, fx)."""
    t = np.zeros((32, 32, 4), np.int32)
    s = np.float32(1.0 / 32)
    for fy in range(32):
        for fx in range(32):
            ay, ax = np.float32(fy) * s, np.float32(fx) * s
            cy, cx = (np.float32(1) - ay, ay), (np.float32(1) - ax, ax)
            it = [int(np.rint(np.float32(np.float32(cy[a] * cx[b]) * np.float32(32768)))) for a in range(2) for b in range(2)]
            diff = sum(it) - 32768
            if diff != 0:
                mk = Mk = 0
                for k in range(4):
                    if it[k] < it[mk]:
                        mk = k
                    elif it[k] > it[Mk]:
                        Mk = k
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            t[fy, fx] = it
    return t


def remap_linear_u8(img: np.ndarray, iu: np.ndarray, iv: np.ndarray) -> np.ndarray:
    h, w = img.shape
    sx = np.clip(iu >> 5, -32768, 32767)
    sy = np.clip(iv >> 5, -32768, 32767)
    W = bilinear_weight_table()[iv & 31, iu & 31].astype(np.int64)

    def px(y, x):
        ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        return np.where(ok, img[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)].astype(np.int64), 0)
    val = (px(sy, sx) * W[..., 0] + px(sy, sx + 1) * W[..., 1] + px(sy + 1, sx) * W[..., 2] + px(sy + 1, sx + 1) * W[..., 3] + (1 << 14)) >> 15
    return np.clip(val, 0, 255).astype(np.uint8)


def ingest(image: np.ndarray, K, dist, newK) -> np.ndarray:
    """BGR (H, W, 3) or grey (H, W) u8 -> undistorted grey u8, as the reference's ros_img_msg_to_opencv_image does."""
    grey = bgr_to_gray(image) if image.ndim == 3 else image
    iu, iv = undistort_map(K, dist, newK, grey.shape[1], grey.shape[0])
    return remap_linear_u8(grey, iu, iv)
