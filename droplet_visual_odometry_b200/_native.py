"""ctypes binding of libdvo.so (include/dvo.h).

The shared library is the product: hand-written sm_100a CUDA behind a C ABI.  This module only moves pointers --
PyTorch supplies device buffers and streams -- and it fails loudly when the library is missing or no CUDA device is
usable.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdvo.so")

DVO_MATCH_CROSSCHECK = 0
DVO_MATCH_KNN_RATIO = 1
PAIR_OK, PAIR_TOO_FEW_MATCHES, PAIR_NO_MODEL = 0, 1, 2
FRAME_TIES_TRUNCATED, FRAME_CANDIDATES_TRUNCATED, FRAME_KEYPOINTS_TRUNCATED = 1, 2, 4
DVO_E_CAPACITY = -3


class DvoError(RuntimeError):
    pass


def describe_frame_flags(f):
    names = [(FRAME_TIES_TRUNCATED, "ties at the Harris boundary truncated"), (FRAME_CANDIDATES_TRUNCATED, "FAST candidate list truncated"),
             (FRAME_KEYPOINTS_TRUNCATED, "keypoint capacity exceeded")]
    return ", ".join(n for b, n in names if f & b) or "none"


def triangulate_points(P0, P1, pts0, pts1):
    """cv.triangulatePoints(P0, P1, pts0.T, pts1.T) -> (4, n), computed by libdvo's host restatement of OpenCV's DLT + Jacobi
    SVD: unnormalised homogeneous points with cv2's sign (visual_odometry_v3.py:265).  float64 in, float64 out."""
    lib = load_library()
    P0 = np.ascontiguousarray(np.asarray(P0, dtype=np.float64).reshape(3, 4))
    P1 = np.ascontiguousarray(np.asarray(P1, dtype=np.float64).reshape(3, 4))
    a = np.ascontiguousarray(np.asarray(pts0, dtype=np.float64).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(pts1, dtype=np.float64).reshape(-1, 2))
    if len(a) != len(b):
        raise ValueError("triangulate_points: the two point sets differ in length")
    X = np.zeros((4, len(a)), dtype=np.float64)
    rc = lib.dvo_triangulate_points_host(P0.ctypes.data, P1.ctypes.data, a.ctypes.data, b.ctypes.data, len(a), X.ctypes.data)
    if rc != 0:
        raise DvoError("dvo_triangulate_points_host failed (%d)" % rc)
    return X


PEAK_NAMES = ("fp32_fma_flops", "fp32_mul_add_flops", "fp64_fma_flops", "fp64_mul_add_flops", "popc_per_s", "int8_tensor_ops")


def measure_peaks(device=0):
    """Pipe rates of `device` from libdvo's microbenchmarks (dvo_measure_peaks): dict name -> ops per second."""
    _torch()
    lib = load_library()
    out = np.zeros(len(PEAK_NAMES), dtype=np.float64)
    rc = lib.dvo_measure_peaks(int(device), out.ctypes.data, len(out))
    if rc != 0:
        raise DvoError("dvo_measure_peaks failed (%d)" % rc)
    return dict(zip(PEAK_NAMES, (float(v) for v in out)))


class dvo_config(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("nfeatures", ctypes.c_int), ("nlevels", ctypes.c_int),
                ("fast_threshold", ctypes.c_int), ("max_frames", ctypes.c_int), ("matcher", ctypes.c_int),
                ("ransac_max_iters", ctypes.c_int), ("ransac_prob", ctypes.c_double), ("ransac_threshold", ctypes.c_double),
                ("distance_thresh", ctypes.c_double), ("ratio", ctypes.c_float), ("use_tma", ctypes.c_int), ("pipeline", ctypes.c_int),
                ("ransac_exhaustive", ctypes.c_int), ("nn_engine", ctypes.c_int)]


class dvo_features(ctypes.Structure):
    _fields_ = [("d_pt", ctypes.c_void_p), ("d_size", ctypes.c_void_p), ("d_angle", ctypes.c_void_p),
                ("d_response", ctypes.c_void_p), ("d_octave", ctypes.c_void_p), ("d_desc", ctypes.c_void_p),
                ("d_count", ctypes.c_void_p), ("capacity", ctypes.c_int32)]


class dvo_pair_arrays(ctypes.Structure):
    _fields_ = [("d_matches", ctypes.c_void_p), ("d_pts_prev", ctypes.c_void_p), ("d_pts_cur", ctypes.c_void_p),
                ("d_ransac_mask", ctypes.c_void_p), ("d_pose_mask", ctypes.c_void_p), ("capacity", ctypes.c_int32)]


POSE_DTYPE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,)), ("E", "<f8", (9,)), ("status", "<i4"), ("n_matches", "<i4"),
                       ("n_inliers", "<i4"), ("n_good", "<i4"), ("ransac_iters", "<i4"), ("best_iter", "<i4"),
                       ("candidate", "<i4"), ("n_prev", "<i4"), ("n_cur", "<i4"), ("frame_flags", "<i4")])
assert POSE_DTYPE.itemsize == 208

_lib = None


def load_library():
    """Load libdvo.so; raise if it was not built (``python -c 'import __graft_entry__ as g; g.build()'`` or ./build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DvoError("libdvo.so not found at %s -- build it with ./build.sh (nvcc, sm_100a); there is no CPU fallback" % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.dvo_default_config.argtypes = [ctypes.POINTER(dvo_config)]
    lib.dvo_default_config.restype = None
    lib.dvo_create.argtypes = [ctypes.POINTER(dvo_config), ci, ctypes.POINTER(vp)]
    lib.dvo_destroy.argtypes = [vp]
    lib.dvo_destroy.restype = None
    lib.dvo_last_error.argtypes = [vp]
    lib.dvo_last_error.restype = ctypes.c_char_p
    lib.dvo_version.restype = ctypes.c_char_p
    lib.dvo_max_keypoints.argtypes = [vp]
    lib.dvo_max_frames.argtypes = [vp]
    lib.dvo_kernel_launches.argtypes = [vp]
    lib.dvo_kernel_launches.restype = ctypes.c_longlong
    lib.dvo_load_frames.argtypes = [vp, vp, ci, cs, cs, ci, ci, vp]
    lib.dvo_set_undistort.argtypes = [vp, vp, vp, ci, vp, ci, vp]
    lib.dvo_orb.argtypes = [vp, ci, ci, vp]
    lib.dvo_get_features.argtypes = [vp, ci, ctypes.POINTER(dvo_features), vp]
    lib.dvo_pairs.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.dvo_match.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.dvo_pose_pairs.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.dvo_get_match_count.argtypes = [vp, ci, ctypes.POINTER(ci), vp]
    lib.dvo_get_frame_flags.argtypes = [vp, ci, ci, vp, vp]
    lib.dvo_triangulate_points_host.argtypes = [vp, vp, vp, vp, ci, vp]
    lib.dvo_measure_peaks.argtypes = [ci, vp, ci]
    lib.dvo_set_features.argtypes = [vp, ci, vp, vp, ci, ci, vp]
    lib.dvo_pose_points.argtypes = [vp, ci, vp, vp, ci, vp, ci, vp]
    lib.dvo_get_poses.argtypes = [vp, ci, ci, vp, ci, vp]
    lib.dvo_get_pair_arrays.argtypes = [vp, ci, ctypes.POINTER(dvo_pair_arrays), vp]
    lib.dvo_sequence.argtypes = [vp, vp, ci, cs, cs, vp, vp, ci, vp]
    lib.dvo_sequence_step.argtypes = [vp, vp, ci, cs, cs, vp, vp, ci, ci, vp]
    lib.dvo_sequence_flush.argtypes = [vp, vp]
    lib.dvo_profile_enable.argtypes = [ci]
    lib.dvo_profile_enable.restype = None
    lib.dvo_profile_collect.argtypes = [vp, vp, ci]
    lib.dvo_profile_name.argtypes = [ci]
    lib.dvo_profile_name.restype = ctypes.c_char_p
    lib.dvo_sizeof.argtypes = [ci]
    lib.dvo_level_size.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.dvo_tap_image.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.dvo_tap_candidates.argtypes = [vp, ci, ci, vp, ci, ctypes.POINTER(ci), vp]
    lib.dvo_tap_ransac.argtypes = [vp, ci, vp, vp]
    _lib = lib
    return lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise DvoError("no CUDA device visible: the visual-odometry hot path runs only on the GPU (no CPU fallback)")
    return torch


class Context:
    """One libdvo context bound to one GPU.  All methods enqueue on torch's current stream for that device."""

    def __init__(self, width, height, nfeatures=500, nlevels=8, max_frames=2, matcher=DVO_MATCH_CROSSCHECK,
                 ransac_max_iters=1000, ransac_prob=0.999, ransac_threshold=1.0, distance_thresh=50.0, ratio=0.75,
                 fast_threshold=20, device=0, use_tma=True, pipeline=True, ransac_exhaustive=False, nn_engine=0):
        self.lib = load_library()
        self.torch = _torch()
        cfg = dvo_config()
        self.lib.dvo_default_config(ctypes.byref(cfg))
        cfg.width, cfg.height, cfg.nfeatures, cfg.nlevels = int(width), int(height), int(nfeatures), int(nlevels)
        cfg.max_frames, cfg.matcher, cfg.ransac_max_iters = int(max_frames), int(matcher), int(ransac_max_iters)
        cfg.ransac_prob, cfg.ransac_threshold, cfg.distance_thresh = float(ransac_prob), float(ransac_threshold), float(distance_thresh)
        cfg.ratio, cfg.fast_threshold, cfg.use_tma = float(ratio), int(fast_threshold), int(bool(use_tma))
        cfg.pipeline = int(bool(pipeline))
        cfg.ransac_exhaustive = int(bool(ransac_exhaustive))
        cfg.nn_engine = int(nn_engine)
        self.cfg = cfg
        self.device = int(device)
        self.width, self.height, self.nlevels = int(width), int(height), int(nlevels)
        h = ctypes.c_void_p()
        rc = self.lib.dvo_create(ctypes.byref(cfg), self.device, ctypes.byref(h))
        self._h = h
        if rc != 0:
            msg = self.lib.dvo_last_error(h).decode() if h else "invalid configuration or no device"
            if h:
                self.lib.dvo_destroy(h)
            self._h = None
            raise DvoError("dvo_create failed (%d): %s" % (rc, msg))
        self.max_keypoints = self.lib.dvo_max_keypoints(h)
        self.max_frames = self.lib.dvo_max_frames(h)
        self.tdev = self.torch.device("cuda", self.device)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dvo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise DvoError("%s failed (%d): %s" % (what, rc, self.lib.dvo_last_error(self._h).decode()))

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def flush(self):
        """Join torch's current stream with the pipelined sequence runner's internal streams (dvo_sequence_flush)."""
        self._check(self.lib.dvo_sequence_flush(self._h, self._stream()), "dvo_sequence_flush")

    def sync(self):
        self.flush()
        self.torch.cuda.current_stream(self.tdev).synchronize()

    @property
    def kernel_launches(self):
        return int(self.lib.dvo_kernel_launches(self._h))

    def level_size(self, level):
        w, h, q = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.dvo_level_size(self._h, level, ctypes.byref(w), ctypes.byref(h), ctypes.byref(q)), "dvo_level_size")
        return w.value, h.value, q.value

    # ------------------------------------------------------------------ stages
    def set_undistort(self, K=None, dist=None, new_K=None, channels=1):
        """Ingest while loading: cv.cvtColor(BGR2GRAY) (channels=3) + cv.undistort(K, dist, new_K) on the GPU, bit-exact
        with cv2.  channels=0 (or K=None) switches it off: frames are then grey and already undistorted."""
        if K is None or channels == 0:
            self._check(self.lib.dvo_set_undistort(self._h, None, None, 0, None, 0, self._stream()), "dvo_set_undistort")
            self.channels = 1
            return
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        Nc = np.ascontiguousarray(np.asarray(new_K, dtype=np.float64).reshape(9))
        Dc = np.ascontiguousarray(np.asarray(dist if dist is not None else [], dtype=np.float64).ravel()[:8])
        self._check(self.lib.dvo_set_undistort(self._h, Kc.ctypes.data, Dc.ctypes.data if Dc.size else None, int(Dc.size), Nc.ctypes.data,
                                               int(channels), self._stream()), "dvo_set_undistort")
        self.channels = int(channels)

    def _frame_layout(self, frames):
        """(row pitch, frame stride) in bytes of a contiguous (n, H, W[, 3]) uint8 batch, checked against the ingest mode."""
        ch = getattr(self, "channels", 1)
        want = (self.height, self.width) if ch == 1 else (self.height, self.width, 3)
        assert tuple(frames.shape[1:]) == want, "frames must be (n, %s) uint8" % (", ".join(map(str, want)))
        return self.width * ch, self.width * self.height * ch

    def load_frames(self, frames, slot0=0):
        """frames: (n, H, W) uint8 -- or (n, H, W, 3) BGR after set_undistort(channels=3) -- torch tensor (cuda or cpu) or
        numpy array (host)."""
        t = self.torch
        if isinstance(frames, np.ndarray):
            frames = t.from_numpy(np.ascontiguousarray(frames))
        if frames.dim() == (2 if getattr(self, "channels", 1) == 1 else 3):
            frames = frames[None]
        assert frames.dtype == t.uint8
        frames = frames.contiguous()
        pitch, stride = self._frame_layout(frames)
        kind = 0 if frames.is_cuda else 1
        self._check(self.lib.dvo_load_frames(self._h, frames.data_ptr(), frames.shape[0], pitch, stride,
                                             slot0, kind, self._stream()), "dvo_load_frames")
        if kind == 1:
            self.sync()   # host source must outlive the copy
        return frames.shape[0]

    def orb(self, slot0, n):
        self._check(self.lib.dvo_orb(self._h, slot0, n, self._stream()), "dvo_orb")

    def frame_flags(self, slot0=0, n=1):
        """FRAME_* bits of slots [slot0, slot0+n) (synchronises); non-zero = the slot's keypoint set was truncated."""
        out = np.zeros(n, dtype=np.int32)
        self._check(self.lib.dvo_get_frame_flags(self._h, slot0, n, out.ctypes.data, self._stream()), "dvo_get_frame_flags")
        return out

    def check_frames(self, slot0=0, n=1):
        """Raise (never return a truncated feature set silently) when a slot could not hold everything cv2 keeps."""
        f = self.frame_flags(slot0, n)
        if f.any():
            bad = int(np.flatnonzero(f)[0])
            raise DvoError("DVO_E_CAPACITY (%d): frame slot %d holds a truncated keypoint set (flags 0x%x: %s) -- cv2 keeps every "
                           "tie at the retainBest boundary and this image has more of them than quota + 64 per level"
                           % (DVO_E_CAPACITY, slot0 + bad, int(f[bad]), describe_frame_flags(int(f[bad]))))

    def features(self, slot):
        """Host copies of one slot's detectAndCompute output (synchronises)."""
        t, M = self.torch, self.max_keypoints
        pt = t.empty((M, 2), dtype=t.float32, device=self.tdev)
        size = t.empty(M, dtype=t.float32, device=self.tdev)
        angle = t.empty(M, dtype=t.float32, device=self.tdev)
        resp = t.empty(M, dtype=t.float32, device=self.tdev)
        octave = t.empty(M, dtype=t.int32, device=self.tdev)
        desc = t.empty((M, 32), dtype=t.uint8, device=self.tdev)
        count = t.zeros(1, dtype=t.int32, device=self.tdev)
        f = dvo_features(pt.data_ptr(), size.data_ptr(), angle.data_ptr(), resp.data_ptr(), octave.data_ptr(), desc.data_ptr(),
                         count.data_ptr(), M)
        self._check(self.lib.dvo_get_features(self._h, slot, ctypes.byref(f), self._stream()), "dvo_get_features")
        n = int(count.item())
        self.check_frames(slot, 1)
        return {"pt": pt[:n].cpu().numpy(), "size": size[:n].cpu().numpy(), "angle": angle[:n].cpu().numpy(),
                "response": resp[:n].cpu().numpy(), "octave": octave[:n].cpu().numpy(), "desc": desc[:n].cpu().numpy()}

    def pairs(self, slot0, pair0, n, K):
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        self._check(self.lib.dvo_pairs(self._h, slot0, pair0, n, Kc.ctypes.data, self._stream()), "dvo_pairs")

    def match(self, slot0, pair0, n, K):
        """bf.match + sorted + KeyPoint_convert only (dvo_match); read the result with match_count + pair_arrays."""
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        self._check(self.lib.dvo_match(self._h, slot0, pair0, n, Kc.ctypes.data, self._stream()), "dvo_match")

    def pose(self, slot0, pair0, n, K):
        """findEssentialMat + recoverPose on the correspondences dvo_match left in the pair slots (dvo_pose_pairs)."""
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        self._check(self.lib.dvo_pose_pairs(self._h, slot0, pair0, n, Kc.ctypes.data, self._stream()), "dvo_pose_pairs")

    def match_count(self, pair=0):
        c = ctypes.c_int()
        self._check(self.lib.dvo_get_match_count(self._h, pair, ctypes.byref(c), self._stream()), "dvo_get_match_count")
        return c.value

    def set_features(self, slot, pt, desc):
        """Overwrite a slot with caller keypoint coordinates (n,2) f32 and descriptors (n,32) u8 (host arrays)."""
        pt = np.ascontiguousarray(pt, dtype=np.float32).reshape(-1, 2)
        desc = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
        assert len(pt) == len(desc)
        self._check(self.lib.dvo_set_features(self._h, slot, pt.ctypes.data, desc.ctypes.data, len(pt), 1, self._stream()),
                    "dvo_set_features")

    def pose_points(self, p_prev, p_cur, K, pair=0):
        """findEssentialMat + recoverPose on caller correspondences (host (n,2) arrays or cuda float32 tensors)."""
        t = self.torch
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        if isinstance(p_prev, np.ndarray) or not hasattr(p_prev, "is_cuda"):
            a = np.ascontiguousarray(p_prev, dtype=np.float32).reshape(-1, 2)
            b = np.ascontiguousarray(p_cur, dtype=np.float32).reshape(-1, 2)
            self._check(self.lib.dvo_pose_points(self._h, pair, a.ctypes.data, b.ctypes.data, len(a), Kc.ctypes.data, 1,
                                                 self._stream()), "dvo_pose_points")
            self.sync()
            return len(a)
        a = p_prev.contiguous().to(t.float32)
        b = p_cur.contiguous().to(t.float32)
        self._check(self.lib.dvo_pose_points(self._h, pair, a.data_ptr(), b.data_ptr(), a.shape[0], Kc.ctypes.data, 0,
                                             self._stream()), "dvo_pose_points")
        return a.shape[0]

    def poses(self, pair0, n):
        out = np.zeros(n, dtype=POSE_DTYPE)
        self._check(self.lib.dvo_get_poses(self._h, pair0, n, out.ctypes.data, 1, self._stream()), "dvo_get_poses")
        self.sync()
        return out

    def pair_arrays(self, pair, n_matches):
        t, M = self.torch, self.max_keypoints
        matches = t.empty((M, 3), dtype=t.int32, device=self.tdev)
        p1 = t.empty((M, 2), dtype=t.float32, device=self.tdev)
        p2 = t.empty((M, 2), dtype=t.float32, device=self.tdev)
        rm = t.empty(M, dtype=t.uint8, device=self.tdev)
        pm = t.empty(M, dtype=t.uint8, device=self.tdev)
        a = dvo_pair_arrays(matches.data_ptr(), p1.data_ptr(), p2.data_ptr(), rm.data_ptr(), pm.data_ptr(), M)
        self._check(self.lib.dvo_get_pair_arrays(self._h, pair, ctypes.byref(a), self._stream()), "dvo_get_pair_arrays")
        n = int(n_matches)
        return {"matches": matches[:n].cpu().numpy(), "p_prev": p1[:n].cpu().numpy(), "p_cur": p2[:n].cpu().numpy(),
                "ransac_mask": rm[:n].cpu().numpy(), "pose_mask": pm[:n].cpu().numpy()}

    def sequence(self, frames, K, out=None):
        """Consecutive-pair VO over all frames.  frames: (n, H, W) uint8, cuda tensor (poses come back through a device
        buffer, one D2H at the end) or host tensor/ndarray (copies inside the call).  Returns POSE_DTYPE array (n-1,)."""
        t = self.torch
        if isinstance(frames, np.ndarray):
            frames = t.from_numpy(np.ascontiguousarray(frames))
        frames = frames.contiguous()
        n = frames.shape[0]
        pitch, stride = self._frame_layout(frames)
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        if frames.is_cuda:
            dposes = t.empty((n - 1) * POSE_DTYPE.itemsize, dtype=t.uint8, device=self.tdev) if out is None else out
            self._check(self.lib.dvo_sequence(self._h, frames.data_ptr(), n, pitch, stride, Kc.ctypes.data,
                                              dposes.data_ptr(), 0, self._stream()), "dvo_sequence")
            if out is not None:
                return out
            return dposes.cpu().numpy().view(POSE_DTYPE)
        poses = np.zeros(n - 1, dtype=POSE_DTYPE)
        self._check(self.lib.dvo_sequence(self._h, frames.data_ptr(), n, pitch, stride, Kc.ctypes.data,
                                          poses.ctypes.data, 1, self._stream()), "dvo_sequence")
        return poses

    def sequence_step(self, frames, K, poses_out, first):
        """One batch: frames (n, H, W) uint8 -- cuda tensor with poses_out a cuda uint8 tensor (async), or pinned/pageable
        host tensor with poses_out a POSE_DTYPE ndarray.  Returns the number of records written."""
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        n = frames.shape[0]
        pitch, stride = self._frame_layout(frames)
        if frames.is_cuda:
            rc = self.lib.dvo_sequence_step(self._h, frames.data_ptr(), n, pitch, stride, Kc.ctypes.data,
                                            poses_out.data_ptr(), 0, int(bool(first)), self._stream())
        else:
            rc = self.lib.dvo_sequence_step(self._h, frames.data_ptr(), n, pitch, stride, Kc.ctypes.data,
                                            poses_out.ctypes.data, 1, int(bool(first)), self._stream())
        if rc < 0:
            self._check(rc, "dvo_sequence_step")
        return rc

    def profile(self, on):
        self.lib.dvo_profile_enable(int(bool(on)))

    def profile_collect(self):
        """{kernel name: (total ms, launch groups)} since the previous collect; synchronises."""
        ms = np.zeros(32, dtype=np.float64)
        cnt = np.zeros(32, dtype=np.int32)
        n = self.lib.dvo_profile_collect(ms.ctypes.data, cnt.ctypes.data, 32)
        return {self.lib.dvo_profile_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    # ------------------------------------------------------------------ taps
    def tap_image(self, slot, level, which=0):
        w, h, _ = self.level_size(level)
        dst = self.torch.empty((h, w), dtype=self.torch.uint8, device=self.tdev)
        self._check(self.lib.dvo_tap_image(self._h, slot, level, which, dst.data_ptr(), self._stream()), "dvo_tap_image")
        return dst.cpu().numpy()

    def tap_candidates(self, slot, level):
        w, h, _ = self.level_size(level)
        cap = ((w + 1) // 2) * ((h + 1) // 2)
        dst = self.torch.empty(cap, dtype=self.torch.int32, device=self.tdev)
        cnt = ctypes.c_int()
        self._check(self.lib.dvo_tap_candidates(self._h, slot, level, dst.data_ptr(), cap, ctypes.byref(cnt), self._stream()),
                    "dvo_tap_candidates")
        v = dst[:cnt.value].cpu().numpy().view(np.uint32)
        return np.stack([v & 0xFFF, (v >> 12) & 0xFFF, v >> 24], axis=1).astype(np.int32)

    def tap_ransac(self, pair):
        st = np.zeros(8, dtype=np.int32)
        self._check(self.lib.dvo_tap_ransac(self._h, pair, st.ctypes.data, self._stream()), "dvo_tap_ransac")
        return st
