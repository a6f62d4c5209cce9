"""Sequence-level host logic: consecutive-pair VO over an offline sequence, sharded by frame pair across GPUs, chained
into a trajectory and written in the reference's stamped text formats.

Reference flow being replaced (ROS-free): ``UnitTestingExtractData.compute_all_gt_vo_comparison_list``
(/root/reference/scripts/trajectory_evaluation_dual_process.py:170-252) calls ``visual_odometry_calculations`` once per
consecutive pair, sequentially, chains ``cur = prev.dot(rel)`` (visual_odometry_v3.py:367) and finally writes
``stamped_traj_estimate_{absolute,relative,velocity}.txt`` (:280-290, :307-309).

Here pairs are independent units: rank r of G processes a contiguous block of pairs (one-frame halo), one NCCL
all-gather of the fixed-size per-pair records (R, t, status ...: 208 bytes) follows, and every rank chains the
4x4 products.  No other collective exists on this path.
"""
from __future__ import annotations

import os

import numpy as np

from . import pose_estimation_module as PEM
from . import transformations_lite as transf
from ._native import POSE_DTYPE, PAIR_OK


# ------------------------------------------------------------------------------------------------ partitioning
def shard_pairs(n_pairs: int, world_size: int, rank: int):
    """Contiguous block [start, end) of pair indices for ``rank``: ceil(n/G) pairs per rank, last ranks may be short or
    empty.  Pair i uses frames i and i+1, so the rank touches frames [start, end]."""
    per = (n_pairs + world_size - 1) // world_size if world_size > 0 else n_pairs
    start = min(rank * per, n_pairs)
    end = min(start + per, n_pairs)
    return start, end


# ------------------------------------------------------------------------------------------------ pose algebra (host, a12)
def relative_transform(R, t, scale: float = 1.0, exact_rotation: bool = False):
    """4x4 of one pair.  Default reproduces the reference's construction: euler_from_matrix(R, 'rxyz') then
    euler_matrix(..., 'sxyz') (visual_odometry_v3.py:334, :140 -- equal to R only to first order, quirk B.4 kept);
    ``exact_rotation`` puts R itself in the 4x4."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(3) * scale
    if exact_rotation:
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = t
        return M
    M4 = np.eye(4)
    M4[:3, :3] = R
    euler = transf.euler_from_matrix(M4, "rxyz")
    return transf.translation_matrix(t).dot(transf.euler_matrix(euler[0], euler[1], euler[2], axes="sxyz"))


def chain(relatives, start=None):
    """Absolute poses T_0 = start, T_i = T_{i-1} . rel_i (visual_odometry_v3.py:367).  Returns n+1 matrices."""
    cur = np.eye(4) if start is None else np.asarray(start, dtype=np.float64)
    out = [cur]
    for rel in relatives:
        cur = cur.dot(rel)
        out.append(cur)
    return out


def poses_to_relatives(poses, exact_rotation: bool = False):
    """Per-pair 4x4s from dvo_pose records.  A pair whose status is not OK contributes the identity (the reference
    would have raised inside cv.recoverPose; a batch runner must keep going -- SURVEY.md §5)."""
    rel = []
    for p in poses:
        if int(p["status"]) != PAIR_OK:
            rel.append(np.eye(4))
        else:
            rel.append(relative_transform(p["R"], p["t"], 1.0, exact_rotation))
    return rel


# ------------------------------------------------------------------------------------------------ writers (§8f rank 1)
TRAJ_FILES = {"absolute": "stamped_traj_estimate_absolute.txt", "relative": "stamped_traj_estimate_relative.txt",
              "velocity": "stamped_traj_estimate_velocity.txt", "legacy": "stamped_traj_estimate.txt"}


def write_stamped(path, timestamps, transforms):
    """Truncate, then one appended line per transform: ``ts tx ty tz qx qy qz qw \\n``
    (trajectory_evaluation_dual_process.py:93-100 clear, :280-290 write; Shepperd quaternion [x,y,z,w])."""
    PEM.clear_txt_file_contents(path)
    with open(path, "a") as f:
        for ts, T in zip(timestamps, transforms):
            f.write(PEM.format_stamped_line(ts, PEM.translation_from_transformation_matrix(T), PEM.quaternion_from_transformation_matrix(T)))


def write_trajectory(folder, timestamps, poses, start=None, exact_rotation: bool = False):
    """Write the three stamped files of the current driver plus the legacy single file.

    timestamps: n_frames values; poses: n_frames-1 dvo_pose records.  Rows (as the reference appends them,
    trajectory_evaluation_dual_process.py:197-250): absolute has the seed pose then one row per pair; relative and
    velocity have one row per pair stamped with the pair's second frame."""
    os.makedirs(folder, exist_ok=True)
    rel = poses_to_relatives(poses, exact_rotation)
    absolute = chain(rel, start)
    vel = [PEM.get_velocity_between_timestamps(r, timestamps[i], timestamps[i + 1]) for i, r in enumerate(rel)]
    paths = {k: os.path.join(folder, v) for k, v in TRAJ_FILES.items()}
    write_stamped(paths["absolute"], timestamps, absolute)
    write_stamped(paths["relative"], timestamps[1:], rel)
    write_stamped(paths["velocity"], timestamps[1:], vel)
    write_stamped(paths["legacy"], timestamps, absolute)
    return paths


# ------------------------------------------------------------------------------------------------ multi-GPU
def gather_poses(local: np.ndarray, n_pairs: int, world_size: int, rank: int, group=None, device=None):
    """All-gather the per-rank blocks of dvo_pose records into the full (n_pairs,) array on every rank.

    One collective: every rank contributes ceil(n/G) records (short blocks are zero-padded), as raw bytes.  With an
    NCCL group the buffers live on ``device`` (NVLink / NVSwitch); with gloo they stay on the host (CPU tests)."""
    import torch
    import torch.distributed as dist
    per = (n_pairs + world_size - 1) // world_size
    rec = POSE_DTYPE.itemsize
    buf = np.zeros(per, dtype=POSE_DTYPE)
    buf[:len(local)] = local
    send = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world_size * per * rec, dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    allp = recv.cpu().numpy().view(POSE_DTYPE)
    out = np.zeros(n_pairs, dtype=POSE_DTYPE)
    for r in range(world_size):
        s, e = shard_pairs(n_pairs, world_size, r)
        out[s:e] = allp[r * per:r * per + (e - s)]
    return out


def run_sharded(n_frames: int, compute_block, world_size: int = 1, rank: int = 0, group=None, device=None):
    """``compute_block(first_frame, last_frame_inclusive) -> dvo_pose records`` for the rank's frames; returns the full
    (n_frames-1,) record array on every rank.  world_size == 1 needs no process group."""
    n_pairs = n_frames - 1
    s, e = shard_pairs(n_pairs, world_size, rank)
    local = compute_block(s, e) if e > s else np.zeros(0, dtype=POSE_DTYPE)
    if len(local) != e - s:
        raise ValueError("compute_block returned %d records for %d pairs" % (len(local), e - s))
    if world_size == 1:
        return local
    return gather_poses(local, n_pairs, world_size, rank, group, device)


class SequenceRunner:
    """Consecutive-pair VO over device- or host-resident frames on one GPU (the unit each rank runs)."""

    def __init__(self, width, height, K, nfeatures=500, batch=32, device=0, matcher=0, **kw):
        from ._native import Context
        self.K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        self.ctx = Context(width, height, nfeatures=nfeatures, max_frames=batch + 1, matcher=matcher, device=device, **kw)

    def run(self, frames):
        """frames: (n, H, W) uint8 cuda tensor / host tensor / ndarray -> (n-1,) POSE_DTYPE.  Raises when a frame's keypoint set
        had to be truncated (dvo_pose.frame_flags): a trajectory built on features cv2 would not return is not written."""
        rec = self.ctx.sequence(frames, self.K)
        bad = np.flatnonzero(rec["frame_flags"])
        if len(bad):
            from ._native import DvoError, describe_frame_flags
            raise DvoError("DVO_E_CAPACITY: %d frame pair(s) (first: %d) used a truncated keypoint set (%s)"
                           % (len(bad), int(bad[0]), describe_frame_flags(int(rec["frame_flags"][bad[0]]))))
        return rec

    def block_fn(self, frames):
        """compute_block for run_sharded over a frame container indexable by slice (frames[a:b])."""
        def fn(first, last):
            return self.run(frames[first:last + 1])
        return fn


# ------------------------------------------------------------------------------------------------ ROS-free input (§8f rank 4)
class FrameFolder:
    """Frames of an offline sequence from a folder of images (or .npy arrays), in name order -- the ROS-free stand-in for
    the rosbag pairing of get_valid_message_stream.py:21-87 and for CollectImagePaths
    (utilities_folder/traj_eval_unit_vis_odom.py:23-34).  Indexing with a slice decodes just those frames, so a rank only
    touches its own block.  Decoding (cv.imread) is host work outside the hot path; ``color=True`` keeps BGR for the GPU
    ingest (Context.set_undistort(channels=3))."""
    EXT = (".png", ".jpg", ".jpeg", ".bmp", ".pgm", ".tif", ".tiff", ".npy")

    def __init__(self, folder, color: bool = False):
        self.folder = folder
        self.color = color
        self.paths = sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.lower().endswith(self.EXT))
        if not self.paths:
            raise FileNotFoundError("no frames in " + folder)

    def __len__(self):
        return len(self.paths)

    @property
    def timestamps(self):
        """The file stem as a float when it parses (frames saved as <stamp>.png), else the frame index."""
        out = []
        for i, p in enumerate(self.paths):
            stem = os.path.splitext(os.path.basename(p))[0]
            try:
                out.append(float(stem))
            except ValueError:
                out.append(float(i))
        return out

    def _read(self, path):
        if path.lower().endswith(".npy"):
            a = np.load(path)
        else:
            import cv2 as cv      # file decoding only (host ingest, as cv.imdecode in the reference :127); never on the hot path
            a = cv.imread(path, cv.IMREAD_COLOR if self.color else cv.IMREAD_GRAYSCALE)
            if a is None:
                raise IOError("cannot decode " + path)
        a = np.ascontiguousarray(a, dtype=np.uint8)
        if not self.color and a.ndim == 3:
            raise ValueError(path + ": colour frame in a grey sequence (pass color=True and use the GPU ingest)")
        return a

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return np.stack([self._read(p) for p in self.paths[idx]])
        return self._read(self.paths[idx])


def extract_trajectory(frames, K, out_dir, timestamps=None, nfeatures=500, batch=32, device=0, world_size=1, rank=0, group=None,
                       undistort=None, start=None):
    """The VO half of the reference's offline driver (trajectory_evaluation_dual_process.py:170-290) without ROS:
    consecutive-pair VO over ``frames`` (FrameFolder, ndarray or tensor, indexable by slice), sharded by frame pair when
    world_size > 1, chained, and written as the stamped_traj_estimate_* files (rank 0).  ``undistort=(K_raw, dist, new_K)``
    switches the GPU ingest on (frames are then distorted; K must be new_K).  Returns (records, paths or None)."""
    n = len(frames)
    first = frames[0]
    h, w = first.shape[0], first.shape[1]
    runner = SequenceRunner(w, h, K, nfeatures=nfeatures, batch=batch, device=device)
    if undistort is not None:
        runner.ctx.set_undistort(undistort[0], undistort[1], undistort[2], channels=3 if first.ndim == 3 else 1)
    torch_dev = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch_dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else None
    records = run_sharded(n, runner.block_fn(frames), world_size, rank, group, torch_dev)
    paths = None
    if rank == 0:
        ts = list(timestamps) if timestamps is not None else (frames.timestamps if hasattr(frames, "timestamps") else [float(i) for i in range(n)])
        paths = write_trajectory(out_dir, ts, records, start=start)
    runner.ctx.close()
    return records, paths
