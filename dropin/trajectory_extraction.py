#!/usr/bin/env python
"""ROS-free counterpart of the reference's offline driver ``trajectory_evaluation_dual_process.py`` (VO half, :170-290):
a folder of frames instead of a rosbag, the same stamped_traj_estimate_{absolute,relative,velocity}.txt files out.

    python dropin/trajectory_extraction.py <frames_folder> <calibration.yaml> <out_dir> [--nfeatures 500] [--batch 32]
                                           [--controlled] [--undistort] [--color]
    torchrun --nproc-per-node N dropin/trajectory_extraction.py ...      (frame pairs sharded over N GPUs)

``calibration.yaml`` uses the reference's two schemas (visual_odometry_v3.py:145-166): ``intrinsic_coeffs`` /
``distortion_coeffs`` (default) or ``camera_matrix`` / ``distortion_coefficients`` (``--controlled``).  ``--undistort``
feeds distorted frames through the GPU ingest (cv.undistort with getOptimalNewCameraMatrix(alpha=1), as
ros_img_msg_to_opencv_image does, :115-135); without it frames are taken as already undistorted, like the images the
reference hands to visual_odometry_calculations.  Like the reference (:297-306), findEssentialMat / recoverPose get the
ORIGINAL camera matrix even after undistorting to the new one; ``--pose-with-new-camera-matrix`` uses the new matrix
instead (geometrically consistent, but not what the reference computes).
"""
import argparse
import os
import sys

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200 import sequence as S  # noqa: E402


def read_calibration(path, controlled):
    with open(path) as f:
        data = yaml.safe_load(f)
    if controlled:
        K = np.array(data["camera_matrix"]["data"], dtype=np.float64).reshape(3, 3)
        D = np.array(data["distortion_coefficients"]["data"], dtype=np.float64).ravel()
    else:
        K = np.array(data["intrinsic_coeffs"][0], dtype=np.float64).reshape(3, 3)
        D = np.array(data["distortion_coeffs"][0], dtype=np.float64).ravel()
    return K, D


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("frames")
    ap.add_argument("calibration")
    ap.add_argument("out_dir")
    ap.add_argument("--nfeatures", type=int, default=500)      # cv.ORB_create() default, visual_odometry_v3.py:96
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--controlled", action="store_true")
    ap.add_argument("--undistort", action="store_true")
    ap.add_argument("--color", action="store_true", help="decode frames as BGR (grey conversion then happens on the GPU)")
    ap.add_argument("--pose-with-new-camera-matrix", action="store_true",
                    help="with --undistort: pass getOptimalNewCameraMatrix's result to the pose stage (the reference passes the original K)")
    a = ap.parse_args(argv)

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, D = read_calibration(a.calibration, a.controlled)
    frames = S.FrameFolder(a.frames, color=a.color)
    undistort, Kpose = None, K
    if a.undistort:
        import cv2 as cv      # one-off host call, as the reference makes it (:117-123); the per-frame remap runs on the GPU
        h, w = frames[0].shape[:2]
        new_K, _ = cv.getOptimalNewCameraMatrix(K, D, (w, h), 1, (w, h))
        undistort = (K, D, new_K)
        if a.pose_with_new_camera_matrix:
            Kpose = new_K
    elif a.color:
        raise SystemExit("--color needs --undistort (the GPU ingest does the grey conversion)")
    rec, paths = S.extract_trajectory(frames, Kpose, a.out_dir, nfeatures=a.nfeatures, batch=a.batch, device=local,
                                      world_size=world, rank=rank, group=group, undistort=undistort)
    if rank == 0:
        ok = int((rec["status"] == 0).sum())
        print("%d frames, %d/%d pairs solved; wrote %s" % (len(frames), ok, len(rec), ", ".join(sorted(paths.values()))))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
