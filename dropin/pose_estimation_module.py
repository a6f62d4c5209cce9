"""Import shim for ``import pose_estimation_module as PEM``
(/root/reference/scripts/visual_odometry_v3.py:14, trajectory_evaluation_dual_process.py:23)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200.pose_estimation_module import *  # noqa: F401,F403,E402
