"""Import shim: put this directory on sys.path (or copy the two shim files next to the reference's scripts) and the
reference's own ``from visual_odometry_v3 import VisualOdometry``
(/root/reference/scripts/trajectory_evaluation_dual_process.py:21) resolves to the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200.visual_odometry_v3 import *  # noqa: F401,F403,E402
from droplet_visual_odometry_b200.visual_odometry_v3 import VisualOdometry, PairEngine, KeyPoint, DMatch  # noqa: F401,E402
