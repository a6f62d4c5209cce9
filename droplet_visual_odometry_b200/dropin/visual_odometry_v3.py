"""Put this directory on sys.path to make ``from visual_odometry_v3 import VisualOdometry`` (the reference's import,
/root/reference/scripts/trajectory_evaluation_dual_process.py:21) resolve to the B200 implementation."""
from droplet_visual_odometry_b200.visual_odometry_v3 import *  # noqa: F401,F403
from droplet_visual_odometry_b200.visual_odometry_v3 import VisualOdometry, KeyPoint, DMatch, PairEngine  # noqa: F401
