"""Put this directory on sys.path to make ``import pose_estimation_module as PEM`` (the reference's import,
/root/reference/scripts/visual_odometry_v3.py:14) resolve to the ROS-free restatement."""
from droplet_visual_odometry_b200.pose_estimation_module import *  # noqa: F401,F403
