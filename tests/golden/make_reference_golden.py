"""Records outputs of the REFERENCE MODULE ITSELF (/root/reference/scripts/visual_odometry_v3.py, imported unmodified by
oracle/ref_loader.py with stub ROS/plot modules) on a small synthetic sequence with fiducial corners, so that the GPU box --
where /root/reference does not exist -- can compare the CUDA drop-in with what the reference's own code returned here.

    python tests/golden/make_reference_golden.py        ->  tests/golden/golden_reference_module.npz

Recorded per consecutive pair of a 3-frame 480x360 sequence (controlled=True, ORB_create() defaults = 500 features,
real_marker_length 0.4): both frames' keypoints/descriptors from compute_current_image_elements (:370-379), the sorted
bf.match list (:219-221), essential_matrix (:297-300), the 4x4 returned by get_transformation_between_two_frames (:293-345,
marker scale and euler round trip included), the chained pose from previous_current_matching (:367) and the projection
matrix carried to the next pair (:344).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from droplet_visual_odometry_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

W, H, N_FRAMES, MARKER = 480, 360, 3, 0.4


def build_inputs():
    frames, poses, K = synth.render_sequence(N_FRAMES + 1, width=W, height=H, device="cpu")
    frames = frames[1:].numpy()
    poses = poses[1:]
    rng = np.random.default_rng(20261019)
    corners = np.stack([synth.marker_corners(p, W, H, MARKER) + rng.normal(scale=0.2, size=(4, 2)) for p in poses])
    return frames, K, corners


def run_reference(frames, K, corners):
    with tempfile.TemporaryDirectory() as d:
        calib = os.path.join(d, "controlled.yaml")
        ref_loader.write_controlled_calibration(calib, K)
        vo = ref_loader.make_vo(calib, MARKER)
        pose = np.array(vo.robot_curr_position)
        out = []
        for i in range(len(frames) - 1):
            r = ref_loader.pair_through_reference(vo, frames[i], frames[i + 1], pose, corners[i], corners[i + 1])
            pose = r["cur_pose"]
            out.append(r)
    return out


def main():
    frames, K, corners = build_inputs()
    res = run_reference(frames, K, corners)
    out = {"frames": frames, "K": K, "corners": corners, "real_marker_length": np.float64(MARKER)}
    for i, r in enumerate(res):
        for k in ("matches", "E", "cur_pose", "rel", "projection"):
            out["p%d_%s" % (i, k)] = r[k]
        for side in ("feats_prev", "feats_cur"):
            for k, v in r[side].items():
                out["p%d_%s_%s" % (i, side, k)] = v
    path = os.path.join(HERE, "golden_reference_module.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", [len(r["matches"]) for r in res], "matches;",
          "scaled |t| =", [float(np.linalg.norm(r["rel"][:3, 3])) for r in res])


if __name__ == "__main__":
    main()
