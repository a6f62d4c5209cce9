"""SURVEY 8e correctness test on hardware: the trajectory-extraction driver under torchrun with 2 ranks (frame pairs sharded, NCCL
all-gather of the per-pair records, chain on rank 0) writes stamped_traj_estimate_{absolute,relative,velocity}.txt byte-identical
to the 1-GPU run (reference chain: /root/reference/scripts/visual_odometry_v3.py:367; writers:
trajectory_evaluation_dual_process.py:280-290).  Skipped on boxes with fewer than 2 GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nproc", [2])
def test_sharded_trajectory_files_equal_the_single_gpu_files(tmp_path, nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    import yaml
    from droplet_visual_odometry_b200 import synth
    n, w, h = 45, 640, 480
    frames, _, K = synth.render_sequence(n, width=w, height=h, device="cuda", start_index=120)
    folder = tmp_path / "frames"
    folder.mkdir()
    for i, f in enumerate(frames.cpu().numpy()):
        np.save(str(folder / ("frame_%04d.npy" % i)), f)
    calib = tmp_path / "calib.yaml"
    yaml.safe_dump({"intrinsic_coeffs": [[float(v) for v in K.ravel()]], "distortion_coeffs": [[0.0, 0.0, 0.0, 0.0, 0.0]]}, open(str(calib), "w"))
    script = os.path.join(ROOT, "dropin", "trajectory_extraction.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    out1, outn = tmp_path / "one", tmp_path / "many"
    subprocess.check_call([sys.executable, script, str(folder), str(calib), str(out1), "--batch", "8"], env=env, timeout=600)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
                           "--master-port", str(_free_port()), script, str(folder), str(calib), str(outn), "--batch", "8"], env=env, timeout=900)
    for name in ("stamped_traj_estimate_absolute.txt", "stamped_traj_estimate_relative.txt", "stamped_traj_estimate_velocity.txt"):
        a, b = (out1 / name).read_bytes(), (outn / name).read_bytes()
        assert len(a) > 0 and a == b, name
    assert len((out1 / "stamped_traj_estimate_absolute.txt").read_text().splitlines()) == n
