"""CPU: the multi-GPU path of the sequence runner on a world_size-2 (and 3) gloo group: contiguous pair blocks, one
all-gather of the fixed-size per-pair records, identical trajectory files on every rank and equal to the 1-rank result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from droplet_visual_odometry_b200 import sequence as S
    from test_host_logic import _fake_poses
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allp = _fake_poses(n_frames - 1, seed=5)
    calls = []

    def block(first, last):
        calls.append((first, last))
        return allp[first:last]        # a rank only ever produces its own block
    out = S.run_sharded(n_frames, block, world, rank)
    assert np.array_equal(out, allp), "gathered records differ on rank %d" % rank
    s, e = S.shard_pairs(n_frames - 1, world, rank)
    assert calls == ([(s, e)] if e > s else [])
    S.write_trajectory(os.path.join(outdir, "rank%d" % rank), [0.05 * i for i in range(n_frames)], out)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_frames, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    ref = None
    for r in range(world):
        txt = open(os.path.join(str(tmp_path), "rank%d" % r, "stamped_traj_estimate_absolute.txt")).read()
        ref = txt if ref is None else ref
        assert txt == ref
    return ref


def test_world2_equals_world1(tmp_path):
    a = _run(2, 24, tmp_path / "w2")
    b = _run(1, 24, tmp_path / "w1")
    assert a == b and a.count("\n") == 24


def test_world3_ragged_blocks(tmp_path):
    assert _run(3, 8, tmp_path / "w3").count("\n") == 8       # 7 pairs over 3 ranks: 3 + 3 + 1
