#!/usr/bin/env python
"""Timing of the parity-test configurations that are not bench lines (BASELINE.json configs[0], [3], [4]) on one GPU:
CUDA events around the C-ABI calls, inputs resident.  Informational; prints one JSON object."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200 import synth, _native


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
# configs[0]: one 1280x1024 pair, ORB 500
frames, _, K = synth.render_sequence(2, device="cuda")
ctx = _native.Context(1280, 1024, nfeatures=500, max_frames=2)
def c0():
    ctx.load_frames(frames, 0); ctx.orb(0, 2); ctx.pairs(0, 0, 1, K)
out["config0_two_frame_pair_ms"] = round(timed(c0, 20), 3)
ctx.close()
# configs[3]: 2448x2048, ORB 10000, kNN ratio + reverse check
frames, _, K = synth.render_sequence(5, width=2448, height=2048, device="cuda")
ctx = _native.Context(2448, 2048, nfeatures=10000, max_frames=5, matcher=_native.DVO_MATCH_KNN_RATIO)
def c3():
    ctx.load_frames(frames, 0); ctx.orb(0, 5); ctx.pairs(0, 0, 4, K)
out["config3_high_density_ms_per_pair"] = round(timed(c3, 5) / 4, 3)
p = ctx.poses(0, 4)
out["config3_matches_median"] = float(np.median(p["n_matches"]))
ctx.close()
# configs[4]: RANSAC-heavy, 40 % outliers, maxIters 4096
for n in (1000, 5000, 20000, 50000):
    p1, p2, K, R, t, truth = synth.synthetic_correspondences(n, 0.4, 0.3, seed=n)
    ctx = _native.Context(64, 64, nfeatures=n, max_frames=2, ransac_max_iters=4096)
    a, b = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    out["config4_ransac_%d_ms" % n] = round(timed(lambda: ctx.pose_points(a, b, K), 5), 3)
    out["config4_ransac_%d_iters" % n] = int(ctx.poses(0, 1)[0]["ransac_iters"])
    ctx.close()
    ctx = _native.Context(64, 64, nfeatures=n, max_frames=2, ransac_max_iters=4096, ransac_exhaustive=True)
    out["config4_exhaustive4096_%d_ms" % n] = round(timed(lambda: ctx.pose_points(a, b, K), 5), 3)
    ctx.close()
print(json.dumps(out))
