"""BASELINE configs[1] at its stated size on the B200: a synthetic 1000-frame 1280x1024 sequence, ORB 2000 features, EVERY one of
the 999 consecutive pairs compared with the reference's cv2 chain (oracle/cv2_chain.py, run in a process pool with each
frame's features computed once per worker chunk).

Bars: matches (queryIdx, trainIdx, distance) array_equal for every pair; rotation <= 0.1 deg, translation direction <= 0.5 deg,
RANSAC-mask IoU >= 0.95 and recoverPose-mask IoU >= 0.95 for every pair on which both RANSAC loops end on the same model.  Winner
selection is discrete and cv2 returns the unrefined minimal model: where the winners differ the pair can leave the tolerance,
and the count of such pairs is bounded by what the reference does to ITSELF when one intrinsic moves by one ulp (measured in the
same run).  Printed: pairs in tolerance, pairs with IDENTICAL masks, pairs with E equal to 1e-4, and the same for cv2 vs cv2.  Reference loop being replaced:
/root/reference/scripts/trajectory_evaluation_dual_process.py:170-252 (one visual_odometry_calculations call per pair).
"""
import os
import tempfile

import numpy as np
import pytest

from conftest import rot_err_deg, dir_err_deg, mask_iou

pytestmark = pytest.mark.gpu

N_FRAMES = int(os.environ.get("DVO_FULLSEQ_FRAMES", "1000"))
NF, W, H = 2000, 1280, 1024


def _worker(args):
    path, kpath, lo, hi = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import cv2_chain
    F = np.load(path, mmap_mode="r")
    K = np.load(kpath)
    out = []
    prev = cv2_chain.orb_features(np.ascontiguousarray(F[lo]), NF)
    for i in range(lo, hi):
        cur = cv2_chain.orb_features(np.ascontiguousarray(F[i + 1]), NF)
        r = cv2_chain.frame_pair(None, None, K, NF, feats_prev=prev, feats_cur=cur)
        o = {k: r[k] for k in ("status", "matches", "E", "R", "t", "ransac_mask", "pose_mask", "good")}
        # the reference against ITSELF with fx moved by one ulp: how much of its answer is numerical chance
        K1 = K.copy()
        K1[0, 0] = np.nextafter(K1[0, 0], 1e9)
        s = cv2_chain.pose_from_points(r["p_prev"], r["p_cur"], K1)
        o["self"] = {k: s[k] for k in ("E", "R", "t", "ransac_mask", "pose_mask")}
        out.append(o)
        prev = cur
    return lo, out


def _reference_pairs(frames_u8, K):
    import multiprocessing as mp
    workers = max(1, os.cpu_count() or 1)
    n_pairs = len(frames_u8) - 1
    with tempfile.TemporaryDirectory(prefix="dvo_fullseq_") as d:
        path, kpath = os.path.join(d, "frames.npy"), os.path.join(d, "K.npy")
        np.save(path, frames_u8)
        np.save(kpath, np.asarray(K, dtype=np.float64))
        chunk = max(4, -(-n_pairs // (workers * 3)))
        jobs = [(path, kpath, lo, min(lo + chunk, n_pairs)) for lo in range(0, n_pairs, chunk)]
        with mp.get_context("spawn").Pool(workers) as pool:
            parts = pool.map(_worker, jobs, chunksize=1)
    ref = [None] * n_pairs
    for lo, out in parts:
        ref[lo:lo + len(out)] = out
    return ref


def test_every_pair_of_the_1000_frame_sequence_against_cv2():
    import time
    import torch
    from droplet_visual_odometry_b200 import synth, _native
    from oracle import cv2_chain
    if not cv2_chain.available():
        pytest.skip("cv2 not importable")
    frames, _, K = synth.render_sequence(N_FRAMES, W, H, device="cuda")
    n_pairs = N_FRAMES - 1
    t0 = time.perf_counter()
    ref = _reference_pairs(frames.cpu().numpy(), K)
    t_ref = time.perf_counter() - t0

    B = 148
    ctx = _native.Context(W, H, nfeatures=NF, max_frames=B + 1)
    t0 = time.perf_counter()
    seq = ctx.sequence(frames, K)               # the product's sequence runner (pipelined batches, carried frame)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    assert len(seq) == n_pairs and (seq["status"] == 0).all() and (seq["frame_flags"] == 0).all()

    # per-pair arrays (matches, masks) through the stage entry points, batch by batch; records must equal the runner's
    stats = {"mask_identical": 0, "pose_mask_identical": 0, "E_1e-4": 0, "worst_rot": 0.0, "worst_tdir": 0.0, "worst_iou": 1.0,
             "worst_pose_iou": 1.0, "self_in_tolerance": 0, "self_mask_identical": 0}
    bad, bad_same_model = [], []
    for lo in range(0, n_pairs, B):
        hi = min(lo + B, n_pairs)
        ctx.load_frames(frames[lo:hi + 1], 0)
        ctx.orb(0, hi - lo + 1)
        ctx.pairs(0, 0, hi - lo, K)
        ps = ctx.poses(0, hi - lo)
        for j in range(hi - lo):
            i = lo + j
            p, r = ps[j], ref[i]
            assert np.array_equal(p["R"], seq[i]["R"]) and np.array_equal(p["t"], seq[i]["t"]) and p["n_inliers"] == seq[i]["n_inliers"], i
            arr = ctx.pair_arrays(j, p["n_matches"])
            assert r["status"] == 0, i
            assert np.array_equal(arr["matches"], r["matches"]), "pair %d: match list differs from cv2" % i
            E = p["E"].reshape(3, 3)
            e_ok = min(np.abs(E - r["E"]).max(), np.abs(E + r["E"]).max()) < 1e-4
            re, de = rot_err_deg(p["R"], r["R"]), dir_err_deg(p["t"], r["t"])
            iou, piou = mask_iou(arr["ransac_mask"], r["ransac_mask"]), mask_iou(arr["pose_mask"], r["pose_mask"])
            stats["E_1e-4"] += int(e_ok)
            stats["mask_identical"] += int(np.array_equal(arr["ransac_mask"] > 0, r["ransac_mask"] > 0))
            stats["pose_mask_identical"] += int(np.array_equal(arr["pose_mask"] > 0, r["pose_mask"] > 0))
            stats["worst_rot"], stats["worst_tdir"] = max(stats["worst_rot"], re), max(stats["worst_tdir"], de)
            stats["worst_iou"], stats["worst_pose_iou"] = min(stats["worst_iou"], iou), min(stats["worst_pose_iou"], piou)
            if not (re <= 0.1 and de <= 0.5 and iou >= 0.95 and piou >= 0.95):
                bad.append((i, round(re, 4), round(de, 4), round(iou, 4), round(piou, 4), bool(e_ok)))
                if e_ok:
                    bad_same_model.append(i)
            sf = r["self"]
            stats["self_in_tolerance"] += int(rot_err_deg(sf["R"], r["R"]) <= 0.1 and dir_err_deg(sf["t"], r["t"]) <= 0.5 and
                                              mask_iou(sf["ransac_mask"], r["ransac_mask"]) >= 0.95 and
                                              mask_iou(sf["pose_mask"], r["pose_mask"]) >= 0.95)
            stats["self_mask_identical"] += int(np.array_equal(sf["ransac_mask"] > 0, r["ransac_mask"] > 0))
    ctx.close()
    print("\nconfigs[1] full size: %d pairs; cv2 chain %.1f s on %d cores; GPU sequence runner %.2f s (%.0f pairs/s incl. first-call set-up)"
          % (n_pairs, t_ref, os.cpu_count() or 1, t_gpu, n_pairs / t_gpu))
    print("in tolerance: %d/%d; RANSAC mask identical: %d; recoverPose mask identical: %d; E within 1e-4: %d; worst rot %.2e deg, "
          "t-dir %.2e deg, mask IoU %.4f, pose-mask IoU %.4f" % (n_pairs - len(bad), n_pairs, stats["mask_identical"], stats["pose_mask_identical"],
                                                               stats["E_1e-4"], stats["worst_rot"], stats["worst_tdir"], stats["worst_iou"], stats["worst_pose_iou"]))
    print("cv2 against itself with fx moved by ONE ulp: in tolerance %d/%d, RANSAC mask identical %d"
          % (stats["self_in_tolerance"], n_pairs, stats["self_mask_identical"]))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        json.dump({"pairs": n_pairs, "in_tolerance": n_pairs - len(bad), "bad": bad, **stats, "cv2_seconds": t_ref, "gpu_seconds": t_gpu,
                   "cores": os.cpu_count()}, open(os.path.join(out_dir, "full_sequence_parity.json"), "w"))
    # Hard bars.  (1) The integer stage is exact on every pair (asserted above).  (2) Whenever the two RANSAC loops end on the same
    # model, everything downstream is inside the north-star tolerance.  (3) cv2 returns the UNREFINED best minimal model, and on this
    # piecewise-planar scene a few per cent of the minimal samples are ill-conditioned (coplanar / clustered points): the solver's
    # answer for them depends on the last bit of its input, cv2's own answer included -- the line above measures that.  The pairs where
    # the winners differ must therefore stay within twice the reference's own one-ulp disagreement (plus 1 % of the pairs).
    assert not bad_same_model, "same E as cv2 but pose/masks outside the tolerance: pairs %s" % bad_same_model[:20]
    self_bad = n_pairs - stats["self_in_tolerance"]
    assert len(bad) <= 2 * self_bad + n_pairs // 100, \
        "%d pairs outside the tolerance, cv2 disagrees with itself on %d (index, rot, tdir, iou, pose iou, E ok): %s" % (len(bad), self_bad, bad[:20])
