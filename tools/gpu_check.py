"""Stage-by-stage parity report of the CUDA path against the oracle on a GPU box (debug aid; tests/ hold the real gates)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200 import synth, _native
from oracle import orb_np as O, pose_np as P, chain_np as N
try:
    from oracle import cv2_chain as C
    HAVE_CV2 = C.available()
except Exception:
    HAVE_CV2 = False
print("cv2 available:", HAVE_CV2, "| torch", torch.__version__, "| gpu", torch.cuda.get_device_name(0))

def main():
    use_tma = os.environ.get("DVO_NO_TMA") is None
    nf = int(os.environ.get("NF", "2000"))
    frames, poses, K = synth.render_sequence(3, device="cuda")
    fh = frames.cpu().numpy()
    ctx = _native.Context(1280, 1024, nfeatures=nf, max_frames=4, use_tma=use_tma)
    print("max_keypoints", ctx.max_keypoints)
    ctx.load_frames(frames, 0)
    ctx.orb(0, 3)
    ctx.sync()
    print("orb ran; launches", ctx.kernel_launches)
    ref = O.orb_detect_and_compute(fh[0], nf)
    for L in range(8):
        pyr = ctx.tap_image(0, L, 0)
        d = int((pyr != ref["pyramid"][L]).sum())
        cand = ctx.tap_candidates(0, L)
        rc = ref["levels"][L]["candidates"]
        eq = cand.shape == rc.shape and np.array_equal(cand, rc)
        blur = ctx.tap_image(0, L, 1)
        bd = int((blur != O.gaussian_blur_7x7(ref["pyramid"][L])).sum())
        print("level %d: pyramid diff %d | candidates %d vs %d equal %s | blur diff %d" % (L, d, len(cand), len(rc), eq, bd))
    f = ctx.features(0)
    print("features:", len(f["pt"]), "vs", len(ref["pt"]))
    if len(f["pt"]) == len(ref["pt"]):
        for k in ("pt", "size", "angle", "response", "octave", "desc"):
            print("   %s equal: %s" % (k, np.array_equal(f[k], ref[k])))
        if not np.array_equal(f["pt"], ref["pt"]):
            same_set = set(map(tuple, f["pt"].tolist())) == set(map(tuple, ref["pt"].tolist()))
            print("   same set of pts:", same_set)
    if HAVE_CV2:
        fc = C.orb_features(fh[1], nf)
        f1 = ctx.features(1)
        print("cv2 frame1: count", len(fc["pt"]), len(f1["pt"]), "all equal:",
              all(np.array_equal(f1[k], fc[k]) for k in ("pt", "size", "angle", "response", "octave", "desc")) if len(fc["pt"]) == len(f1["pt"]) else False)
    # pairs
    ctx.pairs(0, 0, 2, K)
    ps = ctx.poses(0, 2)
    for p in range(2):
        r = (C.frame_pair if HAVE_CV2 else N.frame_pair)(fh[p], fh[p + 1], K, nf)
        a = ctx.pair_arrays(p, ps[p]["n_matches"])
        print("pair %d: status %d matches %d vs %d equal %s" % (p, ps[p]["status"], ps[p]["n_matches"], len(r["matches"]),
              a["matches"].shape == r["matches"].shape and np.array_equal(a["matches"], r["matches"])))
        print("   ransac state", ctx.tap_ransac(p), "iters", ps[p]["ransac_iters"], "best_iter", ps[p]["best_iter"])
        E = ps[p]["E"].reshape(3, 3)
        print("   E diff %.3e | inliers %d vs %d | mask equal %s" % (min(np.abs(E - r["E"]).max(), np.abs(E + r["E"]).max()),
              ps[p]["n_inliers"], int(r["ransac_mask"].sum()), np.array_equal(a["ransac_mask"], r["ransac_mask"])))
        R = ps[p]["R"].reshape(3, 3); t = ps[p]["t"]
        ang = np.degrees(np.arccos(np.clip((np.trace(R @ r["R"].T) - 1) / 2, -1, 1)))
        tang = np.degrees(np.arccos(np.clip(float(t @ r["t"][:, 0]), -1, 1)))
        print("   R err %.5f deg | t err %.5f deg | good %d vs %d | pose mask equal %s | cand %d" % (ang, tang, ps[p]["n_good"], r["good"],
              np.array_equal(a["pose_mask"], r["pose_mask"]), ps[p]["candidate"]))
    # timing
    torch.cuda.synchronize()
    for name, fn in (("orb x3", lambda: ctx.orb(0, 3)), ("pairs x2", lambda: ctx.pairs(0, 0, 2, K))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
        print("%s: %.3f ms" % (name, (time.perf_counter() - t0) * 100))
    # sequence
    seq = ctx.sequence(frames, K)
    print("sequence poses equal pairwise:", all(np.array_equal(seq[i]["R"], ps[i]["R"]) for i in range(2)))
    seqh = ctx.sequence(fh, K)
    print("host sequence equal:", all(np.array_equal(seqh[i]["R"], ps[i]["R"]) for i in range(2)))

main()
