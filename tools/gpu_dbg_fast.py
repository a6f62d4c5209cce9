import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from droplet_visual_odometry_b200 import synth, _native
from oracle import orb_np as O
frames, poses, K = synth.render_sequence(1, device="cuda")
fh = frames.cpu().numpy()
ctx = _native.Context(1280, 1024, nfeatures=2000, max_frames=2, use_tma=os.environ.get("DVO_NO_TMA") is None)
ctx.load_frames(frames, 0); ctx.orb(0, 1); ctx.sync()
pyr = O.build_pyramid(fh[0])
for L in (0, 4):
    lev = pyr[L]
    sc = O.fast_score_map(lev); keep = O.fast_nms(sc)
    h, w = lev.shape
    m = np.zeros_like(keep); m[31:h-31, 31:w-31] = True
    ref = np.where(keep & m, sc, 0).astype(np.uint8)
    got = ctx.tap_image(0, L, 2)
    d = got != ref
    print("level", L, "map diffs", int(d.sum()), "ref nz", int((ref>0).sum()), "got nz", int((got>0).sum()))
    ys, xs = np.nonzero(d)
    for y, x in list(zip(ys, xs))[:12]:
        print("   (x=%d,y=%d) ref %d got %d  raw score ref %d" % (x, y, ref[y, x], got[y, x], sc[y, x]))
    if len(ys):
        print("   diff x range", xs.min(), xs.max(), "y range", ys.min(), ys.max(), "x%128 hist", np.bincount(xs % 128, minlength=128)[:8], "y%32 hist", np.bincount(ys % 32, minlength=32))
    cand = ctx.tap_candidates(0, L)
    print("   candCount", len(cand), "map nz", int((got>0).sum()))
