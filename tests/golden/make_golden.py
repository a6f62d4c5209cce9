"""Generates the golden vectors under tests/golden from the REFERENCE's arithmetic engine, cv2 4.13.0, in the build
container (the reference repository ships no fixtures of its own -- SURVEY.md §4).  Where the reference's own code can
run (``VisualOdometry.compute_current_image_elements`` imports and works once ROS-only modules are stubbed), it is
imported from /root/reference and its output recorded next to the direct cv2 calls, so the fixtures pin the reference's
function, not just our reading of it.  Run:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from droplet_visual_odometry_b200 import synth  # noqa: E402

W, H, NF = 480, 360, 300


def reference_orb(img, calib):
    """Call the reference's own compute_current_image_elements (visual_odometry_v3.py:370-379) if importable."""
    ref_scripts = "/root/reference/scripts"
    if not os.path.isdir(ref_scripts):
        return None
    for name in ("transformations", "tf", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "transformations":
                from droplet_visual_odometry_b200 import transformations_lite as tl
                m.euler_matrix, m.translation_matrix, m.euler_from_matrix = tl.euler_matrix, tl.translation_matrix, tl.euler_from_matrix
            if name == "mpl_toolkits.mplot3d":
                m.axes3d = m.Axes3D = None
            sys.modules[name] = m
    sys.path.insert(0, ref_scripts)
    try:
        import visual_odometry_v3 as ref   # the reference module itself
        vo = ref.VisualOdometry(mode="orb", calibration_file_path=calib)
        vo.feature_detector = cv2.ORB_create(nfeatures=NF)   # reference literal is ORB_create(); nfeatures per config
        kps, desc, _ = vo.compute_current_image_elements(img)
        return np.array([k.pt for k in kps], np.float32), desc
    finally:
        sys.path.remove(ref_scripts)


def main():
    frames, poses, K = synth.render_sequence(3, width=W, height=H, device="cpu")
    a, b = frames[1].numpy(), frames[2].numpy()
    out = {"frame0": a, "frame1": b, "K": K, "nfeatures": np.int32(NF)}
    orb = cv2.ORB_create(nfeatures=NF)
    feats = []
    for i, img in enumerate((a, b)):
        kps, desc = orb.detectAndCompute(img, None)
        f = {"pt": np.array([k.pt for k in kps], np.float32), "size": np.array([k.size for k in kps], np.float32),
             "angle": np.array([k.angle for k in kps], np.float32), "response": np.array([k.response for k in kps], np.float32),
             "octave": np.array([k.octave for k in kps], np.int32), "desc": desc}
        feats.append(f)
        for k, v in f.items():
            out["f%d_%s" % (i, k)] = v
    r = reference_orb(a, "/root/reference/Parameters/camera_calibration.yaml")
    if r is not None:
        assert np.array_equal(r[0], feats[0]["pt"]) and np.array_equal(r[1], feats[0]["desc"]), "reference module disagrees with direct cv2"
        out["reference_module_checked"] = np.int32(1)
        print("reference's compute_current_image_elements == direct cv2 call: OK")
    # stage oracles
    out["level1"] = cv2.resize(a, (400, 300), interpolation=cv2.INTER_LINEAR_EXACT)
    fast = cv2.FastFeatureDetector_create(20, True).detect(a)
    out["fast0"] = np.array([[k.pt[0], k.pt[1], k.response] for k in fast], np.int32)
    k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["blur0"] = cv2.sepFilter2D(a, -1, k32, k32, borderType=cv2.BORDER_REFLECT_101)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    ms = sorted(bf.match(feats[0]["desc"], feats[1]["desc"]), key=lambda m: m.distance)
    m = np.array([(x.queryIdx, x.trainIdx, int(x.distance)) for x in ms], np.int32)
    out["matches"] = m
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(feats[0]["desc"], feats[1]["desc"], k=2)
    out["knn_idx"] = np.array([[p[0].trainIdx, p[1].trainIdx] for p in knn], np.int32)
    out["knn_dist"] = np.array([[p[0].distance, p[1].distance] for p in knn], np.int32)
    p1, p2 = feats[0]["pt"][m[:, 0]], feats[1]["pt"][m[:, 1]]
    E, mask = cv2.findEssentialMat(p1, p2, cameraMatrix=K, method=cv2.RANSAC, prob=0.999, threshold=1.0)
    good, R, t, pmask = cv2.recoverPose(E, p1, p2, cameraMatrix=K)
    out.update(E=E, ransac_mask=mask[:, 0], R=R, t=t, pose_mask=pmask[:, 0], good=np.int32(good))
    # RANSAC-heavy correspondences (config 5 shape, small)
    q1, q2, Kc, Rg, tg, truth = synth.synthetic_correspondences(600, 0.4, 0.3, seed=11)
    E2, mask2 = cv2.findEssentialMat(q1, q2, cameraMatrix=Kc, method=cv2.RANSAC, prob=0.999, threshold=1.0)
    good2, R2, t2, pm2 = cv2.recoverPose(E2, q1, q2, cameraMatrix=Kc)
    out.update(c5_p1=q1, c5_p2=q2, c5_K=Kc, c5_E=E2, c5_mask=mask2[:, 0], c5_R=R2, c5_t=t2, c5_pose_mask=pm2[:, 0], c5_good=np.int32(good2))
    # minimal solver: exactly 5 points -> stacked 3k x 3 solutions
    s1, s2, Ks, _, _, _ = synth.synthetic_correspondences(5, 0.0, 0.3, seed=4)
    Es, _ = cv2.findEssentialMat(s1, s2, cameraMatrix=Ks, method=cv2.RANSAC)
    out.update(five_p1=s1, five_p2=s2, five_K=Ks, five_E=Es)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "golden_480x360.npz"), **out)
    print("wrote golden_480x360.npz:", len(feats[0]["pt"]), "kps,", len(m), "matches, good", good, "| c5 inliers", int(mask2.sum()))


if __name__ == "__main__":
    main()
