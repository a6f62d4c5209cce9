#!/usr/bin/env python
"""Headline benchmark: frame-pairs/sec of the per-frame-pair VO hot path (ORB detect/describe -> BF Hamming match ->
E-RANSAC -> recoverPose) on synthetic textured 1280x1024 mono frames, ORB 2000 features (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1: one rank per GPU)
  python bench.py --impl reference ...                     the reference's cv2 CPU path on the box's host cores

A step = one batch of --batch consecutive new frames of the sequence (= --batch frame pairs; the last frame of the
previous batch is carried, its features are not recomputed).  Timed region: K steps bracketed by barrier + sync, CUDA
events on the launching stream, max over ranks; frames are resident in HBM (value) or in pinned host memory (e2e).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (ORB+match+E-RANSAC+pose)"
UNIT = "frame-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=148, help="new frames (= pairs) per step (default: one k_ransac CTA per SM)")
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--nfeatures", type=int, default=2000)
    ap.add_argument("--ref-pairs-per-step", type=int, default=0, help="reference arm: pairs per step (0 = 2 x workers, min 8)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--matcher-engine", default="tensor", choices=["tensor", "popc"],
                    help="cross-check matcher: int8 tcgen05 GEMM (default) or the XOR+POPC integer-pipe kernel (BASELINE north_star item 4 as written)")
    ap.add_argument("--regions", type=int, default=5,
                    help="timed regions of --steps steps each, back to back, every one bracketed by barrier + sync; the line reports the MEDIAN region "
                         "(all of them are listed under config.region_ms)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE configs[0], [3], [4] timings (N=1 only)")
    ap.add_argument("--ingest", default="none", choices=["none", "grey", "bgr"],
                    help="widened path (SURVEY 8f row 2): frames arrive distorted (grey or BGR) and cv.cvtColor + cv.undistort "
                         "run inside the timed region on both arms; default: the BASELINE workload (undistorted mono frames)")
    return ap.parse_args()


def workload_name(a):
    base = "synthetic sequence %dx%d mono, ORB %d feats, 8 levels, crossCheck BF-Hamming, findEssentialMat RANSAC(0.999, 1px, 1000 it), recoverPose; consecutive-pair VO" % (
        a.width, a.height, a.nfeatures)
    if a.ingest != "none":
        base += "; ingest: %s frames, cvtColor+undistort (reference calibration distortion) inside the timed region" % a.ingest
    return base


# distortion of /root/reference/Parameters/camera_calibration.yaml:25, used when --ingest is on
INGEST_DIST = np.array([-0.296079, 0.099771, 0.000222, 0.000109, 0.0])


def ingest_new_K(K):
    """newCameraMatrix for the ingest runs: the focal lengths scaled like getOptimalNewCameraMatrix(alpha=1) does for this
    distortion (0.81), principal point kept -- any matrix is valid input for cv.undistort / k_ingest."""
    newK = np.array(K, dtype=np.float64).copy()
    newK[0, 0] *= 0.81
    newK[1, 1] *= 0.81
    return newK


# ================================================================================================ CPU reference (cv2)
def _ref_worker_init(path, kpath, nf, ingest="none"):
    global _F, _K, _NF, _ING
    import cv2
    cv2.setNumThreads(1)
    _F = np.load(path, mmap_mode="r")
    _K = np.load(kpath)
    _NF = nf
    _ING = ingest


def _ref_ingest(img):
    import cv2
    if _ING == "none":
        return np.ascontiguousarray(img)
    if _ING == "bgr":        # the reference decodes to BGR, converts, then undistorts (visual_odometry_v3.py:127-133)
        img = cv2.cvtColor(np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2)), cv2.COLOR_BGR2GRAY)
    return cv2.undistort(np.ascontiguousarray(img), _K, INGEST_DIST, None, ingest_new_K(_K))


def _ref_worker_pair(i):
    from oracle import cv2_chain
    K = _K if _ING == "none" else ingest_new_K(_K)
    r = cv2_chain.frame_pair(_ref_ingest(_F[i]), _ref_ingest(_F[i + 1]), K, _NF)     # both frames' ingest + ORB per pair,
    return int(len(r["matches"]))                                                   # as visual_odometry_v3.py:387-392


class CpuReference:
    """The reference's per-pair chain through cv2 (oracle/cv2_chain.py), one worker process per host core over
    independent pairs, cv2 internal threading off in each worker.  Falls back to the numpy port only if cv2 is absent."""

    def __init__(self, frames_u8: np.ndarray, K: np.ndarray, nfeatures: int, workers: int | None = None, ingest: str = "none"):
        import multiprocessing as mp
        from oracle import cv2_chain
        self.kind = "reference" if cv2_chain.available() else "port"
        self.n_pairs = len(frames_u8) - 1
        self.workers = workers or (os.cpu_count() or 1)
        if self.kind == "port":
            self.workers = 1
            self.frames, self.K, self.nf = frames_u8, K, nfeatures
            return
        self.tmp = tempfile.mkdtemp(prefix="dvo_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.path, self.kpath = os.path.join(self.tmp, "frames.npy"), os.path.join(self.tmp, "K.npy")
        np.save(self.path, frames_u8)
        np.save(self.kpath, np.asarray(K, dtype=np.float64))
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_ref_worker_init, initargs=(self.path, self.kpath, nfeatures, ingest))
        self.pool.map(_ref_worker_pair, [0] * self.workers)     # spin every worker up (imports, cv2 init)

    def run_pairs(self, idx):
        if self.kind == "port":
            from oracle import chain_np
            for i in idx:
                chain_np.frame_pair(self.frames[i], self.frames[i + 1], self.K, self.nf)
            return
        self.pool.map(_ref_worker_pair, list(idx), chunksize=1)

    def close(self):
        if self.kind == "reference":
            self.pool.close()
            self.pool.join()
            for f in (self.path, self.kpath):
                os.remove(f)
            os.rmdir(self.tmp)


def render_host_frames(n, a, start_index=0):
    import torch
    from droplet_visual_odometry_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    frames, _, K = synth.render_sequence(n, a.width, a.height, device=dev, start_index=start_index)
    return frames.cpu().numpy(), K


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    pps = a.ref_pairs_per_step or max(8, 2 * workers)
    pps = min(pps, 256)
    frames, K = render_host_frames(pps + 1, a)
    ref = CpuReference(frames, K, a.nfeatures, workers, a.ingest)
    for _ in range(max(a.warmup, 0)):
        ref.run_pairs(range(min(pps, workers)))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ref.run_pairs(range(pps))
    dt = time.perf_counter() - t0
    ref.close()
    value = a.steps * pps / dt
    sample = "%d steps x %d independent frame pairs (both frames' ORB recomputed per pair), %d worker processes, cv2 threads=1 each" % (
        a.steps, pps, ref.workers)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
           "data": "synthetic", "config": {"workload": workload_name(a), "pairs_per_step": pps, "cpu": cpu_model()},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.workers, "kind": ref.kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ================================================================================================ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (every 2 ms -- the timed region of
    a default run is under 100 ms, too short for nvidia-smi's start-up), nvidia-smi -lms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index):
        self.p, self.f, self.thread = None, None, None
        self.sm, self.mx, self.reasons = [], [], set()
        try:
            import pynvml, threading
            pynvml.nvmlInit()
            self.h = None
            try:                                           # CUDA_VISIBLE_DEVICES may renumber: look the device up by UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except Exception:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self._poll()                                   # fail here, not in the thread, if a query is unsupported
            self.halt = threading.Event()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _poll(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        for name, bit in self.BITS:
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        while not self.halt.is_set():
            try:
                self._poll()
            except Exception:
                break
            self.halt.wait(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.halt.set()
            self.thread.join(timeout=2)
            sm = self.sm[1:] or self.sm                    # the first sample predates the timed region
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(self.mx)), reasons=sorted(self.reasons), samples=len(sm),
                       source="nvml")
            return out
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


# ================================================================================================ B200 arm
def level_pixel_counts(ctx):
    return [ctx.level_size(L)[0] * ctx.level_size(L)[1] for L in range(ctx.nlevels)]


def _bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this process (and therefore the pinned host buffers it allocates, first touch) on the CPU
    socket the GPU hangs off, so that eight ranks' H2D streams do not all cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (max(os.sched_getaffinity(0)) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def other_configs(dev, with_cpu=True):
    """BASELINE.json configs[0], [3], [4] -- parity-test cases, timed here so that the driver's line carries them: CUDA events around
    the C-ABI calls (inputs resident) and, beside each, the cv2 chain in this process with cv2's default threading."""
    import torch
    from droplet_visual_odometry_b200 import synth, _native
    from droplet_visual_odometry_b200.visual_odometry_v3 import VisualOdometry
    from oracle import cv2_chain
    have_cv = with_cpu and cv2_chain.available()
    if have_cv:
        import cv2
        cv2.setNumThreads(-1)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def cpu_ms(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {}
    # ---- configs[0]: one 1280x1024 pair, ORB 500 (the reference's literal): stage calls on resident frames, and the per-call
    # latency of the drop-in method with host images in and the 4x4 out
    frames, _, K = synth.render_sequence(2, device=dev)
    fh = frames.cpu().numpy()
    ctx = _native.Context(1280, 1024, nfeatures=500, max_frames=2, device=dev.index)

    def c0():
        ctx.load_frames(frames, 0)
        ctx.orb(0, 2)
        ctx.pairs(0, 0, 1, K)
    c = {"workload": "one 1280x1024 pair, ORB 500, crossCheck + RANSAC + recoverPose", "gpu_ms_resident": round(timed(c0, 20), 3)}
    ctx.close()
    vo = VisualOdometry(mode="orb", camera_matrix=K, nfeatures=500, device=dev.index)
    T = np.eye(4)
    vo.visual_odometry_calculations(fh[0], fh[1], T)
    t0 = time.perf_counter()
    for _ in range(20):
        vo.visual_odometry_calculations(fh[0], fh[1], T)
    c["gpu_ms_per_call_host_images"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
    if have_cv:
        c["cv2_ms"] = round(cpu_ms(lambda: cv2_chain.frame_pair(fh[0], fh[1], K, 500), 5), 2)
        c["cv2_threads"] = int(cv2.getNumThreads())
    out["config0_two_frame"] = c
    del vo
    # ---- configs[3]: 2448x2048, ORB 10000, kNN k=2 ratio + reverse check
    frames, _, K = synth.render_sequence(5, width=2448, height=2048, device=dev)
    ctx = _native.Context(2448, 2048, nfeatures=10000, max_frames=5, matcher=_native.DVO_MATCH_KNN_RATIO, device=dev.index)

    def c3():
        ctx.load_frames(frames, 0)
        ctx.orb(0, 5)
        ctx.pairs(0, 0, 4, K)
    c = {"workload": "2448x2048, ORB 10000, knnMatch(k=2) + 0.75 ratio + reverse check, RANSAC + recoverPose; 5 frames = 4 pairs",
         "gpu_ms_per_pair": round(timed(c3, 5) / 4, 3)}
    c["matches_median"] = float(np.median(ctx.poses(0, 4)["n_matches"]))
    ctx.close()
    if have_cv:
        fh = frames[:2].cpu().numpy()
        c["cv2_ms_per_pair"] = round(cpu_ms(lambda: cv2_chain.frame_pair(fh[0], fh[1], K, 10000, matcher="knn"), 1), 1)
        c["cv2_threads"] = int(cv2.getNumThreads())
    out["config3_high_density"] = c
    del frames
    # ---- configs[4]: RANSAC-heavy, 40 % outliers, maxIters 4096, points only
    sweep = {}
    for n in (1000, 5000, 20000, 50000):
        p1, p2, K, R, t, truth = synth.synthetic_correspondences(n, 0.4, 0.3, seed=n)
        a_, b_ = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
        ctx = _native.Context(64, 64, nfeatures=n, max_frames=2, ransac_max_iters=4096, device=dev.index)
        e = {"gpu_ms_cv2_stop_rule": round(timed(lambda: ctx.pose_points(a_, b_, K), 5), 3), "iterations": int(ctx.poses(0, 1)[0]["ransac_iters"])}
        ctx.close()
        ctx = _native.Context(64, 64, nfeatures=n, max_frames=2, ransac_max_iters=4096, ransac_exhaustive=True, device=dev.index)
        e["gpu_ms_all_4096_scored"] = round(timed(lambda: ctx.pose_points(a_, b_, K), 5), 3)
        ctx.close()
        if have_cv:
            e["cv2_ms"] = round(cpu_ms(lambda: cv2_chain.pose_from_points(p1, p2, K, max_iters=4096), 2), 2)
        sweep[str(n)] = e
    out["config4_ransac_heavy"] = {"workload": "synthetic correspondences, 40 % outliers, findEssentialMat(maxIters=4096) + recoverPose", "by_match_count": sweep}
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    from droplet_visual_odometry_b200 import synth, _native
    from droplet_visual_odometry_b200._native import POSE_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = _bind_to_gpu_numa_node(local) if world > 1 else None     # pinned host frames land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K_steps, W_steps = a.batch, max(a.steps, 1), max(a.warmup, 3)      # timing rule: at least 3 untimed warm-up steps
    n_frames = (K_steps + W_steps) * B + 1
    # ONE synthetic sequence of world * (n_frames - 1) + 1 frames, sharded by frame pair: rank r owns the contiguous block of pairs
    # [r * P, (r + 1) * P), P = n_frames - 1, i.e. frames r * P .. (r + 1) * P -- its last frame is the next rank's first (one-frame
    # halo, recomputed on both, as sequence.shard_pairs / run_sharded do for a trajectory extraction)
    frames, _, Kmat = synth.render_sequence(n_frames, a.width, a.height, device=dev, start_index=rank * (n_frames - 1))
    ctx = _native.Context(a.width, a.height, nfeatures=a.nfeatures, max_frames=B + 1, device=local,
                          nn_engine=0 if a.matcher_engine == "tensor" else 1)
    Kpose = Kmat
    if a.ingest != "none":
        Kpose = ingest_new_K(Kmat)
        ctx.set_undistort(Kmat, INGEST_DIST, Kpose, channels=3 if a.ingest == "bgr" else 1)
        if a.ingest == "bgr":
            frames = frames[:, :, :, None].expand(-1, -1, -1, 3).contiguous()
    rec = POSE_DTYPE.itemsize
    poses_dev = torch.zeros((K_steps + W_steps) * B * rec, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world * K_steps * B * rec, dtype=torch.uint8, device=dev) if world > 1 else None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_pass(first_step, nsteps, fresh):
        """nsteps batches from HBM-resident frames; returns #pairs"""
        pairs = 0
        for s in range(first_step, first_step + nsteps):
            lo = s * B + (0 if (fresh and s == first_step) else 1)
            hi = (s + 1) * B + 1
            out = poses_dev[pairs * rec + first_step * B * rec:]
            pairs += ctx.sequence_step(frames[lo:hi], Kpose, out, first=(fresh and s == first_step))
        return pairs

    # ---- warm-up (also primes the carry slot so every timed step is B pairs); the path's one collective is issued here too, so
    # that NCCL's lazy channel set-up for all-gather is not inside a timed region
    send = poses_dev[W_steps * B * rec:(W_steps + K_steps) * B * rec]
    device_pass(0, W_steps, fresh=True)
    ctx.flush()
    if world > 1:
        for _ in range(2):
            dist.all_gather_into_tensor(gathered, send)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    R_regions = max(a.regions, 1)
    region_ms, gather_ms = [], []
    npairs = 0
    for r in range(R_regions):
        # every region times the same K steps (3.9 GB of frames, far beyond L2).  The step before it is replayed untimed so
        # that the carried frame is the true predecessor of the region's first frame.
        device_pass(W_steps - 1, 1, fresh=False)
        ctx.flush()
        barrier()
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        launches0 = ctx.kernel_launches
        ev0.record()
        npairs = device_pass(W_steps, K_steps, fresh=False)
        ctx.flush()     # join torch's stream with the pipelined runner's internal streams before the closing event
        ev1.record()
        launches = ctx.kernel_launches - launches0      # kernels of ONE timed region
        if world > 1:   # the path's one exchange step: all-gather of the per-pair (R, t, status) records, on the compute stream
            dist.all_gather_into_tensor(gathered, send)
        ev2.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev2), ev1.elapsed_time(ev2)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_ms.append(float(t[0].item()))
        gather_ms.append(float(t[1].item()))
    clocks = sampler.stop() if sampler else None
    assert npairs == K_steps * B
    ms = float(np.median(region_ms))
    value = world * npairs / (ms / 1e3)
    host_poses = poses_dev[W_steps * B * rec:(W_steps + K_steps) * B * rec].cpu().numpy().view(POSE_DTYPE)
    chain_ms, chain_end = None, None
    if rank == 0:      # what the all-gather is for: chain every rank's (R, t) records into one trajectory (visual_odometry_v3.py:367)
        from droplet_visual_odometry_b200 import sequence as S
        allrec = gathered.cpu().numpy().view(POSE_DTYPE) if world > 1 else host_poses
        t0 = time.perf_counter()
        traj = S.chain(S.poses_to_relatives(allrec))
        chain_ms = (time.perf_counter() - t0) * 1e3
        chain_end = [round(float(v), 4) for v in traj[-1][:3, 3]]
        assert len(traj) == world * K_steps * B + 1
    ok_frac = float(np.mean(host_poses["status"] == 0))
    pair_stats = {"ransac_iters_median": float(np.median(host_poses["ransac_iters"])), "ransac_iters_max": int(host_poses["ransac_iters"].max()),
                  "matches_median": float(np.median(host_poses["n_matches"])),
                  "inlier_ratio_median": float(np.median(host_poses["n_inliers"] / np.maximum(host_poses["n_matches"], 1)))}

    # ---- end to end through the public API: pinned host frames in, pose records out, every step
    # (at most 20 steps' worth of distinct frames, copied device -> pinned host directly: 8 ranks share one host's RAM)
    E_steps = min(K_steps, 20)
    src = frames[W_steps * B:(W_steps + E_steps) * B + 1]
    e2e_frames = torch.empty(src.shape, dtype=torch.uint8, pin_memory=True)
    e2e_frames.copy_(src)
    del src
    poses_host_t = torch.empty(E_steps * B * rec, dtype=torch.uint8).pin_memory()
    poses_host = poses_host_t.numpy().view(POSE_DTYPE)

    def host_pass(nsteps, fresh_first=True):
        pairs = 0
        for s in range(nsteps):
            lo = s * B + (0 if s == 0 else 1)
            pairs += ctx.sequence_step(e2e_frames[lo:(s + 1) * B + 1], Kpose, poses_host[pairs:], first=(s == 0))
        ctx.sync()
        return pairs
    host_pass(min(2, E_steps))
    barrier()
    t0 = time.perf_counter()
    e2e_pairs = host_pass(E_steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * e2e_pairs / e2e_s

    # ---- what the platform allows for that: pinned host -> device copies of the same frames, nothing else, all ranks at once
    h2d_dst = torch.empty_like(e2e_frames, device=dev)
    h2d_dst.copy_(e2e_frames, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h2d_dst.copy_(e2e_frames, non_blocking=True)
    torch.cuda.synchronize(dev)
    h2d_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([h2d_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_s = float(t.item())
    h2d_gbs_per_rank = 3 * e2e_frames.numel() / h2d_s / 1e9
    h2d_pairs_ceiling = world * h2d_gbs_per_rank * 1e9 / (a.width * a.height * (3 if a.ingest == "bgr" else 1))
    del h2d_dst

    # ---- roofline pass: the same K steps again with per-kernel CUDA events on the launching stream
    roofline, stages = None, None
    if rank == 0:
        ctx.profile(True)
        ctx.profile_collect()
        device_pass(W_steps, K_steps, fresh=False)
        prof = ctx.profile_collect()
        ctx.profile(False)
        px = level_pixel_counts(ctx)
        total_px = sum(px)
        pyr_bytes = sum(px[L - 1] + px[L] for L in range(1, len(px)))
        alg_bytes_per_frame = {"k_pyr_down": pyr_bytes, "k_fast_nms": total_px, "k_blur": 2 * total_px,
                               "k_load_or_ingest": px[0] * ((3 if a.ingest == "bgr" else 1) + 1)}
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        stages = {}
        nkp = a.nfeatures
        pipe = _native.measure_peaks(local)       # FP32 / FP64 (fused and unfused) / POPC / int8-tensor rates of THIS device, now
        hp = host_poses
        n_prev, n_cur = hp["n_prev"].astype(np.float64), hp["n_cur"].astype(np.float64)
        nm, it = hp["n_matches"].astype(np.float64), hp["ransac_iters"].astype(np.float64)
        # algorithmic work per timed pass (K_steps batches), SURVEY 8d: Sampson scoring 33 flop per (model, match), 4.4 models per
        # hypothesis, ~3e4 flop per 5-point solve; recoverPose: 2 DLT triangulations per match (the other two follow by
        # symmetry), ~3000 flop each (4x4 one-sided Jacobi, ~6 sweeps) + the cheirality tests
        ransac_flop = float(np.sum(it * (4.4 * nm * 33.0 + 3.0e4)))
        cheir_flop = float(np.sum(nm * 2 * 3000.0))
        spec_iters = float(np.sum(np.ceil(it / 16.0) * 16.0))      # k_ransac solves and scores whole chunks of 16 hypotheses
        for name, (tms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            if cnt == 0:
                continue
            st = {"ms_total": round(tms, 3), "share": round(tms / tot_ms, 4), "launch_groups": cnt}
            if name in alg_bytes_per_frame:
                # one launch group covers B frames (k_pyr_down: 7 launches per batch, summed)
                groups = cnt / (7 if name == "k_pyr_down" else 1)
                bytes_total = alg_bytes_per_frame[name] * B * groups
                st["achieved_GBps"] = round(bytes_total / (tms * 1e-3) / 1e9, 2)
                st["frac_of_hbm_peak"] = round(st["achieved_GBps"] / peak, 4)
            if name == "k_nn" and a.matcher_engine == "tensor":
                # cross-check matcher = int8 GEMM on the tensor cores (k_expand_desc + k_nn_tensor).  Algorithmic work: the
                # n_prev x n_cur x 256 distance matrix ONCE, 2 ops per multiply-add (the kernel computes it twice -- the reverse
                # direction is the transposed product -- which is its own inefficiency, not algorithmic work).  Peak: the
                # tcgen05.mma.kind::i8 issue rate measured by libdvo's microbenchmark on this device.
                ops = 2.0 * float(np.sum(n_prev * n_cur)) * 256
                st["int8_tops"] = round(ops / (tms * 1e-3) / 1e12, 1)
                st["int8_peak_tops_measured"] = round(pipe["int8_tensor_ops"] / 1e12, 1)
                st["frac_of_int8_tensor_peak"] = round(ops / (tms * 1e-3) / max(pipe["int8_tensor_ops"], 1.0), 4)
                st["executed_over_algorithmic"] = 2.0
            if name == "k_nn" and a.matcher_engine == "popc":
                # every distance is computed once (row and column minima from the same popcounts): n_prev*n_cur*8 POPC32 per pair
                # against the measured POPC rate (the kernel executes 6 per distance: two carry-save adders replace two)
                pc = float(np.sum(n_prev * n_cur)) * 8
                st["gpopc_per_s"] = round(pc / (tms * 1e-3) / 1e9, 1)
                st["popc_peak_measured_g"] = round(pipe["popc_per_s"] / 1e9, 1)
                st["frac_of_popc_peak"] = round(pc / (tms * 1e-3) / max(pipe["popc_per_s"], 1.0), 4)
            if name == "k_ransac":
                st["algorithmic_gflop"] = round(ransac_flop / 1e9, 3)
                st["fp64_tflops"] = round(ransac_flop / (tms * 1e-3) / 1e12, 3)
                st["fp64_mul_add_peak_tflops_measured"] = round(pipe["fp64_mul_add_flops"] / 1e12, 2)
                st["frac_of_fp64_peak"] = round(ransac_flop / (tms * 1e-3) / max(pipe["fp64_mul_add_flops"], 1.0), 4)
                st["iterations_needed"] = int(np.sum(it))
                st["iterations_speculated"] = int(spec_iters)
                st["iterations_wasted_frac"] = round(1.0 - float(np.sum(it)) / max(spec_iters, 1.0), 4)
            if name == "k_cheirality":
                st["algorithmic_gflop"] = round(cheir_flop / 1e9, 3)
                st["fp64_tflops"] = round(cheir_flop / (tms * 1e-3) / 1e12, 3)
                st["frac_of_fp64_peak"] = round(cheir_flop / (tms * 1e-3) / max(pipe["fp64_mul_add_flops"], 1.0), 4)
            if name in ("k_angle_pack", "k_brief"):
                # gather stages: per keypoint the un-blurred 31x31 disc (709 B, k_angle_pack) or the blurred 37x48 steering window
                # (1776 B, k_brief) plus the 32-B record written; the windows overlap and live in L2, so this is an L2-gather figure
                per_kp = (709 + 32) if name == "k_angle_pack" else (37 * 48 + 32)
                gbytes = float(np.sum(n_cur)) * per_kp
                st["achieved_GBps"] = round(gbytes / (tms * 1e-3) / 1e9, 2)
                st["frac_of_hbm_peak"] = round(st["achieved_GBps"] / peak, 4)
                st["note"] = "L2 gather of overlapping windows; latency-bound"
            if name == "k_select":
                # two retainBest passes over ~1 survivor per 130 px (3 comparisons per element for nth_element + partition) and the
                # Harris response of 2 x quota survivors per level (7x7 block of Sobel products: ~1200 integer ops each)
                cand = total_px / 130.0
                iops = (3.0 * cand + 3.0 * 2 * nkp + 2 * nkp * 1200.0) * B * cnt
                st["algorithmic_gop"] = round(iops / 1e9, 3)
                st["frac_of_fp32_pipe_peak"] = round(iops / (tms * 1e-3) / max(pipe["fp32_mul_add_flops"] / 2.0, 1.0), 4)
            stages[name] = st
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
        dms, dcnt = prof[dom]
        if dom in alg_bytes_per_frame:
            groups = dcnt / (7 if dom == "k_pyr_down" else 1)
            alg = alg_bytes_per_frame[dom] * B
            achieved = alg * groups / (dms * 1e-3) / 1e9
        else:
            # not a streaming kernel: its algorithmic HBM bytes are the records it must read and write per batch
            per_pair = {"k_nn": 2 * nkp * 32 * 2 + nkp * 16, "k_select": total_px // 256 * 4, "k_ransac": nkp * 33,
                        "k_cheirality": nkp * 33}.get(dom, nkp * 32)
            alg = per_pair * B
            achieved = alg * dcnt / (dms * 1e-3) / 1e9
        # DRAM traffic per frame of the streaming kernels from the committed `ncu --set full` capture (profiles/, 1280x1024,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch / frames per launch); null for other sizes / kernels
        ncu_traffic_per_frame = {}
        try:        # written from the committed ncu --set full capture by tools/ncu_summary.py
            ncu_traffic_per_frame = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_per_frame.json"))).get("%dx%d" % (a.width, a.height), {})
        except Exception:
            pass
        traffic = int(ncu_traffic_per_frame[dom] * B) if dom in ncu_traffic_per_frame else None
        # the roofline that actually binds the image kernels is instruction issue (4 warp instructions per clock per SM): warp
        # instructions per frame from the same ncu capture x frames per launch / live launch time / (SMs x 4 x SM clock)
        issue = None
        try:
            wi = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_per_frame.json")))["warp_instructions_per_frame"]["%dx%d" % (a.width, a.height)]
            if dom in wi:
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
                slots = sms * 4 * mhz * 1e6
                rate = wi[dom] * B / (dms / max(dcnt, 1) * 1e-3)
                issue = {"warp_instr_per_launch": int(wi[dom] * B), "achieved_G_per_s": round(rate / 1e9, 1), "peak_G_per_s": round(slots / 1e9, 1),
                         "frac": round(rate / slots, 4), "note": "this, not HBM, bounds the kernel (ncu: issue active 81 %, DRAM 5 %)"}
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 6), "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": round(dms / tot_ms, 4), "algorithmic_bytes_per_launch": int(alg),
                    "avg_launch_ms": round(dms / max(dcnt, 1), 4), "issue": issue,
                    "measured_pipe_peaks": {k: round(v / 1e12, 3) for k, v in pipe.items()}, "measured_pipe_peaks_unit": "T op/s"}

    # ---- CPU baseline beside it (rank 0, bounded sample of the same workload)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:      # the CPU leg is an N=1 measurement
        os.sched_setaffinity(0, all_cpus)
        workers = os.cpu_count() or 1
        pps = min(max(8, 2 * workers), 128)
        fr = (frames[:pps + 1, :, :, 0] if a.ingest == "bgr" else frames[:pps + 1]).contiguous().cpu().numpy()
        ref = CpuReference(fr, Kmat, a.nfeatures, workers, a.ingest)
        t0 = time.perf_counter()
        done = 0
        while True:
            ref.run_pairs(range(pps))
            done += pps
            if time.perf_counter() - t0 >= a.cpu_baseline_seconds:
                break
        dt = time.perf_counter() - t0
        ref.close()
        cpu = {"value": done / dt, "unit": UNIT, "cores": ref.workers, "kind": ref.kind, "cpu": cpu_model(),
               "sample": "%d frame pairs of the same sequence in %.1f s: cv2 chain, both frames' ORB recomputed per pair as the reference does, one worker process per core (cv2 threads=1 each)" % (done, dt)}
        if ref.kind == "reference":
            # BASELINE.md section 3 figure (i), reference-faithful: ONE process, cv2's default thread count, pairs in sequence
            import cv2
            from oracle import cv2_chain
            cv2.setNumThreads(-1)
            cv2_chain.frame_pair(fr[0], fr[1], Kmat, a.nfeatures)
            t0, n1 = time.perf_counter(), 0
            while time.perf_counter() - t0 < 5.0 and n1 + 1 < len(fr):
                cv2_chain.frame_pair(fr[n1], fr[n1 + 1], Kmat, a.nfeatures)
                n1 += 1
            cpu["single_process"] = {"value": n1 / (time.perf_counter() - t0), "unit": UNIT, "cv2_threads": int(cv2.getNumThreads()),
                                     "sample": "%d pairs in sequence in one process, cv2 default threading (as the reference runs)" % n1}

    frames_numel = frames.numel()
    other = None
    if rank == 0 and world == 1 and not a.no_other_configs and a.ingest == "none":
        del frames
        torch.cuda.empty_cache()
        other = other_configs(dev, with_cpu=not a.no_cpu_baseline)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_steps, "warmup": W_steps,
               "ms_per_step": ms / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32/f64",
               "data": "synthetic",
               "config": {"workload": workload_name(a), "pairs_per_step": B, "matcher_engine": a.matcher_engine, "numa_bound_cpus": numa, "frames_resident_MB": round(frames_numel / 1e6, 1),
                          "timing": "median of %d back-to-back regions of %d steps, each bracketed by barrier + synchronize, CUDA events, max over ranks" % (R_regions, K_steps),
                          "region_ms": [round(v, 3) for v in region_ms], "allgather_ms": [round(v, 3) for v in gather_ms],
                          "l2_policy": "inputs larger than L2: every step reads %d new frames (%.0f MB) from a %.0f MB HBM-resident sequence" % (
                              B, B * a.width * a.height / 1e6, frames_numel / 1e6),
                          "parallelism": ("one sequence sharded by frame pair: contiguous blocks per rank with a one-frame halo, no data-path collective, "
                                          "one NCCL all-gather of the per-pair (R,t) records per region, then the 4x4 chain on rank 0") if world > 1 else "single GPU",
                          "chain": {"host_ms": None if chain_ms is None else round(chain_ms, 2), "pairs": world * K_steps * B, "end_position": chain_end},
                          "pairs_ok_fraction": ok_frac, "pair_stats": pair_stats},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * a.width * a.height * (3 if a.ingest == "bgr" else 1),
                       "d2h_bytes_per_step": B * rec, "steps": E_steps,
                       "h2d_ceiling": {"GBps_per_rank": round(h2d_gbs_per_rank, 2), "GBps_all_ranks": round(world * h2d_gbs_per_rank, 2),
                                       "pairs_per_s": round(h2d_pairs_ceiling, 1),
                                       "how": "pinned host -> device copy of the same %d MB of frames, 3 times, all %d rank(s) at once, nothing else running" % (
                                           e2e_frames.numel() // 1000000, world)},
                       "frac_of_h2d_ceiling": round(e2e_value / h2d_pairs_ceiling, 4),
                       "frac_of_resident": round(e2e_value / value, 4)},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages, "cpu_baseline": cpu,
               "other_configs": other}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    # Libraries (NCCL's version banner, torchrun notices) may write to fd 1: keep stdout for the ONE JSON line by pointing
    # fd 1 at stderr while the benchmark runs and restoring it just before printing.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _lines = []
    _print = print

    def print(*a, **k):        # noqa: A001  (collect the JSON line; emitted on the real stdout at the end)
        _lines.append(" ".join(str(x) for x in a))

    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        sys.stdout.flush()
        try:                      # C stdio buffers too (NCCL's banner is an fprintf(stdout)): flush them into stderr now
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
        for ln in _lines:
            _print(ln, flush=True)
