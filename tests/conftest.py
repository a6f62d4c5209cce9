import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_480x360.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def hostsim():
    """CPU build of the kernels' per-element math (tests/hostsim) -- test harness only."""
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    lib = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(ROOT, "droplet_visual_odometry_b200", "csrc", f) for f in ("mathcore.cuh", "select.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", lib, src])
    L = ctypes.CDLL(lib)
    L.hs_harris.restype = ctypes.c_float
    L.hs_fast_atan2.restype = ctypes.c_float
    L.hs_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
    L.hs_sampson.restype = ctypes.c_float
    L.hs_sampson.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
    L.hs_update_iters.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    L.hs_cheirality.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_double] * 5
    return L


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rot_err_deg(Ra, Rb):
    return float(np.degrees(np.arccos(np.clip((np.trace(np.asarray(Ra).reshape(3, 3) @ np.asarray(Rb).reshape(3, 3).T) - 1) / 2, -1, 1))))


def dir_err_deg(ta, tb):
    ta, tb = np.asarray(ta, float).ravel(), np.asarray(tb, float).ravel()
    return float(np.degrees(np.arccos(np.clip(ta @ tb / (np.linalg.norm(ta) * np.linalg.norm(tb)), -1, 1))))


def mask_iou(a, b):
    a, b = np.asarray(a) > 0, np.asarray(b) > 0
    u = np.logical_or(a, b).sum()
    return 1.0 if u == 0 else float(np.logical_and(a, b).sum() / u)
