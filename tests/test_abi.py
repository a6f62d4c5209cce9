"""CPU: libdvo.so loads and exports every function include/dvo.h declares; error paths that need no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from droplet_visual_odometry_b200 import _native


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dvo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvo_[a-z_0-9]+)\s*\(", src)))


def test_header_lists_the_entry_points():
    names = declared_functions()
    for must in ("dvo_create", "dvo_destroy", "dvo_orb", "dvo_pairs", "dvo_sequence", "dvo_pose_points", "dvo_get_poses",
                 "dvo_get_features", "dvo_last_error", "dvo_tap_image"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), "libdvo.so does not export " + name
    assert b"sm_100a" in lib.dvo_version()


def test_struct_layouts_match_header():
    lib = _native.load_library()
    assert ctypes.sizeof(_native.dvo_config) == lib.dvo_sizeof(0)
    assert _native.POSE_DTYPE.itemsize == lib.dvo_sizeof(1) == 208
    assert ctypes.sizeof(_native.dvo_features) == lib.dvo_sizeof(2)
    assert ctypes.sizeof(_native.dvo_pair_arrays) == lib.dvo_sizeof(3)


def test_create_rejects_bad_config_without_touching_a_device():
    lib = _native.load_library()
    cfg = _native.dvo_config()
    lib.dvo_default_config(ctypes.byref(cfg))
    assert (cfg.width, cfg.height, cfg.nfeatures, cfg.nlevels, cfg.ransac_max_iters) == (1280, 1024, 500, 8, 1000)
    cfg.width = 16
    h = ctypes.c_void_p()
    assert lib.dvo_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1 and not h
    assert lib.dvo_create(None, 0, ctypes.byref(h)) == -1


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.DvoError):
        _native.Context(640, 480)
    from droplet_visual_odometry_b200.visual_odometry_v3 import VisualOdometry
    with pytest.raises(_native.DvoError):
        VisualOdometry(camera_matrix=[[1, 0, 0], [0, 1, 0], [0, 0, 1]])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "droplet_visual_odometry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "import cv2" not in text.replace("import cv2 as cv", "") or \
                    f == "visual_odometry_v3.py", f


def test_shipped_kernels_carry_the_blackwell_instructions():
    """The built library is sm_100a code and the kernels DESIGN.md describes as TMA-staged / tensor-core really are:
    cuobjdump -sass must show UTMALDG (TMA tile load) in k_fast_nms<1> and k_pyr_down<1>, and UTCIMMA (tcgen05.mma int8),
    LDTM (tcgen05.ld) and UBLKCP (bulk copy) in k_nn_tensor.  Static check, no GPU needed."""
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "droplet_visual_odometry_b200", "libdvo.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    body, cur = {}, None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            body[cur].append(line)

    def ops(substr):
        names = [k for k in body if substr in k]
        assert names, substr
        return "\n".join("\n".join(body[k]) for k in names)
    assert "UTMALDG" in ops("k_fast_nmsILb1")
    assert "UTMALDG" in ops("k_pyr_downILb1")
    nn = ops("k_nn_tensorILb0")
    for op in ("UTCIMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS"):
        assert op in nn, op
    assert "POPC" in ops("4k_nnILb1")                   # the integer-pipe matcher engine is shipped too
