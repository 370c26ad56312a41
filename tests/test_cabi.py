"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU, exports every symbol include/uvo_c.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "uvo_c.h")).read()
    return sorted(set(re.findall(r"UVO_API\s+[\w\s\*]+?\b(uvo_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ergo_uvo_b200 as U
    lib = U.load()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_only_declared_symbols_are_exported():
    import ergo_uvo_b200 as U
    out = subprocess.check_output(["nm", "-D", "--defined-only", U.SO_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {e for e in exported if e.startswith("uvo_")} == set(_declared())


def test_sass_is_sm100a_only():
    import ergo_uvo_b200 as U
    out = subprocess.run(["cuobjdump", "-lelf", U.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header():
    import ergo_uvo_b200 as U
    import numpy as np
    assert U.KEYPOINT_DTYPE.itemsize == 28 and U.DMATCH_DTYPE.itemsize == 16  # cv::KeyPoint / cv::DMatch
    assert C.sizeof(U.Camera) == 96
    p = U.default_params(True)
    assert (p.clip_limit, p.lowe_ratio, p.reprojection_tolerance, p.surf_min_hessian) == (8, 0.8, 3.0, 1500)
    assert p.stereo_gate == 0 and U.default_params(False).stereo_gate == 0  # the reference has no stereo gate
    m = U.default_params(False)
    assert (m.clip_limit, m.lowe_ratio, m.reprojection_tolerance, m.surf_min_hessian, m.min_num_features) == \
        (3, 0.7, 0.1, 50, 20)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ergo_uvo_b200 as U
    with pytest.raises(U.UvoError) as e:
        U.Context(0)
    assert e.value.code == -1  # UVO_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ergo_uvo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "surf.cu" or "the oracle" in txt.lower(), f
                assert "import oracle" not in txt and "from oracle" not in txt and "uvo_oracle" not in txt, f
