"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/aar_cuda.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def binding():
    from aar_b200 import binding as b
    b.build()
    return b


def test_library_exports_every_declared_symbol(binding):
    hdr = open(os.path.join(ROOT, "include", "aar_cuda.h")).read()
    declared = set(re.findall(r"\b(aar_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(binding.EXPORTS)
    L = C.CDLL(binding.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    # the initialisation path (include/aar_init.h)
    hdr = open(os.path.join(ROOT, "include", "aar_init.h")).read()
    declared = set(re.findall(r"\b(aar_init_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(binding.INIT_EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header(binding):
    # sizeof() of the C structs in include/aar_cuda.h on LP64 (checked with gcc)
    assert C.sizeof(binding.LmParams) == 56
    assert C.sizeof(binding.LmTrace) == 40
    assert C.sizeof(binding.LmReport) == 48
    assert C.sizeof(binding.Desc) == 176
    assert C.sizeof(binding.InitDesc) == 120


def test_no_cpu_fallback_without_device(binding):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aar_b200 import synth
    rig = synth.make_rig(C=2, M=3, F=5, obs_per_frame=4.0, seed=1)
    with pytest.raises(binding.AarError, match="CUDA"):
        binding.Problem(rig)
    with pytest.raises(binding.AarError, match="CUDA"):
        binding.Initializer.from_rig(rig)
    with pytest.raises(binding.AarError, match="CUDA"):
        import numpy as np
        binding.init_consensus(0.05, np.eye(4)[None], np.eye(4)[None], np.eye(4)[None])


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "automatic-ar_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in src and "libaar_oracle" not in src and "mcm_oracle" not in src and "init_oracle" not in src, f
