"""The C-ABI library: loads, exports every symbol include/snapgpu.h declares, and fails LOUDLY without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def header_functions():
    src = open(os.path.join(ROOT, "include", "snapgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snapgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from snap_b200 import engine
    L = engine.lib()
    names = header_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), "libsnapgpu.so does not export " + n
    assert sorted(engine.EXPORTS) == names
    assert L.snapgpu_abi_version() == engine.ABI_VERSION


def test_struct_sizes_match_header():
    from snap_b200 import engine
    assert engine.RESULT_DTYPE.itemsize == 88
    p = engine.default_params()
    assert p.struct_size == C.sizeof(engine.Params)
    assert (p.maxHits, p.maxDist, p.numSeedsFromCommandLine, p.extraSearchDepth, p.minReadLength) == (300, 14, 25, 1, 50)
    assert (p.matchReward, p.subPenalty, p.gapOpenPenalty, p.gapExtendPenalty, p.fivePrimeEndBonus, p.threePrimeEndBonus) == (1, 4, 6, 1, 10, 7)


def test_no_cpu_fallback_without_device(tmp_path):
    from snap_b200 import engine
    L = engine.lib()
    if L.snapgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.SnapGpuError) as e:
        engine.Index.open(str(tmp_path))
    assert "no usable CUDA device" in str(e.value)
    with pytest.raises(engine.SnapGpuError):
        import numpy as np
        engine.Index.build(np.frombuffer(b"ACGT" * 100, dtype=np.uint8), [0])


def test_product_never_imports_oracle():
    """The product package and the CUDA sources must not reference oracle/ or the host-simulation build."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snap_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f
                assert "hostsim" not in src.lower(), f
