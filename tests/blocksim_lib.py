"""ctypes binding of tests/_build/libblocksim.so: the TEST-ONLY build of snap_b200/csrc/sg_deflate.h run by a block of real host threads
(tests/blocksim/blocksim.cpp).  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libblocksim.so")
SRC = os.path.join(HERE, "blocksim", "blocksim.cpp")
DEPS = [SRC, os.path.join(ROOT, "snap_b200", "csrc", "sg_deflate.h"), os.path.join(ROOT, "snap_b200", "csrc", "sg_common.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", SO, SRC], check=True)
    return SO


_lib = None


def bgzf_deflate(data: np.ndarray, threads: int):
    """(BGZF stream, member sizes) of `data` (uint8), every member compressed by a block of `threads` host threads."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.bs_bgzf_deflate.restype = C.c_int64
        _lib.bs_bgzf_deflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    n = int(data.size)
    n_members = (n + 0xff00 - 1) // 0xff00
    out = np.zeros(n + 64 * n_members + 64, dtype=np.uint8)
    sizes = np.zeros(max(1, n_members), dtype=np.uint32)
    used = _lib.bs_bgzf_deflate(data.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), out.size, sizes.ctypes.data_as(C.c_void_p), threads)
    if used < 0:
        raise RuntimeError("bs_bgzf_deflate failed: %d" % used)
    return out[:used].copy(), sizes[:n_members]
