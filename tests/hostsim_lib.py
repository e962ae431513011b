"""ctypes binding of tests/_build/libhostsim.so: the TEST-ONLY g++ build of snap_b200/csrc/sg_*.h.

Lets the CPU suite diff the engine's algorithm headers against the compiled reference without a GPU.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libhostsim.so")
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
CSRC = os.path.join(ROOT, "snap_b200", "csrc")


def build(force: bool = False) -> str:
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC]
    subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hs_last_error.restype = C.c_char_p
        L.hs_index_open.restype = C.c_void_p
        L.hs_index_open.argtypes = [C.c_char_p]
        L.hs_index_close.argtypes = [C.c_void_p]
        L.hs_index_relayout.argtypes = [C.c_void_p, C.c_double]
        L.hs_lookup_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_tables.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hs_mapq.restype = C.c_int
        L.hs_mapq.argtypes = [C.c_double, C.c_double, C.c_int]
        L.hs_lv_batch.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
        L.hs_ag_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p]
        L.hs_lv_cigar_batch.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
        L.hs_cigar_lv_batch.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
        L.hs_ag_cigar_global_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
        L.hs_ag_cigar_norm_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
        L.hs_cigar_ag_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
        L.hs_bam_single_batch.restype = C.c_int64
        L.hs_bam_single_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 10 + [C.c_int64]
        L.hs_sam_single_batch.restype = C.c_int64
        L.hs_sam_single_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 12 + [C.c_int64]
        L.hs_aligner_create.restype = C.c_void_p
        L.hs_aligner_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.hs_aligner_destroy.argtypes = [C.c_void_p]
        L.hs_aligner_set_two_pass.argtypes = [C.c_void_p, C.c_int]
        L.hs_aligner_deferred.restype = C.c_int64
        L.hs_aligner_deferred.argtypes = [C.c_void_p]
        L.hs_align_single.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.hs_bgzf_deflate.restype = C.c_int64
        L.hs_bgzf_deflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.hs_align_single_secondary.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_paired_create.restype = C.c_void_p
        L.hs_paired_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.hs_paired_retried.restype = C.c_int64
        L.hs_paired_retried.argtypes = [C.c_void_p]
        L.hs_paired_destroy.argtypes = [C.c_void_p]
        L.hs_paired_set_staged.argtypes = [C.c_void_p, C.c_int]
        L.hs_align_paired.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HsIndex:
    def __init__(self, directory: str):
        self.handle = lib().hs_index_open(directory.encode())
        if not self.handle:
            raise RuntimeError(lib().hs_last_error().decode())

    def relayout(self, load: float = 0.0) -> "HsIndex":
        """Switches to the sector-bucket layout (snap_b200/csrc/sg_bucket.h), built from the loaded reference-format tables."""
        if lib().hs_index_relayout(self.handle, load) != 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return self

    def lookup(self, seeds: np.ndarray, n: int, max_hits: int = 512):
        nh = np.zeros(2 * n, dtype=np.int64)
        hits = np.zeros(2 * n * max_hits, dtype=np.uint32)
        probes = np.zeros(n, dtype=np.uint32)
        lib().hs_lookup_seeds(self.handle, _p(seeds), n, max_hits, _p(nh), _p(hits), _p(probes))
        return nh.reshape(n, 2), hits.reshape(n, 2, max_hits), probes


class HsAligner:
    def __init__(self, index: HsIndex, params, max_read_len: int = 400):
        import ctypes
        self.index = index
        self.handle = lib().hs_aligner_create(index.handle, ctypes.byref(params), max_read_len)
        if not self.handle:
            raise RuntimeError(lib().hs_last_error().decode())

    def set_two_pass(self, on: bool):
        lib().hs_aligner_set_two_pass(self.handle, 1 if on else 0)

    @property
    def deferred(self) -> int:
        return lib().hs_aligner_deferred(self.handle)

    def align(self, batch, result_dtype, n_counters):
        res = np.zeros(batch.n, dtype=result_dtype)
        ctr = np.zeros(n_counters, dtype=np.int64)
        rc = lib().hs_align_single(self.handle, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens), _p(res), _p(ctr))
        if rc != 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return res, ctr

    def align_secondary(self, batch, result_dtype, n_counters, max_edit_dist, max_secondary=0x7fffffff, max_per_contig=-1, capacity=64, raw_cap=32):
        res = np.zeros(batch.n, dtype=result_dtype)
        sec = np.zeros((batch.n, capacity), dtype=result_dtype)
        nsec = np.zeros(batch.n, dtype=np.int32)
        ctr = np.zeros(n_counters, dtype=np.int64)
        rc = lib().hs_align_single_secondary(self.handle, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens), _p(res),
                                             max_secondary, max_per_contig, max_edit_dist, capacity, _p(sec), _p(nsec), raw_cap, _p(ctr))
        if rc != 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return res, sec, nsec, ctr


class HsPairedAligner:
    def __init__(self, index: HsIndex, params, paired_params, max_read_len: int = 400, pool_cap: int = 0, cand_cap: int = 0):
        import ctypes
        self.index = index
        self.handle = lib().hs_paired_create(index.handle, ctypes.byref(params), ctypes.byref(paired_params), max_read_len, pool_cap, cand_cap)
        if not self.handle:
            raise RuntimeError(lib().hs_last_error().decode())

    def align(self, batch, result_dtype):
        n_pairs = batch.n // 2
        res = np.zeros(n_pairs, dtype=result_dtype)
        nlv = np.zeros(1, dtype=np.int64)
        nag = np.zeros(1, dtype=np.int64)
        rc = lib().hs_align_paired(self.handle, n_pairs, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens), _p(res), _p(nlv), _p(nag))
        if rc != 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return res, int(nlv[0]), int(nag[0])

    def retried(self) -> int:
        return int(lib().hs_paired_retried(self.handle))

    def set_staged(self, on: bool):
        lib().hs_paired_set_staged(self.handle, 1 if on else 0)

    def __del__(self):
        if getattr(self, "handle", None):
            lib().hs_paired_destroy(self.handle)
            self.handle = None


def lv_cigar_batch(text, pat, jobs, out_dtype):
    out = np.zeros(jobs.size, dtype=out_dtype)
    lib().hs_lv_cigar_batch(_p(text), _p(pat), _p(np.ascontiguousarray(jobs)), jobs.size, _p(out))
    return out


CIGAR_OUT_DTYPE = np.dtype([("kind", "<i4"), ("editDistance", "<i4"), ("addFrontClipping", "<i4"), ("refSpan", "<i4"), ("nOps", "<i4"), ("ops", "<u4", (40,))])


def cigar_lv_batch(index, data, jobs):
    out = np.zeros(jobs.size, dtype=CIGAR_OUT_DTYPE)
    lib().hs_cigar_lv_batch(index.handle, _p(data), _p(np.ascontiguousarray(jobs)), jobs.size, _p(out))
    return out


def ag_cigar_global_batch(text, pat, qual, jobs, out_dtype, params=(1, 4, 6, 1)):
    out = np.zeros(jobs.size, dtype=out_dtype)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().hs_ag_cigar_global_batch(_p(prm), _p(text), _p(pat), _p(qual), _p(np.ascontiguousarray(jobs)), jobs.size, _p(out))
    return out


def ag_cigar_norm_batch(text, pat, qual, jobs, out_dtype, params=(1, 4, 6, 1)):
    out = np.zeros(jobs.size, dtype=out_dtype)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().hs_ag_cigar_norm_batch(_p(prm), _p(text), _p(pat), _p(qual), _p(np.ascontiguousarray(jobs)), jobs.size, _p(out))
    return out


def cigar_ag_batch(index, data, qual, jobs, params=(1, 4, 6, 1)):
    out = np.zeros(jobs.size, dtype=CIGAR_OUT_DTYPE)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().hs_cigar_ag_batch(index.handle, _p(prm), _p(data), _p(qual), _p(np.ascontiguousarray(jobs)), jobs.size, _p(out))
    return out


def sam_single(index, batch, ids, results, use_m=True, use_affine_gap=True, params=(1, 4, 6, 1), paired=False, front_clipped=None, clipped_lens=None) -> bytes:
    """SAM records (text) of a batch from its result records; ids: list of bytes (one per read).  paired: results are pair records."""
    prm = np.ascontiguousarray(params, dtype=np.int32)
    id_buf = np.frombuffer(b"".join(ids), dtype=np.uint8).copy()
    id_lens = np.array([len(x) for x in ids], dtype=np.uint32)
    id_offs = np.concatenate([[0], np.cumsum(id_lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    cap = int(batch.n) * 4096 + 8192
    out = np.zeros(cap, dtype=np.uint8)
    res = np.ascontiguousarray(results)
    n = lib().hs_sam_single_batch(index.handle, _p(prm), 1 if use_m else 0, 1 if use_affine_gap else 0, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets),
                                  _p(batch.lens), _p(id_buf), _p(id_offs), _p(id_lens), None if paired else _p(res), _p(res) if paired else None,
                                  None if front_clipped is None else _p(np.ascontiguousarray(front_clipped, dtype=np.uint32)),
                                  None if clipped_lens is None else _p(np.ascontiguousarray(clipped_lens, dtype=np.uint32)), _p(out), cap)
    assert n >= 0
    return out[:n].tobytes()


def bam_single(index, batch, ids, results, use_m=True, use_affine_gap=True, params=(1, 4, 6, 1), paired=False) -> bytes:
    """BAM records (binary, back to back, no header) of an unpaired batch from its result records."""
    prm = np.ascontiguousarray(params, dtype=np.int32)
    id_buf = np.frombuffer(b"".join(ids), dtype=np.uint8).copy()
    id_lens = np.array([len(x) for x in ids], dtype=np.uint32)
    id_offs = np.concatenate([[0], np.cumsum(id_lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    cap = int(batch.n) * 4096 + 8192
    out = np.zeros(cap, dtype=np.uint8)
    res = np.ascontiguousarray(results)
    n = lib().hs_bam_single_batch(index.handle, _p(prm), 1 if use_m else 0, 1 if use_affine_gap else 0, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets),
                                  _p(batch.lens), _p(id_buf), _p(id_offs), _p(id_lens), None if paired else _p(res), _p(res) if paired else None, _p(out), cap)
    assert n >= 0
    return out[:n].tobytes()


def tables(seed_len=20, n_indel=1200, n_perfect=1001):
    phred = np.zeros(256)
    indel = np.zeros(n_indel)
    perfect = np.zeros(n_perfect)
    thr = np.zeros(72)
    wrap = np.zeros(33, dtype=np.uint32)
    lib().hs_tables(seed_len, _p(phred), _p(indel), n_indel, _p(perfect), n_perfect, _p(thr), _p(wrap))
    return phred, indel, perfect, thr, wrap


def lv_batch(text, pat, qual, jobs, out_dtype):
    out = np.zeros(jobs.size, dtype=out_dtype)
    lib().hs_lv_batch(_p(text), _p(pat), _p(qual), _p(jobs), jobs.size, _p(out))
    return out


def ag_batch(text, pat, qual, jobs, out_dtype, params):
    out = np.zeros(jobs.size, dtype=out_dtype)
    poisoned = np.zeros(jobs.size, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.int32)
    lib().hs_ag_batch(_p(params), _p(text), _p(pat), _p(qual), _p(jobs), jobs.size, _p(out), _p(poisoned))
    return out, poisoned


def bgzf_deflate(data: np.ndarray):
    """(BGZF stream, member sizes) of `data` (uint8) through the host build of sg_deflate.h."""
    n = int(data.size)
    n_members = (n + 0xff00 - 1) // 0xff00
    out = np.zeros(n + 64 * n_members + 64, dtype=np.uint8)
    sizes = np.zeros(max(1, n_members), dtype=np.uint32)
    used = lib().hs_bgzf_deflate(_p(data), n, _p(out), out.size, _p(sizes))
    if used < 0:
        raise RuntimeError("hs_bgzf_deflate: output buffer too small")
    return out[:used].copy(), sizes[:n_members]
