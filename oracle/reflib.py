"""ctypes binding of oracle/_ref/libsnapref.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (snap_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsnapref.so")                # the reference's own flags (-O3 -msse)
REF_SO_V3 = os.path.join(HERE, "_ref", "libsnapref_v3.so")          # -O3 -march=x86-64-v3 (second build of the CPU baseline)
SNAP_ALIGNER = os.path.join(HERE, "_ref", "snap-aligner")


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("maxHits", C.c_uint32), ("maxDist", C.c_uint32),
        ("numSeedsFromCommandLine", C.c_uint32), ("seedCoverage", C.c_double),
        ("minWeightToCheck", C.c_uint32), ("extraSearchDepth", C.c_uint32), ("minReadLength", C.c_uint32),
        ("useAffineGap", C.c_int32), ("matchReward", C.c_int32), ("subPenalty", C.c_int32),
        ("gapOpenPenalty", C.c_int32), ("gapExtendPenalty", C.c_int32), ("fivePrimeEndBonus", C.c_int32),
        ("threePrimeEndBonus", C.c_int32), ("noUkkonen", C.c_int32), ("noOrderedEvaluation", C.c_int32),
        ("noTruncation", C.c_int32), ("noEditDistance", C.c_int32), ("noBandedAffineGap", C.c_int32),
        ("altAwareness", C.c_int32), ("maxScoreGapToPreferNonAltAlignment", C.c_int32),
        ("explorePopularSeeds", C.c_int32), ("stopOnFirstHit", C.c_int32),
        ("maxSecondaryAlignmentAdditionalEditDistance", C.c_int32), ("ignoreAlignmentAdjustmentsForOm", C.c_int32),
    ]


def default_params(**kw) -> Params:
    """`snap single` 2.0.5 defaults (reference AlignerOptions.cpp:39-121)."""
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.maxHits = 300
    p.maxDist = 14
    p.numSeedsFromCommandLine = 25
    p.seedCoverage = 0.0
    p.minWeightToCheck = 1
    p.extraSearchDepth = 1
    p.minReadLength = 50
    p.useAffineGap = 1
    p.matchReward, p.subPenalty, p.gapOpenPenalty, p.gapExtendPenalty = 1, 4, 6, 1
    p.fivePrimeEndBonus, p.threePrimeEndBonus = 10, 7
    p.altAwareness = 1
    p.maxScoreGapToPreferNonAltAlignment = 64
    p.maxSecondaryAlignmentAdditionalEditDistance = -1
    p.ignoreAlignmentAdjustmentsForOm = 1
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("direction", "<i4"), ("location", "<i8"), ("origLocation", "<i8"),
    ("score", "<i4"), ("scorePriorToClipping", "<i4"), ("mapq", "<i4"), ("clippingForReadAdjustment", "<i4"),
    ("usedAffineGapScoring", "<i4"), ("basesClippedBefore", "<i4"), ("basesClippedAfter", "<i4"),
    ("agScore", "<i4"), ("supplementary", "<i4"), ("seedOffset", "<i4"),
    ("matchProbability", "<f8"), ("probabilityAllCandidates", "<f8"),
    ("popularSeedsSkipped", "<u4"), ("reserved", "<u4"),
])
assert RESULT_DTYPE.itemsize == 88

COUNTER_FIELDS = ["totalReads", "uselessReads", "singleHits", "multiHits", "notFound", "nHashTableLookups",
                  "nHashEntriesProbed", "nOverflowWordsRead", "lvCalls", "affineGapCalls",
                  "nHitsIgnoredBecauseOfTooHighPopularity"]
N_COUNTERS = len(COUNTER_FIELDS) + 71

LV_JOB_DTYPE = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"),
                         ("k", "<i4"), ("dir", "<i4")])
LV_OUT_DTYPE = np.dtype([("score", "<i4"), ("netIndel", "<i4"), ("totalIndels", "<i4"), ("textSpan", "<i4"),
                         ("matchProbability", "<f8")])
AG_JOB_DTYPE = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"),
                         ("w", "<i4"), ("scoreInit", "<i4"), ("dir", "<i4"), ("isRC", "<i4"), ("banded", "<i4"),
                         ("useClippingOptimizations", "<i4")])
AG_OUT_DTYPE = np.dtype([("agScore", "<i4"), ("textOffset", "<i4"), ("patternOffset", "<i4"), ("nEdits", "<i4"),
                         ("matchProbability", "<f8")])
AG_PARAMS_DEFAULT = np.array([1, 4, 6, 1, 10, 7], dtype=np.int32)


def counters_dict(arr: np.ndarray) -> dict:
    d = {k: int(arr[i]) for i, k in enumerate(COUNTER_FIELDS)}
    d["mapqHistogram"] = [int(x) for x in arr[len(COUNTER_FIELDS):]]
    return d


def available() -> bool:
    return os.path.exists(REF_SO)


_lib = None
_libs = {}


def available_v3() -> bool:
    """The AVX2 build exists and this host can run it."""
    if not os.path.exists(REF_SO_V3):
        return False
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f in flags for f in (" avx2", " bmi2", " fma"))
    except Exception:
        return False


def lib(build: str = "stock"):
    """build: "stock" (the reference's own flags) or "v3" (-march=x86-64-v3).  Each is its own shared object with its own globals."""
    global _lib
    if build == "v3":
        if "v3" not in _libs:
            if not available_v3():
                raise RuntimeError("oracle/_ref/libsnapref_v3.so missing or this CPU lacks AVX2")
            _libs["v3"] = _load(REF_SO_V3)
        return _libs["v3"]
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libsnapref.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = _load(REF_SO)
    return _lib


def _load(path):
    L = C.CDLL(path)
    L.ref_index_load.restype = C.c_void_p
    L.ref_index_load.argtypes = [C.c_char_p]
    L.ref_single_create.restype = C.c_void_p
    L.ref_single_create.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.ref_single_destroy.argtypes = [C.c_void_p]
    L.ref_single_align.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    L.ref_single_create_om.restype = C.c_void_p
    L.ref_single_create_om.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
    L.ref_single_align_om.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_single_align_mt.restype = C.c_double
    L.ref_single_align_mt.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int64] + [C.c_void_p] * 6
    L.ref_lookup_seed.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    L.ref_wrapped_seed.restype = C.c_uint
    L.ref_wrapped_seed.argtypes = [C.c_uint, C.c_uint]
    L.ref_mapq.restype = C.c_int
    L.ref_mapq.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    L.ref_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ref_lv_batch.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
    L.ref_ag_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
    L.ref_lv_cigar_batch.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
    L.ref_cigar_lv_batch.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
    L.ref_ag_cigar_global_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
    L.ref_ag_cigar_norm_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
    L.ref_cigar_ag_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
    L.ref_write_reads_batch.restype = C.c_int64
    L.ref_write_reads_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 9 + [C.c_int64]
    L.ref_write_pairs_batch.restype = C.c_int64
    L.ref_write_pairs_batch.argtypes = L.ref_write_reads_batch.argtypes
    L.ref_decode_cigar.restype = C.c_int
    L.ref_decode_cigar.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    L.ref_index_info.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_index_load_ex.restype = C.c_void_p
    L.ref_index_load_ex.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.ref_numa_interleave.restype = C.c_int
    L.ref_single_align_mt_reps.restype = C.c_double
    L.ref_single_align_mt_reps.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 6
    L.ref_paired_align_mt_reps.restype = C.c_double
    L.ref_paired_align_mt_reps.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(PairedParams), C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 7
    L.ref_paired_create.restype = C.c_void_p
    L.ref_paired_create.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(PairedParams)]
    L.ref_paired_destroy.argtypes = [C.c_void_p]
    L.ref_paired_align.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    L.ref_paired_align_mt.restype = C.c_double
    L.ref_paired_align_mt.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(PairedParams), C.c_int, C.c_int64] + [C.c_void_p] * 7
    L.ref_init()
    return L


def set_thread_pinning(on: bool, builds=("stock", "v3")) -> None:
    """Stock SNAP's -b for the harness's multi-threaded runs: thread t -> logical CPU t (see ref_set_thread_pinning)."""
    for b in builds:
        if b == "stock" or (b in _libs):
            lib(b).ref_set_thread_pinning(1 if on else 0)


def numa_interleave() -> int:
    """MPOL_INTERLEAVE over all memory nodes for this thread and the threads it creates (see ref_numa_interleave): memory nodes
    interleaved over, 1 when there is a single node, -1 if refused."""
    return int(lib().ref_numa_interleave())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefIndex:
    def __init__(self, directory: str, build: str = "stock", map_files: bool = False, prefetch: bool = False):
        """map_files / prefetch: GenomeIndex::loadFromDirectory's own arguments (stock SNAP defaults to -map -pre)."""
        self.build = build
        self.L = lib(build)
        self.handle = self.L.ref_index_load_ex(directory.encode(), 1 if map_files else 0, 1 if prefetch else 0)
        if not self.handle:
            raise RuntimeError("reference failed to load index " + directory)
        self.directory = directory

    def lookup(self, seed: bytes, max_out: int = 512):
        nh = np.zeros(2, dtype=np.int64)
        hits = np.zeros(2 * max_out, dtype=np.uint32)
        extra = np.zeros(1, dtype=np.int64)
        lib().ref_lookup_seed(self.handle, seed, _p(nh), _p(hits), max_out, _p(extra))
        return (int(nh[0]), int(nh[1]), hits[:min(int(nh[0]), max_out)].copy(),
                hits[max_out:max_out + min(int(nh[1]), max_out)].copy(), int(extra[0]))


class RefSingleAligner:
    def __init__(self, index: RefIndex, params: Params):
        self.index = index
        self.params = params
        self.handle = lib().ref_single_create(index.handle, C.byref(params))

    def align(self, batch):
        res = np.zeros(batch.n, dtype=RESULT_DTYPE)
        ctr = np.zeros(N_COUNTERS, dtype=np.int64)
        lib().ref_single_align(self.handle, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens),
                               _p(res), _p(ctr))
        return res, counters_dict(ctr)

    def close(self):
        if self.handle:
            lib().ref_single_destroy(self.handle)
            self.handle = None


class RefSecondaryAligner(RefSingleAligner):
    """`snap single -om` (params.maxSecondaryAlignmentAdditionalEditDistance >= 0): AlignRead with a secondaryResults buffer that doubles on
    overflow, as SingleAligner.cpp:137-142 / :250-263."""

    def __init__(self, index: RefIndex, params: Params, max_secondary: int = 0x7fffffff, max_per_contig: int = -1):
        self.index = index
        self.params = params
        self.max_secondary = max_secondary
        self.handle = lib().ref_single_create_om(index.handle, C.byref(params), max_per_contig)

    def align(self, batch, capacity: int = 64):
        res = np.zeros(batch.n, dtype=RESULT_DTYPE)
        sec = np.zeros((batch.n, capacity), dtype=RESULT_DTYPE)
        nsec = np.zeros(batch.n, dtype=np.int32)
        ctr = np.zeros(N_COUNTERS, dtype=np.int64)
        lib().ref_single_align_om(self.handle, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens), _p(res),
                                  self.max_secondary, capacity, _p(sec), _p(nsec), _p(ctr))
        return res, sec, nsec, counters_dict(ctr)


def align_mt(index: RefIndex, params: Params, batch, threads: int, reps: int = 1):
    """All-threads run: the threads exist and wait at a barrier before the clock starts; `reps` passes over the batch inside the
    timed region (seconds returned are for all of them; counters are those of one pass)."""
    res = np.zeros(batch.n, dtype=RESULT_DTYPE)
    ctr = np.zeros(N_COUNTERS, dtype=np.int64)
    secs = index.L.ref_single_align_mt_reps(index.handle, C.byref(params), threads, reps, batch.n, _p(batch.bases), _p(batch.quals),
                                            _p(batch.offsets), _p(batch.lens), _p(res), _p(ctr))
    return res, counters_dict(ctr), float(secs)


class PairedParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("minSpacing", C.c_int32), ("maxSpacing", C.c_uint32),
                ("intersectingAlignerMaxHits", C.c_uint32), ("maxCandidatePoolSize", C.c_uint32), ("maxSeedsSingleEnd", C.c_uint32),
                ("maxDistForIndels", C.c_uint32), ("forceSpacing", C.c_int32), ("minScoreRealignment", C.c_int32),
                ("minScoreGapRealignmentALT", C.c_int32), ("minAGScoreImprovement", C.c_int32),
                ("enableHammingScoringBaseAligner", C.c_int32), ("useSoftClipping", C.c_int32), ("flattenMAPQAtOrBelow", C.c_int32)]


def default_paired_params(**kw) -> PairedParams:
    """`snap paired` 2.0.5 defaults (PairedAligner.cpp:55-57, 228-243; AlignerOptions.cpp:101-102)."""
    p = PairedParams()
    p.struct_size = C.sizeof(PairedParams)
    p.minSpacing, p.maxSpacing = 0, 1000
    p.intersectingAlignerMaxHits = 4000
    p.maxCandidatePoolSize = 1000000
    p.maxSeedsSingleEnd = 25
    p.maxDistForIndels = 40
    p.minScoreRealignment, p.minScoreGapRealignmentALT, p.minAGScoreImprovement = 3, 3, 24
    p.enableHammingScoringBaseAligner = 1
    p.useSoftClipping = 1
    p.flattenMAPQAtOrBelow = 3
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def default_params_paired(**kw) -> Params:
    """The shared options with their paired-end defaults: -n 8 (AlignerOptions.cpp:107-111)."""
    kw.setdefault("numSeedsFromCommandLine", 8)
    return default_params(**kw)


PAIRED_RESULT_DTYPE = np.dtype([
    ("status", "<i4", 2), ("direction", "<i4", 2), ("location", "<i8", 2), ("origLocation", "<i8", 2), ("score", "<i4", 2),
    ("scorePriorToClipping", "<i4", 2), ("mapq", "<i4", 2), ("clippingForReadAdjustment", "<i4", 2), ("usedAffineGapScoring", "<i4", 2),
    ("basesClippedBefore", "<i4", 2), ("basesClippedAfter", "<i4", 2), ("agScore", "<i4", 2), ("supplementary", "<i4", 2),
    ("seedOffset", "<i4", 2), ("lvIndels", "<i4", 2), ("usedGaplessClipping", "<i4", 2), ("refSpan", "<i4", 2), ("liftover", "<i4", 2),
    ("popularSeedsSkipped", "<u4", 2), ("alignedAsPair", "<i4"), ("agForcedSingleAlignerCall", "<i4"),
    ("matchProbability", "<f8", 2), ("probabilityAllPairs", "<f8"),
])
assert PAIRED_RESULT_DTYPE.itemsize == 200


class RefPairedAligner:
    def __init__(self, index: RefIndex, params: Params, pparams: PairedParams):
        L = index.L
        self.index = index
        self.handle = L.ref_paired_create(index.handle, C.byref(params), C.byref(pparams))

    def align(self, batch):
        n = batch.n // 2
        res = np.zeros(n, dtype=PAIRED_RESULT_DTYPE)
        lvag = np.zeros(2, dtype=np.int64)
        rc = lib().ref_paired_align(self.handle, n, _p(batch.bases), _p(batch.quals), _p(batch.offsets), _p(batch.lens), _p(res),
                                    C.c_void_p(lvag.ctypes.data), C.c_void_p(lvag.ctypes.data + 8))
        if rc != 0:
            raise RuntimeError("ref_paired_align failed")
        return res, {"lvCalls": int(lvag[0]), "affineGapCalls": int(lvag[1])}

    def close(self):
        if self.handle:
            lib().ref_paired_destroy(self.handle)
            self.handle = None


def paired_align_mt(index: RefIndex, params: Params, pparams: PairedParams, batch, n_threads: int, reps: int = 1):
    """All-cores paired run (one aligner stack per thread): (results, {lvCalls, affineGapCalls}, seconds).  `reps` as in align_mt."""
    n = batch.n // 2
    res = np.zeros(n, dtype=PAIRED_RESULT_DTYPE)
    lvag = np.zeros(2, dtype=np.int64)
    secs = index.L.ref_paired_align_mt_reps(index.handle, C.byref(params), C.byref(pparams), n_threads, reps, n, _p(batch.bases), _p(batch.quals), _p(batch.offsets),
                                            _p(batch.lens), _p(res), C.c_void_p(lvag.ctypes.data), C.c_void_p(lvag.ctypes.data + 8))
    if secs < 0:
        raise RuntimeError("ref_paired_align_mt failed")
    return res, {"lvCalls": int(lvag[0]), "affineGapCalls": int(lvag[1])}, float(secs)


def fastq_parse(text: np.ndarray, clipping: int = 2, max_reads: int | None = None):
    """FASTQReader::getReadFromBuffer + Read::clip over a whole buffer (SURVEY 8f N2 oracle).  Returns
    (bases, quals, offsets, lens, id_offsets, id_lens, front_clipped, bytes_consumed)."""
    L = lib()
    L.ref_fastq_parse.restype = C.c_int64
    L.ref_fastq_parse.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.POINTER(C.c_int64)] + [C.c_void_p] * 7
    buf = np.concatenate([np.ascontiguousarray(text, dtype=np.uint8), np.zeros(16, dtype=np.uint8)])
    cap = max_reads if max_reads is not None else int((text == 10).sum()) // 4 + 1
    bases = np.zeros(text.size + 16, dtype=np.uint8)
    quals = np.zeros(text.size + 16, dtype=np.uint8)
    offs = np.zeros(cap, dtype=np.uint64); lens = np.zeros(cap, dtype=np.uint32)
    ido = np.zeros(cap, dtype=np.uint64); idl = np.zeros(cap, dtype=np.uint32); fc = np.zeros(cap, dtype=np.uint32)
    n = C.c_int64(0)
    used = L.ref_fastq_parse(_p(buf), text.size, clipping, cap, C.byref(n), _p(bases), _p(quals), _p(offs), _p(lens), _p(ido), _p(idl), _p(fc))
    r = n.value
    total = int(offs[r - 1] + lens[r - 1]) if r else 0
    return bases[:total], quals[:total], offs[:r], lens[:r], ido[:r], idl[:r], fc[:r], int(used)


def lv_batch(text: np.ndarray, pat: np.ndarray, qual: np.ndarray, jobs: np.ndarray) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=LV_OUT_DTYPE)
    lib().ref_lv_batch(_p(text), _p(pat), _p(qual), _p(jobs), jobs.size, _p(out))
    return out


# LandauVishkinWithCigar (output stage): test-local POD records, see oracle/ref_harness.cpp
LVC_JOB_DTYPE = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"), ("k", "<i4"), ("useM", "<i4")])
LVC_OUT_DTYPE = np.dtype([("score", "<i4"), ("nOps", "<i4"), ("textUsed", "<i4"), ("netIndel", "<i4"), ("normalizedScore", "<i4"),
                          ("addFrontClipping", "<i4"), ("ops", "<u4", (32,))])


def lv_cigar_batch(text: np.ndarray, pat: np.ndarray, jobs: np.ndarray) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=LVC_OUT_DTYPE)
    lib().ref_lv_cigar_batch(_p(text), _p(pat), _p(np.ascontiguousarray(jobs, dtype=LVC_JOB_DTYPE)), jobs.size, _p(out))
    return out


def decode_cigar(ops: np.ndarray, n_ops: int) -> str:
    buf = C.create_string_buffer(512)
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    ok = lib().ref_decode_cigar(_p(ops), int(n_ops), buf, 512)
    assert ok
    return buf.value.decode()


# SAMFormat::computeCigarString (LV overload): test-local POD records, see oracle/ref_harness.cpp
CIGAR_JOB_DTYPE = np.dtype([("dataOff", "<u8"), ("location", "<i8"), ("dataLength", "<i4"), ("basesClippedBefore", "<i4"), ("extraBasesClippedBefore", "<i4"),
                            ("basesClippedAfter", "<i4"), ("frontHardClipping", "<i4"), ("backHardClipping", "<i4"), ("direction", "<i4"), ("useM", "<i4")])
CIGAR_REF_OUT_DTYPE = np.dtype([("kind", "<i4"), ("editDistance", "<i4"), ("addFrontClipping", "<i4"), ("refSpan", "<i4"), ("cigar", "S240")])


def cigar_lv_batch(index, data: np.ndarray, jobs: np.ndarray) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=CIGAR_REF_OUT_DTYPE)
    lib().ref_cigar_lv_batch(index.handle, _p(data), _p(np.ascontiguousarray(jobs, dtype=CIGAR_JOB_DTYPE)), jobs.size, _p(out))
    return out


# AffineGapVectorizedWithCigar::computeGlobalScore: test-local POD records, see oracle/ref_harness.cpp
AGC_JOB_DTYPE = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"), ("w", "<i4"), ("useM", "<i4")])
AGC_OUT_DTYPE = np.dtype([("score", "<i4"), ("nOps", "<i4"), ("netDel", "<i4"), ("tailIns", "<i4"), ("ops", "<u4", (64,))])


def ag_cigar_global_batch(text, pat, qual, jobs, params=(1, 4, 6, 1)) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=AGC_OUT_DTYPE)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().ref_ag_cigar_global_batch(_p(prm), _p(text), _p(pat), _p(qual), _p(np.ascontiguousarray(jobs, dtype=AGC_JOB_DTYPE)), jobs.size, _p(out))
    return out


AGC_NORM_OUT_DTYPE = np.dtype([("score", "<i4"), ("nOps", "<i4"), ("netDel", "<i4"), ("tailIns", "<i4"), ("addFrontClipping", "<i4"), ("ops", "<u4", (64,))])


def ag_cigar_norm_batch(text, pat, qual, jobs, params=(1, 4, 6, 1)) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=AGC_NORM_OUT_DTYPE)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().ref_ag_cigar_norm_batch(_p(prm), _p(text), _p(pat), _p(qual), _p(np.ascontiguousarray(jobs, dtype=AGC_JOB_DTYPE)), jobs.size, _p(out))
    return out


CIGAR_AG_JOB_DTYPE = np.dtype(CIGAR_JOB_DTYPE.descr + [("score", "<i4"), ("pad", "<i4")])


def cigar_ag_batch(index, data: np.ndarray, qual: np.ndarray, jobs: np.ndarray, params=(1, 4, 6, 1)) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=CIGAR_REF_OUT_DTYPE)
    prm = np.ascontiguousarray(params, dtype=np.int32)
    lib().ref_cigar_ag_batch(index.handle, _p(prm), _p(data), _p(qual), _p(np.ascontiguousarray(jobs, dtype=CIGAR_AG_JOB_DTYPE)), jobs.size, _p(out))
    return out


def write_reads(index, batch, ids, results, use_m=True, use_affine_gap=True, params=(1, 4, 6, 1), paired=False) -> bytes:
    """SimpleReadWriter::writeReads (SAM) for every read of the batch with the given result records; returns the text."""
    prm = np.ascontiguousarray(params, dtype=np.int32)
    id_buf = np.frombuffer(b"".join(ids) + b"\0", dtype=np.uint8).copy()
    id_lens = np.array([len(x) for x in ids], dtype=np.uint32)
    id_offs = np.concatenate([[0], np.cumsum(id_lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    cap = int(batch.n) * 4096 + 8192
    out = np.zeros(cap, dtype=np.uint8)
    res = np.ascontiguousarray(results)
    fn = lib().ref_write_pairs_batch if paired else lib().ref_write_reads_batch
    n = fn(index.handle, _p(prm), 1 if use_m else 0, 1 if use_affine_gap else 0, batch.n, _p(batch.bases), _p(batch.quals), _p(batch.offsets),
                                    _p(batch.lens), _p(id_buf), _p(id_offs), _p(id_lens), _p(res), _p(out), cap)
    assert n >= 0
    return out[:n].tobytes()


def ag_batch(text: np.ndarray, pat: np.ndarray, qual: np.ndarray, jobs: np.ndarray, params=AG_PARAMS_DEFAULT) -> np.ndarray:
    out = np.zeros(jobs.size, dtype=AG_OUT_DTYPE)
    params = np.ascontiguousarray(params, dtype=np.int32)
    lib().ref_ag_batch(_p(params), _p(text), _p(pat), _p(qual), _p(jobs), jobs.size, _p(out))
    return out


def tables(n_indel: int = 1200, n_perfect: int = 1001):
    phred = np.zeros(256)
    indel = np.zeros(n_indel)
    perfect = np.zeros(n_perfect)
    lib().ref_tables(_p(phred), _p(indel), n_indel, _p(perfect), n_perfect)
    return phred, indel, perfect


def build_reference_index(snap_aligner: str, fasta: str, out_dir: str, seed_len: int = 20, large: bool = False,
                          threads: int = 8) -> None:
    """Runs the stock reference CLI `snap-aligner index` (oracle/_ref/snap-aligner). Test infrastructure only."""
    import subprocess
    os.makedirs(out_dir, exist_ok=True)
    cmd = [snap_aligner, "index", fasta, out_dir, "-s", str(seed_len), "-t%d" % threads]
    if large:
        cmd.append("-large")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0 or not os.path.exists(os.path.join(out_dir, "GenomeIndexHash")):
        raise RuntimeError("snap-aligner index failed:\n" + res.stdout)
