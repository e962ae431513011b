"""Parity tests proper: the CUDA path, called through the C ABI (snap_b200.engine -> libsnapgpu.so), against the
compiled reference (oracle/_ref, when its prebuilt files travelled to the box) and against the committed golden
fixtures.  Bit-exact for every field; doubles compared bitwise."""
import json
import os

import numpy as np
import pytest

import jobs as J
from conftest import OPTION_SETS, PAIRED_OPTION_SETS, differing, differing_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from snap_b200 import engine as e
    assert e.lib().snapgpu_device_count() >= 1, "no CUDA device: the CUDA path has no fallback"
    return e


@pytest.fixture(scope="module")
def gidx(engine, small_cfg):
    """The reference-built index directory in HBM, in the library's default layout: 32-byte sector buckets (sg_bucket.h)."""
    ix = engine.Index.open(small_cfg.idx)
    assert ix.info().reserved == 1
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def gidx_snap(engine, small_cfg):
    """The same directory kept in the reference's own table layout (SNAPGPU_INDEX_LAYOUT=snap): probe-for-probe parity of
    SNAPHashTable::GetFirstValueForKey."""
    old = os.environ.get("SNAPGPU_INDEX_LAYOUT")
    os.environ["SNAPGPU_INDEX_LAYOUT"] = "snap"
    try:
        ix = engine.Index.open(small_cfg.idx)
    finally:
        if old is None:
            del os.environ["SNAPGPU_INDEX_LAYOUT"]
        else:
            os.environ["SNAPGPU_INDEX_LAYOUT"] = old
    assert ix.info().reserved == 0
    yield ix
    ix.close()


def test_unit_vectors_cuda(engine, golden_dir):
    """The reference's own LV / affine-gap known-answer vectors on the CUDA leaves."""
    v = json.load(open(os.path.join(golden_dir, "ref_unit_vectors.json")))
    for e in v["lv"]:
        text = np.frombuffer(b"n" * 64 + e["text"].encode() + b"\0" * 64, dtype=np.uint8)
        pat = np.frombuffer(e["pattern"].encode() + b"\0" * 16, dtype=np.uint8)
        qual = np.full(pat.size, ord("5"), dtype=np.uint8)
        job = np.zeros(1, dtype=J.LV_JOB)
        job[0] = (64, 0, e["textLen"], e["patternLen"], e["k"], 1)
        assert int(engine.test_lv(text, pat, qual, job, J.LV_OUT)[0]["score"]) == e["expected"], e
    for e in v["ag"]:
        text = np.frombuffer(b"n" * 64 + e["text"].encode() + b"n" * 64, dtype=np.uint8)
        pat = np.frombuffer(e["pattern"].encode() + b"A" * 8, dtype=np.uint8)
        qual = np.full(pat.size, ord("2"), dtype=np.uint8)
        job = np.zeros(1, dtype=J.AG_JOB)
        job[0] = (64, 0, e["textLen"], e["patternLen"], e["w"], e["scoreInit"], 1, 0, 0, 0)
        assert int(engine.test_ag(text, pat, qual, job, J.AG_OUT, v["ag_params"])[0]["agScore"]) == e["expected"], e


def test_leaf_golden_lv(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "leaf_lv.npz"))
    got = engine.test_lv(g["text"], g["pat"], g["qual"], g["jobs"], J.LV_OUT)
    assert J.same_out(g["out"], got).all()


def test_leaf_golden_ag(engine, golden_dir):
    """One job per thread, each thread with its own (initially zero) traceback array: jobs whose reference answer depends
    on bits left by the *previous* job of a sequential run are not comparable here and are skipped (they are covered,
    with matching history, by the CPU-side fuzz and by the whole-read tests)."""
    g = np.load(os.path.join(golden_dir, "leaf_ag.npz"))
    got = engine.test_ag(g["text"], g["pat"], g["qual"], g["jobs"], J.AG_OUT, [1, 4, 6, 1, 10, 7])
    same = J.same_out(g["out"], got)
    # exactly the history-dependent jobs may differ (flagged by the scalar restatement's poison check), nothing else
    hist = _history_dependent(g["text"], g["pat"], g["qual"], g["jobs"], [1, 4, 6, 1, 10, 7])
    assert hist.sum() <= 0.01 * hist.size
    assert same[~hist].all()
    assert (g["out"]["agScore"] == got["agScore"]).all()
    assert (g["out"]["textOffset"] == got["textOffset"]).all() and (g["out"]["patternOffset"] == got["patternOffset"]).all()


def _history_dependent(t, p, q, jb, params):
    """Jobs whose REFERENCE answer depends on what earlier calls left in its never-cleared traceback array (the traceback
    steps onto a cell this call's DP did not write).  Flagged by the host build of the scalar restatement (test-only)."""
    import hostsim_lib as hs
    _, poisoned = hs.ag_batch(t, p, q, jb, J.AG_OUT, params)
    return poisoned != 0


def test_warp_leaves_golden(engine, golden_dir):
    """The warp-cooperative forms (what the alignment kernels run).  LV is a pure function: any number of warps.
    Affine gap: every job whose reference answer is a function of its inputs must match bit for bit, on one warp or many.
    The history-dependent ones (see _history_dependent; < 1 %) still have to agree on score and offsets."""
    g = np.load(os.path.join(golden_dir, "leaf_lv.npz"))
    got = engine.test_lv(g["text"], g["pat"], g["qual"], g["jobs"], J.LV_OUT, warps=512)
    assert J.same_out(g["out"], got).all()
    g = np.load(os.path.join(golden_dir, "leaf_ag.npz"))
    params = [1, 4, 6, 1, 10, 7]
    hist = _history_dependent(g["text"], g["pat"], g["qual"], g["jobs"], params)
    assert hist.sum() <= 0.01 * hist.size
    for warps in (1, 256):
        got = engine.test_ag(g["text"], g["pat"], g["qual"], g["jobs"], J.AG_OUT, params, warps=warps)
        same = J.same_out(g["out"], got)
        assert same[~hist].all(), warps
        assert (g["out"]["agScore"] == got["agScore"]).all() and (g["out"]["textOffset"] == got["textOffset"]).all()
        assert (g["out"]["patternOffset"] == got["patternOffset"]).all()


@pytest.mark.parametrize("packed", ["1", "0", "2", "5", "13"])
def test_warp_leaves_fuzz_vs_scalar_and_reference(engine, reflib, monkeypatch, packed):
    """packed = 1: unbanded jobs take the s16x2 (DPX) form used by the paired kernel (2: its unrolled instantiation); 0: the int form;
    5: narrow bands take the experimental two-units-per-step form of sg_warp_ag_duo.cuh (13: with its per-round H in the arena)."""
    monkeypatch.setenv("SNAPGPU_TEST_AG_PACKED", packed)
    for seed in (301, 302):
        t, p, q, jb = J.lv_jobs(3000, seed)
        want = reflib.lv_batch(t, p, q, jb.astype(reflib.LV_JOB_DTYPE))
        assert J.same_out(want, engine.test_lv(t, p, q, jb, J.LV_OUT, warps=256)).all()
        assert J.same_out(want, engine.test_lv(t, p, q, jb, J.LV_OUT)).all()
        t, p, q, jb = J.ag_jobs(2500, seed + 10)
        want = reflib.ag_batch(t, p, q, jb.astype(reflib.AG_JOB_DTYPE))
        hist = _history_dependent(t, p, q, jb, reflib.AG_PARAMS_DEFAULT)
        got = engine.test_ag(t, p, q, jb, J.AG_OUT, reflib.AG_PARAMS_DEFAULT, warps=64)
        assert J.same_out(want, got)[~hist].all()
        assert (want["agScore"] == got["agScore"]).all() and (want["textOffset"] == got["textOffset"]).all()


@pytest.mark.parametrize("layout", ["bucket", "snap"])
def test_lookup_matches_reference(engine, gidx, gidx_snap, small_cfg, reflib, layout):
    """Hit sets (and their order) of both layouts against the reference's lookupSeed32; on the reference layout also the number of
    entries examined, probe for probe."""
    ix = gidx if layout == "bucket" else gidx_snap
    ridx = reflib.RefIndex(small_cfg.idx)
    rb = small_cfg.reads["noisy150"]
    seeds = []
    for i in range(600):
        b = rb.read(i)[0]
        if len(b) >= 60:
            seeds += [b[0:20], b[37:57]]
    from snap_b200 import synth
    c = small_cfg.contigs[1]
    seeds += [c[o:o + 20].tobytes() for o in range(0, 4000)] + [synth.revcomp(c[o:o + 20]).tobytes() for o in range(0, 4000, 3)]
    arr = np.frombuffer(b"".join(seeds), dtype=np.uint8)
    nh, hits, probes = ix.lookup_seeds(arr, len(seeds), 512)
    for i, s in enumerate(seeds):
        a = ridx.lookup(s, 512)
        assert (a[0], a[1]) == (nh[i, 0], nh[i, 1]), (i, s)
        assert np.array_equal(a[2], hits[i, 0, :min(a[0], 512)]) and np.array_equal(a[3], hits[i, 1, :min(a[1], 512)])
        if b"N" not in s:
            if layout == "snap":
                assert a[4] + 2 == probes[i]
            else:
                assert probes[i] >= 4 and probes[i] % 4 == 0          # whole 32-byte buckets
    if layout == "bucket":
        assert np.mean(probes[probes > 0]) < 6.0                     # ~1.1 sectors per lookup at the default load


def test_reference_layout_counts_entries_like_the_reference(engine, gidx_snap, small_cfg, reflib):
    """On the reference's own table layout the work counters match the reference's probe for probe (nHashEntriesProbed included)."""
    for opt in ("default_d14", "ne_d20"):
        kw = OPTION_SETS[opt]
        al = engine.SingleAligner(gidx_snap, engine.default_params(**kw), 4096)
        ridx = reflib.RefIndex(small_cfg.idx)
        for name in ("std150", "noisy150"):
            rb = small_cfg.reads[name]
            want, wctr = reflib.RefSingleAligner(ridx, reflib.default_params(**kw)).align(rb)
            got, g = al.align(rb)
            assert differing(want, got) == [], (opt, name)
            for k in ("totalReads", "nHashTableLookups", "nHashEntriesProbed", "lvCalls", "affineGapCalls", "mapqHistogram"):
                assert wctr[k] == g[k], (opt, name, k)
        al.close()


def test_device_built_bucket_index_saves_in_the_reference_format(engine, small_cfg, reflib, tmp_path):
    """snapgpu_index_save of a device-built sector-bucket index: the directory is the reference's own format (its tables), loadable by
    the unmodified reference, with the same hit sets."""
    bases, starts = small_cfg.padded_bases()
    bix = engine.Index.build(bases, starts, seed_len=20, chromosome_padding=2000)
    assert bix.info().reserved == 1
    d = str(tmp_path / "saved")
    bix.save(d)
    ridx, r0 = reflib.RefIndex(d), reflib.RefIndex(small_cfg.idx)
    rb = small_cfg.reads["std150"]
    for i in range(0, 300):
        b = rb.read(i)[0]
        for o in (0, 41, 120):
            a, w = ridx.lookup(b[o:o + 20], 256), r0.lookup(b[o:o + 20], 256)
            assert a[0] == w[0] and a[1] == w[1] and np.array_equal(a[2], w[2]) and np.array_equal(a[3], w[3])
    p = reflib.default_params(maxDist=14)
    want, _ = reflib.RefSingleAligner(r0, p).align(small_cfg.reads["noisy150"])
    got, _ = reflib.RefSingleAligner(ridx, p).align(small_cfg.reads["noisy150"])
    assert differing(want, got) == []
    bix.close()


@pytest.mark.parametrize("shared_hist", ["1", "0"])
def test_device_built_index_equals_reference_index(engine, gidx, small_cfg, monkeypatch, shared_hist):
    """snapgpu_index_build: same hit sets in the same (descending) order as the reference-built directory, for every
    seed of a few hundred reads and their reverse complements; and whole-read results identical.  (shared_hist: the two forms
    of the counting pass -- per-block shared histogram, or straight to the global one when there are too many tables.)"""
    monkeypatch.setenv("SNAPGPU_BUILD_SHARED_HIST", shared_hist)
    bases, starts = small_cfg.padded_bases()
    bix = engine.Index.build(bases, starts, seed_len=20, chromosome_padding=2000)
    a, b = gidx.info(), bix.info()
    assert (a.countOfBases, a.seedLen, a.nHashTables, a.overflowTableSize) == (b.countOfBases, b.seedLen, b.nHashTables, b.overflowTableSize)
    rb = small_cfg.reads["std150"]
    seeds = [rb.read(i)[0][o:o + 20] for i in range(400) for o in (0, 33, 77, 130)]
    arr = np.frombuffer(b"".join(seeds), dtype=np.uint8)
    n1, h1, _ = gidx.lookup_seeds(arr, len(seeds), 400)
    n2, h2, _ = bix.lookup_seeds(arr, len(seeds), 400)
    assert np.array_equal(n1, n2)
    for i in range(len(seeds)):
        for d in range(2):
            k = min(int(n1[i, d]), 400)
            assert np.array_equal(h1[i, d, :k], h2[i, d, :k])
    p = engine.default_params(maxDist=14)
    r1, _ = engine.SingleAligner(gidx, p, 4096).align(small_cfg.reads["noisy150"])
    r2, _ = engine.SingleAligner(bix, p, 4096).align(small_cfg.reads["noisy150"])
    assert differing(r1, r2) == []
    bix.close()


@pytest.mark.parametrize("opt", list(OPTION_SETS))
def test_whole_reads_match_reference(engine, gidx, small_cfg, reflib, opt):
    kw = OPTION_SETS[opt]
    al = engine.SingleAligner(gidx, engine.default_params(**kw), 4096)
    ridx = reflib.RefIndex(small_cfg.idx)
    for name, rb in small_cfg.reads.items():
        ral = reflib.RefSingleAligner(ridx, reflib.default_params(**kw))
        want, wctr = ral.align(rb)
        ral.close()
        got, g = al.align(rb)
        assert differing(want, got) == [], (opt, name)
        for k in ("totalReads", "uselessReads", "singleHits", "multiHits", "notFound", "nHashTableLookups",
                  "lvCalls", "affineGapCalls", "nHitsIgnoredBecauseOfTooHighPopularity", "mapqHistogram"):
            assert wctr[k] == g[k], (opt, name, k)
    assert al.launch_count() >= len(small_cfg.reads)
    al.close()


@pytest.mark.parametrize("min_hits", ["2", "8", "300"])
def test_hit_lists_staged_with_bulk_copies_give_the_same_results(engine, gidx, small_cfg, reflib, monkeypatch, min_hits):
    """SNAPGPU_TMA_MIN_HITS: overflow lists of at least that many hits are staged into shared memory with cp.async.bulk + mbarrier
    (sg_warp_stage_hits) instead of being read from HBM word by word -- lists that start at any 4-byte offset, longer than the
    staging buffer (chunked), and every list of a repeat-rich read set.  Records and counters must not change."""
    monkeypatch.setenv("SNAPGPU_TMA_MIN_HITS", min_hits)
    ridx = reflib.RefIndex(small_cfg.idx)
    for opt in ("default_d14", "d8_h20"):
        kw = OPTION_SETS[opt]
        if min_hits == "300":
            kw = dict(kw, maxHits=2000)          # lists longer than the 256-word staging buffer
        al = engine.SingleAligner(gidx, engine.default_params(**kw), 4096)
        for name in ("std150", "noisy150"):
            rb = small_cfg.reads[name]
            want, wctr = reflib.RefSingleAligner(ridx, reflib.default_params(**kw)).align(rb)
            got, g = al.align(rb)
            assert differing(want, got) == [], (opt, name)
            for k in ("nHashTableLookups", "lvCalls", "affineGapCalls", "nHitsIgnoredBecauseOfTooHighPopularity", "mapqHistogram"):
                assert wctr[k] == g[k], (opt, name, k)
        al.close()


def test_hit_lists_longer_than_the_staging_buffer(engine, reflib, tmp_path, monkeypatch):
    """600 copies of a 150 bp unit: every seed of a read from the unit has a 600-entry overflow list, which -h 2000 lets through and
    the bulk-copy staging takes in three chunks (the buffer holds 256 words); compared with the reference and with staging off."""
    from snap_b200 import synth
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, size=150)]
    parts = []
    for i in range(600):
        parts.append(acgt[rng.integers(0, 4, size=int(rng.integers(40, 90)))])
        parts.append(unit)
    contig = np.concatenate(parts + [acgt[rng.integers(0, 4, size=5000)]])
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta(fa, [contig])
    d = str(tmp_path / "idx")
    reflib.build_reference_index(reflib.SNAP_ALIGNER, fa, d)
    reads = synth.make_reads([contig], 300, 100, seed=3)
    kw = dict(maxDist=8, maxHits=2000)
    want, wctr = reflib.RefSingleAligner(reflib.RefIndex(d), reflib.default_params(**kw)).align(reads)
    ix = engine.Index.open(d)
    out = {}
    for tma in ("0", "16"):
        monkeypatch.setenv("SNAPGPU_TMA_MIN_HITS", tma)
        al = engine.SingleAligner(ix, engine.default_params(**kw), 1024)
        got, g = al.align(reads)
        al.close()
        assert differing(want, got) == [], tma
        out[tma] = g
        assert g["lvCalls"] == wctr["lvCalls"] and g["nHashTableLookups"] == wctr["nHashTableLookups"]
    assert out["0"]["nOverflowWordsRead"] == out["16"]["nOverflowWordsRead"] > 300 * 600
    ix.close()


def test_large_index(engine, small_cfg, reflib):
    ix = engine.Index.open(small_cfg.idx_large)
    assert ix.info().largeHashTable == 1
    rb = small_cfg.reads["noisy150"]
    got, _ = engine.SingleAligner(ix, engine.default_params(maxDist=14), 4096).align(rb)
    want, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx_large), reflib.default_params(maxDist=14)).align(rb)
    assert differing(want, got) == []
    ix.close()


def test_golden_e2e_fixture(engine, golden_dir):
    """No reference needed: the index is built on the device from the fixture's genome."""
    from snap_b200 import synth
    g = np.load(os.path.join(golden_dir, "e2e_small.npz"))
    parts, starts, pos = [], [], 0
    for c in (g["contig0"], g["contig1"]):
        parts.append(np.full(2000, ord("n"), dtype=np.uint8)); pos += 2000
        starts.append(pos); parts.append(c); pos += c.size
    parts.append(np.full(2000, ord("n"), dtype=np.uint8))
    ix = engine.Index.build(np.concatenate(parts), starts)
    reads = synth.ReadBatch(g["bases"], g["quals"], g["offsets"], g["lens"])
    for name in ("default_d14", "noag_d14", "ne_d20"):
        got, ctr = engine.SingleAligner(ix, engine.default_params(**OPTION_SETS[name]), 1024).align(reads)
        assert differing(g["res_" + name], got) == [], name
        want_ctr = dict(zip(["totalReads", "uselessReads", "singleHits", "multiHits", "notFound", "nHashTableLookups", "nHashEntriesProbed",
                             "nOverflowWordsRead", "lvCalls", "affineGapCalls", "nHitsIgnoredBecauseOfTooHighPopularity"], g["ctr_" + name]))
        for k in ("singleHits", "multiHits", "notFound", "nHashTableLookups", "lvCalls", "affineGapCalls"):
            assert int(want_ctr[k]) == ctr[k], (name, k)
    ix.close()


def test_edge_reads(engine, gidx, small_cfg, reflib):
    from snap_b200 import synth
    c = small_cfg.contigs[0]
    L = 400
    reads = [
        (b"ACGTACGTAC", b"5" * 10), (c[100:149].tobytes(), b"I" * 49), (c[100:150].tobytes(), b"I" * 50),
        (b"N" * 100, b"5" * 100),
        (bytes(c[1000:1100].tobytes()[:9] + b"N" + c[1010:1100].tobytes()), b"5" * 100),
        (b"".join(c[2000 + 20 * i: 2000 + 20 * i + 19].tobytes() + b"N" for i in range(5)), b"5" * 100),
        (c[5000:5000 + L].tobytes(), b"H" * L),
        (synth.revcomp(c[7000:7150]).tobytes(), b"#" * 150),
        (c[0:150].tobytes(), b"5" * 150), (c[c.size - 150:].tobytes(), b"5" * 150),
        (np.concatenate([c[c.size - 80:], small_cfg.contigs[1][:70]]).tobytes(), b"5" * 150),
    ]
    rb = synth.ReadBatch.from_lists(reads)
    p = dict(maxDist=14)
    want, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params(**p)).align(rb)
    al = engine.SingleAligner(gidx, engine.default_params(**p), 64)
    got, _ = al.align(rb)
    assert differing(want, got) == []
    empty = synth.ReadBatch.from_lists([])
    res, ctr = al.align(empty)
    assert len(res) == 0 and ctr["totalReads"] == 0
    with pytest.raises(engine.SnapGpuError):
        al.align(synth.ReadBatch.from_lists([(b"A" * 401, b"5" * 401)]))     # longer than the configured maximum
    al.close()


def test_properties_at_scale(engine, small_cfg):
    """Size-independent properties on a larger batch (no oracle): (1) results do not depend on batch order or on how the
    batch is split across launches; (2) reads cut from the reference align back to where they came from;
    (3) a read and its reverse complement land on the same location with opposite direction; (4) the MAPQ histogram
    sums to the number of aligned reads."""
    from snap_b200 import synth
    bases, starts = small_cfg.padded_bases()
    ix = engine.Index.build(bases, starts)
    n = 60000
    rb = synth.make_reads(small_cfg.contigs, n, 150, seed=77)
    al = engine.SingleAligner(ix, engine.default_params(maxDist=14), 1 << 16)
    r1, c1 = al.align(rb)
    perm = np.random.default_rng(5).permutation(n)
    shuffled = synth.ReadBatch.from_lists([rb.read(int(i)) for i in perm])
    r2, _ = al.align(shuffled)
    inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
    assert differing(r1, r2[inv]) == []
    small = engine.SingleAligner(ix, engine.default_params(maxDist=14), 7000)
    r3, c3 = small.align(rb)
    assert differing(r1, r3) == [] and c1["lvCalls"] == c3["lvCalls"] and small.launch_count() == 2 * 2 * ((n + 6999) // 7000)      # calls of at most 7000 reads; each is cut into a small first pipeline stage and the rest (align_host_impl), two-pass launch per stage
    aligned = r1["status"] != 0
    assert aligned.mean() > 0.995
    start = np.array(starts)[rb.truth_contig] + rb.truth_pos
    uniq = r1["status"] == 1
    ok = (np.abs(r1["location"][uniq] - start[uniq]) <= 30) & (r1["direction"][uniq] == rb.truth_rc[uniq])
    assert ok.mean() > 0.999
    sub = rb.slice(0, 4000)
    rc = synth.ReadBatch.from_lists([(synth.revcomp(np.frombuffer(b, dtype=np.uint8)).tobytes(), q[::-1]) for b, q in (sub.read(i) for i in range(sub.n))])
    ra, _ = al.align(sub); rb2, _ = al.align(rc)
    both = (ra["status"] == 1) & (rb2["status"] == 1) & (ra["score"] == 0) & (rb2["score"] == 0)
    assert both.sum() > 500
    assert (ra["location"][both] == rb2["location"][both]).all() and (ra["direction"][both] != rb2["direction"][both]).all()
    assert sum(c1["mapqHistogram"]) == int(aligned.sum()) == c1["singleHits"] + c1["multiHits"]
    al.close(); small.close()
    ix.close()


@pytest.mark.parametrize("opt", list(PAIRED_OPTION_SETS))
def test_pairs_match_reference(engine, gidx, small_cfg, reflib, opt):
    """snapgpu_align_paired (one warp per pair) vs ChimericPairedEndAligner(IntersectingPairedEndAligner) of the compiled reference."""
    kw, pkw = PAIRED_OPTION_SETS[opt]
    al = engine.PairedAligner(gidx, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 2048)
    p, pp = reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)
    ridx = reflib.RefIndex(small_cfg.idx)
    for name, pb in small_cfg.pairs.items():
        ral = reflib.RefPairedAligner(ridx, p, pp)
        want, _ = ral.align(pb)
        ral.close()
        got, ctr = al.align(pb)
        assert differing_pairs(want, got) == [], (opt, name)
        assert ctr["totalReads"] == pb.n
    al.close()


def test_pairs_large_index_and_split_batches(engine, small_cfg, reflib):
    kw, pkw = PAIRED_OPTION_SETS["default_d14"]
    ix = engine.Index.open(small_cfg.idx_large)
    pb = small_cfg.pairs["clipped150"]
    want, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx_large), reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)).align(pb)
    big = engine.PairedAligner(ix, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 4096)
    got, c1 = big.align(pb)
    assert differing_pairs(want, got) == []
    small = engine.PairedAligner(ix, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 100)
    got2, c2 = small.align(pb)
    assert differing_pairs(want, got2) == [] and c1["lvCalls"] == c2["lvCalls"] and small.launch_count() == 4 * 2 * ((pb.n // 2 + 99) // 100)      # calls of at most 100 pairs, each in two pipeline stages (a small first one, then the rest); three stages + retry launch per pipeline stage
    big.close(); small.close(); ix.close()


def test_paired_edge_cases(engine, gidx, small_cfg, reflib):
    from snap_b200 import synth
    c = small_cfg.contigs[0]
    rc = synth.revcomp
    def q(n): return b"5" * n
    pairs = [
        (c[1000:1150].tobytes(), rc(c[1300:1450]).tobytes()),              # clean FR pair
        (rc(c[1300:1450]).tobytes(), c[1000:1150].tobytes()),              # same, ends swapped
        (c[1000:1150].tobytes(), c[1300:1450].tobytes()),                  # FF: not a proper pair
        (c[1000:1150].tobytes(), rc(small_cfg.contigs[1][5000:5150]).tobytes()),   # mates on different contigs
        (c[1000:1150].tobytes(), rc(c[9000:9150]).tobytes()),              # too far apart
        (c[1000:1150].tobytes(), b"ACGTACGTAC"),                           # mate shorter than a seed
        (c[1000:1040].tobytes(), rc(c[1300:1450]).tobytes()),              # first end < minReadLength
        (b"N" * 150, rc(c[1300:1450]).tobytes()),                          # all-N end
        (c[1000:1040].tobytes(), c[1300:1330].tobytes()),                  # both ends too short
        (c[0:150].tobytes(), rc(c[250:400]).tobytes()),                    # at the very start of a contig
        (c[c.size - 400:c.size - 250].tobytes(), rc(c[c.size - 150:]).tobytes()),   # at the very end
        (c[1000:1150].tobytes(), rc(c[1020:1170]).tobytes()),              # overlapping mates
        (c[1000:1150].tobytes(), rc(c[1000:1150]).tobytes()),              # identical span
        (c[1000:1090].tobytes() + rc(c[40000:40060]).tobytes(), rc(c[1300:1450]).tobytes()),   # junk tail on one end: soft clip
        (rc(c[40000:40070]).tobytes() + c[1070:1150].tobytes(), rc(c[1300:1450]).tobytes()),   # junk head
        (c[1000:1075].tobytes() + c[60000:60075].tobytes(), rc(c[1300:1375]).tobytes() + c[70000:70075].tobytes()),   # both ends half junk
    ]
    rb = synth.ReadBatch.from_lists([(x, q(len(x))) for pr in pairs for x in pr])
    kw, pkw = PAIRED_OPTION_SETS["default_d14"]
    want, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)).align(rb)
    al = engine.PairedAligner(gidx, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 64)
    got, _ = al.align(rb)
    assert differing_pairs(want, got) == []
    res, ctr = al.align(synth.ReadBatch.from_lists([]))
    assert len(res) == 0 and ctr["totalReads"] == 0
    with pytest.raises(engine.SnapGpuError):
        al.align(synth.ReadBatch.from_lists([(b"A" * 401, b"5" * 401), (b"A" * 100, b"5" * 100)]))
    al.close()


def test_paired_properties_at_scale(engine, small_cfg):
    """No oracle: results independent of batch order; pairs cut from the reference come back as proper pairs at their
    source positions; swapping the ends of every pair swaps the result columns."""
    from snap_b200 import synth
    bases, starts = small_cfg.padded_bases()
    ix = engine.Index.build(bases, starts)
    kw, pkw = PAIRED_OPTION_SETS["default_d14"]
    n = 20000
    pb = synth.make_pairs(small_cfg.contigs, n, 150, seed=91)
    al = engine.PairedAligner(ix, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 1 << 15)
    r1, c1 = al.align(pb)
    perm = np.random.default_rng(6).permutation(n)
    lst = []
    for i in perm:
        lst += [pb.read(2 * int(i)), pb.read(2 * int(i) + 1)]
    r2, _ = al.align(synth.ReadBatch.from_lists(lst))
    inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
    assert differing_pairs(r1, r2[inv]) == []
    assert r1["alignedAsPair"].mean() > 0.97
    proper = (r1["alignedAsPair"] == 1) & (r1["status"][:, 0] != 0) & (r1["status"][:, 1] != 0)
    d = np.abs(r1["location"][proper, 0] - r1["location"][proper, 1])
    assert (d <= 1000).all() and (r1["direction"][proper, 0] != r1["direction"][proper, 1]).all()
    assert c1["totalReads"] == 2 * n and sum(c1["mapqHistogram"]) == int((r1["status"] != 0).sum())
    al.close(); ix.close()


def test_pairs_tiny_pool_caps_take_the_retry_pass(engine, gidx, small_cfg, reflib, monkeypatch):
    """Per-warp candidate pools far too small for most pairs: everything that overflows is re-aligned by the full-size
    retry launch, and the results must still be the reference's."""
    monkeypatch.setenv("SNAPGPU_PAIRED_POOL_CAP", "16")
    monkeypatch.setenv("SNAPGPU_PAIRED_CAND_CAP", "4")
    kw, pkw = PAIRED_OPTION_SETS["default_d14"]
    al = engine.PairedAligner(gidx, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), 2048)
    p, pp = reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)
    for name in ("noisy150", "clipped150"):
        pb = small_cfg.pairs[name]
        want, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx), p, pp).align(pb)
        got, _ = al.align(pb)
        assert differing_pairs(want, got) == [], name
    al.close()


_CTR_KEYS = ("totalReads", "uselessReads", "singleHits", "multiHits", "notFound", "nHashTableLookups", "nHashEntriesProbed", "nOverflowWordsRead",
             "lvCalls", "affineGapCalls", "nHitsIgnoredBecauseOfTooHighPopularity", "mapqHistogram")


@pytest.mark.parametrize("opt", ["default_d14", "ag_d20", "noag_d14", "ne_d20", "stopfirst"])
def test_single_launch_forms_agree(engine, gidx, small_cfg, reflib, monkeypatch, opt):
    """The two-pass launch (first pass without the affine-gap code, deferred reads again from scratch) and the one-launch
    form give the same records AND the same counters; with affine gap off the second pass has nothing to do, under -ne
    (where the library would not choose the two-pass form itself) every read is deferred."""
    p = engine.default_params(**OPTION_SETS[opt])
    rb = small_cfg.reads["noisy150"]
    want, wctr = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params(**OPTION_SETS[opt])).align(rb)
    out = {}
    monkeypatch.setenv("SNAPGPU_OVERLAP_MIN_READS", "1")
    # "ov": the overlapped form of the two-pass launch (both passes resident at once, the second consuming the first's list as it grows)
    for form, two_pass, overlap in (("ov", "1", "1"), ("1", "1", "0"), ("0", "0", "0")):
        monkeypatch.setenv("SNAPGPU_TWO_PASS", two_pass)
        monkeypatch.setenv("SNAPGPU_OVERLAP", overlap)
        al = engine.SingleAligner(gidx, p, 1 << 14)      # (the overlapped form needs the arenas of 8 CTAs per SM: a handle for at least 9472 reads)
        got, ctr = al.align(rb)
        out[form] = (got, ctr, al.launch_count())
        al.close()
        assert differing(want, got) == [], (opt, form)
        for k in ("totalReads", "singleHits", "multiHits", "notFound", "nHashTableLookups", "lvCalls", "affineGapCalls", "mapqHistogram"):
            assert wctr[k] == ctr[k], (opt, form, k)
    for k in _CTR_KEYS:
        assert out["1"][1][k] == out["0"][1][k], (opt, k)
        assert out["ov"][1][k] == out["0"][1][k], (opt, k)
    assert out["0"][2] == 1 and out["1"][2] == 2
    assert out["ov"][2] == 3      # first pass, second pass beside it, the second pass's fourth CTA per SM behind the first


@pytest.mark.parametrize("opt", ["default_d27", "hc_d14", "hc_noag", "hc_forcespacing"])
def test_paired_launch_forms_agree(engine, gidx, small_cfg, reflib, monkeypatch, opt):
    """Staged launch (three stage kernels + retry pass) vs the one-launch form: same records, same counters."""
    kw, pkw = PAIRED_OPTION_SETS[opt]
    p, pp = engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw)
    out = {}
    for form in ("1", "0"):
        monkeypatch.setenv("SNAPGPU_PAIRED_STAGED", form)
        al = engine.PairedAligner(gidx, p, pp, 2048)
        res = [al.align(small_cfg.pairs[name]) for name in ("noisy150", "clipped150")]
        out[form] = res
        al.close()
    for (a, ca), (b, cb) in zip(out["1"], out["0"]):
        assert differing_pairs(a, b) == [], opt
        for k in _CTR_KEYS:
            assert ca[k] == cb[k], (opt, k)


@pytest.mark.parametrize("per_pair,cand_cap", [("1", "512"), ("4", "8")])
def test_paired_handoff_pool_exhaustion_takes_the_retry_pass(engine, gidx, small_cfg, reflib, monkeypatch, per_pair, cand_cap):
    """Stage 1 hands its phase-4 candidates over through a pool of `per_pair` records per pair of the batch.  A batch sized
    1 gives a pool too small for almost any pair with candidates: those pairs must come back, unchanged, from the retry pass
    (which runs the whole function); so must pairs that overflow a small candidate buffer in stage 3."""
    monkeypatch.setenv("SNAPGPU_PAIRED_HANDOFF_PER_PAIR", per_pair)
    monkeypatch.setenv("SNAPGPU_PAIRED_CAND_CAP", cand_cap)
    kw, pkw = PAIRED_OPTION_SETS["default_d14"]
    p, pp = reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)
    pb = small_cfg.pairs["clipped150"]
    want, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx), p, pp).align(pb)
    n_batch = 1 if per_pair == "1" else 2048
    al = engine.PairedAligner(gidx, engine.default_params(**{"numSeedsFromCommandLine": 8, **kw}), engine.default_paired_params(**pkw), n_batch)
    if n_batch == 1:
        from snap_b200 import synth
        sub = pb.slice(0, 120)
        got = np.concatenate([al.align(sub.slice(2 * i, 2 * i + 2))[0] for i in range(sub.n // 2)])
        assert differing_pairs(want[:sub.n // 2], got) == []
    else:
        got, _ = al.align(pb)
        assert differing_pairs(want, got) == []
    al.close()


def test_large_arenas_fall_back_to_fewer_resident_warps(engine, gidx, small_cfg, reflib, monkeypatch):
    """With the longest supported reads the per-warp arenas are ~4x larger; the extra CTAs per SM of the first pass are then
    given up rather than taking most of the HBM.  Results unchanged."""
    monkeypatch.setenv("SNAPGPU_MAX_READ_LEN", "1000")
    p = engine.default_params(maxDist=14)
    al = engine.SingleAligner(gidx, p, 1 << 16)
    rb = small_cfg.reads["noisy150"]
    want, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params(maxDist=14)).align(rb)
    got, _ = al.align(rb)
    assert differing(want, got) == []
    al.close()



def _same_fastq(want, got):
    names = ("bases", "quals", "offsets", "lens", "id_offsets", "id_lens", "front_clipped")
    for n, a, b in zip(names, want[:7], got[:7]):
        assert np.array_equal(np.asarray(a), np.asarray(b)), n
    assert want[7] == got[7], "bytes consumed"


def test_fastq_ingest_matches_reference_reader(engine, small_cfg, reflib, golden_dir):
    """snapgpu_fastq_parse vs FASTQReader::getReadFromBuffer + Read::clip: fixture, then fresh decorated text (CRLF, lower
    case, '.', '#' heads/tails, comments, truncated last record), every clipping mode; and the parsed reads align exactly
    like the same reads handed over directly."""
    from snap_b200 import synth
    g = np.load(os.path.join(golden_dir, "fastq_small.npz"))
    fq = engine.FastqParser(max_bytes=64 << 20, max_reads=200000)
    for clip in (0, 1, 2, 3):
        got = fq.parse(g["text"], clip)
        want = (g["bases%d" % clip], g["quals%d" % clip], g["offsets%d" % clip], g["lens%d" % clip], g["idoff%d" % clip], g["idlen%d" % clip],
                g["front%d" % clip], int(g["used%d" % clip][0]))
        _same_fastq(want, got)
    rb = synth.make_reads(small_cfg.contigs, 30000, 150, seed=81, short_frac=0.1, n_run_frac=0.05)
    text = synth.make_fastq_text(rb, 82, crlf_frac=0.1, lower_frac=0.2, dot_frac=0.1, hash_tail_frac=0.3, hash_head_frac=0.1, comment_frac=0.5,
                                 plus_id_frac=0.2, truncate_last=True)
    for clip in (2, 3, 0):
        _same_fastq(reflib.fastq_parse(text, clip), fq.parse(text, clip))
    # max_reads smaller than the buffer holds: stops after max_reads records, reports where
    small = engine.FastqParser(max_bytes=64 << 20, max_reads=1000)
    got = small.parse(text, 2)
    want = reflib.fastq_parse(text, 2, max_reads=1000)
    _same_fastq(want, got)
    assert len(got[3]) == 1000
    # empty buffer, buffer without a complete record
    assert len(fq.parse(np.zeros(0, dtype=np.uint8), 2)[3]) == 0
    assert fq.parse(np.frombuffer(b"@r\nACGT\n+\n", dtype=np.uint8), 2)[7] == 0
    for bad in (b"@r\nACGT\n\nIIII\n", b"r\nACGT\n+\nIIII\n", b"@r\n1CGT\n+\nIIII\n", b"@r\nACGT\n-\nIIII\n"):
        with pytest.raises(engine.SnapGpuError):
            fq.parse(np.frombuffer(bad, dtype=np.uint8), 2)
    # parsed reads == the reads: align both ways
    plain = synth.make_fastq_text(rb, 83)
    b, q, off, ln, *_ = fq.parse(plain, 2)
    ix = engine.Index.open(small_cfg.idx)
    al = engine.SingleAligner(ix, engine.default_params(maxDist=14), 1 << 15)
    r1, _ = al.align(synth.ReadBatch(b, q, off, ln))
    r2, _ = al.align(rb)
    assert r1.tobytes() == r2.tobytes()
    al.close(); ix.close(); fq.close(); small.close()


def test_pinned_host_buffers_give_the_same_results(engine, gidx, small_cfg):
    """Page-locked inputs / results are DMA'd directly (no staging copies): same records as the pageable path."""
    import torch
    from snap_b200 import synth
    rb = synth.make_reads(small_cfg.contigs, 20000, 150, seed=88)
    al = engine.SingleAligner(gidx, engine.default_params(maxDist=14), 1 << 15)
    want, wc = al.align(rb)
    pb = synth.ReadBatch(torch.from_numpy(rb.bases).pin_memory().numpy(), torch.from_numpy(rb.quals).pin_memory().numpy(), rb.offsets, rb.lens)
    out = torch.empty((rb.n * engine.RESULT_DTYPE.itemsize,), dtype=torch.uint8).pin_memory().numpy().view(engine.RESULT_DTYPE)
    got, gc = al.align(pb, out=out)
    assert got.tobytes() == want.tobytes() and gc == wc
    al.close()


def _reference_sam_lines(reflib, argv):
    import subprocess
    r = subprocess.run([reflib.SNAP_ALIGNER] + argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    return [l for l in open(argv[argv.index("-o") + 1], "rb").read().split(b"\n") if l and not l.startswith(b"@")]


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("indel100", ["-="]), ("noisy150", ["-G-"])])
def test_sam_records_from_the_device_equal_reference_binary(engine, gidx, small_cfg, reflib, tmp_path, name, extra):
    """Output stage on the device (snapgpu_sam_format_single over the engine's own result records) vs the SAM file
    `snap-aligner single ... -o out.sam -t 1` writes for the same reads: record for record, byte for byte."""
    rb = small_cfg.reads[name]
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.sam")
    rb.write_fastq(fq)
    want = _reference_sam_lines(reflib, ["single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"] + extra)
    p = engine.default_params(maxDist=14, useAffineGap=0 if "-G-" in extra else 1)
    al = engine.SingleAligner(gidx, p, 4096)
    res, _ = al.align(rb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096, use_m=("-=" not in extra))
    got = [l for l in fmt.format(rb, [b"r%d" % i for i in range(rb.n)], res).split(b"\n") if l]
    fmt.close()
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("clipped150", ["-="])])
def test_sam_pair_records_from_the_device_equal_reference_binary(engine, gidx, small_cfg, reflib, tmp_path, name, extra):
    """The same for pairs (snapgpu_sam_format_paired): write order, mate fields, template length, QS."""
    pb = small_cfg.pairs[name]
    f1 = str(tmp_path / "p1.fq"); f2 = str(tmp_path / "p2.fq"); out = str(tmp_path / "o.sam")
    ids = []
    with open(f1, "wb") as a, open(f2, "wb") as b:
        for i in range(pb.n // 2):
            x, q = pb.read(2 * i); a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, x, q))
            x, q = pb.read(2 * i + 1); b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, x, q))
            ids += [b"p%d/1" % i, b"p%d/2" % i]
    want = _reference_sam_lines(reflib, ["paired", small_cfg.idx, f1, f2, "-o", out, "-t", "1"] + extra)
    p, pp = engine.default_params(maxDist=27, numSeedsFromCommandLine=8), engine.default_paired_params()
    al = engine.PairedAligner(gidx, p, pp, 2048)
    res, _ = al.align(pb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096, use_m=("-=" not in extra))
    got = [l for l in fmt.format(pb, ids, res, paired=True).split(b"\n") if l]
    fmt.close()
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


def test_fastq_text_to_sam_text_without_leaving_the_device(engine, gidx, small_cfg, reflib, tmp_path):
    """The whole chain resident in HBM: FASTQ text -> snapgpu_fastq_parse_device -> snapgpu_align_single_device ->
    snapgpu_sam_format_single_device (ids taken from the FASTQ text itself, records from the aligner's device buffer, packed on the device);
    the text that comes out is the record set of the SAM file the reference binary writes for that FASTQ file."""
    import torch
    from snap_b200 import synth
    rb = small_cfg.reads["indel100"]
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.sam")
    rb.write_fastq(fq)
    want = _reference_sam_lines(reflib, ["single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"])
    text = np.frombuffer(open(fq, "rb").read(), dtype=np.uint8)
    dev = torch.device("cuda", 0)
    n = rb.n
    d_text = torch.from_numpy(text.copy()).to(dev)
    d_b = torch.empty((text.size // 2 + 64,), dtype=torch.uint8, device=dev); d_q = torch.empty_like(d_b)
    d_off = torch.empty((n + 8,), dtype=torch.int64, device=dev); d_len = torch.empty((n + 8,), dtype=torch.int32, device=dev)
    d_ido = torch.empty((n + 8,), dtype=torch.int64, device=dev); d_idl = torch.empty((n + 8,), dtype=torch.int32, device=dev)
    d_fc = torch.empty((n + 8,), dtype=torch.int32, device=dev)
    d_res = torch.empty((n, engine.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    cap = n * 1024
    d_sam = torch.empty((cap,), dtype=torch.uint8, device=dev)
    p = engine.default_params(maxDist=14)
    fqp = engine.FastqParser(max_bytes=int(text.size) + 64, max_reads=n + 8)
    al = engine.SingleAligner(gidx, p, n)
    fmt = engine.SamFormatter(gidx, p, n)
    st = torch.cuda.Stream(dev)
    nr, used = fqp.parse_device(d_text.data_ptr(), int(text.size), 2, d_b.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_ido.data_ptr(),
                                d_idl.data_ptr(), d_fc.data_ptr(), st.cuda_stream)
    assert nr == n and used == text.size
    al.align_device(nr, d_b.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_res.data_ptr(), 0, st.cuda_stream)
    nbytes = fmt.format_device(nr, 100, d_b.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_text.data_ptr(), d_ido.data_ptr(), d_idl.data_ptr(),
                               d_res.data_ptr(), d_sam.data_ptr(), cap, stream=st.cuda_stream)
    got = [l for l in d_sam[:nbytes].cpu().numpy().tobytes().split(b"\n") if l]
    fmt.close(); al.close(); fqp.close()
    assert len(want) == len(got) == n
    bad = [i for i in range(n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])
    # a text buffer that is too small is reported, not overrun
    fmt2 = engine.SamFormatter(gidx, p, n)
    with pytest.raises(engine.SnapGpuError):
        fmt2.format_device(nr, 100, d_b.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_text.data_ptr(), d_ido.data_ptr(), d_idl.data_ptr(),
                           d_res.data_ptr(), d_sam.data_ptr(), 1000, stream=st.cuda_stream)
    fmt2.close()


def _split_bam(blob):
    import struct
    out, p = [], 0
    while p < len(blob):
        b = struct.unpack("<i", blob[p:p + 4])[0]
        out.append(blob[p:p + 4 + b]); p += 4 + b
    return out


def _reference_bam_records(reflib, argv, out):
    """The alignment records of the BAM file the reference binary writes (BGZF inflated, header and reference table skipped)."""
    import gzip, struct, subprocess
    r = subprocess.run([reflib.SNAP_ALIGNER] + argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    raw = gzip.open(out, "rb").read()
    assert raw[:4] == b"BAM\x01"
    p = 8 + struct.unpack("<i", raw[4:8])[0]
    n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
    for _ in range(n_ref):
        l = struct.unpack("<i", raw[p:p + 4])[0]; p += 8 + l
    return _split_bam(raw[p:])


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("indel100", ["-="]), ("noisy150", ["-G-"])])
def test_bam_records_from_the_device_equal_reference_binary(engine, gidx, small_cfg, reflib, tmp_path, name, extra):
    """SNAPGPU_FORMAT_BAM: the uncompressed BAM alignment records formatted on the device vs the records of the .bam file
    `snap-aligner single ... -o out.bam -t 1` writes (its BGZF blocks inflated): every record, byte for byte."""
    rb = small_cfg.reads[name]
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.bam")
    rb.write_fastq(fq)
    want = _reference_bam_records(reflib, ["single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"] + extra, out)
    p = engine.default_params(maxDist=14, useAffineGap=0 if "-G-" in extra else 1)
    al = engine.SingleAligner(gidx, p, 4096)
    res, _ = al.align(rb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096, use_m=("-=" not in extra))
    fmt.set_format(bam=True)
    got = _split_bam(fmt.format(rb, [b"r%d" % i for i in range(rb.n)], res))
    fmt.close()
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


def test_bam_pair_records_from_the_device_and_bgzf(engine, gidx, small_cfg, reflib, tmp_path):
    """Pairs as BAM records on the device (mate fields, bins, template lengths, QS) vs the reference binary's .bam; then the record stream
    wrapped into BGZF members on the device (snapgpu_bgzf_device: stored deflate blocks + CRC-32) inflates, with the standard gzip reader, to
    exactly that stream."""
    import gzip, io
    import torch
    pb = small_cfg.pairs["noisy150"]
    f1 = str(tmp_path / "p1.fq"); f2 = str(tmp_path / "p2.fq"); out = str(tmp_path / "o.bam")
    ids = []
    with open(f1, "wb") as a, open(f2, "wb") as b:
        for i in range(pb.n // 2):
            x, q = pb.read(2 * i); a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, x, q))
            x, q = pb.read(2 * i + 1); b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, x, q))
            ids += [b"p%d/1" % i, b"p%d/2" % i]
    want = _reference_bam_records(reflib, ["paired", small_cfg.idx, f1, f2, "-o", out, "-t", "1"], out)
    p, pp = engine.default_params(maxDist=27, numSeedsFromCommandLine=8), engine.default_paired_params()
    al = engine.PairedAligner(gidx, p, pp, 2048)
    res, _ = al.align(pb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096)
    fmt.set_format(bam=True)
    blob = fmt.format(pb, ids, res, paired=True)
    fmt.close()
    got = _split_bam(blob)
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])
    # BGZF on the device: several members (the stream is > 65280 bytes), the last one short
    dev = torch.device("cuda", 0)
    payload = np.frombuffer(blob, dtype=np.uint8)
    assert payload.size > 2 * 65280
    d_in = torch.from_numpy(payload.copy()).to(dev)
    cap = payload.size + 31 * (payload.size // 65280 + 2)
    d_out = torch.empty((cap,), dtype=torch.uint8, device=dev)
    used = engine.bgzf_device(d_in.data_ptr(), payload.size, d_out.data_ptr(), cap)
    torch.cuda.synchronize()
    z = d_out[:used].cpu().numpy().tobytes()
    assert used == payload.size + 31 * ((payload.size + 65279) // 65280)
    assert z[:4] == b"\x1f\x8b\x08\x04" and z[12:14] == b"BC"
    assert gzip.GzipFile(fileobj=io.BytesIO(z)).read() == blob
    with pytest.raises(engine.SnapGpuError):
        engine.bgzf_device(d_in.data_ptr(), payload.size, d_out.data_ptr(), payload.size)


def test_sam_records_of_quality_clipped_reads_from_the_device(engine, gidx, small_cfg, reflib, tmp_path):
    """Reads with '#' quality tails: the aligner gets the clipped view (Read::clip, default -C-+), snapgpu_sam_format_single the whole
    read plus the clip counts; the records (whole read, S operations) must be the reference binary's."""
    from snap_b200 import synth
    rng = np.random.default_rng(5)
    rb = small_cfg.reads["noisy150"]
    reads = []; fronts = []; clens = []
    for i in range(600):
        b, q = rb.read(i)
        q = bytearray(q)
        tail = int(rng.choice([0, 0, 3, 11, 40])) if len(b) >= 70 else 0
        for k in range(tail):
            q[len(q) - 1 - k] = ord("#")
        reads.append((b, bytes(q))); fronts.append(0); clens.append(len(b) - tail)
    full = synth.ReadBatch.from_lists(reads)
    clipped = synth.ReadBatch.from_lists([(b[:n], q[:n]) for (b, q), n in zip(reads, clens)])
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.sam")
    full.write_fastq(fq)
    want = _reference_sam_lines(reflib, ["single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"])
    p = engine.default_params(maxDist=14)
    al = engine.SingleAligner(gidx, p, 4096)
    res, _ = al.align(clipped)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096)
    got = [l for l in fmt.format(full, [b"r%d" % i for i in range(full.n)], res, front_clipped=fronts, clipped_lens=clens).split(b"\n") if l]
    fmt.close()
    assert len(want) == len(got) == full.n
    bad = [i for i in range(full.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


def _sorted_on_device(engine, fmt, n_records, cap):
    import torch
    dev = torch.device("cuda", 0)
    d_sorted = torch.empty((cap,), dtype=torch.uint8, device=dev)
    d_keys = torch.empty((n_records,), dtype=torch.int64, device=dev)
    d_offs = torch.empty((n_records,), dtype=torch.int64, device=dev)
    used = fmt.sort_device(0, d_sorted.data_ptr(), cap, d_keys.data_ptr(), d_offs.data_ptr())
    keys = d_keys.cpu().numpy().view(np.uint64)
    offs = d_offs.cpu().numpy()
    return d_sorted[:used].cpu().numpy().tobytes(), keys, offs


def test_records_sorted_on_the_device_equal_reference_sorted_output(engine, gidx, small_cfg, reflib, tmp_path):
    """Row N4, the sort (snapgpu_sam_sort_device = SortedDataFilter's stable sort of a write batch, SortedDataWriter.cpp:905-1010): the records of
    `snap-aligner single ... -so -S d -o sorted.sam -t 1` (one write batch, so its merge has one run), in the file's order -- unaligned
    records last, equal keys in input order.  Then the same with BAM records against the sorted .bam."""
    rb = small_cfg.reads["noisy150"]            # 212 of 1500 reads stay unaligned
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "sorted.sam")
    rb.write_fastq(fq)
    want = _reference_sam_lines(reflib, ["single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14", "-so", "-S", "d"])
    p = engine.default_params(maxDist=14)
    al = engine.SingleAligner(gidx, p, 2048)
    res, _ = al.align(rb)
    al.close()
    ids = [b"r%d" % i for i in range(rb.n)]
    fmt = engine.SamFormatter(gidx, p, 2048)
    unsorted = [l for l in fmt.format(rb, ids, res).split(b"\n") if l]
    assert fmt.last_record_count() == rb.n
    blob, keys, offs = _sorted_on_device(engine, fmt, rb.n, rb.n * 1024)
    got = [l for l in blob.split(b"\n") if l]
    assert sorted(got) == sorted(unsorted) and got != unsorted
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]][:120], got[bad[0]][:120])
    assert (np.diff(keys.astype(np.float64)) >= 0).all() and (keys >> np.uint64(32) == np.uint64(0xffffffff)).sum() == sum(1 for l in got if l.split(b"\t")[2] == b"*")
    assert offs[0] == 0 and all(blob[o - 1:o] == b"\n" for o in offs[1:])
    # BAM records of the same batch against the reference's sorted BAM
    outb = str(tmp_path / "sorted.bam")
    wantb = _reference_bam_records(reflib, ["single", small_cfg.idx, fq, "-o", outb, "-t", "1", "-d", "14", "-so", "-S", "di"], outb)
    fmt.set_format(bam=True)
    fmt.format(rb, ids, res)
    blob, keys, offs = _sorted_on_device(engine, fmt, rb.n, rb.n * 1024)
    fmt.close()
    gotb = _split_bam(blob)
    assert len(wantb) == len(gotb) == rb.n
    bad = [i for i in range(rb.n) if wantb[i] != gotb[i]]
    assert bad == [], (len(bad), wantb[bad[0]][:60], gotb[bad[0]][:60])


def test_pair_records_sorted_on_the_device_equal_reference_sorted_output(engine, gidx, small_cfg, reflib, tmp_path):
    """The same for pairs: every record is filed under its own final location, an unaligned end under its mate's (ReadWriter.cpp:601-606)."""
    pb = small_cfg.pairs["std150"]               # chimeric mates, unaligned ends
    f1 = str(tmp_path / "p1.fq"); f2 = str(tmp_path / "p2.fq"); out = str(tmp_path / "sorted.sam")
    ids = []
    with open(f1, "wb") as a, open(f2, "wb") as b:
        for i in range(pb.n // 2):
            x, q = pb.read(2 * i); a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, x, q))
            x, q = pb.read(2 * i + 1); b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, x, q))
            ids += [b"p%d/1" % i, b"p%d/2" % i]
    want = _reference_sam_lines(reflib, ["paired", small_cfg.idx, f1, f2, "-o", out, "-t", "1", "-so", "-S", "d"])
    p, pp = engine.default_params(maxDist=27, numSeedsFromCommandLine=8), engine.default_paired_params()
    al = engine.PairedAligner(gidx, p, pp, 2048)
    res, _ = al.align(pb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096)
    unsorted = [l for l in fmt.format(pb, ids, res, paired=True).split(b"\n") if l]
    assert fmt.last_record_count() == pb.n
    blob, keys, offs = _sorted_on_device(engine, fmt, pb.n, pb.n * 1024)
    fmt.close()
    got = [l for l in blob.split(b"\n") if l]
    assert sorted(got) == sorted(unsorted)
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]][:120], got[bad[0]][:120])


def test_secondary_alignments_match_reference(engine, gidx, small_cfg, reflib):
    """`snap single -om / -omax / -mpc` through snapgpu_align_single_secondary: primary results, counts and every secondary record in the
    reference's buffer order; a raw buffer that starts too small (SNAPGPU_SECONDARY_RAW_CAP=4) is grown by the call itself; a capacity that
    is too small reports minus the count; the plain entry point refuses an -om handle."""
    from test_hostsim_parity import SECONDARY_SETS, differing_secondary
    ridx = reflib.RefIndex(small_cfg.idx)
    old = os.environ.get("SNAPGPU_SECONDARY_RAW_CAP")
    for opt, (kw, omax, mpc) in SECONDARY_SETS.items():
        if opt == "om3_esd3":
            os.environ["SNAPGPU_SECONDARY_RAW_CAP"] = "4"
        try:
            al = engine.SingleAligner(gidx, engine.default_params(**kw), 4096)
            for name, rb in small_cfg.reads.items():
                ral = reflib.RefSecondaryAligner(ridx, reflib.default_params(**kw), omax, mpc)
                want, wsec, wn, wctr = ral.align(rb, capacity=256)
                ral.close()
                got, gsec, gn, g = al.align_secondary(rb, capacity=256, max_secondary=omax, max_per_contig=mpc)
                bad_p, bad_s = differing_secondary(want, wsec, wn, got, gsec, gn)
                assert bad_p == [] and bad_s == [], (opt, name, bad_p[:5], bad_s[:5])
                for k in ("totalReads", "uselessReads", "singleHits", "multiHits", "notFound", "mapqHistogram"):
                    assert wctr[k] == g[k], (opt, name, k)
            if opt == "om3_esd3":
                rb = small_cfg.reads["std150"]
                _, _, n_big, _ = al.align_secondary(rb, capacity=256, max_secondary=omax, max_per_contig=mpc)
                _, _, n_one, _ = al.align_secondary(rb, capacity=1, max_secondary=omax, max_per_contig=mpc)
                assert (n_one < -1).any() and np.array_equal(np.where(n_big > 1, -n_big, n_big), n_one)
                with pytest.raises(engine.SnapGpuError):
                    al.align(rb)
            al.close()
        finally:
            if old is None:
                os.environ.pop("SNAPGPU_SECONDARY_RAW_CAP", None)
            else:
                os.environ["SNAPGPU_SECONDARY_RAW_CAP"] = old
    # a handle without -om refuses the secondary call
    al = engine.SingleAligner(gidx, engine.default_params(maxDist=14), 4096)
    with pytest.raises(engine.SnapGpuError):
        al.align_secondary(small_cfg.reads["std150"])
    al.close()
