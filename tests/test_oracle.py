"""The oracle pinned against the reference's own golden vectors, and the plain-C restatement
(oracle/port) pinned against the compiled reference (oracle/_ref)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import jobs as J

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def vectors(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_unit_vectors.json")))


@pytest.fixture(scope="module")
def port():
    so = os.path.join(ROOT, "oracle", "liboracle_port.so")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    L = C.CDLL(so)
    L.port_hash.restype = C.c_uint64
    L.port_hash.argtypes = [C.c_uint64]
    L.port_mapq.restype = C.c_int
    L.port_mapq.argtypes = [C.c_double, C.c_double, C.c_int]
    L.port_lv.restype = C.c_int
    L.port_lv.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    L.port_lv_batch.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
    L.port_seed_pack.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p]
    L.port_probe.restype = C.c_int64
    L.port_probe.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
    return L


def _lv_score(fn, v):
    q = b"5" * (v["patternLen"] + 8)
    mp = C.c_double(); a = C.c_int(); b = C.c_int(); c = C.c_int()
    return fn(1, v["text"].encode() + b"\0" * 16, v["textLen"], v["pattern"].encode() + b"\0" * 16, q, v["patternLen"], v["k"],
              C.byref(mp), C.byref(a), C.byref(b), C.byref(c))


def test_reference_lv_known_answers(vectors, reflib, port):
    """tests/LandauVishkinTest.cpp:11-32: the 11 computeEditDistance vectors, on the compiled reference and the port."""
    L = reflib.lib()
    L.ref_lv.restype = C.c_int
    L.ref_lv.argtypes = port.port_lv.argtypes
    for v in vectors["lv"]:
        assert _lv_score(L.ref_lv, v) == v["expected"], v
        assert _lv_score(port.port_lv, v) == v["expected"], v


def test_reference_ag_known_answers(vectors, reflib):
    """tests/AffineGapVectorizedTest.cpp:40-67: the 10 computeScore vectors (params 1,4,6,1,10,5; quality all '2')."""
    for v in vectors["ag"]:
        text = np.frombuffer(b"n" * 64 + v["text"].encode() + b"n" * 64, dtype=np.uint8)
        pat = np.frombuffer(v["pattern"].encode() + b"A" * 8, dtype=np.uint8)
        qual = np.full(pat.size, ord("2"), dtype=np.uint8)
        job = np.zeros(1, dtype=reflib.AG_JOB_DTYPE)
        job[0] = (64, 0, v["textLen"], v["patternLen"], v["w"], v["scoreInit"], 1, 0, 0, 0)
        out = reflib.ag_batch(text, pat, qual, job, params=vectors["ag_params"])
        assert int(out[0]["agScore"]) == v["expected"], v


def test_reference_datatest_end_to_end(vectors, reflib, tmp_path):
    """tests/datatest: `snap index` + `snap single` on the 202 bp reference; SAM columns flag/pos/mapq/cigar of
    correct-fq-datatest.sam:4-5 (columns 1-11 are still what 2.0.5 produces)."""
    from snap_b200 import synth
    d = vectors["datatest"]
    contig = np.frombuffer(d["contig"].encode(), dtype=np.uint8)
    fasta = str(tmp_path / "ref.fa")
    with open(fasta, "w") as f:
        f.write(">%s\n%s\n" % (d["contig_name"], d["contig"]))
    reflib.build_reference_index(reflib.SNAP_ALIGNER, fasta, str(tmp_path / "idx"))
    reads = synth.ReadBatch.from_lists([(r["bases"].encode(), r["quals"].encode()) for r in d["reads"]])
    # the FASTQ reader clips trailing '#' qualities (FASTQ.cpp:294); these reads have none
    idx = reflib.RefIndex(str(tmp_path / "idx"))
    al = reflib.RefSingleAligner(idx, reflib.default_params())
    res, _ = al.align(reads)
    for r, e in zip(res, d["expected"]):
        assert r["status"] == 1 and r["direction"] == (1 if e["flag"] & 16 else 0)
        assert int(r["location"]) - 2000 + 1 == e["pos"]          # 2000 bases of padding precede the contig
        assert int(r["mapq"]) == e["mapq"]
        assert e["cigar"] == "%d=" % len(d["reads"][0]["bases"]) and r["score"] == 0
    assert contig.size == 202


def test_port_tables_and_leaves_match_reference(reflib, port):
    rp, ri, rf = reflib.tables(1100, 1001)
    pp, pi, pf = np.zeros(256), np.zeros(1100), np.zeros(1001)
    port.port_tables(pp.ctypes.data_as(C.c_void_p), pi.ctypes.data_as(C.c_void_p), 1100, pf.ctypes.data_as(C.c_void_p), 1001)
    for a, b in ((rp, pp), (ri, pi), (rf, pf)):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    for sl in range(8, 33):
        offs = (C.c_uint * 64)()
        port.port_seed_sequencer(sl, offs)
        assert [offs[i] for i in range(sl)] == [reflib.lib().ref_wrapped_seed(sl, w) for w in range(sl)]
    rng = np.random.default_rng(3)
    for _ in range(20000):
        pb = rng.random()
        pa = pb + rng.random() * 10.0 ** int(rng.integers(-14, 1))
        pop = int(rng.integers(0, 30))
        assert port.port_mapq(pa, pb, pop) == reflib.lib().ref_mapq(pa, pb, 0, pop)


def test_port_lv_matches_reference_on_random_jobs(reflib, port):
    t, p, q, jb = J.lv_jobs(3000, 77)
    want = reflib.lv_batch(t, p, q, jb.astype(reflib.LV_JOB_DTYPE))
    got = np.zeros(jb.size, dtype=J.LV_OUT)
    port.port_lv_batch(t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
                       jb.ctypes.data_as(C.c_void_p), jb.size, got.ctypes.data_as(C.c_void_p))
    assert J.same_out(want, got).all()


def test_port_probe_and_seed_match_reference(reflib, port, small_cfg):
    """port_seed_pack + port_probe over the raw GenomeIndexHash tables reproduce lookupSeed32's hit counts."""
    import struct
    raw = open(os.path.join(small_cfg.idx, "GenomeIndexHash"), "rb").read()
    tables, off = [], 0
    while off < len(raw):
        magic, tsz, used, ks, vs, vc, inval = struct.unpack_from("<IQQIIII", raw, off)
        assert magic == 0xb111b010 and ks == 4 and vs == 4 and vc == 1
        off += 36
        tables.append(np.frombuffer(raw, dtype=np.uint32, count=tsz * 2, offset=off).copy())
        off += tsz * 8
    idx = reflib.RefIndex(small_cfg.idx)
    n_bases = sum(c.size for c in small_cfg.contigs) + 2000 * (len(small_cfg.contigs) + 1)
    rb = small_cfg.reads["std150"]
    for i in range(300):
        seed = rb.read(i)[0][10:30]
        b = C.c_uint64(); rc = C.c_uint64()
        assert port.port_seed_pack(seed, 20, C.byref(b), C.byref(rc)) == 1
        want = idx.lookup(seed)
        total = 0
        for d, s in enumerate((b.value, rc.value)):
            t = tables[s >> 32]
            ex = C.c_uint()
            slot = port.port_probe(t.ctypes.data_as(C.c_void_p), t.size // 2, s & 0xffffffff, 0xffffffff, C.byref(ex))
            total += ex.value
            if slot < 0:
                assert want[d] == 0
            elif t[2 * slot] < n_bases:
                assert want[d] == 1 and want[2 + d][0] == t[2 * slot]
            else:
                assert want[d] >= 2
        assert total == want[4] + 2      # entries examined: first slot of each chain + the reference's extra-probe counter


def test_fastq_reader_known_answers(reflib):
    """The oracle's FASTQ path (the reference's own FASTQReader::getReadFromBuffer + Read::clip) on hand-checked records."""
    txt = (b"@r1 comment\nACGTNacgt.\n+\nIIII#IIII#\n"        # lower case and '.', one '#' at the back
           b"@r2\r\nGGGG\r\n+r2\r\n####\r\n"                   # CRLF; everything clipped by ClipBack
           b"@r3\nGGAC\n+\n##II\n"                              # '#' qualities in front
           b"@r4\nAC\n+\nI")                                    # incomplete: not consumed
    t = np.frombuffer(txt, dtype=np.uint8)
    b, q, off, ln, ido, idl, fc, used = reflib.fastq_parse(t, 2)
    assert used == len(txt) - len(b"@r4\nAC\n+\nI") and list(ln) == [9, 0, 4]
    assert bytes(b[:9]) == b"ACGTNACGT" and bytes(q[:9]) == b"IIII#IIII"
    assert [bytes(t[int(o):int(o) + int(l)]) for o, l in zip(ido, idl)] == [b"r1", b"r2", b"r3"]
    b, q, off, ln, ido, idl, fc, used = reflib.fastq_parse(t, 3)
    assert list(ln) == [9, 0, 2] and list(fc) == [0, 0, 2] and bytes(b[9:11]) == b"AC"
    b, q, off, ln, ido, idl, fc, used = reflib.fastq_parse(t, 0)
    assert list(ln) == [10, 4, 4] and bytes(b[:10]) == b"ACGTNACGTN"


def test_fastq_golden_fixture_matches_oracle(reflib, golden_dir):
    g = np.load(os.path.join(golden_dir, "fastq_small.npz"))
    for clip in (0, 1, 2, 3):
        b, q, off, ln, ido, idl, fc, used = reflib.fastq_parse(g["text"], clip)
        assert np.array_equal(b, g["bases%d" % clip]) and np.array_equal(q, g["quals%d" % clip]) and np.array_equal(ln, g["lens%d" % clip])
        assert np.array_equal(off, g["offsets%d" % clip]) and np.array_equal(ido, g["idoff%d" % clip]) and np.array_equal(idl, g["idlen%d" % clip])
        assert np.array_equal(fc, g["front%d" % clip]) and used == int(g["used%d" % clip][0])
