"""Output stage, first piece: LandauVishkinWithCigar restated in snap_b200/csrc/sg_lv_cigar.h.

Host-side only so far (the algorithm header built for the host, like the other sg_*.h headers): (1) the compiled reference and
the restatement both reproduce the 30 CIGAR known answers of the reference's own tests (tests/LandauVishkinTest.cpp:34-129);
(2) the restatement equals the compiled reference on random problems: operations, edit distance, text used, net indel,
and the front-clipping verdict of computeEditDistanceNormalized."""
import numpy as np
import pytest

import hostsim_lib as hs

# `snap paired` stock options (-d 27, soft clipping on) and `-hc` (soft clipping off: end bonuses 5/5, minAGScoreImprovement 15)
PAIRED_DEFAULT = (dict(maxDist=27), dict())
PAIRED_HC = (dict(maxDist=27, fivePrimeEndBonus=5, threePrimeEndBonus=5), dict(useSoftClipping=0, minAGScoreImprovement=15))

def _fixture(golden_dir):
    import gzip, json, os
    return json.load(gzip.open(os.path.join(golden_dir, "output_small.json.gz"), "rt"))


def _known_jobs(reflib, known):
    # C string literals: NUL after the last character, as in the reference's test (the routine looks one character past the ends)
    text = bytearray(); pat = bytearray(); jobs = []
    for v in known:
        jobs.append((len(text), len(pat), v["textLen"], v["patternLen"], v["k"], 1 if v["useM"] else 0))
        text += v["text"].encode() + b"\0" * 16
        pat += v["pattern"].encode() + b"\0" * 16
    return (np.frombuffer(bytes(text), dtype=np.uint8).copy(), np.frombuffer(bytes(pat), dtype=np.uint8).copy(),
            np.array(jobs, dtype=reflib.LVC_JOB_DTYPE))


@pytest.mark.parametrize("impl", ["reference", "restatement"])
def test_known_cigars(reflib, golden_dir, impl):
    """The 30 CIGAR known answers of the reference's own unit test (tests/golden/output_small.json.gz, parsed out of
    tests/LandauVishkinTest.cpp:34-129 by make_golden.py) through the compiled reference and through the restatement."""
    known = _fixture(golden_dir)["lv_cigar"]
    assert len(known) == 30
    text, pat, jobs = _known_jobs(reflib, known)
    out = reflib.lv_cigar_batch(text, pat, jobs) if impl == "reference" else hs.lv_cigar_batch(text, pat, jobs, reflib.LVC_OUT_DTYPE)
    for i, v in enumerate(known):
        assert out[i]["score"] >= 0, v
        assert reflib.decode_cigar(out[i]["ops"], int(out[i]["nOps"])) == v["cigar"], v


def test_committed_sam_and_bam_records(reflib, golden_dir, tmp_path):
    """Committed fixture: the SAM and BAM records the reference binary wrote for the reads of e2e_small.npz (made by
    tests/golden/make_golden.py output) vs sg_sam.h / sg_bam.h over that fixture's committed result records."""
    import os, struct
    from snap_b200 import synth
    fx = _fixture(golden_dir)
    g = np.load(os.path.join(golden_dir, "e2e_small.npz"))
    synth.write_fasta(str(tmp_path / "ref.fa"), [g["contig0"], g["contig1"]])
    reflib.build_reference_index(reflib.SNAP_ALIGNER, str(tmp_path / "ref.fa"), str(tmp_path / "idx"))
    reads = synth.ReadBatch(g["bases"], g["quals"], g["offsets"], g["lens"])
    hidx = hs.HsIndex(str(tmp_path / "idx"))
    ids = [b"r%d" % i for i in range(reads.n)]
    res = g["res_default_d14"]
    got = [l for l in hs.sam_single(hidx, reads, ids, res).decode().split("\n") if l]
    assert got == fx["sam"]
    blob = hs.bam_single(hidx, reads, ids, res)
    recs = []; p = 0
    while p < len(blob):
        b = struct.unpack("<i", blob[p:p + 4])[0]
        recs.append(blob[p:p + 4 + b].hex()); p += 4 + b
    assert recs == fx["bam"]


def _fuzz_jobs(reflib, n, seed):
    """Patterns cut from random text with substitutions / insertions / deletions (few or many, so that some exceed k), text
    windows with >= k characters of slack and a shifted start now and then (leading deletions / insertions)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = alphabet[rng.integers(0, 4, size=400000)].copy()
    pats = []; jobs = []; pat_off = 0
    for _ in range(n):
        plen = int(rng.integers(20, 260))
        start = int(rng.integers(200, text.size - 600))
        src = list(text[start:start + plen + 40])
        n_edit = int(rng.choice([0, 1, 1, 2, 2, 3, 4, 6, 9, 14, 30]))
        out = []; si = 0
        edits = set(int(x) for x in rng.integers(0, plen, size=n_edit))
        while len(out) < plen:
            pos = len(out)
            if pos in edits:
                kind = int(rng.integers(0, 4))
                if kind <= 1:
                    out.append(int(alphabet[(int(np.searchsorted(alphabet, src[si])) + 1 + int(rng.integers(0, 3))) % 4])); si += 1
                elif kind == 2:
                    out.append(int(alphabet[rng.integers(0, 4)]))             # insertion
                else:
                    si += int(rng.integers(1, 4)); out.append(src[si]); si += 1   # deletion
                edits.discard(pos)
            else:
                out.append(src[si]); si += 1
        p = np.array(out[:plen], dtype=np.uint8)
        shift = int(rng.choice([0, 0, 0, 0, 1, 2, -1, -2, 5]))
        k = int(rng.choice([3, 8, 15, 28, 60, 126]))
        pats.append(p); pats.append(np.zeros(16, dtype=np.uint8))
        jobs.append((start + shift, pat_off, plen + 127, plen, k, int(rng.integers(0, 2))))
        pat_off += plen + 16
    return text, np.concatenate(pats), np.array(jobs, dtype=reflib.LVC_JOB_DTYPE)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_equals_reference(reflib, seed):
    text, pat, jobs = _fuzz_jobs(reflib, 4000, seed)
    want = reflib.lv_cigar_batch(text, pat, jobs)
    got = hs.lv_cigar_batch(text, pat, jobs, reflib.LVC_OUT_DTYPE)
    for f in ("score", "nOps", "textUsed", "netIndel", "normalizedScore", "addFrontClipping"):
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, (f, int(bad[0]), want[int(bad[0])], got[int(bad[0])])
    assert (want["ops"] == got["ops"]).all()
    solved = want["score"] >= 0
    assert 0.5 < solved.mean() < 1.0                      # both outcomes occur
    assert (want["netIndel"] != 0).sum() > 100 and (want["addFrontClipping"] != 0).sum() > 10


_CODES = "MIDNSHP=X"


def _ops_to_text(ops, n):
    return "".join("%d%s" % (int(o) >> 4, _CODES[int(o) & 15]) for o in ops[:n])


def _record_jobs(reflib, small_cfg, seed):
    """Reads at the locations the (affine-gap-free) reference aligner put them, in alignment direction; reads placed by hand
    across contig ends (with and without indels), starts shifted by a few bases (leading insertion / deletion => the
    front-clipping retry protocol), soft and hard clips."""
    from snap_b200 import synth
    rng = np.random.default_rng(seed)
    bases, starts = small_cfg.padded_bases()
    data = []; jobs = []; off = 0

    def add(read, loc, direction, cb=0, xb=0, ca=0, fh=0, bh=0, use_m=0):
        nonlocal off
        data.append(np.asarray(read, dtype=np.uint8)); data.append(np.zeros(16, dtype=np.uint8))
        jobs.append((off + cb, loc, len(read) - cb - ca, cb, xb, ca, fh, bh, direction, use_m))
        off += len(read) + 16

    ridx = reflib.RefIndex(small_cfg.idx)
    p = reflib.default_params(maxDist=14, useAffineGap=0)
    for name in ("noisy150", "indel100", "long250"):
        rb = small_cfg.reads[name]
        res, _ = reflib.RefSingleAligner(ridx, p).align(rb)
        for i in range(rb.n):
            if res[i]["status"] == 0:
                continue
            b, _q = rb.read(i)
            r = np.frombuffer(b, dtype=np.uint8)
            if res[i]["direction"] == 1:
                r = synth.revcomp(r)
            shift = int(rng.choice([0, 0, 0, 0, 0, 1, -1, 2, -3]))
            cb = int(rng.choice([0, 0, 0, 3, 7])); ca = int(rng.choice([0, 0, 0, 2, 9]))
            add(r, int(res[i]["location"]) + cb + shift, int(res[i]["direction"]), cb=cb, ca=ca,
                fh=int(rng.choice([0, 0, 0, 5])), bh=int(rng.choice([0, 0, 0, 4])), use_m=int(rng.integers(0, 2)))
    # reads hanging over a contig's end by 1..40 bases, clean / with a deletion / with an insertion near the end
    for c, contig in enumerate(small_cfg.contigs):
        end_loc = int(starts[c]) + contig.size
        for over in (1, 2, 5, 17, 40):
            for kind in ("clean", "del", "ins", "sub"):
                src = contig[contig.size - 150 + over - 8: contig.size].copy()
                if kind == "del":
                    src = np.delete(src, [100, 101])
                elif kind == "ins":
                    src = np.insert(src, 90, [ord("A"), ord("C"), ord("A")])
                elif kind == "sub":
                    src[60] = ord("A") if src[60] != ord("A") else ord("C")
                tail = np.frombuffer(bytes(rng.choice(list(b"ACGT"), size=over)), dtype=np.uint8)
                read = np.concatenate([src, tail])
                loc = end_loc - (150 - over + 8) if kind != "del" else end_loc - (150 - over + 8)
                add(read, loc, 0, use_m=int(rng.integers(0, 2)))
                add(read, loc, 0, xb=3, use_m=int(rng.integers(0, 2)))
    return (np.concatenate(data), np.array(jobs, dtype=reflib.CIGAR_JOB_DTYPE), ridx)


@pytest.mark.parametrize("seed", [11, 12])
def test_record_cigars_equal_reference(reflib, small_cfg, seed):
    """sg_cigar.h vs SAMFormat::computeCigarString (LV overload) on the reference's own alignments and on hand-placed reads
    that hang over contig ends: kind (retry / '*' / CIGAR), the CIGAR text incl. soft and hard clips, edit distance, the
    front-clipping verdict and the reference span."""
    data, jobs, ridx = _record_jobs(reflib, small_cfg, seed)
    want = reflib.cigar_lv_batch(ridx, data, jobs)
    got = hs.cigar_lv_batch(hs.HsIndex(small_cfg.idx), data, jobs)
    for f in ("kind", "addFrontClipping"):
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, (f, int(bad[0]), jobs[int(bad[0])], want[int(bad[0])], got[int(bad[0])])
    ok = want["kind"] == 2
    for i in np.nonzero(ok)[0]:
        assert want[i]["cigar"].decode() == _ops_to_text(got[i]["ops"], int(got[i]["nOps"])), (int(i), jobs[i])
    assert (want["editDistance"][ok] == got["editDistance"][ok]).all() and (want["refSpan"][ok] == got["refSpan"][ok]).all()
    assert ok.sum() > 1500 and (want["kind"] == 0).sum() > 50
    texts = [w.decode() for w in want["cigar"][ok]]
    assert sum("S" in t for t in texts) > 300 and sum("H" in t for t in texts) > 100 and sum("D" in t or "I" in t for t in texts) > 100


def _agc_jobs(reflib, n, seed):
    """Global affine-gap problems: patterns cut from random text with substitutions / insertions / deletions (incl. runs, near the
    ends, homopolymer neighbourhoods for the 'flip' heuristics), text = the window the SAM writer would pass (pattern + slack),
    low and high qualities (the heuristics look at quality < 65)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = alphabet[rng.integers(0, 4, size=300000)].copy()
    # homopolymer-rich stretches
    for _ in range(3000):
        p = int(rng.integers(0, text.size - 10)); text[p:p + int(rng.integers(2, 6))] = text[p]
    pats = []; quals = []; jobs = []; off = 0
    for _ in range(n):
        plen = int(rng.integers(12, 260))
        start = int(rng.integers(100, text.size - 700))
        src = text[start:start + plen + 60].tolist()
        out = []; si = 0
        n_edit = int(rng.choice([0, 1, 1, 2, 3, 5, 8]))
        edits = set(int(x) for x in rng.integers(0, plen, size=n_edit))
        while len(out) < plen:
            pos = len(out)
            if pos in edits:
                kind = int(rng.integers(0, 4))
                if kind <= 1:
                    out.append(int(alphabet[(int(np.searchsorted(alphabet, src[si])) + 1 + int(rng.integers(0, 3))) % 4])); si += 1
                elif kind == 2:
                    for _k in range(int(rng.integers(1, 4))):
                        out.append(int(alphabet[rng.integers(0, 4)]))
                else:
                    si += int(rng.integers(1, 5)); out.append(src[si]); si += 1
                edits.discard(pos)
            else:
                out.append(src[si]); si += 1
        pats.append(np.array(out[:plen], dtype=np.uint8)); pats.append(np.zeros(16, dtype=np.uint8))
        q = rng.integers(40, 75, size=plen).astype(np.uint8)
        quals.append(q); quals.append(np.zeros(16, dtype=np.uint8))
        slack = int(rng.choice([0, 3, 10, 27, 127]))
        jobs.append((start + int(rng.choice([0, 0, 0, 1, 2])), off, plen + slack, plen, int(rng.choice([5, 15, 27, 60])), int(rng.integers(0, 2))))
        off += plen + 16
    return text, np.concatenate(pats), np.concatenate(quals), np.array(jobs, dtype=reflib.AGC_JOB_DTYPE)


@pytest.mark.parametrize("seed,params", [(21, (1, 4, 6, 1)), (22, (1, 4, 6, 1)), (23, (2, 3, 4, 2))])
def test_affine_gap_global_cigar_equals_reference(reflib, seed, params):
    """sg_ag_cigar.h vs AffineGapVectorizedWithCigar::computeGlobalScore: BAM operations, edit count, net deletions, tail insertions."""
    text, pat, qual, jobs = _agc_jobs(reflib, 3000, seed)
    want = reflib.ag_cigar_global_batch(text, pat, qual, jobs, params)
    got = hs.ag_cigar_global_batch(text, pat, qual, jobs, reflib.AGC_OUT_DTYPE, params)
    for f in ("score", "nOps", "netDel", "tailIns"):
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, (f, int(bad[0]), jobs[int(bad[0])], want[int(bad[0])], got[int(bad[0])])
    assert (want["ops"] == got["ops"]).all()
    assert (want["score"] >= 0).mean() > 0.9 and (want["netDel"] > 0).sum() > 100 and (want["tailIns"] > 0).sum() > 20


@pytest.mark.parametrize("seed", [31, 32])
def test_affine_gap_normalized_cigar_equals_reference(reflib, seed):
    """computeGlobalScoreNormalized: the banded form (scores shifted by MAX_READ_LENGTH) with its fall-back to the unbanded one, and
    the front-clipping verdict.  One job sequence through one reference object and one restatement scratch: where the banded
    traceback steps onto cells the call did not write, both read what the same earlier calls left."""
    text, pat, qual, jobs = _agc_jobs(reflib, 3000, seed)
    jobs = jobs.copy()
    jobs["w"] = np.random.default_rng(seed).choice([3, 8, 14, 27], size=jobs.size)       # k: small enough that most patterns take the banded form
    want = reflib.ag_cigar_norm_batch(text, pat, qual, jobs)
    got = hs.ag_cigar_norm_batch(text, pat, qual, jobs, reflib.AGC_NORM_OUT_DTYPE)
    for f in ("score", "addFrontClipping", "nOps", "netDel", "tailIns"):
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, (f, int(bad[0]), bad.size, jobs[int(bad[0])], want[int(bad[0])], got[int(bad[0])])
    assert (want["ops"] == got["ops"]).all()
    banded = jobs["patternLen"] >= 3 * (2 * jobs["w"] + 1)
    assert banded.mean() > 0.5 and (want["score"] > 0).sum() > 1000 and (want["addFrontClipping"] != 0).sum() > 10


def _record_jobs_ag(reflib, small_cfg, seed):
    """Reads the reference aligner rescored with affine gap, at its locations with its clip counts and scores; starts shifted now and
    then; reads hanging over contig ends."""
    from snap_b200 import synth
    rng = np.random.default_rng(seed)
    bases, starts = small_cfg.padded_bases()
    data = []; qual = []; jobs = []; off = 0

    def add(read, q, loc, direction, score, cb=0, xb=0, ca=0, fh=0, bh=0, use_m=0):
        nonlocal off
        data.append(np.asarray(read, dtype=np.uint8)); data.append(np.zeros(16, dtype=np.uint8))
        qual.append(np.asarray(q, dtype=np.uint8)); qual.append(np.zeros(16, dtype=np.uint8))
        jobs.append((off + cb, loc, len(read) - cb - ca, cb, xb, ca, fh, bh, direction, use_m, score, 0))
        off += len(read) + 16

    ridx = reflib.RefIndex(small_cfg.idx)
    p = reflib.default_params(maxDist=14)
    for name in ("noisy150", "indel100", "long250"):
        rb = small_cfg.reads[name]
        res, _ = reflib.RefSingleAligner(ridx, p).align(rb)
        for i in range(rb.n):
            if res[i]["status"] == 0 or res[i]["usedAffineGapScoring"] == 0:
                continue
            b, q = rb.read(i)
            r = np.frombuffer(b, dtype=np.uint8); qq = np.frombuffer(q, dtype=np.uint8)
            if res[i]["direction"] == 1:
                r = synth.revcomp(r); qq = qq[::-1]
            shift = int(rng.choice([0, 0, 0, 0, 0, 0, 1, -1, 2]))
            add(r, qq, int(res[i]["location"]) + shift, int(res[i]["direction"]), int(res[i]["score"]) + int(rng.choice([0, 0, 1, 3])),
                cb=int(res[i]["basesClippedBefore"]), ca=int(res[i]["basesClippedAfter"]),
                fh=int(rng.choice([0, 0, 0, 5])), bh=int(rng.choice([0, 0, 0, 4])), use_m=int(rng.integers(0, 2)))
    for c, contig in enumerate(small_cfg.contigs):
        end_loc = int(starts[c]) + contig.size
        for over in (1, 2, 5, 17, 40):
            for kind in ("clean", "del", "ins", "sub"):
                src = contig[contig.size - 150 + over - 8: contig.size].copy()
                if kind == "del":
                    src = np.delete(src, [100, 101])
                elif kind == "ins":
                    src = np.insert(src, 90, [ord("A"), ord("C"), ord("A")])
                elif kind == "sub":
                    src[60] = ord("A") if src[60] != ord("A") else ord("C")
                tail = np.frombuffer(bytes(rng.choice(list(b"ACGT"), size=over)), dtype=np.uint8)
                read = np.concatenate([src, tail])
                q = rng.integers(40, 75, size=read.size).astype(np.uint8)
                loc = end_loc - (150 - over + 8)
                for score in (4, 10, 25):
                    add(read, q, loc, 0, score, use_m=int(rng.integers(0, 2)))
    return (np.concatenate(data), np.concatenate(qual), np.array(jobs, dtype=reflib.CIGAR_AG_JOB_DTYPE), ridx)


@pytest.mark.parametrize("seed", [41, 42])
def test_record_cigars_affine_gap_equal_reference(reflib, small_cfg, seed):
    """sg_cigar_ag vs SAMFormat::computeCigarString (affine-gap overload) on the reference aligner's own affine-gap results and on
    hand-placed reads across contig ends."""
    data, qual, jobs, ridx = _record_jobs_ag(reflib, small_cfg, seed)
    want = reflib.cigar_ag_batch(ridx, data, qual, jobs)
    got = hs.cigar_ag_batch(hs.HsIndex(small_cfg.idx), data, qual, jobs)
    for f in ("kind", "addFrontClipping"):
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, (f, int(bad[0]), bad.size, jobs[int(bad[0])], want[int(bad[0])], got[int(bad[0])])
    ok = want["kind"] == 2
    for i in np.nonzero(ok)[0]:
        assert want[i]["cigar"].decode() == _ops_to_text(got[i]["ops"], int(got[i]["nOps"])), (int(i), jobs[i], want[i], got[i])
    assert (want["editDistance"][ok] == got["editDistance"][ok]).all() and (want["refSpan"][ok] == got["refSpan"][ok]).all()
    texts = [w.decode() for w in want["cigar"][ok]]
    assert ok.sum() > 500 and sum("D" in t or "I" in t for t in texts) > 150 and sum("S" in t for t in texts) > 50


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("indel100", []), ("std150", ["-=" ]), ("noisy150", ["-G-"])])
def test_sam_records_equal_reference_binary(reflib, small_cfg, tmp_path, name, extra):
    """End to end for the output stage: the SAM file `snap-aligner single ... -o out.sam -t 1` writes vs sg_sam.h run over the
    result records of the same reads (record for record, byte for byte, headers aside).  `-=`: =/X CIGARs instead of M;
    `-G-`: affine gap off (every CIGAR from the Landau-Vishkin routine)."""
    import subprocess
    rb = small_cfg.reads[name]
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.sam")
    rb.write_fastq(fq)
    r = subprocess.run([reflib.SNAP_ALIGNER, "single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    want = [l for l in open(out, "rb").read().split(b"\n") if l and not l.startswith(b"@")]
    use_ag = "-G-" not in extra
    p = reflib.default_params(maxDist=14, useAffineGap=1 if use_ag else 0)
    res, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), p).align(rb)
    text = hs.sam_single(hs.HsIndex(small_cfg.idx), rb, [b"r%d" % i for i in range(rb.n)], res, use_m=("-=" not in extra), use_affine_gap=use_ag)
    got = [l for l in text.split(b"\n") if l]
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])
    cig = [l.split(b"\t")[5] for l in want]
    assert sum(b"I" in c or b"D" in c for c in cig) > 20
    if name == "noisy150":
        assert sum(c == b"*" for c in cig) >= 1                # unaligned reads are written too


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("clipped150", []), ("std150", ["-="]), ("len100", ["-hc"])])
def test_sam_pair_records_equal_reference_binary(reflib, small_cfg, tmp_path, name, extra):
    """The same for pairs: `snap-aligner paired ... -o out.sam -t 1` vs sg_sam_write_pair over the pair result records (mate
    fields, template length, write order, QS)."""
    import subprocess
    pb = small_cfg.pairs[name]
    f1 = str(tmp_path / "p1.fq"); f2 = str(tmp_path / "p2.fq"); out = str(tmp_path / "o.sam")
    with open(f1, "wb") as a, open(f2, "wb") as b:
        for i in range(pb.n // 2):
            x, q = pb.read(2 * i); a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, x, q))
            x, q = pb.read(2 * i + 1); b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, x, q))
    r = subprocess.run([reflib.SNAP_ALIGNER, "paired", small_cfg.idx, f1, f2, "-o", out, "-t", "1"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    want = [l for l in open(out, "rb").read().split(b"\n") if l and not l.startswith(b"@")]
    kw, pkw = PAIRED_DEFAULT if "-hc" not in extra else PAIRED_HC
    p, pp = reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)
    res, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx), p, pp).align(pb)
    ids = []
    for i in range(pb.n // 2):
        ids += [b"p%d/1" % i, b"p%d/2" % i]
    text = hs.sam_single(hs.HsIndex(small_cfg.idx), pb, ids, res, use_m=("-=" not in extra), use_affine_gap=True, paired=True)
    got = [l for l in text.split(b"\n") if l]
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


def test_sam_records_do_not_depend_on_call_history(reflib, small_cfg, tmp_path, monkeypatch):
    """AffineGapVectorizedWithCigar never clears its action array, and its banded traceback can step onto cells the current call
    did not write.  The device form gives every thread its own (uninitialised) array, so it matters whether any record of a real
    run depends on that: fill the array with junk before every read and compare with the reference binary's file again."""
    monkeypatch.setenv("HS_SAM_SCRUB", "1")
    test_sam_records_equal_reference_binary(reflib, small_cfg, tmp_path, "noisy150", [])
    test_sam_pair_records_equal_reference_binary(reflib, small_cfg, tmp_path, "noisy150", [])


@pytest.mark.parametrize("clip", ["back", "both"])
def test_sam_records_with_quality_clipping_equal_reference_binary(reflib, small_cfg, tmp_path, clip):
    """Reads with '#' quality tails (and heads): the aligner sees the clipped view (Read::clip, default -C-+ = back only; -C++ = both
    ends), the SAM record carries the whole read with S operations."""
    import subprocess
    from snap_b200 import synth
    rng = np.random.default_rng(5)
    rb = small_cfg.reads["noisy150"]
    reads = []; fronts = []; clens = []
    for i in range(600):
        b, q = rb.read(i)
        q = bytearray(q)
        tail = int(rng.choice([0, 0, 3, 11, 40])); head = int(rng.choice([0, 0, 0, 5])) if clip == "both" else 0
        if len(b) < 70:
            tail = head = 0
        for k in range(tail):
            q[len(q) - 1 - k] = ord("#")
        for k in range(head):
            q[k] = ord("#")
        reads.append((b, bytes(q))); fronts.append(head); clens.append(len(b) - head - tail)
    full = synth.ReadBatch.from_lists(reads)
    clipped = synth.ReadBatch.from_lists([(b[f:f + n], q[f:f + n]) for (b, q), f, n in zip(reads, fronts, clens)])
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.sam")
    full.write_fastq(fq)
    argv = [reflib.SNAP_ALIGNER, "single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"] + (["-C++"] if clip == "both" else [])
    r = subprocess.run(argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    want = [l for l in open(out, "rb").read().split(b"\n") if l and not l.startswith(b"@")]
    res, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params(maxDist=14)).align(clipped)
    text = hs.sam_single(hs.HsIndex(small_cfg.idx), full, [b"r%d" % i for i in range(full.n)], res, front_clipped=fronts, clipped_lens=clens)
    got = [l for l in text.split(b"\n") if l]
    assert len(want) == len(got) == full.n
    bad = [i for i in range(full.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])
    assert sum(b"S" in l.split(b"\t")[5] for l in want) > 200


def _bam_records(path):
    """The alignment records of a BAM file (BGZF blocks inflated, header and reference table skipped), as a list of bytes."""
    import gzip, struct
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\x01"
    p = 8 + struct.unpack("<i", raw[4:8])[0]
    n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
    for _ in range(n_ref):
        l = struct.unpack("<i", raw[p:p + 4])[0]; p += 8 + l
    recs = []
    while p < len(raw):
        b = struct.unpack("<i", raw[p:p + 4])[0]
        recs.append(raw[p:p + 4 + b]); p += 4 + b
    return recs


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("indel100", ["-="]), ("noisy150", ["-G-"])])
def test_bam_records_equal_reference_binary(reflib, small_cfg, tmp_path, name, extra):
    """`snap-aligner single ... -o out.bam -t 1` vs sg_bam.h over the result records of the same reads: every record, byte for
    byte (fixed fields, bin, CIGAR words, 4-bit sequence, qualities, the binary tags)."""
    import subprocess
    rb = small_cfg.reads[name]
    fq = str(tmp_path / "r.fq"); out = str(tmp_path / "o.bam")
    rb.write_fastq(fq)
    r = subprocess.run([reflib.SNAP_ALIGNER, "single", small_cfg.idx, fq, "-o", out, "-t", "1", "-d", "14"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    want = _bam_records(out)
    use_ag = "-G-" not in extra
    res, _ = reflib.RefSingleAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params(maxDist=14, useAffineGap=1 if use_ag else 0)).align(rb)
    blob = hs.bam_single(hs.HsIndex(small_cfg.idx), rb, [b"r%d" % i for i in range(rb.n)], res, use_m=("-=" not in extra), use_affine_gap=use_ag)
    import struct
    got = []; p = 0
    while p < len(blob):
        b = struct.unpack("<i", blob[p:p + 4])[0]
        got.append(blob[p:p + 4 + b]); p += 4 + b
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


@pytest.mark.parametrize("name,extra", [("noisy150", []), ("clipped150", ["-="]), ("std150", [])])
def test_bam_pair_records_equal_reference_binary(reflib, small_cfg, tmp_path, name, extra):
    """`snap-aligner paired ... -o out.bam -t 1` vs sg_bam_write_pair: both records of every pair, byte for byte (mate fields, bin of an
    unmapped end at its mate's position, template length, QS)."""
    import subprocess, struct
    pb = small_cfg.pairs[name]
    f1 = str(tmp_path / "p1.fq"); f2 = str(tmp_path / "p2.fq"); out = str(tmp_path / "o.bam")
    ids = []
    with open(f1, "wb") as a, open(f2, "wb") as b:
        for i in range(pb.n // 2):
            x, q = pb.read(2 * i); a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, x, q))
            x, q = pb.read(2 * i + 1); b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, x, q))
            ids += [b"p%d/1" % i, b"p%d/2" % i]
    r = subprocess.run([reflib.SNAP_ALIGNER, "paired", small_cfg.idx, f1, f2, "-o", out, "-t", "1"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    want = _bam_records(out)
    kw, pkw = PAIRED_DEFAULT
    res, _ = reflib.RefPairedAligner(reflib.RefIndex(small_cfg.idx), reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)).align(pb)
    blob = hs.bam_single(hs.HsIndex(small_cfg.idx), pb, ids, res, use_m=("-=" not in extra), paired=True)
    got = []; p = 0
    while p < len(blob):
        b = struct.unpack("<i", blob[p:p + 4])[0]
        got.append(blob[p:p + 4 + b]); p += 4 + b
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), want[bad[0]], got[bad[0]])


@pytest.mark.parametrize("use_ag", [True, False])
def test_write_reads_retry_loop_equals_reference(reflib, small_cfg, use_ag):
    """The loop around the record formatter that applies a front-clipping verdict (move the start for a leading deletion, soft-clip a
    leading insertion, give the read up at a contig boundary) hardly ever runs on real placements (0 of ~3000 in the datasets above).
    Here the reference aligner's placements are shifted by a few bases either way, moved next to contig ends, or left alone, and
    SimpleReadWriter::writeReads itself (into memory, through oracle/ref_harness.cpp) is the oracle for sg_sam_write_single."""
    rng = np.random.default_rng(77)
    rb = small_cfg.reads["noisy150"]
    bases, starts = small_cfg.padded_bases()
    ridx = reflib.RefIndex(small_cfg.idx)
    res, _ = reflib.RefSingleAligner(ridx, reflib.default_params(maxDist=14, useAffineGap=1 if use_ag else 0)).align(rb)
    res = res.copy()
    for i in range(rb.n):
        if res[i]["status"] == 0:
            continue
        how = int(rng.integers(0, 6))
        if how <= 2:
            res[i]["location"] += int(rng.choice([-3, -2, -1, 1, 2, 3, 5]))
        elif how == 3:
            c = int(rng.integers(0, len(starts)))
            end = int(starts[c]) + small_cfg.contigs[c].size
            res[i]["location"] = end - int(rng.integers(1, 150))             # hangs over / sits at the end of a contig
        elif how == 4:
            c = int(rng.integers(0, len(starts)))
            res[i]["location"] = int(starts[c]) - int(rng.integers(1, 40))   # starts before its contig
    ids = [b"r%d" % i for i in range(rb.n)]
    want = [l for l in reflib.write_reads(ridx, rb, ids, res, use_affine_gap=use_ag).split(b"\n") if l]
    got = [l for l in hs.sam_single(hs.HsIndex(small_cfg.idx), rb, ids, res, use_affine_gap=use_ag).split(b"\n") if l]
    assert len(want) == len(got) == rb.n
    bad = [i for i in range(rb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), res[bad[0]], want[bad[0]], got[bad[0]])
    flags = [int(l.split(b"\t")[1]) for l in want]
    gave_up = sum(1 for i in range(rb.n) if res[i]["status"] != 0 and (flags[i] & 4))
    assert gave_up > 5                                                       # reads given up at a contig boundary


def test_write_pairs_retry_loops_equal_reference(reflib, small_cfg):
    """The same for pairs: SimpleReadWriter::writePairs itself (into memory) vs sg_sam_write_pair, with one or both ends of the
    reference aligner's placements shifted, moved to a contig end, or put before a contig (re-ordering of the two records, ends given
    up, mates pointing at each other)."""
    rng = np.random.default_rng(78)
    pb = small_cfg.pairs["noisy150"]
    bases, starts = small_cfg.padded_bases()
    ridx = reflib.RefIndex(small_cfg.idx)
    kw, pkw = PAIRED_DEFAULT
    res, _ = reflib.RefPairedAligner(ridx, reflib.default_params_paired(**kw), reflib.default_paired_params(**pkw)).align(pb)
    res = res.copy()
    for i in range(pb.n // 2):
        for w in range(2):
            if res[i]["status"][w] == 0:
                continue
            how = int(rng.integers(0, 8))
            if how <= 2:
                res[i]["location"][w] += int(rng.choice([-3, -2, -1, 1, 2, 3, 5]))
            elif how == 3:
                c = int(rng.integers(0, len(starts)))
                res[i]["location"][w] = int(starts[c]) + small_cfg.contigs[c].size - int(rng.integers(1, 150))
            elif how == 4:
                c = int(rng.integers(0, len(starts)))
                res[i]["location"][w] = int(starts[c]) - int(rng.integers(1, 40))
    ids = []
    for i in range(pb.n // 2):
        ids += [b"p%d/1" % i, b"p%d/2" % i]
    want = [l for l in reflib.write_reads(ridx, pb, ids, res, paired=True).split(b"\n") if l]
    got = [l for l in hs.sam_single(hs.HsIndex(small_cfg.idx), pb, ids, res, paired=True).split(b"\n") if l]
    assert len(want) == len(got) == pb.n
    bad = [i for i in range(pb.n) if want[i] != got[i]]
    assert bad == [], (len(bad), res[bad[0] // 2], want[bad[0]], got[bad[0]])


def test_output_stage_on_an_index_with_alt_contigs(reflib, tmp_path):
    """A FASTA whose ALT contigs are interleaved with the primary ones: the reference's indexer moves them to the end (internal != original contig numbers).
    SAM records (RNAME by name) and the file header (@SQ in original order, AH:* on ALT contigs) are byte-identical to the reference binary's.  BAM records
    are identical EXCEPT refID: the writer files the internal contig number where the reference files the original one -- the library refuses BAM on such
    an index (snapgpu_sam_set_format) until the mapping is in the kernel; this test pins the difference to exactly that field."""
    import ctypes as C
    import struct
    import subprocess
    import sorted_data
    from snap_b200 import synth
    P = lambda name: str(tmp_path / name)
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    primary = synth.make_contigs(2, 100_000, seed=71, repeat_frac=0.15)

    def haplo(src, lo, hi, div):
        c = src[lo:hi].copy()
        m = rng.random(c.size) < div
        c[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        return c

    names = [b"chr1", b"chr1_KI270_alt", b"chr2", b"HLA-A*01:01", b"chrUn_decoy_alt"]
    contigs = [primary[0], haplo(primary[0], 20_000, 50_000, 0.01), primary[1], haplo(primary[1], 60_000, 80_000, 0.03), acgt[rng.integers(0, 4, 15_000)]]
    with open(P("alt.fa"), "wb") as f:
        for n, c in zip(names, contigs):
            f.write(b">" + n + b"\n")
            for i in range(0, c.size, 100):
                f.write(c[i:i + 100].tobytes() + b"\n")
    reflib.build_reference_index(reflib.SNAP_ALIGNER, P("alt.fa"), P("idx"))
    rb = synth.make_reads(contigs, 1500, 150, seed=78, sub_rate=0.01, ins_rate=0.003, del_rate=0.003)
    rb.write_fastq(P("r.fq"))
    ids = [b"r%d" % i for i in range(rb.n)]
    res, _ = reflib.RefSingleAligner(reflib.RefIndex(P("idx")), reflib.default_params(maxDist=14)).align(rb)
    for fmt in ("sam", "bam"):
        r = subprocess.run([reflib.SNAP_ALIGNER, "single", P("idx"), P("r.fq"), "-o", P("o." + fmt), "-t", "1", "-d", "14"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-500:]
    # SAM records and header
    lines = open(P("o.sam"), "rb").read().split(b"\n")
    want = [l for l in lines if l and not l.startswith(b"@")]
    got = [l for l in hs.sam_single(hs.HsIndex(P("idx")), rb, ids, res).split(b"\n") if l]
    assert want == got
    assert sum(1 for l in want if l.split(b"\t")[2] in (b"chr1_KI270_alt", b"HLA-A*01:01", b"chrUn_decoy_alt")) > 200
    hdr = b"".join(l + b"\n" for l in lines if l.startswith(b"@"))
    pg = [l for l in hdr.split(b"\n") if l.startswith(b"@PG\tID:SNAP\t")][0]
    rg = [l for l in hdr.split(b"\n") if l.startswith(b"@RG")][0]
    L = C.CDLL(hs.build())
    L.hs_index_open.restype = C.c_void_p
    L.hs_index_open.argtypes = [C.c_char_p]
    L.hs_sam_header.restype = C.c_int64
    L.hs_sam_header.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64]
    buf = (C.c_uint8 * (1 << 20))()
    n = L.hs_sam_header(L.hs_index_open(P("idx").encode()), 0, 0, pg.split(b"\tCL:")[1].rsplit(b"\tVN:", 1)[0], pg.rsplit(b"\tVN:", 1)[1], rg, buf, 1 << 20)
    assert bytes(buf[:n]) == hdr and hdr.count(b"\tAH:*\n") == 3
    # BAM records: identical but for refID (bytes 4..8 of a record), which holds the internal contig number
    refs, want, _ = sorted_data.load_bam(P("o.bam"))
    assert [r[0] for r in refs] == names                       # the header's reference table is in ORIGINAL order
    blob = hs.bam_single(hs.HsIndex(P("idx")), rb, ids, res)
    got, q = [], 0
    while q < len(blob):
        b = struct.unpack("<i", blob[q:q + 4])[0]; got.append(blob[q:q + 4 + b]); q += 4 + b
    assert len(got) == len(want)
    internal_to_original = {0: 0, 1: 2, 2: 1, 3: 3, 4: 4, -1: -1}      # chr1, chr2, then the three ALT contigs
    n_diff = 0
    for w, g in zip(want, got):
        assert w[:4] == g[:4] and w[8:] == g[8:]
        rw, rg_ = struct.unpack("<i", w[4:8])[0], struct.unpack("<i", g[4:8])[0]
        assert rw == internal_to_original[rg_]
        n_diff += rw != rg_
    assert n_diff > 300
