#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/.  Run HERE (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

1. ref_unit_vectors.json -- the known-answer vectors of the reference's OWN unit tests, parsed out of
   /root/reference/tests/LandauVishkinTest.cpp:11-32 and AffineGapVectorizedTest.cpp:40-67, plus the 2-read
   end-to-end data test (tests/datatest/datatest.{fa,fq} -> correct-fq-datatest.sam columns 1-5).
2. e2e_small.npz -- a seeded synthetic genome + reads and the reference's SingleAlignmentResult for each read under
   three option sets, produced by the compiled reference (oracle/_ref/libsnapref.so).
3. leaf_lv.npz / leaf_ag.npz -- random LV / affine-gap jobs (tests/jobs.py) with the reference's outputs.
4. output_small.json.gz (`make_golden.py output`) -- the reference unit test's 30 CIGAR known answers and the SAM / BAM records the
   reference binary writes for the reads of e2e_small.npz.
The GPU box has no /root/reference; the `-m gpu` tests compare the CUDA path against these files (and, when the
prebuilt oracle/_ref travelled along, against it directly).
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import reflib  # noqa: E402
from snap_b200 import synth  # noqa: E402
import jobs as J  # noqa: E402

REF = "/root/reference"

E2E_CONFIGS = {
    "default_d14": dict(maxDist=14),
    "noag_d14": dict(maxDist=14, useAffineGap=0),
    "ne_d20": dict(maxDist=20, noEditDistance=1, useAffineGap=0),     # `-ne` (AlignerOptions.cpp:964-967)
}


def unit_vectors():
    out = {"lv": [], "ag": [], "source": "amplab/snap tests/LandauVishkinTest.cpp:11-32, tests/AffineGapVectorizedTest.cpp:40-67"}
    src = open(os.path.join(REF, "tests", "LandauVishkinTest.cpp")).read()
    for m in re.finditer(r'ASSERT_EQ\((-?\d+), lv\.computeEditDistance\("([^"]*)", (\d+), "([^"]*)", (\d+), (\d+)\)\)', src):
        out["lv"].append(dict(expected=int(m.group(1)), text=m.group(2), textLen=int(m.group(3)), pattern=m.group(4),
                              patternLen=int(m.group(5)), k=int(m.group(6))))
    src = open(os.path.join(REF, "tests", "AffineGapVectorizedTest.cpp")).read()
    src = src.split("/* Edit distance tests */")[0]
    for m in re.finditer(r'^\s*ASSERT_EQ\((-?\d+), computeScore\("([^"]*)", (\d+), "([^"]*)", (?:NULL|"[^"]*"), (\d+), (\d+), (\d+)\)\)', src, re.M):
        out["ag"].append(dict(expected=int(m.group(1)), text=m.group(2), textLen=int(m.group(3)), pattern=m.group(4),
                              patternLen=int(m.group(5)), w=int(m.group(6)), scoreInit=int(m.group(7))))
    out["ag_params"] = [1, 4, 6, 1, 10, 5]            # AffineGapVectorizedTest fixture: ag(1, 4, 6, 1, 10, 5)
    # data test: reference + 2 reads + expected SAM columns
    fa = open(os.path.join(REF, "tests", "datatest", "datatest.fa")).read().split("\n")
    fq = open(os.path.join(REF, "tests", "datatest", "datatest.fq")).read().split("\n")
    sam = [l.split("\t") for l in open(os.path.join(REF, "tests", "datatest", "correct-fq-datatest.sam")) if not l.startswith("@")]
    out["datatest"] = dict(contig_name=fa[0][1:], contig="".join(fa[1:]).strip(),
                           reads=[dict(bases=fq[1], quals=fq[3]), dict(bases=fq[5], quals=fq[7])],
                           expected=[dict(flag=int(s[1]), pos=int(s[3]), mapq=int(s[4]), cigar=s[5]) for s in sam])
    assert len(out["lv"]) == 11 and len(out["ag"]) == 10, (len(out["lv"]), len(out["ag"]))
    return out


def e2e_small(tmp):
    contigs = synth.make_contigs(2, 60_000, seed=101, repeat_frac=0.25)
    fasta = os.path.join(tmp, "ref.fa")
    synth.write_fasta(fasta, contigs)
    reflib.build_reference_index(reflib.SNAP_ALIGNER, fasta, os.path.join(tmp, "idx"))
    sets = [
        synth.make_reads(contigs, 250, 150, seed=102),
        synth.make_reads(contigs, 250, 150, seed=103, sub_rate=0.03, ins_rate=0.004, del_rate=0.004, n_run_frac=0.1, short_frac=0.1, random_frac=0.05),
        synth.make_reads(contigs, 150, 100, seed=104, sub_rate=0.01, ins_rate=0.008, del_rate=0.008),
    ]
    reads = synth.ReadBatch.from_lists([s.read(i) for s in sets for i in range(s.n)])
    ridx = reflib.RefIndex(os.path.join(tmp, "idx"))
    arrays = dict(contig0=contigs[0], contig1=contigs[1], bases=reads.bases, quals=reads.quals, offsets=reads.offsets, lens=reads.lens)
    for name, kw in E2E_CONFIGS.items():
        al = reflib.RefSingleAligner(ridx, reflib.default_params(**kw))
        res, ctr = al.align(reads)
        al.close()
        arrays["res_" + name] = res
        arrays["ctr_" + name] = np.array([ctr[k] for k in reflib.COUNTER_FIELDS], dtype=np.int64)
    return arrays


def fastq_small():
    """FASTQ text with CRLF ends, lower case, '.', '#' heads/tails, comments, a truncated last record, and the reference
    reader's output (FASTQReader::getReadFromBuffer + Read::clip via oracle/_ref) for every clipping mode."""
    contigs = synth.make_contigs(1, 60_000, seed=51)
    rb = synth.make_reads(contigs, 400, 120, seed=52, short_frac=0.25, n_run_frac=0.1)
    text = synth.make_fastq_text(rb, 53, crlf_frac=0.2, lower_frac=0.3, dot_frac=0.2, hash_tail_frac=0.35, hash_head_frac=0.2, comment_frac=0.3,
                                 plus_id_frac=0.3, truncate_last=True)
    out = {"text": text}
    for clip in (0, 1, 2, 3):
        b, q, off, ln, ido, idl, fc, used = reflib.fastq_parse(text, clip)
        out.update({"bases%d" % clip: b, "quals%d" % clip: q, "offsets%d" % clip: off, "lens%d" % clip: ln, "idoff%d" % clip: ido, "idlen%d" % clip: idl,
                    "front%d" % clip: fc, "used%d" % clip: np.array([used], dtype=np.int64)})
    return out


def output_small():
    """Output stage: (1) the CIGAR known answers of the reference's own unit test (tests/LandauVishkinTest.cpp:34-129), parsed
    out of its source; (2) for the reads of e2e_small.npz, the alignment records of the SAM and BAM files the reference BINARY writes
    (`snap-aligner single <idx> r.fq -o out.sam|out.bam -t 1 -d 14`; BAM records as hex, BGZF inflated)."""
    import gzip, struct, subprocess, tempfile
    src = open(os.path.join(REF, "tests", "LandauVishkinTest.cpp")).read()
    vec = []
    for m in re.finditer(r'lvc\.computeEditDistance\("([^"]*)", (\d+), "([^"]*)", (\d+), (\d+), cigarBuf, bufLen, (true|false)\);\s*ASSERT_STREQ\("([^"]*)", cigarBuf\);', src):
        vec.append(dict(text=m.group(1), textLen=int(m.group(2)), pattern=m.group(3), patternLen=int(m.group(4)), k=int(m.group(5)), useM=m.group(6) == "true",
                        cigar=m.group(7)))
    assert len(vec) == 30, len(vec)
    g = np.load(os.path.join(HERE, "e2e_small.npz"))
    reads = synth.ReadBatch(g["bases"], g["quals"], g["offsets"], g["lens"])
    out = {"source": "amplab/snap tests/LandauVishkinTest.cpp:34-129; snap-aligner single -d 14 -t 1 over the reads of e2e_small.npz", "lv_cigar": vec}
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "ref.fa"); synth.write_fasta(fa, [g["contig0"], g["contig1"]])
        idx = os.path.join(tmp, "idx"); reflib.build_reference_index(reflib.SNAP_ALIGNER, fa, idx)
        fq = os.path.join(tmp, "r.fq"); reads.write_fastq(fq)
        for ext in ("sam", "bam"):
            o = os.path.join(tmp, "o." + ext)
            r = subprocess.run([reflib.SNAP_ALIGNER, "single", idx, fq, "-o", o, "-t", "1", "-d", "14"], capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
            if ext == "sam":
                out["sam"] = [l for l in open(o).read().split("\n") if l and not l.startswith("@")]
            else:
                raw = gzip.open(o, "rb").read()
                p = 8 + struct.unpack("<i", raw[4:8])[0]
                n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
                for _ in range(n_ref):
                    p += 8 + struct.unpack("<i", raw[p:p + 4])[0]
                recs = []
                while p < len(raw):
                    b = struct.unpack("<i", raw[p:p + 4])[0]
                    recs.append(raw[p:p + 4 + b].hex()); p += 4 + b
                out["bam"] = recs
    assert len(out["sam"]) == len(out["bam"]) == reads.n
    return out


def main():
    import tempfile
    if len(sys.argv) > 1 and sys.argv[1] == "output":          # only the output-stage fixture (the others are left as committed)
        import gzip
        with gzip.open(os.path.join(HERE, "output_small.json.gz"), "wt") as f:
            json.dump(output_small(), f)
        print("output_small.json.gz %d bytes" % os.path.getsize(os.path.join(HERE, "output_small.json.gz")))
        return
    with open(os.path.join(HERE, "ref_unit_vectors.json"), "w") as f:
        json.dump(unit_vectors(), f, indent=1)
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(HERE, "e2e_small.npz"), **e2e_small(tmp))
    t, p, q, jb = J.lv_jobs(1500, 201)
    np.savez_compressed(os.path.join(HERE, "leaf_lv.npz"), text=t, pat=p, qual=q, jobs=jb,
                        out=reflib.lv_batch(t, p, q, jb.astype(reflib.LV_JOB_DTYPE)))
    t, p, q, jb = J.ag_jobs(1200, 202)
    np.savez_compressed(os.path.join(HERE, "leaf_ag.npz"), text=t, pat=p, qual=q, jobs=jb,
                        out=reflib.ag_batch(t, p, q, jb.astype(reflib.AG_JOB_DTYPE)))
    np.savez_compressed(os.path.join(HERE, "fastq_small.npz"), **fastq_small())
    for f in sorted(os.listdir(HERE)):
        print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
