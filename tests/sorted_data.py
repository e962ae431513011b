"""Duplicate-rich reads for the tests of row N4 (duplicate marking, BAM index): FASTQ files, and the coordinate-sorted BAM files the reference
binary writes for them -- `-so -S d` (sorted, duplicates not marked: the input of our stage) and `-so` (duplicates marked + .bai: the answer)."""
from __future__ import annotations

import gzip
import os
import struct
import sys
import subprocess

import numpy as np

from snap_b200 import synth


def _name(k: int, style: int) -> bytes:
    if style == 0:
        return b"r%d" % k
    if style == 1:      # 7-element (CASAVA 1.8) names: tile, x, y are elements 5-7
        return b"M01:%d:FC:1:%d:%d:%d" % (k % 7, 1000 + (k * 7) % 5, 100 + (k * 13) % 50, 200 + (k * 17) % 60)
    return b"HW:3:%d:%d:%d" % (10 + k % 3, k % 40, (k * 3) % 70)      # 5-element names: elements 3-5


def write_dup_fastqs(d: str, contigs, seed: int, dense: bool):
    """se.fq (single-end) and p1.fq / p2.fq: a third of the fragments come in 2-4 copies (same 5' end and strand, other errors, other qualities,
    some shorter so that reverse-strand copies start elsewhere), some copies of a pair lose a mate; `dense` packs everything into 5 kbp so that
    the reference's overlapping runs all hold many records."""
    rng = np.random.default_rng(seed)

    def mutate(w, rate=0.01):
        w = w.copy(); m = rng.random(w.size) < rate
        w[m] = synth.ACGT[rng.integers(0, 4, size=int(m.sum()))]
        return w

    def qual(n):
        return (33 + rng.integers(20, 41, size=n)).astype(np.uint8).tobytes()

    recs, k = [], 0
    for g in range(700 if dense else 500):
        c = int(rng.integers(0, len(contigs))); L = 150
        pos = int(rng.integers(1000, 6000)) if dense else int(rng.integers(0, contigs[c].size - 400))
        rc = bool(rng.integers(0, 2)); copies = 1 if g % 3 else int(rng.integers(2, 5)); style = g % 3
        for j in range(copies):
            ln = L if j == 0 or rng.random() < 0.6 else int(rng.integers(100, 150))
            w = mutate(contigs[c][pos + L - ln: pos + L] if rc else contigs[c][pos: pos + ln])
            recs.append((_name(k, style), (synth.revcomp(w) if rc else w).tobytes(), qual(ln))); k += 1
        if g % 10 == 0:
            recs.append((_name(k, style), mutate(contigs[c][pos + 3: pos + 3 + L]).tobytes(), qual(L))); k += 1
    with open(os.path.join(d, "se.fq"), "wb") as f:
        for i in rng.permutation(len(recs)):
            n, b, q = recs[i]; f.write(b"@" + n + b"\n" + b + b"\n+\n" + q + b"\n")
    pairs, k = [], 0
    for g in range(600 if dense else 400):
        c = int(rng.integers(0, len(contigs))); L = 100
        ins = int(rng.integers(250, 500))
        pos = int(rng.integers(1000, 6000)) if dense else int(rng.integers(0, contigs[c].size - 700))
        flip = bool(rng.integers(0, 2)); copies = 1 if g % 3 else int(rng.integers(2, 5)); style = (g // 2) % 3
        for j in range(copies):
            a = mutate(contigs[c][pos: pos + L]); b = synth.revcomp(mutate(contigs[c][pos + ins - L: pos + ins]))
            if j and g % 12 == 0:
                b = synth.ACGT[rng.integers(0, 4, size=L)]       # this copy's mate cannot be aligned
            r1, r2 = (b, a) if flip else (a, b)
            pairs.append((_name(k, style), r1.tobytes(), qual(L), r2.tobytes(), qual(L))); k += 1
    with open(os.path.join(d, "p1.fq"), "wb") as f1, open(os.path.join(d, "p2.fq"), "wb") as f2:
        for i in rng.permutation(len(pairs)):
            n, b1, q1, b2, q2 = pairs[i]
            f1.write(b"@" + n + b"/1\n" + b1 + b"\n+\n" + q1 + b"\n"); f2.write(b"@" + n + b"/2\n" + b2 + b"\n+\n" + q2 + b"\n")


def load_bam(path: str):
    """(reference table [(name, length)], alignment records, uncompressed size of everything before the first record)."""
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\x01"
    p = 8 + struct.unpack("<i", raw[4:8])[0]
    n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l = struct.unpack("<i", raw[p:p + 4])[0]
        refs.append((raw[p + 4:p + 4 + l - 1], struct.unpack("<i", raw[p + 4 + l:p + 8 + l])[0])); p += 8 + l
    hdr, recs = p, []
    while p < len(raw):
        b = struct.unpack("<i", raw[p:p + 4])[0]; recs.append(raw[p:p + 4 + b]); p += 4 + b
    return refs, recs, hdr


def bgzf_blocks(path: str):
    """[(compressed start, uncompressed start, uncompressed size)] of a BGZF file's members."""
    raw = open(path, "rb").read(); p = u = 0; out = []
    while p < len(raw):
        bsize = struct.unpack("<H", raw[p + 16:p + 18])[0] + 1
        isize = struct.unpack("<I", raw[p + bsize - 4:p + bsize])[0]
        out.append((p, u, isize)); p += bsize; u += isize
    return out


def our_blocks(total: int):
    """The members snapgpu_bgzf_device makes of `total` bytes (0xff00 payload bytes each, 31 of framing) + the 28-byte end-of-file member."""
    n = (total + 0xff00 - 1) // 0xff00
    return [(k * (0xff00 + 31), k * 0xff00, min(0xff00, total - k * 0xff00)) for k in range(n)] + [(n * 31 + total, total, 0)]


def parse_bai(raw: bytes, blocks):
    """A .bai with its virtual offsets turned back into uncompressed offsets: [({bin: [(start, end)]}, [linear index])] per reference; the second
    chunk of the metadata bin 37450 holds counts and is kept as it is; never-set linear entries (0) -> None."""
    cstart = {c: u for c, u, _ in blocks}
    endc, endu = blocks[-1][0] + 28, blocks[-1][1] + blocks[-1][2]

    def log(v):
        if v == 0:
            return None
        c, dlt = v >> 16, v & 0xffff
        if c not in cstart:
            assert c == endc and dlt == 0, (c, dlt)
            return endu
        return cstart[c] + dlt
    assert raw[:4] == b"BAI\x01"
    n_ref = struct.unpack("<i", raw[4:8])[0]; p = 8; refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack("<i", raw[p:p + 4])[0]; p += 4; bins = {}
        for _ in range(n_bin):
            b, nc = struct.unpack("<Ii", raw[p:p + 8]); p += 8; ch = []
            for k in range(nc):
                s, e = struct.unpack("<QQ", raw[p:p + 16]); p += 16
                ch.append((s, e) if (b == 37450 and k == 1) else (log(s), log(e)))
            assert b not in bins
            bins[b] = ch
        n_intv = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
        refs.append((bins, [log(v) for v in struct.unpack("<%dQ" % n_intv, raw[p:p + 8 * n_intv])])); p += 8 * n_intv
    assert p == len(raw)
    return refs


class SortedCase:
    """One FASTQ set through the reference twice: sorted without / with duplicate marking."""

    def __init__(self, d: str, tag: str, aligner: str, argv: list[str]):
        for suffix, extra in (("nodup", ["-S", "d"]), ("dup", [])):
            out = os.path.join(d, "%s_%s.bam" % (tag, suffix))
            # The reference binary itself dies now and then at the end of a `-so` run (seen: about 1 run in 50 under load, after the statistics header is
            # printed, nothing on stderr -- its sorting writer's threads): it is the oracle here, so a failed run is simply repeated.
            for attempt in range(4):
                r = subprocess.run([aligner] + argv + ["-o", out, "-t", "1", "-so"] + extra, capture_output=True, text=True)
                if r.returncode == 0:
                    break
                sys.stderr.write("sorted_data: reference run failed (rc %d, attempt %d): %s\n" % (r.returncode, attempt, " ".join(argv[:1] + extra)))
            assert r.returncode == 0, "rc %d: " % r.returncode + r.stdout[-500:] + r.stderr[-500:]
        self.refs, self.unmarked, _ = load_bam(os.path.join(d, tag + "_nodup.bam"))
        _, self.marked, self.header_bytes = load_bam(os.path.join(d, tag + "_dup.bam"))
        self.bam = os.path.join(d, tag + "_dup.bam")
        self.bai = open(self.bam + ".bai", "rb").read()

    def contig_starts(self, padding: int = 2000):
        """beginningLocation of every contig as SNAP lays a genome out (chromosome padding before each contig, FASTA.cpp:362-391)."""
        cs, p = [], 0
        for _, ln in self.refs:
            p += padding; cs.append(p); p += ln
        return np.array(cs, dtype=np.int64)


def make_cases(d: str, contigs, idx: str, aligner: str):
    out = {}
    for dense in (False, True):
        sub = os.path.join(d, "dense" if dense else "spread"); os.makedirs(sub, exist_ok=True)
        write_dup_fastqs(sub, contigs, seed=5 if dense else 3, dense=dense)
        t = "_dense" if dense else ""
        out["single" + t] = SortedCase(sub, "se", aligner, ["single", idx, os.path.join(sub, "se.fq"), "-d", "14"])
        out["paired" + t] = SortedCase(sub, "pe", aligner, ["paired", idx, os.path.join(sub, "p1.fq"), os.path.join(sub, "p2.fq")])
    return out
