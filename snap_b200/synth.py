"""Deterministic synthetic references and reads (SURVEY.md 8d "Synthetic inputs").

CPU/numpy generator used by the tests, the golden-fixture script and the small bench configs.
Reference: uniform i.i.d. ACGT contigs.  Reads: uniform start, 50% reverse-complemented, per-base
substitution / insertion / deletion, qualities uniform Phred 20-40 ('5'..'I'), optional N runs and
'#'-quality tails (the reader would clip those: FASTQ.cpp:294, Read.h:567-619 -- we emit the clipped view).
"""
from __future__ import annotations

import os
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def make_contigs(n_contigs: int, contig_len: int, seed: int, repeat_frac: float = 0.0) -> list[np.ndarray]:
    """Uniform random contigs; with repeat_frac>0, that fraction of bases comes from a small repeat
    library at 0-5% divergence (exercises overflow lists / popular seeds / merge logic)."""
    rng = np.random.default_rng(seed)
    contigs = []
    lib = None
    if repeat_frac > 0:
        lib = [ACGT[rng.integers(0, 4, size=int(rng.integers(200, 2000)))] for _ in range(8)]
    for _ in range(n_contigs):
        c = ACGT[rng.integers(0, 4, size=contig_len)]
        if lib is not None:
            placed = 0
            while placed < repeat_frac * contig_len:
                unit = lib[int(rng.integers(0, len(lib)))].copy()
                div = rng.random() * 0.05
                mut = rng.random(unit.size) < div
                unit[mut] = ACGT[rng.integers(0, 4, size=int(mut.sum()))]
                pos = int(rng.integers(0, max(1, contig_len - unit.size)))
                n = min(unit.size, contig_len - pos)
                c[pos:pos + n] = unit[:n]
                placed += n
        contigs.append(c)
    return contigs


def write_fasta(path: str, contigs: list[np.ndarray], width: int = 100) -> None:
    with open(path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">chr%d\n" % (i + 1))
            n = c.size
            full = (n // width) * width
            if full:
                body = c[:full].reshape(-1, width)
                out = np.empty((body.shape[0], width + 1), dtype=np.uint8)
                out[:, :width] = body
                out[:, width] = ord("\n")
                f.write(out.tobytes())
            if n > full:
                f.write(c[full:].tobytes() + b"\n")


class ReadBatch:
    """Concatenated reads in the layout the C ABI takes (include/snapgpu.h snapgpu_align_single)."""

    def __init__(self, bases: np.ndarray, quals: np.ndarray, offsets: np.ndarray, lens: np.ndarray,
                 truth_contig=None, truth_pos=None, truth_rc=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.quals = np.ascontiguousarray(quals, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.lens = np.ascontiguousarray(lens, dtype=np.uint32)
        self.truth_contig = truth_contig
        self.truth_pos = truth_pos
        self.truth_rc = truth_rc

    @property
    def n(self) -> int:
        return int(self.lens.size)

    def read(self, i: int) -> tuple[bytes, bytes]:
        o, l = int(self.offsets[i]), int(self.lens[i])
        return self.bases[o:o + l].tobytes(), self.quals[o:o + l].tobytes()

    def slice(self, lo: int, hi: int) -> "ReadBatch":
        o0 = int(self.offsets[lo])
        o1 = int(self.offsets[hi - 1]) + int(self.lens[hi - 1]) if hi > lo else o0
        return ReadBatch(self.bases[o0:o1], self.quals[o0:o1], self.offsets[lo:hi] - np.uint64(o0), self.lens[lo:hi])

    def write_fastq(self, path: str) -> None:
        with open(path, "wb") as f:
            for i in range(self.n):
                b, q = self.read(i)
                f.write(b"@r%d\n%s\n+\n%s\n" % (i, b, q))

    @staticmethod
    def from_lists(reads: list[tuple[bytes, bytes]]) -> "ReadBatch":
        lens = np.array([len(b) for b, _ in reads], dtype=np.uint32)
        offsets = np.zeros(len(reads), dtype=np.uint64)
        if len(reads) > 1:
            offsets[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        bases = np.frombuffer(b"".join(b for b, _ in reads), dtype=np.uint8)
        quals = np.frombuffer(b"".join(q for _, q in reads), dtype=np.uint8)
        return ReadBatch(bases, quals, offsets, lens)


def make_reads(contigs: list[np.ndarray], n: int, read_len: int, seed: int, sub_rate: float = 0.01,
               ins_rate: float = 0.0005, del_rate: float = 0.0005, rc_frac: float = 0.5,
               n_run_frac: float = 0.0, short_frac: float = 0.0, random_frac: float = 0.0) -> ReadBatch:
    """n reads of read_len bases.  n_run_frac: fraction of reads given a run of 1-12 'N's;
    short_frac: fraction emitted at a shorter length (as if '#'-clipped); random_frac: fraction of
    reads that are pure noise (unalignable)."""
    rng = np.random.default_rng(seed)
    nc = len(contigs)
    clen = np.array([c.size for c in contigs])
    margin = read_len + 64
    reads = []
    t_contig = np.zeros(n, dtype=np.int32)
    t_pos = np.zeros(n, dtype=np.int64)
    t_rc = np.zeros(n, dtype=np.int8)
    for i in range(n):
        ci = int(rng.integers(0, nc))
        pos = int(rng.integers(0, max(1, clen[ci] - margin)))
        window = contigs[ci][pos:pos + margin]
        if rng.random() < random_frac:
            seq = ACGT[rng.integers(0, 4, size=read_len)]
        else:
            # walk the window applying edits until read_len bases are produced
            r = rng.random(margin)
            ops_sub = r < sub_rate
            ops_ins = (r >= sub_rate) & (r < sub_rate + ins_rate)
            ops_del = (r >= sub_rate + ins_rate) & (r < sub_rate + ins_rate + del_rate)
            if not ops_ins.any() and not ops_del.any():
                seq = window[:read_len].copy()
                m = ops_sub[:read_len]
                k = int(m.sum())
                if k:
                    seq[m] = ACGT[(np.searchsorted(ACGT, seq[m]) + rng.integers(1, 4, size=k)) % 4]
            else:
                out = []
                j = 0
                while len(out) < read_len and j < margin:
                    if ops_del[j]:
                        j += 1
                        continue
                    if ops_ins[j]:
                        out.append(int(ACGT[rng.integers(0, 4)]))
                        if len(out) >= read_len:
                            break
                    b = int(window[j])
                    if ops_sub[j]:
                        b = int(ACGT[(int(np.searchsorted(ACGT, b)) + int(rng.integers(1, 4))) % 4])
                    out.append(b)
                    j += 1
                seq = np.array(out[:read_len], dtype=np.uint8)
                if seq.size < read_len:
                    seq = np.concatenate([seq, ACGT[rng.integers(0, 4, size=read_len - seq.size)]])
        rc = rng.random() < rc_frac
        if rc:
            seq = revcomp(seq)
        if rng.random() < n_run_frac:
            run = int(rng.integers(1, 13))
            at = int(rng.integers(0, read_len - run))
            seq = seq.copy()
            seq[at:at + run] = ord("N")
        L = read_len
        if rng.random() < short_frac:
            L = int(rng.integers(20, read_len))
        qual = rng.integers(20, 41, size=L).astype(np.uint8) + 33
        reads.append((seq[:L].tobytes(), qual.tobytes()))
        t_contig[i], t_pos[i], t_rc[i] = ci, pos, rc
    rb = ReadBatch.from_lists(reads)
    rb.truth_contig, rb.truth_pos, rb.truth_rc = t_contig, t_pos, t_rc
    return rb


def make_pairs(contigs: list[np.ndarray], n_pairs: int, read_len: int, seed: int, insert_mean: float = 400.0, insert_sd: float = 40.0,
               sub_rate: float = 0.01, ins_rate: float = 0.0005, del_rate: float = 0.0005, chimeric_frac: float = 0.0,
               n_run_frac: float = 0.0, short_frac: float = 0.0) -> ReadBatch:
    """FR pairs (insert N(400,40)): read 2i = forward strand of the fragment start, read 2i+1 = reverse
    complement of the fragment end; half the fragments are flipped.  chimeric_frac: mate drawn from an unrelated place."""
    rng = np.random.default_rng(seed)
    nc = len(contigs)
    clen = np.array([c.size for c in contigs])
    reads = []

    def mutate(window):
        out = []
        j = 0
        while len(out) < read_len and j < window.size:
            r = rng.random()
            if r < del_rate:
                j += 1
                continue
            if r < del_rate + ins_rate:
                out.append(int(ACGT[rng.integers(0, 4)]))
                if len(out) >= read_len:
                    break
            b = int(window[j])
            if r >= del_rate + ins_rate and r < del_rate + ins_rate + sub_rate:
                b = int(ACGT[(int(np.searchsorted(ACGT, b)) + int(rng.integers(1, 4))) % 4])
            out.append(b)
            j += 1
        seq = np.array(out[:read_len], dtype=np.uint8)
        if seq.size < read_len:
            seq = np.concatenate([seq, ACGT[rng.integers(0, 4, size=read_len - seq.size)]])
        return seq

    for _ in range(n_pairs):
        ci = int(rng.integers(0, nc))
        ins = int(max(read_len + 10, rng.normal(insert_mean, insert_sd)))
        pos = int(rng.integers(0, max(1, clen[ci] - ins - 64)))
        frag = contigs[ci][pos:pos + ins + 48]
        r1 = mutate(frag[:read_len + 32])
        tail = frag[max(0, ins - read_len - 32):ins]
        r2 = mutate(revcomp(tail))
        if rng.random() < chimeric_frac:
            cj = int(rng.integers(0, nc))
            p2 = int(rng.integers(0, max(1, clen[cj] - read_len - 64)))
            r2 = mutate(revcomp(contigs[cj][p2:p2 + read_len + 32]))
        if rng.random() < 0.5:
            r1, r2 = r2, r1
        pair = []
        for seq in (r1, r2):
            if rng.random() < n_run_frac:
                run = int(rng.integers(1, 13))
                at = int(rng.integers(0, read_len - run))
                seq = seq.copy()
                seq[at:at + run] = ord("N")
            L = read_len
            if rng.random() < short_frac:
                L = int(rng.integers(20, read_len))
            qual = (rng.integers(20, 41, size=L) + 33).astype(np.uint8)
            pair.append((seq[:L].tobytes(), qual.tobytes()))
        reads.extend(pair)
    return ReadBatch.from_lists(reads)


def make_fastq_text(batch: ReadBatch, seed: int = 0, crlf_frac: float = 0.0, lower_frac: float = 0.0, dot_frac: float = 0.0, hash_tail_frac: float = 0.0,
                    hash_head_frac: float = 0.0, comment_frac: float = 0.0, plus_id_frac: float = 0.0, truncate_last: bool = False) -> np.ndarray:
    """FASTQ text (uint8 array) for the reads of `batch`, decorated with the things FASTQReader / Read::clip must cope with:
    CRLF line ends, lower-case bases and '.', '#'-quality heads and tails (clipped by the reader), a comment after the id,
    the id repeated on the '+' line, and optionally a last record cut in the middle."""
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(batch.n):
        b, q = batch.read(i)
        b = bytearray(b); q = bytearray(q)
        L = len(b)
        if L and rng.random() < lower_frac:
            for k in rng.integers(0, L, size=max(1, L // 10)):
                b[int(k)] = ord(chr(b[int(k)]).lower())
        if L and rng.random() < dot_frac:
            b[int(rng.integers(0, L))] = ord(".")
        if L and rng.random() < hash_tail_frac:
            k = int(rng.integers(1, min(L, 40) + 1)); q[L - k:] = b"#" * k
        if L and rng.random() < hash_head_frac:
            k = int(rng.integers(1, min(L, 20) + 1)); q[:k] = b"#" * k
        name = b"read%d" % i
        line0 = b"@" + name + (b" 1:N:0:ACGT extra" if rng.random() < comment_frac else b"")
        line2 = b"+" + (name if rng.random() < plus_id_frac else b"")
        eol = b"\r\n" if rng.random() < crlf_frac else b"\n"
        parts.append(line0 + eol + bytes(b) + eol + line2 + eol + bytes(q) + eol)
    text = b"".join(parts)
    if truncate_last and parts:
        text = text[:len(text) - len(parts[-1]) // 2]
    return np.frombuffer(text, dtype=np.uint8).copy()
