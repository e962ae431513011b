"""Synthetic reference and reads generated directly in HBM with torch (bench.py plumbing for the hg-sized configs,
SURVEY 8d: 24 contigs x 125 Mbp uniform ACGT; reads with uniform start, 50% reverse-complemented, 1% substitutions,
0.05% insertions + 0.05% deletions per base (at most one indel event per read here), qualities uniform Phred 20-40).
Deterministic for a given seed and torch version.  Nothing here is on the measured path."""
from __future__ import annotations

import numpy as np
import torch

CHROMOSOME_PADDING = 2000        # FASTA.cpp:362-391: 'n' x padding before each contig and at the end


def make_genome(n_contigs: int, contig_len: int, seed: int, device) -> tuple[torch.Tensor, np.ndarray]:
    """Returns (bases uint8 [nBases] with SNAP's lowercase-'n' contig padding, contig start locations)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    n_bases = n_contigs * (contig_len + CHROMOSOME_PADDING) + CHROMOSOME_PADDING
    bases = torch.full((n_bases,), ord("n"), dtype=torch.uint8, device=device)
    starts = []
    pos = 0
    for _ in range(n_contigs):
        pos += CHROMOSOME_PADDING
        starts.append(pos)
        chunk = 1 << 28
        for o in range(0, contig_len, chunk):
            m = min(chunk, contig_len - o)
            bases[pos + o: pos + o + m] = lut[torch.randint(0, 4, (m,), generator=g, device=device, dtype=torch.int64)]
        pos += contig_len
    return bases, np.array(starts, dtype=np.int64)


def plant_repeats(bases: torch.Tensor, contig_starts: np.ndarray, contig_len: int, frac: float, seed: int, unit_len: int = 10_000,
                  n_units: int = 32, max_div: float = 0.05, tandem_frac: float = 0.1) -> None:
    """SURVEY 8d's stress variant, in place: `frac` of the bases are overwritten with copies of a small library of `unit_len`-bp
    repeat units at 0..max_div divergence per copy (dispersed repeats: overflow lists, maxHits skips, merge logic), a tenth of
    them as short tandem arrays of a 2-50 bp motif."""
    device = bases.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    starts = torch.from_numpy(contig_starts).to(device)
    n_contigs = len(contig_starts)
    lib = lut[torch.randint(0, 4, (n_units, unit_len), generator=g, device=device)]
    n_tandem = max(1, n_units // 8)
    for t in range(n_tandem):                     # tandem units: a short motif repeated over the whole unit
        m = int(torch.randint(2, 51, (1,), generator=g, device=device).item())
        motif = lib[t, :m].clone()
        lib[t] = motif.repeat((unit_len + m - 1) // m)[:unit_len]
    n_copies = int(frac * n_contigs * contig_len / unit_len)
    ar = torch.arange(unit_len, device=device, dtype=torch.int64)[None, :]
    chunk = 2048
    for o in range(0, n_copies, chunk):
        m = min(chunk, n_copies - o)
        tandem = torch.rand((m,), generator=g, device=device) < tandem_frac
        u = torch.where(tandem, torch.randint(0, n_tandem, (m,), generator=g, device=device),
                        torch.randint(n_tandem, n_units, (m,), generator=g, device=device))
        ci = torch.randint(0, n_contigs, (m,), generator=g, device=device)
        pos = torch.randint(0, contig_len - unit_len, (m,), generator=g, device=device)
        div = torch.rand((m, 1), generator=g, device=device) * max_div
        copy = lib[u]
        mut = torch.rand((m, unit_len), generator=g, device=device) < div
        copy = torch.where(mut, lut[torch.randint(0, 4, (m, unit_len), generator=g, device=device)], copy)
        dst = (starts[ci] + pos)[:, None] + ar
        bases[dst.reshape(-1)] = copy.reshape(-1)


_COMP = None


def _comp_table(device):
    global _COMP
    if _COMP is None or _COMP.device != torch.device(device):
        t = torch.full((256,), ord("N"), dtype=torch.uint8)
        for a, b in zip(b"ACGT", b"TGCA"):
            t[a] = b
        _COMP = t.to(device)
    return _COMP


def _mutated_windows(bases, loc, read_len, g, lut, sub_rate, indel_rate):
    """[m, read_len] reads cut at genome locations `loc` (forward strand) with substitutions and at most one indel each."""
    device = bases.device
    m = loc.numel()
    ar = torch.arange(read_len, device=device, dtype=torch.int64)[None, :]
    # at most one indel event per read: P(event) = 1 - (1 - indel_rate)^read_len, half insertions half deletions
    ev = torch.rand((m,), generator=g, device=device) < (1.0 - (1.0 - indel_rate) ** read_len)
    is_ins = torch.rand((m,), generator=g, device=device) < 0.5
    p = torch.randint(1, read_len - 1, (m,), generator=g, device=device)
    shift = torch.zeros((m, read_len), dtype=torch.int64, device=device)
    dele = (ev & ~is_ins)[:, None] & (ar >= p[:, None])
    ins = (ev & is_ins)[:, None] & (ar > p[:, None])
    shift = shift + dele.to(torch.int64) - ins.to(torch.int64)
    win = bases[(loc[:, None] + ar + shift).reshape(-1)].reshape(m, read_len)
    inserted = (ev & is_ins)[:, None] & (ar == p[:, None])
    rnd = lut[torch.randint(0, 4, (m, read_len), generator=g, device=device)]
    win = torch.where(inserted, rnd, win)
    sub = torch.rand((m, read_len), generator=g, device=device) < sub_rate
    # substitute with a *different* base: rotate through ACGT by 1..3
    code = torch.zeros_like(win, dtype=torch.int64)
    code[win == ord("C")] = 1; code[win == ord("G")] = 2; code[win == ord("T")] = 3
    rot = (code + torch.randint(1, 4, (m, read_len), generator=g, device=device)) % 4
    return torch.where(sub, lut[rot], win)


def make_reads(bases: torch.Tensor, contig_starts: np.ndarray, contig_len: int, n: int, read_len: int, seed: int,
               sub_rate: float = 0.01, indel_rate: float = 0.001, rc_frac: float = 0.5):
    """Returns device tensors (bases [n*read_len] u8, quals [n*read_len] u8, offsets [n] u64-as-i64, lens [n] i32->u32)
    plus truth (start location int64 [n], rc bool [n])."""
    device = bases.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    starts = torch.from_numpy(contig_starts).to(device)
    out_b = torch.empty((n, read_len), dtype=torch.uint8, device=device)
    truth_loc = torch.empty((n,), dtype=torch.int64, device=device)
    truth_rc = torch.empty((n,), dtype=torch.bool, device=device)
    chunk = 1 << 20
    for o in range(0, n, chunk):
        m = min(chunk, n - o)
        ci = torch.randint(0, len(contig_starts), (m,), generator=g, device=device)
        pos = torch.randint(0, contig_len - read_len - 8, (m,), generator=g, device=device)
        loc = starts[ci] + pos
        win = _mutated_windows(bases, loc, read_len, g, lut, sub_rate, indel_rate)
        rc = torch.rand((m,), generator=g, device=device) < rc_frac
        rcwin = _comp_table(device)[win.flip(1).to(torch.int64)]
        win = torch.where(rc[:, None], rcwin, win)
        out_b[o:o + m] = win
        truth_loc[o:o + m] = loc
        truth_rc[o:o + m] = rc
    quals = (torch.randint(20, 41, (n, read_len), generator=g, device=device) + 33).to(torch.uint8)
    offsets = (torch.arange(n, device=device, dtype=torch.int64) * read_len)
    lens = torch.full((n,), read_len, dtype=torch.int32, device=device)
    return out_b.reshape(-1), quals.reshape(-1), offsets, lens, truth_loc, truth_rc


def make_pairs(bases: torch.Tensor, contig_starts: np.ndarray, contig_len: int, n_pairs: int, read_len: int, seed: int,
               insert_mean: float = 400.0, insert_sd: float = 40.0, sub_rate: float = 0.01, indel_rate: float = 0.001):
    """FR pairs (insert N(400, 40)): reads 2i / 2i+1 are the two ends of fragment i, one end forward
    from the fragment start, the other the reverse complement of the fragment end; which comes first is a coin flip.
    Returns (bases [2n*read_len], quals, offsets [2n], lens [2n], fragment start location [n], insert [n])."""
    device = bases.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    starts = torch.from_numpy(contig_starts).to(device)
    out_b = torch.empty((n_pairs, 2, read_len), dtype=torch.uint8, device=device)
    truth_loc = torch.empty((n_pairs,), dtype=torch.int64, device=device)
    truth_ins = torch.empty((n_pairs,), dtype=torch.int64, device=device)
    chunk = 1 << 19
    for o in range(0, n_pairs, chunk):
        m = min(chunk, n_pairs - o)
        ci = torch.randint(0, len(contig_starts), (m,), generator=g, device=device)
        ins = (torch.randn((m,), generator=g, device=device) * insert_sd + insert_mean).round().to(torch.int64).clamp(read_len + 10, 900)
        pos = torch.randint(0, contig_len - 1000 - 8, (m,), generator=g, device=device)
        loc = starts[ci] + pos
        left = _mutated_windows(bases, loc, read_len, g, lut, sub_rate, indel_rate)
        right = _mutated_windows(bases, loc + ins - read_len, read_len, g, lut, sub_rate, indel_rate)
        right = _comp_table(device)[right.flip(1).to(torch.int64)]
        flip = torch.rand((m,), generator=g, device=device) < 0.5
        out_b[o:o + m, 0] = torch.where(flip[:, None], right, left)
        out_b[o:o + m, 1] = torch.where(flip[:, None], left, right)
        truth_loc[o:o + m] = loc
        truth_ins[o:o + m] = ins
    n = 2 * n_pairs
    quals = (torch.randint(20, 41, (n, read_len), generator=g, device=device) + 33).to(torch.uint8)
    offsets = (torch.arange(n, device=device, dtype=torch.int64) * read_len)
    lens = torch.full((n,), read_len, dtype=torch.int32, device=device)
    return out_b.reshape(-1), quals.reshape(-1), offsets, lens, truth_loc, truth_ins
