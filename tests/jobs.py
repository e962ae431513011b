"""Random leaf-kernel jobs (LV / affine gap) shaped like the call sites in BaseAligner::score
(reference SNAPLib/BaseAligner.cpp:1160-1276): tail = forward text after the seed, head = reversed read against
text walked backwards.  Shared by the CPU (hostsim vs reference) and GPU (CUDA vs reference / golden) tests."""
from __future__ import annotations

import numpy as np

LV_JOB = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"), ("k", "<i4"), ("dir", "<i4")])
LV_OUT = np.dtype([("score", "<i4"), ("netIndel", "<i4"), ("totalIndels", "<i4"), ("textSpan", "<i4"), ("matchProbability", "<f8")])
AG_JOB = np.dtype([("textOff", "<u8"), ("patOff", "<u8"), ("textLen", "<i4"), ("patternLen", "<i4"), ("w", "<i4"), ("scoreInit", "<i4"),
                   ("dir", "<i4"), ("isRC", "<i4"), ("banded", "<i4"), ("useClippingOptimizations", "<i4")])
AG_OUT = np.dtype([("agScore", "<i4"), ("textOffset", "<i4"), ("patternOffset", "<i4"), ("nEdits", "<i4"), ("matchProbability", "<f8")])
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SLACK = 200


def _mutate(rng, seq, sub, ins, dele, n_rate=0.0):
    out = []
    for b in seq:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + ins:
            out.append(int(ACGT[rng.integers(0, 4)]))
        if r < dele + ins + sub:
            out.append(int(ACGT[rng.integers(0, 4)]))
        elif rng.random() < n_rate:
            out.append(ord("N"))
        else:
            out.append(int(b))
    return np.array(out, dtype=np.uint8)


def make_pairs(n: int, seed: int, max_len: int = 150, err_scale: float = 1.0, low_complexity: float = 0.2):
    """Returns (textBuf, patBuf, qualBuf, list of (textOff, patOff, patLen, dir)).  For dir=+1 the text starts at
    textOff; for dir=-1 textOff is one past the first text character (the reference's convention)."""
    rng = np.random.default_rng(seed)
    texts, pats, quals, meta = [], [], [], []
    toff = SLACK
    poff = SLACK
    for _ in range(n):
        L = int(rng.integers(1, max_len + 1))
        span = L + 140
        if rng.random() < low_complexity:
            unit = ACGT[rng.integers(0, 4, size=int(rng.integers(1, 5)))]
            region = np.tile(unit, span // unit.size + 1)[:span].copy()
            flips = rng.random(span) < 0.05
            region[flips] = ACGT[rng.integers(0, 4, size=int(flips.sum()))]
        else:
            region = ACGT[rng.integers(0, 4, size=span)]
        if rng.random() < 0.05:
            region[int(rng.integers(0, span))] = ord("N")
        if rng.random() < 0.03:
            region[int(rng.integers(0, span)):] = ord("n")      # contig padding
        d = 1 if rng.random() < 0.5 else -1
        e = rng.random() ** 2 * 0.12 * err_scale
        sub, ins, dele = e * 0.6, e * 0.2, e * 0.2
        if d == 1:
            src = region[:L + 20]
        else:
            src = region[::-1][:L + 20]
        pat = _mutate(rng, src, sub, ins, dele, n_rate=0.002)[:L]
        if pat.size < L:
            pat = np.concatenate([pat, ACGT[rng.integers(0, 4, size=L - pat.size)]])
        if rng.random() < 0.1:
            pat = ACGT[rng.integers(0, 4, size=L)]              # unrelated
        q = (rng.integers(2, 41, size=L) + 33).astype(np.uint8)
        if rng.random() < 0.3:
            q[:] = rng.integers(66, 74)                          # high-quality run (clipping heuristics)
        texts.append(region)
        pats.append(pat)
        quals.append(q)
        meta.append((toff if d == 1 else toff + span, poff, L, d, span))
        toff += span + 16
        poff += L + 16
    textBuf = np.full(toff + SLACK, ord("n"), dtype=np.uint8)
    patBuf = np.full(poff + SLACK, ord("A"), dtype=np.uint8)
    qualBuf = np.full(poff + SLACK, ord("5"), dtype=np.uint8)
    t = SLACK
    p = SLACK
    for region, pat, q in zip(texts, pats, quals):
        textBuf[t:t + region.size] = region
        patBuf[p:p + pat.size] = pat
        qualBuf[p:p + q.size] = q
        t += region.size + 16
        p += pat.size + 16
    return textBuf, patBuf, qualBuf, meta


def lv_jobs(n: int, seed: int, **kw):
    textBuf, patBuf, qualBuf, meta = make_pairs(n, seed, **kw)
    rng = np.random.default_rng(seed + 1)
    jobs = np.zeros(n, dtype=LV_JOB)
    for i, (toff, poff, L, d, span) in enumerate(meta):
        jobs[i]["textOff"] = toff
        jobs[i]["patOff"] = poff
        jobs[i]["patternLen"] = L
        jobs[i]["textLen"] = L + 127 if rng.random() < 0.8 else int(rng.integers(max(1, L - 5), L + 130))
        jobs[i]["textLen"] = min(int(jobs[i]["textLen"]), span)
        jobs[i]["k"] = int(rng.integers(-1, 31)) if rng.random() < 0.9 else int(rng.integers(30, 127))
        jobs[i]["dir"] = d
    return textBuf, patBuf, qualBuf, jobs


def ag_jobs(n: int, seed: int, **kw):
    textBuf, patBuf, qualBuf, meta = make_pairs(n, seed, **kw)
    rng = np.random.default_rng(seed + 2)
    jobs = np.zeros(n, dtype=AG_JOB)
    for i, (toff, poff, L, d, span) in enumerate(meta):
        w = int(rng.integers(0, 31)) if rng.random() < 0.9 else int(rng.integers(30, 127))
        if rng.random() < 0.02:
            w = -1
        banded = 1 if (L >= 3 * (2 * w + 1) and w >= 0 and rng.random() < 0.9) else 0
        jobs[i]["textOff"] = toff
        jobs[i]["patOff"] = poff
        jobs[i]["patternLen"] = L
        if d == 1:
            tl = L + 127
        else:
            tl = L + max(w, 0)
        jobs[i]["textLen"] = min(tl, span - 4)
        jobs[i]["w"] = w
        jobs[i]["scoreInit"] = L + int(rng.integers(20, 131))
        jobs[i]["dir"] = d
        jobs[i]["isRC"] = int(rng.integers(0, 2))
        jobs[i]["banded"] = banded
        jobs[i]["useClippingOptimizations"] = int(rng.random() < 0.3)
    return textBuf, patBuf, qualBuf, jobs


def same_out(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Boolean mask of records whose bytes are identical (doubles compared bitwise)."""
    return np.array([x.tobytes() == y.tobytes() for x, y in zip(a, b)])
