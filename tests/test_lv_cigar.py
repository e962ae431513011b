"""Output stage, first piece: LandauVishkinWithCigar restated in snap_b200/csrc/sg_lv_cigar.h.

Host-side only so far (the algorithm header built for the host, like the other sg_*.h headers): (1) the compiled reference and
the restatement both reproduce the 30 CIGAR known answers of the reference's own tests (tests/LandauVishkinTest.cpp:34-129);
(2) the restatement equals the compiled reference on random problems: operations, edit distance, text used, net indel,
and the front-clipping verdict of computeEditDistanceNormalized."""
import numpy as np
import pytest

import hostsim_lib as hs

# (text, pattern, k, cigar with =/X, cigar with M) -- reference tests/LandauVishkinTest.cpp:38-128
KNOWN = [
    ("abcde", "abcde", 2, "5=", "5M"),
    ("abcdef", "abcde", 2, "5=", "5M"),
    ("abcde", "abcdX", 2, "4=1X", "5M"),
    ("abcde", "Xbcde", 2, "1X4=", "5M"),
    ("abcde", "abde", 2, "2=1D2=", "2M1D2M"),
    ("abcde", "bcde", 2, "1D4=", "1D4M"),
    ("abcde", "abcXde", 2, "3=1I2=", "3M1I2M"),
    ("abcde", "abXXe", 2, "2=2X1=", "5M"),
    ("abcde", "abcXXde", 3, "3=2I2=", "3M2I2M"),
    ("ttttc", "tttc", 3, "3=1X", "4M"),
    ("tttcc", "ttttc", 3, "3=1X1=", "5M"),
    ("tttcc", "tttaa", 3, "3=2X", "5M"),
    ("atctcag", "acttcag", 3, "1=2X4=", "7M"),
    ("abc", "abcde", 3, "3=2X", "5M"),
    ("abc", "abXde", 3, "2=3X", "5M"),
]


def _known_jobs(reflib):
    # C string literals: NUL after the last character, as in the reference's test (the routine looks one character past the ends)
    text = bytearray(); pat = bytearray(); jobs = []
    for t, p, k, _, _ in KNOWN:
        for use_m in (0, 1):
            jobs.append((len(text), len(pat), len(t), len(p), k, use_m))
        text += t.encode() + b"\0" * 16
        pat += p.encode() + b"\0" * 16
    return (np.frombuffer(bytes(text), dtype=np.uint8).copy(), np.frombuffer(bytes(pat), dtype=np.uint8).copy(),
            np.array(jobs, dtype=reflib.LVC_JOB_DTYPE))


@pytest.mark.parametrize("impl", ["reference", "restatement"])
def test_known_cigars(reflib, impl):
    text, pat, jobs = _known_jobs(reflib)
    out = reflib.lv_cigar_batch(text, pat, jobs) if impl == "reference" else hs.lv_cigar_batch(text, pat, jobs, reflib.LVC_OUT_DTYPE)
    i = 0
    for t, p, k, eqx, m in KNOWN:
        for want in (eqx, m):
            assert out[i]["score"] >= 0, (t, p)
            assert reflib.decode_cigar(out[i]["ops"], int(out[i]["nOps"])) == want, (t, p, want)
            i += 1


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
