"""AG leaf micro-benchmark: realistic jobs (read tails vs the reference window), warp form, packed vs generic builds."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import jobs as J
from snap_b200 import engine, synth
rng = np.random.default_rng(3)
contigs = synth.make_contigs(1, 400_000, seed=9)
g = contigs[0]
n = int(os.environ.get("AGB_JOBS", "40000"))
w = int(os.environ.get("AGB_W", "29")); banded = int(os.environ.get("AGB_BANDED", "0"))
texts, pats, jb = [], [], np.zeros(n, dtype=J.AG_JOB)
toff = poff = 0
for i in range(n):
    L = int(rng.integers(90, 131))
    pos = int(rng.integers(1000, g.size - 1000))
    pat = g[pos:pos + L].copy()
    for _ in range(int(rng.integers(3, 7))):          # a few substitutions
        pat[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
    if rng.random() < 0.3:                              # one deletion in the read
        k = int(rng.integers(10, L - 10)); pat = np.concatenate([pat[:k], pat[k + 1:], g[pos + L:pos + L + 1]])
    tl = L + 127
    text = g[pos:pos + tl + 8]
    jb[i] = (toff + 64, poff, tl, L, w, 150, 1, 0, banded, 0)
    texts.append(np.full(64, ord('n'), dtype=np.uint8)); texts.append(text); toff += 64 + text.size
    pats.append(pat); pats.append(np.zeros(8, dtype=np.uint8)); poff += L + 8
t = np.concatenate(texts + [np.full(64, ord('n'), dtype=np.uint8)]); p = np.concatenate(pats); q = np.full(p.size, ord('5'), dtype=np.uint8)
out = engine.test_ag(t, p, q, jb, J.AG_OUT, [1, 4, 6, 1, 10, 7], warps=int(os.environ.get("AGB_WARPS", "4096")))
print("lib", os.environ.get("SNAPGPU_LIB", "default"), "banded", banded, "w", w, "mean agScore", out["agScore"].mean(), "nEdits", out["nEdits"].mean(), "checksum", int(out["agScore"].sum()), int(out["nEdits"].sum()))
