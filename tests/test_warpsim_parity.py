"""The warp-cooperative affine-gap forms (snap_b200/csrc/sg_warp_ag*.cuh: shuffles, ballots, DPX s16x2) run on the 32-lane SIMT
emulator of tests/hostsim/warpsim.h, against the compiled reference: the code the alignment kernels run, checked without a GPU.
(The same comparisons run on the device in tests/test_gpu_parity.py.)"""
import numpy as np
import pytest

import hostsim_lib as hs
import jobs as J
import warpsim_lib as ws


def _check(reflib, t, p, q, jb, packed):
    want = reflib.ag_batch(t, p, q, jb.astype(reflib.AG_JOB_DTYPE))
    _, stale = hs.ag_batch(t, p, q, jb, J.AG_OUT, reflib.AG_PARAMS_DEFAULT)        # history-dependent jobs (see test_ag_fuzz)
    got = ws.ag_batch(t, p, q, jb, J.AG_OUT, reflib.AG_PARAMS_DEFAULT, packed)
    same = J.same_out(want, got)
    bad = np.nonzero(~same & (stale == 0))[0]
    assert bad.size == 0, (packed, jb[bad[0]], want[bad[0]], got[bad[0]])
    for f in ("agScore", "textOffset", "patternOffset"):
        assert (want[f] == got[f]).all(), f
    return stale


@pytest.mark.parametrize("packed", [1, 2, 5, 13, 0])
def test_warp_ag_forms_fuzz(reflib, packed):
    """1: packed rows (unbanded) + the one-cell-per-lane narrow-band form -- what the alignment kernels run; 2: the unrolled packed
    instantiation; 5: the experimental two-units-per-step narrow-band form (sg_warp_ag_duo.cuh; 13: its per-round H in the arena instead
    of the shared-memory block); 0: the general int form."""
    t, p, q, jb = J.ag_jobs(1000, 311 + packed)
    stale = _check(reflib, t, p, q, jb, packed)
    assert (stale != 0).sum() <= 0.01 * jb.size
    assert ((jb["banded"] != 0) & (jb["w"] <= 15) & (jb["w"] >= 0)).sum() > 200


@pytest.mark.parametrize("seed,packed", [(1, 1), (2, 1), (3, 5), (4, 5)])
def test_narrow_band_fuzz(reflib, seed, packed):
    """Every job banded with w <= 15 (numVec 1..4, up to 13 segments), among them hopeless ones (tiny scoreInit: the row loop ends on
    an all-zero row while the next row's first unit is already in flight) and texts cut short."""
    t, p, q, jb = J.ag_jobs(1500, 900 + seed)
    rng = np.random.default_rng(seed)
    for i in range(jb.size):
        L = int(jb[i]["patternLen"])
        wmax = min(15, (L // 3 - 1) // 2)
        if wmax < 0:
            jb[i]["banded"] = 0
            continue
        jb[i]["w"] = int(rng.integers(0, wmax + 1)) if rng.random() < 0.5 else wmax
        jb[i]["banded"] = 1
        if rng.random() < 0.15:
            jb[i]["scoreInit"] = int(rng.integers(1, 12))
        if rng.random() < 0.2:
            jb[i]["textLen"] = max(1, int(jb[i]["textLen"]) - int(rng.integers(0, 40)))
    _check(reflib, t, p, q, jb, packed)
    assert (jb["banded"] != 0).sum() > 1200
