"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host-simulated engine headers vs
the compiled reference, C-ABI surface, gloo sharding.  `-m gpu` runs on a B200 and goes through the C ABI."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reflib():
    from oracle import reflib as r
    if not r.available():
        pytest.skip("oracle/_ref/libsnapref.so not built (needs /root/reference once: make -C oracle ref)")
    return r


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(HERE, "golden")


def _clipped_pairs(synth, contigs, n, seed):
    """Pairs where ~45% of the ends carry an adapter-like junk head / tail (20-110 bases) or a burst of substitutions:
    the inputs the soft-clipping (Hamming / gapless) passes exist for."""
    rng = np.random.default_rng(seed)
    base = synth.make_pairs(contigs, n, 150, seed=seed, sub_rate=0.01)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(2 * n):
        b, q = base.read(i)
        b = bytearray(b)
        r = rng.random()
        if r < 0.35:
            k = int(rng.integers(20, 110))
            junk = bytes(rng.choice(acgt, size=k))
            if rng.random() < 0.5:
                b[150 - k:] = junk
            else:
                b[:k] = junk
        elif r < 0.45:
            for _ in range(int(rng.integers(15, 40))):
                b[int(rng.integers(0, 150))] = b"ACGT"[int(rng.integers(0, 4))]
        out.append((bytes(b), q))
    return synth.ReadBatch.from_lists(out)


class SmallCfg:
    """A small repeat-bearing reference with a reference-built index directory (default + -large) and read sets."""

    def __init__(self, tmp, reflib):
        from snap_b200 import synth
        self.dir = str(tmp)
        self.contigs = synth.make_contigs(3, 120_000, seed=31, repeat_frac=0.25)
        self.fasta = os.path.join(self.dir, "ref.fa")
        synth.write_fasta(self.fasta, self.contigs)
        self.idx = os.path.join(self.dir, "idx")
        self.idx_large = os.path.join(self.dir, "idxL")
        reflib.build_reference_index(reflib.SNAP_ALIGNER, self.fasta, self.idx)
        reflib.build_reference_index(reflib.SNAP_ALIGNER, self.fasta, self.idx_large, large=True)
        self.reads = {
            "std150": synth.make_reads(self.contigs, 1500, 150, seed=32),
            "noisy150": synth.make_reads(self.contigs, 1500, 150, seed=33, sub_rate=0.04, ins_rate=0.006, del_rate=0.006,
                                         n_run_frac=0.1, short_frac=0.1, random_frac=0.05),
            "indel100": synth.make_reads(self.contigs, 1000, 100, seed=34, sub_rate=0.01, ins_rate=0.01, del_rate=0.01),
            "long250": synth.make_reads(self.contigs, 500, 250, seed=35, sub_rate=0.02, ins_rate=0.002, del_rate=0.002),
        }

        self.pairs = {
            "clipped150": _clipped_pairs(synth, self.contigs, 500, 45),
            "std150": synth.make_pairs(self.contigs, 600, 150, seed=41, chimeric_frac=0.03, n_run_frac=0.03, short_frac=0.03),
            "noisy150": synth.make_pairs(self.contigs, 600, 150, seed=42, sub_rate=0.04, ins_rate=0.004, del_rate=0.004,
                                         chimeric_frac=0.05, n_run_frac=0.05, short_frac=0.05),
            "len100": synth.make_pairs(self.contigs, 400, 100, seed=43, sub_rate=0.02, ins_rate=0.002, del_rate=0.002),
            "len250": synth.make_pairs(self.contigs, 300, 250, seed=44, sub_rate=0.02, ins_rate=0.003, del_rate=0.003, insert_mean=500),
        }

    def padded_bases(self):
        """Genome as SNAP lays it out: 2000 'n' before each contig and at the end (FASTA.cpp:362-391)."""
        parts, starts, pos = [], [], 0
        for c in self.contigs:
            parts.append(np.full(2000, ord("n"), dtype=np.uint8)); pos += 2000
            starts.append(pos)
            parts.append(c); pos += c.size
        parts.append(np.full(2000, ord("n"), dtype=np.uint8))
        return np.concatenate(parts), np.array(starts, dtype=np.int64)


@pytest.fixture(scope="session")
def small_cfg(tmp_path_factory, reflib):
    return SmallCfg(tmp_path_factory.mktemp("smallcfg"), reflib)


def differing(want: np.ndarray, got: np.ndarray) -> list[int]:
    """Indices whose result records differ bytewise (doubles bit-for-bit).  Reads both sides report NotFound for are
    equal regardless of the remaining fields (the reference leaves them uninitialised on its early returns,
    BaseAligner.cpp:360-366, :398-402)."""
    out = []
    for i in range(len(want)):
        if want[i]["status"] == 0 and got[i]["status"] == 0:
            continue
        if want[i].tobytes() != got[i].tobytes():
            out.append(i)
    return out


OPTION_SETS = {
    "default_d14": dict(maxDist=14),
    "noag_d14": dict(maxDist=14, useAffineGap=0),
    "ne_d20": dict(maxDist=20, noEditDistance=1, useAffineGap=0),
    "d8_h20": dict(maxDist=8, maxHits=20),
    "coverage": dict(maxDist=14, numSeedsFromCommandLine=0, seedCoverage=4.0),
    "esd3_ms2": dict(maxDist=14, extraSearchDepth=3, minWeightToCheck=2),
    "nobanded": dict(maxDist=14, noBandedAffineGap=1),
    "stopfirst": dict(maxDist=14, stopOnFirstHit=1),
    "ag_d20": dict(maxDist=20, useAffineGap=1),          # BASELINE configs[3]: `snap single -G -d 20`
    "d8": dict(maxDist=8),                                # BASELINE configs[0]
}


def differing_pairs(want: np.ndarray, got: np.ndarray) -> list[int]:
    """Indices of pairs whose result records differ (doubles bit-for-bit).  mapq / scorePriorToClipping of an end that is
    reported NotFound are excluded: ChimericPairedEndAligner copies them from a stack SingleAlignmentResult that
    BaseAligner::AlignRead leaves unwritten on its NotFound paths (ChimericPairedEndAligner.cpp:268, :417-426), i.e. the
    reference's own values there are whatever the previous pair left on the stack."""
    w, g = want.copy(), got.copy()
    for arr in (w, g):
        nf = arr["status"] == 0
        arr["mapq"][nf] = 0
        arr["scorePriorToClipping"][nf] = 0
    return [i for i in range(len(w)) if w[i].tobytes() != g[i].tobytes()]


# `snap paired` defaults (soft clipping on: the Hamming / gapless passes run) and `snap paired -hc` style option sets (soft clipping off: bonuses 5/5, minAGScoreImprovement 15, PairedAligner.cpp:380-392)
_HC = dict(fivePrimeEndBonus=5, threePrimeEndBonus=5)
_HCP = dict(useSoftClipping=0, minAGScoreImprovement=15)
PAIRED_OPTION_SETS = {
    "default_d14": (dict(maxDist=14), dict()),
    "default_d27": (dict(maxDist=27), dict()),          # `snap paired` really defaults to -d 27
    "default_coverage": (dict(maxDist=14, numSeedsFromCommandLine=0, seedCoverage=2.0), dict()),
    "default_no_eh": (dict(maxDist=14), dict(enableHammingScoringBaseAligner=0)),
    "default_noopt": (dict(maxDist=10, noUkkonen=1, noOrderedEvaluation=1, noTruncation=1), dict()),
    "default_spacing_300_450": (dict(maxDist=14), dict(minSpacing=300, maxSpacing=450)),
    "hc_d14": (dict(maxDist=14, **_HC), dict(**_HCP)),
    "hc_d27": (dict(maxDist=27, **_HC), dict(**_HCP)),
    "hc_h20_H50": (dict(maxDist=14, maxHits=20, **_HC), dict(intersectingAlignerMaxHits=50, **_HCP)),
    "hc_noag": (dict(maxDist=14, useAffineGap=0), dict(**_HCP)),
    "hc_forcespacing": (dict(maxDist=14, **_HC), dict(forceSpacing=1, **_HCP)),
    "hc_spacing_300_450": (dict(maxDist=14, **_HC), dict(minSpacing=300, maxSpacing=450, **_HCP)),
    "hc_nobanded_esd4": (dict(maxDist=14, noBandedAffineGap=1, extraSearchDepth=4, **_HC), dict(maxDistForIndels=20, **_HCP)),
}
