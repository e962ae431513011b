"""Row N4 after the sort -- duplicate marking (BAMDupMarkFilter) and the BAM index (BAMIndexSupplier) -- on the host build of
snap_b200/csrc/sg_bampost.h, against the files the reference binary writes with `-so`."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
import sorted_data


@pytest.fixture(scope="module")
def cases(tmp_path_factory, small_cfg, reflib):
    return sorted_data.make_cases(str(tmp_path_factory.mktemp("sorted")), small_cfg.contigs, small_cfg.idx, reflib.SNAP_ALIGNER)


@pytest.fixture(scope="module")
def hs():
    L = C.CDLL(hostsim_lib.build())
    L.hs_bam_markdup.restype = C.c_int64
    L.hs_bam_index.restype = C.c_int64
    return L


@pytest.mark.parametrize("name", ["single", "paired", "single_dense", "paired_dense"])
def test_duplicates_marked_like_the_reference(cases, hs, name):
    """The reference's sorted, unmarked records in; its sorted, marked records out -- the whole stream byte for byte (a third of the fragments
    come in several copies; forward and reverse strands, copies of different lengths, pairs that lost a mate, three read-name styles)."""
    c = cases[name]
    assert len(c.unmarked) == len(c.marked) and sum(1 for r in c.marked if r[19] & 4) > len(c.marked) // 5
    blob = bytearray(b"".join(c.unmarked))
    cs = c.contig_starts()
    buf = (C.c_uint8 * len(blob)).from_buffer(blob)
    marked = hs.hs_bam_markdup(buf, C.c_int64(len(blob)), cs.ctypes.data_as(C.c_void_p), C.c_int32(len(cs)))
    assert marked == sum(1 for r in c.marked if r[19] & 4)
    assert bytes(blob) == b"".join(c.marked)
    # idempotent: marked records in, nothing more marked
    assert hs.hs_bam_markdup(buf, C.c_int64(len(blob)), cs.ctypes.data_as(C.c_void_p), C.c_int32(len(cs))) == 0


@pytest.mark.parametrize("name", ["single", "paired", "single_dense", "paired_dense"])
def test_bam_index_equals_the_reference_index(cases, hs, name):
    """.bai of header ‖ records: the same bins, chunks, mapped / unmapped counts and linear index as the reference's .bam.bai once both files'
    virtual offsets are turned back into uncompressed offsets (the reference deflates, we store: the block boundaries differ, the content does not)."""
    c = cases[name]
    blob = b"".join(c.marked)
    bai = (C.c_uint8 * (1 << 22))()
    n = hs.hs_bam_index(blob, C.c_int64(len(blob)), C.c_int64(c.header_bytes), C.c_int32(len(c.refs)), bai, C.c_int64(1 << 22))
    assert n > 8
    got = sorted_data.parse_bai(bytes(bai[:n]), sorted_data.our_blocks(c.header_bytes + len(blob)))
    want = sorted_data.parse_bai(c.bai, sorted_data.bgzf_blocks(c.bam))
    assert got == want
    assert sum(len(b) for b, _ in want) > 3 * len(c.refs) or "dense" in name
