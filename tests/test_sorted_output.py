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
    L.hs_bam_index_members.restype = C.c_int64
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


def _members(raw: bytes):
    out, p = [], 0
    while p < len(raw):
        assert raw[p:p + 4] == bytes([31, 139, 8, 4]) and raw[p + 12:p + 14] == b"BC"
        size = (raw[p + 16] | (raw[p + 17] << 8)) + 1
        out.append(raw[p:p + size]); p += size
    return out


def test_deflate_members_inflate_to_the_payload():
    """sg_deflate.h (host build): every member is a gzip member any inflater accepts -- header, BSIZE, CRC-32, ISIZE -- and gives back its slice of the
    payload; incompressible data is stored; low-entropy data gets close to (or beats) zlib -6."""
    import gzip, os, zlib
    rng = np.random.default_rng(1)
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "DESIGN.md"), "rb").read()
    lines = []
    for i in range(3000):
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 150).tobytes()
        q = rng.choice(np.frombuffer(b"FFFFFFF:,#", dtype=np.uint8), 150).tobytes()
        lines.append(b"read%07d\t0\tchr1\t%d\t60\t150M\t*\t0\t0\t%s\t%s\tPG:Z:SNAP\tNM:i:0\tRG:Z:FASTQ\n" % (i, 1000 + i * 37, seq, q))
    payloads = {"empty": b"", "random": rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(), "zeros": bytes(200000),
                "acgt": rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300000).tobytes(), "text": text * 3, "exact": (text * 3)[:0xff00], "plus1": (text * 3)[:0xff00 + 1],
                "sam": b"".join(lines),
                "skewed": bytes(rng.choice(np.arange(256, dtype=np.uint8), 100000, p=(lambda w: w / w.sum())(1.0 / np.arange(1, 257) ** 3)))}
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 100):
        payloads["tiny%d" % n] = bytes(rng.integers(65, 69, n, dtype=np.uint8))
    for name, data in payloads.items():
        z, sizes = hostsim_lib.bgzf_deflate(np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(0, dtype=np.uint8))
        ms = _members(z.tobytes())
        assert [len(m) for m in ms] == [int(x) for x in sizes], name
        assert len(ms) == (len(data) + 0xff00 - 1) // 0xff00, name
        for k, m in enumerate(ms):
            assert gzip.decompress(m) == data[k * 0xff00:(k + 1) * 0xff00], (name, k)
        if name == "random":
            assert len(z) == len(data) + 31 * len(ms)
        if name in ("acgt", "text", "sam", "skewed", "zeros"):
            assert len(z) < 1.3 * len(zlib.compress(data, 6)) + 2000, (name, len(z), len(zlib.compress(data, 6)))


@pytest.mark.parametrize("name", ["single", "paired_dense"])
def test_bam_index_of_the_compressed_file_equals_the_reference_index(cases, hs, name):
    """header ‖ records deflated member by member (sg_deflate.h, host build): the file inflates to the reference file's content and the .bai composed for
    its member offsets says what the reference's .bai says."""
    import gzip
    c = cases[name]
    blob = b"".join(c.marked)
    header = gzip.open(c.bam, "rb").read()[:c.header_bytes]
    z, sizes = hostsim_lib.bgzf_deflate(np.frombuffer(header + blob, dtype=np.uint8).copy())
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    zfile = z.tobytes() + eof
    assert gzip.decompress(zfile) == header + blob and len(zfile) < 0.8 * (len(header) + len(blob))
    offs = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))]).astype(np.uint64)
    bai = (C.c_uint8 * (1 << 22))()
    n = hs.hs_bam_index_members(blob, C.c_int64(len(blob)), C.c_int64(c.header_bytes), C.c_int32(len(c.refs)), offs.ctypes.data_as(C.c_void_p), bai, C.c_int64(1 << 22))
    assert n > 8
    blocks, u = [], 0
    for k, m in enumerate(_members(zfile)):
        isize = int.from_bytes(m[-4:], "little")
        blocks.append((int(offs[k]) if k < len(offs) - 1 else int(offs[-1]), u, isize)); u += isize
    assert sorted_data.parse_bai(bytes(bai[:n]), blocks) == sorted_data.parse_bai(c.bai, sorted_data.bgzf_blocks(c.bam))


def test_deflate_fuzz_and_the_code_length_limit():
    """Structured random payloads of every size class, and one whose byte counts are Fibonacci numbers (a Huffman tree 21 levels deep: the 15-bit
    limit and the Kraft repair must kick in) -- every member inflates to its slice."""
    import gzip
    rng = np.random.default_rng(11)
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    deep = np.concatenate([np.full(c, 40 + k, dtype=np.uint8) for k, c in enumerate(fib)])
    rng.shuffle(deep)
    payloads = [deep.tobytes()]
    for trial in range(60):
        n = int(rng.choice([rng.integers(0, 300), rng.integers(300, 70000), rng.integers(65000, 66000), rng.integers(130000, 200000)]))
        kind = trial % 6
        if kind == 0:
            a = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            a = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), n, p=[0.3, 0.2, 0.2, 0.29, 0.01])
        elif kind == 2:      # runs
            a = np.repeat(rng.integers(0, 256, n // 7 + 1, dtype=np.uint8), rng.integers(1, 600, n // 7 + 1))[:n]
        elif kind == 3:      # repeats of a motif with mutations, periods around the stride and the window
            m = rng.integers(0, 256, int(rng.choice([3, 5, 64, 1024, 1500, 33000])), dtype=np.uint8)
            a = np.tile(m, n // m.size + 1)[:n].copy()
            if n:
                a[rng.integers(0, n, n // 50)] = 0
        elif kind == 4:      # two alphabets glued
            a = np.concatenate([rng.integers(0, 4, n // 2, dtype=np.uint8), rng.integers(0, 256, n - n // 2, dtype=np.uint8)])
        else:
            a = rng.integers(0, 2, n, dtype=np.uint8) * 255
        payloads.append(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
    for k, data in enumerate(payloads):
        z, sizes = hostsim_lib.bgzf_deflate(np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(0, dtype=np.uint8))
        ms = _members(z.tobytes())
        assert len(ms) == (len(data) + 0xff00 - 1) // 0xff00 and [len(m) for m in ms] == [int(x) for x in sizes], k
        assert b"".join(gzip.decompress(m) for m in ms) == data, k
        assert all(len(m) <= 0xff00 + 31 for m in ms), k
    # the deep one really was compressed with codes (not stored), close to its entropy
    z, _ = hostsim_lib.bgzf_deflate(np.frombuffer(payloads[0], dtype=np.uint8).copy())
    assert len(z) < 0.5 * len(payloads[0])


def _header_via_hostsim(idx_dir, bam, sorted_, cl, vn, rg):
    L = C.CDLL(hostsim_lib.build())
    L.hs_index_open.restype = C.c_void_p
    L.hs_index_open.argtypes = [C.c_char_p]
    L.hs_sam_header.restype = C.c_int64
    L.hs_sam_header.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64]
    ix = L.hs_index_open(idx_dir.encode())
    assert ix
    buf = (C.c_uint8 * (1 << 20))()
    n = L.hs_sam_header(ix, 1 if bam else 0, 1 if sorted_ else 0, cl, vn, rg, buf, C.c_int64(1 << 20))
    assert n > 0
    return bytes(buf[:n])


def _pg_fields(text: bytes):
    pg = [l for l in text.split(b"\n") if l.startswith(b"@PG\tID:SNAP\t")][0]
    cl = pg.split(b"\tCL:")[1].rsplit(b"\tVN:", 1)[0]
    vn = pg.rsplit(b"\tVN:", 1)[1]
    rg = [l for l in text.split(b"\n") if l.startswith(b"@RG")][0]
    return cl, vn, rg


def test_file_headers_equal_the_reference_files(cases, small_cfg, reflib, tmp_path):
    """sg_samheader.h: the BAM header block (magic, text, reference table) of the reference binary's sorted .bam, and the header lines of its unsorted .sam,
    byte for byte -- given the command line, version and read-group line those files carry."""
    import gzip, subprocess
    c = cases["single"]
    want = gzip.open(c.bam, "rb").read()[:c.header_bytes]
    l_text = int.from_bytes(want[4:8], "little")
    cl, vn, rg = _pg_fields(want[8:8 + l_text])
    assert _header_via_hostsim(small_cfg.idx, True, True, cl, vn, rg) == want
    fq = str(tmp_path / "r.fq")
    small_cfg.reads["std150"].write_fastq(fq)
    sam = str(tmp_path / "o.sam")
    r = subprocess.run([reflib.SNAP_ALIGNER, "single", small_cfg.idx, fq, "-o", sam, "-t", "1"], capture_output=True, text=True)
    assert r.returncode == 0
    lines = open(sam, "rb").read().split(b"\n")
    hdr = b"".join(l + b"\n" for l in lines if l.startswith(b"@"))
    cl, vn, rg = _pg_fields(hdr)
    assert _header_via_hostsim(small_cfg.idx, False, False, cl, vn, rg) == hdr


@pytest.mark.parametrize("threads", [32, 96, 256, 1024])
def test_deflate_run_by_a_block_of_host_threads(threads):
    """sg_deflate.h with its thread block made of REAL threads (tests/blocksim: threadIdx = a thread-local, __syncthreads = a pthread barrier, atomics =
    relaxed atomic builtins): the per-thread ranges, per-warp histograms, strided loops and barrier placement the one-thread host build cannot exercise.
    Every member inflates to its slice and every thread returns the same member size; partial warps' worth of threads (96) and the kernel's own 1024."""
    import gzip, os
    import blocksim_lib
    rng = np.random.default_rng(5)
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "DESIGN.md"), "rb").read()
    payloads = [text[:0xff00 + 700], rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 70000).tobytes(), bytes(40000),
                rng.integers(0, 256, 30000, dtype=np.uint8).tobytes(), b"A", b"ACGTACGTAA", rng.choice(np.frombuffer(b"FFFFFFF:,#", dtype=np.uint8), 0xff00).tobytes()]
    if threads == 1024:
        payloads = payloads[:3] + payloads[4:6]
    for k, data in enumerate(payloads):
        z, sizes = blocksim_lib.bgzf_deflate(np.frombuffer(data, dtype=np.uint8).copy(), threads)
        ms = _members(z.tobytes())
        assert [len(m) for m in ms] == [int(x) for x in sizes] and len(ms) == (len(data) + 0xff00 - 1) // 0xff00, (threads, k)
        assert b"".join(gzip.decompress(m) for m in ms) == data, (threads, k)
