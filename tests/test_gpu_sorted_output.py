"""Row N4 after the sort, on the device through the C ABI: snapgpu_bam_markdup_device / snapgpu_bam_index_device against the files the reference binary
writes with `-so` (duplicates marked, .bai), and the whole chain align -> BAM records -> sort -> mark -> BGZF + .bai on the device."""
import gzip
import io
import os
import struct

import numpy as np
import pytest

import sorted_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from snap_b200 import engine as e
    assert e.lib().snapgpu_device_count() >= 1, "no CUDA device: the CUDA path has no fallback"
    return e


@pytest.fixture(scope="module")
def gidx(engine, small_cfg):
    ix = engine.Index.open(small_cfg.idx)
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def cases(tmp_path_factory, small_cfg, reflib):
    return sorted_data.make_cases(str(tmp_path_factory.mktemp("sorted_gpu")), small_cfg.contigs, small_cfg.idx, reflib.SNAP_ALIGNER)


def _upload(records):
    import torch
    dev = torch.device("cuda", 0)
    blob = b"".join(records)
    offs = np.zeros(len(records), dtype=np.int64)
    offs[1:] = np.cumsum([len(r) for r in records])[:-1]
    d_rec = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    d_off = torch.from_numpy(offs).to(dev)
    return blob, d_rec, d_off


@pytest.mark.parametrize("name", ["single", "paired", "single_dense", "paired_dense"])
def test_duplicates_marked_on_the_device_like_the_reference(engine, gidx, cases, name):
    """The reference's sorted, unmarked record stream in HBM in; the reference's marked stream out, byte for byte; a second pass marks nothing."""
    c = cases[name]
    fmt = engine.SamFormatter(gidx, engine.default_params(maxDist=14), 4096)
    blob, d_rec, d_off = _upload(c.unmarked)
    marked = fmt.markdup_device(d_rec.data_ptr(), d_off.data_ptr(), len(c.unmarked))
    got = d_rec.cpu().numpy().tobytes()
    want = b"".join(c.marked)
    assert marked == sum(1 for r in c.marked if r[19] & 4) and marked > len(c.marked) // 5
    bad = [i for i, (a, b) in enumerate(zip(sorted_data_split(got), c.marked)) if a != b]
    assert got == want, (len(bad), bad[:5])
    assert fmt.markdup_device(d_rec.data_ptr(), d_off.data_ptr(), len(c.unmarked)) == 0
    fmt.close()


def sorted_data_split(blob):
    out, p = [], 0
    while p < len(blob):
        b = struct.unpack("<i", blob[p:p + 4])[0]; out.append(blob[p:p + 4 + b]); p += 4 + b
    return out


@pytest.mark.parametrize("name", ["single", "paired", "single_dense", "paired_dense"])
def test_bam_index_from_the_device_equals_the_reference_index(engine, gidx, cases, name):
    c = cases[name]
    fmt = engine.SamFormatter(gidx, engine.default_params(maxDist=14), 4096)
    blob, d_rec, d_off = _upload(c.marked)
    bai = fmt.index_device(d_rec.data_ptr(), d_off.data_ptr(), len(c.marked), len(blob), c.header_bytes)
    fmt.close()
    got = sorted_data.parse_bai(bai, sorted_data.our_blocks(c.header_bytes + len(blob)))
    want = sorted_data.parse_bai(c.bai, sorted_data.bgzf_blocks(c.bam))
    assert got == want


def test_reads_to_sorted_marked_indexed_bam_on_the_device(engine, gidx, small_cfg, cases, tmp_path):
    """The whole `snap single ... -so -o out.bam` output stage on the device: reads aligned, BAM records formatted, coordinate-sorted, duplicates marked,
    the file's BGZF members and its .bai made -- the records equal the reference file's (inflated), the index equals the reference's."""
    import torch
    from snap_b200 import synth
    c = cases["single"]
    lines = open(c.bam.replace("se_dup.bam", "se.fq"), "rb").read().split(b"\n")
    ids = [lines[i][1:] for i in range(0, len(lines) - 1, 4)]
    rb = synth.ReadBatch.from_lists([(lines[i + 1], lines[i + 3]) for i in range(0, len(lines) - 1, 4)])
    p = engine.default_params(maxDist=14)
    al = engine.SingleAligner(gidx, p, 4096)
    res, _ = al.align(rb)
    al.close()
    fmt = engine.SamFormatter(gidx, p, 4096)
    fmt.set_format(bam=True)
    fmt.format(rb, ids, res)
    n = fmt.last_record_count()
    assert n == rb.n == len(c.marked)
    dev = torch.device("cuda", 0)
    cap = n * 1024
    d_sorted = torch.empty((cap,), dtype=torch.uint8, device=dev)
    d_offs = torch.empty((n,), dtype=torch.int64, device=dev)
    used = fmt.sort_device(0, d_sorted.data_ptr(), cap, 0, d_offs.data_ptr())
    marked = fmt.markdup_device(d_sorted.data_ptr(), d_offs.data_ptr(), n)
    records = d_sorted[:used].cpu().numpy().tobytes()
    assert marked == sum(1 for r in c.marked if r[19] & 4)
    assert records == b"".join(c.marked)
    # the file: the reference file's own header bytes, then our records, as BGZF members made on the device + the end-of-file member
    header = gzip.open(c.bam, "rb").read()[:c.header_bytes]
    d_all = torch.cat([torch.frombuffer(bytearray(header), dtype=torch.uint8).to(dev), d_sorted[:used]])
    total = len(header) + used
    n_members = (total + 0xff00 - 1) // 0xff00
    d_out = torch.empty((n_members * (0xff00 + 31),), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    out_bytes = engine.bgzf_device(d_all.data_ptr(), total, d_out.data_ptr(), d_out.numel())
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    file_bytes = d_out[:out_bytes].cpu().numpy().tobytes() + eof
    assert gzip.GzipFile(fileobj=io.BytesIO(file_bytes)).read() == header + records
    bai = fmt.index_device(d_sorted.data_ptr(), d_offs.data_ptr(), n, used, len(header))
    blocks = sorted_data.our_blocks(total)
    assert blocks[-1][0] + 28 == len(file_bytes) and [b[0] for b in blocks] == _member_starts(file_bytes)
    assert sorted_data.parse_bai(bai, blocks) == sorted_data.parse_bai(c.bai, sorted_data.bgzf_blocks(c.bam))
    # the same file with the members COMPRESSED on the device (snapgpu_bgzf_deflate_device): smaller, inflates to the same content, and the .bai made
    # for its member offsets again says what the reference's says
    d_out2 = torch.empty((n_members * (0xff00 + 31),), dtype=torch.uint8, device=dev)
    z_bytes, member_offsets = fmt.bgzf_deflate_device(d_all.data_ptr(), total, d_out2.data_ptr(), d_out2.numel())
    zfile = d_out2[:z_bytes].cpu().numpy().tobytes() + eof
    assert z_bytes < out_bytes * 0.8 and int(member_offsets[-1]) == z_bytes
    assert gzip.GzipFile(fileobj=io.BytesIO(zfile)).read() == header + records
    assert _member_starts(zfile) == [int(x) for x in member_offsets]
    zbai = fmt.index_device(d_sorted.data_ptr(), d_offs.data_ptr(), n, used, len(header), member_offsets=member_offsets)
    zpath = str(tmp_path / "z.bam")
    open(zpath, "wb").write(zfile)
    assert sorted_data.parse_bai(zbai, sorted_data.bgzf_blocks(zpath)) == sorted_data.parse_bai(c.bai, sorted_data.bgzf_blocks(c.bam))
    fmt.close()


def test_bgzf_deflate_on_the_device_inflates_to_the_payload(engine, gidx):
    """snapgpu_bgzf_deflate_device on payloads of every kind: each member a valid gzip member (CRC-32, ISIZE, BSIZE) that inflates to its 65280-byte slice;
    incompressible data falls back to stored members; sizes around the member boundary."""
    import torch
    dev = torch.device("cuda", 0)
    fmt = engine.SamFormatter(gidx, engine.default_params(maxDist=14), 1024)
    rng = np.random.default_rng(7)
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "DESIGN.md"), "rb").read()
    payloads = {
        "one": b"A", "three": b"ACG", "tiny": b"ACGTACGTAC", "zeros": bytes(200000), "random": rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(),
        "acgt": rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300000).tobytes(), "text": text * 3, "exact": (text * 3)[:0xff00], "plus1": (text * 3)[:0xff00 + 1],
        "quals": rng.choice(np.frombuffer(b"FFFFFFF:,#", dtype=np.uint8), 500000).tobytes(),
    }
    for name, data in payloads.items():
        d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
        n_members = (len(data) + 0xff00 - 1) // 0xff00
        d_out = torch.zeros((len(data) + 31 * n_members + 64,), dtype=torch.uint8, device=dev)
        used, offs = fmt.bgzf_deflate_device(d_in.data_ptr(), len(data), d_out.data_ptr(), d_out.numel())
        z = d_out[:used].cpu().numpy().tobytes()
        assert _member_starts(z + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))[:-1] == [int(x) for x in offs[:-1]], name
        assert int(offs[-1]) == used, name
        for k in range(n_members):
            assert gzip.decompress(z[int(offs[k]):int(offs[k + 1])]) == data[k * 0xff00:(k + 1) * 0xff00], (name, k)
        if name == "random":
            assert used == len(data) + 31 * n_members           # stored
        if name in ("zeros", "acgt", "text", "quals"):
            assert used < 0.55 * len(data), (name, used)
    fmt.close()


def _member_starts(raw):
    p, out = 0, []
    while p < len(raw):
        out.append(p); p += struct.unpack("<H", raw[p + 16:p + 18])[0] + 1
    return out


def test_file_header_composed_by_the_library_equals_the_reference_header(engine, gidx, cases):
    """snapgpu_sam_header (host code behind the C ABI): the BAM header block of the reference binary's sorted .bam, byte for byte, given the command line,
    version and read-group line that file carries; and the SAM form of the same header."""
    c = cases["single"]
    header = gzip.open(c.bam, "rb").read()[:c.header_bytes]
    text = header[8:8 + int.from_bytes(header[4:8], "little")]
    pg = [l for l in text.split(b"\n") if l.startswith(b"@PG\tID:SNAP\t")][0]
    rg = [l for l in text.split(b"\n") if l.startswith(b"@RG")][0]
    cl, vn = pg.split(b"\tCL:")[1].rsplit(b"\tVN:", 1)[0], pg.rsplit(b"\tVN:", 1)[1]
    fmt = engine.SamFormatter(gidx, engine.default_params(maxDist=14), 1024)
    assert fmt.header(cl, vn, rg, sorted_=True) == text          # SAM form
    fmt.set_format(bam=True)
    assert fmt.header(cl, vn, rg, sorted_=True) == header
    fmt.close()
