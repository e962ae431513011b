"""N>1 path on CPU: world_size-2 gloo.  Read ranges partition the batch exactly; per-rank counters reduce to the
single-process totals; the per-shard results, concatenated, equal the unsharded run (host-simulated engine)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, idx_dir, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch.distributed as dist
    from oracle import reflib
    from snap_b200 import shard, synth
    import hostsim_lib as hs
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = np.load(os.path.join(out_dir, "reads.npz"))
    rb = synth.ReadBatch(g["bases"], g["quals"], g["offsets"], g["lens"])
    lo, hi = shard.shard_range(rb.n, rank, world)
    al = hs.HsAligner(hs.HsIndex(idx_dir), reflib.default_params(maxDist=14))
    res, ctr = al.align(rb.slice(lo, hi), reflib.RESULT_DTYPE, reflib.N_COUNTERS)
    total = shard.allreduce_counters(ctr)
    tmax = shard.max_over_ranks(float(rank + 1))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), res=res, total=total, lo=lo, hi=hi, tmax=tmax)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from snap_b200 import shard
    for n in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_rank_gloo_matches_single_process(small_cfg, reflib, tmp_path):
    import torch.multiprocessing as mp
    import hostsim_lib as hs
    rb = small_cfg.reads["noisy150"]
    np.savez(str(tmp_path / "reads.npz"), bases=rb.bases, quals=rb.quals, offsets=rb.offsets, lens=rb.lens)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, small_cfg.idx, str(tmp_path)), nprocs=2, join=True)
    whole, wctr = hs.HsAligner(hs.HsIndex(small_cfg.idx), reflib.default_params(maxDist=14)).align(rb, reflib.RESULT_DTYPE, reflib.N_COUNTERS)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(2)]
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == rb.n
    cat = np.concatenate([parts[0]["res"], parts[1]["res"]])
    assert cat.tobytes() == whole.tobytes()
    assert np.array_equal(parts[0]["total"], wctr) and np.array_equal(parts[1]["total"], wctr)
    assert parts[0]["tmax"] == parts[1]["tmax"] == 2.0
