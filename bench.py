#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the reference's headline workload (BASELINE.json):
aligned reads/s for `snap single`, 150 bp synthetic reads vs a 3 Gbp hg-sized synthetic reference, seed 20,
maxDist 14 (configs[1]), read-sharded over N GPUs.

A "step" is one pass of the hot path (seed lookup + LV / affine-gap scoring, BaseAligner::AlignRead semantics) over
one batch of synthetic reads.  `value` = reads of all ranks per second with the inputs already resident in HBM,
timed with CUDA events around exactly K steps (max over ranks).  `e2e` = the same metric through the C ABI call a
SNAP extension makes (snapgpu_align_single) with HOST buffers: host->device copies of the reads and device->host
copy of the results inside the timed region.  `roofline` is for the step's alignment launches taken together (two for
sg_align_kernel's two-pass form, four for the staged sg_align_paired_kernel), timed with CUDA events on their stream;
`seed_phase` is the seed-lookup kernel run in isolation over every seed of the batch (BASELINE's second metric).
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference (oracle/_ref, compiled from /root/reference) on the
host cores, on a bounded sample of the same workload, against the same index written out in the reference's own
directory format.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
SEED_LEN = 20
MAX_DIST = 14
ALG_BYTES_PER_CANDIDATE = READ_LEN - SEED_LEN + 2 * (MAX_DIST + 1)     # (readLen - seedLen + 2*k_used) reference bytes
ALG_BYTES_PER_READ_IO = 2 * READ_LEN + 88                              # bases + qualities in, result record out
# stock `snap paired` (BASELINE configs[2] names no options): -d 27, -n 8, -H 4000, -s 0 1000, soft clipping on
PAIRED_MAX_DIST = 27
PAIRED_KW = dict(maxDist=PAIRED_MAX_DIST, numSeedsFromCommandLine=8)
PAIRED_PKW = dict()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mbp", type=int, default=int(os.environ.get("SNAPGPU_BENCH_GENOME_MBP", "3000")),
                    help="total reference size in Mbp (24 contigs); 3000 = BASELINE configs[1]")
    ap.add_argument("--batch-reads", type=int, default=int(os.environ.get("SNAPGPU_BENCH_BATCH", str(1 << 20))),
                    help="reads per step per GPU")
    ap.add_argument("--cpu-sample-reads", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-seed-phase", action="store_true")
    ap.add_argument("--workload", default=os.environ.get("SNAPGPU_BENCH_WORKLOAD", "single"), choices=["single", "paired"],
                    help="single = BASELINE configs[1] (the default, headline); paired = configs[2] shape: stock `snap paired`, "
                         "--batch-reads/2 FR pairs per step, insert N(400,40)")
    ap.add_argument("--phase", default="", choices=["", "sam"], help=argparse.SUPPRESS)      # child-process mode of run_ours' sam_phase leg
    return ap.parse_args()


def measured_traffic(workload, batch_reads, genome_mbp):
    """DRAM bytes of ONE step's alignment launches (summed) from the committed `ncu --set full` capture of this same workload
    (profiles/traffic.json, written by profiles/extract_traffic.py); None when the capture is of another configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%dreads_%dmbp" % (workload, batch_reads, genome_mbp)
        return int(t[key]["dram_bytes_per_launch"]) if key in t else None
    except Exception:
        return None


def measured_instructions(workload, batch_reads, genome_mbp):
    """Warp instructions one step's alignment launches execute (smsp__inst_executed.sum of the same committed capture)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%dreads_%dmbp" % (workload, batch_reads, genome_mbp)
        return int(t[key]["warp_instructions_per_step"]) if key in t and "warp_instructions_per_step" in t[key] else None
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [l.strip().split(",") for l in open(self.path) if l.strip()]
            sm = [float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit()]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nme in enumerate(names):
                if any(len(r) > 4 + k and "Active" in r[4 + k] and "Not" not in r[4 + k] for r in rows):
                    out["reasons"].append(nme)
            out["samples"] = len(rows)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def build_workload(args, device, rank, world):
    """Genome + index in this rank's HBM (replicated) and W+K batches of reads for this rank's shard."""
    import torch
    from snap_b200 import engine, synth_device
    n_contigs = 24
    contig_len = args.genome_mbp * 1_000_000 // n_contigs
    t0 = time.time()
    bases, starts = synth_device.make_genome(n_contigs, contig_len, seed=20260924, device=device)
    torch.cuda.synchronize()
    t1 = time.time()
    idx = engine.Index.build_device(bases.data_ptr(), bases.numel(), starts, seed_len=SEED_LEN, chromosome_padding=2000, device=device.index or 0)
    torch.cuda.synchronize()
    t2 = time.time()
    batches = []
    nb = args.warmup + args.steps
    for b in range(nb):
        # every (rank, step) gets its own reads: the global batch of step b is the concatenation over ranks
        if args.workload == "paired":
            batches.append(synth_device.make_pairs(bases, starts, contig_len, args.batch_reads // 2, READ_LEN, seed=1000 + b * 64 + rank))
        else:
            batches.append(synth_device.make_reads(bases, starts, contig_len, args.batch_reads, READ_LEN, seed=1000 + b * 64 + rank))
    torch.cuda.synchronize()
    t3 = time.time()
    info = idx.info()
    setup = {"genome_s": round(t1 - t0, 2), "index_build_s": round(t2 - t1, 2), "reads_s": round(t3 - t2, 2),
             "index_hbm_gb": round(info.hbmBytes / 1e9, 2), "overflow_words": int(info.overflowTableSize)}
    return bases, starts, contig_len, idx, batches, setup


def run_ours(args):
    import torch
    import torch.distributed as dist
    from snap_b200 import engine, shard
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=device)
    W, K, B = args.warmup, args.steps, args.batch_reads
    bases, starts, contig_len, idx, batches, setup = build_workload(args, device, rank, world)
    paired = args.workload == "paired"
    if paired:
        B = args.batch_reads = (B // 2) * 2
        al = engine.PairedAligner(idx, engine.default_params(**PAIRED_KW), engine.default_paired_params(**PAIRED_PKW), max_batch_pairs=B // 2)
        result_bytes_per_read = engine.PAIRED_RESULT_DTYPE.itemsize // 2
    else:
        al = engine.SingleAligner(idx, engine.default_params(maxDist=MAX_DIST), max_batch_reads=B)
        result_bytes_per_read = engine.RESULT_DTYPE.itemsize
    # a dedicated (non-default) stream: the kernels are launched on it through the C ABI and the CUDA events that time
    # them are recorded on the same stream (a NULL stream argument would mean "the aligner's own stream")
    stream = torch.cuda.Stream(device)
    assert stream.cuda_stream != 0
    res = torch.empty((B, result_bytes_per_read), dtype=torch.uint8, device=device)
    d_ctr = torch.zeros((engine.N_COUNTERS,), dtype=torch.int64, device=device)

    def step(b):
        rb, rq, ro, rl = batches[b][0], batches[b][1], batches[b][2], batches[b][3]
        al.align_device(B // 2 if paired else B, rb.data_ptr(), rq.data_ptr(), ro.data_ptr(), rl.data_ptr(), res.data_ptr(), d_ctr.data_ptr(),
                        stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-ups, then exactly K steps between barriers, CUDA events on the launch stream ----
    for b in range(W):
        step(b)
    barrier()
    d_ctr.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = al.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(K):
        step(W + k)
        ev[k + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms_total = ev[0].elapsed_time(ev[K])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    launches = al.launch_count() - launches0
    if paired:
        al.check(stream.cuda_stream)        # raises if a kernel latched an error (pool overflow)
    ctr = d_ctr.cpu().numpy()
    ms_max = shard.max_over_ranks(ms_total, device) if world > 1 else ms_total
    ctr_all = shard.allreduce_counters(ctr, device) if world > 1 else ctr
    c = engine.counters_dict(ctr_all)
    reads_total = B * K * world
    value = reads_total / (ms_max / 1e3)

    # ---- end to end through the C ABI with host buffers (H2D + D2H inside the timed region) ----
    host_batches = []
    from snap_b200 import synth
    for b in range(min(2, W + K)):
        rb, rq, ro, rl = batches[b][:4]
        # inputs of the end-to-end leg live in pinned host memory (the C ABI then DMAs straight out of them)
        hb = synth.ReadBatch(rb.cpu().pin_memory().numpy(), rq.cpu().pin_memory().numpy(), ro.cpu().numpy().astype(np.uint64), rl.cpu().numpy().astype(np.uint32))
        host_batches.append(hb)
    # ... and so does the result array (torch is only used to get page-locked memory)
    res_dtype = engine.PAIRED_RESULT_DTYPE if paired else engine.RESULT_DTYPE
    n_units = B // 2 if paired else B
    res_host = torch.empty((n_units * res_dtype.itemsize,), dtype=torch.uint8).pin_memory().numpy().view(res_dtype)
    al.align(host_batches[0], out=res_host)         # warm-up
    barrier()
    t0 = time.perf_counter()
    n_e2e = 0
    e2e_steps = max(1, min(K, 3))
    for k in range(e2e_steps):
        r, _ = al.align(host_batches[k % len(host_batches)], out=res_host)
        n_e2e += len(r) * (2 if paired else 1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = shard.max_over_ranks(e2e_s, device) if world > 1 else e2e_s
    e2e_value = n_e2e * world / e2e_s

    # ---- roofline of the step's alignment kernels (sg_align_kernel: pass 1 + pass 2 launches; sg_align_paired_kernel: three
    #      stage launches + the retry launch), timed together with CUDA events: algorithmic bytes of the step / its duration ----
    peak, peak_src = measured_peaks()
    per_launch = 1.0 / (K * world)
    alg_bytes = (c["nHashEntriesProbed"] * 8 + c["nOverflowWordsRead"] * 4 + (c["lvCalls"] + c["affineGapCalls"]) * ALG_BYTES_PER_CANDIDATE
                 + c["totalReads"] * (2 * READ_LEN + result_bytes_per_read)) * per_launch
    kernel_ms_avg = float(np.mean(kernel_ms))
    achieved = alg_bytes / (kernel_ms_avg / 1e3) / 1e9
    roofline = {"kernel": ("sg_align_paired_kernel<.,1..3> (staged launch: seed/LV, affine gap, single-end fallback) + retry pass" if paired else
                           "sg_align_kernel<.,1> + <.,2> (two-pass launch: without affine gap, then the deferred reads)"),
                "launches_per_step": round(launches / K, 2), "bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 6), "traffic": measured_traffic(args.workload, B, args.genome_mbp), "peak_source": peak_src,
                "algorithmic_bytes_per_step": int(alg_bytes), "avg_step_ms": round(kernel_ms_avg, 3),
                "note": "latency/issue-bound integer state machine: ~%.0f B of index+reference+read traffic per read" % (alg_bytes / B)}

    # what actually bounds these kernels is instruction issue, so that utilisation is reported beside the (required) HBM roofline:
    # warp instructions per step (from the committed ncu capture of this workload) / step time, against SMs x 4 schedulers x SM clock
    n_inst = measured_instructions(args.workload, B, args.genome_mbp)
    if n_inst is not None and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        issue_peak = sms * 4 * float(clocks["sm_mhz"]) * 1e6
        issue_ach = n_inst / (kernel_ms_avg / 1e3)
        roofline["issue_slots"] = {"warp_instructions_per_step": n_inst, "achieved_per_s": round(issue_ach, 1), "peak_per_s": round(issue_peak, 1),
                                   "frac": round(issue_ach / issue_peak, 4), "source": "profiles/traffic.json (ncu smsp__inst_executed.sum)"}

    out = {
        "metric": "aligned reads/s", "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_max / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 (+f64 match probabilities)", "data": "synthetic",
        "config": {"workload": ("snap paired (stock options), 2 x %d x %d bp synthetic FR pairs (insert N(400,40)) per step per GPU vs %d Mbp "
                                "synthetic reference (24 contigs), seed %d, maxDist %d, IntersectingPairedEndAligner + chimeric single-end fallback "
                                "(BASELINE configs[2] shape)" % (B // 2, READ_LEN, args.genome_mbp, SEED_LEN, PAIRED_MAX_DIST)) if paired else
                               ("snap single, %d x %d bp synthetic reads per step per GPU vs %d Mbp synthetic reference (24 contigs), "
                                "seed %d, maxDist %d, affine gap on (BASELINE configs[1] shape)" % (B, READ_LEN, args.genome_mbp, SEED_LEN, MAX_DIST)),
                   "reads_per_step": B * world, "read_len": READ_LEN, "genome_mbp": args.genome_mbp, "seed_len": SEED_LEN,
                   "max_dist": PAIRED_MAX_DIST if paired else MAX_DIST,
                   "parallelism": "read-sharded x%d, index replicated per GPU" % world,
                   "l2": "each step uses fresh reads (%.0f MB/step > L2) against a %.1f GB index" % (B * 2 * READ_LEN / 1e6, setup["index_hbm_gb"])},
        "e2e": {"value": round(e2e_value, 1), "unit": "reads/s", "h2d_bytes_per_step": int(B * (2 * READ_LEN + 12)),
                "d2h_bytes_per_step": int(B * result_bytes_per_read + engine.N_COUNTERS * 8), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "per_read": {"lookups": round(c["nHashTableLookups"] / max(1, c["totalReads"]), 3),
                     "hash_entries_per_lookup": round(c["nHashEntriesProbed"] / max(1, c["nHashTableLookups"]), 3),
                     "lv_locations": round(c["lvCalls"] / max(1, c["totalReads"]), 3),
                     "ag_locations": round(c["affineGapCalls"] / max(1, c["totalReads"]), 3),
                     "aligned_frac": round((c["singleHits"] + c["multiHits"]) / max(1, c["totalReads"]), 5)},
        "setup": setup,
    }

    # ---- seed-lookup phase in isolation (rank 0, N=1) ----
    if paired:
        out["per_read"]["aligned_as_pair_frac"] = None
    if rank == 0 and not args.no_seed_phase and not paired:
        try:
            out["seed_phase"] = seed_phase(args, idx, batches, device, peak, peak_src)
        except Exception as e:  # pragma: no cover
            out["seed_phase"] = {"error": str(e)[:200]}

    # ---- FASTQ ingest in isolation (rank 0) ----
    if rank == 0 and not args.no_seed_phase and not paired:
        try:
            out["ingest_phase"] = ingest_phase(args, host_batches[0], device, peak, peak_src)
        except Exception as e:  # pragma: no cover
            out["ingest_phase"] = {"error": str(e)[:200]}

    # ---- CPU baseline: the unmodified reference on the host cores, bounded sample (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            got0, _ = al.align(host_batches[0])          # the engine's records for the reads the reference is about to be timed on
            out["cpu_baseline"] = cpu_baseline(args, idx, host_batches[0], check_against=got0)
            if paired:
                out["per_read"]["aligned_as_pair_frac"] = out["cpu_baseline"].pop("aligned_as_pair_frac", None)
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"error": str(e)[:300]}

    # ---- output stage in isolation (rank 0, N=1; SURVEY 8f N1, first device form).  The formatter was verified on the GPU in the
    #      last minutes of its round and has not been through a full-size run yet, so it gets a process of its own: whatever happens
    #      there, this process still prints its line.  (Our aligner's arena is released first to leave the child room in HBM.) ----
    if rank == 0 and world == 1 and not args.no_seed_phase:
        try:
            al.close()
            cmd = [sys.executable, os.path.abspath(__file__), "--phase", "sam", "--workload", args.workload, "--genome-mbp", str(args.genome_mbp),
                   "--batch-reads", str(args.batch_reads), "--steps", "1", "--warmup", "0"]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            lines = [l for l in r.stdout.strip().split("\n") if l.startswith("{")]
            out["sam_phase"] = json.loads(lines[-1]) if lines else {"error": "child exited %d: %s" % (r.returncode, r.stderr[-300:])}
        except Exception as e:  # pragma: no cover
            out["sam_phase"] = {"error": str(e)[:300]}

    if rank == 0:
        emit(json.dumps(out))
    try:
        al.close()
        idx.close()
    except Exception:  # pragma: no cover
        pass
    if world > 1:
        dist.destroy_process_group()


def seed_phase(args, idx, batches, device, peak, peak_src):
    """sg_lookup_kernel over the 7 non-overlapping seeds of every read of one batch: algorithmic bytes (entries examined
    x 8 B + overflow words) per second, against the streaming peak and a measured random-8-byte-gather rate."""
    import torch
    B = args.batch_reads
    rb = batches[0][0].reshape(B, READ_LEN)
    offs = [0, 20, 40, 60, 80, 100, 120]
    seeds = torch.stack([rb[:, o:o + SEED_LEN] for o in offs], dim=1).contiguous()       # [B, 7, 20]
    n = B * len(offs)
    nh = torch.empty((n * 2,), dtype=torch.int64, device=device)
    probes = torch.empty((n,), dtype=torch.int32, device=device)
    st = torch.cuda.Stream(device)
    torch.cuda.synchronize()
    for _ in range(2):
        idx.lookup_seeds_device(seeds.data_ptr(), n, nh.data_ptr(), 0, probes.data_ptr(), 0, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(st)
    for _ in range(reps):
        idx.lookup_seeds_device(seeds.data_ptr(), n, nh.data_ptr(), 0, probes.data_ptr(), 0, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    entries = int(probes.to(torch.int64).sum().item())
    multi = nh[nh > 1]
    overflow_words = int((multi + 1).sum().item())
    alg = entries * 8 + overflow_words * 4 + n * SEED_LEN
    # random 8-byte gather rate over a table of the same size (the regime GetFirstValueForKey lives in)
    slots = int(idx.info().hashTableSlots)
    tbl = torch.empty((min(slots, 1 << 31),), dtype=torch.int64, device=device)
    gi = torch.randint(0, tbl.numel(), (1 << 24,), device=device)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        _ = tbl[gi]
        g0.record(st)
        for _ in range(5):
            _ = tbl[gi]
        g1.record(st)
    torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1) / 5
    gather_gbs = gi.numel() * 8 / (gather_ms / 1e3) / 1e9
    gather_sector_gbs = gi.numel() * 32 / (gather_ms / 1e3) / 1e9
    del tbl, gi
    achieved = alg / (ms / 1e3) / 1e9
    sector = entries * 32 / (ms / 1e3) / 1e9
    return {"kernel": "sg_lookup_kernel", "seeds": n, "ms": round(ms, 3), "lookups_per_s": round(n / (ms / 1e3), 1),
            "entries_per_lookup": round(entries / n, 3), "achieved_algorithmic_gbs": round(achieved, 2),
            "achieved_if_every_entry_costs_a_32B_sector_gbs": round(sector, 2), "peak_stream_gbs": peak, "peak_source": peak_src,
            "frac_of_stream_peak": round(achieved / peak, 5),
            "torch_random_8B_gather_gbs": round(gather_gbs, 2), "torch_random_gather_sector_gbs": round(gather_sector_gbs, 2),
            "frac_of_random_gather_rate": round(achieved / gather_gbs, 4)}


def ingest_phase(args, host_batch, device, peak, peak_src):
    """FASTQ ingest: snapgpu_fastq_parse_device over the FASTQ text of one batch, text resident in HBM.
    Algorithmic bytes = text read once + clipped bases and qualities + offsets/lengths/id table written once."""
    import torch
    from snap_b200 import engine
    n = host_batch.n
    L = READ_LEN
    rec = 2 + 8 + 1 + L + 3 + L + 1
    txt = np.empty((n, rec), dtype=np.uint8)
    txt[:, 0] = ord("@"); txt[:, 1] = ord("r")
    ids = np.arange(n, dtype=np.int64)
    for d in range(8):
        txt[:, 2 + 7 - d] = (ids // 10 ** d % 10 + 48).astype(np.uint8)
    txt[:, 10] = 10
    txt[:, 11:11 + L] = host_batch.bases.reshape(n, L)
    txt[:, 11 + L] = 10; txt[:, 12 + L] = ord("+"); txt[:, 13 + L] = 10
    q = host_batch.quals.reshape(n, L).copy()
    q[::7, L - 9:] = ord("#")                       # every 7th read carries a '#' tail the reader clips
    txt[:, 14 + L:14 + 2 * L] = q
    txt[:, 14 + 2 * L] = 10
    text = txt.reshape(-1)
    d_text = torch.from_numpy(text).to(device)
    fq = engine.FastqParser(max_bytes=int(text.size) + 64, max_reads=n + 8, device=device.index or 0)
    d_b = torch.empty((text.size // 2 + 64,), dtype=torch.uint8, device=device); d_q = torch.empty_like(d_b)
    d_off = torch.empty((n + 8,), dtype=torch.int64, device=device); d_len = torch.empty((n + 8,), dtype=torch.int32, device=device)
    d_ido = torch.empty((n + 8,), dtype=torch.int64, device=device); d_idl = torch.empty((n + 8,), dtype=torch.int32, device=device)
    d_fc = torch.empty((n + 8,), dtype=torch.int32, device=device)
    st = torch.cuda.Stream(device)

    def run():
        return fq.parse_device(d_text.data_ptr(), int(text.size), 2, d_b.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_ido.data_ptr(),
                               d_idl.data_ptr(), d_fc.data_ptr(), st.cuda_stream)
    for _ in range(3):
        nr, used = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kept = int(d_len[:nr].to(torch.int64).sum().item())
    alg = int(text.size) + 2 * kept + nr * (8 + 4 + 8 + 4 + 4)
    out = {"kernels": "sg_fastq_count/positions/records/copy + 2 cub scans", "reads": int(nr), "text_bytes": int(text.size), "ms": round(ms, 3),
           "text_gbs": round(text.size / (ms / 1e3) / 1e9, 2), "reads_per_s": round(nr / (ms / 1e3), 1),
           "achieved_algorithmic_gbs": round(alg / (ms / 1e3) / 1e9, 2), "peak_gbs": peak, "peak_source": peak_src,
           "frac": round(alg / (ms / 1e3) / 1e9 / peak, 4), "bytes_consumed_ok": bool(used == text.size)}
    try:
        from oracle import reflib
        if reflib.available():
            m = min(n, 200000)
            t0 = time.perf_counter()
            reflib.fastq_parse(text[:m * rec], 2)
            dt = time.perf_counter() - t0
            out["cpu_reference_1thread_text_gbs"] = round(m * rec / dt / 1e9, 3)
    except Exception as e:  # pragma: no cover
        out["cpu_reference_error"] = str(e)[:100]
    fq.close()
    return out


def sam_phase(args, idx, host_batch, results, paired):
    """Output stage, first device form: snapgpu_sam_format_single / _paired over a bounded sample of the step's reads
    and the engine's own result records, through the host-buffer C ABI (copies and the host-side packing of the records inside the
    timed region).  Beside it the reference's routine for the same job, SAMFormat::computeCigarString, on ONE host thread over a
    smaller sample (oracle/_ref; the reference's writer is single-threaded per output buffer)."""
    from snap_b200 import engine, synth
    n = min(131072, host_batch.n)
    n -= n % 2
    sample = host_batch.slice(0, n)
    res = np.ascontiguousarray(results[:n // 2] if paired else results[:n])
    ids = [(b"p%d/%d" % (i // 2, 1 + i % 2)) if paired else (b"r%d" % i) for i in range(n)]
    p = engine.default_params(**PAIRED_KW) if paired else engine.default_params(maxDist=MAX_DIST)
    fmt = engine.SamFormatter(idx, p, n, use_m=True)
    try:
        id_buf, id_offs, id_lens = fmt.pack_ids(ids)
        sb = synth.ReadBatch(np.ascontiguousarray(sample.bases), np.ascontiguousarray(sample.quals), np.ascontiguousarray(sample.offsets), np.ascontiguousarray(sample.lens))
        buf, used = fmt.format_arrays(sb, id_buf, id_offs, id_lens, res, paired)          # warm-up (and the text buffer)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            buf, used = fmt.format_arrays(sb, id_buf, id_offs, id_lens, res, paired, text=buf)
        dt = (time.perf_counter() - t0) / reps
        text = buf[:used].tobytes()
    finally:
        fmt.close()
    out = {"kernel": "sg_sam_kernel (one thread per read; first form, not optimised)", "reads": int(n), "ms": round(dt * 1e3, 3),
           "reads_per_s": round(n / dt, 1), "text_bytes": len(text), "text_gbs": round(len(text) / dt / 1e9, 3),
           "records_ok": text.count(b"\n") == n}
    if args.genome_mbp <= 300:
        # (needs the index written out in the reference's format: only worth a second export for small genomes; the routine's cost
        #  per read does not depend on the genome -- 0.11 M reads/s on one thread measured on the 0.36 Mbp test genome)
        try:
            out["cpu_reference_1thread_reads_per_s"] = sam_cpu_reference(idx, sample, res, paired)
        except Exception as e:  # pragma: no cover
            out["cpu_reference_error"] = str(e)[:200]
    else:
        out["cpu_reference_1thread_reads_per_s"] = None
        out["cpu_reference_note"] = "SAMFormat::computeCigarString on one host thread: 0.107-0.112 M reads/s (measured on the test genome; not re-timed at this genome size)"
    return out


def run_sam_phase_child(args):
    """`bench.py --phase sam`: builds the same workload, aligns one batch and prints sam_phase's JSON object (parent: run_ours)."""
    import torch
    from snap_b200 import engine, synth
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    bases, starts, contig_len, idx, batches, setup = build_workload(args, device, 0, 1)
    paired = args.workload == "paired"
    B = args.batch_reads
    if paired:
        B = (B // 2) * 2
        al = engine.PairedAligner(idx, engine.default_params(**PAIRED_KW), engine.default_paired_params(**PAIRED_PKW), max_batch_pairs=B // 2)
    else:
        al = engine.SingleAligner(idx, engine.default_params(maxDist=MAX_DIST), max_batch_reads=B)
    rb, rq, ro, rl = batches[0][:4]
    hb = synth.ReadBatch(rb.cpu().numpy(), rq.cpu().numpy(), ro.cpu().numpy().astype(np.uint64), rl.cpu().numpy().astype(np.uint32))
    got, _ = al.align(hb)
    al.close()
    emit(json.dumps(sam_phase(args, idx, hb, got, paired)))


def sam_cpu_reference(idx, sample, res, paired, n_cpu=20000):
    """reads/s of the reference's SAMFormat::computeCigarString (both overloads; the dominant cost of its writer) on one host thread."""
    from oracle import reflib
    from snap_b200 import synth
    if not reflib.available():
        return None
    d = export_index_for_reference(idx)
    try:
        ridx = reflib.RefIndex(d)
        m = min(n_cpu, sample.n)
        data = []; qual = []; jl = []; ja = []; off = 0
        for i in range(m):
            if paired:
                r = res[i // 2]; w = i % 2
                status, loc, direction, used_ag, score, cb, ca = (int(r["status"][w]), int(r["location"][w]), int(r["direction"][w]), int(r["usedAffineGapScoring"][w]),
                                                                   int(r["score"][w]), int(r["basesClippedBefore"][w]), int(r["basesClippedAfter"][w]))
            else:
                r = res[i]
                status, loc, direction, used_ag, score, cb, ca = (int(r["status"]), int(r["location"]), int(r["direction"]), int(r["usedAffineGapScoring"]), int(r["score"]),
                                                                   int(r["basesClippedBefore"]), int(r["basesClippedAfter"]))
            if status == 0:
                continue
            b, q = sample.read(i)
            x = np.frombuffer(b, dtype=np.uint8); y = np.frombuffer(q, dtype=np.uint8)
            if direction == 1:
                x = synth.revcomp(x); y = y[::-1]
            data.append(x); data.append(np.zeros(16, dtype=np.uint8)); qual.append(y); qual.append(np.zeros(16, dtype=np.uint8))
            if used_ag or score > 0:
                ja.append((off + cb, loc, len(b) - cb - ca, cb, 0, ca, 0, 0, direction, 1, score, 0))
            else:
                jl.append((off, loc, len(b), 0, 0, 0, 0, 0, direction, 1))
            off += len(b) + 16
        data = np.concatenate(data); qual = np.concatenate(qual)
        jl = np.array(jl, dtype=reflib.CIGAR_JOB_DTYPE); ja = np.array(ja, dtype=reflib.CIGAR_AG_JOB_DTYPE)
        t0 = time.perf_counter()
        if jl.size:
            reflib.cigar_lv_batch(ridx, data, jl)
        if ja.size:
            reflib.cigar_ag_batch(ridx, data, qual, ja)
        dt = time.perf_counter() - t0
        return round((jl.size + ja.size) / dt, 1)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def export_index_for_reference(idx):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    d = tempfile.mkdtemp(prefix="snapidx_", dir=base)
    idx.save(d)
    return d


def count_differing(want, got, paired):
    """Result records of the engine vs the reference's for the same reads, bytewise (doubles bit for bit), vectorised.  Same
    exclusions as tests/conftest.py: single-end records both sides report NotFound (the reference leaves the rest
    uninitialised on its early returns); mapq / scorePriorToClipping of a paired end reported NotFound (copied from an
    unwritten stack object by ChimericPairedEndAligner)."""
    w, g = want.copy(), np.asarray(got).view(want.dtype).copy()
    if paired:
        for arr in (w, g):
            nf = arr["status"] == 0
            arr["mapq"][nf] = 0
            arr["scorePriorToClipping"][nf] = 0
        skip = np.zeros(len(w), dtype=bool)
    else:
        skip = (w["status"] == 0) & (g["status"] == 0)
    wb = w.view(np.uint8).reshape(len(w), -1)
    gb = g.view(np.uint8).reshape(len(g), -1)
    return int(((wb != gb).any(axis=1) & ~skip).sum())


def cpu_baseline(args, idx, host_batch, check_against):
    """oracle/_ref (the compiled, unmodified reference) on all host cores over a bounded sample of the same reads.  The
    reference's records for the sample are also compared with the engine's (`check_against`): full-size parity evidence."""
    from oracle import reflib
    if not reflib.available():
        return {"error": "oracle/_ref not built"}
    d = export_index_for_reference(idx)
    try:
        t0 = time.time()
        ridx = reflib.RefIndex(d)
        load_s = time.time() - t0
        cores = os.cpu_count() or 1
        n = min(args.cpu_sample_reads, host_batch.n)
        sample = host_batch.slice(0, n)
        if args.workload == "paired":
            n -= n % 2
            sample = host_batch.slice(0, n)
            p, pp = reflib.default_params(**PAIRED_KW), reflib.default_paired_params(**PAIRED_PKW)
            reflib.paired_align_mt(ridx, p, pp, sample.slice(0, min(n, 20000)), cores)
            res, ctr, secs = reflib.paired_align_mt(ridx, p, pp, sample, cores)
            parity = None if check_against is None else {"pairs_compared": n // 2, "differing": count_differing(res, check_against[:n // 2], True)}
            return {"parity_vs_reference": parity, "value": round(n / secs, 1), "unit": "reads/s", "cores": cores, "kind": "reference",
                    "sample": "%d of the step's %d pairs, %d threads, oracle/_ref ChimericPairedEndAligner(IntersectingPairedEndAligner)::align "
                              "(aligner only, no SAM output); index = ours exported to SNAP's directory format (load %.1fs)"
                              % (n // 2, host_batch.n // 2, cores, load_s),
                    "seconds": round(secs, 3), "aligned_as_pair_frac": round(float(res["alignedAsPair"].mean()), 5)}
        p = reflib.default_params(maxDist=MAX_DIST)
        reflib.align_mt(ridx, p, sample.slice(0, min(n, 20000)), cores)          # warm the page cache / TLB
        res, ctr, secs = reflib.align_mt(ridx, p, sample, cores)
        parity = None if check_against is None else {"reads_compared": n, "differing": count_differing(res, check_against[:n], False)}
        return {"parity_vs_reference": parity, "value": round(n / secs, 1), "unit": "reads/s", "cores": cores, "kind": "reference",
                "sample": "%d of the step's %d reads, %d threads, oracle/_ref BaseAligner::AlignRead (aligner only, no SAM output); "
                          "index = ours exported to SNAP's directory format (load %.1fs)" % (n, host_batch.n, cores, load_s),
                "seconds": round(secs, 3), "lv_per_read": round(ctr["lvCalls"] / n, 3), "ag_per_read": round(ctr["affineGapCalls"] / n, 3)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from snap_b200 import engine, synth
    from oracle import reflib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl reference: the workload's index is generated on the GPU; no CUDA device")
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    W, K = args.warmup, args.steps
    n = args.cpu_sample_reads
    saved = args.batch_reads
    args.batch_reads = n
    bases, starts, contig_len, idx, batches, setup = build_workload(args, device, 0, 1)
    args.batch_reads = saved
    d = export_index_for_reference(idx)
    idx.close()
    del bases
    torch.cuda.empty_cache()
    try:
        ridx = reflib.RefIndex(d)
        cores = os.cpu_count() or 1
        p = reflib.default_params(maxDist=MAX_DIST)
        hb = []
        for b in range(W + K):
            rb, rq, ro, rl = batches[b][:4]
            hb.append(synth.ReadBatch(rb.cpu().numpy(), rq.cpu().numpy(), ro.cpu().numpy().astype(np.uint64), rl.cpu().numpy().astype(np.uint32)))
        paired = args.workload == "paired"
        if paired:
            p, pp = reflib.default_params(**PAIRED_KW), reflib.default_paired_params(**PAIRED_PKW)
            run = lambda batch: reflib.paired_align_mt(ridx, p, pp, batch, cores)
        else:
            run = lambda batch: reflib.align_mt(ridx, p, batch, cores)
        for b in range(W):
            run(hb[b])
        total_s = 0.0
        aligned = 0
        for k in range(K):
            res, ctr, secs = run(hb[W + k])
            total_s += secs
            aligned += int((res["status"] != 0).sum())
        value = n * K / total_s
        sample = "%d reads per step (bounded sample of the %d-read step), %d threads" % (n, saved, cores)
        out = {"impl": "reference", "metric": "aligned reads/s", "value": round(value, 1), "unit": "reads/s", "n_gpus": args.gpus, "steps": K,
               "warmup": W, "ms_per_step": round(total_s / K * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8/int32 (+f64 match probabilities)", "data": "synthetic",
               "config": {"workload": ("snap paired (stock options), 2 x %d x %d bp synthetic FR pairs per step vs %d Mbp synthetic reference (24 contigs), seed %d, maxDist %d"
                                       % (n // 2, READ_LEN, args.genome_mbp, SEED_LEN, PAIRED_MAX_DIST)) if paired else
                                      ("snap single, %d x %d bp synthetic reads per step vs %d Mbp synthetic reference (24 contigs), seed %d, maxDist %d"
                                       % (n, READ_LEN, args.genome_mbp, SEED_LEN, MAX_DIST)), "read_len": READ_LEN, "genome_mbp": args.genome_mbp,
                          "seed_len": SEED_LEN, "max_dist": MAX_DIST},
               "cpu_baseline": {"value": round(value, 1), "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
               "e2e": {"value": round(value, 1), "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "aligned_frac": round(aligned / (n * K), 5),
               "note": "unmodified amplab/snap aligner (BaseAligner::AlignRead / ChimericPairedEndAligner::align) via oracle/_ref on host cores; index = GPU-built, exported to SNAP's format"}
        emit(json.dumps(out))
    finally:
        shutil.rmtree(d, ignore_errors=True)


_REAL_STDOUT = None


def emit(line: str) -> None:
    """The one JSON line goes to the process's real stdout; everything else any library prints (NCCL's version banner, for one,
    goes to fd 1 at communicator creation whatever NCCL_DEBUG says short of unset) has been routed to stderr by main()."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                      # C-level stdout of this process (and its children) -> stderr
    sys.stdout = sys.stderr
    if args.phase == "sam":
        run_sam_phase_child(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
