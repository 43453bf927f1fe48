#!/usr/bin/env python
"""bench.py -- Gbases/s sketched + candidate overlaps/s of the MinHash overlap hot path.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference path on the host cores
Under torchrun (N>1) one rank per GPU.  Reads are sharded; every rank sketches, stores and indexes its own shard
(mhapb_store_add_reads*), then ONE collective library call (mhapb_dist_search_self: NCCL all-gather of the forward
sketch blocks inside libmhap_b200.so, hidden behind K2a/K2b) lets every rank query its index with the forward
sketches of all ranks.  torch.distributed only carries the 128-byte communicator id, the barriers and the max-over-ranks
of the timings.

Workloads (--config): 1 = BASELINE configs[1] (default; weak scaling: 100k reads x 10 kbp PER GPU), 2 = configs[2]
(50k x 8 kbp, --ordered-sketch-size 1000), 3 = configs[3] (-s store -q query, 500k x 500k x 12 kbp in total, meant for
4 GPUs), 4 = configs[4] (1M x 15 kbp in total, --num-hashes 1024, meant for 8 GPUs).

A "step" is one full pass over the synthetic read set: K1 sketch (both strands) -> K2a index -> K2b probe/count ->
K2c ordered filter -> hits on the host (config 3: store pass incl. its self search, then the query file).
  value  : bases of all ranks / step time, reads already resident in HBM when the step starts.
  e2e    : the same through the host-buffer C-ABI calls, host->device copy of the reads and device->host copy of the
           hits inside the timed region.
The reference (marbl/MHAP, Java) cannot run here (no JVM in the image; Guava/fastutil un-vendored), so
--impl reference and cpu_baseline time oracle/ -- a C restatement of the Java path that avoids the
JVM's per-k-mer allocation, i.e. a faster-than-reference baseline ("kind": "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(name="BASELINE configs[1]", reads=100_000, read_len=10_000, num_hashes=512, ordered_sketch_size=1536, mode="self", per_gpu=True),
    2: dict(name="BASELINE configs[2]", reads=50_000, read_len=8_000, num_hashes=512, ordered_sketch_size=1000, mode="self", per_gpu=True),
    3: dict(name="BASELINE configs[3]", reads=500_000, query_reads=500_000, read_len=12_000, num_hashes=512, ordered_sketch_size=1536, mode="query", per_gpu=False),
    4: dict(name="BASELINE configs[4]", reads=1_000_000, read_len=15_000, num_hashes=1024, ordered_sketch_size=1536, mode="self", per_gpu=False),
}


def alg_bytes_per_read(L, H, S, ok=12, strands=2):
    """per read ceil(L/4)+8 bytes in, strands*(4H + 8*min(S, L-ok+1)) bytes out."""
    return (L + 3) // 4 + 8 + strands * (4 * H + 8 * min(S, L - ok + 1))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=None, help="override: reads per GPU")
    ap.add_argument("--query-reads", type=int, default=None, help="override (config 3): query reads per GPU")
    ap.add_argument("--read-len", type=int, default=None)
    ap.add_argument("--num-hashes", type=int, default=None)
    ap.add_argument("--ordered-sketch-size", type=int, default=None)
    ap.add_argument("--err", type=float, default=0.15)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target length of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (single-GPU re-run for N<=2, oracle sample)")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    world = max(1, a.gpus)
    per = (lambda tot: tot if c["per_gpu"] else -(-tot // world))
    a.mode = c["mode"]
    a.reads = a.reads if a.reads is not None else per(c["reads"])
    a.query_reads = a.query_reads if a.query_reads is not None else (per(c["query_reads"]) if c["mode"] == "query" else 0)
    a.read_len = a.read_len or c["read_len"]
    a.num_hashes = a.num_hashes or c["num_hashes"]
    a.ordered_sketch_size = a.ordered_sketch_size or c["ordered_sketch_size"]
    a.config_name = c["name"]
    a.weak = c["per_gpu"]
    return a


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Number of samples taken so far (the timed region keeps the rows between two marks)."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        self.rows = self.rows[first:last]
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        hot = sorted(sm)[len(sm) // 2:]  # median of the upper half ~ under load
        return {"sm_mhz": float(np.median(hot)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def oracle_sample_run(bases, offsets, n_reads, L, H, S, threads, want_hits=False):
    """One pass of the CPU path (oracle port, pthread pool like Executors.newFixedThreadPool) over a sample."""
    from oracle import oracle as orc
    st = orc.Store(k=16, num_hashes=H, ordered_k=12, ordered_size=S)
    t0 = time.perf_counter()
    st.add_reads(bases[: n_reads * L], offsets[: n_reads + 1], threads=threads)
    t1 = time.perf_counter()
    res = st.search_self(threads=threads)
    t2 = time.perf_counter()
    st.close()
    return t1 - t0, t2 - t1, res.stats, (res.hits if want_hits else None)


def cpu_sample_size(args, cores):
    # ~1.5 ns per XORShift-min step per core for the C port; both strands
    per_read_s = 2.0 * (args.read_len - 15) * args.num_hashes * 1.5e-9
    n = int(args.cpu_seconds * cores / per_read_s)
    return max(cores, min(args.reads, n))


def make_genome(args, world):
    from mhap_b200 import synth
    total_reads = args.reads * world
    glen = max(args.read_len + 1, int(total_reads * args.read_len / 20))
    return synth.genome(args.seed, glen)


READ_SEED = lambda seed: (seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
QUERY_SEED = lambda seed: ((seed + 1) * 0xC2B2AE3D27D4EB4F) & 0xFFFFFFFFFFFFFFFF


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mhap_b200 import synth
    cores = os.cpu_count() or 1
    n = cpu_sample_size(args, cores)
    total_reads = args.reads * args.gpus
    g = make_genome(args, args.gpus)
    bases, offsets = synth.reads(g, READ_SEED(args.seed), 0, n, args.read_len, args.err)
    times = []
    stats = None
    for it in range(args.warmup + args.steps):
        ts, tq, stats, _ = oracle_sample_run(bases, offsets, n, args.read_len, args.num_hashes, args.ordered_sketch_size, cores)
        if it >= args.warmup:
            times.append((ts, tq))
    ts = float(np.mean([t[0] for t in times])); tq = float(np.mean([t[1] for t in times]))
    gb = n * args.read_len / 1e9
    val = gb / (ts + tq)
    sample = f"first {n} of {total_reads} reads x {args.read_len} bp (same generator/seed), sketch+index+self-search, {cores} threads"
    out = {
        "impl": "reference", "metric": "gbases_per_s_sketched_and_overlapped", "value": val, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (ts + tq) * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.weak else "fixed job", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "sketch_gbases_per_s": gb / ts, "overlaps_per_s": stats["fully_compared"] / tq if tq > 0 else None,
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the Java path (oracle/); the JVM reference cannot run in this image"},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=real_stdout, flush=True)


def workload_config(args, world):
    what = (f"{args.reads} synthetic PacBio-shape reads x {args.read_len} bp per GPU" if args.mode == "self" else
            f"-s store ({args.reads} reads per GPU) vs -q query ({args.query_reads} reads per GPU, forward only) x {args.read_len} bp, store self pass included (MhapMain.java:516-521)")
    return {"workload": f"{args.config_name}: {what}, k=16, --num-hashes {args.num_hashes}, "
                        f"{'self-overlap' if args.mode == 'self' else 'store-vs-query'} (both strands of the store sketched)",
            "mode": args.mode, "reads_per_gpu": args.reads, "query_reads_per_gpu": args.query_reads, "read_len": args.read_len, "k": 16,
            "num_hashes": args.num_hashes, "ordered_kmer_size": 12, "ordered_sketch_size": args.ordered_sketch_size, "error_rate": args.err,
            "coverage": 20, "seed": args.seed,
            "l2_policy": f"inputs larger than L2 ({args.reads * args.read_len / 1e9:.2f} GB of reads per GPU per step, sketches {2 * args.reads * (4 * args.num_hashes + 8 * args.ordered_sketch_size) / 1e9:.2f} GB)",
            "parallelism": f"reads sharded over {world} GPU(s) (two folded half-shards per rank: equal K2c load); each rank indexes its shard; NCCL all-gather of the forward sketch blocks inside the library; every rank queries its index with all forward sketches"}


def main():
    args = parse()
    # NCCL (and anything else) may print to fd 1; the contract is ONE JSON line on stdout, so route
    # fd 1 to stderr for the duration of the run and keep the real stdout for the result line.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _main(args, real_stdout)
    finally:
        real_stdout.flush()


def _main(args, real_stdout):
    if args.impl == "reference":
        run_reference(args, real_stdout)
        return

    import torch
    import torch.distributed as dist
    from mhap_b200 import native, synth
    from mhap_b200.distributed import bootstrap_comm, folded_shard_ranges, gather_hits, hits_digest

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    eng = native.Engine(local)
    bootstrap_comm(eng, dist if world > 1 else None)        # NCCL communicator inside the library
    nccl_version = eng.comm_info()[2] if world > 1 else None
    L, H, S = args.read_len, args.num_hashes, args.ordered_sketch_size
    p = native.SketchParams(16, H, 12, S, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1)
    n_local, nq_local = args.reads, args.query_reads
    total_reads, total_queries = n_local * world, nq_local * world
    g = make_genome(args, world)
    # this rank's reads: two half-shards of the job's read stream, folded so that every rank stores the same number of
    # (lower id, higher id) pairs to score -- a contiguous shard leaves rank 0 with ~2x the K2c work of one GPU and the last
    # rank with none (distributed.folded_shard_ranges); ids are the global 1-based positions, the job's hit set is unchanged
    parts = folded_shard_ranges(total_reads, rank, world)
    assert sum(c for _, c in parts) == n_local, "reads per GPU must be even"
    host = torch.empty(n_local * L, dtype=torch.uint8).pin_memory()
    bases = host.numpy()
    at = 0
    for first, cnt in parts:
        synth.reads(g, READ_SEED(args.seed), first, cnt, L, args.err, out=bases[at * L:(at + cnt) * L])
        at += cnt
    offsets = np.arange(n_local + 1, dtype=np.uint64) * np.uint64(L)
    ids = np.concatenate([np.arange(first + 1, first + cnt + 1, dtype=np.int64) for first, cnt in parts])
    qbases = qoffsets = qids = d_qbases = None
    if args.mode == "query":
        qhost = torch.empty(nq_local * L, dtype=torch.uint8).pin_memory()
        qbases = qhost.numpy()
        _, qoffsets = synth.reads(g, QUERY_SEED(args.seed), rank * nq_local, nq_local, L, args.err, out=qbases)
        qids = np.arange(1, nq_local + 1, dtype=np.int64) + rank * nq_local + total_reads      # MhapMain.java:537: ids continue after the store's
        d_qbases = qhost.to("cuda", non_blocking=True)
    d_bases = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()
    del g

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    TKEYS = ("sketch_total_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms", "index_ms", "probe_ms", "filter_ms", "gather_ms", "h2d_ms", "d2h_ms")
    acc = {k: 0.0 for k in TKEYS}
    launches = {"n": 0}
    last = {}

    def search_self():
        return eng.dist_search_self(sp) if world > 1 else eng.search_self(sp)

    def step(resident=True, record=False):
        """One pass of the hot path over this rank's shard."""
        t_a = time.perf_counter()
        eng.store_reset(p)
        if resident:
            eng.store_add_reads_device(d_bases.data_ptr(), offsets, ids)       # K1 straight into the rank's store
        else:
            eng.store_add_reads(bases, offsets, ids)                           # H2D of the reads inside
        tm = eng.timing()
        xsteps = tm["xorshift_steps"]
        t_b = time.perf_counter()
        hits, stats = search_self()                                            # exchange + K2a + K2b + K2c, hits to the host
        tms = eng.timing()
        n_launch = tms["kernel_launches"]          # K1 launches of the add + K2 launches of the search (the counter restarts with every sketch call)
        qhits = qstats = None
        if args.mode == "query":
            if resident:
                qhits, qstats = eng.dist_search_query_reads_device(sp, d_qbases.data_ptr(), qoffsets, qids)
            elif world > 1:
                qhits, qstats = eng.dist_search_query_reads(sp, qbases, qoffsets, qids)
            else:
                qhits, qstats = eng.search_query_reads(sp, qbases, qoffsets, qids)
            tq = eng.timing()
            n_launch += tq["kernel_launches"]
            xsteps += tq["xorshift_steps"]
        t_c = time.perf_counter()
        last["tm_add"] = {k: tm[k] for k in ("sketch_total_ms", "h2d_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms")}
        last["wall_add"] = (t_b - t_a) * 1e3
        if record:
            for k in ("sketch_total_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms", "h2d_ms"):
                acc[k] += tm[k]
            for k in ("index_ms", "probe_ms", "filter_ms", "gather_ms", "d2h_ms"):
                acc[k] += tms[k]
            if args.mode == "query":
                for k in ("sketch_total_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms", "probe_ms", "filter_ms", "gather_ms"):
                    acc[k] += tq[k]
            launches["n"] += n_launch
            last.update(stats=stats, qstats=qstats, steps=xsteps, wall_sketch=t_b - t_a, wall_search=t_c - t_b, n_store=eng.store_size())
        return hits, stats, qhits, qstats

    # ---- device-resident leg ----
    # the clock sampler is started before the warm-up (nvidia-smi takes a few hundred ms to come up and would otherwise do
    # so inside the timed region); only the samples taken between the two marks below are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    mark0 = sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits, stats, qhits, qstats = step(record=True)
    barrier()
    dt = time.perf_counter() - t0
    ev1.record(); torch.cuda.synchronize()
    dt_dev = ev0.elapsed_time(ev1) * 1e-3      # CUDA events around the same region (every library call ends synchronised)
    clocks = sampler.stop(mark0, sampler.mark()) if rank == 0 else None
    tt = torch.tensor([dt, dt_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_s = float(tt[0].item()) / args.steps
    kmax = torch.tensor([acc[k] / args.steps for k in TKEYS], dtype=torch.float64, device="cuda")     # balance check: slowest rank per kernel
    if world > 1:
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    k_ms_max = {k: float(v) for k, v in zip(TKEYS, kmax.tolist())}
    step_dev_s = float(tt[1].item()) / args.steps
    free_b, total_b = torch.cuda.mem_get_info()
    mem = torch.tensor([float(total_b - free_b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    res_hits, res_qhits = hits, qhits

    # ---- end-to-end leg: host buffers through the C ABI ----
    step(resident=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_hits, e_stats, e_qhits, e_qstats = step(resident=False)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / args.steps
    e2e_add = dict(last["tm_add"], wall_ms=last["wall_add"])      # K1 + H2D of the last host-buffer step (rank 0)

    # ---- parity block: proof next to the speed ----
    # digest = sha256 over the sorted integer fields + score bits of the job's whole hit set (order-independent);
    # resident and host-buffer legs must agree; for N<=2 rank 0 repeats the whole job on ONE GPU and compares digests
    # and counters; at N=1 a sample of the batch is also compared hit by hit with the CPU oracle.
    parity = None
    if not args.no_parity:
        allh = gather_hits(res_hits, dist if world > 1 else None)
        alle = gather_hits(e_hits, dist if world > 1 else None)
        allq = gather_hits(res_qhits, dist if world > 1 else None) if args.mode == "query" else None
        if rank == 0:
            parity = {"hits": int(len(allh)), "digest": hits_digest(allh), "digest_e2e": hits_digest(alle), "legs_agree": hits_digest(allh) == hits_digest(alle),
                      "counters": stats, "counters_e2e_equal": stats == e_stats}
            if allq is not None:
                parity.update(query_hits=int(len(allq)), query_digest=hits_digest(allq), query_counters=qstats)
        if world == 2 and args.mode == "self":
            # the same 2n reads on ONE GPU (rank 0), outside every timed region
            barrier()
            if rank == 0:
                eng1 = native.Engine(local)
                g2 = make_genome(args, world)
                b2, o2 = synth.reads(g2, READ_SEED(args.seed), 0, total_reads, L, args.err)
                eng1.store_reset(p)
                eng1.store_add_reads(b2, o2)
                h1, s1 = eng1.search_self(sp)
                eng1.close()
                parity.update(single_gpu_digest=hits_digest(h1), single_gpu_counters=s1,
                              single_gpu_equal=(hits_digest(h1) == parity["digest"] and s1 == stats))
                del b2, g2
            barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_scalar, peak_bs = eng.xorshift_peaks()
    peak_steps = max(peak_scalar, peak_bs)
    total_bases = (total_reads + total_queries) * L
    k_ms = {k: v / args.steps for k, v in acc.items()}
    stats = last["stats"]
    n_chunks = 1      # K1b is launched once per sketch call (one super-chunk: the key scratch of the whole shard fits)
    alg_bytes = alg_bytes_per_read(L, H, S) * n_local + (alg_bytes_per_read(L, H, S, strands=1) * nq_local if args.mode == "query" else 0)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic of one K1b launch: from the committed ncu --set full capture of the same kernel and shape, if there is one
    traffic = traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "k1b_dram_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("num_hashes") == H and tj.get("read_len") == L:
            traffic = tj["dram_bytes_per_read"] * (n_local + 0.5 * nq_local) / n_chunks
            traffic_src = tj.get("source")
    except Exception:
        pass
    ach = alg_bytes / (k_ms["minhash_ms"] * 1e-3) / 1e9
    steps_per_s = last["steps"] / (k_ms["minhash_ms"] * 1e-3)
    search_ms = k_ms["probe_ms"] + k_ms["filter_ms"]
    compared = stats["fully_compared"] + (last["qstats"]["fully_compared"] if last["qstats"] else 0)
    out = {
        "metric": "gbases_per_s_sketched_and_overlapped", "value": total_bases / step_s / 1e9, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "ms_per_step_cuda_events": step_dev_s * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.weak else "fixed job", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, world),
        "sketch_gbases_per_s": (n_local + nq_local) * L / (k_ms["sketch_total_ms"] * 1e-3) / 1e9 * world,
        "overlaps_per_s": compared / (search_ms * 1e-3) if search_ms > 0 else None,
        "kernel_ms_per_step_rank0": k_ms, "kernel_ms_per_step_max_over_ranks": k_ms_max,
        "wall_ms_rank0": {"sketch": last["wall_sketch"] * 1e3, "exchange_index_search": last["wall_search"] * 1e3},
        "counters": stats, "query_counters": last["qstats"], "n_store_rank0": last["n_store"],
        "hbm_high_water_gb_max_rank": float(mem.item()) / 1e9,
        "roofline": {"kernel": "k_minhash_bs2 (K1b)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "alg_bytes_per_launch": alg_bytes / n_chunks, "launches_per_step": n_chunks,
                     "note": "K1b is integer-issue bound, not HBM bound (about 2*H XORShift-min steps per base against ~3 bytes); see int_issue"},
        "int_issue": {"kernel": "k_minhash_bs2 (K1b)", "achieved_steps_per_s": steps_per_s, "peak_steps_per_s": peak_steps,
                      "frac": steps_per_s / peak_steps, "unit": "XORShift steps/s",
                      "peak_scalar_steps_per_s": peak_scalar, "peak_bitsliced_steps_per_s": peak_bs,
                      "peak_source": "mhapb_xorshift_peaks: the bare recurrence alone (no compare, no memory), scalar and bit-sliced forms, measured in this run; peak = the faster"},
        "e2e": {"value": total_bases / e2e_s / 1e9, "unit": "Gbases/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int((n_local + nq_local) * L + 8 * (n_local + nq_local + 2)) * world,
                "d2h_bytes_per_step": int((len(e_hits) + (len(e_qhits) if e_qhits is not None else 0)) * (12 + 32) * world + 64),
                "store_add_reads_rank0": e2e_add,
                "api": ("mhapb_store_add_reads + mhapb_search_self" if world == 1 else "mhapb_store_add_reads + mhapb_dist_search_self (NCCL inside the library)") + " (host buffers)"},
        "gpu_launches": int(launches["n"]), "clocks": clocks, "n_hits_rank0": int(len(hits)), "nccl_version": nccl_version,
        "parity": parity,
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = cpu_sample_size(args, cores)
        ts, tq, cst, chits = oracle_sample_run(bases, offsets, n, L, H, S, cores, want_hits=not args.no_parity)
        out["cpu_baseline"] = {"value": n * L / 1e9 / (ts + tq), "unit": "Gbases/s", "cores": cores, "kind": "port",
                               "sample": f"first {n} of {total_reads} reads x {L} bp, sketch+index+self-search, {cores} threads, {ts + tq:.1f} s",
                               "sketch_gbases_per_s": n * L / 1e9 / ts, "overlaps_per_s": cst["fully_compared"] / tq if tq > 0 else None}
        if parity is not None:
            # the same sample (the first n reads of rank 0's shard, a self-contained job at this config's shape) through the CUDA
            # path: complete hit set + counters against the oracle's
            e1 = native.Engine(local)
            e1.store_reset(p)
            e1.store_add_reads(bases[: n * L], offsets[: n + 1])
            gh, gs = e1.search_self(sp)
            e1.close()
            parity.update(oracle_sample_reads=n, oracle_sample_hits=int(len(chits)),
                          oracle_sample_equal=bool(gs == cst and hits_digest(gh) == hits_digest(chits)))
    print(json.dumps(out), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
