#!/usr/bin/env python
"""bench.py -- Gbases/s sketched + candidate overlaps/s of the MinHash overlap hot path.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference path on the host cores
Under torchrun (N>1) one rank per GPU; reads are sharded (weak scaling: every rank holds one
BASELINE configs[1] worth of reads), every rank sketches and indexes its own shard, sketch blocks are
all-gathered over NCCL, and every rank queries its index with the forward sketches of all ranks.

A "step" is one full self-overlap pass over the synthetic read set: K1 sketch (both strands) ->
K2a index -> K2b probe/count -> K2c ordered filter -> hits on the host.
  value  : bases of all ranks / step time, reads already resident in HBM when the step starts.
  e2e    : the same through the host-buffer C-ABI calls (mhapb_store_add_reads + mhapb_search_self),
           host->device copy of the reads and device->host copy of the hits inside the timed region.
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

ALG_BYTES_PER_BASE = {  # (ceil(L/4)+8 in + strands*(4H + 8*min(S,no)) out) / L
    "note": "per read: ceil(L/4)+8 bytes in, 2*(4H+8*min(S,L-ok+1)) bytes out",
}


# DRAM traffic of one K1b launch per read, from the ncu --set full capture of k_minhash_bs2<16>
# (profiles/r1r_k_minhash_bs2_ncu_summary.txt, one launch = 12 500 reads x 10 kbp: dram__bytes_read 1.8560 GB +
# dram__bytes_write 0.0497 GB): the 8-byte k-mer keys it streams in, the min-hash rows it writes
K1B_DRAM_BYTES_PER_READ = (1.854622e9 + 47.591424e6) / 12500.0


def alg_bytes_per_read(L, H, S, ok=12, strands=2):
    return (L + 3) // 4 + 8 + strands * (4 * H + 8 * min(S, L - ok + 1))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000, help="reads per GPU (BASELINE configs[1]: 100k)")
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--num-hashes", type=int, default=512)
    ap.add_argument("--ordered-sketch-size", type=int, default=1536)
    ap.add_argument("--err", type=float, default=0.15)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target length of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
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


def oracle_sample_run(bases, offsets, n_reads, L, H, S, threads):
    """One pass of the CPU path (oracle port, pthread pool like Executors.newFixedThreadPool) over a sample."""
    from oracle import oracle as orc
    st = orc.Store(k=16, num_hashes=H, ordered_k=12, ordered_size=S)
    t0 = time.perf_counter()
    st.add_reads(bases[: n_reads * L], offsets[: n_reads + 1], threads=threads)
    t1 = time.perf_counter()
    res = st.search_self(threads=threads)
    t2 = time.perf_counter()
    st.close()
    return t1 - t0, t2 - t1, res.stats


def cpu_sample_size(args, cores):
    # ~1.5 ns per XORShift-min step per core for the C port; both strands
    per_read_s = 2.0 * (args.read_len - 15) * args.num_hashes * 1.5e-9
    n = int(args.cpu_seconds * cores / per_read_s)
    return max(cores, min(args.reads, n))


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mhap_b200 import synth
    cores = os.cpu_count() or 1
    n = cpu_sample_size(args, cores)
    total_reads = args.reads * args.gpus
    glen = max(args.read_len + 1, int(total_reads * args.read_len / 20))
    g = synth.genome(args.seed, glen)
    bases, offsets = synth.reads(g, (args.seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, 0, n, args.read_len, args.err)
    times = []
    stats = None
    for it in range(args.warmup + args.steps):
        ts, tq, stats = oracle_sample_run(bases, offsets, n, args.read_len, args.num_hashes, args.ordered_sketch_size, cores)
        if it >= args.warmup:
            times.append((ts, tq))
    ts = float(np.mean([t[0] for t in times])); tq = float(np.mean([t[1] for t in times]))
    gb = n * args.read_len / 1e9
    val = gb / (ts + tq)
    sample = f"first {n} of {total_reads} reads x {args.read_len} bp (same generator/seed), sketch+index+self-search, {cores} threads"
    out = {
        "impl": "reference", "metric": "gbases_per_s_sketched_and_overlapped", "value": val, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (ts + tq) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args),
        "sketch_gbases_per_s": gb / ts, "overlaps_per_s": stats["fully_compared"] / tq if tq > 0 else None,
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the Java path (oracle/); the JVM reference cannot run in this image"},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=real_stdout, flush=True)


def workload_config(args):
    return {"workload": f"BASELINE configs[1]: {args.reads} synthetic PacBio-shape reads x {args.read_len} bp per GPU, k=16, "
                        f"--num-hashes {args.num_hashes}, self-overlap (both strands sketched)",
            "reads_per_gpu": args.reads, "read_len": args.read_len, "k": 16, "num_hashes": args.num_hashes,
            "ordered_kmer_size": 12, "ordered_sketch_size": args.ordered_sketch_size, "error_rate": args.err,
            "coverage": 20, "seed": args.seed, "l2_policy": "inputs larger than L2 (1 GB of reads per GPU per step, 2.9 GB of sketches)",
            "parallelism": f"reads sharded over {args.gpus} GPU(s); each rank indexes its shard; all-gather of sketch blocks; every rank queries its index with all forward sketches"}


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
    from mhap_b200.distributed import GpuBackend, all_gather_blocks

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
    L, H, S = args.read_len, args.num_hashes, args.ordered_sketch_size
    p = native.SketchParams(16, H, 12, S, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1)
    n_local = args.reads
    total_reads = n_local * world
    glen = max(L + 1, int(total_reads * L / 20))
    g = synth.genome(args.seed, glen)
    host = torch.empty(n_local * L, dtype=torch.uint8).pin_memory()
    bases = host.numpy()
    _, offsets = synth.reads(g, (args.seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, rank * n_local, n_local, L, args.err, out=bases)
    ids = np.arange(1, n_local + 1, dtype=np.int64) + rank * n_local
    del g

    be = GpuBackend(eng, p, sp)
    be.upload(bases)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    acc = {k: 0.0 for k in ("sketch_total_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms", "index_ms", "probe_ms", "filter_ms")}
    launches = {"n": 0, "minhash": 0}
    last = {}

    def device_step(resident=True, record=False):
        t_a = time.perf_counter()
        block = be.store_shard(bases, offsets, ids, resident=resident, build_index=False)      # K1 on the local shard
        # the one exchange step; K2a (index build over the rank's own min-hashes) runs behind the collectives
        gblock, counts = all_gather_blocks(block, dist if world > 1 else None, overlap=be.index_build)
        tm = eng.timing()
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        hits, stats = be.search_all(gblock)                                  # K2b + K2c: all forward sketches vs local index
        tm_s = eng.timing()
        if world > 1:
            keys = [k for k in sorted(stats) if k != "sequences_searched"]
            t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            stats = dict(stats, **{k: int(v) for k, v in zip(keys, t.tolist())})
        t_c = time.perf_counter()
        if record:
            for k in ("sketch_total_ms", "hash_dedup_ms", "minhash_ms", "ordered_ms", "index_ms"):
                acc[k] += tm[k]
            acc["probe_ms"] += tm_s["probe_ms"]; acc["filter_ms"] += tm_s["filter_ms"]
            launches["n"] += tm_s["kernel_launches"]
            last.update(stats=stats, n_hits=len(hits), steps=tm["xorshift_steps"], wall_sketch=t_b - t_a, wall_search=t_c - t_b,
                        n_store=gblock.n)
        return hits, stats

    # ---- device-resident leg ----
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        device_step(record=True)
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_s = float(tt.item()) / args.steps

    # ---- end-to-end leg: host buffers through the C ABI ----
    def e2e_step():
        if world == 1:
            eng.store_reset(p)
            eng.store_add_reads(bases, offsets, ids)        # H2D of the reads inside
            return eng.search_self(sp)                       # D2H of candidates/overlaps inside
        return device_step(resident=False)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_hits = 0
    for _ in range(args.steps):
        hits, stats = e2e_step()
        n_hits = len(hits)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_scalar, peak_bs = eng.xorshift_peaks()
    peak_steps = max(peak_scalar, peak_bs)
    total_bases = total_reads * L
    k_ms = {k: v / args.steps for k, v in acc.items()}
    stats = last["stats"]
    n_chunks = max(1, -(-(n_local * 2 * (L - 15)) // (256 << 20)))
    alg_bytes = alg_bytes_per_read(L, H, S) * n_local
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = alg_bytes / (k_ms["minhash_ms"] * 1e-3) / 1e9
    steps_per_s = last["steps"] / (k_ms["minhash_ms"] * 1e-3)
    out = {
        "metric": "gbases_per_s_sketched_and_overlapped", "value": total_bases / step_s / 1e9, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args),
        "sketch_gbases_per_s": n_local * L / (k_ms["sketch_total_ms"] * 1e-3) / 1e9 * world,
        "overlaps_per_s": stats["fully_compared"] / ((k_ms["probe_ms"] + k_ms["filter_ms"]) * 1e-3) if k_ms["probe_ms"] + k_ms["filter_ms"] > 0 else None,
        "kernel_ms_per_step_rank0": k_ms,
        "wall_ms_rank0": {"sketch_index_gather": last["wall_sketch"] * 1e3, "search": last["wall_search"] * 1e3},
        "counters": stats, "n_store": last["n_store"],
        "roofline": {"kernel": "k_minhash (K1b)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "traffic": K1B_DRAM_BYTES_PER_READ * n_local / n_chunks, "traffic_source": "ncu --set full, profiles/r1r_k_minhash_bs2_ncu_summary.txt: dram read+write of one launch / its reads",
                     "alg_bytes_per_launch": alg_bytes / n_chunks, "launches_per_step": n_chunks,
                     "note": "K1b is integer-issue bound, not HBM bound (about 2*H XORShift-min steps per base against ~3 bytes); see int_issue"},
        "int_issue": {"kernel": "k_minhash (K1b)", "achieved_steps_per_s": steps_per_s, "peak_steps_per_s": peak_steps,
                      "frac": steps_per_s / peak_steps, "unit": "XORShift steps/s",
                      "peak_scalar_steps_per_s": peak_scalar, "peak_bitsliced_steps_per_s": peak_bs,
                      "peak_source": "mhapb_xorshift_peaks: the bare recurrence alone (no compare, no memory), scalar and bit-sliced forms, measured in this run; peak = the faster"},
        "e2e": {"value": total_bases / e2e_s / 1e9, "unit": "Gbases/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(n_local * L + 8 * (n_local + 1)) * world,
                "d2h_bytes_per_step": int(stats["fully_compared"] * (12 + 32) + 24),
                "api": "mhapb_store_add_reads + mhapb_search_self (host buffers)" if world == 1 else "H2D + sketch_device + NCCL all-gather + store/search"},
        "gpu_launches": int(launches["n"]), "clocks": clocks, "n_hits_rank0": n_hits,
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = cpu_sample_size(args, cores)
        ts, tq, cst = oracle_sample_run(bases, offsets, n, L, H, S, cores)
        out["cpu_baseline"] = {"value": n * L / 1e9 / (ts + tq), "unit": "Gbases/s", "cores": cores, "kind": "port",
                               "sample": f"first {n} of {total_reads} reads x {L} bp, sketch+index+self-search, {cores} threads, {ts + tq:.1f} s",
                               "sketch_gbases_per_s": n * L / 1e9 / ts, "overlaps_per_s": cst["fully_compared"] / tq if tq > 0 else None}
    print(json.dumps(out), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
