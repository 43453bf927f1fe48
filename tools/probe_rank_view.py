#!/usr/bin/env python
"""One rank's K2b / K2c load in an N-rank job, reproduced on ONE GPU (for ncu; no collectives involved).

In a job of `world` ranks a rank indexes its own `--store` reads and probes that index with the forward sketches of ALL
ranks.  What matters for the probe is (a) the ratio queries : index and (b) the genome the reads come from (its size sets
how many chance min-hash collisions a query meets): here the genome is that of the whole job (world x 100 k reads at
coverage 20), the store is the first `--store` reads of the job's stream and the queries are the next `--queries` reads.
Prints the library's own CUDA-event timings per query.  `--dry` stops before touching the GPU (data generation only).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--reads-per-rank", type=int, default=100000)
    ap.add_argument("--store", type=int, default=50000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--num-hashes", type=int, default=512)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--l2-modes", default="", help="comma list of MHAPB_K2B_L2 values to compare (0 off, 1 cache-policy loads, 2 persisting window, 3 both)")
    ap.add_argument("--dry", action="store_true")
    a = ap.parse_args()
    from mhap_b200 import synth
    L = a.read_len
    t0 = time.perf_counter()
    g = synth.genome(2, max(L + 1, a.world * a.reads_per_rank * L // 20))
    seed = (2 * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    sb, so = synth.reads(g, seed, 0, a.store, L, 0.15)
    qb, qo = synth.reads(g, seed, a.store, a.queries, L, 0.15)
    qids = np.arange(1, a.queries + 1, dtype=np.int64) + a.store
    print(f"data: genome {g.size / 1e6:.0f} Mbp, store {a.store} reads, queries {a.queries} reads, {time.perf_counter() - t0:.1f} s", flush=True)
    if a.dry:
        return
    from mhap_b200 import native
    eng = native.Engine(0)
    p = native.SketchParams(16, a.num_hashes, 12, 1536, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1)
    eng.store_reset(p)
    eng.store_add_reads(sb, so)
    from mhap_b200.distributed import hits_digest
    modes = [m for m in a.l2_modes.split(",") if m != ""] or [None]
    digests = set()
    for mode in modes:
        if mode is not None:
            os.environ["MHAPB_K2B_L2"] = mode
        for it in range(a.repeat):
            hits, st = eng.search_query_reads(sp, qb, qo, qids)
            t = eng.timing()
            nq = a.queries
            print(f"MHAPB_K2B_L2={mode} run {it}: index {t['index_ms']:.2f} ms  probe {t['probe_ms']:.2f} ms ({t['probe_ms'] * 1e6 / nq:.0f} ns/query)  "
                  f"filter {t['filter_ms']:.2f} ms  hits {len(hits)}  elements/query {st['elements_processed'] / nq:.1f}  "
                  f"distinct targets/query {st['sequences_hit'] / nq:.1f}  candidates/query {st['fully_compared'] / nq:.2f}", flush=True)
        digests.add((hits_digest(hits), tuple(sorted(st.items()))))
    print("hit sets and counters identical across modes:", len(digests) == 1, flush=True)
    eng.close()


if __name__ == "__main__":
    main()
