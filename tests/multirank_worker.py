"""Worker of tests/test_gpu_multirank.py: one rank per GPU (launched with torch.distributed.run).

Every rank stores its shard of the reads and calls the library's collective search (NCCL inside libmhap_b200.so);
torch.distributed (gloo) only carries the 128-byte communicator id and nothing else.  Per-rank hits and the job-wide
counters are written to the output directory for the parent test to merge and compare with the 1-GPU result.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_dir, n_total, L, nq_total = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    import torch.distributed as dist
    from mhap_b200 import native, synth
    from mhap_b200.distributed import GpuBackend, bootstrap_comm, folded_shard_ranges, shard_range, sharded_query_overlap, sharded_self_overlap

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    eng = native.Engine(local)
    assert bootstrap_comm(eng, dist) == (rank, world)
    r, n, ver = eng.comm_info()
    assert (r, n) == (rank, world) and ver > 0

    p = native.SketchParams(16, 256, 12, 1536, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)       # keep_all: also the rejected candidates
    g = synth.genome(11, max(L + 1, n_total * L // 12))
    # the store shard: two folded stripes (what bench.py uses; equal K2c load per rank), ragged when 2*world does not divide n_total
    parts = folded_shard_ranges(n_total, rank, world)
    n_mine = sum(c for _, c in parts)
    bases = np.empty(n_mine * L, dtype=np.uint8)
    at = 0
    for first, cnt in parts:
        synth.reads(g, 77, first, cnt, L, 0.12, out=bases[at * L:(at + cnt) * L])
        at += cnt
    offs = np.arange(n_mine + 1, dtype=np.uint64) * np.uint64(L)
    ids = np.concatenate([np.arange(first + 1, first + cnt + 1, dtype=np.int64) for first, cnt in parts])
    be = GpuBackend(eng, p, sp)
    hits, stats = sharded_self_overlap(be, bases, offs, ids, dist)
    np.save(os.path.join(out_dir, f"self_{rank}.npy"), hits)
    t = eng.timing()
    # store-vs-query on the same store: the query file is sharded too, ids continue after the store's (MhapMain.java:537)
    qf, qc = shard_range(nq_total, rank, world)
    qb, qo = synth.reads(g, 78, qf, qc, L, 0.12)
    qids = np.arange(qf + 1, qf + qc + 1, dtype=np.int64) + n_total
    qhits, qstats = be.search_queries(qb, qo, qids)
    np.save(os.path.join(out_dir, f"query_{rank}.npy"), qhits)
    with open(os.path.join(out_dir, f"stats_{rank}.json"), "w") as f:
        json.dump(dict(self=stats, query=qstats, n_store=eng.store_size(), gather_ms=t["gather_ms"], nccl=ver), f)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
