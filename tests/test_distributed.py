"""CPU suite, part 3: the multi-GPU host logic with world_size 2 and 3 over gloo -- shard plan, communicator bootstrap
(rank 0's 128-byte id reaches every rank), the sharded self / store-vs-query plan reproducing the single-process result,
disjoint per-rank hit sets, merged hits and digests.  The compute behind the backend protocol is the oracle here
(tests/dist_standin.py); on GPUs the same callers run GpuBackend, whose exchange + search is one library call over NCCL
(tests/test_gpu_multirank.py checks that path on hardware)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mhap_b200 import native
from mhap_b200.distributed import bootstrap_comm, folded_shard_ranges, gather_hits, hits_digest, shard_range, sharded_query_overlap, sharded_self_overlap
from oracle import oracle as orc
from tests.dist_standin import OracleBackend, SketchBlock, all_gather_blocks

H, S = 64, 200


class FakeEngine:
    """Records what bootstrap_comm hands to mhapb_comm_init_rank (no GPU in this container)."""

    def comm_init_rank(self, comm_id, rank, nranks):
        self.joined = (bytes(comm_id), rank, nranks)


def _reads(n, L, seed, genome_seed=None):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=6000)
    if genome_seed is not None:          # reads of another file drawn from the same genome (store vs query)
        g = np.random.default_rng(genome_seed).integers(0, 4, size=6000)
    out = []
    for i in range(n):
        ln = L if i % 7 else 90          # some reads below min-olap are skipped (ragged shard sizes)
        st = int(rng.integers(0, g.size - L))
        r = g[st:st + ln].copy()
        mut = rng.random(ln) < 0.04
        r[mut] = rng.integers(0, 4, size=int(mut.sum()))
        out.append(bytes(np.frombuffer(b"ACGT", np.uint8)[r]))
    return out


def _key(h):
    return tuple(int(h[k]) for k in ("from_id", "to_id", "to_fwd", "hit_count", "a1", "a2", "b1", "b2", "valid_count", "intersect", "kmin", "accepted"))


def _worker(rank, world, port, n_reads, folded, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = _reads(n_reads, 500, 5)
    parts = folded_shard_ranges(n_reads, rank, world) if folded else [shard_range(n_reads, rank, world)]
    bases, offs = orc.pack_reads([r for first, cnt in parts for r in reads[first:first + cnt]])
    ids = np.concatenate([np.arange(first + 1, first + cnt + 1, dtype=np.int64) for first, cnt in parts])
    eng = FakeEngine()
    assert bootstrap_comm(eng, dist, make_id=lambda: bytes(range(128))) == (rank, world)
    assert eng.joined == (bytes(range(128)), rank, world)        # rank 0's id reached every rank
    be = OracleBackend(H, S, dist)
    hits, stats = sharded_self_overlap(be, bases, offs, ids, dist)
    merged = gather_hits(hits, dist)
    info = dict(be.info, digest=hits_digest(merged) if rank == 0 else None, n_merged=len(merged) if rank == 0 else None)
    q.put((rank, sorted(_key(h) for h in hits), stats, info))
    dist.barrier()
    dist.destroy_process_group()


def _worker_query(rank, world, port, n_store, n_query, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store, query = _reads(n_store, 500, 5, genome_seed=99), _reads(n_query, 500, 6, genome_seed=99)
    f, c = shard_range(n_store, rank, world)
    sb, so = orc.pack_reads(store[f:f + c])
    sids = np.arange(f + 1, f + c + 1, dtype=np.int64)
    f2, c2 = shard_range(n_query, rank, world)
    qb, qo = orc.pack_reads(query[f2:f2 + c2])
    qids = np.arange(f2 + 1, f2 + c2 + 1, dtype=np.int64) + n_store      # main/MhapMain.java:537 id offset of the query file
    be = OracleBackend(H, S, dist)
    hits, stats = sharded_query_overlap(be, (sb, so, sids), (qb, qo, qids), dist)
    q.put((rank, sorted(_key(h) for h in hits), stats, be.info))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,folded", [(2, False), (3, False), (2, True), (3, True)])
def test_sharded_self_overlap_equals_single_process(world, folded):
    n_reads = 41
    reads = _reads(n_reads, 500, 5)
    st = orc.Store(num_hashes=H, ordered_size=S)
    st.add_reads(*orc.pack_reads(reads))
    ref = st.search_self(keep_all=True)
    assert len(ref.hits) > 20
    n_fwd = sum(1 for i in range(len(st)) if st.get(i)["is_fwd"])

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_reads, folded, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_hits = sorted(k for _, hits, _, _ in res for k in hits)
    assert all_hits == sorted(_key(h) for h in ref.hits)
    for rank, hits, stats, info in res:
        assert stats == ref.stats                      # all-reduced, job-wide counters on every rank
        assert info["n_queries"] == n_fwd and sum(info["counts"]) == n_fwd      # only forward sketches travel
        if rank == 0:
            assert info["n_merged"] == len(ref.hits) and info["digest"] == hits_digest(ref.hits)
    # hit lists are disjoint by target (toId)
    owners = {}
    for rank, hits, _, _ in res:
        for k in hits:
            assert owners.setdefault(k[1], rank) == rank


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_store_vs_query_equals_single_process(world):
    # BASELINE configs[3] shape (-s store -q query) at oracle scale: the query file's sketches are all-gathered,
    # every rank searches its shard of the store
    n_store, n_query = 37, 23
    store, query = _reads(n_store, 500, 5, genome_seed=99), _reads(n_query, 500, 6, genome_seed=99)
    st = orc.Store(num_hashes=H, ordered_size=S)
    st.add_reads(*orc.pack_reads(store))
    qs = orc.Store(num_hashes=H, ordered_size=S)
    qs.add_reads(*orc.pack_reads(query), ids=np.arange(1, n_query + 1, dtype=np.int64) + n_store, both_strands=False)
    ref = st.search_query(qs, keep_all=True)
    assert len(ref.hits) > 20

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_query, args=(r, world, port, n_store, n_query, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(k for _, hits, _, _ in res for k in hits) == sorted(_key(h) for h in ref.hits)
    for rank, hits, stats, info in res:
        assert stats == ref.stats
        assert info["n_queries"] == len(qs) and sum(info["counts"]) == len(qs)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_folded_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 100, 100000):
        for w in (1, 2, 3, 8):
            parts = [folded_shard_ranges(n, r, w) for r in range(w)]
            flat = sorted(p for pr in parts for p in pr)
            assert sum(c for _, c in flat) == n
            assert all(flat[i][0] + flat[i][1] == flat[i + 1][0] for i in range(len(flat) - 1))      # disjoint, no hole
            assert all(pr[0][0] + pr[0][1] <= pr[1][0] for pr in parts)                                # ids ascend within a rank
    # the point of the folding: every rank stores the same number of (stored id < other id) pairs
    n, w = 80000, 8
    load = [sum(cnt * (n - first) - cnt * (cnt + 1) // 2 for first, cnt in folded_shard_ranges(n, r, w)) for r in range(w)]
    assert max(load) - min(load) <= 1e-9 * max(load)
    contiguous = [cnt * (n - first) - cnt * (cnt + 1) // 2 for first, cnt in (shard_range(n, r, w) for r in range(w))]
    assert max(contiguous) > 1.8 * (sum(contiguous) / w)


def test_all_gather_blocks_single_process_is_identity():
    b = SketchBlock(*(torch.zeros(3, dtype=d) for d in (torch.int64, torch.uint8, torch.int32, torch.int32, torch.int32)),
                    minhash=torch.zeros((3, 4), dtype=torch.int32), ord=torch.zeros((3, 5, 2), dtype=torch.int32))
    g, counts = all_gather_blocks(b, None)
    assert g is b and counts == [3]


def test_bootstrap_without_process_group_is_single_rank():
    assert bootstrap_comm(FakeEngine(), None) == (0, 1)


def test_hits_digest_is_order_independent():
    rng = np.random.default_rng(1)
    h = np.zeros(50, dtype=native.HIT_DTYPE)
    for f in ("from_id", "to_id", "a1", "a2", "b1", "b2", "hit_count"):
        h[f] = rng.integers(0, 1000, 50)
    h["score"] = rng.random(50)
    d = hits_digest(h)
    assert hits_digest(h[rng.permutation(50)]) == d
    h2 = h.copy(); h2["a1"][7] += 1
    assert hits_digest(h2) != d
