"""CPU suite, part 3: the multi-GPU host logic (shard plan, padded all-gather of sketch blocks, per-rank
query ranges, counter all-reduce) with world_size 2 and 3 over gloo.  The compute behind the backend
protocol is the oracle here (tests may use it); on GPUs the same code runs GpuBackend over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mhap_b200.distributed import SketchBlock, all_gather_blocks, shard_range, sharded_query_overlap, sharded_self_overlap
from oracle import oracle as orc

H, S = 64, 200


class OracleBackend:
    """Oracle stand-in for GpuBackend: the local shard is stored and indexed, every rank's sketches query it."""

    def __init__(self):
        self.store = None

    def store_shard(self, bases, offsets, ids, build_index=True):
        st = orc.Store(num_hashes=H, ordered_size=S)
        st.add_reads(bases, offsets, ids=ids)
        self.store = st
        self.indexed = build_index
        return self._block(st)

    def index_build(self):
        self.indexed = True

    def sketch_queries(self, bases, offsets, ids):
        qs = orc.Store(num_hashes=H, ordered_size=S)
        qs.add_reads(bases, offsets, ids=ids, both_strands=False)
        return self._block(qs)

    @staticmethod
    def _block(st):
        rows = [st.get(i) for i in range(len(st))]
        n = len(rows)
        od = np.zeros((n, S, 2), np.int32)
        for i, r in enumerate(rows):
            od[i, :r["ord"].shape[0]] = r["ord"]
        t = torch.from_numpy
        return SketchBlock(ids=t(np.array([r["id"] for r in rows], np.int64)), is_fwd=t(np.array([r["is_fwd"] for r in rows], np.uint8)),
                           seq_len=t(np.array([r["seq_len"] for r in rows], np.int32)),
                           seq_len_kmers=t(np.array([r["seq_len_kmers"] for r in rows], np.int32)),
                           ord_n=t(np.array([r["ord"].shape[0] for r in rows], np.int32)),
                           minhash=t(np.stack([r["minhash"] for r in rows]) if n else np.zeros((0, H), np.int32)), ord=t(od))

    def search_all(self, g, to_self=True):
        assert self.indexed, "search before the index build"
        qs = orc.Store(num_hashes=H, ordered_size=S)
        for i in range(g.n):
            qs.add_sketch(int(g.ids[i]), bool(g.is_fwd[i]), int(g.seq_len[i]), g.minhash[i].numpy(), int(g.seq_len_kmers[i]),
                          g.ord[i, :int(g.ord_n[i])].numpy())
        if len(self.store) == 0:
            return np.zeros(0, orc.HIT_DTYPE), dict(elements_processed=0, sequences_hit=0, fully_compared=0, matches_processed=0,
                                                   sequences_searched=int(g.is_fwd.sum()))
        r = self.store.search_query(qs, keep_all=True, to_self=to_self)
        return r.hits, r.stats


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


def _worker(rank, world, port, n_reads, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = _reads(n_reads, 500, 5)
    first, cnt = shard_range(n_reads, rank, world)
    bases, offs = orc.pack_reads(reads[first:first + cnt])
    ids = np.arange(first + 1, first + cnt + 1, dtype=np.int64)
    hits, stats, info = sharded_self_overlap(OracleBackend(), bases, offs, ids, dist)
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
    hits, stats, info = sharded_query_overlap(OracleBackend(), (sb, so, sids), (qb, qo, qids), dist)
    q.put((rank, sorted(_key(h) for h in hits), stats, info))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_self_overlap_equals_single_process(world):
    n_reads = 41
    reads = _reads(n_reads, 500, 5)
    st = orc.Store(num_hashes=H, ordered_size=S)
    st.add_reads(*orc.pack_reads(reads))
    ref = st.search_self(keep_all=True)
    assert len(ref.hits) > 20

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_reads, q)) for r in range(world)]
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
        assert info["n_store"] == len(st) and sum(info["counts"]) == len(st)
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
        assert info["n_queries"] == len(qs) and sum(info["query_counts"]) == len(qs)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_all_gather_blocks_single_process_is_identity():
    b = SketchBlock(*(torch.zeros(3, dtype=d) for d in (torch.int64, torch.uint8, torch.int32, torch.int32, torch.int32)),
                    minhash=torch.zeros((3, 4), dtype=torch.int32), ord=torch.zeros((3, 5, 2), dtype=torch.int32))
    g, counts = all_gather_blocks(b, None)
    assert g is b and counts == [3]
