"""Multi-GPU self-overlap and store-vs-query: shard reads, sketch + index locally, all-gather sketch blocks, query everything.

The reference is one JVM ("Distributed communication backend: none"); its users
partition by hand.  Here the path shards naturally with ONE exchange step:

  1. reads are partitioned contiguously over the ranks; K1 runs per shard, no communication; every
     rank stores and indexes ONLY its own shard (K2a is 1/N of the job per rank, not replicated);
  2. the per-shard sketch blocks (min-hashes [n][H], ordered sketches [n][S][2] and the small
     per-sketch columns) are all-gathered -- NCCL over NVLink on GPUs, gloo in the CPU tests;
  3. every rank queries its local index with the forward sketches of ALL ranks under the self-search
     id rules (MinHashSearch.java:200,215-225), so each overlap (query, target) is found exactly once,
     on the rank that owns the target; the order-independent counters (MhapMain.java:572-590) are
     additive over target shards and are summed with an all-reduce.

Store-vs-query mode (`-s store -q query`, AbstractMatchSearch.findMatches(streamer) :203-285; BASELINE configs[3]) shards
the same way: every rank indexes its shard of the STORE, sketches its shard of the QUERY file (forward strand only),
the query blocks are all-gathered and every rank runs all queries against its local index without the self-search id
rules; a (query, target) pair is again found exactly once, on the target's rank.

The compute is behind a small backend protocol so the host logic can be exercised with gloo on CPU
(tests use an oracle-backed stand-in; the product backend is GpuBackend over the C ABI).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import native


@dataclass
class SketchBlock:
    """Compact sketches of one shard (valid strands only), tensors on the exchange device."""
    ids: torch.Tensor        # int64 [n]
    is_fwd: torch.Tensor     # uint8 [n]
    seq_len: torch.Tensor    # int32 [n]
    seq_len_kmers: torch.Tensor  # int32 [n]
    ord_n: torch.Tensor      # int32 [n]
    minhash: torch.Tensor    # int32 [n, H]
    ord: torch.Tensor        # int32 [n, S, 2]

    @property
    def n(self) -> int:
        return int(self.ids.shape[0])


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of [0, n_items): returns (first, count) of `rank`."""
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    if t.shape[0] == rows:
        return t.contiguous()
    out = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[: t.shape[0]] = t
    return out


def all_gather_blocks(block: SketchBlock, dist=None, overlap=None) -> tuple[SketchBlock, list[int]]:
    """All-gather the shard blocks in rank order.  Returns (global block, per-rank counts).

    overlap: optional callable run while the collectives are in flight (the local index build: it only reads the
    rank's own min-hashes, so it hides behind the gather of the 14 KB-per-sketch blocks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        if overlap is not None:
            overlap()
        return block, [block.n]
    world = dist.get_world_size()
    dev = block.minhash.device
    cnt = torch.tensor([block.n], dtype=torch.int64, device=dev)
    counts_t = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = [int(c) for c in counts_t.tolist()]
    mx = max(counts)

    pending = []

    def gather(t: torch.Tensor):
        padded = _pad_rows(t, mx)
        out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pending.append((dist.all_gather_into_tensor(out, padded, async_op=True), padded))   # keep `padded` alive
        return out

    def trim(out: torch.Tensor) -> torch.Tensor:
        if all(c == mx for c in counts):
            return out
        return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)

    raw = dict(ord=gather(block.ord), minhash=gather(block.minhash), ids=gather(block.ids), is_fwd=gather(block.is_fwd),
               seq_len=gather(block.seq_len), seq_len_kmers=gather(block.seq_len_kmers), ord_n=gather(block.ord_n))
    if overlap is not None:
        overlap()
    for w, _ in pending:
        w.wait()
    g = SketchBlock(**{k: trim(v) for k, v in raw.items()})
    return g, counts


def sharded_self_overlap(backend, bases, offsets, ids, dist=None):
    """One rank's part of a sharded self-overlap.  `bases/offsets/ids` are THIS rank's reads.

    Returns (hits whose target lives on this rank, job-wide stats dict, info dict)."""
    block = backend.store_shard(bases, offsets, ids, build_index=False)      # sketch + store the local shard
    gblock, counts = all_gather_blocks(block, dist, overlap=backend.index_build)   # the one exchange step, K2a behind it
    hits, stats = backend.search_all(gblock)                  # all forward sketches vs the local index
    stats = _reduce_stats(stats, gblock.ids.device, dist)
    return hits, stats, dict(counts=counts, n_store=gblock.n)


def _reduce_stats(stats, dev, dist):
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        keys = [k for k in sorted(stats) if k != "sequences_searched"]
        t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        stats = dict(stats, **{k: int(v) for k, v in zip(keys, t.tolist())})   # every rank searched every query once
    return stats


def sharded_query_overlap(backend, store_reads, query_reads, dist=None):
    """One rank's part of a sharded store-vs-query run.  store_reads / query_reads = (bases, offsets, ids) of THIS
    rank's shard of the store file and of the query file.

    Returns (hits whose target lives on this rank, job-wide stats dict, info dict)."""
    block = backend.store_shard(*store_reads)                 # sketch + store + index the local store shard
    qblock = backend.sketch_queries(*query_reads)             # forward-only sketches of the local query shard
    gq, counts = all_gather_blocks(qblock, dist)              # the one exchange step
    hits, stats = backend.search_all(gq, to_self=False)       # all queries vs the local index
    stats = _reduce_stats(stats, gq.ids.device, dist)
    return hits, stats, dict(query_counts=counts, n_queries=gq.n, n_store_local=block.n)


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuBackend:
    """The product backend: K1/K2 through the C ABI on this rank's GPU.  torch is only the view over
    the store's device blocks and the NCCL plumbing."""

    def __init__(self, engine: native.Engine, params: native.SketchParams, search: native.SearchParams):
        self.e, self.p, self.sp = engine, params, search
        self.dev = torch.device("cuda", engine.device)
        self.d_bases = None

    def upload(self, bases: np.ndarray):
        """H2D of this rank's read characters (pinned source recommended) for the device-resident entry point."""
        t = torch.from_numpy(bases)
        if self.d_bases is None or self.d_bases.numel() < t.numel():
            self.d_bases = torch.empty(max(1, t.numel()), dtype=torch.uint8, device=self.dev)
        self.d_bases[: t.numel()].copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()

    def index_build(self):
        self.e.index_build()

    def store_shard(self, bases, offsets, ids, resident=False, build_index=True) -> SketchBlock:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        k, ok = self.p.kmer_size, self.p.ordered_kmer_size
        ids = np.arange(1, n + 1, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        self.e.store_reset(self.p)
        if resident:
            # reads already in HBM (bench's device-resident leg): sketch into torch blocks, then hand them to the store
            H, S = self.p.num_hashes, self.p.ordered_sketch_size
            mh = torch.empty((2 * n, H), dtype=torch.int32, device=self.dev)
            od = torch.empty((2 * n, S, 2), dtype=torch.int32, device=self.dev)
            on = torch.empty(2 * n, dtype=torch.int32, device=self.dev)
            torch.cuda.current_stream(self.dev).synchronize()
            self.e.sketch_device(self.d_bases.data_ptr(), offsets, self.p, True, mh.data_ptr(), od.data_ptr(), on.data_ptr())
        else:
            self.e.store_add_reads(bases, offsets, ids)          # H2D + K1 straight into the store
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        ok_read = (lens >= self.p.min_olap_length) & (lens - k + 1 >= 1) & (lens - ok + 1 >= 1)
        v = np.nonzero(ok_read)[0]
        meta = dict(ids=np.repeat(ids[v], 2), is_fwd=np.tile(np.array([1, 0], np.uint8), v.size),
                    seq_len=np.repeat(lens[v], 2).astype(np.int32), seq_len_kmers=np.repeat(lens[v] - ok + 1, 2).astype(np.int32))
        if resident:
            if not ok_read.all():
                keep = torch.from_numpy(np.repeat(ok_read, 2)).to(self.dev)
                mh, od, on = mh[keep].contiguous(), od[keep].contiguous(), on[keep].contiguous()
            self.e.store_add_sketches_device(meta["ids"], meta["is_fwd"], meta["seq_len"], meta["seq_len_kmers"],
                                             mh.data_ptr(), od.data_ptr(), on.cpu().numpy())
        d_mh, d_od, d_on, ns, H, S = self.e.store_device_ptrs()
        assert ns == 2 * v.size
        if build_index:
            self.e.index_build()
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.dev)
        view = lambda ptr, shape: torch.as_tensor(_DevArray(ptr, shape, "<i4"), device=self.dev) if ns else torch.empty(shape, dtype=torch.int32, device=self.dev)
        return SketchBlock(ids=to(meta["ids"], np.int64), is_fwd=to(meta["is_fwd"], np.uint8), seq_len=to(meta["seq_len"], np.int32),
                           seq_len_kmers=to(meta["seq_len_kmers"], np.int32), ord_n=view(d_on, (ns,)),
                           minhash=view(d_mh, (ns, H)), ord=view(d_od, (ns, S, 2)))

    def sketch_queries(self, bases, offsets, ids) -> SketchBlock:
        """Forward-only sketches of a query shard (SequenceSketchStreamer with fwdOnly, AbstractMatchSearch.java:214),
        left on the device for the all-gather."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        k, ok = self.p.kmer_size, self.p.ordered_kmer_size
        H, S = self.p.num_hashes, self.p.ordered_sketch_size
        ids = np.arange(1, n + 1, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        self.upload(np.ascontiguousarray(bases, dtype=np.uint8))
        mh = torch.empty((n, H), dtype=torch.int32, device=self.dev)
        od = torch.empty((n, S, 2), dtype=torch.int32, device=self.dev)
        on = torch.empty(n, dtype=torch.int32, device=self.dev)
        st = torch.empty(n, dtype=torch.int32, device=self.dev)
        torch.cuda.current_stream(self.dev).synchronize()
        if n:
            self.e.sketch_device(self.d_bases.data_ptr(), offsets, self.p, False, mh.data_ptr(), od.data_ptr(), on.data_ptr(), st.data_ptr())
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        keep_np = st.cpu().numpy() == 0 if n else np.zeros(0, bool)     # status 0: long enough and (with a -f filter) not emptied
        v = np.nonzero(keep_np)[0]
        if v.size != n:
            keep = torch.from_numpy(keep_np).to(self.dev)
            mh, od, on = mh[keep].contiguous(), od[keep].contiguous(), on[keep].contiguous()
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.dev)
        return SketchBlock(ids=to(ids[v], np.int64), is_fwd=to(np.ones(v.size, np.uint8), np.uint8), seq_len=to(lens[v], np.int32),
                           seq_len_kmers=to(lens[v] - ok + 1, np.int32), ord_n=on, minhash=mh, ord=od)

    def search_all(self, g: SketchBlock, to_self: bool = True):
        torch.cuda.synchronize(self.dev)
        return self.e.search_sketches_device(self.sp, to_self, g.ids.cpu().numpy(), g.is_fwd.cpu().numpy(), g.seq_len.cpu().numpy(),
                                             g.seq_len_kmers.cpu().numpy(), g.minhash.data_ptr(), g.ord.data_ptr(),
                                             g.ord_n.data_ptr(), int(g.ord.shape[1]))
