"""Multi-GPU self-overlap: shard reads, sketch locally, all-gather sketch blocks, query own shard.

The reference is one JVM ("Distributed communication backend: none"); its users
partition by hand.  Here the path shards naturally with ONE exchange step:

  1. reads are partitioned contiguously over the ranks; K1 runs per shard, no communication;
  2. the per-shard sketch blocks (min-hashes [n][H], ordered sketches [n][S][2] and the small
     per-sketch columns) are all-gathered -- NCCL over NVLink on GPUs, gloo in the CPU tests;
  3. every rank builds the full inverted index (replicated) and queries only its own shard's
     forward sketches, so the hit lists are disjoint by fromId and the order-independent counters
     (MhapMain.java:572-590) are summed with an all-reduce.

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


def all_gather_blocks(block: SketchBlock, dist=None) -> tuple[SketchBlock, list[int]]:
    """All-gather the shard blocks in rank order.  Returns (global block, per-rank counts)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return block, [block.n]
    world = dist.get_world_size()
    dev = block.minhash.device
    cnt = torch.tensor([block.n], dtype=torch.int64, device=dev)
    counts_t = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = [int(c) for c in counts_t.tolist()]
    mx = max(counts)

    def gather(t: torch.Tensor) -> torch.Tensor:
        padded = _pad_rows(t, mx)
        out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, padded)
        if all(c == mx for c in counts):
            return out
        return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)

    g = SketchBlock(ids=gather(block.ids), is_fwd=gather(block.is_fwd), seq_len=gather(block.seq_len),
                    seq_len_kmers=gather(block.seq_len_kmers), ord_n=gather(block.ord_n),
                    minhash=gather(block.minhash), ord=gather(block.ord))
    return g, counts


def sharded_self_overlap(backend, bases, offsets, ids, dist=None):
    """One rank's part of a sharded self-overlap.  `bases/offsets/ids` are THIS rank's reads.

    Returns (hits of this rank's queries, job-wide stats dict, info dict)."""
    block = backend.sketch_shard(bases, offsets, ids)
    gblock, counts = all_gather_blocks(block, dist)
    rank = dist.get_rank() if (dist is not None and dist.is_initialized()) else 0
    backend.load_store(gblock)
    first = sum(counts[:rank])
    hits, stats = backend.search_range(first, counts[rank])
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        keys = sorted(stats)
        t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=gblock.ids.device)
        dist.all_reduce(t)
        stats = {k: int(v) for k, v in zip(keys, t.tolist())}
    return hits, stats, dict(counts=counts, first=first, n_store=gblock.n)


class GpuBackend:
    """The product backend: K1/K2 through the C ABI on this rank's GPU, blocks in torch CUDA tensors
    (torch is only the allocator and the NCCL plumbing here)."""

    def __init__(self, engine: native.Engine, params: native.SketchParams, search: native.SearchParams):
        self.e, self.p, self.sp = engine, params, search
        self.dev = torch.device("cuda", engine.device)
        self.d_bases = None

    def upload(self, bases: np.ndarray):
        """H2D of this rank's read characters (pinned source recommended)."""
        t = torch.from_numpy(bases)
        if self.d_bases is None or self.d_bases.numel() < t.numel():
            self.d_bases = torch.empty(max(1, t.numel()), dtype=torch.uint8, device=self.dev)
        self.d_bases[: t.numel()].copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()

    def sketch_shard(self, bases, offsets, ids, resident=False) -> SketchBlock:
        if not resident:
            self.upload(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        H, S, k, ok = self.p.num_hashes, self.p.ordered_sketch_size, self.p.kmer_size, self.p.ordered_kmer_size
        mh = torch.empty((2 * n, H), dtype=torch.int32, device=self.dev)
        od = torch.empty((2 * n, S, 2), dtype=torch.int32, device=self.dev)
        on = torch.empty(2 * n, dtype=torch.int32, device=self.dev)
        torch.cuda.current_stream(self.dev).synchronize()
        self.e.sketch_device(self.d_bases.data_ptr(), offsets, self.p, True, mh.data_ptr(), od.data_ptr(), on.data_ptr())
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        ok_read = (lens >= self.p.min_olap_length) & (lens - k + 1 >= 1) & (lens - ok + 1 >= 1)
        ids = np.arange(1, n + 1, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        if not ok_read.all():
            keep = torch.from_numpy(np.repeat(ok_read, 2)).to(self.dev)
            mh, od, on = mh[keep].contiguous(), od[keep].contiguous(), on[keep].contiguous()
        v = np.nonzero(ok_read)[0]
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.dev)
        return SketchBlock(ids=to(np.repeat(ids[v], 2), np.int64), is_fwd=to(np.tile([1, 0], v.size), np.uint8),
                           seq_len=to(np.repeat(lens[v], 2), np.int32), seq_len_kmers=to(np.repeat(lens[v] - ok + 1, 2), np.int32),
                           ord_n=on, minhash=mh, ord=od)

    def load_store(self, g: SketchBlock):
        self.e.store_reset(self.p)
        torch.cuda.synchronize(self.dev)
        self.e.store_add_sketches_device(g.ids.cpu().numpy(), g.is_fwd.cpu().numpy(), g.seq_len.cpu().numpy(),
                                         g.seq_len_kmers.cpu().numpy(), g.minhash.data_ptr(), g.ord.data_ptr(), g.ord_n.cpu().numpy())
        self.e.index_build()

    def search_range(self, first: int, count: int):
        sp = native.SearchParams(self.sp.num_min_matches, self.sp.min_store_length, self.sp.max_shift, self.sp.accept_score,
                                 self.sp.keep_all, 0, first, count)
        return self.e.search_self(sp)
