"""Test infrastructure: an oracle-backed stand-in for mhap_b200.distributed.GpuBackend.

On GPUs the exchange and the sharded search are ONE library call (mhapb_dist_search_self, csrc/dist.cu: NCCL).  This
stand-in restates that plan on CPU so that the host logic around it (shard plan, communicator bootstrap, hit merging,
digests) can run with world_size > 1 over gloo: every rank stores and indexes its shard in an oracle Store, the forward
sketches of all ranks are all-gathered with torch.distributed, every rank runs all of them against its local store and
the counters are all-reduced.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from oracle import oracle as orc


@dataclass
class SketchBlock:
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


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    if t.shape[0] == rows:
        return t.contiguous()
    out = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype)
    out[: t.shape[0]] = t
    return out


def all_gather_blocks(block: SketchBlock, dist=None):
    """All-gather the shard blocks in rank order (ragged shards are padded, then trimmed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return block, [block.n]
    world = dist.get_world_size()
    cnt = torch.tensor([block.n], dtype=torch.int64)
    counts_t = torch.zeros(world, dtype=torch.int64)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = [int(c) for c in counts_t.tolist()]
    mx = max(counts)

    def gather(t):
        out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype)
        dist.all_gather_into_tensor(out, _pad_rows(t, mx))
        return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)

    g = SketchBlock(**{k: gather(getattr(block, k)) for k in ("ids", "is_fwd", "seq_len", "seq_len_kmers", "ord_n", "minhash", "ord")})
    return g, counts


class OracleBackend:
    """Same three methods as GpuBackend (store_shard / search_self / search_queries)."""

    def __init__(self, H, S, dist=None):
        self.H, self.S, self.dist = H, S, dist
        self.store = None
        self.info = {}

    def store_shard(self, bases, offsets, ids):
        st = orc.Store(num_hashes=self.H, ordered_size=self.S)
        n = st.add_reads(bases, offsets, ids=ids)
        self.store = st
        return n

    def _block(self, st, fwd_only):
        rows = [st.get(i) for i in range(len(st))]
        if fwd_only:
            rows = [r for r in rows if r["is_fwd"]]
        n = len(rows)
        od = np.zeros((n, self.S, 2), np.int32)
        for i, r in enumerate(rows):
            od[i, :r["ord"].shape[0]] = r["ord"]
        t = torch.from_numpy
        return SketchBlock(ids=t(np.array([r["id"] for r in rows], np.int64)), is_fwd=t(np.array([r["is_fwd"] for r in rows], np.uint8)),
                           seq_len=t(np.array([r["seq_len"] for r in rows], np.int32)),
                           seq_len_kmers=t(np.array([r["seq_len_kmers"] for r in rows], np.int32)),
                           ord_n=t(np.array([r["ord"].shape[0] for r in rows], np.int32)),
                           minhash=t(np.stack([r["minhash"] for r in rows]) if n else np.zeros((0, self.H), np.int32)), ord=t(od))

    def _search_all(self, g, counts, to_self):
        qs = orc.Store(num_hashes=self.H, ordered_size=self.S)
        for i in range(g.n):
            qs.add_sketch(int(g.ids[i]), bool(g.is_fwd[i]), int(g.seq_len[i]), g.minhash[i].numpy(), int(g.seq_len_kmers[i]),
                          g.ord[i, :int(g.ord_n[i])].numpy())
        if len(self.store) == 0:
            hits, stats = np.zeros(0, orc.HIT_DTYPE), dict(elements_processed=0, sequences_hit=0, fully_compared=0, matches_processed=0,
                                                         sequences_searched=g.n)
        else:
            r = self.store.search_query(qs, keep_all=True, to_self=to_self)
            hits, stats = r.hits, r.stats
        dist = self.dist
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            keys = [k for k in sorted(stats) if k != "sequences_searched"]
            t = torch.tensor([stats[k] for k in keys], dtype=torch.int64)
            dist.all_reduce(t)
            stats = dict(stats, **{k: int(v) for k, v in zip(keys, t.tolist())})   # every rank searched every query once
        self.info = dict(counts=counts, n_queries=g.n)
        return hits, stats

    def search_self(self):
        g, counts = all_gather_blocks(self._block(self.store, fwd_only=True), self.dist)
        return self._search_all(g, counts, to_self=True)

    def search_queries(self, bases, offsets, ids):
        qs = orc.Store(num_hashes=self.H, ordered_size=self.S)
        qs.add_reads(bases, offsets, ids=ids, both_strands=False)
        g, counts = all_gather_blocks(self._block(qs, fwd_only=True), self.dist)
        return self._search_all(g, counts, to_self=False)
