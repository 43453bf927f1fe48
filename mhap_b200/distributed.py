"""One-rank-per-GPU callers of the library's multi-GPU entry points (include/mhap_b200.h, "multi-GPU").

The data plane is NOT here: the NCCL communicator, the all-gather of sketch blocks and the sharded search live in
libmhap_b200.so (csrc/dist.cu; mhapb_comm_*, mhapb_dist_*).  What a launcher has to do per rank is small:

  1. hand rank 0's communicator id to every rank (`bootstrap_comm`: 128 bytes over whatever the launcher has --
     torch.distributed here, gloo or nccl backend alike);
  2. give the rank ITS shard of the reads (`shard_range`) and call mhapb_store_add_reads on its context;
  3. call the collective search (mhapb_dist_search_self / mhapb_dist_search_query_reads): hits whose target the rank
     stores come back, with the job-wide counters of MhapMain.outputFinalStat (main/MhapMain.java:572-590);
  4. (optional) merge the per-rank hit sets on one rank for output (`gather_hits`) or compare them by digest (`hits_digest`).

The reference is one JVM ("Distributed communication backend: none"); its users partition by hand.
The backend protocol (store_shard / search_self / search_queries) exists so that steps 1, 2 and 4 -- the host logic --
can be exercised with gloo on CPU: tests/dist_standin.py implements it over the oracle with a gloo all-gather.
"""
from __future__ import annotations

import hashlib

import numpy as np

from . import native

KEY_FIELDS = ("from_id", "to_id", "from_fwd", "to_fwd", "hit_count", "a1", "a2", "b1", "b2", "valid_count", "intersect", "kmin",
              "from_len", "to_len", "accepted")


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of [0, n_items): returns (first, count) of `rank`."""
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def folded_shard_ranges(n_items: int, rank: int, world: int) -> list[tuple[int, int]]:
    """Two (first, count) ranges for `rank`: stripes `rank` and `2*world-1-rank` of [0, n_items) cut into 2*world stripes.

    Why not one contiguous range: in a self-overlap the pair {a, b} is reported once, by the rank that STORES the read with
    the lower id (MinHashSearch.java:215-219 keeps `to` < `from`), so the rank holding the lowest ids would score almost
    all of its reads' overlaps and the rank holding the highest ids almost none.  Stripe s stores reads whose share of
    higher-id partners is 1-(s+.5)/(2*world); stripes r and 2*world-1-r sum to exactly 1, every rank gets the same K2c load.
    Ids stay ascending within the rank (lower stripe first), the hit set of the job does not depend on the partition."""
    cuts = [n_items * s // (2 * world) for s in range(2 * world + 1)]
    lo, hi = rank, 2 * world - 1 - rank
    return [(cuts[lo], cuts[lo + 1] - cuts[lo]), (cuts[hi], cuts[hi + 1] - cuts[hi])]


def bootstrap_comm(engine, dist=None, make_id=native.comm_unique_id):
    """Join this rank's context to the job's communicator.  Rank 0 makes the NCCL unique id inside the library
    (mhapb_comm_unique_id); `dist` (an initialised torch.distributed, any backend) only carries its 128 bytes."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 1
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = torch.zeros(native.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
    dist.broadcast(buf, src=0)
    engine.comm_init_rank(bytes(buf.cpu().numpy().tobytes()), rank, world)
    return rank, world


def sorted_hits(h: np.ndarray) -> np.ndarray:
    """Hits in the canonical order of their integer key."""
    return h[np.lexsort(tuple(h[f] for f in reversed(KEY_FIELDS)))]


def hits_digest(h: np.ndarray) -> str:
    """Order-independent digest of a hit set: sha256 over the sorted integer keys and the score bits."""
    s = sorted_hits(h)
    m = hashlib.sha256()
    for f in KEY_FIELDS:
        m.update(np.ascontiguousarray(s[f]).astype("<i8").tobytes())
    m.update(np.ascontiguousarray(s["score"]).astype("<f8").tobytes())
    return m.hexdigest()[:16]


def gather_hits(hits: np.ndarray, dist=None) -> np.ndarray | None:
    """Union of the per-rank hit sets on rank 0 (None elsewhere).  The sets are disjoint by construction: a pair is
    reported by the rank that stores its target."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return hits
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(hits, parts, dst=0)
    if dist.get_rank() != 0:
        return None
    return np.concatenate(parts) if parts else hits


def sharded_self_overlap(backend, bases, offsets, ids, dist=None):
    """One rank's part of a sharded self-overlap.  `bases/offsets/ids` are THIS rank's reads.
    Returns (hits whose target lives on this rank, job-wide stats dict)."""
    backend.store_shard(bases, offsets, ids)
    return backend.search_self()


def sharded_query_overlap(backend, store_reads, query_reads, dist=None):
    """One rank's part of a sharded store-vs-query run (`-s store -q query`, AbstractMatchSearch.java:203-285).
    store_reads / query_reads = (bases, offsets, ids) of THIS rank's shard of the store file and of the query file."""
    backend.store_shard(*store_reads)
    return backend.search_queries(*query_reads)


class GpuBackend:
    """The product backend: this rank's context; every method is one call of the C ABI."""

    def __init__(self, engine: native.Engine, params: native.SketchParams, search: native.SearchParams):
        self.e, self.p, self.sp = engine, params, search
        _, self.world, _ = engine.comm_info()

    def store_shard(self, bases, offsets, ids):
        self.e.store_reset(self.p)
        return self.e.store_add_reads(bases, offsets, ids)       # H2D + K1 straight into the rank's store

    def search_self(self):
        if self.world == 1:
            return self.e.search_self(self.sp)
        return self.e.dist_search_self(self.sp)                   # exchange + K2a/K2b/K2c, collective

    def search_queries(self, bases, offsets, ids):
        if self.world == 1:
            return self.e.search_query_reads(self.sp, bases, offsets, ids)
        return self.e.dist_search_query_reads(self.sp, bases, offsets, ids)
