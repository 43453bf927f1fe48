"""-m gpu: K2 parity -- index build, probe + hit counting, second-stage filter, through the C ABI,
against the CPU oracle: identical hit sets (all integers + score), identical printed lines, identical
order-independent counters."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from mhap_b200 import native, synth
from tests.gpu_common import assert_same_hits, engine, hit_key, nasty_seq, rand_seq
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _run_self(reads, H=256, S=1536, k=16, ok=12, m=3, thr=0.78, min_store=0, max_shift=0.2, unweighted=False, keep_all=True, ids=None):
    bases, offs = native.pack_reads(reads)
    p = native.SketchParams(k, H, ok, S, int(unweighted), 116)
    e = engine()
    e.store_reset(p)
    n_added = e.store_add_reads(bases, offs, ids)
    sp = native.SearchParams(m, min_store, max_shift, thr, int(keep_all), 0, 0, -1)
    hits, stats = e.search_self(sp)
    st = orc.Store(k=k, num_hashes=H, ordered_k=ok, ordered_size=S, unweighted=unweighted)
    assert st.add_reads(bases, offs, ids) == n_added == e.store_size()
    res = st.search_self(num_min_matches=m, min_store_length=min_store, max_shift=max_shift, accept_score=thr, threads=8, keep_all=keep_all)
    assert_same_hits(hits, res.hits, stats, res.stats)
    return hits, stats, st


def test_config1_self_overlap():
    # BASELINE configs[0]: 1k synthetic reads x 1 kbp, k=16, --num-hashes 256, self-vs-self
    bases, offs = synth.dataset(1000, 1000, seed=1)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(1000)]
    hits, stats, _ = _run_self(reads, H=256)
    assert stats["sequences_searched"] == 1000 and stats["fully_compared"] > 0


def test_candidate_rich_low_error():
    bases, offs = synth.dataset(300, 3000, seed=7, err=0.05)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(300)]
    hits, stats, _ = _run_self(reads, H=512, keep_all=False)
    assert stats["matches_processed"] > 300
    assert all(h["accepted"] for h in hits)


def test_ordered_sketch_1000_and_fast_preset_shape():
    bases, offs = synth.dataset(200, 8000, seed=3, err=0.10)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(200)]
    _run_self(reads, H=512, S=1000)
    _run_self(reads[:80], H=256, S=1000, ok=14, thr=0.80)            # --settings 1 (MhapMain.java:137-160)
    _run_self(reads[:80], H=768, m=2, thr=0.73)                      # --settings 3


def test_min_store_length_filters_and_custom_ids():
    rng = random.Random(3)
    g = rand_seq(rng, 4000)
    reads = []
    for i in range(60):
        L = rng.choice([300, 600, 1200])
        st = rng.randrange(0, len(g) - L)
        r = g[st:st + L]
        reads.append(r if rng.random() < 0.5 else orc.rc(r).decode())
    ids = np.array(rng.sample(range(1, 1000), 60), dtype=np.int64)
    for ms in (0, 500, 1000, 5000):
        _run_self(reads, H=128, S=300, m=2, thr=0.5, min_store=ms, ids=ids)


def test_repeat_rich_queries_overflow_to_dense_counting():
    # every read shares a long tandem repeat, so each query hits every stored sketch (> 3072 distinct targets)
    rng = random.Random(4)
    rep = "ACGGTCATTG" * 40
    reads = [rand_seq(rng, 150) + rep + rand_seq(rng, 150) for _ in range(3200)]
    hits, stats, _ = _run_self(reads, H=32, S=64, m=3, thr=0.9, keep_all=False)
    assert stats["sequences_hit"] > 3072 * 3200


def test_nasty_reads_with_duplicate_ordered_hashes():
    rng = random.Random(8)
    base = nasty_seq(rng, 3000)
    reads = []
    for _ in range(50):
        st = rng.randrange(0, 1500)
        r = list(base[st:st + 1500])
        for i in range(len(r)):
            if rng.random() < 0.03:
                r[i] = rng.choice("ACGT")
        reads.append("".join(r))
    _run_self(reads, H=128, S=400, m=2, thr=0.3, max_shift=0.3)


def test_query_mode_against_store():
    # -s store -q query: queries sketched forward only, no id-order filters, id offset = store size
    g = synth.genome(4, 40000)
    sb, so = synth.reads(g, 44, 0, 150, 2000, 0.08)
    qb, qo = synth.reads(g, 55, 0, 120, 2000, 0.08)
    p = native.SketchParams(16, 256, 12, 1536, 0, 116)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(sb, so)
    qids = np.arange(1, 121, dtype=np.int64) + 150
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    hits, stats = e.search_query_reads(sp, qb, qo, qids)
    st = orc.Store(num_hashes=256)
    st.add_reads(sb, so, threads=8)
    qs = orc.Store(num_hashes=256)
    qs.add_reads(qb, qo, ids=qids, both_strands=False, threads=8)
    res = st.search_query(qs, keep_all=True, threads=8)
    assert_same_hits(hits, res.hits, stats, res.stats)
    assert stats["matches_processed"] > 0
    # the same queries as pre-computed sketches (a .dat query file)
    mh, od, on, stt = e.sketch(qb, qo, p, both_strands=False)
    assert not stt.any()
    hits2, stats2 = e.search_query_sketches(sp, qids, np.ones(120, np.uint8), np.full(120, 2000, np.int32), np.full(120, 1989, np.int32), mh, od, on)
    assert_same_hits(hits2, res.hits, stats2, res.stats)


def test_store_from_dat_sketches_equals_store_from_reads():
    bases, offs = synth.dataset(120, 2000, seed=9, err=0.08)
    p = native.SketchParams(16, 128, 12, 500, 0, 116)
    e = engine()
    blob, nrec = e.sketch_to_dat(bases, offs, None, p)
    d = native.dat_decode(blob)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    e.store_reset(p)
    e.store_add_sketches(d["ids"], d["is_fwd"], d["seq_len"], d["seq_len_kmers"], d["minhash"], d["ord"], d["ord_n"])
    h1, s1 = e.search_self(sp)
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    h2, s2 = e.search_self(sp)
    assert_same_hits(h1, h2, s1, s2)
    got = e.store_get(5, 128, 500)
    assert got["id"] == 3 and got["is_fwd"] is False and (got["minhash"] == d["minhash"][5]).all()
    assert (got["ord"] == d["ord"][5, :d["ord_n"][5]]).all()
    assert e.minhash_equal_count(4, 4) == 128
    assert e.minhash_equal_count(4, 5) == int((d["minhash"][4] == d["minhash"][5]).sum())


def test_duplicate_ids_and_empty_store_errors():
    e = engine()
    p = native.SketchParams(16, 64, 12, 100, 0, 116)
    e.store_reset(p)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1)
    with pytest.raises(native.MhapError) as ei:
        e.search_self(sp)
    assert ei.value.code == -6
    b, o = native.pack_reads(["ACGT" * 50, "TTGCA" * 40])
    e.store_add_reads(b, o, np.array([7, 8], np.int64))
    with pytest.raises(native.MhapError) as ei:
        e.store_add_reads(b, o, np.array([9, 7], np.int64))     # MinHashSearch.java:112-117
    assert ei.value.code == -5 and e.store_size() == 4


def test_query_range_partitions_self_search():
    # multi-GPU sharding rule: disjoint query ranges over the same index reproduce the full result
    bases, offs = synth.dataset(200, 2000, seed=12, err=0.08)
    p = native.SketchParams(16, 128, 12, 500, 0, 116)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    full_h, full_s = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1))
    parts, tot = [], {k: 0 for k in full_s}
    for first in range(0, 400, 100):
        h, s = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, 1, 0, first, 100))
        parts.append(h)
        for k in s:
            tot[k] += s[k]
    assert_same_hits(np.concatenate(parts), full_h, tot, full_s)


def test_identical_reads_overflow_the_shared_match_buffer():
    # exact copies share all 1536 ordered k-mers: more match records than the warp kernel keeps in shared
    # memory, so these pairs are resolved by the thread-per-candidate kernel (same exact result)
    rng = random.Random(21)
    base = [rand_seq(rng, 4000) for _ in range(6)]
    reads = []
    for b in base:
        reads += [b, b, orc.rc(b).decode(), b[:3500] + rand_seq(rng, 500)]
    hits, stats, _ = _run_self(reads, H=128, S=1536, m=3, thr=0.5)
    assert max(int(h["valid_count"]) for h in hits) > 1024   # > kFwRecCap


def test_large_ordered_sketch_4096():
    bases, offs = synth.dataset(60, 9000, seed=31, err=0.05)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(60)]
    _run_self(reads, H=64, S=4096, m=2, thr=0.6)


def test_gpu_backend_sharded_paths_single_rank():
    # the product backend of mhap_b200/distributed.py (world size 1: no collective) against the oracle, both modes
    from mhap_b200.distributed import GpuBackend, sharded_query_overlap, sharded_self_overlap
    g = synth.genome(31, 40000)
    sb, so = synth.reads(g, 1, 0, 300, 1500, 0.06)
    qb, qo = synth.reads(g, 2, 0, 120, 1500, 0.06)
    qo = qo.copy(); qo[5:] -= 1400; qb = np.concatenate([qb[:int(qo[5])], qb[int(qo[5]) + 1400:]])   # query 4 is 100 bp: skipped
    p = native.SketchParams(16, 128, 12, 400, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    be = GpuBackend(engine(), p, sp)
    sids = np.arange(1, 301, dtype=np.int64)
    qids = np.arange(1, 121, dtype=np.int64) + 300
    ost = orc.Store(num_hashes=128, ordered_size=400)
    ost.add_reads(sb, so, ids=sids, threads=4)
    oq = orc.Store(num_hashes=128, ordered_size=400)
    oq.add_reads(qb, qo, ids=qids, both_strands=False, threads=4)
    hits, stats = sharded_query_overlap(be, (sb, so, sids), (qb, qo, qids))
    res = ost.search_query(oq, keep_all=True, threads=4)
    assert stats["sequences_searched"] == len(oq) == 119
    assert_same_hits(hits, res.hits, stats, res.stats)
    assert len(hits) > 50
    hits, stats = sharded_self_overlap(be, sb, so, sids)
    res = ost.search_self(keep_all=True, threads=4)
    assert_same_hits(hits, res.hits, stats, res.stats)
    # the collective entry points on a context without a communicator = a job of one rank (same path, no collectives)
    e = engine()
    hits, stats = e.dist_search_self(sp)
    assert_same_hits(hits, res.hits, stats, res.stats)
    hits, stats = e.dist_search_query_reads(sp, qb, qo, qids)
    res = ost.search_query(oq, keep_all=True, threads=4)
    assert_same_hits(hits, res.hits, stats, res.stats)


def test_capacity_hints_change_nothing():
    # mhapb_store_reserve / mhapb_sketch_reserve are pure optimisations (the streaming FASTA producer uses them)
    bases, offs = synth.dataset(400, 1500, seed=21, err=0.06)
    p = native.SketchParams(16, 128, 12, 300, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    h0, s0 = e.search_self(sp)
    e.store_reset(p)
    e.store_reserve(2000)
    e._ck(e.L.mhapb_sketch_reserve(e.h, C.byref(p), 1 << 20, 1000, 1))
    half = 200
    e.store_add_reads(bases[: int(offs[half])], offs[: half + 1])
    e.store_add_reads(bases[int(offs[half]):], offs[half:] - offs[half], ids=np.arange(half + 1, 401, dtype=np.int64))
    h1, s1 = e.search_self(sp)
    assert s0 == s1 and sorted(map(hit_key, h0)) == sorted(map(hit_key, h1)) and len(h0) > 100


def test_more_queries_than_stored_sketches():
    # the multi-GPU shape seen from one rank: all ranks' queries against one rank's (small) shard
    g = synth.genome(41, 30000)
    sb, so = synth.reads(g, 3, 0, 60, 2000, 0.05)
    qb, qo = synth.reads(g, 4, 0, 500, 2000, 0.05)
    p = native.SketchParams(16, 128, 12, 500, 0, 116)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(sb, so)
    qids = np.arange(1, 501, dtype=np.int64) + 60
    hits, stats = e.search_query_reads(native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1), qb, qo, ids=qids)
    ost = orc.Store(num_hashes=128, ordered_size=500)
    ost.add_reads(sb, so, threads=4)
    oq = orc.Store(num_hashes=128, ordered_size=500)
    oq.add_reads(qb, qo, ids=qids, both_strands=False, threads=4)
    res = ost.search_query(oq, keep_all=True, threads=4)
    assert_same_hits(hits, res.hits, stats, res.stats)
    assert len(hits) > 500


def test_index_rebuilds_when_the_optimistic_table_size_overflows():
    # MHAPB_INDEX_OPTIMISTIC=1: K2a sizes a word's sub-table for one slot per two sketches (low-error reads at sequencing coverage
    # hold few distinct values per word) and verifies; here every sketch is unrelated to every other, so a word has as many
    # distinct values as there are sketches: the first build overflows, the search rebuilds with two slots per sketch and must
    # give the oracle's answer.  A fresh context: the option is read at creation, the safe size is remembered once needed.
    rng = random.Random(77)
    reads = [rand_seq(rng, 400) for _ in range(1400)] + [rand_seq(rng, 400)] * 3      # 2806 sketches, and one true overlap group
    bases, offs = native.pack_reads(reads)
    p = native.SketchParams(16, 64, 12, 200, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    os.environ["MHAPB_INDEX_OPTIMISTIC"] = "1"
    try:
        e = native.Engine(0)
    finally:
        del os.environ["MHAPB_INDEX_OPTIMISTIC"]
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    hits, stats = e.search_self(sp)
    st = orc.Store(num_hashes=64, ordered_size=200)
    st.add_reads(bases, offs)
    res = st.search_self(keep_all=True, threads=8)
    assert_same_hits(hits, res.hits, stats, res.stats)
    assert len(hits) >= 3
    # the same context again (now on the safe size), and a store that fits the optimistic size
    hits2, stats2 = e.search_self(sp)
    assert stats2 == stats and len(hits2) == len(hits)
    e.close()
