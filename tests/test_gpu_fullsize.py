"""-m gpu: BASELINE-size runs checked through size-independent properties (the oracle would need ~10 CPU
minutes per config) plus oracle parity on a sample of the same batch.

  configs[1] 100k x 10 kbp, H=512          : strand symmetry, sampled sketch parity, candidate symmetry
  configs[2] 50k x 8 kbp, S=1000 (sampled) : parity with --ordered-sketch-size 1000
  configs[3] store/query mode (scaled)     : query-mode parity with id offsets
  configs[4] 15 kbp, H=1024 (sampled)      : parity at the human-shard sketch shape
"""
import numpy as np
import pytest

from mhap_b200 import native, synth
import os

from tests.gpu_common import assert_same_hits, assert_same_hits_bulk, engine
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _sample_parity(bases, offs, L, idx, mh, od, on, H, S):
    for i in idx:
        r = bytes(bases[int(offs[i]):int(offs[i + 1])])
        for s, seq in enumerate((r, orc.rc(r))):
            j = 2 * i + s
            assert (mh[j] == orc.minhash_sketch(seq, 16, H)).all(), (i, s)
            eod, _ = orc.bottom_sketch(seq, 12, S)
            assert on[j] == eod.shape[0] and (od[j, :on[j]] == eod).all(), (i, s)


def test_config2_full_size_properties_and_sampled_parity():
    n, L, H, S = 100_000, 10_000, 512, 1536
    bases, offs = synth.dataset(n, L, seed=2)
    p = native.SketchParams(16, H, 12, S, 0, 116)
    e = engine()
    e.store_reset(p)
    assert e.store_add_reads(bases, offs) == 2 * n
    t = e.timing()
    assert t["xorshift_steps"] == 2 * n * (L - 15) * H
    rng = np.random.default_rng(0)
    idx = sorted(rng.choice(n, size=6, replace=False).tolist())
    got = [e.store_get(2 * i + s, H, S) for i in idx for s in (0, 1)]
    for k, i in enumerate(idx):
        r = bytes(bases[int(offs[i]):int(offs[i + 1])])
        for s, seq in enumerate((r, orc.rc(r))):
            g = got[2 * k + s]
            assert g["id"] == i + 1 and g["is_fwd"] == (s == 0) and g["seq_len"] == L
            assert (g["minhash"] == orc.minhash_sketch(seq, 16, H)).all()
            assert (g["ord"] == orc.bottom_sketch(seq, 12, S)[0]).all()
    hits, stats = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, 0, 0, 0, -1))
    assert stats["sequences_searched"] == n and stats["matches_processed"] == len(hits) > 10_000
    assert stats["fully_compared"] >= stats["matches_processed"] and stats["sequences_hit"] >= stats["fully_compared"]
    # every reported overlap obeys the self-mode id rule and has sane coordinates
    assert (hits["to_id"] < hits["from_id"]).all() and (hits["from_fwd"] == 1).all()
    assert (hits["a1"] >= 0).all() and (hits["a2"] <= L - 11).all() and (hits["a1"] <= hits["a2"]).all()
    assert (hits["score"] >= 0.78).all() and (hits["hit_count"] >= 3).all()
    # each query's hit count equals MinHashSketch.jaccard's numerator computed pairwise (a10)
    for h in hits[:: max(1, len(hits) // 40)][:40]:
        qi = 2 * (int(h["from_id"]) - 1)
        ti = 2 * (int(h["to_id"]) - 1) + (0 if h["to_fwd"] else 1)
        assert e.minhash_equal_count(qi, ti) == h["hit_count"]
    # a sample of the hits against the oracle's second stage on the very same sketches
    for h in hits[:: max(1, len(hits) // 25)][:25]:
        qi = 2 * (int(h["from_id"]) - 1)
        ti = 2 * (int(h["to_id"]) - 1) + (0 if h["to_fwd"] else 1)
        a, b = e.store_get(qi, H, S), e.store_get(ti, H, S)
        o = orc.overlap_info(a["ord"], a["seq_len_kmers"], b["ord"], b["seq_len_kmers"], 12, 0.2)
        assert (o.a1, o.a2, o.b1, o.b2, o.valid_count, o.intersect, o.kmin) == tuple(int(h[k]) for k in ("a1", "a2", "b1", "b2", "valid_count", "intersect", "kmin"))
        assert abs(o.score - h["score"]) < 1e-15


def _full_parity(n, L, H, S, seed, keep_all):
    """Whole BASELINE config against the oracle: every stored sketch (min-hashes + ordered sketch), the complete hit set
    and the five counters of MhapMain.outputFinalStat (main/MhapMain.java:572-590).  The oracle runs one thread per host
    core (its pool mirrors Executors.newFixedThreadPool, AbstractMatchSearch.java:70,124)."""
    threads = os.cpu_count() or 1
    bases, offs = synth.dataset(n, L, seed=seed)
    p = native.SketchParams(16, H, 12, S, 0, 116)
    e = engine()
    e.store_reset(p)
    assert e.store_add_reads(bases, offs) == 2 * n
    hits, stats = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, int(keep_all), 0, 0, -1))
    st = orc.Store(num_hashes=H, ordered_size=S)
    assert st.add_reads(bases, offs, threads=threads) == 2 * n
    # every sketch, in blocks of 4096 rows
    for first in range(0, 2 * n, 4096):
        cnt = min(4096, 2 * n - first)
        g = e.store_get_range(first, cnt)
        for j in range(cnt):
            o = st.get(first + j)
            assert o["id"] == g["ids"][j] and o["is_fwd"] == bool(g["is_fwd"][j]) and o["seq_len"] == g["seq_len"][j]
            assert (o["minhash"] == g["minhash"][j]).all(), first + j
            assert o["ord"].shape[0] == g["ord_n"][j] and (o["ord"] == g["ord"][j, :g["ord_n"][j]]).all(), first + j
    res = st.search_self(threads=threads, keep_all=keep_all)
    st.close()
    assert_same_hits_bulk(hits, res.hits, stats, res.stats)
    return stats


def test_config2_full_parity():
    # BASELINE configs[1] at its own size: 100k reads x 10 kbp, H=512, S=1536 -- the numbers bench.py reports
    stats = _full_parity(100_000, 10_000, 512, 1536, seed=2, keep_all=False)
    assert stats["sequences_searched"] == 100_000 and stats["matches_processed"] > 10_000


def test_config3_full_parity():
    # BASELINE configs[2] at its own size: 50k reads x 8 kbp, --ordered-sketch-size 1000; keep_all: also the candidates that
    # were fully compared and rejected by the threshold
    stats = _full_parity(50_000, 8_000, 512, 1000, seed=3, keep_all=True)
    assert stats["sequences_searched"] == 50_000 and stats["fully_compared"] > stats["matches_processed"] > 0


def test_config3_shape_ordered_sketch_1000():
    n, L = 50_000, 8_000
    bases, offs = synth.dataset(n, L, seed=3, count=1500)
    p = native.SketchParams(16, 512, 12, 1000, 0, 116)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    hits, stats = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1))
    st = orc.Store(num_hashes=512, ordered_size=1000)
    st.add_reads(bases, offs, threads=16)
    res = st.search_self(threads=16, keep_all=True)
    assert_same_hits(hits, res.hits, stats, res.stats)


def test_config4_shape_store_vs_query():
    g = synth.genome(4, 600_000)
    sb, so = synth.reads(g, 4, 0, 700, 12_000, 0.15)
    qb, qo = synth.reads(g, 5, 0, 500, 12_000, 0.15)
    p = native.SketchParams(16, 512, 12, 1536, 0, 116)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(sb, so)
    qids = np.arange(1, 501, dtype=np.int64) + 700
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    hs, ss = e.search_self(sp)
    hq, sq = e.search_query_reads(sp, qb, qo, qids)
    st = orc.Store(num_hashes=512)
    st.add_reads(sb, so, threads=16)
    qs = orc.Store(num_hashes=512)
    qs.add_reads(qb, qo, ids=qids, both_strands=False, threads=16)
    rs = st.search_self(threads=16, keep_all=True)
    rq = st.search_query(qs, threads=16, keep_all=True)
    assert_same_hits(hs, rs.hits, ss, rs.stats)
    assert_same_hits(hq, rq.hits, sq, rq.stats)
    assert sq["fully_compared"] > 0


def test_config5_shape_h1024_15kbp():
    bases, offs = synth.dataset(1_000_000, 15_000, seed=6, count=400)
    p = native.SketchParams(16, 1024, 12, 1536, 0, 116)
    mh, od, on, st = engine().sketch(bases, offs, p)
    assert not st.any()
    _sample_parity(bases, offs, 15_000, [0, 7, 399], mh, od, on, 1024, 1536)
    # strand symmetry on the whole batch: sketch(rc(read)).fwd == sketch(read).rc
    rc = np.concatenate([np.frombuffer(orc.rc(bytes(bases[int(offs[i]):int(offs[i + 1])])), np.uint8) for i in range(400)])
    mh2, od2, on2, _ = engine().sketch(rc, offs, p)
    assert (mh2[0::2] == mh[1::2]).all() and (mh2[1::2] == mh[0::2]).all() and (od2[0::2] == od[1::2]).all()
