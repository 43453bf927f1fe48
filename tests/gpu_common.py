"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import random

import numpy as np

from mhap_b200 import native
from oracle import oracle as orc

_engine = None


def engine():
    global _engine
    if _engine is None:
        _engine = native.Engine(0)
    return _engine


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def nasty_seq(rng, n):
    s = []
    while len(s) < n:
        r = rng.random()
        if r < 0.12:
            s.extend(rand_seq(rng, rng.randint(1, 6)) * rng.randint(3, 25))
        elif r < 0.2:
            s.extend(rng.choice("ACGT") * rng.randint(5, 60))
        elif r < 0.25:
            s.extend(rand_seq(rng, rng.randint(1, 5), "NRYKMSWBDHVnacgtx*-"))
        else:
            s.extend(rand_seq(rng, rng.randint(5, 80)))
    return "".join(s[:n])


def oracle_sketches(reads, k, H, ok, S, unweighted, both, min_olap=116):
    """Per slot (read*per+strand): (minhash|None, ord|None, status)."""
    out, status = [], []
    per = 2 if both else 1
    for r in reads:
        b = r.encode("latin-1") if isinstance(r, str) else bytes(r)
        up = b.upper()
        if len(up) < min_olap:
            status.append(2); out.extend([None] * per); continue
        if len(up) - k + 1 < 1 or len(up) - ok + 1 < 1:
            status.append(1); out.extend([None] * per); continue
        status.append(0)
        for s in range(per):
            seq = up if s == 0 else orc.rc(up)
            mh = orc.minhash_sketch(seq, k, H, unweighted)
            od, _ = orc.bottom_sketch(seq, ok, S)
            out.append((mh, od))
    return out, status


def check_sketch_parity(reads, k=16, H=512, ok=12, S=1536, unweighted=False, both=True, min_olap=116):
    p = native.SketchParams(k, H, ok, S, int(unweighted), min_olap)
    bases, offs = native.pack_reads(reads)
    mh, od, on, st = engine().sketch(bases, offs, p, both_strands=both)
    exp, est = oracle_sketches(reads, k, H, ok, S, unweighted, both, min_olap)
    assert st.tolist() == est
    for j, e in enumerate(exp):
        if e is None:
            assert not mh[j].any() and on[j] == 0 and not od[j].any(), j
            continue
        emh, eod = e
        assert (mh[j] == emh).all(), (j, np.nonzero(mh[j] != emh)[0][:8], len(reads[j // (2 if both else 1)]))
        assert on[j] == eod.shape[0], (j, on[j], eod.shape)
        assert (od[j, :on[j]] == eod).all(), (j, np.nonzero((od[j, :on[j]] != eod).any(1))[0][:8])
        assert not od[j, on[j]:].any()
    return mh, od, on, st


def hit_key(h):
    return (int(h["from_id"]), int(h["to_id"]), int(h["from_fwd"]), int(h["to_fwd"]), int(h["hit_count"]), int(h["a1"]), int(h["a2"]),
            int(h["b1"]), int(h["b2"]), int(h["valid_count"]), int(h["intersect"]), int(h["kmin"]), int(h["from_len"]), int(h["to_len"]),
            int(h["accepted"]))


def assert_same_hits(got, exp, stats_got, stats_exp):
    assert stats_got == stats_exp, (stats_got, stats_exp)
    g = sorted(hit_key(h) for h in got)
    e = sorted(hit_key(h) for h in exp)
    assert len(g) == len(e), (len(g), len(e))
    assert g == e
    sg = {hit_key(h): float(h["score"]) for h in got}
    for h in exp:
        assert abs(sg[hit_key(h)] - float(h["score"])) <= 1e-15
    # printed lines (MatchResult.toString) as sorted sets
    lg = sorted(native.format_match(h) for h in got if h["accepted"])
    le = sorted(orc.format_match(h) for h in exp if h["accepted"])
    assert lg == le


_KEY_FIELDS = ("from_id", "to_id", "from_fwd", "to_fwd", "hit_count", "a1", "a2", "b1", "b2", "valid_count", "intersect", "kmin",
               "from_len", "to_len", "accepted")


def sorted_hits(h):
    """Hits in the canonical order of their integer key (vectorised: full-size hit sets)."""
    order = np.lexsort(tuple(h[f] for f in reversed(_KEY_FIELDS)))
    return h[order]


def assert_same_hits_bulk(got, exp, stats_got, stats_exp, lines=2000):
    """assert_same_hits for hit sets of 10^4..10^6 rows: every integer field bit-exact on the sorted sets, scores within
    1e-15, the five counters equal, and the printed MatchResult lines of an evenly spaced sample."""
    assert stats_got == stats_exp, (stats_got, stats_exp)
    assert len(got) == len(exp), (len(got), len(exp))
    g, e = sorted_hits(got), sorted_hits(exp)
    for f in _KEY_FIELDS:
        bad = np.nonzero(g[f] != e[f])[0]
        assert bad.size == 0, (f, bad[:5], g[bad[:3]], e[bad[:3]])
    assert np.all(np.abs(g["score"] - e["score"]) <= 1e-15)
    step = max(1, len(g) // max(1, lines))
    for a, b in zip(g[::step], e[::step]):
        if a["accepted"]:
            assert native.format_match(a) == orc.format_match(b)


def hits_digest(h) -> str:
    """Order-independent digest of a hit set: sha256 over the sorted integer keys (bench.py prints the same)."""
    import hashlib
    s = sorted_hits(h)
    m = hashlib.sha256()
    for f in _KEY_FIELDS:
        m.update(np.ascontiguousarray(s[f]).astype("<i8").tobytes())
    return m.hexdigest()[:16]
