"""CPU suite, part 2: the C-ABI library loads without a GPU and exports every symbol the header declares;
its host-only entry points (.dat codec, MatchResult formatting) agree with the oracle.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mhap_b200 import native
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mhap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mhapb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = native.load()
    names = _declared()
    assert len(names) >= 20
    assert sorted(native.EXPORTS) == names
    for n in names:
        assert getattr(L, n) is not None, n
    assert b"mhap-b200" in L.mhapb_version()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.MhapError) as ei:
        native.Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def _rand_sketch(rng, H, n):
    mh = rng.integers(-2**31, 2**31, size=H).astype(np.int32)
    od = np.stack([np.sort(rng.integers(-2**31, 2**31, size=n)).astype(np.int32), rng.integers(0, 10000, size=n).astype(np.int32)], 1)
    return mh, od


def test_dat_encode_matches_oracle_and_round_trips():
    rng = np.random.default_rng(5)
    blob, recs = b"", []
    for i in range(7):
        H, n = 16, int(rng.integers(0, 40))
        mh, od = _rand_sketch(rng, H, n)
        id_, fwd, sl = int(rng.integers(1, 2**40)), bool(i & 1), int(rng.integers(100, 20000))
        a = native.dat_encode(id_, fwd, sl, mh, sl - 11, 12, od)
        assert a == orc.dat_encode(id_, fwd, sl, mh, sl - 11, 12, od)
        blob += a
        recs.append((id_, fwd, sl, mh, od))
    d = native.dat_decode(blob, id_offset=1000)
    assert d["num_hashes"] == 16 and d["ordered_kmer_size"] == 12 and len(d["ids"]) == 7
    for i, (id_, fwd, sl, mh, od) in enumerate(recs):
        assert d["ids"][i] == id_ + 1000 and bool(d["is_fwd"][i]) == fwd and d["seq_len"][i] == sl
        assert d["seq_len_kmers"][i] == sl - 11 and d["ord_n"][i] == od.shape[0]
        assert (d["minhash"][i] == mh).all() and (d["ord"][i, :od.shape[0]] == od).all()
    with pytest.raises(native.MhapError):
        native.dat_decode(blob[:-3])           # truncated stream
    assert len(native.dat_decode(b"")["ids"]) == 0
    assert native.dat_encode(5, True, 10, np.zeros(2, np.int32), 1, 12, np.zeros((0, 2), np.int32), header="read/5")[16:22] == b"read/5"


def test_format_match_matches_oracle():
    rng = np.random.default_rng(9)
    for _ in range(50):
        h = np.zeros(1, dtype=native.HIT_DTYPE)[0]
        o = np.zeros(1, dtype=orc.HIT_DTYPE)[0]
        vals = dict(from_id=int(rng.integers(1, 10**6)), to_id=int(rng.integers(1, 10**6)), from_fwd=int(rng.integers(0, 2)),
                    to_fwd=int(rng.integers(0, 2)), a1=int(rng.integers(0, 500)), a2=int(rng.integers(500, 9000)),
                    b1=int(rng.integers(0, 500)), b2=int(rng.integers(500, 9000)), from_len=10000, to_len=9500,
                    valid_count=int(rng.integers(3, 400)), score=float(rng.uniform(0.7, 1.05)))
        for k, v in vals.items():
            h[k] = v; o[k] = v
        assert native.format_match(h) == orc.format_match(o)
