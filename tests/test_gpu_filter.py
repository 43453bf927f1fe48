"""-m gpu: the -f k-mer filter on the device (K1a weight rule + Bloom/repeat-map lookups, K1b uniform and per-key weights)
against the oracle's FrequencyCounts restatement, through the C ABI."""
import random

import numpy as np
import pytest

from mhap_b200 import native, synth
from oracle import oracle as orc
from tests.filter_common import SETTINGS, make_reads_and_filter
from tests.gpu_common import assert_same_hits, engine

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _clear_filter():
    yield
    engine().filter_clear()


def _expected(reads, f, rw, k, H, ok, S, both, min_olap=0):
    """per read: status and the oracle sketches of the strands the reference would keep"""
    out = []
    for r in reads:
        up = r.upper().encode("latin-1")
        if len(up) < min_olap:
            out.append((2, None, None)); continue
        if len(up) - k + 1 < 1 or len(up) - ok + 1 < 1:
            out.append((1, None, None)); continue
        fwd = orc.minhash_sketch_filtered(up, k, H, rw, f)
        if fwd is None:
            out.append((1, None, None)); continue
        rev = orc.minhash_sketch_filtered(orc.rc(up), k, H, rw, f) if both else None
        out.append((3 if (both and rev is None) else 0, fwd, rev))
    return out


@pytest.mark.parametrize("rw,sn,notf", SETTINGS)
def test_sketch_parity_with_filter(rw, sn, notf):
    reads, text = make_reads_and_filter(n_reads=10, read_len=300)
    reads += [reads[0][:40], "ACGT" * 30, reads[1].lower()]
    e = engine()
    n_rep = e.filter_load_text(text, repeat_weight=rw, supress_noise=sn, no_tf=notf)
    f = orc.KmerFilter(text, repeat_weight=rw, supress_noise=sn, no_tf=notf)
    assert n_rep == len(f)
    H, S = 96, 64
    p = native.SketchParams(16, H, 12, S, int(rw < 0), 0)
    mh, od, on, st = e.sketch(*native.pack_reads(reads), p, both_strands=True)
    exp = _expected(reads, f, rw, 16, H, 12, S, True)
    assert st.tolist() == [x[0] for x in exp]
    for i, (status, fwd, rev) in enumerate(exp):
        if fwd is not None:
            assert (mh[2 * i] == fwd).all(), (i, "fwd")
            assert (od[2 * i, :on[2 * i]] == orc.bottom_sketch(reads[i].upper(), 12, S)[0]).all()
        if rev is not None:
            assert (mh[2 * i + 1] == rev).all(), (i, "rc")


@pytest.mark.parametrize("rw,sn", [(0.9, 0), (0.9, 2), (-1.0, 0), (0.4, 1)])
def test_pacbio_shape_reads_with_filter_take_the_bit_sliced_path(rw, sn):
    # 10 kbp reads: light keys (weight round(range) = 3 under tf-idf) run through the bit-sliced lock-step kernel with
    # three compared steps per word; repeat k-mers and duplicated k-mers carry their own weights on the scalar path
    bases, offs = synth.dataset(100000, 10000, seed=2, count=4)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]).decode() for i in range(4)]
    rng = random.Random(4)
    kmers = sorted({r[i:i + 16] for r in reads for i in range(0, 9000, 3)})
    lines = [f"{len(kmers)} {len(kmers)}"] + [f"{km} {rng.choice([2e-6, 2e-5, 1e-4, 1e-3, 0.01])}" for km in kmers]
    text = "\n".join(lines) + "\n"
    e = engine()
    e.filter_load_text(text, repeat_weight=rw, supress_noise=sn)
    f = orc.KmerFilter(text, repeat_weight=rw, supress_noise=sn)
    H = 512
    p = native.SketchParams(16, H, 12, 1536, int(rw < 0), 116)
    mh, _, _, st = e.sketch(*native.pack_reads(reads), p, both_strands=True, want_ord=False)
    assert st.tolist() == [0] * 4
    for i, r in enumerate(reads):
        assert (mh[2 * i] == orc.minhash_sketch_filtered(r, 16, H, rw, f)).all(), i
        assert (mh[2 * i + 1] == orc.minhash_sketch_filtered(orc.rc(r), 16, H, rw, f)).all(), i


@pytest.mark.parametrize("rw,sn", [(0.9, 0), (0.9, 1), (-1.0, 0)])
def test_store_and_self_search_with_filter(rw, sn):
    bases, offs = synth.dataset(150, 1500, seed=7, err=0.06)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]).decode() for i in range(150)]
    rng = random.Random(8)
    # supress-noise 1 keeps only listed k-mers: list most of them so that reads survive, but leave a few reads uncovered
    kmers = sorted({r[i:i + 16] for r in reads[:140] for i in range(0, 1485, 2 if sn == 1 else 5)})
    text = "\n".join([f"{len(kmers)} {len(kmers)}"] + [f"{km} {rng.choice([2e-6, 5e-5, 1e-3])}" for km in kmers]) + "\n"
    e = engine()
    e.filter_load_text(text, repeat_weight=rw, supress_noise=sn)
    f = orc.KmerFilter(text, repeat_weight=rw, supress_noise=sn)
    p = native.SketchParams(16, 128, 12, 400, int(rw < 0), 116)
    e.store_reset(p)
    n_added = e.store_add_reads(bases, offs)
    ost = orc.Store(num_hashes=128, ordered_size=400, unweighted=rw < 0)
    ost.set_filter(f, rw)
    assert ost.add_reads(bases, offs, threads=4) == n_added == len(ost)
    for j in (0, 1, n_added // 2, n_added - 1):
        g, o = e.store_get(j, 128, 400), ost.get(j)
        assert g["id"] == o["id"] and g["is_fwd"] == o["is_fwd"] and (g["minhash"] == o["minhash"]).all()
    hits, stats = e.search_self(native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1))
    res = ost.search_self(threads=4, keep_all=True)
    assert_same_hits(hits, res.hits, stats, res.stats)
    assert len(hits) > 0


def test_filter_set_from_arrays_equals_load_text_and_clear_restores():
    reads, text = make_reads_and_filter(n_reads=6, read_len=300)
    e = engine()
    p = native.SketchParams(16, 64, 12, 64, 0, 0)
    plain = e.sketch(*native.pack_reads(reads), p)[0].copy()
    f = orc.KmerFilter(text, supress_noise=2)
    e.filter_load_text(text, supress_noise=2)
    a = e.sketch(*native.pack_reads(reads), p)[0].copy()
    h, fr = f.export()
    words, bits, nfun = f.bloom()
    e.filter_set(h, fr, supress_noise=2, bloom_words=words, bloom_bits=bits, bloom_nfun=nfun)
    b = e.sketch(*native.pack_reads(reads), p)[0].copy()
    assert (a == b).all() and (a != plain).any()
    e.filter_clear()
    assert (e.sketch(*native.pack_reads(reads), p)[0] == plain).all()


def test_filter_argument_errors():
    e = engine()
    with pytest.raises(native.MhapError):
        e.filter_load_text("1 1\nACGTACGTACGTACGT 0.5\n", supress_noise=3)
    with pytest.raises(native.MhapError):
        e.filter_load_text("not numbers\nACGTACGTACGTACGT 0.5\n")
    with pytest.raises(native.MhapError):
        e.filter_set(np.zeros(1, np.int64), np.ones(1), supress_noise=1)          # Bloom bits missing
    e.filter_load_text("1 1\nACGTACGTACGTACGT 0.5\n", repeat_weight=-1.0)
    with pytest.raises(native.MhapError):                                            # unweighted flag must agree
        e.sketch(*native.pack_reads(["ACGT" * 50]), native.SketchParams(16, 8, 12, 8, 0, 0))


def test_long_read_h1024_with_filter():
    # strands above 16384 k-mers take K1a's global-table path, H = 1024 runs as two 512-word blocks (virtual strands, MULTI); both with tf-idf
    # weights (light weight 3) and a few repeat k-mers carrying their own weights
    rng = random.Random(12)
    g = "".join(rng.choice("ACGT") for _ in range(21000))
    reads = [g[:20000], g[500:18000] + g[500:3000]]           # the second read repeats 2.5 kbp of itself (tf > 1)
    kmers = sorted({r[i:i + 16] for r in reads for i in range(0, 15000, 5)})
    text = "\n".join([f"{len(kmers)} {len(kmers)}"] + [f"{km} {rng.choice([2e-6, 2e-5, 1e-3])}" for km in kmers]) + "\n"
    e = engine()
    e.filter_load_text(text)
    f = orc.KmerFilter(text)
    H = 1024
    p = native.SketchParams(16, H, 12, 1536, 0, 116)
    mh, _, _, st = e.sketch(*native.pack_reads(reads), p, both_strands=True, want_ord=False)
    assert st.tolist() == [0, 0]
    for i, r in enumerate(reads):
        assert (mh[2 * i] == orc.minhash_sketch_filtered(r, 16, H, 0.9, f)).all(), i
        assert (mh[2 * i + 1] == orc.minhash_sketch_filtered(orc.rc(r), 16, H, 0.9, f)).all(), i
