"""-m gpu: K1 parity -- the CUDA sketch path (through the C ABI) against the CPU oracle, bit-exact."""
import random

import numpy as np
import pytest

from mhap_b200 import native, synth
from tests.gpu_common import check_sketch_parity, engine, nasty_seq, rand_seq
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_survey_vectors_on_gpu():
    s = "ACGTACGTTGCAAGGCTTAACGGTACCATGCATGCAAATTACGTACGTTGCAAGGCTTAA"
    p = native.SketchParams(16, 8, 12, 5, 0, 0)
    mh, od, on, st = engine().sketch(*native.pack_reads([s]), p, both_strands=False)
    assert mh[0].tolist() == [134182658, 1650722871, 219875990, -522100345, -29195125, 425860115, 1362635161, -156129382]
    p.unweighted = 1
    mh, _, _, _ = engine().sketch(*native.pack_reads([s]), p, both_strands=False)
    assert mh[0].tolist() == [134182658, 1650722871, -1881521518, -522100345, 754473341, 425860115, 1362635161, -156129382]
    s2 = "ACGTACGTTGCAAGGCTTAACGGTACCATGCATGCAAATTTCCCGGGACGTACGTTGCAAGG"
    _, od, on, _ = engine().sketch(*native.pack_reads([s2]), p, both_strands=False)
    assert on[0] == 5
    assert od[0, :5].tolist() == [[-2119918638, 38], [-2091752154, 7], [-2061838574, 12], [-2004994854, 1], [-2004994854, 48]]


def test_config1_shape_reads():
    # BASELINE configs[0] shape (1 kbp reads, k=16, H=256), a 64-read sample the oracle finishes in seconds
    bases, offs = synth.dataset(1000, 1000, seed=1, count=64)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(64)]
    check_sketch_parity(reads, k=16, H=256, S=1536)


@pytest.mark.parametrize("H,S", [(512, 1536), (1024, 1536), (768, 1000)])
def test_pacbio_shape_reads(H, S):
    bases, offs = synth.dataset(100000, 10000, seed=2, count=6)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(6)]
    check_sketch_parity(reads, H=H, S=S)


@pytest.mark.parametrize("seed", range(4))
def test_ragged_nasty_reads(seed):
    rng = random.Random(seed)
    reads = []
    for _ in range(40):
        n = rng.choice([0, 1, 11, 12, 15, 16, 17, 115, 116, 117, 200, 500, 1537, 1547, 1548, 3000])
        reads.append(nasty_seq(rng, n) if rng.random() < 0.7 else rand_seq(rng, n))
    check_sketch_parity(reads, H=rng.choice([64, 100, 256]), S=rng.choice([1536, 100]), min_olap=rng.choice([0, 116]))
    check_sketch_parity(reads, H=64, S=50, unweighted=True, both=False, min_olap=0)


@pytest.mark.parametrize("k,ok,H", [(12, 5, 32), (21, 14, 40), (7, 13, 1), (32, 12, 33), (17, 16, 96), (8, 9, 2048)])
def test_other_kmer_sizes_and_hash_counts(k, ok, H):
    rng = random.Random(k * 100 + ok)
    reads = [nasty_seq(rng, rng.randint(10, 900)) for _ in range(12)]
    check_sketch_parity(reads, k=k, H=H, ok=ok, S=64, min_olap=0)


def test_every_tiny_length():
    # every strand length from below k up to a few hundred: the dedup table is fitted per strand, so its capacity
    # crosses every small value (a table smaller than the largest probe stride once indexed out of bounds)
    rng = random.Random(5)
    reads = [rand_seq(rng, n) for n in range(10, 330)]
    check_sketch_parity(reads, H=32, S=64, min_olap=0, both=False)
    check_sketch_parity(reads, H=32, S=64, min_olap=0, both=False, unweighted=True)


def test_homopolymer_and_tandem_repeat_reads():
    # weights in the hundreds (tf weighting, MinHashSketch.java:98-125) and ordered hashes that all tie
    reads = ["A" * 400, "AC" * 300, "ACG" * 250 + "T" * 50, "ACGTTGCA" * 100 + "N" * 30 + "ACGTTGCA" * 20]
    check_sketch_parity(reads, H=64, S=100)
    check_sketch_parity(reads, H=64, S=1536, unweighted=True)


def test_long_reads_take_the_global_table_path():
    rng = random.Random(11)
    g = rand_seq(rng, 60000)
    reads = [g[:40000], g[10000:30000] + g[10000:30000], nasty_seq(rng, 17000), g[:16399], g[:16400], g[:16500]]
    check_sketch_parity(reads, H=32, S=1536)


def test_empty_batch_and_all_skipped():
    p = native.SketchParams(16, 64, 12, 100, 0, 116)
    mh, od, on, st = engine().sketch(np.zeros(0, np.uint8), np.zeros(1, np.uint64), p)
    assert mh.shape == (0, 64) and st.size == 0
    mh, od, on, st = engine().sketch(*native.pack_reads(["ACGT", ""]), p)
    assert st.tolist() == [2, 2] and not mh.any() and not on.any()


def test_bad_parameters_are_rejected():
    e = engine()
    for p in (native.SketchParams(0, 64, 12, 100, 0, 0), native.SketchParams(16, 0, 12, 100, 0, 0),
              native.SketchParams(16, 4096, 12, 100, 0, 0), native.SketchParams(16, 64, 12, 0, 0, 0)):
        with pytest.raises(native.MhapError) as ei:
            e.sketch(*native.pack_reads(["ACGT" * 50]), p)
        assert ei.value.code == -1


def test_sketch_rc_rc_is_identity_property():
    # size-independent property: the rc strand of rc(read) equals the forward strand of read
    bases, offs = synth.dataset(2000, 8000, seed=3, count=8)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(8)]
    p = native.SketchParams(16, 512, 12, 1000, 0, 116)
    a = engine().sketch(*native.pack_reads(reads), p)
    b = engine().sketch(*native.pack_reads([orc.rc(r) for r in reads]), p)
    assert (a[0][0::2] == b[0][1::2]).all() and (a[0][1::2] == b[0][0::2]).all()
    assert (a[1][0::2] == b[1][1::2]).all() and (a[1][1::2] == b[1][0::2]).all()


def test_dat_records_from_gpu_match_oracle_encoding():
    rng = random.Random(5)
    reads = [nasty_seq(rng, rng.randint(100, 700)) for _ in range(9)]
    p = native.SketchParams(16, 64, 12, 80, 0, 116)
    ids = np.arange(1, 10, dtype=np.int64) + 41
    blob, nrec = engine().sketch_to_dat(*native.pack_reads(reads), ids, p)
    exp = b""
    n_ok = 0
    for i, r in enumerate(reads):
        up = r.upper().encode("latin-1")
        if len(up) < 116:
            continue
        for fwd in (True, False):
            s = up if fwd else orc.rc(up)
            od, slk = orc.bottom_sketch(s, 12, 80)
            exp += orc.dat_encode(int(ids[i]), fwd, len(s), orc.minhash_sketch(s, 16, 64), slk, 12, od)
            n_ok += 1
    assert nrec == n_ok and blob == exp
    d = native.dat_decode(blob)
    assert d["ids"].tolist() == [int(ids[i]) for i, r in enumerate(reads) if len(r) >= 116 for _ in (0, 1)]
