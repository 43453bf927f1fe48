"""CPU suite: the -f k-mer filter (sketch/FrequencyCounts.java + the weight rule of sketch/MinHashSketch.java:95-130).

PARITY UNPINNED by the reference (no tests, no JVM): what pins the restatement is two independently written versions
(oracle/mhap_oracle.c and oracle/pyref.py) agreeing over every option combination, a committed golden fixture derived from
them, and the library's host-side k-mer hash agreeing with both."""
import json
import os

import numpy as np
import pytest

from mhap_b200 import native
from oracle import oracle as orc, pyref
from tests.filter_common import SETTINGS, make_reads_and_filter

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_vectors.json")


@pytest.mark.parametrize("rw,sn,notf", SETTINGS)
def test_c_and_python_restatements_agree(rw, sn, notf):
    reads, text = make_reads_and_filter(extra_lines=("acgtacgtacgtacgt 0.5", "ACGTNNNNACGTACGT notanumber", "TTTTTTTTTTTTTTTT"))
    f = orc.KmerFilter(text, repeat_weight=rw, supress_noise=sn, no_tf=notf)
    off = rw if 0.0 <= rw < 1.0 else 0.0
    pf = pyref.FrequencyCounts(text, 1.0e-5, off, sn, notf, 3.0, True)
    assert len(f) == len(pf.fraction)
    h, fr = f.export()
    assert dict(zip(h.tolist(), fr.tolist())) == pf.fraction
    for x in h.tolist()[:40]:
        assert f.scaled_idf(x) == pytest.approx(pf.scaled_idf(x), rel=0, abs=1e-15)
    for r in reads:
        a = orc.minhash_sketch_filtered(r, 16, 24, rw, f)
        b = pyref.minhash_sketch_filtered(r, 16, 24, rw, pf)
        assert (a is None) == (b is None)
        if a is not None:
            assert a.tolist() == b


def test_filter_changes_the_sketch_and_weights_triple():
    # tf-idf default: a once-seen k-mer outside the repeat map weighs round(1 * 3.0) = 3, repeats weigh 1..3
    reads, text = make_reads_and_filter()
    f = orc.KmerFilter(text)
    hs = orc.kmer_hashes_long(reads[0], 16)
    idf = [f.scaled_idf(int(x)) for x in hs]
    assert max(idf) == 3.0 and 1.0 <= min(idf) < 3.0
    assert orc.minhash_sketch_filtered(reads[0], 16, 64, 0.9, f).tolist() != orc.minhash_sketch(reads[0], 16, 64).tolist()
    # repeat-weight >= 1 ignores the idf: same as no filter
    f2 = orc.KmerFilter(text, repeat_weight=1.0)
    assert orc.minhash_sketch_filtered(reads[0], 16, 64, 1.0, f2).tolist() == orc.minhash_sketch(reads[0], 16, 64).tolist()


def test_supress_noise_1_can_empty_a_read():
    reads, text = make_reads_and_filter()
    f = orc.KmerFilter(text, supress_noise=1)
    words, bits, nfun = f.bloom()
    assert bits % 64 == 0 and nfun >= 1 and words.any()
    none = [orc.minhash_sketch_filtered(r, 16, 8, 0.9, f) is None for r in reads]
    assert any(none) and not all(none)


def test_filter_canonicalises_file_kmers_but_not_reads():
    # main/MhapMain.java:359 passes doReverseCompliment to FrequencyCounts; SequenceSketch.java:112 passes false for reads
    km = "TTTTGGGGCCCCAAAC"
    rc = orc.rc(km).decode()
    assert rc < km
    f = orc.KmerFilter(f"1 1\n{km} 0.5\n")
    assert f.is_popular(int(orc.kmer_hashes_long(rc, 16)[0])) and not f.is_popular(int(orc.kmer_hashes_long(km, 16)[0]))


def test_host_kmer_hash_of_the_library_matches_the_oracle():
    for s in ("ACGTACGTACGTACGT", "acgtnnRYacgtACGT", "T", "TTTTGGGGCCCCAAAC", "GATTACA" * 5):
        for canon in (False, True):
            assert native.kmer_hash(s, canon) == int(orc.kmer_hashes_long(s, len(s), 0, canon)[0])
    assert native.kmer_hash("ACGTACGTACGTACGT", False) == 0x77cc6caa6c67a9a4   # SURVEY 8c vector


def test_golden_filter_vectors():
    g = json.load(open(GOLD))
    for case in g["cases"]:
        f = orc.KmerFilter(g["filter_text"], repeat_weight=case["repeat_weight"], supress_noise=case["supress_noise"],
                           no_tf=case["no_tf"])
        for r, exp in zip(g["reads"], case["minhash"]):
            got = orc.minhash_sketch_filtered(r, 16, g["num_hashes"], case["repeat_weight"], f)
            assert (None if got is None else got.tolist()) == exp
