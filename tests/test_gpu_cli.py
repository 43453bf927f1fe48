"""-m gpu: the host driver (mhap-b200, C++ over the C ABI) against the oracle -- the reference's command
lines for the hot path: -s self overlap, -s/-q store vs query files, -p FASTA -> .dat, and .dat inputs.
Output lines are compared as sorted sets (the reference's order is thread-timing dependent)."""
import os
import random
import subprocess

import numpy as np
import pytest

from mhap_b200 import native, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "mhap_b200", "mhap-b200")


def _write_fasta(path, reads, width=70, lower_every=3):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            s = r.decode() if isinstance(r, bytes) else r
            if lower_every and i % lower_every == 0:
                s = s.lower()                         # FastaData.java:194 upper-cases
            f.write(f">read_{i} some description\n")
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + "\n")


def _reads(n, L, seed, err=0.06, genome_seed=77):
    g = synth.genome(genome_seed, 30000)
    b, o = synth.reads(g, seed, 0, n, L, err)
    out = [bytes(b[int(o[i]):int(o[i + 1])]) for i in range(n)]
    rng = random.Random(seed)
    for i in range(0, n, 9):
        out[i] = out[i][: rng.choice([40, 100, 115])]   # below --min-olap-length: id consumed, read skipped
    return out


def _run(args, env=None):
    p = subprocess.run([CLI] + args, capture_output=True, text=True, timeout=600, env=None if env is None else dict(os.environ, **env))
    assert p.returncode == 0, p.stderr[-2000:]
    return sorted(l for l in p.stdout.splitlines() if l.strip()), p.stderr


def _oracle_lines(hits):
    return sorted(orc.format_match(h) for h in hits)


def test_self_overlap_from_fasta(tmp_path):
    reads = _reads(150, 2000, 5)
    fa = tmp_path / "store.fasta"
    _write_fasta(fa, reads)
    got, err = _run(["-s", str(fa), "--num-hashes", "256", "--num-threads", "4"])
    st = orc.Store(num_hashes=256)
    st.add_reads(*orc.pack_reads([r.upper() for r in reads]), threads=8)
    res = st.search_self(threads=8)
    assert got == _oracle_lines(res.hits) and len(got) > 50
    assert f"Total matches found: {res.stats['matches_processed']}" in err
    assert f"Stored {len(st)} sequences in the index." in err


def test_store_vs_query_directory_and_no_self(tmp_path):
    store = _reads(90, 1500, 6)
    q1, q2 = _reads(40, 1500, 7), _reads(30, 1500, 8)
    fa = tmp_path / "s.fa"
    qd = tmp_path / "queries"
    qd.mkdir()
    _write_fasta(fa, store)
    _write_fasta(qd / "b_second.fa", q2)
    _write_fasta(qd / "a_first.fa", q1)
    (qd / ".hidden").write_text("ignored")
    args = ["-s", str(fa), "-q", str(qd), "--settings", "2", "--num-min-matches", "2", "--threshold", "0.7"]
    got, _ = _run(args)
    kw = dict(num_hashes=256, ordered_k=14, ordered_size=1000)
    st = orc.Store(**kw)
    st.add_reads(*orc.pack_reads(store), threads=8)
    sp = dict(num_min_matches=2, accept_score=0.7, threads=8)
    exp = list(st.search_self(**sp).hits)
    offset = len(st) // 2                                   # MhapMain.java:462
    for q in (q1, q2):                                      # alphabetical order (:512)
        qs = orc.Store(**kw)
        qs.add_reads(*orc.pack_reads(q), both_strands=False, id_offset=offset, threads=8)
        exp += list(st.search_query(qs, **sp).hits)
        offset += len(qs)                                   # :537
    assert got == _oracle_lines(exp) and len(got) > 30
    got_ns, _ = _run(args + ["--no-self"])
    assert got_ns == _oracle_lines(exp[len(st.search_self(**sp).hits):])


def test_dat_round_trip_equals_fasta_path(tmp_path):
    store, query = _reads(80, 1500, 9), _reads(50, 1500, 10)
    src = tmp_path / "fa"; out = tmp_path / "dat"
    src.mkdir(); out.mkdir()
    _write_fasta(src / "store.fasta", store)
    _write_fasta(src / "query.fasta", query)
    _, err = _run(["-p", str(src), "-q", str(out), "--num-hashes", "128"])
    assert "Processed" in err and (out / "store.dat").exists() and (out / "query.dat").exists()
    # .dat records = oracle encoding of the same sketches (record order here is file order)
    blob = (out / "store.dat").read_bytes()
    exp = b""
    for i, r in enumerate(store):
        if len(r) < 116:
            continue
        for fwd in (True, False):
            s = r if fwd else orc.rc(r)
            od, slk = orc.bottom_sketch(s, 12, 1536)
            exp += orc.dat_encode(i + 1, fwd, len(s), orc.minhash_sketch(s, 16, 128), slk, 12, od)
    assert blob == exp
    via_dat, _ = _run(["-s", str(out / "store.dat"), "-q", str(out / "query.dat"), "--num-hashes", "128"])
    # the reference prints the header string stored in the record: .dat queries print file-local ids
    st = orc.Store(num_hashes=128)
    st.add_reads(*orc.pack_reads(store), threads=8)
    qs = orc.Store(num_hashes=128)
    qs.add_reads(*orc.pack_reads(query), both_strands=False, threads=8)       # file-local ids
    exp_hits = list(st.search_self(threads=8).hits) + list(st.search_query(qs, threads=8).hits)
    assert via_dat == _oracle_lines(exp_hits) and len(via_dat) > 20


@pytest.mark.parametrize("extra,kw", [([], {}), (["--supress-noise", "2", "--repeat-idf-scale", "4"], dict(supress_noise=2, idf_scale=4.0)),
                                      (["--repeat-weight", "-1"], dict(repeat_weight=-1.0))])
def test_self_overlap_with_kmer_filter_file(tmp_path, extra, kw):
    # -f: main/MhapMain.java:340-372 + sketch/FrequencyCounts.java (what Canu runs: tf-idf weights from a repeat k-mer list)
    reads = _reads(120, 2000, 6)
    fa = tmp_path / "store.fasta"
    _write_fasta(fa, reads)
    rng = random.Random(2)
    kmers = sorted({r[i:i + 16].decode().upper() for r in reads[:60] for i in range(0, max(0, len(r) - 16), 11)})
    text = "\n".join([f"{len(kmers)} {len(kmers)}"] + [f"{km}\t{rng.choice([5e-6, 4e-5, 1e-3, 0.02])}" for km in kmers]) + "\n"
    if extra:                                    # compressed filter files go through the same reader (utils/Utils.java getFile)
        import gzip
        ff = tmp_path / "repeats.txt.gz"
        with gzip.open(ff, "wt") as g:
            g.write(text)
    else:
        ff = tmp_path / "repeats.txt"
        ff.write_text(text)
    got, err = _run(["-s", str(fa), "-f", str(ff), "--num-hashes", "256"] + extra)
    rw = kw.get("repeat_weight", 0.9)
    f = orc.KmerFilter(text, **kw)
    st = orc.Store(num_hashes=256, unweighted=rw < 0)
    st.set_filter(f, rw)
    st.add_reads(*orc.pack_reads([r.upper() for r in reads]), threads=8)
    res = st.search_self(threads=8)
    assert got == _oracle_lines(res.hits) and len(got) > 20
    assert f"Read in k-mer filter with {len(f)} repeat k-mers." in err


def test_streamed_batches_and_gzip_input_equal_the_single_batch_run(tmp_path):
    # the FASTA producer cuts the file into batches at record starts (64 KB here: ~6 batches per file); ids, the store, the
    # query offsets and the .dat files must not depend on where the cuts fall.  Also: .gz input through zlib.
    import gzip
    store, query = _reads(160, 2000, 11), _reads(90, 2000, 12)
    fa, qa = tmp_path / "store.fasta", tmp_path / "query.fasta"
    _write_fasta(fa, store)
    _write_fasta(qa, query)
    gz = tmp_path / "store_gz.fasta.gz"
    with open(fa, "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    args = ["--num-hashes", "256", "--num-threads", "3"]
    one, err1 = _run(["-s", str(fa), "-q", str(qa)] + args)
    many, err2 = _run(["-s", str(fa), "-q", str(qa)] + args, env={"MHAPB_FASTA_CHUNK_KB": "64"})
    zipped, _ = _run(["-s", str(gz), "-q", str(qa)] + args, env={"MHAPB_FASTA_CHUNK_KB": "64"})
    assert one == many == zipped and len(one) > 100
    pick = lambda e: [l for l in e.splitlines() if l.startswith(("Stored", "Processed", "Total matches"))]
    assert pick(err1) == pick(err2)
    d1, d2 = tmp_path / "d1", tmp_path / "d2"
    d1.mkdir(); d2.mkdir()
    _run(["-p", str(fa), "-q", str(d1)] + args)
    _run(["-p", str(gz), "-q", str(d2)] + args, env={"MHAPB_FASTA_CHUNK_KB": "64"})
    assert (d1 / "store.dat").read_bytes() == (d2 / "store_gz.fasta.dat").read_bytes()   # only the last extension is stripped (MhapMain.java:431-434)


def test_store_full_id_prints_fasta_names(tmp_path):
    # --store-full-id (main/MhapMain.java:301-304, impl/FastaData.java:155-156): the first token of the header line replaces the
    # file position in both id columns; query ids continue after the store's
    store, query = _reads(100, 2000, 21), _reads(60, 2000, 22)
    fa, qa = tmp_path / "store.fasta", tmp_path / "query.fasta"
    _write_fasta(fa, store)        # headers ">read_<i> some description"
    _write_fasta(qa, query)
    num, _ = _run(["-s", str(fa), "-q", str(qa), "--num-hashes", "256", "--no-self"])
    named, _ = _run(["-s", str(fa), "-q", str(qa), "--num-hashes", "256", "--no-self", "--store-full-id"], env={"MHAPB_FASTA_CHUNK_KB": "64"})
    # query ids start after the number of STORED sequences (main/MhapMain.java:462,537), store ids are file positions: the two
    # ranges overlap when short store reads were skipped, so each column is renamed through its own file
    offset = sum(len(r) >= 116 for r in store)
    assert offset < 100
    def rename(line):
        f = line.split(" ")
        return " ".join([f"read_{int(f[0]) - offset - 1}", f"read_{int(f[1]) - 1}"] + f[2:])
    self_num, _ = _run(["-s", str(fa), "--num-hashes", "256"])
    self_named, _ = _run(["-s", str(fa), "--num-hashes", "256", "--store-full-id"])
    def rename_self(line):
        f = line.split(" ")
        return " ".join([f"read_{int(f[0]) - 1}", f"read_{int(f[1]) - 1}"] + f[2:])
    assert sorted(map(rename_self, self_num)) == self_named and len(self_named) > 20
    assert sorted(map(rename, num)) == named and len(named) > 50
    # -p with --store-full-id: the header string stored in every .dat record is the FASTA name (SequenceId.getHeader,
    # impl/SequenceId.java:102-108; SequenceSketch.getAsByteArray writes it with writeUTF), without it the decimal id
    d1, d2 = tmp_path / "dat_num", tmp_path / "dat_named"
    d1.mkdir(); d2.mkdir()
    _run(["-p", str(fa), "-q", str(d1), "--num-hashes", "64"])
    _run(["-p", str(fa), "-q", str(d2), "--num-hashes", "64", "--store-full-id"])
    def headers(blob):
        out, off = [], 0
        while off < len(blob):
            size = int.from_bytes(blob[off + 1:off + 5], "big")
            hl = int.from_bytes(blob[off + 5 + 9:off + 5 + 11], "big")
            out.append(blob[off + 5 + 11:off + 5 + 11 + hl].decode())
            off += 5 + size
        return out
    hn, hf = headers((d1 / "store.dat").read_bytes()), headers((d2 / "store.dat").read_bytes())
    kept = [i for i, r in enumerate(store) if len(r) >= 116]
    assert hn == [str(i + 1) for i in kept for _ in (0, 1)]
    assert hf == [f"read_{i}" for i in kept for _ in (0, 1)]


def test_bad_arguments_exit_like_the_reference(tmp_path):
    p = subprocess.run([CLI], capture_output=True, text=True)
    assert p.returncode == 1 and "Please set the -s or the -p options." in p.stdout
    p = subprocess.run([CLI, "-s", str(tmp_path / "missing.fa")], capture_output=True, text=True)
    assert p.returncode == 1 and "Could not find requested file/folder" in p.stdout
    bad = tmp_path / "bad.fa"
    bad.write_text("ACGT\n")
    p = subprocess.run([CLI, "-s", str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and "Next sequence does not start with >" in p.stderr
