"""CPU suite: the host driver's streaming FASTA producer (mhap_b200/host/fasta_stream.hpp) against a
line-by-line Python statement of impl/FastaData.java:125-204, over chunk sizes that cut records everywhere, several parser
threads, CRLF, blank lines, '>' inside sequence lines, records larger than a chunk, an empty record, and gzip input."""
import gzip
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("fs") / "fasta_stream_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(out), os.path.join(ROOT, "tests", "cpp", "fasta_stream_test.cpp"), "-lz", "-ldl"])
    return str(out)


def reference_parse(text: bytes):
    """FastaData.enqueueNextSequenceInFile, as the driver's former getline parser stated it."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    lines = [l[:-1] if l.endswith(b"\r") else l for l in lines]
    recs, i = [], 0
    HEADERS.clear()
    if not lines:
        return recs, False, None
    while i < len(lines):
        if not lines[i] or lines[i][:1] != b">":
            return recs, False, "Next sequence does not start with >. Invalid format."
        import re
        HEADERS.append(re.split(rb"[\s,]+", lines[i][1:], maxsplit=1)[0])      # FastaData.java:155-156
        i += 1
        seq = b""
        while i < len(lines) and lines[i][:1] != b">":
            seq += lines[i]
            i += 1
        if not seq:
            return recs, True, None
        recs.append(seq)
    return recs, False, None


HEADERS = []


def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def run(harness, path, threads, chunk, headers=False):
    out = subprocess.run([harness, path, str(threads), str(chunk)] + (["h"] if headers else []), capture_output=True, text=True, timeout=120).stdout.splitlines()
    return out


def check(harness, tmp_path, text, name="x.fa", configs=((1, 1 << 16), (3, 1 << 16), (4, 100000), (2, 1 << 22))):
    p = tmp_path / name
    if name.endswith(".gz"):
        with gzip.open(p, "wb") as f:
            f.write(text)
    elif name.endswith(".bz2"):
        import bz2
        p.write_bytes(bz2.compress(text))
    else:
        p.write_bytes(text)
    recs, ended, err = reference_parse(text)
    for threads, chunk in configs:
        out = run(harness, str(p), threads, chunk)
        if err:
            assert out[-1] == "ERROR " + err, (threads, chunk)
            continue
        assert out[-1] == f"END {len(recs)} {int(ended)}", (threads, chunk, out[-1])
        assert out[:-1] == [f"{i} {len(r)} {fnv(r)}" for i, r in enumerate(recs)], (threads, chunk)
    if not err:
        out = run(harness, str(p), 2, 1 << 16, headers=True)
        assert out[:-1] == [f"{i} {len(r)} {fnv(r)} [{HEADERS[i].decode()}]" for i, r in enumerate(recs)]


def make(rng, n, lo, hi, width=70, crlf=False, weird=True):
    nl = b"\r\n" if crlf else b"\n"
    parts = []
    for i in range(n):
        L = rng.randint(lo, hi)
        s = bytes(rng.choice(b"ACGTacgtN") for _ in range(L))
        parts.append((b">read_%d some > text" % i if i % 3 else b">r%d,x y" % i) + nl)
        for j in range(0, L, width):
            line = s[j:j + width]
            if weird and rng.random() < 0.02 and len(line) > 2:
                line = line[:1] + b">" + line[2:]          # '>' inside a sequence line is sequence
            parts.append(line + nl)
            if weird and rng.random() < 0.01:
                parts.append(nl)                            # blank line inside a record
    return b"".join(parts)


def test_random_files_all_chunkings(harness, tmp_path):
    rng = random.Random(1)
    check(harness, tmp_path, make(rng, 400, 50, 3000))
    check(harness, tmp_path, make(rng, 300, 50, 3000, crlf=True), name="crlf.fa")
    check(harness, tmp_path, make(rng, 2000, 1, 40, width=7), name="tiny.fa")
    check(harness, tmp_path, make(rng, 50, 1, 2000)[:-1], name="no_final_newline.fa")


def test_record_larger_than_chunk_and_gzip(harness, tmp_path):
    rng = random.Random(2)
    text = make(rng, 5, 100, 200) + make(rng, 1, 400000, 400000, weird=False) + make(rng, 20, 100, 5000)
    check(harness, tmp_path, text, name="big.fa")
    check(harness, tmp_path, text, name="big.fa.gz", configs=((2, 1 << 16), (1, 1 << 20)))
    check(harness, tmp_path, text, name="big.fa.bz2", configs=((2, 1 << 16), (1, 1 << 20)))


def test_empty_record_ends_the_file_and_bad_first_line(harness, tmp_path):
    rng = random.Random(3)
    a, b = make(rng, 30, 100, 900, weird=False), make(rng, 30, 100, 900, weird=False)
    check(harness, tmp_path, a + b">empty\n" + b, name="empty_mid.fa")
    check(harness, tmp_path, a + b">empty_at_end\n", name="empty_end.fa")
    check(harness, tmp_path, b"ACGT\n" + a, name="bad.fa")
    check(harness, tmp_path, b"\n" + a, name="blank_first.fa")
    check(harness, tmp_path, b"", name="zero.fa")
