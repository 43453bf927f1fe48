#!/usr/bin/env python3
"""Pin the oracle against the real reference on a box that has a JDK (this image has none: DESIGN.md section 2).

  python tests/golden/check_against_jvm.py --write-cases            # writes tests/golden/pin_cases.txt (+ pin_filter.txt)
  javac -cp mhap-2.1.3.jar integration/PinOracle.java ; java -cp mhap-2.1.3.jar:integration PinOracle tests/golden/pin_cases.txt > jvm.txt
  python tests/golden/check_against_jvm.py jvm.txt                  # compares line by line with oracle/mhap_oracle.c

Without arguments it prints what the ORACLE says for the same cases (the format PinOracle prints), which is what
tests/test_oracle.py::test_pin_cases_self_consistent checks against the committed pin_expected.txt."""
import os
import random
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from tests.filter_common import make_reads_and_filter  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES, FILTER, EXPECTED = (os.path.join(HERE, n) for n in ("pin_cases.txt", "pin_filter.txt", "pin_expected.txt"))


def make_cases():
    rng = random.Random(2024)
    rs = lambda n, alpha="ACGT": "".join(rng.choice(alpha) for _ in range(n))
    g = rs(1500)
    lines = []
    for s in ("ACGTACGTACGTACGT", rs(40), rs(33, "ACGTN"), "ACGTTGCA" * 6):
        lines.append(f"H 16 {s}")
        lines.append(f"H 12 {s}")
    for s in (g[:300], g[100:500], "ACGTTGCA" * 40, "A" * 60 + rs(100), rs(200, "ACGTNRY")):
        for rw in (0.9, -1.0):
            lines.append(f"M 16 32 {rw} {s}")
        lines.append(f"B 12 40 {s}")
        lines.append(f"B 12 1536 {s}")
    mut = lambda s: "".join(c if rng.random() > 0.04 else rng.choice("ACGT") for c in s)
    for a, b in ((g[:900], mut(g[300:1200])), (g[200:1000], mut(g[:800])), (g[:600], g[:600]), (g[:500], rs(500)), ("ACGTTGCA" * 60, "ACGTTGCA" * 50)):
        lines.append(f"O 12 200 0.2 {a} {b}")
    reads, text = make_reads_and_filter(seed=9, n_reads=6, read_len=200)
    for rw, sn, notf in ((0.9, 0, 0), (0.9, 1, 0), (0.9, 2, 1), (-1.0, 0, 0), (0.3, 0, 0), (1.0, 1, 0)):
        for r in reads:
            lines.append(f"F 16 16 {rw} 1e-05 {sn} {notf} 3.0 tests/golden/pin_filter.txt {r}")
    return lines, text


def oracle_line(line, filter_text):
    f = line.split(" ")
    j = lambda a: "".join(f" {int(v)}" for v in a)
    if f[0] == "H":
        k = int(f[1])
        return [f"H64{j(orc.kmer_hashes_long(f[2], k))}", f"H64C{j(orc.kmer_hashes_long(f[2], k, 0, True))}", f"H32{j(orc.kmer_hashes_int(f[2], k))}"]
    if f[0] == "M":
        m = orc.minhash_sketch_filtered(f[4], int(f[1]), int(f[2]), float(f[3]), None)
        return ["M ZERO" if m is None else f"M{j(m)}"]
    if f[0] == "B":
        od, slen = orc.bottom_sketch(f[3], int(f[1]), int(f[2]))
        if od is None:
            return ["B ZERO"]
        return [f"B {slen} {int(f[1])} {od.shape[0]}{j(od.reshape(-1))}"]   # getAsByteArray: seqLength, kmerSize, n, pairs
    if f[0] == "O":
        ok, S = int(f[1]), int(f[2])
        a, la = orc.bottom_sketch(f[4], ok, S)
        b, lb = orc.bottom_sketch(f[5], ok, S)
        o = orc.overlap_info(a, la, b, lb, ok, float(f[3]))
        if o.empty:
            return [f"O 0 0 0 0 0 {struct.unpack('<q', struct.pack('<d', 0.0))[0]}"]
        return [f"O {o.a1} {o.a2} {o.b1} {o.b2} {o.valid_count} {struct.unpack('<q', struct.pack('<d', o.score))[0]}"]
    if f[0] == "F":
        rw = float(f[3])
        kf = orc.KmerFilter(filter_text, repeat_weight=rw, filter_cutoff=float(f[4]), supress_noise=int(f[5]), no_tf=f[6] == "1", idf_scale=float(f[7]))
        m = orc.minhash_sketch_filtered(f[9], int(f[1]), int(f[2]), rw, kf)
        return ["F ZERO" if m is None else f"F{j(m)}"]
    return [f"? {f[0]}"]


def oracle_output():
    lines, text = make_cases()
    out = []
    for l in lines:
        out.extend(oracle_line(l, text))
    return lines, text, out


def main():
    lines, text, out = oracle_output()
    if len(sys.argv) > 1 and sys.argv[1] == "--write-cases":
        open(CASES, "w").write("\n".join(lines) + "\n")
        open(FILTER, "w").write(text)
        open(EXPECTED, "w").write("\n".join(out) + "\n")
        print(f"wrote {len(lines)} cases, {len(out)} expected lines")
        return 0
    if len(sys.argv) > 1:
        jvm = [l.rstrip("\n") for l in open(sys.argv[1]) if l.strip()]
        bad = [(i, a, b) for i, (a, b) in enumerate(zip(out, jvm)) if a != b]
        # the score is compared as raw double bits: a 1-ulp libm difference (Math.log/exp vs C log/exp) shows up here and nowhere else
        print(f"{len(out)} oracle lines, {len(jvm)} JVM lines, {len(bad)} differ")
        for i, a, b in bad[:20]:
            print(f"line {i}:\n  oracle {a[:200]}\n  jvm    {b[:200]}")
        return 1 if bad or len(out) != len(jvm) else 0
    print("\n".join(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
