#!/usr/bin/env python
"""End-to-end run of the host driver on a synthetic FASTA (BASELINE configs[1] shape by default): FASTA parse (streamed, pinned
batches) + H2D + K1 + index + self search + output lines, as a user of `mhap-b200 -s reads.fasta` sees it."""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mhap_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100_000)
ap.add_argument("--read-len", type=int, default=10_000)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--width", type=int, default=0, help="FASTA line width (0: one line per sequence)")
a = ap.parse_args()
bases, offs = synth.dataset(a.reads, a.read_len, seed=2)
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
fa = os.path.join(d, "reads.fasta")
t0 = time.time()
with open(fa, "wb") as f:
    for i in range(a.reads):
        f.write(b">%d\n" % (i + 1))
        r = bases[int(offs[i]):int(offs[i + 1])]
        if a.width:
            for j in range(0, r.size, a.width):
                f.write(r[j:j + a.width].tobytes()); f.write(b"\n")
        else:
            f.write(r.tobytes()); f.write(b"\n")
t_write = time.time() - t0
out = os.path.join(d, "out.ovl")
res = {}
for label, env in (("warmup", {}), ("streamed", {}), ("single_batch", {"MHAPB_FASTA_CHUNK_KB": str(4 << 20)})):
    t0 = time.time()
    with open(out, "wb") as fo:
        p = subprocess.run([os.path.join(ROOT, "mhap_b200", "mhap-b200"), "-s", fa, "--num-hashes", "512", "--num-threads", str(a.threads)],
                           stdout=fo, stderr=subprocess.PIPE, text=True, env=dict(os.environ, **env))
    wall = time.time() - t0
    assert p.returncode == 0, p.stderr[-2000:]
    times = {l.split(":")[0].strip(): float(l.split(":")[1]) for l in p.stderr.splitlines() if l.startswith("Time (s)") or l.startswith("Total")}
    res[label] = dict(wall_s=wall, read_and_hash_s=times.get("Time (s) to read and hash from file"),
                      score_s=times.get("Time (s) to score and output to self"), total_s=times.get("Total time (s)"),
                      gbases_per_s_read_and_hash=a.reads * a.read_len / times["Time (s) to read and hash from file"] / 1e9,
                      gbases_per_s_total=a.reads * a.read_len / times["Total time (s)"] / 1e9, overlaps=sum(1 for _ in open(out, "rb")))
print(json.dumps(dict(reads=a.reads, read_len=a.read_len, fasta_bytes=os.path.getsize(fa), fasta_write_s=t_write, threads=a.threads, **res)))
