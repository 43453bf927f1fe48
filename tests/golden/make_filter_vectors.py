#!/usr/bin/env python3
"""Writes tests/golden/filter_vectors.json.  The reference ships no vectors for the -f path and cannot run here (no JVM):
these are DERIVED FROM THE RESTATEMENT (oracle/pyref.py, the pure-Python one) and serve to cross-check independent
implementations and to catch regressions, not as ground truth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from tests.filter_common import make_reads_and_filter  # noqa: E402

reads, text = make_reads_and_filter(seed=9, n_reads=6, read_len=200)
H = 16
cases = []
for rw, sn, notf in [(0.9, 0, False), (0.9, 1, False), (0.9, 2, True), (-1.0, 0, False), (0.3, 0, False), (1.0, 1, False)]:
    off = rw if 0.0 <= rw < 1.0 else 0.0
    fc = pyref.FrequencyCounts(text, 1.0e-5, off, sn, notf, 3.0, True)
    cases.append({"repeat_weight": rw, "supress_noise": sn, "no_tf": notf,
                  "minhash": [pyref.minhash_sketch_filtered(r, 16, H, rw, fc) for r in reads]})
json.dump({"source": "oracle/pyref.py (restatement only, not a JVM run)", "num_hashes": H, "reads": reads, "filter_text": text,
           "cases": cases}, open(os.path.join(ROOT, "tests", "golden", "filter_vectors.json"), "w"), indent=0)
