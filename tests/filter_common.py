"""Synthetic -f k-mer filter files for the filter tests (format: docs/source/utilities... none shipped; the layout is the one
FrequencyCounts' constructor parses, sketch/FrequencyCounts.java:91-207)."""
import random

FRACTIONS = [2e-6, 1e-5, 3e-5, 1e-4, 5e-4, 0.002]


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def make_reads_and_filter(seed=3, n_reads=14, read_len=400, k=16, every=7, extra_lines=()):
    """Reads cut from one random genome (so they overlap) plus two repeat-rich ones, and a filter file listing every
    `every`-th k-mer of the first reads with fractions on both sides of the 1e-5 cutoff."""
    rng = random.Random(seed)
    g = rand_seq(rng, read_len * 8)
    step = max(1, (len(g) - read_len) // max(1, n_reads - 2))
    reads = [g[i:i + read_len] for i in range(0, len(g) - read_len + 1, step)][:n_reads - 2]
    reads += [("ACGTTGCA" * (read_len // 8 + 1))[:read_len], "A" * (read_len // 4) + rand_seq(rng, read_len - read_len // 4)]
    kmers = set()
    for r in reads[: max(2, len(reads) // 2)]:
        for i in range(0, len(r) - k, every):
            kmers.add(r[i:i + k])
    kmers = sorted(kmers)
    lines = [f"{len(kmers) + 1 + len(extra_lines)} {len(kmers)}"]
    for i, km in enumerate(kmers):
        lines.append(f"{km}\t{FRACTIONS[i % len(FRACTIONS)]}\t{i}")
    lines.append("AAAAAAAAAAAAAAAA 0.01 7")
    lines.extend(extra_lines)
    return reads, "\n".join(lines) + "\n"


SETTINGS = [(rw, sn, notf) for rw in (0.9, 0.0, 0.5, -1.0, 1.0, 2.0) for sn in (0, 1, 2) for notf in (False, True)]
