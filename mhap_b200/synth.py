"""Deterministic synthetic PacBio-shape reads: ctypes over csrc/synth.c.

Genome = uniform ACGT from splitmix64(seed); read i = window at a uniform start, per-base error
`err` split ins/del/sub = 0.792/0.122/0.086 (the reference simulator's mix,
main/KmerStatSimulator.java:230), exactly L bases, random strand.  Each read has its own PRNG
stream, so any rank can generate just its shard.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libmhap_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} not built: run `make -C mhap_b200/csrc`")
        L = C.CDLL(_PATH)
        L.mhapb_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        L.mhapb_synth_genome.restype = None
        L.mhapb_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                        C.c_double, C.c_int, C.c_void_p]
        L.mhapb_synth_reads.restype = None
        _lib = L
    return _lib


def genome(seed: int, length: int) -> np.ndarray:
    g = np.empty(length, dtype=np.uint8)
    _load().mhapb_synth_genome(seed, length, g.ctypes.data)
    return g


def reads(g: np.ndarray, read_seed: int, first: int, n: int, L: int, err: float = 0.15, threads: int = 0, out=None):
    """Returns (uint8 bases [n*L], uint64 offsets [n+1]) for reads [first, first+n)."""
    if threads <= 0:
        threads = min(64, os.cpu_count() or 1)
    if out is None:
        out = np.empty(n * L, dtype=np.uint8)
    _load().mhapb_synth_reads(g.ctypes.data, g.size, read_seed, first, n, L, err, threads, out.ctypes.data)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    return out, offsets


def dataset(n_reads: int, L: int, seed: int, coverage: float = 20.0, err: float = 0.15, first: int = 0, count: int | None = None,
            genome_seed: int | None = None):
    """The SURVEY 8d recipe: genome of n_reads*L/coverage bases (seeded by genome_seed or seed)."""
    glen = max(L + 1, int(n_reads * L / coverage))
    g = genome(seed if genome_seed is None else genome_seed, glen)
    return reads(g, seed * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF, first, n_reads - first if count is None else count, L, err)
