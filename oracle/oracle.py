"""ctypes binding of oracle/libmhap_oracle.so (the C restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- never imported by mhap_b200/.  PARITY UNPINNED (see mhap_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmhap_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mhap_oracle.c")
    hdr = os.path.join(_HERE, "mhap_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmhap_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Overlap(C.Structure):
    _fields_ = [("empty", C.c_int32), ("a1", C.c_int32), ("a2", C.c_int32), ("b1", C.c_int32), ("b2", C.c_int32),
                ("valid_count", C.c_int32), ("intersect", C.c_int32), ("kmin", C.c_int32), ("score", C.c_double)]


class SketchParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("num_hashes", C.c_int32), ("ordered_kmer_size", C.c_int32),
                ("ordered_sketch_size", C.c_int32), ("unweighted", C.c_int32), ("min_olap_length", C.c_int32)]


class FilterParams(C.Structure):
    _fields_ = [("filter_cutoff", C.c_double), ("offset", C.c_double), ("range", C.c_double),
                ("remove_unique", C.c_int32), ("no_tf", C.c_int32), ("canonical", C.c_int32)]


class SearchParams(C.Structure):
    _fields_ = [("num_min_matches", C.c_int32), ("min_store_length", C.c_int32), ("max_shift", C.c_double),
                ("accept_score", C.c_double)]


class Hit(C.Structure):
    _fields_ = [("from_id", C.c_int64), ("to_id", C.c_int64), ("from_fwd", C.c_int32), ("to_fwd", C.c_int32),
                ("hit_count", C.c_int32), ("a1", C.c_int32), ("a2", C.c_int32), ("b1", C.c_int32), ("b2", C.c_int32),
                ("valid_count", C.c_int32), ("intersect", C.c_int32), ("kmin", C.c_int32), ("from_len", C.c_int32),
                ("to_len", C.c_int32), ("score", C.c_double), ("accepted", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("elements_processed", C.c_int64), ("sequences_hit", C.c_int64), ("fully_compared", C.c_int64),
                ("matches_processed", C.c_int64), ("sequences_searched", C.c_int64)]


HIT_DTYPE = np.dtype([("from_id", "<i8"), ("to_id", "<i8"), ("from_fwd", "<i4"), ("to_fwd", "<i4"),
                      ("hit_count", "<i4"), ("a1", "<i4"), ("a2", "<i4"), ("b1", "<i4"), ("b2", "<i4"),
                      ("valid_count", "<i4"), ("intersect", "<i4"), ("kmin", "<i4"), ("from_len", "<i4"),
                      ("to_len", "<i4"), ("score", "<f8"), ("accepted", "<i4"), ("_pad", "<i4")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.mo_murmur3_x64_128.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint64)]
        L.mo_murmur3_x86_32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
        L.mo_murmur3_x86_32.restype = C.c_uint32
        L.mo_rc.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
        L.mo_quick_select.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.mo_quick_select.restype = C.c_int32
        L.mo_kmer_hashes_long.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_uint32, C.c_int, C.c_void_p]
        L.mo_kmer_hashes_long.restype = C.c_int64
        L.mo_kmer_hashes_int.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.mo_kmer_hashes_int.restype = C.c_int64
        L.mo_minhash_sketch.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mo_filter_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(FilterParams)]
        L.mo_filter_parse.restype = C.c_void_p
        L.mo_filter_free.argtypes = [C.c_void_p]
        L.mo_filter_free.restype = None
        L.mo_filter_size.argtypes = [C.c_void_p]
        L.mo_filter_size.restype = C.c_int64
        L.mo_filter_max_value.argtypes = [C.c_void_p]
        L.mo_filter_max_value.restype = C.c_double
        L.mo_filter_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mo_filter_export.restype = C.c_int64
        L.mo_filter_bloom_export.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
        L.mo_filter_bloom_export.restype = C.c_int64
        L.mo_filter_is_popular.argtypes = [C.c_void_p, C.c_int64]
        L.mo_filter_keep_kmer.argtypes = [C.c_void_p, C.c_int64]
        L.mo_filter_scaled_idf.argtypes = [C.c_void_p, C.c_int64]
        L.mo_filter_scaled_idf.restype = C.c_double
        L.mo_minhash_sketch_filtered.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.mo_store_set_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.mo_store_set_filter.restype = None
        L.mo_bottom_sketch.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32)]
        L.mo_bottom_sketch.restype = C.c_int32
        L.mo_overlap_info.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int,
                                      C.c_double, C.POINTER(Overlap)]
        L.mo_overlap_info.restype = None
        L.mo_jaccard_to_identity.argtypes = [C.c_double, C.c_int]
        L.mo_jaccard_to_identity.restype = C.c_double
        L.mo_store_new.argtypes = [C.POINTER(SketchParams)]
        L.mo_store_new.restype = C.c_void_p
        L.mo_store_free.argtypes = [C.c_void_p]
        L.mo_store_add_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.mo_store_add_reads.restype = C.c_int64
        L.mo_store_add_sketch.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_int32]
        L.mo_store_size.argtypes = [C.c_void_p]
        L.mo_store_size.restype = C.c_int64
        L.mo_store_get.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
        L.mo_store_build_index.argtypes = [C.c_void_p]
        L.mo_store_build_index.restype = None
        L.mo_search_self.argtypes = [C.c_void_p, C.POINTER(SearchParams), C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int64), C.POINTER(Stats)]
        L.mo_search_self_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(SearchParams), C.c_int, C.c_int,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Stats)]
        L.mo_search_query.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(SearchParams), C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Stats)]
        L.mo_search_query_self.argtypes = L.mo_search_query.argtypes
        L.mo_free.argtypes = [C.c_void_p]
        L.mo_free.restype = None
        L.mo_format_match.argtypes = [C.POINTER(Hit), C.c_char_p, C.c_size_t]
        L.mo_dat_encode.argtypes = [C.c_int64, C.c_int, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.mo_dat_encode.restype = C.c_int64
        _lib = L
    return _lib


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1")


def murmur3_x64_128(data: bytes, seed: int = 0):
    out = (C.c_uint64 * 2)()
    lib().mo_murmur3_x64_128(data, len(data), seed & 0xFFFFFFFF, out)
    return int(out[0]), int(out[1])


def murmur3_x86_32(data: bytes, seed: int = 0) -> int:
    return int(lib().mo_murmur3_x86_32(data, len(data), seed & 0xFFFFFFFF))


def rc(seq) -> bytes:
    s = _b(seq)
    out = C.create_string_buffer(len(s))
    lib().mo_rc(s, len(s), out)
    return out.raw


def quick_select(arr, k: int) -> int:
    a = np.ascontiguousarray(arr, dtype=np.int32).copy()
    return int(lib().mo_quick_select(a.ctypes.data, k, a.size))


def kmer_hashes_long(seq, k: int, seed: int = 0, canonical: bool = False) -> np.ndarray:
    s = _b(seq)
    n = max(0, len(s) - k + 1)
    out = np.zeros(n, dtype=np.int64)
    lib().mo_kmer_hashes_long(s, len(s), k, seed, int(canonical), out.ctypes.data)
    return out


def kmer_hashes_int(seq, k: int, canonical: bool = False) -> np.ndarray:
    s = _b(seq)
    n = max(0, len(s) - k + 1)
    out = np.zeros(n, dtype=np.int32)
    lib().mo_kmer_hashes_int(s, len(s), k, int(canonical), out.ctypes.data)
    return out


def minhash_sketch(seq, k: int, num_hashes: int, unweighted: bool = False):
    """Returns int32[H] or None when the reference would throw ZeroNGramsFoundException."""
    s = _b(seq)
    out = np.zeros(max(1, num_hashes), dtype=np.int32)
    st = lib().mo_minhash_sketch(s, len(s), k, num_hashes, int(unweighted), out.ctypes.data)
    return None if st else out


class KmerFilter:
    """FrequencyCounts (sketch/FrequencyCounts.java) built from the text of a -f filter file.

    offset follows main/MhapMain.java:348-350: repeat_weight when 0 <= repeat_weight < 1, else 0."""

    def __init__(self, text, repeat_weight=0.9, filter_cutoff=1.0e-5, idf_scale=3.0, supress_noise=0, no_tf=False, canonical=True):
        t = _b(text)
        self.repeat_weight = float(repeat_weight)
        offset = self.repeat_weight if 0.0 <= self.repeat_weight < 1.0 else 0.0
        self.params = FilterParams(filter_cutoff, offset, idf_scale, supress_noise, int(no_tf), int(canonical))
        self._h = lib().mo_filter_parse(t, len(t), C.byref(self.params))

    def close(self):
        if self._h:
            lib().mo_filter_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(lib().mo_filter_size(self._h))

    @property
    def max_value(self):
        return float(lib().mo_filter_max_value(self._h))

    def export(self):
        """(int64 hashes, float64 fractions) of the repeat map."""
        n = len(self)
        h = np.zeros(max(n, 1), dtype=np.int64); f = np.zeros(max(n, 1), dtype=np.float64)
        lib().mo_filter_export(self._h, h.ctypes.data, f.ctypes.data)
        return h[:n], f[:n]

    def bloom(self):
        """(uint64 words, bit_size, num_hash_functions) of the Guava-layout Bloom filter, or (None, 0, 0)."""
        w = C.c_void_p(); nh = C.c_int32()
        bits = int(lib().mo_filter_bloom_export(self._h, C.byref(w), C.byref(nh)))
        if not bits:
            return None, 0, 0
        words = np.ctypeslib.as_array(C.cast(w, C.POINTER(C.c_uint64)), shape=(bits // 64,)).copy()
        return words, bits, int(nh.value)

    def is_popular(self, h):
        return bool(lib().mo_filter_is_popular(self._h, int(h)))

    def keep_kmer(self, h):
        return bool(lib().mo_filter_keep_kmer(self._h, int(h)))

    def scaled_idf(self, h):
        return float(lib().mo_filter_scaled_idf(self._h, int(h)))


def minhash_sketch_filtered(seq, k: int, num_hashes: int, repeat_weight: float = 0.9, kmer_filter: "KmerFilter | None" = None):
    """MinHashSketch.java:51-179 with the -f filter; None when the reference would throw ZeroNGramsFoundException."""
    s = _b(seq)
    out = np.zeros(max(1, num_hashes), dtype=np.int32)
    st = lib().mo_minhash_sketch_filtered(s, len(s), k, num_hashes, float(repeat_weight), kmer_filter._h if kmer_filter else None,
                                          out.ctypes.data)
    return None if st else out


def bottom_sketch(seq, ok: int, sketch_size: int):
    """Returns (int32[n,2] (hash,pos), seq_len_kmers) or (None, seq_len_kmers)."""
    s = _b(seq)
    n = max(0, min(sketch_size, len(s) - ok + 1))
    out = np.zeros((max(n, 1), 2), dtype=np.int32)
    sl = C.c_int32(0)
    r = lib().mo_bottom_sketch(s, len(s), ok, sketch_size, out.ctypes.data, C.byref(sl))
    if r < 0:
        return None, int(sl.value)
    return out[:r].copy(), int(sl.value)


def overlap_info(a: np.ndarray, a_len: int, b: np.ndarray, b_len: int, ok: int = 12, max_shift: float = 0.2) -> Overlap:
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    out = Overlap()
    lib().mo_overlap_info(a.ctypes.data, a.shape[0], a_len, b.ctypes.data, b.shape[0], b_len, ok, max_shift, C.byref(out))
    return out


def jaccard_to_identity(j: float, ok: int) -> float:
    return float(lib().mo_jaccard_to_identity(j, ok))


@dataclass
class SearchResult:
    hits: np.ndarray  # HIT_DTYPE
    stats: dict


def pack_reads(reads):
    """list[bytes|str] -> (uint8 bases, uint64 offsets[n+1])."""
    bs = [_b(r) for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return bases, offsets


class Store:
    """Mirror of MinHashSearch's stored sketches + index (oracle side)."""

    def __init__(self, k=16, num_hashes=512, ordered_k=12, ordered_size=1536, unweighted=False, min_olap_length=116):
        self.params = SketchParams(k, num_hashes, ordered_k, ordered_size, int(unweighted), min_olap_length)
        self._h = lib().mo_store_new(C.byref(self.params))

    def close(self):
        if self._h:
            lib().mo_store_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_filter(self, kmer_filter: "KmerFilter | None", repeat_weight: float = 0.9):
        """Sketches added afterwards use the -f filter and the full --repeat-weight semantics."""
        self._filter = kmer_filter   # keep it alive
        lib().mo_store_set_filter(self._h, kmer_filter._h if kmer_filter else None, float(repeat_weight))

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray, ids=None, both_strands=True, threads=1, id_offset=0) -> int:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        if ids is None:
            ids = np.arange(1, n + 1, dtype=np.int64) + id_offset  # FastaData.java:181 (1-based)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        if bases.size == 0:
            bases = np.zeros(1, dtype=np.uint8)
        return int(lib().mo_store_add_reads(self._h, bases.ctypes.data, offsets.ctypes.data, ids.ctypes.data, n,
                                            int(both_strands), threads))

    def add_sketch(self, id_, is_fwd, seq_len, minhash, seq_len_kmers, ord_hp):
        mh = np.ascontiguousarray(minhash, dtype=np.int32)
        oh = np.ascontiguousarray(ord_hp, dtype=np.int32)
        lib().mo_store_add_sketch(self._h, id_, int(is_fwd), seq_len, mh.ctypes.data, seq_len_kmers, oh.ctypes.data, oh.shape[0])

    def __len__(self):
        return int(lib().mo_store_size(self._h))

    def get(self, idx):
        id_ = C.c_int64(); fwd = C.c_int32(); sl = C.c_int32(); mh = C.c_void_p(); slk = C.c_int32()
        oh = C.c_void_p(); on = C.c_int32()
        r = lib().mo_store_get(self._h, idx, C.byref(id_), C.byref(fwd), C.byref(sl), C.byref(mh), C.byref(slk),
                               C.byref(oh), C.byref(on))
        if r:
            raise IndexError(idx)
        H = self.params.num_hashes
        minhash = np.ctypeslib.as_array(C.cast(mh, C.POINTER(C.c_int32)), shape=(H,)).copy()
        if on.value > 0:
            ord_hp = np.ctypeslib.as_array(C.cast(oh, C.POINTER(C.c_int32)), shape=(on.value, 2)).copy()
        else:
            ord_hp = np.zeros((0, 2), dtype=np.int32)
        return dict(id=id_.value, is_fwd=bool(fwd.value), seq_len=sl.value, minhash=minhash, seq_len_kmers=slk.value,
                    ord=ord_hp)

    def build_index(self):
        lib().mo_store_build_index(self._h)

    def _collect(self, out, n, st):
        if n.value:
            buf = C.string_at(out.value, n.value * C.sizeof(Hit))
            hits = np.frombuffer(buf, dtype=HIT_DTYPE).copy()
        else:
            hits = np.zeros(0, dtype=HIT_DTYPE)
        lib().mo_free(out)
        return SearchResult(hits, {f: int(getattr(st, f)) for f, _ in Stats._fields_})

    def search_self(self, num_min_matches=3, min_store_length=0, max_shift=0.2, accept_score=0.78, threads=1, keep_all=False):
        sp = SearchParams(num_min_matches, min_store_length, max_shift, accept_score)
        out = C.c_void_p(); n = C.c_int64(); st = Stats()
        lib().mo_search_self(self._h, C.byref(sp), threads, int(keep_all), C.byref(out), C.byref(n), C.byref(st))
        return self._collect(out, n, st)

    def search_self_range(self, first, count, num_min_matches=3, min_store_length=0, max_shift=0.2, accept_score=0.78,
                          threads=1, keep_all=False):
        sp = SearchParams(num_min_matches, min_store_length, max_shift, accept_score)
        out = C.c_void_p(); n = C.c_int64(); st = Stats()
        lib().mo_search_self_range(self._h, first, count, C.byref(sp), threads, int(keep_all), C.byref(out), C.byref(n), C.byref(st))
        return self._collect(out, n, st)

    def search_query(self, queries: "Store", num_min_matches=3, min_store_length=0, max_shift=0.2, accept_score=0.78,
                     threads=1, keep_all=False, to_self=False):
        sp = SearchParams(num_min_matches, min_store_length, max_shift, accept_score)
        out = C.c_void_p(); n = C.c_int64(); st = Stats()
        (lib().mo_search_query_self if to_self else lib().mo_search_query)(self._h, queries._h, C.byref(sp), threads, int(keep_all), C.byref(out), C.byref(n), C.byref(st))
        return self._collect(out, n, st)


def format_match(hit_row) -> str:
    h = Hit()
    for f, _ in Hit._fields_:
        setattr(h, f, hit_row[f].item() if hasattr(hit_row[f], "item") else hit_row[f])
    buf = C.create_string_buffer(256)
    lib().mo_format_match(C.byref(h), buf, 256)
    return buf.value.decode()


def dat_encode(id_, is_fwd, seq_len, minhash, seq_len_kmers, ordered_k, ord_hp, header=None) -> bytes:
    mh = np.ascontiguousarray(minhash, dtype=np.int32)
    oh = np.ascontiguousarray(ord_hp, dtype=np.int32).reshape(-1, 2)
    hdr = None if header is None else _b(header)
    n = lib().mo_dat_encode(id_, int(is_fwd), hdr, seq_len, mh.ctypes.data, mh.size, seq_len_kmers, ordered_k,
                            oh.ctypes.data, oh.shape[0], None)
    buf = C.create_string_buffer(n)
    lib().mo_dat_encode(id_, int(is_fwd), hdr, seq_len, mh.ctypes.data, mh.size, seq_len_kmers, ordered_k,
                        oh.ctypes.data, oh.shape[0], buf)
    return buf.raw
