"""ctypes binding of libmhap_b200.so -- the C ABI declared in include/mhap_b200.h.

This is the same binding a JNI shim would make; Python is only the harness
language of this image (no JVM here).  There is no CPU fallback: loading fails loudly when the
library is missing and `Engine()` fails when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmhap_b200.so")

# every symbol include/mhap_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "mhapb_version", "mhapb_create", "mhapb_destroy", "mhapb_last_error", "mhapb_free", "mhapb_get_timing",
    "mhapb_host_alloc", "mhapb_host_free", "mhapb_xorshift_peak", "mhapb_xorshift_peaks", "mhapb_sketch", "mhapb_sketch_device", "mhapb_sketch_to_dat", "mhapb_sketch_to_dat_named",
    "mhapb_dat_encode", "mhapb_dat_decode", "mhapb_store_reset", "mhapb_store_add_reads", "mhapb_store_add_reads_device",
    "mhapb_store_add_sketches", "mhapb_store_add_sketches_device", "mhapb_store_size", "mhapb_store_get",
    "mhapb_store_get_range", "mhapb_store_params", "mhapb_store_device_ptrs", "mhapb_index_build", "mhapb_search_self", "mhapb_search_query_reads",
    "mhapb_search_query_sketches", "mhapb_search_sketches_device", "mhapb_format_match", "mhapb_minhash_equal_count",
    "mhapb_store_reserve", "mhapb_sketch_reserve", "mhapb_kmer_hash", "mhapb_filter_set", "mhapb_filter_load_text", "mhapb_filter_clear",
    "mhapb_comm_unique_id", "mhapb_comm_init_rank", "mhapb_comm_init_all", "mhapb_comm_destroy", "mhapb_comm_info",
    "mhapb_dist_search_self", "mhapb_dist_search_query_reads", "mhapb_dist_search_query_reads_device",
    "mhapb_multi_create", "mhapb_multi_destroy", "mhapb_multi_last_error", "mhapb_multi_n_devices", "mhapb_multi_ctx",
    "mhapb_multi_store_reset", "mhapb_multi_store_reserve", "mhapb_multi_store_add_reads", "mhapb_multi_store_add_sketches",
    "mhapb_multi_store_size", "mhapb_multi_search_self", "mhapb_multi_search_query_reads", "mhapb_multi_search_query_sketches",
]
COMM_ID_BYTES = 128


class MhapError(RuntimeError):
    """Mirrors the reference's unchecked MhapRuntimeException / SketchRuntimeException."""

    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class SketchParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("num_hashes", C.c_int32), ("ordered_kmer_size", C.c_int32),
                ("ordered_sketch_size", C.c_int32), ("unweighted", C.c_int32), ("min_olap_length", C.c_int32)]


class SearchParams(C.Structure):
    _fields_ = [("num_min_matches", C.c_int32), ("min_store_length", C.c_int32), ("max_shift", C.c_double),
                ("accept_score", C.c_double), ("keep_all", C.c_int32), ("reserved", C.c_int32),
                ("query_first", C.c_int64), ("query_count", C.c_int64)]


class FilterParams(C.Structure):
    """mhapb_filter_params: the FrequencyCounts constructor arguments (sketch/FrequencyCounts.java:63)."""
    _fields_ = [("filter_cutoff", C.c_double), ("repeat_weight", C.c_double), ("idf_scale", C.c_double),
                ("supress_noise", C.c_int32), ("no_tf", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("from_id", C.c_int64), ("to_id", C.c_int64), ("from_fwd", C.c_int32), ("to_fwd", C.c_int32),
                ("hit_count", C.c_int32), ("a1", C.c_int32), ("a2", C.c_int32), ("b1", C.c_int32), ("b2", C.c_int32),
                ("valid_count", C.c_int32), ("intersect", C.c_int32), ("kmin", C.c_int32), ("from_len", C.c_int32),
                ("to_len", C.c_int32), ("score", C.c_double), ("accepted", C.c_int32), ("pad_", C.c_int32)]


HIT_DTYPE = np.dtype([("from_id", "<i8"), ("to_id", "<i8"), ("from_fwd", "<i4"), ("to_fwd", "<i4"),
                      ("hit_count", "<i4"), ("a1", "<i4"), ("a2", "<i4"), ("b1", "<i4"), ("b2", "<i4"),
                      ("valid_count", "<i4"), ("intersect", "<i4"), ("kmin", "<i4"), ("from_len", "<i4"),
                      ("to_len", "<i4"), ("score", "<f8"), ("accepted", "<i4"), ("pad_", "<i4")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit)


class Stats(C.Structure):
    _fields_ = [("elements_processed", C.c_int64), ("sequences_hit", C.c_int64), ("fully_compared", C.c_int64),
                ("matches_processed", C.c_int64), ("sequences_searched", C.c_int64)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("hash_dedup_ms", C.c_float), ("minhash_ms", C.c_float),
                ("ordered_ms", C.c_float), ("index_ms", C.c_float), ("probe_ms", C.c_float), ("filter_ms", C.c_float),
                ("kernel_launches", C.c_int64), ("xorshift_steps", C.c_int64), ("sketch_total_ms", C.c_float),
                ("search_total_ms", C.c_float), ("kmers_hashed", C.c_int64), ("gather_ms", C.c_float), ("pad_", C.c_float)]


_lib = None


def load():
    """dlopen libmhap_b200.so; raises if it has not been built (python __graft_entry__.py / make)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MhapError(-2, f"{LIB_PATH} not built: run `make -C mhap_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    P = C.POINTER
    L.mhapb_version.restype = C.c_char_p
    L.mhapb_create.argtypes = [C.c_int, P(vp)]
    L.mhapb_destroy.argtypes = [vp]; L.mhapb_destroy.restype = None
    L.mhapb_last_error.argtypes = [vp]; L.mhapb_last_error.restype = C.c_char_p
    L.mhapb_free.argtypes = [vp]; L.mhapb_free.restype = None
    L.mhapb_get_timing.argtypes = [vp, P(Timing)]
    L.mhapb_host_alloc.argtypes = [C.c_size_t, P(vp)]
    L.mhapb_host_free.argtypes = [vp]; L.mhapb_host_free.restype = None
    L.mhapb_xorshift_peak.argtypes = [vp, P(C.c_double)]
    L.mhapb_xorshift_peaks.argtypes = [vp, P(C.c_double), P(C.c_double)]
    L.mhapb_sketch.argtypes = [vp, P(SketchParams), vp, vp, u32, C.c_int, vp, vp, vp, vp]
    L.mhapb_sketch_device.argtypes = [vp, P(SketchParams), vp, vp, u32, C.c_int, vp, vp, vp, vp]
    L.mhapb_sketch_to_dat.argtypes = [vp, P(SketchParams), vp, vp, vp, u32, C.c_int, P(vp), P(u64), P(u32)]
    L.mhapb_sketch_to_dat_named.argtypes = [vp, P(SketchParams), vp, vp, vp, vp, u32, C.c_int, P(vp), P(u64), P(u32)]
    L.mhapb_dat_encode.argtypes = [i64, C.c_int, C.c_char_p, i32, vp, i32, i32, i32, vp, i32, vp]
    L.mhapb_dat_encode.restype = i64
    L.mhapb_dat_decode.argtypes = [vp, u64, i64, P(u32), P(i32), P(i32), P(i32), vp, vp, vp, vp, vp, vp, vp]
    L.mhapb_store_reset.argtypes = [vp, P(SketchParams)]
    L.mhapb_store_add_reads.argtypes = [vp, vp, vp, vp, u32, C.c_int, P(i64)]
    L.mhapb_store_add_reads_device.argtypes = [vp, vp, vp, vp, u32, C.c_int, P(i64)]
    L.mhapb_store_add_sketches.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, u32]
    L.mhapb_store_add_sketches_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u32]
    L.mhapb_store_size.argtypes = [vp]; L.mhapb_store_size.restype = i64
    L.mhapb_store_reserve.argtypes = [vp, i64]
    L.mhapb_sketch_reserve.argtypes = [vp, P(SketchParams), u64, u32, C.c_int]
    L.mhapb_store_get.argtypes = [vp, i64, P(i64), P(i32), P(i32), P(i32), vp, vp, P(i32)]
    L.mhapb_store_params.argtypes = [vp, P(SketchParams)]
    L.mhapb_store_get_range.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.mhapb_store_device_ptrs.argtypes = [vp, P(vp), P(vp), P(vp), P(i64), P(i32), P(i32)]
    L.mhapb_index_build.argtypes = [vp]
    L.mhapb_search_self.argtypes = [vp, P(SearchParams), P(vp), P(u64), P(Stats)]
    L.mhapb_search_query_reads.argtypes = [vp, P(SearchParams), vp, vp, vp, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_search_query_sketches.argtypes = [vp, P(SearchParams), vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_search_sketches_device.argtypes = [vp, P(SearchParams), C.c_int, vp, vp, vp, vp, vp, vp, vp, i32, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_format_match.argtypes = [P(Hit), C.c_char_p, C.c_size_t]
    L.mhapb_minhash_equal_count.argtypes = [vp, i64, i64, P(i32)]
    L.mhapb_kmer_hash.argtypes = [C.c_char_p, i32, C.c_int, P(i64)]
    L.mhapb_filter_set.argtypes = [vp, P(FilterParams), vp, vp, u64, vp, u64, i32]
    L.mhapb_filter_load_text.argtypes = [vp, P(FilterParams), C.c_char_p, u64, C.c_int, P(i64)]
    L.mhapb_filter_clear.argtypes = [vp]
    L.mhapb_comm_unique_id.argtypes = [vp]
    L.mhapb_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
    L.mhapb_comm_init_all.argtypes = [P(vp), C.c_int]
    L.mhapb_comm_destroy.argtypes = [vp]
    L.mhapb_comm_info.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int)]
    L.mhapb_dist_search_self.argtypes = [vp, P(SearchParams), P(vp), P(u64), P(Stats)]
    L.mhapb_dist_search_query_reads.argtypes = [vp, P(SearchParams), vp, vp, vp, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_dist_search_query_reads_device.argtypes = [vp, P(SearchParams), vp, vp, vp, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_multi_create.argtypes = [P(C.c_int), C.c_int, P(vp)]
    L.mhapb_multi_destroy.argtypes = [vp]; L.mhapb_multi_destroy.restype = None
    L.mhapb_multi_last_error.argtypes = [vp]; L.mhapb_multi_last_error.restype = C.c_char_p
    L.mhapb_multi_n_devices.argtypes = [vp]
    L.mhapb_multi_ctx.argtypes = [vp, C.c_int]; L.mhapb_multi_ctx.restype = vp
    L.mhapb_multi_store_reset.argtypes = [vp, P(SketchParams)]
    L.mhapb_multi_store_reserve.argtypes = [vp, i64]
    L.mhapb_multi_store_add_reads.argtypes = [vp, vp, vp, vp, u32, C.c_int, P(i64)]
    L.mhapb_multi_store_add_sketches.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, u32]
    L.mhapb_multi_store_size.argtypes = [vp]; L.mhapb_multi_store_size.restype = i64
    L.mhapb_multi_search_self.argtypes = [vp, P(SearchParams), P(vp), P(u64), P(Stats)]
    L.mhapb_multi_search_query_reads.argtypes = [vp, P(SearchParams), vp, vp, vp, u32, P(vp), P(u64), P(Stats)]
    L.mhapb_multi_search_query_sketches.argtypes = [vp, P(SearchParams), vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, u32, P(vp), P(u64), P(Stats)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def kmer_hash(kmer, canonical=True) -> int:
    """mhapb_kmer_hash: the key a filter-file k-mer is stored under (HashUtils.computeSequenceHashesLong)."""
    b = kmer if isinstance(kmer, (bytes, bytearray)) else kmer.encode("latin-1")
    out = C.c_int64()
    rc = load().mhapb_kmer_hash(b, len(b), int(canonical), C.byref(out))
    if rc:
        raise MhapError(rc, "mhapb_kmer_hash")
    return int(out.value)


def pack_reads(reads):
    """list[bytes|str] -> (uint8 bases, uint64 offsets[n+1])."""
    bs = [r if isinstance(r, (bytes, bytearray)) else r.encode("latin-1") for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return bases, offsets


class Engine:
    """One mhapb_ctx = one GPU.  Thin, 1:1 over the C ABI."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.mhapb_create(device, C.byref(h))
        if rc:
            raise MhapError(rc, (self.L.mhapb_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.mhapb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise MhapError(rc, (self.L.mhapb_last_error(self.h) or b"").decode())
        return rc

    def timing(self) -> dict:
        t = Timing()
        self._ck(self.L.mhapb_get_timing(self.h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in Timing._fields_}

    def xorshift_peaks(self) -> tuple[float, float]:
        a = C.c_double(); b = C.c_double()
        self._ck(self.L.mhapb_xorshift_peaks(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def xorshift_peak(self) -> float:
        v = C.c_double()
        self._ck(self.L.mhapb_xorshift_peak(self.h, C.byref(v)))
        return v.value

    # ---- the -f k-mer filter (FrequencyCounts) ----
    def filter_load_text(self, text, repeat_weight=0.9, filter_cutoff=1.0e-5, idf_scale=3.0, supress_noise=0, no_tf=False,
                         canonical=True) -> int:
        """Parse the text of a -f filter file and install it; returns the number of repeat k-mers kept."""
        t = text if isinstance(text, (bytes, bytearray)) else text.encode("latin-1")
        fp = FilterParams(filter_cutoff, repeat_weight, idf_scale, supress_noise, int(no_tf))
        n = C.c_int64()
        self._ck(self.L.mhapb_filter_load_text(self.h, C.byref(fp), t, len(t), int(canonical), C.byref(n)))
        return int(n.value)

    def filter_set(self, hashes, fractions, repeat_weight=0.9, filter_cutoff=1.0e-5, idf_scale=3.0, supress_noise=0, no_tf=False,
                   bloom_words=None, bloom_bits=0, bloom_nfun=0):
        h = np.ascontiguousarray(hashes, dtype=np.int64)
        f = np.ascontiguousarray(fractions, dtype=np.float64)
        bw = None if bloom_words is None else np.ascontiguousarray(bloom_words, dtype=np.uint64)
        fp = FilterParams(filter_cutoff, repeat_weight, idf_scale, supress_noise, int(no_tf))
        self._ck(self.L.mhapb_filter_set(self.h, C.byref(fp), _ptr(h) if h.size else None, _ptr(f) if f.size else None, h.size,
                                         _ptr(bw), bloom_bits, bloom_nfun))

    def filter_clear(self):
        self._ck(self.L.mhapb_filter_clear(self.h))

    # ---- K1 ----
    def sketch(self, bases, offsets, params: SketchParams, both_strands=True, want_ord=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        per = 2 if both_strands else 1
        H, S = params.num_hashes, params.ordered_sketch_size
        mh = np.empty((n * per, H), dtype=np.int32)
        od = np.empty((n * per, S, 2), dtype=np.int32) if want_ord else None
        on = np.empty(n * per, dtype=np.int32) if want_ord else None
        st = np.empty(n, dtype=np.int32)
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_sketch(self.h, C.byref(params), _ptr(b), _ptr(offsets), n, int(both_strands),
                                     _ptr(mh), _ptr(od), _ptr(on), _ptr(st)))
        return mh, od, on, st

    def sketch_device(self, d_bases: int, offsets, params: SketchParams, both_strands, d_minhash: int, d_ord: int,
                      d_ord_n: int, d_status: int = 0):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        self._ck(self.L.mhapb_sketch_device(self.h, C.byref(params), d_bases, _ptr(offsets), n, int(both_strands),
                                            d_minhash or None, d_ord or None, d_ord_n or None, d_status or None))

    def sketch_to_dat(self, bases, offsets, ids, params: SketchParams, both_strands=True) -> tuple[bytes, int]:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        out = C.c_void_p(); ln = C.c_uint64(); nrec = C.c_uint32()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_sketch_to_dat(self.h, C.byref(params), _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1,
                                            int(both_strands), C.byref(out), C.byref(ln), C.byref(nrec)))
        data = C.string_at(out.value, ln.value)
        self.L.mhapb_free(out)
        return data, nrec.value

    # ---- store / index ----
    def store_reset(self, params: SketchParams):
        self._ck(self.L.mhapb_store_reset(self.h, C.byref(params)))

    def store_add_reads(self, bases, offsets, ids=None, both_strands=True) -> int:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        added = C.c_int64()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_store_add_reads(self.h, _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1,
                                              int(both_strands), C.byref(added)))
        return added.value

    def store_add_reads_device(self, d_bases: int, offsets, ids=None, both_strands=True) -> int:
        """Reads already resident in HBM at device address d_bases (mhapb_store_add_reads_device)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        added = C.c_int64()
        self._ck(self.L.mhapb_store_add_reads_device(self.h, d_bases, _ptr(offsets), _ptr(ids), offsets.size - 1, int(both_strands), C.byref(added)))
        return added.value

    def store_add_sketches(self, ids, is_fwd, seq_len, seq_len_kmers, minhash, ord_hp, ord_n, ordered_kmer_size=12):
        ids = np.ascontiguousarray(ids, dtype=np.int64); is_fwd = np.ascontiguousarray(is_fwd, dtype=np.uint8)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32); slk = np.ascontiguousarray(seq_len_kmers, dtype=np.int32)
        minhash = np.ascontiguousarray(minhash, dtype=np.int32); ord_hp = np.ascontiguousarray(ord_hp, dtype=np.int32)
        ord_n = np.ascontiguousarray(ord_n, dtype=np.int32)
        self._ck(self.L.mhapb_store_add_sketches(self.h, _ptr(ids), _ptr(is_fwd), _ptr(seq_len), _ptr(slk), _ptr(minhash),
                                                 minhash.shape[1] if minhash.ndim == 2 else 0, _ptr(ord_hp), _ptr(ord_n), ord_hp.shape[1],
                                                 ordered_kmer_size, ids.size))

    def store_add_sketches_device(self, ids, is_fwd, seq_len, seq_len_kmers, d_minhash: int, d_ord: int, ord_n):
        ids = np.ascontiguousarray(ids, dtype=np.int64); is_fwd = np.ascontiguousarray(is_fwd, dtype=np.uint8)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32); slk = np.ascontiguousarray(seq_len_kmers, dtype=np.int32)
        ord_n = np.ascontiguousarray(ord_n, dtype=np.int32)
        self._ck(self.L.mhapb_store_add_sketches_device(self.h, _ptr(ids), _ptr(is_fwd), _ptr(seq_len), _ptr(slk),
                                                        d_minhash, d_ord, _ptr(ord_n), ids.size))

    def store_size(self) -> int:
        return int(self.L.mhapb_store_size(self.h))

    def store_reserve(self, n_sketches: int):
        self._ck(self.L.mhapb_store_reserve(self.h, int(n_sketches)))

    def store_get(self, idx: int, H: int, S: int) -> dict:
        id_ = C.c_int64(); fwd = C.c_int32(); sl = C.c_int32(); slk = C.c_int32(); on = C.c_int32()
        mh = np.zeros(H, dtype=np.int32); od = np.zeros((S, 2), dtype=np.int32)
        self._ck(self.L.mhapb_store_get(self.h, idx, C.byref(id_), C.byref(fwd), C.byref(sl), C.byref(slk), _ptr(mh), _ptr(od), C.byref(on)))
        return dict(id=id_.value, is_fwd=bool(fwd.value), seq_len=sl.value, seq_len_kmers=slk.value, minhash=mh, ord=od[:on.value].copy())

    def store_get_range(self, first: int, count: int, want_ord=True) -> dict:
        """Stored sketches [first, first+count) as flat arrays (mhapb_store_get_range)."""
        _, _, _, _, H, S = self.store_device_ptrs()
        out = dict(ids=np.zeros(count, np.int64), is_fwd=np.zeros(count, np.uint8), seq_len=np.zeros(count, np.int32),
                   seq_len_kmers=np.zeros(count, np.int32), minhash=np.zeros((count, H), np.int32),
                   ord=np.zeros((count, S, 2), np.int32) if want_ord else None, ord_n=np.zeros(count, np.int32))
        self._ck(self.L.mhapb_store_get_range(self.h, first, count, _ptr(out["ids"]), _ptr(out["is_fwd"]), _ptr(out["seq_len"]),
                                              _ptr(out["seq_len_kmers"]), _ptr(out["minhash"]), _ptr(out["ord"]), _ptr(out["ord_n"])))
        return out

    def store_device_ptrs(self):
        a = C.c_void_p(); b = C.c_void_p(); c = C.c_void_p(); n = C.c_int64(); H = C.c_int32(); S = C.c_int32()
        self._ck(self.L.mhapb_store_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(n), C.byref(H), C.byref(S)))
        return a.value, b.value, c.value, n.value, H.value, S.value

    def index_build(self):
        self._ck(self.L.mhapb_index_build(self.h))

    # ---- search ----
    def _collect(self, out, n, st):
        if n.value:
            raw = (C.c_char * (n.value * C.sizeof(Hit))).from_address(out.value)
            hits = np.frombuffer(raw, dtype=HIT_DTYPE).copy()      # one copy out of the library's buffer
        else:
            hits = np.zeros(0, dtype=HIT_DTYPE)
        self.L.mhapb_free(out)
        return hits, {f: int(getattr(st, f)) for f, _ in Stats._fields_}

    def search_self(self, sp: SearchParams):
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_search_self(self.h, C.byref(sp), C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def search_query_reads(self, sp: SearchParams, bases, offsets, ids=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_search_query_reads(self.h, C.byref(sp), _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1,
                                                 C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def search_query_sketches(self, sp: SearchParams, ids, is_fwd, seq_len, seq_len_kmers, minhash, ord_hp, ord_n, ordered_kmer_size=12):
        ids = np.ascontiguousarray(ids, dtype=np.int64); is_fwd = np.ascontiguousarray(is_fwd, dtype=np.uint8)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32); slk = np.ascontiguousarray(seq_len_kmers, dtype=np.int32)
        minhash = np.ascontiguousarray(minhash, dtype=np.int32); ord_hp = np.ascontiguousarray(ord_hp, dtype=np.int32)
        ord_n = np.ascontiguousarray(ord_n, dtype=np.int32)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_search_query_sketches(self.h, C.byref(sp), _ptr(ids), _ptr(is_fwd), _ptr(seq_len), _ptr(slk),
                                                    _ptr(minhash), minhash.shape[1] if minhash.ndim == 2 else 0, _ptr(ord_hp), _ptr(ord_n),
                                                    ord_hp.shape[1], ordered_kmer_size, ids.size, C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def search_sketches_device(self, sp: SearchParams, to_self: bool, ids, is_fwd, seq_len, seq_len_kmers, d_minhash: int, d_ord: int,
                               d_ord_n: int, ord_stride: int):
        ids = np.ascontiguousarray(ids, dtype=np.int64); is_fwd = np.ascontiguousarray(is_fwd, dtype=np.uint8)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32); slk = np.ascontiguousarray(seq_len_kmers, dtype=np.int32)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_search_sketches_device(self.h, C.byref(sp), int(to_self), _ptr(ids), _ptr(is_fwd), _ptr(seq_len), _ptr(slk),
                                                     d_minhash, d_ord, d_ord_n, ord_stride, ids.size, C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    # ---- multi-GPU: one context per rank, NCCL inside the library ----
    def comm_init_rank(self, comm_id: bytes, rank: int, nranks: int):
        """Join the job's communicator (collective).  comm_id = comm_unique_id() of rank 0, handed over by the launcher."""
        assert len(comm_id) == COMM_ID_BYTES
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._ck(self.L.mhapb_comm_init_rank(self.h, buf, rank, nranks))

    def comm_info(self):
        r = C.c_int(); n = C.c_int(); v = C.c_int()
        self._ck(self.L.mhapb_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v)))
        return r.value, n.value, v.value

    def dist_search_self(self, sp: SearchParams):
        """COLLECTIVE: hits whose target this rank stores + job-wide counters (mhapb_dist_search_self)."""
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_dist_search_self(self.h, C.byref(sp), C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def dist_search_query_reads(self, sp: SearchParams, bases, offsets, ids=None):
        """COLLECTIVE: this rank's shard of the query reads; all queries meet every store shard."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_dist_search_query_reads(self.h, C.byref(sp), _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1,
                                                      C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def dist_search_query_reads_device(self, sp: SearchParams, d_bases: int, offsets, ids=None):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_dist_search_query_reads_device(self.h, C.byref(sp), d_bases, _ptr(offsets), _ptr(ids), offsets.size - 1,
                                                             C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def minhash_equal_count(self, i: int, j: int) -> int:
        out = C.c_int32()
        self._ck(self.L.mhapb_minhash_equal_count(self.h, i, j, C.byref(out)))
        return out.value


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it; the launcher distributes the bytes)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = load().mhapb_comm_unique_id(buf)
    if rc:
        raise MhapError(rc, (load().mhapb_last_error(None) or b"").decode())
    return bytes(buf)


class MultiEngine:
    """mhapb_multi: one process driving several GPUs (what a single JVM would bind)."""

    def __init__(self, devices):
        self.L = load()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.L.mhapb_multi_create(devs, len(devices), C.byref(h))
        if rc:
            raise MhapError(rc, (self.L.mhapb_last_error(None) or b"").decode())
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.L.mhapb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise MhapError(rc, (self.L.mhapb_multi_last_error(self.h) or b"").decode())
        return rc

    def store_reset(self, params: SketchParams):
        self._ck(self.L.mhapb_multi_store_reset(self.h, C.byref(params)))

    def store_add_reads(self, bases, offsets, ids=None, both_strands=True) -> int:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        added = C.c_int64()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_multi_store_add_reads(self.h, _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1, int(both_strands), C.byref(added)))
        return added.value

    def store_size(self) -> int:
        return int(self.L.mhapb_multi_store_size(self.h))

    def _collect(self, out, n, st):
        if n.value:
            raw = (C.c_char * (n.value * C.sizeof(Hit))).from_address(out.value)
            hits = np.frombuffer(raw, dtype=HIT_DTYPE).copy()
        else:
            hits = np.zeros(0, dtype=HIT_DTYPE)
        self.L.mhapb_free(out)
        return hits, {f: int(getattr(st, f)) for f, _ in Stats._fields_}

    def search_self(self, sp: SearchParams):
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        self._ck(self.L.mhapb_multi_search_self(self.h, C.byref(sp), C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)

    def search_query_reads(self, sp: SearchParams, bases, offsets, ids=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        out = C.c_void_p(); n = C.c_uint64(); st = Stats()
        b = bases if bases.size else np.zeros(1, dtype=np.uint8)
        self._ck(self.L.mhapb_multi_search_query_reads(self.h, C.byref(sp), _ptr(b), _ptr(offsets), _ptr(ids), offsets.size - 1,
                                                       C.byref(out), C.byref(n), C.byref(st)))
        return self._collect(out, n, st)


def format_match(hit_row) -> str:
    """impl/MatchResult.java:98-113 through the library (one line, no newline)."""
    h = Hit()
    for f, _ in Hit._fields_:
        v = hit_row[f]
        setattr(h, f, v.item() if hasattr(v, "item") else v)
    buf = C.create_string_buffer(256)
    load().mhapb_format_match(C.byref(h), buf, 256)
    return buf.value.decode()


def dat_encode(id_, is_fwd, seq_len, minhash, seq_len_kmers, ordered_k, ord_hp, header=None) -> bytes:
    mh = np.ascontiguousarray(minhash, dtype=np.int32)
    oh = np.ascontiguousarray(ord_hp, dtype=np.int32).reshape(-1, 2)
    hdr = None if header is None else (header if isinstance(header, bytes) else header.encode())
    L = load()
    n = L.mhapb_dat_encode(id_, int(is_fwd), hdr, seq_len, _ptr(mh), mh.size, seq_len_kmers, ordered_k, _ptr(oh), oh.shape[0], None)
    if n < 0:
        raise MhapError(n, "dat_encode")
    buf = C.create_string_buffer(n)
    L.mhapb_dat_encode(id_, int(is_fwd), hdr, seq_len, _ptr(mh), mh.size, seq_len_kmers, ordered_k, _ptr(oh), oh.shape[0], buf)
    return buf.raw


def dat_decode(data: bytes, id_offset: int = 0) -> dict:
    """SequenceSketchStreamer.java:278-320 + SequenceSketch.fromByteStream: .dat bytes -> flat arrays."""
    L = load()
    buf = np.frombuffer(data, dtype=np.uint8)
    n = C.c_uint32(); H = C.c_int32(); mo = C.c_int32(); ok = C.c_int32()
    rc = L.mhapb_dat_decode(_ptr(buf) if buf.size else None, buf.size, id_offset, C.byref(n), C.byref(H), C.byref(mo), C.byref(ok),
                            None, None, None, None, None, None, None)
    if rc:
        raise MhapError(rc, "corrupt .dat stream")
    nr, h, s = n.value, H.value, max(mo.value, 1)
    out = dict(ids=np.zeros(nr, np.int64), is_fwd=np.zeros(nr, np.uint8), seq_len=np.zeros(nr, np.int32),
               seq_len_kmers=np.zeros(nr, np.int32), minhash=np.zeros((nr, h), np.int32), ord=np.zeros((nr, s, 2), np.int32),
               ord_n=np.zeros(nr, np.int32), ordered_kmer_size=ok.value, num_hashes=h)
    mo2 = C.c_int32(s)
    rc = L.mhapb_dat_decode(_ptr(buf) if buf.size else None, buf.size, id_offset, C.byref(n), C.byref(H), C.byref(mo2), C.byref(ok),
                            _ptr(out["ids"]), _ptr(out["is_fwd"]), _ptr(out["seq_len"]), _ptr(out["seq_len_kmers"]),
                            _ptr(out["minhash"]), _ptr(out["ord"]), _ptr(out["ord_n"]))
    if rc:
        raise MhapError(rc, "corrupt .dat stream")
    return out
