"""Second, independent restatement of the MHAP path in pure Python (slow; tiny inputs only).

TEST INFRASTRUCTURE ONLY.  Written from the Java source independently of mhap_oracle.c so a
transcription slip in either cannot hide: tests assert the two agree on random inputs that
include N's, IUPAC letters, tandem repeats and duplicate ordered hashes.  PARITY UNPINNED by the
reference (no JVM here, no reference tests).  Citations relative to
/root/reference/src/main/java/edu/umd/marbl/mhap/ .
"""
from __future__ import annotations

import math

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def _s64(x):
    x &= M64
    return x - (1 << 64) if x >> 63 else x


def _s32(x):
    x &= M32
    return x - (1 << 32) if x >> 31 else x


def _rotl64(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _rotl32(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _fmix64(k):
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & M64
    k ^= k >> 33
    return k


def murmur3_x64_128(data: bytes, seed: int = 0):
    """Guava Hashing.murmur3_128(seed) == MurmurHash3_x64_128."""
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    h1 = h2 = seed & M32
    n = len(data)
    nb = n // 16
    for i in range(nb):
        k1 = int.from_bytes(data[16 * i:16 * i + 8], "little")
        k2 = int.from_bytes(data[16 * i + 8:16 * i + 16], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl64(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
        h1 = _rotl64(h1, 27); h1 = (h1 + h2) & M64; h1 = (h1 * 5 + 0x52DCE729) & M64
        k2 = (k2 * c2) & M64; k2 = _rotl64(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
        h2 = _rotl64(h2, 31); h2 = (h2 + h1) & M64; h2 = (h2 * 5 + 0x38495AB5) & M64
    tail = data[16 * nb:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & M64; k2 = _rotl64(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl64(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
    h1 ^= n; h2 ^= n
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    h1 = _fmix64(h1); h2 = _fmix64(h2)
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    return h1, h2


def murmur3_x86_32(data: bytes, seed: int = 0) -> int:
    """Guava Hashing.murmur3_32(seed) == MurmurHash3_x86_32 (unsigned result)."""
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & M32
    n = len(data)
    nb = n // 4
    for i in range(nb):
        k = int.from_bytes(data[4 * i:4 * i + 4], "little")
        k = (k * c1) & M32; k = _rotl32(k, 15); k = (k * c2) & M32
        h ^= k; h = _rotl32(h, 13); h = (h * 5 + 0xE6546B64) & M32
    tail = data[4 * nb:]
    if tail:
        k = int.from_bytes(tail, "little")
        k = (k * c1) & M32; k = _rotl32(k, 15); k = (k * c2) & M32
        h ^= k
    h ^= n
    h ^= h >> 16; h = (h * 0x85EBCA6B) & M32; h ^= h >> 13; h = (h * 0xC2B2AE35) & M32; h ^= h >> 16
    return h


_COMP = dict(zip("ABCDGHKMNRSTVWY", "TVGHCDMKNYSABWR"))  # utils/Utils.java:84-114


def rc(s: str) -> str:  # utils/Utils.java:496-507
    return "".join(_COMP.get(c.upper(), c.upper()) for c in reversed(s))


def _utf16(s: str) -> bytes:  # Hasher.putUnencodedChars
    return s.encode("utf-16-le")


def kmer_hashes_long(seq: str, k: int):  # sketch/HashUtils.java:237-258
    return [_s64(murmur3_x64_128(_utf16(seq[i:i + k]), 0)[0]) for i in range(len(seq) - k + 1)]


def kmer_hashes_int(seq: str, k: int):  # sketch/HashUtils.java:213-235
    return [_s32(murmur3_x86_32(_utf16(seq[i:i + k]), 0)) for i in range(len(seq) - k + 1)]


def minhash_sketch(seq: str, k: int, H: int, unweighted: bool = False):  # sketch/MinHashSketch.java:51-179
    if len(seq) - k + 1 < 1:
        return None
    counts = {}  # dict keeps insertion order == Long2ObjectLinkedOpenHashMap
    for h in kmer_hashes_long(seq, k):
        counts[h] = counts.get(h, 0) + 1
    best = [(1 << 63) - 1] * H
    out = [0] * max(1, H)
    for key, cnt in counts.items():
        w = 1 if unweighted else cnt
        x = key & M64
        for word in range(H):
            for _ in range(w):
                x ^= (x << 21) & M64
                x ^= x >> 35
                x ^= (x << 4) & M64
                sx = _s64(x)
                if sx < best[word]:
                    best[word] = sx
                    out[word] = _s32(key & M32) if word % 2 == 0 else _s32((key & M64) >> 32)
    return out


class FrequencyCounts:  # sketch/FrequencyCounts.java:63-320
    """The -f k-mer filter, written from the Java source independently of mhap_oracle.c.

    The Bloom filter follows Guava 19.0's published BloomFilter.create / BloomFilterStrategies.MURMUR128_MITZ_64."""

    def __init__(self, text: str, filter_cutoff=1.0e-5, offset=0.0, remove_unique=0, no_tf=False, range_=3.0, canonical=True):
        self.cutoff, self.offset, self.remove_unique, self.no_tf, self.range = filter_cutoff, offset, remove_unique, no_tf, range_
        lines = text.split("\n")
        first = lines[0].strip().split() if lines and lines[0].strip() else []
        size_bloom = int(first[0]) if first else 1
        if size_bloom == 0:
            size_bloom = 1
        self.fraction = {}
        self.max_value = -math.inf
        self.bits = None
        if remove_unique > 0:
            nbits = int(-size_bloom * math.log(1.0e-5) / (math.log(2) * math.log(2)))   # optimalNumOfBits
            self.nfun = max(1, _jround(nbits / size_bloom * math.log(2)))                 # optimalNumOfHashFunctions
            self.bit_size = max(1, -(-nbits // 64)) * 64
            self.bits = bytearray(self.bit_size // 8)
        for line in lines[1:]:
            tok = line.split(None, 2) if not line[:1].isspace() else [""]   # Java split keeps a leading empty token
            if not tok or not tok[0]:
                continue
            kmer = tok[0]
            if canonical:                                                  # HashUtils.java:246-251
                r = rc(kmer)                                               # Utils.rc upper-cases
                if r < kmer:
                    kmer = r
            h = _s64(murmur3_x64_128(_utf16(kmer), 0)[0])
            if len(tok) >= 2:
                try:
                    pct = float(tok[1])
                except ValueError:
                    continue
                if pct >= filter_cutoff:
                    self.max_value = max(self.max_value, pct)
                    self.fraction[h] = pct
            if self.bits is not None:
                for bit in self._bloom_bits(h):
                    self.bits[bit >> 3] |= 1 << (bit & 7)
        self.min_value = filter_cutoff
        self.min_idf = self._idf(self.max_value)
        self.max_idf = self._idf(self.min_value)

    def _idf(self, freq):
        try:
            return math.log(self.max_value / freq - self.offset)
        except (ValueError, ZeroDivisionError):
            return math.nan

    def _bloom_bits(self, h):
        h1, h2 = murmur3_x64_128((h & M64).to_bytes(8, "little"), 0)
        c = h1
        for _ in range(self.nfun):
            yield (c & ((1 << 63) - 1)) % self.bit_size
            c = (c + h2) & M64

    def might_contain(self, h):
        return all(self.bits[b >> 3] >> (b & 7) & 1 for b in self._bloom_bits(h))

    def keep_kmer(self, h):
        return self.might_contain(h) if self.remove_unique == 1 else True

    def is_popular(self, h):
        return h in self.fraction

    def scaled_idf(self, h):
        if self.remove_unique == 2 and self.bits is not None and not self.might_contain(h):
            return 1.0
        if h not in self.fraction:
            return self.range
        idf = self._idf(self.fraction[h])
        d = self.range - 1.0
        scale = (self.max_idf - self.min_idf) / d if d != 0.0 else math.inf
        return 1.0 + (idf - self.min_idf) / scale if scale != 0.0 else math.nan


def minhash_sketch_filtered(seq: str, k: int, H: int, repeat_weight: float, fc: "FrequencyCounts | None"):
    """sketch/MinHashSketch.java:51-179 with kmerFilter != null paths."""
    if len(seq) - k + 1 < 1:
        return None
    counts = {}
    for h in kmer_hashes_long(seq, k):
        if fc is not None and not fc.keep_kmer(h):
            continue
        counts[h] = counts.get(h, 0) + 1
    if not counts:
        return None
    best = [(1 << 63) - 1] * H
    out = [0] * max(1, H)
    valid = 0
    for key, cnt in counts.items():
        w = cnt
        if repeat_weight < 0.0:
            w = 0 if (fc is not None and fc.is_popular(key)) else 1
        elif fc is not None and 0.0 <= repeat_weight < 1.0:
            tf = 1.0 if fc.no_tf else float(cnt)
            v = tf * fc.scaled_idf(key)
            w = 0 if v != v else _jround(v)
            if w < 1:
                w = 1
        if w <= 0:
            continue
        valid += 1
        x = key & M64
        for word in range(H):
            for _ in range(w):
                x ^= (x << 21) & M64
                x ^= x >> 35
                x ^= (x << 4) & M64
                sx = _s64(x)
                if sx < best[word]:
                    best[word] = sx
                    out[word] = _s32(key & M32) if word % 2 == 0 else _s32((key & M64) >> 32)
    return out if valid > 0 else None


def bottom_sketch(seq: str, ok: int, size: int):  # sketch/BottomOverlapSketch.java:525-559
    n = len(seq) - ok + 1
    if n <= 0:
        return None, n
    hs = kmer_hashes_int(seq, ok)
    order = sorted(range(n), key=lambda i: (hs[i], i))  # signed ascending, stable
    return [(hs[i], i) for i in order[:min(size, n)]], n


def _jround(x: float) -> int:  # Java 8 Math.round
    f = math.floor(x)
    return int(f) + (1 if x - f >= 0.5 else 0)


def _median_absmax(shifts, len1, len2, max_shift):  # MatchData.performUpdate :191-215
    if shifts:
        med = sorted(shifts)[len(shifts) // 2]  # == Utils.quickSelect(copy, count/2, count)
        left = max(0, -med)
        right = min(len1, len2 - med)
        ov = max(10, right - left)
        return med, min(max(len1, len2), int(float(ov) * max_shift))
    return 0, max(len1, len2) + 1


def _record(A, B, len1, len2, med, amax):  # recordMatchingKmers :397-516
    v1lo, v1hi = max(0, -med - amax), min(len1, len2 - med + amax)
    v2lo, v2hi = max(0, med - amax), min(len2, len1 + med + amax)
    rec = []
    i = j = 0
    while i < len(A) and j < len(B):
        h1, p1 = A[i]
        h2, p2 = B[j]
        if h1 < h2 or p1 < v1lo or p1 >= v1hi:
            i += 1
        elif h2 < h1 or p2 < v2lo or p2 >= v2hi:
            j += 1
        else:
            d = (p2 - p1) - med
            if d > amax:
                i += 1
            elif d < -amax:
                j += 1
            else:
                rec.append((p1, p2, p2 - p1))
                il = i
                while il + 1 < len(A) and A[il + 1][0] == h1 and v1lo <= A[il + 1][1] < v1hi:
                    il += 1
                jl = j
                while jl + 1 < len(B) and B[jl + 1][0] == h2 and v2lo <= B[jl + 1][1] < v2hi:
                    jl += 1
                if il != i or jl != j:
                    rec.append((A[il][1], B[jl][1], B[jl][1] - A[il][1]))
                    i, j = il + 1, jl + 1
                else:
                    i += 1
                    j += 1
    return rec


def overlap_info(A, len1, B, len2, ok=12, max_shift=0.2):  # getOverlapInfo :592-630
    """A, B: lists of (hash,pos).  Returns None for OverlapInfo.EMPTY, else dict."""
    med, amax = _median_absmax([], len1, len2, max_shift)
    rec = _record(A, B, len1, len2, med, amax)
    if not rec:
        return None
    med, amax = _median_absmax([r[2] for r in rec], len1, len2, max_shift)
    rec = _record(A, B, len1, len2, med, amax)
    if not rec:
        return None
    med, amax = _median_absmax([r[2] for r in rec], len1, len2, max_shift)
    red = []  # optimizeShifts :156-189
    for r in rec:
        if red and red[-1][0] == r[0]:
            if abs(red[-1][2] - med) > abs(r[2] - med):
                red[-1] = r
        else:
            red.append(r)
    med, amax = _median_absmax([r[2] for r in red], len1, len2, max_shift)
    ok_rec = [r for r in red if abs(r[2] - med) <= amax]  # computeEdges :90-137
    n = len(ok_rec)
    if n < 3:
        return None
    l1, r1 = min(r[0] for r in ok_rec), max(r[0] for r in ok_rec)
    l2, r2 = min(r[1] for r in ok_rec), max(r[1] for r in ok_rec)
    a1 = max(0, _jround(float(n * l1 - r1) / float(n - 1)))
    a2 = min(len1, _jround(float(n * r1 - l1) / float(n - 1)))
    b1 = max(0, _jround(float(n * l2 - r2) / float(n - 1)))
    b2 = min(len2, _jround(float(n * r2 - l2) / float(n - 1)))
    h1 = [h for h, p in A if a1 <= p <= a2]  # computeKBottomSketchJaccard :304-364
    h2 = [h for h, p in B if b1 <= p <= b2]
    k = min(len(h1), len(h2))
    inter = 0
    jac = 0.0
    if k:
        i = j = u = 0
        while u < k:
            if h1[i] < h2[j]:
                i += 1
            elif h1[i] > h2[j]:
                j += 1
            else:
                inter += 1
                i += 1
                j += 1
            u += 1
        jac = inter / k
    if jac > 0:
        score = math.exp(-(-1.0 / float(ok) * math.log(2.0 * jac / (1.0 + jac))))  # :391-395
    else:
        score = 0.0
    return dict(a1=a1, a2=a2, b1=b1, b2=b2, valid_count=n, intersect=inter, kmin=k, score=score)


def sketch_reads(reads, k=16, H=512, ok=12, S=1536, unweighted=False, both=True, min_olap=116, id_offset=0):
    """impl/SequenceSketchStreamer.java:123-156 -> list of sketch dicts (read i fwd, read i rev)."""
    out = []
    for i, r in enumerate(reads):
        r = r.upper()
        if len(r) < min_olap:
            continue
        for fwd in ((True, False) if both else (True,)):
            s = r if fwd else rc(r)
            mh = minhash_sketch(s, k, H, unweighted)
            od, n = bottom_sketch(s, ok, S)
            if mh is None or od is None:
                break
            out.append(dict(id=i + 1 + id_offset, is_fwd=fwd, seq_len=len(s), minhash=mh, ord=od, seq_len_kmers=n))
    return out


def search(store, queries, to_self, m=3, min_store=0, max_shift=0.2, accept=0.78, ok=12):
    """impl/MinHashSearch.java:150-251 over every forward query.  Returns (hits, stats)."""
    H = len(store[0]["minhash"]) if store else 0
    index = {}
    for si, s in enumerate(store):  # addSequence :101-147
        for w in range(H):
            index.setdefault((w, s["minhash"][w]), []).append(si)
    hits = []
    st = dict(elements_processed=0, sequences_hit=0, fully_compared=0, matches_processed=0, sequences_searched=0)
    for q in queries:
        if not q["is_fwd"]:
            continue
        st["sequences_searched"] += 1
        cnt = {}
        for w in range(H):
            lst = index.get((w, q["minhash"][w]))
            if lst:
                st["elements_processed"] += len(lst)
                for si in lst:
                    cnt[si] = cnt.get(si, 0) + 1
        st["sequences_hit"] += len(cnt)
        for si, c in cnt.items():
            t = store[si]
            if to_self and t["id"] == q["id"]:
                continue
            if c < m:
                continue
            if t["seq_len"] < min_store and q["seq_len"] < min_store:
                continue
            if to_self and t["id"] > q["id"] and t["seq_len"] >= min_store and q["seq_len"] >= min_store:
                continue
            if to_self and t["seq_len"] < min_store and q["seq_len"] >= min_store:
                continue
            ov = overlap_info(q["ord"], q["seq_len_kmers"], t["ord"], t["seq_len_kmers"], ok, max_shift)
            st["fully_compared"] += 1
            score = ov["score"] if ov else 0.0
            acc = score >= accept
            if acc:
                st["matches_processed"] += 1
            hits.append(dict(from_id=q["id"], to_id=t["id"], from_fwd=q["is_fwd"], to_fwd=t["is_fwd"], hit_count=c,
                             ov=ov, score=score, accepted=acc, from_len=q["seq_len"], to_len=t["seq_len"]))
    return hits, st
