/*
 * mhap_oracle.c -- CPU restatement of MHAP 2.1.3's sketch + overlap-search path.
 *
 * TEST INFRASTRUCTURE ONLY (see mhap_oracle.h).  PARITY UNPINNED by the reference's own tests
 * (it has none); pinned by public MurmurHash3 known-answer vectors and cross-checked against an
 * independent pure-Python restatement (oracle/pyref.py).
 *
 * Citations are relative to /root/reference/src/main/java/edu/umd/marbl/mhap/ .
 * Java semantics kept on purpose: signed 64/32-bit compares, >>> as unsigned shift,
 * (int) casts truncating toward zero, Math.round = round-half-up.
 */
#define _GNU_SOURCE
#include "mhap_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================== */
/* MurmurHash3 (Austin Appleby, public domain algorithm) as used by Guava 19.0              */
/* ======================================================================================== */

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t le64(const uint8_t *p)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}

/* Guava Hashing.murmur3_128(seed): MurmurHash3_x64_128, h1 = h2 = seed; HashCode.asLong() = h1. */
void mo_murmur3_x64_128(const uint8_t *data, size_t len, uint32_t seed, uint64_t out[2])
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    size_t nblocks = len / 16;
    for (size_t i = 0; i < nblocks; i++) {
        uint64_t k1 = le64(data + 16 * i), k2 = le64(data + 16 * i + 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    switch (len & 15) {
    case 15: k2 ^= (uint64_t)tail[14] << 48; /* fallthrough */
    case 14: k2 ^= (uint64_t)tail[13] << 40; /* fallthrough */
    case 13: k2 ^= (uint64_t)tail[12] << 32; /* fallthrough */
    case 12: k2 ^= (uint64_t)tail[11] << 24; /* fallthrough */
    case 11: k2 ^= (uint64_t)tail[10] << 16; /* fallthrough */
    case 10: k2 ^= (uint64_t)tail[9] << 8;   /* fallthrough */
    case 9:  k2 ^= (uint64_t)tail[8];
             k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= (uint64_t)tail[7] << 56;  /* fallthrough */
    case 7:  k1 ^= (uint64_t)tail[6] << 48;  /* fallthrough */
    case 6:  k1 ^= (uint64_t)tail[5] << 40;  /* fallthrough */
    case 5:  k1 ^= (uint64_t)tail[4] << 32;  /* fallthrough */
    case 4:  k1 ^= (uint64_t)tail[3] << 24;  /* fallthrough */
    case 3:  k1 ^= (uint64_t)tail[2] << 16;  /* fallthrough */
    case 2:  k1 ^= (uint64_t)tail[1] << 8;   /* fallthrough */
    case 1:  k1 ^= (uint64_t)tail[0];
             k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* Guava Hashing.murmur3_32(seed): MurmurHash3_x86_32; HashCode.asInt() = h. */
uint32_t mo_murmur3_x86_32(const uint8_t *data, size_t len, uint32_t seed)
{
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    size_t nblocks = len / 4;
    for (size_t i = 0; i < nblocks; i++) {
        const uint8_t *p = data + 4 * i;
        uint32_t k = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        k *= c1; k = rotl32(k, 15); k *= c2;
        h ^= k; h = rotl32(h, 13); h = h * 5 + 0xe6546b64u;
    }
    const uint8_t *tail = data + nblocks * 4;
    uint32_t k = 0;
    switch (len & 3) {
    case 3: k ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1: k ^= tail[0];
            k *= c1; k = rotl32(k, 15); k *= c2; h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

/* ======================================================================================== */
/* utils/Utils.java                                                                          */
/* ======================================================================================== */

/* Utils.java:84-114 Translate.lookup; unknown characters are returned unchanged (upper-cased,
 * Utils.java:503). */
static inline char translate(char c)
{
    if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
    switch (c) {
    case 'A': return 'T'; case 'B': return 'V'; case 'C': return 'G'; case 'D': return 'H';
    case 'G': return 'C'; case 'H': return 'D'; case 'K': return 'M'; case 'M': return 'K';
    case 'N': return 'N'; case 'R': return 'Y'; case 'S': return 'S'; case 'T': return 'A';
    case 'V': return 'B'; case 'W': return 'W'; case 'Y': return 'R';
    default:  return c;
    }
}

/* Utils.java:496-507 */
void mo_rc(const char *seq, int64_t len, char *out)
{
    for (int64_t i = 0; i < len; i++) out[i] = translate(seq[len - 1 - i]);
}

/* Utils.java:445-494, literal (including the in-place partition). */
int32_t mo_quick_select(int32_t *array, int32_t k, int32_t length)
{
    if (array == NULL || length <= k) return INT32_MAX;
    int32_t from = 0, to = length - 1;
    while (from < to) {
        int32_t r = from, w = to;
        int32_t mid = array[(r + w) / 2];
        while (r < w) {
            if (array[r] >= mid) {
                int32_t tmp = array[w];
                array[w] = array[r];
                array[r] = tmp;
                w--;
            } else {
                r++;
            }
        }
        if (array[r] > mid) r--;
        if (k <= r) to = r; else from = r + 1;
    }
    return array[k];
}

/* ======================================================================================== */
/* sketch/HashUtils.java                                                                      */
/* ======================================================================================== */

/* Hasher.putUnencodedChars: each Java char -> 2 bytes, little-endian.  Input restricted to
 * single-byte characters (ASCII / Latin-1), so the high byte is always 0. */
static inline void utf16le(const char *s, int k, uint8_t *out)
{
    for (int i = 0; i < k; i++) { out[2 * i] = (uint8_t)s[i]; out[2 * i + 1] = 0; }
}

/* String.compareTo on single-byte chars == unsigned byte lexicographic order. */
static const char *canonical_kmer(const char *kmer, int k, char *scratch)
{
    mo_rc(kmer, k, scratch);
    for (int i = 0; i < k; i++) {
        unsigned char a = (unsigned char)scratch[i], b = (unsigned char)kmer[i];
        if (a != b) return a < b ? scratch : kmer;
    }
    return kmer;
}

/* HashUtils.java:237-258 */
int64_t mo_kmer_hashes_long(const char *seq, int64_t len, int k, uint32_t seed, int canonical,
                            int64_t *out)
{
    int64_t n = len - k + 1;
    if (n < 1) return 0;
    uint8_t *buf = (uint8_t *)malloc((size_t)2 * k);
    char *scratch = (char *)malloc((size_t)k);
    for (int64_t i = 0; i < n; i++) {
        const char *str = seq + i;
        if (canonical) str = canonical_kmer(str, k, scratch);
        utf16le(str, k, buf);
        uint64_t h[2];
        mo_murmur3_x64_128(buf, (size_t)2 * k, seed, h);
        out[i] = (int64_t)h[0];
    }
    free(buf); free(scratch);
    return n;
}

/* HashUtils.java:213-235 */
int64_t mo_kmer_hashes_int(const char *seq, int64_t len, int k, int canonical, int32_t *out)
{
    int64_t n = len - k + 1;
    if (n < 1) return 0;
    uint8_t *buf = (uint8_t *)malloc((size_t)2 * k);
    char *scratch = (char *)malloc((size_t)k);
    for (int64_t i = 0; i < n; i++) {
        const char *str = seq + i;
        if (canonical) str = canonical_kmer(str, k, scratch);
        utf16le(str, k, buf);
        out[i] = (int32_t)mo_murmur3_x86_32(buf, (size_t)2 * k, 0);
    }
    free(buf); free(scratch);
    return n;
}

/* ======================================================================================== */
/* sketch/MinHashSketch.java:51-179                                                          */
/* ======================================================================================== */

/* ---- sketch/FrequencyCounts.java ------------------------------------------------------------ */
struct mo_filter {
    mo_filter_params p;
    /* fractionCounts: Long2DoubleOpenHashMap, here open addressing over (hash, fraction) */
    int64_t *mkey; double *mval; uint8_t *mused; size_t mcap; int64_t mn;
    double max_value, min_value, min_idf, max_idf;       /* :223-229 */
    /* validMers: Guava BloomFilter<Long> (only when remove_unique > 0) */
    uint64_t *bits; int64_t bit_size; int32_t num_hash_functions;
};

static void filter_map_put(mo_filter *f, int64_t key, double val)
{
    if ((size_t)(f->mn + 1) * 2 > f->mcap) {
        size_t nc = f->mcap ? f->mcap * 2 : 64;
        int64_t *ok = f->mkey; double *ov = f->mval; uint8_t *ou = f->mused; size_t oc = f->mcap;
        f->mkey = (int64_t *)calloc(nc, sizeof(int64_t)); f->mval = (double *)calloc(nc, sizeof(double));
        f->mused = (uint8_t *)calloc(nc, 1); f->mcap = nc; f->mn = 0;
        for (size_t i = 0; i < oc; i++) if (ou[i]) filter_map_put(f, ok[i], ov[i]);
        free(ok); free(ov); free(ou);
    }
    size_t q = (size_t)fmix64((uint64_t)key) & (f->mcap - 1);
    while (f->mused[q] && f->mkey[q] != key) q = (q + 1) & (f->mcap - 1);
    if (!f->mused[q]) { f->mused[q] = 1; f->mkey[q] = key; f->mn++; }
    f->mval[q] = val;                                     /* Map.put: a repeated k-mer keeps the last value */
}

static int filter_map_get(const mo_filter *f, int64_t key, double *val)
{
    if (!f->mcap) return 0;
    size_t q = (size_t)fmix64((uint64_t)key) & (f->mcap - 1);
    while (f->mused[q]) {
        if (f->mkey[q] == key) { *val = f->mval[q]; return 1; }
        q = (q + 1) & (f->mcap - 1);
    }
    return 0;
}

/* Guava 19.0 BloomFilterStrategies.MURMUR128_MITZ_64 with Funnel (value, sink) -> sink.putLong(value):
 * murmur3_128(seed 0) of the 8 little-endian bytes; hash1/hash2 = lower/upper eight bytes; bit i =
 * ((hash1 + i*hash2) & Long.MAX_VALUE) % bitSize. */
static void bloom_hashes(int64_t value, uint64_t h[2])
{
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)((uint64_t)value >> (8 * i));
    mo_murmur3_x64_128(b, 8, 0, h);
}
static void bloom_put(mo_filter *f, int64_t value)
{
    uint64_t h[2]; bloom_hashes(value, h);
    uint64_t c = h[0];
    for (int i = 0; i < f->num_hash_functions; i++) {
        uint64_t bit = (c & 0x7fffffffffffffffull) % (uint64_t)f->bit_size;
        f->bits[bit >> 6] |= 1ull << (bit & 63);
        c += h[1];
    }
}
static int bloom_might_contain(const mo_filter *f, int64_t value)
{
    uint64_t h[2]; bloom_hashes(value, h);
    uint64_t c = h[0];
    for (int i = 0; i < f->num_hash_functions; i++) {
        uint64_t bit = (c & 0x7fffffffffffffffull) % (uint64_t)f->bit_size;
        if (!(f->bits[bit >> 6] & (1ull << (bit & 63)))) return 0;
        c += h[1];
    }
    return 1;
}

static double filter_idf(const mo_filter *f, double freq) { return log(f->max_value / freq - f->p.offset); }   /* :250-254 */

/* FrequencyCounts ctor :63-230 */
mo_filter *mo_filter_parse(const char *text, int64_t len, const mo_filter_params *p)
{
    mo_filter *f = (mo_filter *)calloc(1, sizeof(mo_filter));
    f->p = *p;
    f->max_value = -INFINITY;
    const char *cur = text, *end = text + len;
    int64_t size_bloom = 1;
    /* first line: "<sizeBloom> <sizeRepeat>" (:91-117) */
    {
        const char *nl = memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        if (le > cur) {
            char tmp[128]; size_t n = (size_t)(le - cur) < sizeof(tmp) - 1 ? (size_t)(le - cur) : sizeof(tmp) - 1;
            memcpy(tmp, cur, n); tmp[n] = 0;
            long long a = 0, b = 0;
            if (sscanf(tmp, "%lld %lld", &a, &b) >= 1) size_bloom = a;
            if (size_bloom == 0) size_bloom = 1;
        }
        cur = nl ? nl + 1 : end;
    }
    if (p->remove_unique > 0) {
        /* BloomFilter.create(funnel, expectedInsertions, fpp): optimalNumOfBits / optimalNumOfHashFunctions */
        const double fpp = 1.0e-5;
        int64_t n = size_bloom;
        int64_t num_bits = (int64_t)(-(double)n * log(fpp) / (log(2.0) * log(2.0)));
        int nh = (int)floor((double)num_bits / (double)n * log(2.0) + 0.5);
        f->num_hash_functions = nh < 1 ? 1 : nh;
        int64_t words = (num_bits + 63) / 64;             /* LongMath.divide(bits, 64, CEILING) */
        if (words < 1) words = 1;
        f->bits = (uint64_t *)calloc((size_t)words, 8);
        f->bit_size = words * 64;                         /* BitArray.bitSize() */
    }
    char kbuf[4096], rcbuf[4096];
    while (cur < end) {
        const char *nl = memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        const char *q = cur;
        /* String.split("\\s+", 3): a leading separator yields an empty first token (k-mer of length 0: no hash, caught) */
        const char *t0 = q;
        while (q < le && !(*q == ' ' || *q == '\t' || *q == '\r' || *q == '\f' || *q == '\v')) q++;
        size_t klen = (size_t)(q - t0);
        if (klen >= 1 && klen < sizeof(kbuf)) {
            memcpy(kbuf, t0, klen);
            const char *str = kbuf;
            if (p->canonical) str = canonical_kmer(kbuf, (int)klen, rcbuf);     /* HashUtils.java:246-251 */
            uint8_t u16[2 * sizeof(kbuf)];
            utf16le(str, (int)klen, u16);
            uint64_t h[2];
            mo_murmur3_x64_128(u16, 2 * klen, 0, h);
            while (q < le && (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\f' || *q == '\v')) q++;
            if (q < le) {                                 /* str.length >= 2 (:178-193) */
                char num[64]; size_t n = 0;
                while (q < le && n < sizeof(num) - 1 && !(*q == ' ' || *q == '\t' || *q == '\r')) num[n++] = *q++;
                num[n] = 0;
                char *ep; double percent = strtod(num, &ep);
                if (ep != num && *ep == 0) {
                    if (percent >= p->filter_cutoff) {
                        if (percent > f->max_value) f->max_value = percent;
                        filter_map_put(f, (int64_t)h[0], percent);
                    }
                } else { cur = nl ? nl + 1 : end; continue; }   /* NumberFormatException: the line is skipped (:204-207) */
            }
            if (p->remove_unique > 0) bloom_put(f, (int64_t)h[0]);
        }
        cur = nl ? nl + 1 : end;
    }
    f->min_value = p->filter_cutoff;                      /* :226 */
    f->min_idf = filter_idf(f, f->max_value);             /* :228 */
    f->max_idf = filter_idf(f, f->min_value);             /* :229 */
    return f;
}

void mo_filter_free(mo_filter *f)
{
    if (!f) return;
    free(f->mkey); free(f->mval); free(f->mused); free(f->bits); free(f);
}
int64_t mo_filter_size(const mo_filter *f) { return f->mn; }
double mo_filter_max_value(const mo_filter *f) { return f->max_value; }
int64_t mo_filter_export(const mo_filter *f, int64_t *hashes, double *fractions)
{
    int64_t n = 0;
    for (size_t i = 0; i < f->mcap; i++) if (f->mused[i]) { hashes[n] = f->mkey[i]; fractions[n] = f->mval[i]; n++; }
    return n;
}
int64_t mo_filter_bloom_export(const mo_filter *f, const uint64_t **words, int32_t *nh)
{
    *words = f->bits; *nh = f->num_hash_functions; return f->bits ? f->bit_size : 0;
}
int mo_filter_is_popular(const mo_filter *f, int64_t hash) { double v; return filter_map_get(f, hash, &v); }
int mo_filter_keep_kmer(const mo_filter *f, int64_t hash)
{
    if (f->p.remove_unique == 1) return bloom_might_contain(f, hash);
    return 1;
}
double mo_filter_scaled_idf(const mo_filter *f, int64_t hash)
{
    if (f->p.remove_unique == 2 && f->bits && !bloom_might_contain(f, hash)) return 1.0;   /* :292-293 */
    double val;
    if (!filter_map_get(f, hash, &val)) return f->p.range;                                /* :295-297 */
    double idf = filter_idf(f, val);
    double scale = (f->max_idf - f->min_idf) / (f->p.range - 1.0);
    return 1.0 + (idf - f->min_idf) / scale;
}

int mo_minhash_sketch_filtered(const char *seq, int64_t len, int k, int num_hashes, double repeat_weight,
                               const mo_filter *f, int32_t *hashes)
{
    int64_t n = len - k + 1;
    if (n < 1) return 1; /* :55-56 ZeroNGramsFoundException */

    int64_t *kmer = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    mo_kmer_hashes_long(seq, len, k, 0, 0, kmer); /* :62 (doReverseCompliment=false, SequenceSketch.java:112) */

    /* :66-81 Long2ObjectLinkedOpenHashMap<HitCounter>: insertion-ordered multiset count. */
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    int64_t *slot = (int64_t *)malloc(sizeof(int64_t) * cap); /* index into keys[], -1 empty */
    for (size_t i = 0; i < cap; i++) slot[i] = -1;
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int32_t *counts = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int64_t ndistinct = 0;
    for (int64_t i = 0; i < n; i++) {
        if (f && !mo_filter_keep_kmer(f, kmer[i])) continue;   /* :70-71 */
        uint64_t h = fmix64((uint64_t)kmer[i]);
        size_t p = (size_t)h & (cap - 1);
        for (;;) {
            if (slot[p] < 0) {
                slot[p] = ndistinct; keys[ndistinct] = kmer[i]; counts[ndistinct] = 1; ndistinct++;
                break;
            }
            if (keys[slot[p]] == kmer[i]) { counts[slot[p]]++; break; }
            p = (p + 1) & (cap - 1);
        }
    }
    free(slot); free(kmer);
    if (ndistinct == 0) { free(keys); free(counts); return 1; }   /* :84-85 */

    /* :87-90 */
    int nh = num_hashes < 1 ? 1 : num_hashes;
    for (int i = 0; i < nh; i++) hashes[i] = 0;
    int64_t *best = (int64_t *)malloc(sizeof(int64_t) * (size_t)(num_hashes > 0 ? num_hashes : 1));
    for (int i = 0; i < num_hashes; i++) best[i] = INT64_MAX;

    /* :95-154 */
    int64_t number_valid = 0;
    for (int64_t e = 0; e < ndistinct; e++) {
        int64_t key = keys[e];
        int32_t weight = counts[e];
        if (repeat_weight < 0.0) {                        /* :101-107 original MHAP */
            weight = 1;
            if (f && mo_filter_is_popular(f, key)) weight = 0;
        } else if (f) {
            if (repeat_weight >= 0.0 && repeat_weight < 1.0) {   /* :111-124 tf-idf */
                double tf = f->p.no_tf ? 1.0 : (double)weight;   /* FrequencyCounts.tfWeight :312-318 */
                double idf = mo_filter_scaled_idf(f, key);
                double r = floor(tf * idf + 0.5);                /* Math.round */
                weight = (r != r) ? 0 : (r > 2147483647.0 ? 2147483647 : (int32_t)r);
                if (weight < 1) weight = 1;
            }
        }
        if (weight <= 0) continue;
        number_valid++;
        uint64_t x = (uint64_t)key;
        for (int word = 0; word < num_hashes; word++) {
            for (int c = 0; c < weight; c++) {
                x ^= x << 21;
                x ^= x >> 35; /* >>> */
                x ^= x << 4;
                if ((int64_t)x < best[word]) {
                    best[word] = (int64_t)x;
                    if (word % 2 == 0) hashes[word] = (int32_t)(uint32_t)(uint64_t)key;
                    else               hashes[word] = (int32_t)(uint32_t)((uint64_t)key >> 32);
                }
            }
        }
    }
    free(best); free(keys); free(counts);
    return number_valid <= 0 ? 1 : 0;                     /* :156-157 */
}

int mo_minhash_sketch(const char *seq, int64_t len, int k, int num_hashes, int unweighted,
                      int32_t *hashes)
{
    return mo_minhash_sketch_filtered(seq, len, k, num_hashes, unweighted ? -1.0 : 0.9, NULL, hashes);
}

/* ======================================================================================== */
/* sketch/BottomOverlapSketch.java                                                            */
/* ======================================================================================== */

typedef struct { int32_t hash; int32_t pos; } hp_t;

static int hp_cmp(const void *a, const void *b)
{
    const hp_t *x = (const hp_t *)a, *y = (const hp_t *)b;
    if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1; /* signed ascending */
    return (x->pos > y->pos) - (x->pos < y->pos);             /* stable == ties by position */
}

/* :525-559.  fastutil IntArrays.radixSortIndirect(perm, hashes, stable=true): ascending signed
 * int order, equal keys keep ascending perm (=position) order. */
int32_t mo_bottom_sketch(const char *seq, int64_t len, int ok, int sketch_size,
                         int32_t *out_hash_pos, int32_t *seq_len_kmers)
{
    int64_t n = len - ok + 1;
    if (seq_len_kmers) *seq_len_kmers = (int32_t)n;
    if (n <= 0) return -1; /* :530-531 */
    int32_t *h = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    mo_kmer_hashes_int(seq, len, ok, 0, h);
    hp_t *v = (hp_t *)malloc(sizeof(hp_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) { v[i].hash = h[i]; v[i].pos = (int32_t)i; }
    qsort(v, (size_t)n, sizeof(hp_t), hp_cmp);
    int32_t kk = sketch_size < n ? sketch_size : (int32_t)n; /* :548 */
    if (kk < 0) kk = 0;
    for (int32_t i = 0; i < kk; i++) { out_hash_pos[2 * i] = v[i].hash; out_hash_pos[2 * i + 1] = v[i].pos; }
    free(v); free(h);
    return kk;
}

/* Math.round(double) of Java 8: nearest, ties toward +inf. */
static inline int64_t java_round(double x)
{
    double f = floor(x);
    return (int64_t)f + ((x - f) >= 0.5 ? 1 : 0);
}
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t iabs(int32_t a) { return a < 0 ? -a : a; }

/* MatchData :64-298 */
typedef struct {
    int32_t abs_max_shift, count, median_shift, need_recompute;
    double max_shift_percent;
    int32_t *pos1, *pos2, *shift;
    int32_t cap;
    int32_t len1, len2;
} match_data;

static void md_perform_update(match_data *m) /* :191-215 */
{
    if (m->need_recompute) {
        if (m->count > 0) {
            int32_t *copy = (int32_t *)malloc(sizeof(int32_t) * (size_t)m->count);
            memcpy(copy, m->shift, sizeof(int32_t) * (size_t)m->count);
            m->median_shift = mo_quick_select(copy, m->count / 2, m->count);
            free(copy);
            int32_t left = imax(0, -m->median_shift);
            int32_t right = imin(m->len1, m->len2 - m->median_shift);
            int32_t overlap = imax(10, right - left);
            m->abs_max_shift = imin(imax(m->len1, m->len2), (int32_t)((double)overlap * m->max_shift_percent));
        } else {
            m->median_shift = 0;
            m->abs_max_shift = imax(m->len1, m->len2) + 1;
        }
    }
    m->need_recompute = 0;
}
static int32_t md_median(match_data *m) { md_perform_update(m); return m->median_shift; }
static int32_t md_absmax(match_data *m) { md_perform_update(m); return m->abs_max_shift; }

static void md_record(match_data *m, int32_t p1, int32_t p2, int32_t sh) /* :217-233 */
{
    if (m->cap <= m->count) {
        m->cap *= 2;
        m->pos1 = (int32_t *)realloc(m->pos1, sizeof(int32_t) * (size_t)m->cap);
        m->pos2 = (int32_t *)realloc(m->pos2, sizeof(int32_t) * (size_t)m->cap);
        m->shift = (int32_t *)realloc(m->shift, sizeof(int32_t) * (size_t)m->cap);
    }
    m->shift[m->count] = sh; m->pos1[m->count] = p1; m->pos2[m->count] = p2;
    m->count++;
    m->need_recompute = 1;
}

/* recordMatchingKmers :397-516 */
static void record_matching_kmers(match_data *m, const int32_t *s1, int32_t n1, const int32_t *s2, int32_t n2)
{
    int32_t median = md_median(m), absmax = md_absmax(m);
    int32_t v1lo = imax(0, -median - absmax);                 /* :246-252 */
    int32_t v2lo = imax(0, median - absmax);                  /* :262-268 */
    int32_t v1hi = imin(m->len1, m->len2 - median + absmax);  /* :254-260 */
    int32_t v2hi = imin(m->len2, m->len1 + median + absmax);  /* :270-276 */
    int32_t i1 = 0, i2 = 0;
    m->count = 0; m->need_recompute = 1; /* reset :235-239 */
    for (;;) {
        if (i1 >= n1) break;
        if (i2 >= n2) break;
        int32_t hash1 = s1[2 * i1], pos1 = s1[2 * i1 + 1];
        int32_t hash2 = s2[2 * i2], pos2 = s2[2 * i2 + 1];
        if (hash1 < hash2 || pos1 < v1lo || pos1 >= v1hi) i1++;
        else if (hash2 < hash1 || pos2 < v2lo || pos2 >= v2hi) i2++;
        else {
            int32_t curr = pos2 - pos1;
            int32_t diff = curr - median;
            if (diff > absmax) i1++;
            else if (diff < -absmax) i2++;
            else {
                md_record(m, pos1, pos2, curr);
                int32_t i1last = i1, i1try = i1 + 1;
                if (i1try < n1) {
                    int32_t h = s1[2 * i1try], p = s1[2 * i1try + 1];
                    while (h == hash1 && p >= v1lo && p < v1hi) {
                        i1last = i1try;
                        i1try++;
                        if (i1try >= n1) break;
                        h = s1[2 * i1try]; p = s1[2 * i1try + 1];
                    }
                }
                int32_t i2last = i2, i2try = i2 + 1;
                if (i2try < n2) {
                    int32_t h = s2[2 * i2try], p = s2[2 * i2try + 1];
                    while (h == hash2 && p >= v2lo && p < v2hi) {
                        i2last = i2try;
                        i2try++;
                        if (i2try >= n2) break;
                        h = s2[2 * i2try]; p = s2[2 * i2try + 1];
                    }
                }
                if (i1 != i1last || i2 != i2last) {
                    int32_t p1n = s1[2 * i1last + 1], p2n = s2[2 * i2last + 1];
                    md_record(m, p1n, p2n, p2n - p1n);
                    i1 = i1last + 1; i2 = i2last + 1;
                } else { i1++; i2++; }
            }
        }
    }
}

/* optimizeShifts :156-189 */
static void md_optimize_shifts(match_data *m)
{
    if (m->count <= 0) return;
    int32_t reduced = -1;
    int32_t median = md_median(m);
    for (int32_t it = 0; it < m->count; it++) {
        if (reduced >= 0 && m->pos1[reduced] == m->pos1[it]) {
            if (iabs(m->shift[reduced] - median) > iabs(m->shift[it] - median)) {
                m->pos1[reduced] = m->pos1[it]; m->pos2[reduced] = m->pos2[it]; m->shift[reduced] = m->shift[it];
            }
        } else {
            reduced++;
            m->pos1[reduced] = m->pos1[it]; m->pos2[reduced] = m->pos2[it]; m->shift[reduced] = m->shift[it];
        }
    }
    m->count = reduced + 1;
    m->need_recompute = 1;
}

/* jaccardToIdentity :391-395 */
double mo_jaccard_to_identity(double score, int kmer_size)
{
    double d = -1.0 / (double)kmer_size * log(2.0 * score / (1.0 + score));
    return exp(-d);
}

/* getOverlapInfo :592-630 */
void mo_overlap_info(const int32_t *A, int32_t nA, int32_t lenA, const int32_t *B, int32_t nB, int32_t lenB,
                     int ok, double max_shift, mo_overlap *out)
{
    memset(out, 0, sizeof(*out));
    out->empty = 1;
    match_data m;
    m.len1 = lenA; m.len2 = lenB;
    m.cap = imax(nA, nB) / 4 + 1; /* :77-81 */
    m.pos1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)m.cap);
    m.pos2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)m.cap);
    m.shift = (int32_t *)malloc(sizeof(int32_t) * (size_t)m.cap);
    m.max_shift_percent = max_shift;
    m.count = 0; m.need_recompute = 1; m.median_shift = 0; m.abs_max_shift = 0;

    record_matching_kmers(&m, A, nA, B, nB);        /* :601 */
    if (m.count <= 0) goto done;
    record_matching_kmers(&m, A, nA, B, nB);        /* :607 */
    if (m.count <= 0) goto done;
    md_optimize_shifts(&m);                          /* :612 */
    if (m.count <= 0) goto done;
    {
        /* computeEdges :90-137 */
        int32_t le1 = INT32_MAX, le2 = INT32_MAX, re1 = INT32_MIN, re2 = INT32_MIN, valid = 0;
        int32_t median = md_median(&m), absmax = md_absmax(&m);
        for (int32_t it = 0; it < m.count; it++) {
            int32_t p1 = m.pos1[it], p2 = m.pos2[it];
            if (iabs(m.shift[it] - median) > absmax) continue;
            if (p1 < le1) le1 = p1;
            if (p2 < le2) le2 = p2;
            if (p1 > re1) re1 = p1;
            if (p2 > re2) re2 = p2;
            valid++;
        }
        if (valid < 3) goto done;
        /* Java int arithmetic wraps; operands here are < 2^31 / 3072 so no overflow in practice,
         * but compute in int32 with wraparound to stay literal. */
        int32_t n = valid;
        int32_t a1 = imax(0, (int32_t)java_round((double)(int32_t)((uint32_t)n * (uint32_t)le1 - (uint32_t)re1) / (double)(n - 1)));
        int32_t a2 = imin(lenA, (int32_t)java_round((double)(int32_t)((uint32_t)n * (uint32_t)re1 - (uint32_t)le1) / (double)(n - 1)));
        int32_t b1 = imax(0, (int32_t)java_round((double)(int32_t)((uint32_t)n * (uint32_t)le2 - (uint32_t)re2) / (double)(n - 1)));
        int32_t b2 = imin(lenB, (int32_t)java_round((double)(int32_t)((uint32_t)n * (uint32_t)re2 - (uint32_t)le2) / (double)(n - 1)));

        /* computeKBottomSketchJaccard :304-364 */
        int32_t *h1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nA > 0 ? nA : 1));
        int32_t *h2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nB > 0 ? nB : 1));
        int32_t s1 = 0, s2 = 0;
        for (int32_t i = 0; i < nA; i++) { int32_t p = A[2 * i + 1]; if (p >= a1 && p <= a2) h1[s1++] = A[2 * i]; }
        for (int32_t j = 0; j < nB; j++) { int32_t p = B[2 * j + 1]; if (p >= b1 && p <= b2) h2[s2++] = B[2 * j]; }
        int32_t k = imin(s1, s2);
        int32_t inter = 0;
        double jac = 0.0;
        if (k != 0) {
            int32_t i = 0, j = 0, uni = 0;
            while (uni < k) {
                if (h1[i] < h2[j]) i++;
                else if (h1[i] > h2[j]) j++;
                else { inter++; i++; j++; }
                uni++;
            }
            jac = (double)inter / (double)k;
        }
        free(h1); free(h2);
        out->empty = 0;
        out->a1 = a1; out->a2 = a2; out->b1 = b1; out->b2 = b2;
        out->valid_count = valid; out->intersect = inter; out->kmin = k;
        out->score = mo_jaccard_to_identity(jac, ok);
    }
done:
    free(m.pos1); free(m.pos2); free(m.shift);
}

/* ======================================================================================== */
/* store + inverted index + search                                                            */
/* ======================================================================================== */

typedef struct {
    int64_t id; int32_t is_fwd; int32_t seq_len;
    int32_t *minhash;
    int32_t seq_len_kmers; int32_t ord_n; int32_t *ord; /* [ord_n][2] */
} sketch_t;

struct mo_store {
    mo_sketch_params p;
    sketch_t *sk; int64_t n, cap;
    /* index: open-addressed (word,value) -> CSR bucket of sketch indices */
    uint64_t *tkey; uint32_t *tcount; uint64_t *toff; size_t tcap;
    int32_t *postings;
    int indexed_n;
    const mo_filter *filter; double repeat_weight; int has_filter_cfg;
};

mo_store *mo_store_new(const mo_sketch_params *p)
{
    mo_store *s = (mo_store *)calloc(1, sizeof(mo_store));
    s->p = *p;
    return s;
}

static void sketch_free(sketch_t *k) { free(k->minhash); free(k->ord); }

static void index_free(mo_store *s)
{
    free(s->tkey); free(s->tcount); free(s->toff); free(s->postings);
    s->tkey = NULL; s->tcount = NULL; s->toff = NULL; s->postings = NULL; s->tcap = 0;
}

void mo_store_free(mo_store *s)
{
    if (!s) return;
    for (int64_t i = 0; i < s->n; i++) sketch_free(&s->sk[i]);
    free(s->sk);
    index_free(s);
    free(s);
}

void mo_store_set_filter(mo_store *s, const mo_filter *f, double repeat_weight)
{
    s->filter = f; s->repeat_weight = repeat_weight; s->has_filter_cfg = 1;
}

static void store_reserve(mo_store *s, int64_t extra)
{
    if (s->n + extra > s->cap) {
        int64_t nc = s->cap ? s->cap : 64;
        while (nc < s->n + extra) nc *= 2;
        s->sk = (sketch_t *)realloc(s->sk, sizeof(sketch_t) * (size_t)nc);
        s->cap = nc;
    }
}

/* SequenceSketch.java:106-116 : returns 0 ok, 1 zero n-grams */
static int make_sketch(const mo_sketch_params *p, const mo_filter *filter, double repeat_weight, const char *seq, int64_t len, int64_t id, int is_fwd, sketch_t *out)
{
    memset(out, 0, sizeof(*out));
    out->id = id; out->is_fwd = is_fwd; out->seq_len = (int32_t)len;
    out->minhash = (int32_t *)malloc(sizeof(int32_t) * (size_t)(p->num_hashes > 0 ? p->num_hashes : 1));
    if (mo_minhash_sketch_filtered(seq, len, p->kmer_size, p->num_hashes, repeat_weight, filter, out->minhash)) {
        free(out->minhash); out->minhash = NULL; return 1;
    }
    int64_t no = len - p->ordered_kmer_size + 1;
    int64_t kk = no < p->ordered_sketch_size ? no : p->ordered_sketch_size;
    if (kk < 1) { free(out->minhash); out->minhash = NULL; return 1; }
    out->ord = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)kk);
    out->ord_n = mo_bottom_sketch(seq, len, p->ordered_kmer_size, p->ordered_sketch_size, out->ord, &out->seq_len_kmers);
    if (out->ord_n < 0) { free(out->minhash); free(out->ord); out->minhash = NULL; out->ord = NULL; return 1; }
    return 0;
}

typedef struct {
    const mo_sketch_params *p; const char *bases; const uint64_t *offsets; const int64_t *ids;
    int64_t n_reads; int both; sketch_t *tmp; uint8_t *ok; int64_t *next; pthread_mutex_t *mu;
    const mo_filter *filter; double repeat_weight;
} add_job;

static void *add_worker(void *arg)
{
    add_job *j = (add_job *)arg;
    int per = j->both ? 2 : 1;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int64_t i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n_reads) break;
        const char *seq = j->bases + j->offsets[i];
        int64_t len = (int64_t)(j->offsets[i + 1] - j->offsets[i]);
        if (len < j->p->min_olap_length) continue; /* SequenceSketchStreamer.java:129-133 */
        /* FastaData.java:194 upper-cases */
        char *up = (char *)malloc((size_t)len + 1);
        for (int64_t c = 0; c < len; c++) { char ch = seq[c]; up[c] = (ch >= 'a' && ch <= 'z') ? (char)(ch - 32) : ch; }
        if (make_sketch(j->p, j->filter, j->repeat_weight, up, len, j->ids[i], 1, &j->tmp[per * i]) == 0) {
            j->ok[per * i] = 1;
            if (j->both) {
                char *r = (char *)malloc((size_t)len + 1);
                mo_rc(up, len, r); /* Sequence.java:75-78 */
                if (make_sketch(j->p, j->filter, j->repeat_weight, r, len, j->ids[i], 0, &j->tmp[per * i + 1]) == 0) j->ok[per * i + 1] = 1;
                free(r);
            }
        }
        free(up);
    }
    return NULL;
}

int64_t mo_store_add_reads(mo_store *s, const char *bases, const uint64_t *offsets, const int64_t *ids,
                           int64_t n_reads, int both_strands, int threads)
{
    int per = both_strands ? 2 : 1;
    sketch_t *tmp = (sketch_t *)calloc((size_t)(n_reads * per + 1), sizeof(sketch_t));
    uint8_t *ok = (uint8_t *)calloc((size_t)(n_reads * per + 1), 1);
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    add_job job = { &s->p, bases, offsets, ids, n_reads, both_strands, tmp, ok, &next, &mu,
                    s->filter, s->has_filter_cfg ? s->repeat_weight : (s->p.unweighted ? -1.0 : 0.9) };
    if (threads < 1) threads = 1;
    if (threads == 1) add_worker(&job);
    else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, add_worker, &job);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    int64_t added = 0;
    store_reserve(s, n_reads * per);
    for (int64_t i = 0; i < n_reads * per; i++)
        if (ok[i]) { s->sk[s->n++] = tmp[i]; added++; }
    free(tmp); free(ok);
    return added;
}

int mo_store_add_sketch(mo_store *s, int64_t id, int is_fwd, int32_t seq_len, const int32_t *minhash,
                        int32_t seq_len_kmers, const int32_t *ord, int32_t ord_n)
{
    store_reserve(s, 1);
    sketch_t *k = &s->sk[s->n++];
    k->id = id; k->is_fwd = is_fwd; k->seq_len = seq_len;
    k->minhash = (int32_t *)malloc(sizeof(int32_t) * (size_t)s->p.num_hashes);
    memcpy(k->minhash, minhash, sizeof(int32_t) * (size_t)s->p.num_hashes);
    k->seq_len_kmers = seq_len_kmers; k->ord_n = ord_n;
    k->ord = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(ord_n > 0 ? ord_n : 1));
    memcpy(k->ord, ord, sizeof(int32_t) * 2 * (size_t)ord_n);
    return 0;
}

int64_t mo_store_size(const mo_store *s) { return s->n; }

int mo_store_get(const mo_store *s, int64_t idx, int64_t *id, int32_t *is_fwd, int32_t *seq_len,
                 const int32_t **minhash, int32_t *seq_len_kmers, const int32_t **ord, int32_t *ord_n)
{
    if (idx < 0 || idx >= s->n) return -1;
    const sketch_t *k = &s->sk[idx];
    if (id) *id = k->id;
    if (is_fwd) *is_fwd = k->is_fwd;
    if (seq_len) *seq_len = k->seq_len;
    if (minhash) *minhash = k->minhash;
    if (seq_len_kmers) *seq_len_kmers = k->seq_len_kmers;
    if (ord) *ord = k->ord;
    if (ord_n) *ord_n = k->ord_n;
    return 0;
}

static inline size_t tslot(const mo_store *s, uint64_t key) { return (size_t)fmix64(key) & (s->tcap - 1); }

/* MinHashSearch.java:101-147 : every sketch (fwd and rev) is posted under each of its H values,
 * one map per word => key = (word, value). */
void mo_store_build_index(mo_store *s)
{
    index_free(s);
    int H = s->p.num_hashes;
    uint64_t P = (uint64_t)s->n * (uint64_t)H;
    size_t cap = 64;
    while (cap < P * 2) cap <<= 1;
    s->tcap = cap;
    s->tkey = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    s->tcount = (uint32_t *)calloc(cap, sizeof(uint32_t));
    s->toff = (uint64_t *)calloc(cap, sizeof(uint64_t));
    memset(s->tkey, 0xff, sizeof(uint64_t) * cap);
    for (int64_t i = 0; i < s->n; i++)
        for (int w = 0; w < H; w++) {
            uint64_t key = ((uint64_t)(uint32_t)w << 32) | (uint32_t)s->sk[i].minhash[w];
            size_t p = tslot(s, key);
            while (s->tkey[p] != UINT64_MAX && s->tkey[p] != key) p = (p + 1) & (cap - 1);
            s->tkey[p] = key; s->tcount[p]++;
        }
    uint64_t off = 0;
    for (size_t p = 0; p < cap; p++) { s->toff[p] = off; off += s->tcount[p]; s->tcount[p] = 0; }
    s->postings = (int32_t *)malloc(sizeof(int32_t) * (size_t)(P ? P : 1));
    for (int64_t i = 0; i < s->n; i++)
        for (int w = 0; w < H; w++) {
            uint64_t key = ((uint64_t)(uint32_t)w << 32) | (uint32_t)s->sk[i].minhash[w];
            size_t p = tslot(s, key);
            while (s->tkey[p] != key) p = (p + 1) & (cap - 1);
            s->postings[s->toff[p] + s->tcount[p]++] = (int32_t)i;
        }
    s->indexed_n = 1;
}

typedef struct { mo_hit *v; int64_t n, cap; } hitvec;
static void hv_push(hitvec *h, const mo_hit *x)
{
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 256; h->v = (mo_hit *)realloc(h->v, sizeof(mo_hit) * (size_t)h->cap); }
    h->v[h->n++] = *x;
}

/* MinHashSearch.findMatches(SequenceSketch,boolean) :150-251 */
static void find_matches(const mo_store *s, const sketch_t *q, int to_self, const mo_search_params *sp, int keep_all,
                         int32_t *cnt, int32_t *touched, hitvec *out, mo_stats *st)
{
    int H = s->p.num_hashes;
    int64_t ntouched = 0, processed = 0;
    for (int w = 0; w < H; w++) { /* :166-181 */
        uint64_t key = ((uint64_t)(uint32_t)w << 32) | (uint32_t)q->minhash[w];
        size_t p = tslot(s, key);
        while (s->tkey[p] != UINT64_MAX && s->tkey[p] != key) p = (p + 1) & (s->tcap - 1);
        if (s->tkey[p] != key) continue;
        uint32_t c = s->tcount[p];
        processed += c;
        const int32_t *lst = s->postings + s->toff[p];
        for (uint32_t e = 0; e < c; e++) {
            int32_t t = lst[e];
            if (cnt[t]++ == 0) touched[ntouched++] = t;
        }
    }
    st->elements_processed += processed; /* :188 */
    st->sequences_hit += ntouched;       /* :189 */
    for (int64_t e = 0; e < ntouched; e++) { /* :194-243 (iteration order is unspecified in Java) */
        int32_t t = touched[e];
        int32_t count = cnt[t];
        cnt[t] = 0;
        const sketch_t *m = &s->sk[t];
        if (to_self && m->id == q->id) continue;                                       /* :200 */
        if (count < sp->num_min_matches) continue;                                      /* :204 */
        if (m->seq_len < sp->min_store_length && q->seq_len < sp->min_store_length) continue; /* :211 */
        if (to_self && m->id > q->id && m->seq_len >= sp->min_store_length && q->seq_len >= sp->min_store_length) continue; /* :215-219 */
        if (to_self && m->seq_len < sp->min_store_length && q->seq_len >= sp->min_store_length) continue; /* :222-225 */
        mo_overlap ov;
        mo_overlap_info(q->ord, q->ord_n, q->seq_len_kmers, m->ord, m->ord_n, m->seq_len_kmers,
                        s->p.ordered_kmer_size, sp->max_shift, &ov); /* :228 */
        int accept = ov.score >= sp->accept_score; /* :229 (EMPTY has score 0.0) */
        st->fully_compared++; /* :232 */
        if (accept) st->matches_processed++;
        if (accept || keep_all) {
            mo_hit h;
            memset(&h, 0, sizeof(h));
            h.from_id = q->id; h.to_id = m->id; h.from_fwd = q->is_fwd; h.to_fwd = m->is_fwd;
            h.hit_count = count;
            h.a1 = ov.a1; h.a2 = ov.a2; h.b1 = ov.b1; h.b2 = ov.b2;
            h.valid_count = ov.valid_count; h.intersect = ov.intersect; h.kmin = ov.kmin;
            h.from_len = q->seq_len; h.to_len = m->seq_len;
            h.score = ov.score; h.accepted = accept;
            hv_push(out, &h);
        }
    }
}

typedef struct {
    const mo_store *s; const sketch_t *queries; int64_t nq; int fwd_only_queries; int to_self;
    const mo_search_params *sp; int keep_all; int64_t *next; pthread_mutex_t *mu;
    hitvec hits; mo_stats st;
} search_job;

static void *search_worker(void *arg)
{
    search_job *j = (search_job *)arg;
    int32_t *cnt = (int32_t *)calloc((size_t)(j->s->n + 1), sizeof(int32_t));
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(j->s->n + 1));
    for (;;) {
        pthread_mutex_lock(j->mu);
        int64_t b = *j->next; *j->next += 16;
        pthread_mutex_unlock(j->mu);
        if (b >= j->nq) break;
        int64_t e = b + 16 < j->nq ? b + 16 : j->nq;
        for (int64_t i = b; i < e; i++) {
            const sketch_t *q = &j->queries[i];
            if (j->fwd_only_queries && !q->is_fwd) continue; /* AbstractMatchSearch.java:128-129 / :225 */
            find_matches(j->s, q, j->to_self, j->sp, j->keep_all, cnt, touched, &j->hits, &j->st);
            j->st.sequences_searched++;
        }
    }
    free(cnt); free(touched);
    return NULL;
}

static int run_search(mo_store *s, const sketch_t *queries, int64_t nq, int to_self, const mo_search_params *sp,
                      int threads, int keep_all, mo_hit **out, int64_t *n_out, mo_stats *stats)
{
    if (!s->indexed_n) mo_store_build_index(s);
    if (threads < 1) threads = 1;
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    search_job *jobs = (search_job *)calloc((size_t)threads, sizeof(search_job));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].s = s; jobs[t].queries = queries; jobs[t].nq = nq; jobs[t].fwd_only_queries = 1;
        jobs[t].to_self = to_self; jobs[t].sp = sp; jobs[t].keep_all = keep_all; jobs[t].next = &next; jobs[t].mu = &mu;
    }
    if (threads == 1) search_worker(&jobs[0]);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, search_worker, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    int64_t total = 0;
    mo_stats st; memset(&st, 0, sizeof(st));
    for (int t = 0; t < threads; t++) {
        total += jobs[t].hits.n;
        st.elements_processed += jobs[t].st.elements_processed; st.sequences_hit += jobs[t].st.sequences_hit;
        st.fully_compared += jobs[t].st.fully_compared; st.matches_processed += jobs[t].st.matches_processed;
        st.sequences_searched += jobs[t].st.sequences_searched;
    }
    mo_hit *all = (mo_hit *)malloc(sizeof(mo_hit) * (size_t)(total ? total : 1));
    int64_t o = 0;
    for (int t = 0; t < threads; t++) {
        if (jobs[t].hits.n) memcpy(all + o, jobs[t].hits.v, sizeof(mo_hit) * (size_t)jobs[t].hits.n);
        o += jobs[t].hits.n;
        free(jobs[t].hits.v);
    }
    free(jobs); free(th);
    *out = all; *n_out = total;
    if (stats) *stats = st;
    return 0;
}

int mo_search_self(mo_store *s, const mo_search_params *sp, int threads, int keep_all, mo_hit **out, int64_t *n_out, mo_stats *stats)
{
    return run_search(s, s->sk, s->n, 1, sp, threads, keep_all, out, n_out, stats);
}

/* Self search restricted to stored sketches [first, first+count) as queries: the sharding rule of
 * the multi-GPU path (each rank queries its own shard against the full index). */
int mo_search_self_range(mo_store *s, int64_t first, int64_t count, const mo_search_params *sp, int threads, int keep_all,
                         mo_hit **out, int64_t *n_out, mo_stats *stats)
{
    if (first < 0) first = 0;
    if (first > s->n) first = s->n;
    if (count < 0 || first + count > s->n) count = s->n - first;
    return run_search(s, s->sk + first, count, 1, sp, threads, keep_all, out, n_out, stats);
}

int mo_search_query(mo_store *s, const mo_store *q, const mo_search_params *sp, int threads, int keep_all,
                    mo_hit **out, int64_t *n_out, mo_stats *stats)
{
    return run_search(s, q->sk, q->n, 0, sp, threads, keep_all, out, n_out, stats);
}

/* Queries from another store under the self-search id rules (toSelf = true): the sharded multi-GPU job, where a
 * rank's store holds only its shard and the queries are every rank's forward sketches. */
int mo_search_query_self(mo_store *s, const mo_store *q, const mo_search_params *sp, int threads, int keep_all,
                         mo_hit **out, int64_t *n_out, mo_stats *stats)
{
    return run_search(s, q->sk, q->n, 1, sp, threads, keep_all, out, n_out, stats);
}

void mo_free(void *p) { free(p); }

/* MatchResult.java:46-65,98-113 */
int mo_format_match(const mo_hit *h, char *buf, size_t buflen)
{
    int32_t a1 = h->from_fwd ? h->a1 : h->from_len - h->a2 - 1;
    int32_t a2 = h->from_fwd ? h->a2 : h->from_len - h->a1 - 1;
    int32_t b1 = h->to_fwd ? h->b1 : h->to_len - h->b2 - 1;
    int32_t b2 = h->to_fwd ? h->b2 : h->to_len - h->b1 - 1;
    double score = h->score > 1.0 ? 1.0 : h->score;
    return snprintf(buf, buflen, "%lld %lld %.6f %.6f %d %d %d %d %d %d %d %d",
                    (long long)h->from_id, (long long)h->to_id, 1.0 - score, (double)h->valid_count,
                    h->from_fwd ? 0 : 1, a1, a2, h->from_len, h->to_fwd ? 0 : 1, b1, b2, h->to_len);
}

/* ---- .dat record ------------------------------------------------------------------------- */
static inline uint8_t *put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; return p + 4; }
static inline uint8_t *put64(uint8_t *p, uint64_t v) { p = put32(p, (uint32_t)(v >> 32)); return put32(p, (uint32_t)v); }

int64_t mo_dat_encode(int64_t id, int is_fwd, const char *header, int32_t seq_len, const int32_t *minhash, int32_t H,
                      int32_t seq_len_kmers, int32_t ok, const int32_t *ord, int32_t ord_n, uint8_t *buf)
{
    char idbuf[32];
    if (!header) { snprintf(idbuf, sizeof idbuf, "%lld", (long long)id); header = idbuf; }
    size_t hl = strlen(header); /* ASCII => modified UTF-8 == bytes */
    int64_t payload = 1 + 8 + 2 + (int64_t)hl + 4 + 4 + 4 * (int64_t)H + 12 + 8 * (int64_t)ord_n;
    int64_t total = 1 + 4 + payload;
    if (!buf) return total;
    uint8_t *p = buf;
    *p++ = is_fwd ? 1 : 0;                   /* SequenceSketchStreamer.java:352-356 */
    p = put32(p, (uint32_t)payload);
    *p++ = is_fwd ? 1 : 0;                   /* SequenceSketch.java:135 */
    p = put64(p, (uint64_t)id);
    *p++ = (uint8_t)(hl >> 8); *p++ = (uint8_t)hl; memcpy(p, header, hl); p += hl;
    p = put32(p, (uint32_t)seq_len);
    p = put32(p, (uint32_t)H);               /* MinHashSketch.java:218-230 */
    for (int i = 0; i < H; i++) p = put32(p, (uint32_t)minhash[i]);
    p = put32(p, (uint32_t)seq_len_kmers);   /* BottomOverlapSketch.java:561-585 */
    p = put32(p, (uint32_t)ok);
    p = put32(p, (uint32_t)ord_n);
    for (int i = 0; i < 2 * ord_n; i++) p = put32(p, (uint32_t)ord[i]);
    return total;
}
