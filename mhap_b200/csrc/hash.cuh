// hash.cuh -- device restatements of the two Guava hash functions MHAP calls, specialised to the
// byte stream Hasher.putUnencodedChars produces (each Java char = 2 bytes little-endian, high byte 0
// for the Latin-1 alphabet FASTA uses).
//   murmur3_128(seed 0).asLong()  -> sketch/HashUtils.java:237-258 (computeSequenceHashesLong)
//   murmur3_32(seed 0).asInt()    -> sketch/HashUtils.java:213-235 (computeSequenceHashes)
// Written from the published MurmurHash3 algorithm (Appleby), not from Guava source.
#pragma once
#include <stdint.h>

namespace mhapb {

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// utils/Utils.java:84-114 (Translate.lookup) + :503 toUpperCase: complement of an upper-cased char;
// characters outside the IUPAC table are returned unchanged.
__host__ __device__ __forceinline__ uint8_t complement_char(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'B': return 'V'; case 'C': return 'G'; case 'D': return 'H';
    case 'G': return 'C'; case 'H': return 'D'; case 'K': return 'M'; case 'M': return 'K';
    case 'R': return 'Y'; case 'T': return 'A'; case 'V': return 'B'; case 'Y': return 'R';
    default:  return c;   // N, S, W map to themselves; unknown characters are kept
    }
}
__host__ __device__ __forceinline__ uint8_t upper_char(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// h1 of MurmurHash3_x64_128 (seed 0) over k chars as UTF-16LE.  C is any callable int -> uint8_t char.
template <class C>
__host__ __device__ __forceinline__ uint64_t murmur3_128_h1_chars(const C &ch, int k)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = k >> 3;            // 16-byte blocks = 8 chars
    int p = 0;
#pragma unroll 4
    for (int b = 0; b < nblocks; b++, p += 8) {
        uint64_t k1 = (uint64_t)ch(p) | ((uint64_t)ch(p + 1) << 16) | ((uint64_t)ch(p + 2) << 32) | ((uint64_t)ch(p + 3) << 48);
        uint64_t k2 = (uint64_t)ch(p + 4) | ((uint64_t)ch(p + 5) << 16) | ((uint64_t)ch(p + 6) << 32) | ((uint64_t)ch(p + 7) << 48);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int m = k & 7;                   // tail chars (2m tail bytes)
    if (m) {
        uint64_t k1 = 0, k2 = 0;
        for (int j = 0; j < m && j < 4; j++) k1 |= (uint64_t)ch(p + j) << (16 * j);
        for (int j = 4; j < m; j++) k2 |= (uint64_t)ch(p + j) << (16 * (j - 4));
        if (m >= 5) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    const uint64_t len = (uint64_t)(2 * k);
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

#ifdef __CUDACC__
// The same hash for k = 16 with the (upper-cased) characters in shared memory: the 16 characters are fetched as five aligned
// 32-bit words and shifted into place with funnel shifts, and each group of four characters is widened to UTF-16LE with two
// byte permutes -- 17 instructions instead of 16 byte loads and two dozen 64-bit shift/or pairs.  chars must be 4-byte aligned
// and readable up to 3 bytes past character i+15.
__device__ __forceinline__ uint64_t murmur3_128_h1_16chars_smem(const uint8_t *chars, int i)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(chars + (i & ~3));
    const uint32_t sh = (uint32_t)(i & 3) * 8u;
    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
    const uint32_t c[4] = {__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh)};
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        uint64_t k1 = ((uint64_t)__byte_perm(c[2 * b], 0, 0x4342) << 32) | __byte_perm(c[2 * b], 0, 0x4140);
        uint64_t k2 = ((uint64_t)__byte_perm(c[2 * b + 1], 0, 0x4342) << 32) | __byte_perm(c[2 * b + 1], 0, 0x4140);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    h1 ^= 32; h2 ^= 32;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}
#endif

// Guava 19.0 BloomFilterStrategies.MURMUR128_MITZ_64 hashes a Long through Funnel (v, sink) -> sink.putLong(v):
// MurmurHash3_x64_128 (seed 0) of the key's 8 little-endian bytes (tail-only input: one k1 lane).
// hash1 = lower eight bytes (h1), hash2 = upper eight (h2); probe i tests bit ((h1 + i*h2) & Long.MAX_VALUE) % bitSize.
__host__ __device__ __forceinline__ void bloom_hash_pair(uint64_t key, uint64_t *o1, uint64_t *o2)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t k1 = key * c1; k1 = rotl64(k1, 31); k1 *= c2;
    uint64_t h1 = k1, h2 = 0;
    h1 ^= 8; h2 ^= 8;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    *o1 = h1; *o2 = h2;
}

// MurmurHash3_x86_32 (seed 0) over k chars as UTF-16LE.
template <class C>
__host__ __device__ __forceinline__ uint32_t murmur3_32_chars(const C &ch, int k)
{
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = 0;
    const int nblocks = k >> 1;            // 4-byte blocks = 2 chars
    int p = 0;
#pragma unroll 8
    for (int b = 0; b < nblocks; b++, p += 2) {
        uint32_t kk = (uint32_t)ch(p) | ((uint32_t)ch(p + 1) << 16);
        kk *= c1; kk = rotl32(kk, 15); kk *= c2;
        h ^= kk; h = rotl32(h, 13); h = h * 5 + 0xe6546b64u;
    }
    if (k & 1) {                           // 2 tail bytes: (char, 0)
        uint32_t kk = (uint32_t)ch(p);
        kk *= c1; kk = rotl32(kk, 15); kk *= c2; h ^= kk;
    }
    h ^= (uint32_t)(2 * k);
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

} // namespace mhapb
