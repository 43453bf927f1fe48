// sketch.cu -- K1: per-read MinHashSketch + BottomOverlapSketch construction on sm_100a.
//
// Replaces (paths relative to /root/reference/src/main/java/edu/umd/marbl/mhap/):
//   K1a k_hash_dedup : sketch/HashUtils.java:237-258 + the multiset count of
//                      sketch/MinHashSketch.java:66-81
//   K1b k_minhash    : the weighted XORShift-min loop, sketch/MinHashSketch.java:95-154
//   K1c k_ordered    : sketch/HashUtils.java:213-235 + sketch/BottomOverlapSketch.java:525-559
//
// Pure integer work, no tensor cores.  K1b dominates (nk*H XORShift steps per strand) and is
// bound by integer issue, not HBM -- see DESIGN.md.
#include "engine.h"
#include "hash.cuh"
#include "bs_step.cuh"
#include "bs_filter.cuh"

#include <algorithm>
#include <cstdlib>

namespace mhapb {

static constexpr uint64_t kEmptyKey = ~0ull;
static constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// character sources
// ---------------------------------------------------------------------------------------------
struct CharsSmem {
    const uint8_t *s;
    __device__ __forceinline__ uint8_t operator()(int i) const { return s[i]; }
};
// Direct from HBM with FastaData's upper-casing and Utils.rc applied on the fly (long reads).
struct CharsGlobal {
    const uint8_t *g; uint32_t len; uint32_t rc;
    __device__ __forceinline__ uint8_t operator()(int i) const
    {
        uint8_t c = upper_char(rc ? __ldg(g + (len - 1 - (uint32_t)i)) : __ldg(g + i));
        return rc ? complement_char(c) : c;
    }
};

// 256-entry tables of FastaData's upper-casing [0] and of upper-casing followed by Utils.rc's complement [1]: one shared-memory
// look-up per character instead of a compare/subtract and a 12-way switch (the byte-wise arithmetic was 14 % of K1a's instructions)
__device__ __forceinline__ void stage_lut_init(uint8_t (*lut)[256])
{
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const uint8_t u = upper_char((uint8_t)(i & 255));
        lut[i >> 8][i & 255] = (i >> 8) ? complement_char(u) : u;
    }
}

__device__ __forceinline__ void stage_chars(uint8_t *dst, const uint8_t *g, uint32_t len, uint32_t rc, const uint8_t (*lut)[256])
{
    // 16-byte loads from the aligned window around the read (its first byte sits anywhere in the batch), one byte store per
    // character with FastaData's upper-casing and, for the reverse strand, Utils.rc applied on the way.  The byte-wise
    // version showed up with 25 % of K1c's stall samples (profiles/r2c): a CTA stages one strand at a time and waits for it.
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15;
    const int head = (int)(reinterpret_cast<uintptr_t>(g) - a0);
    const uint4 *src = reinterpret_cast<const uint4 *>(a0);
    const int nvec = (head + (int)len + 15) >> 4;
    const uint8_t *t = lut[rc ? 1 : 0];
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint4 q = __ldg(src + v);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const int i0 = v * 16 - head;
        if (i0 >= 0 && i0 + 16 <= (int)len) {            // the whole vector lies inside the read
            uint8_t *d = rc ? dst + (len - 1 - i0) : dst + i0;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint8_t c = t[(w[b >> 2] >> ((b & 3) * 8)) & 0xffu];
                if (rc) d[-b] = c; else d[b] = c;
            }
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const int i = i0 + b;
                if (i >= 0 && i < (int)len) {
                    const uint8_t c = t[(w[b >> 2] >> ((b & 3) * 8)) & 0xffu];
                    if (rc) dst[len - 1 - i] = c; else dst[i] = c;
                }
            }
        }
    }
}

// warp-aggregated cursor bump on a shared-memory counter; every lane of the warp must call it
__device__ __forceinline__ int warp_alloc(int *counter, bool want)
{
    unsigned m = __ballot_sync(kFull, want);
    if (m == 0) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(kFull, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

// ---------------------------------------------------------------------------------------------
// the -f k-mer filter on the device (sketch/FrequencyCounts.java)
// ---------------------------------------------------------------------------------------------
// Guava 19.0 BloomFilterStrategies.MURMUR128_MITZ_64.mightContain (hash pair: hash.cuh bloom_hash_pair)
__device__ __forceinline__ bool bloom_might_contain(const KmerFilterView &f, uint64_t key)
{
    uint64_t h1, h2;
    bloom_hash_pair(key, &h1, &h2);
    uint64_t c = h1;
    for (int i = 0; i < f.bloom_nfun; i++) {
        const uint64_t bit = (c & 0x7fffffffffffffffULL) % f.bloom_bits;
        if (!((f.bloom[bit >> 6] >> (bit & 63)) & 1ull)) return false;
        c += h2;
    }
    return true;
}

// fractionCounts.get(key) turned into scaledIdf on the host; returns false when the key is not a repeat k-mer
__device__ __forceinline__ bool filter_lookup(const KmerFilterView &f, uint64_t key, double *idf)
{
    if (!f.map_keys) return false;
    uint32_t q = (uint32_t)fmix64(key) & f.map_mask;
    for (;;) {
        if (!((f.map_used[q >> 5] >> (q & 31)) & 1u)) return false;
        if (f.map_keys[q] == key) { *idf = f.map_idf[q]; return true; }
        q = (q + 1) & f.map_mask;
    }
}

// sketch/MinHashSketch.java:95-130: the weight of a distinct k-mer seen `count` times (0 = not used)
__device__ __forceinline__ uint32_t kmer_weight(const KmerFilterView &f, uint64_t key, uint32_t count)
{
    if (f.mode == 0 || f.mode == 3) return count;
    double idf = 0.0;
    if (f.mode == 1) return filter_lookup(f, key, &idf) ? 0u : 1u;        // :101-107 isPopular
    // :109-124 tf-idf
    const double tf = f.no_tf ? 1.0 : (double)count;                       // FrequencyCounts.tfWeight
    if (f.remove_unique == 2 && f.bloom && !bloom_might_contain(f, key)) idf = 1.0;   // scaledIdf :292-293
    else if (!filter_lookup(f, key, &idf)) idf = f.range;                 // :295-297
    const double r = floor(tf * idf + 0.5);                               // Math.round
    if (!(r >= 1.0)) return 1u;                                           // NaN or < 1 -> 1 (:122-123)
    return r > 2147483647.0 ? 2147483647u : (uint32_t)r;
}

// ---------------------------------------------------------------------------------------------
// K1a: k-mer hashing + exact de-duplication with counts
// ---------------------------------------------------------------------------------------------
// One CTA per strand (persistent, work queue).  Output per strand: keys[koff .. koff+nlight) the hashes of the launch's
// light weight, keys[koff+nk-1 .. ] (downwards) the others with their weights.  Order is irrelevant: the XORShift map is a
// bijection, so two distinct keys never tie in the min.
//
// Two de-duplication schemes:
//  * TABLE (round 1; long strands, and the fallback for low-complexity short ones): every hash goes into an open-addressed
//    table (64-bit CAS), a repeat occurrence bumps a per-slot counter, then the table is scanned.
//  * BITMAP (short strands, the default): duplicates are rare, so they are isolated first.  Pass 1 hashes every k-mer once
//    into a shared array and sets one bit per hash in a bitmap of >= 6 bits per k-mer; a k-mer that finds its bit already set
//    is LATE (a repeat of an earlier k-mer, or a chance collision: n^2/2m, ~7 % of the k-mers).  Only the late ones go into a
//    small exact table with counts (pass 2); every other k-mer then looks its hash up in that table (pass 3): found = an
//    earlier occurrence of a repeated k-mer (counted), not found = unique, written out at once.  The table scan emits the
//    repeated k-mers with their counts.  Per k-mer this is one 32-bit atomicOr + one table look-up instead of a CAS insert
//    into a 12 k-slot table and a scan of that table with two cursor bumps per slot.

// the TABLE scheme on one strand; table / dupmask in shared (LONG = false) or global (LONG = true) memory, characters through `src`
template <bool LONG, int KC, class Src>
__device__ __forceinline__ void dedup_table_strand(const Src &src, const StrandDesc &d, int s, int k, int unweighted, uint64_t *table, uint32_t *dupmask,
                                                   uint32_t *dupcnt, uint32_t table_cap, const SketchScratch &sc, const KmerFilterView &flt,
                                                   int *s_nlight, int *s_nheavy, int *s_special)
{
    const int nk = (int)d.len - k + 1;
    // table sized for this strand (load factor <= 0.8)
    uint32_t C = dedup_table_slots((uint32_t)nk);
    if (C > table_cap) C = table_cap;
    for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) table[i] = kEmptyKey;
    if (!LONG) for (uint32_t i = threadIdx.x; i < (C + 31) / 32; i += blockDim.x) dupmask[i] = 0;
    __syncthreads();
    // uniform trip count + __syncwarp: the CAS probe below is a data-dependent loop and without an
    // explicit reconvergence point the lanes of a warp stay split for the rest of the strand
    for (int base = 0; base < nk; base += blockDim.x) {
        const int i = base + (int)threadIdx.x;
        if (i < nk) {
            const uint64_t h = murmur3_128_h1_chars([&](int j) { return src(i + j); }, KC ? KC : k);
            if (flt.remove_unique == 1 && !bloom_might_contain(flt, h)) { /* keepKmer false (MinHashSketch.java:70-71) */ }
            else if (h == kEmptyKey) atomicAdd(s_special, 1);
            else {
                // open addressing with double hashing: the stride is a power of two picked by three hash bits (C is odd,
                // so every stride is a full cycle); a plain load looks at the slot first and the CAS is only issued on
                // an empty one.
                uint32_t slot = (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * C) >> 32);
                uint32_t stride = 1u << ((uint32_t)h & 7u);
                if (stride >= C) stride = 1u;   // tiny strands: keep slot + stride - C inside the table
                for (;;) {
                    unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&table[slot]);
                    if (old == kEmptyKey)
                        old = atomicCAS(reinterpret_cast<unsigned long long *>(&table[slot]), (unsigned long long)kEmptyKey, (unsigned long long)h);
                    if (old == kEmptyKey) break;
                    if (old == h) {
                        if (!unweighted) {
                            if (!LONG) atomicOr(&dupmask[slot >> 5], 1u << (slot & 31));
                            atomicAdd(&dupcnt[slot], 1u);
                        }
                        break;
                    }
                    slot += stride;
                    if (slot >= C) slot -= C;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    uint64_t *keys = sc.keys + d.koff;
    uint32_t *wts = sc.wts + d.koff;
    const uint32_t Cr = (C + 31u) & ~31u;
    for (uint32_t i = threadIdx.x; i < Cr; i += blockDim.x) {
        uint64_t key = (i < C) ? table[i] : kEmptyKey;
        bool occ = key != kEmptyKey;
        uint32_t extra = 0;
        if (occ && !unweighted) {
            if (LONG) extra = dupcnt[i];
            else if (dupmask[i >> 5] & (1u << (i & 31))) extra = dupcnt[i];
            if (extra) dupcnt[i] = 0;   // leave the scratch zeroed for the next strand
        }
        // weight rule (MinHashSketch.java:95-130); keys of the launch's light weight go to the front, the rest
        // (with their weights) to the back, weight 0 (a popular k-mer under --repeat-weight < 0) is dropped
        uint32_t w = 0;
        if (occ) w = kmer_weight(flt, key, extra + 1);
        const bool light = occ && w == flt.light_weight, heavy = occ && w != 0 && w != flt.light_weight;
        int pl = warp_alloc(s_nlight, light);
        int ph = warp_alloc(s_nheavy, heavy);
        if (light) keys[pl] = key;
        else if (heavy) { keys[nk - 1 - ph] = key; wts[nk - 1 - ph] = w; }
    }
    (void)s;
}

// shared-memory plan of the BITMAP scheme for strands of up to nkmax k-mers (bytes, 16-byte aligned regions)
struct DedupPlan { uint32_t bits, l2cap, tslots, off_bm, off_l2, off_chars, off_tab, off_cnt, total; };
__host__ __device__ inline DedupPlan dedup_plan(uint32_t nkmax, uint32_t chars_cap)
{
    DedupPlan p;
    p.bits = 1024; while (p.bits < 6u * nkmax) p.bits <<= 1;          // >= 6 bits per k-mer: late k-mers ~ n / (2 * 6.5)
    p.l2cap = (nkmax / 8u + 128u + 7u) & ~7u;                          // late k-mers the small table is sized for (1.8x the expectation;
                                                                       // 10 kbp strands: the whole plan is 110 KB, two CTAs per SM)
    p.tslots = ((p.l2cap + p.l2cap / 3u + 8u) | 1u);                   // load <= 0.75; odd: power-of-two probe strides cycle
    auto al = [](uint32_t x) { return (x + 15u) & ~15u; };
    p.off_bm = al(nkmax * 8u);
    p.off_l2 = p.off_bm + al(p.bits / 8u);
    p.off_chars = p.off_l2 + al(p.l2cap * 2u);
    // the small table reuses the characters' region (dead after pass 1) and extends past it
    p.off_tab = p.off_chars;
    p.off_cnt = p.off_tab + al(p.tslots * 8u);
    p.total = p.off_cnt + al(p.tslots * 4u);
    if (p.total < p.off_chars + al(chars_cap)) p.total = p.off_chars + al(chars_cap);
    return p;
}

template <bool LONG, int KC /* compile-time k, 0 = runtime */>
__global__ void __launch_bounds__(1024)
k_hash_dedup(const uint8_t *__restrict__ bases, const StrandDesc *__restrict__ desc, int s_begin, int s_end,
             int k, int unweighted, uint32_t table_cap, uint32_t chars_cap, uint32_t nkmax, SketchScratch sc, uint32_t *queue,
             const KmerFilterView flt)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_strand, s_nlight, s_nheavy, s_special, s_nl2;
    __shared__ uint8_t s_lut[2][256];

    uint32_t *dupcnt = sc.dupcnt + (size_t)blockIdx.x * table_cap;
    const DedupPlan pl = LONG ? DedupPlan{} : dedup_plan(nkmax, chars_cap);
    if (!LONG) stage_lut_init(s_lut);
    const bool plain_weight = flt.mode == 0 || flt.mode == 3;      // weight = occurrence count (kmer_weight's first line), hoisted

    for (;;) {
        if (threadIdx.x == 0) {
            s_strand = s_begin + (int)atomicAdd(queue, 1u);
            s_nlight = 0; s_nheavy = 0; s_special = 0; s_nl2 = 0;
        }
        __syncthreads();
        const int s = s_strand;
        if (s >= s_end) break;
        const StrandDesc d = desc[s];
        const int nk = (int)d.len - k + 1;
        const CharsGlobal gsrc{bases + d.base_off, d.len, d.rc};
        uint64_t *keys = sc.keys + d.koff;
        uint32_t *wts = sc.wts + d.koff;

        if constexpr (LONG) {
            dedup_table_strand<true, KC>(gsrc, d, s, k, unweighted, sc.gtable + (size_t)blockIdx.x * table_cap, nullptr, dupcnt, table_cap, sc, flt,
                                         &s_nlight, &s_nheavy, &s_special);
        } else {
            uint64_t *Hh = reinterpret_cast<uint64_t *>(smem_raw);
            uint32_t *bm = reinterpret_cast<uint32_t *>(smem_raw + pl.off_bm);
            uint16_t *L2 = reinterpret_cast<uint16_t *>(smem_raw + pl.off_l2);
            uint8_t *chars = smem_raw + pl.off_chars;
            uint64_t *T = reinterpret_cast<uint64_t *>(smem_raw + pl.off_tab);
            uint32_t *Tc = reinterpret_cast<uint32_t *>(smem_raw + pl.off_cnt);
            // ---- pass 1: hash once, one bit per hash; a k-mer whose bit is already set is late ----
            for (uint32_t i = threadIdx.x; i < pl.bits / 32; i += blockDim.x) bm[i] = 0;
            stage_chars(chars, bases + d.base_off, d.len, d.rc, s_lut);
            __syncthreads();
            for (int i = threadIdx.x; i < nk; i += blockDim.x) {
                uint64_t h;
                if constexpr (KC == 16) h = murmur3_128_h1_16chars_smem(chars, i);
                else h = murmur3_128_h1_chars([&](int j) { return chars[i + j]; }, KC ? KC : k);
                if (flt.remove_unique == 1 && !bloom_might_contain(flt, h)) h = kEmptyKey;     // keepKmer false (MinHashSketch.java:70-71)
                else if (h == kEmptyKey) atomicAdd(&s_special, 1);                             // a hash equal to the marker (p = 2^-64)
                else {
                    const uint32_t b = (uint32_t)(h >> 13) & (pl.bits - 1u);
                    const uint32_t old = atomicOr(&bm[b >> 5], 1u << (b & 31));
                    if ((old >> (b & 31)) & 1u) { const int q = atomicAdd(&s_nl2, 1); if (q < (int)pl.l2cap) L2[q] = (uint16_t)i; }
                }
                Hh[i] = h;
            }
            __syncthreads();
            const int nl2 = s_nl2;
            if (nl2 > (int)pl.l2cap) {
                // low complexity: more repeats than the small table holds -> the TABLE scheme for this strand (characters from global
                // memory: the table takes the whole shared window)
                __syncthreads();
                if (threadIdx.x == 0) s_special = 0;
                __syncthreads();
                dedup_table_strand<false, KC>(gsrc, d, s, k, unweighted, reinterpret_cast<uint64_t *>(smem_raw),
                                              reinterpret_cast<uint32_t *>(smem_raw + (size_t)table_cap * 8), dupcnt, table_cap, sc, flt,
                                              &s_nlight, &s_nheavy, &s_special);
            } else {
                // ---- pass 2: the late k-mers into the small exact table, with counts ----
                for (uint32_t i = threadIdx.x; i < pl.tslots; i += blockDim.x) { T[i] = kEmptyKey; Tc[i] = 0; }
                __syncthreads();
                const uint32_t Tn = pl.tslots;
                for (int j = threadIdx.x; j < nl2; j += blockDim.x) {
                    const int idx = L2[j];
                    const uint64_t h = Hh[idx];
                    Hh[idx] = kEmptyKey;                                // consumed: pass 3 skips it
                    uint32_t slot = (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * Tn) >> 32);
                    uint32_t stride = 1u << ((uint32_t)h & 7u);
                    if (stride >= Tn) stride = 1u;
                    for (;;) {
                        unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&T[slot]);
                        if (old == kEmptyKey) old = atomicCAS(reinterpret_cast<unsigned long long *>(&T[slot]), (unsigned long long)kEmptyKey, (unsigned long long)h);
                        if (old == kEmptyKey || old == h) { atomicAdd(&Tc[slot], 1u); break; }
                        slot += stride;
                        if (slot >= Tn) slot -= Tn;
                    }
                }
                __syncthreads();
                // ---- pass 3: every other k-mer: an earlier occurrence of a repeated k-mer (counted), or unique (written out) ----
                const int nkr = (nk + 31) & ~31;
                for (int i = threadIdx.x; i < nkr; i += blockDim.x) {
                    const uint64_t h = i < nk ? Hh[i] : kEmptyKey;
                    bool uniq = false;
                    if (h != kEmptyKey) {
                        uniq = true;
                        if (nl2 > 0) {
                            uint32_t slot = (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * Tn) >> 32);
                            uint32_t stride = 1u << ((uint32_t)h & 7u);
                            if (stride >= Tn) stride = 1u;
                            for (;;) {
                                const uint64_t t = T[slot];
                                if (t == kEmptyKey) break;
                                if (t == h) { atomicAdd(&Tc[slot], 1u); uniq = false; break; }
                                slot += stride;
                                if (slot >= Tn) slot -= Tn;
                            }
                        }
                    }
                    // weight rule (MinHashSketch.java:95-130) for a k-mer seen once
                    uint32_t w = 0;
                    if (uniq) w = plain_weight ? 1u : kmer_weight(flt, h, 1u);
                    const bool light = uniq && w == flt.light_weight, heavy = uniq && w != 0 && w != flt.light_weight;
                    const int pl_ = warp_alloc(&s_nlight, light);
                    const int ph = warp_alloc(&s_nheavy, heavy);
                    if (light) keys[pl_] = h;
                    else if (heavy) { keys[nk - 1 - ph] = h; wts[nk - 1 - ph] = w; }
                }
                __syncthreads();
                // ---- the repeated k-mers with their counts ----
                const uint32_t Tr = (Tn + 31u) & ~31u;
                for (uint32_t i = threadIdx.x; i < Tr; i += blockDim.x) {
                    const uint64_t key = i < Tn ? T[i] : kEmptyKey;
                    const bool occ = key != kEmptyKey;
                    uint32_t w = 0;
                    if (occ) w = plain_weight ? (unweighted ? 1u : Tc[i]) : kmer_weight(flt, key, unweighted ? 1u : Tc[i]);
                    const bool light = occ && w == flt.light_weight, heavy = occ && w != 0 && w != flt.light_weight;
                    const int pl_ = warp_alloc(&s_nlight, light);
                    const int ph = warp_alloc(&s_nheavy, heavy);
                    if (light) keys[pl_] = key;
                    else if (heavy) { keys[nk - 1 - ph] = key; wts[nk - 1 - ph] = w; }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int nl = s_nlight, nh = s_nheavy;
            if (s_special > 0) {   // a k-mer whose hash equals the tables' empty marker (p = 2^-64)
                const uint32_t w = kmer_weight(flt, kEmptyKey, unweighted ? 1u : (uint32_t)s_special);
                if (w == flt.light_weight) { keys[nl] = kEmptyKey; nl++; }
                else if (w != 0) { keys[nk - 1 - nh] = kEmptyKey; wts[nk - 1 - nh] = w; nh++; }
            }
            sc.nlight[s] = nl;
            sc.nheavy[s] = nh;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K1b: XORShift-min
// ---------------------------------------------------------------------------------------------
// One warp per strand, systolic: lane l owns words [l*B, (l+1)*B) with their running minima in
// registers; distinct k-mers enter at lane 0 and their chain state x is handed lane to lane by
// shuffle, so every lane always advances a *different* k-mer through its own B words and the
// chain (x_{j+1} = M x_j, MinHashSketch.java:140-143) is never recomputed.  The common case per
// step is the 3-shift XORShift plus one signed compare of the high word against the lane's
// register-resident minimum; the full 64-bit compare and the update run only on the rare path.
__device__ __forceinline__ uint64_t xorshift_step(uint64_t x)
{
    x ^= x << 21;
    x ^= x >> 35;
    x ^= x << 4;
    return x;
}

template <int B>
struct LaneMins {
    int32_t hi[B];
    uint32_t lo[B];
    int32_t out[B];
};

// The step on 32-bit halves.  tools/ubench_xorshift.cu measured eight instruction mixes for this
// recurrence on B200 (shifts as SHF on the alu pipe vs as IMAD / IMAD.WIDE / IMAD.HI on the fma pipe):
// the plain form below -- ptxas emits 3 SHF + 4 LOP3 (alu) + 2 IMAD.SHL (fma) -- is the fastest at
// 14.1 cycles per warp-step (alu-pipe bound, 7 x 2 cycles); every multiply-based rewrite was slower
// (14.2-17.7), wide/high multiplies being far below the alu rate (profiles/ubench_xorshift_r1.txt).
__device__ __forceinline__ void xorshift_step32(uint32_t &lo, uint32_t &hi)
{
    uint64_t x = ((uint64_t)hi << 32) | lo;
    x = xorshift_step(x);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
}

// keys: the k-mer hashes (what ends up in the sketch); xkeys: the chain states the H words of this pass start from -- the same
// array for words 0..511, the hashes advanced by 512*weight steps per earlier pass for a later block of words (k_advance_keys)
template <int B, bool WEIGHTED, bool XK = false /* xkeys differ from keys */>
__device__ __forceinline__ void minhash_pipeline(LaneMins<B> &m, const uint64_t *__restrict__ keys, const uint64_t *__restrict__ xkeys,
                                                 const uint32_t *__restrict__ wts,
                                                 int n, int dir /* +1 light, -1 heavy */, uint64_t *kring, uint64_t *xring, uint32_t *wbuf, int lane,
                                                 uint32_t uniform_w = 1 /* WEIGHTED with wts == nullptr: every key has this weight */)
{
    constexpr int G = B < 4 ? B : 4;   // steps per rare-path check
    static_assert(B % G == 0, "B must be a multiple of the check group");
    // kring: the last 64 keys that entered the pipeline (two batches), so the rare path can fetch the
    // key of the k-mer a lane is working on (e = t - lane) with one LDS instead of keeping a global
    // address live on the hot path.
    uint32_t xl = 0, xh = 0;
    uint32_t w = 1;
    const int total = n + 31;
    for (int t0 = 0; t0 < total; t0 += 32) {
        {
            int e = t0 + lane;
            uint64_t mk = 0, mx = 0; uint32_t mw = 1;
            if (e < n) { mk = keys[(long long)dir * e]; if (XK) mx = xkeys[(long long)dir * e]; if (WEIGHTED) mw = wts ? wts[(long long)dir * e] : uniform_w; }
            kring[e & 63] = mk;
            if (XK) xring[e & 63] = mx;
            if (WEIGHTED) wbuf[lane] = mw;
        }
        __syncwarp();
        const int jn = min(32, total - t0);
#pragma unroll 1
        for (int j = 0; j < jn; j++) {
            const uint32_t il = __shfl_up_sync(kFull, xl, 1), ih = __shfl_up_sync(kFull, xh, 1);
            const uint64_t kin = XK ? xring[(t0 + j) & 63] : kring[(t0 + j) & 63];
            xl = lane == 0 ? (uint32_t)kin : il;
            xh = lane == 0 ? (uint32_t)(kin >> 32) : ih;
            if (WEIGHTED) {
                uint32_t win = __shfl_up_sync(kFull, w, 1);
                w = lane == 0 ? wbuf[j] : win;
            }
            const int e = t0 + j - lane;
            if (e >= 0 && e < n) {
                if (!WEIGHTED) {
#pragma unroll
                    for (int b0 = 0; b0 < B; b0 += G) {
                        uint32_t l[G], h[G];
                        bool any = false;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            xorshift_step32(xl, xh);
                            l[g] = xl; h[g] = xh;
                            any |= (int32_t)xh <= m.hi[b0 + g];
                        }
                        if (__builtin_expect(any, 0)) {   // rare: ~ln(n) times per word per strand
                            const uint64_t key = kring[e & 63];
#pragma unroll
                            for (int g = 0; g < G; g++) {
                                // signed 64-bit x < best[word]  (MinHashSketch.java:144)
                                if ((int32_t)h[g] < m.hi[b0 + g] || ((int32_t)h[g] == m.hi[b0 + g] && l[g] < m.lo[b0 + g])) {
                                    m.hi[b0 + g] = (int32_t)h[g]; m.lo[b0 + g] = l[g];
                                    // :146-149 even word -> (int)key, odd word -> (int)(key>>>32)
                                    m.out[b0 + g] = ((lane * B + b0 + g) & 1) ? (int32_t)(key >> 32) : (int32_t)key;
                                }
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int b = 0; b < B; b++) {
                        // register arrays need static indices: weighted k-mers (rare) go through a switch-free
                        // unrolled select below instead of indexing m.hi[b] dynamically
                        int32_t bh = 0x7fffffff; uint32_t bl = 0xffffffffu;
                        bool hit = false;
                        for (uint32_t c = 0; c < w; c++) {
                            xorshift_step32(xl, xh);
                            if ((int32_t)xh < bh || ((int32_t)xh == bh && xl < bl)) { bh = (int32_t)xh; bl = xl; hit = true; }
                        }
                        if (hit) {
                            const uint64_t key = kring[e & 63];
#pragma unroll
                            for (int bb = 0; bb < B; bb++) {
                                if (bb == b && (bh < m.hi[bb] || (bh == m.hi[bb] && bl < m.lo[bb]))) {
                                    m.hi[bb] = bh; m.lo[bb] = bl;
                                    m.out[bb] = ((lane * B + bb) & 1) ? (int32_t)(key >> 32) : (int32_t)key;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

template <int B>
__global__ void __launch_bounds__(256)
k_minhash(const StrandDesc *__restrict__ desc, int n_strands, int k, int H, SketchScratch sc,
          int32_t *__restrict__ minhash, uint32_t *queue, uint32_t light_w)
{
    __shared__ uint64_t s_kbuf[8][64], s_xbuf[8][64];
    __shared__ uint32_t s_wbuf[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (;;) {
        int s = 0;
        if (lane == 0) s = (int)atomicAdd(queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= n_strands) break;
        const StrandDesc d = desc[s];
        const int nk = (int)d.len - k + 1;
        LaneMins<B> m;
#pragma unroll
        for (int b = 0; b < B; b++) { m.hi[b] = 0x7fffffff; m.lo[b] = 0xffffffffu; m.out[b] = 0; }   // Long.MAX_VALUE
        const uint64_t *keys = sc.keys + d.koff;
        const int nl = sc.nlight[s], nh = sc.nheavy[s];
        if (light_w == 1) minhash_pipeline<B, false>(m, keys, keys, nullptr, nl, +1, s_kbuf[wib], s_xbuf[wib], s_wbuf[wib], lane);
        else              minhash_pipeline<B, true>(m, keys, keys, nullptr, nl, +1, s_kbuf[wib], s_xbuf[wib], s_wbuf[wib], lane, light_w);
        if (nh > 0)
            minhash_pipeline<B, true>(m, keys + (nk - 1), keys + (nk - 1), sc.wts + d.koff + (nk - 1), nh, -1, s_kbuf[wib], s_xbuf[wib], s_wbuf[wib], lane);
        int32_t *row = minhash + (size_t)d.row * H;
#pragma unroll
        for (int b = 0; b < B; b++) {
            int word = lane * B + b;
            if (word < H) row[word] = m.out[b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1b, bit-sliced variant (the default): 32 k-mers per thread
// ---------------------------------------------------------------------------------------------
// The scalar pipeline above is bound by the alu pipe at 7 LOP3/SHF per XORShift step (3 of them shifts).
// Bit-slicing removes the shifts: a thread keeps 64 registers R[i] = bit i of the chain states of 32
// different k-mers (a "bundle"), so x ^= x << 21 is R[i] ^= R[i-21] -- a register rename -- and one step
// of 32 chains is 132 two-input XORs (4.1 LOP3 per chain step; tools/ubench_bitslice.cu measures
// 5.1e12 steps/s bare against 2.6e12 for the scalar form).
//
// Structure (one warp per strand, still systolic): lane l owns words [l*B,(l+1)*B); bundles enter at
// lane 0 and move lane to lane each hop (64 shuffles per B*32 chain steps).  The running minima are exact
// 64-bit values in shared memory; per word-step a lane only asks "does any of my 32 chains have its top
// F bits (of the sign-biased value) all zero?", with F = a lower bound of the leading zeros of the word's
// current minimum, which is an OR over the top planes.  Flagged chains (a superset of the true updates,
// about 2x) are resolved exactly: the lane publishes its 64 planes to shared memory and the warp
// re-assembles the flagged chain's 64-bit value with two ballots.
// The first kBsScalarKeys keys of a strand (where a running minimum changes most often: the expected
// number of updates of word w after t keys is the harmonic sum) and the keys with weight > 1 go through
// the scalar pipeline first; so do strands too short to fill bundles.
constexpr int kBsScalarKeys = 512;   // floor; whole-set rounding in k_minhash_bs2 adds up to 1023 (10 kbp strands: 769)
constexpr int kBsTaps = 9;
constexpr int kBsStage = 16;   // bundles transposed per staging round (lanes 0..15 transpose one each)
__device__ __constant__ int c_bs_tap_bits[kBsTaps] = {0, 4, 8, 10, 12, 14, 16, 18, 20};

// The plain plane form of the step: 43 + 29 + 60 = 132 two-input XORs.  Kept as the statement of what
// bs_step() (bs_step.cuh, generated: the same linear map as 92 three-input XORs) must equal.
__device__ __forceinline__ void bs_step_plain(uint32_t (&R)[64])
{
#pragma unroll
    for (int i = 63; i >= 21; i--) R[i] ^= R[i - 21];      // x ^= x << 21
#pragma unroll
    for (int i = 0; i <= 28; i++) R[i] ^= R[i + 35];       // x ^= x >>> 35
#pragma unroll
    for (int i = 63; i >= 4; i--) R[i] ^= R[i - 4];        // x ^= x << 4
}

// 32x32 bit-matrix transpose in registers.  With this butterfly t[i] bit j == a[31-j] bit (31-i): bit p of
// key c ends up in t[31-p] at bit position 31-c.
template <int J>
__device__ __forceinline__ void transpose_stage(uint32_t (&a)[32], uint32_t m)
{
#pragma unroll
    for (int k = 0; k < 32; k++) {
        if ((k & J) == 0) {
            uint32_t t = (a[k] ^ (a[k + J] >> J)) & m;
            a[k] ^= t; a[k + J] ^= t << J;
        }
    }
}
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
    transpose_stage<16>(a, 0x0000ffffu); transpose_stage<8>(a, 0x00ff00ffu); transpose_stage<4>(a, 0x0f0f0f0fu);
    transpose_stage<2>(a, 0x33333333u); transpose_stage<1>(a, 0x55555555u);
}

// tap index for a word whose current minimum has biased high word tu (leading zeros F): largest tap <= F
__device__ __forceinline__ uint32_t bs_tap_of(uint32_t hi_signed)
{
    const uint32_t tu = hi_signed ^ 0x80000000u;
    const int F = tu ? __clz(tu) : 32;
    return F < 4 ? 0 : F < 8 ? 1 : F < 10 ? 2 : F < 12 ? 3 : F < 14 ? 4 : F < 16 ? 5 : F < 18 ? 6 : F < 20 ? 7 : 8;
}

struct BsState {           // lane-private exact state in shared memory, element b at [b * 32]
    uint32_t *hi, *lo, *out, *tap;
};

#ifdef MHAPB_AB_KERNELS   // the superseded bit-sliced systolic kernel, kept for A/B runs only (make AB=1)
template <int B>
__device__ __forceinline__ void bs_phase(const BsState &st, const uint64_t *__restrict__ keys /* 32*nb keys */, int nb,
                                         uint32_t *stage /* [64][kBsStage] */, uint32_t *scratch /* [64] */, int lane)
{
    uint32_t R[64];
#pragma unroll
    for (int i = 0; i < 64; i++) R[i] = 0;
    const int total = nb + 31;
    for (int t = 0; t < total; t++) {
        if ((t & (kBsStage - 1)) == 0) {
            // stage the next kBsStage bundles: lane j transposes bundle t+j (low words, then high words)
            __syncwarp();
            const int bj = t + lane;
            if (lane < kBsStage && bj < nb) {
                const uint64_t *kp = keys + (size_t)bj * 32;   // only 8-byte aligned (a strand's key region starts anywhere)
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 32; c++) w[c] = (uint32_t)kp[c];
                transpose32(w);
#pragma unroll
                for (int pbit = 0; pbit < 32; pbit++) stage[pbit * kBsStage + lane] = w[31 - pbit];
#pragma unroll
                for (int c = 0; c < 32; c++) w[c] = (uint32_t)(kp[c] >> 32);
                transpose32(w);
#pragma unroll
                for (int pbit = 0; pbit < 32; pbit++) stage[(32 + pbit) * kBsStage + lane] = w[31 - pbit];
            }
            __syncwarp();
        }
        // inject: lane 31 picks up bundle t, then everything rotates one lane up (lane 0 <- lane 31)
        if (lane == 31 && t < nb) {
#pragma unroll
            for (int i = 0; i < 64; i++) R[i] = stage[i * kBsStage + (t & (kBsStage - 1))];
        }
#pragma unroll
        for (int i = 0; i < 64; i++) R[i] = __shfl_sync(kFull, R[i], (lane + 31) & 31);
        const int bi = t - lane;                       // bundle this lane now holds
        const bool valid = bi >= 0 && bi < nb;
#pragma unroll 1
        for (int b = 0; b < B; b++) {
            bs_step(R);
            // candidates: chains whose sign-biased value has its top tap_bits all zero
            const uint32_t tap = st.tap[b * 32];
            const uint32_t o1 = ~R[63] | R[62] | R[61] | R[60];
            const uint32_t o2 = o1 | R[59] | R[58] | R[57] | R[56];
            const uint32_t o3 = o2 | R[55] | R[54];
            const uint32_t o4 = o3 | R[53] | R[52];
            const uint32_t o5 = o4 | R[51] | R[50];
            const uint32_t o6 = o5 | R[49] | R[48];
            const uint32_t o7 = o6 | R[47] | R[46];
            const uint32_t o8 = o7 | R[45] | R[44];
            const uint32_t lo4 = (tap & 2) ? ((tap & 1) ? o3 : o2) : ((tap & 1) ? o1 : 0u);
            const uint32_t hi4 = (tap & 2) ? ((tap & 1) ? o7 : o6) : ((tap & 1) ? o5 : o4);
            const uint32_t o = (tap & 8) ? o8 : ((tap & 4) ? hi4 : lo4);
            uint32_t cand = valid ? ~o : 0u;
            unsigned evm = __ballot_sync(kFull, cand != 0);
            while (evm) {                               // warp-uniform
                const int L = __ffs(evm) - 1;
                evm &= evm - 1;
                // high planes first: about half of the flagged chains are rejected on the high word alone
                if (lane == L) {
#pragma unroll
                    for (int i = 0; i < 32; i++) scratch[32 + i] = R[32 + i];
                }
                __syncwarp();
                const uint32_t p_hi = scratch[32 + lane];
                uint32_t cm = __shfl_sync(kFull, cand, L);
                const int bL = __shfl_sync(kFull, bi, L);
                const int32_t bhL = __shfl_sync(kFull, (int32_t)st.hi[b * 32], L);
                uint32_t keep = 0;                      // flagged chains that pass the high-word test
                for (uint32_t c2 = cm; c2; c2 &= c2 - 1) {
                    const int sb = __ffs(c2) - 1;
                    const uint32_t xh = __ballot_sync(kFull, (p_hi >> sb) & 1u);
                    if ((int32_t)xh <= bhL) keep |= 1u << sb;
                }
                cm = keep;
                if (cm) {
                    if (lane == L) {
#pragma unroll
                        for (int i = 0; i < 32; i++) scratch[i] = R[i];
                    }
                    __syncwarp();
                }
                const uint32_t p_lo = cm ? scratch[lane] : 0u;
                while (cm) {                            // warp-uniform
                    const int sb = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const uint32_t xh = __ballot_sync(kFull, (p_hi >> sb) & 1u);
                    const uint32_t xl = __ballot_sync(kFull, (p_lo >> sb) & 1u);
                    if (lane == L) {
                        const int32_t bh = (int32_t)st.hi[b * 32];
                        if ((int32_t)xh < bh || ((int32_t)xh == bh && xl < st.lo[b * 32])) {   // MinHashSketch.java:144
                            const uint64_t key = keys[(size_t)bL * 32 + (31 - sb)];
                            st.hi[b * 32] = xh; st.lo[b * 32] = xl;
                            st.out[b * 32] = ((lane * B + b) & 1) ? (uint32_t)(key >> 32) : (uint32_t)key;   // :146-149
                            st.tap[b * 32] = bs_tap_of(xh);
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
}

#endif

// ---- lock-step variant of the bit-sliced phase (the default) --------------------------------------
// Instead of handing bundles from lane to lane, every lane keeps its own bundle for the whole chain and the
// 32 lanes walk the H words in lock step, so at any moment the whole warp works on ONE word: its exact
// minimum (and the filter depth derived from it) is a single warp-uniform value read by a broadcast LDS.
// That removes the 64 shuffles per hop, the pipeline fill/drain (31 of 311 hops for a 10 kbp strand) and the
// per-lane tap select (the OR over the top planes is a fall-through switch on a uniform value, 1-bit
// granularity).  Flagged chains are resolved one at a time against the shared exact state as before.
struct BsShared { uint32_t *hi, *lo, *out, *depth; };   // [Hpad] each, indexed by word; depth holds the filter code (bs_filter.cuh)

// shared-window accesses by 32-bit address (the lock-step word loop carries one such address; everything the rare path
// touches is at a constant offset from it, so nothing has to be re-derived from %tid inside the loop)
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
template <int OFF>   // immediate offset: one address register serves all sixteen stores of a publish
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0+%5], {%1, %2, %3, %4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_off(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory"); return v; }

// MULTI: every key advances light_w (> 1) steps per word, each of them compared (a uniform tf-idf weight, MinHashSketch.java:138)
template <bool MULTI>
__device__ __forceinline__ void bs_phase_lockstep(const BsShared &st, int H, const uint64_t *__restrict__ keys /* 32*nb chain starts */,
                                                  const uint64_t *__restrict__ okeys /* the k-mer hashes behind them */, int nb,
                                                  uint32_t hpb /* bytes between the per-word arrays hi|lo|out|depth|scratch */, int lane, int light_w)
{
    uint32_t R[64];
    for (int r0 = 0; r0 < nb; r0 += 32) {
        const int bi = r0 + lane;
        const bool valid = bi < nb;
        if (valid) {   // load this lane's 32 keys and transpose them into bit planes (low words, then high words)
            const uint64_t *kp = keys + (size_t)bi * 32;
            uint32_t w[32];
#pragma unroll
            for (int c = 0; c < 32; c++) w[c] = (uint32_t)kp[c];
            transpose32(w);
#pragma unroll
            for (int pbit = 0; pbit < 32; pbit++) R[pbit] = w[31 - pbit];
#pragma unroll
            for (int c = 0; c < 32; c++) w[c] = (uint32_t)(kp[c] >> 32);
            transpose32(w);
#pragma unroll
            for (int pbit = 0; pbit < 32; pbit++) R[32 + pbit] = w[31 - pbit];
        } else {
            // a lane without a bundle carries 32 all-zero chains: 0 is a fixed point of the XORShift map and, sign-biased,
            // has its top bit set, so these chains pass no prefix filter of depth >= 1 and need no mask on the hot path
            // (depth 0 flags everything; the rare path below drops them)
#pragma unroll
            for (int i = 0; i < 64; i++) R[i] = 0;
        }
        // the word loop carries a shared-window address and a countdown instead of (base + 4*wd, wd < H): written as C
        // the compiler re-derives the base from %tid every iteration (three S2R and six integer ops per word step)
        uint32_t daddr = (uint32_t)__cvta_generic_to_shared(st.depth);
        asm volatile("" : "+r"(daddr));   // opaque: otherwise daddr - 4*wd is folded back into the %tid expression
        int rep = light_w;           // MULTI: steps left in the current word
#pragma unroll 1
        for (int left = H; left > 0;) {
            bs_step(R);
            // chains whose sign-biased value has its top `depth` bits all zero (warp-uniform depth)
            uint32_t cand;
            const int wd = H - left;
            cand = bs_prefix_filter(R, lds_u32(daddr));
            if (__any_sync(kFull, cand != 0)) {
            if (!valid) cand = 0;
            unsigned evm = __ballot_sync(kFull, cand != 0);
            // per-word state of this word: hi | lo | out | code at daddr - 3,2,1,0 * hpb; the warp's scratch after the code array
            const uint32_t sbase = daddr - 4u * (uint32_t)wd + hpb;
            while (evm) {                               // warp-uniform
                const int L = __ffs(evm) - 1;
                evm &= evm - 1;
                // the filter compares the top depth+3 bits exactly, so ~9 of 10 flagged chains are real updates: the lane
                // publishes all 64 planes at once and every flagged chain is re-assembled with two ballots
                if (lane == L) {
                    sts_v4<0>(sbase, R[0], R[1], R[2], R[3]);       sts_v4<16>(sbase, R[4], R[5], R[6], R[7]);
                    sts_v4<32>(sbase, R[8], R[9], R[10], R[11]);    sts_v4<48>(sbase, R[12], R[13], R[14], R[15]);
                    sts_v4<64>(sbase, R[16], R[17], R[18], R[19]);  sts_v4<80>(sbase, R[20], R[21], R[22], R[23]);
                    sts_v4<96>(sbase, R[24], R[25], R[26], R[27]);  sts_v4<112>(sbase, R[28], R[29], R[30], R[31]);
                    sts_v4<128>(sbase, R[32], R[33], R[34], R[35]); sts_v4<144>(sbase, R[36], R[37], R[38], R[39]);
                    sts_v4<160>(sbase, R[40], R[41], R[42], R[43]); sts_v4<176>(sbase, R[44], R[45], R[46], R[47]);
                    sts_v4<192>(sbase, R[48], R[49], R[50], R[51]); sts_v4<208>(sbase, R[52], R[53], R[54], R[55]);
                    sts_v4<224>(sbase, R[56], R[57], R[58], R[59]); sts_v4<240>(sbase, R[60], R[61], R[62], R[63]);
                }
                __syncwarp();
                const uint32_t p_lo = lds_u32_off<0>(sbase + 4u * lane), p_hi = lds_u32_off<128>(sbase + 4u * lane);
                uint32_t cm = __shfl_sync(kFull, cand, L);
                while (cm) {                            // warp-uniform
                    const int sb = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const uint32_t xh = __ballot_sync(kFull, (p_hi >> sb) & 1u);
                    const uint32_t xl = __ballot_sync(kFull, (p_lo >> sb) & 1u);
                    const int32_t bh = (int32_t)lds_u32(daddr - 3u * hpb);
                    if ((int32_t)xh < bh || ((int32_t)xh == bh && xl < lds_u32(daddr - 2u * hpb))) {   // MinHashSketch.java:144, uniform
                        __syncwarp();
                        if (lane == L) {
                            const uint64_t key = okeys[(size_t)(r0 + L) * 32 + (31 - sb)];
                            sts_u32(daddr - 3u * hpb, xh); sts_u32(daddr - 2u * hpb, xl);
                            sts_u32(daddr - hpb, (wd & 1) ? (uint32_t)(key >> 32) : (uint32_t)key);   // :146-149
                            sts_u32(daddr, bs_code_of(xh));
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
            }
            if constexpr (MULTI) {
                if (--rep > 0) continue;
                rep = light_w;
            }
            left--; daddr += 4;
        }
    }
}

// virt != 0: the descriptors are VIRTUAL strands, one per block of <= 512 words of a sketch wider than 512 (--num-hashes 1024:
// two per strand).  Block p starts its chains from the k-mer hashes advanced by 512*p*weight steps (k_advance_keys), so it
// is an ordinary 512-word job that runs in the B = 16 instantiation (94 registers, 20 warps per SM) instead of the B = 32 one
// (158 registers, 12 warps per SM, int_issue 0.56).  Field use in a virtual descriptor: koff = the advanced keys, base_off =
// offset of the strand's hashes and weights, rc = words in this block, slot = first word, row as usual; nlight / nheavy are
// those of strand s % n_real.
template <int B, bool MULTI, bool virt>
__global__ void __launch_bounds__(128, B <= 16 ? 5 : (B <= 32 ? 3 : 1))
k_minhash_bs2(const StrandDesc *__restrict__ desc, int n_strands, int k, int H, SketchScratch sc,
              int32_t *__restrict__ minhash, uint32_t *queue, int scalar_keys, int light_w, int n_real, int hstride)
{
    // per warp: state [4][B*32] | scratch [64] ; static: key rings + weights for the scalar pipeline
    extern __shared__ __align__(16) uint32_t s_dyn[];
    __shared__ uint64_t s_kbuf[4][64], s_xbuf[virt ? 4 : 1][64];
    __shared__ uint32_t s_wbuf[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int HP = B * 32;
    constexpr int kPerWarp = 4 * HP + 64;
    uint32_t *wbase = s_dyn + (size_t)wib * kPerWarp;
    BsShared st;
    st.hi = wbase; st.lo = wbase + HP; st.out = wbase + 2 * HP; st.depth = wbase + 3 * HP;
    for (;;) {
        int s = 0;
        if (lane == 0) s = (int)atomicAdd(queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= n_strands) break;
        const StrandDesc d = desc[s];
        const int nk = (int)d.len - k + 1;
        const int hloc = virt ? (int)d.rc : H;                               // words of this job
        const int si = virt ? s % n_real : s;
        const uint64_t *keys = sc.keys + d.koff;                             // chain starts
        const uint64_t *okeys = virt ? sc.keys + d.base_off : keys;          // the k-mer hashes
        const uint32_t *wts = sc.wts + (virt ? d.base_off : d.koff);
        const int nl = sc.nlight[si], nh = sc.nheavy[si];
        // full bundles, taken from the end; at least scalar_keys keys stay scalar, and when more than one warp-wide set
        // of 32 bundles is available only whole sets are taken (a partial set costs a whole set's plane steps), so
        // between scalar_keys and scalar_keys + 1023 keys go through the scalar pipeline
        int nb = nl > scalar_keys ? (nl - scalar_keys) / 32 : 0;
        if (nb >= 32) nb &= ~31;
        const int n_sc = nl - 32 * nb;
        LaneMins<B> m;
#pragma unroll
        for (int b = 0; b < B; b++) { m.hi[b] = 0x7fffffff; m.lo[b] = 0xffffffffu; m.out[b] = 0; }   // Long.MAX_VALUE
        uint64_t *xr = s_xbuf[virt ? wib : 0];
        if constexpr (!MULTI) minhash_pipeline<B, false, virt>(m, okeys, keys, nullptr, n_sc, +1, s_kbuf[wib], xr, s_wbuf[wib], lane);
        else                  minhash_pipeline<B, true, virt>(m, okeys, keys, nullptr, n_sc, +1, s_kbuf[wib], xr, s_wbuf[wib], lane, (uint32_t)light_w);
        if (nh > 0)
            minhash_pipeline<B, true, virt>(m, okeys + (nk - 1), keys + (nk - 1), wts + (nk - 1), nh, -1, s_kbuf[wib], xr, s_wbuf[wib], lane);
        int32_t *row = minhash + (size_t)d.row * hstride + (virt ? (int)d.slot : 0);
        if (nb > 0) {
            __syncwarp();
#pragma unroll
            for (int b = 0; b < B; b++) {
                const int word = lane * B + b;
                st.hi[word] = (uint32_t)m.hi[b]; st.lo[word] = m.lo[b]; st.out[word] = (uint32_t)m.out[b];
                st.depth[word] = bs_code_of((uint32_t)m.hi[b]);
            }
            __syncwarp();
            bs_phase_lockstep<MULTI>(st, hloc, keys + n_sc, okeys + n_sc, nb, (uint32_t)HP * 4u, lane, light_w);
            __syncwarp();
            for (int word = lane; word < hloc; word += 32) row[word] = (int32_t)st.out[word];
            __syncwarp();
        } else {
#pragma unroll
            for (int b = 0; b < B; b++) { int word = lane * B + b; if (word < hloc) row[word] = m.out[b]; }
        }
    }
}

// Chain states for the later word blocks of a wide sketch: block p of a key of weight w starts at step^(512*p*w)(hash).  step^512
// is a fixed linear map over GF(2): eight table look-ups (one per byte of the state) and seven XORs (c_t512, built on the host).
__global__ void k_advance_keys(const StrandDesc *__restrict__ desc, int n_strands, int k, SketchScratch sc, const uint64_t *__restrict__ t512,
                               uint64_t total_k, int passes, uint32_t light_w)
{
    for (int s = blockIdx.x; s < n_strands; s += gridDim.x) {
        const StrandDesc d = desc[s];
        const int nk = (int)d.len - k + 1;
        const int nl = sc.nlight[s], nh = sc.nheavy[s];
        for (int j = threadIdx.x; j < nl + nh; j += blockDim.x) {
            const size_t idx = d.koff + (j < nl ? (size_t)j : (size_t)(nk - 1 - (j - nl)));
            const uint32_t w = j < nl ? light_w : sc.wts[idx];
            uint64_t x = sc.keys[idx];
            for (int p = 1; p < passes; p++) {
                for (uint32_t t = 0; t < w; t++) {
                    uint64_t y = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) y ^= __ldg(&t512[b * 256 + ((x >> (8 * b)) & 0xff)]);
                    x = y;
                }
                sc.keys[(size_t)p * total_k + idx] = x;
            }
        }
    }
}

#ifdef MHAPB_AB_KERNELS
template <int B>
__global__ void __launch_bounds__(128, B <= 16 ? 4 : 2)
k_minhash_bs(const StrandDesc *__restrict__ desc, int n_strands, int k, int H, SketchScratch sc,
             int32_t *__restrict__ minhash, uint32_t *queue, int scalar_keys)
{
    // per warp: state [4][B][32] | stage [64][32] | scratch [64] ; static: key ring + weights for the scalar pipeline
    extern __shared__ __align__(16) uint32_t s_dyn[];
    __shared__ uint64_t s_kbuf[4][64];
    __shared__ uint32_t s_wbuf[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int kPerWarp = 4 * B * 32 + 64 * kBsStage + 64;
    uint32_t *wbase = s_dyn + (size_t)wib * kPerWarp;
    BsState st;
    st.hi = wbase + lane; st.lo = wbase + B * 32 + lane; st.out = wbase + 2 * B * 32 + lane; st.tap = wbase + 3 * B * 32 + lane;
    uint32_t *stage = wbase + 4 * B * 32, *scratch = stage + 64 * kBsStage;
    for (;;) {
        int s = 0;
        if (lane == 0) s = (int)atomicAdd(queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= n_strands) break;
        const StrandDesc d = desc[s];
        const int nk = (int)d.len - k + 1;
        const uint64_t *keys = sc.keys + d.koff;
        const int nl = sc.nlight[s], nh = sc.nheavy[s];
        const int nb = nl > scalar_keys ? (nl - scalar_keys) / 32 : 0;   // full bundles, taken from the end
        const int n_sc = nl - 32 * nb;
        LaneMins<B> m;
#pragma unroll
        for (int b = 0; b < B; b++) { m.hi[b] = 0x7fffffff; m.lo[b] = 0xffffffffu; m.out[b] = 0; }   // Long.MAX_VALUE
        minhash_pipeline<B, false>(m, keys, keys, nullptr, n_sc, +1, s_kbuf[wib], s_kbuf[wib], s_wbuf[wib], lane);
        if (nh > 0)
            minhash_pipeline<B, true>(m, keys + (nk - 1), keys + (nk - 1), sc.wts + d.koff + (nk - 1), nh, -1, s_kbuf[wib], s_kbuf[wib], s_wbuf[wib], lane);
        int32_t *row = minhash + (size_t)d.row * H;
        if (nb > 0) {
#pragma unroll
            for (int b = 0; b < B; b++) {
                st.hi[b * 32] = (uint32_t)m.hi[b]; st.lo[b * 32] = m.lo[b]; st.out[b * 32] = (uint32_t)m.out[b];
                st.tap[b * 32] = bs_tap_of((uint32_t)m.hi[b]);
            }
            __syncwarp();
            bs_phase<B>(st, keys + n_sc, nb, stage, scratch, lane);
            __syncwarp();
#pragma unroll
            for (int b = 0; b < B; b++) { int word = lane * B + b; if (word < H) row[word] = (int32_t)st.out[b * 32]; }
            __syncwarp();
        } else {
#pragma unroll
            for (int b = 0; b < B; b++) { int word = lane * B + b; if (word < H) row[word] = m.out[b]; }
        }
    }
}

#endif

// ---------------------------------------------------------------------------------------------
// K1c: ordered bottom-S sketch
// ---------------------------------------------------------------------------------------------
// One CTA per strand.  Hash every ordered k-mer, radix-select the S smallest 64-bit keys
// (hash biased to unsigned order << 32 | position) -- which is exactly "ascending signed hash,
// ties by ascending position" of fastutil's stable radixSortIndirect -- then bitonic-sort the
// S survivors in shared memory.
constexpr int kOrdBinBits = 11, kOrdBins = 1 << kOrdBinBits;   // buckets of the select/sort histogram (top bits of the hash)
constexpr int kOrdMaxBin = 64;                                  // a fuller bucket among the selected ones -> generic path

template <bool LONG, int KC /* compile-time ordered k, 0 = runtime */>
__global__ void __launch_bounds__(512)
k_ordered(const uint8_t *__restrict__ bases, const StrandDesc *__restrict__ desc, int s_begin, int s_end, int ok, int S,
          int ord_stride, uint32_t len_cap, uint32_t sel_cap, SketchScratch sc, int32_t *__restrict__ ord,
          int32_t *__restrict__ ord_n, uint32_t *queue)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_strand, s_nsel;
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_digit, s_before, s_bin, s_bstar, s_maxbin, s_nbl, s_wsum[16];
    __shared__ uint64_t s_T, s_bl[kOrdMaxBin];
    __shared__ uint8_t s_lut[2][256];
    if (!LONG) stage_lut_init(s_lut);

    uint64_t *sel = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *oh;
    uint32_t *hist2k = nullptr;
    uint8_t *chars = nullptr;
    if (LONG) oh = sc.ohash + (size_t)blockIdx.x * len_cap;
    else {
        hist2k = reinterpret_cast<uint32_t *>(smem_raw + (size_t)sel_cap * 8);
        oh = hist2k + kOrdBins;
        chars = reinterpret_cast<uint8_t *>(oh + len_cap);
    }

    for (;;) {
        if (threadIdx.x == 0) { s_strand = s_begin + (int)atomicAdd(queue, 1u); s_nsel = 0; }
        __syncthreads();
        const int s = s_strand;
        if (s >= s_end) break;
        const StrandDesc d = desc[s];
        const int no = (int)d.len - ok + 1;
        if (!LONG) { stage_chars(chars, bases + d.base_off, d.len, d.rc, s_lut); __syncthreads(); }
        const CharsGlobal gsrc{bases + d.base_off, d.len, d.rc};
        if constexpr (!LONG && KC > 0 && (KC & 1) == 0) {
            // MurmurHash3_x86_32 block b of k-mer i is the pair of chars (i+2b, i+2b+1), and its mixed value
            // rotl(P*c1,15)*c2 depends on the pair alone -- the same pair serves KC/2 overlapping k-mers.  Every thread
            // hashes a RUN of consecutive k-mers with a sliding window of KC mixed pairs in registers: one new pair
            // mix (and one char load) per k-mer instead of KC/2 mixes and KC loads.
            const int run = (no + (int)blockDim.x - 1) / (int)blockDim.x;
            const int i0 = (int)threadIdx.x * run, i1 = min(no, i0 + run);
            if (i0 < i1) {
                auto mixp = [](uint32_t p) { p *= 0xcc9e2d51u; p = rotl32(p, 15); return p * 0x1b873593u; };
                uint32_t W[KC];                          // W[(u+q) % KC] = mix(pair(i+q)) while hashing k-mer i = base+u
                uint32_t cprev = chars[i0];
#pragma unroll
                for (int q = 0; q < KC - 1; q++) {       // pairs i0 .. i0+KC-2 all lie inside k-mer i0
                    const uint32_t cn = chars[i0 + q + 1];
                    W[q] = mixp(cprev | (cn << 16));
                    cprev = cn;
                }
                for (int base = i0; base < i1; base += KC) {
#pragma unroll
                    for (int u = 0; u < KC; u++) {
                        const int i = base + u;
                        if (i < i1) {
                            uint32_t h = 0;
#pragma unroll
                            for (int b = 0; b < KC / 2; b++) { h ^= W[(u + 2 * b) % KC]; h = rotl32(h, 13); h = h * 5 + 0xe6546b64u; }
                            h ^= (uint32_t)(2 * KC);
                            h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
                            oh[i] = h ^ 0x80000000u;     // signed order -> unsigned order
                            // slide: pair (i+KC-1, i+KC) is the last pair of k-mer i+1 (cprev = char i+KC-1; the staged
                            // buffer is padded, a read past the strand only feeds k-mers that do not exist)
                            const uint32_t cn = chars[i + KC];
                            W[(u + KC - 1) % KC] = mixp(cprev | (cn << 16));
                            cprev = cn;
                        }
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < no; i += blockDim.x) {
                uint32_t h;
                if (LONG) h = murmur3_32_chars([&](int j) { return gsrc(i + j); }, KC ? KC : ok);
                else      h = murmur3_32_chars([&](int j) { return chars[i + j]; }, KC ? KC : ok);
                oh[i] = h ^ 0x80000000u;   // signed order -> unsigned order
            }
        }
        __syncthreads();

        const int nsel = min(S, no);
        bool generic = LONG;
        if constexpr (!LONG) {
            // ---- bucket select + bucket sort (the common case) -------------------------------------------------------
            // ONE histogram over the top 11 bits of the (order-biased) hash serves both steps: its running sum locates the
            // bucket b* in which the S-th smallest key falls (everything below is selected, b* itself is split exactly by
            // ranking its few keys), and the same running sum is where each bucket starts in the sorted output, so the
            // selected keys are scattered straight to their bucket and every bucket (2-5 keys for a 10 kbp strand) is
            // finished by one thread with an insertion sort.  This replaced an 8-bit radix select (2-3 passes over all
            // hashes) + compaction + a 66-stage bitonic sort of 2048 keys, which was 36 % of the kernel's instructions.
            // Degenerate strands (low complexity: hundreds of equal hashes in one bucket) take the generic path below.
            uint32_t *cur = hist2k;
            for (int i = threadIdx.x; i < kOrdBins; i += blockDim.x) cur[i] = 0;
            if (threadIdx.x == 0) { s_bstar = kOrdBins; s_before = 0; s_bin = 0; s_maxbin = 0; s_nbl = 0; }
            __syncthreads();
            for (int i = threadIdx.x; i < no; i += blockDim.x) atomicAdd(&cur[oh[i] >> (32 - kOrdBinBits)], 1u);
            __syncthreads();
            {   // exclusive scan over the buckets, kOrdBins / blockDim.x per thread
                constexpr int PER = kOrdBins / 512;
                uint32_t c[PER], sum = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) { c[q] = cur[threadIdx.x * PER + q]; sum += c[q]; }
                uint32_t incl = sum;
                const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
                if (lane == 31) s_wsum[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    uint32_t x = lane < 16 ? s_wsum[lane] : 0, xi = x;
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) { uint32_t v = __shfl_up_sync(kFull, xi, o); if (lane >= o) xi += v; }
                    if (lane < 16) s_wsum[lane] = xi - x;
                }
                __syncthreads();
                uint32_t run = incl - sum + s_wsum[wid];
                uint32_t mx = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const uint32_t before = run;
                    run += c[q];
                    cur[threadIdx.x * PER + q] = before;                               // bucket start = scatter cursor
                    if (before < (uint32_t)nsel) mx = max(mx, c[q]);                    // buckets that receive selected keys
                    if (before < (uint32_t)nsel && (uint32_t)nsel <= run && no > nsel) { s_bstar = threadIdx.x * PER + q; s_before = before; s_bin = c[q]; }
                }
                if (mx > kOrdMaxBin) atomicMax(&s_maxbin, mx);
            }
            __syncthreads();
            const uint32_t bstar = s_bstar;
            generic = s_maxbin > kOrdMaxBin;
            uint64_t Tb = ~0ull;                    // keys of bucket b* are selected iff <= Tb
            if (!generic && bstar < (uint32_t)kOrdBins && s_bin > (uint32_t)nsel - s_before) {
                // split b* exactly: its keys (<= kOrdMaxBin of them), ranked by (hash, position)
                for (int i = threadIdx.x; i < no; i += blockDim.x)
                    if ((oh[i] >> (32 - kOrdBinBits)) == bstar) { const uint32_t p = atomicAdd(&s_nbl, 1u); s_bl[p] = ((uint64_t)oh[i] << 32) | (uint32_t)i; }
                __syncthreads();
                const uint32_t nb_ = s_nbl, need = (uint32_t)nsel - s_before;
                if (threadIdx.x < nb_) {
                    const uint64_t me = s_bl[threadIdx.x];
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < nb_; j++) rank += s_bl[j] < me;
                    if (rank == need - 1) s_T = me;
                }
                __syncthreads();
                Tb = s_T;
            }
            if (!generic) {
                for (int i = threadIdx.x; i < no; i += blockDim.x) {
                    const uint32_t h = oh[i], bin = h >> (32 - kOrdBinBits);
                    const uint64_t key = ((uint64_t)h << 32) | (uint32_t)i;
                    if (bin < bstar || (bin == bstar && key <= Tb)) sel[atomicAdd(&cur[bin], 1u)] = key;
                }
                __syncthreads();
                // cur[b] is now the END of bucket b (b <= b*), its start the end of the bucket before
                const uint32_t last = min(bstar, (uint32_t)kOrdBins - 1);
                for (uint32_t bkt = threadIdx.x; bkt <= last; bkt += blockDim.x) {
                    const uint32_t beg = bkt ? cur[bkt - 1] : 0u, end = min(cur[bkt], (uint32_t)nsel);
                    for (uint32_t x = beg + 1; x < end; x++) {
                        const uint64_t v = sel[x];
                        uint32_t y = x;
                        while (y > beg && sel[y - 1] > v) { sel[y] = sel[y - 1]; y--; }
                        sel[y] = v;
                    }
                }
                __syncthreads();
            }
        }
        if (generic) {
        uint64_t T = ~0ull;   // select keys <= T
        if (no > S) {
            uint64_t prefix = 0, mask = 0;
            uint32_t remaining = (uint32_t)S;
            for (int pass = 7; pass >= 0; pass--) {
                const int shift = pass * 8;
                if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
                __syncthreads();
                for (int i = threadIdx.x; i < no; i += blockDim.x) {
                    uint64_t key = ((uint64_t)oh[i] << 32) | (uint32_t)i;
                    if ((key & mask) == prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (threadIdx.x < 32) {
                    // lane l owns bins [8l, 8l+8): find the bin where the running count crosses `remaining`
                    uint32_t c[8], sum = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) { c[q] = s_hist[threadIdx.x * 8 + q]; sum += c[q]; }
                    uint32_t incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(kFull, incl, o); if ((int)threadIdx.x >= o) incl += v; }
                    uint32_t excl = incl - sum;
                    if (excl < remaining && remaining <= incl) {
                        uint32_t run = excl;
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            if (run < remaining && remaining <= run + c[q]) { s_digit = threadIdx.x * 8 + q; s_before = run; s_bin = c[q]; }
                            run += c[q];
                        }
                    }
                }
                __syncthreads();
                prefix |= (uint64_t)s_digit << shift;
                mask |= 0xffull << shift;
                remaining -= s_before;
                if (s_bin == remaining) {   // the whole bin is selected: no need to split it further
                    T = prefix | ((shift == 0) ? 0ull : ((1ull << shift) - 1));
                    break;
                }
                // (after pass 0 the bin holds exactly one key, so the branch above always fires)
            }
        }
        __syncthreads();
        const int nor = (no + 31) & ~31;
        for (int i = threadIdx.x; i < nor; i += blockDim.x) {
            uint64_t key = (i < no) ? (((uint64_t)oh[i] << 32) | (uint32_t)i) : ~0ull;
            bool want = (i < no) && key <= T;
            int p = warp_alloc(&s_nsel, want);
            if (want) sel[p] = key;
        }
        uint32_t P = 1; while ((int)P < nsel) P <<= 1;
        __syncthreads();
        for (uint32_t i = nsel + threadIdx.x; i < P; i += blockDim.x) sel[i] = ~0ull;
        __syncthreads();
        for (uint32_t size = 2; size <= P; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                    uint32_t i = 2 * t - (t & (stride - 1));
                    uint32_t j = i + stride;
                    bool up = (i & size) == 0;
                    uint64_t a = sel[i], b = sel[j];
                    if ((a > b) == up) { sel[i] = b; sel[j] = a; }
                }
                __syncthreads();
            }
        }
        }
        int2 *row = reinterpret_cast<int2 *>(ord) + (size_t)d.row * ord_stride;
        for (int i = threadIdx.x; i < nsel; i += blockDim.x) {
            uint64_t key = sel[i];
            row[i] = make_int2((int32_t)((uint32_t)(key >> 32) ^ 0x80000000u), (int32_t)(uint32_t)key);
        }
        if (threadIdx.x == 0) ord_n[d.row] = nsel;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// roofline denominator for K1b: the same pipe-balanced recurrence with nothing else (no compare,
// no memory, 4 independent chains/thread)
// ---------------------------------------------------------------------------------------------
constexpr int kPeakIters = 4096, kPeakIlp = 4;
__global__ void __launch_bounds__(256) k_xorshift_peak(unsigned long long *sink)
{
    uint32_t xl[kPeakIlp], xh[kPeakIlp];
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) {
        uint64_t x = 0x9E3779B97F4A7C15ull * (blockIdx.x * 256ull + threadIdx.x + 1) + i;
        xl[i] = (uint32_t)x; xh[i] = (uint32_t)(x >> 32);
    }
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < kPeakIlp; i++) xorshift_step32(xl[i], xh[i]);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) acc ^= xl[i] ^ xh[i];
    if (acc == 0x1234567u) atomicAdd(sink, 1ull);   // keeps the chains live
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count()
{
    if (!g_sm_count) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

static constexpr uint32_t kShortTableCap = (kShortMaxKmers + kShortMaxKmers / 4 + 8) | 1;   // 20489 slots = 160 KB

// the bit-sliced recurrence alone (bs_step.cuh: 92 three-input XORs per 32 chain steps), nothing else
__global__ void __launch_bounds__(256) k_xorshift_peak_bs(unsigned long long *sink)
{
    uint32_t R[64];
#pragma unroll
    for (int i = 0; i < 64; i++) R[i] = 0x9E3779B9u * (blockIdx.x * 256u + threadIdx.x + 1u) + i * 0x85EBCA6Bu;
    for (int it = 0; it < kPeakIters / 2; it++) {
#pragma unroll 1
        for (int u = 0; u < 8; u++) bs_step(R);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) acc ^= R[i];
    if (acc == 0x1234567u) atomicAdd(sink, 1ull);
}

cudaError_t launch_xorshift_peak_bs(cudaStream_t st, unsigned long long *d_sink, double *steps)
{
    const int grid = sm_count() * 2;
    k_xorshift_peak_bs<<<grid, 256, 0, st>>>(d_sink);
    *steps = (double)grid * 256.0 * (kPeakIters / 2) * 8.0 * 32.0;
    return cudaGetLastError();
}

cudaError_t launch_xorshift_peak(cudaStream_t st, unsigned long long *d_sink, double *steps)
{
    const int grid = sm_count() * 8;
    k_xorshift_peak<<<grid, 256, 0, st>>>(d_sink);
    *steps = (double)grid * 256.0 * kPeakIters * 8.0 * kPeakIlp;
    return cudaGetLastError();
}

int hash_dedup_grid() { return sm_count(); }          // 1 CTA/SM (shared-memory table)
int ordered_grid() { return sm_count() * 2; }
size_t dedup_table_cap_short() { return kShortTableCap; }

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

cudaError_t launch_hash_dedup(cudaStream_t st, const uint8_t *d_bases, const StrandDesc *d_desc, int s_base, int n_strands,
                              int first_long, int max_kmers_short, int max_kmers_long, int k, int unweighted,
                              const KmerFilterView &filter, const SketchScratch &sc, uint32_t *queues, int *launches)
{
    cudaError_t e;
    if (first_long > 0) {
        uint32_t cap = dedup_table_slots((uint32_t)max_kmers_short);
        uint32_t chars_cap = (uint32_t)align16((size_t)max_kmers_short + k);
        // shared window: the BITMAP scheme's plan, and room for the TABLE scheme a low-complexity strand falls back to
        const DedupPlan plan = dedup_plan((uint32_t)max_kmers_short, chars_cap);
        size_t smem = std::max<size_t>(plan.total, (size_t)cap * 8 + (size_t)((cap + 31) / 32) * 4);
        auto kern = k == 16 ? k_hash_dedup<false, 16> : k_hash_dedup<false, 0>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int grid = hash_dedup_grid();
        if (smem <= 113 * 1024) grid *= 2;
        if (grid > first_long) grid = first_long;
        // dupcnt rows are strided by the *launch's* cap so both variants can share the buffer
        kern<<<grid, 1024, smem, st>>>(d_bases, d_desc, s_base, s_base + first_long, k, unweighted, cap, chars_cap, (uint32_t)max_kmers_short, sc, queues + 0, filter);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (first_long < n_strands) {
        uint32_t cap = dedup_table_slots((uint32_t)max_kmers_long);
        int grid = hash_dedup_grid();
        if (grid > n_strands - first_long) grid = n_strands - first_long;
        k_hash_dedup<true, 0><<<grid, 1024, 0, st>>>(d_bases, d_desc, s_base + first_long, s_base + n_strands, k, unweighted, cap, 0, 0, sc, queues + 1, filter);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static int k1b_variant()
{
    // MHAPB_K1B=scalar selects the scalar systolic kernel, =bitsliced-systolic the first bit-sliced kernel (both kept
    // for A/B measurements); default: bit-sliced lock-step
    static int v = -1;
    if (v < 0) { const char *e = getenv("MHAPB_K1B"); v = (e && e[0] == 's') ? 0 : (e && e[0] == 'b') ? 1 : 2; }
    return v;
}

template <int B>
static cudaError_t launch_minhash_b(cudaStream_t st, const StrandDesc *d_desc, int n_strands, int k, int H,
                                    const SketchScratch &sc, int32_t *d_minhash, uint32_t light_w, uint32_t *queue,
                                    int virt = 0, int n_real = 0, int hstride = 0)
{
    if (!hstride) hstride = H;
    cudaError_t e;
    int per_sm = 0;
    // the two A/B variants (MHAPB_K1B) only implement light weight 1; a uniform tf-idf weight takes the default kernels
    const int variant = light_w == 1 ? k1b_variant() : 2;
    if (variant == 2 && B <= 32) {
        const size_t smem = (size_t)4 * (4 * B * 32 + 64) * 4;
        auto kern = virt ? (light_w == 1 ? k_minhash_bs2<B, false, B == 16> : k_minhash_bs2<B, true, B == 16>)
                         : (light_w == 1 ? k_minhash_bs2<B, false, false> : k_minhash_bs2<B, true, false>);
        if (virt && B != 16) return cudaErrorInvalidValue;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int grid = sm_count() * per_sm;
        int need = (n_strands + 3) / 4;
        if (grid > need) grid = need;
        static int scalar_keys2 = -1;
        if (scalar_keys2 < 0) { const char *ev = getenv("MHAPB_BS_SCALAR_KEYS"); scalar_keys2 = ev ? atoi(ev) : kBsScalarKeys; if (scalar_keys2 < 0) scalar_keys2 = 0; }
        kern<<<grid, 128, smem, st>>>(d_desc, n_strands, k, H, sc, d_minhash, queue, scalar_keys2, (int)light_w, n_real, hstride);
        return cudaGetLastError();
    }
#ifdef MHAPB_AB_KERNELS
    if (variant == 1 && B <= 32) {
        const size_t smem = (size_t)4 * (4 * B * 32 + 64 * kBsStage + 64) * 4;
        e = cudaFuncSetAttribute(k_minhash_bs<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_minhash_bs<B>, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int grid = sm_count() * per_sm;
        int need = (n_strands + 3) / 4;
        if (grid > need) grid = need;
        static int scalar_keys = -1;
        if (scalar_keys < 0) { const char *ev = getenv("MHAPB_BS_SCALAR_KEYS"); scalar_keys = ev ? atoi(ev) : kBsScalarKeys; if (scalar_keys < 0) scalar_keys = 0; }
        k_minhash_bs<B><<<grid, 128, smem, st>>>(d_desc, n_strands, k, H, sc, d_minhash, queue, scalar_keys);
        return cudaGetLastError();
    }
#endif
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_minhash<B>, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count() * per_sm;
    int need = (n_strands + 7) / 8;
    if (grid > need) grid = need;
    k_minhash<B><<<grid, 256, 0, st>>>(d_desc, n_strands, k, H, sc, d_minhash, queue, light_w);
    return cudaGetLastError();
}

// the later word blocks' chain starts (see k_minhash_bs2 / k_advance_keys)
cudaError_t launch_advance_keys(cudaStream_t st, const StrandDesc *d_desc, int n_strands, int k, const SketchScratch &sc, const uint64_t *d_t512,
                                uint64_t total_k, int passes, uint32_t light_w, int *launches)
{
    if (n_strands <= 0 || passes < 2) return cudaSuccess;
    k_advance_keys<<<std::min(n_strands, sm_count() * 16), 256, 0, st>>>(d_desc, n_strands, k, sc, d_t512, total_k, passes, light_w);
    (*launches)++;
    return cudaGetLastError();
}

// virtual strands: n_virtual descriptors of <= 512 words each, rows of hstride words (always the B = 16 instantiation)
cudaError_t launch_minhash_virtual(cudaStream_t st, const StrandDesc *d_vdesc, int n_virtual, int n_real, int k, int hstride,
                                   const SketchScratch &sc, int32_t *d_minhash, uint32_t light_w, uint32_t *queue, int *launches)
{
    if (n_virtual <= 0) return cudaSuccess;
    (*launches)++;
    return launch_minhash_b<16>(st, d_vdesc, n_virtual, k, 512, sc, d_minhash, light_w, queue, 1, n_real, hstride);
}

cudaError_t launch_minhash(cudaStream_t st, const StrandDesc *d_desc, int n_strands, int k, int H,
                           const SketchScratch &sc, int32_t *d_minhash, uint32_t light_w, uint32_t *queue, int *launches)
{
    if (n_strands <= 0) return cudaSuccess;
    (*launches)++;
    const int b = (H + 31) / 32;
    if (b <= 1) return launch_minhash_b<1>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 2) return launch_minhash_b<2>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 4) return launch_minhash_b<4>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 8) return launch_minhash_b<8>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 12) return launch_minhash_b<12>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 16) return launch_minhash_b<16>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 24) return launch_minhash_b<24>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 32) return launch_minhash_b<32>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    if (b <= 48) return launch_minhash_b<48>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
    return launch_minhash_b<64>(st, d_desc, n_strands, k, H, sc, d_minhash, light_w, queue);
}

cudaError_t launch_ordered(cudaStream_t st, const uint8_t *d_bases, const StrandDesc *d_desc, int s_base, int n_strands,
                           int first_long, int max_len_short, int max_len_long, int ok, int S, int ord_stride,
                           const SketchScratch &sc, int32_t *d_ord, int32_t *d_ord_n, int max_ctas_per_sm, uint32_t *queues, int *launches)
{
    cudaError_t e;
    uint32_t sel_cap = 1; while ((int)sel_cap < S) sel_cap <<= 1;
    if (first_long > 0) {
        uint32_t len_cap = (uint32_t)align16((size_t)max_len_short + 16);
        // a short read may have fewer ordered k-mers than S, but never more than len_cap
        size_t smem = (size_t)sel_cap * 8 + (size_t)kOrdBins * 4 + (size_t)len_cap * 4 + len_cap;
        auto kern = ok == 12 ? k_ordered<false, 12> : k_ordered<false, 0>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 512, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (max_ctas_per_sm > 0 && per_sm > max_ctas_per_sm) per_sm = max_ctas_per_sm;
        int grid = sm_count() * per_sm;
        if (grid > first_long) grid = first_long;
        kern<<<grid, 512, smem, st>>>(d_bases, d_desc, s_base, s_base + first_long, ok, S, ord_stride, len_cap, sel_cap, sc, d_ord, d_ord_n, queues + 0);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (first_long < n_strands) {
        uint32_t len_cap = (uint32_t)align16((size_t)max_len_long + 16);
        size_t smem = (size_t)sel_cap * 8;
        e = cudaFuncSetAttribute(k_ordered<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int grid = ordered_grid();
        if (grid > n_strands - first_long) grid = n_strands - first_long;
        k_ordered<true, 0><<<grid, 512, smem, st>>>(d_bases, d_desc, s_base + first_long, s_base + n_strands, ok, S, ord_stride, len_cap, sel_cap, sc, d_ord, d_ord_n, queues + 1);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace mhapb
