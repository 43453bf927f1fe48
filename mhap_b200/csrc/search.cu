// search.cu -- K2: MinHashSearch on sm_100a: inverted-index build, probe + hit counting, and the
// second-stage ordered-sketch filter.
//
// Replaces (paths relative to /root/reference/src/main/java/edu/umd/marbl/mhap/):
//   K2a k_index_*   : impl/MinHashSearch.java:101-147 (addSequence: H maps value -> list of ids)
//   K2b k_probe     : impl/MinHashSearch.java:150-225 (findMatches: bucket walk, hit counts, filters)
//   K2c k_filter    : sketch/BottomOverlapSketch.java:592-630 (getOverlapInfo) with MatchData :64-298,
//                     recordMatchingKmers :397-516, computeKBottomSketchJaccard :304-364 and
//                     utils/Utils.java:445-494 (quickSelect)
// HBM / L2 random-access bound integer work; no tensor cores.
#include "engine.h"

#include <algorithm>
#include <cstdlib>

namespace mhapb {

static constexpr uint64_t kEmptySlot = ~0ull;
static constexpr unsigned kFull = 0xffffffffu;
static constexpr uint32_t kLastFlag = 0x80000000u;

__device__ __forceinline__ uint32_t slot_hash(uint32_t value, int log2capw)
{
    return (value * 0x9E3779B1u) >> (32 - log2capw);   // Fibonacci hashing into one word's sub-table
}

// ---------------------------------------------------------------------------------------------
// K2a: index build (count -> scan -> fill -> pack)
// ---------------------------------------------------------------------------------------------
// Sub-table w (capw slots, capw = pow2 >= 2*n_store) holds the distinct values of min-hash word w,
// i.e. it is the reference's hashes.get(w) map.  During the count pass a slot is (value | cnt<<32).
// Work order: a CTA owns one (word group, sketch range) tile, word groups outermost.  A group is 8 consecutive words = the 32-byte
// sector of a min-hash row, so every sector read from HBM is fully used, and the sub-tables the CTAs resident at any moment
// insert into are those of one or two word groups (8 x capw x 8 B = 32 MB for 2*10^5 sketches): they stay in the 126 MB L2 and
// the CAS / add traffic never goes to DRAM.  (The first version walked the rows linearly, word innermost: consecutive threads
// hit 512 different sub-tables -- 2 GB of working set, one DRAM sector per atomic; ncu: 31-38 % of DRAM throughput for 8-byte slots.)
constexpr int kIdxMaxProbe = 128;     // longest probe sequence the build accepts (see IndexView::overflow)
constexpr int kIdxGroup = 8;          // words per group
constexpr int kIdxTileRows = 256;     // sketches per tile (small: the CTAs resident at one moment span one or two word groups)

__device__ __forceinline__ bool index_tile(int64_t tile, int64_t tiles_per_group, int64_t n_store, int H, int *w0, int *nw, int64_t *r0, int64_t *r1)
{
    const int64_t g = tile / tiles_per_group, t = tile % tiles_per_group;
    *w0 = (int)g * kIdxGroup;
    *nw = min(kIdxGroup, H - *w0);
    *r0 = t * kIdxTileRows;
    *r1 = min(n_store, *r0 + kIdxTileRows);
    return *w0 < H;
}

__global__ void __launch_bounds__(256) k_index_count(const int32_t *__restrict__ minhash, int64_t n_store, int H, uint64_t *slots, int log2capw, uint32_t *overflow)
{
    const uint32_t capmask = (1u << log2capw) - 1;
    const int64_t tiles_per_group = (n_store + kIdxTileRows - 1) / kIdxTileRows;
    const int64_t n_tiles = tiles_per_group * ((H + kIdxGroup - 1) / kIdxGroup);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int w0, nw; int64_t r0, r1;
        index_tile(tile, tiles_per_group, n_store, H, &w0, &nw, &r0, &r1);
        // thread -> (row, word in group): 8 consecutive threads read one 32-byte sector
        for (int64_t e = threadIdx.x; e < (r1 - r0) * kIdxGroup; e += blockDim.x) {
            const int64_t row = r0 + (e >> 3);
            const int wi = (int)(e & 7);
            if (wi >= nw) continue;
            const int w = w0 + wi;
            const uint32_t v = (uint32_t)minhash[row * H + w];
            uint64_t *sub = slots + ((size_t)w << log2capw);
            uint32_t p = slot_hash(v, log2capw);
            int steps = 0;
            for (; steps < kIdxMaxProbe; steps++) {
                unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&sub[p]);
                if (cur == kEmptySlot) {
                    cur = atomicCAS(reinterpret_cast<unsigned long long *>(&sub[p]), (unsigned long long)kEmptySlot, (unsigned long long)v);
                    if (cur == kEmptySlot) cur = v;
                }
                if ((uint32_t)cur == v) {
                    atomicAdd(reinterpret_cast<unsigned int *>(&sub[p]) + 1, 1u);   // cnt lives in the high word
                    break;
                }
                p = (p + 1) & capmask;
            }
            if (steps == kIdxMaxProbe) *overflow = 1u;   // the sub-table is too small for this store: the host rebuilds
        }
    }
}

// three-kernel exclusive scan of the slot counts (empty slots count 0)
constexpr int kScanThreads = 512, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t slot_cnt(uint64_t s) { return s == kEmptySlot ? 0u : (uint32_t)(s >> 32); }

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t s_w[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < kScanThreads / 32 ? s_w[lane] : 0, xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFull, xi, o); if (lane >= o) xi += t; }
        if (lane < kScanThreads / 32) s_w[lane] = xi - x;
        if (lane == 31 && total) *total = xi;
    }
    __syncthreads();
    uint32_t r = incl - v + s_w[w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const uint64_t *__restrict__ slots, size_t n, uint32_t *block_sums)
{
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) if (base + q < n) sum += slot_cnt(slots[base + q]);
    __shared__ uint32_t s_total;
    block_exclusive_scan(sum, &s_total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(uint32_t *block_sums, size_t nb)
{
    __shared__ uint32_t s_carry, s_total;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nb; b0 += kScanThreads) {
        size_t i = b0 + threadIdx.x;
        uint32_t v = i < nb ? block_sums[i] : 0;
        uint32_t ex = block_exclusive_scan(v, &s_total);
        if (i < nb) block_sums[i] = ex + s_carry;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint64_t *__restrict__ slots, size_t n, const uint32_t *__restrict__ block_sums, uint32_t *start)
{
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t c[kScanItems], sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { c[q] = base + q < n ? slot_cnt(slots[base + q]) : 0; sum += c[q]; }
    uint32_t run = block_exclusive_scan(sum, nullptr) + block_sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < kScanItems; q++) { if (base + q < n) start[base + q] = run; run += c[q]; }
}

__global__ void __launch_bounds__(256) k_index_fill(const int32_t *__restrict__ minhash, int64_t n_store, int H, const uint64_t *__restrict__ slots,
                                                    int log2capw, uint32_t *start, uint32_t *postings)
{
    const uint32_t capmask = (1u << log2capw) - 1;
    const int64_t tiles_per_group = (n_store + kIdxTileRows - 1) / kIdxTileRows;
    const int64_t n_tiles = tiles_per_group * ((H + kIdxGroup - 1) / kIdxGroup);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int w0, nw; int64_t r0, r1;
        index_tile(tile, tiles_per_group, n_store, H, &w0, &nw, &r0, &r1);
        for (int64_t e = threadIdx.x; e < (r1 - r0) * kIdxGroup; e += blockDim.x) {
            const int64_t row = r0 + (e >> 3);
            const int wi = (int)(e & 7);
            if (wi >= nw) continue;
            const int w = w0 + wi;
            const uint32_t v = (uint32_t)minhash[row * H + w];
            const size_t sub = (size_t)w << log2capw;
            uint32_t p = slot_hash(v, log2capw);
            int steps = 0;
            for (; steps < kIdxMaxProbe; steps++) { const uint64_t sl = slots[sub + p]; if (sl != kEmptySlot && (uint32_t)sl == v) break; p = (p + 1) & capmask; }
            if (steps == kIdxMaxProbe) continue;          // only after an overflow (the index is discarded)
            const uint32_t pos = atomicAdd(&start[sub + p], 1u);
            postings[pos] = (uint32_t)row;
        }
    }
}

// after the fill start[] holds each bucket's end: flag the last posting, rewrite slot as (value | begin<<32)
__global__ void k_index_pack(uint64_t *slots, size_t n, const uint32_t *__restrict__ start, uint32_t *postings, uint32_t *present)
{
    // n is a multiple of 32 and the stride a multiple of the warp size: a warp always covers 32 consecutive slots
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t s = slots[i];
        const unsigned occ = __ballot_sync(kFull, s != kEmptySlot);
        if ((threadIdx.x & 31) == 0) present[i >> 5] = occ;
        if (s == kEmptySlot) continue;
        uint32_t cnt = (uint32_t)(s >> 32), end = start[i];
        postings[end - 1] |= kLastFlag;
        slots[i] = (uint64_t)(uint32_t)s | ((uint64_t)(end - cnt) << 32);
    }
}

cudaError_t launch_index_build(cudaStream_t st, const int32_t *d_minhash, int64_t n_store, int H, IndexView iv,
                               uint32_t *d_tmp_start, uint32_t *d_block_sums, int *launches)
{
    const size_t nslots = (size_t)H << iv.log2capw;
    cudaError_t e = cudaMemsetAsync(iv.slots, 0xff, nslots * 8, st);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    const int64_t n_tiles = ((n_store + kIdxTileRows - 1) / kIdxTileRows) * ((H + kIdxGroup - 1) / kIdxGroup);
    const int tgrid = (int)std::min<int64_t>(n_tiles, 1 << 30);   // one tile per CTA, in tile order: short-lived CTAs also let the collectives' kernels in
    e = cudaMemsetAsync(iv.overflow, 0, 4, st);
    if (e != cudaSuccess) return e;
    k_index_count<<<tgrid, 256, 0, st>>>(d_minhash, n_store, H, iv.slots, iv.log2capw, iv.overflow);
    const size_t nb = (nslots + kScanTile - 1) / kScanTile;
    k_scan_tiles<<<(unsigned)nb, kScanThreads, 0, st>>>(iv.slots, nslots, d_block_sums);
    k_scan_sums<<<1, kScanThreads, 0, st>>>(d_block_sums, nb);
    k_scan_apply<<<(unsigned)nb, kScanThreads, 0, st>>>(iv.slots, nslots, d_block_sums, d_tmp_start);
    k_index_fill<<<tgrid, 256, 0, st>>>(d_minhash, n_store, H, iv.slots, iv.log2capw, d_tmp_start, iv.postings);
    k_index_pack<<<grid, 256, 0, st>>>(iv.slots, nslots, d_tmp_start, iv.postings, iv.present);
    *launches += 6;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K2b: probe + hit counting
// ---------------------------------------------------------------------------------------------
// One CTA per query.  The H buckets are walked by the CTA's threads and the per-target hit counts
// (the reference's bestSequenceHit map) live in a shared-memory open-addressed table.
// Two launches share the kernel template: the first pass gives every query a SMALL table (1024 slots, 8 KB: 16 CTAs
// per SM instead of 6, and an 8 KB clear per query instead of 32 KB -- with the index sharded over N GPUs a query meets
// ~1/N of its hits per rank, the clear was most of the per-query cost); a query that touches more distinct targets than
// the small table holds is appended to an overflow list and redone by the second pass with the 4096-slot table, which
// in turn falls back to dense per-range counters (exact, a few extra bucket walks; only repeat-rich queries get there).
constexpr int kProbeThreads = 128;
constexpr int kHitCapSmall = 1024, kHitMaxSmall = 704;
constexpr int kHitCap = 4096;          // slots in the big table (32 KB)
constexpr int kHitMaxDistinct = 3072;  // switch to the dense path above this many distinct targets
constexpr int kDenseRange = 2 * kHitCap * 2;   // u16 counters in the same 32 KB

struct HitSlot { uint32_t t; uint32_t c; };

__device__ __forceinline__ bool find_bucket(const IndexView &iv, int w, uint32_t v, uint32_t *begin)
{
    const uint64_t *sub = iv.slots + ((size_t)w << iv.log2capw);
    const uint32_t *pres = iv.present + ((size_t)w << (iv.log2capw - 5));
    const uint32_t capmask = (1u << iv.log2capw) - 1;
    uint32_t p = slot_hash(v, iv.log2capw);
    for (;;) {
        if (iv.use_present && !((__ldg(&pres[p >> 5]) >> (p & 31)) & 1u)) return false;   // empty slot: answered from the L2-resident bitmap
        uint64_t s = __ldg(&sub[p]);
        if (s == kEmptySlot) return false;
        if ((uint32_t)s == v) { *begin = (uint32_t)(s >> 32); return true; }
        p = (p + 1) & capmask;
    }
}

__device__ __forceinline__ bool pass_filters(const ProbeArgs &a, int64_t qid, int32_t qlen, uint32_t t, uint32_t count)
{
    const int64_t tid = a.t_id[t];
    if (a.to_self && tid == qid) return false;                                   // MinHashSearch.java:200
    if ((int)count < a.num_min_matches) return false;                            // :204
    const int32_t tlen = a.t_len[t];
    const int ms = a.min_store_length;
    if (tlen < ms && qlen < ms) return false;                                    // :211
    if (a.to_self && tid > qid && tlen >= ms && qlen >= ms) return false;        // :215-219
    if (a.to_self && tlen < ms && qlen >= ms) return false;                      // :222-225
    return true;
}

__device__ __forceinline__ void emit_candidate(const ProbeArgs &a, uint32_t q, uint32_t t, uint32_t count)
{
    unsigned long long p = atomicAdd(&a.counters[0], 1ull);
    if (p < a.cand_cap) { Candidate c; c.q = q; c.t = t; c.count = count; a.cand[p] = c; }
}

// SECOND = false: queries a.q_list[0..nq_list) (or 0..nq_list), table of CAP slots, overflowing queries -> a.ovf_q / counters[3]
// SECOND = true : queries a.ovf_q[0..counters[3]), big table + dense fallback
template <int CAP, int MAXD, bool SECOND>
__global__ void __launch_bounds__(kProbeThreads)
k_probe(IndexView iv, ProbeArgs a)
{
    __shared__ HitSlot s_tab[CAP];
    __shared__ int s_distinct, s_overflow;
    __shared__ unsigned long long s_elements;

    if (*iv.overflow) return;                                 // incomplete index (sub-tables sized too small): the host rebuilds and searches again
    const int64_t n_work = SECOND ? (int64_t)a.counters[3] : a.nq_list;
    for (int64_t qi = blockIdx.x; qi < n_work; qi += gridDim.x) {
        const uint32_t q = SECOND ? a.ovf_q[qi] : (a.q_list ? a.q_list[qi] : (uint32_t)qi);
        const int32_t *qmh = a.q_minhash + (size_t)q * iv.H;
        const int64_t qid = a.q_id[q];
        const int32_t qlen = a.q_len[q];

        for (int i = threadIdx.x; i < CAP; i += blockDim.x) { s_tab[i].t = 0xffffffffu; s_tab[i].c = 0; }
        if (threadIdx.x == 0) { s_distinct = 0; s_overflow = 0; s_elements = 0; }
        __syncthreads();

        unsigned long long elements = 0;
        for (int w = threadIdx.x; w < iv.H; w += blockDim.x) {
            uint32_t begin;
            if (!find_bucket(iv, w, (uint32_t)qmh[w], &begin)) continue;
            if (!SECOND && *reinterpret_cast<volatile int *>(&s_overflow)) break;   // the second pass redoes this query from scratch
            for (uint32_t p = begin;; p++) {
                const uint32_t raw = __ldg(&iv.postings[p]);
                const uint32_t t = raw & ~kLastFlag;
                elements++;
                if (!*reinterpret_cast<volatile int *>(&s_overflow)) {
                    uint32_t hs = (t * 0x9E3779B1u) >> (32 - (CAP == 4096 ? 12 : 10));
                    static_assert(CAP == 4096 || CAP == 1024, "table sizes");
                    for (;;) {
                        uint32_t old = atomicCAS(&s_tab[hs].t, 0xffffffffu, t);
                        if (old == 0xffffffffu) {
                            if (atomicAdd(&s_distinct, 1) + 1 > MAXD) s_overflow = 1;
                            old = t;
                        }
                        if (old == t) { atomicAdd(&s_tab[hs].c, 1u); break; }
                        hs = (hs + 1) & (CAP - 1);
                    }
                }
                if (raw & kLastFlag) break;
            }
        }
        atomicAdd(&s_elements, elements);
        __syncthreads();

        if (!s_overflow) {
            for (int i = threadIdx.x; i < CAP; i += blockDim.x) {
                const uint32_t t = s_tab[i].t;
                if (t == 0xffffffffu) continue;
                if (pass_filters(a, qid, qlen, t, s_tab[i].c)) emit_candidate(a, q, t, s_tab[i].c);
            }
            if (threadIdx.x == 0) { atomicAdd(&a.counters[1], s_elements); atomicAdd(&a.counters[2], (unsigned long long)s_distinct); }
        } else if (!SECOND) {
            if (threadIdx.x == 0) { const unsigned long long p = atomicAdd(&a.counters[3], 1ull); a.ovf_q[p] = q; }
        } else {
            // dense recount: targets [r0, r0+kDenseRange) per pass, u16 counters (count <= H <= 2048)
            uint16_t *cnt = reinterpret_cast<uint16_t *>(s_tab);
            unsigned long long distinct = 0;
            __syncthreads();
            for (int64_t r0 = 0; r0 < iv.n_store; r0 += kDenseRange) {
                for (int i = threadIdx.x; i < kDenseRange / 2; i += blockDim.x) reinterpret_cast<uint32_t *>(cnt)[i] = 0;
                __syncthreads();
                for (int w = threadIdx.x; w < iv.H; w += blockDim.x) {
                    uint32_t begin;
                    if (!find_bucket(iv, w, (uint32_t)qmh[w], &begin)) continue;
                    for (uint32_t p = begin;; p++) {
                        const uint32_t raw = __ldg(&iv.postings[p]);
                        const int64_t rel = (int64_t)(raw & ~kLastFlag) - r0;
                        if (rel >= 0 && rel < kDenseRange) {
                            // 16-bit increment through a 32-bit atomic on the containing word
                            atomicAdd(reinterpret_cast<uint32_t *>(cnt) + (rel >> 1), (rel & 1) ? 0x10000u : 1u);
                        }
                        if (raw & kLastFlag) break;
                    }
                }
                __syncthreads();
                for (int i = threadIdx.x; i < kDenseRange; i += blockDim.x) {
                    const uint32_t c = cnt[i];
                    if (!c) continue;
                    distinct++;
                    const uint32_t t = (uint32_t)(r0 + i);
                    if (pass_filters(a, qid, qlen, t, c)) emit_candidate(a, q, t, c);
                }
                __syncthreads();
            }
            atomicAdd(&a.counters[2], distinct);
            if (threadIdx.x == 0) atomicAdd(&a.counters[1], s_elements);
        }
        __syncthreads();
    }
}

cudaError_t launch_probe(cudaStream_t st, IndexView iv, ProbeArgs a, int *launches)
{
    if (a.nq_list <= 0) return cudaSuccess;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (int64_t)sms * 16;
    if (grid > a.nq_list) grid = a.nq_list;
    k_probe<kHitCapSmall, kHitMaxSmall, false><<<(unsigned)grid, kProbeThreads, 0, st>>>(iv, a);
    // second pass over the overflow list (count read on the device: no host round trip; empty list = an idle launch)
    int64_t grid2 = std::min<int64_t>((int64_t)sms * 6, a.nq_list);
    k_probe<kHitCap, kHitMaxDistinct, true><<<(unsigned)grid2, kProbeThreads, 0, st>>>(iv, a);
    *launches += 2;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K2c: second-stage ordered-sketch filter
// ---------------------------------------------------------------------------------------------
// One thread per candidate pair; the merge is inherently serial (window tests interleaved with the
// hash compare, "first and last match" handling of duplicate hashes).  Matches are recorded in an
// HBM scratch laid out [entry][thread] so a warp's records coalesce.
struct Scratch {
    int32_t *p1, *p2, *tmp; uint32_t stride;
    __device__ __forceinline__ int32_t &P1(int i) const { return p1[(size_t)i * stride]; }
    __device__ __forceinline__ int32_t &P2(int i) const { return p2[(size_t)i * stride]; }
    __device__ __forceinline__ int32_t &T(int i) const { return tmp[(size_t)i * stride]; }
};

// utils/Utils.java:445-494 on the scratch copy
__device__ int32_t quick_select(const Scratch &sc, int k, int length)
{
    int from = 0, to = length - 1;
    while (from < to) {
        int r = from, w = to;
        const int32_t mid = sc.T((r + w) / 2);
        while (r < w) {
            if (sc.T(r) >= mid) { int32_t tmp = sc.T(w); sc.T(w) = sc.T(r); sc.T(r) = tmp; w--; }
            else r++;
        }
        if (sc.T(r) > mid) r--;
        if (k <= r) to = r; else from = r + 1;
    }
    return sc.T(k);
}

struct MatchState { int32_t count, median, absmax, len1, len2; double max_shift; };

// MatchData.performUpdate, sketch/BottomOverlapSketch.java:191-215
__device__ void perform_update(MatchState &m, const Scratch &sc)
{
    if (m.count > 0) {
        for (int i = 0; i < m.count; i++) sc.T(i) = sc.P2(i) - sc.P1(i);
        m.median = quick_select(sc, m.count / 2, m.count);
        const int32_t left = max(0, -m.median);
        const int32_t right = min(m.len1, m.len2 - m.median);
        const int32_t overlap = max(10, right - left);
        m.absmax = min(max(m.len1, m.len2), (int32_t)((double)overlap * m.max_shift));
    } else {
        m.median = 0;
        m.absmax = max(m.len1, m.len2) + 1;
    }
}

// recordMatchingKmers, sketch/BottomOverlapSketch.java:397-516
__device__ void record_matching(MatchState &m, const Scratch &sc, const int2 *__restrict__ s1, int n1, const int2 *__restrict__ s2, int n2)
{
    const int32_t median = m.median, absmax = m.absmax;
    const int32_t v1lo = max(0, -median - absmax);
    const int32_t v2lo = max(0, median - absmax);
    const int32_t v1hi = min(m.len1, m.len2 - median + absmax);
    const int32_t v2hi = min(m.len2, m.len1 + median + absmax);
    int i1 = 0, i2 = 0, count = 0;
    if (n1 > 0 && n2 > 0) {
        int2 e1 = __ldg(&s1[0]), e2 = __ldg(&s2[0]);
        for (;;) {
            const int32_t hash1 = e1.x, pos1 = e1.y, hash2 = e2.x, pos2 = e2.y;
            if (hash1 < hash2 || pos1 < v1lo || pos1 >= v1hi) { if (++i1 >= n1) break; e1 = __ldg(&s1[i1]); }
            else if (hash2 < hash1 || pos2 < v2lo || pos2 >= v2hi) { if (++i2 >= n2) break; e2 = __ldg(&s2[i2]); }
            else {
                const int32_t diff = (pos2 - pos1) - median;
                if (diff > absmax) { if (++i1 >= n1) break; e1 = __ldg(&s1[i1]); }
                else if (diff < -absmax) { if (++i2 >= n2) break; e2 = __ldg(&s2[i2]); }
                else {
                    sc.P1(count) = pos1; sc.P2(count) = pos2; count++;
                    int i1last = i1, i1try = i1 + 1;
                    int32_t p1last = pos1;
                    while (i1try < n1) {
                        int2 t = __ldg(&s1[i1try]);
                        if (!(t.x == hash1 && t.y >= v1lo && t.y < v1hi)) break;
                        i1last = i1try; p1last = t.y; i1try++;
                    }
                    int i2last = i2, i2try = i2 + 1;
                    int32_t p2last = pos2;
                    while (i2try < n2) {
                        int2 t = __ldg(&s2[i2try]);
                        if (!(t.x == hash2 && t.y >= v2lo && t.y < v2hi)) break;
                        i2last = i2try; p2last = t.y; i2try++;
                    }
                    if (i1 != i1last || i2 != i2last) {
                        sc.P1(count) = p1last; sc.P2(count) = p2last; count++;
                        i1 = i1last + 1; i2 = i2last + 1;
                    } else { i1++; i2++; }
                    if (i1 >= n1 || i2 >= n2) break;
                    e1 = __ldg(&s1[i1]); e2 = __ldg(&s2[i2]);
                }
            }
        }
    }
    m.count = count;
}

__device__ __forceinline__ int32_t java_round_div(int32_t num, int32_t den)
{
    // (int) Math.round((double) num / (double) den): round half up (floor(x + 0.5) semantics)
    double x = (double)num / (double)den;
    double f = floor(x);
    return (int32_t)((long long)f + ((x - f) >= 0.5 ? 1 : 0));
}

__global__ void __launch_bounds__(128) k_filter(FilterArgs a)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= a.n_threads) return;
    Scratch sc;
    sc.stride = a.n_threads;
    sc.p1 = a.scratch + tid;
    sc.p2 = a.scratch + (size_t)a.scratch_entries * a.n_threads + tid;
    sc.tmp = a.scratch + 2 * (size_t)a.scratch_entries * a.n_threads + tid;

    uint64_t n_work = a.sel ? a.n_sel : a.n_cand;
    if (a.sel && a.n_sel_dev) n_work = *a.n_sel_dev;                      // the warp kernel's overflow cursor
    else if (!a.sel && a.n_cand_dev) n_work = min((uint64_t)*a.n_cand_dev, a.cand_cap);
    for (uint64_t wi = tid; wi < n_work; wi += a.n_threads) {
        const uint64_t ci = a.sel ? a.sel[wi] : wi;
        const Candidate c = a.cand[ci];
        const int2 *A = reinterpret_cast<const int2 *>(a.q_ord) + (size_t)c.q * a.q_stride;
        const int2 *Bs = reinterpret_cast<const int2 *>(a.t_ord) + (size_t)c.t * a.t_stride;
        const int nA = a.q_ord_n[c.q], nB = a.t_ord_n[c.t];
        OverlapOut o; o.a1 = o.a2 = o.b1 = o.b2 = o.valid = o.inter = o.kmin = 0; o.empty = 1;
        MatchState m; m.count = 0; m.len1 = a.q_lenk[c.q]; m.len2 = a.t_lenk[c.t]; m.max_shift = a.max_shift;

        perform_update(m, sc);                       // count == 0: median 0, absmax = max(len)+1
        record_matching(m, sc, A, nA, Bs, nB);       // getOverlapInfo :601
        if (m.count > 0) {
            perform_update(m, sc);
            record_matching(m, sc, A, nA, Bs, nB);   // :607
        }
        if (m.count > 0) {
            // optimizeShifts :156-189
            perform_update(m, sc);
            int reduced = -1;
            for (int it = 0; it < m.count; it++) {
                const int32_t p1 = sc.P1(it), p2 = sc.P2(it);
                if (reduced >= 0 && sc.P1(reduced) == p1) {
                    const int32_t sr = sc.P2(reduced) - sc.P1(reduced);
                    if (abs(sr - m.median) > abs((p2 - p1) - m.median)) { sc.P1(reduced) = p1; sc.P2(reduced) = p2; }
                } else { reduced++; sc.P1(reduced) = p1; sc.P2(reduced) = p2; }
            }
            m.count = reduced + 1;
            perform_update(m, sc);
            // computeEdges :90-137
            int32_t le1 = INT32_MAX, le2 = INT32_MAX, re1 = INT32_MIN, re2 = INT32_MIN, valid = 0;
            for (int it = 0; it < m.count; it++) {
                const int32_t p1 = sc.P1(it), p2 = sc.P2(it);
                if (abs((p2 - p1) - m.median) > m.absmax) continue;
                le1 = min(le1, p1); le2 = min(le2, p2); re1 = max(re1, p1); re2 = max(re2, p2);
                valid++;
            }
            if (valid >= 3) {
                const int32_t n = valid;
                o.a1 = max(0, java_round_div(n * le1 - re1, n - 1));
                o.a2 = min(m.len1, java_round_div(n * re1 - le1, n - 1));
                o.b1 = max(0, java_round_div(n * le2 - re2, n - 1));
                o.b2 = min(m.len2, java_round_div(n * re2 - le2, n - 1));
                o.valid = valid;
                // computeKBottomSketchJaccard :304-364, streamed: s1/s2 first, then the bottom-k merge
                int s1 = 0, s2 = 0;
                for (int i = 0; i < nA; i++) { int32_t p = __ldg(&A[i]).y; s1 += (p >= o.a1 && p <= o.a2); }
                for (int j = 0; j < nB; j++) { int32_t p = __ldg(&Bs[j]).y; s2 += (p >= o.b1 && p <= o.b2); }
                const int k = min(s1, s2);
                int inter = 0;
                if (k > 0) {
                    int i = 0, j = 0, uni = 0;
                    // advance to the first in-window entry of each side
                    int2 ea = __ldg(&A[0]); while (!(ea.y >= o.a1 && ea.y <= o.a2)) ea = __ldg(&A[++i]);
                    int2 eb = __ldg(&Bs[0]); while (!(eb.y >= o.b1 && eb.y <= o.b2)) eb = __ldg(&Bs[++j]);
                    while (uni < k) {
                        bool adv_a = false, adv_b = false;
                        if (ea.x < eb.x) adv_a = true;
                        else if (ea.x > eb.x) adv_b = true;
                        else { inter++; adv_a = adv_b = true; }
                        uni++;
                        if (uni >= k) break;
                        // the reference indexes the filtered arrays; entries past the k-th union element are never read
                        if (adv_a) { do { ++i; if (i >= nA) break; ea = __ldg(&A[i]); } while (!(ea.y >= o.a1 && ea.y <= o.a2)); }
                        if (adv_b) { do { ++j; if (j >= nB) break; eb = __ldg(&Bs[j]); } while (!(eb.y >= o.b1 && eb.y <= o.b2)); }
                    }
                }
                o.inter = inter; o.kmin = k; o.empty = 0;
            }
        }
        a.out[ci] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// K2c, warp-per-candidate (the default)
// ---------------------------------------------------------------------------------------------
// recordMatchingKmers is a merge-join on hash whose state never carries across hash values (elements of either
// sketch with a hash the other side lacks are skipped without a record; the first/last handling of duplicate hashes
// stays inside one hash value), so the hash axis is cut into 32 ranges at values taken from sketch A and every lane
// runs the reference's sequential loop on its range, straight from L2/L1; concatenating the lanes' records in lane
// order reproduces the reference's record order.  Median = exact k-th smallest (rank by counting for <= 32 records,
// 8-bit radix select above); optimizeShifts = ordered compaction of pos1 runs; the bottom-k merge is partitioned the
// same way, the lane in which the union count crosses k finishing it sequentially.
//
// The first version of this kernel executed 28 700 warp instructions per candidate with 12.6 of 32 threads active
// (profiles/r2a_k2_probe_filter_ncu_summary.txt): every `||` of the reference's if / else-if chain became a branch and the
// three outcomes of a merge step ran one after the other.  Here a merge step is straight-line code -- the two window tests
// are one unsigned compare each, "advance A" / "advance B" are predicates, the element that moved is reloaded under its
// predicate -- and only an actual hash match (a few per lane) takes a branch; the two counting loops and the walk of the
// bottom-k intersect are one fused pass.
constexpr int kFwRecCap = 512;      // match records kept in shared memory per warp
constexpr int kFwLaneCap = kFwRecCap / 2 / 32;   // private first-try slot per lane (8 records)
// (a pair with more match records than kFwRecCap goes to the thread-per-candidate kernel)

// MatchData.valid1Lower/valid1Upper/valid2Lower/valid2Upper as (lower, width): pos is inside iff (unsigned)(pos - lo) < width
struct FwWindow { int32_t lo1, w1, lo2, w2, median, absmax; };

__device__ __forceinline__ FwWindow fw_window(int32_t median, int32_t absmax, int32_t len1, int32_t len2)
{
    FwWindow w;
    w.median = median; w.absmax = absmax;
    w.lo1 = max(0, -median - absmax); w.lo2 = max(0, median - absmax);
    w.w1 = max(0, min(len1, len2 - median + absmax) - w.lo1);
    w.w2 = max(0, min(len2, len1 + median + absmax) - w.lo2);
    return w;
}

// Element i of a sketch.  SM = false: straight from global memory (L1/L2).  SM = true: from the warp's staged copy in shared
// memory, laid out at i + (i >> 5): the lanes walk ranges that start about n/32 elements apart (48 for S = 1536, i.e. 96 words
// = a multiple of the 32 banks), so without the skew every lane's loads would hit the same bank.
template <bool SM>
__device__ __forceinline__ int2 fw_ld(const int2 *__restrict__ p, int i) { if (SM) return p[i + (i >> 5)]; else return __ldg(p + i); }

template <bool SM>
__device__ __forceinline__ int fw_lower_bound(const int2 *__restrict__ s, int n, int32_t h)
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fw_ld<SM>(s, mid).x < h) lo = mid + 1; else hi = mid; }
    return lo;
}

// the reference loop (sketch/BottomOverlapSketch.java:428-515) on A[i1..e1) x B[i2..e2): returns the number of match
// records and stores the first `cap` of them in out.
// A step is straight-line: both window tests are one unsigned compare each, "advance A" / "advance B" are predicates and
// the element that moved is reloaded under its predicate (A / Bs are plain register pointers, so the reload is one
// IMAD.WIDE + one LDG; the first version re-derived the row address from the constant bank with five instructions per
// load and branched around it -- 40 instructions per step).  [Keeping the next element of each side in flight one step
// ahead was measured slower: 19.2 vs 14.9 ms, 72 registers and twice the loads.]
template <bool SM>
__device__ __forceinline__ int fw_merge_range(const int2 *__restrict__ A, int i1, const int e1, const int2 *__restrict__ Bs, int i2, const int e2,
                                              const FwWindow &w, int2 *out, const int cap)
{
    int count = 0;
    if (i1 >= e1 || i2 >= e2) return 0;
    int2 a = fw_ld<SM>(A, i1), b = fw_ld<SM>(Bs, i2);
    for (;;) {
        const bool aout = (uint32_t)(a.y - w.lo1) >= (uint32_t)w.w1;
        const bool bout = (uint32_t)(b.y - w.lo2) >= (uint32_t)w.w2;
        bool adv1 = (a.x < b.x) | aout;                        // :438
        bool adv2 = !adv1 & ((b.x < a.x) | bout);              // :440
        if (!(adv1 | adv2)) {                                  // equal hashes, both positions inside their windows
            const int32_t diff = (b.y - a.y) - w.median;
            if (diff > w.absmax) adv1 = true;
            else if (diff < -w.absmax) adv2 = true;
            else {
                if (count < cap) out[count] = make_int2(a.y, b.y);
                count++;
                int i1last = i1, i2last = i2;
                int32_t p1 = a.y, p2 = b.y;
                for (int t = i1 + 1; t < e1; t++) { const int2 x = fw_ld<SM>(A, t); if (!(x.x == a.x && (uint32_t)(x.y - w.lo1) < (uint32_t)w.w1)) break; i1last = t; p1 = x.y; }
                for (int t = i2 + 1; t < e2; t++) { const int2 x = fw_ld<SM>(Bs, t); if (!(x.x == b.x && (uint32_t)(x.y - w.lo2) < (uint32_t)w.w2)) break; i2last = t; p2 = x.y; }
                if (i1 != i1last || i2 != i2last) {
                    if (count < cap) out[count] = make_int2(p1, p2);
                    count++;
                    i1 = i1last + 1; i2 = i2last + 1;
                } else { i1++; i2++; }
                if (i1 >= e1 || i2 >= e2) break;
                a = fw_ld<SM>(A, i1); b = fw_ld<SM>(Bs, i2);
                continue;
            }
        }
        i1 += adv1; i2 += adv2;
        if (i1 >= e1 || i2 >= e2) break;
        if (adv1) a = fw_ld<SM>(A, i1);
        if (adv2) b = fw_ld<SM>(Bs, i2);
    }
    return count;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int *total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    *total = __shfl_sync(kFull, incl, 31);
    return incl - v;
}

// exact k-th smallest (0-based) of the shifts rec[i].y - rec[i].x, i < n
__device__ int32_t fw_select_shift(const int2 *rec, int n, int k, uint32_t *hist, int lane)
{
    if (n <= 32) {   // rank by counting: one record per lane, n shuffles
        int32_t v = 0;
        if (lane < n) { const int2 r = rec[lane]; v = r.y - r.x; }
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const int32_t vj = __shfl_sync(kFull, v, j);
            rank += (vj < v) | ((vj == v) & (j < lane));
        }
        const unsigned who = __ballot_sync(kFull, lane < n && rank == k);
        return __shfl_sync(kFull, v, __ffs(who) - 1);
    }
    // MSB-first 8-bit radix select
    uint32_t prefix = 0, mask = 0, remaining = (uint32_t)k + 1;
    for (int pass = 3; pass >= 0; pass--) {
        const int shift = pass * 8;
        for (int i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const uint32_t key = (uint32_t)(rec[i].y - rec[i].x) ^ 0x80000000u;
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncwarp();
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { c[q] = hist[lane * 8 + q]; sum += c[q]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
        const uint32_t excl = incl - sum;
        uint32_t digit = 0, before = 0;
        const bool mine = excl < remaining && remaining <= incl;
        if (mine) {
            uint32_t run = excl;
#pragma unroll
            for (int q = 0; q < 8; q++) { if (run < remaining && remaining <= run + c[q]) { digit = lane * 8 + q; before = run; } run += c[q]; }
        }
        const unsigned who = __ballot_sync(kFull, mine);
        const int src = __ffs(who) - 1;
        digit = __shfl_sync(kFull, digit, src); before = __shfl_sync(kFull, before, src);
        prefix |= digit << shift; mask |= 0xffu << shift; remaining -= before;
        __syncwarp();
    }
    return (int32_t)(prefix ^ 0x80000000u);
}

// MatchData.performUpdate (sketch/BottomOverlapSketch.java:191-215) for count > 0
__device__ __forceinline__ void fw_update(const int2 *rec, int count, int32_t len1, int32_t len2, double max_shift, uint32_t *hist, int lane,
                                          int32_t *median, int32_t *absmax)
{
    const int32_t med = fw_select_shift(rec, count, count / 2, hist, lane);
    const int32_t left = max(0, -med), right = min(len1, len2 - med), overlap = max(10, right - left);
    *median = med;
    *absmax = min(max(len1, len2), (int32_t)((double)overlap * max_shift));
}

// SM = true: both sketches of a pair are first copied into the warp's shared memory with coalesced 8-byte cp.async
// (capA / capB skewed slots each) and the three passes over them (two merges, the bottom-k walk) run from there.
// Straight from global memory every lane walks its own contiguous range, so one warp-wide load touches 32 different
// lines, and with the index sharded over several GPUs each pair brings a query sketch nobody else touches:
// 15 ms at N=1, 34 ms at N=8 for the same 3*10^5 pairs per rank (profiles/r2i_bench_config1_n8.json).
template <bool SM>
__global__ void __launch_bounds__(SM ? 224 : 128) k_filter_warp(FilterArgs a, int capA, int capB)
{
    extern __shared__ __align__(16) uint8_t fw_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)(kFwRecCap + capA + capB) * 8 + 256 * 4;
    uint8_t *base = fw_smem + (size_t)wib * per_warp;
    int2 *rec = reinterpret_cast<int2 *>(base);
    uint32_t *hist = reinterpret_cast<uint32_t *>(rec + kFwRecCap);
    int2 *sA = reinterpret_cast<int2 *>(hist + 256), *sB = sA + capA;
    const uint64_t warp_id = (uint64_t)blockIdx.x * wpb + wib, n_warps = (uint64_t)gridDim.x * wpb;
    // the candidate count is read on the device (it is K2b's cursor): no host round trip between probe and filter
    uint64_t n_cand = a.n_cand;
    if (a.n_cand_dev) { n_cand = *a.n_cand_dev; if (n_cand > a.cand_cap) n_cand = a.cand_cap; }

    for (uint64_t ci = warp_id; ci < n_cand; ci += n_warps) {
        const Candidate c = a.cand[ci];
        if (a.prefetch && ci + n_warps < n_cand) {
            // pull the NEXT pair's two sketches towards L2 while this pair is merged: the merges are chains of dependent
            // loads, and with the queries in a gathered block of several GB (multi-GPU) most of them missed L2
            const Candidate cn = a.cand[ci + n_warps];
            const char *pa = reinterpret_cast<const char *>(reinterpret_cast<const int2 *>(a.q_ord) + (size_t)cn.q * a.q_stride);
            const char *pb = reinterpret_cast<const char *>(reinterpret_cast<const int2 *>(a.t_ord) + (size_t)cn.t * a.t_stride);
            for (int o = lane * 128; o < a.q_stride * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(pa + o));
            for (int o = lane * 128; o < a.t_stride * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(pb + o));
        }
        const int nA = a.q_ord_n[c.q], nB = a.t_ord_n[c.t];
        const int32_t len1 = a.q_lenk[c.q], len2 = a.t_lenk[c.t];
        const int2 *A = reinterpret_cast<const int2 *>(a.q_ord) + (size_t)c.q * a.q_stride;
        const int2 *Bs = reinterpret_cast<const int2 *>(a.t_ord) + (size_t)c.t * a.t_stride;
        asm volatile("" : "+l"(A), "+l"(Bs));   // keep the two row addresses in registers: element i is then one IMAD.WIDE away
        __syncwarp();
        if (SM) {
            const uint32_t da = (uint32_t)__cvta_generic_to_shared(sA), db = (uint32_t)__cvta_generic_to_shared(sB);
            for (int i = lane; i < nA; i += 32) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(da + 8u * (uint32_t)(i + (i >> 5))), "l"(A + i) : "memory");
            for (int i = lane; i < nB; i += 32) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(db + 8u * (uint32_t)(i + (i >> 5))), "l"(Bs + i) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            A = sA; Bs = sB;
        }
        OverlapOut o; o.a1 = o.a2 = o.b1 = o.b2 = o.valid = o.inter = o.kmin = 0; o.empty = 1;
        bool overflow = false;

        // hash ranges: lane l takes hashes in [h_l, h_{l+1}), h_l = A[l*nA/32].x (h_0 = -inf)
        int a0 = 0, b0 = 0;
        if (lane > 0 && nA > 0) {
            a0 = (int)(((long long)lane * nA) >> 5);
            const int32_t h = fw_ld<SM>(A, a0).x;
            while (a0 > 0 && fw_ld<SM>(A, a0 - 1).x == h) a0--;        // = lower_bound(A, h): equal hashes are adjacent
            b0 = fw_lower_bound<SM>(Bs, nB, h);
        }
        int a1 = __shfl_down_sync(kFull, a0, 1), b1 = __shfl_down_sync(kFull, b0, 1);
        if (lane == 31) { a1 = nA; b1 = nB; }
        if (a1 < a0) a1 = a0;   // equal splitters give empty ranges
        if (b1 < b0) b1 = b0;

        int count = 0;
        int32_t median = 0, absmax = max(len1, len2) + 1;     // empty MatchData (BottomOverlapSketch.java:207-211)
        for (int pass = 0; pass < 2 && !overflow; pass++) {
            const FwWindow w = fw_window(median, absmax, len1, len2);
            // one merge in the common case: every lane records into a private slot of kFwLaneCap entries in the upper half
            // of the buffer, then the slots are packed in lane order into the lower half; a lane with more matches (or a
            // pair with more than half the buffer) repeats the merge writing at its final offset
            int2 *slot = rec + kFwRecCap / 2 + lane * kFwLaneCap;
            const int mine = fw_merge_range<SM>(A, a0, a1, Bs, b0, b1, w, slot, kFwLaneCap);
            int total;
            const int off = warp_excl_scan(mine, lane, &total);
            count = total;
            // Fewer than three records cannot end as an overlap: computeEdges needs >= 3 valid ones (:125) and neither the second
            // pass nor optimizeShifts adds any -- a hash value yields one record, or two when one side holds a run of equal hashes
            // inside its window, and the second pass only narrows the windows (its matching hashes and runs are subsets of the
            // first pass's, whose shift test rejects nothing).  Most candidates of unrelated reads stop here, after one merge.
            if (total < 3) { count = 0; break; }
            if (total > kFwRecCap) { overflow = true; break; }
            const bool fits = __all_sync(kFull, mine <= kFwLaneCap) && total <= kFwRecCap / 2;
            __syncwarp();
            if (fits) { for (int j = 0; j < mine; j++) rec[off + j] = slot[j]; }
            else fw_merge_range<SM>(A, a0, a1, Bs, b0, b1, w, rec + off, mine);
            __syncwarp();
            fw_update(rec, count, len1, len2, a.max_shift, hist, lane, &median, &absmax);
        }
        if (overflow) {
            if (lane == 0) { unsigned long long p = atomicAdd(a.ovf_count, 1ull); a.ovf_list[p] = (uint32_t)ci; }
            continue;
        }
        if (count > 0) {
            // optimizeShifts (:156-189): within a run of equal pos1 keep the entry closest to the median (first on ties)
            int out_n = 0;
            for (int base_i = 0; base_i < count; base_i += 32) {
                const int i = base_i + lane;
                bool start = false;
                int2 best = make_int2(0, 0);
                if (i < count) {
                    const int2 r = rec[i];
                    start = i == 0 || rec[i - 1].x != r.x;
                    if (start) {
                        best = r;
                        int32_t bd = abs((r.y - r.x) - median);
                        for (int j = i + 1; j < count; j++) {
                            const int2 x = rec[j];
                            if (x.x != r.x) break;
                            const int32_t d = abs((x.y - x.x) - median);
                            if (bd > d) { bd = d; best = x; }
                        }
                    }
                }
                const unsigned m = __ballot_sync(kFull, start);
                __syncwarp();
                if (start) rec[out_n + __popc(m & ((1u << lane) - 1))] = best;
                out_n += __popc(m);
                __syncwarp();
            }
            count = out_n;
            fw_update(rec, count, len1, len2, a.max_shift, hist, lane, &median, &absmax);
            // computeEdges (:90-137)
            int32_t le1 = INT32_MAX, le2 = INT32_MAX, re1 = INT32_MIN, re2 = INT32_MIN, valid = 0;
            for (int i = lane; i < count; i += 32) {
                const int2 r = rec[i];
                if (abs((r.y - r.x) - median) > absmax) continue;
                le1 = min(le1, r.x); le2 = min(le2, r.y); re1 = max(re1, r.x); re2 = max(re2, r.y); valid++;
            }
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) {
                le1 = min(le1, __shfl_xor_sync(kFull, le1, ofs)); le2 = min(le2, __shfl_xor_sync(kFull, le2, ofs));
                re1 = max(re1, __shfl_xor_sync(kFull, re1, ofs)); re2 = max(re2, __shfl_xor_sync(kFull, re2, ofs));
                valid += __shfl_xor_sync(kFull, valid, ofs);
            }
            if (valid >= 3) {
                const int32_t n = valid;
                o.a1 = max(0, java_round_div(n * le1 - re1, n - 1));
                o.a2 = min(len1, java_round_div(n * re1 - le1, n - 1));
                o.b1 = max(0, java_round_div(n * le2 - re2, n - 1));
                o.b2 = min(len2, java_round_div(n * re2 - le2, n - 1));
                o.valid = valid;
                // computeKBottomSketchJaccard (:304-364) over the entries whose position is inside [a1,a2] / [b1,b2].
                // One fused pass over this lane's two ranges: entries outside their window are skipped, the others are
                // counted (sa, sbn) as the two-pointer walk consumes them; equal hashes are consumed together (inter_l).
                const int32_t wa_lo = o.a1, wb_lo = o.b1;
                const uint32_t wa_w = (uint32_t)max(0, o.a2 - o.a1 + 1), wb_w = (uint32_t)max(0, o.b2 - o.b1 + 1);
                int sa = 0, sbn = 0, inter_l = 0;
                {
                    int i = a0, j = b0;
                    int2 ea = make_int2(0, 0), eb = make_int2(0, 0);
                    if (i < a1) ea = fw_ld<SM>(A, i);
                    if (j < b1) eb = fw_ld<SM>(Bs, j);
                    while (i < a1 || j < b1) {
                        const bool ha = i < a1, hb = j < b1;
                        const bool ain = ha & ((uint32_t)(ea.y - wa_lo) < wa_w), bin = hb & ((uint32_t)(eb.y - wb_lo) < wb_w);
                        const bool skipa = ha & !ain, skipb = hb & !bin, noskip = !(skipa | skipb);
                        const bool takea = noskip & ain & (!bin | (ea.x <= eb.x));
                        const bool takeb = noskip & bin & (!ain | (eb.x <= ea.x));
                        sa += takea; sbn += takeb; inter_l += takea & takeb;
                        const bool adva = skipa | takea, advb = skipb | takeb;
                        i += adva; j += advb;
                        if (adva && i < a1) ea = fw_ld<SM>(A, i);
                        if (advb && j < b1) eb = fw_ld<SM>(Bs, j);
                    }
                }
                const int union_l = sa + sbn - inter_l;
                int s1, s2, utot;
                warp_excl_scan(sa, lane, &s1);
                warp_excl_scan(sbn, lane, &s2);
                const int ubefore = warp_excl_scan(union_l, lane, &utot);
                const int k = min(s1, s2);
                int inter = 0;
                if (k > 0) {
                    if (ubefore + union_l <= k) inter = inter_l;                 // whole range inside the first k union steps
                    else if (ubefore < k) {                                       // the range in which the walk stops
                        int i = a0, j = b0, uni = ubefore;
                        while (uni < k) {
                            while (i < a1 && !((uint32_t)(fw_ld<SM>(A, i).y - wa_lo) < wa_w)) i++;
                            while (j < b1 && !((uint32_t)(fw_ld<SM>(Bs, j).y - wb_lo) < wb_w)) j++;
                            if (i >= a1) { j++; }                                 // only B entries left in range: each is one union step
                            else if (j >= b1) { i++; }
                            else {
                                const int32_t ha = fw_ld<SM>(A, i).x, hb = fw_ld<SM>(Bs, j).x;
                                if (ha < hb) i++; else if (ha > hb) j++; else { inter++; i++; j++; }
                            }
                            uni++;
                        }
                    }
#pragma unroll
                    for (int ofs = 16; ofs > 0; ofs >>= 1) inter += __shfl_xor_sync(kFull, inter, ofs);
                }
                o.inter = inter; o.kmin = k; o.empty = 0;
            }
        }
        if (lane == 0) a.out[ci] = o;
    }
}

cudaError_t launch_filter_warp(cudaStream_t st, FilterArgs a, int *launches)
{
    if (a.n_cand == 0 && !a.n_cand_dev) return cudaSuccess;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // tunables (A/B runs): MHAPB_K2C_STAGE=0 keeps the sketches in global memory; MHAPB_K2C_CTAS bounds the pairs in flight
    static int max_ctas = -1, pf = -1, stage = -1;
    if (max_ctas < 0) { const char *e = getenv("MHAPB_K2C_CTAS"); max_ctas = e ? atoi(e) : 0; }
    if (pf < 0) { const char *e = getenv("MHAPB_K2C_PREFETCH"); pf = e ? atoi(e) : 0; }
    // staging is OFF by default: measured slower on 1xB200 (profiles/r2k_k2c_stage_sweep.txt: 18.8 vs 13.1 ms at configs[1],
    // 130 vs 94 ms on 2.1 M pairs) -- 30 KB of shared memory per warp leaves 7 warps per SM against 24 without it, and the
    // merges are chains of dependent loads that need the warps more than they need the bandwidth
    if (stage < 0) { const char *e = getenv("MHAPB_K2C_STAGE"); stage = e ? atoi(e) : 0; }
    const int capA = (a.q_stride + a.q_stride / 32 + 2) & ~1, capB = (a.t_stride + a.t_stride / 32 + 2) & ~1;
    const size_t staged_per_warp = (size_t)(kFwRecCap + capA + capB) * 8 + 256 * 4;
    if (stage && staged_per_warp <= 48 * 1024) {
        // as many warps as one SM's shared memory holds, in one CTA
        int wpb = (int)((size_t)(227 - 2) * 1024 / staged_per_warp);
        if (wpb > 7) wpb = 7;
        const size_t smem = staged_per_warp * wpb;
        static size_t attr_smem = 0;
        if (smem > attr_smem) {
            cudaError_t e = cudaFuncSetAttribute(k_filter_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attr_smem = smem;
        }
        int per_sm = 1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_filter_warp<true>, wpb * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        uint64_t grid = (uint64_t)sms * per_sm;
        if (!a.n_cand_dev) { const uint64_t need = (a.n_cand + wpb - 1) / wpb; if (grid > need) grid = need; }
        a.prefetch = 0;
        k_filter_warp<true><<<(unsigned)grid, wpb * 32, smem, st>>>(a, capA, capB);
        (*launches)++;
        return cudaGetLastError();
    }
    const size_t per_warp = (size_t)kFwRecCap * 8 + 256 * 4;
    const int wpb = 4;
    const size_t smem = per_warp * wpb;
    static int per_sm = 0;
    if (!per_sm) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_filter_warp<false>, wpb * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    // measured (profiles/r2j_k2c_sweep.txt, 2.1 M pairs): 9 CTAs/SM 111.9 ms, 6 -> 103.6, 4 -> 130.8, 2 -> 221; prefetching the
    // next pair towards L2 118.0 vs 111.9 without -- the pairs in flight already fill L2
    int use = per_sm;
    const int cap = max_ctas > 0 ? max_ctas : 6;
    if (use > cap) use = cap;
    a.prefetch = pf;
    uint64_t grid = (uint64_t)sms * use;
    if (!a.n_cand_dev) { const uint64_t need = (a.n_cand + wpb - 1) / wpb; if (grid > need) grid = need; }
    k_filter_warp<false><<<(unsigned)grid, wpb * 32, smem, st>>>(a, 0, 0);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_filter(cudaStream_t st, FilterArgs a, int *launches)
{
    if (a.n_threads == 0) return cudaSuccess;
    unsigned grid = (a.n_threads + 127) / 128;
    k_filter<<<grid, 128, 0, st>>>(a);
    (*launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// result compaction: only pairs that can still pass the threshold travel to the host
// ---------------------------------------------------------------------------------------------
__global__ void k_compact_hits(const Candidate *__restrict__ cand, const OverlapOut *__restrict__ ovl, uint64_t n, const unsigned long long *n_dev,
                               double jmin, int keep_all, Candidate *cand_out, OverlapOut *ovl_out, unsigned long long *count)
{
    const int lane = threadIdx.x & 31;
    if (n_dev) n = min((uint64_t)*n_dev, n);   // n = the capacity, *n_dev = K2b's candidate cursor
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        bool keep = false;
        OverlapOut o;
        if (i < n) {
            o = ovl[i];
            keep = keep_all || (!o.empty && o.kmin > 0 && (double)o.inter >= jmin * (double)o.kmin);
        }
        const unsigned m = __ballot_sync(kFull, keep);
        if (m) {
            unsigned long long p = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) p = atomicAdd(count, (unsigned long long)__popc(m));
            p = __shfl_sync(kFull, p, leader) + __popc(m & ((1u << lane) - 1));
            if (keep) { cand_out[p] = cand[i]; ovl_out[p] = o; }
        }
    }
}

cudaError_t launch_compact_hits(cudaStream_t st, const Candidate *cand, const OverlapOut *ovl, uint64_t n, const unsigned long long *n_dev,
                                double jmin, int keep_all, Candidate *cand_out, OverlapOut *ovl_out, unsigned long long *d_count, int *launches)
{
    if (n == 0) return cudaSuccess;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t grid = std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 8);
    k_compact_hits<<<(unsigned)grid, 256, 0, st>>>(cand, ovl, n, n_dev, jmin, keep_all, cand_out, ovl_out, d_count);
    (*launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// a10: MinHashSketch.jaccard numerator (sketch/MinHashSketch.java:237-263)
// ---------------------------------------------------------------------------------------------
__global__ void k_equal_count(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int H, int32_t *out)
{
    int c = 0;
    for (int i = threadIdx.x; i < H; i += blockDim.x) c += a[i] == b[i];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

cudaError_t launch_equal_count(cudaStream_t st, const int32_t *a, const int32_t *b, int H, int32_t *d_out, int *launches)
{
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    k_equal_count<<<1, 128, 0, st>>>(a, b, H, d_out);
    (*launches)++;
    return cudaGetLastError();
}

} // namespace mhapb
