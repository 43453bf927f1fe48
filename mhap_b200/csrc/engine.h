// engine.h -- internal interface between the C-ABI layer (api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mhapb {

// One strand of one read to be sketched.  `row` is the output row (sketch slot / store index).
struct StrandDesc {
    uint64_t base_off;   // offset of the read's first char in the bases buffer
    uint64_t koff;       // offset of this strand's region in the k-mer key scratch
    uint32_t len;        // read length in bases
    uint32_t row;        // output row
    uint32_t rc;         // 1: sketch the reverse complement (Sequence.getReverseCompliment)
    uint32_t slot;       // caller's sketch slot (read * strands + strand), for per-strand status read-back
};

// ---- K1 ------------------------------------------------------------------------------------
// Strands with nk = len-k+1 <= kShortMaxKmers use the shared-memory dedup table; longer ones the
// global-memory table.
constexpr int kShortMaxKmers = 16384;
constexpr int kMaxNumHashes = 2048;
constexpr int kMaxOrderedSketch = 4096;
// slots of the K1a de-duplication table for a strand of nk k-mers: load factor <= 0.8, so two 10 kbp CTAs
// (table + staged characters) fit one SM's shared memory and overlap each other's barrier phases
// (odd, so that the power-of-two probe strides of the double hashing below visit every slot)
__host__ __device__ inline uint32_t dedup_table_slots(uint32_t nk) { return (nk + nk / 4 + 8) | 1u; }

struct SketchScratch {
    uint64_t *keys;      // [cap_kmers] distinct k-mer hashes per strand: light from the front, heavy from the back
    uint32_t *wts;       // [cap_kmers] weights of the heavy keys (same index as keys)
    int32_t  *nlight;    // [n_strands]
    int32_t  *nheavy;    // [n_strands]
    uint32_t *dupcnt;    // [grid][table_cap] zero between uses
    uint64_t *gtable;    // [grid_long][long_cap] global dedup tables (long strands only)
    uint32_t *ohash;     // [grid][ohash_cap] ordered-hash staging for long strands
    uint32_t *counters;  // work-queue counters
};

// Device view of the -f k-mer filter (sketch/FrequencyCounts.java) and of the weight rule of
// sketch/MinHashSketch.java:95-130.  mode: 0 no filter (weight = count, or 1 when unweighted),
// 1 repeatWeight<0 with a filter (weight 1, popular k-mers dropped :101-107), 2 tf-idf (:109-124),
// 3 repeatWeight>=1 with a filter (tf only).  The map holds scaledIdf(key) (FrequencyCounts.java:285-309)
// precomputed on the host in double precision; absent keys get `range`.
struct KmerFilterView {
    const uint64_t *map_keys;   // open addressing (linear probing), empty slots hold kFilterEmpty with map_used bit clear
    const double   *map_idf;
    const uint32_t *map_used;   // bitmap of occupied slots
    uint32_t map_mask;          // capacity - 1 (power of two); 0 entries => map_keys == nullptr
    const uint64_t *bloom;      // Guava BloomFilter bit array (LockFreeBitArray words), nullptr when --supress-noise 0
    uint64_t bloom_bits;
    int32_t bloom_nfun;
    int32_t mode;
    int32_t remove_unique;      // --supress-noise
    int32_t no_tf;
    double  range;              // --repeat-idf-scale
    uint32_t light_weight;      // the weight of a k-mer seen once that is not in the map: keys of this weight are "light"
};

// K1a: hash every k-mer (MurmurHash3_x64_128 h1), de-duplicate with counts, apply the weight rule.
// descs [s_base, s_base + n_strands) of d_desc; queues: two zeroed work-queue counters (short, long) of this launch pair
cudaError_t launch_hash_dedup(cudaStream_t st, const uint8_t *d_bases, const StrandDesc *d_desc, int s_base, int n_strands,
                              int first_long /* descs [s_base + first_long, s_base + n) are long strands */, int max_kmers_short,
                              int max_kmers_long, int k, int unweighted, const KmerFilterView &filter,
                              const SketchScratch &sc, uint32_t *queues, int *launches);
// K1b: H-step XORShift chain per distinct k-mer (light keys advance light_weight steps per word, heavy keys their
// own weight), per-word signed minimum -> minhash rows.
cudaError_t launch_minhash(cudaStream_t st, const StrandDesc *d_desc, int n_strands, int k, int H,
                           const SketchScratch &sc, int32_t *d_minhash, uint32_t light_weight, uint32_t *queue, int *launches);
// sketches wider than 512 words: chain starts of the later word blocks + the launch over virtual strands (sketch.cu)
cudaError_t launch_advance_keys(cudaStream_t st, const StrandDesc *d_desc, int n_strands, int k, const SketchScratch &sc, const uint64_t *d_t512,
                                uint64_t total_k, int passes, uint32_t light_w, int *launches);
cudaError_t launch_minhash_virtual(cudaStream_t st, const StrandDesc *d_vdesc, int n_virtual, int n_real, int k, int hstride,
                                   const SketchScratch &sc, int32_t *d_minhash, uint32_t light_w, uint32_t *queue, int *launches);
// K1c: MurmurHash3_x86_32 of every ordered k-mer, bottom-S by (signed hash, position), sorted.
cudaError_t launch_ordered(cudaStream_t st, const uint8_t *d_bases, const StrandDesc *d_desc, int s_base, int n_strands,
                           int first_long, int max_len_short, int max_len_long, int ok, int S, int ord_stride,
                           const SketchScratch &sc, int32_t *d_ord, int32_t *d_ord_n, int max_ctas_per_sm /* 0: default */, uint32_t *queues, int *launches);
// independent XORShift chains at full occupancy: the integer-issue ceiling K1b is measured against
cudaError_t launch_xorshift_peak(cudaStream_t st, unsigned long long *d_sink, double *steps);
cudaError_t launch_xorshift_peak_bs(cudaStream_t st, unsigned long long *d_sink, double *steps);
int hash_dedup_grid();
int ordered_grid();
size_t dedup_table_cap_short();   // slots per block in SketchScratch.dupcnt

// ---- K2 ------------------------------------------------------------------------------------
struct IndexView {
    uint64_t *slots;      // [H][capw]  (value | begin<<32), empty = ~0
    uint32_t *postings;   // [n_store*H] sketch index, MSB set on the last entry of a bucket
    uint32_t *present;    // [H*capw/32] one bit per slot: occupied.  32 MB for 2*10^5 sketches -- it stays in L2, so a probe
                          // that lands on an empty slot (most probes once the index is sharded over several GPUs) costs no DRAM
    int       log2capw;
    int       H;
    int64_t   n_store;
    uint32_t *overflow;      // set by the build when a sub-table was sized too small (a bucket needed > kIdxMaxProbe probes): the
                             // index is then incomplete, K2b does nothing, and the host rebuilds with the safe size (2 slots per sketch)
    int       use_present;   // probe: consult `present` first (off when the store queries itself: every probe hits, the bitmap is pure overhead)
};

cudaError_t launch_index_build(cudaStream_t st, const int32_t *d_minhash, int64_t n_store, int H, IndexView iv,
                               uint32_t *d_tmp_start /*[H*capw]*/, uint32_t *d_block_sums, int *launches);

struct Candidate { uint32_t q; uint32_t t; uint32_t count; };

struct ProbeArgs {
    const int32_t *q_minhash;  // [nq][H]
    const int64_t *q_id;       // [nq]
    const int32_t *q_len;      // [nq] bases
    const uint32_t *q_list;    // [nq_list] indices into the query arrays (NULL: 0..nq-1)
    int64_t nq_list;
    const int64_t *t_id;       // [n_store]
    const int32_t *t_len;      // [n_store]
    int to_self, num_min_matches, min_store_length;
    Candidate *cand; uint64_t cand_cap;
    unsigned long long *counters;  // [0]=n_cand [1]=elements_processed [2]=sequences_hit [3]=n queries in ovf_q
    uint32_t *ovf_q;               // [nq_list] queries whose hit table overflowed in the first pass
};
cudaError_t launch_probe(cudaStream_t st, IndexView iv, ProbeArgs a, int *launches);

struct OverlapOut { int32_t a1, a2, b1, b2, valid, inter, kmin, empty; };

struct FilterArgs {
    const Candidate *cand; uint64_t n_cand;
    const unsigned long long *n_cand_dev; uint64_t cand_cap;   // when set: the count is read on the device (min(*n_cand_dev, cand_cap))
    const unsigned long long *n_sel_dev;                       // thread-per-candidate kernel: count of sel[] read on the device
    const int32_t *q_ord; const int32_t *q_ord_n; const int32_t *q_lenk; int q_stride;   // [nq][stride][2]
    const int32_t *t_ord; const int32_t *t_ord_n; const int32_t *t_lenk; int t_stride;
    double max_shift;
    int32_t *scratch; uint32_t scratch_entries; uint32_t n_threads;   // 3 arrays [entries][n_threads] (thread-per-candidate kernel)
    OverlapOut *out;
    const uint32_t *sel; uint64_t n_sel;          // thread-per-candidate kernel: only candidates sel[0..n_sel) (NULL: all)
    uint32_t *ovf_list; unsigned long long *ovf_count;   // warp kernel: candidates it hands to the thread-per-candidate kernel
    int prefetch;                                        // warp kernel: pull the next pair's sketches towards L2 (set by the launcher)
};
// thread-per-candidate kernel (any sketch size, any match count; serial merge per thread)
cudaError_t launch_filter(cudaStream_t st, FilterArgs a, int *launches);
// warp-per-candidate kernel (hash-range-partitioned merge, any sketch size; pairs with more than 512 match records are
// handed to launch_filter through ovf_list / ovf_count)
cudaError_t launch_filter_warp(cudaStream_t st, FilterArgs a, int *launches);

// Compact the (candidate, overlap) pairs that can still pass the score threshold: non-empty overlaps whose bottom-k
// jaccard inter/kmin is >= jmin (a bound the caller lowers by a safety margin; the exact double-precision score
// test stays on the host).  keep_all copies everything.  d_count: 64-bit cursor, zero on entry.
cudaError_t launch_compact_hits(cudaStream_t st, const Candidate *cand, const OverlapOut *ovl, uint64_t n, const unsigned long long *n_dev,
                                double jmin, int keep_all, Candidate *cand_out, OverlapOut *ovl_out, unsigned long long *d_count, int *launches);

cudaError_t launch_equal_count(cudaStream_t st, const int32_t *a, const int32_t *b, int H, int32_t *d_out, int *launches);

} // namespace mhapb
