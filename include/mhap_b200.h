/*
 * mhap_b200.h -- C ABI of the B200-native MinHash overlap engine (libmhap_b200.so).
 *
 * The reference (marbl/MHAP 2.1.3, pure Java) has no FFI; its seams for this path are Java
 * abstract methods and static deserialisers.  Each entry point below names the reference
 * interface it replaces; paths are relative to
 * /root/reference/src/main/java/edu/umd/marbl/mhap/ .  INTEGRATION.md shows the JNI shim and the
 * Java subclass (GpuMinHashSearch extends AbstractMatchSearch) that bind these symbols.
 *
 * Conventions: extern "C"; plain pointers and sizes; every call returns 0 on success and a
 * negative MHAPB_E* code on failure (text via mhapb_last_error); no exception crosses the ABI;
 * the caller owns every buffer it passes in; buffers the library returns are released with
 * mhapb_free.  One context drives one GPU.  Multi-GPU jobs use one context per GPU joined by an NCCL
 * communicator that the library owns (mhapb_comm_*): either one process per GPU (mhapb_comm_init_rank, the
 * collective calls mhapb_dist_*) or one process driving all GPUs (mhapb_multi_*, what a single JVM binds).
 * Calls on one context are serialised by an internal mutex, so the reference's pool threads
 * (AbstractMatchSearch.java:70,124,206) may call concurrently.
 *
 * There is NO CPU fallback: if no CUDA device is usable mhapb_create fails with MHAPB_ENODEV.
 */
#ifndef MHAP_B200_H
#define MHAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHAPB_OK        0
#define MHAPB_EINVAL   -1   /* bad argument (the reference throws MhapRuntimeException / SketchRuntimeException) */
#define MHAPB_ENODEV   -2   /* no usable CUDA device / kernels not loadable */
#define MHAPB_ECUDA    -3   /* CUDA runtime error */
#define MHAPB_ENOMEM   -4
#define MHAPB_EDUPID   -5   /* "Sequence ID already exists in the hash table." impl/MinHashSearch.java:112-117 */
#define MHAPB_ESTATE   -6   /* call out of order (e.g. search before any sequence was added) */
#define MHAPB_ECOMM    -7   /* NCCL not loadable / a collective failed */

typedef struct mhapb_ctx mhapb_ctx;

/* Sketch parameters: the constructor arguments of impl/SequenceSketch.java:106-116
 * (kmerSize, numHashes, orderedKmerSize, orderedSketchSize, repeatWeight) plus the streamer's
 * minOlapLength read filter (impl/SequenceSketchStreamer.java:129-133).  No -f filter file. */
typedef struct {
    int32_t kmer_size;            /* -k, 16 */
    int32_t num_hashes;           /* --num-hashes, 512 (1..2048) */
    int32_t ordered_kmer_size;    /* --ordered-kmer-size, 12 */
    int32_t ordered_sketch_size;  /* --ordered-sketch-size, 1536 */
    int32_t unweighted;           /* 1 <=> --repeat-weight < 0 (classic MinHash, weight 1) */
    int32_t min_olap_length;      /* --min-olap-length, 116: shorter reads are skipped */
} mhapb_sketch_params;

/* Search parameters: impl/MinHashSearch.java:63-76 constructor arguments. */
typedef struct {
    int32_t num_min_matches;      /* --num-min-matches, 3 */
    int32_t min_store_length;     /* --min-store-length, 0 */
    double  max_shift;            /* --max-shift, 0.2 */
    double  accept_score;         /* --threshold, 0.78 */
    int32_t keep_all;             /* 1: also return fully-compared candidates that fail the threshold */
    int32_t reserved;
    int64_t query_first;          /* self search: restrict queries to stored sketches            */
    int64_t query_count;          /*   [query_first, query_first+query_count); count<0 => all    */
} mhapb_search_params;

/* One overlap = the integers behind impl/MatchResult.java:46-65 + impl/OverlapInfo.java:40-50.
 * a1..b2 are in ordered-k-mer units before MatchResult's strand flip; score is
 * BottomOverlapSketch.jaccardToIdentity(intersect/kmin) (sketch/BottomOverlapSketch.java:391-395)
 * evaluated on the host in double precision. */
typedef struct {
    int64_t from_id, to_id;        /* header ids */
    int32_t from_fwd, to_fwd;
    int32_t hit_count;             /* shared min-hashes (HitCounter.count, MinHashSearch.java:176) */
    int32_t a1, a2, b1, b2;
    int32_t valid_count;           /* rawScore */
    int32_t intersect, kmin;       /* bottom-k jaccard numerator / k */
    int32_t from_len, to_len;      /* bases */
    double  score;
    int32_t accepted;              /* score >= accept_score (MinHashSearch.java:229) */
    int32_t pad_;
} mhapb_hit;

/* The order-independent counters MhapMain.outputFinalStat prints (main/MhapMain.java:572-590). */
typedef struct {
    int64_t elements_processed;    /* getNumberElementsProcessed   MinHashSearch.java:188 */
    int64_t sequences_hit;         /* getNumberSequencesHit        MinHashSearch.java:189 */
    int64_t fully_compared;        /* getNumberSequencesFullyCompared MinHashSearch.java:232 */
    int64_t matches_processed;     /* getMatchesProcessed          AbstractMatchSearch.java:161 */
    int64_t sequences_searched;    /* getNumberSequencesSearched   AbstractMatchSearch.java:152 */
} mhapb_stats;

/* Device-time breakdown of the last sketch / search call (CUDA events, milliseconds). */
typedef struct {
    float h2d_ms, d2h_ms;
    float hash_dedup_ms, minhash_ms, ordered_ms;      /* K1a, K1b, K1c */
    float index_ms, probe_ms, filter_ms;              /* K2a, K2b, K2c */
    int64_t kernel_launches;                          /* launches of this library's own kernels */
    int64_t xorshift_steps;                           /* algorithmic XORShift-min steps of the last sketch call */
    float sketch_total_ms;                            /* first K1 launch -> last K1 kernel end, all chunks */
    float search_total_ms;                            /* K2b start -> K2c end (device work only) */
    int64_t kmers_hashed;                             /* k-mers (MinHash k) of the last sketch call */
    float gather_ms;                                  /* multi-GPU: first to last collective of the last mhapb_dist_* call (its own stream) */
    float pad_;
} mhapb_timing;

/* ---- housekeeping ------------------------------------------------------------------------ */
const char *mhapb_version(void);
/* device_id: CUDA ordinal.  Replaces nothing in the reference (single JVM); one per rank. */
int  mhapb_create(int device_id, mhapb_ctx **out);
void mhapb_destroy(mhapb_ctx *ctx);
const char *mhapb_last_error(const mhapb_ctx *ctx);   /* ctx may be NULL: error of the last failed create */
void mhapb_free(void *p);
int  mhapb_get_timing(mhapb_ctx *ctx, mhapb_timing *out);
/* cudaHostAlloc / cudaFreeHost, so callers (JNI direct buffers, the benchmark) can stage reads
 * in pinned memory. */
int  mhapb_host_alloc(size_t bytes, void **out);
void mhapb_host_free(void *p);
/* Micro-benchmarks for the roofline denominator of K1b: the bare XORShift recurrence
 * (x ^= x<<21; x ^= x>>>35; x ^= x<<4) at full occupancy, no compare, no memory, in its two
 * formulations -- scalar (one chain per thread op) and bit-sliced (32 chains per thread, 132 XORs per
 * step).  Sustained chain steps/s of this GPU at its current clocks; _peak returns the larger. */
int  mhapb_xorshift_peaks(mhapb_ctx *ctx, double *scalar_steps_per_s, double *bitsliced_steps_per_s);
int  mhapb_xorshift_peak(mhapb_ctx *ctx, double *steps_per_s);

/* ---- the -f k-mer filter -------------------------------------------------------------------
 * Replaces sketch/FrequencyCounts.java:63-320 (built by main/MhapMain.java:340-372 from the -f file) and the weight
 * rule it feeds, sketch/MinHashSketch.java:66-130:
 *   repeat_weight < 0        weight 1, k-mers present in the repeat map are dropped           (:101-107)
 *   0 <= repeat_weight < 1   weight = max(1, round(tf * scaledIdf(k-mer))), offset = repeat_weight  (:109-124)
 *   repeat_weight >= 1       weight = tf (the filter only acts through --supress-noise 1)
 *   supress_noise 1          k-mers not in the file are removed before counting (keepKmer, :70-71)
 *   supress_noise 2          k-mers not in the file get idf 1 (FrequencyCounts.java:292-293)
 * The filter belongs to the context and applies to every later sketching call (mhapb_sketch*, mhapb_store_add_reads,
 * mhapb_search_query_reads) until mhapb_filter_clear; those calls must pass unweighted = (repeat_weight < 0).
 * With a filter a strand can lose all its k-mers (ZeroNGramsFoundException, MinHashSketch.java:84,156): status 1 when
 * the forward strand is empty (the read is skipped, impl/SequenceSketchStreamer.java:225-240), status 3 when only the
 * reverse strand is (the forward sketch alone is kept).
 * The membership test of --supress-noise is Guava 19.0's BloomFilter (create(funnel putLong, n, 1e-5), strategy
 * MURMUR128_MITZ_64), an un-vendored dependency restated from its published algorithm. */
typedef struct {
    double  filter_cutoff;     /* --filter-threshold, 1e-5: fractions below it are not repeats */
    double  repeat_weight;     /* --repeat-weight, 0.9 */
    double  idf_scale;         /* --repeat-idf-scale, 3.0 (>= 1) */
    int32_t supress_noise;     /* --supress-noise, 0 */
    int32_t no_tf;             /* --no-tf */
} mhapb_filter_params;

/* HashUtils.computeSequenceHashesLong(kmer, kmer.length(), 0, doReverseCompliment)[0] (sketch/HashUtils.java:237-258):
 * the key the filter file's k-mers are stored under (canonical = !--no-rc; reads themselves are never canonicalised). */
int mhapb_kmer_hash(const char *kmer, int32_t len, int canonical, int64_t *out_hash);
/* The state FrequencyCounts' constructor leaves behind: (hash, fraction) pairs of the file (entries below the cutoff
 * are ignored like :183) and, for supress_noise > 0, the Bloom filter's bit array (bloom_bits a multiple of 64). */
int mhapb_filter_set(mhapb_ctx *ctx, const mhapb_filter_params *p, const int64_t *hashes, const double *fractions, uint64_t n,
                     const uint64_t *bloom_words, uint64_t bloom_bits, int32_t bloom_num_hash_functions);
/* Same from the text of a filter file ("<sizeBloom> <sizeRepeat>" then "<k-mer> <fraction> ..." lines): the parsing
 * half of the FrequencyCounts constructor.  n_repeat (optional) = distinct k-mers at or above the cutoff (fractionCounts.size()). */
int mhapb_filter_load_text(mhapb_ctx *ctx, const mhapb_filter_params *p, const char *text, uint64_t len, int canonical, int64_t *n_repeat);
int mhapb_filter_clear(mhapb_ctx *ctx);

/* ---- K1: sketching ------------------------------------------------------------------------
 * Replaces SequenceSketchStreamer.getSketch (impl/SequenceSketchStreamer.java:262-266) applied
 * to a batch of reads, i.e. new SequenceSketch(seq, k, H, ok, os, filter=null, true, repeatWeight)
 * (impl/SequenceSketch.java:106-116) for the read and, when both_strands, its reverse complement
 * (SequenceSketchStreamer.java:147-155, utils/Utils.java:496-507).
 *
 * bases: the reads' characters concatenated (ASCII; lower case is upper-cased like
 * impl/FastaData.java:194; any letter is hashed as its UTF-16 code unit exactly as
 * sketch/HashUtils.java:213-258 does, so N / IUPAC need no special handling);
 * offsets[n_reads+1] delimit them.  Sketch slot j = read*(both_strands?2:1) + (0 fwd | 1 rc).
 *
 * out_minhash    [n_slots][H]      MinHashSketch.minHashes           (sketch/MinHashSketch.java:51-179)
 * out_ord        [n_slots][S][2]   BottomOverlapSketch.orderedHashes (hash,pos) (sketch/BottomOverlapSketch.java:525-559)
 * out_ord_n      [n_slots]         orderedHashes.length = min(S, L-ok+1)
 * out_status     [n_reads]         0 ok; 1 = ZeroNGramsFoundException (read shorter than a k-mer,
 *                                  MinHashSketch.java:55-56 / BottomOverlapSketch.java:530-531);
 *                                  2 = skipped, shorter than min_olap_length.
 * Slots of reads with status != 0 are zero-filled.  Any out pointer may be NULL. */
int mhapb_sketch(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases,
                 const uint64_t *offsets, uint32_t n_reads, int both_strands,
                 int32_t *out_minhash, int32_t *out_ord, int32_t *out_ord_n, int32_t *out_status);

/* Same computation with the bases already resident in HBM and the sketches left there
 * (d_* are device pointers of the sizes above; h_offsets stays on the host).  This is what
 * bench.py times for the device-resident figure and what the multi-GPU path calls per shard
 * before the NCCL all-gather. */
int mhapb_sketch_device(mhapb_ctx *ctx, const mhapb_sketch_params *p, const void *d_bases,
                        const uint64_t *h_offsets, uint32_t n_reads, int both_strands,
                        void *d_minhash, void *d_ord, void *d_ord_n, void *d_status);

/* K1 + big-endian .dat records, the byte format SequenceSketch.fromByteStream reads
 * (impl/SequenceSketch.java:61-96,123-148; framing impl/SequenceSketchStreamer.java:349-360):
 * this is how sketches are handed to Java, whose MinHashSketch(int[]) /
 * BottomOverlapSketch(int,int,int[][]) constructors are private.  ids[i] is the SequenceId
 * number (1-based file position, impl/FastaData.java:181); the header string written is its
 * decimal form (SequenceId.getHeader without --store-full-id).  *out is malloc'd (mhapb_free). */
int mhapb_sketch_to_dat(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases,
                        const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                        int both_strands, uint8_t **out, uint64_t *out_len, uint32_t *n_records);

/* Same with the header string of every record given (headers[i] for read i; NULL entries / NULL array = decimal id):
 * --store-full-id, where SequenceId.getHeader is the FASTA name (impl/SequenceId.java:102-108, impl/FastaData.java:155-156). */
int mhapb_sketch_to_dat_named(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases,
                              const uint64_t *offsets, const int64_t *ids, const char *const *headers, uint32_t n_reads,
                              int both_strands, uint8_t **out, uint64_t *out_len, uint32_t *n_records);

/* ---- .dat codec (host only) ---------------------------------------------------------------
 * Encode one record (returns bytes; buf==NULL sizes it).  header==NULL => decimal id. */
int64_t mhapb_dat_encode(int64_t id, int is_fwd, const char *header, int32_t seq_len,
                         const int32_t *minhash, int32_t num_hashes, int32_t seq_len_kmers,
                         int32_t ordered_kmer_size, const int32_t *ord_hash_pos, int32_t ord_n,
                         uint8_t *buf);
/* Decode a .dat byte stream (SequenceSketchStreamer.java:278-320 + SequenceSketch.fromByteStream)
 * into flat arrays sized for n records with H hashes and at most S ordered entries each; first
 * call with all outputs NULL to get *n_records, *num_hashes, *max_ord.  id_offset is added to the
 * ids like fromByteStream(input, offset). */
int mhapb_dat_decode(const uint8_t *buf, uint64_t len, int64_t id_offset, uint32_t *n_records,
                     int32_t *num_hashes, int32_t *max_ord, int32_t *ordered_kmer_size,
                     int64_t *ids, uint8_t *is_fwd, int32_t *seq_len, int32_t *seq_len_kmers,
                     int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n);

/* ---- K2a: the sketch store and its inverted index -------------------------------------------
 * Replaces MinHashSearch's constructor + addSequence (impl/MinHashSearch.java:63-147): every
 * stored sketch (forward and reverse) is posted under each of its H min-hashes in an
 * open-addressed (word,value) -> posting-list table in HBM. */
int mhapb_store_reset(mhapb_ctx *ctx, const mhapb_sketch_params *p);
/* Sketch reads on the GPU and append them to the store without leaving HBM
 * (enqueueFullFile + addData, MinHashSearch.java:80,95).  ids[i] = SequenceId number of read i. */
int mhapb_store_add_reads(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets,
                          const int64_t *ids, uint32_t n_reads, int both_strands,
                          int64_t *n_added);
/* Same with the reads' characters already resident in HBM (d_bases: device pointer; h_offsets stays on the host): the
 * device-resident figure of bench.py, and callers that stage reads on the GPU themselves. */
int mhapb_store_add_reads_device(mhapb_ctx *ctx, const void *d_bases, const uint64_t *h_offsets, const int64_t *ids,
                                 uint32_t n_reads, int both_strands, int64_t *n_added);
/* Append pre-computed sketches from host arrays (the .dat path, addSequence per record).  num_hashes and
 * ordered_kmer_size describe the arrays (a .dat file does not record the run's parameters, only each sketch's
 * own): a mismatch with the store's parameters fails like the reference does -- "Number of MinHashes of the
 * sequence does not match current settings." (impl/MinHashSearch.java:105-106), "Sketch k-mer size does not
 * match between the two sequences." (sketch/BottomOverlapSketch.java:594-595). */
int mhapb_store_add_sketches(mhapb_ctx *ctx, const int64_t *ids, const uint8_t *is_fwd,
                             const int32_t *seq_len, const int32_t *seq_len_kmers,
                             const int32_t *minhash, int32_t num_hashes, const int32_t *ord_hash_pos,
                             const int32_t *ord_n, int32_t ord_stride, int32_t ordered_kmer_size, uint32_t n);
/* Same with the two big blocks (minhash [n][H], ord [n][S][2]) already on this GPU -- the
 * landing buffers of the NCCL all-gather of sketch blocks; the small per-sketch columns stay on
 * the host. */
int mhapb_store_add_sketches_device(mhapb_ctx *ctx, const int64_t *ids, const uint8_t *is_fwd,
                                    const int32_t *seq_len, const int32_t *seq_len_kmers,
                                    const void *d_minhash, const void *d_ord, const int32_t *ord_n,
                                    uint32_t n);
/* Scratch hint for callers that feed reads in batches (the streaming FASTA producer of the host driver): pre-sizes the
 * per-call device scratch (copy of the reads, k-mer key/weight scratch, strand descriptors) for calls of up to max_bases
 * characters in max_reads reads, so that growing batch sizes do not re-allocate it.  Purely an optimisation. */
int mhapb_sketch_reserve(mhapb_ctx *ctx, const mhapb_sketch_params *p, uint64_t max_bases, uint32_t max_reads, int both_strands);
/* Capacity hint (total sketches the store is expected to hold), like the size argument of the reference's hash maps
 * (impl/MinHashSearch.java:83-90): batches appended later do not have to grow-and-copy the sketch blocks. */
int mhapb_store_reserve(mhapb_ctx *ctx, int64_t n_sketches);
int64_t mhapb_store_size(mhapb_ctx *ctx);                          /* AbstractMatchSearch.size() */
/* Copy stored sketch idx back to the host (getStoredSequenceHash, AbstractMatchSearch.java:314). */
int mhapb_store_get(mhapb_ctx *ctx, int64_t idx, int64_t *id, int32_t *is_fwd, int32_t *seq_len,
                    int32_t *seq_len_kmers, int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n);
/* The parameters the store was configured with (mhapb_store_reset; ordered_kmer_size may come from a .dat store). */
int mhapb_store_params(mhapb_ctx *ctx, mhapb_sketch_params *out);
/* Bulk form: stored sketches [first, first+count) into flat host arrays (minhash [count][H], ord [count][ord_stride][2]
 * with ord_stride from mhapb_store_device_ptrs, rows zero-padded past ord_n).  Any pointer may be NULL. */
int mhapb_store_get_range(mhapb_ctx *ctx, int64_t first, int64_t count, int64_t *ids, uint8_t *is_fwd, int32_t *seq_len,
                          int32_t *seq_len_kmers, int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n);
/* Device views of the store's blocks (for the all-gather): minhash [n][H] and ord [n][S][2]. */
int mhapb_store_device_ptrs(mhapb_ctx *ctx, void **d_minhash, void **d_ord, void **d_ord_n, int64_t *n,
                            int32_t *num_hashes, int32_t *ord_stride);
int mhapb_index_build(mhapb_ctx *ctx);

/* ---- K2b + K2c: search --------------------------------------------------------------------
 * findMatches() to self (impl/AbstractMatchSearch.java:121-199 driving
 * MinHashSearch.findMatches(sketch,true), impl/MinHashSearch.java:150-251): every stored forward
 * sketch queries the index; hit counting, the id / length filters, then
 * BottomOverlapSketch.getOverlapInfo (sketch/BottomOverlapSketch.java:592-630) per candidate.
 * *out is malloc'd (mhapb_free); order of hits is unspecified, as in the reference. */
int mhapb_search_self(mhapb_ctx *ctx, const mhapb_search_params *sp, mhapb_hit **out,
                      uint64_t *n_out, mhapb_stats *stats);
/* findMatches(SequenceSketchStreamer) (AbstractMatchSearch.java:203-285,
 * MinHashSearch.findMatches(sketch,false)): query reads are sketched forward only (:225). */
int mhapb_search_query_reads(mhapb_ctx *ctx, const mhapb_search_params *sp, const char *bases,
                             const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                             mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
/* Same with query sketches from host arrays (a .dat query file).  num_hashes / ordered_kmer_size as in
 * mhapb_store_add_sketches: "Number of hashes does not match. Stored size S, input size I."
 * (impl/MinHashSearch.java:157-159) and the ordered k-mer size check of getOverlapInfo. */
int mhapb_search_query_sketches(mhapb_ctx *ctx, const mhapb_search_params *sp, const int64_t *ids,
                                const uint8_t *is_fwd, const int32_t *seq_len,
                                const int32_t *seq_len_kmers, const int32_t *minhash, int32_t num_hashes,
                                const int32_t *ord_hash_pos, const int32_t *ord_n,
                                int32_t ord_stride, int32_t ordered_kmer_size, uint32_t n, mhapb_hit **out,
                                uint64_t *n_out, mhapb_stats *stats);
/* Query sketches whose blocks are already on this GPU (minhash [n][H], ord [n][ord_stride][2], ord_n [n]
 * device pointers; the small columns on the host) against the store.  to_self=1 applies
 * findMatches(sketch, toSelf=true)'s id rules (MinHashSearch.java:200,215-225) -- the multi-GPU self
 * overlap: every rank stores its own shard and queries it with the all-gathered sketch blocks of every
 * rank, so hit lists are disjoint by target and the index is built once across the job, not per rank. */
int mhapb_search_sketches_device(mhapb_ctx *ctx, const mhapb_search_params *sp, int to_self, const int64_t *ids,
                                 const uint8_t *is_fwd, const int32_t *seq_len, const int32_t *seq_len_kmers,
                                 const void *d_minhash, const void *d_ord, const void *d_ord_n, int32_t ord_stride,
                                 uint32_t n, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
/* impl/MatchResult.java:46-65,98-113: one output line, no newline; returns its length. */
int mhapb_format_match(const mhapb_hit *h, char *buf, size_t buflen);

/* a10: MinHashSketch.jaccard (sketch/MinHashSketch.java:237-263) between stored sketches i and j:
 * positional equality count; *out_equal / H is the reference's double. */
int mhapb_minhash_equal_count(mhapb_ctx *ctx, int64_t i, int64_t j, int32_t *out_equal);

/* ---- multi-GPU ---------------------------------------------------------------------------
 * The reference is one JVM with a thread pool (impl/AbstractMatchSearch.java:67-117,121-199); its users scale out by
 * partitioning the reads by hand (docs/source/quickstart.rst:23).  Here reads are sharded over the GPUs of one box:
 * every context sketches, stores and indexes ITS shard (mhapb_store_add_reads), the forward sketches of all shards --
 * the queries of findMatches -- are exchanged with ONE NCCL all-gather over NVLink (columns + min-hashes first, ordered
 * sketches behind them, hidden behind K2a and K2b), and every rank answers for the targets it stores.  A pair
 * (query, target) is therefore found exactly once, on the target's rank; the counters of MhapMain.outputFinalStat are
 * additive over ranks and are all-reduced, so every rank returns the job-wide mhapb_stats.
 * NCCL is loaded at run time (libnccl.so.2; override with MHAPB_NCCL_LIB): MHAPB_ECOMM if it is missing. */
#define MHAPB_COMM_ID_BYTES 128
/* one process per GPU: rank 0 makes the id (ncclGetUniqueId), the launcher hands it to every rank, each rank joins
 * with its own context.  Collective over the ranks. */
int mhapb_comm_unique_id(uint8_t *id /* [MHAPB_COMM_ID_BYTES] */);
int mhapb_comm_init_rank(mhapb_ctx *ctx, const uint8_t *id, int rank, int nranks);
/* one process, n contexts on n different GPUs: rank i = ctxs[i] */
int mhapb_comm_init_all(mhapb_ctx **ctxs, int n);
int mhapb_comm_destroy(mhapb_ctx *ctx);      /* mhapb_destroy does it too */
int mhapb_comm_info(mhapb_ctx *ctx, int *rank, int *nranks, int *nccl_version);
/* COLLECTIVE (every rank of the communicator calls it, from its own thread / process).  Replaces findMatches() to self
 * (impl/AbstractMatchSearch.java:121-199 -> MinHashSearch.findMatches(sketch,true), impl/MinHashSearch.java:150-251)
 * for a store sharded over the ranks: *out = the overlaps whose TARGET this rank stores, *stats = job-wide counters. */
int mhapb_dist_search_self(mhapb_ctx *ctx, const mhapb_search_params *sp, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
/* COLLECTIVE.  findMatches(SequenceSketchStreamer) (AbstractMatchSearch.java:203-285) with the query file sharded too:
 * this rank passes ITS shard of the query reads (sketched forward only, :225); all queries meet every store shard. */
int mhapb_dist_search_query_reads(mhapb_ctx *ctx, const mhapb_search_params *sp, const char *bases, const uint64_t *offsets,
                                  const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);

/* Same with this rank's query reads already resident in HBM (bench.py's device-resident figure). */
int mhapb_dist_search_query_reads_device(mhapb_ctx *ctx, const mhapb_search_params *sp, const void *d_bases, const uint64_t *h_offsets,
                                         const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);

/* One process driving several GPUs -- what MhapMain's single construction site (main/MhapMain.java:554-558) binds:
 * the same calls as the single-GPU store / search entry points, batches split over the devices by the library (contiguous
 * parts balanced by bases), one host thread per device inside each call, results merged.  n_devices == 1 forwards to the
 * single-GPU calls (no NCCL needed). */
typedef struct mhapb_multi mhapb_multi;
int  mhapb_multi_create(const int *device_ids, int n_devices, mhapb_multi **out);
void mhapb_multi_destroy(mhapb_multi *m);
const char *mhapb_multi_last_error(const mhapb_multi *m);
int  mhapb_multi_n_devices(const mhapb_multi *m);
mhapb_ctx *mhapb_multi_ctx(mhapb_multi *m, int i);      /* e.g. to install the -f filter on every device, or to read timings */
int  mhapb_multi_store_reset(mhapb_multi *m, const mhapb_sketch_params *p);
int  mhapb_multi_store_reserve(mhapb_multi *m, int64_t n_sketches);
int  mhapb_multi_store_add_reads(mhapb_multi *m, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                                 int both_strands, int64_t *n_added);
int  mhapb_multi_store_add_sketches(mhapb_multi *m, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len,
                                    const int32_t *seq_len_kmers, const int32_t *minhash, int32_t num_hashes,
                                    const int32_t *ord_hash_pos, const int32_t *ord_n, int32_t ord_stride,
                                    int32_t ordered_kmer_size, uint32_t n);
int64_t mhapb_multi_store_size(mhapb_multi *m);
int  mhapb_multi_search_self(mhapb_multi *m, const mhapb_search_params *sp, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
int  mhapb_multi_search_query_reads(mhapb_multi *m, const mhapb_search_params *sp, const char *bases, const uint64_t *offsets,
                                    const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
int  mhapb_multi_search_query_sketches(mhapb_multi *m, const mhapb_search_params *sp, const int64_t *ids, const uint8_t *is_fwd,
                                       const int32_t *seq_len, const int32_t *seq_len_kmers, const int32_t *minhash,
                                       int32_t num_hashes, const int32_t *ord_hash_pos, const int32_t *ord_n, int32_t ord_stride,
                                       int32_t ordered_kmer_size, uint32_t n, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
