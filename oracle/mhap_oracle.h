/*
 * mhap_oracle.h -- CPU restatement of MHAP's sketch + overlap-search path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (mhap_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (marbl/MHAP 2.1.3, Java) ships no tests or golden vectors
 * for this path, no JVM exists in this image, and the hash / sort arithmetic lives in two
 * un-vendored Maven dependencies (Guava 19.0: Hashing.murmur3_128 / murmur3_32,
 * Hasher.putUnencodedChars; fastutil 7.0.12: IntArrays.radixSortIndirect(stable),
 * Long2ObjectLinkedOpenHashMap).  Their published algorithms are restated here and pinned by
 * the public MurmurHash3 known-answer vectors (tests/golden/murmur3_kat.json); everything
 * else follows the reference Java source line by line (citations at each function, paths
 * relative to /root/reference/src/main/java/edu/umd/marbl/mhap/).
 */
#ifndef MHAP_ORACLE_H
#define MHAP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- hashing (Guava 19.0 semantics) ---------------------------------------------------- */
void     mo_murmur3_x64_128(const uint8_t *data, size_t len, uint32_t seed, uint64_t out[2]);
uint32_t mo_murmur3_x86_32(const uint8_t *data, size_t len, uint32_t seed);

/* utils/Utils.java:84-114,496-507 : reverse complement (IUPAC aware, unknown chars kept). */
void mo_rc(const char *seq, int64_t len, char *out);
/* utils/Utils.java:445-494 : literal quickSelect (modifies array). */
int32_t mo_quick_select(int32_t *array, int32_t k, int32_t length);

/* sketch/HashUtils.java:237-258 : h1 of murmur3_128(seed) over each k-mer's UTF-16LE bytes.
 * canonical!=0 enables the doReverseCompliment branch (dead on the main path). */
int64_t mo_kmer_hashes_long(const char *seq, int64_t len, int k, uint32_t seed, int canonical,
                            int64_t *out);
/* sketch/HashUtils.java:213-235 : murmur3_32(0) over each k-mer's UTF-16LE bytes. */
int64_t mo_kmer_hashes_int(const char *seq, int64_t len, int k, int canonical, int32_t *out);

/* sketch/MinHashSketch.java:51-179 (no -f filter).  unweighted!=0 <=> repeatWeight<0.
 * returns 0 ok, 1 = ZeroNGramsFoundException. */
int mo_minhash_sketch(const char *seq, int64_t len, int k, int num_hashes, int unweighted,
                      int32_t *out_hashes);

/* ---- the -f k-mer filter (sketch/FrequencyCounts.java:63-320) --------------------------- */
/* Guava 19.0 BloomFilter (create(funnel putLong, expectedInsertions, 1e-5), strategy MURMUR128_MITZ_64) is a
 * further un-vendored dependency, restated from its published algorithm for --supress-noise 1/2. */
typedef struct mo_filter mo_filter;
typedef struct {
    double  filter_cutoff;   /* --filter-threshold (1e-5), FrequencyCounts ctor :63 */
    double  offset;          /* main/MhapMain.java:348-350: repeatWeight when 0<=repeatWeight<1, else 0 */
    double  range;           /* --repeat-idf-scale (3.0) */
    int32_t remove_unique;   /* --supress-noise 0/1/2 */
    int32_t no_tf;           /* --no-tf */
    int32_t canonical;       /* doReverseCompliment = !--no-rc (filter k-mers ARE canonicalised, reads are not) */
} mo_filter_params;
/* text = the whole filter file: first line "<sizeBloom> <sizeRepeat>", then "<k-mer> <fraction> [..]" lines. */
mo_filter *mo_filter_parse(const char *text, int64_t len, const mo_filter_params *p);
void       mo_filter_free(mo_filter *f);
int64_t    mo_filter_size(const mo_filter *f);                           /* entries with fraction >= cutoff */
double     mo_filter_max_value(const mo_filter *f);
int64_t    mo_filter_export(const mo_filter *f, int64_t *hashes, double *fractions);   /* the map, any order */
int64_t    mo_filter_bloom_export(const mo_filter *f, const uint64_t **words, int32_t *num_hash_functions); /* returns bit size (0 = none) */
int        mo_filter_is_popular(const mo_filter *f, int64_t hash);      /* :262 */
int        mo_filter_keep_kmer(const mo_filter *f, int64_t hash);       /* :267 */
double     mo_filter_scaled_idf(const mo_filter *f, int64_t hash);      /* :285-309 */
/* sketch/MinHashSketch.java:51-179 with a filter (f may be NULL) and the full repeatWeight semantics. */
int mo_minhash_sketch_filtered(const char *seq, int64_t len, int k, int num_hashes, double repeat_weight,
                               const mo_filter *f, int32_t *out_hashes);

/* sketch/BottomOverlapSketch.java:525-559.  out_hash_pos = [n][2]; returns n (or -1 zero n-grams),
 * *seq_len_kmers = len-ok+1. */
int32_t mo_bottom_sketch(const char *seq, int64_t len, int ok, int sketch_size,
                         int32_t *out_hash_pos, int32_t *seq_len_kmers);

/* sketch/BottomOverlapSketch.java:592-630 result, integers only plus the score. */
typedef struct {
    int32_t empty;       /* 1 => OverlapInfo.EMPTY */
    int32_t a1, a2, b1, b2;
    int32_t valid_count; /* rawScore */
    int32_t intersect;   /* bottom-k intersection count */
    int32_t kmin;        /* k of the bottom-k jaccard (0 => jaccard 0) */
    double  score;       /* jaccardToIdentity(intersect/kmin, ok) */
} mo_overlap;

void mo_overlap_info(const int32_t *a_hash_pos, int32_t a_n, int32_t a_seqlen,
                     const int32_t *b_hash_pos, int32_t b_n, int32_t b_seqlen,
                     int ordered_kmer_size, double max_shift, mo_overlap *out);
double mo_jaccard_to_identity(double jaccard, int kmer_size);

/* ---- sketch store + search (impl/MinHashSearch.java, impl/AbstractMatchSearch.java) ------ */
typedef struct mo_store mo_store;

typedef struct {
    int32_t kmer_size, num_hashes, ordered_kmer_size, ordered_sketch_size;
    int32_t unweighted;        /* --repeat-weight < 0 */
    int32_t min_olap_length;   /* --min-olap-length (116) */
} mo_sketch_params;

typedef struct {
    int32_t num_min_matches;   /* 3 */
    int32_t min_store_length;  /* 0 */
    double  max_shift;         /* 0.2 */
    double  accept_score;      /* 0.78 */
} mo_search_params;

typedef struct {
    int64_t from_id, to_id;        /* header ids (1-based file positions) */
    int32_t from_fwd, to_fwd;
    int32_t hit_count;             /* shared min-mers */
    int32_t a1, a2, b1, b2;        /* k-mer units, before MatchResult strand flip */
    int32_t valid_count, intersect, kmin;
    int32_t from_len, to_len;      /* bases */
    double  score;
    int32_t accepted;              /* score >= accept_score */
} mo_hit;

typedef struct {
    int64_t elements_processed;    /* numberElementsProcessed */
    int64_t sequences_hit;         /* numberSequencesHit */
    int64_t fully_compared;        /* numberSequencesFullyCompared */
    int64_t matches_processed;     /* matchesProcessed */
    int64_t sequences_searched;    /* sequencesSearched */
} mo_stats;

mo_store *mo_store_new(const mo_sketch_params *p);
void      mo_store_free(mo_store *s);
/* sketches added afterwards use the filter and repeat_weight (the store does not own f). */
void      mo_store_set_filter(mo_store *s, const mo_filter *f, double repeat_weight);
/* Sketch reads (concatenated ASCII, offsets[n+1]); ids[i] = header id.  both_strands: fwd then rc
 * per read (SequenceSketchStreamer.java:123-156).  Reads shorter than min_olap_length are skipped.
 * threads>1 splits reads over a pthread pool.  Returns number of sketches appended. */
int64_t   mo_store_add_reads(mo_store *s, const char *bases, const uint64_t *offsets,
                             const int64_t *ids, int64_t n_reads, int both_strands, int threads);
/* Append one pre-computed sketch (the .dat path). */
int       mo_store_add_sketch(mo_store *s, int64_t id, int is_fwd, int32_t seq_len,
                              const int32_t *minhash, int32_t seq_len_kmers,
                              const int32_t *ord_hash_pos, int32_t ord_n);
int64_t   mo_store_size(const mo_store *s);
/* accessors */
int       mo_store_get(const mo_store *s, int64_t idx, int64_t *id, int32_t *is_fwd, int32_t *seq_len,
                       const int32_t **minhash, int32_t *seq_len_kmers, const int32_t **ord_hash_pos,
                       int32_t *ord_n);
/* Build the inverted index over every sketch currently in the store (MinHashSearch.java:101-147). */
void      mo_store_build_index(mo_store *s);
/* findMatches() to self (AbstractMatchSearch.java:121-199): queries = forward sketches of the store.
 * keep_all!=0 returns every fully-compared candidate (accepted flag set), else only accepted ones.
 * Result array is malloc'd; free with mo_free. */
int       mo_search_self(mo_store *s, const mo_search_params *sp, int threads, int keep_all,
                         mo_hit **out, int64_t *n_out, mo_stats *stats);
int       mo_search_self_range(mo_store *s, int64_t first, int64_t count, const mo_search_params *sp, int threads,
                               int keep_all, mo_hit **out, int64_t *n_out, mo_stats *stats);
/* findMatches(streamer) (AbstractMatchSearch.java:203-285): queries come from another store. */
int       mo_search_query(mo_store *s, const mo_store *queries, const mo_search_params *sp, int threads,
                          int keep_all, mo_hit **out, int64_t *n_out, mo_stats *stats);
int       mo_search_query_self(mo_store *s, const mo_store *queries, const mo_search_params *sp, int threads,
                               int keep_all, mo_hit **out, int64_t *n_out, mo_stats *stats);
void      mo_free(void *p);

/* impl/MatchResult.java:46-65,98-113 : one output line (no newline); returns strlen. */
int mo_format_match(const mo_hit *h, char *buf, size_t buflen);

/* .dat record codec (SequenceSketchStreamer.java:349-360, SequenceSketch.java:123-148,
 * MinHashSketch.java:218-230, BottomOverlapSketch.java:561-585).  Returns bytes written
 * (call with buf==NULL to size). header==NULL => decimal id. */
int64_t mo_dat_encode(int64_t id, int is_fwd, const char *header, int32_t seq_len,
                      const int32_t *minhash, int32_t num_hashes, int32_t seq_len_kmers,
                      int32_t ordered_kmer_size, const int32_t *ord_hash_pos, int32_t ord_n,
                      uint8_t *buf);

#ifdef __cplusplus
}
#endif
#endif
