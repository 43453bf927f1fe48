/*
 * synth.c -- deterministic synthetic PacBio-shape reads.
 *
 * Not part of the reference path: the reference's simulator (main/KmerStatSimulator.java:230,
 * utils/RandomSequenceGenerator.java:93-96) only lends its error mix (ins/del/sub =
 * 0.792/0.122/0.086 of the per-base error rate).  The generator is ours: genome = uniform ACGT
 * from splitmix64(genome_seed); read i = window at a uniform start, per-base errors, exactly L
 * bases, strand by coin flip.  Each read has its own PRNG stream so any shard of reads can be
 * generated independently (multi-GPU ranks generate only their shard).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

void mhapb_synth_genome(uint64_t seed, uint64_t len, char *out)
{
    static const char B[4] = { 'A', 'C', 'G', 'T' };
    uint64_t s = seed;
    uint64_t i = 0;
    while (i < len) {
        uint64_t r = splitmix64(&s);
        for (int j = 0; j < 32 && i < len; j++, i++) { out[i] = B[r & 3]; r >>= 2; }
    }
}

static void one_read(const char *genome, uint64_t glen, uint64_t read_seed, uint64_t idx, uint32_t L, double err, char *out)
{
    static const char B[4] = { 'A', 'C', 'G', 'T' };
    uint64_t s = read_seed ^ ((idx + 1) * 0xD1B54A32D192ED03ULL);
    (void)splitmix64(&s);
    uint64_t pos = splitmix64(&s) % glen;
    int rev = (int)(splitmix64(&s) & 1);
    uint32_t n = 0;
    while (n < L) {
        char src = genome[pos];
        double u = u01(&s);
        if (u < err) {
            double v = u / err; /* reuse the draw: uniform in [0,1) given u<err */
            if (v < 0.792) { /* insertion: emit a random base, keep the source base */
                out[n++] = B[splitmix64(&s) & 3];
                continue;
            } else if (v < 0.792 + 0.122) { /* deletion */
                pos = pos + 1 == glen ? 0 : pos + 1;
                continue;
            } else { /* substitution by one of the other three bases */
                int c = (src == 'A') ? 0 : (src == 'C') ? 1 : (src == 'G') ? 2 : 3;
                out[n++] = B[(c + 1 + (int)(splitmix64(&s) % 3)) & 3];
                pos = pos + 1 == glen ? 0 : pos + 1;
                continue;
            }
        }
        out[n++] = src;
        pos = pos + 1 == glen ? 0 : pos + 1;
    }
    if (rev) {
        for (uint32_t i = 0, j = L - 1; i < j; i++, j--) { char t = out[i]; out[i] = out[j]; out[j] = t; }
        for (uint32_t i = 0; i < L; i++) {
            char c = out[i];
            out[i] = (c == 'A') ? 'T' : (c == 'C') ? 'G' : (c == 'G') ? 'C' : 'A';
        }
    }
}

typedef struct {
    const char *genome; uint64_t glen, read_seed, first, n; uint32_t L; double err; char *out;
    int tid, nthreads;
} synth_job;

static void *synth_worker(void *arg)
{
    synth_job *j = (synth_job *)arg;
    for (uint64_t i = (uint64_t)j->tid; i < j->n; i += (uint64_t)j->nthreads)
        one_read(j->genome, j->glen, j->read_seed, j->first + i, j->L, j->err, j->out + i * (uint64_t)j->L);
    return NULL;
}

/* out must hold n_reads*L bytes; reads [first_read, first_read+n_reads) of the stream read_seed. */
void mhapb_synth_reads(const char *genome, uint64_t genome_len, uint64_t read_seed, uint64_t first_read,
                       uint64_t n_reads, uint32_t L, double err, int threads, char *out)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    synth_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < threads; t++) {
        synth_job j = { genome, genome_len, read_seed, first_read, n_reads, L, err, out, t, threads };
        jobs[t] = j;
    }
    if (threads == 1) { synth_worker(&jobs[0]); return; }
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, synth_worker, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}
