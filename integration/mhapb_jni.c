/*
 * mhapb_jni.c -- the JNI shim between marbl/MHAP's Java host and libmhap_b200.so (include/mhap_b200.h).
 *
 * One C function per `static native` of integration/MhapB200.java (class edu.umd.marbl.mhap.impl.MhapB200); every one
 * is a thin marshalling layer over ONE family of C-ABI calls, named next to it.  The handle is an mhapb_multi*: one
 * process (the JVM) driving one or several GPUs, NCCL inside the library.
 *
 * Build against a JDK:   gcc -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *                            mhapb_jni.c -L../mhap_b200 -lmhap_b200 -o libmhapb_jni.so
 * In this image there is no JDK: the file is compiled against tests/cpp/jni_stub/jni.h (same names and signatures, see
 * its header) by tests/test_jni_shim.py, and its functions are driven on the GPU through a fake JNIEnv by
 * tests/cpp/jni_harness.c (tests/test_gpu_jni_shim.py).  It has never been loaded into a JVM.
 *
 * Data formats across the boundary:
 *   reads     : direct ByteBuffer of ASCII bases (best from MhapB200.hostAlloc = pinned memory) + long[n+1] offsets + long[n] ids
 *   sketches  : framed .dat records (byte isFwd, int32 payloadBytes, payload = SequenceSketch.getAsByteArray(),
 *               impl/SequenceSketchStreamer.java:349-360) -- what SequenceSketch.fromByteStream reads
 *   overlaps  : byte[] of packed little-endian mhapb_hit structs (80 bytes each), decoded by GpuMinHashSearch.decode
 *   counters  : long[5] = elements processed, sequences hit, fully compared, matches, sequences searched
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "mhap_b200.h"

#define H(h) ((mhapb_multi *)(intptr_t)(h))

static void throw_msg(JNIEnv *env, const char *msg)
{
    /* the reference's unchecked exception (impl/MhapRuntimeException.java:50) */
    jclass ex = (*env)->FindClass(env, "edu/umd/marbl/mhap/impl/MhapRuntimeException");
    if (ex) (*env)->ThrowNew(env, ex, msg ? msg : "mhap-b200 error");
}
static void throw_multi(JNIEnv *env, mhapb_multi *m) { throw_msg(env, mhapb_multi_last_error(m)); }

static jbyteArray hits_to_bytes(JNIEnv *env, mhapb_hit *hits, uint64_t n, const mhapb_stats *st, jlongArray stats)
{
    const uint64_t bytes = n * sizeof(mhapb_hit);
    jbyteArray out = NULL;
    if (bytes > 0x7fffffffull) { mhapb_free(hits); throw_msg(env, "more than 2^31 bytes of overlaps in one call: search in query ranges"); return NULL; }
    out = (*env)->NewByteArray(env, (jsize)bytes);
    if (out && bytes) (*env)->SetByteArrayRegion(env, out, 0, (jsize)bytes, (const jbyte *)hits);
    mhapb_free(hits);
    if (stats && st) {
        jlong s[5];
        s[0] = st->elements_processed; s[1] = st->sequences_hit; s[2] = st->fully_compared; s[3] = st->matches_processed; s[4] = st->sequences_searched;
        (*env)->SetLongArrayRegion(env, stats, 0, 5, s);
    }
    return out;
}

/* framed .dat records -> flat arrays (mhapb_dat_decode); caller frees d->* */
typedef struct {
    uint32_t n; int32_t H, max_ord, ok;
    int64_t *ids; uint8_t *fwd; int32_t *len, *lenk, *mh, *ord, *ordn;
} dat_arrays;

static void dat_free(dat_arrays *d) { free(d->ids); free(d->fwd); free(d->len); free(d->lenk); free(d->mh); free(d->ord); free(d->ordn); memset(d, 0, sizeof *d); }

static int dat_parse(const uint8_t *buf, uint64_t len, int64_t id_offset, dat_arrays *d)
{
    memset(d, 0, sizeof *d);
    if (mhapb_dat_decode(buf, len, id_offset, &d->n, &d->H, &d->max_ord, &d->ok, NULL, NULL, NULL, NULL, NULL, NULL, NULL)) return -1;
    if (d->max_ord < 1) d->max_ord = 1;
    const size_t n = d->n ? d->n : 1;
    d->ids = malloc(n * 8); d->fwd = malloc(n); d->len = malloc(n * 4); d->lenk = malloc(n * 4); d->ordn = malloc(n * 4);
    d->mh = malloc(n * (size_t)(d->H ? d->H : 1) * 4); d->ord = malloc(n * (size_t)d->max_ord * 8);
    if (!d->ids || !d->fwd || !d->len || !d->lenk || !d->ordn || !d->mh || !d->ord) { dat_free(d); return -1; }
    if (mhapb_dat_decode(buf, len, id_offset, &d->n, &d->H, &d->max_ord, &d->ok, d->ids, d->fwd, d->len, d->lenk, d->mh, d->ord, d->ordn)) { dat_free(d); return -1; }
    return 0;
}

/* ---- lifecycle ---------------------------------------------------------------------------------------------- */
/* mhapb_multi_create: one context per listed GPU, NCCL communicator inside the library when there are several */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_create(JNIEnv *env, jclass c, jintArray devices)
{
    (void)c;
    mhapb_multi *m = NULL;
    const jsize n = (*env)->GetArrayLength(env, devices);
    jint *d = (*env)->GetIntArrayElements(env, devices, NULL);
    int rc = mhapb_multi_create((const int *)d, (int)n, &m);
    (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
    if (rc) { throw_msg(env, mhapb_last_error(NULL)); return 0; }
    return (jlong)(intptr_t)m;
}

JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_destroy(JNIEnv *env, jclass c, jlong h)
{
    (void)env; (void)c;
    mhapb_multi_destroy(H(h));
}

/* cudaHostAlloc'd direct buffer: reads staged here go to the GPU without an extra host copy (mhapb_host_alloc) */
JNIEXPORT jobject JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_hostAlloc(JNIEnv *env, jclass c, jlong bytes)
{
    (void)c;
    void *p = NULL;
    if (bytes < 0 || mhapb_host_alloc((size_t)bytes, &p)) { throw_msg(env, "pinned host allocation failed"); return NULL; }
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}

JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_hostFree(JNIEnv *env, jclass c, jobject buf)
{
    (void)c;
    mhapb_host_free((*env)->GetDirectBufferAddress(env, buf));
}

/* ---- store (MinHashSearch constructor + addSequence, impl/MinHashSearch.java:63-147) ---------------------------- */
JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeReset(JNIEnv *env, jclass c, jlong h,
        jint k, jint numHashes, jint ok, jint os, jboolean unweighted, jint minOlap)
{
    (void)c;
    mhapb_sketch_params p = { k, numHashes, ok, os, unweighted ? 1 : 0, minOlap };
    if (mhapb_multi_store_reset(H(h), &p)) throw_multi(env, H(h));
}

/* mhapb_multi_store_add_reads: a batch of FASTA reads, sketched on the GPUs and appended to their stores */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeAddReads(JNIEnv *env, jclass c, jlong h,
        jobject bases, jlongArray offsets, jlongArray ids, jint n, jboolean bothStrands)
{
    (void)c;
    const char *b = (const char *)(*env)->GetDirectBufferAddress(env, bases);
    if (!b && n) { throw_msg(env, "storeAddReads needs a direct ByteBuffer"); return 0; }
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    jlong *id = (*env)->GetLongArrayElements(env, ids, NULL);
    int64_t added = 0;
    int rc = mhapb_multi_store_add_reads(H(h), b, (const uint64_t *)off, (const int64_t *)id, (uint32_t)n, bothStrands ? 1 : 0, &added);
    (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, ids, id, JNI_ABORT);
    if (rc) { throw_multi(env, H(h)); return 0; }
    return (jlong)added;
}

/* mhapb_dat_decode + mhapb_multi_store_add_sketches: sketches that already exist (a .dat store, or addSequence(SequenceSketch)) */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeAddDat(JNIEnv *env, jclass c, jlong h, jbyteArray records, jlong idOffset)
{
    (void)c;
    const jsize len = (*env)->GetArrayLength(env, records);
    jbyte *buf = (*env)->GetByteArrayElements(env, records, NULL);
    dat_arrays d;
    int rc = dat_parse((const uint8_t *)buf, (uint64_t)len, idOffset, &d);
    (*env)->ReleaseByteArrayElements(env, records, buf, JNI_ABORT);
    if (rc) { throw_msg(env, "Unexpected data read error."); return 0; }   /* impl/SequenceSketch.java:80 */
    rc = mhapb_multi_store_add_sketches(H(h), d.ids, d.fwd, d.len, d.lenk, d.mh, d.H, d.ord, d.ordn, d.max_ord, d.ok, d.n);
    const jlong n = d.n;
    dat_free(&d);
    if (rc) { throw_multi(env, H(h)); return 0; }
    return n;
}

JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeSize(JNIEnv *env, jclass c, jlong h)
{
    (void)env; (void)c;
    return (jlong)mhapb_multi_store_size(H(h));
}

/* ids / strands of every stored sketch, device after device (getStoredForwardSequenceIds, AbstractMatchSearch.java:312) */
JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeIds(JNIEnv *env, jclass c, jlong h, jlongArray ids, jbyteArray fwd)
{
    (void)c;
    mhapb_multi *m = H(h);
    jsize at = 0;
    for (int dev = 0; dev < mhapb_multi_n_devices(m); dev++) {
        mhapb_ctx *ctx = mhapb_multi_ctx(m, dev);
        const int64_t n = mhapb_store_size(ctx);
        if (n <= 0) continue;
        int64_t *i64 = malloc((size_t)n * 8); uint8_t *f = malloc((size_t)n);
        if (!i64 || !f || mhapb_store_get_range(ctx, 0, n, i64, f, NULL, NULL, NULL, NULL, NULL)) { free(i64); free(f); throw_msg(env, mhapb_last_error(ctx)); return; }
        (*env)->SetLongArrayRegion(env, ids, at, (jsize)n, (const jlong *)i64);
        (*env)->SetByteArrayRegion(env, fwd, at, (jsize)n, (const jbyte *)f);
        free(i64); free(f);
        at += (jsize)n;
    }
}

/* stored sketch `index` (numbering of storeIds) as one framed .dat record for SequenceSketch.fromByteStream
 * (getStoredSequenceHash, AbstractMatchSearch.java:314): mhapb_store_get + mhapb_dat_encode */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeGetDat(JNIEnv *env, jclass c, jlong h, jlong index)
{
    (void)c;
    mhapb_multi *m = H(h);
    for (int dev = 0; dev < mhapb_multi_n_devices(m); dev++) {
        mhapb_ctx *ctx = mhapb_multi_ctx(m, dev);
        const int64_t n = mhapb_store_size(ctx);
        if (index >= n) { index -= n; continue; }
        int32_t num_hashes = 0, stride = 0;
        if (mhapb_store_device_ptrs(ctx, NULL, NULL, NULL, NULL, &num_hashes, &stride)) break;
        int32_t *mh = malloc((size_t)num_hashes * 4), *ord = malloc((size_t)(stride ? stride : 1) * 8);
        int64_t id = 0; int32_t fwd = 0, len = 0, lenk = 0, on = 0;
        mhapb_sketch_params sp;
        jbyteArray out = NULL;
        if (mh && ord && !mhapb_store_get(ctx, index, &id, &fwd, &len, &lenk, mh, ord, &on) && !mhapb_store_params(ctx, &sp)) {
            const int32_t okk = sp.ordered_kmer_size;
            const int64_t bytes = mhapb_dat_encode(id, fwd, NULL, len, mh, num_hashes, lenk, okk, ord, on, NULL);
            uint8_t *rec = bytes > 0 ? malloc((size_t)bytes) : NULL;
            if (rec) {
                mhapb_dat_encode(id, fwd, NULL, len, mh, num_hashes, lenk, okk, ord, on, rec);
                out = (*env)->NewByteArray(env, (jsize)bytes);
                if (out) (*env)->SetByteArrayRegion(env, out, 0, (jsize)bytes, (const jbyte *)rec);
                free(rec);
            }
        }
        free(mh); free(ord);
        if (!out) throw_msg(env, mhapb_last_error(ctx));
        return out;
    }
    throw_msg(env, "stored sketch index out of range");
    return NULL;
}

/* ---- search ----------------------------------------------------------------------------------------------------- */
/* mhapb_multi_search_self: findMatches() to self (impl/AbstractMatchSearch.java:121-199) */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_searchSelf(JNIEnv *env, jclass c, jlong h,
        jint m, jint minStore, jdouble maxShift, jdouble accept, jlongArray stats)
{
    (void)c;
    mhapb_search_params sp = { m, minStore, maxShift, accept, 0, 0, 0, -1 };
    mhapb_hit *hits = NULL; uint64_t n = 0; mhapb_stats st;
    if (mhapb_multi_search_self(H(h), &sp, &hits, &n, &st)) { throw_multi(env, H(h)); return NULL; }
    return hits_to_bytes(env, hits, n, &st, stats);
}

/* mhapb_search_self restricted to stored sketches [first, first+count) of a single-GPU store:
 * findMatches(sketch, toSelf=true) for ONE stored sequence (impl/MinHashSearch.java:150-251) */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_searchSelfRange(JNIEnv *env, jclass c, jlong h,
        jint m, jint minStore, jdouble maxShift, jdouble accept, jlong first, jlong count, jlongArray stats)
{
    (void)c;
    if (mhapb_multi_n_devices(H(h)) != 1) { throw_msg(env, "per-sequence self search is only available on a single device; use findMatches()"); return NULL; }
    mhapb_ctx *ctx = mhapb_multi_ctx(H(h), 0);
    mhapb_search_params sp = { m, minStore, maxShift, accept, 0, 0, first, count };
    mhapb_hit *hits = NULL; uint64_t n = 0; mhapb_stats st;
    if (mhapb_search_self(ctx, &sp, &hits, &n, &st)) { throw_msg(env, mhapb_last_error(ctx)); return NULL; }
    return hits_to_bytes(env, hits, n, &st, stats);
}

/* mhapb_multi_search_query_reads: findMatches(SequenceSketchStreamer) for a batch of FASTA query reads, sketched
 * forward-only on the GPUs (impl/AbstractMatchSearch.java:203-285) */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_searchQueryReads(JNIEnv *env, jclass c, jlong h,
        jint m, jint minStore, jdouble maxShift, jdouble accept, jobject bases, jlongArray offsets, jlongArray ids, jint n, jlongArray stats)
{
    (void)c;
    const char *b = (const char *)(*env)->GetDirectBufferAddress(env, bases);
    if (!b && n) { throw_msg(env, "searchQueryReads needs a direct ByteBuffer"); return NULL; }
    mhapb_search_params sp = { m, minStore, maxShift, accept, 0, 0, 0, -1 };
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    jlong *id = (*env)->GetLongArrayElements(env, ids, NULL);
    mhapb_hit *hits = NULL; uint64_t nh = 0; mhapb_stats st;
    int rc = mhapb_multi_search_query_reads(H(h), &sp, b, (const uint64_t *)off, (const int64_t *)id, (uint32_t)n, &hits, &nh, &st);
    (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, ids, id, JNI_ABORT);
    if (rc) { throw_multi(env, H(h)); return NULL; }
    return hits_to_bytes(env, hits, nh, &st, stats);
}

/* mhapb_dat_decode + mhapb_multi_search_query_sketches: queries that are already sketches (a .dat query file, or
 * SequenceSketch objects dequeued from any SequenceSketchStreamer) */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_searchQueryDat(JNIEnv *env, jclass c, jlong h,
        jint m, jint minStore, jdouble maxShift, jdouble accept, jbyteArray records, jlong idOffset, jlongArray stats)
{
    (void)c;
    const jsize len = (*env)->GetArrayLength(env, records);
    jbyte *buf = (*env)->GetByteArrayElements(env, records, NULL);
    dat_arrays d;
    int rc = dat_parse((const uint8_t *)buf, (uint64_t)len, idOffset, &d);
    (*env)->ReleaseByteArrayElements(env, records, buf, JNI_ABORT);
    if (rc) { throw_msg(env, "Unexpected data read error."); return NULL; }
    mhapb_search_params sp = { m, minStore, maxShift, accept, 0, 0, 0, -1 };
    mhapb_hit *hits = NULL; uint64_t nh = 0; mhapb_stats st;
    rc = mhapb_multi_search_query_sketches(H(h), &sp, d.ids, d.fwd, d.len, d.lenk, d.mh, d.H, d.ord, d.ordn, d.max_ord, d.ok, d.n, &hits, &nh, &st);
    dat_free(&d);
    if (rc) { throw_multi(env, H(h)); return NULL; }
    return hits_to_bytes(env, hits, nh, &st, stats);
}

/* ---- sketches for Java (-p mode; SequenceSketchStreamer.getSketch, impl/SequenceSketchStreamer.java:262-266) -------- */
/* mhapb_sketch_to_dat on the first device: framed .dat records of the batch, in read order */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_sketchToDat(JNIEnv *env, jclass c, jlong h,
        jint k, jint numHashes, jint ok, jint os, jboolean unweighted, jint minOlap,
        jobject bases, jlongArray offsets, jlongArray ids, jint n, jboolean bothStrands)
{
    (void)c;
    mhapb_ctx *ctx = mhapb_multi_ctx(H(h), 0);
    const char *b = (const char *)(*env)->GetDirectBufferAddress(env, bases);
    if (!b && n) { throw_msg(env, "sketchToDat needs a direct ByteBuffer"); return NULL; }
    mhapb_sketch_params p = { k, numHashes, ok, os, unweighted ? 1 : 0, minOlap };
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    jlong *id = (*env)->GetLongArrayElements(env, ids, NULL);
    uint8_t *blob = NULL; uint64_t len = 0; uint32_t nrec = 0;
    int rc = mhapb_sketch_to_dat(ctx, &p, b, (const uint64_t *)off, (const int64_t *)id, (uint32_t)n, bothStrands ? 1 : 0, &blob, &len, &nrec);
    (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, ids, id, JNI_ABORT);
    if (rc) { throw_msg(env, mhapb_last_error(ctx)); return NULL; }
    jbyteArray out = NULL;
    if (len > 0x7fffffffull) throw_msg(env, "more than 2^31 bytes of .dat records in one call: use smaller batches");
    else {
        out = (*env)->NewByteArray(env, (jsize)len);
        if (out && len) (*env)->SetByteArrayRegion(env, out, 0, (jsize)len, (const jbyte *)blob);
    }
    mhapb_free(blob);
    return out;
}

/* ---- the -f k-mer filter (main/MhapMain.java:340-372, sketch/FrequencyCounts.java) -------------------------------- */
/* mhapb_filter_load_text on every device; text = the (decompressed) bytes of the filter file */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_filterLoadText(JNIEnv *env, jclass c, jlong h, jbyteArray text,
        jdouble filterCutoff, jdouble repeatWeight, jdouble idfScale, jint supressNoise, jboolean noTf, jboolean canonical)
{
    (void)c;
    mhapb_filter_params p = { filterCutoff, repeatWeight, idfScale, supressNoise, noTf ? 1 : 0 };
    const jsize n = (*env)->GetArrayLength(env, text);
    jbyte *b = (*env)->GetByteArrayElements(env, text, NULL);
    int64_t n_repeat = 0;
    int rc = 0;
    mhapb_ctx *bad = NULL;
    for (int dev = 0; dev < mhapb_multi_n_devices(H(h)) && !rc; dev++) {
        bad = mhapb_multi_ctx(H(h), dev);
        rc = mhapb_filter_load_text(bad, &p, (const char *)b, (uint64_t)n, canonical ? 1 : 0, &n_repeat);
    }
    (*env)->ReleaseByteArrayElements(env, text, b, JNI_ABORT);
    if (rc) { throw_msg(env, mhapb_last_error(bad)); return 0; }
    return (jlong)n_repeat;
}

JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_filterClear(JNIEnv *env, jclass c, jlong h)
{
    (void)env; (void)c;
    for (int dev = 0; dev < mhapb_multi_n_devices(H(h)); dev++) mhapb_filter_clear(mhapb_multi_ctx(H(h), dev));
}
