/* JNI shim stub for libmhap_b200.so -- see INTEGRATION.md.  NOT compiled in this image (no JDK / jni.h). */
#include <jni.h>
#include "mhap_b200.h"

static void throw_mhap(JNIEnv *env, mhapb_ctx *ctx) {
    jclass ex = (*env)->FindClass(env, "edu/umd/marbl/mhap/impl/MhapRuntimeException");
    (*env)->ThrowNew(env, ex, mhapb_last_error(ctx));
}

JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_create(JNIEnv *env, jclass c, jint dev) {
    mhapb_ctx *ctx = NULL;
    if (mhapb_create(dev, &ctx)) { throw_mhap(env, NULL); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeReset(JNIEnv *env, jclass c, jlong h,
        jint k, jint H, jint ok, jint os, jboolean unweighted, jint minOlap) {
    mhapb_sketch_params p = { k, H, ok, os, unweighted ? 1 : 0, minOlap };
    if (mhapb_store_reset((mhapb_ctx *)(intptr_t)h, &p)) throw_mhap(env, (mhapb_ctx *)(intptr_t)h);
}

/* bases: direct ByteBuffer (ideally from mhapb_host_alloc so the H2D copy is pinned); offsets: long[n+1] */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_storeAddReads(JNIEnv *env, jclass c, jlong h,
        jobject bases, jlongArray offsets, jlongArray ids, jint n, jboolean bothStrands) {
    mhapb_ctx *ctx = (mhapb_ctx *)(intptr_t)h;
    const char *b = (*env)->GetDirectBufferAddress(env, bases);
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    jlong *id = (*env)->GetLongArrayElements(env, ids, NULL);
    int64_t added = 0;
    int rc = mhapb_store_add_reads(ctx, b, (const uint64_t *)off, (const int64_t *)id, (uint32_t)n, bothStrands, &added);
    (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, ids, id, JNI_ABORT);
    if (rc) { throw_mhap(env, ctx); return 0; }
    return added;
}

/* returns the hits as one byte[] of packed mhapb_hit structs (80 bytes each, little-endian);
 * stats[5] receives the counters */
JNIEXPORT jbyteArray JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_searchSelf(JNIEnv *env, jclass c, jlong h,
        jint m, jint minStore, jdouble maxShift, jdouble accept, jlongArray stats) {
    mhapb_ctx *ctx = (mhapb_ctx *)(intptr_t)h;
    mhapb_search_params sp = { m, minStore, maxShift, accept, 0, 0, 0, -1 };
    mhapb_hit *hits = NULL; uint64_t n = 0; mhapb_stats st;
    if (mhapb_search_self(ctx, &sp, &hits, &n, &st)) { throw_mhap(env, ctx); return NULL; }
    jbyteArray out = (*env)->NewByteArray(env, (jsize)(n * sizeof(mhapb_hit)));
    (*env)->SetByteArrayRegion(env, out, 0, (jsize)(n * sizeof(mhapb_hit)), (const jbyte *)hits);
    mhapb_free(hits);
    jlong s[5] = { st.elements_processed, st.sequences_hit, st.fully_compared, st.matches_processed, st.sequences_searched };
    (*env)->SetLongArrayRegion(env, stats, 0, 5, s);
    return out;
}
/* searchQueryReads, sketchToDat, storeSize, destroy: same pattern over
 * mhapb_search_query_reads, mhapb_sketch_to_dat, mhapb_store_size, mhapb_destroy. */

/* main/MhapMain.java:340-372: the -f k-mer filter.  text = the (decompressed) bytes of the filter file. */
JNIEXPORT jlong JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_filterLoadText(JNIEnv *env, jclass c, jlong h, jbyteArray text,
        jdouble filterCutoff, jdouble repeatWeight, jdouble idfScale, jint supressNoise, jboolean noTf, jboolean canonical) {
    mhapb_ctx *ctx = (mhapb_ctx *)(intptr_t)h;
    mhapb_filter_params p = { filterCutoff, repeatWeight, idfScale, supressNoise, noTf ? 1 : 0 };
    jsize n = (*env)->GetArrayLength(env, text);
    jbyte *b = (*env)->GetByteArrayElements(env, text, NULL);
    int64_t n_repeat = 0;
    int rc = mhapb_filter_load_text(ctx, &p, (const char *)b, (uint64_t)n, canonical ? 1 : 0, &n_repeat);
    (*env)->ReleaseByteArrayElements(env, text, b, JNI_ABORT);
    if (rc) { throw_mhap(env, ctx); return 0; }
    return n_repeat;
}

JNIEXPORT void JNICALL Java_edu_umd_marbl_mhap_impl_MhapB200_filterClear(JNIEnv *env, jclass c, jlong h) {
    mhapb_filter_clear((mhapb_ctx *)(intptr_t)h);
}
