/*
 * jni_harness.c -- drives every native of integration/mhapb_jni.c through a FAKE JNIEnv (tests/cpp/jni_stub/jni.h) on the
 * GPU and checks the results against direct calls of the C ABI on the same data.  Test infrastructure: there is no JVM in
 * this image, so this is how the shim's marshalling (arrays, direct buffers, exceptions, the packed hit records, the
 * framed .dat records) is exercised end to end.  Built and run by tests/test_gpu_jni_shim.py.
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mhap_b200.h"

/* ---- the natives under test (integration/mhapb_jni.c) ---- */
#define N(name) Java_edu_umd_marbl_mhap_impl_MhapB200_##name
jlong N(create)(JNIEnv *, jclass, jintArray);
void N(destroy)(JNIEnv *, jclass, jlong);
jobject N(hostAlloc)(JNIEnv *, jclass, jlong);
void N(hostFree)(JNIEnv *, jclass, jobject);
void N(storeReset)(JNIEnv *, jclass, jlong, jint, jint, jint, jint, jboolean, jint);
jlong N(storeAddReads)(JNIEnv *, jclass, jlong, jobject, jlongArray, jlongArray, jint, jboolean);
jlong N(storeAddDat)(JNIEnv *, jclass, jlong, jbyteArray, jlong);
jlong N(storeSize)(JNIEnv *, jclass, jlong);
void N(storeIds)(JNIEnv *, jclass, jlong, jlongArray, jbyteArray);
jbyteArray N(storeGetDat)(JNIEnv *, jclass, jlong, jlong);
jbyteArray N(searchSelf)(JNIEnv *, jclass, jlong, jint, jint, jdouble, jdouble, jlongArray);
jbyteArray N(searchSelfRange)(JNIEnv *, jclass, jlong, jint, jint, jdouble, jdouble, jlong, jlong, jlongArray);
jbyteArray N(searchQueryReads)(JNIEnv *, jclass, jlong, jint, jint, jdouble, jdouble, jobject, jlongArray, jlongArray, jint, jlongArray);
jbyteArray N(searchQueryDat)(JNIEnv *, jclass, jlong, jint, jint, jdouble, jdouble, jbyteArray, jlong, jlongArray);
jbyteArray N(sketchToDat)(JNIEnv *, jclass, jlong, jint, jint, jint, jint, jboolean, jint, jobject, jlongArray, jlongArray, jint, jboolean);
jlong N(filterLoadText)(JNIEnv *, jclass, jlong, jbyteArray, jdouble, jdouble, jdouble, jint, jboolean, jboolean);
void N(filterClear)(JNIEnv *, jclass, jlong);

/* ---- fake JVM objects ---- */
typedef struct { int kind; jsize len; void *data; jlong cap; } fobj;   /* kind: 1 byte[], 2 int[], 3 long[], 4 direct buffer, 5 class */
static char g_exc[512];
static int g_exc_set = 0;

static fobj *mk(int kind, jsize len, size_t elt) { fobj *o = calloc(1, sizeof *o); o->kind = kind; o->len = len; o->data = calloc(len ? len : 1, elt); return o; }
static void rel(void *o) { if (o) { fobj *f = o; if (f->kind != 4 && f->kind != 5) free(f->data); free(f); } }

static jclass f_FindClass(JNIEnv *e, const char *name) { (void)e; fobj *o = calloc(1, sizeof *o); o->kind = 5; o->data = (void *)name; return (jclass)o; }
static jint f_ThrowNew(JNIEnv *e, jclass c, const char *msg) { (void)e; (void)c; snprintf(g_exc, sizeof g_exc, "%s", msg); g_exc_set = 1; return 0; }
static jboolean f_ExceptionCheck(JNIEnv *e) { (void)e; return (jboolean)g_exc_set; }
static jsize f_GetArrayLength(JNIEnv *e, jarray a) { (void)e; return ((fobj *)a)->len; }
static jbyteArray f_NewByteArray(JNIEnv *e, jsize len) { (void)e; return (jbyteArray)mk(1, len, 1); }
static jbyte *f_GetByteArrayElements(JNIEnv *e, jbyteArray a, jboolean *c) { (void)e; if (c) *c = 0; return ((fobj *)a)->data; }
static void f_ReleaseByteArrayElements(JNIEnv *e, jbyteArray a, jbyte *p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void f_SetByteArrayRegion(JNIEnv *e, jbyteArray a, jsize s, jsize l, const jbyte *b) { (void)e; memcpy((jbyte *)((fobj *)a)->data + s, b, (size_t)l); }
static jint *f_GetIntArrayElements(JNIEnv *e, jintArray a, jboolean *c) { (void)e; if (c) *c = 0; return ((fobj *)a)->data; }
static void f_ReleaseIntArrayElements(JNIEnv *e, jintArray a, jint *p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static jlong *f_GetLongArrayElements(JNIEnv *e, jlongArray a, jboolean *c) { (void)e; if (c) *c = 0; return ((fobj *)a)->data; }
static void f_ReleaseLongArrayElements(JNIEnv *e, jlongArray a, jlong *p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void f_SetLongArrayRegion(JNIEnv *e, jlongArray a, jsize s, jsize l, const jlong *b) { (void)e; memcpy((jlong *)((fobj *)a)->data + s, b, (size_t)l * 8); }
static jobject f_NewDirectByteBuffer(JNIEnv *e, void *addr, jlong cap) { (void)e; fobj *o = calloc(1, sizeof *o); o->kind = 4; o->data = addr; o->cap = cap; return (jobject)o; }
static void *f_GetDirectBufferAddress(JNIEnv *e, jobject b) { (void)e; return b ? ((fobj *)b)->data : NULL; }
static jlong f_GetDirectBufferCapacity(JNIEnv *e, jobject b) { (void)e; return b ? ((fobj *)b)->cap : -1; }

static const struct JNINativeInterface_ g_table = {
    NULL, f_FindClass, f_ThrowNew, f_ExceptionCheck, f_GetArrayLength, f_NewByteArray, f_GetByteArrayElements, f_ReleaseByteArrayElements,
    f_SetByteArrayRegion, f_GetIntArrayElements, f_ReleaseIntArrayElements, f_GetLongArrayElements, f_ReleaseLongArrayElements,
    f_SetLongArrayRegion, f_NewDirectByteBuffer, f_GetDirectBufferAddress, f_GetDirectBufferCapacity,
};
static JNIEnv g_env = &g_table;
#define ENV (&g_env)

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "jni_harness: CHECK failed at line %d: %s (exception: %s)\n", __LINE__, #cond, g_exc_set ? g_exc : "none"); exit(1); } } while (0)
#define NOEXC() CHECK(!g_exc_set)

/* ---- synthetic overlapping reads ---- */
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }

static void make_reads(char *genome, int glen, char *bases, jlong *off, int n, int L, uint64_t seed)
{
    rng_state = seed;
    for (int i = 0; i < n; i++) {
        const int st = (int)(rnd() % (uint32_t)(glen - L));
        for (int j = 0; j < L; j++) {
            char c = genome[st + j];
            if (rnd() % 100 < 4) c = "ACGT"[rnd() & 3];
            bases[(size_t)i * L + j] = c;
        }
        off[i] = (jlong)i * L;
    }
    off[n] = (jlong)n * L;
}

static int cmp_hit(const void *a, const void *b) { return memcmp(a, b, sizeof(mhapb_hit)); }

static fobj *long_array(const jlong *v, int n) { fobj *o = mk(3, n, 8); memcpy(o->data, v, (size_t)n * 8); return o; }

int main(void)
{
    enum { NR = 600, NQ = 150, L = 2000, GLEN = 150000, K = 16, NH = 128, OK = 12, OS = 400, MINOLAP = 116 };
    char *genome = malloc(GLEN);
    rng_state = 12345;
    for (int i = 0; i < GLEN; i++) genome[i] = "ACGT"[rnd() & 3];

    /* ---- through the shim ---- */
    fobj *devs = mk(2, 1, 4); ((jint *)devs->data)[0] = 0;
    const jlong h = N(create)(ENV, NULL, (jintArray)devs); NOEXC(); CHECK(h != 0);
    N(storeReset)(ENV, NULL, h, K, NH, OK, OS, 0, MINOLAP); NOEXC();
    jobject buf = N(hostAlloc)(ENV, NULL, (jlong)NR * L); NOEXC();
    CHECK(f_GetDirectBufferCapacity(ENV, buf) == (jlong)NR * L);
    jlong *off = malloc((NR + 1) * 8), *ids = malloc(NR * 8);
    make_reads(genome, GLEN, f_GetDirectBufferAddress(ENV, buf), off, NR, L, 777);
    for (int i = 0; i < NR; i++) ids[i] = i + 1;
    fobj *joff = long_array(off, NR + 1), *jids = long_array(ids, NR), *jstats = mk(3, 5, 8);
    CHECK(N(storeAddReads)(ENV, NULL, h, buf, (jlongArray)joff, (jlongArray)jids, NR, 1) == 2 * NR); NOEXC();
    CHECK(N(storeSize)(ENV, NULL, h) == 2 * NR);
    fobj *raw = (fobj *)N(searchSelf)(ENV, NULL, h, 3, 0, 0.2, 0.78, (jlongArray)jstats); NOEXC();
    const size_t n_hits = (size_t)raw->len / sizeof(mhapb_hit);
    CHECK(raw->len % (jsize)sizeof(mhapb_hit) == 0 && n_hits > 50);

    /* ---- the same through the C ABI directly ---- */
    mhapb_ctx *ctx = NULL;
    CHECK(mhapb_create(0, &ctx) == 0);
    mhapb_sketch_params p = { K, NH, OK, OS, 0, MINOLAP };
    mhapb_search_params sp = { 3, 0, 0.2, 0.78, 0, 0, 0, -1 };
    CHECK(mhapb_store_reset(ctx, &p) == 0);
    int64_t added = 0;
    CHECK(mhapb_store_add_reads(ctx, f_GetDirectBufferAddress(ENV, buf), (const uint64_t *)off, (const int64_t *)ids, NR, 1, &added) == 0 && added == 2 * NR);
    mhapb_hit *dh = NULL; uint64_t dn = 0; mhapb_stats dst;
    CHECK(mhapb_search_self(ctx, &sp, &dh, &dn, &dst) == 0);
    CHECK(dn == n_hits);
    qsort(dh, dn, sizeof *dh, cmp_hit); qsort(raw->data, n_hits, sizeof(mhapb_hit), cmp_hit);
    CHECK(memcmp(dh, raw->data, dn * sizeof *dh) == 0);
    const jlong *s = jstats->data;
    CHECK(s[0] == dst.elements_processed && s[1] == dst.sequences_hit && s[2] == dst.fully_compared && s[3] == dst.matches_processed && s[4] == dst.sequences_searched);
    CHECK(s[3] == (jlong)n_hits && s[4] == NR);

    /* ---- storeIds / storeGetDat: the record SequenceSketch.fromByteStream would read ---- */
    fobj *jall = mk(3, 2 * NR, 8), *jfwd = mk(1, 2 * NR, 1);
    N(storeIds)(ENV, NULL, h, (jlongArray)jall, (jbyteArray)jfwd); NOEXC();
    CHECK(((jlong *)jall->data)[0] == 1 && ((jbyte *)jfwd->data)[0] == 1 && ((jlong *)jall->data)[1] == 1 && ((jbyte *)jfwd->data)[1] == 0);
    CHECK(((jlong *)jall->data)[2 * NR - 1] == NR);
    fobj *rec = (fobj *)N(storeGetDat)(ENV, NULL, h, 7); NOEXC();
    {
        uint32_t n = 0; int32_t H = 0, mo = 0, okk = 0;
        CHECK(mhapb_dat_decode(rec->data, (uint64_t)rec->len, 0, &n, &H, &mo, &okk, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == 0);
        CHECK(n == 1 && H == NH && okk == OK && mo == OS);
        int64_t id = 0; uint8_t fwd = 0; int32_t len = 0, lenk = 0, on = 0; int32_t mh[NH], *od = malloc((size_t)mo * 8);
        CHECK(mhapb_dat_decode(rec->data, (uint64_t)rec->len, 0, &n, &H, &mo, &okk, &id, &fwd, &len, &lenk, mh, od, &on) == 0);
        int64_t id2 = 0; int32_t fwd2 = 0, len2 = 0, lenk2 = 0, on2 = 0; int32_t mh2[NH], *od2 = malloc((size_t)OS * 8);
        CHECK(mhapb_store_get(ctx, 7, &id2, &fwd2, &len2, &lenk2, mh2, od2, &on2) == 0);
        CHECK(id == id2 && fwd == fwd2 && len == len2 && lenk == lenk2 && on == on2 && id == 4 && fwd == 0);
        CHECK(memcmp(mh, mh2, sizeof mh) == 0 && memcmp(od, od2, (size_t)on * 8) == 0);
        free(od); free(od2);
    }

    /* ---- sketchToDat -> storeAddDat into a second handle: the .dat route gives the same overlaps ---- */
    fobj *dat = (fobj *)N(sketchToDat)(ENV, NULL, h, K, NH, OK, OS, 0, MINOLAP, buf, (jlongArray)joff, (jlongArray)jids, NR, 1); NOEXC();
    const jlong h2 = N(create)(ENV, NULL, (jintArray)devs); NOEXC();
    N(storeReset)(ENV, NULL, h2, K, NH, OK, OS, 0, MINOLAP); NOEXC();
    CHECK(N(storeAddDat)(ENV, NULL, h2, (jbyteArray)dat, 0) == 2 * NR); NOEXC();
    fobj *raw2 = (fobj *)N(searchSelf)(ENV, NULL, h2, 3, 0, 0.2, 0.78, (jlongArray)jstats); NOEXC();
    CHECK(raw2->len == raw->len);
    qsort(raw2->data, n_hits, sizeof(mhapb_hit), cmp_hit);
    CHECK(memcmp(raw2->data, raw->data, (size_t)raw->len) == 0);

    /* ---- query mode: FASTA reads sketched on the GPU == the same queries handed over as .dat records ---- */
    jobject qbuf = N(hostAlloc)(ENV, NULL, (jlong)NQ * L); NOEXC();
    jlong *qoff = malloc((NQ + 1) * 8), *qids = malloc(NQ * 8);
    make_reads(genome, GLEN, f_GetDirectBufferAddress(ENV, qbuf), qoff, NQ, L, 4242);
    for (int i = 0; i < NQ; i++) qids[i] = NR + i + 1;
    fobj *jqoff = long_array(qoff, NQ + 1), *jqids = long_array(qids, NQ), *jqstats = mk(3, 5, 8), *jqstats2 = mk(3, 5, 8);
    fobj *q1 = (fobj *)N(searchQueryReads)(ENV, NULL, h, 3, 0, 0.2, 0.78, qbuf, (jlongArray)jqoff, (jlongArray)jqids, NQ, (jlongArray)jqstats); NOEXC();
    fobj *qdat = (fobj *)N(sketchToDat)(ENV, NULL, h, K, NH, OK, OS, 0, MINOLAP, qbuf, (jlongArray)jqoff, (jlongArray)jqids, NQ, 0); NOEXC();
    fobj *q2 = (fobj *)N(searchQueryDat)(ENV, NULL, h, 3, 0, 0.2, 0.78, (jbyteArray)qdat, 0, (jlongArray)jqstats2); NOEXC();
    CHECK(q1->len == q2->len && q1->len > 20 * (jsize)sizeof(mhapb_hit));
    qsort(q1->data, (size_t)q1->len / sizeof(mhapb_hit), sizeof(mhapb_hit), cmp_hit);
    qsort(q2->data, (size_t)q2->len / sizeof(mhapb_hit), sizeof(mhapb_hit), cmp_hit);
    CHECK(memcmp(q1->data, q2->data, (size_t)q1->len) == 0);
    CHECK(memcmp(jqstats->data, jqstats2->data, 40) == 0 && ((jlong *)jqstats->data)[4] == NQ);

    /* ---- per-sequence self search = the slice of the full self search whose query is that sequence ---- */
    fobj *jst3 = mk(3, 5, 8);
    fobj *r1 = (fobj *)N(searchSelfRange)(ENV, NULL, h, 3, 0, 0.2, 0.78, 2 * 300, 2, (jlongArray)jst3); NOEXC();   /* stored rows 600,601 = read id 301 */
    size_t expect = 0;
    for (size_t i = 0; i < n_hits; i++) expect += ((mhapb_hit *)raw->data)[i].from_id == 301;
    CHECK((size_t)r1->len / sizeof(mhapb_hit) == expect && ((jlong *)jst3->data)[4] == 1);

    /* ---- the reference's own error text crosses the boundary as the exception message ---- */
    CHECK(N(storeAddReads)(ENV, NULL, h, buf, (jlongArray)joff, (jlongArray)jids, NR, 1) == 0);
    CHECK(g_exc_set && strcmp(g_exc, "Sequence ID already exists in the hash table.") == 0);   /* impl/MinHashSearch.java:112-117 */
    g_exc_set = 0;
    CHECK(N(storeSize)(ENV, NULL, h) == 2 * NR);                                                /* the failed add left nothing behind */

    /* ---- -f filter: parsed by the library, installed on every device ---- */
    const char *ftxt = "4 2\nACGTACGTACGTACGT 0.5\nTTTTTTTTTTTTTTTT 0.25\nGGGGGGGGGGGGGGGG 0.000001\n";
    fobj *jtxt = mk(1, (jsize)strlen(ftxt), 1); memcpy(jtxt->data, ftxt, strlen(ftxt));
    CHECK(N(filterLoadText)(ENV, NULL, h, (jbyteArray)jtxt, 1.0e-5, 0.9, 3.0, 0, 0, 1) == 2); NOEXC();
    N(filterClear)(ENV, NULL, h); NOEXC();

    const size_t n_qhits = (size_t)q1->len / sizeof(mhapb_hit);
    N(hostFree)(ENV, NULL, buf); N(hostFree)(ENV, NULL, qbuf);
    N(destroy)(ENV, NULL, h); N(destroy)(ENV, NULL, h2);
    mhapb_free(dh); mhapb_destroy(ctx);
    rel(raw); rel(raw2); rel(dat); rel(rec); rel(q1); rel(q2); rel(qdat); rel(r1);
    printf("JNI_HARNESS_OK hits=%zu query_hits=%zu\n", n_hits, n_qhits);
    return 0;
}
