// dist.cu -- the multi-GPU half of the C ABI: one context per GPU, the contexts of a job joined by an NCCL
// communicator that lives INSIDE the library (no torch, no Python in the data plane).
//
// The reference is one JVM ("Distributed communication backend: none"); its users partition the reads by
// hand (docs/source/quickstart.rst:23).  Here the path shards with ONE exchange step:
//   1. reads are partitioned over the ranks; K1 runs per shard with no communication and every rank stores and
//      indexes ONLY its own shard (mhapb_store_add_reads on its context);
//   2. the forward sketches of every shard -- the queries of MinHashSearch.findMatches, AbstractMatchSearch.java:128-129 --
//      are all-gathered: small columns + min-hashes first, the ordered sketches behind them, on a separate
//      high-priority stream, so that K2a (index build over the rank's own min-hashes) hides the first part and
//      K2b (probe, needs only min-hashes) the second;
//   3. every rank queries its local index with the forward sketches of ALL ranks under the self-search id rules
//      (MinHashSearch.java:200,215-225): each overlap (query, target) is found exactly once, on the rank that owns
//      the target; the order-independent counters (MhapMain.java:572-590) are additive over target shards and are
//      summed with an all-reduce.
// Store-vs-query mode (AbstractMatchSearch.findMatches(streamer), :203-285) shards the same way: every rank sketches
// its shard of the query file forward-only, the query blocks are all-gathered, every rank runs all queries against
// its local index.
//
// NCCL is bound at run time (dlopen): a single-GPU user needs no NCCL installed, and inside a process that already
// loaded a libnccl.so.2 (PyTorch ships one) the same copy is used.
#include "ctx.h"

#include <nccl.h>     // types and enums only; every symbol is resolved with dlsym below
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace mhapb;

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // NCCL writes its banner / debug lines to stdout unless told otherwise; stdout is the overlap stream of the
        // command-line driver (MatchResult lines), so they go to stderr unless the user chose a file
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        const char *names[] = {getenv("MHAPB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) { api.err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        bool ok = true;
        auto sym = [&](const char *name) { void *p = dlsym(api.h, name); if (!p) { ok = false; api.err = std::string("NCCL symbol missing: ") + name; } return p; };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.h); api.h = nullptr; }
    });
    return &api;
}

// NCCL prints its version banner with printf when the first communicator is created (NCCL_DEBUG=VERSION/WARN/INFO); stdout
// is the overlap stream of the command-line driver, so fd 1 points at stderr for the duration of the initialisation
struct StdoutToStderr {
    int saved;
    StdoutToStderr() { fflush(stdout); saved = dup(1); if (saved >= 0) dup2(2, 1); }
    ~StdoutToStderr() { if (saved >= 0) { fflush(stdout); dup2(saved, 1); close(saved); } }
};

#define NC(ctx, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return fail(ctx, MHAPB_ECOMM, "%s: %s (%s:%d)", #call, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); } while (0)

static_assert(sizeof(ncclUniqueId) == MHAPB_COMM_ID_BYTES, "mhapb_comm_unique_id carries an ncclUniqueId");

// rows[i] of a [n][row_u4 * 16 bytes] block -> dst row i (rows == NULL: identity)
__global__ void k_gather_rows(const uint4 *__restrict__ src, uint4 *__restrict__ dst, const uint32_t *__restrict__ rows, int64_t n, int row_u4)
{
    const int64_t total = n * row_u4;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / row_u4; const int c = (int)(e % row_u4);
        const int64_t sr = rows ? rows[r] : r;
        dst[e] = src[sr * row_u4 + c];
    }
}
__global__ void k_gather_i32(const int32_t *__restrict__ src, int32_t *__restrict__ dst, const uint32_t *__restrict__ rows, int64_t n)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) dst[r] = src[rows ? rows[r] : r];
}

// the collective entry points also serve a job of ONE rank (no communicator): same path, no collectives
int comm_ready(mhapb_ctx *ctx)
{
    if (!ctx->comm_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = greatest priority: collectives get SMs as soon as a CTA of K2a / K2b retires
        CU(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
        for (auto &e : ctx->comm_ev) CU(ctx, cudaEventCreate(&e));
    }
    return MHAPB_OK;
}

// in-place all-gather of per-rank segments of `unit`-byte rows: rank r's rows live at buf + off[r]*unit
int gather_segments(mhapb_ctx *ctx, void *buf, const std::vector<int64_t> &cnt, const std::vector<int64_t> &off, size_t unit, bool equal)
{
    NcclApi *N = nccl_api();
    char *b = static_cast<char *>(buf);
    if (equal) {
        NC(ctx, N->AllGather(b + (size_t)off[ctx->rank] * unit, b, (size_t)cnt[ctx->rank] * unit, ncclChar, ctx->comm, ctx->comm_stream));
    } else {   // ragged shards: one broadcast per rank, fused by the group
        for (int r = 0; r < ctx->nranks; r++)
            if (cnt[r]) NC(ctx, N->Broadcast(b + (size_t)off[r] * unit, b + (size_t)off[r] * unit, (size_t)cnt[r] * unit, ncclChar, r, ctx->comm, ctx->comm_stream));
    }
    return MHAPB_OK;
}

// this rank's queries before the exchange
struct LocalQueries {
    int64_t n = 0;
    const int32_t *d_minhash = nullptr, *d_ord = nullptr, *d_ordn = nullptr;   // device blocks, rows picked by d_rows
    const uint32_t *d_rows = nullptr;                                           // NULL: rows 0..n-1
    int ord_stride = 0;
    std::vector<int64_t> id; std::vector<int32_t> len, lenk;                    // host columns, n entries
};

// the collective search: all-gather the queries of every rank, run them all against the local index
int dist_search(mhapb_ctx *ctx, const mhapb_search_params *sp, const LocalQueries &lq, int to_self,
                mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    NcclApi *N = nccl_api();
    Store &s = ctx->store;
    const int R = ctx->nranks, me = ctx->rank;
    const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.ord_stride;
    if ((size_t)lq.ord_stride != S && lq.n) return fail(ctx, MHAPB_EINVAL, "query ordered-sketch stride %d differs from the store's %d", lq.ord_stride, (int)S);
    if ((H * 4) % 16 || (S * 8) % 16) return fail(ctx, MHAPB_EINVAL, "multi-GPU search needs num_hashes a multiple of 4 and an even ordered sketch size");

    // 1. shard sizes (one 8-byte all-gather; the only host round trip before the search)
    CU(ctx, ctx->g_small.ensure((size_t)R * 8 + 64));
    std::vector<int64_t> cnt(R, 0), off(R + 1, 0);
    if (!ctx->comm) cnt[0] = lq.n;      // no communicator: a job of one rank, same code path without the collectives
    else {
        int64_t mine = lq.n;
        int64_t *d = ctx->g_small.as<int64_t>();
        CU(ctx, cudaMemcpyAsync(d + me, &mine, 8, cudaMemcpyHostToDevice, ctx->comm_stream));
        NC(ctx, N->AllGather(d + me, d, 1, ncclInt64, ctx->comm, ctx->comm_stream));
        CU(ctx, cudaMemcpyAsync(cnt.data(), d, (size_t)R * 8, cudaMemcpyDeviceToHost, ctx->comm_stream));
        CU(ctx, cudaStreamSynchronize(ctx->comm_stream));
    }
    bool equal = true;
    for (int r = 0; r < R; r++) { off[r + 1] = off[r] + cnt[r]; equal = equal && cnt[r] == cnt[0]; }
    const int64_t Q = off[R];
    if (Q >= 0x7fffffff) return fail(ctx, MHAPB_EINVAL, "%lld queries in the job exceed 32-bit query indices", (long long)Q);
    if (s.n == 0 && Q == 0) { if (stats) *stats = mhapb_stats{}; if (n_out) *n_out = 0; if (out) *out = (mhapb_hit *)malloc(sizeof(mhapb_hit)); return MHAPB_OK; }

    // 2. landing buffers; this rank's segment is packed in place (forward rows of the store, or the query block)
    const size_t q = (size_t)std::max<int64_t>(Q, 1);
    CU(ctx, ctx->g_minhash.ensure(q * H * 4)); CU(ctx, ctx->g_ord.ensure(q * S * 8)); CU(ctx, ctx->g_ordn.ensure(q * 4));
    CU(ctx, ctx->g_lenk.ensure(q * 4)); CU(ctx, ctx->g_len.ensure(q * 4)); CU(ctx, ctx->g_id.ensure(q * 8));
    CU(ctx, ctx->g_h_id.ensure(q * 8)); CU(ctx, ctx->g_h_len.ensure(q * 4));
    int launches = 0;
    const int64_t o = off[me];
    if (lq.n) {
        const int grid = 148 * 8;
        k_gather_rows<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(lq.d_minhash), reinterpret_cast<uint4 *>(ctx->g_minhash.as<int32_t>() + (size_t)o * H), lq.d_rows, lq.n, (int)(H * 4 / 16));
        k_gather_rows<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(lq.d_ord), reinterpret_cast<uint4 *>(ctx->g_ord.as<int32_t>() + (size_t)o * S * 2), lq.d_rows, lq.n, (int)(S * 8 / 16));
        k_gather_i32<<<(unsigned)std::min<int64_t>((lq.n + 255) / 256, grid), 256, 0, ctx->stream>>>(lq.d_ordn, ctx->g_ordn.as<int32_t>() + o, lq.d_rows, lq.n);
        launches += 3;
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(ctx->g_id.as<int64_t>() + o, lq.id.data(), (size_t)lq.n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->g_len.as<int32_t>() + o, lq.len.data(), (size_t)lq.n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->g_lenk.as<int32_t>() + o, lq.lenk.data(), (size_t)lq.n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx, cudaEventRecord(ctx->comm_ev[2], ctx->stream));

    // 3. the exchange, on the communication stream: columns + min-hashes, then the ordered sketches
    CU(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[2], 0));
    CU(ctx, cudaEventRecord(ctx->comm_ev[3], ctx->comm_stream));
    int rc = MHAPB_OK;
    if (ctx->comm) {
        NC(ctx, N->GroupStart());
        if (!rc) rc = gather_segments(ctx, ctx->g_id.p, cnt, off, 8, equal);
        if (!rc) rc = gather_segments(ctx, ctx->g_len.p, cnt, off, 4, equal);
        if (!rc) rc = gather_segments(ctx, ctx->g_lenk.p, cnt, off, 4, equal);
        if (!rc) rc = gather_segments(ctx, ctx->g_ordn.p, cnt, off, 4, equal);
        if (!rc) rc = gather_segments(ctx, ctx->g_minhash.p, cnt, off, H * 4, equal);
        NC(ctx, N->GroupEnd());
        if (rc) return rc;
    }
    CU(ctx, cudaEventRecord(ctx->comm_ev[0], ctx->comm_stream));
    if (ctx->comm) {
        NC(ctx, N->GroupStart());
        rc = gather_segments(ctx, ctx->g_ord.p, cnt, off, S * 8, equal);
        NC(ctx, N->GroupEnd());
        if (rc) return rc;
    }
    CU(ctx, cudaEventRecord(ctx->comm_ev[1], ctx->comm_stream));

    // 4. K2a over the rank's own min-hashes runs behind the exchange; K2b waits for the min-hashes, K2c for the ordered sketches
    if (s.n > 0) { rc = index_build(ctx); if (rc) return rc; }
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[0], 0));
    CU(ctx, cudaMemcpyAsync(ctx->g_h_id.p, ctx->g_id.p, (size_t)Q * 8, cudaMemcpyDeviceToHost, ctx->stream));     // host copies for the hit records
    CU(ctx, cudaMemcpyAsync(ctx->g_h_len.p, ctx->g_len.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    mhapb_stats st{};
    mhapb_hit *hits = nullptr; uint64_t nh = 0;
    if (s.n > 0 && Q > 0) {
        QuerySet qs;
        qs.d_minhash = ctx->g_minhash.as<int32_t>(); qs.d_ord = ctx->g_ord.as<int32_t>(); qs.d_ordn = ctx->g_ordn.as<int32_t>();
        qs.d_lenk = ctx->g_lenk.as<int32_t>(); qs.d_len = ctx->g_len.as<int32_t>(); qs.d_id = ctx->g_id.as<int64_t>(); qs.ord_stride = (int)S;
        qs.h_id = ctx->g_h_id.as<int64_t>(); qs.h_fwd = nullptr; qs.h_len = ctx->g_h_len.as<int32_t>();
        qs.list_all = true; qs.n_all = Q;
        qs.ord_ready = ctx->comm_ev[1];
        rc = search_core(ctx, sp, qs, to_self, &hits, &nh, &st);
        if (rc) return rc;
    } else {
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[1], 0));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    {
        float ms = 0;
        if (cudaEventSynchronize(ctx->comm_ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->comm_ev[3], ctx->comm_ev[1]) == cudaSuccess) ctx->timing.gather_ms = ms;
    }
    ctx->timing.kernel_launches += launches;

    // 5. job-wide counters: additive over target shards; every rank searched every query once
    st.sequences_searched = Q;
    if (ctx->comm) {
        int64_t v[4] = {st.elements_processed, st.sequences_hit, st.fully_compared, st.matches_processed};
        int64_t *d = ctx->g_small.as<int64_t>();
        CU(ctx, cudaMemcpyAsync(d, v, sizeof v, cudaMemcpyHostToDevice, ctx->comm_stream));
        NC(ctx, N->AllReduce(d, d, 4, ncclInt64, ncclSum, ctx->comm, ctx->comm_stream));
        CU(ctx, cudaMemcpyAsync(v, d, sizeof v, cudaMemcpyDeviceToHost, ctx->comm_stream));
        CU(ctx, cudaStreamSynchronize(ctx->comm_stream));
        st.elements_processed = v[0]; st.sequences_hit = v[1]; st.fully_compared = v[2]; st.matches_processed = v[3];
    }
    if (stats) *stats = st;
    if (n_out) *n_out = nh;
    if (out) { if (!hits) hits = (mhapb_hit *)malloc(sizeof(mhapb_hit)); *out = hits; } else free(hits);
    return MHAPB_OK;
}

} // namespace

namespace mhapb {

void comm_release(mhapb_ctx *ctx)
{
    if (ctx->comm) { NcclApi *N = nccl_api(); if (N->h) N->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    for (auto &e : ctx->comm_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    ctx->g_h_id.release(); ctx->g_h_len.release();
    ctx->rank = 0; ctx->nranks = 1;
}

static int comm_attach(mhapb_ctx *ctx, ncclComm_t comm, int rank, int nranks)
{
    ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
    return comm_ready(ctx);
}

} // namespace mhapb

extern "C" {

int mhapb_comm_unique_id(uint8_t *id)
{
    if (!id) return MHAPB_EINVAL;
    NcclApi *N = nccl_api();
    if (!N->h) return fail(nullptr, MHAPB_ECOMM, "%s", N->err.c_str());
    ncclUniqueId u;
    if (N->GetUniqueId(&u) != ncclSuccess) return fail(nullptr, MHAPB_ECOMM, "ncclGetUniqueId failed");
    memcpy(id, &u, sizeof u);
    return MHAPB_OK;
}

int mhapb_comm_init_rank(mhapb_ctx *ctx, const uint8_t *id, int rank, int nranks)
{
    if (!ctx || !id) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (rank < 0 || nranks < 1 || rank >= nranks) return fail(ctx, MHAPB_EINVAL, "rank %d of %d", rank, nranks);
    if (ctx->comm) return fail(ctx, MHAPB_ESTATE, "context already belongs to a communicator");
    NcclApi *N = nccl_api();
    if (!N->h) return fail(ctx, MHAPB_ECOMM, "%s", N->err.c_str());
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    {
        StdoutToStderr guard;
        NC(ctx, N->CommInitRank(&comm, nranks, u, rank));
    }
    return comm_attach(ctx, comm, rank, nranks);
}

int mhapb_comm_init_all(mhapb_ctx **ctxs, int n)
{
    if (!ctxs || n < 1) return MHAPB_EINVAL;
    for (int i = 0; i < n; i++) if (!ctxs[i]) return MHAPB_EINVAL;
    mhapb_ctx *c0 = ctxs[0];
    for (int i = 0; i < n; i++) if (ctxs[i]->comm) return fail(c0, MHAPB_ESTATE, "context %d already belongs to a communicator", i);
    NcclApi *N = nccl_api();
    if (!N->h) return fail(c0, MHAPB_ECOMM, "%s", N->err.c_str());
    ncclUniqueId u;
    NC(c0, N->GetUniqueId(&u));
    std::vector<ncclComm_t> comms(n, nullptr);
    {
        StdoutToStderr guard;
        NC(c0, N->GroupStart());
        for (int i = 0; i < n; i++) {
            CU(c0, cudaSetDevice(ctxs[i]->device));
            NC(c0, N->CommInitRank(&comms[i], n, u, i));
        }
        NC(c0, N->GroupEnd());
    }
    for (int i = 0; i < n; i++) {
        CU(c0, cudaSetDevice(ctxs[i]->device));
        int rc = comm_attach(ctxs[i], comms[i], i, n);
        if (rc) return rc;
    }
    return MHAPB_OK;
}

int mhapb_comm_destroy(mhapb_ctx *ctx)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    comm_release(ctx);
    return MHAPB_OK;
}

int mhapb_comm_info(mhapb_ctx *ctx, int *rank, int *nranks, int *nccl_version)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->comm ? ctx->nranks : 1;
    if (nccl_version) { *nccl_version = 0; NcclApi *N = nccl_api(); if (N->h) N->GetVersion(nccl_version); }
    return MHAPB_OK;
}

int mhapb_dist_search_self(mhapb_ctx *ctx, const mhapb_search_params *sp, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = comm_ready(ctx);
    if (rc) return rc;
    Store &s = ctx->store;
    if (!s.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    LocalQueries lq;
    std::vector<uint32_t> rows;
    for (int64_t i = 0; i < s.n; i++)
        if (s.h_fwd[i]) { rows.push_back((uint32_t)i); lq.id.push_back(s.h_id[i]); lq.len.push_back(s.h_len[i]); lq.lenk.push_back(s.h_lenk[i]); }
    lq.n = (int64_t)rows.size();
    if (!s.fwd_list_valid) {
        CU(ctx, s.fwd_list.ensure(rows.size() * 4 + 4));
        CU(ctx, cudaMemcpyAsync(s.fwd_list.p, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        s.fwd_list_n = lq.n; s.fwd_list_valid = true;
    }
    lq.d_rows = s.fwd_list.as<uint32_t>();
    lq.d_minhash = s.minhash.as<int32_t>(); lq.d_ord = s.ord.as<int32_t>(); lq.d_ordn = s.ord_n.as<int32_t>(); lq.ord_stride = s.ord_stride;
    return dist_search(ctx, sp, lq, 1, out, n_out, stats);
}

static int dist_query_reads(mhapb_ctx *ctx, const mhapb_search_params *sp, const char *bases, const uint8_t *d_bases, const uint64_t *offsets,
                            const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = comm_ready(ctx);
    if (rc) return rc;
    Store &s = ctx->store;
    if (!s.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    if (!offsets || (!bases && !d_bases && n_reads && offsets[n_reads] > offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    reset_sketch_timing(ctx);
    LocalQueries lq;
    rc = sketch_query_reads(ctx, bases, offsets, ids, n_reads, &lq.id, &lq.len, &lq.lenk, d_bases);   // forward only (AbstractMatchSearch.java:225)
    if (rc) return rc;
    lq.n = (int64_t)lq.id.size();
    lq.d_minhash = ctx->q_minhash.as<int32_t>(); lq.d_ord = ctx->q_ord.as<int32_t>(); lq.d_ordn = ctx->q_ordn.as<int32_t>();
    lq.ord_stride = s.p.ordered_sketch_size;
    if (lq.ord_stride != s.ord_stride && lq.n) return fail(ctx, MHAPB_EINVAL, "store stride %d differs from --ordered-sketch-size %d", s.ord_stride, lq.ord_stride);
    return dist_search(ctx, sp, lq, 0, out, n_out, stats);
}

int mhapb_dist_search_query_reads(mhapb_ctx *ctx, const mhapb_search_params *sp, const char *bases, const uint64_t *offsets,
                                  const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    return dist_query_reads(ctx, sp, bases, nullptr, offsets, ids, n_reads, out, n_out, stats);
}

int mhapb_dist_search_query_reads_device(mhapb_ctx *ctx, const mhapb_search_params *sp, const void *d_bases, const uint64_t *h_offsets,
                                         const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    return dist_query_reads(ctx, sp, nullptr, (const uint8_t *)d_bases, h_offsets, ids, n_reads, out, n_out, stats);
}

} // extern "C"

// ---- one process, several GPUs --------------------------------------------------------------------------------
struct mhapb_multi {
    std::vector<mhapb_ctx *> ctx;
    std::string err;
    mhapb_sketch_params p{};
    int next_dev = 0;        // round-robin start for batches smaller than the device count
};

static int multi_fail(mhapb_multi *m, int code, const std::string &msg) { if (m) m->err = msg; return code; }

// run f(i) on one host thread per device; first failure wins
template <class F>
static int multi_run(mhapb_multi *m, F f)
{
    const int n = (int)m->ctx.size();
    std::vector<int> rc(n, 0);
    if (n == 1) rc[0] = f(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++) th.emplace_back([&, i] { rc[i] = f(i); });
        for (auto &t : th) t.join();
    }
    // a single device reports the library's (= the reference's) message verbatim; several prefix the device
    for (int i = 0; i < n; i++) if (rc[i]) return multi_fail(m, rc[i], n == 1 ? std::string(mhapb_last_error(m->ctx[i])) : std::string("device ") + std::to_string(m->ctx[i]->device) + ": " + mhapb_last_error(m->ctx[i]));
    return MHAPB_OK;
}

// contiguous parts of a batch, balanced by bases (K1 cost is linear in bases)
static std::vector<uint32_t> split_reads(const uint64_t *offsets, uint32_t n_reads, int parts)
{
    std::vector<uint32_t> cut(parts + 1, n_reads);
    cut[0] = 0;
    const uint64_t b0 = offsets[0], total = offsets[n_reads] - b0;
    uint32_t r = 0;
    for (int i = 1; i < parts; i++) {
        const uint64_t want = b0 + total * (uint64_t)i / (uint64_t)parts;
        while (r < n_reads && offsets[r] < want) r++;
        cut[i] = r;
    }
    return cut;
}

extern "C" {

int mhapb_multi_create(const int *device_ids, int n_devices, mhapb_multi **out)
{
    if (!out) return MHAPB_EINVAL;
    *out = nullptr;
    if (!device_ids || n_devices < 1) return fail(nullptr, MHAPB_EINVAL, "need at least one device id");
    for (int i = 0; i < n_devices; i++) for (int j = 0; j < i; j++) if (device_ids[i] == device_ids[j]) return fail(nullptr, MHAPB_EINVAL, "device %d listed twice", device_ids[i]);
    mhapb_multi *m = new mhapb_multi();
    for (int i = 0; i < n_devices; i++) {
        mhapb_ctx *c = nullptr;
        int rc = mhapb_create(device_ids[i], &c);
        if (rc) { for (auto x : m->ctx) mhapb_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
    if (n_devices > 1) {
        int rc = mhapb_comm_init_all(m->ctx.data(), n_devices);
        if (rc) { fail(nullptr, rc, "%s", mhapb_last_error(m->ctx[0])); for (auto x : m->ctx) mhapb_destroy(x); delete m; return rc; }
    }
    *out = m;
    return MHAPB_OK;
}

void mhapb_multi_destroy(mhapb_multi *m)
{
    if (!m) return;
    for (auto c : m->ctx) mhapb_destroy(c);
    delete m;
}

const char *mhapb_multi_last_error(const mhapb_multi *m) { return m ? m->err.c_str() : mhapb_last_error(nullptr); }
int mhapb_multi_n_devices(const mhapb_multi *m) { return m ? (int)m->ctx.size() : 0; }
mhapb_ctx *mhapb_multi_ctx(mhapb_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int mhapb_multi_store_reset(mhapb_multi *m, const mhapb_sketch_params *p)
{
    if (!m || !p) return MHAPB_EINVAL;
    m->p = *p;
    return multi_run(m, [&](int i) { return mhapb_store_reset(m->ctx[i], p); });
}

int mhapb_multi_store_reserve(mhapb_multi *m, int64_t n_sketches)
{
    if (!m) return MHAPB_EINVAL;
    const int64_t per = n_sketches / (int64_t)m->ctx.size() + 64;
    return multi_run(m, [&](int i) { return mhapb_store_reserve(m->ctx[i], per); });
}

int mhapb_multi_store_add_reads(mhapb_multi *m, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                                int both_strands, int64_t *n_added)
{
    if (!m) return MHAPB_EINVAL;
    if (!offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return multi_fail(m, MHAPB_EINVAL, "null bases/offsets");
    const int n = (int)m->ctx.size();
    const std::vector<uint32_t> cut = split_reads(offsets, n_reads, n);
    std::vector<int64_t> added(n, 0);
    // a batch's parts go to the devices in rotation, so that small batches do not all land on device 0
    const int rot = m->next_dev; m->next_dev = (m->next_dev + 1) % n;
    int rc = multi_run(m, [&](int i) {
        const int part = (i + n - rot) % n;
        const uint32_t r0 = cut[part], r1 = cut[part + 1];
        if (r1 <= r0) return (int)MHAPB_OK;
        std::vector<uint64_t> off(r1 - r0 + 1);
        for (uint32_t r = r0; r <= r1; r++) off[r - r0] = offsets[r] - offsets[r0];
        std::vector<int64_t> idv;
        if (!ids) { idv.resize(r1 - r0); for (uint32_t r = r0; r < r1; r++) idv[r - r0] = (int64_t)r + 1; }
        return mhapb_store_add_reads(m->ctx[i], bases + offsets[r0], off.data(), ids ? ids + r0 : idv.data(), r1 - r0, both_strands, &added[i]);
    });
    if (n_added) { *n_added = 0; for (int64_t a : added) *n_added += a; }
    return rc;
}

int mhapb_multi_store_add_sketches(mhapb_multi *m, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len, const int32_t *seq_len_kmers,
                                   const int32_t *minhash, int32_t num_hashes, const int32_t *ord_hash_pos, const int32_t *ord_n,
                                   int32_t ord_stride, int32_t ordered_kmer_size, uint32_t n_sk)
{
    if (!m) return MHAPB_EINVAL;
    const int n = (int)m->ctx.size();
    return multi_run(m, [&](int i) {
        const uint32_t r0 = (uint32_t)((uint64_t)n_sk * i / n), r1 = (uint32_t)((uint64_t)n_sk * (i + 1) / n);
        if (r1 <= r0) return (int)MHAPB_OK;
        return mhapb_store_add_sketches(m->ctx[i], ids + r0, is_fwd + r0, seq_len + r0, seq_len_kmers + r0, minhash + (size_t)r0 * num_hashes, num_hashes,
                                        ord_hash_pos + (size_t)r0 * ord_stride * 2, ord_n + r0, ord_stride, ordered_kmer_size, r1 - r0);
    });
}

int64_t mhapb_multi_store_size(mhapb_multi *m)
{
    if (!m) return MHAPB_EINVAL;
    int64_t t = 0;
    for (auto c : m->ctx) t += mhapb_store_size(c);
    return t;
}

// concatenate per-device hit arrays; counters: job-wide values are already on every rank (dist) or are summed here (host-array queries)
static int multi_merge(mhapb_multi *m, std::vector<mhapb_hit *> &h, std::vector<uint64_t> &nh, std::vector<mhapb_stats> &st, bool stats_are_global,
                       mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    uint64_t total = 0;
    for (uint64_t x : nh) total += x;
    mhapb_hit *all = (mhapb_hit *)malloc(sizeof(mhapb_hit) * std::max<uint64_t>(1, total));
    if (!all) { for (auto p : h) free(p); return multi_fail(m, MHAPB_ENOMEM, "malloc hits"); }
    uint64_t w = 0;
    for (size_t i = 0; i < h.size(); i++) { if (nh[i]) memcpy(all + w, h[i], sizeof(mhapb_hit) * nh[i]); w += nh[i]; free(h[i]); }
    mhapb_stats s = st[0];
    if (!stats_are_global)
        for (size_t i = 1; i < st.size(); i++) {
            s.elements_processed += st[i].elements_processed; s.sequences_hit += st[i].sequences_hit;
            s.fully_compared += st[i].fully_compared; s.matches_processed += st[i].matches_processed;
        }
    if (stats) *stats = s;
    if (n_out) *n_out = total;
    if (out) *out = all; else free(all);
    return MHAPB_OK;
}

int mhapb_multi_search_self(mhapb_multi *m, const mhapb_search_params *sp, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!m || !sp) return MHAPB_EINVAL;
    const int n = (int)m->ctx.size();
    if (n == 1) { int rc = mhapb_search_self(m->ctx[0], sp, out, n_out, stats); if (rc) multi_fail(m, rc, mhapb_last_error(m->ctx[0])); return rc; }
    std::vector<mhapb_hit *> h(n, nullptr); std::vector<uint64_t> nh(n, 0); std::vector<mhapb_stats> st(n);
    int rc = multi_run(m, [&](int i) { return mhapb_dist_search_self(m->ctx[i], sp, &h[i], &nh[i], &st[i]); });
    if (rc) { for (auto p : h) free(p); return rc; }
    return multi_merge(m, h, nh, st, true, out, n_out, stats);
}

int mhapb_multi_search_query_reads(mhapb_multi *m, const mhapb_search_params *sp, const char *bases, const uint64_t *offsets, const int64_t *ids,
                                   uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!m || !sp) return MHAPB_EINVAL;
    if (!offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return multi_fail(m, MHAPB_EINVAL, "null bases/offsets");
    const int n = (int)m->ctx.size();
    if (n == 1) { int rc = mhapb_search_query_reads(m->ctx[0], sp, bases, offsets, ids, n_reads, out, n_out, stats); if (rc) multi_fail(m, rc, mhapb_last_error(m->ctx[0])); return rc; }
    const std::vector<uint32_t> cut = split_reads(offsets, n_reads, n);
    std::vector<mhapb_hit *> h(n, nullptr); std::vector<uint64_t> nh(n, 0); std::vector<mhapb_stats> st(n);
    int rc = multi_run(m, [&](int i) {
        const uint32_t r0 = cut[i], r1 = cut[i + 1];
        std::vector<uint64_t> off(r1 - r0 + 1, 0);
        for (uint32_t r = r0; r <= r1 && r1 > r0; r++) off[r - r0] = offsets[r] - offsets[r0];
        std::vector<int64_t> idv;
        if (!ids) { idv.resize(r1 - r0); for (uint32_t r = r0; r < r1; r++) idv[r - r0] = (int64_t)r + 1; }
        return mhapb_dist_search_query_reads(m->ctx[i], sp, bases + offsets[r0], off.data(), ids ? ids + r0 : idv.data(), r1 - r0, &h[i], &nh[i], &st[i]);
    });
    if (rc) { for (auto p : h) free(p); return rc; }
    return multi_merge(m, h, nh, st, true, out, n_out, stats);
}

int mhapb_multi_search_query_sketches(mhapb_multi *m, const mhapb_search_params *sp, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len,
                                      const int32_t *seq_len_kmers, const int32_t *minhash, int32_t num_hashes, const int32_t *ord_hash_pos,
                                      const int32_t *ord_n, int32_t ord_stride, int32_t ordered_kmer_size, uint32_t n_sk,
                                      mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!m || !sp) return MHAPB_EINVAL;
    const int n = (int)m->ctx.size();
    // the query sketches are host arrays: every device gets all of them (no exchange needed) and answers for its shard of the store
    std::vector<mhapb_hit *> h(n, nullptr); std::vector<uint64_t> nh(n, 0); std::vector<mhapb_stats> st(n);
    int rc = multi_run(m, [&](int i) {
        if (mhapb_store_size(m->ctx[i]) == 0) { st[i] = mhapb_stats{}; return (int)MHAPB_OK; }
        return mhapb_search_query_sketches(m->ctx[i], sp, ids, is_fwd, seq_len, seq_len_kmers, minhash, num_hashes, ord_hash_pos, ord_n, ord_stride,
                                           ordered_kmer_size, n_sk, &h[i], &nh[i], &st[i]);
    });
    if (rc) { for (auto p : h) free(p); return rc; }
    int64_t searched = 0;
    for (auto &x : st) searched = std::max(searched, x.sequences_searched);
    st[0].sequences_searched = searched;
    return multi_merge(m, h, nh, st, false, out, n_out, stats);
}

} // extern "C"
