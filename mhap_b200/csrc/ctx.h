// ctx.h -- the context object behind include/mhap_b200.h and the host-side helpers shared by api.cu (single-GPU
// entry points) and dist.cu (multi-GPU entry points: NCCL communicator, sharded search).
#pragma once
#include "../../include/mhap_b200.h"
#include "engine.h"

#include <cstdarg>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

struct ncclComm;   // NCCL is bound at run time (dist.cu); the context only carries the handle

namespace mhapb {


extern thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow keeping the first keep_bytes
    cudaError_t grow(size_t bytes, size_t keep_bytes, cudaStream_t st)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t want = std::max(bytes, cap * 2);
        void *np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&np, want); }
        if (e != cudaSuccess) return e;
        if (p && keep_bytes) { e = cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; e = cudaStreamSynchronize(st); }
        if (p) cudaFree(p);
        p = np; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Store {
    mhapb_sketch_params p{};
    bool configured = false;
    int64_t n = 0;
    int ord_stride = 0;
    DevBuf minhash, ord, ord_n, lenk, len, id;
    std::vector<int64_t> h_id; std::vector<uint8_t> h_fwd; std::vector<int32_t> h_len, h_lenk, h_ordn;
    // duplicate-id detection ("Sequence ID already exists in the hash table.", MinHashSearch.java:112-117).  FASTA ids arrive
    // in increasing order, so the common case is one compare per sketch; the hash set is only materialised when an id
    // arrives out of order (200 k set inserts cost 10+ ms of host time in front of every K1 launch).
    bool ids_monotonic = true; uint64_t last_key = 0; bool any_key = false;
    std::unordered_set<uint64_t> seen;
    // device list of the forward rows (the queries of a self search), rebuilt when the store changes
    DevBuf fwd_list; int64_t fwd_list_n = 0; bool fwd_list_valid = false;
    // index
    bool indexed = false;
    DevBuf slots, postings, present, idx_flag;
    bool index_safe = false;          // a build with the optimistic sub-table size overflowed once: use 2 slots per sketch from now on
    int log2capw = 0;
};


} // namespace mhapb

struct mhapb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // K1c runs here, concurrently with K1b of the same chunk
    std::mutex mu;
    std::string err;
    mhapb_timing timing{};
    // sketch scratch
    mhapb::DevBuf bases, desc, keys, wts, nlight, nheavy, dupcnt, gtable, ohash, counters;
    mhapb::DevBuf out_minhash, out_ord, out_ordn;
    // search scratch
    mhapb::DevBuf ovf_q;
    mhapb::PinnedBuf h_cand, h_ovl;               // pinned landing buffers of the surviving pairs
    std::vector<mhapb::StrandDesc> plan_all;      // the strands of the current sketch call (sketch_core)
    std::vector<uint32_t> plan_nk;                // their k-mer counts
    uint64_t super_cap = 0;                       // k-mers of key scratch one K1b launch may cover (sized once from the free memory)
    mhapb::PinnedBuf h_desc, h_vdesc;             // pinned plan of the strand descriptors (sketch_core)
    mhapb::DevBuf t512, vdesc;                    // step^512 tables and the virtual strands of sketches wider than 512 words
    uint64_t cand_cap_hint = 0; uint32_t ovf_threads_hint = 0;   // sizes the previous search needed
    mhapb::DevBuf qlist, cand, ovl, cand2, ovl2, ovf_list, fscratch, scounters, tmp_start, block_sums, q_minhash, q_ord, q_ordn, q_lenk, q_len, q_id, eq;
    mhapb::Store store;
    cudaEvent_t ev[12]{};                     // [8],[9]: K2a, read lazily (index_timing_pending)
    bool index_timing_pending = false;
    bool index_optimistic = false;            // MHAPB_INDEX_OPTIMISTIC at creation: start K2a with the small sub-table size (index_build)
    // the -f k-mer filter (FrequencyCounts); view.mode == 0 when none is set
    mhapb::DevBuf f_keys, f_idf, f_used, f_bloom;
    mhapb::KmerFilterView filter{};
    mhapb_filter_params filter_params{};
    bool filter_set = false;
    // multi-GPU (dist.cu): this context's rank in a communicator of nranks contexts (one per GPU)
    ncclComm *comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t comm_stream = nullptr;       // collectives run here, behind K2a / K2b on `stream`
    cudaEvent_t comm_ev[4]{};
    mhapb::DevBuf g_minhash, g_ord, g_ordn, g_lenk, g_len, g_id, g_pack, g_small;   // landing buffers of the all-gather
    mhapb::PinnedBuf g_h_id, g_h_len;         // host copies of the gathered id / length columns (for the hit records)
};

namespace mhapb {

int fail(mhapb_ctx *ctx, int code, const char *fmt, ...);

#define CU(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return mhapb::fail(ctx, e__ == cudaErrorMemoryAllocation ? MHAPB_ENOMEM : MHAPB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

// a set of query sketches on this GPU (device blocks + host columns) and the rows of it to search with
struct QuerySet {
    const int32_t *d_minhash = nullptr, *d_ord = nullptr, *d_ordn = nullptr, *d_lenk = nullptr, *d_len = nullptr; const int64_t *d_id = nullptr; int ord_stride = 0;
    const int64_t *h_id = nullptr; const uint8_t *h_fwd = nullptr /* NULL: all forward */; const int32_t *h_len = nullptr;
    std::vector<uint32_t> list;          // indices into the query arrays ...
    const uint32_t *d_list = nullptr; int64_t n_list = 0;   // ... or a list already on the device
    bool list_all = false; int64_t n_all = 0;   // ... or simply rows 0..n_all-1
    cudaEvent_t minhash_ready = nullptr; // multi-GPU: K2b waits for the gathered min-hash block,
    cudaEvent_t ord_ready = nullptr;     //            K2c for the gathered ordered sketches (dist.cu)
};

int check_sketch_params(mhapb_ctx *ctx, const mhapb_sketch_params *p);
int read_status(const mhapb_sketch_params &p, uint64_t len);
KmerFilterView filter_view(const mhapb_ctx *ctx);
bool filter_can_empty(const KmerFilterView &v);
void reset_sketch_timing(mhapb_ctx *ctx);
int h2d_bases(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads);
int sketch_core(mhapb_ctx *ctx, const mhapb_sketch_params &p, const uint8_t *d_bases, const uint64_t *h_offsets,
                uint32_t n_reads, int both, const std::vector<int64_t> &row_of_slot, int32_t *d_minhash,
                int32_t *d_ord, int ord_stride, int32_t *d_ord_n, std::vector<uint8_t> *slot_valid = nullptr,
                const char *h_bases = nullptr /* host copy of the reads: each chunk's characters are copied to d_bases (= ctx->bases) on the
                                                 second stream right before its K1a, so the H2D of chunk i+1 runs under K1 of chunk i */,
                const std::function<int()> *on_enqueued = nullptr /* run once everything is enqueued, before the final synchronisation;
                                                 a non-zero result is returned after the GPU work has completed */);
int index_build(mhapb_ctx *ctx);
int search_core(mhapb_ctx *ctx, const mhapb_search_params *sp, const QuerySet &q, int to_self,
                mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats);
// forward-only sketches of the valid reads of a query batch, compacted into ctx->q_minhash / q_ord / q_ordn; the host
// columns of the kept reads are appended to id / len / lenk (AbstractMatchSearch.java:225 dequeue(true))
int sketch_query_reads(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                       std::vector<int64_t> *id, std::vector<int32_t> *len, std::vector<int32_t> *lenk,
                       const uint8_t *d_resident /* non-NULL: the reads are already in HBM there, no H2D */);
void comm_release(mhapb_ctx *ctx);   // dist.cu

} // namespace mhapb
