// api.cu -- the C ABI of include/mhap_b200.h: context, device memory, batching, and the host-side
// halves of the reference interfaces (status per read, id filters' inputs, score, MatchResult).
// No CPU fallback: every compute entry point launches the kernels in sketch.cu / search.cu.
#include "ctx.h"
#include "hash.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

using namespace mhapb;

namespace mhapb { thread_local std::string g_create_error; }


namespace mhapb {

int fail(mhapb_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}


int check_sketch_params(mhapb_ctx *ctx, const mhapb_sketch_params *p)
{
    if (!p) return fail(ctx, MHAPB_EINVAL, "null sketch params");
    if (p->kmer_size < 1 || p->kmer_size > 4096) return fail(ctx, MHAPB_EINVAL, "kmer_size %d out of range", p->kmer_size);
    if (p->num_hashes < 1 || p->num_hashes > kMaxNumHashes) return fail(ctx, MHAPB_EINVAL, "num_hashes %d out of range 1..%d", p->num_hashes, kMaxNumHashes);
    if (p->ordered_kmer_size < 1 || p->ordered_kmer_size > 4096) return fail(ctx, MHAPB_EINVAL, "ordered_kmer_size %d out of range", p->ordered_kmer_size);
    if (p->ordered_sketch_size < 1 || p->ordered_sketch_size > kMaxOrderedSketch) return fail(ctx, MHAPB_EINVAL, "ordered_sketch_size %d out of range 1..%d", p->ordered_sketch_size, kMaxOrderedSketch);
    return MHAPB_OK;
}

// Per-read status from lengths alone (SequenceSketchStreamer.java:129-133, MinHashSketch.java:55-56,
// BottomOverlapSketch.java:530-531).
int read_status(const mhapb_sketch_params &p, uint64_t len)
{
    if ((int64_t)len < (int64_t)p.min_olap_length) return 2;
    if ((int64_t)len - p.kmer_size + 1 < 1) return 1;
    if ((int64_t)len - p.ordered_kmer_size + 1 < 1) return 1;
    return 0;
}

struct Elapsed { float hash = 0, minhash = 0, ordered = 0; };

// Sketch reads whose characters are at d_bases (device).  row_of_slot[slot] (slot = read*per+strand)
// gives the output row or -1 to skip.  Outputs are device arrays.
// The weight rule in force for a sketch call: the context's filter (if any) under the call's repeat-weight class.
// Returns the view by value; light_weight is the weight of a once-seen k-mer outside the repeat map.
KmerFilterView filter_view(const mhapb_ctx *ctx)
{
    KmerFilterView v = ctx->filter;
    if (!ctx->filter_set) { v = KmerFilterView{}; v.mode = 0; v.light_weight = 1; }
    return v;
}
// k-mers can be dropped (and a strand end up with none: ZeroNGramsFoundException, MinHashSketch.java:84,156)
bool filter_can_empty(const KmerFilterView &v) { return v.mode == 1 || v.remove_unique == 1; }
// dup counts are not needed when the weight ignores tf
int filter_unweighted(const KmerFilterView &v, int unweighted) { return v.mode == 0 ? unweighted : (v.mode == 1 || (v.mode == 2 && v.no_tf)) ? 1 : 0; }

// slot_valid (optional, size n_reads*per): set to 0 for strands whose every k-mer was filtered out.
int sketch_core(mhapb_ctx *ctx, const mhapb_sketch_params &p, const uint8_t *d_bases, const uint64_t *h_offsets,
                uint32_t n_reads, int both, const std::vector<int64_t> &row_of_slot, int32_t *d_minhash,
                int32_t *d_ord, int ord_stride, int32_t *d_ord_n, std::vector<uint8_t> *slot_valid, const char *h_bases,
                const std::function<int()> *on_enqueued)
{
    static const bool trace = getenv("MHAPB_TRACE") != nullptr;   // host-side time stamps of the enqueue path, to stderr
    const auto t_enter = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count(); };
    double t_plan = 0, t_first = 0, t_enq = 0;
    const KmerFilterView flt = filter_view(ctx);
    if (ctx->filter_set && (ctx->filter_params.repeat_weight < 0.0) != (p.unweighted != 0))
        return fail(ctx, MHAPB_EINVAL, "sketch params unweighted=%d contradict the filter's repeat_weight %g", p.unweighted, ctx->filter_params.repeat_weight);
    const bool want_valid = slot_valid && filter_can_empty(flt);
    std::vector<int32_t> h_nl, h_nh;
    const int per = both ? 2 : 1;
    const int k = p.kmer_size, ok = p.ordered_kmer_size, H = p.num_hashes, S = p.ordered_sketch_size;
    // The strands of the call, materialised LAZILY (chunk by chunk, see fill_until): the first K1a launch should not wait for
    // 2*10^5 descriptors to be written -- on a host whose cores are busy or throttled everything before the first launch is
    // GPU idle time.
    std::vector<StrandDesc> &all = ctx->plan_all;    // kept across calls: a fresh 6 MB vector per call is a page fault per 4 KB
    all.clear();
    all.reserve((size_t)n_reads * per);
    uint32_t gen_r = 0; int gen_s = 0;
    auto gen_next = [&](StrandDesc *out) -> bool {
        while (gen_r < n_reads) {
            const uint64_t len = h_offsets[gen_r + 1] - h_offsets[gen_r];
            if (!read_status(p, len)) {
                while (gen_s < per) {
                    const int st = gen_s++;
                    const int64_t row = row_of_slot[(size_t)gen_r * per + st];
                    if (row < 0) continue;
                    out->base_off = h_offsets[gen_r]; out->koff = 0; out->len = (uint32_t)len; out->row = (uint32_t)row; out->rc = (uint32_t)st;
                    out->slot = (uint32_t)((size_t)gen_r * per + st);
                    return true;
                }
            }
            gen_r++; gen_s = 0;
        }
        return false;
    };
    auto fill_until = [&](size_t need) { StrandDesc d; while (all.size() < need && gen_next(&d)) all.push_back(d); };
    // sizes only: k-mers of strand i of the call, in order (the same walk as gen_next, without writing descriptors)
    std::vector<uint32_t> &nk_of = ctx->plan_nk;
    nk_of.clear();
    nk_of.reserve((size_t)n_reads * per);
    for (uint32_t r = 0; r < n_reads; r++) {
        const uint64_t len = h_offsets[r + 1] - h_offsets[r];
        if (len > 0x7fffff00ull) return fail(ctx, MHAPB_EINVAL, "read %u longer than 2^31 bases", r);
        if (read_status(p, len)) continue;
        for (int st = 0; st < per; st++) if (row_of_slot[(size_t)r * per + st] >= 0) nk_of.push_back((uint32_t)(len - k + 1));
    }
    const size_t n_total = nk_of.size();
    ctx->timing.xorshift_steps = 0;
    ctx->timing.kmers_hashed = 0;
    if (!n_total) { if (on_enqueued) return (*on_enqueued)(); return MHAPB_OK; }

    // Work plan.  The strands are cut into CHUNKS of <= 256 M k-mers: a chunk is the unit of the host->device copy (its
    // characters travel on the copy stream while the previous chunk is hashed) and of the K1a / K1c launches.  K1b -- 76 % of
    // the step -- is launched ONCE over all chunks of a SUPER-CHUNK: its persistent warps take strands from a queue and a
    // strand is ~4 ms of warp time, so every launch ends with a tail in which the SMs run dry one by one; eight launches per
    // 100 k reads paid that tail eight times (profiles/r2l: 267 -> 25x ms).  The price is key scratch for the whole
    // super-chunk (12 bytes per k-mer: 24 GB for 100 k x 10 kbp reads, both strands), bounded by MHAPB_K1_SUPER_GB (default:
    // a third of the free memory, at most 64 GB).
    const uint64_t chunk_cap = 256ull << 20;
    const int max_chunk_strands = 1 << 20;
    uint64_t super_cap = ctx->super_cap;
    if (!super_cap) {   // once per context: cudaMemGetInfo is a driver round trip (0.1-1 ms, occasionally far more)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t have = ctx->keys.cap + ctx->wts.cap;
        double gb = std::min(64.0, (double)(free_b + have) / 3.0 / 1e9);
        if (const char *e = getenv("MHAPB_K1_SUPER_GB")) gb = atof(e);
        super_cap = ctx->super_cap = std::max<uint64_t>(chunk_cap, (uint64_t)(gb * 1e9 / 12.0));
    }
    t_plan = since();
    int launches = 0;
    struct Sub { int begin, n, first_long, max_k_short, max_k_long, max_len_short, max_len_long; uint64_t lo, hi; };
    auto is_short_len = [&](uint64_t len) { return (int64_t)len - k + 1 <= kShortMaxKmers && (int64_t)len - ok + 1 <= kShortMaxKmers + 64; };
    auto is_short = [&](const StrandDesc &d) { return is_short_len(d.len); };
    // events of the call; destroyed on every way out of the function
    struct EventBag : std::vector<cudaEvent_t> { ~EventBag() { for (cudaEvent_t e : *this) cudaEventDestroy(e); } };
    EventBag evs;     // per chunk: [before K1a, after K1a, after K1c]; per super-chunk: [before K1b, after K1b]
    EventBag hevs, cpevs;
    auto ev_new = [&](EventBag &v, cudaStream_t st) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); v.push_back(e); };
    std::vector<size_t> k1a_ev, k1b_ev;
    uint64_t copied_hi = 0;
    CU(ctx, cudaStreamSynchronize(ctx->stream2));   // the previous call's copies are long done; keeps the stream's order simple
    cudaEvent_t first_ev = nullptr, last_ev = nullptr;

    size_t pos = 0;
    // strand descriptors are planned into PINNED memory: a copy from pageable memory makes the host wait for the stream, and
    // then the next chunk's characters are not on their way while this chunk is hashed
    std::vector<uint32_t> slot_of;                       // want_valid: caller's slot of every strand of the super-chunk
    while (pos < n_total) {
        // ---- extent of one super-chunk (sizes only: the first K1a launch should not wait for the whole plan) ----
        size_t send = pos; uint64_t total_k = 0; size_t n_chunks = 0;
        int max_k_short = 1, max_k_long = 1, max_len_long = 1; bool any_short = false, any_long = false;
        {
            uint64_t in_chunk = 0; size_t in_chunk_n = 0;
            while (send < n_total) {
                const uint64_t nk = nk_of[send];
                if (send > pos && total_k + nk > super_cap) break;
                if (in_chunk_n == 0 || in_chunk + nk > chunk_cap || (int)in_chunk_n >= max_chunk_strands) { n_chunks++; in_chunk = 0; in_chunk_n = 0; }
                in_chunk += nk; in_chunk_n++;
                total_k += nk;
                const uint64_t len_s = nk + (uint64_t)k - 1;
                if (is_short_len(len_s)) { any_short = true; max_k_short = std::max(max_k_short, (int)nk); }
                else { any_long = true; max_k_long = std::max(max_k_long, (int)nk); max_len_long = std::max(max_len_long, (int)len_s); }
                send++;
            }
        }
        const int n_all = (int)(send - pos);
        // ---- scratch for the super-chunk ----
        CU(ctx, ctx->desc.ensure((size_t)n_all * sizeof(StrandDesc)));
        CU(ctx, ctx->h_desc.ensure((size_t)n_all * sizeof(StrandDesc)));
        StrandDesc *hd = ctx->h_desc.as<StrandDesc>();
        // sketches wider than 512 words run as `passes` blocks of <= 512 words (virtual strands, see k_minhash_bs2): block p needs
        // its own copy of the keys, advanced by 512*p*weight steps
        static int multipass = -1;
        if (multipass < 0) { const char *e = getenv("MHAPB_K1B_PASSES"); multipass = e ? atoi(e) : 1; }
        const int passes = (multipass && H > 512 && H <= kMaxNumHashes && d_minhash) ? (H + 511) / 512 : 1;
        {   // the key scratch is the big allocation: if the memory is not there (the store grew since the bound was sized), halve the
            // super-chunk and plan again instead of failing
            cudaError_t ek = ctx->keys.ensure((size_t)total_k * 8 * (size_t)passes);
            if (ek == cudaSuccess) ek = ctx->wts.ensure((size_t)total_k * 4);
            if (ek == cudaErrorMemoryAllocation && super_cap > chunk_cap && n_chunks > 1) {
                (void)cudaGetLastError();
                super_cap = ctx->super_cap = std::max<uint64_t>(chunk_cap, super_cap / 2);
                continue;
            }
            CU(ctx, ek);
        }
        CU(ctx, ctx->nlight.ensure((size_t)n_all * 4));
        CU(ctx, ctx->nheavy.ensure((size_t)n_all * 4));
        const size_t n_counters = (n_chunks + 2) * 4 + 4;        // per chunk: K1a short, K1a long, K1c short, K1c long; + K1b (after the chunks' blocks)
        CU(ctx, ctx->counters.ensure(n_counters * 4));
        const int g1 = hash_dedup_grid() * 2, g2 = ordered_grid();
        size_t dup_need = 0;
        if (any_short) dup_need = (size_t)g1 * dedup_table_slots((uint32_t)max_k_short) * 4;
        if (any_long) {
            const size_t cap_long = dedup_table_slots((uint32_t)max_k_long);
            dup_need = std::max(dup_need, (size_t)g1 * cap_long * 4);
            CU(ctx, ctx->gtable.ensure((size_t)g1 * cap_long * 8));
            CU(ctx, ctx->ohash.ensure((size_t)g2 * ((size_t)max_len_long + 32) * 4));
        }
        if (dup_need > ctx->dupcnt.cap) {
            CU(ctx, ctx->dupcnt.ensure(dup_need));
            CU(ctx, cudaMemsetAsync(ctx->dupcnt.p, 0, ctx->dupcnt.cap, ctx->stream));   // invariant: zero between uses
        }
        CU(ctx, cudaMemsetAsync(ctx->counters.p, 0, n_counters * 4, ctx->stream));
        {   // the copy stream starts this super-chunk after everything queued so far on the compute stream (buffers may have moved)
            ev_new(cpevs, ctx->stream);
            CU(ctx, cudaStreamWaitEvent(ctx->stream2, cpevs.back(), 0));
        }
        SketchScratch sc;
        sc.keys = ctx->keys.as<uint64_t>(); sc.wts = ctx->wts.as<uint32_t>();
        sc.nlight = ctx->nlight.as<int32_t>(); sc.nheavy = ctx->nheavy.as<int32_t>();
        sc.dupcnt = ctx->dupcnt.as<uint32_t>(); sc.gtable = ctx->gtable.as<uint64_t>();
        sc.ohash = ctx->ohash.as<uint32_t>(); sc.counters = ctx->counters.as<uint32_t>();
        StrandDesc *dd = ctx->desc.as<StrandDesc>();
        if (want_valid) slot_of.assign((size_t)n_all, 0);
        // ---- per chunk: plan, descriptors + characters in, K1a, K1c (the host plans chunk i+1 while the GPU hashes chunk i) ----
        uint64_t koff = 0;
        size_t ci = 0;
        while (pos < send) {
            uint64_t tot = 0;
            size_t q = pos;
            while (q < send && (int)(q - pos) < max_chunk_strands) {
                const uint64_t nk = nk_of[q];
                if (q > pos && tot + nk > chunk_cap) break;
                tot += nk; q++;
            }
            fill_until(q);                                        // this chunk's descriptors (the later ones while the GPU hashes this chunk)
            const int c0 = n_all - (int)(send - pos);             // index of this chunk's first strand in the super-chunk
            StrandDesc *cd = hd + c0;
            const int cn = (int)(q - pos);
            std::copy(all.begin() + pos, all.begin() + q, cd);
            pos = q;
            std::stable_partition(cd, cd + cn, is_short);         // short strands first, long (global-table) strands after
            Sub sb{c0, cn, cn, 1, 1, 1, 1, ~0ull, 0};
            for (int i = 0; i < sb.n; i++) {
                StrandDesc &d = cd[i];
                d.koff = koff; koff += d.len - k + 1;
                const int nk = (int)d.len - k + 1;
                if (is_short(d)) { sb.max_k_short = std::max(sb.max_k_short, nk); sb.max_len_short = std::max(sb.max_len_short, (int)d.len); }
                else { if (sb.first_long == sb.n) sb.first_long = i; sb.max_k_long = std::max(sb.max_k_long, nk); sb.max_len_long = std::max(sb.max_len_long, (int)d.len); }
                sb.lo = std::min<uint64_t>(sb.lo, d.base_off); sb.hi = std::max<uint64_t>(sb.hi, d.base_off + d.len);
                ctx->timing.xorshift_steps += (int64_t)nk * H;
                ctx->timing.kmers_hashed += nk;
                if (want_valid) slot_of[(size_t)c0 + i] = d.slot;
            }
            // Descriptors and characters of the chunk travel on the COPY stream, the compute stream waits for their event.
            // (A descriptor copy on the compute stream sits behind the previous chunk's kernels in stream order but in front
            // of this chunk's characters in the copy engine's queue: head-of-line blocking made every chunk's H2D wait for the
            // previous chunk's K1a/K1c -- 18 ms of exposed copies per 1 GB of reads, found with MHAPB_TRACE.)
            CU(ctx, cudaMemcpyAsync(dd + c0, cd, (size_t)cn * sizeof(StrandDesc), cudaMemcpyHostToDevice, ctx->stream2));
            if (h_bases) {
                const uint64_t lo = std::max(sb.lo, copied_hi);   // a read whose two strands straddle two chunks was copied with the first
                if (sb.hi > lo) {
                    ev_new(hevs, ctx->stream2);
                    CU(ctx, cudaMemcpyAsync(const_cast<uint8_t *>(d_bases) + lo, h_bases + lo, (size_t)(sb.hi - lo), cudaMemcpyHostToDevice, ctx->stream2));
                    ev_new(hevs, ctx->stream2);
                    copied_hi = sb.hi;
                }
            }
            ev_new(cpevs, ctx->stream2);
            CU(ctx, cudaStreamWaitEvent(ctx->stream, cpevs.back(), 0));
            uint32_t *qs = sc.counters + ci * 4;
            ci++;
            k1a_ev.push_back(evs.size());
            ev_new(evs, ctx->stream);                                     // before K1a
            if (!first_ev) { first_ev = evs.back(); t_first = since(); }
            // table capacities: the super-chunk's maxima (dupcnt rows are strided by the launch's capacity)
            CU(ctx, launch_hash_dedup(ctx->stream, d_bases, dd, sb.begin, sb.n, sb.first_long, max_k_short, max_k_long, k, filter_unweighted(flt, p.unweighted), flt, sc, qs, &launches));
            ev_new(evs, ctx->stream);                                     // after K1a
            if (d_ord) CU(ctx, launch_ordered(ctx->stream, d_bases, dd, sb.begin, sb.n, sb.first_long, sb.max_len_short, std::max(sb.max_len_long, max_len_long), ok, S, ord_stride, sc, d_ord, d_ord_n, 0, qs + 2, &launches));
            ev_new(evs, ctx->stream);                                     // after K1c
        }
        if (want_valid) {   // which strands kept at least one k-mer
            h_nl.resize(n_all); h_nh.resize(n_all);
            CU(ctx, cudaMemcpyAsync(h_nl.data(), sc.nlight, (size_t)n_all * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaMemcpyAsync(h_nh.data(), sc.nheavy, (size_t)n_all * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < n_all; i++) if (h_nl[i] + h_nh[i] == 0) (*slot_valid)[slot_of[(size_t)i]] = 0;
        }
        // ---- K1b: one launch over the whole super-chunk ----
        k1b_ev.push_back(evs.size());
        ev_new(evs, ctx->stream);
        if (d_minhash && passes > 1) {
            if (!ctx->t512.p) {   // step^512 as eight byte-indexed tables (the map is linear over GF(2))
                std::vector<uint64_t> t(8 * 256);
                for (int b = 0; b < 8; b++)
                    for (int v = 0; v < 256; v++) {
                        uint64_t x = (uint64_t)v << (8 * b);
                        for (int i = 0; i < 512; i++) { x ^= x << 21; x ^= x >> 35; x ^= x << 4; }   // MinHashSketch.java:140-143
                        t[(size_t)b * 256 + v] = x;
                    }
                CU(ctx, ctx->t512.ensure(t.size() * 8));
                CU(ctx, cudaMemcpyAsync(ctx->t512.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
                CU(ctx, cudaStreamSynchronize(ctx->stream));
            }
            // virtual strands, block-major: v = p * n_all + i
            CU(ctx, ctx->h_vdesc.ensure((size_t)n_all * passes * sizeof(StrandDesc)));
            CU(ctx, ctx->vdesc.ensure((size_t)n_all * passes * sizeof(StrandDesc)));
            StrandDesc *hv = ctx->h_vdesc.as<StrandDesc>();
            for (int pp = 0; pp < passes; pp++)
                for (int i = 0; i < n_all; i++) {
                    StrandDesc v = hd[i];
                    v.base_off = hd[i].koff;                              // the strand's hashes and weights
                    v.koff = (uint64_t)pp * total_k + hd[i].koff;         // this block's chain starts
                    v.rc = (uint32_t)std::min(512, H - 512 * pp);         // words in the block
                    v.slot = (uint32_t)(512 * pp);                        // first word
                    hv[(size_t)pp * n_all + i] = v;
                }
            CU(ctx, cudaMemcpyAsync(ctx->vdesc.p, hv, (size_t)n_all * passes * sizeof(StrandDesc), cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, launch_advance_keys(ctx->stream, dd, n_all, k, sc, ctx->t512.as<uint64_t>(), total_k, passes, flt.light_weight, &launches));
            CU(ctx, launch_minhash_virtual(ctx->stream, ctx->vdesc.as<StrandDesc>(), n_all * passes, n_all, k, H, sc, d_minhash, flt.light_weight,
                                           sc.counters + (n_chunks + 2) * 4, &launches));
        } else if (d_minhash) CU(ctx, launch_minhash(ctx->stream, dd, n_all, k, H, sc, d_minhash, flt.light_weight, sc.counters + (n_chunks + 2) * 4, &launches));
        ev_new(evs, ctx->stream);
        last_ev = evs.back();
        if (pos < n_total) CU(ctx, cudaStreamSynchronize(ctx->stream));   // the next super-chunk reuses desc / keys / the pinned plan
    }
    t_enq = since();
    int cb_rc = MHAPB_OK;
    if (on_enqueued) cb_rc = (*on_enqueued)();      // host work of the caller that can run while the GPU sketches (store metadata)
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (trace) fprintf(stderr, "[mhapb] sketch_core host ms: strands planned %.2f, first K1a enqueued %.2f, all enqueued %.2f, synced %.2f (h2d %s)\n",
                       t_plan, t_first, t_enq, since(), h_bases ? "per chunk" : "none");
    for (size_t i : k1a_ev) {
        float a = 0, c = 0;
        cudaEventElapsedTime(&a, evs[i], evs[i + 1]); cudaEventElapsedTime(&c, evs[i + 1], evs[i + 2]);
        ctx->timing.hash_dedup_ms += a; ctx->timing.ordered_ms += c;
    }
    for (size_t i : k1b_ev) { float b = 0; cudaEventElapsedTime(&b, evs[i], evs[i + 1]); ctx->timing.minhash_ms += b; }
    for (size_t i = 0; i + 1 < hevs.size(); i += 2) { float c = 0; cudaEventElapsedTime(&c, hevs[i], hevs[i + 1]); ctx->timing.h2d_ms += c; }
    if (first_ev && last_ev) { float t = 0; cudaEventElapsedTime(&t, first_ev, last_ev); ctx->timing.sketch_total_ms += t; }
    ctx->timing.kernel_launches += launches;
    return cb_rc;
}

void reset_sketch_timing(mhapb_ctx *ctx)
{
    ctx->timing.h2d_ms = ctx->timing.d2h_ms = 0;
    ctx->timing.hash_dedup_ms = ctx->timing.minhash_ms = ctx->timing.ordered_ms = 0;
    ctx->timing.sketch_total_ms = 0;
    ctx->timing.kernel_launches = 0;
}

double jaccard_to_identity(double score, int kmer_size)
{
    // sketch/BottomOverlapSketch.java:391-395
    double d = -1.0 / (double)kmer_size * std::log(2.0 * score / (1.0 + score));
    return std::exp(-d);
}

int store_configure(mhapb_ctx *ctx, const mhapb_sketch_params *p)
{
    int rc = check_sketch_params(ctx, p);
    if (rc) return rc;
    Store &s = ctx->store;
    s.p = *p; s.configured = true; s.n = 0; s.ord_stride = p->ordered_sketch_size; s.indexed = false;
    s.h_id.clear(); s.h_fwd.clear(); s.h_len.clear(); s.h_lenk.clear(); s.h_ordn.clear(); s.seen.clear(); s.ids_monotonic = true; s.any_key = false; s.last_key = 0; s.fwd_list_valid = false;
    return MHAPB_OK;
}

int store_reserve(mhapb_ctx *ctx, int64_t extra)
{
    Store &s = ctx->store;
    const size_t n = (size_t)s.n, want = (size_t)(s.n + extra);
    const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.ord_stride;
    CU(ctx, s.minhash.grow(want * H * 4, n * H * 4, ctx->stream));
    CU(ctx, s.ord.grow(want * S * 8, n * S * 8, ctx->stream));
    CU(ctx, s.ord_n.grow(want * 4, n * 4, ctx->stream));
    CU(ctx, s.lenk.grow(want * 4, n * 4, ctx->stream));
    CU(ctx, s.len.grow(want * 4, n * 4, ctx->stream));
    CU(ctx, s.id.grow(want * 8, n * 8, ctx->stream));
    return MHAPB_OK;
}

// push the per-sketch host columns of rows [from, n) to the device
int store_sync_columns(mhapb_ctx *ctx, int64_t from)
{
    Store &s = ctx->store;
    const size_t cnt = (size_t)(s.n - from);
    if (!cnt) return MHAPB_OK;
    CU(ctx, cudaMemcpyAsync(s.lenk.as<int32_t>() + from, s.h_lenk.data() + from, cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(s.len.as<int32_t>() + from, s.h_len.data() + from, cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(s.id.as<int64_t>() + from, s.h_id.data() + from, cnt * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MHAPB_OK;
}

int store_push_meta(mhapb_ctx *ctx, int64_t id, int fwd, int32_t len, int32_t lenk, int32_t ordn)
{
    Store &s = ctx->store;
    // order: (id, forward) before (id, reverse) before (id+1, forward), as the streamer produces them
    const uint64_t key = (((uint64_t)id ^ 0x8000000000000000ull) << 1) | (uint64_t)(fwd ? 0 : 1);
    if (s.ids_monotonic) {
        if (!s.any_key || key > s.last_key) { s.last_key = key; s.any_key = true; }
        else {   // out of order: fall back to the set, built once from what is stored
            s.ids_monotonic = false;
            s.seen.clear();
            for (size_t i = 0; i < s.h_id.size(); i++) s.seen.insert((((uint64_t)s.h_id[i] ^ 0x8000000000000000ull) << 1) | (uint64_t)(s.h_fwd[i] ? 0 : 1));
        }
    }
    if (!s.ids_monotonic && !s.seen.insert(key).second) return fail(ctx, MHAPB_EDUPID, "Sequence ID already exists in the hash table.");
    s.h_id.push_back(id); s.h_fwd.push_back((uint8_t)(fwd ? 1 : 0)); s.h_len.push_back(len); s.h_lenk.push_back(lenk); s.h_ordn.push_back(ordn);
    return MHAPB_OK;
}

// drop the host columns pushed since meta0 (an add that failed after store_push_meta must leave the store as it was)
void store_rollback_meta(Store &s, size_t meta0)
{
    if (!s.ids_monotonic) for (size_t i = meta0; i < s.h_id.size(); i++) s.seen.erase((((uint64_t)s.h_id[i] ^ 0x8000000000000000ull) << 1) | (uint64_t)(s.h_fwd[i] ? 0 : 1));
    s.h_id.resize(meta0); s.h_fwd.resize(meta0); s.h_len.resize(meta0); s.h_lenk.resize(meta0); s.h_ordn.resize(meta0);
    if (s.ids_monotonic) {
        s.any_key = meta0 > 0;
        if (meta0 > 0) s.last_key = (((uint64_t)s.h_id[meta0 - 1] ^ 0x8000000000000000ull) << 1) | (uint64_t)(s.h_fwd[meta0 - 1] ? 0 : 1);
    }
}

int index_build(mhapb_ctx *ctx)
{
    Store &s = ctx->store;
    if (!s.configured || s.n == 0) return fail(ctx, MHAPB_ESTATE, "index build on an empty store");
    if (s.indexed) return MHAPB_OK;
    const int H = s.p.num_hashes;
    if ((uint64_t)s.n * (uint64_t)H >= 0xfffffff0ull || s.n >= 0x7fffffff) return fail(ctx, MHAPB_EINVAL, "store too large for 32-bit postings (%lld sketches x %d)", (long long)s.n, H);
    // Sub-table size: two slots per sketch, the size that can never overflow (a word's sub-table holds its DISTINCT values, at
    // most one per sketch).  With noisy long reads that bound is nearly reached: at 15 % error only 7 % of the 16-mers are
    // error-free, so the minimum of a read is almost always a k-mer nobody shares (measured on configs[1]: the optimistic size of
    // one slot per two sketches overflowed).  MHAPB_INDEX_OPTIMISTIC=1 (read when the context is created) starts with that
    // smaller size for low-error data -- 4x fewer slots to memset, scan and pack -- and verifies: an insert that needs more
    // than 128 probes raises IndexView::overflow, K2b then does nothing, and the search rebuilds with the safe size.
    const long long want = (ctx->index_optimistic && !s.index_safe) ? std::max<long long>(s.n / 2, 1024) : 2 * s.n;
    int lg = 5; while ((1ll << lg) < want) lg++;
    s.log2capw = lg;
    const size_t nslots = (size_t)H << lg;
    CU(ctx, s.slots.ensure(nslots * 8));
    CU(ctx, s.postings.ensure((size_t)s.n * H * 4));
    CU(ctx, s.present.ensure(nslots / 8 + 64));
    CU(ctx, s.idx_flag.ensure(64));
    CU(ctx, ctx->tmp_start.ensure(nslots * 4));
    CU(ctx, ctx->block_sums.ensure(((nslots + 4095) / 4096 + 1) * 4));
    IndexView iv{s.slots.as<uint64_t>(), s.postings.as<uint32_t>(), s.present.as<uint32_t>(), lg, H, s.n, s.idx_flag.as<uint32_t>(), 0};
    int launches = 0;
    cudaEventRecord(ctx->ev[8], ctx->stream);
    CU(ctx, launch_index_build(ctx->stream, s.minhash.as<int32_t>(), s.n, H, iv, ctx->tmp_start.as<uint32_t>(), ctx->block_sums.as<uint32_t>(), &launches));
    cudaEventRecord(ctx->ev[9], ctx->stream);
    ctx->index_timing_pending = true;     // no synchronisation here: the probe is enqueued right behind the build
    ctx->timing.kernel_launches += launches;
    s.indexed = true;
    return MHAPB_OK;
}

// K2b + K2c + result compaction as ONE enqueue: the candidate count, the warp kernel's overflow count and the number
// of surviving pairs stay on the device (every consumer reads its count from the producer's cursor), so the host
// synchronises twice per search -- once for the counters, once for the surviving pairs -- instead of after every kernel.
// Capacities are guesses that the counters validate afterwards; an overrun (rare: repeat-rich data) repeats the
// search with exact sizes.
int search_core(mhapb_ctx *ctx, const mhapb_search_params *sp, const QuerySet &q, int to_self,
                mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    Store &s = ctx->store;
    int rc = index_build(ctx);
    if (rc) return rc;
    const int H = s.p.num_hashes;
    IndexView iv{s.slots.as<uint64_t>(), s.postings.as<uint32_t>(), s.present.as<uint32_t>(), s.log2capw, H, s.n, s.idx_flag.as<uint32_t>(),
                 q.d_minhash != s.minhash.as<int32_t>() ? 1 : 0};
    const int64_t nq = q.list_all ? q.n_all : (q.d_list ? q.n_list : (int64_t)q.list.size());
    mhapb_stats st{};
    st.sequences_searched = nq;
    mhapb_hit *hit_arr = nullptr; size_t n_hits = 0;   // malloc'd result, filled by the scoring threads
    int launches = 0;
    ctx->timing.probe_ms = ctx->timing.filter_ms = 0;
    if (nq > 0) {
        const uint32_t *d_qlist = q.d_list;
        if (!q.list_all && !d_qlist) {
            CU(ctx, ctx->qlist.ensure((size_t)nq * 4));
            CU(ctx, cudaMemcpyAsync(ctx->qlist.p, q.list.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, ctx->stream));
            d_qlist = ctx->qlist.as<uint32_t>();
        }
        CU(ctx, ctx->scounters.ensure(128));
        CU(ctx, ctx->ovf_q.ensure((size_t)nq * 4));
        // counters (u64): [0] candidates [1] elements processed [2] sequences hit [3] probe overflow queries
        //                 [4] K2c overflow pairs [5] surviving pairs
        unsigned long long cnt[8] = {0};
        uint64_t cand_cap = std::max<uint64_t>(std::max<uint64_t>(1 << 16, (uint64_t)nq * 16), ctx->cand_cap_hint);
        uint32_t ovf_threads = std::max<uint32_t>(4096u, ctx->ovf_threads_hint);
        const int ok = s.p.ordered_kmer_size;
        // Only pairs that can still reach the threshold travel to the host.  score >= accept  <=>  jaccard >= T/(2-T) with
        // T = accept^ok (jaccardToIdentity is increasing); the device test uses that bound lowered by 1e-9 relative, the
        // exact double-precision decision (MinHashSearch.java:229) is taken below on the survivors.
        double jmin = 0.0;
        if (sp->accept_score > 0.0) {
            const double T = std::pow(sp->accept_score, (double)ok);
            jmin = T < 2.0 ? (T / (2.0 - T)) * (1.0 - 1e-9) - 1e-12 : 2.0;
            if (jmin < 0.0) jmin = 0.0;
        }
        const int keep_all = sp->keep_all || sp->accept_score <= 0.0;
        for (int attempt = 0; attempt < 3; attempt++) {
            CU(ctx, ctx->cand.ensure(cand_cap * sizeof(Candidate)));
            CU(ctx, ctx->ovl.ensure(cand_cap * sizeof(OverlapOut)));
            CU(ctx, ctx->ovf_list.ensure(cand_cap * 4 + 16));
            CU(ctx, ctx->cand2.ensure(cand_cap * sizeof(Candidate)));
            CU(ctx, ctx->ovl2.ensure(cand_cap * sizeof(OverlapOut)));
            const uint32_t entries = 2u * (uint32_t)std::max(q.ord_stride, s.ord_stride) + 2u;
            CU(ctx, ctx->fscratch.ensure((size_t)3 * entries * ovf_threads * 4));
            CU(ctx, cudaMemsetAsync(ctx->scounters.p, 0, 128, ctx->stream));
            unsigned long long *dc = ctx->scounters.as<unsigned long long>();
            ProbeArgs a{};
            a.q_minhash = q.d_minhash; a.q_id = q.d_id; a.q_len = q.d_len; a.q_list = d_qlist; a.nq_list = nq;
            a.t_id = s.id.as<int64_t>(); a.t_len = s.len.as<int32_t>();
            a.to_self = to_self; a.num_min_matches = sp->num_min_matches; a.min_store_length = sp->min_store_length;
            a.cand = ctx->cand.as<Candidate>(); a.cand_cap = cand_cap; a.counters = dc; a.ovf_q = ctx->ovf_q.as<uint32_t>();
            if (q.minhash_ready) CU(ctx, cudaStreamWaitEvent(ctx->stream, q.minhash_ready, 0));
            cudaEventRecord(ctx->ev[0], ctx->stream);
            CU(ctx, launch_probe(ctx->stream, iv, a, &launches));
            cudaEventRecord(ctx->ev[1], ctx->stream);
            if (q.ord_ready) CU(ctx, cudaStreamWaitEvent(ctx->stream, q.ord_ready, 0));
            FilterArgs f{};
            f.cand = ctx->cand.as<Candidate>(); f.n_cand = 0; f.n_cand_dev = dc + 0; f.cand_cap = cand_cap;
            f.q_ord = q.d_ord; f.q_ord_n = q.d_ordn; f.q_lenk = q.d_lenk; f.q_stride = q.ord_stride;
            f.t_ord = s.ord.as<int32_t>(); f.t_ord_n = s.ord_n.as<int32_t>(); f.t_lenk = s.lenk.as<int32_t>(); f.t_stride = s.ord_stride;
            f.max_shift = sp->max_shift;
            f.out = ctx->ovl.as<OverlapOut>();
            f.ovf_list = ctx->ovf_list.as<uint32_t>(); f.ovf_count = dc + 4;
            cudaEventRecord(ctx->ev[2], ctx->stream);
            // warp-per-candidate kernel; pairs with more match records than its shared buffer holds (near-identical reads)
            // are listed for the thread-per-candidate kernel, which reads the list length on the device
            CU(ctx, launch_filter_warp(ctx->stream, f, &launches));
            f.sel = ctx->ovf_list.as<uint32_t>(); f.n_sel_dev = dc + 4; f.n_sel = 0;
            f.scratch = ctx->fscratch.as<int32_t>(); f.scratch_entries = entries; f.n_threads = ovf_threads;
            CU(ctx, launch_filter(ctx->stream, f, &launches));
            cudaEventRecord(ctx->ev[3], ctx->stream);
            CU(ctx, launch_compact_hits(ctx->stream, ctx->cand.as<Candidate>(), ctx->ovl.as<OverlapOut>(), cand_cap, dc + 0, jmin, keep_all,
                                        ctx->cand2.as<Candidate>(), ctx->ovl2.as<OverlapOut>(), dc + 5, &launches));
            uint32_t idx_overflow = 0;
            CU(ctx, cudaMemcpyAsync(cnt, dc, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaMemcpyAsync(&idx_overflow, s.idx_flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));                      // sync 1 of 2: the counters
            if (idx_overflow) {   // the optimistic sub-table size was too small for this store: rebuild safely, search again
                if (s.index_safe) return fail(ctx, MHAPB_ECUDA, "index overflow with the safe table size");
                s.index_safe = true; s.indexed = false;
                rc = index_build(ctx);
                if (rc) return rc;
                iv = IndexView{s.slots.as<uint64_t>(), s.postings.as<uint32_t>(), s.present.as<uint32_t>(), s.log2capw, H, s.n, s.idx_flag.as<uint32_t>(), iv.use_present};
                attempt = -1;
                continue;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->timing.probe_ms += ms;
            cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->timing.filter_ms += ms;
            if (cnt[0] <= cand_cap) break;
            cand_cap = cnt[0];   // rare: more candidates than guessed; rerun with the exact size
        }
        if (cnt[0] > cand_cap) return fail(ctx, MHAPB_ECUDA, "candidate buffer overrun after resize");
        ctx->cand_cap_hint = std::min<uint64_t>(cnt[0] + cnt[0] / 8, 1ull << 31);
        if (cnt[4] > ovf_threads) ctx->ovf_threads_hint = (uint32_t)std::min<uint64_t>((cnt[4] + 127) & ~127ull, 148ull * 1024);
        st.elements_processed = (int64_t)cnt[1];
        st.sequences_hit = (int64_t)cnt[2];
        st.fully_compared = (int64_t)cnt[0];
        const unsigned long long nkeep = cnt[5];
        if (nkeep > 0) {
            CU(ctx, ctx->h_cand.ensure(nkeep * sizeof(Candidate)));
            CU(ctx, ctx->h_ovl.ensure(nkeep * sizeof(OverlapOut)));
            const Candidate *hc = ctx->h_cand.as<Candidate>();
            const OverlapOut *ho = ctx->h_ovl.as<OverlapOut>();
            cudaEventRecord(ctx->ev[6], ctx->stream);
            CU(ctx, cudaMemcpyAsync(ctx->h_cand.p, ctx->cand2.p, nkeep * sizeof(Candidate), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaMemcpyAsync(ctx->h_ovl.p, ctx->ovl2.p, nkeep * sizeof(OverlapOut), cudaMemcpyDeviceToHost, ctx->stream));
            cudaEventRecord(ctx->ev[7], ctx->stream);
            CU(ctx, cudaStreamSynchronize(ctx->stream));                      // sync 2 of 2: the surviving pairs
            cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[6], ctx->ev[7]);
            // score + MatchResult fields, split over host threads (order of hits is unspecified, as in the reference)
            hit_arr = (mhapb_hit *)malloc(sizeof(mhapb_hit) * (size_t)nkeep);
            if (!hit_arr) return fail(ctx, MHAPB_ENOMEM, "malloc hits");
            const unsigned nthr = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), nkeep / 16384 + 1));
            std::vector<int64_t> acc(nthr, 0);
            std::vector<uint64_t> kept(nthr, 0);
            // every thread scores a contiguous slice in place (slot i of the result for pair i); rejected pairs leave
            // holes that are closed afterwards
            auto work = [&](unsigned t) {
                const uint64_t lo = nkeep * t / nthr, hi = nkeep * (t + 1) / nthr;
                uint64_t w = lo;
                for (uint64_t i = lo; i < hi; i++) {
                    const OverlapOut &o = ho[i];
                    double score = 0.0;   // OverlapInfo.EMPTY
                    if (!o.empty) {
                        double jac = o.kmin ? (double)o.inter / (double)o.kmin : 0.0;
                        score = jaccard_to_identity(jac, ok);
                    }
                    const bool accept = score >= sp->accept_score;   // MinHashSearch.java:229
                    if (accept) acc[t]++;
                    if (!accept && !sp->keep_all) continue;
                    mhapb_hit h{};
                    const uint32_t qi = hc[i].q, ti = hc[i].t;
                    h.from_id = q.h_id[qi]; h.to_id = s.h_id[ti];
                    h.from_fwd = q.h_fwd ? q.h_fwd[qi] : 1; h.to_fwd = s.h_fwd[ti];
                    h.hit_count = (int32_t)hc[i].count;
                    h.a1 = o.a1; h.a2 = o.a2; h.b1 = o.b1; h.b2 = o.b2;
                    h.valid_count = o.valid; h.intersect = o.inter; h.kmin = o.kmin;
                    h.from_len = q.h_len[qi]; h.to_len = s.h_len[ti];
                    h.score = score; h.accepted = accept ? 1 : 0;
                    hit_arr[w++] = h;
                }
                kept[t] = w - lo;
            };
            if (nthr == 1) work(0);
            else {
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nthr; t++) th.emplace_back(work, t);
                for (auto &x : th) x.join();
            }
            size_t total = 0;
            for (unsigned t = 0; t < nthr; t++) {
                const uint64_t lo = nkeep * t / nthr;
                if (total != lo && kept[t]) memmove(hit_arr + total, hit_arr + lo, sizeof(mhapb_hit) * kept[t]);
                total += kept[t]; st.matches_processed += acc[t];
            }
            n_hits = total;
        }
    }
    ctx->timing.kernel_launches += launches;
    ctx->timing.search_total_ms = ctx->timing.probe_ms + ctx->timing.filter_ms;
    if (stats) *stats = st;
    if (n_out) *n_out = n_hits;
    if (out) {
        if (!hit_arr) hit_arr = (mhapb_hit *)malloc(sizeof(mhapb_hit));
        if (!hit_arr) return fail(ctx, MHAPB_ENOMEM, "malloc hits");
        *out = hit_arr;
    } else free(hit_arr);
    return MHAPB_OK;
}

int h2d_bases(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads)
{
    const uint64_t total = offsets[n_reads] - offsets[0];
    CU(ctx, ctx->bases.ensure((size_t)offsets[n_reads] + 64));
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (total) CU(ctx, cudaMemcpyAsync(ctx->bases.as<uint8_t>() + offsets[0], bases + offsets[0], (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev[5], ctx->stream);
    return MHAPB_OK;
}

// sketch reads and append them to the store (shared by store_add_reads and the tests' store_get path)
int add_reads_locked(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads, int both, int64_t *n_added,
                     const uint8_t *d_resident = nullptr /* the reads are already in HBM at this address: no H2D */)
{
    Store &s = ctx->store;
    if (!s.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    const int per = both ? 2 : 1;
    std::vector<int64_t> rows((size_t)n_reads * per, -1);
    const int64_t n0 = s.n;
    int64_t next = n0;
    // with a filter that can drop k-mers a strand may be left with none (ZeroNGramsFoundException): a K1a-only pass
    // finds those before rows are assigned.  Forward strand empty => the read is skipped; only the reverse strand
    // empty => the forward sketch alone is stored (SequenceSketchStreamer.java:123-156,225-240).
    std::vector<uint8_t> valid((size_t)n_reads * per, 1);
    bool bases_on_device = d_resident != nullptr;
    if (d_resident) { cudaEventRecord(ctx->ev[4], ctx->stream); cudaEventRecord(ctx->ev[5], ctx->stream); }
    auto dev_bases = [&]() { return d_resident ? d_resident : ctx->bases.as<uint8_t>(); };
    if (filter_can_empty(filter_view(ctx)) && n_reads) {
        std::vector<int64_t> ident((size_t)n_reads * per);
        for (size_t i = 0; i < ident.size(); i++) ident[i] = (int64_t)i;
        int rc0 = bases_on_device ? MHAPB_OK : h2d_bases(ctx, bases, offsets, n_reads);
        if (rc0) return rc0;
        bases_on_device = true;
        rc0 = sketch_core(ctx, s.p, dev_bases(), offsets, n_reads, both, ident, nullptr, nullptr, 0, nullptr, &valid);
        if (rc0) return rc0;
    }
    auto strand_kept = [&](uint32_t r, int st) { return valid[(size_t)r * per] && valid[(size_t)r * per + st]; };
    for (uint32_t r = 0; r < n_reads; r++) {
        uint64_t len = offsets[r + 1] - offsets[r];
        if (read_status(s.p, len)) continue;
        for (int st = 0; st < per; st++) if (strand_kept(r, st)) rows[(size_t)r * per + st] = next++;
    }
    const int64_t added = next - n0;
    if (n_added) *n_added = added;
    if (!added) return MHAPB_OK;
    // The per-sketch host columns (ids, lengths, the duplicate-id check of MinHashSearch.java:112-117) are filled WHILE the GPU
    // sketches: sketch_core runs this once everything is enqueued.  A duplicate id then surfaces after the device work, which
    // wrote rows beyond s.n that nobody reads; the store is left exactly as it was (s.n is only advanced at the end).
    const size_t meta0 = s.h_id.size();
    const std::function<int()> push_meta = [&]() -> int {
        for (uint32_t r = 0; r < n_reads; r++) {
            uint64_t len = offsets[r + 1] - offsets[r];
            if (read_status(s.p, len)) continue;
            for (int st = 0; st < per; st++) {
                if (!strand_kept(r, st)) continue;
                int32_t no = (int32_t)len - s.p.ordered_kmer_size + 1;
                int rc = store_push_meta(ctx, ids ? ids[r] : (int64_t)r + 1, st == 0, (int32_t)len, no, std::min(no, s.p.ordered_sketch_size));
                if (rc) return rc;
            }
        }
        return MHAPB_OK;
    };
    auto device_part = [&]() -> int {
        int rc = store_reserve(ctx, added);
        if (rc) return rc;
        const bool stream_h2d = !bases_on_device;      // copy chunk by chunk under K1 instead of all reads up front
        if (stream_h2d) CU(ctx, ctx->bases.ensure((size_t)offsets[n_reads] + 64));
        const size_t S = (size_t)s.ord_stride;
        CU(ctx, cudaMemsetAsync(s.ord.as<int32_t>() + (size_t)n0 * S * 2, 0, (size_t)added * S * 8, ctx->stream));
        rc = sketch_core(ctx, s.p, dev_bases(), offsets, n_reads, both, rows, s.minhash.as<int32_t>(), s.ord.as<int32_t>(), s.ord_stride, s.ord_n.as<int32_t>(),
                         nullptr, stream_h2d ? bases : nullptr, &push_meta);
        if (rc) return rc;
        if (!stream_h2d && !d_resident) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->timing.h2d_ms += ms; }
        s.n = next;
        s.indexed = false; s.fwd_list_valid = false;
        rc = store_sync_columns(ctx, n0);
        if (rc) s.n = n0;
        return rc;
    };
    const int rc = device_part();
    if (rc) store_rollback_meta(s, meta0);
    return rc;
}

int sketch_query_reads(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                       std::vector<int64_t> *qid, std::vector<int32_t> *qlen, std::vector<int32_t> *qlenk, const uint8_t *d_resident)
{
    Store &s = ctx->store;
    std::vector<int64_t> rows(n_reads, -1);
    int64_t nq = 0;
    std::vector<uint8_t> valid(n_reads, 1);
    bool bases_on_device = d_resident != nullptr;
    auto dev_bases = [&]() { return d_resident ? d_resident : ctx->bases.as<uint8_t>(); };
    if (filter_can_empty(filter_view(ctx)) && n_reads) {   // see add_reads_locked
        std::vector<int64_t> ident(n_reads);
        for (size_t i = 0; i < ident.size(); i++) ident[i] = (int64_t)i;
        int rc0 = bases_on_device ? MHAPB_OK : h2d_bases(ctx, bases, offsets, n_reads);
        if (rc0) return rc0;
        bases_on_device = true;
        rc0 = sketch_core(ctx, s.p, dev_bases(), offsets, n_reads, 0, ident, nullptr, nullptr, 0, nullptr, &valid);
        if (rc0) return rc0;
    }
    for (uint32_t r = 0; r < n_reads; r++) {
        uint64_t len = offsets[r + 1] - offsets[r];
        if (read_status(s.p, len) || !valid[r]) continue;
        rows[r] = nq++;
        qid->push_back(ids ? ids[r] : (int64_t)r + 1); qlen->push_back((int32_t)len);
        qlenk->push_back((int32_t)len - s.p.ordered_kmer_size + 1);
    }
    const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.p.ordered_sketch_size;
    CU(ctx, ctx->q_minhash.ensure((size_t)nq * H * 4 + 16));
    CU(ctx, ctx->q_ord.ensure((size_t)nq * S * 8 + 16));
    CU(ctx, ctx->q_ordn.ensure((size_t)nq * 4 + 16));
    if (nq) {
        int rc = bases_on_device ? MHAPB_OK : h2d_bases(ctx, bases, offsets, n_reads);
        if (rc) return rc;
        CU(ctx, cudaMemsetAsync(ctx->q_ord.p, 0, (size_t)nq * S * 8, ctx->stream));
        rc = sketch_core(ctx, s.p, dev_bases(), offsets, n_reads, 0, rows, ctx->q_minhash.as<int32_t>(), ctx->q_ord.as<int32_t>(), (int)S, ctx->q_ordn.as<int32_t>());
        if (rc) return rc;
    }
    return MHAPB_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

const char *mhapb_version(void) { return "mhap-b200 0.1 (sm_100a; reference marbl/MHAP 2.1.3)"; }

int mhapb_create(int device_id, mhapb_ctx **out)
{
    if (!out) return fail(nullptr, MHAPB_EINVAL, "null out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(nullptr, MHAPB_ENODEV, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(nullptr, MHAPB_ENODEV, "device %d out of range (%d visible)", device_id, n);
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return fail(nullptr, MHAPB_ENODEV, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return fail(nullptr, MHAPB_ENODEV, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, MHAPB_ENODEV, "device %d is sm_%d%d; this build carries sm_100a code only", device_id, prop.major, prop.minor);
    mhapb_ctx *c = new mhapb_ctx();
    c->device = device_id;
    { const char *e = getenv("MHAPB_INDEX_OPTIMISTIC"); c->index_optimistic = e && atoi(e) != 0; }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return fail(nullptr, MHAPB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    if ((e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return fail(nullptr, MHAPB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    *out = c;
    return MHAPB_OK;
}

void mhapb_destroy(mhapb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->bases, &ctx->desc, &ctx->keys, &ctx->wts, &ctx->nlight, &ctx->nheavy, &ctx->dupcnt, &ctx->gtable, &ctx->ohash,
                      &ctx->counters, &ctx->out_minhash, &ctx->out_ord, &ctx->out_ordn, &ctx->qlist, &ctx->cand, &ctx->ovl, &ctx->cand2, &ctx->ovl2, &ctx->ovf_list, &ctx->fscratch,
                      &ctx->scounters, &ctx->tmp_start, &ctx->block_sums, &ctx->q_minhash, &ctx->q_ord, &ctx->q_ordn, &ctx->q_lenk, &ctx->q_len,
                      &ctx->q_id, &ctx->eq, &ctx->store.minhash, &ctx->store.ord, &ctx->store.ord_n, &ctx->store.lenk, &ctx->store.len,
                      &ctx->store.id, &ctx->store.slots, &ctx->store.postings, &ctx->f_keys, &ctx->f_idf, &ctx->f_used, &ctx->f_bloom};
    for (auto b : bufs) b->release();
    DevBuf *more[] = {&ctx->ovf_q, &ctx->store.fwd_list, &ctx->store.present, &ctx->store.idx_flag, &ctx->g_minhash, &ctx->g_ord, &ctx->g_ordn, &ctx->g_lenk, &ctx->g_len, &ctx->g_id, &ctx->g_pack, &ctx->g_small};
    for (auto b : more) b->release();
    ctx->h_cand.release(); ctx->h_ovl.release(); ctx->h_desc.release(); ctx->h_vdesc.release(); ctx->t512.release(); ctx->vdesc.release();
    comm_release(ctx);
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

const char *mhapb_last_error(const mhapb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void mhapb_free(void *p) { free(p); }

int mhapb_get_timing(mhapb_ctx *ctx, mhapb_timing *out)
{
    if (!ctx || !out) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->index_timing_pending) {
        cudaSetDevice(ctx->device);
        if (cudaEventSynchronize(ctx->ev[9]) == cudaSuccess) cudaEventElapsedTime(&ctx->timing.index_ms, ctx->ev[8], ctx->ev[9]);
        ctx->index_timing_pending = false;
    }
    *out = ctx->timing;
    return MHAPB_OK;
}

int mhapb_host_alloc(size_t bytes, void **out)
{
    if (!out) return MHAPB_EINVAL;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? MHAPB_OK : MHAPB_ENOMEM;
}
void mhapb_host_free(void *p) { if (p) cudaFreeHost(p); }

static int xorshift_peak_one(mhapb_ctx *ctx, int bitsliced, double *steps_per_s)
{
    CU(ctx, ctx->eq.ensure(16));
    float best = 1e30f; double steps = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        if (bitsliced) CU(ctx, launch_xorshift_peak_bs(ctx->stream, ctx->eq.as<unsigned long long>(), &steps));
        else CU(ctx, launch_xorshift_peak(ctx->stream, ctx->eq.as<unsigned long long>(), &steps));
        cudaEventRecord(ctx->ev[1], ctx->stream);
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        if (rep > 0 && ms < best) best = ms;
    }
    *steps_per_s = steps / (best * 1e-3);
    return MHAPB_OK;
}

int mhapb_xorshift_peaks(mhapb_ctx *ctx, double *scalar_steps_per_s, double *bitsliced_steps_per_s)
{
    if (!ctx || !scalar_steps_per_s || !bitsliced_steps_per_s) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = xorshift_peak_one(ctx, 0, scalar_steps_per_s);
    if (rc) return rc;
    return xorshift_peak_one(ctx, 1, bitsliced_steps_per_s);
}

int mhapb_xorshift_peak(mhapb_ctx *ctx, double *steps_per_s)
{
    double a = 0, b = 0;
    if (!steps_per_s) return MHAPB_EINVAL;
    int rc = mhapb_xorshift_peaks(ctx, &a, &b);
    *steps_per_s = a > b ? a : b;
    return rc;
}

int mhapb_sketch_device(mhapb_ctx *ctx, const mhapb_sketch_params *p, const void *d_bases, const uint64_t *h_offsets,
                        uint32_t n_reads, int both_strands, void *d_minhash, void *d_ord, void *d_ord_n, void *d_status)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = check_sketch_params(ctx, p);
    if (rc) return rc;
    if (!h_offsets || (!d_bases && n_reads && h_offsets[n_reads] > h_offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    if (d_ord && !d_ord_n) return fail(ctx, MHAPB_EINVAL, "d_ord needs d_ord_n");
    reset_sketch_timing(ctx);
    const int per = both_strands ? 2 : 1;
    const size_t slots = (size_t)n_reads * per, H = (size_t)p->num_hashes, S = (size_t)p->ordered_sketch_size;
    std::vector<int64_t> rows(slots);
    std::vector<int32_t> status(n_reads);
    for (uint32_t r = 0; r < n_reads; r++) {
        status[r] = read_status(*p, h_offsets[r + 1] - h_offsets[r]);
        for (int s = 0; s < per; s++) rows[(size_t)r * per + s] = (int64_t)r * per + s;
    }
    if (d_minhash && slots) CU(ctx, cudaMemsetAsync(d_minhash, 0, slots * H * 4, ctx->stream));
    if (d_ord && slots) CU(ctx, cudaMemsetAsync(d_ord, 0, slots * S * 8, ctx->stream));
    if (d_ord_n && slots) CU(ctx, cudaMemsetAsync(d_ord_n, 0, slots * 4, ctx->stream));
    std::vector<uint8_t> valid(slots, 1);
    rc = sketch_core(ctx, *p, (const uint8_t *)d_bases, h_offsets, n_reads, both_strands, rows, (int32_t *)d_minhash, (int32_t *)d_ord, (int)S, (int32_t *)d_ord_n, &valid);
    if (rc) return rc;
    for (uint32_t r = 0; r < n_reads; r++)
        if (!status[r]) status[r] = !valid[(size_t)r * per] ? 1 : (per == 2 && !valid[(size_t)r * per + 1]) ? 3 : 0;
    if (d_status && n_reads) CU(ctx, cudaMemcpyAsync(d_status, status.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MHAPB_OK;
}

int mhapb_sketch(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases, const uint64_t *offsets, uint32_t n_reads,
                 int both_strands, int32_t *out_minhash, int32_t *out_ord, int32_t *out_ord_n, int32_t *out_status)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = check_sketch_params(ctx, p);
    if (rc) return rc;
    if (!offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    reset_sketch_timing(ctx);
    const int per = both_strands ? 2 : 1;
    const size_t slots = (size_t)n_reads * per, H = (size_t)p->num_hashes, S = (size_t)p->ordered_sketch_size;
    std::vector<int64_t> rows(slots);
    for (uint32_t r = 0; r < n_reads; r++) {
        int stt = read_status(*p, offsets[r + 1] - offsets[r]);
        if (out_status) out_status[r] = stt;
        for (int s = 0; s < per; s++) rows[(size_t)r * per + s] = (int64_t)r * per + s;
    }
    if (!slots) return MHAPB_OK;
    rc = h2d_bases(ctx, bases, offsets, n_reads);
    if (rc) return rc;
    const bool want_ord = out_ord || out_ord_n;
    if (out_minhash) { CU(ctx, ctx->out_minhash.ensure(slots * H * 4)); CU(ctx, cudaMemsetAsync(ctx->out_minhash.p, 0, slots * H * 4, ctx->stream)); }
    if (want_ord) {
        CU(ctx, ctx->out_ord.ensure(slots * S * 8)); CU(ctx, cudaMemsetAsync(ctx->out_ord.p, 0, slots * S * 8, ctx->stream));
        CU(ctx, ctx->out_ordn.ensure(slots * 4)); CU(ctx, cudaMemsetAsync(ctx->out_ordn.p, 0, slots * 4, ctx->stream));
    }
    std::vector<uint8_t> valid(slots, 1);
    rc = sketch_core(ctx, *p, ctx->bases.as<uint8_t>(), offsets, n_reads, both_strands, rows,
                     out_minhash ? ctx->out_minhash.as<int32_t>() : nullptr, want_ord ? ctx->out_ord.as<int32_t>() : nullptr, (int)S,
                     want_ord ? ctx->out_ordn.as<int32_t>() : nullptr, &valid);
    if (rc) return rc;
    if (out_status)
        for (uint32_t r = 0; r < n_reads; r++)
            if (!out_status[r]) out_status[r] = !valid[(size_t)r * per] ? 1 : (per == 2 && !valid[(size_t)r * per + 1]) ? 3 : 0;
    cudaEventRecord(ctx->ev[6], ctx->stream);
    if (out_minhash) CU(ctx, cudaMemcpyAsync(out_minhash, ctx->out_minhash.p, slots * H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ord) CU(ctx, cudaMemcpyAsync(out_ord, ctx->out_ord.p, slots * S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ord_n) CU(ctx, cudaMemcpyAsync(out_ord_n, ctx->out_ordn.p, slots * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[7], ctx->stream);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[6], ctx->ev[7]);
    return MHAPB_OK;
}

// ---- .dat ------------------------------------------------------------------------------------
static inline uint8_t *put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; return p + 4; }
static inline uint32_t get32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int64_t mhapb_dat_encode(int64_t id, int is_fwd, const char *header, int32_t seq_len, const int32_t *minhash, int32_t H,
                         int32_t seq_len_kmers, int32_t ok, const int32_t *ord, int32_t ord_n, uint8_t *buf)
{
    char idbuf[32];
    if (!header) { snprintf(idbuf, sizeof idbuf, "%lld", (long long)id); header = idbuf; }
    const size_t hl = strlen(header);   // ASCII headers: modified UTF-8 == the bytes
    if (hl > 65535 || H < 0 || ord_n < 0) return MHAPB_EINVAL;
    const int64_t payload = 1 + 8 + 2 + (int64_t)hl + 4 + 4 + 4 * (int64_t)H + 12 + 8 * (int64_t)ord_n;
    const int64_t total = 1 + 4 + payload;
    if (!buf) return total;
    uint8_t *p = buf;
    *p++ = is_fwd ? 1 : 0;                                  // SequenceSketchStreamer.java:352-356
    p = put32(p, (uint32_t)payload);
    *p++ = is_fwd ? 1 : 0;                                  // SequenceSketch.java:135 writeBoolean
    p = put32(p, (uint32_t)((uint64_t)id >> 32)); p = put32(p, (uint32_t)(uint64_t)id);   // writeLong
    *p++ = (uint8_t)(hl >> 8); *p++ = (uint8_t)hl; memcpy(p, header, hl); p += hl;        // writeUTF
    p = put32(p, (uint32_t)seq_len);
    p = put32(p, (uint32_t)H);                              // MinHashSketch.java:218-230
    for (int i = 0; i < H; i++) p = put32(p, (uint32_t)minhash[i]);
    p = put32(p, (uint32_t)seq_len_kmers);                  // BottomOverlapSketch.java:561-585
    p = put32(p, (uint32_t)ok);
    p = put32(p, (uint32_t)ord_n);
    for (int i = 0; i < 2 * ord_n; i++) p = put32(p, (uint32_t)ord[i]);
    return total;
}

int mhapb_dat_decode(const uint8_t *buf, uint64_t len, int64_t id_offset, uint32_t *n_records, int32_t *num_hashes,
                     int32_t *max_ord, int32_t *ordered_kmer_size, int64_t *ids, uint8_t *is_fwd, int32_t *seq_len,
                     int32_t *seq_len_kmers, int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n)
{
    if (!buf && len) return MHAPB_EINVAL;
    // pass 1: shape
    uint32_t n = 0; int32_t H = -1, mo = 0, okk = -1;
    uint64_t off = 0;
    while (off + 5 <= len) {
        const uint32_t payload = get32(buf + off + 1);
        if (off + 5 + payload > len || payload < 1 + 8 + 2 + 4 + 4 + 12) return MHAPB_EINVAL;   // truncated / corrupt
        const uint8_t *p = buf + off + 5;
        const uint32_t hl = ((uint32_t)p[9] << 8) | p[10];
        if (11 + hl + 8 > payload) return MHAPB_EINVAL;
        const int32_t h = (int32_t)get32(p + 11 + hl + 4);
        if (h < 0 || 11ull + hl + 8 + 4ull * h + 12 > payload) return MHAPB_EINVAL;
        if (H < 0) H = h; else if (H != h) return MHAPB_EINVAL;   // MinHashSearch.java:105
        const uint8_t *o = p + 11 + hl + 8 + 4ull * h;
        const int32_t kk = (int32_t)get32(o + 4), on = (int32_t)get32(o + 8);
        if (on < 0 || 11ull + hl + 8 + 4ull * h + 12 + 8ull * on != payload) return MHAPB_EINVAL;
        if (okk < 0) okk = kk; else if (okk != kk) return MHAPB_EINVAL;   // BottomOverlapSketch.java:594
        mo = std::max(mo, on);
        n++; off += 5 + payload;
    }
    if (off != len) return MHAPB_EINVAL;
    const bool sizing = !ids && !is_fwd && !seq_len && !seq_len_kmers && !minhash && !ord_hash_pos && !ord_n;
    const int32_t stride = (max_ord && !sizing && *max_ord > 0) ? *max_ord : mo;
    if (!sizing && stride < mo) return MHAPB_EINVAL;
    if (n_records) *n_records = n;
    if (num_hashes) *num_hashes = H < 0 ? 0 : H;
    if (ordered_kmer_size) *ordered_kmer_size = okk < 0 ? 0 : okk;
    if (sizing) { if (max_ord) *max_ord = mo; return MHAPB_OK; }
    off = 0;
    for (uint32_t r = 0; r < n; r++) {
        const uint32_t payload = get32(buf + off + 1);
        const uint8_t *p = buf + off + 5;
        if (is_fwd) is_fwd[r] = p[0] ? 1 : 0;
        if (ids) ids[r] = (int64_t)(((uint64_t)get32(p + 1) << 32) | get32(p + 5)) + id_offset;
        const uint32_t hl = ((uint32_t)p[9] << 8) | p[10];
        const uint8_t *q = p + 11 + hl;
        if (seq_len) seq_len[r] = (int32_t)get32(q);
        const int32_t h = (int32_t)get32(q + 4);
        if (minhash) for (int i = 0; i < h; i++) minhash[(size_t)r * h + i] = (int32_t)get32(q + 8 + 4ull * i);
        const uint8_t *o = q + 8 + 4ull * h;
        if (seq_len_kmers) seq_len_kmers[r] = (int32_t)get32(o);
        const int32_t on = (int32_t)get32(o + 8);
        if (ord_n) ord_n[r] = on;
        if (ord_hash_pos) {
            int32_t *dst = ord_hash_pos + (size_t)r * stride * 2;
            for (int i = 0; i < 2 * on; i++) dst[i] = (int32_t)get32(o + 12 + 4ull * i);
            for (int i = 2 * on; i < 2 * stride; i++) dst[i] = 0;
        }
        off += 5 + payload;
    }
    return MHAPB_OK;
}

int mhapb_sketch_to_dat(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases, const uint64_t *offsets,
                        const int64_t *ids, uint32_t n_reads, int both_strands, uint8_t **out, uint64_t *out_len, uint32_t *n_records)
{
    return mhapb_sketch_to_dat_named(ctx, p, bases, offsets, ids, nullptr, n_reads, both_strands, out, out_len, n_records);
}

int mhapb_sketch_to_dat_named(mhapb_ctx *ctx, const mhapb_sketch_params *p, const char *bases, const uint64_t *offsets,
                              const int64_t *ids, const char *const *headers, uint32_t n_reads, int both_strands,
                              uint8_t **out, uint64_t *out_len, uint32_t *n_records)
{
    if (!ctx || !out || !out_len) return MHAPB_EINVAL;
    int rc = check_sketch_params(ctx, p);
    if (rc) return rc;
    const int per = both_strands ? 2 : 1;
    const size_t slots = (size_t)n_reads * per, H = (size_t)p->num_hashes, S = (size_t)p->ordered_sketch_size;
    std::vector<int32_t> mh(slots * H), ord(slots * S * 2), on(slots), status(n_reads);
    rc = mhapb_sketch(ctx, p, bases, offsets, n_reads, both_strands, mh.data(), ord.data(), on.data(), status.data());
    if (rc) return rc;
    uint64_t total = 0; uint32_t nrec = 0;
    for (uint32_t r = 0; r < n_reads; r++) {
        if (status[r] && status[r] != 3) continue;
        for (int s = 0; s < (status[r] == 3 ? 1 : per); s++) {
            size_t j = (size_t)r * per + s;
            int64_t id = ids ? ids[r] : (int64_t)r + 1;
            const int64_t sz = mhapb_dat_encode(id, s == 0, headers ? headers[r] : nullptr, 0, nullptr, (int32_t)H, 0, 0, nullptr, on[j], nullptr);
            if (sz < 0) { return fail(ctx, MHAPB_EINVAL, "header of read %u is longer than 65535 bytes", r); }
            total += (uint64_t)sz;
            nrec++;
        }
    }
    uint8_t *buf = (uint8_t *)malloc(total ? total : 1);
    if (!buf) return fail(ctx, MHAPB_ENOMEM, "malloc .dat buffer");
    uint8_t *w = buf;
    for (uint32_t r = 0; r < n_reads; r++) {
        if (status[r] && status[r] != 3) continue;
        const int32_t len = (int32_t)(offsets[r + 1] - offsets[r]);
        for (int s = 0; s < (status[r] == 3 ? 1 : per); s++) {
            size_t j = (size_t)r * per + s;
            int64_t id = ids ? ids[r] : (int64_t)r + 1;
            w += mhapb_dat_encode(id, s == 0, headers ? headers[r] : nullptr, len, mh.data() + j * H, (int32_t)H, len - p->ordered_kmer_size + 1,
                                  p->ordered_kmer_size, ord.data() + j * S * 2, on[j], w);
        }
    }
    *out = buf; *out_len = total;
    if (n_records) *n_records = nrec;
    return MHAPB_OK;
}

// ---- store -------------------------------------------------------------------------------------
int mhapb_store_reset(mhapb_ctx *ctx, const mhapb_sketch_params *p)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return store_configure(ctx, p);
}

int mhapb_store_add_reads(mhapb_ctx *ctx, const char *bases, const uint64_t *offsets, const int64_t *ids, uint32_t n_reads,
                          int both_strands, int64_t *n_added)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    if (!offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    reset_sketch_timing(ctx);
    return add_reads_locked(ctx, bases, offsets, ids, n_reads, both_strands, n_added);
}

// the reference's own checks on sketches that did not come from this context's parameters (.dat records):
// MinHashSearch.java:105-106 / :157-159 (number of hashes) and BottomOverlapSketch.java:594-595 (ordered k-mer size)
static int check_sketch_shape(mhapb_ctx *ctx, int32_t num_hashes, int32_t ordered_kmer_size, bool query)
{
    const Store &s = ctx->store;
    if (num_hashes != s.p.num_hashes)
        return query ? fail(ctx, MHAPB_EINVAL, "Number of hashes does not match. Stored size %d, input size %d.", s.p.num_hashes, num_hashes)
                     : fail(ctx, MHAPB_EINVAL, "Number of MinHashes of the sequence does not match current settings.");
    if (ordered_kmer_size != s.p.ordered_kmer_size)
        return fail(ctx, MHAPB_EINVAL, "Sketch k-mer size does not match between the two sequences.");
    return MHAPB_OK;
}

int mhapb_store_add_reads_device(mhapb_ctx *ctx, const void *d_bases, const uint64_t *h_offsets, const int64_t *ids, uint32_t n_reads,
                                 int both_strands, int64_t *n_added)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    if (!h_offsets || (!d_bases && n_reads && h_offsets[n_reads] > h_offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    reset_sketch_timing(ctx);
    return add_reads_locked(ctx, nullptr, h_offsets, ids, n_reads, both_strands, n_added, (const uint8_t *)d_bases);
}

static int add_sketches_common(mhapb_ctx *ctx, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len,
                               const int32_t *seq_len_kmers, const void *minhash, const void *ord, const int32_t *ord_n,
                               int32_t ord_stride, uint32_t n, cudaMemcpyKind kind)
{
    Store &s = ctx->store;
    if (!s.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    if (!ids || !is_fwd || !seq_len || !seq_len_kmers || !minhash || !ord || !ord_n) return fail(ctx, MHAPB_EINVAL, "null sketch column");
    if (ord_stride < 1) return fail(ctx, MHAPB_EINVAL, "ord_stride %d", ord_stride);
    if (!n) return MHAPB_OK;
    int32_t max_on = 0;
    for (uint32_t i = 0; i < n; i++) { if (ord_n[i] < 0 || ord_n[i] > ord_stride) return fail(ctx, MHAPB_EINVAL, "ord_n[%u]=%d exceeds stride %d", i, ord_n[i], ord_stride); max_on = std::max(max_on, ord_n[i]); }
    if (max_on > s.ord_stride) {
        if (s.n) return fail(ctx, MHAPB_EINVAL, "ordered sketch of %d entries exceeds the store's stride %d", max_on, s.ord_stride);
        s.ord_stride = max_on;
    }
    const size_t meta0 = s.h_id.size();
    for (uint32_t i = 0; i < n; i++) {
        int rc = store_push_meta(ctx, ids[i], is_fwd[i] != 0, seq_len[i], seq_len_kmers[i], ord_n[i]);
        if (rc) { store_rollback_meta(s, meta0); return rc; }
    }
    const int64_t n0 = s.n;
    auto device_part = [&]() -> int {
        int rc = store_reserve(ctx, n);
        if (rc) return rc;
        const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.ord_stride;
        CU(ctx, cudaMemcpyAsync(s.minhash.as<int32_t>() + (size_t)n0 * H, minhash, (size_t)n * H * 4, kind, ctx->stream));
        if ((size_t)ord_stride == S) CU(ctx, cudaMemcpyAsync(s.ord.as<int32_t>() + (size_t)n0 * S * 2, ord, (size_t)n * S * 8, kind, ctx->stream));
        else {
            CU(ctx, cudaMemsetAsync(s.ord.as<int32_t>() + (size_t)n0 * S * 2, 0, (size_t)n * S * 8, ctx->stream));
            CU(ctx, cudaMemcpy2DAsync(s.ord.as<int32_t>() + (size_t)n0 * S * 2, S * 8, ord, (size_t)ord_stride * 8, std::min(S, (size_t)ord_stride) * 8, n, kind, ctx->stream));
        }
        CU(ctx, cudaMemcpyAsync(s.ord_n.as<int32_t>() + n0, ord_n, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        s.n += n;
        s.indexed = false; s.fwd_list_valid = false;
        rc = store_sync_columns(ctx, n0);
        if (rc) s.n = n0;
        return rc;
    };
    const int rc = device_part();
    if (rc) store_rollback_meta(s, meta0);
    return rc;
}

int mhapb_store_add_sketches(mhapb_ctx *ctx, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len,
                             const int32_t *seq_len_kmers, const int32_t *minhash, int32_t num_hashes, const int32_t *ord_hash_pos,
                             const int32_t *ord_n, int32_t ord_stride, int32_t ordered_kmer_size, uint32_t n)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    if (!ctx->store.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    if (n) { int rc = check_sketch_shape(ctx, num_hashes, ordered_kmer_size, false); if (rc) return rc; }
    return add_sketches_common(ctx, ids, is_fwd, seq_len, seq_len_kmers, minhash, ord_hash_pos, ord_n, ord_stride, n, cudaMemcpyHostToDevice);
}

int mhapb_store_add_sketches_device(mhapb_ctx *ctx, const int64_t *ids, const uint8_t *is_fwd, const int32_t *seq_len,
                                    const int32_t *seq_len_kmers, const void *d_minhash, const void *d_ord, const int32_t *ord_n, uint32_t n)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return add_sketches_common(ctx, ids, is_fwd, seq_len, seq_len_kmers, d_minhash, d_ord, ord_n, ctx->store.p.ordered_sketch_size, n, cudaMemcpyDeviceToDevice);
}

int mhapb_sketch_reserve(mhapb_ctx *ctx, const mhapb_sketch_params *p, uint64_t max_bases, uint32_t max_reads, int both_strands)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = check_sketch_params(ctx, p);
    if (rc) return rc;
    const uint64_t per = both_strands ? 2 : 1;
    const uint64_t kmers = std::min<uint64_t>(max_bases * per, 256ull << 20);   // sketch_core's chunk cap
    CU(ctx, ctx->bases.ensure((size_t)max_bases + 64));
    CU(ctx, ctx->keys.ensure((size_t)kmers * 8));
    CU(ctx, ctx->wts.ensure((size_t)kmers * 4));
    CU(ctx, ctx->desc.ensure((size_t)max_reads * per * sizeof(StrandDesc)));
    CU(ctx, ctx->nlight.ensure((size_t)max_reads * per * 4));
    CU(ctx, ctx->nheavy.ensure((size_t)max_reads * per * 4));
    return MHAPB_OK;
}

int mhapb_store_reserve(mhapb_ctx *ctx, int64_t n_sketches)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (!s.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    if (n_sketches <= s.n) return MHAPB_OK;
    return store_reserve(ctx, n_sketches - s.n);
}

int64_t mhapb_store_size(mhapb_ctx *ctx)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return ctx->store.n;
}

int mhapb_store_get(mhapb_ctx *ctx, int64_t idx, int64_t *id, int32_t *is_fwd, int32_t *seq_len, int32_t *seq_len_kmers,
                    int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (idx < 0 || idx >= s.n) return fail(ctx, MHAPB_EINVAL, "store index %lld out of range", (long long)idx);
    if (id) *id = s.h_id[idx];
    if (is_fwd) *is_fwd = s.h_fwd[idx];
    if (seq_len) *seq_len = s.h_len[idx];
    if (seq_len_kmers) *seq_len_kmers = s.h_lenk[idx];
    if (ord_n) *ord_n = s.h_ordn[idx];
    const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.ord_stride;
    if (minhash) CU(ctx, cudaMemcpyAsync(minhash, s.minhash.as<int32_t>() + (size_t)idx * H, H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (ord_hash_pos) CU(ctx, cudaMemcpyAsync(ord_hash_pos, s.ord.as<int32_t>() + (size_t)idx * S * 2, (size_t)s.h_ordn[idx] * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MHAPB_OK;
}

int mhapb_store_params(mhapb_ctx *ctx, mhapb_sketch_params *out)
{
    if (!ctx || !out) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->store.configured) return fail(ctx, MHAPB_ESTATE, "mhapb_store_reset must be called first");
    *out = ctx->store.p;
    return MHAPB_OK;
}

int mhapb_store_get_range(mhapb_ctx *ctx, int64_t first, int64_t count, int64_t *ids, uint8_t *is_fwd, int32_t *seq_len,
                          int32_t *seq_len_kmers, int32_t *minhash, int32_t *ord_hash_pos, int32_t *ord_n)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (first < 0 || count < 0 || first + count > s.n) return fail(ctx, MHAPB_EINVAL, "store range [%lld, %lld) out of range", (long long)first, (long long)(first + count));
    if (!count) return MHAPB_OK;
    const size_t H = (size_t)s.p.num_hashes, S = (size_t)s.ord_stride, c = (size_t)count;
    if (ids) memcpy(ids, s.h_id.data() + first, c * 8);
    if (is_fwd) memcpy(is_fwd, s.h_fwd.data() + first, c);
    if (seq_len) memcpy(seq_len, s.h_len.data() + first, c * 4);
    if (seq_len_kmers) memcpy(seq_len_kmers, s.h_lenk.data() + first, c * 4);
    if (ord_n) memcpy(ord_n, s.h_ordn.data() + first, c * 4);
    if (minhash) CU(ctx, cudaMemcpyAsync(minhash, s.minhash.as<int32_t>() + (size_t)first * H, c * H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (ord_hash_pos) CU(ctx, cudaMemcpyAsync(ord_hash_pos, s.ord.as<int32_t>() + (size_t)first * S * 2, c * S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MHAPB_OK;
}

int mhapb_store_device_ptrs(mhapb_ctx *ctx, void **d_minhash, void **d_ord, void **d_ord_n, int64_t *n, int32_t *num_hashes, int32_t *ord_stride)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    Store &s = ctx->store;
    if (d_minhash) *d_minhash = s.minhash.p;
    if (d_ord) *d_ord = s.ord.p;
    if (d_ord_n) *d_ord_n = s.ord_n.p;
    if (n) *n = s.n;
    if (num_hashes) *num_hashes = s.p.num_hashes;
    if (ord_stride) *ord_stride = s.ord_stride;
    return MHAPB_OK;
}

int mhapb_index_build(mhapb_ctx *ctx)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return index_build(ctx);
}

// ---- search ------------------------------------------------------------------------------------
int mhapb_search_self(mhapb_ctx *ctx, const mhapb_search_params *sp, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (!s.configured || s.n == 0) return fail(ctx, MHAPB_ESTATE, "search on an empty store");
    ctx->store.ord_n.as<int32_t>();
    // ord_n column may have been produced on the device (add_reads): the host mirror is exact by construction
    QuerySet q{};
    q.d_minhash = s.minhash.as<int32_t>(); q.d_ord = s.ord.as<int32_t>(); q.d_ordn = s.ord_n.as<int32_t>();
    q.d_lenk = s.lenk.as<int32_t>(); q.d_len = s.len.as<int32_t>(); q.d_id = s.id.as<int64_t>(); q.ord_stride = s.ord_stride;
    q.h_id = s.h_id.data(); q.h_fwd = s.h_fwd.data(); q.h_len = s.h_len.data();
    int64_t first = std::max<int64_t>(0, sp->query_first);
    int64_t last = sp->query_count < 0 ? s.n : std::min<int64_t>(s.n, first + sp->query_count);
    if (first == 0 && last == s.n) {   // every forward sketch queries: the list lives on the device, rebuilt only when the store changed
        if (!s.fwd_list_valid) {
            std::vector<uint32_t> l;
            for (int64_t i = 0; i < s.n; i++) if (s.h_fwd[i]) l.push_back((uint32_t)i);   // AbstractMatchSearch.java:128-129
            CU(ctx, s.fwd_list.ensure(l.size() * 4 + 4));
            CU(ctx, cudaMemcpyAsync(s.fwd_list.p, l.data(), l.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            s.fwd_list_n = (int64_t)l.size(); s.fwd_list_valid = true;
        }
        q.d_list = s.fwd_list.as<uint32_t>(); q.n_list = s.fwd_list_n;
    } else {
        for (int64_t i = first; i < last; i++) if (s.h_fwd[i]) q.list.push_back((uint32_t)i);
    }
    return search_core(ctx, sp, q, 1, out, n_out, stats);
}

static int search_query_sketches_locked(mhapb_ctx *ctx, const mhapb_search_params *sp, const int64_t *ids, const uint8_t *is_fwd,
                                        const int32_t *seq_len, const int32_t *seq_len_kmers, const int32_t *d_minhash,
                                        const int32_t *d_ord, const int32_t *d_ordn, int32_t ord_stride, uint32_t n,
                                        mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats, int to_self = 0)
{
    CU(ctx, ctx->q_lenk.ensure((size_t)n * 4 + 4));
    CU(ctx, ctx->q_len.ensure((size_t)n * 4 + 4));
    CU(ctx, ctx->q_id.ensure((size_t)n * 8 + 8));
    if (n) {
        CU(ctx, cudaMemcpyAsync(ctx->q_lenk.p, seq_len_kmers, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->q_len.p, seq_len, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->q_id.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    QuerySet q{};
    q.d_minhash = d_minhash; q.d_ord = d_ord; q.d_ordn = d_ordn;
    q.d_lenk = ctx->q_lenk.as<int32_t>(); q.d_len = ctx->q_len.as<int32_t>(); q.d_id = ctx->q_id.as<int64_t>(); q.ord_stride = ord_stride;
    q.h_id = ids; q.h_fwd = is_fwd; q.h_len = seq_len;
    bool all_fwd = true;
    for (uint32_t i = 0; i < n && all_fwd; i++) all_fwd = is_fwd[i] != 0;
    if (all_fwd) { q.list_all = true; q.n_all = n; }
    else for (uint32_t i = 0; i < n; i++) if (is_fwd[i]) q.list.push_back(i);   // AbstractMatchSearch.java:225 dequeue(true)
    return search_core(ctx, sp, q, to_self, out, n_out, stats);
}

int mhapb_search_query_sketches(mhapb_ctx *ctx, const mhapb_search_params *sp, const int64_t *ids, const uint8_t *is_fwd,
                                const int32_t *seq_len, const int32_t *seq_len_kmers, const int32_t *minhash, int32_t num_hashes,
                                const int32_t *ord_hash_pos, const int32_t *ord_n, int32_t ord_stride, int32_t ordered_kmer_size,
                                uint32_t n, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (!s.configured || s.n == 0) return fail(ctx, MHAPB_ESTATE, "search on an empty store");
    if (n && (!ids || !is_fwd || !seq_len || !seq_len_kmers || !minhash || !ord_hash_pos || !ord_n)) return fail(ctx, MHAPB_EINVAL, "null query column");
    if (ord_stride < 1) return fail(ctx, MHAPB_EINVAL, "ord_stride %d", ord_stride);
    if (n) { int rc = check_sketch_shape(ctx, num_hashes, ordered_kmer_size, true); if (rc) return rc; }
    const size_t H = (size_t)s.p.num_hashes;
    CU(ctx, ctx->q_minhash.ensure((size_t)n * H * 4 + 4));
    CU(ctx, ctx->q_ord.ensure((size_t)n * ord_stride * 8 + 8));
    CU(ctx, ctx->q_ordn.ensure((size_t)n * 4 + 4));
    if (n) {
        CU(ctx, cudaMemcpyAsync(ctx->q_minhash.p, minhash, (size_t)n * H * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->q_ord.p, ord_hash_pos, (size_t)n * ord_stride * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->q_ordn.p, ord_n, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return search_query_sketches_locked(ctx, sp, ids, is_fwd, seq_len, seq_len_kmers, ctx->q_minhash.as<int32_t>(),
                                        ctx->q_ord.as<int32_t>(), ctx->q_ordn.as<int32_t>(), ord_stride, n, out, n_out, stats);
}

int mhapb_search_sketches_device(mhapb_ctx *ctx, const mhapb_search_params *sp, int to_self, const int64_t *ids, const uint8_t *is_fwd,
                                 const int32_t *seq_len, const int32_t *seq_len_kmers, const void *d_minhash, const void *d_ord,
                                 const void *d_ord_n, int32_t ord_stride, uint32_t n, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (!s.configured || s.n == 0) return fail(ctx, MHAPB_ESTATE, "search on an empty store");
    if (n && (!ids || !is_fwd || !seq_len || !seq_len_kmers || !d_minhash || !d_ord || !d_ord_n)) return fail(ctx, MHAPB_EINVAL, "null query column");
    return search_query_sketches_locked(ctx, sp, ids, is_fwd, seq_len, seq_len_kmers, (const int32_t *)d_minhash, (const int32_t *)d_ord,
                                        (const int32_t *)d_ord_n, ord_stride, n, out, n_out, stats, to_self ? 1 : 0);
}

int mhapb_search_query_reads(mhapb_ctx *ctx, const mhapb_search_params *sp, const char *bases, const uint64_t *offsets,
                             const int64_t *ids, uint32_t n_reads, mhapb_hit **out, uint64_t *n_out, mhapb_stats *stats)
{
    if (!ctx || !sp) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (!s.configured || s.n == 0) return fail(ctx, MHAPB_ESTATE, "search on an empty store");
    if (!offsets || (!bases && n_reads && offsets[n_reads] > offsets[0])) return fail(ctx, MHAPB_EINVAL, "null bases/offsets");
    reset_sketch_timing(ctx);
    std::vector<int64_t> qid; std::vector<int32_t> qlen, qlenk;
    int rc = sketch_query_reads(ctx, bases, offsets, ids, n_reads, &qid, &qlen, &qlenk, nullptr);
    if (rc) return rc;
    const std::vector<uint8_t> qfwd(qid.size(), 1);
    return search_query_sketches_locked(ctx, sp, qid.data(), qfwd.data(), qlen.data(), qlenk.data(), ctx->q_minhash.as<int32_t>(),
                                        ctx->q_ord.as<int32_t>(), ctx->q_ordn.as<int32_t>(), s.p.ordered_sketch_size, (uint32_t)qid.size(), out, n_out, stats);
}

int mhapb_format_match(const mhapb_hit *h, char *buf, size_t buflen)
{
    if (!h || !buf) return MHAPB_EINVAL;
    // impl/MatchResult.java:54-57 strand flip, :61-64 clamp, :98-113 format
    const int32_t a1 = h->from_fwd ? h->a1 : h->from_len - h->a2 - 1;
    const int32_t a2 = h->from_fwd ? h->a2 : h->from_len - h->a1 - 1;
    const int32_t b1 = h->to_fwd ? h->b1 : h->to_len - h->b2 - 1;
    const int32_t b2 = h->to_fwd ? h->b2 : h->to_len - h->b1 - 1;
    const double score = h->score > 1.0 ? 1.0 : h->score;
    return snprintf(buf, buflen, "%lld %lld %.6f %.6f %d %d %d %d %d %d %d %d", (long long)h->from_id, (long long)h->to_id,
                    1.0 - score, (double)h->valid_count, h->from_fwd ? 0 : 1, a1, a2, h->from_len, h->to_fwd ? 0 : 1, b1, b2, h->to_len);
}

int mhapb_minhash_equal_count(mhapb_ctx *ctx, int64_t i, int64_t j, int32_t *out_equal)
{
    if (!ctx || !out_equal) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    Store &s = ctx->store;
    if (i < 0 || j < 0 || i >= s.n || j >= s.n) return fail(ctx, MHAPB_EINVAL, "store index out of range");
    CU(ctx, ctx->eq.ensure(16));
    const size_t H = (size_t)s.p.num_hashes;
    int launches = 0;
    CU(ctx, launch_equal_count(ctx->stream, s.minhash.as<int32_t>() + (size_t)i * H, s.minhash.as<int32_t>() + (size_t)j * H, (int)H, ctx->eq.as<int32_t>(), &launches));
    CU(ctx, cudaMemcpyAsync(out_equal, ctx->eq.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.kernel_launches += launches;
    return MHAPB_OK;
}

} // extern "C"

// ---- the -f k-mer filter ----------------------------------------------------------------------
namespace {

// HashUtils.computeSequenceHashesLong(kmer, kmer.length(), 0, doReverseCompliment)[0] on the host
uint64_t host_kmer_hash(const char *kmer, int len, int canonical)
{
    std::string fwd(kmer, (size_t)len);
    const std::string *use = &fwd;
    std::string rc;
    if (canonical) {                                   // HashUtils.java:246-251: the smaller of k-mer and Utils.rc(k-mer)
        rc.resize((size_t)len);
        for (int i = 0; i < len; i++) rc[(size_t)i] = (char)complement_char(upper_char((uint8_t)kmer[len - 1 - i]));
        if (std::lexicographical_compare(rc.begin(), rc.end(), fwd.begin(), fwd.end(),
                                         [](char a, char b) { return (unsigned char)a < (unsigned char)b; })) use = &rc;
    }
    const std::string &u = *use;
    return murmur3_128_h1_chars([&](int j) { return (uint8_t)u[(size_t)j]; }, len);
}

// Installs the filter: scaledIdf per repeat k-mer in double precision (FrequencyCounts.java:223-229,250-254,285-309),
// an open-addressed device map, and the Bloom bit array.
int filter_install(mhapb_ctx *ctx, const mhapb_filter_params *p, const int64_t *hashes, const double *fractions, uint64_t n,
                   const uint64_t *bloom_words, uint64_t bloom_bits, int32_t bloom_nfun, int64_t *n_kept = nullptr)
{
    if (!p) return fail(ctx, MHAPB_EINVAL, "null filter params");
    if (p->supress_noise < 0 || p->supress_noise > 2) return fail(ctx, MHAPB_EINVAL, "The --supress-noise parameter must be in [0,2].");
    if (p->idf_scale < 1.0) return fail(ctx, MHAPB_EINVAL, "--repeat-idf-scale must be >= 1");
    if (n && (!hashes || !fractions)) return fail(ctx, MHAPB_EINVAL, "null filter arrays");
    if (p->supress_noise > 0 && (!bloom_words || bloom_bits == 0 || (bloom_bits & 63) || bloom_nfun < 1))
        return fail(ctx, MHAPB_EINVAL, "--supress-noise %d needs the Bloom filter bit array", p->supress_noise);
    const double rw = p->repeat_weight;
    const double offset = (rw >= 0.0 && rw < 1.0) ? rw : 0.0;                       // MhapMain.java:348-350
    // fractionCounts: entries at or above the cutoff, a repeated k-mer keeps its last value (Map.put)
    std::unordered_map<int64_t, double> kept;
    double max_value = -INFINITY;
    for (uint64_t i = 0; i < n; i++)
        if (fractions[i] >= p->filter_cutoff) { kept[hashes[i]] = fractions[i]; if (fractions[i] > max_value) max_value = fractions[i]; }
    if (n_kept) *n_kept = (int64_t)kept.size();
    const double min_idf = log(max_value / max_value - offset);                    // idf(maxValue) :228
    const double max_idf = log(max_value / p->filter_cutoff - offset);              // idf(minValue) :229
    const double scale = (max_idf - min_idf) / (p->idf_scale - 1.0);
    size_t cap = 64;
    while (cap < kept.size() * 2) cap <<= 1;
    std::vector<uint64_t> keys(cap, 0);
    std::vector<double> idf(cap, 0.0);
    std::vector<uint32_t> used(cap / 32, 0u);
    for (const auto &kv : kept) {
        const double v = 1.0 + (log(max_value / kv.second - offset) - min_idf) / scale;   // scaledIdf :300-308
        size_t q = (size_t)fmix64((uint64_t)kv.first) & (cap - 1);
        while ((used[q >> 5] >> (q & 31)) & 1u) q = (q + 1) & (cap - 1);
        used[q >> 5] |= 1u << (q & 31); keys[q] = (uint64_t)kv.first; idf[q] = v;
    }
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    KmerFilterView v{};
    if (!kept.empty()) {
        CU(ctx, ctx->f_keys.ensure(cap * 8)); CU(ctx, ctx->f_idf.ensure(cap * 8)); CU(ctx, ctx->f_used.ensure(cap / 8));
        CU(ctx, cudaMemcpy(ctx->f_keys.p, keys.data(), cap * 8, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->f_idf.p, idf.data(), cap * 8, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->f_used.p, used.data(), cap / 8, cudaMemcpyHostToDevice));
        v.map_keys = ctx->f_keys.as<uint64_t>(); v.map_idf = ctx->f_idf.as<double>(); v.map_used = ctx->f_used.as<uint32_t>();
        v.map_mask = (uint32_t)(cap - 1);
    }
    if (p->supress_noise > 0) {
        CU(ctx, ctx->f_bloom.ensure(bloom_bits / 8));
        CU(ctx, cudaMemcpy(ctx->f_bloom.p, bloom_words, bloom_bits / 8, cudaMemcpyHostToDevice));
        v.bloom = ctx->f_bloom.as<uint64_t>(); v.bloom_bits = bloom_bits; v.bloom_nfun = bloom_nfun;
    }
    v.mode = rw < 0.0 ? 1 : (rw < 1.0 ? 2 : 3);
    v.remove_unique = p->supress_noise;
    v.no_tf = p->no_tf ? 1 : 0;
    v.range = p->idf_scale;
    v.light_weight = 1;
    if (v.mode == 2) {                                 // weight of a once-seen k-mer outside the repeat map: round(1 * range)
        const double r = floor(p->idf_scale + 0.5);
        v.light_weight = r >= 1.0 ? (r > 2147483647.0 ? 2147483647u : (uint32_t)r) : 1u;
    }
    ctx->filter = v; ctx->filter_params = *p; ctx->filter_set = true;
    return MHAPB_OK;
}

} // namespace

extern "C" {

int mhapb_kmer_hash(const char *kmer, int32_t len, int canonical, int64_t *out_hash)
{
    if (!kmer || len < 1 || !out_hash) return MHAPB_EINVAL;
    *out_hash = (int64_t)host_kmer_hash(kmer, len, canonical);
    return MHAPB_OK;
}

int mhapb_filter_set(mhapb_ctx *ctx, const mhapb_filter_params *p, const int64_t *hashes, const double *fractions, uint64_t n,
                     const uint64_t *bloom_words, uint64_t bloom_bits, int32_t bloom_num_hash_functions)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return filter_install(ctx, p, hashes, fractions, n, bloom_words, bloom_bits, bloom_num_hash_functions);
}

int mhapb_filter_clear(mhapb_ctx *ctx)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->filter = KmerFilterView{}; ctx->filter_set = false;
    return MHAPB_OK;
}

int mhapb_filter_load_text(mhapb_ctx *ctx, const mhapb_filter_params *p, const char *text, uint64_t len, int canonical, int64_t *n_repeat)
{
    if (!ctx) return MHAPB_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!p || (!text && len)) return fail(ctx, MHAPB_EINVAL, "null filter text/params");
    if (p->supress_noise < 0 || p->supress_noise > 2) return fail(ctx, MHAPB_EINVAL, "The --supress-noise parameter must be in [0,2].");
    const char *cur = text, *end = text + len;
    auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; };
    // first line: "<sizeBloom> <sizeRepeat>" (FrequencyCounts.java:91-117)
    long long size_bloom = 1;
    {
        const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        if (le > cur) {
            std::string first(cur, le);
            long long a = 0, b = 0;
            if (sscanf(first.c_str(), "%lld %lld", &a, &b) < 2 || a < 0 || b < 0)
                return fail(ctx, MHAPB_EINVAL, "K-mer filter file first line must contain estimated number of k-mers in the file (long).");
            size_bloom = a ? a : 1;
        }
        cur = nl ? nl + 1 : end;
    }
    std::vector<uint64_t> bloom;
    uint64_t bloom_bits = 0; int32_t nfun = 0;
    if (p->supress_noise > 0) {
        // Guava BloomFilter.create(funnel, expectedInsertions, 1e-5): optimalNumOfBits, optimalNumOfHashFunctions, BitArray
        const double fpp = 1.0e-5;
        const long long num_bits = (long long)(-(double)size_bloom * log(fpp) / (log(2.0) * log(2.0)));
        const int nh = (int)floor((double)num_bits / (double)size_bloom * log(2.0) + 0.5);
        nfun = nh < 1 ? 1 : nh;
        long long words = (num_bits + 63) / 64;
        if (words < 1) words = 1;
        bloom.assign((size_t)words, 0ull);
        bloom_bits = (uint64_t)words * 64;
    }
    std::vector<int64_t> hashes; std::vector<double> fractions;
    while (cur < end) {
        const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        const char *q = cur;
        while (q < le && !is_ws(*q)) q++;
        const int klen = (int)(q - cur);
        if (klen >= 1) {
            const uint64_t h = host_kmer_hash(cur, klen, canonical);
            while (q < le && is_ws(*q)) q++;
            bool skip = false;
            if (q < le) {                               // second column: the fraction (:178-193)
                const char *t = q;
                while (q < le && !is_ws(*q)) q++;
                std::string num(t, q);
                char *ep = nullptr;
                const double pct = strtod(num.c_str(), &ep);
                if (ep == num.c_str() || *ep != 0) skip = true;     // NumberFormatException: the line is dropped (:204-207)
                else if (pct >= p->filter_cutoff) { hashes.push_back((int64_t)h); fractions.push_back(pct); }
            }
            if (!skip && p->supress_noise > 0) {        // validMers.put(hash) (:196-200)
                uint64_t h1, h2; bloom_hash_pair(h, &h1, &h2);
                uint64_t c = h1;
                for (int i = 0; i < nfun; i++) { const uint64_t bit = (c & 0x7fffffffffffffffULL) % bloom_bits; bloom[bit >> 6] |= 1ull << (bit & 63); c += h2; }
            }
        }
        cur = nl ? nl + 1 : end;
    }
    return filter_install(ctx, p, hashes.data(), fractions.data(), hashes.size(), bloom.empty() ? nullptr : bloom.data(), bloom_bits, nfun, n_repeat);
}

} // extern "C"
