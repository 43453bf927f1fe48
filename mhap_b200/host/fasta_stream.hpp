// fasta_stream.hpp -- streaming FASTA producer for the host driver.
//
// Replaces impl/FastaData.java:125-204 (enqueueNextSequenceInFile) as a pipeline: at >= 1 Gbases/s of K1 a
// single-threaded getline parser is the wall-clock bottleneck, so
//   reader thread  : reads the file (plain or .gz through zlib) in large text chunks cut at record starts ("\n>"),
//   parser threads : turn a chunk into a batch -- sequence lines concatenated into a PINNED buffer
//                    (mhapb_host_alloc, so the H2D copy of mhapb_store_add_reads is a straight DMA) + offsets,
//   consumer       : takes batches in file order (ids are file positions, FastaData.java:181,190-192) and calls the
//                    library, so parsing batch i+1 overlaps H2D + K1 of batch i.
// Semantics kept from the reference: records start with '>', the header text is ignored (numeric ids), sequence lines
// are concatenated, '\r' is dropped, a first line that is not a header is "Next sequence does not start with >. Invalid
// format.", and an EMPTY record ends the file (enqueueNextSequenceInFile returns false on a zero-length sequence).
// Upper-casing (FastaData.java:194) happens on the GPU.  .bz2 goes through libbz2's high-level API, bound with dlopen
// (the image ships libbz2.so.1.0 but no bzlib.h; the three prototypes used are part of its stable C ABI).
#pragma once
#include "../../include/mhap_b200.h"

#include <dlfcn.h>
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mhapb_host {

struct FastaBatch {
    char *bases = nullptr;            // pinned, cap bytes
    size_t cap = 0, len = 0;
    std::vector<uint64_t> offsets;    // n+1
    std::vector<std::string> headers; // first token of every header line, only when the stream was asked to keep them
    bool ended = false;               // an empty record was met: nothing after this batch is read
    std::string error;                // non-empty: the reference's exception text
    uint64_t seq = 0;
    char *text = nullptr;             // the chunk this batch is parsed from (owned here so pools stay paired)
    size_t text_cap = 0, text_len = 0;
    double t_read = 0, t_parse = 0, t_alloc = 0;   // seconds, for MHAPB_FASTA_TRACE
    bool first_chunk = false;
    uint32_t n_reads() const { return offsets.empty() ? 0u : (uint32_t)(offsets.size() - 1); }
};

class FastaStream {
public:
    FastaStream(const std::string &path, int parser_threads, size_t chunk_bytes, bool keep_headers = false)
        : path_(path), chunk_(chunk_bytes < (1u << 16) ? (1u << 16) : chunk_bytes), keep_headers_(keep_headers)
    {
        const size_t n = path.size();
        gz_ = n > 3 && path.compare(n - 3, 3, ".gz") == 0;
        bz_ = n > 4 && path.compare(n - 4, 4, ".bz2") == 0;
        if (gz_) { gzf_ = gzopen(path.c_str(), "rb"); if (gzf_) gzbuffer(gzf_, 1u << 20); }
        else if (bz_) {
            // BZFILE *BZ2_bzopen(const char *path, const char *mode); int BZ2_bzread(BZFILE *, void *, int); void BZ2_bzclose(BZFILE *)
            for (const char *lib : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) if ((bzlib_ = dlopen(lib, RTLD_NOW))) break;
            if (bzlib_) {
                bz_open_ = (void *(*)(const char *, const char *))dlsym(bzlib_, "BZ2_bzopen");
                bz_read_ = (int (*)(void *, void *, int))dlsym(bzlib_, "BZ2_bzread");
                bz_close_ = (void (*)(void *))dlsym(bzlib_, "BZ2_bzclose");
                if (bz_open_ && bz_read_ && bz_close_) bzf_ = bz_open_(path.c_str(), "rb");
            }
        }
        else fp_ = fopen(path.c_str(), "rb");
        if (!gzf_ && !fp_ && !bzf_) { open_failed_ = true; return; }
        if (parser_threads < 1) parser_threads = 1;
        // batches in flight: one being read, one per parser, one with the consumer -- but no more than 4, so that the
        // pinned buffers are reused instead of page-locking the whole file once
        const int pool = std::min(parser_threads + 2, 4);
        for (int i = 0; i < pool; i++) free_.push_back(new FastaBatch());
        reader_ = std::thread([this] { read_loop(); });
        for (int i = 0; i < parser_threads; i++) parsers_.emplace_back([this] { parse_loop(); });
    }
    ~FastaStream()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        if (reader_.joinable()) reader_.join();
        for (auto &t : parsers_) t.join();
        auto drop = [](FastaBatch *b) { if (b->bases) mhapb_host_free(b->bases); free(b->text); delete b; };
        for (auto *b : free_) drop(b);
        for (auto &kv : done_) drop(kv.second);
        for (auto *b : work_) drop(b);
        if (held_) drop(held_);
        if (gzf_) gzclose(gzf_);
        if (bzf_) bz_close_(bzf_);
        if (bzlib_) dlclose(bzlib_);
        if (fp_) fclose(fp_);
    }
    bool open_failed() const { return open_failed_; }

    // Next batch in file order, or nullptr at the end of the file (or after an empty record).  The previous batch
    // is recycled by this call.
    FastaBatch *next()
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (held_) { held_->len = 0; held_->offsets.clear(); held_->headers.clear(); held_->ended = false; held_->error.clear(); free_.push_back(held_); held_ = nullptr; cv_.notify_all(); }
        if (finished_) return nullptr;
        cv_.wait(lk, [this] { return done_.count(next_seq_) || (eof_ && next_seq_ >= eof_seq_); });
        auto it = done_.find(next_seq_);
        if (it == done_.end()) { finished_ = true; return nullptr; }
        held_ = it->second;
        done_.erase(it);
        next_seq_++;
        if (held_->ended || !held_->error.empty()) { finished_ = true; stop_ = true; cv_.notify_all(); }
        return held_;
    }

private:
    size_t read_some(char *dst, size_t want)
    {
        if (gz_) { int r = gzread(gzf_, dst, (unsigned)std::min<size_t>(want, 1u << 30)); return r > 0 ? (size_t)r : 0; }
        if (bz_) { int r = bz_read_(bzf_, dst, (int)std::min<size_t>(want, 1u << 30)); return r > 0 ? (size_t)r : 0; }
        return fread(dst, 1, want, fp_);
    }

    // Fill chunks that end just before a record start; the tail after the last "\n>" is carried to the next chunk.
    void read_loop()
    {
        std::vector<char> carry;
        uint64_t seq = 0;
        bool at_eof = false;
        while (!at_eof) {
            FastaBatch *b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !free_.empty(); });
                if (stop_) break;
                b = free_.back(); free_.pop_back();
            }
            const auto t0 = std::chrono::steady_clock::now();
            auto ensure = [&](size_t want) {
                if (b->text_cap >= want) return;
                b->text = (char *)realloc(b->text, want);     // no zero fill: every byte below `len` is written by the read
                b->text_cap = want;
            };
            // the first chunks are small (1/8, 1/4, 1/2 of the chunk size) so that the GPU starts early
            size_t target = chunk_;
            if (seq < 3) target = std::max<size_t>(1u << 16, chunk_ >> (3 - seq));
            size_t limit = target + carry.size();
            ensure(limit);
            char *t = b->text;
            size_t len = carry.size();
            if (len) memcpy(t, carry.data(), len);
            carry.clear();
            size_t cut = 0;
            for (;;) {
                const size_t got = read_some(t + len, limit - len);
                len += got;
                if (got == 0) { at_eof = true; cut = len; break; }
                if (len < limit) continue;                          // short read: keep filling
                // full buffer: cut at the last record start
                size_t p = len;
                while (p > 1) { const void *q = memrchr(t, '>', p - 1); if (!q) { p = 0; break; } p = (size_t)((const char *)q - t); if (p > 0 && t[p - 1] == '\n') break; }
                if (p > 1) { cut = p; break; }
                limit *= 2;                                         // one record larger than the chunk: grow and go on
                ensure(limit);
                t = b->text;
            }
            if (!at_eof) carry.assign(t + cut, t + len);
            b->t_read = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            b->text_len = cut; b->seq = seq; b->first_chunk = seq == 0;
            seq++;
            {
                std::lock_guard<std::mutex> lk(mu_);
                work_.push_back(b);
            }
            cv_.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            eof_ = true; eof_seq_ = seq;
        }
        cv_.notify_all();
    }

    void parse_loop()
    {
        for (;;) {
            FastaBatch *b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !work_.empty() || (eof_ && work_.empty()); });
                if (work_.empty()) { if (stop_ || eof_) return; continue; }
                b = work_.front(); work_.pop_front();
            }
            parse(*b, keep_headers_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                done_[b->seq] = b;
            }
            cv_.notify_all();
        }
    }

    static void parse(FastaBatch &b, bool keep_headers)
    {
        const auto t00 = std::chrono::steady_clock::now();
        parse_impl(b, keep_headers);
        b.t_parse = std::chrono::duration<double>(std::chrono::steady_clock::now() - t00).count();
    }

    static void parse_impl(FastaBatch &b, bool keep_headers)
    {
        const auto t0 = std::chrono::steady_clock::now();
        const char *p = b.text, *end = p + b.text_len;
        b.t_alloc = 0;
        if (b.cap < b.text_len + 64) {
            if (b.bases) mhapb_host_free(b.bases);
            void *q = nullptr;
            b.cap = b.text_len + b.text_len / 8 + 64;
            if (mhapb_host_alloc(b.cap, &q)) { b.error = "cudaHostAlloc failed for a FASTA batch"; b.bases = nullptr; b.cap = 0; return; }
            b.bases = (char *)q;
            b.t_alloc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        b.len = 0; b.offsets.clear(); b.offsets.push_back(0); b.headers.clear(); b.ended = false; b.error.clear();
        if (p == end) return;
        if (*p != '>') {
            // only the first line of the file can fail this test (a chunk always starts at a record start)
            if (b.first_chunk) b.error = "Next sequence does not start with >. Invalid format.";
            return;
        }
        while (p < end) {
            // header line; --store-full-id keeps line.substring(1).split("[\\s,]+", 2)[0] (FastaData.java:155-156)
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *hs = p + 1, *he = nl ? nl : end;
            p = nl ? nl + 1 : end;
            const size_t start = b.len;
            while (p < end && *p != '>') {
                nl = (const char *)memchr(p, '\n', (size_t)(end - p));
                const char *le = nl ? nl : end;
                size_t n = (size_t)(le - p);
                if (n && p[n - 1] == '\r') n--;
                memcpy(b.bases + b.len, p, n);
                b.len += n;
                p = nl ? nl + 1 : end;
            }
            if (b.len == start) { b.ended = true; return; }     // empty record: the reference stops reading here
            b.offsets.push_back(b.len);
            if (keep_headers) {
                const char *q = hs;
                while (q < he && !(*q == ' ' || *q == '\t' || *q == '\r' || *q == '\f' || *q == '\v' || *q == ',')) q++;
                b.headers.emplace_back(hs, (size_t)(q - hs));
            }
        }
    }

    std::string path_;
    size_t chunk_;
    bool gz_ = false, bz_ = false, open_failed_ = false, keep_headers_ = false;
    void *bzlib_ = nullptr, *bzf_ = nullptr;
    void *(*bz_open_)(const char *, const char *) = nullptr;
    int (*bz_read_)(void *, void *, int) = nullptr;
    void (*bz_close_)(void *) = nullptr;
    gzFile gzf_ = nullptr;
    FILE *fp_ = nullptr;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<FastaBatch *> free_;
    std::deque<FastaBatch *> work_;
    std::map<uint64_t, FastaBatch *> done_;
    FastaBatch *held_ = nullptr;
    uint64_t next_seq_ = 0, eof_seq_ = 0;
    bool eof_ = false, stop_ = false, finished_ = false;
    std::thread reader_;
    std::vector<std::thread> parsers_;
};

} // namespace mhapb_host
