// mhap-b200 -- host driver above the C ABI, mirroring the reference's command line for the hot path.
//
// The reference's driver is Java (main/MhapMain.java); this image has no JVM, so the host side above
// libmhap_b200.so is written in C++ with the same flags, the same stdout (MatchResult lines) and the
// same stderr bookkeeping, for the modes on the hot path:
//   -s <fasta|dat>                 self overlap                      (MhapMain.computeMain :452-477)
//   -s <fasta|dat> -q <file|dir>   store vs query files              (:478-541), --no-self
//   -p <fasta|dir> -q <outdir>     FASTA -> .dat sketch files        (:384-451)
//   -f <filter file> [--filter-threshold --repeat-idf-scale --supress-noise --no-tf]   k-mer filter / tf-idf
//                                  weights (main/MhapMain.java:340-372, sketch/FrequencyCounts.java); plain, .gz or .bz2
// FASTA input (plain, .gz or .bz2) is streamed: a reader thread + --num-threads parser threads fill pinned batches while the
// GPU works on the previous one (fasta_stream.hpp).  --store-full-id prints the first token of the FASTA headers instead of
// the file positions (impl/FastaData.java:155-156, impl/SequenceId.java; ignored for .dat input like the reference).
// Paths cited are relative to /root/reference/src/main/java/edu/umd/marbl/mhap/.
#include "../../include/mhap_b200.h"
#include "fasta_stream.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

[[noreturn]] void die(const std::string &msg)
{
    // the reference throws MhapRuntimeException; the JVM prints it and exits non-zero
    fprintf(stderr, "MhapRuntimeException: %s\n", msg.c_str());
    exit(1);
}

struct Options {
    std::string s, q, p, f;
    int k = 16, num_hashes = 512, num_min_matches = 3, num_threads = 1, ordered_kmer = 12, ordered_sketch = 1536;
    int min_store_length = 0, min_olap_length = 116, settings = 0, device = 0;
    std::vector<int> devices;   // --devices 0,1,2 / 0-7: the reads are sharded over these GPUs (mhapb_multi_*, NCCL inside the library)
    double threshold = 0.78, max_shift = 0.2, repeat_weight = 0.9, filter_threshold = 1.0e-5, repeat_idf_scale = 3.0;
    int supress_noise = 0;
    bool no_self = false, store_full_id = false, no_rc = false, no_tf = false;
    std::map<std::string, bool> set;
};

// utils/ParseOptions.java:327-368: flag and value are separate argv tokens, booleans are presence flags
Options parse(int argc, char **argv)
{
    Options o;
    auto need = [&](int &i) -> const char * { if (i + 1 >= argc) { printf("Missing value for option %s\n", argv[i]); exit(1); } return argv[++i]; };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        o.set[a] = true;
        if (a == "-s") o.s = need(i);
        else if (a == "-q") o.q = need(i);
        else if (a == "-p") o.p = need(i);
        else if (a == "-f") o.f = need(i);
        else if (a == "-k") o.k = atoi(need(i));
        else if (a == "--num-hashes") o.num_hashes = atoi(need(i));
        else if (a == "--threshold") o.threshold = atof(need(i));
        else if (a == "--max-shift") o.max_shift = atof(need(i));
        else if (a == "--num-min-matches") o.num_min_matches = atoi(need(i));
        else if (a == "--num-threads") o.num_threads = atoi(need(i));
        else if (a == "--repeat-weight") o.repeat_weight = atof(need(i));
        else if (a == "--ordered-kmer-size") o.ordered_kmer = atoi(need(i));
        else if (a == "--ordered-sketch-size") o.ordered_sketch = atoi(need(i));
        else if (a == "--min-store-length") o.min_store_length = atoi(need(i));
        else if (a == "--min-olap-length") o.min_olap_length = atoi(need(i));
        else if (a == "--settings") o.settings = atoi(need(i));
        else if (a == "--device") o.device = atoi(need(i));
        else if (a == "--devices") {
            std::string v = need(i);
            size_t pos = 0;
            while (pos <= v.size()) {
                size_t c = v.find(',', pos);
                std::string tok = v.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
                size_t dash = tok.find('-');
                if (tok.empty()) die("--devices: empty entry");
                if (dash != std::string::npos && dash > 0) { for (int d = atoi(tok.substr(0, dash).c_str()); d <= atoi(tok.substr(dash + 1).c_str()); d++) o.devices.push_back(d); }
                else o.devices.push_back(atoi(tok.c_str()));
                if (c == std::string::npos) break;
                pos = c + 1;
            }
        }
        else if (a == "--no-self") o.no_self = true;
        else if (a == "--no-rc") o.no_rc = true;   // main/MhapMain.java: does not stop rc sketches being stored (MinHashSearch.java:80)
        else if (a == "--store-full-id") o.store_full_id = true;
        else if (a == "--filter-threshold") o.filter_threshold = atof(need(i));
        else if (a == "--repeat-idf-scale") o.repeat_idf_scale = atof(need(i));
        else if (a == "--supress-noise") o.supress_noise = atoi(need(i));
        else if (a == "--no-tf") o.no_tf = true;
        else if (a == "-h" || a == "--help" || a == "--version") { printf("%s\n", mhapb_version()); exit(0); }
        else { printf("Unknown option %s\n", a.c_str()); exit(1); }
    }
    // main/MhapMain.java:130-198 presets fill only options the user did not set
    if (o.settings < 0 || o.settings > 3) { printf("Please enter valid --settings flag.\n"); exit(1); }
    auto unset = [&](const char *n) { return !o.set.count(n); };
    if (o.settings >= 1) {
        const int m[4] = {0, 3, 3, 2}, h[4] = {0, 512, 256, 768}, os[4] = {0, 1536, 1000, 1536}, ok[4] = {0, 12, 14, 12};
        const double thr[4] = {0, .78, .80, .73};
        if (unset("-k")) o.k = 16;
        if (unset("--num-min-matches")) o.num_min_matches = m[o.settings];
        if (unset("--num-hashes")) o.num_hashes = h[o.settings];
        if (unset("--threshold")) o.threshold = thr[o.settings];
        if (unset("--ordered-sketch-size")) o.ordered_sketch = os[o.settings];
        if (unset("--ordered-kmer-size")) o.ordered_kmer = ok[o.settings];
    }
    // main/MhapMain.java:200-300 validation, same messages
    struct stat st;
    if (o.s.empty() && o.p.empty()) { printf("Please set the -s or the -p options.\n"); exit(1); }
    if (!o.p.empty() && o.q.empty()) { printf("Please set the -q option.\n"); exit(1); }
    for (const std::string *f : {&o.p, &o.s, &o.q, &o.f})
        if (!f->empty() && stat(f->c_str(), &st) != 0) { printf("Could not find requested file/folder: %s\n", f->c_str()); exit(1); }
    if (o.num_threads <= 0) { printf("Number of threads must be positive.\n"); exit(1); }
    if (o.k <= 0) { printf("k-mer size must be positive.\n"); exit(1); }
    if (o.num_min_matches <= 0) { printf("Minimum number of matches must be positive.\n"); exit(1); }
    if (o.min_store_length < 0) { printf("The minimum read length stored must be >=0.\n"); exit(1); }
    if (o.max_shift < -1.0) { printf("The minimum shift must be greater than -1.\n"); exit(1); }
    if (o.threshold < 0.0 || o.threshold > 1.0) { printf("The second stage filter threshold must be 0<=threshold<=1.0.\n"); exit(1); }
    if (o.repeat_idf_scale < 1.0) { printf("The --repeat-idf-scale parameter must be >=1.0.\n"); exit(1); }               // :272-276
    if (o.supress_noise < 0 || o.supress_noise > 2) { printf("The --supress-noise parameter must be in [0,2].\n"); exit(1); }   // :293-297
    return o;
}

bool ends_with(const std::string &s, const char *suf)
{
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

bool is_dir(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

// directory listing without dot files, sorted (main/MhapMain.java:408-425,495-512)
std::vector<std::string> list_files(const std::string &path)
{
    std::vector<std::string> out;
    if (!is_dir(path)) { out.push_back(path); return out; }
    DIR *d = opendir(path.c_str());
    if (!d) die("Cannot list directory " + path);
    while (dirent *e = readdir(d)) if (e->d_name[0] != '.') out.push_back(path + "/" + e->d_name);
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

// impl/FastaData.java:125-204 through the streaming producer: fn(batch, ids) is called once per batch in file order;
// ids are the 1-based positions of the records in the file (+offset).  Returns the number of records read.
// batch size of the FASTA producer in text bytes: 64 MB (6.7 k reads of 10 kbp keep every SM busy and the pinned buffers small)
size_t fasta_chunk_bytes(const std::string &path)
{
    struct stat st;
    size_t chunk = 64u << 20;
    if (const char *e = getenv("MHAPB_FASTA_CHUNK_KB")) chunk = (size_t)std::max(64, atoi(e)) << 10;
    if (!ends_with(path, ".gz") && !ends_with(path, ".bz2") && stat(path.c_str(), &st) == 0 && (size_t)st.st_size + 4096 < chunk) chunk = (size_t)st.st_size + 4096;
    return chunk;
}

// --store-full-id: SequenceId.getHeader() of the sequences of one FASTA file, indexed by position in the file.  Kept per
// file, not per id: the ids of a query file start at the number of STORED sequences (main/MhapMain.java:462,537), which can
// overlap the file positions of the store when short reads were skipped.
typedef std::vector<std::string> Names;
bool g_full_ids = false;

template <class F>
int64_t for_each_fasta_batch(const std::string &path, int64_t offset, int threads, Names *names, F fn)
{
    const size_t chunk = fasta_chunk_bytes(path);
    mhapb_host::FastaStream fs(path, std::max(1, std::min(threads, 8)), chunk, g_full_ids);
    if (fs.open_failed()) die("Could not open " + path);
    int64_t n = 0;
    std::vector<int64_t> ids;
    const bool trace = getenv("MHAPB_FASTA_TRACE") != nullptr;
    for (;;) {
        const double tw = now_s();
        mhapb_host::FastaBatch *b = fs.next();
        if (!b) break;
        const double t0 = now_s();
        if (!b->error.empty()) die(b->error);
        const uint32_t nb = b->n_reads();
        if (nb) {
            ids.resize(nb);
            for (uint32_t i = 0; i < nb; i++) ids[i] = n + i + 1 + offset;
            if (g_full_ids && names) names->insert(names->end(), b->headers.begin(), b->headers.end());
            fn(*b, ids);
            n += nb;
        }
        if (trace) fprintf(stderr, "[fasta] batch %llu: %u reads, %.1f MB text; read %.3f s, parse %.3f s (pinned alloc %.3f s), waited %.3f s, library call %.3f s\n",
                           (unsigned long long)b->seq, nb, b->text_len / 1048576.0, b->t_read, b->t_parse, b->t_alloc, t0 - tw, now_s() - t0);
    }
    return n;
}

std::vector<uint8_t> read_file(const std::string &path)
{
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) die("Could not open " + path);
    std::vector<uint8_t> buf((size_t)in.tellg());
    in.seekg(0);
    in.read((char *)buf.data(), (std::streamsize)buf.size());
    return buf;
}

// utils/Utils.java getFile: plain, .gz or .bz2 text (the -f filter file)
std::vector<uint8_t> read_text_any(const std::string &path)
{
    if (!ends_with(path, ".gz") && !ends_with(path, ".bz2")) return read_file(path);
    std::vector<uint8_t> out;
    std::vector<uint8_t> buf(1u << 20);
    if (ends_with(path, ".gz")) {
        gzFile f = gzopen(path.c_str(), "rb");
        if (!f) die("Could not open " + path);
        for (int r; (r = gzread(f, buf.data(), (unsigned)buf.size())) > 0;) out.insert(out.end(), buf.begin(), buf.begin() + r);
        gzclose(f);
        return out;
    }
    void *lib = nullptr;
    for (const char *l : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) if ((lib = dlopen(l, RTLD_NOW))) break;
    if (!lib) die("libbz2 not found, cannot read " + path);
    auto bzopen = (void *(*)(const char *, const char *))dlsym(lib, "BZ2_bzopen");
    auto bzread = (int (*)(void *, void *, int))dlsym(lib, "BZ2_bzread");
    auto bzclose = (void (*)(void *))dlsym(lib, "BZ2_bzclose");
    void *f = (bzopen && bzread && bzclose) ? bzopen(path.c_str(), "rb") : nullptr;
    if (!f) die("Could not open " + path);
    for (int r; (r = bzread(f, buf.data(), (int)buf.size())) > 0;) out.insert(out.end(), buf.begin(), buf.begin() + r);
    bzclose(f);
    dlclose(lib);
    return out;
}

struct DatSketches {
    uint32_t n = 0; int32_t H = 0, max_ord = 0, ok = 0;
    std::vector<int64_t> ids; std::vector<uint8_t> fwd; std::vector<int32_t> len, lenk, mh, ord, ordn;
};

// impl/SequenceSketchStreamer.java:278-320 + impl/SequenceSketch.java:61-96 (ids get +offset)
DatSketches read_dat(const std::string &path, int64_t offset)
{
    std::vector<uint8_t> buf = read_file(path);
    DatSketches d;
    if (mhapb_dat_decode(buf.data(), buf.size(), offset, &d.n, &d.H, &d.max_ord, &d.ok, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))
        die("Unexpected data read error.");
    int32_t stride = std::max(1, d.max_ord);
    d.ids.resize(d.n); d.fwd.resize(d.n); d.len.resize(d.n); d.lenk.resize(d.n); d.ordn.resize(d.n);
    d.mh.resize((size_t)d.n * d.H); d.ord.resize((size_t)d.n * stride * 2);
    d.max_ord = stride;
    if (mhapb_dat_decode(buf.data(), buf.size(), offset, &d.n, &d.H, &d.max_ord, &d.ok, d.ids.data(), d.fwd.data(), d.len.data(), d.lenk.data(),
                         d.mh.data(), d.ord.data(), d.ordn.data()))
        die("Unexpected data read error.");
    return d;
}

void ck(mhapb_ctx *ctx, int rc) { if (rc) die(mhapb_last_error(ctx)); }
void ckm(mhapb_multi *m, int rc) { if (rc) die(mhapb_multi_last_error(m)); }

struct Totals { mhapb_stats st{}; };

// from_sub: sketches read from a .dat file print the header string stored in the record, i.e. the
// file-local id without the run's offset (impl/SequenceId.java:102-108, SequenceSketch.java:75)
void emit(mhapb_hit *hits, uint64_t n, const mhapb_stats &st, Totals &tot, int64_t from_sub = 0,
          const Names *from_names = nullptr, int64_t from_offset = 0, const Names *to_names = nullptr)
{
    // AbstractMatchSearch.outputResults :316-338: one MatchResult.toString() per line on stdout
    char line[256];
    auto name = [](const Names *t, int64_t pos, int64_t id) {
        return (t && pos >= 0 && (size_t)pos < t->size()) ? (*t)[(size_t)pos] : std::to_string(id);
    };
    for (uint64_t i = 0; i < n; i++) {
        hits[i].from_id -= from_sub;
        mhapb_format_match(&hits[i], line, sizeof line);
        if (g_full_ids && (from_names || to_names)) {
            // MatchResult.toString prints fromId.getHeader() / toId.getHeader(): the FASTA names where the sequences came from
            // FASTA files, the decimal ids for sketches read from .dat records
            const char *rest = strchr(line, ' ');
            rest = rest ? strchr(rest + 1, ' ') : nullptr;
            printf("%s %s%s\n", name(from_names, hits[i].from_id - 1 - from_offset, hits[i].from_id).c_str(),
                   name(to_names, hits[i].to_id - 1, hits[i].to_id).c_str(), rest ? rest : "");
        } else puts(line);
    }
    fflush(stdout);
    mhapb_free(hits);
    tot.st.elements_processed += st.elements_processed; tot.st.sequences_hit += st.sequences_hit;
    tot.st.fully_compared += st.fully_compared; tot.st.matches_processed += st.matches_processed;
    tot.st.sequences_searched += st.sequences_searched;
}

} // namespace

int main(int argc, char **argv)
{
    Options o = parse(argc, argv);
    g_full_ids = o.store_full_id;
    const double t_total = now_s();
    // one context per GPU behind mhapb_multi (a single device forwards to the single-GPU calls); ctx = the first device,
    // which also serves the -p mode
    if (o.devices.empty()) o.devices.push_back(o.device);
    mhapb_multi *multi = nullptr;
    if (mhapb_multi_create(o.devices.data(), (int)o.devices.size(), &multi)) die(mhapb_last_error(nullptr));
    const int n_dev = mhapb_multi_n_devices(multi);
    mhapb_ctx *ctx = mhapb_multi_ctx(multi, 0);
    mhapb_sketch_params p{o.k, o.num_hashes, o.ordered_kmer, o.ordered_sketch, o.repeat_weight < 0.0 ? 1 : 0, o.min_olap_length};

    if (!o.f.empty()) {   // main/MhapMain.java:340-372: read the k-mer filter set
        const double t0 = now_s();
        fprintf(stderr, "Reading in filter file %s.\n", o.f.c_str());
        std::vector<uint8_t> text = read_text_any(o.f);
        mhapb_filter_params fp{o.filter_threshold, o.repeat_weight, o.repeat_idf_scale, o.supress_noise, o.no_tf ? 1 : 0};
        int64_t n_repeat = 0;
        for (int d = 0; d < n_dev; d++)   // every device sketches with the same filter
            if (mhapb_filter_load_text(mhapb_multi_ctx(multi, d), &fp, (const char *)text.data(), text.size(), o.no_rc ? 0 : 1, &n_repeat))
                die(std::string("Could not parse k-mer filter file. ") + mhapb_last_error(mhapb_multi_ctx(multi, d)));
        fprintf(stderr, "Time (s) to read filter file: %g\n", now_s() - t0);
        fprintf(stderr, "Read in k-mer filter with %lld repeat k-mers.\n", (long long)n_repeat);
    }

    if (!o.p.empty()) {   // main/MhapMain.java:384-451
        fprintf(stderr, "Processing FASTA files for binary compression...\n");
        if (!is_dir(o.q)) die("Target directory doesn't exit.");
        for (const std::string &pf : list_files(o.p)) {
            const double t0 = now_s();
            std::string name = pf.substr(pf.find_last_of('/') == std::string::npos ? 0 : pf.find_last_of('/') + 1);
            size_t dot = name.find_last_of('.');   // main/MhapMain.java:431-434: only the last extension goes (reads.fasta.gz -> reads.fasta.dat)
            if (dot != std::string::npos && dot > 0) name = name.substr(0, dot);
            std::string outp = o.q + "/" + name + ".dat";
            std::ofstream out(outp, std::ios::binary);
            if (!out) die("Could not open " + outp);
            uint32_t nrec = 0;
            for_each_fasta_batch(pf, 0, o.num_threads, nullptr, [&](mhapb_host::FastaBatch &b, const std::vector<int64_t> &ids) {
                uint8_t *blob = nullptr; uint64_t len = 0; uint32_t nr = 0;
                // --store-full-id: the record's header string is the FASTA name (SequenceId.getHeader), otherwise the decimal id
                std::vector<const char *> hdr;
                if (g_full_ids) for (const std::string &h : b.headers) hdr.push_back(h.c_str());
                ck(ctx, mhapb_sketch_to_dat_named(ctx, &p, b.bases, b.offsets.data(), ids.data(), hdr.size() == b.n_reads() ? hdr.data() : nullptr,
                                                  b.n_reads(), 1, &blob, &len, &nr));
                out.write((const char *)blob, (std::streamsize)len);
                mhapb_free(blob);
                nrec += nr;
            });
            fprintf(stderr, "Processed %u sequences (fwd and rev).\n", nrec);
            fprintf(stderr, "Read, hashed, and stored file %s to %s.\n", pf.c_str(), outp.c_str());
            fprintf(stderr, "Time (s): %g\n", now_s() - t0);
        }
        fprintf(stderr, "Total time (s): %g\n", now_s() - t_total);
        mhapb_multi_destroy(multi);
        return 0;
    }

    fprintf(stderr, "Processing files for storage in reverse index...\n");
    const double t_proc = now_s();
    int64_t n_sketches = 0;
    int32_t store_ok = o.ordered_kmer;   // the ordered k-mer size the stored sketches carry
    Names store_names;   // empty for .dat stores: ids print as numbers
    if (ends_with(o.s, ".dat")) {
        DatSketches d = read_dat(o.s, 0);
        if (d.n && d.H != o.num_hashes) die("Number of MinHashes of the sequence does not match current settings.");   // MinHashSearch.java:105
        p.ordered_sketch_size = std::max(p.ordered_sketch_size, d.max_ord);
        // a stored sketch scores with the k-mer size recorded in it (BottomOverlapSketch.kmerSize, :391-395,:613), whatever
        // --ordered-kmer-size says; sketches made with another size then fail the check of :594-595 when compared
        if (d.n) { store_ok = d.ok; p.ordered_kmer_size = d.ok; }
        ckm(multi, mhapb_multi_store_reset(multi, &p));
        ckm(multi, mhapb_multi_store_add_sketches(multi, d.ids.data(), d.fwd.data(), d.len.data(), d.lenk.data(), d.mh.data(), d.H, d.ord.data(), d.ordn.data(), d.max_ord, d.ok, d.n));
        n_sketches = d.n;
    } else {
        ckm(multi, mhapb_multi_store_reset(multi, &p));
        {   // one allocation of the per-call scratch for the largest batch instead of one per ramp step
            const size_t chunk = fasta_chunk_bytes(o.s) / (size_t)n_dev + 65536;
            for (int d = 0; d < n_dev; d++)
                ck(mhapb_multi_ctx(multi, d), mhapb_sketch_reserve(mhapb_multi_ctx(multi, d), &p, chunk, (uint32_t)std::min<size_t>(chunk / 500 + 64, 1u << 24), 1));
        }
        struct stat fst;
        const double file_bytes = (!ends_with(o.s, ".gz") && !ends_with(o.s, ".bz2") && stat(o.s.c_str(), &fst) == 0) ? (double)fst.st_size : 0.0;
        for_each_fasta_batch(o.s, 0, o.num_threads, &store_names, [&](mhapb_host::FastaBatch &b, const std::vector<int64_t> &ids) {
            if (b.seq == 0 && file_bytes > 0 && b.text_len > 0 && (double)b.text_len < file_bytes) {
                // size the store once from the first batch's record density instead of growing it batch by batch
                const double est_reads = file_bytes / (double)b.text_len * b.n_reads() * 1.03 + 64;
                ckm(multi, mhapb_multi_store_reserve(multi, (int64_t)(2 * est_reads)));
            }
            int64_t added = 0;
            ckm(multi, mhapb_multi_store_add_reads(multi, b.bases, b.offsets.data(), ids.data(), b.n_reads(), 1, &added));
            n_sketches += added;
        });
    }
    if (n_sketches > 0 && n_dev == 1) ck(ctx, mhapb_index_build(ctx));   // several devices: every rank builds its index behind the exchange
    fprintf(stderr, "Stored %lld sequences in the index.\n", (long long)n_sketches);
    int64_t seq_number_processed = n_sketches / 2;   // main/MhapMain.java:462
    fprintf(stderr, "Processed %lld unique sequences (fwd and rev).\n", (long long)n_sketches);
    fprintf(stderr, "Time (s) to read and hash from file: %g\n", now_s() - t_proc);

    const double t_score = now_s();
    mhapb_search_params sp{o.num_min_matches, o.min_store_length, o.max_shift, o.threshold, 0, 0, 0, -1};
    Totals tot;
    auto self = [&]() {
        const double t0 = now_s();
        if (n_sketches > 0) {
            mhapb_hit *hits = nullptr; uint64_t n = 0; mhapb_stats st{};
            ckm(multi, mhapb_multi_search_self(multi, &sp, &hits, &n, &st));
            emit(hits, n, st, tot, 0, store_names.empty() ? nullptr : &store_names, 0, store_names.empty() ? nullptr : &store_names);
        }
        fprintf(stderr, "Time (s) to score and output to self: %g\n", now_s() - t0);
    };
    if (o.q.empty()) self();
    else {
        if (!o.no_self) self();
        for (const std::string &cf : list_files(o.q)) {   // :525-541
            const double t0 = now_s();
            fprintf(stderr, "Opened fasta file %s.\n", cf.c_str());
            mhapb_hit *hits = nullptr; uint64_t n = 0; mhapb_stats st{};
            int64_t processed = 0, from_sub = 0;
            if (n_sketches == 0) { /* nothing stored: nothing can match */ }
            else if (ends_with(cf, ".dat")) {
                DatSketches d = read_dat(cf, seq_number_processed);
                // the library repeats the reference's checks: "Number of hashes does not match..." (MinHashSearch.java:157-159),
                // "Sketch k-mer size does not match between the two sequences." (BottomOverlapSketch.java:594-595)
                ckm(multi, mhapb_multi_search_query_sketches(multi, &sp, d.ids.data(), d.fwd.data(), d.len.data(), d.lenk.data(), d.mh.data(), d.H, d.ord.data(),
                                                    d.ordn.data(), d.max_ord, d.ok, d.n, &hits, &n, &st));
                processed = st.sequences_searched;
                from_sub = seq_number_processed;
            } else {
                if (store_ok != o.ordered_kmer) die("Sketch k-mer size does not match between the two sequences.");   // .dat store made with another --ordered-kmer-size
                Names query_names;
                for_each_fasta_batch(cf, seq_number_processed, o.num_threads, &query_names, [&](mhapb_host::FastaBatch &b, const std::vector<int64_t> &ids) {
                    mhapb_hit *bh = nullptr; uint64_t bn = 0; mhapb_stats bst{};
                    ckm(multi, mhapb_multi_search_query_reads(multi, &sp, b.bases, b.offsets.data(), ids.data(), b.n_reads(), &bh, &bn, &bst));
                    emit(bh, bn, bst, tot, 0, &query_names, seq_number_processed, store_names.empty() ? nullptr : &store_names);
                    processed += bst.sequences_searched;
                });
            }
            if (hits) emit(hits, n, st, tot, from_sub);
            seq_number_processed += processed;   // :537 counts the sketched (forward) query sequences
            fprintf(stderr, "Processed %lld to sequences.\n", (long long)processed);
            fprintf(stderr, "Time (s) to score, hash to-file, and output: %g\n", now_s() - t0);
        }
    }
    fprintf(stderr, "Total scoring time (s): %g\n", now_s() - t_score);
    fprintf(stderr, "Total time (s): %g\n", now_s() - t_total);
    // main/MhapMain.java:572-590
    const mhapb_stats &s = tot.st;
    const double size = (double)mhapb_multi_store_size(multi);
    fprintf(stderr, "Total matches found: %lld\n", (long long)s.matches_processed);
    fprintf(stderr, "Average number of matches per lookup: %g\n", (double)s.matches_processed / (double)s.sequences_searched);
    fprintf(stderr, "Average number of table elements processed per lookup: %g\n", (double)s.elements_processed / (double)s.sequences_searched);
    fprintf(stderr, "Average number of table elements processed per match: %g\n", (double)s.elements_processed / (double)s.matches_processed);
    fprintf(stderr, "Average %% of hashed sequences hit per lookup: %g\n", (double)s.sequences_hit / (size * (double)s.sequences_searched) * 100.0);
    fprintf(stderr, "Average %% of hashed sequences hit that are matches: %g\n", (double)s.matches_processed / (double)s.sequences_hit * 100.0);
    fprintf(stderr, "Average %% of hashed sequences fully compared that are matches: %g\n", (double)s.matches_processed / (double)s.fully_compared * 100.0);
    mhapb_multi_destroy(multi);
    return 0;
}
