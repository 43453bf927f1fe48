// CPU harness for mhap_b200/host/fasta_stream.hpp: pinned allocation is stubbed with malloc so the parser can be
// exercised without a GPU.  Prints one line per record: "<index> <length> <fnv1a64 of the sequence>", then "END <n> <ended>"
// or "ERROR <text>".
#include "../../mhap_b200/host/fasta_stream.hpp"

#include <cstdlib>

extern "C" int mhapb_host_alloc(size_t bytes, void **out) { *out = malloc(bytes ? bytes : 1); return *out ? 0 : -4; }
extern "C" void mhapb_host_free(void *p) { free(p); }

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    const bool headers = argc > 4;
    mhapb_host::FastaStream fs(argv[1], atoi(argv[2]), (size_t)atoll(argv[3]), headers);
    if (fs.open_failed()) { printf("ERROR open\n"); return 1; }
    long long n = 0; int ended = 0;
    while (mhapb_host::FastaBatch *b = fs.next()) {
        if (!b->error.empty()) { printf("ERROR %s\n", b->error.c_str()); return 0; }
        for (uint32_t i = 0; i < b->n_reads(); i++) {
            unsigned long long h = 1469598103934665603ull;
            for (uint64_t j = b->offsets[i]; j < b->offsets[i + 1]; j++) { h ^= (unsigned char)b->bases[j]; h *= 1099511628211ull; }
            if (headers) printf("%lld %llu %llu [%s]\n", n++, (unsigned long long)(b->offsets[i + 1] - b->offsets[i]), h, b->headers[i].c_str());
            else printf("%lld %llu %llu\n", n++, (unsigned long long)(b->offsets[i + 1] - b->offsets[i]), h);
        }
        ended |= b->ended;
    }
    printf("END %lld %d\n", n, ended);
    return 0;
}
