// ubench_bitslice.cu -- bit-sliced evaluation of MHAP's XORShift step: 32 k-mers per thread, bit i of all
// 32 chain states in register R[i]; shifts become register renames.  Two forms: the plain plane form (132 two-input
// XORs, which nvcc fuses to ~117 LOP3) and the generated 92-gate form of mhap_b200/csrc/bs_step.cuh (inline-PTX lop3).
// Checks both against the scalar recurrence and measures their throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_bitslice ubench_bitslice.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../mhap_b200/csrc/bs_step.cuh"   // bs_step(): the generated 92-gate form

__device__ __forceinline__ void bs_step_plain(uint32_t (&R)[64])
{
#pragma unroll
    for (int i = 63; i >= 21; i--) R[i] ^= R[i - 21];      // x ^= x << 21
#pragma unroll
    for (int i = 0; i <= 28; i++) R[i] ^= R[i + 35];       // x ^= x >>> 35
#pragma unroll
    for (int i = 63; i >= 4; i--) R[i] ^= R[i - 4];        // x ^= x << 4
}

// 32x32 bit-matrix transpose in registers (a[j] bit i <-> a[i] bit j), Hacker's Delight 7-3
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= m << j) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            uint32_t t = (a[k] ^ (a[k + j] >> j)) & m;
            a[k] ^= t; a[k + j] ^= t << j;
        }
    }
}

template <int GEN> __device__ __forceinline__ void step(uint32_t (&R)[64]) { if (GEN) bs_step(R); else bs_step_plain(R); }

template <int GEN> __global__ void k_check(const uint64_t *keys, uint64_t *out, int steps)
{
    // one thread: 32 keys -> planes -> steps -> back
    uint32_t lo[32], hi[32];
    for (int c = 0; c < 32; c++) { lo[c] = (uint32_t)keys[c]; hi[c] = (uint32_t)(keys[c] >> 32); }
    transpose32(lo); transpose32(hi);
    uint32_t R[64];
    // after the transpose lo[i] holds bit i of every key?  (verified below by the round trip)
    for (int i = 0; i < 32; i++) { R[i] = lo[31 - i]; R[32 + i] = hi[31 - i]; }   // plane p = t[31-p]; key c sits at bit 31-c
    for (int s = 0; s < steps; s++) step<GEN>(R);
    for (int i = 0; i < 32; i++) { lo[31 - i] = R[i]; hi[31 - i] = R[32 + i]; }
    transpose32(lo); transpose32(hi);
    for (int c = 0; c < 32; c++) out[c] = ((uint64_t)hi[c] << 32) | lo[c];
}

constexpr int ITERS = 256;
template <int TAPS, int GEN> __global__ void __launch_bounds__(256) k_bs(unsigned long long *sink, uint32_t seed)
{
    uint32_t R[64];
#pragma unroll
    for (int i = 0; i < 64; i++) R[i] = seed * (threadIdx.x + 1 + blockIdx.x * 256) + i * 0x9e3779b9u;
    uint32_t acc = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
        for (int b = 0; b < 16; b++) {
            step<GEN>(R);
            if (TAPS) {
                uint32_t o = ~R[63] | R[62] | R[61]; o |= R[60] | R[59]; o |= R[58] | R[57]; o |= R[56] | R[55];
                o |= R[54] | R[53]; o |= R[52] | R[51];
                if (__builtin_expect(~o != 0, 0)) acc += __popc(~o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 64; i++) acc ^= R[i];
    if (acc == 0x12345u) atomicAdd(sink, 1ull);
}

int main()
{
    uint64_t *keys, *out;
    cudaMallocManaged(&keys, 32 * 8); cudaMallocManaged(&out, 32 * 8);
    uint64_t s = 88172645463325252ull;
    for (int c = 0; c < 32; c++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; keys[c] = s; }
    const int steps = 77;
    for (int gen = 0; gen < 2; gen++) {
        if (gen) k_check<1><<<1, 1>>>(keys, out, steps); else k_check<0><<<1, 1>>>(keys, out, steps);
        cudaDeviceSynchronize();
        for (int c = 0; c < 32; c++) {
            uint64_t x = keys[c];
            for (int i = 0; i < steps; i++) { x ^= x << 21; x ^= x >> 35; x ^= x << 4; }
            if (x != out[c]) { printf("MISMATCH (%s) at %d: %016llx vs %016llx\n", gen ? "92-gate" : "plain", c, (unsigned long long)x, (unsigned long long)out[c]); return 1; }
        }
        printf("%s bit-sliced step == scalar recurrence on 32 keys x %d steps\n", gen ? "92-gate" : "plain", steps);
    }
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bpsm : {1, 2, 3}) {
        const int grid = sms * bpsm;
        const double stepsd = (double)grid * 256 * 32.0 * ITERS * 16;
        for (int v = 0; v < 4; v++) {
            const int taps = v & 1, gen = v >> 1;
            float best = 1e30f;
            for (int r = 0; r < 4; r++) {
                cudaEventRecord(e0);
                if (gen) { if (taps) k_bs<1, 1><<<grid, 256>>>(sink, 12345u); else k_bs<0, 1><<<grid, 256>>>(sink, 12345u); }
                else     { if (taps) k_bs<1, 0><<<grid, 256>>>(sink, 12345u); else k_bs<0, 0><<<grid, 256>>>(sink, 12345u); }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
            }
            printf("%s bit-sliced %s  blocks/SM %d: %.3e k-mer-steps/s  (%.2f cycles per 32-k-mer thread-step-warp @1.965GHz)\n", gen ? "92-gate" : "plain  ", taps ? "with 13-bit prefix filter" : "bare", bpsm,
                   stepsd / (best * 1e-3), 148.0 * 4 * 1.965e9 / (stepsd / 32 / 32 / (best * 1e-3)));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
