// ubench_xorshift.cu -- which instruction mix runs the MHAP XORShift step fastest on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_xorshift ubench_xorshift.cu
// Every variant computes the identical recurrence (checked against the plain C form on the host);
// they differ only in which shifts are expressed as multiplies (fma pipe) vs shifts (alu pipe).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define MUL_WIDE(w, a, imm) asm("mul.wide.u32 %0, %1, " #imm ";" : "=l"(w) : "r"(a))
#define MUL_LO(r, a, imm)   asm("mul.lo.u32 %0, %1, " #imm ";" : "=r"(r) : "r"(a))
#define MUL_HI(r, a, imm)   asm("mul.hi.u32 %0, %1, " #imm ";" : "=r"(r) : "r"(a))

template <int V> __device__ __forceinline__ void step(uint32_t &lo, uint32_t &hi)
{
    if (V == 0) {          // plain C: compiler's choice (7 alu + 2 IMAD.SHL)
        uint64_t x = ((uint64_t)hi << 32) | lo;
        x ^= x << 21; x ^= x >> 35; x ^= x << 4;
        lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
    } else if (V == 1) {   // X: hi1>>3 via mul.hi
        uint32_t a = lo << 21, f = __funnelshift_l(lo, hi, 21);
        uint32_t hi1 = hi ^ f, t2; MUL_HI(t2, hi1, 0x20000000);
        uint32_t lo2 = lo ^ a ^ t2;
        uint32_t g = __funnelshift_l(lo2, hi1, 4);
        lo = lo2 ^ (lo2 << 4); hi = hi1 ^ g;
    } else if (V == 2) {   // Y: second half via mul.wide, hi1>>3 via mul.hi
        uint32_t a; MUL_LO(a, lo, 0x200000);
        uint32_t f = __funnelshift_l(lo, hi, 21);
        uint32_t hi1 = hi ^ f, t2; MUL_HI(t2, hi1, 0x20000000);
        uint32_t lo2 = lo ^ a ^ t2;
        uint64_t w2; MUL_WIDE(w2, lo2, 16);
        uint32_t hs2; MUL_LO(hs2, hi1, 16);
        lo = lo2 ^ (uint32_t)w2; hi = hi1 ^ hs2 ^ (uint32_t)(w2 >> 32);
    } else if (V == 3) {   // Z: every shift a multiply
        uint64_t w1, w2; uint32_t hs, t2, hs2;
        MUL_WIDE(w1, lo, 0x200000); MUL_LO(hs, hi, 0x200000);
        uint32_t hi1 = hi ^ hs ^ (uint32_t)(w1 >> 32);
        MUL_HI(t2, hi1, 0x20000000);
        uint32_t lo2 = lo ^ (uint32_t)w1 ^ t2;
        MUL_WIDE(w2, lo2, 16); MUL_LO(hs2, hi1, 16);
        lo = lo2 ^ (uint32_t)w2; hi = hi1 ^ hs2 ^ (uint32_t)(w2 >> 32);
    } else if (V == 4) {   // Y with hi1>>3 as a shift
        uint32_t a; MUL_LO(a, lo, 0x200000);
        uint32_t f = __funnelshift_l(lo, hi, 21);
        uint32_t hi1 = hi ^ f;
        uint32_t lo2 = lo ^ a ^ (hi1 >> 3);
        uint64_t w2; MUL_WIDE(w2, lo2, 16);
        uint32_t hs2; MUL_LO(hs2, hi1, 16);
        lo = lo2 ^ (uint32_t)w2; hi = hi1 ^ hs2 ^ (uint32_t)(w2 >> 32);
    } else if (V == 5) {   // first half via mul.wide, rest shifts
        uint64_t w1; MUL_WIDE(w1, lo, 0x200000); uint32_t hs; MUL_LO(hs, hi, 0x200000);
        uint32_t hi1 = hi ^ hs ^ (uint32_t)(w1 >> 32);
        uint32_t lo2 = lo ^ (uint32_t)w1 ^ (hi1 >> 3);
        uint32_t g = __funnelshift_l(lo2, hi1, 4);
        uint32_t b; MUL_LO(b, lo2, 16);
        lo = lo2 ^ b; hi = hi1 ^ g;
    } else if (V == 6) {   // only the two left shifts of the low word as multiplies + hi<<4 split
        uint32_t a; MUL_LO(a, lo, 0x200000);
        uint32_t f = __funnelshift_l(lo, hi, 21);
        uint32_t hi1 = hi ^ f;
        uint32_t lo2 = lo ^ a ^ (hi1 >> 3);
        uint32_t b; MUL_LO(b, lo2, 16); uint32_t hs2; MUL_LO(hs2, hi1, 16);
        lo = lo2 ^ b; hi = hi1 ^ hs2 ^ (lo2 >> 28);
    } else if (V == 7) {   // like 6 but hi1>>3 via mul.hi
        uint32_t a; MUL_LO(a, lo, 0x200000);
        uint32_t f = __funnelshift_l(lo, hi, 21);
        uint32_t hi1 = hi ^ f, t2; MUL_HI(t2, hi1, 0x20000000);
        uint32_t lo2 = lo ^ a ^ t2;
        uint32_t b; MUL_LO(b, lo2, 16); uint32_t hs2; MUL_LO(hs2, hi1, 16);
        lo = lo2 ^ b; hi = hi1 ^ hs2 ^ (lo2 >> 28);
    }
}

constexpr int ITERS = 2048, ILP = 4, UNR = 8;

// bare recurrence
template <int V> __global__ void __launch_bounds__(256) k_bare(unsigned long long *sink, uint64_t seed)
{
    uint32_t xl[ILP], xh[ILP];
    for (int i = 0; i < ILP; i++) { uint64_t x = seed * (blockIdx.x * 256ull + threadIdx.x + 1) + i; xl[i] = (uint32_t)x; xh[i] = (uint32_t)(x >> 32); }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < UNR; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) step<V>(xl[i], xh[i]);
    uint64_t acc = 0;
    for (int i = 0; i < ILP; i++) acc ^= ((uint64_t)xh[i] << 32) | xl[i];
    if (sink) sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

// with the per-step high-word compare against 16 register thresholds, checked in groups of 4 (as K1b)
template <int V> __global__ void __launch_bounds__(256) k_min(unsigned long long *sink, uint64_t seed)
{
    int hi[16]; uint32_t lo[16];
#pragma unroll
    for (int b = 0; b < 16; b++) { hi[b] = (int)0x80000400 + b; lo[b] = 0; }
    uint64_t x = seed * (blockIdx.x * 256ull + threadIdx.x + 1);
    uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    unsigned cnt = 0;
    for (int it = 0; it < ITERS * ILP * UNR / 16; it++) {
#pragma unroll
        for (int b0 = 0; b0 < 16; b0 += 4) {
            uint32_t l[4], h[4]; bool any = false;
#pragma unroll
            for (int g = 0; g < 4; g++) { step<V>(xl, xh); l[g] = xl; h[g] = xh; any |= (int)xh <= hi[b0 + g]; }
            if (__builtin_expect(any, 0)) {
#pragma unroll
                for (int g = 0; g < 4; g++)
                    if ((int)h[g] < hi[b0 + g] || ((int)h[g] == hi[b0 + g] && l[g] < lo[b0 + g])) { hi[b0 + g] = (int)h[g]; lo[b0 + g] = l[g]; cnt++; }
            }
        }
    }
    uint64_t acc = cnt;
#pragma unroll
    for (int b = 0; b < 16; b++) acc ^= ((uint64_t)(uint32_t)hi[b] << 32) | lo[b];
    acc ^= ((uint64_t)xh << 32) | xl;
    if (sink) sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int V> __global__ void k_check(uint64_t *out, uint64_t x0, int n)
{
    uint32_t lo = (uint32_t)x0, hi = (uint32_t)(x0 >> 32);
    for (int i = 0; i < n; i++) step<V>(lo, hi);
    out[V] = ((uint64_t)hi << 32) | lo;
}

template <int V> void run(unsigned long long *sink, int sms, cudaEvent_t e0, cudaEvent_t e1, int bpsm)
{
    const int grid = sms * bpsm;
    const double steps = (double)grid * 256 * ITERS * ILP * UNR;
    float best_b = 1e30f, best_m = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k_bare<V><<<grid, 256>>>(sink, 0x9E3779B97F4A7C15ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best_b) best_b = ms;
        cudaEventRecord(e0); k_min<V><<<grid, 256>>>(sink, 0x9E3779B97F4A7C15ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best_m) best_m = ms;
    }
    printf("variant %d  blocks/SM %d  bare %.3e steps/s (%.2f cyc/warp-step @1.965GHz)   with-min %.3e steps/s (%.2f cyc)\n", V, bpsm,
           steps / (best_b * 1e-3), 148.0 * 4 * 1.965e9 * 32 / (steps / (best_b * 1e-3)), steps / (best_m * 1e-3), 148.0 * 4 * 1.965e9 * 32 / (steps / (best_m * 1e-3)));
}

int main()
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *sink; cudaMalloc(&sink, (size_t)sms * 16 * 256 * 8);
    uint64_t *chk; cudaMallocManaged(&chk, 64 * 8);
    k_check<0><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000); k_check<1><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000);
    k_check<2><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000); k_check<3><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000);
    k_check<4><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000); k_check<5><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000);
    k_check<6><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000); k_check<7><<<1, 1>>>(chk, 0x0123456789abcdefull, 1000);
    cudaDeviceSynchronize();
    uint64_t x = 0x0123456789abcdefull; for (int i = 0; i < 1000; i++) { x ^= x << 21; x ^= x >> 35; x ^= x << 4; }
    for (int v = 0; v < 8; v++) if (chk[v] != x) { printf("variant %d WRONG\n", v); return 1; }
    printf("all variants agree with the plain recurrence\n");
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bpsm : {4, 8}) {
        run<0>(sink, sms, e0, e1, bpsm); run<1>(sink, sms, e0, e1, bpsm); run<2>(sink, sms, e0, e1, bpsm); run<3>(sink, sms, e0, e1, bpsm);
        run<4>(sink, sms, e0, e1, bpsm); run<5>(sink, sms, e0, e1, bpsm); run<6>(sink, sms, e0, e1, bpsm); run<7>(sink, sms, e0, e1, bpsm);
    }
    return 0;
}
