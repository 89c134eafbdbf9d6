// tools/arith_bench.cu -- register-level throughput of candidate 64-bit butterfly formulations on
// B200 (the "integer roof" DESIGN.md quotes) plus raw pipe rates.  Each probe keeps 16 coefficients
// per thread in registers and runs radix-16 rounds (4 butterfly stages, 32 butterflies) in a loop,
// reading its 15 twiddles from shared memory every round like the real kernels do.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpu_ntt_b200/csrc -o tools/arith_bench tools/arith_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "modarith.cuh"
using namespace gpuntt_b200;

template <typename TW> TW make_tw(uint64_t w, uint64_t p);
template <typename M, typename TW, bool INV> __device__ __forceinline__ void round16(uint64_t (&e)[16], const TW* tw, const M& m)
{
    if constexpr (!INV)
    {
#pragma unroll
        for (int it = 0; it < 4; it++)
        {
            const int ab = 3 - it;
#pragma unroll
            for (int x = 0; x < (16 >> (ab + 1)); x++)
            {
                const TW w = tw[(16 >> (ab + 1)) - 1 + x];
#pragma unroll
                for (int y = 0; y < (1 << ab); y++)
                {
                    const int a0 = (x << (ab + 1)) | y;
                    m.ct(e[a0], e[a0 | (1 << ab)], w);
                }
            }
        }
    }
    else
    {
#pragma unroll
        for (int ab = 0; ab < 4; ab++)
#pragma unroll
            for (int x = 0; x < (16 >> (ab + 1)); x++)
            {
                const TW w = tw[(16 >> (ab + 1)) - 1 + x];
#pragma unroll
                for (int y = 0; y < (1 << ab); y++)
                {
                    const int a0 = (x << (ab + 1)) | y;
                    m.gs(e[a0], e[a0 | (1 << ab)], w);
                }
            }
    }
}

template <typename M, typename TW, bool INV, bool RENORM, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) probe(uint64_t* out, const TW* gtw, uint64_t p, int iters)
{
    __shared__ TW stw[64 * 15];
    for (int i = threadIdx.x; i < 64 * 15; i += blockDim.x) stw[i] = gtw[i];
    __syncthreads();
    const M m(p);
    uint64_t e[16];
#pragma unroll
    for (int i = 0; i < 16; i++) e[i] = (out[(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i]) % p;
    for (int it = 0; it < iters; it++)
    {
        round16<M, TW, INV>(e, stw + ((it + (threadIdx.x >> 5)) & 63) * 15, m);
        if constexpr (RENORM)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = m.renorm(e[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) out[(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] = m.canon_fwd(e[i]);
}

// ---------------------------------------------------------------- raw pipe probes
// Inline PTX so that ptxas cannot re-balance the instruction mix (it rewrites IADD3 as IMAD.IADD /
// IMAD.X and splits mad.wide with a 64-bit addend when it sees plain C).
template <int KIND> __global__ void __launch_bounds__(256, 2) pipe(uint32_t* out, uint32_t seed, int iters)
{
    uint32_t a[8], l[8], b = seed | 1, c = seed * 3 + 7;
    uint64_t w[8];
    double d[8], dm = 1.0 + seed * 1e-9, dc = seed * 1e-3;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        a[i] = threadIdx.x * 8 + i + seed;
        l[i] = a[i] * 3;
        w[i] = ((uint64_t) a[i] << 32) | (a[i] * 77u);
        d[i] = (double) a[i];
    }
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (KIND == 0 || KIND == 6 || KIND == 8 || KIND == 10 || KIND == 12) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 2 || KIND == 9) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; lop3.b32 %1, %1, hi, lo, 0x96; mul.wide.u32 %0, lo, %2;}" : "+l"(w[i]), "+r"(l[i]) : "r"(b));
            if (KIND == 3) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b));
            if (KIND == 4 || KIND == 7 || KIND == 8 || KIND == 9 || KIND == 12) asm volatile("fma.rm.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(dc));
            if (KIND == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dc));
            if (KIND == 6 || KIND == 7 || KIND == 10 || KIND == 11 || KIND == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(b), "r"(c));
            if (KIND == 10 || KIND == 11 || KIND == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x69;" : "+r"(l[i]) : "r"(c), "r"(b));
            if (KIND == 13) asm volatile("{.reg .u32 lo, hi; .reg .f64 t; cvt.rn.f64.u32 t, %0; mov.b64 {lo, hi}, t; lop3.b32 %0, %0, hi, lo, 0x96; }" : "+r"(a[i]));
            if (KIND == 14) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; add.cc.u32 lo, lo, %1; addc.u32 hi, hi, %2; mov.b64 %0, {lo, hi};}" : "+l"(w[i]) : "r"(b), "r"(c));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ l[i] ^ (uint32_t) w[i] ^ (uint32_t) (w[i] >> 32) ^ (uint32_t) __double2loint(d[i]) ^ (uint32_t) __double2hiint(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static int g_sms = 0;
template <int KIND> void run_pipe(const char* name, double ops)
{
    const int iters = 8192, threads = 256, blocks = g_sms * 8;
    uint32_t* out;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pipe<KIND><<<blocks, threads>>>(out, 1, iters);
    cudaEventRecord(e0);
    pipe<KIND><<<blocks, threads>>>(out, 2, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double) blocks * threads * iters * 8 * ops;
    printf("pipe  %-44s %8.3f ms  %7.2f Tops/s   (%.1f lane-ops/clk/SM at 1.90 GHz)\n", name, ms, total / ms / 1e9,
           total / ms / 1e9 * 1e12 / (g_sms * 1.90e9) );
    cudaFree(out);
}


template <typename M, typename TW, bool INV, bool RENORM, int MINB = 2> void run_bfly(const char* name, uint64_t p, const uint64_t* h_w)
{
    const int iters = 2048, threads = 256, blocks = g_sms * 8;
    uint64_t* out;
    TW* tw;
    const size_t n = (size_t) threads * blocks * 16;
    cudaMalloc(&out, n * 8);
    cudaMemset(out, 0x5a, n * 8);
    TW* h = new TW[64 * 15];
    for (int i = 0; i < 64 * 15; i++) h[i] = make_tw<TW>(h_w[i], p);
    cudaMalloc(&tw, sizeof(TW) * 64 * 15);
    cudaMemcpy(tw, h, sizeof(TW) * 64 * 15, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<M, TW, INV, RENORM, MINB><<<blocks, threads>>>(out, tw, p, iters);
    cudaEventRecord(e0);
    probe<M, TW, INV, RENORM, MINB><<<blocks, threads>>>(out, tw, p, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double) blocks * threads * iters * 32.0;
    printf("bfly  %-44s %8.3f ms  %7.3f T butterflies/s  -> %.2f M NTT/s (N=2^16)  %s\n", name, ms, total / ms / 1e9,
           total / ms / 1e9 * 1e12 / 524288.0 / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(out);
    cudaFree(tw);
    delete[] h;
}

// ---------------------------------------------------------------- correctness of every variant's multiply
template <typename M, typename TW> __global__ void check_mul(const uint64_t* ys, const TW* tws, uint64_t* rs, uint64_t p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const M m(p);
    rs[i] = m.mul(ys[i], tws[i]);
}
template <typename M, typename TW> void run_check(const char* name, uint64_t p, uint64_t bound_mult)
{
    const int n = 1 << 20;
    uint64_t *ys = new uint64_t[n], *ws = new uint64_t[n], *rs = new uint64_t[n];
    TW* tws = new TW[n];
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < n; i++)
    {
        ys[i] = rnd();
        ws[i] = rnd() % p;
        if (i < 64) { ys[i] = (i & 1) ? ~0ull : 0ull; if (i & 2) ys[i] = (uint64_t) 0xffffffffu << ((i & 4) ? 32 : 0); }
        if (i >= 64 && i < 128) ws[i] = (i & 1) ? p - 1 : (i & 2 ? 1 : 0);
        tws[i] = make_tw<TW>(ws[i], p);
    }
    uint64_t *dy, *dr;
    TW* dt;
    cudaMalloc(&dy, n * 8); cudaMalloc(&dr, n * 8); cudaMalloc(&dt, n * sizeof(TW));
    cudaMemcpy(dy, ys, n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dt, tws, n * sizeof(TW), cudaMemcpyHostToDevice);
    check_mul<M, TW><<<n / 256, 256>>>(dy, dt, dr, p, n);
    cudaMemcpy(rs, dr, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0; uint64_t maxq = 0;
    for (int i = 0; i < n; i++)
    {
        const unsigned __int128 prod = (unsigned __int128) ys[i] * ws[i];
        const uint64_t want = (uint64_t) (prod % p);
        if (rs[i] % p != want || rs[i] >= bound_mult * p) { if (bad < 5) printf("   BAD i=%d y=%llu w=%llu got=%llu want=%llu\n", i, (unsigned long long) ys[i], (unsigned long long) ws[i], (unsigned long long) rs[i], (unsigned long long) want); bad++; }
        if (rs[i] / p > maxq) maxq = rs[i] / p;
    }
    printf("check %-44s %s  (max r/p = %llu, bound %llu)\n", name, bad ? "FAILED" : "ok", (unsigned long long) maxq, (unsigned long long) bound_mult);
    cudaFree(dy); cudaFree(dr); cudaFree(dt);
    delete[] ys; delete[] ws; delete[] rs; delete[] tws;
}

#include "arith_variants.cuh"

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    g_sms = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, g_sms);
    run_pipe<0>("IMAD", 1);
    run_pipe<1>("IMAD.HI", 1);
    run_pipe<2>("IMAD.WIDE (no addend) + LOP3 (counted as 1)", 1);
    run_pipe<3>("mad.wide + 64-bit addend (see SASS)", 1);
    run_pipe<4>("DFMA.RM", 1);
    run_pipe<5>("DADD", 1);
    run_pipe<14>("add64 via add.cc/addc (2 ops)", 2);
    run_pipe<11>("LOP3 x2 (2 ops)", 2);
    run_pipe<6>("IMAD + LOP3 (2 ops)", 2);
    run_pipe<10>("IMAD + 2 LOP3 (3 ops)", 3);
    run_pipe<7>("DFMA + LOP3 (2 ops)", 2);
    run_pipe<8>("DFMA + IMAD (2 ops)", 2);
    run_pipe<9>("DFMA + IMAD.WIDE + LOP3 (counted as 2)", 2);
    run_pipe<12>("DFMA + IMAD + 2 LOP3 (4 ops)", 4);
    run_pipe<13>("I2F.F64.U32 + LOP3 (counted as 1)", 1);
    const uint64_t p = 576460756061519873ull;
    uint64_t hw[64 * 15];
    uint64_t s = 1234567;
    for (int i = 0; i < 64 * 15; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; hw[i] = s % p; }
    run_variants(p, hw);
    return 0;
}
