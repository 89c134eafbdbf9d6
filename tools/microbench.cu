// tools/microbench.cu -- integer-pipe throughput probes for B200 (informational roof for DESIGN.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int KIND> __global__ void probe(uint32_t* out, uint32_t seed, long long* cyc)
{
    uint32_t a[8], b = seed | 1, c = seed * 3 + 7;
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 8 + i + seed; w[i] = ((uint64_t) a[i] << 32) | (a[i] * 77u); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (KIND == 0) a[i] = a[i] * b + c;                                 // IMAD
            if (KIND == 1) a[i] = __umulhi(a[i], b) + c;                        // IMAD.HI
            if (KIND == 2) w[i] = (uint64_t) (uint32_t) w[i] * b + w[i];        // IMAD.WIDE.U32 with 64-bit addend
            if (KIND == 3) a[i] = (a[i] + a[(i + 1) & 7]) ^ c;                              // IADD3 + LOP3 (alu)
            if (KIND == 4) w[i] = w[i] + w[(i + 1) & 7];            // 64-bit add: IADD3 + IADD3.X
            if (KIND == 5) { a[i] = a[i] * b + c; w[i] = w[i] + w[(i + 1) & 7]; } // IMAD + 2 ALU mixed
            if (KIND == 6) w[i] = __umul64hi(w[i], ((uint64_t) b << 32) | c) + w[i]; // full mulhi64
            if (KIND == 7) w[i] = w[i] * (((uint64_t) b << 32) | c) + 12345;    // mullo64
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ (uint32_t) w[i] ^ (uint32_t) (w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND> void run(const char* name, double ops_per_iter)
{
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t* out; long long* cyc;
    const int threads = 1024, blocks = sms * 2;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND><<<blocks, threads>>>(out, 1, cyc);
    cudaEventRecord(e0);
    probe<KIND><<<blocks, threads>>>(out, 2, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double total = (double) blocks * threads * ITERS * 8 * ops_per_iter;
    // 2 resident blocks of 1024 threads per SM run concurrently: per-SM lane-ops per clock of block 0's span
    double per_clk_sm = (double) 2 * threads * ITERS * 8 * ops_per_iter / (double) c;
    printf("%-34s %8.3f ms  %7.2f Tops/s  %6.1f lane-ops/clk/SM  (eff. clock %.0f MHz)\n", name, ms, total / ms / 1e9,
           per_clk_sm, (double) c / ms / 1e3);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("IMAD (32-bit mad.lo)", 1);
    run<1>("IMAD.HI.U32", 1);
    run<2>("IMAD.WIDE.U32 (+64-bit addend)", 1);
    run<3>("IADD3+LOP3 (2 alu ops)", 2);
    run<4>("64-bit add (IADD3+IADD3.X)", 2);
    run<5>("IMAD + 64-bit add (1 fma + 2 alu)", 3);
    run<6>("umul64hi + add64 (as 1 op)", 1);
    run<7>("mullo64 + add (as 1 op)", 1);
    return 0;
}
