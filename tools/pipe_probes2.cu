// raw pipe probes, second batch: multiply forms with 64-bit addends / carry-out, three-operand adds, SEL
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND> __global__ void __launch_bounds__(256, 2) pipe(uint32_t* out, uint32_t seed, int iters)
{
    uint32_t a[8], l[8], m[8], b = seed | 1, c = seed * 3 + 7, d = seed * 5 + 11;
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 8 + i + seed; l[i] = a[i] * 3; m[i] = a[i] * 7; w[i] = ((uint64_t) a[i] << 32) | (a[i] * 77u); }
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            // 0: IMAD.WIDE no addend (baseline)  1: WIDE + 64-bit addend (natural pair)  2: mad.lo.cc+madc.hi.cc (64-bit mad with carry-out, both halves used)
            // 3: same, only high half + carry used (ptxas: IMAD.HI with 64-bit addend?)  4: IMAD.HI no addend  5: IADD3 3 regs  6: IADD3.X chain 3 regs  7: SEL
            if (KIND == 0) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %1;}" : "+l"(w[i]) : "r"(b));
            if (KIND == 1) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b));
            if (KIND == 2) asm volatile("{.reg .u32 lo, hi, cy; mov.b64 {lo, hi}, %0; mad.lo.cc.u32 lo, %2, %1, lo; madc.hi.cc.u32 hi, %2, %1, hi; addc.u32 %2, %2, 0; mov.b64 %0, {lo, hi};}" : "+l"(w[i]), "+r"(a[i]), "+r"(l[i]) : "r"(b));
            if (KIND == 3) asm volatile("{.reg .u32 lo, hi, d0; mov.b64 {lo, hi}, %0; mad.lo.cc.u32 d0, %2, %1, lo; madc.hi.cc.u32 hi, %2, %1, hi; addc.u32 %2, %2, 0; mov.b64 %0, {lo, hi};}" : "+l"(w[i]), "+r"(a[i]), "+r"(l[i]) : "r"(b));
            if (KIND == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (KIND == 5) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(l[i]), "r"(m[i]));
            if (KIND == 6) asm volatile("{add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3; add.cc.u32 %0, %0, %4; addc.u32 %1, %1, %5;}" : "+r"(a[i]), "+r"(l[i]) : "r"(m[i]), "r"(b), "r"(c), "r"(d));
            if (KIND == 7) asm volatile("{.reg .pred P; setp.gt.u32 P, %0, %2; selp.b32 %1, %3, %1, P; selp.b32 %0, %1, %0, P;}" : "+r"(a[i]), "+r"(l[i]) : "r"(c), "r"(d));
            // 8: IMAD with all-distinct register operands (no reuse possible)
            if (KIND == 8) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(l[i]), "r"(m[i]));
            // 9: IMAD distinct + IADD3 distinct (co-issue, register bandwidth)
            if (KIND == 9) asm volatile("{mad.lo.u32 %0, %1, %2, %0; .reg .u32 t; add.u32 t, %1, %2; add.u32 %1, t, %0;}" : "+r"(a[i]), "+r"(l[i]) : "r"(m[i]));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ l[i] ^ m[i] ^ (uint32_t) w[i] ^ (uint32_t) (w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
static int g_sms = 0;
template <int KIND> void run_pipe(const char* name, double ops)
{
    const int iters = 8192, threads = 256, blocks = g_sms * 8;
    uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe<KIND><<<blocks, threads>>>(out, 1, iters);
    cudaEventRecord(e0);
    pipe<KIND><<<blocks, threads>>>(out, 2, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = (double) blocks * threads * iters * 8;
    printf("pipe2 %-2d %-58s %8.3f ms  %6.2f cycles per warp-group per SMSP (1.965 GHz)\n", KIND, name, ms, ms * 1e-3 * 1.965e9 * g_sms * 4 / (groups / 32));
    cudaFree(out);
}
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); g_sms = pr.multiProcessorCount;
    run_pipe<0>("IMAD.WIDE no addend (+LOP3)", 1);
    run_pipe<1>("IMAD.WIDE + 64-bit addend (natural pair)", 1);
    run_pipe<2>("mad.lo.cc + madc.hi.cc + addc (64-bit mad, carry out)", 1);
    run_pipe<3>("same, high half + carry only", 1);
    run_pipe<4>("IMAD.HI no addend", 1);
    run_pipe<5>("add3 (two adds -> IADD3 3 regs?)", 1);
    run_pipe<6>("64-bit three-operand add (2x add.cc/addc)", 1);
    run_pipe<7>("ISETP + 2 SEL", 1);
    run_pipe<8>("IMAD three distinct registers", 1);
    run_pipe<9>("IMAD + add3, distinct registers", 1);
    return 0;
}
