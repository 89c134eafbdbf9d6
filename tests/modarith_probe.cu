// tests/modarith_probe.cu -- host-side check of the __host__ __device__ helpers of gpu_ntt_b200/csrc/modarith.cuh that the
// kernels rely on for exactness: div_128_by_64 / recip_mu64 (per-segment reciprocal of the RNS kernels) against the compiler's
// own 128-bit division, over every bit length a modulus can have.  Built and run by tests/test_modarith_host.py (no GPU needed).
#include <cstdint>
#include <cstdio>
#include <random>

#include "modarith.cuh"

using namespace gpuntt_b200;

int main()
{
    std::mt19937_64 g(12345);
    long bad = 0, n = 0;
    for (int it = 0; it < 4000000; it++)
    {
        const int bits = 2 + (int) (g() % 61); // 2 .. 62
        uint64_t p = (g() >> (64 - bits)) | (1ull << (bits - 1)) | 1ull;
        if (bits == 2) p = 3;
        if ((p & (p - 1)) == 0) continue;
        const unsigned __int128 m = (((unsigned __int128) 1) << (63 + bits)) / p;
        const uint64_t want = (m >> 64) ? ~0ull : (uint64_t) m;
        if (recip_mu64(p, bits) != want) bad++;
        const uint64_t u1 = g() % p, u0 = g();
        const unsigned __int128 U = ((unsigned __int128) u1 << 64) | u0;
        if (div_128_by_64(u1, u0, p) != (uint64_t) (U / p)) bad++;
        const uint64_t w = g() % p;
        if (shoup_companion(w, p) != div_128_by_64(w, 0, p)) bad++;
        n++;
    }
    // the edges: the reference's pooled primes and the policy boundaries
    const uint64_t edge[] = {576460756061519873ull, 576460752303415297ull, 576460753175838721ull, (1ull << 40) + 15, (1ull << 60) - (1ull << 31) - 1,
                             2305843009213693951ull, 4611686018427387847ull, 3, 5, 469762049ull};
    for (uint64_t p : edge)
    {
        const int bits = 64 - __builtin_clzll(p);
        const unsigned __int128 m = (((unsigned __int128) 1) << (63 + bits)) / p;
        if (recip_mu64(p, bits) != (uint64_t) m) bad++;
        n++;
    }
    printf("checked=%ld bad=%ld\n", n, bad);
    return bad != 0;
}
