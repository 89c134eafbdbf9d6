// gpu_ntt_b200/csrc/modarith.cuh -- device modular arithmetic for the B200 NTT engine.
//
// The reference multiplies with a generic Barrett reduction on every butterfly
// (modular_arith.cuh:270-339 in the reference tree: three wide products + shifts, ~57 SASS
// arithmetic instructions per 64-bit butterfly on sm_100).  Here every twiddle w travels with
// a precomputed companion w' = floor(w * 2^BITS / p) (Shoup), so a butterfly multiply is
//      q = mulhi(w', v);   r = w*v - q*p   (mod 2^BITS),   r in [0, 2p)  for ANY v < 2^BITS
// and values are kept lazily reduced between stages (Harvey).  A final correction makes every
// output canonical in [0,p), so results are bit-identical to the reference's
// OPERATOR::mult/add/sub chain.
//
// Two arithmetic policies (struct Mod<T, FAST>):
//   exact (FAST=false): q exact; values in [0,4p) forward / [0,2p) inverse; needs 4p < 2^BITS,
//         i.e. the reference's whole supported modulus range (p < 2^62 / p < 2^30,
//         modular_arith.cuh:66-67).
//   fast  (FAST=true, 64-bit only, p < 2^60.9): B200's IMAD.WIDE / IMAD.HI issue at half the
//         rate of a 32-bit IMAD (tools/microbench.cu), and everything a 64-bit butterfly does is
//         bound by that pipe, so the quotient is taken from three partial products only
//             q~ = a1*y1 + hi32(a1*y0) + hi32(a0*y1)   in {Q-2, Q-1, Q},  r = w*y - q~*p in [0,4p)
//         and r is accumulated with mad chains against -p (no separate subtract).  Values live
//         in [0, 8p + 2^32) forward (range test on the high word only) / [0,4p) inverse.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpuntt_b200
{

    template <typename T> struct Twiddle; // (w, w') pair
    template <> struct __align__(16) Twiddle<uint64_t>
    {
        uint64_t w, wq;
    };
    template <> struct __align__(8) Twiddle<uint32_t>
    {
        uint32_t w, wq;
    };

    __device__ __forceinline__ uint64_t mulhi(uint64_t a, uint64_t b) { return __umul64hi(a, b); }
    __device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

    // x in [0, 2m) -> [0, m)
    template <typename T> __device__ __forceinline__ T csub(T x, T m) { return (x >= m) ? x - m : x; }

    // companion w' = floor(w * 2^BITS / p), w < p
    __host__ __device__ __forceinline__ uint64_t shoup_companion(uint64_t w, uint64_t p)
    {
        return (uint64_t) ((((unsigned __int128) w) << 64) / p);
    }
    __host__ __device__ __forceinline__ uint32_t shoup_companion(uint32_t w, uint32_t p)
    {
        return (uint32_t) ((((uint64_t) w) << 32) / p);
    }

    // ------------------------------------------------------------------ exact policy
    template <typename T, bool FAST> struct Mod
    {
        T p, two_p;
        __device__ __forceinline__ explicit Mod(T p_) : p(p_), two_p(p_ + p_) {}

        // r = w*v - floor(w'*v / 2^BITS)*p  in [0, 2p), any v
        __device__ __forceinline__ T mul(T v, const Twiddle<T>& tw) const
        {
            T q = mulhi(tw.wq, v);
            return tw.w * v - q * p;
        }
        // Cooley-Tukey butterfly (replaces CooleyTukeyUnit, ntt.cuh:69-78 of the reference).
        // In/out: [0,4p).
        __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<T>& tw) const
        {
            T x = csub(X, two_p);
            T t = mul(Y, tw);
            X = x + t;
            Y = x - t + two_p;
        }
        // Gentleman-Sande butterfly (replaces GentlemanSandeUnit, ntt.cuh:80-92). In/out: [0,2p).
        __device__ __forceinline__ void gs(T& X, T& Y, const Twiddle<T>& tw) const
        {
            T s = X + Y;
            T d = X - Y + two_p;
            X = csub(s, two_p);
            Y = mul(d, tw);
        }
        // forward lazy value -> canonical
        __device__ __forceinline__ T canon_fwd(T x) const { return csub(csub(x, two_p), p); }
        // inverse lazy value * n^-1 -> canonical
        __device__ __forceinline__ T canon_inv(T x, const Twiddle<T>& ninv) const { return csub(mul(x, ninv), p); }
    };

    // ------------------------------------------------------------------ fast policy (u64, p < 2^60.9)
    template <> struct Mod<uint64_t, true>
    {
        using T = uint64_t;
        T p, four_p;
        uint32_t n0, n1, f0, f1; // -p mod 2^64 ; 4p
        __device__ __forceinline__ explicit Mod(T p_) : p(p_), four_p(4 * p_)
        {
            const T np = 0 - p_;
            n0 = (uint32_t) np;
            n1 = (uint32_t) (np >> 32);
            f0 = (uint32_t) four_p;
            f1 = (uint32_t) (four_p >> 32);
        }

        // r = w*y - q~*p in [0,4p): 3 IMAD.WIDE + 2 IMAD.HI + 4 IMAD, no subtract
        __device__ __forceinline__ T mul(T y, const Twiddle<T>& tw) const
        {
            const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
            const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
            const uint32_t a0 = (uint32_t) tw.wq, a1 = (uint32_t) (tw.wq >> 32);
            uint32_t r0, r1;
            asm("{\n\t"
                ".reg .u32 h1, h2, q0, q1, t0, t1;\n\t"
                ".reg .u64 q, r;\n\t"
                "mul.hi.u32 h1, %6, %3;\n\t"  // hi(a1*y0)
                "mul.hi.u32 h2, %7, %2;\n\t"  // hi(a0*y1)
                "mul.wide.u32 q, %6, %2;\n\t" // a1*y1
                "mov.b64 {q0, q1}, q;\n\t"
                "add.cc.u32 q0, q0, h1;\n\t"
                "addc.u32 q1, q1, 0;\n\t"
                "add.cc.u32 q0, q0, h2;\n\t"
                "addc.u32 q1, q1, 0;\n\t"
                "mul.wide.u32 r, %4, %3;\n\t" // w0*y0
                "mov.b64 {t0, t1}, r;\n\t"
                "mad.lo.u32 t1, %5, %3, t1;\n\t" // w1*y0
                "mad.lo.u32 t1, %4, %2, t1;\n\t" // w0*y1
                "mov.b64 r, {t0, t1};\n\t"
                "mad.wide.u32 r, q0, %8, r;\n\t" // q0*n0
                "mov.b64 {t0, t1}, r;\n\t"
                "mad.lo.u32 t1, q1, %8, t1;\n\t" // q1*n0
                "mad.lo.u32 %1, q0, %9, t1;\n\t" // q0*n1
                "mov.u32 %0, t0;\n\t"
                "}"
                : "=r"(r0), "=r"(r1)
                : "r"(y1), "r"(y0), "r"(w0), "r"(w1), "r"(a1), "r"(a0), "r"(n0), "r"(n1));
            return ((T) r1 << 32) | r0;
        }
        // In/out: [0, 8p + 2^32).  The range test looks at the high word only: x1 > hi32(4p)
        // implies x >= 4p; otherwise x < 4p + 2^32.
        __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T x = ((uint32_t) (X >> 32) > f1) ? X - four_p : X;
            const T t = mul(Y, tw);
            X = x + t;
            Y = x - t + four_p;
        }
        // In/out: [0,4p)  (exact range test: a high-word-only test would double the slack per stage)
        __device__ __forceinline__ void gs(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T s = X + Y;
            const T d = X - Y + four_p;
            X = csub(s, four_p);
            Y = mul(d, tw);
        }
        __device__ __forceinline__ T canon_fwd(T x) const
        {
            x = csub(csub(x, four_p), four_p); // [0, 8p+2^32) -> [0,4p)
            return csub(csub(x, p + p), p);
        }
        __device__ __forceinline__ T canon_inv(T x, const Twiddle<T>& ninv) const
        {
            return csub(csub(mul(x, ninv), p + p), p);
        }
    };

    // largest modulus the fast policy accepts: 8p + 2^32 < 2^64 with margin
    constexpr uint64_t kFastModulusLimit = (1ull << 60) + (1ull << 59);

} // namespace gpuntt_b200
