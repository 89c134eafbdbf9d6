// gpu_ntt_b200/csrc/modarith.cuh -- device modular arithmetic for the B200 NTT engine.
//
// The reference multiplies with a generic Barrett reduction on every butterfly
// (modular_arith.cuh:270-339 in the reference tree: three wide products + shifts, ~57 SASS
// arithmetic instructions per 64-bit butterfly on sm_100).  Here every twiddle w travels with
// a precomputed companion w' = floor(w * 2^BITS / p) (Shoup), so a butterfly multiply is
//      q = mulhi(w', v);   r = w*v - q*p   (mod 2^BITS),   r in [0, 2p)  for ANY v < 2^BITS
// and values are kept lazily in [0, 4p) between stages (Harvey), which needs 4p < 2^BITS,
// i.e. p < 2^62 (u64) / p < 2^30 (u32): exactly the reference's supported modulus range
// (modular_arith.cuh:66-67).  A final correction makes every output canonical in [0,p), so
// results are bit-identical to the reference's OPERATOR::mult/add/sub chain.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpuntt_b200
{

    template <typename T> struct Twiddle; // (w, w') pair as stored in the per-call companion table
    template <> struct __align__(16) Twiddle<uint64_t>
    {
        uint64_t w, wq;
    };
    template <> struct __align__(8) Twiddle<uint32_t>
    {
        uint32_t w, wq;
    };

    // ---------------------------------------------------------------- 64-bit
    __device__ __forceinline__ uint64_t mulhi(uint64_t a, uint64_t b) { return __umul64hi(a, b); }
    __device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

    // r = w*v - floor(w'*v / 2^BITS)*p  in [0, 2p), for any v.
    template <typename T> __device__ __forceinline__ T shoup_mul_lazy(T v, const Twiddle<T>& tw, T p)
    {
        T q = mulhi(tw.wq, v);
        return tw.w * v - q * p;
    }

    // x in [0, 2m) -> [0, m)
    template <typename T> __device__ __forceinline__ T csub(T x, T m) { return (x >= m) ? x - m : x; }

    // Cooley-Tukey (forward) butterfly, Harvey lazy form.  In: X, Y in [0,4p).  Out: [0,4p).
    // Replaces the reference's CooleyTukeyUnit (ntt.cuh:69-78).
    template <typename T>
    __device__ __forceinline__ void ct_butterfly(T& X, T& Y, const Twiddle<T>& tw, T p, T two_p)
    {
        T x = csub(X, two_p);
        T t = shoup_mul_lazy(Y, tw, p);
        X = x + t;
        Y = x - t + two_p;
    }

    // Gentleman-Sande (inverse) butterfly, lazy form.  In: X, Y in [0,2p).  Out: [0,2p).
    // Replaces the reference's GentlemanSandeUnit (ntt.cuh:80-92).
    template <typename T>
    __device__ __forceinline__ void gs_butterfly(T& X, T& Y, const Twiddle<T>& tw, T p, T two_p)
    {
        T s = X + Y;
        T d = X - Y + two_p;
        X = csub(s, two_p);
        Y = shoup_mul_lazy(d, tw, p);
    }

    // [0,4p) -> [0,p)
    template <typename T> __device__ __forceinline__ T canon4(T x, T p, T two_p)
    {
        return csub(csub(x, two_p), p);
    }

    // companion w' = floor(w * 2^BITS / p), w < p
    __device__ __forceinline__ uint64_t shoup_companion(uint64_t w, uint64_t p)
    {
        return (uint64_t) ((((unsigned __int128) w) << 64) / p);
    }
    __device__ __forceinline__ uint32_t shoup_companion(uint32_t w, uint32_t p)
    {
        return (uint32_t) ((((uint64_t) w) << 32) / p);
    }

} // namespace gpuntt_b200
