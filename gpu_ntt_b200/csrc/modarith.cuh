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
// Arithmetic policies:
//   exact (Mod<T,false>): q exact; values in [0,4p) forward / [0,2p) inverse; needs 4p < 2^BITS,
//         i.e. the reference's whole supported modulus range (p < 2^62 / p < 2^30,
//         modular_arith.cuh:66-67).
//   fast  (Mod<u64,true>, p < 1.25 * 2^60): B200's IMAD.WIDE / IMAD.HI issue at half the rate of a
//         32-bit IMAD and everything a 64-bit butterfly does is bound by that pipe, so the quotient
//         is taken from three partial products only (q~ in {Q-1, Q}, r in [0,3p) for ANY 64-bit
//         input) and r is accumulated with mad chains against -p.  Forward values live in
//         [0, 8p + 2^32) (range test on the high word only; needs p >= 2^36), inverse values in
//         [0, 4p) (exact range test).
//   F60   (ModF60, forward, 2^40 <= p < 2^60 - 2^31): the tuned kernels' policy -- range correction
//         every other stage, bare add/subtract for twiddle-1 butterflies, one-step canonicalisation.
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

    // x in [0, 2m) -> [0, m).  32-bit: min(x, x - m) -- the difference wraps above x exactly when x < m -- is one subtraction and one
    // VIMNMX instead of a compare and a select (the 32-bit kernels are bound by instruction issue, not by the multiplier).
    template <typename T> __device__ __forceinline__ T csub(T x, T m) { return (x >= m) ? x - m : x; }
    template <> __device__ __forceinline__ uint32_t csub<uint32_t>(uint32_t x, uint32_t m) { return min(x, x - m); }

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
        // In/out: [0,4p).  Three-operand sums (see the fast policy below for why).
        __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T t = mul(Y, tw);
            if constexpr (sizeof(T) == 4)
            {
                const T x = csub(X, two_p); // [0,4p) -> [0,2p)
                Y = x + two_p - t;
                X = x + t;
            }
            else
            {
                const bool P = X >= two_p;
                const T g = P ? T(0) - two_p : T(0), k = P ? T(0) : two_p;
                const T Xn = X + g + t;
                Y = X + k - t;
                X = Xn;
            }
        }
        // Gentleman-Sande butterfly (replaces GentlemanSandeUnit, ntt.cuh:80-92). In/out: [0,2p).
        __device__ __forceinline__ void gs(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T d = X + two_p - Y;
            const T s = X + Y;
            if constexpr (sizeof(T) == 4)
                X = csub(s, two_p);
            else
            {
                const T g = (s >= two_p) ? T(0) - two_p : T(0);
                X = X + Y + g;
            }
            Y = mul(d, tw);
        }
        // Gentleman-Sande butterfly with twiddle 1 (the top stages of an inverse X^N-1 transform): no multiply
        __device__ __forceinline__ void gs_one(T& X, T& Y) const
        {
            const T d = X + two_p - Y;
            const T s = X + Y;
            if constexpr (sizeof(T) == 4)
                X = csub(s, two_p);
            else
            {
                const T g = (s >= two_p) ? T(0) - two_p : T(0);
                X = X + Y + g;
            }
            Y = csub(d, two_p);
        }
        // forward lazy value -> canonical
        __device__ __forceinline__ T canon_fwd(T x) const { return csub(csub(x, two_p), p); }
        // inverse lazy value * n^-1 -> canonical
        __device__ __forceinline__ T canon_inv(T x, const Twiddle<T>& ninv) const { return csub(mul(x, ninv), p); }
        // inverse lazy value (n^-1 already folded into the twiddles of the last round, see fast_round) -> canonical
        __device__ __forceinline__ T canon_lazy_inv(T x) const { return csub(x, p); }
    };

    // ------------------------------------------------------------------ fast policy (u64, p < 2^60.5)
    // Everything below is shaped by what tools/arith_bench.cu measured on B200: all integer multiplies
    // issue on the "fmaheavy" pipe (IMAD 64 lanes/clk/SM, IMAD.WIDE / IMAD.HI 32), additions on the alu
    // pipe at 128 lanes/clk/SM, and ptxas likes to move carry additions onto the multiplier pipe
    // (IMAD.X) -- so the multiply is one PTX block whose partial sums ride in IMAD / IMAD.WIDE addends,
    // and the range correction is a predicated subtract keyed on the high word only.
    template <> struct Mod<uint64_t, true>
    {
        using T = uint64_t;
        T p, four_p, neg_four_p;
        uint32_t n0, n1, f0, f1; // -p mod 2^64 ; 4p
        __device__ __forceinline__ explicit Mod(T p_) : p(p_), four_p(4 * p_), neg_four_p(0 - 4 * p_)
        {
            const T np = 0 - p_;
            n0 = (uint32_t) np;
            n1 = (uint32_t) (np >> 32);
            f0 = (uint32_t) four_p;
            f1 = (uint32_t) (four_p >> 32);
        }

        // r = w*y - q~*p in [0,3p) for ANY 64-bit y;  q~ = a1*y1 + floor((a1*y0 + a0*y1) / 2^32) in {Q-1, Q}
        // (only the a0*y0 partial product of the quotient is dropped).  4 IMAD.WIDE + 1 IMAD.HI + 4 IMAD on the
        // multiplier pipe = 28 issue cycles per warp; the formulation is what ptxas turns into exactly that with
        // NO extra IMAD.X / IMAD.MOV / IMAD.IADD (see tools/bfly_lab.cu): the cross terms are a 64-bit
        // multiply-add with carry-out, the high word of r is an IMAD chain that starts from the high word of
        // q0*n0 + w0*y0, so the product needs two additions in total (q += cross terms).
        __device__ __forceinline__ T mul(T y, const Twiddle<T>& tw) const
        {
            const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
            const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
            const uint32_t a0 = (uint32_t) tw.wq, a1 = (uint32_t) (tw.wq >> 32);
            uint32_t r0, r1;
            asm("{\n\t"
                ".reg .u32 q0, q1, c0, c1, d0, d1;\n\t"
                ".reg .u64 q, A, B, C;\n\t"
                "mul.wide.u32 C, %6, %2;\n\t"        // a1*y0
                "mov.b64 {c0, c1}, C;\n\t"
                "mul.wide.u32 q, %6, %3;\n\t"        // a1*y1
                "mov.b64 {q0, q1}, q;\n\t"
                "mad.lo.cc.u32 d0, %7, %3, c0;\n\t"  // a0*y1 + a1*y0 (64-bit, carry out)
                "madc.hi.cc.u32 d1, %7, %3, c1;\n\t"
                "addc.u32 q1, q1, 0;\n\t"
                "add.cc.u32 q0, q0, d1;\n\t"
                "addc.u32 q1, q1, 0;\n\t"
                "mul.wide.u32 A, %4, %2;\n\t"        // w0*y0
                "mad.wide.u32 B, q0, %8, A;\n\t"     // q0*n0 + w0*y0
                "mov.b64 {%0, %1}, B;\n\t"
                "mad.lo.u32 %1, %5, %2, %1;\n\t"     // + w1*y0
                "mad.lo.u32 %1, %4, %3, %1;\n\t"     // + w0*y1
                "mad.lo.u32 %1, q1, %8, %1;\n\t"     // + q1*n0
                "mad.lo.u32 %1, q0, %9, %1;\n\t"     // + q0*n1
                "}"
                : "=r"(r0), "=r"(r1)
                : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(a1), "r"(a0), "r"(n0), "r"(n1));
            return ((T) r1 << 32) | r0;
        }
        // x >= 4p + 2^32 is impossible afterwards: hi32(x) > hi32(4p) implies x >= 4p.
        __device__ __forceinline__ T csub_hi(T x) const
        {
            uint32_t x0 = (uint32_t) x, x1 = (uint32_t) (x >> 32);
            asm("{\n\t"
                ".reg .pred P;\n\t"
                "setp.gt.u32 P, %1, %3;\n\t"
                "@P sub.cc.u32 %0, %0, %2;\n\t"
                "@P subc.u32 %1, %1, %3;\n\t"
                "}"
                : "+r"(x0), "+r"(x1)
                : "r"(f0), "r"(f1));
            return ((T) x1 << 32) | x0;
        }
        // Forward values live in [0, 8p + 2^32).  x = X - 4p when hi32(X) > hi32(4p); X' = x + t, Y' = x - t + 4p are
        // written as three-operand 64-bit sums (X + g + t, X + k - t with g = -4p or 0, k = 0 or 4p): ptxas then
        // emits IADD3 / IADD3.X pairs on the alu pipe and cannot move the carry additions onto the multiplier pipe.
        __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T t = mul(Y, tw);
            uint32_t xh;
            asm("{\n\t.reg .u32 lo;\n\tmov.b64 {lo, %0}, %1;\n\t}" : "=r"(xh) : "l"(X)); // opaque: keeps a 32-bit compare
            const bool P = xh > f1;
            const T g = P ? neg_four_p : T(0), k = P ? T(0) : four_p;
            const T Xn = X + g + t;
            Y = X + k - t;
            X = Xn;
        }
        // butterfly with twiddle 1 (first stages of X^N-1 transforms): no multiply, Y only range-reduced
        __device__ __forceinline__ void ct_one(T& X, T& Y) const
        {
            const T x = csub_hi(X);
            const T t = csub(csub_hi(Y), four_p); // [0, 8p + 2^32) -> [0, 4p), same range as a product
            X = x + t;
            Y = x - t + four_p;
        }
        // Inverse values live in [0, 4p): the range test of the sum is an exact 64-bit compare (a high-word test would
        // leave slack that DOUBLES per stage, because both inputs of a Gentleman-Sande sum are lazy), and the three
        // 64-bit results are three-operand sums so that they stay on the alu pipe.
        __device__ __forceinline__ void gs(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T d = X + four_p - Y;
            const T s = X + Y;
            const T g = (s >= four_p) ? neg_four_p : T(0);
            X = X + Y + g;
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
        // twiddle 1: the difference only needs its range back in [0, 4p)
        __device__ __forceinline__ void gs_one(T& X, T& Y) const
        {
            const T d = X + four_p - Y;
            const T s = X + Y;
            const T g = (s >= four_p) ? neg_four_p : T(0);
            X = X + Y + g;
            Y = csub(d, four_p);
        }
        __device__ __forceinline__ T canon_lazy_inv(T x) const { return csub(csub(x, p + p), p); } // [0, 4p) -> [0, p)
    };

    // ------------------------------------------------------------------ "F60" forward policy (u64, 2^40 <= p < 2^60 - 2^31)
    // What bounds the 64-bit butterfly on B200 is not one pipe but three things at once -- multiplier-pipe cycles,
    // alu-pipe cycles and register-file operand reads (tools/pipe_probes2.cu: an IMAD and an IADD3 with distinct
    // registers co-issue at 3.3 cycles per pair, not 2) -- so the forward path also trims the alu side:
    //   * values live in [0, 12p + 2^32) at round boundaries and the range correction (subtract 8p when the HIGH
    //     WORD exceeds hi32(8p)) runs on every OTHER stage only: stage kind A adds at most 4p (16p + 2^32 < 2^64),
    //     kind B brings the X input back under 8p + 2^32 first;
    //   * the first round of a cyclic transform sees canonical inputs, so its twiddle-1 butterflies are a bare
    //     add / subtract with a constant multiple of p and nothing else (add_sub);
    //   * the final correction is one Barrett step from the top 32 bits (q~ in {q-1, q}) and one conditional
    //     subtraction instead of four.
    struct ModF60 : Mod<uint64_t, true>
    {
        T neg_eight_p;
        uint32_t e1;     // hi32(8p)
        uint32_t red_m;  // floor(2^59 / ((p >> red_sh) + 1))
        int red_sh;      // bit length of p - 28
        __device__ __forceinline__ explicit ModF60(T p_) : Mod<uint64_t, true>(p_), neg_eight_p(0 - 8 * p_)
        {
            e1 = (uint32_t) ((8 * p_) >> 32);
            red_sh = (64 - __clzll((long long) p_)) - 28;
            const uint32_t pt = (uint32_t) (p_ >> red_sh) + 1u;
            red_m = (uint32_t) ((1ull << 59) / pt);
        }
        // kind A: no correction.  In: X < 12p + 2^32, any Y.  Out: < 16p + 2^32.
        __device__ __forceinline__ void ctA(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T t = mul(Y, tw);
            Y = X + four_p - t;
            X = X + t;
        }
        // kind B: x = X - 8p when hi32(X) > hi32(8p).  In: X < 16p + 2^32.  Out: < 12p + 2^32.
        __device__ __forceinline__ void ctB(T& X, T& Y, const Twiddle<T>& tw) const
        {
            const T t = mul(Y, tw);
            uint32_t xh;
            asm("{\n\t.reg .u32 lo;\n\tmov.b64 {lo, %0}, %1;\n\t}" : "=r"(xh) : "l"(X)); // opaque: keeps a 32-bit compare
            const bool P = xh > e1;
            const T g = P ? neg_eight_p : T(0), k = P ? neg_four_p : four_p;
            const T Xn = X + g + t;
            Y = X + k - t;
            X = Xn;
        }
        // x - 8p when hi32(x) > hi32(8p): [0, 16p + 2^32) -> [0, 8p + 2^32)
        __device__ __forceinline__ T csub8_hi(T x) const
        {
            uint32_t xh;
            asm("{\n\t.reg .u32 lo;\n\tmov.b64 {lo, %0}, %1;\n\t}" : "=r"(xh) : "l"(x));
            return xh > e1 ? x + neg_eight_p : x;
        }
        // twiddle-1 butterfly on values below K (K a multiple of p): X' = X + Y, Y' = X - Y + K, both below 2K
        __device__ __forceinline__ void add_sub(T& X, T& Y, T K) const
        {
            const T s = X + Y;
            Y = X + K - Y;
            X = s;
        }
        // bound (multiple of p) of the values before stage it of a first cyclic round with canonical inputs
        // in_bound: the inputs are below in_bound * p (1: canonical; 2: the 4-step column phase leaves values below 2p)
        __device__ __forceinline__ T triv_bound(int it, int in_bound = 1) const
        {
            return (T) (it == 0 ? in_bound : it == 1 ? 2 * in_bound : (in_bound == 1 ? 6 : 8)) * p;
        }
        // [0, 13p) -> [0, p)
        __device__ __forceinline__ T canon_fwd(T x) const
        {
            const uint32_t xt = __funnelshift_rc((uint32_t) x, (uint32_t) (x >> 32), red_sh);
            const uint32_t q = __umulhi(xt, red_m) >> 27;
            T r = (T) q * n0 + x;                 // x - q*p, low word
            r += (T) (q * n1) << 32;
            return csub(r, p);
        }
    };
    // ------------------------------------------------------------------ lazy forward policy for 32-bit data (p <= 2^29)
    // Same idea as F60 with an exact quotient (one IMAD.HI): products in [0,2p), values below 6p at round
    // boundaries and 8p <= 2^32 inside, the range correction (-4p, exact 32-bit compare) on every other stage.
    struct ModL32 : Mod<uint32_t, false>
    {
        using T = uint32_t;
        T four_p, red_m; // red_m = floor(2^32 / p)
        __device__ __forceinline__ explicit ModL32(T p_) : Mod<uint32_t, false>(p_), four_p(4 * p_), red_m((T) ((1ull << 32) / p_)) {}
        __device__ __forceinline__ void ctA(T& X, T& Y, const Twiddle<T>& tw) const // in: X < 6p; out < 8p
        {
            const T t = mul(Y, tw);
            Y = X + two_p - t;
            X = X + t;
        }
        __device__ __forceinline__ void ctB(T& X, T& Y, const Twiddle<T>& tw) const // in: X < 8p; out < 6p
        {
            const T t = mul(Y, tw);
            const T x = csub(X, four_p); // [0,8p) -> [0,4p)
            Y = x + two_p - t;
            X = x + t;
        }
        __device__ __forceinline__ void add_sub(T& X, T& Y, T K) const
        {
            const T s = X + Y;
            Y = X + K - Y;
            X = s;
        }
        // bound (multiple of p) of the values before stage it of a first cyclic round with canonical inputs
        __device__ __forceinline__ T triv_bound(int it, int = 1) const { return (it == 0 ? 1 : it == 1 ? 2 : 4) * p; }
        __device__ __forceinline__ T csub8_hi(T x) const { return x; } // (in_bound > 1 never reaches the 32-bit kernels)
        __device__ __forceinline__ T canon_fwd(T x) const // [0, 8p) -> [0, p)
        {
            const T q = __umulhi(x, red_m); // in {floor(x/p) - 1, floor(x/p)}
            return csub(x - q * p, p);
        }
    };
    constexpr uint32_t kL32ModulusLimit = (1u << 29) + 1; // 8p <= 2^32

    constexpr uint64_t kF60ModulusMin = 1ull << 40;
    constexpr uint64_t kF60ModulusLimit = (1ull << 60) - (1ull << 31);

    // companion w' = floor(w * 2^64 / p) without a 128-bit division: mu = floor(2^(63 + bits) / p) (host-computed,
    // bits = bit length of p), estimate (w * mu) >> (bits - 1) is at most 3 short and is corrected against
    // w * 2^64 - q * p = -(q * p) mod 2^64.
    __device__ __forceinline__ uint64_t shoup_companion_mu(uint64_t w, uint64_t p, uint64_t mu, int bits)
    {
        const uint64_t hi = __umul64hi(w, mu), lo = w * mu;
        const int s = bits - 1;
        uint64_t q = (hi << (64 - s)) | (lo >> s);
        uint64_t r = 0 - q * p;
        while (r >= p)
        {
            r -= p;
            q++;
        }
        return q;
    }

    // 32-bit companion floor(w * 2^32 / p) from mu = floor(2^64 / p): (w * mu) >> 32 is at most 2 short
    __device__ __forceinline__ uint32_t shoup_companion_mu32(uint32_t w, uint32_t p, uint64_t mu)
    {
        uint64_t q = (uint64_t) w * (uint32_t) (mu >> 32) + (((uint64_t) w * (uint32_t) mu) >> 32);
        uint64_t r = ((uint64_t) w << 32) - q * p;
        while (r >= p)
        {
            r -= p;
            q++;
        }
        return (uint32_t) q;
    }

    // floor((u1 * 2^64 + u0) / v) for u1 < v: schoolbook division in base 2^32 (two 64-bit divisions and their
    // corrections) -- the 128-bit division the compiler would emit for unsigned __int128 costs about a microsecond on
    // one thread, which the RNS kernels pay per segment on their critical path.
    __host__ __device__ __forceinline__ uint64_t div_128_by_64(uint64_t u1, uint64_t u0, uint64_t v)
    {
        const uint64_t b = 1ull << 32;
        int s = 0;
        while (!((v << s) >> 63)) s++; // v != 0
        v <<= s;
        const uint64_t vn1 = v >> 32, vn0 = v & 0xffffffffull;
        const uint64_t un32 = s ? ((u1 << s) | (u0 >> (64 - s))) : u1;
        const uint64_t un10 = u0 << s;
        const uint64_t un1 = un10 >> 32, un0 = un10 & 0xffffffffull;
        uint64_t q1 = un32 / vn1, rhat = un32 - q1 * vn1;
        while (q1 >= b || q1 * vn0 > b * rhat + un1)
        {
            q1--;
            rhat += vn1;
            if (rhat >= b) break;
        }
        const uint64_t un21 = un32 * b + un1 - q1 * v;
        uint64_t q0 = un21 / vn1;
        rhat = un21 - q0 * vn1;
        while (q0 >= b || q0 * vn0 > b * rhat + un0)
        {
            q0--;
            rhat += vn1;
            if (rhat >= b) break;
        }
        return q1 * b + q0;
    }
    // mu = floor(2^(63 + bits) / p) for an odd p of bit length `bits` >= 2 (what shoup_companion_mu takes)
    __host__ __device__ __forceinline__ uint64_t recip_mu64(uint64_t p, int bits) { return div_128_by_64(1ull << (bits - 1), 0, p); }

    // ------------------------------------------------------------------ plain Barrett product
    // a * b mod p for canonical a, b with the reference's Modulus constants (bit = bit length of p,
    // mu = floor(2^(2 bit + 1) / p), modular_arith.cuh:28-57): q = ((z >> (bit-2)) * mu) >> (bit+3) is at most
    // 2 short of floor(z / p).  Used where a twiddle arrives without a Shoup companion (4-step W matrix).
    __device__ __forceinline__ uint64_t barrett_mul(uint64_t a, uint64_t b, uint64_t p, uint64_t bit, uint64_t mu)
    {
        const uint64_t hi = __umul64hi(a, b), lo = a * b;
        const int s1 = (int) bit - 2;
        const uint64_t t = (lo >> s1) | (hi << (64 - s1)); // bit >= 3 here; z < 2^(2 bit) so t < 2^(bit+2)
        const uint64_t ph = __umul64hi(t, mu), pl = t * mu;
        const int s2 = (int) bit + 3;
        const uint64_t q = s2 >= 64 ? (ph >> (s2 - 64)) : ((pl >> s2) | (ph << (64 - s2)));
        uint64_t r = lo - q * p;
        r = csub(r, p + p);
        return csub(r, p);
    }
    __device__ __forceinline__ uint32_t barrett_mul(uint32_t a, uint32_t b, uint32_t p, uint32_t bit, uint32_t mu)
    {
        const uint64_t z = (uint64_t) a * b;
        const uint64_t q = ((z >> (bit - 2)) * mu) >> (bit + 3);
        uint32_t r = (uint32_t) (z - q * p);
        r = csub(r, p + p);
        return csub(r, p);
    }

    // largest modulus the fast policy accepts: forward needs 8p + 2^33 < 2^64, inverse 10p + 2^37 < 2^64
    constexpr uint64_t kFastModulusLimit = (1ull << 60) + (1ull << 58);
    // ... and the smallest: the high-word range tests leave up to 2^32 of slack, which the final correction (at most
    // 11p) only absorbs when 2^32 is small against p
    constexpr uint64_t kFastModulusMin = 1ull << 36;

} // namespace gpuntt_b200
