// tools/arith_variants.cuh -- candidate butterfly formulations timed by arith_bench.cu
#pragma once
template <> inline Twiddle<uint64_t> make_tw<Twiddle<uint64_t>>(uint64_t w, uint64_t p)
{
    return Twiddle<uint64_t>{w, shoup_companion(w, p)};
}

// ---- V0/V0x: the current integer policies with a renorm() stub
struct ModV0 : Mod<uint64_t, true>
{
    __device__ explicit ModV0(uint64_t p_) : Mod<uint64_t, true>(p_) {}
    __device__ __forceinline__ uint64_t renorm(uint64_t x) const { return x; }
};
struct ModV0x : Mod<uint64_t, false>
{
    __device__ explicit ModV0x(uint64_t p_) : Mod<uint64_t, false>(p_) {}
    __device__ __forceinline__ uint64_t renorm(uint64_t x) const { return x; }
    __device__ __forceinline__ uint64_t canon_fwd(uint64_t x) const { return csub(csub(x, two_p), p); }
};

// ---- V1: FP64-assisted quotient.  Twiddle record 32 bytes.
struct __align__(16) TwH
{
    uint64_t w;
    uint32_t a1, pad;
    double A0, K01;
};
template <> inline TwH make_tw<TwH>(uint64_t w, uint64_t p)
{
    const uint64_t wq = shoup_companion(w, p);
    TwH t;
    t.w = w;
    t.a1 = (uint32_t) (wq >> 32);
    t.pad = 0;
    const uint64_t a0 = (uint32_t) wq;
    t.A0 = (double) a0;
    t.K01 = 4503599627370496.0 - (double) ((a0 + (uint64_t) t.a1) << 20); // exact: multiple of 2^20 below 2^53
    return t;
}

template <int CSUB_MODE> struct ModH
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1, k0, k1;
    __device__ __forceinline__ explicit ModH(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
        const T k = 0x4330000000000000ull * p_; // q' = q + 0x433<<52  =>  r' = r - C*p  => add C*p back
        k0 = (uint32_t) k;
        k1 = (uint32_t) (k >> 32);
    }
    // r = w*y - q~*p in [0,4p), q~ in {Q-2..Q}; the two cross terms of the quotient come from the FP64 pipe
    __device__ __forceinline__ T mul(T y, const TwH& tw) const
    {
        const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const double A1 = __hiloint2double(0x43300000, (int) tw.a1) - 4503599627370496.0; // = a1
        const double Y0 = __hiloint2double(0x41300000, (int) y0);                           // 2^20 + y0 * 2^-32
        const double Y1 = __hiloint2double(0x41300000, (int) y1);
        const double inner = __fma_rd(tw.A0, Y1, tw.K01);
        const double v = __fma_rd(A1, Y0, inner); // 2^52 + I,  I in (s'-2, s']
        const uint32_t v0 = (uint32_t) __double2loint(v), v1 = (uint32_t) __double2hiint(v);
        uint32_t r0, r1;
        asm("{\n\t"
            ".reg .u32 q0, q1, t0, t1;\n\t"
            ".reg .u64 q, r, vv, kk;\n\t"
            "mov.b64 vv, {%8, %9};\n\t"
            "mad.wide.u32 q, %6, %2, vv;\n\t" // a1*y1 + v bits
            "mov.b64 {q0, q1}, q;\n\t"
            "mov.b64 kk, {%10, %11};\n\t"
            "mad.wide.u32 r, %4, %3, kk;\n\t" // w0*y0 + C*p
            "mov.b64 {t0, t1}, r;\n\t"
            "mad.lo.u32 t1, %5, %3, t1;\n\t" // w1*y0
            "mad.lo.u32 t1, %4, %2, t1;\n\t" // w0*y1
            "mov.b64 r, {t0, t1};\n\t"
            "mad.wide.u32 r, q0, %7, r;\n\t" // q0*n0
            "mov.b64 {t0, t1}, r;\n\t"
            "mad.lo.u32 t1, q1, %7, t1;\n\t" // q1*n0
            "mad.lo.u32 %1, q0, %12, t1;\n\t" // q0*n1
            "mov.u32 %0, t0;\n\t"
            "}"
            : "=r"(r0), "=r"(r1)
            : "r"(y1), "r"(y0), "r"(w0), "r"(w1), "r"(tw.a1), "r"(n0), "r"(v0), "r"(v1), "r"(k0), "r"(k1), "r"(n1));
        return ((T) r1 << 32) | r0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwH& tw) const
    {
        T x;
        if constexpr (CSUB_MODE == 0)
            x = ((uint32_t) (X >> 32) > f1) ? X - four_p : X;
        else if constexpr (CSUB_MODE == 1)
        {
            uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32);
            asm("{\n\t"
                ".reg .pred P;\n\t"
                "setp.gt.u32 P, %1, %3;\n\t"
                "@P sub.cc.u32 %0, %0, %2;\n\t"
                "@P subc.u32 %1, %1, %3;\n\t"
                "}"
                : "+r"(x0), "+r"(x1)
                : "r"(f0), "r"(f1));
            x = ((T) x1 << 32) | x0;
        }
        else
            x = X; // lazy: caller renormalises
        const T t = mul(Y, tw);
        X = x + t;
        Y = x - t + four_p;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const TwH& tw) const
    {
        const T s = X + Y;
        const T d = X - Y + four_p;
        X = csub(s, four_p);
        Y = mul(d, tw);
    }
    __device__ __forceinline__ T renorm(T x) const
    {
        x = csub(x, 4 * four_p);
        x = csub(x, 2 * four_p);
        return csub(x, four_p);
    }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V3: FP64-assisted quotient, 32-bit-limb formulation with the additions arranged for 3-input IADD3
template <int CSUB_MODE> struct ModH3
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1, kq;
    __device__ __forceinline__ explicit ModH3(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
        kq = 0u - 0x43300000u * n0; // q1 carries +0x43300000 (the exponent bits of v); cancel its product with n0
    }
    __device__ __forceinline__ void mul32(uint32_t y0, uint32_t y1, const TwH& tw, uint32_t& r0, uint32_t& r1) const
    {
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const double A1 = __hiloint2double(0x43300000, (int) tw.a1) - 4503599627370496.0;
        const double Y0 = __hiloint2double(0x41300000, (int) y0);
        const double Y1 = __hiloint2double(0x41300000, (int) y1);
        const double inner = __fma_rd(tw.A0, Y1, tw.K01);
        const double v = __fma_rd(A1, Y0, inner);
        const uint64_t q = (uint64_t) tw.a1 * y1 + (uint64_t) __double_as_longlong(v);
        const uint32_t q0 = (uint32_t) q, q1 = (uint32_t) (q >> 32);
        const uint64_t A = (uint64_t) w0 * y0;
        const uint64_t B = (uint64_t) q0 * n0;
        uint32_t u = w1 * y0 + kq;
        u = w0 * y1 + u;
        u = q1 * n0 + u;
        u = q0 * n1 + u;
        const uint64_t s = A + B;
        r0 = (uint32_t) s;
        r1 = (uint32_t) (s >> 32) + u;
    }
    __device__ __forceinline__ T mul(T y, const TwH& tw) const
    {
        uint32_t r0, r1;
        mul32((uint32_t) y, (uint32_t) (y >> 32), tw, r0, r1);
        return ((T) r1 << 32) | r0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwH& tw) const
    {
        T x;
        if constexpr (CSUB_MODE == 1)
        {
            uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32);
            asm("{\n\t"
                ".reg .pred P;\n\t"
                "setp.gt.u32 P, %1, %3;\n\t"
                "@P sub.cc.u32 %0, %0, %2;\n\t"
                "@P subc.u32 %1, %1, %3;\n\t"
                "}"
                : "+r"(x0), "+r"(x1)
                : "r"(f0), "r"(f1));
            x = ((T) x1 << 32) | x0;
        }
        else
            x = X;
        const T t = mul(Y, tw);
        X = x + t;
        Y = x - t + four_p;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const TwH& tw) const
    {
        const T s = X + Y;
        const T d = X - Y + four_p;
        X = csub(s, four_p);
        Y = mul(d, tw);
    }
    __device__ __forceinline__ T renorm(T x) const
    {
        x = csub(x, 4 * four_p);
        return csub(x, 2 * four_p);
    }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V4: whole Cooley-Tukey butterfly in one PTX block (FP64-assisted quotient)
struct ModP4
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1, kq, c41;
    double m52;
    __device__ __forceinline__ explicit ModP4(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
        kq = 0u - 0x43300000u * n0;
        c41 = 0x41300000u;
        m52 = 4503599627370496.0;
    }
    __device__ __forceinline__ T mul(T y, const TwH& tw) const
    {
        T X = 0, Y = y;
        // reuse ct: X = 0 -> X' = t
        ct(X, Y, tw);
        return X;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwH& tw) const
    {
        uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32), y0 = (uint32_t) Y, y1 = (uint32_t) (Y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const double A1 = __hiloint2double(0x43300000, (int) tw.a1) - m52;
        asm("{\n\t"
            ".reg .pred P;\n\t"
            ".reg .u32 q0, q1, u, t0, t1, a0, a1;\n\t"
            ".reg .u64 q, A, B;\n\t"
            ".reg .f64 Y0, Y1, in, v;\n\t"
            "setp.gt.u32 P, %1, %13;\n\t"
            "@P sub.cc.u32 %0, %0, %12;\n\t"
            "@P subc.u32 %1, %1, %13;\n\t"
            "mov.b64 Y0, {%2, %15};\n\t"
            "mov.b64 Y1, {%3, %15};\n\t"
            "fma.rm.f64 in, %8, Y1, %9;\n\t"
            "fma.rm.f64 v, %7, Y0, in;\n\t"
            "mov.b64 q, v;\n\t"
            "mad.wide.u32 q, %6, %3, q;\n\t"   // a1*y1 + v bits
            "mov.b64 {q0, q1}, q;\n\t"
            "mul.wide.u32 A, %4, %2;\n\t"      // w0*y0
            "mad.lo.u32 u, %5, %2, %14;\n\t"   // w1*y0 + kq
            "mad.lo.u32 u, %4, %3, u;\n\t"     // w0*y1
            "mad.lo.u32 u, q1, %10, u;\n\t"    // q1*n0
            "mad.lo.u32 u, q0, %11, u;\n\t"    // q0*n1
            "mad.wide.u32 B, q0, %10, A;\n\t"  // q0*n0 + A
            "mov.b64 {t0, t1}, B;\n\t"
            "add.u32 t1, t1, u;\n\t"
            // Y' = x - t + 4p ; X' = x + t
            "sub.cc.u32 %2, %0, t0;\n\t"
            "subc.u32 %3, %1, t1;\n\t"
            "add.cc.u32 %2, %2, %12;\n\t"
            "addc.u32 %3, %3, %13;\n\t"
            "add.cc.u32 %0, %0, t0;\n\t"
            "addc.u32 %1, %1, t1;\n\t"
            "}"
            : "+r"(x0), "+r"(x1), "+r"(y0), "+r"(y1)
            : "r"(w0), "r"(w1), "r"(tw.a1), "d"(A1), "d"(tw.A0), "d"(tw.K01), "r"(n0), "r"(n1), "r"(f0), "r"(f1), "r"(kq), "r"(c41));
        X = ((T) x1 << 32) | x0;
        Y = ((T) y1 << 32) | y0;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const TwH& tw) const {}
    __device__ __forceinline__ T renorm(T x) const { return x; }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V5: V4 + pipe steering variants (text-assembled PTX)
#define V5_HEAD \
    "{\n\t.reg .pred P;\n\t.reg .u32 q0, q1, u, t0, t1, m0, m1;\n\t.reg .u64 q, A, B;\n\t.reg .f64 Y0, Y1, in, v;\n\t" \
    "setp.gt.u32 P, %1, %13;\n\t@P sub.cc.u32 %0, %0, %12;\n\t@P subc.u32 %1, %1, %13;\n\t"
#define V5_MOVE_PLAIN "mov.b64 Y0, {%2, %15};\n\tmov.b64 Y1, {%3, %15};\n\t"
#define V5_MOVE_LOP "or.b32 m0, %2, %16;\n\tor.b32 m1, %3, %16;\n\tmov.b64 Y0, {m0, %15};\n\tmov.b64 Y1, {m1, %15};\n\t"
#define V5_MUL \
    "fma.rm.f64 in, %8, Y1, %9;\n\tfma.rm.f64 v, %7, Y0, in;\n\tmov.b64 q, v;\n\tmad.wide.u32 q, %6, %3, q;\n\tmov.b64 {q0, q1}, q;\n\t" \
    "mul.wide.u32 A, %4, %2;\n\tmad.lo.u32 u, %5, %2, %14;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %10, u;\n\t" \
    "mad.lo.u32 u, q0, %11, u;\n\tmad.wide.u32 B, q0, %10, A;\n\tmov.b64 {t0, t1}, B;\n\t"
#define V5_TU_PLAIN "add.u32 t1, t1, u;\n\t"
#define V5_Y_PLAIN "sub.cc.u32 %2, %0, t0;\n\tsubc.u32 %3, %1, t1;\n\tadd.cc.u32 %2, %2, %12;\n\taddc.u32 %3, %3, %13;\n\t"
#define V5_X_PLAIN "add.cc.u32 %0, %0, t0;\n\taddc.u32 %1, %1, t1;\n\t"
#define V5_X_3IN "add.cc.u32 %0, %0, t0;\n\taddc.u32 %1, %1, t1;\n\tadd.cc.u32 %0, %0, %16;\n\taddc.u32 %1, %1, %16;\n\t"
// t1 + u folded into the consumers: X1' = x1 + t1 + u + c ; Y1' = x1 - t1 - u + f1 + c
#define V5_TAIL "}"
#define V5_OPERANDS \
    : "+r"(x0), "+r"(x1), "+r"(y0), "+r"(y1) \
    : "r"(w0), "r"(w1), "r"(tw.a1), "d"(A1), "d"(tw.A0), "d"(tw.K01), "r"(n0), "r"(n1), "r"(f0), "r"(f1), "r"(kq), "r"(c41), "r"(z)

template <int STEER> struct ModP5
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1, kq, c41, z;
    double m52;
    __device__ __forceinline__ explicit ModP5(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
        kq = 0u - 0x43300000u * n0;
        c41 = 0x41300000u;
        z = (uint32_t) (p_ >> 63); // opaque zero
        m52 = 4503599627370496.0;
    }
    __device__ __forceinline__ T mul(T y, const TwH& tw) const
    {
        T X = 0, Y = y;
        ct(X, Y, tw);
        return X;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwH& tw) const
    {
        uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32), y0 = (uint32_t) Y, y1 = (uint32_t) (Y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const double A1 = __hiloint2double(0x43300000, (int) tw.a1) - m52;
        if constexpr (STEER == 0)
            asm(V5_HEAD V5_MOVE_PLAIN V5_MUL V5_TU_PLAIN V5_Y_PLAIN V5_X_PLAIN V5_TAIL V5_OPERANDS);
        else if constexpr (STEER == 1)
            asm(V5_HEAD V5_MOVE_LOP V5_MUL V5_TU_PLAIN V5_Y_PLAIN V5_X_PLAIN V5_TAIL V5_OPERANDS);
        else if constexpr (STEER == 2)
            asm(V5_HEAD V5_MOVE_LOP V5_MUL V5_TU_PLAIN V5_Y_PLAIN V5_X_3IN V5_TAIL V5_OPERANDS);
        else
            asm(V5_HEAD V5_MOVE_PLAIN V5_MUL V5_TU_PLAIN V5_Y_PLAIN V5_X_3IN V5_TAIL V5_OPERANDS);
        X = ((T) x1 << 32) | x0;
        Y = ((T) y1 << 32) | y0;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const TwH& tw) const {}
    __device__ __forceinline__ T renorm(T x) const { return x; }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V6: FP64-assisted quotient with cvt (I2F on the XU pipe) instead of the register-pair bit trick
struct __align__(16) TwI
{
    uint64_t w;
    uint32_t a1, pad;
    double A1s, A0s; // a1 * 2^-32, a0 * 2^-32
};
template <> inline TwI make_tw<TwI>(uint64_t w, uint64_t p)
{
    const uint64_t wq = shoup_companion(w, p);
    TwI t;
    t.w = w;
    t.a1 = (uint32_t) (wq >> 32);
    t.pad = 0;
    t.A1s = (double) t.a1 / 4294967296.0;
    t.A0s = (double) (uint32_t) wq / 4294967296.0;
    return t;
}
#define V6_BODY(YPART) \
    "{\n\t.reg .pred P;\n\t.reg .u32 q0, q1, u, t0, t1;\n\t.reg .u64 q, A, B;\n\t.reg .f64 Y0, Y1, in, v;\n\t" \
    "setp.gt.u32 P, %1, %13;\n\t@P sub.cc.u32 %0, %0, %12;\n\t@P subc.u32 %1, %1, %13;\n\t" \
    "cvt.rn.f64.u32 Y0, %2;\n\tcvt.rn.f64.u32 Y1, %3;\n\t" \
    "fma.rm.f64 in, %8, Y1, %9;\n\tfma.rm.f64 v, %7, Y0, in;\n\tmov.b64 q, v;\n\tmad.wide.u32 q, %6, %3, q;\n\tmov.b64 {q0, q1}, q;\n\t" \
    "mul.wide.u32 A, %4, %2;\n\tmad.lo.u32 u, %5, %2, %14;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %10, u;\n\t" \
    "mad.lo.u32 u, q0, %11, u;\n\tmad.wide.u32 B, q0, %10, A;\n\tmov.b64 {t0, t1}, B;\n\tadd.u32 t1, t1, u;\n\t" \
    YPART \
    "add.cc.u32 %0, %0, t0;\n\taddc.u32 %1, %1, t1;\n\t}"
struct ModP6
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1, kq;
    double m52;
    __device__ __forceinline__ explicit ModP6(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
        kq = 0u - 0x43300000u * n0;
        m52 = 4503599627370496.0;
    }
    __device__ __forceinline__ T mul(T y, const TwI& tw) const
    {
        T X = 0, Y = y;
        ct(X, Y, tw);
        return X;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwI& tw) const
    {
        uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32), y0 = (uint32_t) Y, y1 = (uint32_t) (Y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        asm(V6_BODY("sub.cc.u32 %2, %0, t0;\n\tsubc.u32 %3, %1, t1;\n\tadd.cc.u32 %2, %2, %12;\n\taddc.u32 %3, %3, %13;\n\t")
            : "+r"(x0), "+r"(x1), "+r"(y0), "+r"(y1)
            : "r"(w0), "r"(w1), "r"(tw.a1), "d"(tw.A1s), "d"(tw.A0s), "d"(m52), "r"(n0), "r"(n1), "r"(f0), "r"(f1), "r"(kq));
        X = ((T) x1 << 32) | x0;
        Y = ((T) y1 << 32) | y0;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const TwI& tw) const {}
    __device__ __forceinline__ T renorm(T x) const { return x; }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V7: V6's multiply in PTX, additions in C (ptxas fuses x - t + 4p into 3-input IADD3 pairs)
template <int MODE> struct ModP7 : ModP6
{
    __device__ __forceinline__ explicit ModP7(T p_) : ModP6(p_) {}
    __device__ __forceinline__ T mul(T y, const TwI& tw) const
    {
        const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        uint32_t t0, t1;
        asm("{\n\t.reg .u32 q0, q1, u;\n\t.reg .u64 q, A, B;\n\t.reg .f64 Y0, Y1, in, v;\n\t"
            "cvt.rn.f64.u32 Y0, %2;\n\tcvt.rn.f64.u32 Y1, %3;\n\t"
            "fma.rm.f64 in, %8, Y1, %9;\n\tfma.rm.f64 v, %7, Y0, in;\n\tmov.b64 q, v;\n\tmad.wide.u32 q, %6, %3, q;\n\tmov.b64 {q0, q1}, q;\n\t"
            "mul.wide.u32 A, %4, %2;\n\tmad.lo.u32 u, %5, %2, %12;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %10, u;\n\t"
            "mad.lo.u32 u, q0, %11, u;\n\tmad.wide.u32 B, q0, %10, A;\n\tmov.b64 {%0, %1}, B;\n\tadd.u32 %1, %1, u;\n\t}"
            : "=r"(t0), "=r"(t1)
            : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(tw.a1), "d"(tw.A1s), "d"(tw.A0s), "d"(m52), "r"(n0), "r"(n1), "r"(kq));
        return ((T) t1 << 32) | t0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwI& tw) const
    {
        T x;
        if constexpr (MODE == 0)
            x = ((uint32_t) (X >> 32) > f1) ? X - four_p : X;
        else
        {
            uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32);
            asm("{\n\t.reg .pred P;\n\tsetp.gt.u32 P, %1, %3;\n\t@P sub.cc.u32 %0, %0, %2;\n\t@P subc.u32 %1, %1, %3;\n\t}"
                : "+r"(x0), "+r"(x1) : "r"(f0), "r"(f1));
            x = ((T) x1 << 32) | x0;
        }
        const T t = mul(Y, tw);
        X = x + t;
        Y = x - t + four_p;
    }
};


// ---- V8: integer-only 3-product quotient, multiply in PTX with WIDE addends, additions in C, predicated csub
struct ModP8
{
    using T = uint64_t;
    T p, four_p;
    uint32_t n0, n1, f0, f1;
    __device__ __forceinline__ explicit ModP8(T p_) : p(p_), four_p(4 * p_)
    {
        const T np = 0 - p_;
        n0 = (uint32_t) np;
        n1 = (uint32_t) (np >> 32);
        f0 = (uint32_t) four_p;
        f1 = (uint32_t) (four_p >> 32);
    }
    __device__ __forceinline__ T mul(T y, const Twiddle<uint64_t>& tw) const
    {
        const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const uint32_t a0 = (uint32_t) tw.wq, a1 = (uint32_t) (tw.wq >> 32);
        uint32_t t0, t1;
        asm("{\n\t.reg .u32 q0, q1, u, h1, h2, z;\n\t.reg .u64 q, A, B, H;\n\t"
            "mul.hi.u32 h1, %6, %2;\n\tmul.hi.u32 h2, %7, %3;\n\t"
            "mov.u32 z, 0;\n\tadd.cc.u32 h1, h1, h2;\n\taddc.u32 h2, z, z;\n\tmov.b64 H, {h1, h2};\n\t"
            "mad.wide.u32 q, %6, %3, H;\n\tmov.b64 {q0, q1}, q;\n\t"
            "mul.wide.u32 A, %4, %2;\n\tmul.lo.u32 u, %5, %2;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %8, u;\n\t"
            "mad.lo.u32 u, q0, %9, u;\n\tmad.wide.u32 B, q0, %8, A;\n\tmov.b64 {%0, %1}, B;\n\tadd.u32 %1, %1, u;\n\t}"
            : "=r"(t0), "=r"(t1)
            : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(a1), "r"(a0), "r"(n0), "r"(n1));
        return ((T) t1 << 32) | t0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<uint64_t>& tw) const
    {
        uint32_t x0 = (uint32_t) X, x1 = (uint32_t) (X >> 32);
        asm("{\n\t.reg .pred P;\n\tsetp.gt.u32 P, %1, %3;\n\t@P sub.cc.u32 %0, %0, %2;\n\t@P subc.u32 %1, %1, %3;\n\t}"
            : "+r"(x0), "+r"(x1) : "r"(f0), "r"(f1));
        const T x = ((T) x1 << 32) | x0;
        const T t = mul(Y, tw);
        X = x + t;
        Y = x - t + four_p;
    }
    __device__ __forceinline__ void gs(T& X, T& Y, const Twiddle<uint64_t>& tw) const {}
    __device__ __forceinline__ T renorm(T x) const { return x; }
    __device__ __forceinline__ T canon_fwd(T x) const
    {
        x = csub(csub(x, four_p), four_p);
        return csub(csub(x, p + p), p);
    }
};


// ---- V9: V7 with every addition forced onto the alu pipe (3-input forms with an opaque zero), csub via SEL
template <int MODE> struct ModP9 : ModP6
{
    uint32_t z;
    T zz;
    __device__ __forceinline__ explicit ModP9(T p_) : ModP6(p_)
    {
        z = (uint32_t) (p_ >> 63);
        zz = p_ >> 63;
    }
    __device__ __forceinline__ T mul(T y, const TwI& tw) const
    {
        const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        uint32_t t0, t1;
        asm("{\n\t.reg .u32 q0, q1, u;\n\t.reg .u64 q, A, B;\n\t.reg .f64 Y0, Y1, in, v;\n\t"
            "cvt.rn.f64.u32 Y0, %2;\n\tcvt.rn.f64.u32 Y1, %3;\n\t"
            "fma.rm.f64 in, %8, Y1, %9;\n\tfma.rm.f64 v, %7, Y0, in;\n\tmov.b64 q, v;\n\tmad.wide.u32 q, %6, %3, q;\n\tmov.b64 {q0, q1}, q;\n\t"
            "mul.wide.u32 A, %4, %2;\n\tmad.lo.u32 u, %5, %2, %12;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %10, u;\n\t"
            "mad.lo.u32 u, q0, %11, u;\n\tmad.wide.u32 B, q0, %10, A;\n\tmov.b64 {%0, %1}, B;\n\tadd.u32 %1, %1, u;\n\tadd.u32 %1, %1, %13;\n\t}"
            : "=r"(t0), "=r"(t1)
            : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(tw.a1), "d"(tw.A1s), "d"(tw.A0s), "d"(m52), "r"(n0), "r"(n1), "r"(kq), "r"(z));
        return ((T) t1 << 32) | t0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwI& tw) const
    {
        const bool c = (uint32_t) (X >> 32) > f1;
        const T t = mul(Y, tw);
        if constexpr (MODE == 0)
        {
            const T g = c ? four_p : zz;
            const T x = X - g + zz;
            X = x + t + zz;
            Y = x - t + four_p;
        }
        else
        {
            const T g = c ? four_p : zz; // X' = X - g + t ; Y' = X + (4p - g) - t
            const T h = c ? zz : four_p;
            Y = X + h - t;
            X = X - g + t;
        }
    }
};


// ---- V10: whole butterfly in PTX with 64-bit add/sub pairs that ptxas fuses into 3-input IADD3 (alu pipe only)
template <int MODE> struct ModP10 : ModP6
{
    T zz;
    __device__ __forceinline__ explicit ModP10(T p_) : ModP6(p_) { zz = p_ >> 63; }
    __device__ __forceinline__ T mul(T y, const TwI& tw) const
    {
        T X = 0, Y = y;
        ct(X, Y, tw);
        return X;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const TwI& tw) const
    {
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        asm("{\n\t.reg .pred P;\n\t.reg .u32 q0, q1, u, t0, t1, x0, x1, y0, y1, z0, z1;\n\t.reg .u64 q, A, B, g, x, t;\n\t.reg .f64 Y0, Y1, in, v;\n\t"
            "mov.b64 {x0, x1}, %0;\n\tmov.b64 {y0, y1}, %1;\n\tmov.b64 {z0, z1}, %12;\n\t"
            "setp.gt.u32 P, x1, %11;\n\tselp.b64 g, %10, %12, P;\n\t"
            "sub.u64 x, %0, g;\n\tadd.u64 x, x, %12;\n\t"
            "cvt.rn.f64.u32 Y0, y0;\n\tcvt.rn.f64.u32 Y1, y1;\n\t"
            "fma.rm.f64 in, %6, Y1, %7;\n\tfma.rm.f64 v, %5, Y0, in;\n\tmov.b64 q, v;\n\tmad.wide.u32 q, %4, y1, q;\n\tmov.b64 {q0, q1}, q;\n\t"
            "mul.wide.u32 A, %2, y0;\n\tmad.lo.u32 u, %3, y0, %13;\n\tmad.lo.u32 u, %2, y1, u;\n\tmad.lo.u32 u, q1, %8, u;\n\t"
            "mad.lo.u32 u, q0, %9, u;\n\tmad.wide.u32 B, q0, %8, A;\n\tmov.b64 {t0, t1}, B;\n\tadd.u32 t1, t1, u;\n\tadd.u32 t1, t1, z0;\n\t"
            "mov.b64 t, {t0, t1};\n\t"
            "sub.u64 %1, x, t;\n\tadd.u64 %1, %1, %10;\n\t"
            "add.u64 %0, x, t;\n\tadd.u64 %0, %0, %12;\n\t}"
            : "+l"(X), "+l"(Y)
            : "r"(w0), "r"(w1), "r"(tw.a1), "d"(tw.A1s), "d"(tw.A0s), "d"(m52), "r"(n0), "r"(n1), "l"(four_p), "r"(f1), "l"(zz), "r"(kq));
    }
};


// ---- V11: integer 3-product quotient (PTX) + additions forced onto the alu pipe
template <int MODE> struct ModP11 : ModP8
{
    uint32_t z;
    T zz;
    __device__ __forceinline__ explicit ModP11(T p_) : ModP8(p_)
    {
        z = (uint32_t) (p_ >> 63);
        zz = p_ >> 63;
    }
    __device__ __forceinline__ T mul(T y, const Twiddle<uint64_t>& tw) const
    {
        const uint32_t y0 = (uint32_t) y, y1 = (uint32_t) (y >> 32);
        const uint32_t w0 = (uint32_t) tw.w, w1 = (uint32_t) (tw.w >> 32);
        const uint32_t a0 = (uint32_t) tw.wq, a1 = (uint32_t) (tw.wq >> 32);
        uint32_t t0, t1;
        if constexpr (MODE == 0)
            asm("{\n\t.reg .u32 q0, q1, u, h1, h2, hc;\n\t.reg .u64 q, A, B, H;\n\t"
                "mul.hi.u32 h1, %6, %2;\n\tmul.hi.u32 h2, %7, %3;\n\t"
                "add.cc.u32 h1, h1, h2;\n\taddc.u32 hc, %10, %10;\n\tmov.b64 H, {h1, hc};\n\t"
                "mad.wide.u32 q, %6, %3, H;\n\tmov.b64 {q0, q1}, q;\n\t"
                "mul.wide.u32 A, %4, %2;\n\tmul.lo.u32 u, %5, %2;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %8, u;\n\t"
                "mad.lo.u32 u, q0, %9, u;\n\tmad.wide.u32 B, q0, %8, A;\n\tmov.b64 {%0, %1}, B;\n\tadd.u32 %1, %1, u;\n\tadd.u32 %1, %1, %10;\n\t}"
                : "=r"(t0), "=r"(t1)
                : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(a1), "r"(a0), "r"(n0), "r"(n1), "r"(z));
        else // cross terms accumulated in the high word chain instead of a 64-bit addend
            asm("{\n\t.reg .u32 q0, q1, u, h1, h2;\n\t.reg .u64 q, A, B;\n\t"
                "mul.hi.u32 h1, %6, %2;\n\tmul.hi.u32 h2, %7, %3;\n\t"
                "mul.wide.u32 q, %6, %3;\n\tmov.b64 {q0, q1}, q;\n\t"
                "add.cc.u32 q0, q0, h1;\n\taddc.u32 q1, q1, %10;\n\tadd.cc.u32 q0, q0, h2;\n\taddc.u32 q1, q1, %10;\n\t"
                "mul.wide.u32 A, %4, %2;\n\tmul.lo.u32 u, %5, %2;\n\tmad.lo.u32 u, %4, %3, u;\n\tmad.lo.u32 u, q1, %8, u;\n\t"
                "mad.lo.u32 u, q0, %9, u;\n\tmad.wide.u32 B, q0, %8, A;\n\tmov.b64 {%0, %1}, B;\n\tadd.u32 %1, %1, u;\n\tadd.u32 %1, %1, %10;\n\t}"
                : "=r"(t0), "=r"(t1)
                : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(a1), "r"(a0), "r"(n0), "r"(n1), "r"(z));
        return ((T) t1 << 32) | t0;
    }
    __device__ __forceinline__ void ct(T& X, T& Y, const Twiddle<uint64_t>& tw) const
    {
        uint32_t g0, g1;
        asm("{\n\t.reg .pred P;\n\tsetp.gt.u32 P, %2, %4;\n\tselp.b32 %0, %3, %5, P;\n\tselp.b32 %1, %4, %5, P;\n\t}"
            : "=r"(g0), "=r"(g1) : "r"((uint32_t) (X >> 32)), "r"(f0), "r"(f1), "r"(z));
        const T g = ((T) g1 << 32) | g0;
        const T t = mul(Y, tw);
        const T x = X - g + zz;
        X = x + t + zz;
        Y = x - t + four_p;
    }
};

static void run_variants(uint64_t p, const uint64_t* hw)
{
    run_check<ModV0, Twiddle<uint64_t>>("V0  integer 3-product quotient", p, 4);
    run_check<ModH<0>, TwH>("V1  FP64-assisted quotient", p, 4);
    run_bfly<ModV0, Twiddle<uint64_t>, false, false>("V0  fwd integer fast policy (current)", p, hw);
    run_bfly<ModV0x, Twiddle<uint64_t>, false, false>("V0x fwd integer exact policy", p, hw);
    run_bfly<ModH<0>, TwH, false, false>("V1  fwd FP64 quotient, C csub", p, hw);
    run_bfly<ModH<1>, TwH, false, false>("V1p fwd FP64 quotient, predicated csub", p, hw);
    run_bfly<ModH<2>, TwH, false, true>("V2  fwd FP64 quotient, lazy X + renorm/round", p, hw);
    run_check<ModH3<1>, TwH>("V3  FP64 quotient, limb form", p, 4);
    run_bfly<ModH3<1>, TwH, false, false>("V3  fwd FP64 quotient limb form, pred csub", p, hw);
    run_bfly<ModH3<2>, TwH, false, false>("V3n fwd FP64 quotient limb form, NO csub (invalid, cost probe)", p, hw);
    run_check<ModP4, TwH>("V4  whole butterfly in PTX", p, 4);
    run_bfly<ModP4, TwH, false, false>("V4  fwd whole butterfly in PTX", p, hw);
    run_check<ModP5<2>, TwH>("V5.2 steered PTX butterfly", p, 4);
    run_bfly<ModP5<0>, TwH, false, false>("V5.0 fwd PTX butterfly (= V4)", p, hw);
    run_bfly<ModP5<1>, TwH, false, false>("V5.1 fwd PTX, LOP3 moves", p, hw);
    run_bfly<ModP5<2>, TwH, false, false>("V5.2 fwd PTX, LOP3 moves + 3-input X'", p, hw);
    run_bfly<ModP5<3>, TwH, false, false>("V5.3 fwd PTX, 3-input X'", p, hw);
    run_check<ModP6, TwI>("V6  PTX butterfly, cvt-based FP64 quotient", p, 4);
    run_bfly<ModP6, TwI, false, false>("V6  fwd PTX butterfly, cvt-based FP64 quotient", p, hw);
    run_check<ModP7<0>, TwI>("V7  PTX multiply (cvt FP64 quotient), C adds", p, 4);
    run_bfly<ModP7<0>, TwI, false, false>("V7.0 fwd PTX mul + C adds, C csub", p, hw);
    run_bfly<ModP7<1>, TwI, false, false>("V7.1 fwd PTX mul + C adds, pred csub", p, hw);
    run_check<ModP8, Twiddle<uint64_t>>("V8  integer PTX multiply", p, 4);
    run_bfly<ModP8, Twiddle<uint64_t>, false, false>("V8  fwd integer PTX mul + C adds, pred csub", p, hw);
    run_check<ModP9<0>, TwI>("V9  PTX mul, alu-forced adds", p, 4);
    run_bfly<ModP9<0>, TwI, false, false>("V9.0 fwd PTX mul, alu-forced adds (x via SEL)", p, hw);
    run_bfly<ModP9<1>, TwI, false, false>("V9.1 fwd PTX mul, alu-forced adds (g/h SEL)", p, hw);
    run_check<ModP10<0>, TwI>("V10 PTX butterfly, 64-bit fused adds", p, 4);
    run_bfly<ModP10<0>, TwI, false, false>("V10 fwd PTX butterfly, 64-bit fused adds", p, hw);
    run_bfly<ModP7<1>, TwI, false, false, 3>("V7.1 fwd, 3 blocks/SM (<=80 regs)", p, hw);
    run_bfly<ModP7<1>, TwI, false, false, 4>("V7.1 fwd, 4 blocks/SM (<=64 regs)", p, hw);
    run_bfly<ModP8, Twiddle<uint64_t>, false, false, 3>("V8 fwd, 3 blocks/SM", p, hw);
    run_bfly<ModP8, Twiddle<uint64_t>, false, false, 4>("V8 fwd, 4 blocks/SM", p, hw);
    run_bfly<ModV0, Twiddle<uint64_t>, false, false, 3>("V0 fwd, 3 blocks/SM", p, hw);
    run_check<ModP11<0>, Twiddle<uint64_t>>("V11.0 integer PTX mul, alu-forced adds", p, 4);
    run_check<ModP11<1>, Twiddle<uint64_t>>("V11.1 integer PTX mul, alu-forced adds", p, 4);
    run_bfly<ModP11<0>, Twiddle<uint64_t>, false, false>("V11.0 fwd integer PTX mul (addend q), alu-forced adds", p, hw);
    run_bfly<ModP11<1>, Twiddle<uint64_t>, false, false>("V11.1 fwd integer PTX mul (carry q), alu-forced adds", p, hw);
    run_bfly<ModV0, Twiddle<uint64_t>, true, false>("V0  inv integer fast policy (current)", p, hw);
    run_bfly<ModH<0>, TwH, true, false>("V1  inv FP64 quotient", p, hw);
}
