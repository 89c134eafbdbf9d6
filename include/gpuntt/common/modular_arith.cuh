// include/gpuntt/common/modular_arith.cuh -- element types, Modulus<T> and scalar modular arithmetic.
//
// Source-compatible re-creation of the reference header of the same name
// (src/include/gpuntt/common/modular_arith.cuh): identical type names, member order and
// signatures so that GPU-NTT callers recompile unchanged; the bodies are this project's own.
// The batched transforms do NOT use these per-element routines (they run on Shoup twiddle pairs,
// gpu_ntt_b200/csrc/modarith.cuh); OPERATOR / OPERATOR_GPU exist for caller code that does its own
// pointwise arithmetic between transforms.
#ifndef GPUNTT_B200_MODULAR_ARITH_CUH
#define GPUNTT_B200_MODULAR_ARITH_CUH

#include <math.h>

#include <cinttypes>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>
#include <device_launch_parameters.h>

// global-namespace aliases, as in the reference (modular_arith.cuh:18-26)
typedef std::int32_t Data32s;
typedef std::uint32_t Data32;
typedef std::uint32_t Root32;
typedef std::uint32_t Ninverse32;
typedef std::int64_t Data64s;
typedef std::uint64_t Data64;
typedef std::uint64_t Root64;
typedef std::uint64_t Ninverse64;

// {p, bit length of p, floor(2^(2*bit+1) / p)} -- layout is ABI (passed by value to GPU_NTT;
// arrays of it live in device memory for the RNS entry points).  modular_arith.cuh:28-57.
template <typename T1> struct Modulus
{
    T1 value;
    T1 bit;
    T1 mu;

    __host__ Modulus(T1 mod) : value(mod), bit(0), mu(0)
    {
        // integer bit length (the reference takes log2() in floating point, modular_arith.cuh:46;
        // both give the same answer for every modulus below 2^53 and for its pooled primes)
        T1 v = mod;
        while (v)
        {
            bit++;
            v >>= 1;
        }
        using Wide = typename std::conditional<std::is_same<T1, Data32>::value, Data64, unsigned __int128>::type;
        mu = mod ? static_cast<T1>((static_cast<Wide>(1) << (2 * bit + 1)) / mod) : 0;
    }
    __host__ Modulus() : value(0), bit(0), mu(0) {}
};
typedef Modulus<Data32> Modulus32;
typedef Modulus<Data64> Modulus64;

template <typename T>
using Root = typename std::conditional<std::is_same<T, Data32>::value, Root32, Root64>::type;
template <typename T>
using Ninverse = typename std::conditional<std::is_same<T, Data32>::value, Ninverse32, Ninverse64>::type;

namespace modular_operation_cpu
{
    // Host arithmetic on canonical residues (modular_arith.cuh:62-158 of the reference).
    template <typename T1> class BarrettOperations
    {
        using Wide = typename std::conditional<std::is_same<T1, Data32>::value, Data64, unsigned __int128>::type;

      public:
        static __host__ T1 add(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            const T1 s = a + b;
            return s >= m.value ? s - m.value : s;
        }
        static __host__ T1 sub(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            const T1 d = a + m.value - b;
            return d >= m.value ? d - m.value : d;
        }
        static __host__ T1 mult(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            return static_cast<T1>(static_cast<Wide>(a) * static_cast<Wide>(b) % m.value);
        }
        static __host__ T1 exp(T1 base, T1 exponent, const Modulus<T1>& m)
        {
            T1 r = 1 % m.value;
            while (exponent)
            {
                if (exponent & 1) r = mult(r, base, m);
                base = mult(base, base, m);
                exponent >>= 1;
            }
            return r;
        }
        // prime modulus assumed, like the reference (Fermat)
        static __host__ T1 modinv(T1 input, const Modulus<T1>& m) { return exp(input, m.value - 2, m); }
        static __host__ T1 reduce(const T1& a, const Modulus<T1>& m) { return a % m.value; }
    };
} // namespace modular_operation_cpu

template <typename T> using OPERATOR = modular_operation_cpu::BarrettOperations<T>;
typedef OPERATOR<Data32> OPERATOR32;
typedef OPERATOR<Data64> OPERATOR64;

namespace modular_operation_gpu
{
    // Device arithmetic for caller kernels (modular_arith.cuh:174-454 of the reference).  Inputs are
    // canonical residues; results are canonical.  Moduli up to 30 / 62 bits, as in the reference.
    template <typename T1> class BarrettOperations
    {
      public:
        static __device__ __forceinline__ T1 add(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            const T1 s = a + b;
            return s >= m.value ? s - m.value : s;
        }
        static __device__ __forceinline__ T1 sub(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            const T1 d = a + m.value - b;
            return d >= m.value ? d - m.value : d;
        }
        // Barrett with the struct's own constants: q = ((z >> (bit-2)) * mu) >> (bit+3)
        static __device__ __forceinline__ T1 mult(const T1& a, const T1& b, const Modulus<T1>& m)
        {
            if constexpr (std::is_same<T1, Data32>::value)
                return reduce_wide(static_cast<Data64>(a) * b, m);
            else
                return reduce_wide(mulhi64(a, b), a * b, m);
        }
        static __device__ __forceinline__ T1 reduce(const T1& a, const Modulus<T1>& m)
        {
            if constexpr (std::is_same<T1, Data32>::value)
                return reduce_wide(static_cast<Data64>(a), m);
            else
                return reduce_wide(Data64(0), a, m);
        }
        // signed input -> [0, p)   (modular_arith.cuh:372-385)
        static __device__ __forceinline__ T1 reduce(const typename std::make_signed<T1>::type& a, const Modulus<T1>& m)
        {
            return a < 0 ? static_cast<T1>(a) + m.value : static_cast<T1>(a); // two's complement: 2^w + a + p wraps to p - |a|
        }
        // [0, p) -> centred representative (modular_arith.cuh:389-405)
        static __device__ __forceinline__ typename std::make_signed<T1>::type centered_reduction(const T1& a,
                                                                                                const Modulus<T1>& m)
        {
            using S = typename std::make_signed<T1>::type;
            return a > (m.value >> 1) ? static_cast<S>(a - m.value) : static_cast<S>(a);
        }
        static __device__ __forceinline__ T1 reduce_forced(const T1& a, const Modulus<T1>& m)
        {
            T1 r = a;
            while (r >= m.value) r = reduce(r, m);
            return r;
        }
        // two-word input {lo, hi}
        static __device__ __forceinline__ T1 reduce(T1* in, const Modulus<T1>& m)
        {
            if constexpr (std::is_same<T1, Data32>::value)
                return reduce_wide((static_cast<Data64>(in[1]) << 32) | in[0], m);
            else
                return reduce_wide(in[1], in[0], m);
        }

      private:
        // the header also has to parse as plain C++ (host-only translation units include it)
        static __device__ __forceinline__ Data64 mulhi64(Data64 a, Data64 b)
        {
#ifdef __CUDA_ARCH__
            return __umul64hi(a, b);
#else
            return static_cast<Data64>((static_cast<unsigned __int128>(a) * b) >> 64);
#endif
        }
        static __device__ __forceinline__ Data32 reduce_wide(Data64 z, const Modulus<Data32>& m)
        {
            Data64 q = ((z >> (m.bit - 2)) * m.mu) >> (m.bit + 3);
            Data64 r = z - q * m.value;
            while (r >= m.value) r -= m.value;
            return static_cast<Data32>(r);
        }
        static __device__ __forceinline__ Data64 reduce_wide(Data64 hi, Data64 lo, const Modulus<Data64>& m)
        {
            // t = z >> (bit-2)  (fits 64 bits for canonical products: z < 2^(2 bit))
            const int s1 = static_cast<int>(m.bit) - 2;
            const Data64 t = s1 >= 64 ? (hi >> (s1 - 64)) : (s1 == 0 ? lo : ((lo >> s1) | (hi << (64 - s1))));
            // q = (t * mu) >> (bit+3)
            const Data64 ph = mulhi64(t, m.mu), pl = t * m.mu;
            const int s2 = static_cast<int>(m.bit) + 3;
            const Data64 q = s2 >= 64 ? (ph >> (s2 - 64)) : ((pl >> s2) | (ph << (64 - s2)));
            Data64 r = lo - q * m.value;
            while (r >= m.value) r -= m.value;
            return r;
        }
    };
} // namespace modular_operation_gpu

template <typename T> using OPERATOR_GPU = modular_operation_gpu::BarrettOperations<T>;
typedef OPERATOR_GPU<Data32> OPERATOR_GPU_32;
typedef OPERATOR_GPU<Data64> OPERATOR_GPU_64;

#endif // GPUNTT_B200_MODULAR_ARITH_CUH
