// gpu_ntt_b200/csrc/merge_fast_rns.cu -- tuned Merge-NTT entry point for the RNS overloads:
//   fast_merge_rns   GPU_NTT / GPU_INTT with Modulus*, GPU_NTT_Modulus_Ordered, GPU_NTT_Poly_Ordered
// Kernels and the launch helper live in fast_kernels.cuh.
#include "fast_kernels.cuh"

namespace gpuntt_b200
{

    // ------------------------------------------------------------------ RNS form on the tuned kernels
    template <typename T, bool INV, int POL> static cudaError_t launch_strided_rns(int d, const FastArgs<T>& args, cudaStream_t st)
    {
        constexpr int K = sizeof(T) == 8 ? 12 : 13;
        switch (d)
        {
            case 4: return launch_fast<Shape<T, INV, POL, true, 4, 0, K, 0>, false, true>(args, st);
            case 5: return launch_fast<Shape<T, INV, POL, true, 3, 2, K, 0>, false, true>(args, st);
            case 6: return launch_fast<Shape<T, INV, POL, true, 3, 3, K, 0>, false, true>(args, st);
            case 7: return launch_fast<Shape<T, INV, POL, true, 4, 3, K, 0>, false, true>(args, st);
            case 8: return launch_fast<Shape<T, INV, POL, true, 4, 4, K, 0>, false, true>(args, st);
            default: return cudaErrorNotSupported;
        }
    }
    // 64-bit: lazy policy POL and the exact policy in one launch (the device picks)
    template <bool INV, int POL> static cudaError_t launch_strided_rns_dual(int d, const FastArgs<uint64_t>& args, cudaStream_t st)
    {
        using T = uint64_t;
        switch (d)
        {
            case 4: return launch_fast<Shape<T, INV, POL, true, 4, 0, 12, 0>, false, true, Shape<T, INV, 0, true, 4, 0, 12, 0>>(args, st);
            case 5: return launch_fast<Shape<T, INV, POL, true, 3, 2, 12, 0>, false, true, Shape<T, INV, 0, true, 3, 2, 12, 0>>(args, st);
            case 6: return launch_fast<Shape<T, INV, POL, true, 3, 3, 12, 0>, false, true, Shape<T, INV, 0, true, 3, 3, 12, 0>>(args, st);
            case 7: return launch_fast<Shape<T, INV, POL, true, 4, 3, 12, 0>, false, true, Shape<T, INV, 0, true, 4, 3, 12, 0>>(args, st);
            case 8: return launch_fast<Shape<T, INV, POL, true, 4, 4, 12, 0>, false, true, Shape<T, INV, 0, true, 4, 4, 12, 0>>(args, st);
            default: return cudaErrorNotSupported;
        }
    }

    bool fast_small_supported(int n_power, int element_bits); // merge_fast.cu

    // Small rings (64-bit 2^7..2^11, 32-bit 2^8..2^12) in the RNS form: ONE contiguous pass whose tiles hold 2^(K - n) whole
    // polynomials of ONE modulus slot (4-D tensor map {row, rows, slot, polynomial within the slot}: the polynomials of a slot are
    // mod_count apart in the caller's array, ntt.cu:613-619 of the reference), two or three register rounds, a segment per slot.
    template <typename T, bool INV> static cudaError_t launch_small_rns(int n_power, const FastArgs<T>& s, cudaStream_t st)
    {
        if constexpr (sizeof(T) == 8)
        {
            constexpr int PL = INV ? 1 : 2; // lazy inverse / F60 forward; the exact twin is picked on the device
#define GPUNTT_SMALL_RNS64(R1, R2, NP, NT, R3)                                                                                                    \
    launch_fast<Shape<T, INV, PL, false, R1, R2, 12, NP, NT, R3>, false, true, Shape<T, INV, 0, false, R1, R2, 12, NP, NT, R3>>(s, st)
            switch (n_power)
            {
                case 7: return GPUNTT_SMALL_RNS64(3, 4, 5, 7, 0);
                case 8: return GPUNTT_SMALL_RNS64(4, 4, 4, 8, 0);
                case 9: return GPUNTT_SMALL_RNS64(2, 3, 3, 9, 4);
                case 10: return GPUNTT_SMALL_RNS64(3, 3, 2, 10, 4);
                case 11: return GPUNTT_SMALL_RNS64(3, 4, 1, 11, 4);
                default: return cudaErrorNotSupported;
            }
#undef GPUNTT_SMALL_RNS64
        }
        else
        {
            switch (n_power)
            {
                case 8: return launch_fast<Shape<T, INV, 0, false, 3, 5, 13, 5, 8>, false, true>(s, st);
                case 9: return launch_fast<Shape<T, INV, 0, false, 4, 5, 13, 4, 9>, false, true>(s, st);
                case 10: return launch_fast<Shape<T, INV, 0, false, 5, 5, 13, 3, 10>, false, true>(s, st);
                case 11: return launch_fast<Shape<T, INV, 0, false, 3, 3, 13, 2, 11, 5>, false, true>(s, st);
                case 12: return launch_fast<Shape<T, INV, 0, false, 3, 4, 13, 1, 12, 5>, false, true>(s, st);
                default: return cudaErrorNotSupported;
            }
        }
    }

    // GPU_NTT / GPU_INTT RNS overloads (ntt.cu:2560-3058) for the small rings (one pass, above) and the two- and three-pass ring
    // sizes, batch a multiple of mod_count.  The moduli are device data, so for 64-bit every pass is ONE launch of
    // fast_pass_dual_kernel, which holds the lazy-policy and the exact-policy body and picks from the modulus array.
    // flag_ws: unused (kept for the call signature).  *launched = 0 when not covered.
    template <typename T>
    cudaError_t fast_merge_rns(const T* in, T* out, const T* table, const T* mod_dev, const T* ninv_dev, const int* mod_order,
                               const int* poly_order, int mod_count, int n_power, int plus, bool inverse, int batch, int* flag_ws,
                               cudaStream_t st, int* launched,
                               void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        *launched = 0;
        constexpr int bits = (int) sizeof(T) * 8;
        constexpr int K = bits == 64 ? 12 : 13;
        if (fast_small_supported(n_power, bits) && mod_count >= 1 && batch % mod_count == 0 &&
            ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
        {
            FastArgs<T> s{};
            s.in = in;
            s.out = out;
            s.table = table;
            s.p = (T) ((1ull << (bits - 5)) + 1); // placeholder until the first segment reads its modulus
            s.n = n_power;
            s.lo = 0;
            s.plus = plus;
            s.first = 1;
            s.last = 1;
            s.in_bound = 1;
            s.batch = batch / mod_count;
            s.mod_count = mod_count;
            s.mod_dev = mod_dev;
            s.ninv_dev = ninv_dev;
            s.mod_order = mod_order;
            s.poly_order = poly_order;
            s.policy_flag = nullptr;
            const int nplog = K - n_power;
            s.work = (long long) mod_count * (((long long) s.batch + (1 << nplog) - 1) >> nplog);
            prof_begin(1, st);
            const cudaError_t e = inverse ? launch_small_rns<T, true>(n_power, s, st) : launch_small_rns<T, false>(n_power, s, st);
            prof_end(st);
            if (e == cudaErrorNotSupported) return cudaSuccess; // generic path
            if (e != cudaSuccess) return e;
            *launched = 1;
            return cudaSuccess;
        }
        if (!fast_supported(n_power, bits) || mod_count < 1 || batch % mod_count != 0) return cudaSuccess;
        if (((long long) batch << (n_power - (bits == 64 ? 8 : 10))) >= (1LL << 31)) return cudaSuccess;
        const FastPlan pl = make_fast_plan(n_power, bits);
        for (int i = 0; i + 1 < pl.npass; i++)
            if (pl.d[i] < 4 || pl.d[i] > 8) return cudaSuccess; // strided RNS shapes exist for 4..8 stages
        if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        FastArgs<T> a{};
        a.table = table;
        a.p = (T) ((1ull << (bits - 5)) + 1); // placeholder until the first segment reads its modulus
        a.n = n_power;
        a.plus = plus;
        a.batch = batch / mod_count;
        a.mod_count = mod_count;
        a.mod_dev = mod_dev;
        a.ninv_dev = ninv_dev;
        a.mod_order = mod_order;
        a.poly_order = poly_order;
        a.policy_flag = nullptr;
        int kind = 1;
        if (flag_ws != nullptr && pl.npass == 2)
        {
            // one launch, the two passes chained through the L2 (merge_fused.cu); flag_ws = the per-polynomial counters
            FastArgs<T> s = a;
            s.in = in;
            s.out = out;
            const cudaError_t fe = fused_merge_rns<T>(s, pl, inverse, reinterpret_cast<unsigned*>(flag_ws), st, prof_begin, prof_end);
            if (fe == cudaSuccess)
            {
                *launched = 1;
                return cudaSuccess;
            }
            if (fe != cudaErrorNotSupported) return fe;
        }
        for (int k = 0; k < pl.npass; k++)
        {
            const int i = inverse ? pl.npass - 1 - k : k;
            FastArgs<T> s = a;
            s.in = (k == 0) ? in : out;
            s.out = out;
            s.lo = pl.lo[i];
            s.first = (k == 0);
            s.last = (k == pl.npass - 1);
            cudaError_t e = cudaSuccess;
            prof_begin(kind, st);
            if (pl.strided[i])
            {
                const int c = K - pl.d[i];
                // (slot, range) segments: 2^(n - lo - d) ranges per slot (one in a two-pass plan)
                s.work = (((long long) mod_count * s.batch) << (pl.lo[i] - c)) << (n_power - pl.lo[i] - pl.d[i]);
                if constexpr (bits == 64)
                    e = inverse ? launch_strided_rns_dual<true, 1>(pl.d[i], s, st) : launch_strided_rns_dual<false, 2>(pl.d[i], s, st);
                else
                    e = inverse ? launch_strided_rns<T, true, 0>(pl.d[i], s, st) : launch_strided_rns<T, false, 0>(pl.d[i], s, st);
            }
            else
            {
                const long long tpr = (s.batch + 1) >> 1;
                s.work = ((long long) mod_count * tpr) << (n_power - (K - 1));
                if constexpr (bits == 64)
                    e = inverse ? launch_fast<Shape<T, true, 1, false, 4, 4, 12, 1>, false, true, Shape<T, true, 0, false, 4, 4, 12, 1>>(s, st)
                                : launch_fast<Shape<T, false, 2, false, 4, 4, 12, 1>, false, true, Shape<T, false, 0, false, 4, 4, 12, 1>>(s, st);
                else
                    e = inverse ? launch_fast<Shape<T, true, 0, false, 5, 5, 13, 1>, false, true>(s, st)
                                : launch_fast<Shape<T, false, 0, false, 5, 5, 13, 1>, false, true>(s, st);
            }
            prof_end(st);
            kind++;
            if (e == cudaErrorNotSupported && k == 0) return cudaSuccess;
            if (e != cudaSuccess) return e;
        }
        *launched = pl.npass;
        return cudaSuccess;
    }
    template cudaError_t fast_merge_rns<uint64_t>(const uint64_t*, uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, const int*, const int*,
                                                  int, int, int, bool, int, int*, cudaStream_t, int*, void (*)(int, cudaStream_t),
                                                  void (*)(cudaStream_t));
    template cudaError_t fast_merge_rns<uint32_t>(const uint32_t*, uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, const int*, const int*,
                                                  int, int, int, bool, int, int*, cudaStream_t, int*, void (*)(int, cudaStream_t),
                                                  void (*)(cudaStream_t));

} // namespace gpuntt_b200
