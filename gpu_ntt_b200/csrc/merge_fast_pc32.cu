// gpu_ntt_b200/csrc/merge_fast_pc32.cu -- NTTLayout::PerCoefficient for 32-bit data on the tuned kernels (replaces
// ForwardCoreTranspose / InverseCoreTranspose<Data32>, ntt.cu:1554-2074 of the reference; like it: power-of-two batch, n_power <= 9).
// Same idea as the 64-bit form in merge_fast_4step.cu: the buffer is one [2^n_power][2^col_log] row-major matrix whose COLUMNS are
// the transforms, i.e. strided passes over the top n_power index bits of one array of 2^(n_power + col_log) elements.  8192-element
// tiles (2^d rows of 2^(13 - d) adjacent columns); one pass up to 2^8, two for 2^9 (5 + 4 stages).  The pass that ends a forward
// transform canonicalises (SFIN), the one that ends an inverse applies n^-1.  Forward: lazy policy for p <= 2^29, exact above (the
// reference takes Data32 moduli up to 30 bits, modular_arith.cuh:66); inverse: exact.  *launched = 0: the caller takes the generic kernel.
#include "fast_kernels.cuh"

namespace gpuntt_b200
{

    template <int POL> static cudaError_t launch_strided32_final(int d, const FastArgs<uint32_t>& args, cudaStream_t st)
    {
        using T = uint32_t;
        switch (d)
        {
            case 3: return launch_fast<Shape<T, false, POL, true, 3, 0, 13, 0>, false, false, void, true>(args, st);
            case 4: return launch_fast<Shape<T, false, POL, true, 4, 0, 13, 0>, false, false, void, true>(args, st);
            case 5: return launch_fast<Shape<T, false, POL, true, 5, 0, 13, 0>, false, false, void, true>(args, st);
            case 6: return launch_fast<Shape<T, false, POL, true, 3, 3, 13, 0>, false, false, void, true>(args, st);
            case 7: return launch_fast<Shape<T, false, POL, true, 4, 3, 13, 0>, false, false, void, true>(args, st);
            case 8: return launch_fast<Shape<T, false, POL, true, 4, 4, 13, 0>, false, false, void, true>(args, st);
            default: return cudaErrorNotSupported;
        }
    }

    cudaError_t fast_per_coefficient32(const uint32_t* in, uint32_t* out, const uint32_t* table, uint32_t p, uint32_t ninv, int n_power,
                                       int col_log, int plus, bool inverse, int signed_io, cudaStream_t st, int* launched,
                                       void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = uint32_t;
        *launched = 0;
        if (n_power < 3 || n_power > 9 || col_log < 1) return cudaSuccess;
        const int n = n_power + col_log;
        if (n > 40) return cudaSuccess;
        if (p >= (1u << 30) || p < 3) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        // stages of the pass on the high bits / on the low bits (forward order)
        const int da = n_power <= 8 ? n_power : 5, db = n_power - da;
        if (13 - da > col_log + db || (db > 0 && 13 - db > col_log)) return cudaSuccess; // a tile is 2^(13 - stages) adjacent columns wide
        const bool lazy = !inverse && p < kL32ModulusLimit;
        FastArgs<T> a{};
        a.table = table;
        a.p = p;
        a.ninv_w = ninv;
        a.ninv_wq = inverse ? shoup_companion(ninv, p) : 0;
        a.pbits = 32 - __builtin_clz((unsigned) p);
        a.mu = ~0ull / (uint64_t) p;
        a.n = n;
        a.plus = plus;
        a.batch = 1;
        a.in_bound = 1;
        a.signed_io = signed_io;
        auto pass = [&](int d, int lo, bool first, bool last, const T* src) -> cudaError_t
        {
            FastArgs<T> s = a;
            s.in = src;
            s.out = out;
            s.lo = lo;
            s.first = first ? 1 : 0;
            s.last = last ? 1 : 0;
            s.work = (1LL << (lo - (13 - d))) << (n - lo - d);
            s.rr = (n == lo + d && lo > 12) ? 1 : 0;
            if (inverse) return launch_strided32<true>(d, s, st);
            if (last) return lazy ? launch_strided32_final<2>(d, s, st) : launch_strided32_final<0>(d, s, st);
            if (d != 5) return cudaErrorNotSupported; // (a forward pass that does not end the transform: only the 5-stage top pass of 2^9)
            return lazy ? launch_fast<Shape<T, false, 2, true, 5, 0, 13, 0>>(s, st) : launch_fast<Shape<T, false, 0, true, 5, 0, 13, 0>>(s, st);
        };
        cudaError_t e;
        int k = 0;
        if (db == 0)
        {
            prof_begin(++k, st);
            e = pass(da, col_log, true, true, in);
            prof_end(st);
        }
        else if (!inverse)
        {
            prof_begin(++k, st);
            e = pass(da, col_log + db, true, false, in);
            prof_end(st);
            if (e == cudaSuccess)
            {
                prof_begin(++k, st);
                e = pass(db, col_log, false, true, out);
                prof_end(st);
            }
        }
        else
        {
            prof_begin(++k, st);
            e = pass(db, col_log, true, false, in);
            prof_end(st);
            if (e == cudaSuccess)
            {
                prof_begin(++k, st);
                e = pass(da, col_log + db, false, true, out);
                prof_end(st);
            }
        }
        if (e == cudaErrorNotSupported && k == 1) return cudaSuccess; // no tensor maps: generic path
        if (e != cudaSuccess) return e;
        *launched = k;
        return cudaSuccess;
    }

} // namespace gpuntt_b200
