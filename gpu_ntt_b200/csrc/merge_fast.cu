// gpu_ntt_b200/csrc/merge_fast.cu -- tuned Merge-NTT entry points for single-modulus calls (unsigned data, PerPolynomial layout):
//   fast_merge   GPU_NTT / GPU_INTT: 64-bit rings 2^7..2^24, 32-bit 2^8..2^26 (small rings in one pass: fast_small)
// Kernels and the launch helper live in fast_kernels.cuh.
#include <cstdlib>

#include "fast_kernels.cuh"

namespace gpuntt_b200
{

    // which (n_power, element width) the fast path covers.  64-bit: the last forward pass is the contiguous
    // 8-stage pass, the n - 8 stages above it are one (n <= 16), two (n <= 24) or three (n <= 28) strided passes of 4..8 stages.
    // 32-bit: contiguous 10-stage pass (two radix-32 rounds on 8192-element tiles), strided passes of 3..8 stages.
    bool fast_supported(int n_power, int element_bits)
    {
        if (element_bits == 64) return n_power >= 12 && n_power <= 28;
        return n_power >= 13 && n_power <= 28;
    }

    // single-pass small rings (fast_small below)
    bool fast_small_supported(int n_power, int element_bits)
    {
        return element_bits == 64 ? (n_power >= 7 && n_power <= 11) : (n_power >= 8 && n_power <= 12);
    }
    // Rings of exactly one tile (64-bit 2^12, 32-bit 2^13): the whole polynomial in ONE tile, three register rounds, one launch and
    // one HBM round trip with nothing handed from CTA to CTA (every twiddle pair of the transform sits in shared memory: 64 KiB
    // beside the two 32 KiB tile buffers, so one CTA per SM with half the resident warps of the two-pass kernel).  Measured
    // (profiles/r2_one_tile_ab.jsonl): inverse transforms gain at every device-bound batch (64-bit +4 %, 32-bit +22 % at
    // 16384 polynomials), forward ones lose 10 %, and in the launch-bound regime the small-tile two-pass kernel is the
    // quicker of the two.  Mode 1 (default): inverse transforms, 64-bit only above the small-tile range; 0: never; 2: always.
    static int one_tile_default()
    {
        const char* e = getenv("GPUNTT_B200_ONE_TILE"); // (A/B runs of binaries that cannot call gpuntt_b200_tune)
        return e ? atoi(e) : 1;
    }
    static std::atomic<int> g_one_tile_mode{one_tile_default()};
    static std::atomic<int> g_single_poly_tiles{1}; // (A/B: GPUNTT_B200_TUNE_SINGLE_POLY_TILES)
    void fast_set_single_poly_tiles(int v) { g_single_poly_tiles.store(v ? 1 : 0); }
    void fast_set_one_tile_mode(int v) { g_one_tile_mode.store(v); }
    long long fused_small_tile_elems(); // merge_fused.cu
    static bool fast_one_tile(int n_power, int element_bits, long long batch, bool inverse)
    {
        if (n_power != (element_bits == 64 ? 12 : 13)) return false;
        const int mode = g_one_tile_mode.load();
        if (mode == 2) return true;
        if (mode != 1 || !inverse) return false;
        // (32-bit: ahead or level at every batch size once the twiddle build issues its table loads in batches: 0.99-1.21x)
        return element_bits == 32 || (batch << n_power) > fused_small_tile_elems();
    }

    // text form of the tuned plan for gpuntt_b200_describe_plan; returns the number of launches (0: not covered)
    int fast_describe(int n_power, int element_bits, char* buf, size_t len)
    {
        if (fast_small_supported(n_power, element_bits))
        {
            snprintf(buf, len, "pass0{tile=2^%d contiguous, whole transforms of 2^%d in a tile, stages=%d in %d register rounds, TMA persistent; "
                               "batches that do not fill whole %d-element chunks: generic kernel} ",
                     element_bits == 64 ? 12 : 13, n_power, n_power, n_power > (element_bits == 64 ? 8 : 10) ? 3 : 2, element_bits == 64 ? 2048 : 4096);
            return 1;
        }
        if (!fast_supported(n_power, element_bits)) return 0;
        const FastPlan pl = make_fast_plan(n_power, element_bits);
        size_t off = 0;
        for (int i = 0; i < pl.npass && off < len; i++)
            off += (size_t) snprintf(buf + off, len - off, "pass%d{tile=2^%d %s lo=%d stages=%d TMA persistent} ", i, element_bits == 64 ? 12 : 13,
                                     pl.strided[i] ? "strided" : "contiguous", pl.lo[i], pl.d[i]);
        return pl.npass;
    }

    // Small rings (replaces ForwardCoreLowRing / InverseCoreLowRing and the one-launch plans, ntt.cu:11-433 of the
    // reference): 2^7 .. 2^11 (64-bit) / 2^8 .. 2^12 (32-bit).  The [batch][N] array is walked as chunks of one tile row
    // group (2048 / 4096 elements); a tile holds whole transforms, so ONE pass with two or three register rounds does
    // every stage, canonicalises (forward) or applies n^-1 (inverse), and the data makes one HBM round trip.
    // Needs batch * N to be a whole number of chunks; anything else stays on the generic kernel.
    template <typename T, bool INV, int POL> static cudaError_t launch_small(int n_power, const FastArgs<T>& s, cudaStream_t st)
    {
        if constexpr (sizeof(T) == 8)
        {
            switch (n_power)
            {
                case 7: return launch_fast<Shape<T, INV, POL, false, 3, 4, 12, 1, 7>>(s, st);
                case 8: return launch_fast<Shape<T, INV, POL, false, 4, 4, 12, 1, 8>>(s, st);
                case 9: return launch_fast<Shape<T, INV, POL, false, 2, 3, 12, 1, 9, 4>>(s, st);
                case 10: return launch_fast<Shape<T, INV, POL, false, 3, 3, 12, 1, 10, 4>>(s, st);
                case 11: return launch_fast<Shape<T, INV, POL, false, 3, 4, 12, 1, 11, 4>>(s, st);
                case 12: return launch_fast<Shape<T, INV, POL, false, 4, 4, 12, 0, 12, 4>>(s, st);
                default: return cudaErrorNotSupported;
            }
        }
        else
        {
            switch (n_power)
            {
                case 8: return launch_fast<Shape<T, INV, POL, false, 3, 5, 13, 1, 8>>(s, st);
                case 9: return launch_fast<Shape<T, INV, POL, false, 4, 5, 13, 1, 9>>(s, st);
                case 10: return launch_fast<Shape<T, INV, POL, false, 5, 5, 13, 1, 10>>(s, st);
                case 11: return launch_fast<Shape<T, INV, POL, false, 3, 3, 13, 1, 11, 5>>(s, st);
                case 12: return launch_fast<Shape<T, INV, POL, false, 3, 4, 13, 1, 12, 5>>(s, st);
                case 13: return launch_fast<Shape<T, INV, POL, false, 4, 4, 13, 0, 13, 5>>(s, st);
                default: return cudaErrorNotSupported;
            }
        }
    }
    // a: table / p / ninv / mu / pbits / plus filled in by fast_merge
    template <typename T>
    static cudaError_t fast_small(const FastArgs<T>& a, const T* in, T* out, int n_power, bool inverse, int batch, cudaStream_t st, int* launched,
                                  void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        // chunk = one row group of a two-group tile; the one-tile rings: chunk = tile = polynomial
        const bool one_tile = n_power == (sizeof(T) == 8 ? 12 : 13);
        const int KC = (sizeof(T) == 8 ? 11 : 12) + (one_tile ? 1 : 0);
        const long long elems = (long long) batch << n_power;
        if (elems & ((1LL << KC) - 1)) return cudaSuccess;
        const long long chunks = elems >> KC;
        if (chunks > 0x7fffffffLL) return cudaSuccess;
        FastArgs<T> s = a;
        s.in = in;
        s.out = out;
        s.n = KC;
        s.n_tw = n_power;
        s.lo = 0;
        s.first = 1;
        s.last = 1;
        s.batch = (int) chunks;
        s.work = one_tile ? chunks : (chunks + 1) >> 1;
        // arithmetic policy: 64-bit forward F60, inverse lazy; 32-bit forward lazy (p <= 2^29), inverse exact
        bool covered;
        if constexpr (sizeof(T) == 8)
            covered = inverse ? ((uint64_t) a.p < kFastModulusLimit) : ((uint64_t) a.p >= kF60ModulusMin && (uint64_t) a.p < kF60ModulusLimit);
        else
            covered = inverse || (uint32_t) a.p < kL32ModulusLimit;
        cudaError_t e;
        prof_begin(1, st);
        if (!covered) // exact policy: any modulus the reference accepts
            e = inverse ? launch_small<T, true, 0>(n_power, s, st) : launch_small<T, false, 0>(n_power, s, st);
        else if constexpr (sizeof(T) == 8)
            e = inverse ? launch_small<T, true, 1>(n_power, s, st) : launch_small<T, false, 2>(n_power, s, st);
        else
            e = inverse ? launch_small<T, true, 0>(n_power, s, st) : launch_small<T, false, 2>(n_power, s, st);
        prof_end(st);
        if (e == cudaErrorNotSupported) return cudaSuccess; // no tensor maps: generic path
        if (e != cudaSuccess) return e;
        *launched = 1;
        return cudaSuccess;
    }

    // Returns cudaSuccess and sets *launched to the number of kernels, or *launched = 0 if this
    // transform is not covered (the caller then uses the generic path).
    template <typename T>
    cudaError_t fast_merge(const T* in, T* out, const T* table, T p, T ninv, int n_power, int plus, bool inverse,
                           int batch, cudaStream_t st, int* launched, void (*prof_begin)(int, cudaStream_t),
                           void (*prof_end)(cudaStream_t), int in_bound, unsigned* counters, int signed_io)
    {
        *launched = 0;
        if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        if (fast_small_supported(n_power, (int) sizeof(T) * 8) || fast_one_tile(n_power, (int) sizeof(T) * 8, batch, inverse))
        {
            if (sizeof(T) == 4 && ((uint32_t) p >= (1u << 30) || (uint32_t) p < 3)) return cudaSuccess;
            if (sizeof(T) == 8 && ((uint64_t) p >= (1ull << 62) || (uint64_t) p < 3)) return cudaSuccess;
            FastArgs<T> a{};
            a.table = table;
            a.p = p;
            a.ninv_w = ninv;
            a.ninv_wq = inverse ? shoup_companion(ninv, p) : 0;
            if constexpr (sizeof(T) == 8)
            {
                a.pbits = 64 - __builtin_clzll((unsigned long long) p);
                const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
                a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
            }
            else
            {
                a.pbits = 32 - __builtin_clz((unsigned) p);
                a.mu = ~0ull / (uint64_t) p;
            }
            a.plus = plus;
            a.in_bound = 1;
            a.signed_io = signed_io;
            const cudaError_t e = fast_small<T>(a, in, out, n_power, inverse, batch, st, launched, prof_begin, prof_end);
            if (e != cudaSuccess || *launched > 0 || fast_small_supported(n_power, (int) sizeof(T) * 8)) return e;
            // (a one-tile ring the single-tile kernel declined: the two-pass plan below)
        }
        if (!fast_supported(n_power, (int) sizeof(T) * 8)) return cudaSuccess;
        // TMA coordinates are signed 32-bit: the matrix-row index of the finest strided pass must fit
        if (((long long) batch << (n_power - (sizeof(T) == 8 ? 8 : 10))) >= (1LL << 31)) return cudaSuccess;
        if constexpr (sizeof(T) == 8)
        {
            // forward: F60 policy for 2^40 <= p < 2^60 - 2^31, the every-stage lazy policy for the other moduli in
            // [2^36, 1.25 * 2^60) (61-bit primes); inverse: the lazy policy (exact range test, any p below 1.25 * 2^60).
            // Other moduli: exact-policy kernels exist for n = 16 only, the rest goes to the generic kernel.
            const bool f60 = (uint64_t) p >= kF60ModulusMin && (uint64_t) p < kF60ModulusLimit;
            const bool lazy_fwd = !f60 && (uint64_t) p >= kFastModulusMin && (uint64_t) p < kFastModulusLimit;
            const bool fast_inv = (uint64_t) p < kFastModulusLimit;
            const bool fast_arith = inverse ? fast_inv : (f60 || lazy_fwd);
            if (!fast_arith && (uint64_t) p >= (1ull << 62)) return cudaSuccess; // (the exact policy needs 4p < 2^64, like the reference)
            FastArgs<T> a{};
            a.table = table;
            a.p = p;
            a.ninv_w = ninv;
            a.ninv_wq = inverse ? shoup_companion(ninv, p) : 0;
            a.pbits = 64 - __builtin_clzll((unsigned long long) p);
            {
                const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
                a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
            }
            a.n = n_power;
            a.plus = plus;
            a.batch = batch;
            a.in_bound = in_bound;
            a.signed_io = signed_io;
            using Cf = Shape<T, false, 2, false, 4, 4, 12, GPUNTT_FAST_P1_NPLOG>;
            using Ci = Shape<T, true, 1, false, 4, 4, 12, 1>;
            using Cfx = Shape<T, false, 0, false, 4, 4, 12, 1>;
            using Cix = Shape<T, true, 0, false, 4, 4, 12, 1>;
            const FastPlan pl = make_fast_plan(n_power, 64);
            if (counters != nullptr && pl.npass == 2)
            {
                // one launch, the two passes chained through the L2 (merge_fused.cu)
                FastArgs<T> s = a;
                s.in = in;
                s.out = out;
                const cudaError_t e = fused_merge<T>(s, pl, inverse, f60, fast_inv, counters, st, prof_begin, prof_end);
                if (e == cudaSuccess)
                {
                    *launched = 1;
                    return cudaSuccess;
                }
                if (e != cudaErrorNotSupported) return e;
            }
            for (int k = 0; k < pl.npass; k++)
            {
                const int i = inverse ? pl.npass - 1 - k : k; // pass of the forward plan executed k-th
                FastArgs<T> s = a;
                s.in = (k == 0) ? in : out;
                s.out = out;
                s.lo = pl.lo[i];
                s.first = (k == 0);
                s.last = (k == pl.npass - 1);
                cudaError_t e;
                prof_begin(k + 1, st);
                if (pl.strided[i])
                {
                    const int c = 12 - pl.d[i];
                    s.work = ((long long) batch << (pl.lo[i] - c)) << (n_power - pl.lo[i] - pl.d[i]);
                    s.rr = (n_power == pl.lo[i] + pl.d[i]) && pl.lo[i] > 10; // one range, long row stride
                    if (fast_arith)
                        e = inverse ? launch_strided<T, true, 1>(pl.d[i], s, st)
                                    : (f60 ? launch_strided<T, false, 2>(pl.d[i], s, st) : launch_strided<T, false, 1>(pl.d[i], s, st));
                    else
                        e = inverse ? launch_strided<T, true, 0>(pl.d[i], s, st) : launch_strided<T, false, 0>(pl.d[i], s, st);
                }
                else if (batch == 1 && pl.npass >= 3 && (inverse ? fast_inv : f60) && g_single_poly_tiles.load())
                {
                    // ONE polynomial of a large ring (the shape of a ZK prover's transform): every tile of this pass has its own
                    // twiddle set, and half of a two-polynomial tile would be zero fill.  2048-element tiles of one polynomial:
                    // the same 2040 pairs per tile, half the butterflies.
                    s.work = 1LL << (n_power - 11);
                    e = inverse ? launch_fast<Shape<T, true, 1, false, 4, 4, 11, 0>>(s, st) : launch_fast<Shape<T, false, 2, false, 4, 4, 11, 0>>(s, st);
                }
                else
                {
                    const int nplog = (!inverse && f60) ? Cf::NPLOG : 1;
                    const long long tpr = (batch + (1 << nplog) - 1) >> nplog;
                    s.work = tpr << (n_power - (12 - nplog));
                    if (fast_arith)
                        e = inverse ? launch_fast<Ci>(s, st) : (f60 ? launch_fast<Cf>(s, st) : launch_fast<Shape<T, false, 1, false, 4, 4, 12, 1>>(s, st));
                    else
                        e = inverse ? launch_fast<Cix>(s, st) : launch_fast<Cfx>(s, st);
                }
                prof_end(st);
                if (e == cudaErrorNotSupported && k == 0) return cudaSuccess; // no tensor maps: generic path
                if (e != cudaSuccess) return e;
            }
            *launched = pl.npass;
        }
        else
        {
            // 32-bit: exact Harvey butterflies ([0,4p) forward, [0,2p) inverse; p < 2^30 like the reference)
            if ((uint32_t) p >= (1u << 30) || (uint32_t) p < 3) return cudaSuccess;
            FastArgs<T> a{};
            a.table = table;
            a.p = p;
            a.ninv_w = ninv;
            a.ninv_wq = inverse ? shoup_companion(ninv, p) : 0;
            a.pbits = 32 - __builtin_clz((unsigned) p);
            a.mu = ~0ull / (uint64_t) p; // floor((2^64 - 1) / p) = floor(2^64 / p) unless p divides 2^64 (p is odd here)
            a.n = n_power;
            a.plus = plus;
            a.batch = batch;
            a.signed_io = signed_io;
            using Cf = Shape<T, false, 0, false, 5, 5, 13, 1>;
            using Cl = Shape<T, false, 2, false, 5, 5, 13, 1>; // lazy forward policy (p <= 2^29)
            using Ci = Shape<T, true, 0, false, 5, 5, 13, 1>;
            const bool lazy = !inverse && (uint32_t) p < kL32ModulusLimit;
            const FastPlan pl = make_fast_plan(n_power, 32);
            if (counters != nullptr && pl.npass == 2)
            {
                FastArgs<T> s = a;
                s.in = in;
                s.out = out;
                const cudaError_t e = fused_merge<T>(s, pl, inverse, lazy, true, counters, st, prof_begin, prof_end);
                if (e == cudaSuccess)
                {
                    *launched = 1;
                    return cudaSuccess;
                }
                if (e != cudaErrorNotSupported) return e;
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
                cudaError_t e;
                prof_begin(k + 1, st);
                if (pl.strided[i])
                {
                    const int c = 13 - pl.d[i];
                    s.work = ((long long) batch << (pl.lo[i] - c)) << (n_power - pl.lo[i] - pl.d[i]);
                    s.rr = (n_power == pl.lo[i] + pl.d[i]) && pl.lo[i] > 12;
                    e = inverse ? launch_strided32<true>(pl.d[i], s, st)
                                : (lazy ? launch_strided32<false, 2>(pl.d[i], s, st) : launch_strided32<false>(pl.d[i], s, st));
                }
                else if (batch == 1 && pl.npass >= 3 && g_single_poly_tiles.load())
                {
                    // one polynomial of a large ring: 4096-element tiles of that polynomial (see the 64-bit branch)
                    s.work = 1LL << (n_power - 12);
                    e = inverse ? launch_fast<Shape<T, true, 0, false, 5, 5, 12, 0>>(s, st)
                                : (lazy ? launch_fast<Shape<T, false, 2, false, 5, 5, 12, 0>>(s, st) : launch_fast<Shape<T, false, 0, false, 5, 5, 12, 0>>(s, st));
                }
                else
                {
                    const long long tpr = (batch + 1) >> 1;
                    s.work = tpr << (n_power - 12);
                    e = inverse ? launch_fast<Ci>(s, st) : (lazy ? launch_fast<Cl>(s, st) : launch_fast<Cf>(s, st));
                }
                prof_end(st);
                if (e == cudaErrorNotSupported && k == 0) return cudaSuccess;
                if (e != cudaSuccess) return e;
            }
            *launched = pl.npass;
        }
        return cudaSuccess;
    }

    template cudaError_t fast_merge<uint64_t>(const uint64_t*, uint64_t*, const uint64_t*, uint64_t, uint64_t, int, int, bool,
                                              int, cudaStream_t, int*, void (*)(int, cudaStream_t), void (*)(cudaStream_t), int, unsigned*, int);
    template cudaError_t fast_merge<uint32_t>(const uint32_t*, uint32_t*, const uint32_t*, uint32_t, uint32_t, int, int, bool,
                                              int, cudaStream_t, int*, void (*)(int, cudaStream_t), void (*)(cudaStream_t), int, unsigned*, int);

} // namespace gpuntt_b200
