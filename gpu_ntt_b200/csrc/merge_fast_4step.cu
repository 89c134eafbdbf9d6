// gpu_ntt_b200/csrc/merge_fast_4step.cu -- 4-step phases on the tuned kernels:
//   fast_fourstep_columns                forward column phase on the natural matrix, twiddle-matrix product as epilogue
//   fast_fourstep_rows_t                 forward row phase along the n2 x n1 layout (optionally storing the n1 x n2 matrix)
//   fast_fourstep_forward_transposed_in  forward, reference contract: contiguous column phase on the caller's transposed input
//   fast_fourstep_inverse                inverse, either contract (size-n1 pass, product pass, last pass)
//   fast_per_coefficient                 NTTLayout::PerCoefficient as strided passes of the same kernels
// Kernels and the launch helper live in fast_kernels.cuh, the position-major product kernel in merge_wcol.cu.
#include "fast_kernels.cuh"

namespace gpuntt_b200
{

    // (w, w') pairs of the 4-step twiddle matrix, once per call (the batch shares it)
    // t_lo > 0: entry i comes from the transposed index ((i mod 2^t_lo) << t_hi) | (i >> t_lo) (4-step inverse: the data
    // is the n2 x n1 matrix, the reference's inverse twiddle matrix is laid out n1 x n2)
    __global__ void __launch_bounds__(256) w_pairs_kernel(const uint64_t* __restrict__ w, Twiddle<uint64_t>* __restrict__ out, long long count,
                                                          uint64_t p, uint64_t mu, int pbits, int t_lo, int t_hi)
    {
        const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= count) return;
        const long long src = t_lo ? (((i & ((1LL << t_lo) - 1)) << t_hi) | (i >> t_lo)) : i;
        const uint64_t v = w[src];
        out[i] = Twiddle<uint64_t>{v, shoup_companion_mu(v, p, mu, pbits)};
    }

    // The same pairs for the forward pass over the TRANSPOSED input (reference contract; merge_wcol.cu, contiguous form).  The data is
    // the n2 x n1 matrix (element (i, j) of the n1 x n2 twiddle matrix sits at flat offset f = j * n1 + i) and is walked as positions
    // of 4096 consecutive elements; inside a position the table is element-major -- entry a * 256 + item is the pair of local element
    // item * 16 + a -- so the 32 threads of a warp, which hold 32 consecutive items, read 32 consecutive pairs per butterfly input.
    __global__ void __launch_bounds__(256) w_pairs_tile_kernel(const uint64_t* __restrict__ w, Twiddle<uint64_t>* __restrict__ out, long long count,
                                                               uint64_t p, uint64_t mu, int pbits, int lg1, int lg2)
    {
        const long long o = (long long) blockIdx.x * blockDim.x + threadIdx.x;
        if (o >= count) return;
        const int r = (int) (o & 4095);
        const long long f = (o & ~4095LL) | (long long) (((r & 255) << 4) | (r >> 8));
        const long long src = ((f & ((1LL << lg1) - 1)) << lg2) | (f >> lg1);
        const uint64_t v = w[src];
        out[o] = Twiddle<uint64_t>{v, shoup_companion_mu(v, p, mu, pbits)};
    }

    // merge_wcol.cu (cudaErrorNotSupported: batch below 4 or no tensor maps -- the caller takes fast_pass_kernel<..., WMUL>)
    cudaError_t fourstep_columns_resident_pairs(const FastArgs<uint64_t>& a, int lg1, cudaStream_t st);
    cudaError_t fourstep_rows_of_transposed_resident_pairs(const FastArgs<uint64_t>& a, int lg1, cudaStream_t st);
    cudaError_t fourstep_inverse_product_pass_resident_pairs(const FastArgs<uint64_t>& a, int d, bool transposed, cudaStream_t st);
    static std::atomic<int> g_resident_pairs{1};
    void fourstep_set_resident_pairs(int on) { g_resident_pairs.store(on ? 1 : 0); }

    // Forward 4-step column phase on the tuned strided kernel: the first lg1 stages of a size-2^n transform with the
    // n1 table (rows 2^lg2 elements apart), then every element times W[offset] (pairs built into w_pairs_ws, N
    // entries), canonical outputs.  Single modulus, 64-bit, F60 moduli; *launched = 0 when not covered.
    cudaError_t fast_fourstep_columns(const uint64_t* in, uint64_t* out, const uint64_t* n1_table, const uint64_t* w_table,
                                      void* w_pairs_ws, uint64_t p, int n_power, int lg1, int lg2, int batch, cudaStream_t st,
                                      int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t), int w_lazy,
                                      int transposed)
    {
        using T = uint64_t;
        *launched = 0;
        if (transposed && in == out) return cudaSuccess;
        if (lg1 < 5 || lg1 > 8 || lg2 < 12 - lg1 || n_power != lg1 + lg2) return cudaSuccess;
        if (((long long) batch << lg1) >= (1LL << 31)) return cudaSuccess;
        if (!(p >= kF60ModulusMin && p < kF60ModulusLimit)) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        FastArgs<T> a{};
        a.in = in;
        a.out = out;
        a.table = n1_table;
        a.p = p;
        a.pbits = 64 - __builtin_clzll((unsigned long long) p);
        {
            const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
            a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
        }
        a.n = n_power;
        a.lo = lg2;
        a.plus = 0;
        a.first = 1;
        a.last = 0;
        a.batch = batch;
        a.w_pairs = w_pairs_ws;
        a.w_lazy = w_lazy; // products below 2p (the tuned row phase starts from that bound) or canonical
        a.work = (long long) batch << (lg2 - (12 - lg1));
        a.rr = 1; // column-chunk-major round robin: the polynomials of a chunk share the pair fetch through the L2
        const long long N = 1LL << n_power;
        prof_begin(0, st);
        w_pairs_kernel<<<(unsigned) ((N + 255) / 256), 256, 0, st>>>(w_table, reinterpret_cast<Twiddle<T>*>(w_pairs_ws), N, p, a.mu, a.pbits, 0, 0);
        prof_end(st);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        prof_begin(1, st);
        bool resident = false;
        if (transposed && g_resident_pairs.load())
        {
            // position-major kernel with the pairs of a tile position resident in shared memory (merge_wcol.cu)
            e = fourstep_columns_resident_pairs(a, lg1, st);
            resident = e != cudaErrorNotSupported; // (not supported: the per-tile kernel below)
        }
        if (resident)
            ;
        else if (transposed) // the pass writes the n2 x n1 matrix (every column transform becomes a contiguous row)
            switch (lg1)
            {
                case 5: e = launch_fast<Shape<T, false, 2, true, 3, 2, 12, 0>, true, false, void, false, true>(a, st); break;
                case 6: e = launch_fast<Shape<T, false, 2, true, 3, 3, 12, 0>, true, false, void, false, true>(a, st); break;
                case 7: e = launch_fast<Shape<T, false, 2, true, 4, 3, 12, 0>, true, false, void, false, true>(a, st); break;
                default: e = launch_fast<Shape<T, false, 2, true, 4, 4, 12, 0>, true, false, void, false, true>(a, st); break;
            }
        else
            switch (lg1)
            {
                case 5: e = launch_fast<Shape<T, false, 2, true, 3, 2, 12, 0>, true>(a, st); break;
                case 6: e = launch_fast<Shape<T, false, 2, true, 3, 3, 12, 0>, true>(a, st); break;
                case 7: e = launch_fast<Shape<T, false, 2, true, 4, 3, 12, 0>, true>(a, st); break;
                default: e = launch_fast<Shape<T, false, 2, true, 4, 4, 12, 0>, true>(a, st); break;
            }
        prof_end(st);
        if (e == cudaErrorNotSupported) return cudaSuccess; // (the pair table was written for nothing)
        if (e != cudaSuccess) return e;
        *launched = 2;
        return cudaSuccess;
    }

    // Forward 4-step row phase on the TRANSPOSED layout: `buf` holds, per polynomial, the n2 x n1 matrix the transposing
    // column pass wrote (row j = the n1 outputs of column transform j, already multiplied by W).  The size-n2 transforms
    // run along j, i.e. over index bits [lg1, n): one or two strided passes with the n2 table, the last one canonical.
    // Their output order IS the reference's final order (out[j' * n1 + i'], ntt_4step_cpu.cu:33-109 of the reference),
    // so no transpose kernel runs anywhere.  in_bound: the column products are below in_bound * p.
    // split: lg2 = da + db, db in {7, 8} (the low pass needs 2^(12 - db) <= n1 adjacent elements), da in {0, 4..8}.
    bool fast_fourstep_rows_t_supported(int lg1, int lg2)
    {
        if (lg1 < 5 || lg1 > 8) return false;
        if (lg2 == 7 || lg2 == 8) return true;
        const int db = lg2 - 8 >= 4 ? 8 : 7, da = lg2 - db;
        return da >= 4 && da <= 8;
    }
    template <int D, bool SFIN, bool TS = false> static cudaError_t launch_rows_t(const FastArgs<uint64_t>& s, cudaStream_t st)
    {
        using T = uint64_t;
        if constexpr (D == 4) return launch_fast<Shape<T, false, 2, true, 4, 0, 12, 0>, false, false, void, SFIN>(s, st);
        if constexpr (D == 5) return launch_fast<Shape<T, false, 2, true, 3, 2, 12, 0>, false, false, void, SFIN>(s, st);
        if constexpr (D == 6) return launch_fast<Shape<T, false, 2, true, 3, 3, 12, 0>, false, false, void, SFIN>(s, st);
        if constexpr (D == 7) return launch_fast<Shape<T, false, 2, true, 4, 3, 12, 0>, false, false, void, SFIN, TS>(s, st);
        return launch_fast<Shape<T, false, 2, true, 4, 4, 12, 0>, false, false, void, SFIN, TS>(s, st);
    }
    // mid: where the first of two passes leaves its output (the fused contract passes out; a transposing call its own buffer,
    // i.e. that pass runs in place).  transposed_out: the last pass stores the n1 x n2 matrix (reference contract; out != buf, mid).
    cudaError_t fast_fourstep_rows_t(const uint64_t* buf, uint64_t* mid, uint64_t* out, const uint64_t* n2_table, uint64_t p, int n_power,
                                     int lg1, int lg2, int batch, int in_bound, bool transposed_out, int first_kind, cudaStream_t st,
                                     int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = uint64_t;
        *launched = 0;
        if (!fast_fourstep_rows_t_supported(lg1, lg2) || n_power != lg1 + lg2) return cudaSuccess;
        if (transposed_out && (out == buf || out == mid)) return cudaSuccess;
        if (!(p >= kF60ModulusMin && p < kF60ModulusLimit)) return cudaSuccess;
        if (((long long) batch << lg2) >= (1LL << 31)) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(buf) | reinterpret_cast<uintptr_t>(mid) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        FastArgs<T> a{};
        a.table = n2_table;
        a.p = p;
        a.pbits = 64 - __builtin_clzll((unsigned long long) p);
        {
            const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
            a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
        }
        a.n = n_power;
        a.plus = 0;
        a.batch = batch;
        const int db = lg2 <= 8 ? lg2 : (lg2 - 8 >= 4 ? 8 : 7), da = lg2 - db;
        cudaError_t e = cudaSuccess;
        int kind = first_kind;
        if (da > 0)
        {
            FastArgs<T> s = a;
            s.in = buf;
            s.out = mid;
            s.lo = lg1 + db;
            s.first = 1; // opens the size-n2 transforms: twiddle-1 butterflies without a multiply
            s.last = 0;
            s.in_bound = in_bound;
            s.work = (long long) batch << (s.lo - (12 - da));
            s.rr = s.lo > 10 ? 1 : 0;
            prof_begin(kind++, st);
            switch (da)
            {
                case 4: e = launch_rows_t<4, false>(s, st); break;
                case 5: e = launch_rows_t<5, false>(s, st); break;
                case 6: e = launch_rows_t<6, false>(s, st); break;
                case 7: e = launch_rows_t<7, false>(s, st); break;
                default: e = launch_rows_t<8, false>(s, st); break;
            }
            prof_end(st);
            if (e != cudaSuccess) return e;
        }
        {
            FastArgs<T> s = a;
            s.in = da > 0 ? mid : buf;
            s.out = out;
            s.lo = lg1;
            s.first = da > 0 ? 0 : 1;
            s.last = 1;
            s.in_bound = da > 0 ? 1 : in_bound;
            s.work = ((long long) batch << (lg1 - (12 - db))) << da; // 2^da twiddle ranges
            s.rr = 0;
            prof_begin(kind++, st);
            if (transposed_out)
                e = db == 7 ? launch_rows_t<7, true, true>(s, st) : launch_rows_t<8, true, true>(s, st);
            else
                e = db == 7 ? launch_rows_t<7, true>(s, st) : launch_rows_t<8, true>(s, st);
            prof_end(st);
            if (e != cudaSuccess) return e;
        }
        *launched = kind - first_kind;
        return cudaSuccess;
    }

    // Forward 4-step, reference contract, without a transpose kernel: `in_t` is the n2 x n1 matrix the caller's GPU_Transpose made
    // (n2 rows of n1 contiguous elements, so every size-n1 column transform is a contiguous run).
    //   1. contiguous pass with whole size-n1 transforms inside its tiles, the twiddle-matrix product as epilogue (position-major
    //      kernel, pairs resident in shared memory): in_t -> work, same layout;
    //   2. the size-n2 transforms along the rows of that layout (strided passes, fast_fourstep_rows_t); the last one stores the
    //      n1 x n2 matrix the contract asks for (transposing TMA store): work -> out.
    // Replaces FourStepForwardCoreT1-T4 + FourStepPartialForwardCore1/2 (ntt_4step.cu:68-1020 of the reference).  Single modulus,
    // 64-bit, F60 moduli, batch >= 4; *launched = 0 when not covered (the caller transposes and takes the natural-layout passes).
    cudaError_t fast_fourstep_forward_transposed_in(const uint64_t* in_t, uint64_t* work, uint64_t* out, const uint64_t* n1_table,
                                                    const uint64_t* n2_table, const uint64_t* w_table, void* w_pairs_ws, uint64_t p, int n_power,
                                                    int lg1, int lg2, int batch, cudaStream_t st, int* launched,
                                                    void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = uint64_t;
        *launched = 0;
        if (lg1 < 5 || lg1 > 8 || n_power != lg1 + lg2 || n_power < 12 || batch < 4) return cudaSuccess;
        if (!fast_fourstep_rows_t_supported(lg1, lg2)) return cudaSuccess;
        if (!(p >= kF60ModulusMin && p < kF60ModulusLimit)) return cudaSuccess;
        if (((long long) batch << lg2) >= (1LL << 31)) return cudaSuccess;
        if (in_t == work || work == out || in_t == out) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(in_t) | reinterpret_cast<uintptr_t>(work) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        FastArgs<T> a{};
        a.in = in_t;
        a.out = work;
        a.table = n1_table;
        a.p = p;
        a.pbits = 64 - __builtin_clzll((unsigned long long) p);
        {
            const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
            a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
        }
        a.n = n_power;
        a.n_tw = lg1;
        a.lo = 0;
        a.plus = 0;
        a.first = 1;
        a.last = 0;
        a.in_bound = 1;
        a.batch = batch;
        a.w_pairs = w_pairs_ws;
        a.w_lazy = 1; // products below 2p: the row phase starts from that bound
        const long long N = 1LL << n_power;
        prof_begin(0, st);
        w_pairs_tile_kernel<<<(unsigned) ((N + 255) / 256), 256, 0, st>>>(w_table, reinterpret_cast<Twiddle<T>*>(w_pairs_ws), N, p, a.mu, a.pbits, lg1, lg2);
        prof_end(st);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        prof_begin(1, st);
        e = fourstep_rows_of_transposed_resident_pairs(a, lg1, st);
        prof_end(st);
        if (e == cudaErrorNotSupported) return cudaSuccess; // (the pair table was written for nothing)
        if (e != cudaSuccess) return e;
        int rl = 0;
        e = fast_fourstep_rows_t(work, work, out, n2_table, p, n_power, lg1, lg2, batch, 2, true, 2, st, &rl, prof_begin, prof_end);
        if (e != cudaSuccess) return e;
        if (rl == 0) return cudaErrorUnknown; // (every condition was checked above: the column pass has already run)
        *launched = 2 + rl;
        return cudaSuccess;
    }

    // Inverse 4-step on the tuned kernels.  The arithmetic runs on the n2 x n1 matrix A the reference's intt_first_transpose makes
    // (ntt_4step_cpu.cu:287-299: n2 rows of n1 contiguous elements):
    //   1. a size-n1 Gentleman-Sande transform on every row (n1 table, no n^-1);
    //   2. strided inverse passes over the top lg2 index bits with the n2 table: the first multiplies by the inverse twiddle
    //      matrix as it loads (pairs in A's layout, built from the transposed index), the last applies n^-1 and canonicalises.
    // src_is_y (fused contract: the caller passes y, whose transpose is A): step 1 is a STRIDED pass over y -- row r of A is the
    // run y[i + (jb * n1 + c) * n1], c = 0 .. n1 - 1, i.e. the low lg1 row bits of the n2 x n1 view of y -- with a transposing
    // store, y -> work = A after step 1 (needs n1 >= 64: a tile is 2^(12 - lg1) <= n1 adjacent elements wide); otherwise a
    // contiguous pass, src -> work.
    // transposed_out (reference contract: the n1 x n2 matrix whose transpose is intt(y)): the product pass (the low da stages of the
    // size-n2 transforms, i.e. consecutive rows of A) stores transposed, work -> dst, and the remaining db stages run along the rows
    // of that matrix as the top pass of batch * n1 ordinary size-n2 inverse transforms, in place in dst.  Otherwise work -> dst,
    // then dst in place, A's layout throughout (= NTT_4STEP_CPU::intt order).
    // Single modulus, 64-bit, p below the lazy inverse limit, shapes whose pass splits fit the tile; *launched = 0 when the
    // requested form is not covered (the caller transposes and asks for the plain form).
    cudaError_t fast_fourstep_inverse(const uint64_t* src, uint64_t* work, uint64_t* dst, const uint64_t* n1_table,
                                      const uint64_t* n2_table, const uint64_t* w_table, void* w_pairs_ws, uint64_t p, uint64_t ninv,
                                      int n_power, int lg1, int lg2, int batch, bool src_is_y, bool transposed_out, cudaStream_t st,
                                      int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = uint64_t;
        *launched = 0;
        if (lg1 < 5 || lg1 > 8 || n_power != lg1 + lg2 || n_power < 12) return cudaSuccess;
        if (((long long) batch << lg2) >= (1LL << 31)) return cudaSuccess;
        if (!(p < kFastModulusLimit) || p < 5) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(work) | reinterpret_cast<uintptr_t>(dst)) & 15) return cudaSuccess;
        // split of the lg2 strided stages (executed low bits first)
        int da, db;
        if (lg2 <= 8)
        {
            da = lg2;
            db = 0;
            if (da < 4 || 12 - da > lg1) return cudaSuccess;
        }
        else
        {
            da = lg2 - 8;
            if (da < 12 - lg1) da = 12 - lg1;
            if (da < 4) da = 4;
            db = lg2 - da;
            if (da > 8 || db < 4 || db > 8) return cudaSuccess;
        }
        if (src_is_y && (lg1 < 6 || src == work)) return cudaSuccess;
        if (transposed_out)
        {
            if (da < 5 || work == dst) return cudaSuccess;                     // (transposing stores: passes of 5..8 stages, out of place)
            if (db > 0 && da < 12 - db) return cudaSuccess;                    // the last pass needs 2^(12 - db) adjacent elements of a row
            if (db > 0 && ((((long long) batch << lg1) << db) >= (1LL << 31))) return cudaSuccess;
        }
        FastArgs<T> a{};
        a.p = p;
        a.ninv_w = ninv;
        a.ninv_wq = shoup_companion(ninv, p);
        a.pbits = 64 - __builtin_clzll((unsigned long long) p);
        {
            const unsigned __int128 m = (((unsigned __int128) 1) << (63 + a.pbits)) / (unsigned __int128) p;
            a.mu = (m >> 64) ? ~0ull : (uint64_t) m;
        }
        a.plus = 0;
        const long long N = 1LL << n_power;
        cudaError_t e;
        prof_begin(0, st);
        w_pairs_kernel<<<(unsigned) ((N + 255) / 256), 256, 0, st>>>(w_table, reinterpret_cast<Twiddle<T>*>(w_pairs_ws), N, p, a.mu, a.pbits, lg1, lg2);
        prof_end(st);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        int kind = 1;
        if (src_is_y)
        {
            // row phase straight from y: size-n1 transforms over the low lg1 row bits of the n2 x n1 view, transposing store
            FastArgs<T> s = a;
            s.in = src;
            s.out = work;
            s.table = n1_table;
            s.n = n_power;
            s.lo = lg1;
            s.n_tw = 2 * lg1;
            s.tw_fixed = 1;
            s.first = 1;
            s.last = 0;
            s.batch = batch;
            s.work = ((long long) batch << (lg1 - (12 - lg1))) << (n_power - 2 * lg1);
            prof_begin(kind++, st);
            switch (lg1)
            {
                case 6: e = launch_fast<Shape<T, true, 1, true, 3, 3, 12, 0>, false, false, void, false, true>(s, st); break;
                case 7: e = launch_fast<Shape<T, true, 1, true, 4, 3, 12, 0>, false, false, void, false, true>(s, st); break;
                default: e = launch_fast<Shape<T, true, 1, true, 4, 4, 12, 0>, false, false, void, false, true>(s, st); break;
            }
            prof_end(st);
            if (e == cudaErrorNotSupported) return cudaSuccess;
            if (e != cudaSuccess) return e;
        }
        else
        {
            // row phase: the array as (batch * N / 2048) chunks of 2048 elements, transforms of 2^lg1 inside
            FastArgs<T> s = a;
            s.in = src;
            s.out = work;
            s.table = n1_table;
            s.n = 11;
            s.n_tw = lg1;
            s.lo = 0;
            s.first = 1;
            s.last = 0;
            const long long chunks = ((long long) batch << n_power) >> 11;
            if (chunks > 0x7fffffffLL) return cudaSuccess;
            s.batch = (int) chunks;
            s.work = (chunks + 1) >> 1;
            prof_begin(kind++, st);
            switch (lg1)
            {
                case 5: e = launch_fast<Shape<T, true, 1, false, 1, 4, 12, 1, 5>>(s, st); break;
                case 6: e = launch_fast<Shape<T, true, 1, false, 2, 4, 12, 1, 6>>(s, st); break;
                case 7: e = launch_fast<Shape<T, true, 1, false, 3, 4, 12, 1, 7>>(s, st); break;
                default: e = launch_fast<Shape<T, true, 1, false, 4, 4, 12, 1, 8>>(s, st); break;
            }
            prof_end(st);
            if (e == cudaErrorNotSupported) return cudaSuccess;
            if (e != cudaSuccess) return e;
        }
        // the product pass: d = da stages from row stride 2^lg1
        {
            FastArgs<T> s = a;
            s.in = work;
            s.out = dst;
            s.table = n2_table;
            s.n = n_power;
            s.lo = lg1;
            s.first = 0;
            s.last = db == 0 ? 1 : 0;
            s.batch = batch;
            s.w_pairs = w_pairs_ws;
            s.work = ((long long) batch << (lg1 - (12 - da))) << (n_power - lg1 - da);
            s.rr = 0;
            s.cc_major = 1; // the polynomials of a position follow each other: its pairs are fetched from DRAM once
            prof_begin(kind++, st);
            cudaError_t r = g_resident_pairs.load() ? fourstep_inverse_product_pass_resident_pairs(s, da, transposed_out, st) : cudaErrorNotSupported;
            if (r == cudaErrorNotSupported)
            {
                if (transposed_out)
                    switch (da)
                    {
                        case 5: r = launch_fast<Shape<T, true, 1, true, 3, 2, 12, 0>, true, false, void, false, true>(s, st); break;
                        case 6: r = launch_fast<Shape<T, true, 1, true, 3, 3, 12, 0>, true, false, void, false, true>(s, st); break;
                        case 7: r = launch_fast<Shape<T, true, 1, true, 4, 3, 12, 0>, true, false, void, false, true>(s, st); break;
                        default: r = launch_fast<Shape<T, true, 1, true, 4, 4, 12, 0>, true, false, void, false, true>(s, st); break;
                    }
                else
                    switch (da)
                    {
                        case 4: r = launch_fast<Shape<T, true, 1, true, 4, 0, 12, 0>, true>(s, st); break;
                        case 5: r = launch_fast<Shape<T, true, 1, true, 3, 2, 12, 0>, true>(s, st); break;
                        case 6: r = launch_fast<Shape<T, true, 1, true, 3, 3, 12, 0>, true>(s, st); break;
                        case 7: r = launch_fast<Shape<T, true, 1, true, 4, 3, 12, 0>, true>(s, st); break;
                        default: r = launch_fast<Shape<T, true, 1, true, 4, 4, 12, 0>, true>(s, st); break;
                    }
            }
            prof_end(st);
            if (r != cudaSuccess) return r; // (cudaErrorNotSupported cannot appear here: the row pass already built a tensor map)
        }
        if (db > 0)
        {
            FastArgs<T> s = a;
            s.in = dst;
            s.out = dst;
            s.table = n2_table;
            s.first = 0;
            s.last = 1;
            if (transposed_out)
            {
                // dst holds batch * n1 rows of n2: the top db stages of ordinary size-n2 inverse transforms
                s.n = lg2;
                s.lo = da;
                s.batch = batch << lg1;
                s.work = (long long) s.batch << (da - (12 - db));
                s.rr = da > 10 ? 1 : 0;
            }
            else
            {
                s.n = n_power;
                s.lo = lg1 + da;
                s.batch = batch;
                s.work = (long long) batch << (s.lo - (12 - db));
                s.rr = s.lo > 10 ? 1 : 0;
            }
            prof_begin(kind++, st);
            e = launch_strided<T, true, 1>(db, s, st);
            prof_end(st);
            if (e != cudaSuccess) return e;
        }
        *launched = kind;
        return cudaSuccess;
    }

    // NTTLayout::PerCoefficient on the tuned kernels (replaces ForwardCoreTranspose / InverseCoreTranspose, ntt.cu:1554-2074 of the
    // reference; like it: power-of-two batch, n_power <= 9).  The buffer is one [2^n_power][2^col_log] row-major matrix whose
    // COLUMNS are the transforms, i.e. strided passes over the top n_power index bits of one array of 2^(n_power + col_log)
    // elements -- the same kernels as the 4-step column phase, without a transposition anywhere.  One pass up to 2^8, two for 2^9
    // (5 + 4 stages); the pass that ends a forward transform canonicalises (SFIN), the one that ends an inverse applies n^-1.
    // 64-bit, single modulus, F60 moduli forward / lazy-policy moduli inverse, 2^(12 - stages) <= batch; *launched = 0 otherwise.
    cudaError_t fast_per_coefficient(const uint64_t* in, uint64_t* out, const uint64_t* table, uint64_t p, uint64_t ninv, int n_power,
                                     int col_log, int plus, bool inverse, int signed_io, cudaStream_t st, int* launched,
                                     void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = uint64_t;
        *launched = 0;
        if (n_power < 4 || n_power > 9 || col_log < 1) return cudaSuccess;
        const int n = n_power + col_log;
        if (n > 40 || (1LL << n_power) >= (1LL << 31)) return cudaSuccess;
        if (inverse ? !(p < kFastModulusLimit && p >= 5) : !(p >= kF60ModulusMin && p < kF60ModulusLimit)) return cudaSuccess;
        if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaSuccess;
        // stages of the pass on the high bits / on the low bits (forward order)
        const int da = n_power <= 8 ? n_power : 5, db = n_power - da;
        if (12 - da > col_log + db || (db > 0 && 12 - db > col_log)) return cudaSuccess; // a tile is 2^(12 - stages) adjacent columns wide
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
            s.work = (1LL << (lo - (12 - d))) << (n - lo - d);
            s.rr = (n == lo + d && lo > 10) ? 1 : 0;
            if (inverse) return launch_strided<T, true, 1>(d, s, st);
            if (last)
                switch (d)
                {
                    case 4: return launch_rows_t<4, true>(s, st);
                    case 5: return launch_rows_t<5, true>(s, st);
                    case 6: return launch_rows_t<6, true>(s, st);
                    case 7: return launch_rows_t<7, true>(s, st);
                    default: return launch_rows_t<8, true>(s, st);
                }
            return launch_strided<T, false, 2>(d, s, st);
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
