// gpu_ntt_b200/csrc/merge_ntt.cu -- B200 (sm_100a) Merge-NTT kernels and their host dispatcher.
//
// Replaces the reference's ForwardCore / InverseCore / *LowRing kernels and the GPU_NTT /
// GPU_INTT host functions (src/lib/ntt_merge/ntt.cu:11-3097 in the reference tree) with a
// different design:
//   * one generic "pass" kernel: a tile of 2^k elements is staged in XOR-swizzled shared
//     memory, then processed in register rounds of radix 2^R (R <= 4 for 64-bit, <= 5 for
//     32-bit data): every thread owns 2^R elements of a round and performs R butterfly stages
//     on them out of registers, so shared memory is touched once per R stages instead of once
//     per stage, with one __syncthreads per round;
//   * Shoup/Harvey lazy butterflies on (w, w') twiddle pairs prepared per call by a small
//     pre-kernel (modarith.cuh);
//   * 128-bit global and shared accesses; bank-conflict-free swizzle for every round shape;
//   * any n_power 1..28: 1 pass (whole polynomial(s) per tile) up to 2^13 (u64) / 2^14 (u32),
//     2 passes up to 2^22/2^23, 3 passes above; several small polynomials share one tile.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "gpuntt_b200.h"
#include "merge_ntt.cuh"
#include "modarith.cuh"

namespace gpuntt_b200
{

    // ------------------------------------------------------------------ shared-memory swizzle
    // 16-byte chunks are XOR-permuted inside each 128-byte row by the row index (the TMA
    // SWIZZLE_128B pattern), keeping 16-byte vectors intact:
    //   u64: element bits [1:3] ^= bits [4:6]      u32: element bits [2:4] ^= bits [5:7]
    template <typename T> __device__ __forceinline__ int swz(int l)
    {
        if constexpr (sizeof(T) == 8)
            return l ^ (((l >> 4) & 7) << 1);
        else
            return l ^ (((l >> 5) & 7) << 2);
    }

    template <typename T> struct Vec16;
    template <> struct Vec16<uint64_t>
    {
        using type = ulonglong2;
        static constexpr int N = 2;
    };
    template <> struct Vec16<uint32_t>
    {
        using type = uint4;
        static constexpr int N = 4;
    };

    template <typename T> __device__ __forceinline__ Twiddle<T> ld_tw(const Twiddle<T>* p)
    {
        if constexpr (sizeof(T) == 8)
        {
            ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
            return Twiddle<T>{v.x, v.y};
        }
        else
        {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            return Twiddle<T>{v.x, v.y};
        }
    }

    // ------------------------------------------------------------------ one register round
    // Item `item` of a round with local active bits [lb, lb+R): its 2^R elements are
    // l_base | (a << lb).  Forward = Cooley-Tukey, highest active bit first; inverse =
    // Gentleman-Sande, lowest first.  Twiddle for the stage on local bit lb+ab and element a:
    //   index = (plus << s) + (jrow >> (rb+1)) + (a >> (ab+1)),  rb = lb+ab-c, s = n-1-lo-rb
    // (the bit-reversed caller table makes these 2^(R-1-ab) twiddles adjacent in memory).
    // LB0: the round's lowest active bit is local bit 0, i.e. the thread's 2^R elements are adjacent
    // in memory: they are moved with 16-byte shared-memory accesses (conflict-free under the swizzle).
    template <typename T, int R, bool INV, bool LB0, bool FAST>
    __device__ __forceinline__ void do_round(T* sm, int item, int lb, int c, int stage_hi, int jrow,
                                             int plus, const Twiddle<T>* __restrict__ tw,
                                             const Mod<T, FAST>& M)
    {
        constexpr int E = 1 << R;
        using V = typename Vec16<T>::type;
        constexpr int VN = Vec16<T>::N;
        constexpr bool VEC = LB0 && (E >= VN);
        if constexpr (LB0) lb = 0;
        const int l_base = ((item >> lb) << (lb + R)) | (item & ((1 << lb) - 1));
        T e[E];
        if constexpr (VEC)
        {
#pragma unroll
            for (int a = 0; a < E; a += VN)
            {
                V vv = *reinterpret_cast<const V*>(sm + swz<T>(l_base | a));
                memcpy(&e[a], &vv, sizeof(V));
            }
        }
        else
        {
#pragma unroll
            for (int a = 0; a < E; a++) e[a] = sm[swz<T>(l_base | (a << lb))];
        }

        // stage_hi = s of the stage acting on local bit lb (the LOWEST active bit); the stage on
        // bit lb+ab has s = stage_hi - ab.   rb0 = lb - c.
        const int rb0 = lb - c;
        if constexpr (!INV)
        {
#pragma unroll
            for (int it = 0; it < R; it++)
            {
                const int ab = R - 1 - it;
                const int s = stage_hi - ab;
                const Twiddle<T>* t = tw + ((plus << s) + (jrow >> (rb0 + ab + 1)));
#pragma unroll
                for (int x = 0; x < (E >> (ab + 1)); x++)
                {
                    Twiddle<T> w = ld_tw<T>(t + x);
#pragma unroll
                    for (int y = 0; y < (1 << ab); y++)
                    {
                        const int a0 = (x << (ab + 1)) | y;
                        M.ct(e[a0], e[a0 | (1 << ab)], w);
                    }
                }
            }
        }
        else
        {
#pragma unroll
            for (int ab = 0; ab < R; ab++)
            {
                const int s = stage_hi - ab;
                const Twiddle<T>* t = tw + ((plus << s) + (jrow >> (rb0 + ab + 1)));
#pragma unroll
                for (int x = 0; x < (E >> (ab + 1)); x++)
                {
                    Twiddle<T> w = ld_tw<T>(t + x);
#pragma unroll
                    for (int y = 0; y < (1 << ab); y++)
                    {
                        const int a0 = (x << (ab + 1)) | y;
                        M.gs(e[a0], e[a0 | (1 << ab)], w);
                    }
                }
            }
        }
        if constexpr (VEC)
        {
#pragma unroll
            for (int a = 0; a < E; a += VN)
            {
                V vv;
                memcpy(&vv, &e[a], sizeof(V));
                *reinterpret_cast<V*>(sm + swz<T>(l_base | a)) = vv;
            }
        }
        else
        {
#pragma unroll
            for (int a = 0; a < E; a++) sm[swz<T>(l_base | (a << lb))] = e[a];
        }
    }

    // ------------------------------------------------------------------ the pass kernel
    template <typename T, bool INV, bool RNS, bool FAST>
    __global__ void __launch_bounds__(kThreads) merge_pass_kernel(const PassArgs<T> a)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T* sm = reinterpret_cast<T*>(smem_raw);
        using ST = typename std::make_signed<T>::type;
        using V = typename Vec16<T>::type;
        constexpr int VN = Vec16<T>::N;

        const PassPlan& pl = a.plan;
        const int k = pl.tile_log, n = a.n_power, c = pl.c, lo = pl.lo;
        const int tile_elems = 1 << k;
        const int tid = threadIdx.x;

        // ---- which tile is this?
        const long long tile = blockIdx.x;
        long long obase;   // offset inside polynomial poly0 of local index 0
        int jrow_tile;     // (index within polynomial >> lo) of local row 0
        long long poly0;   // polynomial (transform number) of local index 0
        if (lo > 0)
        {
            const int hi = lo + pl.d;
            const int cc_bits = lo - c, pre_bits = n - hi;
            const long long cc = tile & ((1LL << cc_bits) - 1);
            const long long P = (tile >> cc_bits) & ((1LL << pre_bits) - 1);
            poly0 = tile >> (cc_bits + pre_bits);
            obase = (P << hi) + (cc << c);
            jrow_tile = (int) (P << pl.d);
        }
        else
        {
            const long long g0 = tile << k;
            obase = g0 & ((1LL << n) - 1);
            jrow_tile = (int) obase;
            poly0 = g0 >> n;
        }
        const int cmask = (1 << c) - 1;
        const int poly_shift = n - lo; // (l >> c) >> poly_shift = polynomial offset inside the tile
        const long long nmask = (1LL << n) - 1;
        // transform number b -> modulus index / polynomial slot (the *_Ordered entry points of the reference)
        // (transform number, offset in polynomial) -> twiddle slice / modulus index
        auto slice_index = [&](long long b, long long off) -> int
        {
            if (a.col_log > 0) return (int) ((off & ((1LL << a.col_log) - 1)) % a.mod_count); // PerCoefficient: by column
            return (int) ((b >> a.mod_shift) % a.mod_count);
        };
        auto mod_index = [&](long long b, long long off) -> int
        {
            const int mi = slice_index(b, off);
            return a.mod_order ? a.mod_order[mi] : mi;
        };
        // global element index of local element (row, col); false when the transform does not exist (ragged last tile)
        auto locate = [&](int row, int col, long long& g, long long& off, long long& b) -> bool
        {
            const long long o = obase + ((long long) row << lo) + col;
            b = poly0 + (o >> n);
            off = o & nmask;
            if (b >= a.batch) return false;
            const long long slot = a.poly_order ? (long long) a.poly_order[b] : b;
            g = (slot << n) + off;
            return true;
        };
        auto w_index = [&](long long off) -> long long
        {
            return a.w_mode == 2 ? (((off & ((1LL << a.w_lo) - 1)) << a.w_hi) | (off >> a.w_lo)) : off;
        };

        // ---- global -> shared
        // per-element fix-ups on the way in: signed input (first forward pass), 4-step inverse twiddle matrix
        const bool fix_signed = (!INV) && pl.first && a.signed_io;
        const bool w_in = (a.w_mode == 2) && pl.first;
        auto fix_in = [&](T x, long long b, long long off) -> T
        {
            if (!fix_signed && !w_in) return x;
            T p = a.p, bit = a.bar_bit, mu = a.bar_mu;
            const T* w = a.w_table;
            if constexpr (RNS)
            {
                const int mi = mod_index(b, off);
                p = a.mod_values[3 * mi];
                bit = a.mod_values[3 * mi + 1];
                mu = a.mod_values[3 * mi + 2];
                if (w_in && !a.shared_tables) w += ((size_t) mi << n);
            }
            if (fix_signed && (ST) x < 0) x += p; // p - |x|  (modular_arith.cuh:372-385 of the reference)
            if (w_in) x = barrett_mul(x, w[w_index(off)], p, bit, mu);
            return x;
        };
        {
            const T* gin = reinterpret_cast<const T*>(a.in);
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(gin) & 15) == 0) && (c == 0 || c >= 2) && (k >= 2) &&
                                ((1 << n) >= VN);
            for (int l = tid * VN; l < tile_elems; l += kThreads * VN)
            {
                const int row = l >> c, col = l & cmask;
                long long g, off, b;
                T v[VN];
                if (vec_ok && locate(row, col, g, off, b))
                {
                    V vv = *reinterpret_cast<const V*>(gin + g);
                    memcpy(v, &vv, sizeof(V));
#pragma unroll
                    for (int i = 0; i < VN; i++) v[i] = fix_in(v[i], b, off + i);
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < VN; i++)
                        v[i] = locate((l + i) >> c, (l + i) & cmask, g, off, b) ? fix_in(gin[g], b, off) : T(0);
                }
                V vv;
                memcpy(&vv, v, sizeof(V));
                *reinterpret_cast<V*>(sm + swz<T>(l)) = vv;
            }
        }

        // ---- register rounds
        int lb_fwd[kMaxRounds]; // local low bit of each round
        {
            int acc = c + pl.d;
#pragma unroll
            for (int r = 0; r < kMaxRounds; r++)
            {
                if (r < pl.nrounds) acc -= pl.round_bits[r];
                lb_fwd[r] = acc;
            }
        }
#pragma unroll 1
        for (int ri = 0; ri < pl.nrounds; ri++)
        {
            const int r = INV ? (pl.nrounds - 1 - ri) : ri;
            int R = 0, lb = 0;
#pragma unroll
            for (int q = 0; q < kMaxRounds; q++)
                if (q == r)
                {
                    R = pl.round_bits[q];
                    lb = lb_fwd[q];
                }
            __syncthreads();
            const int items = tile_elems >> R;
            const int stage_hi = n - 1 - lo - (lb - c);
#pragma unroll 1
            for (int item = tid; item < items; item += kThreads)
            {
                const int l_base = ((item >> lb) << (lb + R)) | (item & ((1 << lb) - 1));
                const int row = l_base >> c;
                const int jrow = (jrow_tile | row) & ((1 << poly_shift) - 1);
                T p = a.p;
                const Twiddle<T>* tw = reinterpret_cast<const Twiddle<T>*>(a.tw);
                if constexpr (RNS)
                {
                    const long long b = poly0 + (row >> poly_shift);
                    const long long off = obase + (l_base & cmask); // only its column bits matter here
                    p = a.mod_values[3 * mod_index(b, off)];
                    tw += ((size_t) slice_index(b, off) << a.tw_stride_log);
                }
                const Mod<T, FAST> M(p);
                if (lb == 0)
                    switch (R)
                    {
                        case 1: do_round<T, 1, INV, true, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        case 2: do_round<T, 2, INV, true, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        case 3: do_round<T, 3, INV, true, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        case 4: do_round<T, 4, INV, true, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        default:
                            if constexpr (sizeof(T) == 4)
                                do_round<T, 5, INV, true, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M);
                            break;
                    }
                else
                    switch (R)
                    {
                        case 1: do_round<T, 1, INV, false, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        case 2: do_round<T, 2, INV, false, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        case 3: do_round<T, 3, INV, false, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                        default: do_round<T, 4, INV, false, FAST>(sm, item, lb, c, stage_hi, jrow, a.plus, tw, M); break;
                    }
            }
        }
        __syncthreads();

        // ---- shared -> global
        {
            T* gout = a.out;
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(gout) & 15) == 0) && (c == 0 || c >= 2) && (k >= 2) &&
                                ((1 << n) >= VN);
            const bool centre = INV && pl.last && a.signed_io;
            // last pass: canonical form, n^-1 and centred output (inverse), 4-step twiddle matrix (forward)
            auto fix_out = [&](T x, long long b, long long off) -> T
            {
                if (!pl.last) return x;
                T p = a.p, bit = a.bar_bit, mu = a.bar_mu;
                Twiddle<T> ni{a.ninv_w, a.ninv_wq};
                const T* w = a.w_table;
                if constexpr (RNS)
                {
                    const int mi = mod_index(b, off);
                    p = a.mod_values[3 * mi];
                    bit = a.mod_values[3 * mi + 1];
                    mu = a.mod_values[3 * mi + 2];
                    if constexpr (INV) ni = reinterpret_cast<const Twiddle<T>*>(a.ninv_tw)[slice_index(b, off)];
                    if (a.w_mode == 1 && !a.shared_tables) w += ((size_t) mi << n);
                }
                const Mod<T, FAST> M(p);
                if constexpr (INV)
                {
                    x = M.canon_inv(x, ni);
                    if (centre && x > (p >> 1)) x -= p; // modular_arith.cuh:389-405 of the reference
                }
                else
                {
                    x = M.canon_fwd(x);
                    if (a.w_mode == 1) x = barrett_mul(x, w[off], p, bit, mu);
                }
                return x;
            };
            for (int l = tid * VN; l < tile_elems; l += kThreads * VN)
            {
                const int row = l >> c, col = l & cmask;
                long long g, off, b;
                V vv = *reinterpret_cast<const V*>(sm + swz<T>(l));
                T v[VN];
                memcpy(v, &vv, sizeof(V));
                if (vec_ok)
                {
                    if (!locate(row, col, g, off, b)) continue;
#pragma unroll
                    for (int i = 0; i < VN; i++) v[i] = fix_out(v[i], b, off + i);
                    memcpy(&vv, v, sizeof(V));
                    *reinterpret_cast<V*>(gout + g) = vv;
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < VN; i++)
                        if (locate((l + i) >> c, (l + i) & cmask, g, off, b)) gout[g] = fix_out(v[i], b, off);
                }
            }
        }
    }

    // ------------------------------------------------------------------ twiddle companion pre-kernel
    // out[(m << stride_log) + i] = { w, floor(w * 2^BITS / p_m) }  for the caller's table slice m;
    // after the slices: the same pair for n^-1 of every modulus (RNS inverse only).
    template <typename T>
    __global__ void twiddle_prep_kernel(const T* __restrict__ table, Twiddle<T>* __restrict__ out,
                                        const T* __restrict__ mod_values, T p_single, int slices, int in_stride_log,
                                        int out_stride_log, int shared_tables, long long table_len,
                                        const T* __restrict__ ninv_dev, Twiddle<T>* __restrict__ ninv_out,
                                        const int* __restrict__ mod_order, int unit_ninv)
    {
        // output slice m serves the transforms with b % mod_count == m; with mod_order their modulus, table
        // slice and n^-1 are entry mod_order[m] of the caller's arrays (ntt.cu:3117-3118 of the reference)
        const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
        const int m = blockIdx.y;
        const int mi = mod_order ? mod_order[m] : m;
        const T p = mod_values ? mod_values[3 * mi] : p_single;
        if (i < table_len)
        {
            const size_t src = shared_tables ? (size_t) i : (((size_t) mi << in_stride_log) + (size_t) i);
            const T w = table[src];
            out[((size_t) m << out_stride_log) + (size_t) i] = Twiddle<T>{w, shoup_companion(w, p)};
        }
        if ((ninv_dev || unit_ninv) && i == 0)
        {
            const T w = unit_ninv ? T(1) : ninv_dev[mi];
            ninv_out[m] = Twiddle<T>{w, shoup_companion(w, p)};
        }
    }

    // ------------------------------------------------------------------ launch plan
    static void split_rounds(PassPlan& ps, int element_bits)
    {
        // rounds listed from the highest bits down.  The lowest round of a contiguous 32-bit
        // tile is radix-32 so that no round starts at local bit 4 (bank conflicts).
        int d = ps.d, nr = 0;
        int rounds_low_first[kMaxRounds];
        if (element_bits == 32 && ps.lo == 0 && d >= 5)
        {
            rounds_low_first[nr++] = 5;
            d -= 5;
        }
        while (d > 0)
        {
            int r = d >= 4 ? 4 : d;
            rounds_low_first[nr++] = r;
            d -= r;
        }
        ps.nrounds = nr;
        for (int i = 0; i < nr; i++) ps.round_bits[i] = rounds_low_first[nr - 1 - i];
        for (int i = nr; i < kMaxRounds; i++) ps.round_bits[i] = 0;
    }

    PassPlan make_strided_pass(int lo, int d, int element_bits)
    {
        const int kmax = (element_bits == 64) ? 13 : 14;
        const int cmin = (element_bits == 64) ? 4 : 5;
        PassPlan ps{};
        ps.lo = lo;
        ps.d = d;
        int c = 12 - d;
        if (c < cmin) c = cmin;
        if (c > lo) c = lo;
        if (d + c > kmax) c = kmax - d;
        ps.c = c;
        ps.tile_log = d + c;
        split_rounds(ps, element_bits);
        return ps;
    }

    MergePlan make_merge_plan(int n, int element_bits)
    {
        MergePlan mp{};
        const int kmax = (element_bits == 64) ? 13 : 14;   // 64 KiB tiles at most
        const int cmin = (element_bits == 64) ? 4 : 5;     // 128-byte row segments in strided tiles
        const int kpref = 12;                              // preferred tile: 4096 elements
        auto contiguous = [&](int d)
        {
            PassPlan ps{};
            ps.lo = 0;
            ps.c = 0;
            ps.d = d;
            ps.tile_log = d > 11 ? d : 11; // small rings: several polynomials per 2048-element tile
            if (d <= kpref && d > 0 && n > d) ps.tile_log = kpref;
            split_rounds(ps, element_bits);
            return ps;
        };
        auto strided = [&](int lo, int d)
        {
            PassPlan ps{};
            ps.lo = lo;
            ps.d = d;
            int c = kpref - d;
            if (c < cmin) c = cmin;
            if (c > lo) c = lo;
            ps.c = c;
            ps.tile_log = d + c;
            split_rounds(ps, element_bits);
            return ps;
        };
        if (n <= kmax)
        {
            mp.npasses = 1;
            mp.pass[0] = contiguous(n);
        }
        else if (n <= 2 * kmax - cmin)
        {
            int d1 = n / 2;
            if (d1 > kmax - cmin) d1 = kmax - cmin; // strided tile = 2^(d1 + c) elements <= 64 KiB
            int d2 = n - d1;
            mp.npasses = 2;
            mp.pass[0] = strided(d2, d1);
            mp.pass[1] = contiguous(d2);
        }
        else
        {
            int d1 = n / 3, d2 = n / 3, d3 = n - d1 - d2;
            mp.npasses = 3;
            mp.pass[0] = strided(d2 + d3, d1);
            mp.pass[1] = strided(d3, d2);
            mp.pass[2] = contiguous(d3);
        }
        for (int i = 0; i < mp.npasses; i++)
        {
            mp.pass[i].first = 0;
            mp.pass[i].last = 0;
        }
        return mp;
    }

    // ------------------------------------------------------------------ host side
    thread_local std::string g_last_error;
    thread_local int g_last_launches = 0;
    static std::atomic<unsigned long long> g_total_launches{0};

    // Optional per-launch device timing (bench.py's live roofline): when enabled, every launch is
    // bracketed by CUDA events on the launching stream; gpuntt_b200_profile_read() resolves them.
    struct ProfRec
    {
        cudaEvent_t e0, e1;
        int dev;
        int kind; // 0 = twiddle_prep_kernel, 1.. = merge pass number (1-based, execution order)
    };
    static std::atomic<int> g_profiling{0};
    static std::atomic<int> g_force_generic{0};
    static std::mutex g_prof_mutex;
    static std::vector<ProfRec> g_prof;
    static std::map<int, std::vector<cudaEvent_t>> g_prof_pool; // per device; recycled: creating an event costs more host time than recording it
    static cudaEvent_t prof_event()
    {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(g_prof_mutex);
            auto& pool = g_prof_pool[dev];
            if (!pool.empty())
            {
                cudaEvent_t e = pool.back();
                pool.pop_back();
                return e;
            }
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }

    struct ProfScope
    {
        bool on;
        ProfRec r;
        cudaStream_t st;
        ProfScope(int kind, cudaStream_t s) : on(g_profiling.load() != 0), st(s)
        {
            if (!on) return;
            r.kind = kind;
            cudaGetDevice(&r.dev);
            r.e0 = prof_event();
            r.e1 = prof_event();
            cudaEventRecord(r.e0, st);
        }
        ~ProfScope()
        {
            if (!on) return;
            cudaEventRecord(r.e1, st);
            std::lock_guard<std::mutex> lk(g_prof_mutex);
            g_prof.push_back(r);
        }
    };

    static thread_local ProfScope* g_cur_prof = nullptr;
    static void prof_begin(int kind, cudaStream_t st)
    {
        g_cur_prof = new ProfScope(kind, st);
        g_last_launches++;
        g_total_launches++;
    }
    static void prof_end(cudaStream_t)
    {
        delete g_cur_prof;
        g_cur_prof = nullptr;
    }

    // merge_fast.cu
    int fast_describe(int n_power, int element_bits, char* buf, size_t len);
    template <typename T>
    cudaError_t fast_merge_rns(const T* in, T* out, const T* table, const T* mod_dev, const T* ninv_dev, const int* mod_order,
                               const int* poly_order, int mod_count, int n_power, int plus, bool inverse, int batch, int* flag_ws,
                               cudaStream_t st, int* launched,
                               void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    cudaError_t fast_fourstep_inverse(const uint64_t* src, uint64_t* work, uint64_t* dst, const uint64_t* n1_table,
                                      const uint64_t* n2_table, const uint64_t* w_table, void* w_pairs_ws, uint64_t p, uint64_t ninv,
                                      int n_power, int lg1, int lg2, int batch, bool src_is_y, bool transposed_out, cudaStream_t st,
                                      int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    cudaError_t fast_fourstep_forward_transposed_in(const uint64_t* in_t, uint64_t* work, uint64_t* out, const uint64_t* n1_table,
                                                    const uint64_t* n2_table, const uint64_t* w_table, void* w_pairs_ws, uint64_t p, int n_power,
                                                    int lg1, int lg2, int batch, cudaStream_t st, int* launched,
                                                    void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    cudaError_t fast_fourstep_columns(const uint64_t* in, uint64_t* out, const uint64_t* n1_table, const uint64_t* w_table,
                                      void* w_pairs_ws, uint64_t p, int n_power, int lg1, int lg2, int batch, cudaStream_t st,
                                      int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t), int w_lazy,
                                      int transposed = 0);
    bool fast_fourstep_rows_t_supported(int lg1, int lg2);
    cudaError_t fast_fourstep_rows_t(const uint64_t* buf, uint64_t* mid, uint64_t* out, const uint64_t* n2_table, uint64_t p, int n_power,
                                     int lg1, int lg2, int batch, int in_bound, bool transposed_out, int first_kind, cudaStream_t st,
                                     int* launched, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    template <typename T>
    cudaError_t fast_merge(const T* in, T* out, const T* table, T p, T ninv, int n_power, int plus, bool inverse,
                           int batch, cudaStream_t st, int* launched, void (*prof_begin)(int, cudaStream_t),
                           void (*prof_end)(cudaStream_t), int in_bound = 1, unsigned* counters = nullptr, int signed_io = 0);
    cudaError_t fast_per_coefficient(const uint64_t* in, uint64_t* out, const uint64_t* table, uint64_t p, uint64_t ninv, int n_power,
                                     int col_log, int plus, bool inverse, int signed_io, cudaStream_t st, int* launched,
                                     void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    cudaError_t fast_per_coefficient32(const uint32_t* in, uint32_t* out, const uint32_t* table, uint32_t p, uint32_t ninv, int n_power,
                                       int col_log, int plus, bool inverse, int signed_io, cudaStream_t st, int* launched,
                                       void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t)); // merge_fast_pc32.cu
    void fused_set_lag_steps(int v); // merge_fused.cu
    void fused_set_policy(int v);
    void fused_set_small_tile_elems(long long v);
    void fourstep_set_resident_pairs(int on); // merge_fast_4step.cu
    void fast_set_one_tile_mode(int v);       // merge_fast.cu
    void fast_set_single_poly_tiles(int v);
    bool fast_supported(int n_power, int element_bits);
    bool fast_small_supported(int n_power, int element_bits); // merge_fast.cu

    static int fail(int code, const std::string& msg)
    {
        g_last_error = msg;
        return code;
    }
    static int cuda_fail(cudaError_t e, const char* what)
    {
        return fail(GPUNTT_B200_ERR_CUDA,
                    std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
    }

    struct Workspace
    {
        void* ptr = nullptr;
        size_t bytes = 0;
    };
    static std::mutex g_ws_mutex;
    // (device, stream, slot).  Scratch is reused by the calls of one stream, which the stream itself orders.  The handle
    // cudaStreamPerThread names a DIFFERENT stream in every host thread, so for it the key carries the calling thread.
    using WsKey = std::tuple<int, void*, int, size_t>;
    static std::map<WsKey, Workspace> g_ws;
    static WsKey ws_key(int dev, void* stream, int slot)
    {
        size_t th = 0;
        if ((cudaStream_t) stream == cudaStreamPerThread) th = std::hash<std::thread::id>()(std::this_thread::get_id()) | 1;
        return WsKey(dev, stream, slot, th);
    }

    // A scratch buffer is written by one kernel of a call and read by later ones (twiddle companions by the generic passes,
    // the 4-step scratch and pair table by its passes).  The stream orders kernels in ENQUEUE order, so two host threads that
    // issue calls on the same stream handle (the legacy default stream, say) must not interleave their enqueues: every call
    // that uses such scratch holds this per-(device, stream) lock while it enqueues (microseconds; calls never wait for the
    // device under it unless a buffer has to grow).  Recursive: the 4-step call re-enters the merge call for its row phase.
    static std::map<WsKey, std::unique_ptr<std::recursive_mutex>> g_stream_locks; // under g_ws_mutex
    struct StreamEnqueueLock
    {
        std::recursive_mutex* m = nullptr;
        explicit StreamEnqueueLock(void* stream)
        {
            int dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess) return;
            {
                std::lock_guard<std::mutex> lk(g_ws_mutex);
                auto& slot = g_stream_locks[ws_key(dev, stream, -1)];
                if (!slot) slot.reset(new std::recursive_mutex);
                m = slot.get();
            }
            m->lock();
        }
        ~StreamEnqueueLock()
        {
            if (m) m->unlock();
        }
        StreamEnqueueLock(const StreamEnqueueLock&) = delete;
        StreamEnqueueLock& operator=(const StreamEnqueueLock&) = delete;
    };

    // slot 0: twiddle companions; slots 1,2: host-convenience staging buffers
    static cudaError_t get_workspace(void* stream, int slot, size_t bytes, void** out)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        Workspace& w = g_ws[ws_key(dev, stream, slot)];
        if (w.bytes < bytes)
        {
            // growing means cudaMalloc (and a stream synchronisation): neither is legal while the stream is being captured, and
            // attempting them would invalidate the capture -- decline cleanly (run the call once outside the capture first)
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing((cudaStream_t) stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
                return cudaErrorStreamCaptureUnsupported;
            if (w.ptr)
            {
                // in-flight kernels of earlier calls may still read the old buffer
                e = cudaStreamSynchronize((cudaStream_t) stream);
                if (e != cudaSuccess) return e;
                cudaFree(w.ptr);
                w.ptr = nullptr;
                w.bytes = 0;
            }
            size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
            e = cudaMalloc(&w.ptr, want);
            if (e != cudaSuccess) return e;
            w.bytes = want;
        }
        *out = w.ptr;
        return cudaSuccess;
    }

    // Per-polynomial progress counters of the single-launch two-pass kernels (merge_fused.cu): all-zero between calls
    // (the kernels clean up after themselves), so they are zeroed once, when the buffer is allocated -- on the caller's
    // stream, i.e. before the first kernel that uses them.  nullptr: run the passes as separate launches.
    static std::atomic<int> g_fused_enabled{1};
    static std::atomic<int> g_fourstep_transposed{1};
    static std::atomic<int> g_fourstep_modcache{1};
    static std::map<std::tuple<int, const void*, const void*>, std::pair<uint64_t, uint64_t>> g_modcache; // under g_ws_mutex
    static unsigned* fused_counters(void* stream, long long polys)
    {
        if (!g_fused_enabled.load() || polys <= 0) return nullptr;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
        const size_t bytes = ((size_t) polys + 1) * sizeof(unsigned); // + the ticket word
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        Workspace& w = g_ws[ws_key(dev, stream, 8)];
        if (w.bytes < bytes)
        {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing((cudaStream_t) stream, &cap);
            if (cap != cudaStreamCaptureStatusNone) return nullptr; // no allocation inside a capture: separate launches
            if (w.ptr)
            {
                if (cudaStreamSynchronize((cudaStream_t) stream) != cudaSuccess) return nullptr;
                cudaFree(w.ptr);
                w.ptr = nullptr;
                w.bytes = 0;
            }
            size_t want = bytes < (1u << 16) ? (1u << 16) : bytes * 2;
            if (cudaMalloc(&w.ptr, want) != cudaSuccess)
            {
                cudaGetLastError();
                w.ptr = nullptr;
                return nullptr;
            }
            w.bytes = want;
            if (cudaMemsetAsync(w.ptr, 0, want, (cudaStream_t) stream) != cudaSuccess) return nullptr;
        }
        return reinterpret_cast<unsigned*>(w.ptr);
    }

    template <typename K> static cudaError_t allow_smem(K kernel, size_t bytes)
    {
        if (bytes <= 48 * 1024) return cudaSuccess;
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
    }

    template <typename T, bool INV, bool RNS, bool FAST>
    static cudaError_t launch_pass(const PassArgs<T>& args, int batch, cudaStream_t st, int pass_no)
    {
        ProfScope prof(pass_no, st);
        const PassPlan& pl = args.plan;
        const long long total = (long long) batch << args.n_power;
        long long tiles;
        if (pl.lo > 0)
            tiles = total >> pl.tile_log;
        else
            tiles = (total + (1LL << pl.tile_log) - 1) >> pl.tile_log;
        const size_t smem = sizeof(T) << pl.tile_log;
        auto kern = merge_pass_kernel<T, INV, RNS, FAST>;
        cudaError_t e = allow_smem(kern, smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned) tiles, kThreads, smem, st>>>(args);
        g_last_launches++;
        g_total_launches++;
        return cudaGetLastError();
    }

    // ------------------------------------------------------------------ pass sequencer
    // One "core call": a list of passes (already in execution order) over one [batch][2^n_power] array with one
    // root table.  merge_execute_t builds it from make_merge_plan; the 4-step driver builds its own lists.
    template <typename T> struct CoreCall
    {
        const void* in = nullptr;
        T* out = nullptr;
        const T* table = nullptr; // caller's bit-reversed root table (plain residues)
        long long table_len = 0;  // entries per modulus slice
        int stride_log = 0;       // log2 of the slice stride in `table` (RNS); ignored when shared_tables
        int shared_tables = 0;
        int n_power = 0, batch = 0, mod_count = 0, plus = 0, signed_io = 0;
        bool inverse = false;
        T p = 0, ninv = 0;                 // single modulus
        const T* mod_values = nullptr;     // RNS: device Modulus<T>[mod_count]
        const T* ninv_dev = nullptr;       // RNS inverse: device n^-1 per modulus
        const int* mod_order = nullptr;
        const int* poly_order = nullptr;
        int mod_shift = 0;
        int col_log = 0;                   // PerCoefficient: columns of one [2^n][2^col_log] matrix
        bool unit_ninv = false;            // inverse without the final n^-1 (4-step column phase)
        const T* w_table = nullptr;
        int w_mode = 0, w_lo = 0, w_hi = 0;
        cudaStream_t st = nullptr;
        int ws_slot = 0;                   // workspace slot for the (w, w') table
        int npasses = 0;
        PassPlan pass[4];
        bool mark_first = true, mark_last = true; // pass[0] reads caller input / pass[npasses-1] writes final results
    };

    template <typename T> static void barrett_constants(T p, T& bit, T& mu)
    {
        bit = 0;
        for (T v = p; v; v >>= 1) bit++;
        if constexpr (sizeof(T) == 8)
            mu = p ? (T) ((((unsigned __int128) 1) << (2 * bit + 1)) / p) : 0;
        else
            mu = p ? (T) ((((uint64_t) 1) << (2 * bit + 1)) / p) : 0;
    }

    template <typename T> static int run_core(const CoreCall<T>& cc)
    {
        StreamEnqueueLock enqueue_lock((void*) cc.st); // twiddle_prep_kernel -> passes share the companion buffer
        const bool rns = cc.mod_count > 0;
        const int slices = rns ? cc.mod_count : 1;
        const int tw_stride_log = cc.shared_tables ? 0 : cc.stride_log;
        // companion slices are always laid out (1 << lg) apart, lg = ceil(log2(table_len))
        int lg = 0;
        while ((1LL << lg) < cc.table_len) lg++;
        const int out_stride_log = cc.shared_tables ? lg : cc.stride_log;
        const size_t tw_elems = ((size_t) (slices - 1) << out_stride_log) + (size_t) cc.table_len;
        const size_t bytes = (tw_elems + (size_t) slices) * sizeof(Twiddle<T>);
        void* ws = nullptr;
        cudaError_t e = get_workspace((void*) cc.st, cc.ws_slot, bytes, &ws);
        if (e != cudaSuccess) return cuda_fail(e, "workspace allocation");
        Twiddle<T>* tw = reinterpret_cast<Twiddle<T>*>(ws);
        Twiddle<T>* ninv_tw = tw + tw_elems;
        {
            ProfScope prof(0, cc.st);
            const int threads = 256;
            dim3 grid((unsigned) ((cc.table_len + threads - 1) / threads), (unsigned) slices);
            twiddle_prep_kernel<T><<<grid, threads, 0, cc.st>>>(cc.table, tw, cc.mod_values, cc.p, slices, tw_stride_log, out_stride_log,
                                                                cc.shared_tables, cc.table_len,
                                                                (rns && cc.inverse && !cc.unit_ninv) ? cc.ninv_dev : nullptr, ninv_tw,
                                                                cc.mod_order, (rns && cc.inverse && cc.unit_ninv) ? 1 : 0);
            g_last_launches++;
            g_total_launches++;
            e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "twiddle_prep_kernel launch");
        }
        PassArgs<T> args{};
        args.tw = tw;
        args.mod_values = cc.mod_values;
        args.ninv_tw = ninv_tw;
        args.p = cc.p;
        if (cc.inverse && !rns)
        {
            args.ninv_w = cc.unit_ninv ? T(1) : cc.ninv;
            args.ninv_wq = cc.p ? shoup_companion(args.ninv_w, cc.p) : 0;
        }
        args.n_power = cc.n_power;
        args.mod_count = cc.mod_count;
        args.tw_stride_log = out_stride_log;
        args.plus = cc.plus;
        args.signed_io = cc.signed_io;
        args.total_elems = (long long) cc.batch << cc.n_power;
        args.mod_order = cc.mod_order;
        args.poly_order = cc.poly_order;
        args.batch = cc.batch;
        args.mod_shift = cc.mod_shift;
        args.col_log = cc.col_log;
        args.shared_tables = cc.shared_tables;
        args.w_table = cc.w_table;
        args.w_mode = cc.w_mode;
        args.w_lo = cc.w_lo;
        args.w_hi = cc.w_hi;
        if (!rns) barrett_constants<T>(cc.p, args.bar_bit, args.bar_mu);
        bool fast = false;
        if constexpr (sizeof(T) == 8) fast = !rns && (uint64_t) cc.p < kFastModulusLimit && (cc.inverse || (uint64_t) cc.p >= kFastModulusMin);
        for (int i = 0; i < cc.npasses; i++)
        {
            args.plan = cc.pass[i];
            args.plan.first = (i == 0) && cc.mark_first;
            args.plan.last = (i == cc.npasses - 1) && cc.mark_last;
            args.in = (i == 0) ? cc.in : cc.out; // later passes chain through `out` (also out of place)
            args.out = cc.out;
            if constexpr (sizeof(T) == 8)
            {
                if (fast)
                    e = cc.inverse ? launch_pass<T, true, false, true>(args, cc.batch, cc.st, i + 1)
                                   : launch_pass<T, false, false, true>(args, cc.batch, cc.st, i + 1);
            }
            if (!fast)
            {
                if (cc.inverse)
                    e = rns ? launch_pass<T, true, true, false>(args, cc.batch, cc.st, i + 1)
                            : launch_pass<T, true, false, false>(args, cc.batch, cc.st, i + 1);
                else
                    e = rns ? launch_pass<T, false, true, false>(args, cc.batch, cc.st, i + 1)
                            : launch_pass<T, false, false, false>(args, cc.batch, cc.st, i + 1);
            }
            if (e != cudaSuccess) return cuda_fail(e, "merge_pass_kernel launch");
        }
        return GPUNTT_B200_OK;
    }

    constexpr long long kRaggedSplitMinElems = 1LL << 22;
    template <typename T> static int merge_execute_t(const gpuntt_b200_merge_desc* d)
    {
        const int n = d->n_power;
        const bool inv = d->direction == GPUNTT_B200_INVERSE;
        const bool rns = d->mod_count > 0;
        const bool plus = d->reduction_poly == GPUNTT_B200_X_N_PLUS;
        cudaStream_t st = (cudaStream_t) d->stream;
        if (!rns && !g_force_generic.load() && d->ntt_layout == GPUNTT_B200_PER_POLYNOMIAL && fast_small_supported(n, (int) sizeof(T) * 8))
        {
            // Small rings walk the array as whole 2048- / 4096-element chunks (fast_small).  A batch that ends inside a chunk is
            // split: the polynomials of the whole chunks take the one-launch tuned kernel, the (fewer than 16) polynomials of the
            // ragged tail the generic kernel -- instead of the whole batch falling back for one odd polynomial.  Only above 2^22
            // elements: the tail costs two more launches, and a launch-bound call is quicker on the generic kernel alone
            // (profiles/r2_ragged_small_ab.jsonl: 2.1-4.4x for 2^26-element batches, 0.7-0.9x for 2^12 .. 2^20 elements).
            const int chunk_log = (sizeof(T) == 8 ? 11 : 12) - n; // log2 polynomials per chunk (one-tile rings: a polynomial is a chunk)
            const int tail = chunk_log > 0 ? (d->batch_size & ((1 << chunk_log) - 1)) : 0;
            if (tail > 0 && ((long long) (d->batch_size - tail) << n) >= kRaggedSplitMinElems)
            {
                gpuntt_b200_merge_desc part = *d;
                part.batch_size = d->batch_size - tail;
                int rc = merge_execute_t<T>(&part);
                if (rc != GPUNTT_B200_OK) return rc;
                const size_t skip = ((size_t) part.batch_size << n) * sizeof(T);
                part.in = reinterpret_cast<const char*>(d->in) + skip;
                part.out = reinterpret_cast<char*>(d->out) + skip;
                part.batch_size = tail;
                return merge_execute_t<T>(&part);
            }
        }
        if (!rns && !g_force_generic.load() && d->ntt_layout == GPUNTT_B200_PER_POLYNOMIAL)
        {
            // (signed data: same kernels -- the first forward round fixes negative inputs up as it loads, the last inverse
            // round centres its outputs, ntt.cu:481-489, 1178-1186 of the reference)
            int launched = 0;
            cudaError_t fe = fast_merge<T>(reinterpret_cast<const T*>(d->in), reinterpret_cast<T*>(d->out),
                                           reinterpret_cast<const T*>(d->root_of_unity_table), (T) d->modulus_value,
                                           (T) d->mod_inverse_value, n, plus ? 1 : 0, inv, d->batch_size, st, &launched,
                                           prof_begin, prof_end, 1, fused_counters(d->stream, d->batch_size), d->is_signed ? 1 : 0);
            if (fe != cudaSuccess) return cuda_fail(fe, "fast_pass_kernel launch");
            if (launched > 0) return GPUNTT_B200_OK;
        }
        if (rns && !d->is_signed && !g_force_generic.load() && d->ntt_layout == GPUNTT_B200_PER_POLYNOMIAL && !d->modulus_order_dev &&
            !d->poly_order_dev && d->mod_count > 1 && d->batch_size % d->mod_count != 0 &&
            ((long long) (d->batch_size - d->batch_size % d->mod_count) << n) >= kRaggedSplitMinElems)
        {
            // The tuned RNS kernels give every modulus slot the same number of polynomials.  A batch that is not a multiple of
            // mod_count is split the same way as a ragged small-ring batch: the whole rounds of slots on the tuned kernels, the last
            // (fewer than mod_count) polynomials -- whose index modulo mod_count is their index in the tail -- on the generic kernel.
            const int tail = d->batch_size % d->mod_count;
            gpuntt_b200_merge_desc part = *d;
            part.batch_size = d->batch_size - tail;
            int rc = merge_execute_t<T>(&part);
            if (rc != GPUNTT_B200_OK) return rc;
            const size_t skip = ((size_t) part.batch_size << n) * sizeof(T);
            part.in = reinterpret_cast<const char*>(d->in) + skip;
            part.out = reinterpret_cast<char*>(d->out) + skip;
            part.batch_size = tail;
            return merge_execute_t<T>(&part);
        }
        if (rns && !d->is_signed && !g_force_generic.load() && d->ntt_layout == GPUNTT_B200_PER_POLYNOMIAL)
        {
            void* flag = fused_counters(d->stream, d->batch_size); // progress counters of the single-launch kernels (may be null)
            int launched = 0;
            cudaError_t fe = fast_merge_rns<T>(reinterpret_cast<const T*>(d->in), reinterpret_cast<T*>(d->out),
                                   reinterpret_cast<const T*>(d->root_of_unity_table), reinterpret_cast<const T*>(d->modulus_dev),
                                   reinterpret_cast<const T*>(d->mod_inverse_dev), d->modulus_order_dev, d->poly_order_dev, d->mod_count, n,
                                   plus ? 1 : 0, inv, d->batch_size,
                                   reinterpret_cast<int*>(flag), st, &launched, prof_begin, prof_end);
            if (fe != cudaSuccess) return cuda_fail(fe, "fast_pass_kernel (RNS) launch");
            if (launched > 0) return GPUNTT_B200_OK;
        }
        // NTTLayout::PerCoefficient (ForwardCoreTranspose / InverseCoreTranspose, ntt.cu:1554-2074): the buffer is one
        // [2^n_power][batch] row-major matrix and every COLUMN is a transform.  That is a single strided pass over
        // the top n_power index bits of one array of 2^(n_power + log2 batch) elements -- no transposition anywhere.
        int col_log = 0;
        if (d->ntt_layout == GPUNTT_B200_PER_COEFFICIENT)
            while ((1 << col_log) < d->batch_size) col_log++;
        if constexpr (sizeof(T) == 8)
        {
            if (col_log > 0 && !rns && !g_force_generic.load())
            {
                int launched = 0;
                cudaError_t fe = fast_per_coefficient(reinterpret_cast<const uint64_t*>(d->in), reinterpret_cast<uint64_t*>(d->out),
                                                      reinterpret_cast<const uint64_t*>(d->root_of_unity_table), (uint64_t) d->modulus_value,
                                                      (uint64_t) d->mod_inverse_value, n, col_log, plus ? 1 : 0, inv, d->is_signed ? 1 : 0, st,
                                                      &launched, prof_begin, prof_end);
                if (fe != cudaSuccess) return cuda_fail(fe, "fast_pass_kernel (PerCoefficient) launch");
                if (launched > 0) return GPUNTT_B200_OK;
            }
        }
        else
        {
            if (col_log > 0 && !rns && !g_force_generic.load())
            {
                int launched = 0;
                cudaError_t fe = fast_per_coefficient32(reinterpret_cast<const uint32_t*>(d->in), reinterpret_cast<uint32_t*>(d->out),
                                                        reinterpret_cast<const uint32_t*>(d->root_of_unity_table), (uint32_t) d->modulus_value,
                                                        (uint32_t) d->mod_inverse_value, n, col_log, plus ? 1 : 0, inv, d->is_signed ? 1 : 0, st,
                                                        &launched, prof_begin, prof_end);
                if (fe != cudaSuccess) return cuda_fail(fe, "fast_pass_kernel (PerCoefficient, 32-bit) launch");
                if (launched > 0) return GPUNTT_B200_OK;
            }
        }
        CoreCall<T> cc;
        cc.in = d->in;
        cc.out = reinterpret_cast<T*>(d->out);
        cc.table = reinterpret_cast<const T*>(d->root_of_unity_table);
        cc.table_len = plus ? (1LL << n) : (1LL << (n - 1));
        cc.stride_log = n; // the reference's RNS tables are spaced (1 << n_power) apart for both ring types
        cc.n_power = n + col_log;
        cc.batch = col_log > 0 ? 1 : d->batch_size;
        cc.col_log = col_log;
        cc.mod_count = d->mod_count;
        cc.plus = plus ? 1 : 0;
        cc.signed_io = d->is_signed ? 1 : 0;
        cc.inverse = inv;
        cc.p = (T) d->modulus_value;
        cc.ninv = (T) d->mod_inverse_value;
        cc.mod_values = rns ? reinterpret_cast<const T*>(d->modulus_dev) : nullptr;
        cc.ninv_dev = rns ? reinterpret_cast<const T*>(d->mod_inverse_dev) : nullptr;
        cc.mod_order = rns ? d->modulus_order_dev : nullptr;
        cc.poly_order = rns ? d->poly_order_dev : nullptr;
        cc.st = st;
        cc.ws_slot = 0;
        if (col_log > 0)
        {
            cc.npasses = 1;
            cc.pass[0] = make_strided_pass(col_log, n, (int) sizeof(T) * 8);
            return run_core<T>(cc);
        }
        const MergePlan mp = make_merge_plan(n, (int) sizeof(T) * 8);
        cc.npasses = mp.npasses;
        for (int i = 0; i < mp.npasses; i++) cc.pass[i] = mp.pass[inv ? (mp.npasses - 1 - i) : i];
        return run_core<T>(cc);
    }

    static int merge_execute(const gpuntt_b200_merge_desc* d)
    {
        g_last_launches = 0;
        if (!d) return fail(GPUNTT_B200_ERR_ARGUMENT, "null descriptor");
        if (d->ntt_layout != GPUNTT_B200_PER_POLYNOMIAL && d->ntt_layout != GPUNTT_B200_PER_COEFFICIENT)
            return fail(GPUNTT_B200_ERR_LAYOUT, "Invalid ntt_layout!");
        if (d->n_power < 1 || d->n_power > 28) return fail(GPUNTT_B200_ERR_N_POWER, "Invalid n_power range!");
        if (d->ntt_layout == GPUNTT_B200_PER_COEFFICIENT)
        {
            // the reference's limits for this layout (ntt.cu:2230-2233): n_power 1..9; its launch arithmetic also
            // assumes a power-of-two batch (log2(batch_size), ntt.cu:2235) -- anything else is rejected here
            if (d->n_power > 9) return fail(GPUNTT_B200_ERR_N_POWER, "Invalid n_power range!");
            if (d->batch_size & (d->batch_size - 1))
                return fail(GPUNTT_B200_ERR_ARGUMENT, "NTTLayout::PerCoefficient needs a power-of-two batch_size");
            if (d->poly_order_dev || d->modulus_order_dev)
                return fail(GPUNTT_B200_ERR_ARGUMENT, "order arrays do not apply to NTTLayout::PerCoefficient");
        }
        if (d->element_bits != 32 && d->element_bits != 64)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "element_bits must be 32 or 64");
        if (d->batch_size < 0 || d->mod_count < 0) return fail(GPUNTT_B200_ERR_ARGUMENT, "negative batch_size / mod_count");
        if (d->direction != GPUNTT_B200_FORWARD && d->direction != GPUNTT_B200_INVERSE)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "direction must be FORWARD or INVERSE");
        if (d->reduction_poly != GPUNTT_B200_X_N_PLUS && d->reduction_poly != GPUNTT_B200_X_N_MINUS)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "reduction_poly must be X_N_plus or X_N_minus");
        if (d->batch_size == 0) return GPUNTT_B200_OK;
        if (!d->in || !d->out || !d->root_of_unity_table)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "null data / table pointer");
        if (d->mod_count > 0 && !d->modulus_dev) return fail(GPUNTT_B200_ERR_ARGUMENT, "RNS form needs modulus_dev");
        if (d->mod_count > 0 && d->direction == GPUNTT_B200_INVERSE && !d->mod_inverse_dev)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "RNS inverse needs mod_inverse_dev");
        if (d->mod_count == 0 && d->modulus_value < 2) return fail(GPUNTT_B200_ERR_ARGUMENT, "modulus_value < 2");
        if (d->element_bits == 64) return merge_execute_t<uint64_t>(d);
        return merge_execute_t<uint32_t>(d);
    }

#include "fourstep.inl"

} // namespace gpuntt_b200

using namespace gpuntt_b200;

extern "C"
{
    int gpuntt_b200_4step_ntt(const gpuntt_b200_4step_desc* desc) { return fourstep_execute(desc); }
    int gpuntt_b200_4step_shape(int n_power, int* n1, int* n2) { return fourstep_shape(n_power, n1, n2) ? GPUNTT_B200_OK : GPUNTT_B200_ERR_N_POWER; }
    int gpuntt_b200_transpose(int element_bits, const void* in, void* out, int row, int col, int n_power, int batch_size, void* stream)
    {
        return transpose_execute(element_bits, in, out, row, col, n_power, batch_size, stream);
    }

    int gpuntt_b200_merge_ntt(const gpuntt_b200_merge_desc* desc) { return merge_execute(desc); }

    static gpuntt_b200_merge_desc simple_desc(int bits, int dir, const void* in, void* out, const void* table,
                                              uint64_t p, uint64_t ninv, int n_power, int poly, int batch,
                                              void* stream)
    {
        gpuntt_b200_merge_desc d;
        memset(&d, 0, sizeof(d));
        d.element_bits = bits;
        d.direction = dir;
        d.n_power = n_power;
        d.ntt_layout = GPUNTT_B200_PER_POLYNOMIAL;
        d.reduction_poly = poly;
        d.batch_size = batch;
        d.in = in;
        d.out = out;
        d.root_of_unity_table = table;
        d.modulus_value = p;
        d.mod_inverse_value = ninv;
        d.stream = stream;
        return d;
    }

    int gpuntt_b200_ntt_u64(const uint64_t* in, uint64_t* out, const uint64_t* root_table, uint64_t modulus,
                            int n_power, int reduction_poly, int batch_size, void* stream)
    {
        gpuntt_b200_merge_desc d = simple_desc(64, GPUNTT_B200_FORWARD, in, out, root_table, modulus, 0, n_power,
                                               reduction_poly, batch_size, stream);
        return merge_execute(&d);
    }
    int gpuntt_b200_intt_u64(const uint64_t* in, uint64_t* out, const uint64_t* inv_root_table, uint64_t modulus,
                             uint64_t n_inverse, int n_power, int reduction_poly, int batch_size, void* stream)
    {
        gpuntt_b200_merge_desc d = simple_desc(64, GPUNTT_B200_INVERSE, in, out, inv_root_table, modulus, n_inverse,
                                               n_power, reduction_poly, batch_size, stream);
        return merge_execute(&d);
    }
    int gpuntt_b200_ntt_u32(const uint32_t* in, uint32_t* out, const uint32_t* root_table, uint32_t modulus,
                            int n_power, int reduction_poly, int batch_size, void* stream)
    {
        gpuntt_b200_merge_desc d = simple_desc(32, GPUNTT_B200_FORWARD, in, out, root_table, modulus, 0, n_power,
                                               reduction_poly, batch_size, stream);
        return merge_execute(&d);
    }
    int gpuntt_b200_intt_u32(const uint32_t* in, uint32_t* out, const uint32_t* inv_root_table, uint32_t modulus,
                             uint32_t n_inverse, int n_power, int reduction_poly, int batch_size, void* stream)
    {
        gpuntt_b200_merge_desc d = simple_desc(32, GPUNTT_B200_INVERSE, in, out, inv_root_table, modulus, n_inverse,
                                               n_power, reduction_poly, batch_size, stream);
        return merge_execute(&d);
    }

    // Host-buffer pipeline: the batch is cut into chunks of whole polynomials; chunk i's H2D copy, transform and D2H
    // copy run on three engine-owned streams chained by events, so the two DMA directions and the kernels overlap
    // (PCIe is full duplex).  Ordered after everything already enqueued on the caller's stream; synchronises it.
    struct HostPipe
    {
        cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
        cudaEvent_t start = nullptr, done = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_cmp;
    };
    static std::mutex g_pipe_mutex;
    static std::map<int, HostPipe> g_pipes; // per device

    static cudaError_t get_pipe(int nchunks, HostPipe** out)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        HostPipe& hp = g_pipes[dev];
        if (!hp.s_in)
        {
            if ((e = cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&hp.s_cmp, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&hp.start, cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&hp.done, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        while ((int) hp.ev_in.size() < nchunks)
        {
            cudaEvent_t a, b;
            if ((e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming)) != cudaSuccess) return e;
            hp.ev_in.push_back(a);
            hp.ev_cmp.push_back(b);
        }
        *out = &hp;
        return cudaSuccess;
    }

    int gpuntt_b200_merge_ntt_host(const gpuntt_b200_merge_desc* hd, const void* host_root_table,
                                   size_t root_table_elems)
    {
        g_last_launches = 0;
        if (!hd || !hd->in || !hd->out || !host_root_table) return fail(GPUNTT_B200_ERR_ARGUMENT, "null pointer");
        if (hd->mod_count != 0) return fail(GPUNTT_B200_ERR_UNSUPPORTED, "host convenience call is single-modulus only");
        if (hd->n_power < 1 || hd->n_power > 28) return fail(GPUNTT_B200_ERR_N_POWER, "Invalid n_power range!");
        if (hd->element_bits != 32 && hd->element_bits != 64)
            return fail(GPUNTT_B200_ERR_ARGUMENT, "element_bits must be 32 or 64");
        if (hd->batch_size < 0) return fail(GPUNTT_B200_ERR_ARGUMENT, "negative batch_size");
        const size_t esz = hd->element_bits / 8;
        const size_t poly_bytes = ((size_t) 1 << hd->n_power) * esz;
        const size_t data_bytes = (size_t) hd->batch_size * poly_bytes;
        const size_t table_bytes = root_table_elems * esz;
        cudaStream_t st = (cudaStream_t) hd->stream;
        // chunks of ~32 MiB, whole polynomials; PerCoefficient interleaves the batch, so it stays one chunk
        int chunk_polys = hd->batch_size;
        if (hd->ntt_layout == GPUNTT_B200_PER_POLYNOMIAL && hd->batch_size > 1)
        {
            size_t target = (size_t) 32 << 20;
            if (const char* ev = getenv("GPUNTT_B200_HOST_CHUNK_MB")) // (measurement knob: tools/e2e_chunk_sweep.py)
            {
                const long mb = atol(ev);
                if (mb > 0) target = (size_t) mb << 20;
            }
            size_t cp = poly_bytes >= target ? 1 : target / poly_bytes;
            if (cp < (size_t) hd->batch_size) chunk_polys = (int) cp;
        }
        const int nchunks = hd->batch_size == 0 ? 0 : (hd->batch_size + chunk_polys - 1) / chunk_polys;
        std::lock_guard<std::mutex> lk(g_pipe_mutex);
        HostPipe* hp = nullptr;
        cudaError_t e = get_pipe(nchunks, &hp);
        if (e != cudaSuccess) return cuda_fail(e, "host pipeline streams");
        void *dbuf = nullptr, *dtab = nullptr;
        e = get_workspace((void*) hp->s_cmp, 1, data_bytes, &dbuf);
        if (e != cudaSuccess) return cuda_fail(e, "staging allocation");
        e = get_workspace((void*) hp->s_cmp, 2, table_bytes, &dtab);
        if (e != cudaSuccess) return cuda_fail(e, "staging allocation");
        if ((e = cudaEventRecord(hp->start, st)) != cudaSuccess) return cuda_fail(e, "event record");
        cudaStreamWaitEvent(hp->s_in, hp->start, 0);
        cudaStreamWaitEvent(hp->s_cmp, hp->start, 0);
        cudaStreamWaitEvent(hp->s_out, hp->start, 0);
        e = cudaMemcpyAsync(dtab, host_root_table, table_bytes, cudaMemcpyHostToDevice, hp->s_in);
        if (e != cudaSuccess) return cuda_fail(e, "table H2D");
        int launches = 0;
        for (int c = 0; c < nchunks; c++)
        {
            const int b0 = c * chunk_polys;
            const int nb = (b0 + chunk_polys <= hd->batch_size) ? chunk_polys : hd->batch_size - b0;
            unsigned char* dchunk = static_cast<unsigned char*>(dbuf) + (size_t) b0 * poly_bytes;
            e = cudaMemcpyAsync(dchunk, static_cast<const unsigned char*>(hd->in) + (size_t) b0 * poly_bytes,
                                (size_t) nb * poly_bytes, cudaMemcpyHostToDevice, hp->s_in);
            if (e != cudaSuccess) return cuda_fail(e, "data H2D");
            cudaEventRecord(hp->ev_in[c], hp->s_in);
            cudaStreamWaitEvent(hp->s_cmp, hp->ev_in[c], 0);
            gpuntt_b200_merge_desc d = *hd;
            d.in = dchunk;
            d.out = dchunk;
            d.batch_size = nb;
            d.root_of_unity_table = dtab;
            d.stream = (void*) hp->s_cmp;
            int rc = merge_execute(&d);
            if (rc != GPUNTT_B200_OK) return rc;
            launches += g_last_launches;
            cudaEventRecord(hp->ev_cmp[c], hp->s_cmp);
            cudaStreamWaitEvent(hp->s_out, hp->ev_cmp[c], 0);
            e = cudaMemcpyAsync(static_cast<unsigned char*>(hd->out) + (size_t) b0 * poly_bytes, dchunk,
                                (size_t) nb * poly_bytes, cudaMemcpyDeviceToHost, hp->s_out);
            if (e != cudaSuccess) return cuda_fail(e, "data D2H");
        }
        g_last_launches = launches;
        cudaEventRecord(hp->done, hp->s_out);
        cudaStreamWaitEvent(st, hp->done, 0);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return cuda_fail(e, "stream synchronize");
        return GPUNTT_B200_OK;
    }

    int gpuntt_b200_last_launch_count(void) { return g_last_launches; }
    unsigned long long gpuntt_b200_total_launch_count(void) { return g_total_launches.load(); }
    const char* gpuntt_b200_last_error(void) { return g_last_error.c_str(); }

    void gpuntt_b200_release_workspaces(void)
    {
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        for (auto& kv : g_ws)
            if (kv.second.ptr)
            {
                cudaSetDevice(std::get<0>(kv.first));
                cudaDeviceSynchronize();
                cudaFree(kv.second.ptr);
            }
        g_ws.clear();
        g_modcache.clear();
    }

    int gpuntt_b200_describe_plan(int n_power, int element_bits, char* buf, size_t buf_len)
    {
        if (n_power < 1 || n_power > 28 || (element_bits != 32 && element_bits != 64)) return -1;
        MergePlan mp = make_merge_plan(n_power, element_bits);
        std::string s;
        char tmp[256];
        {
            // single-modulus unsigned PerPolynomial calls with a modulus in the lazy-policy range take the tuned kernels
            char fb[512];
            const int nf = fast_describe(n_power, element_bits, fb, sizeof(fb));
            if (nf > 0)
            {
                s = std::string("tuned: ") + fb + "| generic: ";
                if (buf && buf_len) snprintf(buf, buf_len, "%s", s.c_str());
            }
        }
        for (int i = 0; i < mp.npasses; i++)
        {
            const PassPlan& p = mp.pass[i];
            snprintf(tmp, sizeof(tmp), "pass%d{tile=2^%d %s lo=%d d=%d c=%d rounds=", i, p.tile_log,
                     p.lo ? "strided" : "contiguous", p.lo, p.d, p.c);
            s += tmp;
            for (int r = 0; r < p.nrounds; r++)
            {
                snprintf(tmp, sizeof(tmp), "%s%d", r ? "," : "", p.round_bits[r]);
                s += tmp;
            }
            s += "} ";
        }
        if (buf && buf_len)
        {
            strncpy(buf, s.c_str(), buf_len - 1);
            buf[buf_len - 1] = 0;
        }
        return mp.npasses;
    }

    void gpuntt_b200_set_profiling(int on) { g_profiling.store(on ? 1 : 0); }
    void gpuntt_b200_force_generic_path(int on) { g_force_generic.store(on ? 1 : 0); }
    void gpuntt_b200_tune(int knob, int value)
    {
        switch (knob)
        {
            case GPUNTT_B200_TUNE_FUSED_PASSES:
                g_fused_enabled.store(value ? 1 : 0);
                fused_set_policy(value);
                break;
            case GPUNTT_B200_TUNE_FUSED_LAG: fused_set_lag_steps(value); break;
            case GPUNTT_B200_TUNE_4STEP_TRANSPOSED: g_fourstep_transposed.store(value ? 1 : 0); break;
            case GPUNTT_B200_TUNE_4STEP_MODULUS_CACHE: g_fourstep_modcache.store(value); break;
            case GPUNTT_B200_TUNE_4STEP_RESIDENT_PAIRS: fourstep_set_resident_pairs(value); break;
            case GPUNTT_B200_TUNE_ONE_TILE: fast_set_one_tile_mode(value); break;
            case GPUNTT_B200_TUNE_SMALL_TILE_ELEMS: fused_set_small_tile_elems(value); break;
            case GPUNTT_B200_TUNE_SINGLE_POLY_TILES: fast_set_single_poly_tiles(value); break;
            default: break;
        }
    }

    int gpuntt_b200_profile_read(float* ms_out, int* kind_out, int max_records)
    {
        std::vector<ProfRec> recs;
        {
            std::lock_guard<std::mutex> lk(g_prof_mutex);
            recs.swap(g_prof);
        }
        int n = 0;
        for (auto& r : recs)
        {
            float ms = 0.f;
            cudaEventSynchronize(r.e1);
            cudaEventElapsedTime(&ms, r.e0, r.e1);
            if (n < max_records)
            {
                if (ms_out) ms_out[n] = ms;
                if (kind_out) kind_out[n] = r.kind;
                n++;
            }
            std::lock_guard<std::mutex> lk(g_prof_mutex);
            g_prof_pool[r.dev].push_back(r.e0);
            g_prof_pool[r.dev].push_back(r.e1);
        }
        return n;
    }

    static int batch_copy(bool scatter, const void* one, int one_dev, void* const* many, const int* devices, int ndev, size_t poly_bytes,
                          long long batch, int mod_count, void* const* streams)
    {
        if (!one || !many || !devices || ndev < 1 || batch < 0 || mod_count < 0) return fail(GPUNTT_B200_ERR_ARGUMENT, "bad scatter / gather arguments");
        const long long unit = mod_count > 0 ? mod_count : 1;
        if (batch % unit) return fail(GPUNTT_B200_ERR_ARGUMENT, "batch_size must be a multiple of mod_count");
        const long long groups = batch / unit;
        int cur = 0;
        cudaGetDevice(&cur);
        for (int g = 0; g < ndev; g++)
        {
            const long long lo = groups * g / ndev * unit, hi = groups * (g + 1) / ndev * unit;
            if (hi == lo) continue;
            if (!many[g]) return fail(GPUNTT_B200_ERR_ARGUMENT, "null slice pointer");
            if (devices[g] != one_dev)
            {
                // peer access in both directions (an error here only means the copy is staged through the host)
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[g], one_dev) == cudaSuccess && can)
                {
                    cudaSetDevice(devices[g]);
                    if (cudaDeviceEnablePeerAccess(one_dev, 0) != cudaSuccess) cudaGetLastError();
                    cudaSetDevice(one_dev);
                    if (cudaDeviceEnablePeerAccess(devices[g], 0) != cudaSuccess) cudaGetLastError();
                }
            }
            cudaSetDevice(devices[g]);
            cudaStream_t st = streams ? (cudaStream_t) streams[g] : nullptr;
            const unsigned char* whole = static_cast<const unsigned char*>(one) + (size_t) lo * poly_bytes;
            cudaError_t e = scatter ? cudaMemcpyPeerAsync(many[g], devices[g], whole, one_dev, (size_t) (hi - lo) * poly_bytes, st)
                                    : cudaMemcpyPeerAsync(const_cast<unsigned char*>(whole), one_dev, many[g], devices[g],
                                                          (size_t) (hi - lo) * poly_bytes, st);
            if (e != cudaSuccess)
            {
                cudaSetDevice(cur);
                return cuda_fail(e, scatter ? "scatter_batch copy" : "gather_batch copy");
            }
        }
        cudaSetDevice(cur);
        return GPUNTT_B200_OK;
    }
    int gpuntt_b200_scatter_batch(const void* src, int src_device, void* const* dst, const int* devices, int ndev, size_t poly_bytes,
                                  long long batch_size, int mod_count, void* const* streams)
    {
        return batch_copy(true, src, src_device, dst, devices, ndev, poly_bytes, batch_size, mod_count, streams);
    }
    int gpuntt_b200_gather_batch(void* dst, int dst_device, const void* const* src, const int* devices, int ndev, size_t poly_bytes,
                                 long long batch_size, int mod_count, void* const* streams)
    {
        return batch_copy(false, dst, dst_device, const_cast<void* const*>(src), devices, ndev, poly_bytes, batch_size, mod_count, streams);
    }

    void gpuntt_b200_example_input(uint32_t seed, uint64_t modulus, uint64_t count, uint64_t* host_out)
    {
        if (!host_out || modulus == 0) return;
        std::mt19937 gen(seed);
        std::uniform_int_distribution<std::uint64_t> dis(0, modulus - 1);
        for (uint64_t i = 0; i < count; i++) host_out[i] = dis(gen);
    }

    int gpuntt_b200_version(void) { return GPUNTT_B200_VERSION; }
}
