// gpu_ntt_b200/csrc/fast_kernels.cuh -- device code and launch helper shared by the tuned Merge-NTT translation
// units (merge_fast.cu, merge_fast_rns.cu, merge_fast_4step.cu, merge_fused.cu).
//
// One persistent, warp-specialised kernel per pass:
//   * 8 consumer warps do nothing but shared-memory <-> register butterfly rounds;
//   * 1 producer warp moves every coefficient tile with the TMA engine: ONE cp.async.bulk.tensor
//     (UTMALDG) per 32 KiB tile global -> shared for the next tile while the current one is being
//     transformed (two tile buffers, mbarrier full/done hand-shake) and one UTMASTG per finished
//     tile shared -> global.  Strided tiles are 3-D boxes [2^D matrix rows x column blocks x 128 bytes],
//     contiguous tiles [polynomials x rows x 128 bytes] (4-D with a modulus-slot dimension for RNS);
//     out-of-range polynomials are clipped by the hardware;
//   * tiles land in shared memory in the hardware SWIZZLE_128B layout (16-byte chunk index XOR
//     row mod 8), which makes every round shape bank-conflict free (16-byte accesses for the
//     lowest round);
//   * the twiddles of a pass are turned into (w, w') Shoup pairs once per CTA and twiddle range,
//     straight from the caller's table, into a slot-major shared-memory layout -- no scratch memory
//     and no pre-kernel; the contiguous pass pins each CTA to one range of the ring and streams
//     polynomials through it, so the ~N distinct twiddles of the final stages are fetched once per
//     CTA instead of once per polynomial.
// Replaces ForwardCore/InverseCore (src/lib/ntt_merge/ntt.cu:435-1318) and the FourStep*Core kernels
// (src/lib/ntt_4step/ntt_4step.cu:571-2291) of the reference for the shapes the tuned entry points list;
// everything else takes the generic pass kernel in merge_ntt.cu.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <type_traits>

#include "gpuntt_b200.h"
#include "merge_ntt.cuh"
#include "modarith.cuh"

namespace gpuntt_b200
{

#ifndef GPUNTT_FAST_P0_BLOCKS
#define GPUNTT_FAST_P0_BLOCKS 2 // CTAs per SM of the forward strided pass (3 fits at 72 registers but measured 4% slower)
#endif
#ifndef GPUNTT_FAST_P1_BLOCKS
#define GPUNTT_FAST_P1_BLOCKS 2 // CTAs per SM of the forward contiguous pass
#endif
#ifndef GPUNTT_FAST_P1_NPLOG
#define GPUNTT_FAST_P1_NPLOG 1  // log2 polynomials per tile of the forward contiguous pass
#endif
#ifndef GPUNTT_FAST_SP_BLOCKS
#define GPUNTT_FAST_SP_BLOCKS 2 // CTAs per SM of the single-polynomial contiguous pass (one tile per twiddle segment: latency-bound)
#endif
    constexpr int kConsumers = 256;
    constexpr int kFastThreads = kConsumers + 32;

    // ------------------------------------------------------------------ PTX helpers
    __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    }
    __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_arrive(uint32_t bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
    {
        asm volatile("{\n\t"
                     ".reg .pred P;\n\t"
                     "WAIT_%=:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
                     "@P bra DONE_%=;\n\t"
                     "bra WAIT_%=;\n\t"
                     "DONE_%=:\n\t"
                     "}" ::"r"(bar),
                     "r"(parity)
                     : "memory");
    }
    __device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar)
    {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(map), "r"(c0), "r"(c1), "r"(bar)
                     : "memory");
    }
    __device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
    {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                     "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                     : "memory");
    }
    __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar)
    {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                     "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src)
    {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map), "r"(c0),
                     "r"(c1), "r"(c2), "r"(c3), "r"(src)
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src)
    {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                     "r"(c1), "r"(src)
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src)
    {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                     "r"(c1), "r"(c2), "r"(src)
                     : "memory");
    }
    __device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
    {
        asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
    }
    __device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
    __device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
    __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#ifdef GPUNTT_EXPERIMENT_NOSYNC // timing experiment only (wrong results): upper bound of what a barrier-free round structure could gain
    __device__ __forceinline__ void consumer_sync(int = 1) {}
#else
    // named barrier of one consumer group (256 threads); the fused kernels run two groups per CTA (ids 1 and 2)
    __device__ __forceinline__ void consumer_sync(int id = 1) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kConsumers) : "memory"); }
#endif

    // ------------------------------------------------------------------ compile-time pass shape
    // STRIDED: tile = 2^D rows (row stride 2^lo elements) x 2^C adjacent columns, K = D + C.
    // !STRIDED: tile = 2^NPLOG polynomials x 2^KC adjacent elements of the same ring range.
    // Two register rounds: R1 stages on the high bits, R2 on the low bits (D = R1 + R2).
    // POL: arithmetic policy -- 0 exact (any modulus the reference accepts), 1 lazy (inverse; forward for 61-bit moduli
    // above the F60 range: correction on every stage), 2 F60 / L32 (forward).
    // NT: contiguous passes of transforms SHORTER than a tile row group (the 4-step inverse row phase, N = n1 <= 256): twiddles
    // depend on the low NT index bits only and a tile holds several whole transforms.
    // R3: a third register round below the other two (contiguous NT passes only): whole transforms of 2^9 .. 2^12 elements in
    // ONE pass (the small rings, fast_small below).
    template <typename T_, bool INV_, int POL_, bool STRIDED_, int R1_, int R2_, int K_, int NPLOG_, int NT_ = 0, int R3_ = 0> struct Shape
    {
        using T = T_;
        static constexpr int POL = POL_;
        static constexpr bool INV = INV_, FAST = POL_ != 0, STRIDED = STRIDED_;
        static_assert(POL_ == 0 || POL_ == 1 || (POL_ == 2 && !INV_), "policy 2 is forward-only");
        static constexpr int R1 = R1_, R2 = R2_, R3 = R3_, K = K_, NPLOG = NPLOG_;
        static constexpr int D = R1 + R2 + R3;
        static_assert(R3_ == 0 || (!STRIDED_ && NT_ > 0 && R2_ > 0), "three rounds: contiguous whole-transform passes only");
        static constexpr int C = STRIDED ? (K - D) : 0;
        static constexpr int KC = K - NPLOG;          // contiguous elements per polynomial in a tile
        static constexpr int NT = NT_;
        static constexpr int KTW = STRIDED ? K : (NT_ ? NT_ : KC);  // local index bits that select twiddles
        static constexpr int CB = (sizeof(T) == 8) ? 4 : 5; // log2 elements per 128-byte row
        static constexpr int ROWS = (1 << K) >> CB;
        static constexpr int TILE_SMEM = ROWS * 128;
        static constexpr int LB1 = C + R2 + R3, LB2 = C + R3, LB3 = C;   // lowest local bit of the high / low / third round
        static constexpr int G1 = 1 << (KTW - LB1 - R1), G2 = 1 << (KTW - LB2 - R2); // twiddle groups
        static constexpr int G3 = R3 > 0 ? (1 << (KTW - LB3 - R3)) : 1;
        static constexpr int TW1 = ((1 << R1) - 1) * G1, TW2 = ((1 << R2) - 1) * G2, TW3 = ((1 << R3) - 1) * G3;
        // inverse: a second copy of the high round's pairs, multiplied by n^-1 (the last round of an inverse transform folds the
        // scaling into its twiddles, see fast_round)
        // (only shapes that can end a transform: strided passes and whole-transform tiles)
        static constexpr int TW1C = (INV_ && (STRIDED_ || NT_ > 0)) ? TW1 : 0;
        static constexpr int TWN = TW1 + TW2 + TW3 + TW1C;
        static constexpr int TW_SMEM = TWN * (int) sizeof(Twiddle<T>);
        static constexpr int SMEM = 2 * TILE_SMEM + TW_SMEM + 128 + 1024; // barriers, segment constants + slack to align the tiles to 1 KiB
        static_assert(LB2 == 0 || LB2 >= CB, "low round must start at bit 0 or on a row boundary");
        static_assert(LB1 >= CB, "high round must start on a row boundary");
    };

    template <typename T, int POL> struct ModSel
    {
        using type = Mod<T, false>;
    };
    template <> struct ModSel<uint64_t, 1>
    {
        using type = Mod<uint64_t, true>;
    };
    template <> struct ModSel<uint64_t, 2>
    {
        using type = ModF60;
    };
    template <> struct ModSel<uint32_t, 2>
    {
        using type = ModL32;
    };
    template <typename S> struct ModOf
    {
        using type = typename ModSel<typename S::T, S::POL>::type;
    };

    template <typename T> struct FastArgs
    {
        const T* in;
        T* out;
        const T* table; // the caller's bit-reversed root table (w only)
        T p, ninv_w, ninv_wq;
        uint64_t mu; // 64-bit: floor(2^(63 + pbits) / p), 32-bit: floor(2^64 / p) -- companions without a division
        int pbits; // bit length of p
        int n, lo, plus, first, last, batch, rr;
        int in_bound; // forward first pass of a cyclic transform: inputs are below in_bound * p (0/1: canonical)
        int w_lazy;   // WMUL forward kernels: leave the products below 2p (one correction) instead of canonical
        int signed_io; // Data32s / Data64s: forward = signed input (x < 0 -> x + p as the first pass loads, modular_arith.cuh:372-385 of
                       // the reference), inverse = centred signed output (r > p/2 -> r - p after n^-1, modular_arith.cuh:389-405)
        int n_tw; // transform size (log2) for twiddle indexing when it differs from the layout size n (0: n)
        int tw_fixed; // strided passes whose transforms are SHORTER than the row count (size 2^D transforms on every block of 2^D
                      // matrix rows: the 4-step inverse row phase read from the n2 x n1 matrix): one twiddle set for every range
        int cc_major; // strided passes: inside a range the tiles are ordered column-chunk-major, polynomial-minor (default:
                      // polynomial-major), so the tiles that share the twiddle-matrix pairs of a position follow each other
        int cta_per_seg, seg_extra; // cta_per_seg > 0: every CTA works inside ONE twiddle segment; the first seg_extra segments get
                                    // cta_per_seg + 1 CTAs, the others cta_per_seg (set by launch_fast)
        long long work; // total tiles of this pass
        // RNS kernels (polynomial b uses modulus slot b % mod_count, ntt.cu:613-619 of the reference): `batch` is then the
        // number of polynomials PER SLOT, mod_dev the device array of {value, bit, mu} triples, ninv_dev the per-slot
        // N^-1; policy_flag/want_policy: a kernel returns at once unless *policy_flag == want_policy (the moduli live on
        // the device, so the lazy-policy and the exact-policy kernel are both enqueued and the data decides)
        const T* mod_dev;
        const T* ninv_dev;
        const int* policy_flag;
        const int* poly_order; // GPU_NTT_Poly_Ordered: the b-th transform (slot b % mod_count) lives in polynomial poly_order[b]
        const int* mod_order; // GPU_NTT_Modulus_Ordered: slot m uses entry mod_order[m] of the modulus / table / N^-1 arrays
        int mod_count, want_policy;
        const void* w_pairs; // WMUL kernels: Twiddle<T>[N], the 4-step twiddle matrix with Shoup companions
    };

    // byte offset of local element l inside a (1 KiB aligned) tile buffer: TMA SWIZZLE_128B
    template <typename S> __device__ __forceinline__ int tile_off(int l)
    {
        const int b = l * (int) sizeof(typename S::T);
        return b ^ (((b >> 7) & 7) << 4);
    }

    // ------------------------------------------------------------------ one register round
    // TRIV: the twiddle in slot 0 of every stage is 1 (first round of an X^N-1 transform: table[0] = omega^0),
    // so those butterflies skip the multiply (15 of the 32 butterflies of a radix-16 round).
    // WMUL (forward strided passes only): after the stages every element is multiplied by the (w, w') pair at its
    // offset inside the polynomial -- the 4-step twiddle-matrix product (w_pairs: see fast_fourstep_columns) -- and
    // canonicalised.  wtile points at the pair of the tile's first element, lo is the pass's row stride (log2).
    // TS (forward strided passes, the round on the LOWEST row bits): transposed store.  Every thread keeps its results in
    // registers, the consumer group synchronises (all loads of the tile are done), and the results are written in the
    // layout of the transposing TMA store box {16 rows, 2^C columns, 2^(D-4) row blocks}: the 128-byte line
    // (row >> 4) * 2^C + column holds rows (row & ~15) .. +15 of that column, 16-byte chunks XOR (line & 7) like every
    // SWIZZLE_128B tile -- so a thread's consecutive rows are one vector store per pair and the eight lanes of a quarter
    // warp (eight columns) hit eight different bank groups.  This is how the 4-step column phase writes its output as the
    // n2 x n1 matrix without a transpose kernel.
    // WL: the (w, w') pairs of the twiddle-matrix product sit in SHARED memory in tile order (merge_wcol.cu passes lo = C so that
    // the pair of local element l is entry l): plain loads instead of __ldg.
    template <typename S, int R, int LB, int G, bool FINAL, bool TRIV = false, bool WMUL = false, bool TS = false, bool WL = false>
    __device__ __forceinline__ void fast_round(unsigned char* buf, const Twiddle<typename S::T>* __restrict__ tws,
                                               const typename ModOf<S>::type& M, int ctid,
                                               const Twiddle<typename S::T>& ninv,
                                               const Twiddle<typename S::T>* __restrict__ wtile = nullptr, int lo = 0, int in_bound = 1,
                                               bool w_lazy = false, int bar = 1)
    {
        using T = typename S::T;
        constexpr int E = 1 << R;
        constexpr int ITEMS = (1 << S::K) >> R;
        constexpr int VN = 16 / (int) sizeof(T);
        constexpr int ES = (int) sizeof(T);
        // Swizzled address of element a of the item: with b = byte offset of (l_base | a << LB),
        //   addr = b ^ (((b >> 7) & 7) << 4).
        // LB*ES >= 1 KiB rows apart (LB >= CB + 3): the XOR term comes from l_base only.
        // CB <= LB < CB + 3: bits 7..9 of b come from a (l_base has zeros there) -> XOR constant per a.
        // LB == 0: the item is one or more whole 128-byte rows; chunks are 16-byte vectors.
        static_assert(LB == 0 || LB >= S::CB, "round must start at bit 0 or on a row boundary");
        constexpr int NI = TS ? (ITEMS / kConsumers) : 1;
        static_assert(!TS || (S::STRIDED && LB >= S::C && R >= 1 && S::D >= 5 && sizeof(T) == 8 && ITEMS % kConsumers == 0),
                      "transposed store: last executed round of a strided 64-bit pass");
        T keep[NI][TS ? E : 1];
#pragma unroll(TS ? NI : 1)
        for (int item = ctid, ii = 0; item < ITEMS; item += kConsumers, ii++)
        {
            const int l_base = ((item >> LB) << (LB + R)) | (item & ((1 << LB) - 1));
            const int group = (l_base >> (LB + R)) & (G - 1);
            const int b0 = l_base * ES;
            T e[E];
            auto addr = [&](int a) -> unsigned char*
            {
                // a is a compile-time constant after unrolling
                const int b = b0 + ((a << LB) * ES);         // no carries: the a-bits of l_base are zero
                if constexpr (LB >= S::CB + 3)
                    return buf + ((b0 ^ (((b0 >> 7) & 7) << 4)) + ((a << LB) * ES));
                else
                    return buf + (b ^ (((b >> 7) & 7) << 4));
            };
            if constexpr (LB == 0)
            {
#pragma unroll
                for (int a = 0; a < E; a += VN)
                {
                    if constexpr (sizeof(T) == 8)
                    {
                        ulonglong2 v = *reinterpret_cast<const ulonglong2*>(addr(a));
                        e[a] = v.x;
                        e[a + 1] = v.y;
                    }
                    else
                    {
                        uint4 v = *reinterpret_cast<const uint4*>(addr(a));
                        e[a] = v.x;
                        e[a + 1] = v.y;
                        e[a + 2] = v.z;
                        e[a + 3] = v.w;
                    }
                }
            }
            else
            {
#pragma unroll
                for (int a = 0; a < E; a++) e[a] = *reinterpret_cast<const T*>(addr(a));
            }

            // 4-step twiddle-matrix product on the item's elements (forward: epilogue, canonical results; inverse:
            // prologue, lazy results in [0,3p)).  Loads in batches of 8 (32 registers): ptxas otherwise serialises
            // load -> multiply -> load and exposes one global-memory latency per element.
            auto w_product = [&](bool canonical)
            {
                if constexpr (WMUL)
                {
                    static_assert(!WMUL || (S::STRIDED && LB >= S::C) || (WL && !S::STRIDED && LB == 0),
                                  "the twiddle-matrix product belongs to strided passes (contiguous passes: resident pairs, lowest round)");
                    // offset of element a: row (l >> C) * 2^lo + column (l & (2^C - 1)); a only moves the row
                    const Twiddle<T>* wp = wtile + (((long long) (l_base >> S::C)) << lo) + (l_base & ((1 << S::C) - 1));
                    constexpr int WB = WL ? (E < 4 ? E : 4) : (E < 8 ? E : 8); // (shared-memory pairs: short latency, fewer registers)
#pragma unroll
                    for (int h = 0; h < E; h += WB)
                    {
                        ulonglong2 v[WB];
#pragma unroll
                        for (int j = 0; j < WB; j++)
                        {
                            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(wp + (((long long) (h + j) << (LB - S::C)) << lo));
                            // contiguous tiles: the pair table is built element-major inside a tile position (entry a * ITEMS + item is
                            // the pair of local element item * E + a, merge_wcol.cu), so a warp's loads are 32 consecutive pairs
                            if constexpr (WL && !S::STRIDED) src = reinterpret_cast<const ulonglong2*>(wtile + (h + j) * ITEMS + item);
                            if constexpr (WL)
                                v[j] = *src;
                            else
                                v[j] = __ldg(src);
                        }
#pragma unroll
                        for (int j = 0; j < WB; j++)
                        {
                            const Twiddle<T> tw{v[j].x, v[j].y};
                            const T r = M.mul(e[h + j], tw); // any 64-bit value in, [0,3p) out
                            if constexpr (S::INV)
                                e[h + j] = r; // the Gentleman-Sande butterflies take [0,4p)
                            else
                                e[h + j] = canonical ? csub(csub(r, M.p + M.p), M.p) : csub(r, M.p + M.p);
                        }
                    }
                }
            };
            const Twiddle<T>* tg = tws + group;
            if constexpr (!S::INV)
            {
#pragma unroll
                for (int it = 0; it < R; it++)
                {
                    const int ab = R - 1 - it;
#pragma unroll
                    for (int x = 0; x < (E >> (ab + 1)); x++)
                    {
                        if constexpr (S::POL == 2)
                        {
                            // first round of a cyclic transform, canonical inputs: before stage it the values are below
                            // {1, 2, 6, 12}[it] * p, the twiddle-1 butterflies of stages 0..2 are bare add/subtract
                            if (TRIV && it < 3 && x == 0)
                            {
                                const T K = M.triv_bound(it, in_bound);
#pragma unroll
                                for (int y = 0; y < (1 << ab); y++) M.add_sub(e[y], e[y | (1 << ab)], K);
                                continue;
                            }
                            const Twiddle<T> w = tg[((E >> (ab + 1)) - 1 + x) * G];
                            // correction on every other stage, ending each round with one; after the special stages of a
                            // first cyclic round every stage corrects (their outputs may sit at the policy's cap)
                            const bool kindB = TRIV ? (it >= 3) : (ab % 2 == 0);
#pragma unroll
                            for (int y = 0; y < (1 << ab); y++)
                            {
                                const int a0 = (x << (ab + 1)) | y;
                                if (kindB)
                                    M.ctB(e[a0], e[a0 | (1 << ab)], w);
                                else
                                    M.ctA(e[a0], e[a0 | (1 << ab)], w);
                            }
                        }
                        else
                        {
                            const Twiddle<T> w = tg[((E >> (ab + 1)) - 1 + x) * G];
#pragma unroll
                            for (int y = 0; y < (1 << ab); y++)
                            {
                                const int a0 = (x << (ab + 1)) | y;
                                M.ct(e[a0], e[a0 | (1 << ab)], w);
                            }
                        }
                    }
                }
                if constexpr (S::POL == 2 && TRIV && R < 4)
                {
                    // A first cyclic round of fewer than four stages ends on the bare add/subtract stages: from inputs
                    // below 2p (in_bound 2, the 4-step row phase after a lazy column phase) its twiddle-1 outputs reach 16p,
                    // but the next round may open with a kind-A stage, which needs X < 12p + 2^32 (X + 4p - t would wrap
                    // for p above 0.8 * 2^60).  One high-word correction restores the round-boundary invariant.
                    if (in_bound > 1)
                    {
#pragma unroll
                        for (int a = 0; a < E; a++) e[a] = M.csub8_hi(e[a]);
                    }
                }
                if constexpr (FINAL)
                {
#pragma unroll
                    for (int a = 0; a < E; a++) e[a] = M.canon_fwd(e[a]);
                }
                if constexpr (WMUL) w_product(!w_lazy);
            }
            else
            {
                if constexpr (WMUL) w_product(false);
                if constexpr (FINAL)
                {
                    // Last round of the transform: the scaling by n^-1 rides in the twiddles.  A value picks the factor up at the
                    // first stage of this round in which it is the multiplied output of its butterfly (both inputs of a butterfly
                    // whose low index bits y are non-zero already carry it), so only element 0 -- never multiplied -- needs a
                    // product of its own: 1 instead of 2^R extra multiplies per item.  tgc: the round's pairs times n^-1
                    // (build_twiddles).  TRIV (X^N-1, top of the transform): the twiddle in slot 0 of every stage is 1, so the
                    // butterflies with x == 0 and y != 0 need no multiply at all.
                    const Twiddle<T>* tgc = tg + (S::TW1 + S::TW2 + S::TW3);
#pragma unroll
                    for (int ab = 0; ab < R; ab++)
                    {
#pragma unroll
                        for (int x = 0; x < (E >> (ab + 1)); x++)
                        {
                            const int slot = (E >> (ab + 1)) - 1 + x;
                            const Twiddle<T> wc = tgc[slot * G];
                            Twiddle<T> w = wc;
                            if (ab > 0 && !(TRIV && x == 0)) w = tg[slot * G];
#pragma unroll
                            for (int y = 0; y < (1 << ab); y++)
                            {
                                const int a0 = (x << (ab + 1)) | y;
                                if (y == 0)
                                    M.gs(e[a0], e[a0 | (1 << ab)], wc);
                                else if (TRIV && x == 0)
                                    M.gs_one(e[a0], e[a0 | (1 << ab)]);
                                else
                                    M.gs(e[a0], e[a0 | (1 << ab)], w);
                            }
                        }
                    }
                    e[0] = M.canon_inv(e[0], ninv);
#pragma unroll
                    for (int a = 1; a < E; a++) e[a] = M.canon_lazy_inv(e[a]);
                }
                else
                {
#pragma unroll
                    for (int ab = 0; ab < R; ab++)
                    {
#pragma unroll
                        for (int x = 0; x < (E >> (ab + 1)); x++)
                        {
                            const Twiddle<T> w = tg[((E >> (ab + 1)) - 1 + x) * G];
#pragma unroll
                            for (int y = 0; y < (1 << ab); y++)
                            {
                                const int a0 = (x << (ab + 1)) | y;
                                M.gs(e[a0], e[a0 | (1 << ab)], w);
                            }
                        }
                    }
                }
            }

            if constexpr (TS)
            {
#pragma unroll
                for (int a = 0; a < E; a++) keep[ii][a] = e[a];
            }
            else if constexpr (LB == 0)
            {
#pragma unroll
                for (int a = 0; a < E; a += VN)
                {
                    if constexpr (sizeof(T) == 8)
                        *reinterpret_cast<ulonglong2*>(addr(a)) = make_ulonglong2(e[a], e[a + 1]);
                    else
                        *reinterpret_cast<uint4*>(addr(a)) = make_uint4(e[a], e[a + 1], e[a + 2], e[a + 3]);
                }
            }
            else
            {
#pragma unroll
                for (int a = 0; a < E; a++) *reinterpret_cast<T*>(addr(a)) = e[a];
            }
        }
        if constexpr (TS)
        {
            consumer_sync(bar); // every thread of the group has read its elements: the buffer can take the new layout
#pragma unroll
            for (int ii = 0; ii < NI; ii++)
            {
                const int item = ctid + ii * kConsumers;
                const int l_base = ((item >> LB) << (LB + R)) | (item & ((1 << LB) - 1));
                const int col = l_base & ((1 << S::C) - 1), row0 = l_base >> S::C; // row0 has zeros in the R bits of this round
                if constexpr (LB == S::C)
                {
                    // the round on the lowest row bits (forward passes): consecutive rows, one vector store per pair
#pragma unroll
                    for (int a = 0; a < E; a += 2)
                    {
                        const int row = row0 | a;
                        const int line = ((row >> 4) << S::C) + col;
                        const int off = (line << 7) + (((((row & 15) >> 1) ^ (col & 7))) << 4);
                        *reinterpret_cast<ulonglong2*>(buf + off) = make_ulonglong2(keep[ii][a], keep[ii][a + 1]);
                    }
                }
                else
                {
                    // a round on higher row bits (the last round of an INVERSE pass): rows 2^(LB - C) apart, one 8-byte store
                    // each (two-way bank conflicts between columns j and j + 8; 16 stores per thread and tile)
#pragma unroll
                    for (int a = 0; a < E; a++)
                    {
                        const int row = row0 | (a << (LB - S::C));
                        const int line = ((row >> 4) << S::C) + col;
                        const int off = (line << 7) + (((((row & 15) >> 1) ^ (col & 7))) << 4) + ((row & 1) << 3);
                        *reinterpret_cast<T*>(buf + off) = keep[ii][a];
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------ twiddle pairs of one (pass, range)
    // (w, w') pairs for the rounds of shape S, slot-major: entry (slot, group) at slot*G + group; tw1 / tw2 / tw3 are
    // consecutive.  `range`: the index bits above this pass's stage window (see fast_pass_body); lo / plus / n / n_tw as in
    // FastArgs.  Called by every thread of the CTA (thread t of nthreads).
    // scaled (inverse shapes, the pass that ends the transform): the high round's pairs times n^-1 go behind the three tables.
    template <typename S>
    __device__ __forceinline__ void build_twiddles(Twiddle<typename S::T>* tw1, const typename S::T* __restrict__ seg_table, int range, int n,
                                                   int n_tw, int lo, int plus, typename S::T seg_p, uint64_t seg_mu, int seg_pbits, int t,
                                                   int nthreads, bool scaled = false, Twiddle<typename S::T> ninv = Twiddle<typename S::T>{0, 0})
    {
        using T = typename S::T;
        Twiddle<T>* tw2 = tw1 + S::TW1;
        Twiddle<T>* tw3 = tw2 + S::TW2;
        const int j0 = S::STRIDED ? (range << S::D) : (S::NT ? 0 : (range << S::KC)); // index (>> lo) of the tile's first row
        const int ntw = n_tw ? n_tw : n;
        constexpr int TOTAL = S::TW1 + S::TW2 + S::TW3;
        // position of entry i in the caller's table (G1 / G2 / G3 are powers of two)
        auto locate = [&](int i, int& ii, bool& hi, bool& third) -> long long
        {
            hi = i < S::TW1;
            third = S::R3 > 0 && i >= S::TW1 + S::TW2;
            ii = hi ? i : (third ? i - S::TW1 - S::TW2 : i - S::TW1);
            const int R = hi ? S::R1 : (third ? S::R3 : S::R2), LB = hi ? S::LB1 : (third ? S::LB3 : S::LB2);
            const int lg = hi ? (S::KTW - S::LB1 - S::R1) : (third ? (S::KTW - S::LB3 - S::R3) : (S::KTW - S::LB2 - S::R2)); // log2 G
            const int slot = ii >> lg, group = ii & ((1 << lg) - 1);
            // slot -> (ab, x): slot = 2^(R-1-ab) - 1 + x
            const int lvl = 31 - __clz(slot + 1); // = R-1-ab
            const int ab = R - 1 - lvl, x = slot + 1 - (1 << lvl);
            const int rb0 = LB - S::C;
            const int sft = ntw - 1 - lo - (rb0 + ab);
            const int J = j0 | (group << (LB + R - S::C));
            return ((long long) plus << sft) + (J >> (rb0 + ab + 1)) + x;
        };
        // UN table loads are in flight together: a CTA with few tiles (the launch-bound regime) has this loop on its critical
        // path, and one global-memory latency per ENTRY was most of a small call's kernel time
        constexpr int UN = 4;
        for (int i0 = t; i0 < TOTAL; i0 += UN * nthreads)
        {
            T wv[UN];
            int ii[UN];
            bool hi[UN], third[UN];
#pragma unroll
            for (int u = 0; u < UN; u++)
            {
                const int i = i0 + u * nthreads;
                wv[u] = T(0);
                if (i < TOTAL) wv[u] = seg_table[locate(i, ii[u], hi[u], third[u])];
            }
#pragma unroll
            for (int u = 0; u < UN; u++)
            {
                const int i = i0 + u * nthreads;
                if (i >= TOTAL) break;
                Twiddle<T>* dst = hi[u] ? tw1 : (third[u] ? tw3 : tw2);
                if constexpr (sizeof(T) == 8)
                    dst[ii[u]] = Twiddle<T>{wv[u], shoup_companion_mu(wv[u], seg_p, seg_mu, seg_pbits)};
                else
                    dst[ii[u]] = Twiddle<T>{wv[u], shoup_companion_mu32(wv[u], seg_p, seg_mu)};
                if constexpr (S::TW1C > 0)
                {
                    if (scaled && hi[u])
                    {
                        const Mod<T, false> Mx(seg_p);
                        const T wc = csub(Mx.mul(wv[u], ninv), seg_p); // w * n^-1 mod p
                        if constexpr (sizeof(T) == 8)
                            tw3[S::TW3 + ii[u]] = Twiddle<T>{wc, shoup_companion_mu(wc, seg_p, seg_mu, seg_pbits)};
                        else
                            tw3[S::TW3 + ii[u]] = Twiddle<T>{wc, shoup_companion_mu32(wc, seg_p, seg_mu)};
                    }
                }
            }
        }
    }

    // Elementwise sweep over a whole tile buffer (any layout: every byte of the buffer is data).  CENTRE = false: signed input,
    // x < 0 -> x + p (modular_arith.cuh:372-385 of the reference); true: centred output, r > p/2 -> r - p (:389-405).
    template <typename S, bool CENTRE> __device__ __forceinline__ void signed_tile_fixup(unsigned char* buf, typename S::T p, int ctid)
    {
        using T = typename S::T;
        using ST = typename std::make_signed<T>::type;
        constexpr int VN = 16 / (int) sizeof(T);
        const T half = p >> 1;
#pragma unroll 2
        for (int off = ctid * 16; off < S::TILE_SMEM; off += kConsumers * 16)
        {
            T v[VN];
            *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(buf + off);
#pragma unroll
            for (int k = 0; k < VN; k++)
            {
                if constexpr (CENTRE)
                    v[k] = v[k] > half ? v[k] - p : v[k];
                else
                    v[k] = (ST) v[k] < 0 ? v[k] + p : v[k];
            }
            *reinterpret_cast<uint4*>(buf + off) = *reinterpret_cast<const uint4*>(v);
        }
    }

    // ------------------------------------------------------------------ the register rounds of one tile
    // In the merge plans the contiguous pass is always the LAST forward / FIRST inverse pass, so only it canonicalises
    // forward and only a strided pass (the top one) applies n^-1 on the inverse.  SFIN: a forward STRIDED pass that ends
    // the transform (4-step row phase on the transposed layout) canonicalises as well.
    // 4-step twiddle-matrix product: forward = epilogue of the last round, inverse = prologue of the first executed
    // round; either way that is the low round when there are two.
    // LASTC: compile-time knowledge of a.last for inverse passes (-1: read a.last; 0 / 1: never / always the last pass) -- a kernel
    // that never ends a transform then carries no copy of the last-round code (measured: the position-major product kernel lost
    // 12 % with the unused variants compiled in).
    template <typename S, bool WMUL, bool SFIN = false, bool TS = false, bool WL = false, int LASTC = -1>
    __device__ __forceinline__ void tile_rounds(unsigned char* buf, const Twiddle<typename S::T>* tw1, const Twiddle<typename S::T>* tw2,
                                                const Twiddle<typename S::T>* tw3, const typename ModOf<S>::type& M, int tid,
                                                const Twiddle<typename S::T>& ninv, const Twiddle<typename S::T>* wtile,
                                                const FastArgs<typename S::T>& a, bool triv, int bar = 1)
    {
        constexpr bool W1 = WMUL && S::R2 == 0, W2 = WMUL && S::R2 > 0;
        if constexpr (!S::INV)
        {
            // Data32s / Data64s: negative inputs become p - |x| before the first round of the first pass -- a separate sweep
            // over the tile, so that the unsigned path carries no extra instructions
            if (a.signed_io && a.first)
            {
                signed_tile_fixup<S, false>(buf, M.p, tid);
                consumer_sync(bar);
            }
#ifdef GPUNTT_EXPERIMENT_NOCANON // timing experiment only (lazy outputs): what the final canonicalisation costs
            constexpr bool FIN1 = false, FIN2 = false;
#else
            // (a contiguous pass that carries the twiddle-matrix product is not the end of the transform: the product takes any
            // 64-bit value and does the range reduction itself)
            constexpr bool ENDS = SFIN || (!S::STRIDED && !WMUL);
            constexpr bool FIN1 = ENDS && S::R2 == 0, FIN2 = ENDS && S::R3 == 0;
#endif
            constexpr bool TS1 = TS && S::R2 == 0, TS2 = TS && S::R2 > 0;
            if constexpr (S::STRIDED && S::POL == 2 && S::G1 == 1)
            {
                if (triv)
                    fast_round<S, S::R1, S::LB1, S::G1, FIN1, true, W1, TS1, WL>(buf, tw1, M, tid, ninv, wtile, a.lo,
                                                                              a.in_bound > 1 ? a.in_bound : 1, a.w_lazy != 0, bar);
                else
                    fast_round<S, S::R1, S::LB1, S::G1, FIN1, false, W1, TS1, WL>(buf, tw1, M, tid, ninv, wtile, a.lo, 1, a.w_lazy != 0, bar);
            }
            else
                fast_round<S, S::R1, S::LB1, S::G1, FIN1>(buf, tw1, M, tid, ninv, nullptr, 0, 1, false, bar);
            if constexpr (S::R2 > 0)
            {
                consumer_sync(bar);
                fast_round<S, S::R2, S::LB2, S::G2, FIN2, false, W2, TS2, WL>(buf, tw2, M, tid, ninv, wtile, a.lo, 1, a.w_lazy != 0, bar);
            }
            if constexpr (S::R3 > 0)
            {
                consumer_sync(bar);
                fast_round<S, S::R3, S::LB3, S::G3, true>(buf, tw3, M, tid, ninv);
            }
        }
        else
        {
            if constexpr (S::R3 > 0)
            {
                fast_round<S, S::R3, S::LB3, S::G3, false>(buf, tw3, M, tid, ninv);
                consumer_sync(bar);
            }
            if constexpr (S::R2 > 0)
            {
                fast_round<S, S::R2, S::LB2, S::G2, false, false, W2, false, WL>(buf, tw2, M, tid, ninv, wtile, a.lo, 1, false, bar);
                consumer_sync(bar);
            }
            // (triv: X^N-1 and this round is the top of the transform -- its slot-0 twiddles are 1)
            const bool last = LASTC < 0 ? (a.last != 0) : (LASTC != 0);
            if constexpr (S::STRIDED)
            {
                if (last && S::G1 == 1 && triv)
                    fast_round<S, S::R1, S::LB1, S::G1, true, S::G1 == 1, W1, TS, WL>(buf, tw1, M, tid, ninv, wtile, a.lo, 1, false, bar);
                else if (last)
                    fast_round<S, S::R1, S::LB1, S::G1, true, false, W1, TS, WL>(buf, tw1, M, tid, ninv, wtile, a.lo, 1, false, bar);
                else
                    fast_round<S, S::R1, S::LB1, S::G1, false, false, W1, TS, WL>(buf, tw1, M, tid, ninv, wtile, a.lo, 1, false, bar);
            }
            else if constexpr (S::NT > 0)
            {
                // whole transforms in the tile: the top round is the last one of a single-pass inverse (n^-1 there)
                if (last && S::G1 == 1 && triv)
                    fast_round<S, S::R1, S::LB1, S::G1, true, S::G1 == 1>(buf, tw1, M, tid, ninv, nullptr, 0, 1, false, bar);
                else if (last)
                    fast_round<S, S::R1, S::LB1, S::G1, true>(buf, tw1, M, tid, ninv, nullptr, 0, 1, false, bar);
                else
                    fast_round<S, S::R1, S::LB1, S::G1, false>(buf, tw1, M, tid, ninv);
            }
            else
                fast_round<S, S::R1, S::LB1, S::G1, false>(buf, tw1, M, tid, ninv);
            // centred signed output after n^-1 (last pass of the inverse): again a separate sweep
            if (a.signed_io && last)
            {
                consumer_sync(bar);
                signed_tile_fixup<S, true>(buf, M.p, tid);
            }
        }
    }

    // ------------------------------------------------------------------ the persistent pass kernel
    // Work item w of a pass:
    //   STRIDED:   w = poly * 2^(lo-C) + column chunk            (any CTA, any order)
    //   !STRIDED:  w = range * tiles_per_range + polynomial group (range-major, so a CTA's
    //              contiguous share of the work stays inside one or two ranges)
    template <typename T> struct SegConsts // per-segment modulus data of the RNS kernels, in shared memory
    {
        T p, ninv_w, ninv_wq;
        uint64_t mu;
        int pbits, mi;
    };

    template <typename S, bool WMUL, bool RNS, bool SFIN = false, bool TS = false>
    __device__ __forceinline__ void fast_pass_body(const FastArgs<typename S::T>& a, const CUtensorMap& map_in, const CUtensorMap& map_out)
    {
        using T = typename S::T;
        extern __shared__ __align__(128) unsigned char smem_raw[];
        unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
        unsigned char* bufs = smem;                                                    // 2 tile buffers
        Twiddle<T>* tw1 = reinterpret_cast<Twiddle<T>*>(smem + 2 * S::TILE_SMEM);       // high round, slot-major
        Twiddle<T>* tw2 = tw1 + S::TW1;                                                 // low round
        Twiddle<T>* tw3 = tw2 + S::TW2;                                                 // third round (small rings)
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * S::TILE_SMEM + S::TW_SMEM); // full[2], done[2]

        const int tid = threadIdx.x;
        const int n = a.n;
        if constexpr (RNS)
        {
            if (a.policy_flag != nullptr && *a.policy_flag != a.want_policy) return; // the other arithmetic policy's kernel does this pass
        }
        SegConsts<T>* segc = reinterpret_cast<SegConsts<T>*>(bars + 4);
        const int nranges = S::STRIDED ? (1 << (a.n - a.lo - S::D)) : (S::NT ? 1 : (1 << (a.n - S::KC)));
        // Work assignment.  Default: a contiguous share of the work items (a CTA stays inside one or two twiddle ranges).
        // a.rr (strided passes with ONE range and a long row stride): round robin, item(i) = blockIdx + i * grid, with the
        // items ordered column-chunk-major, so the tiles in flight across the chip at any moment are the polynomials
        // of a few ADJACENT column chunks -- whole DRAM pages are consumed together and, for the 4-step column pass,
        // the 16 polynomials of a chunk share one fetch of the twiddle-matrix pairs through the L2.
        // tiles that share one twiddle set ("range" = the index bits above this pass's stage window):
        //   STRIDED: every polynomial x every column chunk of one 2^D-row block;  else: every polynomial group
        const long long tiles_per_range =
            S::STRIDED ? ((long long) a.batch << (a.lo - S::C)) : (long long) ((a.batch + (1 << S::NPLOG) - 1) >> S::NPLOG);
        // a.cta_per_seg: when there are fewer segments than CTA slots, cta_per_seg CTAs split each segment, so nobody
        // builds two twiddle sets for a handful of tiles
        const long long step = a.rr ? (long long) gridDim.x : 1LL;
        const bool pmajor = !a.rr && !a.cc_major; // tile order inside a range: polynomial-major or column-chunk-major
        long long w_begin, w_end;
        if (a.rr)
        {
            w_begin = blockIdx.x;
            w_end = a.work;
        }
        else if (a.cta_per_seg > 0)
        {
            const long long big = (long long) a.seg_extra * (a.cta_per_seg + 1); // CTAs of the segments with one more
            long long sg, j, kk;
            if ((long long) blockIdx.x < big)
            {
                kk = a.cta_per_seg + 1;
                sg = blockIdx.x / kk;
                j = blockIdx.x % kk;
            }
            else
            {
                kk = a.cta_per_seg;
                sg = a.seg_extra + ((long long) blockIdx.x - big) / kk;
                j = ((long long) blockIdx.x - big) % kk;
            }
            w_begin = sg * tiles_per_range + tiles_per_range * j / kk;
            w_end = sg * tiles_per_range + tiles_per_range * (j + 1) / kk;
        }
        else
        {
            w_begin = a.work * blockIdx.x / gridDim.x;
            w_end = a.work * (blockIdx.x + 1) / gridDim.x;
        }

        if (tid == kConsumers)
        {
            tma_prefetch_desc(&map_in);
            tma_prefetch_desc(&map_out);
        }
        if (tid == 0)
        {
            mbar_init(smem_u32(&bars[0]), 1);
            mbar_init(smem_u32(&bars[1]), 1);
            mbar_init(smem_u32(&bars[2]), kConsumers);
            mbar_init(smem_u32(&bars[3]), kConsumers);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fence_async();
        }
        typename ModOf<S>::type M(a.p);
        Twiddle<T> ninv{a.ninv_w, a.ninv_wq};
        const T* seg_table = a.table;
        T seg_p = a.p;
        uint64_t seg_mu = a.mu;
        int seg_pbits = a.pbits;
        // first pass of a cyclic transform: slot 0 of every stage of the high round is table[0]; when that is 1
        // (it is omega^0 in the reference's tables) those butterflies need no multiply
        // (inverse: the LAST round of a cyclic transform -- the top round of the top strided pass, or of a whole-transform tile)
        bool triv = S::INV ? ((S::STRIDED ? (a.lo + S::D == a.n) : S::NT > 0) && !a.plus && a.last && a.table[0] == T(1))
                           : (S::STRIDED && !a.plus && a.first && (a.lo + S::D == a.n) && a.table[0] == T(1)); // inputs canonical by contract
        uint32_t uses0 = 0, uses1 = 0; // how often each buffer has been filled so far (phase tracking)

        long long w = w_begin;
        while (w < w_end)
        {
            // ---- segment: a run of tiles sharing one twiddle set
            long long seg_end = w_end;
            int range = 0, mslot = 0;
            {
                const long long sg = w / tiles_per_range; // RNS: (modulus slot, range), slot-major
                const long long re = (sg + 1) * tiles_per_range;
                if (re < seg_end) seg_end = re;
                if constexpr (RNS)
                {
                    mslot = (int) (sg / nranges);
                    range = (int) (sg % nranges);
                }
                else
                    range = (int) sg;
            }
            const int ntiles = (int) ((seg_end - w + step - 1) / step);

            __syncthreads(); // everybody is done with the previous segment's twiddles and buffers
            const int lane = tid - kConsumers;
            // TMA coordinates of work item ww (innermost first)
            auto issue_load = [&](long long ww, int b)
            {
                if (lane == 0)
                {
                    const uint32_t bar = smem_u32(&bars[b]);
                    const uint32_t dst = smem_u32(bufs + b * S::TILE_SMEM);
                    mbar_expect_tx(bar, S::TILE_SMEM); // out-of-range rows are zero-filled and still counted
                    if constexpr (S::STRIDED)
                    {
                        const int ccb = a.lo - S::C;
                        const long long within = ww % tiles_per_range;
                        const long long poly = pmajor ? within >> ccb : within % a.batch;
                        const long long cc = pmajor ? (within & ((1LL << ccb) - 1)) : within / a.batch;
                        long long gp = RNS ? poly * a.mod_count + mslot : poly; // polynomial in the caller's array
                        if constexpr (RNS)
                            if (a.poly_order) gp = a.poly_order[gp];
                        tma_load_3d(dst, &map_in, 0, (int) (cc << (S::C - S::CB)),
                                    (int) ((gp << (a.n - a.lo)) + ((long long) range << S::D)), bar);
                    }
                    else
                    {
                        const long long grp = ww % tiles_per_range;
                        if constexpr (RNS) // {row, rows of a polynomial, modulus slot, polynomial within the slot}
                        {
                            if (a.poly_order)
                            {
                                // the tile's polynomials sit in arbitrary slots: one box per polynomial (3-D map, box of one
                                // polynomial); a missing last polynomial is an out-of-range coordinate (zero fill)
#pragma unroll
                                for (int q = 0; q < (1 << S::NPLOG); q++)
                                {
                                    const long long kq = (grp << S::NPLOG) + q;
                                    const int slot = kq < a.batch ? a.poly_order[kq * a.mod_count + mslot] : 0x7fffffff;
                                    tma_load_3d(dst + q * (S::TILE_SMEM >> S::NPLOG), &map_in, 0, range << (S::KC - S::CB), slot, bar);
                                }
                            }
                            else
                                tma_load_4d(dst, &map_in, 0, range << (S::KC - S::CB), mslot, (int) (grp << S::NPLOG), bar);
                        }
                        else
                            tma_load_3d(dst, &map_in, 0, range << (S::KC - S::CB), (int) (grp << S::NPLOG), bar);
                    }
                }
            };
            // the first two tiles of the segment start moving now, under the twiddle build
            if (tid >= kConsumers)
                for (int i = 0; i < 2 && i < ntiles; i++) issue_load(w + i * step, (int) ((uses0 + uses1 + i) & 1));
            if constexpr (RNS)
            {
                if (tid == 0)
                {
                    SegConsts<T> c;
                    const int mi = a.mod_order ? a.mod_order[mslot] : mslot;
                    c.mi = mi;
                    c.p = a.mod_dev[3 * mi];
                    if constexpr (sizeof(T) == 8)
                    {
                        c.pbits = 64 - __clzll((long long) c.p);
                        c.mu = (c.p & (c.p - 1)) ? recip_mu64(c.p, c.pbits) : ~0ull; // (a power of two is no modulus; keep the old clamp)
                    }
                    else
                    {
                        c.pbits = 32 - __clz((int) c.p);
                        c.mu = ~0ull / (uint64_t) c.p;
                    }
                    c.ninv_w = S::INV ? a.ninv_dev[mi] : T(0);
                    if constexpr (sizeof(T) == 8)
                        c.ninv_wq = S::INV ? shoup_companion_mu(c.ninv_w, c.p, c.mu, c.pbits) : T(0);
                    else
                        c.ninv_wq = S::INV ? shoup_companion(c.ninv_w, c.p) : T(0);
                    *segc = c;
                }
                __syncthreads();
                seg_p = segc->p;
                seg_mu = segc->mu;
                seg_pbits = segc->pbits;
                ninv = Twiddle<T>{segc->ninv_w, segc->ninv_wq};
                seg_table = a.table + ((size_t) segc->mi << a.n);
                M = typename ModOf<S>::type(seg_p);
                triv = S::INV ? ((S::STRIDED ? (a.lo + S::D == a.n) : S::NT > 0) && !a.plus && a.last && seg_table[0] == T(1))
                              : (S::STRIDED && !a.plus && a.first && (a.lo + S::D == a.n) && seg_table[0] == T(1));
            }
            build_twiddles<S>(tw1, seg_table, a.tw_fixed ? 0 : range, n, a.n_tw, a.lo, a.plus, seg_p, seg_mu, seg_pbits, tid, kFastThreads,
                              S::INV && a.last, ninv);
            __syncthreads();

            if (tid >= kConsumers)
            {
                // =================== producer warp ===================
                // tile t of the CTA's whole stream uses buffer (t & 1); uses0 + uses1 = tiles so far
                for (int i = 0; i < ntiles; i++)
                {
                    const uint32_t t = uses0 + uses1;
                    const int b = t & 1;
                    const uint32_t k = b ? uses1 : uses0;
                    mbar_wait(smem_u32(&bars[2 + b]), k & 1); // consumers finished this tile
                    if (lane == 0)
                    {
                        const uint32_t src = smem_u32(bufs + b * S::TILE_SMEM);
                        const long long ww = w + i * step;
                        if constexpr (S::STRIDED)
                        {
                            const int ccb = a.lo - S::C;
                            const long long within = ww % tiles_per_range;
                            const long long poly = pmajor ? within >> ccb : within % a.batch;
                            const long long cc = pmajor ? (within & ((1LL << ccb) - 1)) : within / a.batch;
                            long long gp = RNS ? poly * a.mod_count + mslot : poly;
                            if constexpr (RNS)
                                if (a.poly_order) gp = a.poly_order[gp];
                            if constexpr (TS) // transposing box {16 rows, 2^C columns, 2^(D-4) row blocks} of the transposed output matrix
                                tma_store_3d(&map_out, 0, (int) ((gp << a.lo) + (cc << S::C)), range << (S::D - 4), src);
                            else
                                tma_store_3d(&map_out, 0, (int) (cc << (S::C - S::CB)),
                                             (int) ((gp << (a.n - a.lo)) + ((long long) range << S::D)), src);
                        }
                        else
                        {
                            const long long grp = ww % tiles_per_range;
                            if constexpr (RNS)
                            {
                                if (a.poly_order)
                                {
#pragma unroll
                                    for (int q = 0; q < (1 << S::NPLOG); q++)
                                    {
                                        const long long kq = (grp << S::NPLOG) + q;
                                        if (kq < a.batch)
                                            tma_store_3d(&map_out, 0, range << (S::KC - S::CB), a.poly_order[kq * a.mod_count + mslot],
                                                         src + q * (S::TILE_SMEM >> S::NPLOG));
                                    }
                                }
                                else
                                    tma_store_4d(&map_out, 0, range << (S::KC - S::CB), mslot, (int) (grp << S::NPLOG), src);
                            }
                            else
                                tma_store_3d(&map_out, 0, range << (S::KC - S::CB), (int) (grp << S::NPLOG), src);
                        }
                        bulk_commit();
                        bulk_wait_read0(); // this buffer may be overwritten again
                    }
                    __syncwarp();
                    if (b) uses1++; else uses0++;
                    if (i + 2 < ntiles) issue_load(w + (i + 2) * step, b);
                }
                if (lane == 0) bulk_wait0();
            }
            else
            {
                // =================== consumer warps ===================
                for (int i = 0; i < ntiles; i++)
                {
                    const uint32_t t = uses0 + uses1;
                    const int b = t & 1;
                    const uint32_t k = b ? uses1 : uses0;
                    unsigned char* buf = bufs + b * S::TILE_SMEM;
                    mbar_wait(smem_u32(&bars[b]), k & 1); // tile landed
                    // In this path the contiguous pass is always the LAST forward / FIRST inverse pass, so only it
                    // canonicalises forward and only a strided pass (the top one) applies n^-1 on the inverse.
                    // 4-step twiddle-matrix product: forward = epilogue of the last round, inverse = prologue of the first
                    // executed round; either way that is the low round when there are two.
                    const Twiddle<T>* wtile = nullptr;
                    if constexpr (WMUL)
                    {
                        // pair of the tile's first element: row block `range`, column chunk cc (same for every polynomial)
                        const long long within = (w + i * step) % tiles_per_range;
                        const long long cc = pmajor ? (within & ((1LL << (a.lo - S::C)) - 1)) : within / a.batch;
                        wtile = reinterpret_cast<const Twiddle<T>*>(a.w_pairs) + ((((long long) range << S::D)) << a.lo) + (cc << S::C);
                        // pull this thread's pairs towards the SM now
                        constexpr int RW = S::R2 > 0 ? S::R2 : S::R1, LBW = S::R2 > 0 ? S::LB2 : S::LB1;
#pragma unroll 1
                        for (int item = tid; item < ((1 << S::K) >> RW); item += kConsumers)
                        {
                            const int l_base = ((item >> LBW) << (LBW + RW)) | (item & ((1 << LBW) - 1));
                            const Twiddle<T>* wp = wtile + (((long long) (l_base >> S::C)) << a.lo) + (l_base & ((1 << S::C) - 1));
#pragma unroll
                            for (int x = 0; x < (1 << RW); x++)
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + (((long long) x << (LBW - S::C)) << a.lo)));
                        }
                    }
                    tile_rounds<S, WMUL, SFIN, TS>(buf, tw1, tw2, tw3, M, tid, ninv, wtile, a, triv);
                    fence_async(); // make the generic-proxy writes visible to the bulk store
                    mbar_arrive(smem_u32(&bars[2 + b]));
                    if (b) uses1++; else uses0++;
                }
            }
            // the producer's counters advance identically
            if (tid >= kConsumers)
            {
                // (already advanced inside the loop)
            }
            w = seg_end;
        }
    }

    template <typename S, bool WMUL = false, bool RNS = false, bool SFIN = false, bool TS = false>
    __global__ void __launch_bounds__(kFastThreads, (!S::STRIDED && S::NPLOG == 0 && S::NT == 0)
                                                        ? GPUNTT_FAST_SP_BLOCKS
                                                        : ((!S::INV && S::POL == 2) ? (S::STRIDED ? GPUNTT_FAST_P0_BLOCKS : GPUNTT_FAST_P1_BLOCKS) : 2))
        fast_pass_kernel(const FastArgs<typename S::T> a, const __grid_constant__ CUtensorMap map_in,
                         const __grid_constant__ CUtensorMap map_out)
    {
        fast_pass_body<S, WMUL, RNS, SFIN, TS>(a, map_in, map_out);
    }

    // RNS calls on 64-bit data: the moduli live on the device, so the choice between the lazy-policy body (SL) and the
    // exact-policy body (SX) of a pass is made HERE, by every warp, from the modulus array itself -- one launch per pass
    // and no pre-kernel (a call used to be a flag kernel plus two launches per pass, one of which returned at once:
    // 5 launches, 32-34 us for a small batch against 18-23 us on the reference's two kernels).
    template <typename SL, typename SX>
    __global__ void __launch_bounds__(kFastThreads, 2)
        fast_pass_dual_kernel(const FastArgs<uint64_t> a, const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out)
    {
        static_assert(SL::INV == SX::INV && SL::STRIDED == SX::STRIDED && SL::SMEM == SX::SMEM && SX::POL == 0 && SL::POL != 0, "same pass, two policies");
        int bad = 0;
        for (int i = threadIdx.x & 31; i < a.mod_count; i += 32)
        {
            const uint64_t p = a.mod_dev[3 * (a.mod_order ? a.mod_order[i] : i)];
            const bool ok = SL::INV ? (p < kFastModulusLimit) : (p >= kF60ModulusMin && p < kF60ModulusLimit);
            bad |= ok ? 0 : 1;
        }
        if (__any_sync(0xffffffffu, bad))
            fast_pass_body<SX, false, true>(a, map_in, map_out);
        else
            fast_pass_body<SL, false, true>(a, map_in, map_out);
    }

    // ------------------------------------------------------------------ host side
    // (function-local static with an initialiser: thread-safe by the language rules)
    static PFN_cuTensorMapEncodeTiled get_encode()
    {
        static const PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled
        {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
            return nullptr;
        }();
        return fn;
    }

    // Tensor map of the [batch][N] array for one pass shape.
    //   STRIDED: 2-D view {2^lo (adjacent elements of a matrix row), batch * 2^D rows}; box {2^C, 2^D}
    //   else:    3-D view {2^CB (one 128-byte row), N / 2^CB rows, batch}; box {2^CB, 2^(KC-CB), 2^NPLOG}
    //   RNS (mod_count > 0; batch = polynomials per slot): strided maps see batch * mod_count polynomials, contiguous
    //   maps are 4-D {row, rows, slot, polynomial within the slot} so a tile holds polynomials of ONE modulus.
    template <typename S> static bool make_map(CUtensorMap* map, const void* base, int n, int lo, int batch, int mod_count = 0, bool one_poly_box = false)
    {
        using T = typename S::T;
        // A tensor map is a pure function of these arguments (an address and a shape, nothing about the memory behind it), and a
        // caller in the launch-bound regime passes the same few buffers over and over: remember the last encodes of this shape
        // on this thread (cuTensorMapEncodeTiled is a microsecond-class driver call, three of them per single-launch transform).
        struct Cached
        {
            const void* base;
            int n, lo, batch, mod_count, opb;
            bool valid;
            CUtensorMap map;
        };
        thread_local static Cached cache[4];
        thread_local static unsigned next_slot = 0;
        for (const Cached& c : cache)
            if (c.valid && c.base == base && c.n == n && c.lo == lo && c.batch == batch && c.mod_count == mod_count && c.opb == (int) one_poly_box)
            {
                *map = c.map;
                return true;
            }
        PFN_cuTensorMapEncodeTiled enc = get_encode();
        if (!enc) return false;
        const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
        cuuint64_t gdim[4], gstride[3];
        cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
        int rank;
        if constexpr (S::STRIDED)
        {
            // {one 128-byte row, column blocks of a matrix row, all matrix rows of all polynomials}
            rank = 3;
            gdim[0] = 1ull << S::CB;
            gdim[1] = 1ull << (lo - S::CB);
            gdim[2] = ((cuuint64_t) batch * (mod_count > 0 ? mod_count : 1)) << (n - lo);
            if (one_poly_box) gdim[2] = 1ull << 31; // Poly_Ordered: the slots may lie anywhere in a larger array of unknown size
            gstride[0] = 128;
            gstride[1] = (cuuint64_t) sizeof(T) << lo;
            box[0] = 1u << S::CB;
            box[1] = 1u << (S::C - S::CB);
            box[2] = 1u << S::D;
        }
        else
        {
            rank = 3;
            gdim[0] = 1ull << S::CB;
            gdim[1] = 1ull << (n - S::CB);
            gstride[0] = 128;
            gstride[1] = (cuuint64_t) sizeof(T) << n;
            box[0] = 1u << S::CB;
            box[1] = 1u << (S::KC - S::CB);
            if (mod_count > 0 && one_poly_box)
            {
                // Poly_Ordered: plain 3-D view of every polynomial, one polynomial per box
                gdim[2] = 1ull << 30; // slots may lie anywhere in a larger array; 0x7fffffff stays out of range (ragged last tile)
                box[2] = 1;
            }
            else if (mod_count > 0)
            {
                rank = 4;
                gdim[2] = (cuuint64_t) mod_count;
                gdim[3] = (cuuint64_t) batch;
                gstride[2] = ((cuuint64_t) sizeof(T) << n) * (cuuint64_t) mod_count;
                box[2] = 1;
                box[3] = 1u << S::NPLOG;
            }
            else
            {
                gdim[2] = (cuuint64_t) batch;
                box[2] = 1u << S::NPLOG;
            }
        }
        CUresult r = enc(map, dt, (cuuint32_t) rank, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
        Cached& c = cache[next_slot++ & 3];
        c.base = base;
        c.n = n;
        c.lo = lo;
        c.batch = batch;
        c.mod_count = mod_count;
        c.opb = (int) one_poly_box;
        c.map = *map;
        c.valid = true;
        return true;
    }

    // Output map of a transposing strided pass (fast_round TS): the pass reads the [rows = 2^D][2^lo] matrix of every
    // polynomial (one range: lo + D == n) and writes its transpose, 2^lo rows of 2^D contiguous elements.  View
    // {16 elements of a transposed row, every transposed row of every polynomial, 2^(D-4) blocks of 16 elements along
    // that row}; box {16, 2^C, 2^(D-4)}.
    // n > lo + D (several twiddle ranges: the pass works on 2^D of the 2^(n - lo) matrix rows at a time): the transposed rows are
    // 2^(n - lo) elements long and range r fills elements [r * 2^D, (r + 1) * 2^D) of each -- third coordinate r * 2^(D - 4).
    template <typename S> static bool make_map_tstore(CUtensorMap* map, const void* base, int lo, int batch, int n)
    {
        using T = typename S::T;
        static_assert(sizeof(T) == 8 && S::STRIDED && S::D >= 5, "transposing store: 64-bit strided passes of 5..8 stages");
        PFN_cuTensorMapEncodeTiled enc = get_encode();
        if (!enc) return false;
        cuuint64_t gdim[3] = {16, (cuuint64_t) batch << lo, 1ull << (n - lo - 4)};
        cuuint64_t gstride[2] = {(cuuint64_t) sizeof(T) << (n - lo), 128};
        cuuint32_t box[3] = {16, 1u << S::C, 1u << (S::D - 4)}, estr[3] = {1, 1, 1};
        CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS;
    }

    // returns cudaErrorNotSupported when the tensor maps cannot be built (caller falls back)
    // SX: exact-policy twin of S for the 64-bit RNS calls (fast_pass_dual_kernel picks on the device)
    // SFIN: forward strided pass that ends the transform (canonical outputs); TS: transposing store (in != out)
    template <typename S, bool WMUL = false, bool RNS = false, typename SX = void, bool SFIN = false, bool TS = false>
    static cudaError_t launch_fast(const FastArgs<typename S::T>& args, cudaStream_t st)
    {
        // per device (the shared-memory opt-in is a per-device function attribute); a race between first callers only
        // repeats idempotent work
        constexpr int kMaxDev = 64;
        static std::atomic<int> cached_bps[kMaxDev];
        static std::atomic<int> cached_sms[kMaxDev];
        using KernT = void (*)(const FastArgs<typename S::T>, const CUtensorMap, const CUtensorMap);
        KernT kern;
        if constexpr (std::is_void<SX>::value)
            kern = fast_pass_kernel<S, WMUL, RNS, SFIN, TS>;
        else
            kern = fast_pass_dual_kernel<S, SX>;
        int dev = 0;
        cudaError_t ge = cudaGetDevice(&dev);
        if (ge != cudaSuccess) return ge;
        if (dev < 0 || dev >= kMaxDev) return cudaErrorNotSupported;
        int blocks_per_sm = cached_bps[dev].load(std::memory_order_acquire), sms = cached_sms[dev].load(std::memory_order_acquire);
        if (blocks_per_sm <= 0 || sms <= 0)
        {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM);
            if (e != cudaSuccess) return e;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int bps = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kFastThreads, S::SMEM);
            if (e != cudaSuccess) return e;
            blocks_per_sm = bps > 0 ? bps : 1;
            cached_sms[dev].store(sms, std::memory_order_release);
            cached_bps[dev].store(blocks_per_sm, std::memory_order_release);
        }
        alignas(64) CUtensorMap map_in, map_out;
        const int mc = RNS ? args.mod_count : 0;
        if constexpr (S::STRIDED)
        {
            // TMA coordinates are signed 32-bit: the row index of the last polynomial must fit
            const long long rows = ((long long) args.batch * (mc > 0 ? mc : 1)) << (args.n - args.lo);
            if (rows >= (1LL << 31)) return cudaErrorNotSupported;
        }
        bool opb = false;
        if constexpr (RNS) opb = args.poly_order != nullptr;
        if (!make_map<S>(&map_in, args.in, args.n, args.lo, args.batch, mc, opb)) return cudaErrorNotSupported;
        if constexpr (TS)
        {
            if (args.in == args.out || args.lo + S::D > args.n) return cudaErrorNotSupported;
            if (!make_map_tstore<S>(&map_out, args.out, args.lo, args.batch, args.n)) return cudaErrorNotSupported;
        }
        else if (args.in == args.out)
            map_out = map_in;
        else if (!make_map<S>(&map_out, args.out, args.n, args.lo, args.batch, mc, opb))
            return cudaErrorNotSupported;
        long long grid = (long long) sms * blocks_per_sm;
        if (grid > args.work) grid = args.work;
        FastArgs<typename S::T> la = args;
        la.cta_per_seg = 0;
        la.seg_extra = 0;
        if (!args.rr)
        {
            // segment-aligned shares when the segments are fewer than the CTA slots
            const long long tpr = S::STRIDED ? ((long long) args.batch << (args.lo - S::C)) : (long long) ((args.batch + (1 << S::NPLOG) - 1) >> S::NPLOG);
            const long long nseg = tpr > 0 ? args.work / tpr : 0;
            if (nseg > 0 && nseg * tpr == args.work && nseg <= grid)
            {
                long long k = grid / nseg, extra = grid % nseg;
                if (k >= tpr)
                {
                    k = tpr; // one tile per CTA
                    extra = 0;
                }
                // cost in tile times, a twiddle build counted as half a tile: aligned shares are uneven (the slowest CTA
                // has ceil(tpr / k) tiles) but build once; contiguous shares are even but usually straddle two segments
                const double aligned = (double) ((tpr + k - 1) / k) + 0.5;
                const double contiguous = (double) ((args.work + grid - 1) / grid) + 1.0;
                if (aligned < contiguous)
                {
                    la.cta_per_seg = (int) k;
                    la.seg_extra = (int) extra;
                    grid = nseg * k + extra;
                }
            }
        }
        kern<<<(unsigned) grid, kFastThreads, S::SMEM, st>>>(la, map_in, map_out);
        return cudaGetLastError();
    }

    bool fast_supported(int n_power, int element_bits); // merge_fast.cu
    struct FastPlan;
    // merge_fused.cu: a two-pass plan in one launch (cudaErrorNotSupported: not covered, launch the passes separately)
    template <typename T>
    cudaError_t fused_merge(const FastArgs<T>& a, const FastPlan& pl, bool inverse, bool f60_or_l32, bool lazy_inv, unsigned* counters,
                            cudaStream_t st, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    template <typename T>
    cudaError_t fused_merge_rns(const FastArgs<T>& a, const FastPlan& pl, bool inverse, unsigned* counters, cudaStream_t st,
                                void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t));
    void fused_set_lag_steps(int v);
    void fused_set_small_tile_elems(long long v); // 64-bit calls of at most this many elements run on 1024-element tiles
    void fused_set_policy(int v); // 1: where measured faster, 2: wherever the shapes allow

    struct FastPlan
    {
        int npass = 0;     // forward order
        int d[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
        bool strided[4] = {false, false, false, false};
    };
    // (an 8-stage contiguous pass for the small 32-bit rings -- a "balanced" split -- measured no faster than 4 + 10:
    // profiles/r1_ab_experiments.txt)
    // The contiguous pass takes the low dc stages; the stages above it go to one, two or three strided passes of at most 8
    // (replaces the reference's two- and three-kernel plans, ntt.cuh:628-697; three strided passes: rings of 2^25 .. 2^28).
    inline FastPlan make_fast_plan(int n, int element_bits)
    {
        FastPlan pl;
        const int dc = element_bits == 64 ? 8 : 10, rest = n - dc;
        const int ns = rest <= 8 ? 1 : (rest <= 16 ? 2 : 3); // strided passes
        pl.npass = ns + 1;
        int left = rest, top = n;
        for (int i = 0; i < ns; i++)
        {
            pl.d[i] = (left + (ns - i) - 1) / (ns - i); // the larger shares first
            left -= pl.d[i];
            top -= pl.d[i];
            pl.lo[i] = top;
            pl.strided[i] = true;
        }
        pl.d[ns] = dc;
        pl.lo[ns] = 0;
        return pl;
    }

    // strided pass of D stages: rounds (D, 0) up to 4 stages, else (ceil(D/2), floor(D/2))
    template <bool INV, int POL = 0> static cudaError_t launch_strided32(int d, const FastArgs<uint32_t>& args, cudaStream_t st)
    {
        using T = uint32_t;
        switch (d)
        {
            case 3: return launch_fast<Shape<T, INV, POL, true, 3, 0, 13, 0>>(args, st);
            case 4: return launch_fast<Shape<T, INV, POL, true, 4, 0, 13, 0>>(args, st);
            case 5: return launch_fast<Shape<T, INV, POL, true, 5, 0, 13, 0>>(args, st);
            case 6: return launch_fast<Shape<T, INV, POL, true, 3, 3, 13, 0>>(args, st);
            case 7: return launch_fast<Shape<T, INV, POL, true, 4, 3, 13, 0>>(args, st);
            case 8: return launch_fast<Shape<T, INV, POL, true, 4, 4, 13, 0>>(args, st);
            default: return cudaErrorNotSupported;
        }
    }
    template <typename T, bool INV, int POL> static cudaError_t launch_strided(int d, const FastArgs<T>& args, cudaStream_t st)
    {
        switch (d)
        {
            case 4: return launch_fast<Shape<T, INV, POL, true, 4, 0, 12, 0>>(args, st);
            case 5: return launch_fast<Shape<T, INV, POL, true, 3, 2, 12, 0>>(args, st);
            case 6: return launch_fast<Shape<T, INV, POL, true, 3, 3, 12, 0>>(args, st);
            case 7: return launch_fast<Shape<T, INV, POL, true, 4, 3, 12, 0>>(args, st);
            case 8: return launch_fast<Shape<T, INV, POL, true, 4, 4, 12, 0>>(args, st);
            default: return cudaErrorNotSupported;
        }
    }


} // namespace gpuntt_b200
