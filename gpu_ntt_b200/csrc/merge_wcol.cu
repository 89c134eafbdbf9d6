// gpu_ntt_b200/csrc/merge_wcol.cu -- the 4-step passes that carry the twiddle-matrix product, with the (W, W') pairs of a tile
// position resident in shared memory.
//
// Every 4-step transform has ONE pass whose tiles are multiplied elementwise by the N-entry twiddle matrix (the reference
// fuses that product into its first row kernel, ntt_4step.cu:745-880; FourStepForwardCoreT4 :571-743 feeds it):
//   forward, fused contract      strided pass over the top log2(n1) index bits of the natural n1 x n2 matrix, product as epilogue,
//                                transposing store (fast_round TS) -> the n2 x n1 matrix
//   forward, reference contract  contiguous pass over the rows of the n2 x n1 matrix the caller's GPU_Transpose made (every
//                                column transform is a contiguous run there), product as epilogue, plain store
//   inverse, either contract     strided Gentleman-Sande pass over the low stages of the size-n2 transforms, product as
//                                prologue; reference contract: transposing store -> the n1 x n2 matrix
// The first version of the forward pass (fast_pass_kernel<..., WMUL>) fetched the pairs of a tile with __ldg behind an L1
// prefetch: 64 KiB of pairs per 32 KiB tile, more than the L1 that is left beside two CTAs' tile buffers, and the ncu
// capture shows the consumers stalled on those loads (profiles/r2_v1_ncu_summary.txt: long_scoreboard the top stall, 57 %
// multiplier-pipe utilisation against 70 % for the row passes; the inverse pass, polynomial-major, re-read the whole pair
// table from DRAM for every polynomial: 6.2 GB of reads for 2.1 GB of data, profiles/r2_v4_kernel_families_ncu.txt).
// This kernel turns the loop around: a CTA owns tile POSITIONS and walks every polynomial of the batch through one position
// before it moves on, so the position's 64 KiB of pairs are loaded ONCE by the TMA engine into shared memory and reused
// batch_size times; the product reads them with LDS.128 (conflict-free: a quarter warp reads 128 contiguous bytes).
// One CTA per SM: two consumer groups of 8 warps, a loader thread, a storer thread, four data-tile buffers (same skeleton
// as merge_fused.cu).  Passes with several twiddle ranges (the inverse) give every CTA a contiguous block of positions and
// keep two twiddle sets: the group that claims the first tile of a new range builds the set the range before last used.
#include "fast_kernels.cuh"

namespace gpuntt_b200
{

    constexpr int kWcolGroups = 2;
    constexpr int kWcolConsumers = kWcolGroups * kConsumers;
    constexpr int kWcolThreads = kWcolConsumers + 64;
    constexpr int kWcolBufs = 4;

    struct WcolCtl
    {
        uint64_t full[kWcolBufs], done[kWcolBufs], free_[kWcolBufs];
        uint64_t pairs_full, pairs_free;
        uint64_t tw_ready[2];
        int next_t;
        int bcast[kWcolGroups][2];
    };

    template <typename S> struct WcolSmem
    {
        static constexpr int TILE = S::TILE_SMEM;
        static constexpr int PAIRS = (1 << S::K) * (int) sizeof(Twiddle<typename S::T>);
        static constexpr int BYTES = kWcolBufs * TILE + PAIRS + 2 * S::TW_SMEM + (int) sizeof(WcolCtl) + 1024;
    };

    // a.n: log2 of the polynomial length; strided shapes: a.lo = row stride, positions = (range, column chunk); contiguous
    // shapes (whole transforms of 2^NT inside a tile, a.lo = 0, a.n_tw = NT): positions = runs of 2^K elements.
    template <typename S, bool TS>
    __global__ void __launch_bounds__(kWcolThreads, 1)
        wcol_kernel(const FastArgs<typename S::T> a, const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                    const __grid_constant__ CUtensorMap map_pairs)
    {
        using T = typename S::T;
        static_assert(sizeof(T) == 8 && S::POL != 0, "64-bit lazy-policy passes");
        static_assert(S::STRIDED || (S::NT > 0 && S::NPLOG == 1 && !TS), "contiguous form: whole transforms inside a tile of two chunks");
        constexpr int TILE = S::TILE_SMEM, NB = kWcolBufs;
        constexpr int TWN = S::TWN;
        extern __shared__ __align__(128) unsigned char smem_raw[];
        unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
        unsigned char* bufs = smem;
        Twiddle<T>* pairs = reinterpret_cast<Twiddle<T>*>(smem + NB * TILE);
        Twiddle<T>* tw = reinterpret_cast<Twiddle<T>*>(smem + NB * TILE + WcolSmem<S>::PAIRS); // two sets of TWN entries
        WcolCtl* ctl = reinterpret_cast<WcolCtl*>(smem + NB * TILE + WcolSmem<S>::PAIRS + 2 * S::TW_SMEM);

        const int tid = threadIdx.x;
        const int batch = a.batch, n = a.n, lo = a.lo;
        const int ncc = S::STRIDED ? (1 << (lo - S::C)) : 1;                     // column chunks of a matrix row block
        const int nranges = S::STRIDED ? (1 << (n - lo - S::D)) : 1;              // twiddle ranges (blocks of 2^D matrix rows)
        const int npos_all = S::STRIDED ? ncc * nranges : (1 << (n - S::K));
        // one range: positions interleaved over the CTAs (neighbouring CTAs read neighbouring 128-byte columns of the same DRAM
        // pages); several: a contiguous block each, so a CTA sees few ranges
        const bool blocked = nranges > 1;
        const int pos0 = blocked ? (int) ((long long) npos_all * blockIdx.x / gridDim.x) : (int) blockIdx.x;
        const int pos_step = blocked ? 1 : (int) gridDim.x;
        const int npos = blocked ? (int) ((long long) npos_all * (blockIdx.x + 1) / gridDim.x) - pos0
                                 : (((int) blockIdx.x < npos_all) ? (npos_all - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0);
        const int total = npos * batch;
        const bool per_range_tw = S::STRIDED && !a.tw_fixed;
        const int range0 = pos0 / ncc;

        if (tid == kWcolConsumers)
        {
            tma_prefetch_desc(&map_in);
            tma_prefetch_desc(&map_pairs);
        }
        if (tid == kWcolConsumers + 32) tma_prefetch_desc(&map_out);
        if (tid == 0)
        {
            for (int b = 0; b < NB; b++)
            {
                mbar_init(smem_u32(&ctl->full[b]), 1);
                mbar_init(smem_u32(&ctl->done[b]), kConsumers);
                mbar_init(smem_u32(&ctl->free_[b]), 1);
            }
            mbar_init(smem_u32(&ctl->pairs_full), 1);
            mbar_init(smem_u32(&ctl->pairs_free), 1);
            mbar_init(smem_u32(&ctl->tw_ready[0]), kConsumers);
            mbar_init(smem_u32(&ctl->tw_ready[1]), kConsumers);
            ctl->next_t = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fence_async();
        }
        __syncthreads();

        if (tid == kWcolConsumers)
        {
            // =================== loader ===================
            for (int k = 0; k < npos; k++)
            {
                const int pos = pos0 + k * pos_step;
                const int range = pos / ncc, cc = pos % ncc;
                if (k > 0) mbar_wait(smem_u32(&ctl->pairs_free), (unsigned) (k - 1) & 1u); // every tile of the previous position is stored
                mbar_expect_tx(smem_u32(&ctl->pairs_full), WcolSmem<S>::PAIRS);
                if constexpr (S::STRIDED)
                    tma_load_2d(smem_u32(pairs), &map_pairs, cc << (S::C + 1), range << S::D, smem_u32(&ctl->pairs_full));
                else
                    tma_load_2d(smem_u32(pairs), &map_pairs, 0, pos << (S::K + 1 - 8), smem_u32(&ctl->pairs_full));
                for (int poly = 0; poly < batch; poly++)
                {
                    const int t = k * batch + poly, b = t % NB;
                    if (t >= NB) mbar_wait(smem_u32(&ctl->free_[b]), (unsigned) (t / NB - 1) & 1u);
                    mbar_expect_tx(smem_u32(&ctl->full[b]), TILE);
                    if constexpr (S::STRIDED)
                        tma_load_3d(smem_u32(bufs + b * TILE), &map_in, 0, cc << (S::C - S::CB),
                                    (int) (((long long) poly << (n - lo)) + ((long long) range << S::D)), smem_u32(&ctl->full[b]));
                    else
                        tma_load_3d(smem_u32(bufs + b * TILE), &map_in, 0, 0,
                                    (int) (((long long) poly << (n - S::KC)) + ((long long) pos << S::NPLOG)), smem_u32(&ctl->full[b]));
                }
            }
        }
        else if (tid == kWcolConsumers + 32)
        {
            // =================== storer ===================
            for (int k = 0; k < npos; k++)
            {
                const int pos = pos0 + k * pos_step;
                const int range = pos / ncc, cc = pos % ncc;
                for (int poly = 0; poly < batch; poly++)
                {
                    const int t = k * batch + poly, b = t % NB;
                    mbar_wait(smem_u32(&ctl->done[b]), (unsigned) (t / NB) & 1u);
                    const uint32_t src = smem_u32(bufs + b * TILE);
                    if constexpr (!S::STRIDED)
                        tma_store_3d(&map_out, 0, 0, (int) (((long long) poly << (n - S::KC)) + ((long long) pos << S::NPLOG)), src);
                    else if constexpr (TS) // transposing box {16 rows, 2^C columns, 2^(D-4) row blocks} of the transposed matrix
                        tma_store_3d(&map_out, 0, (int) (((long long) poly << lo) + ((long long) cc << S::C)), range << (S::D - 4), src);
                    else
                        tma_store_3d(&map_out, 0, cc << (S::C - S::CB), (int) (((long long) poly << (n - lo)) + ((long long) range << S::D)), src);
                    bulk_commit();
                    bulk_wait_read0();
                    mbar_arrive(smem_u32(&ctl->free_[b]));
                }
                mbar_arrive(smem_u32(&ctl->pairs_free)); // (in-order: the position's last tile has been computed and stored)
            }
            bulk_wait0();
        }
        else if (tid < kWcolConsumers)
        {
            // =================== consumer groups ===================
            const int g = tid / kConsumers, ctid = tid % kConsumers;
            typename ModOf<S>::type M(a.p);
            const Twiddle<T> ninv{a.ninv_w, a.ninv_wq};
            const bool triv = S::STRIDED && !a.plus && (S::INV ? a.last : a.first) && (lo + S::D == n) && a.table[0] == T(1);
            FastArgs<T> aw = a;
            aw.lo = S::STRIDED ? S::C : 0; // the pairs of a tile sit in shared memory in tile order: pair of local element l at index l
            if (ctid == 0) ctl->bcast[g][0] = atomicAdd(&ctl->next_t, 1);
            consumer_sync(1 + g);
            int t = ctl->bcast[g][0];
            for (int it = 0; t < total; it++)
            {
                if (ctid == 0) ctl->bcast[g][(it + 1) & 1] = atomicAdd(&ctl->next_t, 1);
                const int b = t % NB, k = t / batch;
                // twiddle set of this tile's range: the i-th range this CTA meets lives in set i & 1 (its (i >> 1)-th fill); the
                // group holding the range's first tile builds it (a range has at least batch >= 4 tiles).
                const int range = per_range_tw ? (pos0 + k * pos_step) / ncc : range0;
                const int ri = range - range0;
                Twiddle<T>* tws = tw + (ri & 1) * TWN;
                const bool opens = (t % batch == 0) && (k == 0 || (per_range_tw && (pos0 + (k - 1) * pos_step) / ncc != range));
                // (the tile has landed => the loader saw tile t - NB stored => every tile still being computed is one of
                // t - 3 .. t, all of this range or the one before: the set about to be overwritten has no readers left)
                mbar_wait(smem_u32(&ctl->full[b]), (unsigned) (t / NB) & 1u);
                if (opens)
                {
                    build_twiddles<S>(tws, a.table, a.tw_fixed ? 0 : range, n, a.n_tw, lo, a.plus, a.p, a.mu, a.pbits, ctid, kConsumers);
                    mbar_arrive(smem_u32(&ctl->tw_ready[ri & 1]));
                }
                mbar_wait(smem_u32(&ctl->tw_ready[ri & 1]), (unsigned) (ri >> 1) & 1u);
                mbar_wait(smem_u32(&ctl->pairs_full), (unsigned) k & 1u); // this position's pairs are in shared memory
                tile_rounds<S, true, false, TS, true, 0>(bufs + b * TILE, tws, tws + S::TW1, tws + S::TW1 + S::TW2, M, ctid, ninv, pairs, aw, triv, 1 + g);
                fence_async();
                mbar_arrive(smem_u32(&ctl->done[b]));
                consumer_sync(1 + g);
                t = ctl->bcast[g][(it + 1) & 1];
            }
        }
    }

    // Pairs array [2^n] of 16-byte (w, w') entries.  Strided passes: viewed as {2^(lo + 1) 64-bit words, 2^(n - lo) rows}, box = one
    // tile position {2^(C + 1) words, 2^D rows}.  Contiguous passes: {256 words, 2^(n + 1 - 8) rows}, box {256, 2^(K + 1 - 8)} = the
    // 2^K pairs of a position (written element-major by w_pairs_kernel, see fast_round).
    template <typename S> static bool make_map_pairs(CUtensorMap* map, const void* base, int n, int lo)
    {
        PFN_cuTensorMapEncodeTiled enc = get_encode();
        if (!enc) return false;
        cuuint64_t gdim[2], gstride[1];
        cuuint32_t box[2], estr[2] = {1, 1};
        if constexpr (S::STRIDED)
        {
            gdim[0] = 2ull << lo;
            gdim[1] = 1ull << (n - lo);
            gstride[0] = (cuuint64_t) 16 << lo;
            box[0] = 2u << S::C;
            box[1] = 1u << S::D;
        }
        else
        {
            gdim[0] = 256;
            gdim[1] = 1ull << (n + 1 - 8);
            gstride[0] = 2048;
            box[0] = 256;
            box[1] = 1u << (S::K + 1 - 8);
        }
        CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS;
    }

    template <typename S, bool TS> static cudaError_t launch_wcol(const FastArgs<uint64_t>& a, cudaStream_t st)
    {
        constexpr int kMaxDev = 64;
        static std::atomic<int> cached_sms[kMaxDev];
        auto kern = wcol_kernel<S, TS>;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDev) return cudaErrorNotSupported;
        int sms = cached_sms[dev].load(std::memory_order_acquire);
        if (sms <= 0)
        {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WcolSmem<S>::BYTES);
            if (e != cudaSuccess) return e;
            int b = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kWcolThreads, WcolSmem<S>::BYTES);
            if (e != cudaSuccess) return e;
            if (b < 1) return cudaErrorNotSupported;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cached_sms[dev].store(sms, std::memory_order_release);
        }
        if (a.batch < 4) return cudaErrorNotSupported; // too little reuse to pay for the position-major order (and see the twiddle sets)
        if (S::INV && a.last) return cudaErrorNotSupported; // (the kernel carries no last-round code: tile_rounds LASTC = 0)
        if (TS && a.in == a.out) return cudaErrorNotSupported;
        alignas(64) CUtensorMap m_in, m_out, m_pairs;
        long long npos;
        if constexpr (S::STRIDED)
        {
            if (a.lo + S::D > a.n || a.lo < S::C) return cudaErrorNotSupported;
            if (((long long) a.batch << (a.n - a.lo)) >= (1LL << 31)) return cudaErrorNotSupported;
            if (!make_map<S>(&m_in, a.in, a.n, a.lo, a.batch)) return cudaErrorNotSupported;
            if constexpr (TS)
            {
                if (!make_map_tstore<S>(&m_out, a.out, a.lo, a.batch, a.n)) return cudaErrorNotSupported;
            }
            else if (a.in == a.out)
                m_out = m_in;
            else if (!make_map<S>(&m_out, a.out, a.n, a.lo, a.batch))
                return cudaErrorNotSupported;
            npos = (1LL << (a.lo - S::C)) << (a.n - a.lo - S::D);
        }
        else
        {
            if (a.n < S::K || a.lo != 0) return cudaErrorNotSupported;
            const long long chunks = (long long) a.batch << (a.n - S::KC);
            if (chunks > 0x7fffffffLL) return cudaErrorNotSupported;
            if (!make_map<S>(&m_in, a.in, S::KC, 0, (int) chunks)) return cudaErrorNotSupported;
            if (a.in == a.out)
                m_out = m_in;
            else if (!make_map<S>(&m_out, a.out, S::KC, 0, (int) chunks))
                return cudaErrorNotSupported;
            npos = 1LL << (a.n - S::K);
        }
        if (!make_map_pairs<S>(&m_pairs, a.w_pairs, a.n, a.lo)) return cudaErrorNotSupported;
        const int grid = npos < sms ? (int) npos : sms;
        kern<<<grid, kWcolThreads, WcolSmem<S>::BYTES, st>>>(a, m_in, m_out, m_pairs);
        return cudaGetLastError();
    }

    // a: everything fast_fourstep_columns fills in (in, out, n1 table, p, mu, pbits, n, lo = lg2, batch, w_pairs, w_lazy, first).
    // cudaErrorNotSupported: the caller uses fast_pass_kernel<..., WMUL, TS>.
    cudaError_t fourstep_columns_resident_pairs(const FastArgs<uint64_t>& a, int lg1, cudaStream_t st)
    {
        using T = uint64_t;
        switch (lg1)
        {
            case 5: return launch_wcol<Shape<T, false, 2, true, 3, 2, 12, 0>, true>(a, st);
            case 6: return launch_wcol<Shape<T, false, 2, true, 3, 3, 12, 0>, true>(a, st);
            case 7: return launch_wcol<Shape<T, false, 2, true, 4, 3, 12, 0>, true>(a, st);
            case 8: return launch_wcol<Shape<T, false, 2, true, 4, 4, 12, 0>, true>(a, st);
            default: return cudaErrorNotSupported;
        }
    }

    // Forward column phase on the TRANSPOSED input (reference contract: the caller's GPU_Transpose made the n2 x n1 matrix, so a
    // column transform is a contiguous run of n1 elements): whole size-2^lg1 transforms inside contiguous tiles, product as
    // epilogue, same layout out.  a.n = log2 N, a.lo = 0, a.n_tw = lg1, a.w_pairs = the pair table in tile order.
    cudaError_t fourstep_rows_of_transposed_resident_pairs(const FastArgs<uint64_t>& a, int lg1, cudaStream_t st)
    {
        using T = uint64_t;
        switch (lg1)
        {
            case 5: return launch_wcol<Shape<T, false, 2, false, 1, 4, 12, 1, 5>, false>(a, st);
            case 6: return launch_wcol<Shape<T, false, 2, false, 2, 4, 12, 1, 6>, false>(a, st);
            case 7: return launch_wcol<Shape<T, false, 2, false, 3, 4, 12, 1, 7>, false>(a, st);
            case 8: return launch_wcol<Shape<T, false, 2, false, 4, 4, 12, 1, 8>, false>(a, st);
            default: return cudaErrorNotSupported;
        }
    }

    // Inverse: the strided Gentleman-Sande pass of d stages that opens the size-n2 transforms (row stride a.lo = lg1, 2^(n - lo - d)
    // twiddle ranges), product as prologue; transposed: the store writes the n1 x n2 matrix (reference contract).
    cudaError_t fourstep_inverse_product_pass_resident_pairs(const FastArgs<uint64_t>& a, int d, bool transposed, cudaStream_t st)
    {
        using T = uint64_t;
        if (transposed)
            switch (d)
            {
                case 5: return launch_wcol<Shape<T, true, 1, true, 3, 2, 12, 0>, true>(a, st);
                case 6: return launch_wcol<Shape<T, true, 1, true, 3, 3, 12, 0>, true>(a, st);
                case 7: return launch_wcol<Shape<T, true, 1, true, 4, 3, 12, 0>, true>(a, st);
                case 8: return launch_wcol<Shape<T, true, 1, true, 4, 4, 12, 0>, true>(a, st);
                default: return cudaErrorNotSupported;
            }
        switch (d)
        {
            case 4: return launch_wcol<Shape<T, true, 1, true, 4, 0, 12, 0>, false>(a, st);
            case 5: return launch_wcol<Shape<T, true, 1, true, 3, 2, 12, 0>, false>(a, st);
            case 6: return launch_wcol<Shape<T, true, 1, true, 3, 3, 12, 0>, false>(a, st);
            case 7: return launch_wcol<Shape<T, true, 1, true, 4, 3, 12, 0>, false>(a, st);
            case 8: return launch_wcol<Shape<T, true, 1, true, 4, 4, 12, 0>, false>(a, st);
            default: return cudaErrorNotSupported;
        }
    }

} // namespace gpuntt_b200
