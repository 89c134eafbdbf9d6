// gpu_ntt_b200/csrc/merge_wcol.cu -- the forward 4-step column phase with the twiddle-matrix pairs resident in shared memory.
//
// The column phase of the reference's 4-step transform (FourStepForwardCoreT4 + the W product fused into its first row
// kernel, ntt_4step.cu:571-743, 745-880 of the reference) is here one strided pass over the top log2(n1) index bits whose
// epilogue multiplies every element by W[offset] and whose store writes the n2 x n1 matrix (fast_round TS).  The first
// version of that pass (fast_pass_kernel<..., WMUL, TS>) fetched the (W, W') pairs of a tile with __ldg behind an L1
// prefetch: 64 KiB of pairs per 32 KiB tile, more than the L1 that is left beside two CTAs' tile buffers, and the ncu
// capture shows the consumers stalled on those loads (profiles/r2_v1_ncu_summary.txt: long_scoreboard the top stall, 57 %
// multiplier-pipe utilisation against 70 % for the row passes).
// This kernel turns the loop around: a CTA owns tile POSITIONS (a block of 2^C matrix columns) and walks every polynomial
// of the batch through one position before it moves on, so the position's 64 KiB of pairs are loaded ONCE by the TMA
// engine into shared memory and reused batch_size times; the epilogue reads them with LDS.128 (conflict-free: a quarter
// warp reads 128 contiguous bytes).  One CTA per SM: two consumer groups of 8 warps, a loader thread, a storer thread,
// four data-tile buffers (same skeleton as merge_fused.cu).
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
        int next_t;
        int bcast[kWcolGroups][2];
    };

    template <typename S> struct WcolSmem
    {
        static constexpr int TILE = S::TILE_SMEM;
        static constexpr int PAIRS = (1 << S::K) * (int) sizeof(Twiddle<typename S::T>);
        static constexpr int BYTES = kWcolBufs * TILE + PAIRS + S::TW_SMEM + (int) sizeof(WcolCtl) + 1024;
    };

    template <typename S>
    __global__ void __launch_bounds__(kWcolThreads, 1)
        wcol_kernel(const FastArgs<typename S::T> a, const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                    const __grid_constant__ CUtensorMap map_pairs)
    {
        using T = typename S::T;
        static_assert(S::STRIDED && !S::INV && sizeof(T) == 8 && S::POL == 2, "forward 64-bit strided column pass");
        constexpr int TILE = S::TILE_SMEM, NB = kWcolBufs;
        extern __shared__ __align__(128) unsigned char smem_raw[];
        unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
        unsigned char* bufs = smem;
        Twiddle<T>* pairs = reinterpret_cast<Twiddle<T>*>(smem + NB * TILE);
        Twiddle<T>* tw1 = reinterpret_cast<Twiddle<T>*>(smem + NB * TILE + WcolSmem<S>::PAIRS);
        WcolCtl* ctl = reinterpret_cast<WcolCtl*>(smem + NB * TILE + WcolSmem<S>::PAIRS + S::TW_SMEM);

        const int tid = threadIdx.x;
        const int batch = a.batch, n = a.n, lo = a.lo;
        const int npos_all = 1 << (lo - S::C);                              // tile positions of the matrix (one twiddle range)
        const int npos = ((int) blockIdx.x < npos_all) ? (npos_all - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0; // this CTA's
        const int total = npos * batch;

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
                const int cc = (int) blockIdx.x + k * (int) gridDim.x;
                if (k > 0) mbar_wait(smem_u32(&ctl->pairs_free), (unsigned) (k - 1) & 1u); // every tile of the previous position is stored
                mbar_expect_tx(smem_u32(&ctl->pairs_full), WcolSmem<S>::PAIRS);
                tma_load_2d(smem_u32(pairs), &map_pairs, cc << (S::C + 1), 0, smem_u32(&ctl->pairs_full));
                for (int poly = 0; poly < batch; poly++)
                {
                    const int t = k * batch + poly, b = t % NB;
                    if (t >= NB) mbar_wait(smem_u32(&ctl->free_[b]), (unsigned) (t / NB - 1) & 1u);
                    mbar_expect_tx(smem_u32(&ctl->full[b]), TILE);
                    tma_load_3d(smem_u32(bufs + b * TILE), &map_in, 0, cc << (S::C - S::CB), (int) ((long long) poly << (n - lo)), smem_u32(&ctl->full[b]));
                }
            }
        }
        else if (tid == kWcolConsumers + 32)
        {
            // =================== storer ===================
            for (int k = 0; k < npos; k++)
            {
                const int cc = (int) blockIdx.x + k * (int) gridDim.x;
                for (int poly = 0; poly < batch; poly++)
                {
                    const int t = k * batch + poly, b = t % NB;
                    mbar_wait(smem_u32(&ctl->done[b]), (unsigned) (t / NB) & 1u);
                    // transposing box {16 rows, 2^C columns, 2^(D-4) row blocks} of the n2 x n1 output matrix
                    tma_store_3d(&map_out, 0, (int) (((long long) poly << lo) + ((long long) cc << S::C)), 0, smem_u32(bufs + b * TILE));
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
            const bool triv = !a.plus && a.first && (lo + S::D == n) && a.table[0] == T(1);
            build_twiddles<S>(tw1, a.table, 0, n, a.n_tw, lo, a.plus, a.p, a.mu, a.pbits, tid, kWcolConsumers);
            asm volatile("bar.sync 3, %0;" ::"n"(kWcolConsumers) : "memory");
            FastArgs<T> aw = a;
            aw.lo = S::C; // the pairs of a tile sit in shared memory in tile order: pair of local element l at index l
            if (ctid == 0) ctl->bcast[g][0] = atomicAdd(&ctl->next_t, 1);
            consumer_sync(1 + g);
            int t = ctl->bcast[g][0];
            for (int it = 0; t < total; it++)
            {
                if (ctid == 0) ctl->bcast[g][(it + 1) & 1] = atomicAdd(&ctl->next_t, 1);
                const int b = t % NB, k = t / batch;
                mbar_wait(smem_u32(&ctl->pairs_full), (unsigned) k & 1u); // this position's pairs are in shared memory
                mbar_wait(smem_u32(&ctl->full[b]), (unsigned) (t / NB) & 1u);
                tile_rounds<S, true, false, true, true>(bufs + b * TILE, tw1, tw1 + S::TW1, tw1 + S::TW1 + S::TW2, M, ctid, ninv, pairs, aw, triv, 1 + g);
                fence_async();
                mbar_arrive(smem_u32(&ctl->done[b]));
                consumer_sync(1 + g);
                t = ctl->bcast[g][(it + 1) & 1];
            }
        }
    }

    // Pairs array [2^n] of 16-byte (w, w') entries viewed as {2^(lo + 1) 64-bit words, 2^D rows}; box = one tile position.
    template <typename S> static bool make_map_pairs(CUtensorMap* map, const void* base, int lo)
    {
        PFN_cuTensorMapEncodeTiled enc = get_encode();
        if (!enc) return false;
        cuuint64_t gdim[2] = {2ull << lo, 1ull << S::D};
        cuuint64_t gstride[1] = {(cuuint64_t) 16 << lo};
        cuuint32_t box[2] = {2u << S::C, 1u << S::D}, estr[2] = {1, 1};
        CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS;
    }

    template <typename S> static cudaError_t launch_wcol(const FastArgs<uint64_t>& a, cudaStream_t st)
    {
        constexpr int kMaxDev = 64;
        static std::atomic<int> cached_sms[kMaxDev];
        auto kern = wcol_kernel<S>;
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
        if (a.in == a.out || a.lo + S::D != a.n || a.lo < S::C) return cudaErrorNotSupported;
        alignas(64) CUtensorMap m_in, m_out, m_pairs;
        if (!make_map<S>(&m_in, a.in, a.n, a.lo, a.batch)) return cudaErrorNotSupported;
        if (!make_map_tstore<S>(&m_out, a.out, a.lo, a.batch)) return cudaErrorNotSupported;
        if (!make_map_pairs<S>(&m_pairs, a.w_pairs, a.lo)) return cudaErrorNotSupported;
        const int npos = 1 << (a.lo - S::C);
        const int grid = npos < sms ? npos : sms;
        kern<<<grid, kWcolThreads, WcolSmem<S>::BYTES, st>>>(a, m_in, m_out, m_pairs);
        return cudaGetLastError();
    }

    // a: everything fast_fourstep_columns fills in (in, out, n1 table, p, mu, pbits, n, lo = lg2, batch, w_pairs, w_lazy, first).
    // cudaErrorNotSupported: the caller uses fast_pass_kernel<..., WMUL, TS>.
    cudaError_t fourstep_columns_resident_pairs(const FastArgs<uint64_t>& a, int lg1, cudaStream_t st)
    {
        using T = uint64_t;
        if (a.batch < 4) return cudaErrorNotSupported; // too little reuse to pay for the position-major order
        switch (lg1)
        {
            case 5: return launch_wcol<Shape<T, false, 2, true, 3, 2, 12, 0>>(a, st);
            case 6: return launch_wcol<Shape<T, false, 2, true, 3, 3, 12, 0>>(a, st);
            case 7: return launch_wcol<Shape<T, false, 2, true, 4, 3, 12, 0>>(a, st);
            case 8: return launch_wcol<Shape<T, false, 2, true, 4, 4, 12, 0>>(a, st);
            default: return cudaErrorNotSupported;
        }
    }

} // namespace gpuntt_b200
