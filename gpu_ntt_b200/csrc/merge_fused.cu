// gpu_ntt_b200/csrc/merge_fused.cu -- two-pass Merge-NTT plans in ONE launch, the passes chained through the L2.
//
// The reference runs a 2^12..2^16-point transform as two kernels (plans [logN-9, 9], ntt.cuh:628-636 of the
// reference; launches at ntt.cu:2104-2141), so every coefficient crosses HBM twice.  The tuned kernels of
// fast_kernels.cuh kept that structure (strided pass + contiguous pass, two launches).  Here both passes live in one
// persistent kernel:
//   * every CTA owns a share of the FIRST pass's tiles and a share of the SECOND pass's tiles and walks them in one
//     merged order; a second-pass tile is only loaded once the first pass has finished every polynomial it touches
//     (one counter word per polynomial in global memory: producers add 1 per stored tile with release semantics, the
//     TMA-issuing thread polls with acquire loads -- never blocking its own stores, see below);
//   * the second pass trails the first by a few tile times (`lag` polynomials), so what it reads was written to
//     the L2 microseconds earlier and is still there (a wave of 2 x 148 tiles is ~10 MiB against 126 MB of L2), and
//     in-place transforms overwrite the same dirty lines before they are evicted: the data crosses HBM ONCE
//     (read by the first pass, written back after the second);
//   * a call is one launch instead of two.
// Deadlock freedom: every tile has a virtual time (first pass: its polynomial; second pass: its last polynomial +
// lag, lag >= polynomials per contiguous tile); every CTA processes its tiles in increasing time and a tile only
// depends on tiles of strictly smaller time.  A tile that has been loaded is always computed, stored and signalled:
// the producer thread polls "dependency satisfied?" and "consumers done?" in one loop and never spins on a dependency
// while a finished tile waits for its store.  The grid never exceeds the number of co-resident CTAs.
// The counters clean up after themselves: every CTA takes a ticket when it is done and the last one zeroes the words
// the call used, so the workspace is all-zero between calls, no memset is enqueued, and a captured graph can be
// replayed any number of times.
#include "fast_kernels.cuh"

namespace gpuntt_b200
{

#ifdef GPUNTT_TIMELINE
    // Lab build only (tools/build_variant.sh timeline -DGPUNTT_TIMELINE): SM cycle counter at the hand-off points of a CTA's FIRST
    // tile, read back by tools/fused_timeline.py -- where the microseconds of a launch-bound call go.
    __device__ long long g_timeline[512][16];
    __device__ __forceinline__ void tl_mark(int slot)
    {
        if (blockIdx.x < 512) g_timeline[blockIdx.x][slot] = clock64();
    }
    __device__ __forceinline__ void tl_mark_global(int slot)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (blockIdx.x < 512) g_timeline[blockIdx.x][slot] = (long long) t;
    }
#define TL(slot) tl_mark(slot)
#define TLG(slot) tl_mark_global(slot)
#else
#define TL(slot)
#define TLG(slot)
#endif

    template <typename T> struct FusedArgs
    {
        FastArgs<T> s, c;     // the strided pass / the contiguous pass (in, out, work, rr, cta_per_seg unused)
        unsigned* counters;   // one word per polynomial: first-pass tiles stored so far
        int fwd;              // 1: strided pass first (forward transform), 0: contiguous pass first (inverse)
        int lag;              // see above
        int g_str;            // CTAs [0, g_str) take the strided tiles, round robin
        int c_off, con_k, con_extra; // CTAs from c_off on take the contiguous tiles: the first con_extra ranges have con_k + 1
                                     // CTAs, the others con_k; CTA j of a range takes tile groups j, j + k, ...
        int tpp_log;          // log2 strided tiles per polynomial
        int nranges;          // contiguous ranges per polynomial
        int ngroups;          // contiguous tile groups = ceil(batch / polynomials per tile)
        long long ticket_off; // counters[ticket_off]: CTAs that have finished (the last one zeroes the counters)
        int g_slot;           // RNS: CTAs per modulus slot (grid = mod_count * g_slot)
    };

    __device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity)
    {
        uint32_t ok;
        asm volatile("{\n\t"
                     ".reg .pred P;\n\t"
                     "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, P;\n\t"
                     "}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        return ok;
    }
    __device__ __forceinline__ unsigned ld_acquire(const unsigned* p)
    {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void red_release_add(unsigned* p, unsigned v)
    {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    __device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

    // A CTA's position in its two tile streams.  next() yields the tiles in the merged (virtual time) order.
    struct FusedCursor
    {
        long long s_next, s_end; // strided tile ids (poly << tpp_log | column chunk), step s_step
        int s_step;
        int q_next, q_end, q_step; // contiguous tile groups
    };
    struct FusedTile
    {
        int kind;      // 0 strided, 1 contiguous, -1 none
        long long id;  // strided tile id / contiguous group
    };

    template <int NPLOG> __device__ __forceinline__ FusedTile fused_peek(const FusedCursor& k, int fwd, int lag, int tpp_log, int batch)
    {
        const bool hasS = k.s_next < k.s_end, hasC = k.q_next < k.q_end;
        FusedTile t;
        t.kind = -1;
        t.id = 0;
        if (!hasS && !hasC) return t;
        const long long polyS = k.s_next >> tpp_log;
        long long pmaxC = (((long long) k.q_next + 1) << NPLOG) - 1;
        if (pmaxC > batch - 1) pmaxC = batch - 1;
        bool pickC;
        if (fwd)
            pickC = hasC && (!hasS || pmaxC + lag <= polyS); // contiguous is the second pass: time pmaxC + lag
        else
            pickC = hasC && (!hasS || pmaxC <= polyS + lag); // strided is the second pass: time polyS + lag
        t.kind = pickC ? 1 : 0;
        t.id = pickC ? (long long) k.q_next : k.s_next;
        return t;
    }
    __device__ __forceinline__ void fused_advance(FusedCursor& k, const FusedTile& t)
    {
        if (t.kind == 1)
            k.q_next += k.q_step;
        else
            k.s_next += k.s_step;
    }

    // One CTA per SM: two consumer groups of 8 warps (each takes the next tile of the CTA's merged order as soon as it is
    // free, so short strided tiles and long contiguous tiles balance out), a LOADER thread and a STORER thread in warps of
    // their own, kFusedBufs tile buffers.  Loads run up to kFusedBufs tiles ahead of the arithmetic.  The loader blocks on
    // dependencies, the storer on finished tiles and on the completion of first-pass stores (which it then signals) --
    // neither ever waits for the other except through the free[] barriers of the buffers, so a finished tile is always
    // stored and signalled.  (A single producer thread doing all of this serially was the bottleneck of the first two
    // versions: the consumers spent 32 % / 55 % of their time waiting for tiles, profiles/r2_fused_v1_ncu_summary.txt.)
    constexpr int kFusedGroups = 2;
    constexpr int kFusedConsumers = kFusedGroups * kConsumers;
    constexpr int kFusedThreads = kFusedConsumers + 64;
    constexpr int kFusedBufs = 5;

    struct FusedCtl
    {
        uint64_t full[kFusedBufs], done[kFusedBufs], free_[kFusedBufs];
        int kind[kFusedBufs]; // what the loader put into each buffer
        int next_t;           // next tile index to be claimed by a consumer group
        int bcast[kFusedGroups][2];
        int is_last;
        // RNS: this CTA's modulus slot
        unsigned long long seg_p, seg_mu, seg_ninv_w, seg_ninv_wq;
        int seg_pbits, seg_mi;
    };

    template <typename SS, typename SC> struct FusedSmem
    {
        static constexpr int TILE = SS::TILE_SMEM;
        static constexpr int TW = SS::TW_SMEM + SC::TW_SMEM;
        static constexpr int BYTES = kFusedBufs * TILE + TW + (int) sizeof(FusedCtl) + 1024; // + slack to align the tiles to 1 KiB
    };

    __device__ __forceinline__ unsigned long long ld_acquire64(const unsigned* p)
    {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }

    // RNS (polynomial b uses modulus slot b % mod_count, ntt.cu:613-619 of the reference): the CTAs are partitioned among the
    // modulus slots (f.g_slot CTAs each), so a CTA has ONE modulus, one strided twiddle set and one contiguous twiddle set
    // like in the single-modulus form; f.s.batch is then the number of polynomials PER SLOT, counters are indexed by the
    // polynomial's position in the caller's array, contiguous tiles come from the 4-D map {row, rows, slot, polynomial}.
    template <typename SS, typename SC, bool RNS>
    __device__ __forceinline__ void fused2_body(const FusedArgs<typename SS::T>& f, const CUtensorMap& mapA_in, const CUtensorMap& mapA_out,
                                                const CUtensorMap& mapB)
    {
        using T = typename SS::T;
        static_assert(SS::STRIDED && !SC::STRIDED && SS::TILE_SMEM == SC::TILE_SMEM && SS::INV == SC::INV, "one strided and one contiguous pass");
        static_assert(SC::NT == 0 && SC::R3 == 0, "merge passes only");
        constexpr int TILE = SS::TILE_SMEM;
        constexpr int NB = kFusedBufs;
        extern __shared__ __align__(128) unsigned char smem_raw[];
        unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
        unsigned char* bufs = smem;
        Twiddle<T>* twS = reinterpret_cast<Twiddle<T>*>(smem + NB * TILE);
        Twiddle<T>* twC = twS + SS::TWN;
        FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem + NB * TILE + SS::TW_SMEM + SC::TW_SMEM);

        const int tid = threadIdx.x;
        const int batch = f.s.batch, n = f.s.n;
        const int fwd = f.fwd, lag = f.lag, tpp_log = f.tpp_log;
        const int mslot = RNS ? (int) blockIdx.x / f.g_slot : 0;          // modulus slot of this CTA
        const int cta = RNS ? (int) blockIdx.x % f.g_slot : (int) blockIdx.x; // index among the CTAs of the slot
        const int mc = RNS ? f.s.mod_count : 1;

        // ---- this CTA's shares
        FusedCursor cur;
        cur.s_step = f.g_str;
        cur.s_next = cta < f.g_str ? (long long) cta : 0;
        cur.s_end = cta < f.g_str ? ((long long) batch << tpp_log) : 0;
        int range = 0;
        cur.q_next = 0;
        cur.q_end = 0;
        cur.q_step = 1;
        if (cta >= f.c_off)
        {
            const int cb = cta - f.c_off;
            const int big = f.con_extra * (f.con_k + 1);
            int kk, j;
            if (cb < big)
            {
                kk = f.con_k + 1;
                range = cb / kk;
                j = cb % kk;
            }
            else
            {
                kk = f.con_k;
                range = f.con_extra + (cb - big) / kk;
                j = (cb - big) % kk;
            }
            if (range < f.nranges)
            {
                cur.q_next = j;
                cur.q_end = f.ngroups;
                cur.q_step = kk;
            }
        }
        if (tid == 0)
        {
            TL(0);
            TLG(15);
        }
        const bool doS = cur.s_next < cur.s_end, doC = cur.q_next < cur.q_end;
        const int total = (doS ? (int) ((cur.s_end - 1 - cur.s_next) / cur.s_step + 1) : 0) + (doC ? (cur.q_end - 1 - cur.q_next) / cur.q_step + 1 : 0);

        if (tid == kFusedConsumers)
        {
            tma_prefetch_desc(&mapA_in);
            tma_prefetch_desc(&mapB);
        }
        if (tid == kFusedConsumers + 32)
        {
            tma_prefetch_desc(&mapA_out);
            tma_prefetch_desc(&mapB);
        }
        if (tid == 0)
        {
            for (int b = 0; b < NB; b++)
            {
                mbar_init(smem_u32(&ctl->full[b]), 1);
                mbar_init(smem_u32(&ctl->done[b]), kConsumers);
                mbar_init(smem_u32(&ctl->free_[b]), 1);
            }
            ctl->next_t = 0;
            ctl->is_last = 0;
            if constexpr (RNS)
            {
                const int mi = f.s.mod_order ? f.s.mod_order[mslot] : mslot;
                const T p = f.s.mod_dev[3 * mi];
                ctl->seg_mi = mi;
                ctl->seg_p = (unsigned long long) p;
                if constexpr (sizeof(T) == 8)
                {
                    ctl->seg_pbits = 64 - __clzll((long long) p);
                    ctl->seg_mu = (p & (p - 1)) ? recip_mu64(p, ctl->seg_pbits) : ~0ull;
                }
                else
                {
                    ctl->seg_pbits = 32 - __clz((int) p);
                    ctl->seg_mu = ~0ull / (unsigned long long) p;
                }
                const T nv = SS::INV ? f.s.ninv_dev[mi] : T(0);
                ctl->seg_ninv_w = (unsigned long long) nv;
                if constexpr (sizeof(T) == 8)
                    ctl->seg_ninv_wq = SS::INV ? shoup_companion_mu(nv, p, ctl->seg_mu, ctl->seg_pbits) : 0ull;
                else
                    ctl->seg_ninv_wq = SS::INV ? (unsigned long long) shoup_companion(nv, p) : 0ull;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fence_async();
        }
        __syncthreads();
        if (tid == 0) TL(1);
        const T seg_p = RNS ? (T) ctl->seg_p : f.s.p;
        const uint64_t seg_mu = RNS ? ctl->seg_mu : f.s.mu;
        const int seg_pbits = RNS ? ctl->seg_pbits : f.s.pbits;
        const T* tabS = RNS ? f.s.table + ((size_t) ctl->seg_mi << n) : f.s.table;
        const T* tabC = RNS ? f.c.table + ((size_t) ctl->seg_mi << n) : f.c.table;

        if (tid == kFusedConsumers)
        {
            // =================== loader ===================
            FusedCursor ldc = cur;
            for (int t = 0; t < total; t++)
            {
                const FusedTile tl = fused_peek<SC::NPLOG>(ldc, fwd, lag, tpp_log, batch);
                fused_advance(ldc, tl);
                const bool second = fwd ? (tl.kind == 1) : (tl.kind == 0);
                const int b = t % NB;
                if (t >= NB) mbar_wait(smem_u32(&ctl->free_[b]), (unsigned) (t / NB - 1) & 1u); // the storer released this buffer
                if (t == 0) TL(2);
                if (second)
                {
                    // every first-pass tile of the polynomials this tile touches has been stored
                    if (tl.kind == 1)
                    {
                        const long long p0 = tl.id << SC::NPLOG;
                        long long p1 = p0 + (1 << SC::NPLOG);
                        if (p1 > batch) p1 = batch;
                        const unsigned need = 1u << tpp_log;
                        if (!RNS && SC::NPLOG == 1 && p1 - p0 == 2)
                        {
                            const unsigned long long want = ((unsigned long long) need << 32) | need;
                            while (ld_acquire64(f.counters + p0) != want) __nanosleep(64);
                        }
                        else
                            for (long long p = p0; p < p1; p++)
                                while (ld_acquire(f.counters + p * mc + mslot) != need) __nanosleep(64);
                    }
                    else
                    {
                        const unsigned* c = f.counters + (tl.id >> tpp_log) * mc + mslot;
                        while (ld_acquire(c) != (unsigned) f.nranges) __nanosleep(64);
                    }
                    asm volatile("fence.proxy.async.global;" ::: "memory"); // the bulk read below is ordered after the acquire loads
                }
                if (t == 0) TL(3);
                const uint32_t bar = smem_u32(&ctl->full[b]);
                const uint32_t dst = smem_u32(bufs + b * TILE);
                ctl->kind[b] = tl.kind; // (released by the arrive below, acquired by the consumers' wait)
                mbar_expect_tx(bar, TILE);
                const CUtensorMap* mp = second ? &mapB : &mapA_in;
                if (tl.kind == 0)
                {
                    const long long poly = (tl.id >> tpp_log) * mc + mslot, cc = tl.id & ((1LL << tpp_log) - 1);
                    tma_load_3d(dst, mp, 0, (int) (cc << (SS::C - SS::CB)), (int) (poly << (n - f.s.lo)), bar);
                }
                else if constexpr (RNS)
                    tma_load_4d(dst, mp, 0, range << (SC::KC - SC::CB), mslot, (int) (tl.id << SC::NPLOG), bar);
                else
                    tma_load_3d(dst, mp, 0, range << (SC::KC - SC::CB), (int) (tl.id << SC::NPLOG), bar);
            }
        }
        else if (tid == kFusedConsumers + 32)
        {
            // =================== storer ===================
            FusedCursor stc = cur;
            for (int t = 0; t < total; t++)
            {
                const FusedTile tl = fused_peek<SC::NPLOG>(stc, fwd, lag, tpp_log, batch);
                fused_advance(stc, tl);
                const bool second = fwd ? (tl.kind == 1) : (tl.kind == 0);
                const int b = t % NB;
                mbar_wait(smem_u32(&ctl->done[b]), (unsigned) (t / NB) & 1u); // a consumer group finished this tile
                if (t == 0) TL(7);
                const uint32_t src = smem_u32(bufs + b * TILE);
                const CUtensorMap* mp = second ? &mapB : &mapA_out;
                if (tl.kind == 0)
                {
                    const long long poly = (tl.id >> tpp_log) * mc + mslot, cc = tl.id & ((1LL << tpp_log) - 1);
                    tma_store_3d(mp, 0, (int) (cc << (SS::C - SS::CB)), (int) (poly << (n - f.s.lo)), src);
                }
                else if constexpr (RNS)
                    tma_store_4d(mp, 0, range << (SC::KC - SC::CB), mslot, (int) (tl.id << SC::NPLOG), src);
                else
                    tma_store_3d(mp, 0, range << (SC::KC - SC::CB), (int) (tl.id << SC::NPLOG), src);
                bulk_commit();
                bulk_wait_read0();
                if (t == 0) TL(8);
                mbar_arrive(smem_u32(&ctl->free_[b])); // the loader may refill the buffer
                if (!second)
                {
                    // first-pass tile: its polynomials advance once the bulk store is COMPLETE (not merely read); wait_group
                    // makes the writes visible to this thread, the release publishes them
                    bulk_wait0();
                    if (tl.kind == 0)
                        red_release_add(f.counters + (tl.id >> tpp_log) * mc + mslot, 1u);
                    else
                    {
                        long long p0 = tl.id << SC::NPLOG, p1 = p0 + (1 << SC::NPLOG);
                        if (p1 > batch) p1 = batch;
                        for (long long p = p0; p < p1; p++) red_release_add(f.counters + p * mc + mslot, 1u);
                    }
                    if (t == 0) TL(9);
                }
            }
            bulk_wait0();
            TL(10);
        }
        else if (tid < kFusedConsumers)
        {
            // =================== consumer groups ===================
            const int g = tid / kConsumers, ctid = tid % kConsumers;
            typename ModOf<SS>::type MS(seg_p);
            typename ModOf<SC>::type MC(seg_p);
            const Twiddle<T> ninv = RNS ? Twiddle<T>{(T) ctl->seg_ninv_w, (T) ctl->seg_ninv_wq} : Twiddle<T>{f.s.ninv_w, f.s.ninv_wq};
            // forward: the strided pass opens a cyclic transform; inverse: it ends one (its last round folds n^-1 into the twiddles)
            const bool triv = !f.s.plus && (SS::INV ? f.s.last : f.s.first) && (f.s.lo + SS::D == n) && tabS[0] == T(1);
            if (doS) build_twiddles<SS>(twS, tabS, 0, n, f.s.n_tw, f.s.lo, f.s.plus, seg_p, seg_mu, seg_pbits, tid, kFusedConsumers, SS::INV && f.s.last, ninv);
            if (doC) build_twiddles<SC>(twC, tabC, range, n, f.c.n_tw, 0, f.c.plus, seg_p, seg_mu, seg_pbits, tid, kFusedConsumers);
            asm volatile("bar.sync 3, %0;" ::"n"(kFusedConsumers) : "memory");
            if (tid == 0) TL(4);
            // tile claims run one ahead: the leader takes the NEXT index before the group starts on the current tile, so the
            // shared-memory atomic and its broadcast are off the critical path
            if (ctid == 0) ctl->bcast[g][0] = atomicAdd(&ctl->next_t, 1);
            consumer_sync(1 + g);
            int t = ctl->bcast[g][0];
            for (int it = 0; t < total; it++)
            {
                if (ctid == 0) ctl->bcast[g][(it + 1) & 1] = atomicAdd(&ctl->next_t, 1);
                const int b = t % NB;
                unsigned char* buf = bufs + b * TILE;
                mbar_wait(smem_u32(&ctl->full[b]), (unsigned) (t / NB) & 1u); // tile landed
                if (t == 0 && ctid == 0) TL(5);
                if (ctl->kind[b] == 0)
                    tile_rounds<SS, false>(buf, twS, twS + SS::TW1, twS + SS::TW1 + SS::TW2, MS, ctid, ninv, nullptr, f.s, triv, 1 + g);
                else
                    tile_rounds<SC, false>(buf, twC, twC + SC::TW1, twC + SC::TW1 + SC::TW2, MC, ctid, ninv, nullptr, f.c, false, 1 + g);
                fence_async(); // make the generic-proxy writes visible to the bulk store
                if (t == 0 && ctid == 0) TL(6);
                mbar_arrive(smem_u32(&ctl->done[b]));
                consumer_sync(1 + g);
                t = ctl->bcast[g][(it + 1) & 1];
            }
        }
        // ---- the counters go back to zero: the last CTA to finish (every other CTA has made all its observations) clears them
        __syncthreads();
        if (tid == 0) TL(11);
        if (tid == 0)
        {
            __threadfence();
            const unsigned old = atomicAdd(f.counters + f.ticket_off, 1u);
            ctl->is_last = (old == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (ctl->is_last)
        {
            for (long long i = tid; i < (long long) batch * mc; i += kFusedThreads) f.counters[i] = 0u;
            if (tid == 0) f.counters[f.ticket_off] = 0u;
        }
        if (tid == 0) TL(12);
    }

    template <typename SS, typename SC>
    __global__ void __launch_bounds__(kFusedThreads, 1)
        fused2_kernel(const FusedArgs<typename SS::T> f, const __grid_constant__ CUtensorMap mapA_in, const __grid_constant__ CUtensorMap mapA_out,
                      const __grid_constant__ CUtensorMap mapB)
    {
        fused2_body<SS, SC, false>(f, mapA_in, mapA_out, mapB);
    }

    // RNS calls: the moduli live on the device, so the arithmetic policy is chosen HERE, per CTA, from the CTA's own modulus
    // (SSL / SCL: the lazy-policy pair of passes, SSX / SCX: the exact pair; both bodies are in the kernel).
    template <typename SSL, typename SCL, typename SSX, typename SCX>
    __global__ void __launch_bounds__(kFusedThreads, 1)
        fused2_rns_kernel(const FusedArgs<typename SSL::T> f, const __grid_constant__ CUtensorMap mapA_in,
                          const __grid_constant__ CUtensorMap mapA_out, const __grid_constant__ CUtensorMap mapB)
    {
        using T = typename SSL::T;
        static_assert(FusedSmem<SSL, SCL>::BYTES == FusedSmem<SSX, SCX>::BYTES, "same pass shapes, two policies");
        if constexpr (std::is_same<SSL, SSX>::value)
            fused2_body<SSL, SCL, true>(f, mapA_in, mapA_out, mapB);
        else
        {
            const int mslot = (int) blockIdx.x / f.g_slot;
            const T p = f.s.mod_dev[3 * (f.s.mod_order ? f.s.mod_order[mslot] : mslot)];
            const bool lazy_ok = SSL::INV ? ((uint64_t) p < kFastModulusLimit) : ((uint64_t) p >= kF60ModulusMin && (uint64_t) p < kF60ModulusLimit);
            if (lazy_ok)
                fused2_body<SSL, SCL, true>(f, mapA_in, mapA_out, mapB);
            else
                fused2_body<SSX, SCX, true>(f, mapA_in, mapA_out, mapB);
        }
    }

    static std::atomic<int> g_fused_lag_steps{4};
    void fused_set_lag_steps(int v) { g_fused_lag_steps.store(v < 0 ? 0 : v); }

    // in / out / table / p / ninv / mu / pbits / n / plus / batch / in_bound of `a` are filled in; lo_s = row stride (log2) of the
    // strided pass.  Returns cudaErrorNotSupported when this call cannot take the fused kernel (the caller launches the
    // two passes separately).
    static std::atomic<int> g_fused_policy{1};
    void fused_set_policy(int v) { g_fused_policy.store(v); }

    // SSX / SCX = void: single modulus.  Otherwise the RNS form (a.mod_count slots, a.batch polynomials PER SLOT, moduli on the
    // device): SSX / SCX are the exact-policy twins of SS / SC (the same types when there is only one policy).
    template <typename SS, typename SC, typename SSX = void, typename SCX = void>
    static cudaError_t launch_fused(const FastArgs<typename SS::T>& a, int lo_s, bool inverse, unsigned* counters, cudaStream_t st,
                                    void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        using T = typename SS::T;
        constexpr bool RNS = !std::is_void<SSX>::value;
        constexpr int kMaxDev = 64;
        static std::atomic<int> cached_bps[kMaxDev];
        static std::atomic<int> cached_sms[kMaxDev];
        constexpr int SMEM = FusedSmem<SS, SC>::BYTES;
        using KernT = void (*)(const FusedArgs<T>, const CUtensorMap, const CUtensorMap, const CUtensorMap);
        KernT kern;
        if constexpr (RNS)
            kern = fused2_rns_kernel<SS, SC, SSX, SCX>;
        else
            kern = fused2_kernel<SS, SC>;
        const int mc = RNS ? a.mod_count : 1;
        if (RNS && (mc < 1 || a.poly_order != nullptr)) return cudaErrorNotSupported;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDev) return cudaErrorNotSupported;
        int bps = cached_bps[dev].load(std::memory_order_acquire), sms = cached_sms[dev].load(std::memory_order_acquire);
        if (bps <= 0 || sms <= 0)
        {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            if (e != cudaSuccess) return e;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int b = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kFusedThreads, SMEM);
            if (e != cudaSuccess) return e;
            if (b < 1) return cudaErrorNotSupported; // (the flag protocol needs every CTA resident)
            bps = 1; // one CTA per SM by design
            cached_sms[dev].store(sms, std::memory_order_release);
            cached_bps[dev].store(bps, std::memory_order_release);
        }
        const int n = a.n, batch = a.batch;
        if (lo_s + SS::D != n || lo_s < SS::C) return cudaErrorNotSupported; // one strided range (two-pass plans)
        const int tpp_log = lo_s - SS::C;
        const int nranges = 1 << (n - SC::KC);
        if (n < SC::KC || tpp_log > 14 || nranges > 16384) return cudaErrorNotSupported;
        const long long slots = ((long long) sms * bps) / mc; // CTAs per modulus slot
        if (nranges > slots) return cudaErrorNotSupported;
        const long long rows = ((long long) batch * mc) << (n - lo_s);
        if (rows >= (1LL << 31)) return cudaErrorNotSupported;

        FusedArgs<T> f{};
        f.s = a;
        f.s.lo = lo_s;
        f.s.first = inverse ? 0 : 1;
        f.s.last = inverse ? 1 : 0;
        f.c = a;
        f.c.lo = 0;
        f.c.first = inverse ? 1 : 0;
        f.c.last = inverse ? 0 : 1;
        f.c.in_bound = 1;
        f.counters = counters;
        f.ticket_off = (long long) batch * mc; // (the workspace holds at least that many words + 1)
        f.fwd = inverse ? 0 : 1;
        f.tpp_log = tpp_log;
        f.nranges = nranges;
        f.ngroups = (batch + (1 << SC::NPLOG) - 1) >> SC::NPLOG;
        const long long n_str = (long long) batch << tpp_log;
        // Measured (profiles/r2_fused_ab.jsonl): the small 32-bit rings gain at every batch size (the data crosses HBM once
        // and their strided pass is bandwidth-bound on its own); the 64-bit transforms are bound by the integer multiplier,
        // so beyond 2^13 the fused kernel only wins while the call is launch-bound (a few tiles per CTA).
        // Round-2 numbers at saturating batch (unfused / fused time): 32-bit 2^13 1.31, 2^14 1.26, 2^15 1.17, 2^16 1.10,
        // 2^17 0.98, 2^18 0.98; 64-bit 2^12 1.09, 2^13 1.01, 2^14 0.98, 2^15 0.98, 2^16 0.90; small batches 1.1-1.5 everywhere.
        if (g_fused_policy.load() == 1)
        {
            const bool always = !RNS && (sizeof(T) == 8 ? n <= 13 : n <= 16);
            if (!always && n_str > 8 * slots) return cudaErrorNotSupported;
        }
        long long grid;
        const long long con_each = (slots - n_str) / nranges; // CTAs per range left over when every strided tile has its own CTA
        if (n_str <= slots / 2 && con_each >= 1)
        {
            // small batch: separate CTAs per pass (the contiguous CTAs build their twiddles while the strided ones compute)
            long long k = con_each < f.ngroups ? con_each : f.ngroups;
            f.g_str = (int) n_str;
            f.c_off = (int) n_str;
            f.con_k = (int) k;
            f.con_extra = 0;
            grid = n_str + k * nranges;
            f.lag = 1 << SC::NPLOG;
        }
        else
        {
            grid = slots;
            f.g_str = (int) grid;
            f.c_off = 0;
            f.con_k = (int) (grid / nranges);
            f.con_extra = (int) (grid % nranges);
            // polynomials the whole grid moves through one pass per tile time
            const long long per_step = (grid << (SS::K)) >> n;
            long long lag = g_fused_lag_steps.load() * (per_step > 0 ? per_step : 1);
            if (lag < (1 << SC::NPLOG)) lag = 1 << SC::NPLOG;
            if (lag > 0x3fffffff) lag = 0x3fffffff;
            f.lag = (int) lag;
        }
        f.g_slot = (int) grid;
        grid *= mc;
        const int mm = RNS ? mc : 0; // (make_map: 0 = single-modulus views)
        alignas(64) CUtensorMap mA_in, mA_out, mB;
        if (!inverse)
        {
            if (!make_map<SS>(&mA_in, a.in, n, lo_s, batch, mm)) return cudaErrorNotSupported;
            if (a.in == a.out)
                mA_out = mA_in;
            else if (!make_map<SS>(&mA_out, a.out, n, lo_s, batch, mm))
                return cudaErrorNotSupported;
            if (!make_map<SC>(&mB, a.out, n, 0, batch, mm)) return cudaErrorNotSupported;
        }
        else
        {
            if (!make_map<SC>(&mA_in, a.in, n, 0, batch, mm)) return cudaErrorNotSupported;
            if (a.in == a.out)
                mA_out = mA_in;
            else if (!make_map<SC>(&mA_out, a.out, n, 0, batch, mm))
                return cudaErrorNotSupported;
            if (!make_map<SS>(&mB, a.out, n, lo_s, batch, mm)) return cudaErrorNotSupported;
        }
        prof_begin(1, st);
        kern<<<(unsigned) grid, kFusedThreads, SMEM, st>>>(f, mA_in, mA_out, mB);
        prof_end(st);
        return cudaGetLastError();
    }

    // Launch-bound calls (64-bit): what such a call costs is the critical path through ONE tile of each pass -- tile load, two
    // register rounds of 32 butterflies per thread, store (profiles/r2_fused_timeline.txt: 5.2 of the 10.5 us of a 2^12 x 8 call
    // are the rounds of one 4096-element tile per CTA, bound by the SM's own multiplier pipes, while 130 SMs idle).  Calls of
    // at most g_small_tile_elems (2^18) elements therefore run on 1024-element tiles (K = 10): four times the CTAs, a quarter of the
    // arithmetic on each tile's critical path.  Rings 2^12 .. 2^14 (a strided tile needs 2^(10 - d) >= 16 adjacent elements).
    static std::atomic<long long> g_small_tile_elems{1LL << 18};
    void fused_set_small_tile_elems(long long v) { g_small_tile_elems.store(v < 0 ? 0 : v); }
    long long fused_small_tile_elems() { return g_small_tile_elems.load(); }
    static bool small_tiles(long long polys, int n, int d) { return d >= 4 && d <= 6 && (polys << n) <= g_small_tile_elems.load(); }

    // Two-pass plans in one launch.  Returns cudaErrorNotSupported when the shape / modulus is not covered.
    template <typename T>
    cudaError_t fused_merge(const FastArgs<T>& a, const FastPlan& pl, bool inverse, bool f60_or_l32, bool lazy_inv, unsigned* counters,
                            cudaStream_t st, void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        if (pl.npass != 2 || !pl.strided[0] || pl.strided[1]) return cudaErrorNotSupported;
        const int d = pl.d[0], lo = pl.lo[0];
        if constexpr (sizeof(T) == 8)
        {
            using Cf = Shape<T, false, 2, false, 4, 4, 12, 1>;
            using Ci = Shape<T, true, 1, false, 4, 4, 12, 1>;
            if (small_tiles(a.batch, a.n, d) && (inverse ? lazy_inv : f60_or_l32))
            {
                using Cfs = Shape<T, false, 2, false, 4, 4, 10, 0>;
                using Cis = Shape<T, true, 1, false, 4, 4, 10, 0>;
                if (!inverse)
                    switch (d)
                    {
                        case 4: return launch_fused<Shape<T, false, 2, true, 4, 0, 10, 0>, Cfs>(a, lo, false, counters, st, prof_begin, prof_end);
                        case 5: return launch_fused<Shape<T, false, 2, true, 3, 2, 10, 0>, Cfs>(a, lo, false, counters, st, prof_begin, prof_end);
                        default: return launch_fused<Shape<T, false, 2, true, 3, 3, 10, 0>, Cfs>(a, lo, false, counters, st, prof_begin, prof_end);
                    }
                switch (d)
                {
                    case 4: return launch_fused<Shape<T, true, 1, true, 4, 0, 10, 0>, Cis>(a, lo, true, counters, st, prof_begin, prof_end);
                    case 5: return launch_fused<Shape<T, true, 1, true, 3, 2, 10, 0>, Cis>(a, lo, true, counters, st, prof_begin, prof_end);
                    default: return launch_fused<Shape<T, true, 1, true, 3, 3, 10, 0>, Cis>(a, lo, true, counters, st, prof_begin, prof_end);
                }
            }
            if (!inverse)
            {
                if (!f60_or_l32) return cudaErrorNotSupported;
                switch (d)
                {
                    case 4: return launch_fused<Shape<T, false, 2, true, 4, 0, 12, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 5: return launch_fused<Shape<T, false, 2, true, 3, 2, 12, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 6: return launch_fused<Shape<T, false, 2, true, 3, 3, 12, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 7: return launch_fused<Shape<T, false, 2, true, 4, 3, 12, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 8: return launch_fused<Shape<T, false, 2, true, 4, 4, 12, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    default: return cudaErrorNotSupported;
                }
            }
            if (!lazy_inv) return cudaErrorNotSupported;
            switch (d)
            {
                case 4: return launch_fused<Shape<T, true, 1, true, 4, 0, 12, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 5: return launch_fused<Shape<T, true, 1, true, 3, 2, 12, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 6: return launch_fused<Shape<T, true, 1, true, 3, 3, 12, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 7: return launch_fused<Shape<T, true, 1, true, 4, 3, 12, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 8: return launch_fused<Shape<T, true, 1, true, 4, 4, 12, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                default: return cudaErrorNotSupported;
            }
        }
        else
        {
            using Cl = Shape<T, false, 2, false, 5, 5, 13, 1>;
            using Cf = Shape<T, false, 0, false, 5, 5, 13, 1>;
            using Ci = Shape<T, true, 0, false, 5, 5, 13, 1>;
            if (!inverse && f60_or_l32)
            {
                switch (d)
                {
                    case 3: return launch_fused<Shape<T, false, 2, true, 3, 0, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 4: return launch_fused<Shape<T, false, 2, true, 4, 0, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 5: return launch_fused<Shape<T, false, 2, true, 5, 0, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 6: return launch_fused<Shape<T, false, 2, true, 3, 3, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 7: return launch_fused<Shape<T, false, 2, true, 4, 3, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 8: return launch_fused<Shape<T, false, 2, true, 4, 4, 13, 0>, Cl>(a, lo, false, counters, st, prof_begin, prof_end);
                    default: return cudaErrorNotSupported;
                }
            }
            if (!inverse)
            {
                switch (d)
                {
                    case 3: return launch_fused<Shape<T, false, 0, true, 3, 0, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 4: return launch_fused<Shape<T, false, 0, true, 4, 0, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 5: return launch_fused<Shape<T, false, 0, true, 5, 0, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 6: return launch_fused<Shape<T, false, 0, true, 3, 3, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 7: return launch_fused<Shape<T, false, 0, true, 4, 3, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    case 8: return launch_fused<Shape<T, false, 0, true, 4, 4, 13, 0>, Cf>(a, lo, false, counters, st, prof_begin, prof_end);
                    default: return cudaErrorNotSupported;
                }
            }
            switch (d)
            {
                case 3: return launch_fused<Shape<T, true, 0, true, 3, 0, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 4: return launch_fused<Shape<T, true, 0, true, 4, 0, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 5: return launch_fused<Shape<T, true, 0, true, 5, 0, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 6: return launch_fused<Shape<T, true, 0, true, 3, 3, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 7: return launch_fused<Shape<T, true, 0, true, 4, 3, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                case 8: return launch_fused<Shape<T, true, 0, true, 4, 4, 13, 0>, Ci>(a, lo, true, counters, st, prof_begin, prof_end);
                default: return cudaErrorNotSupported;
            }
        }
    }
    // RNS form (GPU_NTT / GPU_INTT with Modulus*, GPU_NTT_Modulus_Ordered): a.batch = polynomials per slot, a.mod_count,
    // a.mod_dev, a.ninv_dev, a.mod_order set; the moduli are device data, so 64-bit kernels carry the lazy and the exact pair
    // of passes and every CTA picks from its own modulus.
    template <typename T>
    cudaError_t fused_merge_rns(const FastArgs<T>& a, const FastPlan& pl, bool inverse, unsigned* counters, cudaStream_t st,
                                void (*prof_begin)(int, cudaStream_t), void (*prof_end)(cudaStream_t))
    {
        if (pl.npass != 2 || !pl.strided[0] || pl.strided[1]) return cudaErrorNotSupported;
        const int d = pl.d[0], lo = pl.lo[0];
        if constexpr (sizeof(T) == 8)
        {
            using Cf = Shape<T, false, 2, false, 4, 4, 12, 1>;
            using Cfx = Shape<T, false, 0, false, 4, 4, 12, 1>;
            using Ci = Shape<T, true, 1, false, 4, 4, 12, 1>;
            using Cix = Shape<T, true, 0, false, 4, 4, 12, 1>;
            if (small_tiles((long long) a.batch * a.mod_count, a.n, d))
            {
                using Cfs = Shape<T, false, 2, false, 4, 4, 10, 0>;
                using Cfxs = Shape<T, false, 0, false, 4, 4, 10, 0>;
                using Cis = Shape<T, true, 1, false, 4, 4, 10, 0>;
                using Cixs = Shape<T, true, 0, false, 4, 4, 10, 0>;
#define GPUNTT_FUSED_RNS64S(R1, R2)                                                                                                              \
    (inverse ? launch_fused<Shape<T, true, 1, true, R1, R2, 10, 0>, Cis, Shape<T, true, 0, true, R1, R2, 10, 0>, Cixs>(a, lo, true, counters, st,    \
                                                                                                                      prof_begin, prof_end)       \
             : launch_fused<Shape<T, false, 2, true, R1, R2, 10, 0>, Cfs, Shape<T, false, 0, true, R1, R2, 10, 0>, Cfxs>(a, lo, false, counters, st, \
                                                                                                                        prof_begin, prof_end))
                switch (d)
                {
                    case 4: return GPUNTT_FUSED_RNS64S(4, 0);
                    case 5: return GPUNTT_FUSED_RNS64S(3, 2);
                    default: return GPUNTT_FUSED_RNS64S(3, 3);
                }
#undef GPUNTT_FUSED_RNS64S
            }
#define GPUNTT_FUSED_RNS64(R1, R2)                                                                                                               \
    (inverse ? launch_fused<Shape<T, true, 1, true, R1, R2, 12, 0>, Ci, Shape<T, true, 0, true, R1, R2, 12, 0>, Cix>(a, lo, true, counters, st,      \
                                                                                                                    prof_begin, prof_end)         \
             : launch_fused<Shape<T, false, 2, true, R1, R2, 12, 0>, Cf, Shape<T, false, 0, true, R1, R2, 12, 0>, Cfx>(a, lo, false, counters, st,   \
                                                                                                                      prof_begin, prof_end))
            switch (d)
            {
                case 4: return GPUNTT_FUSED_RNS64(4, 0);
                case 5: return GPUNTT_FUSED_RNS64(3, 2);
                case 6: return GPUNTT_FUSED_RNS64(3, 3);
                case 7: return GPUNTT_FUSED_RNS64(4, 3);
                case 8: return GPUNTT_FUSED_RNS64(4, 4);
                default: return cudaErrorNotSupported;
            }
#undef GPUNTT_FUSED_RNS64
        }
        else
        {
            using Cf = Shape<T, false, 0, false, 5, 5, 13, 1>;
            using Ci = Shape<T, true, 0, false, 5, 5, 13, 1>;
#define GPUNTT_FUSED_RNS32(R1, R2)                                                                                                               \
    (inverse ? launch_fused<Shape<T, true, 0, true, R1, R2, 13, 0>, Ci, Shape<T, true, 0, true, R1, R2, 13, 0>, Ci>(a, lo, true, counters, st,       \
                                                                                                                   prof_begin, prof_end)          \
             : launch_fused<Shape<T, false, 0, true, R1, R2, 13, 0>, Cf, Shape<T, false, 0, true, R1, R2, 13, 0>, Cf>(a, lo, false, counters, st,    \
                                                                                                                     prof_begin, prof_end))
            switch (d)
            {
                case 4: return GPUNTT_FUSED_RNS32(4, 0);
                case 5: return GPUNTT_FUSED_RNS32(5, 0);
                case 6: return GPUNTT_FUSED_RNS32(3, 3);
                case 7: return GPUNTT_FUSED_RNS32(4, 3);
                case 8: return GPUNTT_FUSED_RNS32(4, 4);
                default: return cudaErrorNotSupported;
            }
#undef GPUNTT_FUSED_RNS32
        }
    }
    template cudaError_t fused_merge_rns<uint64_t>(const FastArgs<uint64_t>&, const FastPlan&, bool, unsigned*, cudaStream_t,
                                                   void (*)(int, cudaStream_t), void (*)(cudaStream_t));
    template cudaError_t fused_merge_rns<uint32_t>(const FastArgs<uint32_t>&, const FastPlan&, bool, unsigned*, cudaStream_t,
                                                   void (*)(int, cudaStream_t), void (*)(cudaStream_t));

#ifdef GPUNTT_TIMELINE
    extern "C" int gpuntt_b200_timeline_read(long long* out, int clear)
    {
        if (cudaDeviceSynchronize() != cudaSuccess) return -1;
        if (cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline)) != cudaSuccess) return -1;
        if (clear)
        {
            static long long zeros[512][16];
            cudaMemcpyToSymbol(g_timeline, zeros, sizeof(zeros));
        }
        return 512 * 16;
    }
#endif

    template cudaError_t fused_merge<uint64_t>(const FastArgs<uint64_t>&, const FastPlan&, bool, bool, bool, unsigned*, cudaStream_t,
                                               void (*)(int, cudaStream_t), void (*)(cudaStream_t));
    template cudaError_t fused_merge<uint32_t>(const FastArgs<uint32_t>&, const FastPlan&, bool, bool, bool, unsigned*, cudaStream_t,
                                               void (*)(int, cudaStream_t), void (*)(cudaStream_t));

} // namespace gpuntt_b200
