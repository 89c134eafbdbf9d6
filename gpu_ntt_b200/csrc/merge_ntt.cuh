// gpu_ntt_b200/csrc/merge_ntt.cuh -- launch-plan structures shared by the Merge-NTT kernels
// and their host dispatcher.  B200-native replacement for the reference's KernelConfig tables
// (ntt.cuh:53-67, 606-797 in the reference tree); nothing here is derived from those tables.
#pragma once
#include <cstdint>

namespace gpuntt_b200
{

    constexpr int kMaxRounds = 6;
    constexpr int kThreads = 256;

    // One pass = one kernel launch = every tile makes one HBM round trip.
    // A tile holds 2^tile_log elements in shared memory:
    //   strided tile  (lo > 0): 2^d rows (row stride 2^lo elements) x 2^c adjacent columns,
    //                           local index l = (row << c) | col
    //   contiguous    (lo == 0, c == 0): 2^tile_log adjacent elements (>= 1 whole sub-transform;
    //                           several polynomials when n_power < tile_log)
    // The pass applies the butterfly stages that act on global index bits [lo, lo+d), split
    // into `nrounds` register rounds of round_bits[i] stages; rounds are listed from the
    // HIGHEST bits to the lowest (forward order); the inverse transform walks them backwards.
    struct PassPlan
    {
        int tile_log;
        int lo;
        int d;
        int c;
        int nrounds;
        int round_bits[kMaxRounds];
        int first; // this launch reads the caller's input (signed-input fix-up happens here)
        int last;  // this launch writes final results (canonical form, n^-1, signed output)
    };

    struct MergePlan
    {
        int npasses;
        PassPlan pass[3]; // forward order; inverse executes pass[npasses-1] first
    };

    // element_bits = 32 or 64
    MergePlan make_merge_plan(int n_power, int element_bits);

    template <typename T> struct PassArgs
    {
        const void* in; // T or signed T
        T* out;         // T (or signed T on the last inverse pass)
        const void* tw; // Twiddle<T>[mod_slices << tw_stride_log]  (w, w') pairs, caller's index order
        const T* mod_values; // RNS: device array of p (stride 3 elements = Modulus<T>), else nullptr
        const void* ninv_tw; // RNS inverse: Twiddle<T>[mod_count] for n^-1
        T p;                 // single modulus
        T ninv_w, ninv_wq;   // single modulus, inverse
        int n_power;
        int mod_count;       // 0 => single modulus passed by value
        int tw_stride_log;   // log2 of the per-modulus slice stride in the twiddle table
        int plus;            // reduction polynomial X^N+1 (table index m+i) vs X^N-1 (index i)
        int signed_io;       // first forward pass: signed input; last inverse pass: centred output
        long long total_elems; // batch << n_power
        PassPlan plan;
    };

} // namespace gpuntt_b200
