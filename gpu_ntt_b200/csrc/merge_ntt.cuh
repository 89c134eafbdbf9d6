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
    // one strided pass over index bits [lo, lo+d) (d + column bits <= 13 / 14)
    PassPlan make_strided_pass(int lo, int d, int element_bits);

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
        // RNS indirection (GPU_NTT_Modulus_Ordered / GPU_NTT_Poly_Ordered of the reference): device arrays or null
        const int* mod_order;  // modulus / table slice of polynomial b = mod_order[b % mod_count]
        const int* poly_order; // the b-th transform lives in polynomial slot poly_order[b]
        int batch;
        int mod_shift;         // RNS: transform b belongs to modulus group (b >> mod_shift) % mod_count (4-step row phases)
        int col_log;           // NTTLayout::PerCoefficient: > 0 = the array is ONE [2^n][2^col_log] matrix transformed along
                               // its columns; the modulus group of an element is (column % mod_count)
        int shared_tables;     // RNS: every modulus reads the same (un-offset) table (the reference's 4-step convention)
        // 4-step twiddle matrix, plain residues (no Shoup companion): multiplied with a Barrett reduction
        //   w_mode 1: at store, by w_table[offset in polynomial]            (forward: after the column transforms)
        //   w_mode 2: at load,  by w_table[(offset & (2^w_lo - 1)) << w_hi | offset >> w_lo]   (inverse: transposed index)
        const T* w_table;
        int w_mode, w_lo, w_hi;
        T bar_bit, bar_mu;     // single modulus: Modulus<T>::bit / mu  (RNS: read from mod_values[3*m + 1], [3*m + 2])
        PassPlan plan;
    };

} // namespace gpuntt_b200
