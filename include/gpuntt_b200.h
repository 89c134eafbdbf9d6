/*
 * gpuntt_b200.h -- C ABI of the B200-native (sm_100a) batched NTT/INTT engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point replaces one host entry point of Alisah-Ozcan/GPU-NTT (citations are file:line in the
 * reference tree); the C++17 headers under include/gpuntt/ re-create the reference's template
 * API (GPU_NTT / GPU_INTT / GPU_NTT_Inplace / GPU_INTT_Inplace, ntt_configuration,
 * NTTParameters ...) as thin forwarders onto these functions, so reference callers recompile
 * unchanged; see INTEGRATION.md.
 *
 * Conventions (identical to the reference):
 *  - all data/table/modulus pointers are DEVICE pointers owned by the caller;
 *  - polynomials are rows of a row-major [batch][N] array, N = 1 << n_power;
 *  - forward: natural-order coefficients in, bit-reversed-order evaluations out; inverse: the
 *    opposite, including the multiplication by N^-1 (mod_inverse);
 *  - root_of_unity_table is the BIT-REVERSED power table the reference's
 *    NTTParameters::gpu_root_of_unity_table_generator produces (nttparameters.cu:175-189):
 *    N/2 powers of omega for X^N-1, N powers of psi for X^N+1;
 *  - RNS form: polynomial b uses modulus[b % mod_count], table slice starting at element
 *    (b % mod_count) << n_power, mod_inverse[b % mod_count] (ntt.cu:613-619,672-673);
 *  - calls only enqueue work on `stream` and return; they never synchronise -- with two documented exceptions:
 *    a cached scratch buffer that has to GROW is re-allocated after a stream synchronisation, and the first
 *    gpuntt_b200_4step_ntt call per (device, modulus pointer) in the one-device-modulus RNS form reads that modulus
 *    back (see GPUNTT_B200_TUNE_4STEP_MODULUS_CACHE);
 *  - in == out is allowed (that is all the reference's *_Inplace entry points do).
 *
 * Unlike the reference the engine keeps one small device scratch buffer per (device, stream)
 * for the per-call twiddle companion table; it is allocated on first use,
 * grown on demand and released by gpuntt_b200_release_workspaces().
 *
 * There is no CPU fallback: every function fails with GPUNTT_B200_ERR_CUDA if no usable
 * sm_100 device/context exists.
 */
#ifndef GPUNTT_B200_H
#define GPUNTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUNTT_B200_VERSION 100 /* 0.1.0 */

typedef enum gpuntt_b200_status
{
    GPUNTT_B200_OK = 0,
    GPUNTT_B200_ERR_N_POWER = 1,     /* "Invalid n_power range!"  (ntt.cu:2088-2091) */
    GPUNTT_B200_ERR_LAYOUT = 2,      /* "Invalid ntt_layout!"     (ntt.cu:2253)      */
    GPUNTT_B200_ERR_CUDA = 3,        /* a CUDA runtime call / launch failed (GPUNTT_CUDA_CHECK) */
    GPUNTT_B200_ERR_ARGUMENT = 4,    /* null pointer, negative batch, bad enum ...   */
    GPUNTT_B200_ERR_UNSUPPORTED = 5  /* valid in the reference but not built yet     */
} gpuntt_b200_status;

/* enum values equal the reference's (nttparameters.cuh:19-36) */
enum { GPUNTT_B200_FORWARD = 0, GPUNTT_B200_INVERSE = 1 };
enum { GPUNTT_B200_PER_POLYNOMIAL = 0, GPUNTT_B200_PER_COEFFICIENT = 1 };
enum { GPUNTT_B200_X_N_PLUS = 0, GPUNTT_B200_X_N_MINUS = 1 };

/* Same layout as the reference's Modulus<Data64> / Modulus<Data32> (modular_arith.cuh:28-57):
 * value = p, bit = bit length of p, mu = floor(2^(2*bit+1) / p).  Only `value` is read by this
 * engine (it derives its own reduction constants); bit/mu are carried for ABI compatibility. */
typedef struct gpuntt_b200_modulus64 { uint64_t value, bit, mu; } gpuntt_b200_modulus64;
typedef struct gpuntt_b200_modulus32 { uint32_t value, bit, mu; } gpuntt_b200_modulus32;

/* One Merge-NTT call.  Replaces, depending on the fields,
 *   GPU_NTT  single modulus (ntt.cu:2076-2256)   GPU_NTT  RNS (ntt.cu:2560-2746)
 *   GPU_INTT single modulus (ntt.cu:2258-2558)   GPU_INTT RNS (ntt.cu:2748-3058)
 *   GPU_NTT_Inplace / GPU_INTT_Inplace (ntt.cu:3060-3097): pass in == out.
 * The ntt_configuration / ntt_rns_configuration fields (ntt.cuh:31-51) map 1:1:
 *   n_power, ntt_layout, reduction_poly, mod_inverse(_dev), stream; `direction` is the function
 *   name in the reference (its cfg.ntt_type is ignored there); zero_padding is never read by the
 *   reference and has no field here. */
typedef struct gpuntt_b200_merge_desc
{
    int element_bits;     /* 32 (Data32/Data32s) or 64 (Data64/Data64s)                       */
    int is_signed;        /* T = Data32s/Data64s: signed INPUT on forward (reduced into [0,p)
                             on load, ntt.cu:481-489), centred signed OUTPUT on inverse
                             (ntt.cu:1178-1186)                                               */
    int direction;        /* GPUNTT_B200_FORWARD / GPUNTT_B200_INVERSE                        */
    int n_power;          /* 1..28                                                            */
    int ntt_layout;       /* GPUNTT_B200_PER_POLYNOMIAL (PER_COEFFICIENT: see status codes)   */
    int reduction_poly;   /* GPUNTT_B200_X_N_PLUS / GPUNTT_B200_X_N_MINUS                     */
    int batch_size;       /* number of polynomials                                           */
    int mod_count;        /* 0: single modulus by value; >= 1: RNS form, arrays on device     */
    const void* in;       /* device, [batch][N] elements                                      */
    void* out;            /* device, [batch][N] elements (may equal in)                       */
    const void* root_of_unity_table; /* device, bit-reversed order (see above)                */
    uint64_t modulus_value;          /* single-modulus form: p                                */
    uint64_t mod_inverse_value;      /* single-modulus inverse: N^-1 mod p                    */
    const void* modulus_dev;         /* RNS: device array of gpuntt_b200_modulus{32,64}       */
    const void* mod_inverse_dev;     /* RNS inverse: device array of N^-1 mod p_i (elements)  */
    void* stream;                    /* cudaStream_t                                          */
    /* RNS indirection (GPU_NTT_Modulus_Ordered / GPU_NTT_Poly_Ordered, ntt.cu:3600-3776, 4281-4458); both may
     * be NULL.  modulus_order_dev: device int[mod_count], polynomial b uses modulus / table slice
     * order[b % mod_count].  poly_order_dev: device int[batch_size], the b-th transform reads and writes
     * the polynomial slot order[b] (of `in` and of `out`) with modulus index b % mod_count.             */
    const int* modulus_order_dev;
    const int* poly_order_dev;
} gpuntt_b200_merge_desc;

/* Enqueue one batched Merge-NTT / INTT. Returns a gpuntt_b200_status. */
int gpuntt_b200_merge_ntt(const gpuntt_b200_merge_desc* desc);

/* One 4-step (large ring, n_power 12..24) transform.  Replaces GPU_4STEP_NTT single modulus / RNS
 * (ntt_4step.cu:2767-3232 / 2293-2765) and ntt4step_(rns_)configuration (ntt_4step.cuh:19-33).  N = n1 * n2 with
 * the reference's shapes (gpuntt_b200_4step_shape); tables exactly as NTTParameters4Step produces them and
 * its examples upload them (test_4step_ntt.cu:90-110): n1_table / n2_table = bit-reversed n1/2- and
 * n2/2-entry power tables, w_table = the natural-layout N-entry twiddle matrix (forward:
 * W_root_of_unity_table, inverse: W_inverse_root_of_unity_table and the inverse small tables).
 * io_contract:
 *   GPUNTT_B200_4STEP_REFERENCE  the reference's: forward takes the n2 x n1 matrix GPU_Transpose(x, row=n1,
 *        col=n2) made and leaves the n1 x n2 matrix R whose transpose is NTT_4STEP_CPU::ntt(x); inverse
 *        takes NTT_4STEP_CPU::intt_first_transpose(y) and leaves the n1 x n2 matrix whose transpose is
 *        NTT_4STEP_CPU::intt(y).  Out of place.
 *   GPUNTT_B200_4STEP_FUSED      natural x in, NTT_4STEP_CPU::ntt(x) out / y in, NTT_4STEP_CPU::intt(y) out;
 *        in == out allowed.
 * RNS form (mod_count >= 1): polynomial b uses modulus_dev[b % mod_count]; like the reference's kernels all
 * moduli index the SAME tables (ntt_4step.cu:150-229), which is only meaningful for mod_count == 1 or moduli
 * sharing their roots.  Every call on the tuned kernels (64-bit, one modulus) is a (W, W') pair table + three data passes and no
 * transpose kernel: where a contract asks for a transposed layout a pass STORES transposed.  They use an engine-owned scratch
 * buffer of batch_size * N elements and a pair table of 16 * N bytes (cached per device and stream, see
 * gpuntt_b200_release_workspaces); a reference-contract forward call of fewer than four polynomials needs neither. */
enum { GPUNTT_B200_4STEP_REFERENCE = 0, GPUNTT_B200_4STEP_FUSED = 1 };
typedef struct gpuntt_b200_4step_desc
{
    int element_bits;   /* 32 or 64 */
    int direction;      /* GPUNTT_B200_FORWARD / GPUNTT_B200_INVERSE (cfg.ntt_type) */
    int n_power;        /* 12..24 */
    int batch_size;
    int mod_count;      /* 0: single modulus by value; >= 1: RNS form */
    int io_contract;    /* GPUNTT_B200_4STEP_REFERENCE / GPUNTT_B200_4STEP_FUSED */
    const void* in;     /* device, [batch][N] */
    void* out;          /* device, [batch][N] */
    const void* n1_table;
    const void* n2_table;
    const void* w_table;
    uint64_t modulus_value;
    uint64_t mod_inverse_value;  /* inverse: N^-1 mod p */
    const void* modulus_dev;     /* RNS: device array of gpuntt_b200_modulus{32,64} (bit and mu ARE read here) */
    const void* mod_inverse_dev; /* RNS inverse: device array of N^-1 mod p_i */
    void* stream;
} gpuntt_b200_4step_desc;
int gpuntt_b200_4step_ntt(const gpuntt_b200_4step_desc* desc);
/* n1, n2 of the reference's matrix_dimention() (nttparameters.cu:305-354); GPUNTT_B200_ERR_N_POWER outside 12..24 */
int gpuntt_b200_4step_shape(int n_power, int* n1, int* n2);
/* GPU_Transpose (ntt_4step.cu:36-66): out[x * row + y] = in[y * col + x] for each of batch_size polynomials
 * spaced (1 << n_power) elements apart.  Unlike the reference it takes a stream. */
int gpuntt_b200_transpose(int element_bits, const void* in, void* out, int row, int col, int n_power, int batch_size,
                          void* stream);

/* Convenience forms of the above for the four hot entry points on unsigned 64/32-bit data with
 * a single modulus (GPU_NTT / GPU_INTT, ntt.cuh:315-340; in == out gives the *_Inplace forms). */
int gpuntt_b200_ntt_u64(const uint64_t* in, uint64_t* out, const uint64_t* root_table,
                        uint64_t modulus, int n_power, int reduction_poly, int batch_size,
                        void* stream);
int gpuntt_b200_intt_u64(const uint64_t* in, uint64_t* out, const uint64_t* inv_root_table,
                         uint64_t modulus, uint64_t n_inverse, int n_power, int reduction_poly,
                         int batch_size, void* stream);
int gpuntt_b200_ntt_u32(const uint32_t* in, uint32_t* out, const uint32_t* root_table,
                        uint32_t modulus, int n_power, int reduction_poly, int batch_size,
                        void* stream);
int gpuntt_b200_intt_u32(const uint32_t* in, uint32_t* out, const uint32_t* inv_root_table,
                         uint32_t modulus, uint32_t n_inverse, int n_power, int reduction_poly,
                         int batch_size, void* stream);

/* Host-buffer convenience (used by the end-to-end benchmark and the ctypes tests): copies
 * `in` (host, pinned or pageable) to an internal device buffer, runs the transform and copies the
 * result back to `out` (host); synchronises `stream` before returning. Single modulus, unsigned. */
int gpuntt_b200_merge_ntt_host(const gpuntt_b200_merge_desc* desc_with_host_in_out,
                               const void* host_root_table, size_t root_table_elems);

/* Number of kernel launches the last gpuntt_b200_* call on this thread enqueued (the caller's
 * evidence for "gpu_launches" in bench.py) and cumulative count since load. */
int gpuntt_b200_last_launch_count(void);
unsigned long long gpuntt_b200_total_launch_count(void);

/* Per-launch device timing for the benchmark's live roofline: while enabled, every kernel launch
 * is bracketed by CUDA events on its stream.  gpuntt_b200_profile_read waits for the recorded
 * launches, writes up to max_records (duration in ms, kind: 0 = twiddle companion pre-kernel,
 * k >= 1 = k-th merge pass of a call) and returns how many it wrote; the record list is cleared. */
void gpuntt_b200_set_profiling(int on);
int gpuntt_b200_profile_read(float* ms_out, int* kind_out, int max_records);

/* Testing aid: when on, the tuned persistent kernels (merge_fast.cu) are bypassed and every call
 * takes the generic pass kernel, so both implementations can be checked against the oracle. */
void gpuntt_b200_force_generic_path(int on);

/* Tuning / A-B testing knobs (process-wide; results never depend on them):
 *   FUSED_PASSES  1 (default): two-pass plans run as ONE launch with the passes chained through the L2
 *                 (merge_fused.cu) wherever that measured faster (32-bit data; launch-bound 64-bit calls);
 *                 2: wherever the shapes allow; 0: one launch per pass.
 *   FUSED_LAG     how many tile times the second pass trails the first inside the fused kernel (default 4). */
#define GPUNTT_B200_TUNE_FUSED_PASSES 1
#define GPUNTT_B200_TUNE_FUSED_LAG 2
/*   4STEP_TRANSPOSED  1 (default): the fused-contract forward 4-step writes its column phase as the n2 x n1 matrix
 *                 (transposing TMA store) and runs the row transforms along that layout -- no transpose kernel;
 *                 0: column pass, row Merge-NTT, transpose kernel. */
#define GPUNTT_B200_TUNE_4STEP_TRANSPOSED 3
/*   4STEP_MODULUS_CACHE  how gpuntt_b200_4step_ntt treats the RNS form with ONE device modulus (what the reference's
 *                 examples pass): 1 (default) read the Modulus / n^-1 back once per (device, pointers) -- the first such
 *                 call synchronises the stream, later ones do not -- and run the single-modulus tuned kernels (the value
 *                 behind a cached pointer must not change; gpuntt_b200_release_workspaces() forgets it); 0 never read
 *                 back (device-modulus kernels); 2 read back on every call. */
#define GPUNTT_B200_TUNE_4STEP_MODULUS_CACHE 4
/*   4STEP_RESIDENT_PAIRS  1 (default): the transposing column pass walks the batch through one tile position at a time with
 *                 that position's (W, W') pairs resident in shared memory (merge_wcol.cu); 0: pairs fetched per tile. */
#define GPUNTT_B200_TUNE_4STEP_RESIDENT_PAIRS 5
/*   ONE_TILE      64-bit N = 2^12 and 32-bit N = 2^13 are exactly one tile and can run with the whole transform inside it (one
 *                 launch, one HBM round trip, no hand-off between CTAs; one CTA per SM).  1 (default): inverse transforms -- the
 *                 measured win -- (64-bit: calls above the small-tile range below); 0: never; 2: every call of these sizes. */
#define GPUNTT_B200_TUNE_ONE_TILE 6
/*   SMALL_TILE_ELEMS  64-bit rings 2^12 .. 2^14: calls of at most this many elements in total (default 2^18) run the single-launch
 *                 kernel on 1024-element tiles instead of 4096-element ones -- four times the CTAs, a quarter of the work on the
 *                 critical path of a launch-bound call.  0: never. */
#define GPUNTT_B200_TUNE_SMALL_TILE_ELEMS 7
/*   SINGLE_POLY_TILES  1 (default): a call with ONE polynomial of a ring that takes three or more passes (64-bit: above 2^16,
 *                 32-bit: above 2^18) runs its contiguous pass on 2048- / 4096-element tiles of that polynomial (a two-polynomial
 *                 tile would be half zero fill); 0: the usual tiles. */
#define GPUNTT_B200_TUNE_SINGLE_POLY_TILES 8
void gpuntt_b200_tune(int knob, int value);

/* Batch-slice helpers for callers whose whole batch lives on ONE GPU: device g of ndev owns the
 * contiguous polynomials [groups * g / ndev, groups * (g + 1) / ndev) * unit with unit = max(mod_count, 1) and
 * groups = batch_size / unit (slice boundaries fall on multiples of mod_count, so every device can use the same modulus
 * array -- polynomial b uses modulus b % mod_count, ntt.cu:613 of the reference).  scatter copies slice g of the
 * [batch_size][poly_bytes] array `src` (on src_device) to dst[g] (on devices[g]); gather is the reverse.  The copies are
 * cudaMemcpyPeerAsync on streams[g] (a stream of devices[g]; NULL array = default streams): NVLink / NVSwitch when peer
 * access is available (it is enabled on first use), staged through the host otherwise.  Nothing synchronises. */
int gpuntt_b200_scatter_batch(const void* src, int src_device, void* const* dst, const int* devices, int ndev,
                              size_t poly_bytes, long long batch_size, int mod_count, void* const* streams);
int gpuntt_b200_gather_batch(void* dst, int dst_device, const void* const* src, const int* devices, int ndev,
                             size_t poly_bytes, long long batch_size, int mod_count, void* const* streams);

/* The input recipe of the reference's example drivers (example/ntt_merge/test_merge_ntt.cu:70-85,
 * test_4step_ntt.cu:48-57): std::mt19937 gen(seed); std::uniform_int_distribution<uint64_t> dis(0, modulus - 1);
 * count values, in order, into the HOST array out (uint64_t; callers narrow for 32-bit data).  Used by examples/,
 * bench.py and anything else that wants the reference's seed-0 stream without linking the reference. */
void gpuntt_b200_example_input(uint32_t seed, uint64_t modulus, uint64_t count, uint64_t* host_out);

/* Human-readable message for the last non-OK status returned on this thread. */
const char* gpuntt_b200_last_error(void);

/* Frees every cached scratch buffer (device-synchronising). */
void gpuntt_b200_release_workspaces(void);

/* Describes the launch plan chosen for (n_power, element_bits) as text into buf (for DESIGN.md /
 * profiles); returns the number of kernel launches of the transform itself. */
int gpuntt_b200_describe_plan(int n_power, int element_bits, char* buf, size_t buf_len);

int gpuntt_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUNTT_B200_H */
