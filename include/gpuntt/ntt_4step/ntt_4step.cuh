// include/gpuntt/ntt_4step/ntt_4step.cuh -- the 4-Step NTT host API of GPU-NTT, served by the B200 engine.
//
// Same declarations as the reference (src/include/gpuntt/ntt_4step/ntt_4step.cuh:19-33, 46-49,
// 278-308).  GPU_4STEP_NTT keeps the reference's I/O contract: for the forward transform the input is
// the n2 x n1 (pre-transposed) matrix and the output is the n1 x n2 matrix R whose transpose is
// NTT_4STEP_CPU::ntt(x); the inverse expects NTT_4STEP_CPU::intt_first_transpose to have been applied.
// Additions over the reference: both functions enqueue on cfg.stream (the reference's 4-step ignores
// it and GPU_Transpose has no stream, so that one uses the legacy default stream as there), and
// GPU_4STEP_NTT_Fused runs natural-order input to final-order output with the transposes absorbed.
#ifndef GPUNTT_B200_NTT_4STEP_CORE_CUH
#define GPUNTT_B200_NTT_4STEP_CORE_CUH

#include <cuda_runtime.h>

#include "gpuntt/ntt_4step/ntt_4step_cpu.cuh"

namespace gpuntt
{
    template <typename T> struct ntt4step_configuration
    {
        int n_power;
        type ntt_type;
        Ninverse<T> mod_inverse;
        cudaStream_t stream;
    };

    template <typename T> struct ntt4step_rns_configuration
    {
        int n_power;
        type ntt_type;
        Ninverse<T>* mod_inverse;
        cudaStream_t stream;
    };

    // out[x * row + y] = in[y * col + x] for every polynomial of the batch
    template <typename T>
    __host__ void GPU_Transpose(T* polynomial_in, T* polynomial_out, const int row, const int col, const int n_power,
                                const int batch_size);

    template <typename T>
    __host__ void GPU_4STEP_NTT(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table,
                                Root<T>* n2_root_of_unity_table, Root<T>* W_root_of_unity_table, Modulus<T> modulus,
                                ntt4step_configuration<T> cfg, int batch_size);

    template <typename T>
    __host__ void GPU_4STEP_NTT(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table,
                                Root<T>* n2_root_of_unity_table, Root<T>* W_root_of_unity_table, Modulus<T>* modulus,
                                ntt4step_rns_configuration<T> cfg, int batch_size, int mod_count);

    // B200 addition: natural-order coefficients in, NTT_4STEP_CPU::ntt order out (forward) /
    // NTT_4STEP_CPU::ntt order in, natural-order coefficients out (inverse); no GPU_Transpose calls,
    // no host-side permutation.  device_in == device_out is allowed.
    template <typename T>
    __host__ void GPU_4STEP_NTT_Fused(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table,
                                      Root<T>* n2_root_of_unity_table, Root<T>* W_root_of_unity_table,
                                      Modulus<T> modulus, ntt4step_configuration<T> cfg, int batch_size);

} // namespace gpuntt
#endif // GPUNTT_B200_NTT_4STEP_CORE_CUH
