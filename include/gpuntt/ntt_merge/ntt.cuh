// include/gpuntt/ntt_merge/ntt.cuh -- the Merge-NTT host API of GPU-NTT, served by the B200 engine.
//
// Declarations are spelled exactly like the reference's (src/include/gpuntt/ntt_merge/ntt.cuh:29-51,
// 315-340, 395-421, 495-507, 592-603) -- same template parameter use, same dependent-type spellings
// -- so the explicitly instantiated symbols in libntt have the same mangled names and callers that
// were compiled against GPU-NTT link unchanged.  Every function is a thin forwarder onto the C ABI in
// gpuntt_b200.h (gpu_ntt_b200/cxx/ntt_api.cu); none of the reference's kernels, KernelConfig tables
// or device butterfly helpers exist here.
//
// Behaviour kept from the reference: direction is the function name (cfg.ntt_type is ignored by
// GPU_NTT / GPU_INTT, honoured by the *_Ordered entry points); cfg.zero_padding is ignored; work is
// only enqueued on cfg.stream; std::invalid_argument for a bad n_power / layout, gpuntt::CudaException
// for a failed launch.  Behaviour fixed: single-modulus out-of-place GPU_INTT is correct for every
// n_power (the reference re-reads device_in in every launch, ntt.cu:2367-2390).
#ifndef GPUNTT_B200_NTT_CORE_CUH
#define GPUNTT_B200_NTT_CORE_CUH

#include <cuda_runtime.h>

#include <functional>
#include <type_traits>
#include <unordered_map>

#include "gpuntt/ntt_merge/ntt_cpu.cuh"

typedef std::uint32_t location_t;

namespace gpuntt
{
    template <typename T> struct ntt_configuration
    {
        int n_power;
        type ntt_type;
        NTTLayout ntt_layout;
        ReductionPolynomial reduction_poly;
        bool zero_padding;
        Ninverse<T> mod_inverse;
        cudaStream_t stream;
    };

    template <typename T> struct ntt_rns_configuration
    {
        int n_power;
        type ntt_type;
        NTTLayout ntt_layout;
        ReductionPolynomial reduction_poly;
        bool zero_padding;
        Ninverse<T>* mod_inverse; // device array, one n^-1 per modulus
        cudaStream_t stream;
    };

    // ---- single modulus (by value)
    template <typename T>
    __host__ void GPU_NTT(T* device_in, typename std::make_unsigned<T>::type* device_out,
                          Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                          Modulus<typename std::make_unsigned<T>::type> modulus,
                          ntt_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size);

    template <typename T>
    __host__ void GPU_INTT(typename std::make_unsigned<T>::type* device_in, T* device_out,
                           Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                           Modulus<typename std::make_unsigned<T>::type> modulus,
                           ntt_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size);

    template <typename T>
    __host__ void GPU_NTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T> modulus,
                                  ntt_configuration<T> cfg, int batch_size);

    template <typename T>
    __host__ void GPU_INTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T> modulus,
                                   ntt_configuration<T> cfg, int batch_size);

    // ---- RNS: polynomial b uses modulus[b % mod_count], table slice (b % mod_count) << n_power
    template <typename T>
    __host__ void GPU_NTT(T* device_in, typename std::make_unsigned<T>::type* device_out,
                          Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                          Modulus<typename std::make_unsigned<T>::type>* modulus,
                          ntt_rns_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size,
                          int mod_count);

    template <typename T>
    __host__ void GPU_INTT(typename std::make_unsigned<T>::type* device_in, T* device_out,
                           Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                           Modulus<typename std::make_unsigned<T>::type>* modulus,
                           ntt_rns_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size,
                           int mod_count);

    template <typename T>
    __host__ void GPU_NTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                  ntt_rns_configuration<T> cfg, int batch_size, int mod_count);

    template <typename T>
    __host__ void GPU_INTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                   ntt_rns_configuration<T> cfg, int batch_size, int mod_count);

    // ---- RNS with indirection (direction from cfg.ntt_type).
    // Modulus_Ordered: polynomial b uses modulus / table slice order[b % mod_count].
    // Poly_Ordered:    the b-th transform runs on the polynomial stored at slot order[b] (in that
    //                  slot of device_out), with modulus b % mod_count.   `order` is a device array.
    template <typename T>
    __host__ void GPU_NTT_Modulus_Ordered(T* device_in, T* device_out, Root<T>* root_of_unity_table,
                                          Modulus<T>* modulus, ntt_rns_configuration<T> cfg, int batch_size,
                                          int mod_count, int* order);
    template <typename T>
    __host__ void GPU_NTT_Modulus_Ordered_Inplace(T* device_inout, Root<T>* root_of_unity_table,
                                                  Modulus<T>* modulus, ntt_rns_configuration<T> cfg,
                                                  int batch_size, int mod_count, int* order);
    template <typename T>
    __host__ void GPU_NTT_Poly_Ordered(T* device_in, T* device_out, Root<T>* root_of_unity_table,
                                       Modulus<T>* modulus, ntt_rns_configuration<T> cfg, int batch_size,
                                       int mod_count, int* order);
    template <typename T>
    __host__ void GPU_NTT_Poly_Ordered_Inplace(T* device_inout, Root<T>* root_of_unity_table,
                                               Modulus<T>* modulus, ntt_rns_configuration<T> cfg,
                                               int batch_size, int mod_count, int* order);

} // namespace gpuntt
#endif // GPUNTT_B200_NTT_CORE_CUH
