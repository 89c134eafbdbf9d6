// gpu_ntt_b200/cxx/ntt_4step_api.cu -- GPU_4STEP_NTT / GPU_Transpose / GPU_4STEP_NTT_Fused of
// gpuntt/ntt_4step/ntt_4step.cuh as forwarders onto the C ABI, with the reference's explicit
// instantiations (src/lib/ntt_4step/ntt_4step.cu:3268-3636: Data32 and Data64).
#include <iostream>
#include <stdexcept>
#include <string>

#include "gpuntt/ntt_4step/ntt_4step.cuh"
#include "gpuntt_b200.h"

namespace gpuntt
{
    namespace
    {
        void raise4(int status, const char* file, int line)
        {
            if (status == GPUNTT_B200_OK) return;
            const std::string msg = gpuntt_b200_last_error();
            if (status == GPUNTT_B200_ERR_CUDA) throw CudaException(file, line, msg);
            throw std::invalid_argument(msg);
        }
        template <typename T>
        gpuntt_b200_4step_desc describe4(T* in, T* out, Root<T>* n1, Root<T>* n2, Root<T>* w, int n_power, type ntt_type, cudaStream_t stream,
                                         int batch_size, int contract)
        {
            gpuntt_b200_4step_desc d{};
            d.element_bits = static_cast<int>(sizeof(T)) * 8;
            d.direction = ntt_type == FORWARD ? GPUNTT_B200_FORWARD : GPUNTT_B200_INVERSE;
            d.n_power = n_power;
            d.batch_size = batch_size;
            d.io_contract = contract;
            d.in = in;
            d.out = out;
            d.n1_table = n1;
            d.n2_table = n2;
            d.w_table = w;
            d.stream = stream;
            return d;
        }
        // the reference only prints for sizes it has no plan for (ntt_4step.cu:2529-2532) -- keep that
        bool supported(int n_power)
        {
            if (gpuntt_b200_4step_shape(n_power, nullptr, nullptr) == GPUNTT_B200_OK) return true;
            std::cout << "This ring size is not supported!" << std::endl;
            return false;
        }
    } // namespace

    template <typename T>
    __host__ void GPU_Transpose(T* polynomial_in, T* polynomial_out, const int row, const int col, const int n_power, const int batch_size)
    {
        // no stream parameter in the reference: legacy default stream, as there
        raise4(gpuntt_b200_transpose(static_cast<int>(sizeof(T)) * 8, polynomial_in, polynomial_out, row, col, n_power, batch_size, nullptr),
               __FILE__, __LINE__);
    }

    template <typename T>
    __host__ void GPU_4STEP_NTT(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table, Root<T>* n2_root_of_unity_table,
                                Root<T>* W_root_of_unity_table, Modulus<T> modulus, ntt4step_configuration<T> cfg, int batch_size)
    {
        if (!supported(cfg.n_power)) return;
        gpuntt_b200_4step_desc d = describe4<T>(device_in, device_out, n1_root_of_unity_table, n2_root_of_unity_table, W_root_of_unity_table,
                                                cfg.n_power, cfg.ntt_type, cfg.stream, batch_size, GPUNTT_B200_4STEP_REFERENCE);
        d.modulus_value = modulus.value;
        d.mod_inverse_value = cfg.mod_inverse;
        raise4(gpuntt_b200_4step_ntt(&d), __FILE__, __LINE__);
    }

    template <typename T>
    __host__ void GPU_4STEP_NTT(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table, Root<T>* n2_root_of_unity_table,
                                Root<T>* W_root_of_unity_table, Modulus<T>* modulus, ntt4step_rns_configuration<T> cfg, int batch_size,
                                int mod_count)
    {
        if (!supported(cfg.n_power)) return;
        gpuntt_b200_4step_desc d = describe4<T>(device_in, device_out, n1_root_of_unity_table, n2_root_of_unity_table, W_root_of_unity_table,
                                                cfg.n_power, cfg.ntt_type, cfg.stream, batch_size, GPUNTT_B200_4STEP_REFERENCE);
        d.mod_count = mod_count;
        d.modulus_dev = modulus;
        d.mod_inverse_dev = cfg.mod_inverse;
        raise4(gpuntt_b200_4step_ntt(&d), __FILE__, __LINE__);
    }

    template <typename T>
    __host__ void GPU_4STEP_NTT_Fused(T* device_in, T* device_out, Root<T>* n1_root_of_unity_table, Root<T>* n2_root_of_unity_table,
                                      Root<T>* W_root_of_unity_table, Modulus<T> modulus, ntt4step_configuration<T> cfg, int batch_size)
    {
        if (!supported(cfg.n_power)) return;
        gpuntt_b200_4step_desc d = describe4<T>(device_in, device_out, n1_root_of_unity_table, n2_root_of_unity_table, W_root_of_unity_table,
                                                cfg.n_power, cfg.ntt_type, cfg.stream, batch_size, GPUNTT_B200_4STEP_FUSED);
        d.modulus_value = modulus.value;
        d.mod_inverse_value = cfg.mod_inverse;
        raise4(gpuntt_b200_4step_ntt(&d), __FILE__, __LINE__);
    }

#define GPUNTT_B200_INSTANTIATE_4STEP(T)                                                                                                   \
    template __host__ void GPU_Transpose<T>(T*, T*, const int, const int, const int, const int);                                           \
    template __host__ void GPU_4STEP_NTT<T>(T*, T*, Root<T>*, Root<T>*, Root<T>*, Modulus<T>, ntt4step_configuration<T>, int);             \
    template __host__ void GPU_4STEP_NTT<T>(T*, T*, Root<T>*, Root<T>*, Root<T>*, Modulus<T>*, ntt4step_rns_configuration<T>, int, int);   \
    template __host__ void GPU_4STEP_NTT_Fused<T>(T*, T*, Root<T>*, Root<T>*, Root<T>*, Modulus<T>, ntt4step_configuration<T>, int);
    GPUNTT_B200_INSTANTIATE_4STEP(Data32)
    GPUNTT_B200_INSTANTIATE_4STEP(Data64)
} // namespace gpuntt
