// gpu_ntt_b200/cxx/ntt_api.cu -- GPU_NTT / GPU_INTT / *_Inplace / *_Ordered of gpuntt/ntt_merge/ntt.cuh
// as forwarders onto the C ABI (include/gpuntt_b200.h), plus the explicit instantiations the
// reference library exports (src/lib/ntt_merge/ntt.cu:4948-5082, 5140-5244).
#include <stdexcept>
#include <string>

#include "gpuntt/ntt_merge/ntt.cuh"
#include "gpuntt_b200.h"

namespace gpuntt
{
    namespace
    {
        // status -> the reference's exception convention (ntt.cu:2088-2091, 2253; common.cuh:42-50)
        void raise(int status, const char* file, int line)
        {
            if (status == GPUNTT_B200_OK) return;
            const std::string msg = gpuntt_b200_last_error();
            if (status == GPUNTT_B200_ERR_CUDA) throw CudaException(file, line, msg);
            throw std::invalid_argument(msg);
        }
#define GPUNTT_B200_CALL(expr) raise((expr), __FILE__, __LINE__)

        template <typename TU, typename CFG>
        gpuntt_b200_merge_desc describe(const void* in, void* out, const void* table, const CFG& cfg, int direction, bool is_signed,
                                        int batch_size)
        {
            gpuntt_b200_merge_desc d{};
            d.element_bits = static_cast<int>(sizeof(TU)) * 8;
            d.is_signed = is_signed ? 1 : 0;
            d.direction = direction;
            d.n_power = cfg.n_power;
            d.ntt_layout = static_cast<int>(cfg.ntt_layout);
            d.reduction_poly = static_cast<int>(cfg.reduction_poly);
            d.batch_size = batch_size;
            d.in = in;
            d.out = out;
            d.root_of_unity_table = table;
            d.stream = cfg.stream;
            return d;
        }
        template <typename TU>
        void run_single(const void* in, void* out, const void* table, Modulus<TU> modulus, const ntt_configuration<TU>& cfg, int direction,
                        bool is_signed, int batch_size)
        {
            gpuntt_b200_merge_desc d = describe<TU>(in, out, table, cfg, direction, is_signed, batch_size);
            d.mod_count = 0;
            d.modulus_value = modulus.value;
            d.mod_inverse_value = cfg.mod_inverse;
            GPUNTT_B200_CALL(gpuntt_b200_merge_ntt(&d));
        }
        template <typename TU>
        void run_rns(const void* in, void* out, const void* table, Modulus<TU>* modulus, const ntt_rns_configuration<TU>& cfg, int direction,
                     bool is_signed, int batch_size, int mod_count, const int* modulus_order = nullptr, const int* poly_order = nullptr)
        {
            if (mod_count < 1) throw std::invalid_argument("mod_count must be at least 1");
            gpuntt_b200_merge_desc d = describe<TU>(in, out, table, cfg, direction, is_signed, batch_size);
            d.mod_count = mod_count;
            d.modulus_dev = modulus;
            d.mod_inverse_dev = cfg.mod_inverse;
            d.modulus_order_dev = modulus_order;
            d.poly_order_dev = poly_order;
            GPUNTT_B200_CALL(gpuntt_b200_merge_ntt(&d));
        }
    } // namespace

    // ---------------------------------------------------------------- single modulus
    template <typename T>
    __host__ void GPU_NTT(T* device_in, typename std::make_unsigned<T>::type* device_out,
                          Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                          Modulus<typename std::make_unsigned<T>::type> modulus,
                          ntt_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size)
    {
        using TU = typename std::make_unsigned<T>::type;
        run_single<TU>(device_in, device_out, root_of_unity_table, modulus, cfg, GPUNTT_B200_FORWARD, std::is_signed<T>::value, batch_size);
    }
    template <typename T>
    __host__ void GPU_INTT(typename std::make_unsigned<T>::type* device_in, T* device_out,
                           Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                           Modulus<typename std::make_unsigned<T>::type> modulus,
                           ntt_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size)
    {
        using TU = typename std::make_unsigned<T>::type;
        run_single<TU>(device_in, device_out, root_of_unity_table, modulus, cfg, GPUNTT_B200_INVERSE, std::is_signed<T>::value, batch_size);
    }
    template <typename T>
    __host__ void GPU_NTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T> modulus, ntt_configuration<T> cfg, int batch_size)
    {
        GPU_NTT<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size);
    }
    template <typename T>
    __host__ void GPU_INTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T> modulus, ntt_configuration<T> cfg, int batch_size)
    {
        GPU_INTT<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size);
    }

    // ---------------------------------------------------------------- RNS
    template <typename T>
    __host__ void GPU_NTT(T* device_in, typename std::make_unsigned<T>::type* device_out,
                          Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                          Modulus<typename std::make_unsigned<T>::type>* modulus,
                          ntt_rns_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size, int mod_count)
    {
        using TU = typename std::make_unsigned<T>::type;
        run_rns<TU>(device_in, device_out, root_of_unity_table, modulus, cfg, GPUNTT_B200_FORWARD, std::is_signed<T>::value, batch_size, mod_count);
    }
    template <typename T>
    __host__ void GPU_INTT(typename std::make_unsigned<T>::type* device_in, T* device_out,
                           Root<typename std::make_unsigned<T>::type>* root_of_unity_table,
                           Modulus<typename std::make_unsigned<T>::type>* modulus,
                           ntt_rns_configuration<typename std::make_unsigned<T>::type> cfg, int batch_size, int mod_count)
    {
        using TU = typename std::make_unsigned<T>::type;
        run_rns<TU>(device_in, device_out, root_of_unity_table, modulus, cfg, GPUNTT_B200_INVERSE, std::is_signed<T>::value, batch_size, mod_count);
    }
    template <typename T>
    __host__ void GPU_NTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus, ntt_rns_configuration<T> cfg,
                                  int batch_size, int mod_count)
    {
        GPU_NTT<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size, mod_count);
    }
    template <typename T>
    __host__ void GPU_INTT_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus, ntt_rns_configuration<T> cfg,
                                   int batch_size, int mod_count)
    {
        GPU_INTT<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size, mod_count);
    }

    // ---------------------------------------------------------------- RNS with indirection
    namespace
    {
        template <typename T> void check_ordered(const ntt_rns_configuration<T>& cfg)
        {
            // the reference's ordered entry points only exist for the multi-kernel sizes (ntt.cu:3607-3610)
            if (cfg.n_power <= 9 || cfg.n_power >= 29) throw std::invalid_argument("Invalid n_power range!");
        }
        template <typename T> int direction_of(const ntt_rns_configuration<T>& cfg)
        {
            return cfg.ntt_type == FORWARD ? GPUNTT_B200_FORWARD : GPUNTT_B200_INVERSE;
        }
    } // namespace
    template <typename T>
    __host__ void GPU_NTT_Modulus_Ordered(T* device_in, T* device_out, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                          ntt_rns_configuration<T> cfg, int batch_size, int mod_count, int* order)
    {
        check_ordered(cfg);
        run_rns<T>(device_in, device_out, root_of_unity_table, modulus, cfg, direction_of(cfg), false, batch_size, mod_count, order, nullptr);
    }
    template <typename T>
    __host__ void GPU_NTT_Modulus_Ordered_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                                  ntt_rns_configuration<T> cfg, int batch_size, int mod_count, int* order)
    {
        GPU_NTT_Modulus_Ordered<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size, mod_count, order);
    }
    template <typename T>
    __host__ void GPU_NTT_Poly_Ordered(T* device_in, T* device_out, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                       ntt_rns_configuration<T> cfg, int batch_size, int mod_count, int* order)
    {
        check_ordered(cfg);
        run_rns<T>(device_in, device_out, root_of_unity_table, modulus, cfg, direction_of(cfg), false, batch_size, mod_count, nullptr, order);
    }
    template <typename T>
    __host__ void GPU_NTT_Poly_Ordered_Inplace(T* device_inout, Root<T>* root_of_unity_table, Modulus<T>* modulus,
                                               ntt_rns_configuration<T> cfg, int batch_size, int mod_count, int* order)
    {
        GPU_NTT_Poly_Ordered<T>(device_inout, device_inout, root_of_unity_table, modulus, cfg, batch_size, mod_count, order);
    }

    // ---------------------------------------------------------------- the reference's export list
#define GPUNTT_B200_INSTANTIATE_IO(T)                                                                                                      \
    template __host__ void GPU_NTT<T>(T*, typename std::make_unsigned<T>::type*, Root<typename std::make_unsigned<T>::type>*,              \
                                      Modulus<typename std::make_unsigned<T>::type>,                                                       \
                                      ntt_configuration<typename std::make_unsigned<T>::type>, int);                                       \
    template __host__ void GPU_INTT<T>(typename std::make_unsigned<T>::type*, T*, Root<typename std::make_unsigned<T>::type>*,             \
                                       Modulus<typename std::make_unsigned<T>::type>,                                                      \
                                       ntt_configuration<typename std::make_unsigned<T>::type>, int);                                      \
    template __host__ void GPU_NTT<T>(T*, typename std::make_unsigned<T>::type*, Root<typename std::make_unsigned<T>::type>*,              \
                                      Modulus<typename std::make_unsigned<T>::type>*,                                                      \
                                      ntt_rns_configuration<typename std::make_unsigned<T>::type>, int, int);                              \
    template __host__ void GPU_INTT<T>(typename std::make_unsigned<T>::type*, T*, Root<typename std::make_unsigned<T>::type>*,             \
                                       Modulus<typename std::make_unsigned<T>::type>*,                                                     \
                                       ntt_rns_configuration<typename std::make_unsigned<T>::type>, int, int);
    GPUNTT_B200_INSTANTIATE_IO(Data32)
    GPUNTT_B200_INSTANTIATE_IO(Data64)
    GPUNTT_B200_INSTANTIATE_IO(Data32s)
    GPUNTT_B200_INSTANTIATE_IO(Data64s)

#define GPUNTT_B200_INSTANTIATE_INPLACE(T)                                                                                                 \
    template __host__ void GPU_NTT_Inplace<T>(T*, Root<T>*, Modulus<T>, ntt_configuration<T>, int);                                        \
    template __host__ void GPU_INTT_Inplace<T>(T*, Root<T>*, Modulus<T>, ntt_configuration<T>, int);                                       \
    template __host__ void GPU_NTT_Inplace<T>(T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int);                              \
    template __host__ void GPU_INTT_Inplace<T>(T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int);                             \
    template __host__ void GPU_NTT_Modulus_Ordered<T>(T*, T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int, int*);            \
    template __host__ void GPU_NTT_Modulus_Ordered_Inplace<T>(T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int, int*);        \
    template __host__ void GPU_NTT_Poly_Ordered<T>(T*, T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int, int*);               \
    template __host__ void GPU_NTT_Poly_Ordered_Inplace<T>(T*, Root<T>*, Modulus<T>*, ntt_rns_configuration<T>, int, int, int*);
    GPUNTT_B200_INSTANTIATE_INPLACE(Data32)
    GPUNTT_B200_INSTANTIATE_INPLACE(Data64)

} // namespace gpuntt
