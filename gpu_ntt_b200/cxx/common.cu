// gpu_ntt_b200/cxx/common.cu -- host helpers declared in gpuntt/common/common.cuh
// (behaviour of the reference's src/lib/common/common.cu:5-54).
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "gpuntt/common/common.cuh"
#include "gpuntt/common/nttparameters.cuh"

namespace gpuntt
{
    void customAssert(bool condition, const std::string& errorMessage)
    {
        if (!condition) throw std::invalid_argument(errorMessage);
    }

    void CudaDevice()
    {
        const int device = 0;
        cudaDeviceProp prop;
        GPUNTT_CUDA_CHECK(cudaSetDevice(device));
        GPUNTT_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        std::printf("GPU Device %d: %s (compute capability %d.%d)\n\n", device, prop.name, prop.major, prop.minor);
    }

    template <typename T> bool check_result(T* input1, T* input2, int size)
    {
        for (int i = 0; i < size; i++)
            if (input1[i] != input2[i])
            {
                std::cout << "Error in index: " << i << " -> " << input1[i] << " - " << input2[i] << " " << std::endl;
                return false;
            }
        return true;
    }
    template bool check_result<std::uint64_t>(std::uint64_t*, std::uint64_t*, int);
    template bool check_result<std::uint32_t>(std::uint32_t*, std::uint32_t*, int);
    template bool check_result<std::int64_t>(std::int64_t*, std::int64_t*, int);
    template bool check_result<std::int32_t>(std::int32_t*, std::int32_t*, int);

    int bitreverse(int index, int n_power)
    {
        int r = 0;
        for (int b = 0; b < n_power; b++) r |= ((index >> b) & 1) << (n_power - 1 - b);
        return r;
    }
} // namespace gpuntt
