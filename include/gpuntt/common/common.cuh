// include/gpuntt/common/common.cuh -- error convention and small host helpers.
// Same names and behaviour as the reference's common.cuh (src/include/gpuntt/common/common.cuh:17-56).
#ifndef GPUNTT_B200_COMMON_CUH
#define GPUNTT_B200_COMMON_CUH

#include <cuda_runtime.h>

#include <cassert>
#include <cstdint>
#include <exception>
#include <iostream> // the reference header pulls these in and its callers rely on it (common.cuh:10-15)
#include <stdexcept>
#include <string>

namespace gpuntt
{
    // what(): "CUDA Error in <file> at line <n>: <cudaGetErrorString>"  (common.cuh:20-40 of the reference)
    // Object layout = the reference's (file, line, error, text, in that order): the class is header-inline on both sides, so a
    // caller object compiled against the reference's header and this library must agree on it (one vtable, one what()).
    class CudaException : public std::exception
    {
      public:
        CudaException(const std::string& file, int line, cudaError_t error) : file_(file), line_(line), error_(error) {}
        // engine-side failures arrive as text from the C ABI (gpuntt_b200_last_error)
        CudaException(const std::string& file, int line, const std::string& message)
            : file_(file), line_(line), error_(cudaErrorUnknown),
              m_error_string("CUDA Error in " + file + " at line " + std::to_string(line) + ": " + message)
        {
        }
        const char* what() const noexcept override { return m_error_string.c_str(); }
        cudaError_t code() const noexcept { return error_; }

      private:
        std::string file_;
        int line_;
        cudaError_t error_;
        std::string m_error_string = "CUDA Error in " + file_ + " at line " + std::to_string(line_) + ": " + cudaGetErrorString(error_);
    };

#define GPUNTT_CUDA_CHECK(err)                                                                                         \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t gpuntt_error_ = (err);                                                                             \
        if (gpuntt_error_ != cudaSuccess) throw ::gpuntt::CudaException(__FILE__, __LINE__, gpuntt_error_);            \
    } while (0)

    // throws std::invalid_argument(errorMessage) when the condition is false (common.cu:5-11)
    void customAssert(bool condition, const std::string& errorMessage);

    // selects device 0 and prints its name (common.cu:13-22)
    void CudaDevice();

    // elementwise comparison; prints the first mismatch (common.cu:24-47)
    template <typename T> bool check_result(T* input1, T* input2, int size);

} // namespace gpuntt
#endif // GPUNTT_B200_COMMON_CUH
