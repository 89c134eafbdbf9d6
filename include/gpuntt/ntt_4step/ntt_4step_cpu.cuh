// include/gpuntt/ntt_4step/ntt_4step_cpu.cuh -- host (CPU) 4-step transform.
// Same class as the reference header (src/include/gpuntt/ntt_4step/ntt_4step_cpu.cuh:13-52).
#ifndef GPUNTT_B200_NTT_4STEP_CPU_CUH
#define GPUNTT_B200_NTT_4STEP_CPU_CUH

#include "gpuntt/common/nttparameters.cuh"

namespace gpuntt
{
    template <typename T> class NTT_4STEP_CPU
    {
      public:
        NTTParameters4Step<T> parameters;
        NTT_4STEP_CPU(NTTParameters4Step<T> parameters_);

        std::vector<T> mult(std::vector<T>& input1, std::vector<T>& input2);
        // n1 x n2 matrix view of the input: column transforms, twiddle product, row transforms, transpose
        std::vector<T> ntt(std::vector<T>& input);
        std::vector<T> intt(std::vector<T>& input);
        // the permutation GPU_4STEP_NTT(INVERSE) expects to have been applied to its input
        std::vector<T> intt_first_transpose(const std::vector<T>& input);

      private:
        void core_ntt(std::vector<T>& input, std::vector<T> root_table, int log_size);
        void core_intt(std::vector<T>& input, std::vector<T> root_table, int log_size);
        void product(std::vector<T>& input, std::vector<T> root_table, int log_size);
        std::vector<std::vector<T>> vector_to_matrix(const std::vector<T>& array, int rows, int cols);
        std::vector<std::vector<T>> vector_to_matrix_intt(const std::vector<T>& array, int rows, int cols);
        std::vector<T> matrix_to_vector(const std::vector<std::vector<T>>& originalMatrix);
        std::vector<std::vector<T>> transpose_matrix(const std::vector<std::vector<T>>& originalMatrix);
    };
} // namespace gpuntt
#endif // GPUNTT_B200_NTT_4STEP_CPU_CUH
