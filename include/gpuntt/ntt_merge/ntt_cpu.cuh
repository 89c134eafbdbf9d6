// include/gpuntt/ntt_merge/ntt_cpu.cuh -- host (CPU) transforms of the Merge-NTT family.
// Same classes as the reference header (src/include/gpuntt/ntt_merge/ntt_cpu.cuh:13-33); these are
// what the example programs compare the GPU against.  Host-only utilities: nothing in the GPU entry
// points ever calls them.
#ifndef GPUNTT_B200_NTT_CPU_CUH
#define GPUNTT_B200_NTT_CPU_CUH

#include "gpuntt/common/nttparameters.cuh"

namespace gpuntt
{
    // O(N^2) product of a and b reduced by X^N - 1 or X^N + 1
    template <typename T>
    std::vector<T> schoolbook_poly_multiplication(std::vector<T> a, std::vector<T> b, Modulus<T> modulus,
                                                  ReductionPolynomial reduction_poly);

    template <typename T> class NTTCPU
    {
      public:
        NTTParameters<T> parameters;
        NTTCPU(NTTParameters<T> parameters_);

        std::vector<T> mult(std::vector<T>& input1, std::vector<T>& input2); // pointwise
        std::vector<T> ntt(std::vector<T>& input);                           // natural in, bit-reversed out
        std::vector<T> intt(std::vector<T>& input);                          // bit-reversed in, natural out, times n^-1
    };
} // namespace gpuntt
#endif // GPUNTT_B200_NTT_CPU_CUH
