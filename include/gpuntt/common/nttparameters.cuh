// include/gpuntt/common/nttparameters.cuh -- host-side parameter generators.
//
// Same public surface as the reference header (src/include/gpuntt/common/nttparameters.cuh:17-170):
// enums, NTTFactors<T>, NTTParameters<T>, NTTParameters4Step<T>, bitreverse.  The constants (default
// prime / omega / psi pools, 4-step shapes) are the reference's (nttparameters.cu:84-142, 229-354),
// because callers and the CPU oracles depend on them; table construction is this project's own code
// (gpu_ntt_b200/cxx/nttparameters.cu).
#ifndef GPUNTT_B200_NTT_PARAMETERS_CUH
#define GPUNTT_B200_NTT_PARAMETERS_CUH

#include <type_traits>
#include <vector>

#include "gpuntt/common/common.cuh"
#include "gpuntt/common/modular_arith.cuh"

namespace gpuntt
{
    int bitreverse(int index, int n_power);

    enum type
    {
        FORWARD,
        INVERSE
    };

    enum NTTLayout
    {
        PerPolynomial, // one transform per row of the [batch][N] matrix
        PerCoefficient // one transform per column of it
    };

    enum ReductionPolynomial
    {
        X_N_plus, // negacyclic, psi powers
        X_N_minus // cyclic, omega powers
    };

    template <typename T> struct NTTFactors
    {
        Modulus<T> modulus;
        T omega;
        T psi;
        __host__ NTTFactors(Modulus<T> q_, T omega_, T psi_) : modulus(q_), omega(omega_), psi(psi_) {}
        __host__ NTTFactors() : omega(0), psi(0) {}
    };

    template <typename T> class NTTParameters
    {
      public:
        int logn;
        T n;
        ReductionPolynomial poly_reduction;
        Modulus<T> modulus;
        T omega;
        T psi;
        Ninverse<T> n_inv;
        T root_of_unity;
        T inverse_root_of_unity;
        T root_of_unity_size;
        std::vector<T> forward_root_of_unity_table; // natural order: root^0, root^1, ...
        std::vector<T> inverse_root_of_unity_table;

        NTTParameters(int LOGN, ReductionPolynomial poly_reduce_type);
        NTTParameters(int LOGN, NTTFactors<T> ntt_factors, ReductionPolynomial poly_reduce_type);
        NTTParameters();

        // table[bitreverse(i, log2 size)] -- the order GPU_NTT / GPU_INTT expect on the device
        std::vector<Root<T>> gpu_root_of_unity_table_generator(std::vector<T> table);

      private:
        Modulus<T> modulus_pool();
        T omega_pool();
        T psi_pool();
        void forward_root_of_unity_table_generator();
        void inverse_root_of_unity_table_generator();
        void n_inverse_generator();
    };

    template <typename T> class NTTParameters4Step
    {
      public:
        int logn;
        T n;
        ReductionPolynomial poly_reduction;
        Modulus<T> modulus;
        T omega;
        T psi;
        T n_inv;
        Ninverse<T> n_inv_gpu;
        T root_of_unity;
        T inverse_root_of_unity;
        T root_of_unity_size;
        int n1, n2;
        std::vector<T> n1_based_root_of_unity_table;
        std::vector<T> n2_based_root_of_unity_table;
        std::vector<T> W_root_of_unity_table;
        std::vector<T> n1_based_inverse_root_of_unity_table;
        std::vector<T> n2_based_inverse_root_of_unity_table;
        std::vector<T> W_inverse_root_of_unity_table;

        NTTParameters4Step(int LOGN, ReductionPolynomial poly_reduce_type);
        NTTParameters4Step();

        std::vector<Root<T>> gpu_root_of_unity_table_generator(std::vector<T> table);

      private:
        Modulus<T> modulus_pool();
        T omega_pool();
        T psi_pool();
        std::vector<int> matrix_dimention();
        void small_forward_root_of_unity_table_generator();
        void TW_forward_table_generator();
        void small_inverse_root_of_unity_table_generator();
        void TW_inverse_table_generator();
        void n_inverse_generator();
        void n_inverse_generator_gpu();
    };

} // namespace gpuntt
#endif // GPUNTT_B200_NTT_PARAMETERS_CUH
