// gpu_ntt_b200/cxx/nttparameters.cu -- NTTParameters<T> / NTTParameters4Step<T>.
//
// Produces exactly the values of the reference's generators (src/lib/common/nttparameters.cu:22-471)
// -- the default prime / root pools and the 4-step shapes are data the CPU oracles and callers share
// -- but builds every table with running products (O(size) multiplies) instead of one modular
// exponentiation per entry, which turns the reference's 13 s W-table build at logN = 24 into ~0.2 s.
#include <stdexcept>

#include "gpuntt/common/nttparameters.cuh"

namespace gpuntt
{
    namespace
    {
        template <typename T> std::vector<T> powers(T root, std::size_t count, const Modulus<T>& q)
        {
            std::vector<T> t(count);
            T acc = 1;
            for (std::size_t i = 0; i < count; i++)
            {
                t[i] = acc;
                acc = OPERATOR<T>::mult(acc, root, q);
            }
            return t;
        }
        template <typename T> std::vector<T> bit_reversed_copy(const std::vector<T>& table)
        {
            int lg = 0;
            while ((std::size_t(1) << lg) < table.size()) lg++;
            std::vector<T> out(table.size());
            for (std::size_t i = 0; i < table.size(); i++) out[i] = table[bitreverse(static_cast<int>(i), lg)];
            return out;
        }
        template <typename T> constexpr int max_logn() { return std::is_same<T, Data32>::value ? 25 : 28; }
    } // namespace

    // ---------------------------------------------------------------- NTTParameters
    template <typename T> NTTParameters<T>::NTTParameters(int LOGN, ReductionPolynomial poly_reduce_type)
    {
        logn = LOGN;
        n = T(1) << logn;
        poly_reduction = poly_reduce_type;
        modulus = modulus_pool();
        omega = omega_pool();
        psi = psi_pool();
        root_of_unity = (poly_reduce_type == X_N_minus) ? omega : psi;
        inverse_root_of_unity = OPERATOR<T>::modinv(root_of_unity, modulus);
        root_of_unity_size = (poly_reduce_type == X_N_minus) ? (T(1) << (logn - 1)) : (T(1) << logn);
        forward_root_of_unity_table_generator();
        inverse_root_of_unity_table_generator();
        n_inverse_generator();
    }

    template <typename T>
    NTTParameters<T>::NTTParameters(int LOGN, NTTFactors<T> ntt_factors, ReductionPolynomial poly_reduce_type)
    {
        logn = LOGN;
        n = T(1) << logn;
        poly_reduction = poly_reduce_type;
        modulus = ntt_factors.modulus;
        omega = ntt_factors.omega;
        psi = ntt_factors.psi;
        n_inverse_generator();
        root_of_unity = (poly_reduce_type == X_N_minus) ? omega : psi;
        inverse_root_of_unity = OPERATOR<T>::modinv(root_of_unity, modulus);
        root_of_unity_size = (poly_reduce_type == X_N_minus) ? (T(1) << (logn - 1)) : (T(1) << logn);
        forward_root_of_unity_table_generator();
        inverse_root_of_unity_table_generator();
    }

    template <typename T> NTTParameters<T>::NTTParameters() : logn(0), n(0), poly_reduction(X_N_minus), omega(0), psi(0), n_inv(0), root_of_unity(0), inverse_root_of_unity(0), root_of_unity_size(0) {}

    // default pools: nttparameters.cu:84-142 of the reference (the range checks there are chained
    // comparisons that never fire; here they do)
    template <typename T> Modulus<T> NTTParameters<T>::modulus_pool()
    {
        customAssert(logn > 0 && logn <= max_logn<T>(), "LOGN should be in range 2^0 to 2^" + std::to_string(max_logn<T>()) + ".");
        if constexpr (std::is_same<T, Data32>::value)
            return Modulus32(469762049u);
        else
            return Modulus64(576460756061519873ULL);
    }
    template <typename T> T NTTParameters<T>::omega_pool()
    {
        const Modulus<T> q = modulus_pool();
        if constexpr (std::is_same<T, Data32>::value)
            return OPERATOR32::exp(900u, Data32(1) << (25 - logn), q);
        else
            return OPERATOR64::exp(229929041166717729ULL, Data64(1) << (28 - logn), q);
    }
    template <typename T> T NTTParameters<T>::psi_pool()
    {
        const Modulus<T> q = modulus_pool();
        if constexpr (std::is_same<T, Data32>::value)
            return OPERATOR32::exp(30u, Data32(1) << (25 - logn), q);
        else
            return OPERATOR64::exp(4517306222ULL, Data64(1) << (28 - logn), q);
    }
    template <typename T> void NTTParameters<T>::forward_root_of_unity_table_generator()
    {
        forward_root_of_unity_table = powers<T>(root_of_unity, static_cast<std::size_t>(root_of_unity_size), modulus);
    }
    template <typename T> void NTTParameters<T>::inverse_root_of_unity_table_generator()
    {
        inverse_root_of_unity_table = powers<T>(inverse_root_of_unity, static_cast<std::size_t>(root_of_unity_size), modulus);
    }
    template <typename T> void NTTParameters<T>::n_inverse_generator() { n_inv = OPERATOR<T>::modinv(n, modulus); }
    template <typename T> std::vector<Root<T>> NTTParameters<T>::gpu_root_of_unity_table_generator(std::vector<T> table)
    {
        return bit_reversed_copy(table);
    }

    // ---------------------------------------------------------------- NTTParameters4Step
    namespace
    {
        // nttparameters.cu:229-303 of the reference, index logn - 12
        const Data64 kPrime64[] = {576460752303415297ULL, 576460752303439873ULL, 576460752304439297ULL, 576460752308273153ULL,
                                   576460752308273153ULL, 576460752315482113ULL, 576460752315482113ULL, 576460752340123649ULL,
                                   576460752364240897ULL, 576460752475389953ULL, 576460752597024769ULL, 576460753024843777ULL,
                                   576460753175838721ULL};
        const Data64 kOmega64[] = {288482366111684746ULL, 37048445140799662ULL,  459782973201979845ULL, 64800917766465203ULL,
                                   425015386842055933ULL, 18734847765732801ULL,  119109113519742895ULL, 227584740857897520ULL,
                                   477282059544659462ULL, 570131728462077067ULL, 433594414095420776ULL, 219263994987749328ULL,
                                   189790554094222112ULL};
        const Data64 kPsi64[] = {238394956950829ULL, 54612008597396ULL, 8242615629351ULL, 16141297350887ULL, 3760097055997ULL,
                                 11571974431275ULL,  328867687796ULL,   2298846063117ULL, 731868219707ULL,   409596963254ULL,
                                 189266227206ULL,    31864818375ULL,    92067739764ULL};
        const Data32 kPrime32[] = {268460033u, 268582913u, 268664833u, 268369921u, 269221889u, 269221889u, 270532609u,
                                   270532609u, 270532609u, 377487361u, 377487361u, 469762049u, 469762049u};
        const Data32 kOmega32[] = {36747374u, 249229369u, 4092529u, 175218169u, 10653696u, 238764304u, 240100u,
                                   23104u,    179776u,    19321u,   38809u,     1600u,     169u};
        const Data32 kPsi32[] = {77090u, 15787u, 2023u, 13237u, 3264u, 15452u, 490u, 152u, 424u, 139u, 197u, 40u, 13u};
        // matrix_dimention(): nttparameters.cu:305-354
        const int kN1[] = {32, 32, 32, 64, 128, 32, 32, 32, 32, 64, 128, 128, 256};
        const int kN2[] = {128, 256, 512, 512, 512, 4096, 8192, 16384, 32768, 32768, 32768, 65536, 65536};
        inline void check_4step_logn(int logn) { customAssert(logn >= 12 && logn <= 24, "LOGN should be in range 12 to 24."); }
        inline int ilog2(int v)
        {
            int l = 0;
            while ((1 << l) < v) l++;
            return l;
        }
    } // namespace

    template <typename T> NTTParameters4Step<T>::NTTParameters4Step(int LOGN, ReductionPolynomial poly_reduce_type)
    {
        logn = LOGN;
        check_4step_logn(logn);
        n = T(1) << logn;
        poly_reduction = poly_reduce_type;
        modulus = modulus_pool();
        omega = omega_pool();
        psi = psi_pool();
        root_of_unity = (poly_reduce_type == X_N_minus) ? omega : psi;
        inverse_root_of_unity = OPERATOR<T>::modinv(root_of_unity, modulus);
        root_of_unity_size = (poly_reduce_type == X_N_minus) ? (T(1) << (logn - 1)) : (T(1) << logn);
        const std::vector<int> shape = matrix_dimention();
        n1 = shape[0];
        n2 = shape[1];
        small_forward_root_of_unity_table_generator();
        small_inverse_root_of_unity_table_generator();
        TW_forward_table_generator();
        TW_inverse_table_generator();
        n_inverse_generator();
        n_inverse_generator_gpu();
    }
    template <typename T> NTTParameters4Step<T>::NTTParameters4Step() : logn(0), n(0), poly_reduction(X_N_minus), omega(0), psi(0), n_inv(0), n_inv_gpu(0), root_of_unity(0), inverse_root_of_unity(0), root_of_unity_size(0), n1(0), n2(0) {}

    template <typename T> Modulus<T> NTTParameters4Step<T>::modulus_pool()
    {
        check_4step_logn(logn);
        if constexpr (std::is_same<T, Data32>::value)
            return Modulus32(kPrime32[logn - 12]);
        else
            return Modulus64(kPrime64[logn - 12]);
    }
    template <typename T> T NTTParameters4Step<T>::omega_pool()
    {
        check_4step_logn(logn);
        if constexpr (std::is_same<T, Data32>::value)
            return kOmega32[logn - 12];
        else
            return kOmega64[logn - 12];
    }
    template <typename T> T NTTParameters4Step<T>::psi_pool()
    {
        check_4step_logn(logn);
        if constexpr (std::is_same<T, Data32>::value)
            return kPsi32[logn - 12];
        else
            return kPsi64[logn - 12];
    }
    template <typename T> std::vector<int> NTTParameters4Step<T>::matrix_dimention()
    {
        if (logn < 12 || logn > 24) throw std::runtime_error("Invalid choice.\n");
        return {kN1[logn - 12], kN2[logn - 12]};
    }
    // natural-order tables of n1/2 (n2/2) powers of root^(n/n1) (root^(n/n2)): nttparameters.cu:356-380
    template <typename T> void NTTParameters4Step<T>::small_forward_root_of_unity_table_generator()
    {
        n1_based_root_of_unity_table = powers<T>(OPERATOR<T>::exp(root_of_unity, n / T(n1), modulus), n1 >> 1, modulus);
        n2_based_root_of_unity_table = powers<T>(OPERATOR<T>::exp(root_of_unity, n / T(n2), modulus), n2 >> 1, modulus);
    }
    template <typename T> void NTTParameters4Step<T>::small_inverse_root_of_unity_table_generator()
    {
        const T r1 = OPERATOR<T>::modinv(OPERATOR<T>::exp(root_of_unity, n / T(n1), modulus), modulus);
        const T r2 = OPERATOR<T>::modinv(OPERATOR<T>::exp(root_of_unity, n / T(n2), modulus), modulus);
        n1_based_inverse_root_of_unity_table = powers<T>(r1, n1 >> 1, modulus);
        n2_based_inverse_root_of_unity_table = powers<T>(r2, n2 >> 1, modulus);
    }
    // W[i * n2 + j] = root^(bitreverse(i, log2 n1) * j): nttparameters.cu:382-396
    template <typename T> void NTTParameters4Step<T>::TW_forward_table_generator()
    {
        const int lg1 = ilog2(n1);
        W_root_of_unity_table.resize(static_cast<std::size_t>(n1) * n2);
        for (int i = 0; i < n1; i++)
        {
            const T g = OPERATOR<T>::exp(root_of_unity, static_cast<T>(bitreverse(i, lg1)), modulus);
            T acc = 1;
            T* row = W_root_of_unity_table.data() + static_cast<std::size_t>(i) * n2;
            for (int j = 0; j < n2; j++)
            {
                row[j] = acc;
                acc = OPERATOR<T>::mult(acc, g, modulus);
            }
        }
    }
    // Winv[i * n2 + j] = inverse_root^(bitreverse(j, log2 n2) * i): nttparameters.cu:429-443
    template <typename T> void NTTParameters4Step<T>::TW_inverse_table_generator()
    {
        const int lg2 = ilog2(n2);
        W_inverse_root_of_unity_table.resize(static_cast<std::size_t>(n1) * n2);
        std::vector<T> g(n2);
        for (int j = 0; j < n2; j++) g[j] = OPERATOR<T>::exp(inverse_root_of_unity, static_cast<T>(bitreverse(j, lg2)), modulus);
        T* w = W_inverse_root_of_unity_table.data();
        for (int j = 0; j < n2; j++) w[j] = 1;
        for (int i = 1; i < n1; i++)
            for (int j = 0; j < n2; j++)
                w[static_cast<std::size_t>(i) * n2 + j] = OPERATOR<T>::mult(w[static_cast<std::size_t>(i - 1) * n2 + j], g[j], modulus);
    }
    template <typename T> void NTTParameters4Step<T>::n_inverse_generator() { n_inv = OPERATOR<T>::modinv(n, modulus); }
    template <typename T> void NTTParameters4Step<T>::n_inverse_generator_gpu() { n_inv_gpu = OPERATOR<T>::modinv(n, modulus); }
    template <typename T> std::vector<Root<T>> NTTParameters4Step<T>::gpu_root_of_unity_table_generator(std::vector<T> table)
    {
        return bit_reversed_copy(table);
    }

    template class NTTParameters<Data32>;
    template class NTTParameters<Data64>;
    template class NTTParameters4Step<Data32>;
    template class NTTParameters4Step<Data64>;
} // namespace gpuntt
