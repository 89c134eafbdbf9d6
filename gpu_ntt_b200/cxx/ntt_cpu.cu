// gpu_ntt_b200/cxx/ntt_cpu.cu -- NTTCPU<T>, NTT_4STEP_CPU<T>, schoolbook_poly_multiplication.
//
// Host transforms with the reference's semantics (src/lib/ntt_merge/ntt_cpu.cu:10-185,
// src/lib/ntt_4step/ntt_4step_cpu.cu:33-299): radix-2 Cooley-Tukey forward (natural -> bit-reversed),
// Gentleman-Sande inverse with the final n^-1, twiddle of group i taken at bitreverse(i) of the
// natural-order power table.  Written on flat arrays with the bit-reversed table built once per call.
#include <stdexcept>

#include "gpuntt/ntt_4step/ntt_4step_cpu.cuh"
#include "gpuntt/ntt_merge/ntt_cpu.cuh"

namespace gpuntt
{
    namespace
    {
        inline int ilog2(std::size_t v)
        {
            int l = 0;
            while ((std::size_t(1) << l) < v) l++;
            return l;
        }
        // br[i] = table[bitreverse(i)]
        template <typename T> std::vector<T> reversed(const std::vector<T>& table)
        {
            const int lg = ilog2(table.size());
            std::vector<T> br(table.size());
            for (std::size_t i = 0; i < table.size(); i++) br[i] = table[bitreverse(static_cast<int>(i), lg)];
            return br;
        }
        // forward stages on a[0..n): group i of the stage with m groups uses br[off*m + i]
        // (off = 1: X^N+1 tables of n entries; off = 0: X^N-1 tables of n/2 entries)
        template <typename T> void forward_stages(T* a, std::size_t n, const std::vector<T>& br, int off, const Modulus<T>& q)
        {
            std::size_t t = n;
            for (std::size_t m = 1; m < n; m <<= 1)
            {
                t >>= 1;
                for (std::size_t i = 0; i < m; i++)
                {
                    const T s = br[off * m + i];
                    T* lo = a + 2 * i * t;
                    T* hi = lo + t;
                    for (std::size_t j = 0; j < t; j++)
                    {
                        const T u = lo[j];
                        const T v = OPERATOR<T>::mult(hi[j], s, q);
                        lo[j] = OPERATOR<T>::add(u, v, q);
                        hi[j] = OPERATOR<T>::sub(u, v, q);
                    }
                }
            }
        }
        template <typename T> void inverse_stages(T* a, std::size_t n, const std::vector<T>& br, int off, const Modulus<T>& q)
        {
            std::size_t t = 1;
            for (std::size_t m = n; m > 1; m >>= 1)
            {
                const std::size_t h = m >> 1;
                for (std::size_t i = 0; i < h; i++)
                {
                    const T s = br[off * h + i];
                    T* lo = a + 2 * i * t;
                    T* hi = lo + t;
                    for (std::size_t j = 0; j < t; j++)
                    {
                        const T u = lo[j];
                        const T v = hi[j];
                        lo[j] = OPERATOR<T>::add(u, v, q);
                        hi[j] = OPERATOR<T>::mult(OPERATOR<T>::sub(u, v, q), s, q);
                    }
                }
                t <<= 1;
            }
        }
    } // namespace

    template <typename T>
    std::vector<T> schoolbook_poly_multiplication(std::vector<T> a, std::vector<T> b, Modulus<T> modulus,
                                                  ReductionPolynomial reduction_poly)
    {
        if (reduction_poly != X_N_minus && reduction_poly != X_N_plus) throw std::runtime_error("Poly reduction type is not supported!");
        const std::size_t n = a.size();
        std::vector<T> wide(2 * n, 0);
        for (std::size_t i = 0; i < n; i++)
            for (std::size_t j = 0; j < n; j++) wide[i + j] = OPERATOR<T>::add(wide[i + j], OPERATOR<T>::mult(a[i], b[j], modulus), modulus);
        std::vector<T> r(n);
        for (std::size_t i = 0; i < n; i++)
            r[i] = (reduction_poly == X_N_minus) ? OPERATOR<T>::add(wide[i], wide[i + n], modulus) : OPERATOR<T>::sub(wide[i], wide[i + n], modulus);
        return r;
    }
    template std::vector<Data32> schoolbook_poly_multiplication<Data32>(std::vector<Data32>, std::vector<Data32>, Modulus<Data32>, ReductionPolynomial);
    template std::vector<Data64> schoolbook_poly_multiplication<Data64>(std::vector<Data64>, std::vector<Data64>, Modulus<Data64>, ReductionPolynomial);

    // ---------------------------------------------------------------- NTTCPU
    template <typename T> NTTCPU<T>::NTTCPU(NTTParameters<T> parameters_) : parameters(parameters_) {}

    template <typename T> std::vector<T> NTTCPU<T>::mult(std::vector<T>& input1, std::vector<T>& input2)
    {
        std::vector<T> out(static_cast<std::size_t>(parameters.n));
        for (std::size_t i = 0; i < out.size(); i++) out[i] = OPERATOR<T>::mult(input1[i], input2[i], parameters.modulus);
        return out;
    }
    template <typename T> std::vector<T> NTTCPU<T>::ntt(std::vector<T>& input)
    {
        std::vector<T> a = input;
        const int off = parameters.poly_reduction == X_N_minus ? 0 : 1;
        forward_stages<T>(a.data(), static_cast<std::size_t>(parameters.n), reversed(parameters.forward_root_of_unity_table), off, parameters.modulus);
        return a;
    }
    template <typename T> std::vector<T> NTTCPU<T>::intt(std::vector<T>& input)
    {
        std::vector<T> a = input;
        const int off = parameters.poly_reduction == X_N_minus ? 0 : 1;
        inverse_stages<T>(a.data(), static_cast<std::size_t>(parameters.n), reversed(parameters.inverse_root_of_unity_table), off, parameters.modulus);
        for (T& v : a) v = OPERATOR<T>::mult(v, parameters.n_inv, parameters.modulus);
        return a;
    }
    template class NTTCPU<Data32>;
    template class NTTCPU<Data64>;

    // ---------------------------------------------------------------- NTT_4STEP_CPU
    template <typename T> NTT_4STEP_CPU<T>::NTT_4STEP_CPU(NTTParameters4Step<T> parameters_) : parameters(parameters_) {}

    template <typename T> std::vector<T> NTT_4STEP_CPU<T>::mult(std::vector<T>& input1, std::vector<T>& input2)
    {
        std::vector<T> out(static_cast<std::size_t>(parameters.n));
        for (std::size_t i = 0; i < out.size(); i++) out[i] = OPERATOR<T>::mult(input1[i], input2[i], parameters.modulus);
        return out;
    }
    template <typename T> void NTT_4STEP_CPU<T>::core_ntt(std::vector<T>& input, std::vector<T> root_table, int log_size)
    {
        forward_stages<T>(input.data(), std::size_t(1) << log_size, reversed(root_table), 0, parameters.modulus);
    }
    template <typename T> void NTT_4STEP_CPU<T>::core_intt(std::vector<T>& input, std::vector<T> root_table, int log_size)
    {
        inverse_stages<T>(input.data(), std::size_t(1) << log_size, reversed(root_table), 0, parameters.modulus);
    }
    template <typename T> void NTT_4STEP_CPU<T>::product(std::vector<T>& input, std::vector<T> root_table, int log_size)
    {
        const std::size_t n = std::size_t(1) << log_size;
        for (std::size_t i = 0; i < n; i++) input[i] = OPERATOR<T>::mult(input[i], root_table[i], parameters.modulus);
    }
    template <typename T> std::vector<std::vector<T>> NTT_4STEP_CPU<T>::vector_to_matrix(const std::vector<T>& array, int rows, int cols)
    {
        std::vector<std::vector<T>> m(rows, std::vector<T>(cols));
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) m[i][j] = array[static_cast<std::size_t>(i) * cols + j];
        return m;
    }
    // rows of n1 elements filled in the order array[i + j*rows], i outer (ntt_4step_cpu.cu:227-243 of the reference)
    template <typename T> std::vector<std::vector<T>> NTT_4STEP_CPU<T>::vector_to_matrix_intt(const std::vector<T>& array, int rows, int cols)
    {
        std::vector<std::vector<T>> m(cols);
        for (auto& r : m) r.reserve(rows);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++)
                m[(static_cast<std::size_t>(i) * cols + j) / rows].push_back(array[static_cast<std::size_t>(i) + static_cast<std::size_t>(j) * rows]);
        return m;
    }
    template <typename T> std::vector<T> NTT_4STEP_CPU<T>::matrix_to_vector(const std::vector<std::vector<T>>& originalMatrix)
    {
        std::vector<T> v;
        for (const auto& r : originalMatrix) v.insert(v.end(), r.begin(), r.end());
        return v;
    }
    template <typename T> std::vector<std::vector<T>> NTT_4STEP_CPU<T>::transpose_matrix(const std::vector<std::vector<T>>& originalMatrix)
    {
        const std::size_t rows = originalMatrix.size(), cols = rows ? originalMatrix[0].size() : 0;
        std::vector<std::vector<T>> t(cols, std::vector<T>(rows));
        for (std::size_t i = 0; i < rows; i++)
            for (std::size_t j = 0; j < cols; j++) t[j][i] = originalMatrix[i][j];
        return t;
    }

    template <typename T> std::vector<T> NTT_4STEP_CPU<T>::ntt(std::vector<T>& input)
    {
        const int n1 = parameters.n1, n2 = parameters.n2;
        const int lg1 = ilog2(n1), lg2 = ilog2(n2);
        const Modulus<T>& q = parameters.modulus;
        const std::vector<T> br1 = reversed(parameters.n1_based_root_of_unity_table);
        const std::vector<T> br2 = reversed(parameters.n2_based_root_of_unity_table);
        // columns of the n1 x n2 view, gathered into contiguous vectors
        std::vector<T> cols(static_cast<std::size_t>(n1) * n2);
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++) cols[static_cast<std::size_t>(j) * n1 + i] = input[static_cast<std::size_t>(i) * n2 + j];
        for (int j = 0; j < n2; j++) forward_stages<T>(cols.data() + static_cast<std::size_t>(j) * n1, std::size_t(1) << lg1, br1, 0, q);
        std::vector<T> rows(static_cast<std::size_t>(n1) * n2);
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++)
            {
                const std::size_t k = static_cast<std::size_t>(i) * n2 + j;
                rows[k] = OPERATOR<T>::mult(cols[static_cast<std::size_t>(j) * n1 + i], parameters.W_root_of_unity_table[k], q);
            }
        for (int i = 0; i < n1; i++) forward_stages<T>(rows.data() + static_cast<std::size_t>(i) * n2, std::size_t(1) << lg2, br2, 0, q);
        std::vector<T> out(rows.size());
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++) out[static_cast<std::size_t>(j) * n1 + i] = rows[static_cast<std::size_t>(i) * n2 + j];
        return out;
    }
    template <typename T> std::vector<T> NTT_4STEP_CPU<T>::intt(std::vector<T>& input)
    {
        const int n1 = parameters.n1, n2 = parameters.n2;
        const int lg1 = ilog2(n1), lg2 = ilog2(n2);
        const Modulus<T>& q = parameters.modulus;
        const std::vector<T> br1 = reversed(parameters.n1_based_inverse_root_of_unity_table);
        const std::vector<T> br2 = reversed(parameters.n2_based_inverse_root_of_unity_table);
        std::vector<T> cols = intt_first_transpose(input);
        for (int j = 0; j < n2; j++) inverse_stages<T>(cols.data() + static_cast<std::size_t>(j) * n1, std::size_t(1) << lg1, br1, 0, q);
        std::vector<T> rows(static_cast<std::size_t>(n1) * n2);
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++)
            {
                const std::size_t k = static_cast<std::size_t>(i) * n2 + j;
                rows[k] = OPERATOR<T>::mult(cols[static_cast<std::size_t>(j) * n1 + i], parameters.W_inverse_root_of_unity_table[k], q);
            }
        for (int i = 0; i < n1; i++) inverse_stages<T>(rows.data() + static_cast<std::size_t>(i) * n2, std::size_t(1) << lg2, br2, 0, q);
        std::vector<T> out(rows.size());
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++) out[static_cast<std::size_t>(j) * n1 + i] = OPERATOR<T>::mult(rows[static_cast<std::size_t>(i) * n2 + j], parameters.n_inv, q);
        return out;
    }
    template <typename T> std::vector<T> NTT_4STEP_CPU<T>::intt_first_transpose(const std::vector<T>& input)
    {
        return matrix_to_vector(vector_to_matrix_intt(input, parameters.n1, parameters.n2));
    }
    template class NTT_4STEP_CPU<Data32>;
    template class NTT_4STEP_CPU<Data64>;
} // namespace gpuntt
