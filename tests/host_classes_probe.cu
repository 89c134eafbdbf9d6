// Compiled twice by tests/test_host_classes.py (host only, g++ -x c++): once against the reference's headers together with the
// reference's own CPU sources, once against include/gpuntt and gpu_ntt_b200/lib/libntt-1.0.a.  Prints every public value of the host
// classes a GPU-NTT caller builds its tables with (NTTParameters, NTTParameters4Step: nttparameters.cuh:56-104, nttparameters.cu:22-471
// of the reference) and hashes of what the CPU classes compute (NTTCPU, NTT_4STEP_CPU, schoolbook_poly_multiplication:
// ntt_cpu.cu:81-185, ntt_4step_cpu.cu:33-299); the two outputs must be identical, line for line.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "gpuntt/ntt_4step/ntt_4step_cpu.cuh"
#include "gpuntt/ntt_merge/ntt_cpu.cuh"

using namespace gpuntt;
typedef unsigned long long ull;

template <typename V> static ull fold(const V& v)
{
    ull h = 1469598103934665603ull;
    for (auto x : v) h = h * 1000003ull + (ull) x;
    return h;
}
template <typename T> static std::vector<T> input(size_t count, T p, ull seed)
{
    std::vector<T> v(count);
    ull s = seed * 6364136223846793005ull + 1442695040888963407ull;
    for (auto& x : v)
    {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        x = (T) ((s >> 3) % (ull) p);
    }
    return v;
}

template <typename T> static void merge(const char* name, int logn, ReductionPolynomial poly)
{
    NTTParameters<T> P(logn, poly);
    std::printf("%s merge logn=%d poly=%d | n=%llu p=%llu bit=%llu mu=%llu omega=%llu psi=%llu n_inv=%llu root=%llu iroot=%llu size=%llu\n", name,
                P.logn, (int) P.poly_reduction, (ull) P.n, (ull) P.modulus.value, (ull) P.modulus.bit, (ull) P.modulus.mu, (ull) P.omega,
                (ull) P.psi, (ull) P.n_inv, (ull) P.root_of_unity, (ull) P.inverse_root_of_unity, (ull) P.root_of_unity_size);
    std::printf("  tables %zu %llu | %zu %llu | gpu %llu %llu\n", P.forward_root_of_unity_table.size(), fold(P.forward_root_of_unity_table),
                P.inverse_root_of_unity_table.size(), fold(P.inverse_root_of_unity_table),
                fold(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table)),
                fold(P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table)));
    NTTCPU<T> cpu(P);
    std::vector<T> a = input<T>((size_t) 1 << logn, P.modulus.value, 11 + logn), b = input<T>((size_t) 1 << logn, P.modulus.value, 97 + logn);
    a[0] = P.modulus.value - 1;
    std::vector<T> fa = cpu.ntt(a), fb = cpu.ntt(b);
    std::vector<T> ia = cpu.intt(a);
    std::vector<T> prod = cpu.mult(fa, fb);
    std::vector<T> back = cpu.intt(prod);
    std::printf("  ntt %llu %llu intt %llu mult %llu conv %llu roundtrip %d\n", fold(fa), fold(fb), fold(ia), fold(prod), fold(back),
                (int) (cpu.intt(fa) == a));
    if (logn <= 10)
    {
        std::vector<T> sb = schoolbook_poly_multiplication<T>(a, b, P.modulus, poly);
        sb.resize((size_t) 1 << logn); // (the reduced product)
        std::printf("  schoolbook %llu equals_conv %d\n", fold(sb), (int) (sb == back));
    }
}

template <typename T> static void fourstep(const char* name, int logn, bool transforms = true)
{
    NTTParameters4Step<T> P(logn, X_N_minus);
    std::printf("%s 4step logn=%d | n=%llu p=%llu bit=%llu mu=%llu omega=%llu psi=%llu n_inv=%llu n_inv_gpu=%llu root=%llu iroot=%llu size=%llu n1=%d n2=%d\n",
                name, P.logn, (ull) P.n, (ull) P.modulus.value, (ull) P.modulus.bit, (ull) P.modulus.mu, (ull) P.omega, (ull) P.psi, (ull) P.n_inv,
                (ull) P.n_inv_gpu, (ull) P.root_of_unity, (ull) P.inverse_root_of_unity, (ull) P.root_of_unity_size, P.n1, P.n2);
    std::printf("  tables n1 %zu %llu n2 %zu %llu W %zu %llu | inverse n1 %zu %llu n2 %zu %llu W %zu %llu | gpu %llu\n",
                P.n1_based_root_of_unity_table.size(), fold(P.n1_based_root_of_unity_table), P.n2_based_root_of_unity_table.size(),
                fold(P.n2_based_root_of_unity_table), P.W_root_of_unity_table.size(), fold(P.W_root_of_unity_table),
                P.n1_based_inverse_root_of_unity_table.size(), fold(P.n1_based_inverse_root_of_unity_table),
                P.n2_based_inverse_root_of_unity_table.size(), fold(P.n2_based_inverse_root_of_unity_table),
                P.W_inverse_root_of_unity_table.size(), fold(P.W_inverse_root_of_unity_table),
                fold(P.gpu_root_of_unity_table_generator(P.n2_based_root_of_unity_table)));
    if (!transforms) return; // (the large shapes: parameters and tables only)
    NTT_4STEP_CPU<T> cpu(P);
    std::vector<T> a = input<T>((size_t) 1 << logn, P.modulus.value, 5 + logn), b = input<T>((size_t) 1 << logn, P.modulus.value, 55 + logn);
    a[1] = P.modulus.value - 1;
    std::vector<T> fa = cpu.ntt(a), fb = cpu.ntt(b);
    std::vector<T> ia = cpu.intt(a);
    std::vector<T> prod = cpu.mult(fa, fb);
    std::vector<T> ft = cpu.intt_first_transpose(a);
    std::printf("  ntt %llu %llu intt %llu mult %llu first_transpose %llu roundtrip %d\n", fold(fa), fold(fb), fold(ia), fold(prod), fold(ft),
                (int) (cpu.intt(fa) == a));
}

// Modulus<T>(value) and the host Barrett helpers (modular_arith.cuh:28-158 of the reference) on canonical operands
template <typename T> static void arith(const char* name, ull value)
{
    Modulus<T> m((T) value);
    std::vector<T> a = input<T>(20000, (T) value, value % 1000), b = input<T>(20000, (T) value, value % 777 + 3);
    a[0] = (T) (value - 1);
    b[0] = (T) (value - 1);
    a[1] = 0;
    b[2] = 1;
    ull h_add = 7, h_sub = 7, h_mul = 7, h_red = 7, h_exp = 7, h_inv = 7;
    for (size_t i = 0; i < a.size(); i++)
    {
        h_add = h_add * 1000003ull + (ull) OPERATOR<T>::add(a[i], b[i], m);
        h_sub = h_sub * 1000003ull + (ull) OPERATOR<T>::sub(a[i], b[i], m);
        h_mul = h_mul * 1000003ull + (ull) OPERATOR<T>::mult(a[i], b[i], m);
        h_red = h_red * 1000003ull + (ull) OPERATOR<T>::reduce(a[i], m);
        if (i < 300)
        {
            h_exp = h_exp * 1000003ull + (ull) OPERATOR<T>::exp(a[i], b[i], m);
            if (a[i] != 0) h_inv = h_inv * 1000003ull + (ull) OPERATOR<T>::modinv(a[i], m);
        }
    }
    std::printf("%s modulus %llu bit=%llu mu=%llu | add %llu sub %llu mult %llu reduce %llu exp %llu modinv %llu\n", name, (ull) m.value,
                (ull) m.bit, (ull) m.mu, h_add, h_sub, h_mul, h_red, h_exp, h_inv);
}

// NTTParameters(LOGN, NTTFactors, poly): the caller's own prime and roots (nttparameters.cu:51-78 of the reference)
template <typename T> static void factors(const char* name, int logn, ReductionPolynomial poly)
{
    NTTParameters<T> D(logn, poly);
    NTTFactors<T> f(D.modulus, D.omega, D.psi);
    NTTParameters<T> P(logn, f, poly);
    std::printf("%s factors logn=%d poly=%d | n=%llu p=%llu omega=%llu psi=%llu n_inv=%llu root=%llu iroot=%llu size=%llu tables %llu %llu same_as_default %d\n",
                name, P.logn, (int) P.poly_reduction, (ull) P.n, (ull) P.modulus.value, (ull) P.omega, (ull) P.psi, (ull) P.n_inv,
                (ull) P.root_of_unity, (ull) P.inverse_root_of_unity, (ull) P.root_of_unity_size, fold(P.forward_root_of_unity_table),
                fold(P.inverse_root_of_unity_table), (int) (P.forward_root_of_unity_table == D.forward_root_of_unity_table));
}

int main()
{
    for (ull v : {12289ull, 65537ull, 469762049ull, 536870909ull, 1073741789ull /* largest prime below 2^30 */})
        arith<Data32>("u32", v);
    // (largest primes below 2^60 - 2^20, 2^61 - 2^20, 2^62 - 2^20.  Not closer to the power of two: the reference derives Modulus::bit
    // from a floating-point log2, modular_arith.cuh:46, which rounds values within 2^(k-54) of 2^k up to k + 1 bits -- 2^61 - 1 gets
    // bit = 62 and a mu that overflows 64 bits there; this library uses the exact bit length)
    for (ull v : {12289ull, 469762049ull, 1099511627689ull, 576460756061519873ull, 1152921504605798393ull, 2305843009212645239ull,
                  4611686018426339311ull})
        arith<Data64>("u64", v);
    for (int logn : {4, 11, 14})
        for (ReductionPolynomial poly : {X_N_minus, X_N_plus})
        {
            factors<Data64>("u64", logn, poly);
            factors<Data32>("u32", logn, poly);
        }
    for (int logn : {1, 2, 3, 5, 8, 10, 12, 13})
        for (ReductionPolynomial poly : {X_N_minus, X_N_plus})
        {
            merge<Data64>("u64", logn, poly);
            merge<Data32>("u32", logn, poly);
        }
    for (int logn : {12, 13, 14, 15, 16, 17})
    {
        fourstep<Data64>("u64", logn);
        fourstep<Data32>("u32", logn);
    }
    // every remaining shape of matrix_dimention() (nttparameters.cu:305-354), up to BASELINE config C4 (2^24 = 256 x 65536)
    for (int logn : {18, 19, 20, 21, 22, 23, 24}) fourstep<Data64>("u64", logn, false);
    for (int logn : {18, 20, 22, 24}) fourstep<Data32>("u32", logn, false);
    std::printf("bitreverse %d %d %d\n", bitreverse(1, 4), bitreverse(6, 3), bitreverse(1234, 12));
    return 0;
}
