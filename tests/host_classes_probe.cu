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

template <typename T> static void fourstep(const char* name, int logn)
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

int main()
{
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
    std::printf("bitreverse %d %d %d\n", bitreverse(1, 4), bitreverse(6, 3), bitreverse(1234, 12));
    return 0;
}
