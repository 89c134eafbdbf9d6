// Compiled twice by tests/test_abi_host.py: once against the reference's headers, once against include/gpuntt.
// (1) prints the layout of every struct that crosses the boundary; (2) references every host entry point the
// reference library instantiates, so that `nm -u` of the object lists the mangled names a caller needs.
#include <cstddef>
#include <cstdio>

#include "gpuntt/ntt_4step/ntt_4step.cuh"
#include "gpuntt/ntt_merge/ntt.cuh"

using namespace gpuntt;

#define SHOW(T, m) std::printf(#T "." #m " %zu\n", offsetof(T, m))
template <typename T> void layout(const char* name)
{
    std::printf("%s Modulus %zu %zu | cfg %zu rns %zu c4 %zu r4 %zu\n", name, sizeof(Modulus<T>), alignof(Modulus<T>), sizeof(ntt_configuration<T>),
                sizeof(ntt_rns_configuration<T>), sizeof(ntt4step_configuration<T>), sizeof(ntt4step_rns_configuration<T>));
    std::printf("%s cfg %zu %zu %zu %zu %zu %zu %zu\n", name, offsetof(ntt_configuration<T>, n_power), offsetof(ntt_configuration<T>, ntt_type),
                offsetof(ntt_configuration<T>, ntt_layout), offsetof(ntt_configuration<T>, reduction_poly),
                offsetof(ntt_configuration<T>, zero_padding), offsetof(ntt_configuration<T>, mod_inverse), offsetof(ntt_configuration<T>, stream));
    std::printf("%s rns %zu %zu\n", name, offsetof(ntt_rns_configuration<T>, mod_inverse), offsetof(ntt_rns_configuration<T>, stream));
    std::printf("%s c4 %zu %zu %zu %zu\n", name, offsetof(ntt4step_configuration<T>, n_power), offsetof(ntt4step_configuration<T>, ntt_type),
                offsetof(ntt4step_configuration<T>, mod_inverse), offsetof(ntt4step_configuration<T>, stream));
    std::printf("%s mod %zu %zu %zu\n", name, offsetof(Modulus<T>, value), offsetof(Modulus<T>, bit), offsetof(Modulus<T>, mu));
}

template <typename T, typename TU> void use_io()
{
    T* s = nullptr;
    TU* u = nullptr;
    Modulus<TU> m;
    Modulus<TU>* mp = nullptr;
    ntt_configuration<TU> c{};
    ntt_rns_configuration<TU> r{};
    GPU_NTT<T>(s, u, u, m, c, 0);
    GPU_INTT<T>(u, s, u, m, c, 0);
    GPU_NTT<T>(s, u, u, mp, r, 0, 1);
    GPU_INTT<T>(u, s, u, mp, r, 0, 1);
}
template <typename T> void use_inplace()
{
    T* u = nullptr;
    Modulus<T> m;
    Modulus<T>* mp = nullptr;
    ntt_configuration<T> c{};
    ntt_rns_configuration<T> r{};
    int* order = nullptr;
    GPU_NTT_Inplace<T>(u, u, m, c, 0);
    GPU_INTT_Inplace<T>(u, u, m, c, 0);
    GPU_NTT_Inplace<T>(u, u, mp, r, 0, 1);
    GPU_INTT_Inplace<T>(u, u, mp, r, 0, 1);
    GPU_NTT_Modulus_Ordered<T>(u, u, u, mp, r, 0, 1, order);
    GPU_NTT_Modulus_Ordered_Inplace<T>(u, u, mp, r, 0, 1, order);
    GPU_NTT_Poly_Ordered<T>(u, u, u, mp, r, 0, 1, order);
    GPU_NTT_Poly_Ordered_Inplace<T>(u, u, mp, r, 0, 1, order);
    ntt4step_configuration<T> c4{};
    ntt4step_rns_configuration<T> r4{};
    GPU_Transpose<T>(u, u, 1, 1, 0, 0);
    GPU_4STEP_NTT<T>(u, u, u, u, u, m, c4, 0);
    GPU_4STEP_NTT<T>(u, u, u, u, u, mp, r4, 0, 1);
    NTTParameters<T> P(12, X_N_minus);
    NTTCPU<T> cpu(P);
    std::vector<T> v(1);
    cpu.ntt(v);
    cpu.intt(v);
    cpu.mult(v, v);
    schoolbook_poly_multiplication<T>(v, v, m, X_N_minus);
    NTTParameters4Step<T> P4(12, X_N_minus);
    NTT_4STEP_CPU<T> cpu4(P4);
    cpu4.ntt(v);
    cpu4.intt(v);
    cpu4.intt_first_transpose(v);
    P.gpu_root_of_unity_table_generator(v);
    P4.gpu_root_of_unity_table_generator(v);
    check_result<T>(u, u, 0);
}

int main(int argc, char**)
{
    layout<Data64>("u64");
    layout<Data32>("u32");
    std::printf("enums %d %d %d %d %d %d\n", (int) FORWARD, (int) INVERSE, (int) PerPolynomial, (int) PerCoefficient, (int) X_N_plus, (int) X_N_minus);
    if (argc > 100)
    {
        use_io<Data32, Data32>();
        use_io<Data64, Data64>();
        use_io<Data32s, Data32>();
        use_io<Data64s, Data64>();
        use_inplace<Data32>();
        use_inplace<Data64>();
        CudaDevice();
        customAssert(true, "");
        (void) bitreverse(1, 2);
    }
    return 0;
}
