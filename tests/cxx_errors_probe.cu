// CPU-only (tests/test_host_classes.py): the C++ entry points throw what the reference's throw, before any CUDA work --
// std::invalid_argument("Invalid n_power range!") for n_power outside 1..28 (PerPolynomial) / 1..9 (PerCoefficient) and outside 10..28
// for the *_Ordered entry points, std::invalid_argument("Invalid ntt_layout!") for an unknown layout (ntt.cu:2088-2091, 2230-2233,
// 2253, 3607-3610 of the reference).  Compiled with g++ against include/gpuntt, linked against libntt-1.0.a.
#include <cstdio>
#include <stdexcept>
#include <string>

#include "gpuntt/ntt_merge/ntt.cuh"

using namespace gpuntt;

template <typename F> static int expect(const char* what, const char* message, F&& call)
{
    try
    {
        call();
    }
    catch (const std::invalid_argument& e)
    {
        const bool ok = std::string(e.what()) == message;
        std::printf("%-58s %s (\"%s\")\n", what, ok ? "ok" : "WRONG MESSAGE", e.what());
        return ok ? 0 : 1;
    }
    catch (const std::exception& e)
    {
        std::printf("%-58s WRONG TYPE (\"%s\")\n", what, e.what());
        return 1;
    }
    std::printf("%-58s NO EXCEPTION\n", what);
    return 1;
}

template <typename T, typename TS> static int run(const char* name)
{
    int bad = 0;
    T* u = nullptr;
    TS* s = nullptr;
    Modulus<T> m(static_cast<T>(469762049));
    Modulus<T>* mp = nullptr;
    int* order = nullptr;
    std::printf("-- %s\n", name);
    for (int n_power : {0, 29, -3})
    {
        ntt_configuration<T> c{};
        c.n_power = n_power;
        c.ntt_layout = PerPolynomial;
        c.reduction_poly = X_N_minus;
        ntt_rns_configuration<T> r{};
        r.n_power = n_power;
        r.ntt_layout = PerPolynomial;
        bad += expect("GPU_NTT n_power out of range", "Invalid n_power range!", [&] { GPU_NTT<T>(u, u, u, m, c, 1); });
        bad += expect("GPU_NTT (signed) n_power out of range", "Invalid n_power range!", [&] { GPU_NTT<TS>(s, u, u, m, c, 1); });
        bad += expect("GPU_INTT n_power out of range", "Invalid n_power range!", [&] { GPU_INTT<T>(u, u, u, m, c, 1); });
        bad += expect("GPU_NTT_Inplace n_power out of range", "Invalid n_power range!", [&] { GPU_NTT_Inplace<T>(u, u, m, c, 1); });
        bad += expect("GPU_INTT_Inplace n_power out of range", "Invalid n_power range!", [&] { GPU_INTT_Inplace<T>(u, u, m, c, 1); });
        bad += expect("GPU_NTT RNS n_power out of range", "Invalid n_power range!", [&] { GPU_NTT<T>(u, u, u, mp, r, 2, 2); });
        bad += expect("GPU_INTT RNS n_power out of range", "Invalid n_power range!", [&] { GPU_INTT<T>(u, u, u, mp, r, 2, 2); });
    }
    {
        ntt_configuration<T> c{};
        c.n_power = 10; // PerCoefficient stops at 9
        c.ntt_layout = PerCoefficient;
        bad += expect("GPU_NTT PerCoefficient n_power 10", "Invalid n_power range!", [&] { GPU_NTT<T>(u, u, u, m, c, 4); });
        bad += expect("GPU_INTT PerCoefficient n_power 10", "Invalid n_power range!", [&] { GPU_INTT<T>(u, u, u, m, c, 4); });
        c.n_power = 12;
        c.ntt_layout = static_cast<NTTLayout>(7);
        bad += expect("GPU_NTT unknown layout", "Invalid ntt_layout!", [&] { GPU_NTT<T>(u, u, u, m, c, 4); });
        bad += expect("GPU_INTT unknown layout", "Invalid ntt_layout!", [&] { GPU_INTT<T>(u, u, u, m, c, 4); });
    }
    {
        ntt_rns_configuration<T> r{};
        r.n_power = 9; // the ordered entry points start at 10
        r.ntt_layout = PerPolynomial;
        bad += expect("GPU_NTT_Modulus_Ordered n_power 9", "Invalid n_power range!", [&] { GPU_NTT_Modulus_Ordered<T>(u, u, u, mp, r, 2, 2, order); });
        bad += expect("GPU_NTT_Poly_Ordered_Inplace n_power 9", "Invalid n_power range!", [&] { GPU_NTT_Poly_Ordered_Inplace<T>(u, u, mp, r, 2, 2, order); });
    }
    return bad;
}

// No usable device (the build container): a VALID call cannot run, and there is no CPU fallback -- it must end in
// gpuntt::CudaException ("CUDA Error in <file> at line <n>: ...", common.cuh:20-50 of the reference).  The class is header-inline, so
// this also runs as a MIXED build: this file compiled against the reference's header, linked against this repository's library.
static int no_device()
{
    int count = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess && count > 0)
    {
        std::printf("-- a device is present: no-device section skipped\n");
        return 0;
    }
    static Data64 buf[4096], tab[2048];
    ntt_configuration<Data64> c{};
    c.n_power = 12;
    c.ntt_layout = PerPolynomial;
    c.reduction_poly = X_N_minus;
    c.stream = 0;
    Modulus<Data64> m(576460756061519873ULL);
    try
    {
        GPU_NTT_Inplace<Data64>(buf, tab, m, c, 1);
    }
    catch (const CudaException& e)
    {
        const bool ok = std::string(e.what()).rfind("CUDA Error in ", 0) == 0;
        std::printf("%-58s %s\n", "valid call without a device: gpuntt::CudaException", ok ? "ok" : "WRONG TEXT");
        return ok ? 0 : 1;
    }
    catch (const std::exception& e)
    {
        std::printf("valid call without a device: WRONG TYPE (\"%s\")\n", e.what());
        return 1;
    }
    std::printf("valid call without a device: NO EXCEPTION\n");
    return 1;
}

int main()
{
    int bad = run<Data64, Data64s>("Data64") + run<Data32, Data32s>("Data32") + no_device();
    std::printf("%s\n", bad ? "FAILED" : "cxx errors ok");
    return bad;
}
