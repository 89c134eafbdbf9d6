// oracle/ref_shim.cpp -- extern "C" window onto the UNMODIFIED reference CPU
// implementation (NTTParameters / NTTCPU / NTT_4STEP_CPU), whose sources are
// compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libgpuntt_ref_cpu.so.  TEST INFRASTRUCTURE ONLY: used to pin the C
// restatement (oracle/ntt_oracle.c), to generate tests/golden/, and as the
// "reference" CPU baseline of bench.py.  No reference source is copied here;
// this file only calls the reference's public classes.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "gpuntt/common/nttparameters.cuh"
#include "gpuntt/ntt_4step/ntt_4step_cpu.cuh"
#include "gpuntt/ntt_merge/ntt_cpu.cuh"

using namespace gpuntt;

namespace
{
    template <typename T> ReductionPolynomial rp(int poly)
    {
        return poly == 1 ? ReductionPolynomial::X_N_minus : ReductionPolynomial::X_N_plus;
    }

    template <typename T> void merge_params(int logn, int poly, uint64_t* scal, uint64_t* fwd,
                                            uint64_t* inv, uint64_t* fwd_br, uint64_t* inv_br)
    {
        NTTParameters<T> P(logn, rp<T>(poly));
        scal[0] = P.modulus.value;
        scal[1] = P.modulus.bit;
        scal[2] = P.modulus.mu;
        scal[3] = P.omega;
        scal[4] = P.psi;
        scal[5] = P.n_inv;
        scal[6] = P.root_of_unity;
        scal[7] = P.inverse_root_of_unity;
        scal[8] = P.root_of_unity_size;
        scal[9] = P.n;
        auto fb = P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table);
        auto ib = P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table);
        for (size_t i = 0; i < (size_t) P.root_of_unity_size; i++)
        {
            if (fwd) fwd[i] = P.forward_root_of_unity_table[i];
            if (inv) inv[i] = P.inverse_root_of_unity_table[i];
            if (fwd_br) fwd_br[i] = fb[i];
            if (inv_br) inv_br[i] = ib[i];
        }
    }

    template <typename T>
    void merge_transform(int logn, int poly, int inverse, const uint64_t* in, uint64_t* out,
                         int batch)
    {
        NTTParameters<T> P(logn, rp<T>(poly));
        NTTCPU<T> cpu(P);
        size_t n = (size_t) 1 << logn;
        std::vector<T> v(n);
        for (int b = 0; b < batch; b++)
        {
            for (size_t i = 0; i < n; i++) v[i] = (T) in[b * n + i];
            std::vector<T> r = inverse ? cpu.intt(v) : cpu.ntt(v);
            for (size_t i = 0; i < n; i++) out[b * n + i] = r[i];
        }
    }

    template <typename T> struct FourStepHolder
    {
        NTTParameters4Step<T> P;
        FourStepHolder(int logn, int poly) : P(logn, rp<T>(poly)) {}
    };
    // scal[12] = modulus, bit, mu, omega, psi, n_inv, root, inv_root, root_size, n, n1, n2
    // which: 0 n1 fwd, 1 n2 fwd, 2 W fwd, 3 n1 inv, 4 n2 inv, 5 W inv (natural order)
    template <typename T> static void fs_scal(FourStepHolder<T>* H, uint64_t* scal)
    {
        auto& P = H->P;
        scal[0] = P.modulus.value;
        scal[1] = P.modulus.bit;
        scal[2] = P.modulus.mu;
        scal[3] = P.omega;
        scal[4] = P.psi;
        scal[5] = P.n_inv;
        scal[6] = P.root_of_unity;
        scal[7] = P.inverse_root_of_unity;
        scal[8] = P.root_of_unity_size;
        scal[9] = P.n;
        scal[10] = P.n1;
        scal[11] = P.n2;
    }
    template <typename T>
    static uint64_t fs_table(FourStepHolder<T>* H, int which, uint64_t* out, int bitrev)
    {
        auto& P = H->P;
        std::vector<T>* t = nullptr;
        switch (which)
        {
            case 0: t = &P.n1_based_root_of_unity_table; break;
            case 1: t = &P.n2_based_root_of_unity_table; break;
            case 2: t = &P.W_root_of_unity_table; break;
            case 3: t = &P.n1_based_inverse_root_of_unity_table; break;
            case 4: t = &P.n2_based_inverse_root_of_unity_table; break;
            default: t = &P.W_inverse_root_of_unity_table; break;
        }
        if (out)
        {
            if (bitrev)
            {
                auto br = P.gpu_root_of_unity_table_generator(*t);
                for (size_t i = 0; i < br.size(); i++) out[i] = br[i];
            }
            else
                for (size_t i = 0; i < t->size(); i++) out[i] = (*t)[i];
        }
        return t->size();
    }
    template <typename T>
    static void fs_run(FourStepHolder<T>* H, int op, const uint64_t* in, uint64_t* out)
    {
        NTT_4STEP_CPU<T> cpu(H->P);
        size_t n = (size_t) H->P.n;
        std::vector<T> v(n);
        for (size_t i = 0; i < n; i++) v[i] = (T) in[i];
        std::vector<T> r = (op == 0) ? cpu.ntt(v) : (op == 1) ? cpu.intt(v)
                                                              : cpu.intt_first_transpose(v);
        for (size_t i = 0; i < n; i++) out[i] = r[i];
    }
    // ---- CPU baseline: time NTTCPU<T>::ntt over `count` polynomials on `threads`
    // host threads (one NTTCPU per thread, polynomials sharded contiguously).
    // Returns seconds of wall time for the transform part only.
    template <typename T>
    static double time_merge(int logn, int poly, int count, int threads, uint32_t seed)
    {
        NTTParameters<T> P(logn, rp<T>(poly));
        size_t n = (size_t) 1 << logn;
        std::mt19937 gen(seed);
        std::uniform_int_distribution<std::uint64_t> dis(0, P.modulus.value - 1);
        std::vector<std::vector<T>> in(count, std::vector<T>(n));
        for (auto& v : in)
            for (auto& x : v) x = (T) dis(gen);
        std::vector<uint64_t> sink(threads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back(
                [&, t]()
                {
                    NTTCPU<T> cpu(P);
                    for (int b = t; b < count; b += threads)
                    {
                        std::vector<T> r = cpu.ntt(in[b]);
                        sink[t] += r[0];
                    }
                });
        for (auto& x : th) x.join();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
    // Same as time_merge on caller-provided polynomials, results kept: bench.py's cpu_baseline leg times this and then
    // uses `out` to check the GPU's first step (the parity gate).
    template <typename T>
    static double time_merge_io(int logn, int poly, int count, int threads, const uint64_t* in, uint64_t* out)
    {
        NTTParameters<T> P(logn, rp<T>(poly));
        size_t n = (size_t) 1 << logn;
        std::vector<std::vector<T>> v(count, std::vector<T>(n));
        for (int b = 0; b < count; b++)
            for (size_t i = 0; i < n; i++) v[b][i] = (T) in[(size_t) b * n + i];
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back(
                [&, t]()
                {
                    NTTCPU<T> cpu(P);
                    for (int b = t; b < count; b += threads)
                    {
                        std::vector<T> r = cpu.ntt(v[b]);
                        for (size_t i = 0; i < n; i++) out[(size_t) b * n + i] = r[i];
                    }
                });
        for (auto& x : th) x.join();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
} // namespace

extern "C"
{
    double ref_time_merge_ntt_io(int logn, int poly, int width, int count, int threads, const uint64_t* in, uint64_t* out)
    {
        if (width == 32) return time_merge_io<Data32>(logn, poly, count, threads, in, out);
        return time_merge_io<Data64>(logn, poly, count, threads, in, out);
    }

    // scal[10]; tables may be NULL. width = 32 or 64.
    void ref_merge_params(int logn, int poly, int width, uint64_t* scal, uint64_t* fwd,
                          uint64_t* inv, uint64_t* fwd_br, uint64_t* inv_br)
    {
        if (width == 32)
            merge_params<Data32>(logn, poly, scal, fwd, inv, fwd_br, inv_br);
        else
            merge_params<Data64>(logn, poly, scal, fwd, inv, fwd_br, inv_br);
    }

    void ref_merge_transform(int logn, int poly, int width, int inverse, const uint64_t* in,
                             uint64_t* out, int batch)
    {
        if (width == 32)
            merge_transform<Data32>(logn, poly, inverse, in, out, batch);
        else
            merge_transform<Data64>(logn, poly, inverse, in, out, batch);
    }

    void ref_schoolbook(const uint64_t* a, const uint64_t* b, int n, uint64_t p, int poly,
                        uint64_t* out)
    {
        std::vector<Data64> va(a, a + n), vb(b, b + n);
        auto r = schoolbook_poly_multiplication<Data64>(va, vb, Modulus<Data64>(p),
                                                        rp<Data64>(poly));
        for (int i = 0; i < n; i++) out[i] = r[i];
    }

    uint64_t ref_barrett_mult(uint64_t a, uint64_t b, uint64_t p, int width)
    {
        if (width == 32)
            return OPERATOR<Data32>::mult((Data32) a, (Data32) b, Modulus<Data32>((Data32) p));
        return OPERATOR<Data64>::mult(a, b, Modulus<Data64>(p));
    }

    // std::mt19937(seed) + uniform_int_distribution<uint64_t>(0, p-1), poly-major:
    // exactly the generator of example/ntt_merge/test_merge_ntt.cu:70-85.
    void ref_example_input(uint32_t seed, uint64_t p, uint64_t count, uint64_t* out)
    {
        std::mt19937 gen(seed);
        std::uniform_int_distribution<std::uint64_t> dis(0, p - 1);
        for (uint64_t i = 0; i < count; i++) out[i] = dis(gen);
    }

    // ---- 4-step: an opaque handle because the W table build is expensive ----
    void* ref_4step_new(int logn, int poly, int width)
    {
        if (width == 32) return new FourStepHolder<Data32>(logn, poly);
        return new FourStepHolder<Data64>(logn, poly);
    }
    void ref_4step_free(void* h, int width)
    {
        if (width == 32)
            delete (FourStepHolder<Data32>*) h;
        else
            delete (FourStepHolder<Data64>*) h;
    }
    void ref_4step_scalars(void* h, int width, uint64_t* scal)
    {
        if (width == 32)
            fs_scal((FourStepHolder<Data32>*) h, scal);
        else
            fs_scal((FourStepHolder<Data64>*) h, scal);
    }
    uint64_t ref_4step_table(void* h, int width, int which, uint64_t* out, int bitrev)
    {
        if (width == 32) return fs_table((FourStepHolder<Data32>*) h, which, out, bitrev);
        return fs_table((FourStepHolder<Data64>*) h, which, out, bitrev);
    }
    // op: 0 ntt, 1 intt, 2 intt_first_transpose
    void ref_4step_run(void* h, int width, int op, const uint64_t* in, uint64_t* out)
    {
        if (width == 32)
            fs_run((FourStepHolder<Data32>*) h, op, in, out);
        else
            fs_run((FourStepHolder<Data64>*) h, op, in, out);
    }

    double ref_time_merge_ntt(int logn, int poly, int width, int count, int threads,
                              uint32_t seed)
    {
        if (width == 32) return time_merge<Data32>(logn, poly, count, threads, seed);
        return time_merge<Data64>(logn, poly, count, threads, seed);
    }
    int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }
}
