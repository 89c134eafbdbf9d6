// tools/api_bench.cu -- one caller, two libraries.
//
// A program written ONLY against the public GPU-NTT host API ("gpuntt/ntt_merge/ntt.cuh",
// "gpuntt/ntt_4step/ntt_4step.cuh": NTTParameters, NTTCPU, GPU_NTT_Inplace, GPU_INTT_Inplace, GPU_4STEP_NTT,
// GPU_Transpose).  tools/build_api_bench.sh compiles this same file twice:
//   * against the reference's headers + the reference's own GPU sources built for sm_100 (oracle/_ref/libntt_ref_gpu.a,
//     the "same-box GPU baseline" of BASELINE.md section 4 item 4), and
//   * against include/gpuntt + gpu_ntt_b200/lib/libntt-1.0.a (this repository),
// so the two binaries differ in nothing but the library behind the API.  Each case checks polynomial 0 and the last
// polynomial against the library's own NTTCPU (bit-exact) before it is timed with CUDA events.
//
//   api_bench <label> [c2|c2inv|c3|c4|sweep|small|fhe|latency ...]     one JSON object per line
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "gpuntt/ntt_4step/ntt_4step.cuh"
#include "gpuntt/ntt_4step/ntt_4step_cpu.cuh"
#include "gpuntt/ntt_merge/ntt.cuh"

using namespace gpuntt;

static const char* g_label = "?";

// A/B runs: GPUNTT_TUNE="knob=value,knob=value" is applied through gpuntt_b200_tune when the library behind the API has it
// (this repository's; the reference build leaves the weak symbol null)
extern "C" void gpuntt_b200_tune(int knob, int value) __attribute__((weak));
static void apply_tune_env()
{
    const char* e = getenv("GPUNTT_TUNE");
    if (!e || !gpuntt_b200_tune) return;
    std::string s(e);
    size_t pos = 0;
    while (pos < s.size())
    {
        size_t c = s.find(',', pos);
        if (c == std::string::npos) c = s.size();
        const std::string kv = s.substr(pos, c - pos);
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) gpuntt_b200_tune(atoi(kv.substr(0, eq).c_str()), atoi(kv.substr(eq + 1).c_str()));
        pos = c + 1;
    }
}

#define CK(x)                                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                   \
            exit(2);                                                                                                   \
        }                                                                                                              \
    } while (0)

template <typename F> static double time_ms(F&& fn, int iters, int warm = 3)
{
    for (int i = 0; i < warm; i++)
        fn();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> t;
    for (int i = 0; i < iters; i++)
    {
        CK(cudaEventRecord(e0, 0));
        fn();
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

template <typename T> static std::vector<T> random_poly(size_t n, T p, unsigned seed)
{
    std::mt19937 gen(seed);
    std::uniform_int_distribution<unsigned long long> dis(0, (unsigned long long) p - 1);
    std::vector<T> v(n);
    for (auto& x : v)
        x = (T) dis(gen);
    return v;
}

// Merge-NTT, single modulus, in place (the call BASELINE.json's metric is quoted on).
template <typename T>
static void merge_case(const char* name, int logn, int batch, bool inverse, bool round_trip, int iters,
                       ReductionPolynomial ring = ReductionPolynomial::X_N_minus, int mod_count = 0)
{
    NTTParameters<T> P(logn, ring);
    NTTCPU<T> cpu(P);
    const size_t n = (size_t) 1 << logn;
    std::vector<T> a = random_poly<T>(n, P.modulus.value, 0), b = random_poly<T>(n, P.modulus.value, 1);
    T* d;
    CK(cudaMalloc(&d, sizeof(T) * n * batch));
    auto fill = [&]() {
        for (int j = 0; j < batch; j++)
            CK(cudaMemcpy(d + n * j, (j == batch - 1 ? b : a).data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    };
    fill();
    std::vector<Root<T>> ft = P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table);
    std::vector<Root<T>> it = P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table);
    Root<T>*dft, *dit;
    CK(cudaMalloc(&dft, sizeof(T) * ft.size()));
    CK(cudaMalloc(&dit, sizeof(T) * it.size()));
    CK(cudaMemcpy(dft, ft.data(), sizeof(T) * ft.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dit, it.data(), sizeof(T) * it.size(), cudaMemcpyHostToDevice));
    ntt_configuration<T> cf = {.n_power = logn, .ntt_type = FORWARD, .ntt_layout = PerPolynomial,
                               .reduction_poly = ring, .zero_padding = false, .stream = 0};
    ntt_configuration<T> ci = {.n_power = logn, .ntt_type = INVERSE, .ntt_layout = PerPolynomial,
                               .reduction_poly = ring, .zero_padding = false,
                               .mod_inverse = P.n_inv, .stream = 0};
    // RNS overloads (mod_count > 0): mod_count slots that all hold the pooled prime of this ring size (the timing does not
    // depend on the moduli being distinct; polynomial b uses slot b % mod_count, ntt.cu:613-619), tables 1 << n_power apart
    Modulus<T>* dmod = nullptr;
    Ninverse<T>* dninv = nullptr;
    Root<T>*rft = nullptr, *rit = nullptr;
    if (mod_count > 0)
    {
        std::vector<Modulus<T>> hm(mod_count, P.modulus);
        std::vector<Ninverse<T>> hn(mod_count, P.n_inv);
        CK(cudaMalloc(&dmod, sizeof(Modulus<T>) * mod_count));
        CK(cudaMalloc(&dninv, sizeof(Ninverse<T>) * mod_count));
        CK(cudaMemcpy(dmod, hm.data(), sizeof(Modulus<T>) * mod_count, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dninv, hn.data(), sizeof(Ninverse<T>) * mod_count, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&rft, sizeof(T) * n * mod_count));
        CK(cudaMalloc(&rit, sizeof(T) * n * mod_count));
        for (int m = 0; m < mod_count; m++)
        {
            CK(cudaMemcpy(rft + n * m, ft.data(), sizeof(T) * ft.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(rit + n * m, it.data(), sizeof(T) * it.size(), cudaMemcpyHostToDevice));
        }
    }
    ntt_rns_configuration<T> rf = {.n_power = logn, .ntt_type = FORWARD, .ntt_layout = PerPolynomial,
                                   .reduction_poly = ring, .zero_padding = false, .stream = 0};
    ntt_rns_configuration<T> ri = {.n_power = logn, .ntt_type = INVERSE, .ntt_layout = PerPolynomial,
                                   .reduction_poly = ring, .zero_padding = false, .mod_inverse = dninv, .stream = 0};
    auto fwd = [&]() {
        if (mod_count > 0)
            GPU_NTT_Inplace(d, rft, dmod, rf, batch, mod_count);
        else
            GPU_NTT_Inplace(d, dft, P.modulus, cf, batch);
    };
    auto inv = [&]() {
        if (mod_count > 0)
            GPU_INTT_Inplace(d, rit, dmod, ri, batch, mod_count);
        else
            GPU_INTT_Inplace(d, dit, P.modulus, ci, batch);
    };
    // parity against the library's own CPU transform
    bool ok = true;
    std::vector<T> h(n);
    auto cmp = [&](int j, const std::vector<T>& want) {
        CK(cudaMemcpy(h.data(), d + n * j, sizeof(T) * n, cudaMemcpyDeviceToHost));
        ok = ok && (memcmp(h.data(), want.data(), sizeof(T) * n) == 0);
    };
    if (inverse && !round_trip)
    {
        inv();
        cmp(0, cpu.intt(a));
        cmp(batch - 1, cpu.intt(b));
    }
    else
    {
        fwd();
        cmp(0, cpu.ntt(a));
        cmp(batch - 1, cpu.ntt(b));
        if (round_trip)
        {
            inv();
            cmp(0, a);
            cmp(batch - 1, b);
        }
    }
    double ms = round_trip ? time_ms([&]() { fwd(); inv(); }, iters) : (inverse ? time_ms(inv, iters) : time_ms(fwd, iters));
    double ntts = (double) batch * (round_trip ? 2 : 1);
    double gbs = 2.0 * n * sizeof(T) * ntts / (ms * 1e-3) / 1e9;
    // launch-bound cases: what one call costs the HOST thread (enqueue only, nothing waited for) and what a call costs in a
    // stream of back-to-back calls (the larger of host and device time per call)
    double host_us = 0, stream_us = 0;
    if (batch <= 128 && !round_trip)
    {
        const int reps = 300;
        auto one = [&]() { if (inverse) inv(); else fwd(); };
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, 0));
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; i++) one();
        auto t1 = std::chrono::steady_clock::now();
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float sm = 0;
        CK(cudaEventElapsedTime(&sm, e0, e1));
        host_us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
        stream_us = sm * 1e3 / reps;
    }
    printf("{\"lib\": \"%s\", \"case\": \"%s\", \"bits\": %d, \"logn\": %d, \"batch\": %d, \"ring\": \"%s\", \"mod_count\": %d, \"op\": \"%s\", "
           "\"parity_vs_NTTCPU\": %s, \"host_us_per_call\": %.2f, \"stream_us_per_call\": %.2f, \"ms\": %.4f, \"ntt_per_s\": %.1f, \"alg_GBps\": %.1f}\n",
           g_label, name, (int) sizeof(T) * 8, logn, batch, ring == ReductionPolynomial::X_N_minus ? "X^N-1" : "X^N+1", mod_count, round_trip ? "fwd+inv" : (inverse ? "inv" : "fwd"), ok ? "true" : "false",
           host_us, stream_us, ms, ntts / (ms * 1e-3), gbs);
    fflush(stdout);
    cudaFree(d);
    cudaFree(dft);
    cudaFree(dit);
    cudaFree(dmod);
    cudaFree(dninv);
    cudaFree(rft);
    cudaFree(rit);
}

// 4-step, reference contract: transpose, GPU_4STEP_NTT, transpose (example/ntt_4step/test_4step_ntt.cu:147-154).
static void fourstep_case(const char* name, int logn, int batch, int iters)
{
    typedef Data64 T;
    NTTParameters4Step<T> P(logn, ReductionPolynomial::X_N_minus);
    NTT_4STEP_CPU<T> cpu(P);
    const size_t n = (size_t) 1 << logn;
    std::vector<T> a = random_poly<T>(n, P.modulus.value, 0);
    T *din, *dout;
    CK(cudaMalloc(&din, sizeof(T) * n * batch));
    CK(cudaMalloc(&dout, sizeof(T) * n * batch));
    for (int j = 0; j < batch; j++)
        CK(cudaMemcpy(din + n * j, a.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    std::vector<Root<T>> t1 = P.gpu_root_of_unity_table_generator(P.n1_based_root_of_unity_table);
    std::vector<Root<T>> t2 = P.gpu_root_of_unity_table_generator(P.n2_based_root_of_unity_table);
    Root<T>*d1, *d2, *dw;
    CK(cudaMalloc(&d1, sizeof(T) * (P.n1 >> 1)));
    CK(cudaMalloc(&d2, sizeof(T) * (P.n2 >> 1)));
    CK(cudaMalloc(&dw, sizeof(T) * n));
    CK(cudaMemcpy(d1, t1.data(), sizeof(T) * (P.n1 >> 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, t2.data(), sizeof(T) * (P.n2 >> 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, P.W_root_of_unity_table.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    Modulus<T>* dmod;
    Ninverse<T>* dninv;
    Modulus<T> hm[1] = {P.modulus};
    Ninverse<T> hn[1] = {P.n_inv};
    CK(cudaMalloc(&dmod, sizeof(hm)));
    CK(cudaMalloc(&dninv, sizeof(hn)));
    CK(cudaMemcpy(dmod, hm, sizeof(hm), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dninv, hn, sizeof(hn), cudaMemcpyHostToDevice));
    ntt4step_rns_configuration<T> cfg = {.n_power = logn, .ntt_type = FORWARD, .mod_inverse = dninv, .stream = 0};
    T* dtmp;
    CK(cudaMalloc(&dtmp, sizeof(T) * n * batch));
    auto run = [&]() {
        GPU_Transpose(din, dtmp, P.n1, P.n2, P.logn, batch);
        GPU_4STEP_NTT(dtmp, dout, d1, d2, dw, dmod, cfg, batch, 1);
        GPU_Transpose(dout, dtmp, P.n1, P.n2, P.logn, batch);
    };
    run();
    std::vector<T> h(n), want = cpu.ntt(a);
    bool ok = true;
    for (int j : {0, batch - 1})
    {
        CK(cudaMemcpy(h.data(), dtmp + n * j, sizeof(T) * n, cudaMemcpyDeviceToHost));
        ok = ok && memcmp(h.data(), want.data(), sizeof(T) * n) == 0;
    }
    double ms = time_ms(run, iters);
    double core_ms = time_ms([&]() { GPU_4STEP_NTT(dtmp, dout, d1, d2, dw, dmod, cfg, batch, 1); }, iters);
    double gbs = 2.0 * n * sizeof(T) * batch / (ms * 1e-3) / 1e9;
    printf("{\"lib\": \"%s\", \"case\": \"%s\", \"bits\": 64, \"logn\": %d, \"batch\": %d, \"op\": \"transpose+4step+transpose\", "
           "\"parity_vs_NTT_4STEP_CPU\": %s, \"ms\": %.4f, \"ms_GPU_4STEP_NTT_only\": %.4f, \"ntt_per_s\": %.1f, \"alg_GBps\": %.1f}\n",
           g_label, name, logn, batch, ok ? "true" : "false", ms, core_ms, batch / (ms * 1e-3), gbs);
    fflush(stdout);
    cudaFree(din);
    cudaFree(dout);
    cudaFree(dtmp);
    cudaFree(dw);
}

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        fprintf(stderr, "usage: %s <label> [c2|c2inv|c3|c4|sweep|small|fhe|latency ...]\n", argv[0]);
        return 1;
    }
    g_label = argv[1];
    apply_tune_env();
    CudaDevice();
    CK(cudaSetDevice(0));
    std::vector<std::string> cases;
    for (int i = 2; i < argc; i++)
        cases.push_back(argv[i]);
    if (cases.empty())
        cases = {"c2", "c2inv", "c3", "c4"};
    for (auto& c : cases)
    {
        if (c == "c2")
            merge_case<Data64>("C2", 16, 1024, false, false, 20);
        else if (c == "c2inv")
            merge_case<Data64>("C2-inverse", 16, 1024, true, false, 20);
        else if (c == "c3")
            merge_case<Data32>("C3", 14, 4096, false, true, 20);
        else if (c == "c4")
            fourstep_case("C4", 24, 16, 5);
        else if (c == "sweep")
        {
            for (int l = 12; l <= 24; l += 2)
                merge_case<Data64>("sweep64", l, 1 << (26 - l), false, false, 10);
            for (int l = 12; l <= 24; l += 4)
                merge_case<Data32>("sweep32", l, 1 << (27 - l), false, false, 10);
        }
        else if (c == "small")
        {
            for (int l = 8; l <= 11; l++)
                merge_case<Data64>("small64", l, 32768, false, false, 10);
            for (int l = 8; l <= 13; l++)
                merge_case<Data32>("small32", l, 32768, false, false, 10);
        }
        else if (c == "latency")
        {
            // launch-bound regime: what an RNS-FHE caller issues per ciphertext operation (tens of polynomials per call)
            for (int l = 12; l <= 16; l++)
                for (int b : {8, 32, 128})
                {
                    merge_case<Data64>("latency-rns4", l, b, false, false, 50, ReductionPolynomial::X_N_plus, 4);
                    merge_case<Data64>("latency-rns4", l, b, true, false, 50, ReductionPolynomial::X_N_plus, 4);
                    merge_case<Data64>("latency-single", l, b, false, false, 50, ReductionPolynomial::X_N_plus, 0);
                }
        }
        else if (c == "fhe")
        {
            // what RNS-FHE callers issue: negacyclic ring, RNS overloads, forward and inverse
            for (int l = 12; l <= 17; l++)
            {
                merge_case<Data64>("fhe64-rns4", l, 1 << (26 - l), false, false, 10, ReductionPolynomial::X_N_plus, 4);
                merge_case<Data64>("fhe64-rns4", l, 1 << (26 - l), true, false, 10, ReductionPolynomial::X_N_plus, 4);
            }
            merge_case<Data64>("fhe64-single", 16, 1024, false, false, 10, ReductionPolynomial::X_N_plus, 0);
            merge_case<Data64>("fhe64-single", 16, 1024, true, false, 10, ReductionPolynomial::X_N_plus, 0);
            merge_case<Data32>("fhe32-rns4", 14, 8192, false, false, 10, ReductionPolynomial::X_N_plus, 4);
            merge_case<Data32>("fhe32-rns4", 14, 8192, true, false, 10, ReductionPolynomial::X_N_plus, 4);
            merge_case<Data64>("fhe64-rns4-batch64", 16, 64, false, false, 10, ReductionPolynomial::X_N_plus, 4);
            merge_case<Data64>("fhe64-rns4-batch64", 15, 64, false, false, 10, ReductionPolynomial::X_N_plus, 4);
            merge_case<Data64>("fhe64-rns4-batch64", 13, 64, false, false, 10, ReductionPolynomial::X_N_plus, 4);
        }
        else
            fprintf(stderr, "unknown case %s\n", c.c_str());
    }
    return 0;
}
