// examples/gpu_4step_examples.cu -- self-check of the 4-step entry points through the public C++ API.
//
//   gpu_4step_examples [LOGN 12..24] [BATCH]
//
// Two ways to the same result (NTT_4STEP_CPU::ntt, the reference's example/ntt_4step/test_4step_ntt.cu:147-166):
//   reference contract  GPU_Transpose -> GPU_4STEP_NTT -> GPU_Transpose, exactly the reference's call sequence;
//   fused contract      GPU_4STEP_NTT_Fused: natural order in, final order out, in place, no transposes by the caller.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gpuntt/ntt_4step/ntt_4step.cuh"
#include "gpuntt_b200.h"

using namespace gpuntt;

int main(int argc, char** argv)
{
    CudaDevice();
    const int logn = argc > 1 ? std::atoi(argv[1]) : 12;
    const int batch = argc > 2 ? std::atoi(argv[2]) : 2;
    typedef Data64 T;
    NTTParameters4Step<T> params(logn, ReductionPolynomial::X_N_minus);
    NTT_4STEP_CPU<T> cpu(params);
    const size_t n = (size_t) 1 << logn;
    std::vector<uint64_t> host(n * batch);
    gpuntt_b200_example_input(0, params.modulus.value, host.size(), host.data());

    std::vector<std::vector<T>> want;
    for (int b = 0; b < batch; b++)
    {
        std::vector<T> one(host.begin() + b * n, host.begin() + (b + 1) * n);
        want.push_back(cpu.ntt(one));
    }

    std::vector<Root<T>> t1 = params.gpu_root_of_unity_table_generator(params.n1_based_root_of_unity_table);
    std::vector<Root<T>> t2 = params.gpu_root_of_unity_table_generator(params.n2_based_root_of_unity_table);
    T *d_a = nullptr, *d_b = nullptr;
    Root<T>*d_t1 = nullptr, *d_t2 = nullptr, *d_w = nullptr;
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_a, host.size() * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_b, host.size() * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_t1, t1.size() * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_t2, t2.size() * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_w, n * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_t1, t1.data(), t1.size() * sizeof(T), cudaMemcpyHostToDevice));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_t2, t2.data(), t2.size() * sizeof(T), cudaMemcpyHostToDevice));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_w, params.W_root_of_unity_table.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    ntt4step_configuration<T> cfg = {.n_power = logn, .ntt_type = FORWARD, .mod_inverse = params.n_inv_gpu, .stream = 0};

    auto check = [&](T* dev, const char* what) -> bool
    {
        std::vector<T> got(host.size());
        GPUNTT_CUDA_CHECK(cudaMemcpy(got.data(), dev, got.size() * sizeof(T), cudaMemcpyDeviceToHost));
        for (int b = 0; b < batch; b++)
            for (size_t i = 0; i < n; i++)
                if (got[b * n + i] != want[b][i])
                {
                    std::printf("%s: mismatch at polynomial %d, index %zu\n", what, b, i);
                    return false;
                }
        std::printf("%-20s logN=%d (%d x %d) batch=%d: All Correct.\n", what, logn, params.n1, params.n2, batch);
        return true;
    };

    bool ok = true;
    // the reference's sequence
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_a, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    GPU_Transpose(d_a, d_b, params.n1, params.n2, logn, batch);
    GPU_4STEP_NTT(d_b, d_a, d_t1, d_t2, d_w, params.modulus, cfg, batch);
    GPU_Transpose(d_a, d_b, params.n1, params.n2, logn, batch);
    ok &= check(d_b, "reference contract");
    // fused, in place
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_a, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    GPU_4STEP_NTT_Fused(d_a, d_a, d_t1, d_t2, d_w, params.modulus, cfg, batch);
    ok &= check(d_a, "fused contract");
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_t1);
    cudaFree(d_t2);
    cudaFree(d_w);
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}
