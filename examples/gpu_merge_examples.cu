// examples/gpu_merge_examples.cu -- self-check of the Merge-NTT entry points through the public C++ API.
//
//   gpu_merge_examples [LOGN] [BATCH]
//
// What the reference's gpu_merge_ntt_examples / gpu_merge_intt_examples verify (example/ntt_merge/test_merge_ntt.cu:
// 143-180, test_merge_intt.cu:166-200), written against this repository's headers: Data64 and Data32, both ring types,
// GPU_NTT_Inplace == NTTCPU::ntt word for word on the examples' seed-0 input stream, GPU_INTT_Inplace restores the input.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gpuntt/ntt_merge/ntt.cuh"
#include "gpuntt_b200.h"

using namespace gpuntt;

template <typename T> static bool run(int logn, int batch, ReductionPolynomial poly, const char* name)
{
    NTTParameters<T> params(logn, poly);
    NTTCPU<T> cpu(params);
    const size_t n = (size_t) 1 << logn;
    std::vector<uint64_t> stream(n * batch);
    gpuntt_b200_example_input(0, (uint64_t) params.modulus.value, stream.size(), stream.data());
    std::vector<T> host(stream.begin(), stream.end());

    T* d_data = nullptr;
    Root<T>*d_fwd = nullptr, *d_inv = nullptr;
    std::vector<Root<T>> fwd = params.gpu_root_of_unity_table_generator(params.forward_root_of_unity_table);
    std::vector<Root<T>> inv = params.gpu_root_of_unity_table_generator(params.inverse_root_of_unity_table);
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_data, host.size() * sizeof(T)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_fwd, fwd.size() * sizeof(Root<T>)));
    GPUNTT_CUDA_CHECK(cudaMalloc(&d_inv, inv.size() * sizeof(Root<T>)));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_data, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_fwd, fwd.data(), fwd.size() * sizeof(Root<T>), cudaMemcpyHostToDevice));
    GPUNTT_CUDA_CHECK(cudaMemcpy(d_inv, inv.data(), inv.size() * sizeof(Root<T>), cudaMemcpyHostToDevice));

    ntt_configuration<T> cfg = {.n_power = logn, .ntt_type = FORWARD, .ntt_layout = PerPolynomial, .reduction_poly = poly,
                                .zero_padding = false, .mod_inverse = params.n_inv, .stream = 0};
    GPU_NTT_Inplace(d_data, d_fwd, params.modulus, cfg, batch);
    std::vector<T> got(host.size());
    GPUNTT_CUDA_CHECK(cudaMemcpy(got.data(), d_data, got.size() * sizeof(T), cudaMemcpyDeviceToHost));
    bool ok = true;
    for (int b = 0; b < batch && ok; b++)
    {
        std::vector<T> one(host.begin() + b * n, host.begin() + (b + 1) * n);
        std::vector<T> want = cpu.ntt(one);
        for (size_t i = 0; i < n; i++)
            if (want[i] != got[b * n + i])
            {
                std::printf("%s: forward mismatch at polynomial %d, index %zu\n", name, b, i);
                ok = false;
                break;
            }
    }
    cfg.ntt_type = INVERSE;
    GPU_INTT_Inplace(d_data, d_inv, params.modulus, cfg, batch);
    GPUNTT_CUDA_CHECK(cudaMemcpy(got.data(), d_data, got.size() * sizeof(T), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < got.size() && ok; i++)
        if (got[i] != host[i])
        {
            std::printf("%s: inverse mismatch at word %zu\n", name, i);
            ok = false;
        }
    cudaFree(d_data);
    cudaFree(d_fwd);
    cudaFree(d_inv);
    std::printf("%-28s logN=%d batch=%d: %s\n", name, logn, batch, ok ? "All Correct." : "FAILED");
    return ok;
}

int main(int argc, char** argv)
{
    CudaDevice();
    const int logn = argc > 1 ? std::atoi(argv[1]) : 12;
    const int batch = argc > 2 ? std::atoi(argv[2]) : 4;
    bool ok = true;
    ok &= run<Data64>(logn, batch, ReductionPolynomial::X_N_minus, "Data64 X^N-1");
    ok &= run<Data64>(logn, batch, ReductionPolynomial::X_N_plus, "Data64 X^N+1");
    ok &= run<Data32>(logn, batch, ReductionPolynomial::X_N_minus, "Data32 X^N-1");
    ok &= run<Data32>(logn, batch, ReductionPolynomial::X_N_plus, "Data32 X^N+1");
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}
