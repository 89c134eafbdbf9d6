// Downstream caller compiled against the INSTALLED headers (include/GPUNTT-1.0/gpuntt/...) and linked against
// GPUNTT::ntt: parameter generation and the CPU transform run anywhere; the GPU call needs a B200.
#include <cstdio>
#include <vector>

#include "gpuntt/ntt_merge/ntt.cuh"

using namespace gpuntt;

int main(int argc, char**)
{
    NTTParameters<Data64> params(12, ReductionPolynomial::X_N_minus);
    NTTCPU<Data64> cpu(params);
    std::vector<Data64> x(1 << 12, 1);
    std::vector<Data64> y = cpu.ntt(x);
    std::printf("modulus %llu, NTT(1,...,1)[0] = %llu\n", (unsigned long long) params.modulus.value, (unsigned long long) y[0]);
    if (argc > 1) // only referenced so that the GPU entry point is linked
    {
        ntt_configuration<Data64> cfg = {.n_power = 12, .ntt_type = FORWARD, .ntt_layout = PerPolynomial,
                                         .reduction_poly = ReductionPolynomial::X_N_minus, .zero_padding = false,
                                         .mod_inverse = params.n_inv, .stream = 0};
        GPU_NTT_Inplace<Data64>(nullptr, nullptr, params.modulus, cfg, 0);
    }
    return y[0] == 4096 % params.modulus.value ? 0 : 1;
}
