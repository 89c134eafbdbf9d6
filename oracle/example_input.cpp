// oracle/example_input.cpp -- the input generator of the reference's example drivers
// (example/ntt_merge/test_merge_ntt.cu:70-85: std::mt19937 gen(0);
// std::uniform_int_distribution<uint64_t> dis(0, p-1); filled polynomial-major).
// TEST INFRASTRUCTURE ONLY.  Uses <random> directly so that the stream is the
// libstdc++ one the reference examples produce on this image.
#include <cstdint>
#include <random>

extern "C" void ora_example_input(uint32_t seed, uint64_t p, uint64_t count, uint64_t* out)
{
    std::mt19937 gen(seed);
    std::uniform_int_distribution<std::uint64_t> dis(0, p - 1);
    for (uint64_t i = 0; i < count; i++) out[i] = dis(gen);
}
