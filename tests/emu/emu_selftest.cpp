// TEST INFRASTRUCTURE: self-test of the warp emulator (cuda_host_emu.h) — the collectives must behave like the CUDA
// ones before any kernel result obtained on it means anything.  Exports `emu_selftest()` → 0 when every check passes,
// otherwise the index of the first failing check.
#include "cuda_host_emu.h"

namespace {

std::atomic<int> g_fail{0};
void expect(bool ok, int id) {
    int zero = 0;
    if (!ok) g_fail.compare_exchange_strong(zero, id);
}

void kernel_checks(int* block_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 1. butterfly all-reduce: every lane ends with the warp's sum
    int v = lane + 1;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    expect(v == 32 * 33 / 2, 1);
    // 2. broadcast from an arbitrary lane, float payload
    const float f = __shfl_sync(0xffffffffu, 0.5f * lane, 7);
    expect(f == 3.5f, 2);
    // 3. ballot of a predicate
    const unsigned even = __ballot_sync(0xffffffffu, (lane & 1) == 0);
    expect(even == 0x55555555u, 3);
    // 4. match_any: lanes grouped by value
    const unsigned same = __match_any_sync(0xffffffffu, lane / 8);
    expect(same == (0xffu << (8 * (lane / 8))), 4);
    // 5. collectives among a SUBSET of lanes while the others run ahead to a full-warp collective
    unsigned sub = 0;
    const bool member = lane % 3 == 0;
    const unsigned mask = __ballot_sync(0xffffffffu, member);
    if (member) sub = __match_any_sync(mask, lane % 2);
    const unsigned after = __ballot_sync(0xffffffffu, member && __builtin_popcount(sub) > 0);
    expect(after == mask, 5);
    if (member) {
        unsigned want = 0;
        for (int l = 0; l < 32; ++l) if (l % 3 == 0 && l % 2 == lane % 2) want |= 1u << l;
        expect(sub == want, 6);
    }
    // 6. 64-bit payloads and divergent trip counts in front of a collective
    unsigned long long big = 0x100000000ull * lane + 5;
    for (int i = 0; i < lane % 4; ++i) big += 0;                       // lanes arrive at different times
    const unsigned long long got = __shfl_sync(0xffffffffu, big, 31);
    expect(got == 0x100000000ull * 31 + 5, 7);
    // 7. the block-level pattern the kernels use: every warp publishes, thread 0 of warp 0 consumes after the barrier
    __shared__ int partial[8];
    if (lane == 0) partial[warp] = warp + 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x / 32); ++w) t += partial[w];
        block_sums[blockIdx.x] = t;
    }
}

}  // namespace

extern "C" int emu_selftest() {
    g_fail = 0;
    int sums[3] = {0, 0, 0};
    emu::launch(3, 256, [&] { kernel_checks(sums); });
    for (int b = 0; b < 3; ++b) expect(sums[b] == 36, 8);
    expect(gridDim.x == 3 && blockDim.x == 256, 9);
    return g_fail.load();
}
