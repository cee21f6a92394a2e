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
    // 8. shift shuffles: lanes without a source keep their own value
    const int up = __shfl_up_sync(0xffffffffu, lane, 3), down = __shfl_down_sync(0xffffffffu, lane, 5);
    expect(up == (lane >= 3 ? lane - 3 : lane), 10);
    expect(down == (lane + 5 < 32 ? lane + 5 : lane), 11);
    // 9. all warps of the block are concurrent and __syncthreads() is a real barrier: a two-phase exchange through
    //    shared memory between DIFFERENT warps, several rounds, plus shared and global atomics from every thread
    __shared__ int ring[256];
    __shared__ unsigned hits;
    __shared__ float fsum;
    if (threadIdx.x == 0) { hits = 0u; fsum = 0.f; }
    __syncthreads();
    int token = (int)threadIdx.x;
    for (int round = 0; round < 4; ++round) {
        ring[threadIdx.x] = token;
        __syncthreads();
        token = ring[(threadIdx.x + 37) % blockDim.x] + 1;            // a thread of another warp wrote this slot
        __syncthreads();
    }
    expect(token == (int)((threadIdx.x + 4 * 37) % blockDim.x) + 4, 12);
    atomicAdd(&hits, 1u);
    atomicAdd(&fsum, 0.5f);
    atomicMax(reinterpret_cast<unsigned*>(&ring[0]), threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        expect(hits == blockDim.x, 13);
        expect(fsum == 0.5f * blockDim.x, 14);
        expect((unsigned)ring[0] >= blockDim.x - 1, 15);
    }
    // 10. __syncwarp orders a lane-0 write after every lane's read of the same word (the row optimisers' flag)
    __shared__ int flag[8];
    if (lane == 0) flag[warp] = 1;
    __syncwarp();
    const int seen = flag[warp];
    __syncwarp();
    if (lane == 0) flag[warp] = 0;
    expect(seen == 1, 16);
}

}  // namespace

extern "C" int emu_selftest() {
    g_fail = 0;
    int sums[3] = {0, 0, 0};
    emu::launch(dim3(3), dim3(256), [&] { kernel_checks(sums); });
    for (int b = 0; b < 3; ++b) expect(sums[b] == 36, 8);
    expect(gridDim.x == 3 && blockDim.x == 256, 9);
    // 2-dimensional grids: blockIdx.y / gridDim.y
    std::atomic<int> cells{0};
    emu::launch(dim3(2, 3), dim3(32), [&] { if (threadIdx.x == 0) cells.fetch_add(1 + 10 * (int)blockIdx.y + 100 * (int)(gridDim.y == 3)); });
    expect(cells.load() == 6 + 2 * 10 * (0 + 1 + 2) + 600, 17);
    return g_fail.load();
}
