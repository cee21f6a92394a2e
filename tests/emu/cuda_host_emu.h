// cuda_host_emu.h — TEST INFRASTRUCTURE: a warp-level CUDA execution model on the CPU, so that the product's kernel
// sources (openea_b200/csrc/*.cu, compiled unchanged with -DOEA_HOST_EMU by g++) can be checked against the oracle
// where no GPU exists.  Not a fallback: nothing under openea_b200/ ever loads a library built with it.
//
// Model: one OS thread per lane; the 32 lanes of a warp run concurrently and meet at every warp collective
// (__shfl*_sync, __ballot_sync, __match_any_sync) on a per-mask rendezvous, exactly where real lanes exchange
// registers.  The warps of a block, and the blocks of a grid, run one after another (warp 0 of a block last, so a
// block-level reduction that thread 0 finishes after __syncthreads() sees every warp's partial result);
// __syncthreads() itself is a no-op, which is valid only for kernels whose barriers separate "every warp publishes"
// from "thread 0 consumes" — the only pattern the emulated kernels use (LossAcc::flush).
// "Device" pointers are host pointers; red/atomic adds are plain adds (lanes own disjoint addresses, warps are serial).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

// stream-ordered runtime calls the entry points make between launches: "device" memory is host memory and launches are
// synchronous, so they are plain memset / memcpy
#define cudaMemsetAsync(P, V, N, ST) (memset((P), (V), (N)), cudaSuccess)
#define cudaMemcpyAsync(D, S, N, KIND, ST) (memcpy((D), (S), (N)), cudaSuccess)
#define cudaStreamSynchronize(ST) (cudaSuccess)
#define cudaLaunchCooperativeKernel(...) (cudaErrorNotSupported)        /* grid barriers are not emulated */

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local EmuDim3 threadIdx, blockIdx;
inline EmuDim3 gridDim, blockDim;

#define _COOPERATIVE_GROUPS_H_          /* keep <cooperative_groups.h> out: it needs nvcc */
namespace cooperative_groups {
struct grid_group { void sync() const {} };    // serial blocks cannot honour a grid barrier: kernels using it are not run
inline grid_group this_grid() { return {}; }
}  // namespace cooperative_groups

namespace emu {

struct Rendezvous {
    std::atomic<int> arrived{0};
    std::atomic<unsigned> generation{0};
};
inline std::mutex g_map_mutex;
inline std::map<unsigned, Rendezvous> g_rendezvous;   // one per participation mask
inline uint64_t g_slot[32];
inline thread_local int t_lane = 0;

inline Rendezvous& rendezvous_of(unsigned mask) {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    return g_rendezvous[mask];
}
// all lanes named in `mask` meet here
inline void sync(unsigned mask) {
    Rendezvous& r = rendezvous_of(mask);
    const int expected = __builtin_popcount(mask);
    const unsigned gen = r.generation.load();
    if (r.arrived.fetch_add(1) + 1 == expected) {
        r.arrived.store(0);
        r.generation.fetch_add(1);
    } else {
        while (r.generation.load() == gen) std::this_thread::yield();
    }
}
template <typename T>
inline void publish(T v) {
    static_assert(sizeof(T) <= 8, "exchange slot is 64 bits");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_slot[t_lane] = bits;
}
template <typename T>
inline T peek(int lane) {
    T v;
    memcpy(&v, &g_slot[lane], sizeof(T));
    return v;
}

// Lane scheduling of the next launches.  Concurrent (default): the 32 lanes of a warp are free-running threads that meet
// only at collectives — right for kernels that synchronise explicitly, but a kernel that relies on the hardware keeping a
// converged warp in step between two plain memory accesses (all lanes read a row's `touched` flag, lane 0 clears it after
// its columns: the row optimisers) sees a race real lanes never see.  Serial: the lanes of a warp run one after another,
// lane 31 first and lane 0 last — valid ONLY for kernels without warp collectives, and it gives exactly that order.
inline std::atomic<bool> g_serial_lanes{false};

// run `body` as a grid of blocks of `threads` threads (a multiple of 32)
template <typename Body>
inline void launch(int grid, int threads, Body body) {
    gridDim.x = (unsigned)grid;
    blockDim.x = (unsigned)threads;
    const int warps = threads / 32;
    if (g_serial_lanes.load()) {
        for (int b = 0; b < grid; ++b)
            for (int w = warps - 1; w >= 0; --w)
                for (int l = 31; l >= 0; --l) {
                    t_lane = l;
                    threadIdx.x = (unsigned)(w * 32 + l);
                    blockIdx.x = (unsigned)b;
                    body();
                }
        return;
    }
    for (int b = 0; b < grid; ++b) {
        for (int w = warps - 1; w >= 0; --w) {
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; ++l) {
                lanes.emplace_back([=] {
                    t_lane = l;
                    threadIdx.x = (unsigned)(w * 32 + l);
                    blockIdx.x = (unsigned)b;
                    body();
                });
            }
            for (auto& t : lanes) t.join();
        }
    }
}

}  // namespace emu

// ---- warp collectives ------------------------------------------------------------------------------------------------
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    emu::publish(v);
    emu::sync(mask);
    const T r = emu::peek<T>((emu::t_lane ^ lane_mask) & 31);
    emu::sync(mask);
    return r;
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src) {
    emu::publish(v);
    emu::sync(mask);
    const T r = emu::peek<T>(src & 31);
    emu::sync(mask);
    return r;
}
inline unsigned __ballot_sync(unsigned mask, bool pred) {
    emu::publish<unsigned>(pred ? 1u : 0u);
    emu::sync(mask);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && emu::peek<unsigned>(l)) out |= 1u << l;
    emu::sync(mask);
    return out;
}
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    emu::publish(v);
    emu::sync(mask);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && emu::peek<T>(l) == v) out |= 1u << l;
    emu::sync(mask);
    return out;
}
inline void __syncthreads() {}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::sync(mask); }     // lanes are threads: a real rendezvous

// ---- loads, atomics, intrinsics ----------------------------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline double atomicAdd(double* p, double v) { const double old = *p; *p = old + v; return old; }
inline float atomicAdd(float* p, float v) {          // lanes of one warp may hit the same word (scatter-adds): really atomic
    std::atomic_ref<float> a(*p);
    float old = a.load(std::memory_order_relaxed);
    while (!a.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
    return old;
}
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long expected, unsigned long long desired) {
    std::atomic_ref<unsigned long long> a(*p);
    a.compare_exchange_strong(expected, desired);
    return expected;                                  // the value found at *p, as CUDA's atomicCAS returns
}
using std::min;
using std::max;
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float emu_expf(float x) { return expf(x); }
#define __expf emu_expf
