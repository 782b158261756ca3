// emu_shim.h — TEST INFRASTRUCTURE: run a CUDA kernel's SOURCE on the host, one std::thread per CUDA thread of a block, blocks one
// after the other.  Enough of the execution model for kernels that use threadIdx/blockIdx, __shared__ arrays, __syncthreads,
// shared/global integer atomics, __ffs and warp / quad shuffles: the logic of a kernel (indexing, barriers, cursors,
// fallbacks) can be exercised without a GPU and compared with another kernel run the same way.  It says nothing about
// performance, memory-model subtleties or fused multiply-add placement.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <algorithm>
#include <thread>
#include <vector>

#include <cuda_runtime.h>      // host-side declarations only (vector types, qualifiers as ignored attributes)

// qualifiers: CUDA's host_defines.h turns them into attributes g++ ignores; __shared__ must become block-wide storage
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;
static std::barrier<> *emu_barrier = nullptr;

static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
using std::max;
using std::min;

static inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
// warp shuffles: the participants are the lanes named by `mask` — the full warp (0xffffffff) or one aligned quad (0xf << 4q).  Each
// group has its own reusable barrier, so groups may call a different number of shuffles (quads walk lists of different lengths).
static std::barrier<> *emu_warp_bar[32], *emu_quad_bar[256];
static double emu_slot[1024];
template <typename T>
static inline T emu_shfl(unsigned mask, T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    std::barrier<> *bar;
    if (mask == 0xffffffffu) bar = emu_warp_bar[tid >> 5];
    else if (mask == (0xfu << (lane & ~3u))) bar = emu_quad_bar[tid >> 2];
    else { std::fprintf(stderr, "emu_shfl: unsupported mask %08x on lane %u\n", mask, lane); std::abort(); }
    std::memcpy(&emu_slot[tid], &v, sizeof(T));
    bar->arrive_and_wait();
    T r;
    std::memcpy(&r, &emu_slot[(tid & ~31u) + (src_lane & 31u)], sizeof(T));
    bar->arrive_and_wait();
    return r;
}
static inline double __shfl_xor_sync(unsigned m, double v, int x) { return emu_shfl(m, v, (threadIdx.x & 31u) ^ (unsigned) x); }
static inline int __shfl_xor_sync(unsigned m, int v, int x) { return emu_shfl(m, v, (threadIdx.x & 31u) ^ (unsigned) x); }
static inline double __shfl_sync(unsigned m, double v, int src) { return emu_shfl(m, v, (unsigned) src); }
static inline int __shfl_sync(unsigned m, int v, int src) { return emu_shfl(m, v, (unsigned) src); }

// launch: kernel(args...) for every thread of every block
template <typename K, typename... A>
static void emu_launch(unsigned blocks, unsigned threads, K kernel, A... args) {
    blockDim.x = threads; gridDim.x = blocks;
    for (unsigned b = 0; b < blocks; b++) {
        std::barrier<> bar((std::ptrdiff_t) threads);
        emu_barrier = &bar;
        std::vector<std::unique_ptr<std::barrier<>>> groups;       // (blocks are multiples of 32 threads)
        for (unsigned w = 0; w < threads / 32; w++) { groups.emplace_back(new std::barrier<>(32)); emu_warp_bar[w] = groups.back().get(); }
        for (unsigned q = 0; q < threads / 4; q++) { groups.emplace_back(new std::barrier<>(4)); emu_quad_bar[q] = groups.back().get(); }
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; t++)
            pool.emplace_back([=]() { threadIdx.x = t; blockIdx.x = b; kernel(args...); });
        for (auto &th : pool) th.join();
    }
    emu_barrier = nullptr;
}
