// CPU emulation of the small CUDA subset the foam kernels use -- TEST INFRASTRUCTURE ONLY.
//
// Purpose: run the product's kernel SOURCE (radfoam_b200/csrc/*.cuh, *.cu, unmodified apart from three
// `#ifdef RFB_EMU` spots around inline PTX / dynamic shared memory) on host threads, so the `-m "not gpu"`
// suite can check the kernels' LOGIC (walk, ranked face scan, tape record/replay, warp-aggregated gradient
// cache, host-side cache keys and tape growth) against the oracle in a container without a GPU, and so that
// new kernel variants can be developed against a functional check before they are measured on a B200.
// It says nothing about the compiled SASS (nvcc's contraction, libdevice expf/logf, MUFU.RCP): GPU parity is
// established by tests/test_gpu_*.py only.  It is slow (a fiber per CUDA thread).
//
// The library built from this (tests/emu/libradfoam_b200_emu.so) is never loaded by radfoam_b200/_lib.py and is
// not a fallback: the product has none (tests/test_cabi.py checks both).
//
// Model: the threads of a CTA are fibers on one OS thread; warp collectives (__shfl_sync, __match_any_sync,
// __ballot_sync, __any_sync, __syncwarp) rendezvous the lanes named in the mask (lanes that have exited count as
// arrived); __syncthreads is a barrier over the CTA's live threads; CTAs run in parallel on a few OS threads;
// global atomics are host atomics; "device memory" is host memory and the stream/event API is synchronous.
#pragma once
#define __DEVICE_LAUNCH_PARAMETERS_H__ // keep CUDA's non-thread-local threadIdx/blockIdx declarations out
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <ucontext.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <thread>
#include <vector>

namespace emu {

// ---- execution model: the threads of a CTA are fibers (ucontext) on one OS thread, scheduled round-robin; a
// collective that is not complete yields to the scheduler.  CTAs of a launch run in parallel on a few OS threads
// (function-local __shared__ arrays become thread_local statics, so each OS thread = one CTA at a time).  A full
// scheduler pass without any progress is a deadlock in the kernel's use of the collectives and aborts with a message.
struct Round {
    unsigned arrived = 0, readers = 0;
    bool ready = false;
    uint64_t vals[32];
};

struct Warp {
    std::map<unsigned, Round> rounds; // one rendezvous per participation mask
    unsigned exited = 0;
};

struct Tls {
    uint3 tid, bid;
    dim3 bdim, gdim;
    Warp *warp = nullptr;
    unsigned lane = 0;
    unsigned participants = 0; // lanes that took part in this fiber's last warp collective
};

struct Fiber {
    ucontext_t ctx;
    Tls t;
    bool done = false;
};

struct CtaRun {
    std::vector<Warp> warps;
    std::vector<Fiber> fibers;
    ucontext_t scheduler;
    const std::function<void()> *body = nullptr;
    uint64_t progress = 0;
    unsigned live = 0, barrier_waiting = 0;
    uint64_t barrier_generation = 0;
    float *smem_base = nullptr;
    const void *scheduler_stack = nullptr; // for the sanitizer annotations only
    size_t scheduler_stack_size = 0;
};

struct Counters { // what the kernels would send to L2 / how many warp collectives they ran, since the last reset
    std::atomic<uint64_t> red_v4{0}, red_v2{0}, collectives{0}, votes{0}, votes_skipped{0};
};
extern Counters counters;

extern thread_local CtaRun *run; // the CTA this OS thread is executing
extern thread_local Fiber *cur;  // the fiber (CUDA thread) that is running

inline unsigned popc(unsigned v) { return (unsigned)__builtin_popcount(v); }

// AddressSanitizer has to be told about stack switches
#if defined(__SANITIZE_ADDRESS__)
extern "C" void __sanitizer_start_switch_fiber(void **fake_stack_save, const void *bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void *fake_stack_save, const void **bottom_old, size_t *size_old);
#define RFB_EMU_SWITCH_TO(save, bottom, size) __sanitizer_start_switch_fiber(save, bottom, size)
#define RFB_EMU_SWITCHED(save, bottom_old, size_old) __sanitizer_finish_switch_fiber(save, bottom_old, size_old)
#else
#define RFB_EMU_SWITCH_TO(save, bottom, size) ((void)0)
#define RFB_EMU_SWITCHED(save, bottom_old, size_old) ((void)0)
#endif

inline void yield() {
    Fiber *self = cur;
    void *fake = nullptr;
    RFB_EMU_SWITCH_TO(&fake, run->scheduler_stack, run->scheduler_stack_size);
    swapcontext(&self->ctx, &run->scheduler);
    RFB_EMU_SWITCHED(fake, nullptr, nullptr);
}

// All lanes named in `mask` (minus those that exited) exchange one 64-bit value each.
inline std::array<uint64_t, 32> gather(unsigned mask, uint64_t v) {
    Fiber *self = cur;
    Warp &w = *self->t.warp;
    const unsigned lane = self->t.lane;
    Round &r = w.rounds[mask];
    while (r.ready) // the previous round on this mask is still being read
        yield();
    r.vals[lane] = v;
    r.arrived |= 1u << lane;
    run->progress++;
    if (((r.arrived | w.exited) & mask) == mask) {
        counters.collectives.fetch_add(1, std::memory_order_relaxed);
        r.ready = true;
        r.readers = popc(r.arrived);
    } else {
        while (!r.ready)
            yield();
    }
    std::array<uint64_t, 32> out;
    std::memcpy(out.data(), r.vals, sizeof(r.vals));
    // the lanes whose values are valid: those that ARRIVED -- not "those that have not exited by now" (a fast lane
    // may finish the kernel before a slow one gets to read the round)
    self->t.participants = r.arrived;
    run->progress++;
    if (--r.readers == 0) {
        r.ready = false;
        r.arrived = 0;
    }
    return out;
}

inline void thread_exit() {
    Warp &w = *cur->t.warp;
    w.exited |= 1u << cur->t.lane;
    for (auto &kv : w.rounds) { // rendezvous that were only waiting for this lane
        Round &r = kv.second;
        if (!r.ready && r.arrived && ((r.arrived | w.exited) & kv.first) == kv.first) {
            r.ready = true;
            r.readers = popc(r.arrived);
        }
    }
    CtaRun &c = *run;
    c.live--;
    c.progress++;
    if (c.live && c.barrier_waiting == c.live) { // the barrier was only waiting for this thread
        c.barrier_waiting = 0;
        c.barrier_generation++;
    }
}

inline void syncthreads() {
    CtaRun &c = *run;
    const uint64_t gen = c.barrier_generation;
    c.progress++;
    if (++c.barrier_waiting == c.live) {
        c.barrier_waiting = 0;
        c.barrier_generation++;
    } else {
        while (c.barrier_generation == gen)
            yield();
    }
}

void fiber_entry(); // emu_lib.cpp: runs the kernel body for `cur`, then returns to the scheduler for good

constexpr size_t kFiberStack = 256 * 1024;

inline void run_cta(dim3 grid, dim3 block, uint3 bid, size_t smem_bytes, const std::function<void()> &body) {
    const unsigned nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32;
    struct StackPool { // freed when the worker thread of a launch ends; the calling thread keeps and reuses its own
        std::vector<char *> v;
        ~StackPool() {
            for (char *p : v)
                std::free(p);
        }
    };
    static thread_local StackPool pool;
    static thread_local std::vector<float> smem;
    std::vector<char *> &stacks = pool.v;
    while (stacks.size() < nthreads)
        stacks.push_back(static_cast<char *>(std::aligned_alloc(64, kFiberStack)));
    smem.assign(smem_bytes / sizeof(float) + 8, 0.0f);
    CtaRun c;
    c.warps.resize(nwarps);
    if (nthreads % 32)
        c.warps.back().exited = ~0u << (nthreads % 32);
    c.fibers.resize(nthreads);
    c.body = &body;
    c.live = nthreads;
    c.smem_base = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
    run = &c;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = c.fibers[t];
        f.t.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        f.t.bid = bid;
        f.t.bdim = block;
        f.t.gdim = grid;
        f.t.warp = &c.warps[t / 32];
        f.t.lane = t % 32;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks[t];
        f.ctx.uc_stack.ss_size = kFiberStack;
        f.ctx.uc_link = &c.scheduler;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    // RFB_EMU_SHUFFLE=<seed>: resume the fibers in a different pseudo-random order on every pass, so that a missing
    // __syncwarp / __syncthreads between a shared-memory write and another lane's read shows up as a wrong result
    // (the default lane order 0..31 hides writer-before-reader dependencies)
    static const char *shuffle_env = std::getenv("RFB_EMU_SHUFFLE");
    uint64_t rng = shuffle_env ? (uint64_t)std::atoll(shuffle_env) * 0x9E3779B97F4A7C15ull + bid.x + 1 : 0;
    std::vector<unsigned> order(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
        order[t] = t;
    unsigned remaining = nthreads;
    while (remaining) {
        const uint64_t before = c.progress;
        if (shuffle_env)
            for (unsigned t = nthreads - 1; t > 0; --t) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[t], order[(rng >> 33) % (t + 1)]);
            }
        for (unsigned oi = 0; oi < nthreads; ++oi) {
            const unsigned t = order[oi];
            Fiber &f = c.fibers[t];
            if (f.done)
                continue;
            cur = &f;
            void *fake = nullptr;
            RFB_EMU_SWITCH_TO(&fake, stacks[t], kFiberStack);
            swapcontext(&c.scheduler, &f.ctx);
            RFB_EMU_SWITCHED(fake, &c.scheduler_stack, &c.scheduler_stack_size);
            if (f.done)
                remaining--;
        }
        if (remaining && c.progress == before) {
            std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %u threads wait on collectives that can "
                                 "never complete\n", bid.x, bid.y, bid.z, remaining);
            std::abort();
        }
    }
    run = nullptr;
    cur = nullptr;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    std::atomic<uint64_t> next{0};
    auto worker = [&] {
        for (uint64_t b; (b = next.fetch_add(1)) < nblocks;)
            run_cta(grid, block, make_uint3((unsigned)(b % grid.x), (unsigned)(b / grid.x % grid.y),
                                            (unsigned)(b / ((uint64_t)grid.x * grid.y))),
                    smem_bytes, body);
    };
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned nworkers = (unsigned)std::min<uint64_t>(nblocks, hw ? hw : 4);
    std::vector<std::thread> threads;
    for (unsigned i = 1; i < nworkers; ++i)
        threads.emplace_back(worker);
    worker();
    for (auto &t : threads)
        t.join();
}

template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle of a wide type");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}

// ---- the synchronous "runtime"
inline cudaError_t malloc_(void **p, size_t n) {
    *p = std::aligned_alloc(256, (n + 255) & ~size_t(255));
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t free_(void *p) {
    std::free(p);
    return cudaSuccess;
}

} // namespace emu

// ---------------------------------------------------------------------------------- language surface
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static thread_local /* one CTA at a time per OS thread */
#define threadIdx (emu::cur->t.tid)
#define blockIdx (emu::cur->t.bid)
#define blockDim (emu::cur->t.bdim)
#define gridDim (emu::cur->t.gdim)
#define RFB_UNPAREN(...) __VA_ARGS__
#define RFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=] { (RFB_UNPAREN kernel)(__VA_ARGS__); })

inline float *rfb_emu_dynamic_smem() { return emu::run->smem_base; }

using std::isfinite;

// ---- arithmetic with a stated rounding (the host FPU is IEEE round-to-nearest; build with -ffp-contract=off)
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __double2float_rn(double a) { return (float)a; }
inline float __uint2float_rn(unsigned a) { return (float)a; }
inline float __int_as_float(int a) { return emu::from_bits<float>((uint32_t)a); }
inline float __uint_as_float(unsigned a) { return emu::from_bits<float>(a); }
inline unsigned __float_as_uint(float a) { return (unsigned)emu::to_bits(a); }
inline int __float_as_int(float a) { return (int)emu::to_bits(a); }
inline int __popc(unsigned a) { return __builtin_popcount(a); }
inline int __ffs(unsigned a) { return __builtin_ffs((int)a); }
inline int __ffsll(long long a) { return __builtin_ffsll(a); }
inline float rfb_emu_rcp_approx(float x) { // rcp.approx.ftz: denormal inputs and results flush to zero
    if (std::fabs(x) < 1.17549435e-38f)
        x = std::copysign(0.0f, x);
    float r = 1.0f / x;
    return std::fabs(r) < 1.17549435e-38f ? std::copysign(0.0f, r) : r;
}
template <typename T>
inline T emu_min(T a, T b) { return b < a ? b : a; }
inline unsigned min(unsigned a, unsigned b) { return emu_min(a, b); }
inline unsigned max(unsigned a, unsigned b) { return a < b ? b : a; }
inline int min(int a, int b) { return emu_min(a, b); }
inline int max(int a, int b) { return a < b ? b : a; }

// ---- loads / stores with cache hints
template <typename T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcs(const T *p) { return *p; }
template <typename T>
inline void __stcs(T *p, T v) { *p = v; }

// ---- global atomics
inline float atomicAdd(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        uint32_t want = (uint32_t)emu::to_bits(emu::from_bits<float>(old) + v);
        if (__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
            return emu::from_bits<float>(old);
    }
}
inline __half atomicAdd(__half *p, __half v) {
    uint16_t *u = reinterpret_cast<uint16_t *>(p);
    uint16_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        __half sum = __float2half_rn(__half2float(emu::from_bits<__half>(old)) + __half2float(v));
        uint16_t want = (uint16_t)emu::to_bits(sum);
        if (__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
            return emu::from_bits<__half>(old);
    }
}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline unsigned atomicCAS(unsigned *p, unsigned expected, unsigned desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expected; // the old value either way
}
inline unsigned atomicMax(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
// NVSwitch multicast stand-in: one registered group = `world` equally sized buffers reachable through a fake base
// address; ld_reduce sums the copies in rank order (the switch's order is unspecified), st writes every copy.
namespace emu {
struct MulticastGroup {
    char *fake = nullptr;
    size_t bytes = 0;
    std::vector<char *> copies;
};
inline MulticastGroup &multicast() {
    static MulticastGroup g;
    return g;
}
inline size_t multicast_offset(const void *p) {
    auto &g = multicast();
    const char *c = reinterpret_cast<const char *>(p);
    if (!g.fake || c < g.fake || c >= g.fake + g.bytes)
        std::abort();
    return (size_t)(c - g.fake);
}
} // namespace emu
inline float4 rfb_emu_multimem_ld_reduce_v4(const float *p) {
    const size_t off = emu::multicast_offset(p);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    bool first = true;
    for (char *c : emu::multicast().copies) {
        const float4 v = *reinterpret_cast<const float4 *>(c + off);
        if (first) {
            s = v;
            first = false;
        } else {
            s.x += v.x;
            s.y += v.y;
            s.z += v.z;
            s.w += v.w;
        }
    }
    return s;
}
inline void rfb_emu_multimem_st(void *p, const void *value, size_t n) {
    const size_t off = emu::multicast_offset(p);
    for (char *c : emu::multicast().copies)
        std::memcpy(c + off, value, n);
}
inline void rfb_emu_red_add_v4(float *p, float a, float b, float c, float d) {
    emu::counters.red_v4.fetch_add(1, std::memory_order_relaxed);
    atomicAdd(p, a);
    atomicAdd(p + 1, b);
    atomicAdd(p + 2, c);
    atomicAdd(p + 3, d);
}

inline void rfb_emu_red_add_v2(float *p, float a, float b) {
    emu::counters.red_v2.fetch_add(1, std::memory_order_relaxed);
    atomicAdd(p, a);
    atomicAdd(p + 1, b);
}

// ---- warp collectives
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::gather(mask, 0); }
inline void __syncthreads() { emu::syncthreads(); }
inline unsigned __ballot_sync(unsigned mask, int pred) {
    auto v = emu::gather(mask, pred ? 1 : 0);
    unsigned out = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((emu::cur->t.participants >> l & 1u) && v[l])
            out |= 1u << l;
    return out;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline unsigned __reduce_max_sync(unsigned mask, unsigned value) {
    auto v = emu::gather(mask, value);
    unsigned out = 0;
    for (unsigned l = 0; l < 32; ++l)
        if (emu::cur->t.participants >> l & 1u)
            out = std::max(out, (unsigned)v[l]);
    return out;
}
inline unsigned __match_any_sync(unsigned mask, unsigned value) {
    auto v = emu::gather(mask, value);
    unsigned out = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((emu::cur->t.participants >> l & 1u) && (unsigned)v[l] == value)
            out |= 1u << l;
    return out;
}
template <typename T>
inline T __shfl_sync(unsigned mask, T value, int src, int width = 32) {
    auto v = emu::gather(mask, emu::to_bits(value));
    const unsigned lane = emu::cur->t.lane;
    const unsigned from = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    return (mask >> from & 1u) ? emu::from_bits<T>(v[from]) : value;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T value, int lane_mask, int width = 32) {
    auto v = emu::gather(mask, emu::to_bits(value));
    const unsigned from = emu::cur->t.lane ^ (unsigned)lane_mask;
    (void)width;
    return (mask >> from & 1u) ? emu::from_bits<T>(v[from]) : value;
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T value, unsigned delta, int width = 32) {
    auto v = emu::gather(mask, emu::to_bits(value));
    const unsigned lane = emu::cur->t.lane, from = lane + delta;
    if ((lane & (unsigned)(width - 1)) + delta >= (unsigned)width || !(mask >> from & 1u))
        return value;
    return emu::from_bits<T>(v[from]);
}

template <typename T>
inline T __shfl_up_sync(unsigned mask, T value, unsigned delta, int width = 32) {
    auto v = emu::gather(mask, emu::to_bits(value));
    const unsigned lane = emu::cur->t.lane;
    if ((lane & (unsigned)(width - 1)) < delta || !(mask >> (lane - delta) & 1u))
        return value;
    return emu::from_bits<T>(v[lane - delta]);
}

// ---- runtime API (synchronous stand-ins; every handle is a dummy)
#define cudaMalloc(p, n) emu::malloc_(reinterpret_cast<void **>(p), (n))
#define cudaFree(p) emu::free_(p)
#define cudaMallocHost(p, n) emu::malloc_(reinterpret_cast<void **>(p), (n))
#define cudaFreeHost(p) emu::free_(p)
#define cudaMallocAsync(p, n, s) emu::malloc_(reinterpret_cast<void **>(p), (n))
#define cudaFreeAsync(p, s) emu::free_(p)
#define cudaMemsetAsync(p, v, n, s) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, kind, st) (std::memcpy((d), (s), (n)), cudaSuccess)
#define cudaStreamSynchronize(s) (cudaSuccess)
#define cudaStreamWaitEvent(s, e, f) (cudaSuccess)
#define cudaStreamIsCapturing(s, st) (*(st) = cudaStreamCaptureStatusNone, cudaSuccess)
#define cudaMemGetInfo(f, t) (*(f) = (size_t)1 << 40, *(t) = (size_t)1 << 40, cudaSuccess)
#define cudaGetLastError() (cudaSuccess)
#define cudaGetErrorString(e) ("emulated CUDA runtime error")
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#define cudaEventCreate(e) (*(e) = nullptr, cudaSuccess)
#define cudaEventCreateWithFlags(e, f) (*(e) = nullptr, cudaSuccess)
#define cudaEventRecord(e, s) (cudaSuccess)
#define cudaEventQuery(e) (cudaSuccess)
#define cudaEventSynchronize(e) (cudaSuccess)
#define cudaEventDestroy(e) (cudaSuccess)
#define cudaEventElapsedTime(ms, a, b) (*(ms) = 0.0f, cudaSuccess)
#define cudaFuncSetAttribute(f, a, v) (cudaSuccess)
#define cudaDeviceGetDefaultMemPool(p, d) (*(p) = nullptr, cudaSuccess)
#define cudaMemPoolSetAttribute(p, a, v) (cudaSuccess)
