// CPU execution of the CUDA kernel SOURCES for logic tests (tests only -- never part of the product).
//
// A CTA's threads run as fibers (ucontext) on ONE OS thread.  A fiber runs until it reaches a synchronisation
// point (__syncthreads, a warp collective) or returns; the scheduler resumes the fibers in a pseudo-random order,
// so a missing barrier shows up as a wrong result instead of passing by luck.  Dynamic shared memory is filled
// with a NaN pattern before every CTA (real shared memory is not zeroed either).  What this cannot show:
// performance, real memory-model races inside a barrier interval, anything about SASS.
#pragma once
#include <cuda_runtime.h>   // float2, dim3, ... (host-side declarations only)
#include <cufft.h>
#include <ucontext.h>
#include <algorithm>
#include <cstdarg>
#include <string>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __launch_bounds__(...)

namespace simt {

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 128 * 1024;
constexpr size_t kDynSmem = 232448;

struct Fiber {
    ucontext_t uc;
    char *stack = nullptr;
    bool done = true;
};

inline ucontext_t sched_uc;
inline Fiber fibers[kMaxThreads];
inline int cur = 0, nthreads = 0;
inline std::function<void()> body;
alignas(128) inline unsigned char dyn_smem_buf[kDynSmem];
inline unsigned char *dyn_smem = dyn_smem_buf;
inline int bar_count = 0, bar_gen = 0;
inline int wbar_count[kMaxThreads / 32], wbar_gen[kMaxThreads / 32];
inline unsigned long long wbuf[kMaxThreads / 32][32];
inline unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
inline long long switches = 0;
inline long long progress = 0;     // barrier releases + fiber exits: lets the scheduler see that a burst is stuck

inline void yield() {
    ++switches;
    swapcontext(&fibers[cur].uc, &sched_uc);
}

inline void cta_barrier() {
    const int g = bar_gen;
    if (++bar_count == nthreads) { bar_count = 0; ++bar_gen; ++progress; }
    else while (bar_gen == g) yield();
}

inline int warp_size_of(int w) { return std::min(32, nthreads - 32 * w); }

inline void warp_barrier() {
    const int w = cur >> 5;
    const int g = wbar_gen[w];
    if (++wbar_count[w] == warp_size_of(w)) { wbar_count[w] = 0; ++wbar_gen[w]; ++progress; }
    else while (wbar_gen[w] == g) yield();
}

// every lane publishes v; returns the value of lane `src` (own value if src is outside the warp)
template <typename T>
inline T exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "exchange of up to 8 bytes");
    const int w = cur >> 5, l = cur & 31;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    wbuf[w][l] = raw;
    warp_barrier();
    T r = v;
    if (src >= 0 && src < warp_size_of(w)) std::memcpy(&r, &wbuf[w][src], sizeof(T));
    warp_barrier();
    return r;
}

inline void trampoline() {
    body();
    fibers[cur].done = true;
    swapcontext(&fibers[cur].uc, &sched_uc);
}

}  // namespace simt

// ---- built-in variables -------------------------------------------------------------------------
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace simt {

// runs kernel body `fn` for every CTA of the grid (1-D grids and blocks)
inline void launch(unsigned grid, unsigned block, std::function<void()> fn) {
    if (block > (unsigned)kMaxThreads) { std::fprintf(stderr, "simt: block too large\n"); std::abort(); }
    body = std::move(fn);
    gridDim = dim3(grid, 1, 1);
    blockDim = dim3(block, 1, 1);
    nthreads = (int)block;
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx = {b, 0, 0};
        std::memset(dyn_smem_buf, 0xff, sizeof(dyn_smem_buf));      // NaN / huge-int pattern
        bar_count = 0;
        for (int w = 0; w < kMaxThreads / 32; ++w) wbar_count[w] = 0;
        for (int t = 0; t < nthreads; ++t) {
            Fiber &f = fibers[t];
            if (!f.stack) f.stack = (char *)std::malloc(kStack);
            getcontext(&f.uc);
            f.uc.uc_stack.ss_sp = f.stack;
            f.uc.uc_stack.ss_size = kStack;
            f.uc.uc_link = nullptr;
            makecontext(&f.uc, (void (*)())trampoline, 0);
            f.done = false;
        }
        int alive = nthreads;
        const int nwarps = (nthreads + 31) / 32;
        int burst_warp = -1;
        while (alive > 0) {
            // One pass resumes the fibers of some warps; each fiber runs to its next synchronisation point.  Two
            // modes, chosen at random: a random SUBSET of the warps (they drift apart step by step), or a BURST in
            // which one warp alone runs on, pass after pass, until it is stuck behind a CTA barrier or done (so it
            // gets whole phases ahead of the others).  Only a CTA barrier holds warps together -- which is what makes
            // a MISSING barrier visible as a wrong result.  Inside a warp the lane order varies from pass to pass.
            rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
            unsigned long long pick = rng_state >> 16;
            if (burst_warp < 0 && ((rng_state >> 12) & 7) == 0) burst_warp = (int)((rng_state >> 40) % (unsigned)nwarps);
            if (burst_warp >= 0) pick = 1ull << (burst_warp & 31);
            else if (nwarps <= 32 && (pick & ((1ull << nwarps) - 1ull)) == 0ull) pick = ~0ull;
            const int start = (int)((rng_state >> 33) % (unsigned)nthreads);
            const int dir = (rng_state >> 20) & 1 ? 1 : nthreads - 1;
            const long long before = progress;
            for (int k = 0, t = start; k < nthreads; ++k, t = (t + dir) % nthreads) {
                if (fibers[t].done || !((pick >> ((t >> 5) & 31)) & 1ull)) continue;
                cur = t;
                threadIdx = {(unsigned)t, 0, 0};
                swapcontext(&sched_uc, &fibers[t].uc);
                if (fibers[t].done) { --alive; ++progress; }
            }
            if (burst_warp >= 0 && progress == before) burst_warp = -1;     // stuck (CTA barrier) or finished
        }
    }
}

}  // namespace simt

// ---- synchronisation and warp collectives --------------------------------------------------------
inline void __syncthreads() { simt::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return simt::exchange(v, src & 31); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return simt::exchange(v, (simt::cur & 31) - (int)d); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    const int s = (simt::cur & 31) + (int)d;
    return simt::exchange(v, s < 32 ? s : -1);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return simt::exchange(v, (simt::cur & 31) ^ m); }
inline unsigned __ballot_sync(unsigned, int pred) {
    const int w = simt::cur >> 5, l = simt::cur & 31;
    simt::wbuf[w][l] = pred ? 1ull : 0ull;
    simt::warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < simt::warp_size_of(w); ++i) m |= (unsigned)(simt::wbuf[w][i] & 1ull) << i;
    simt::warp_barrier();
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) {
    return __ballot_sync(mask, pred) == (simt::warp_size_of(simt::cur >> 5) == 32 ? 0xffffffffu : (1u << simt::warp_size_of(simt::cur >> 5)) - 1u);
}
template <typename T> inline unsigned __match_any_sync(unsigned, T v) {
    const int w = simt::cur >> 5, l = simt::cur & 31;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    simt::wbuf[w][l] = raw;
    simt::warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < simt::warp_size_of(w); ++i) m |= (unsigned)(simt::wbuf[w][i] == raw) << i;
    simt::warp_barrier();
    return m;
}

// ---- atomics (one OS thread: plain read-modify-write) ----------------------------------------------
template <typename T, typename U> inline T atomicAdd(T *p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> inline T atomicMax(T *p, U v) { const T old = *p; if ((T)v > old) *p = (T)v; return old; }

// ---- loads, conversions, arithmetic intrinsics -------------------------------------------------------
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
inline int __float2int_rn(float f) { return (int)std::nearbyintf(f); }      // round to nearest even, like cvt.rni
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; std::memcpy(&d, &b, 8); return d;
}
inline int __double2loint(double d) { long long i; std::memcpy(&i, &d, 8); return (int)(i & 0xffffffffll); }
inline int __double2hiint(double d) { long long i; std::memcpy(&i, &d, 8); return (int)(i >> 32); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

// CUDA's min / max accept mixed integer types
template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b) {
    using C = typename std::common_type<A, B>::type;
    return (C)a < (C)b ? (C)a : (C)b;
}
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b) {
    using C = typename std::common_type<A, B>::type;
    return (C)a > (C)b ? (C)a : (C)b;
}

// last: libstdc++ spells its own attribute __attribute__((__noinline__)), so this macro must not be visible to it
#undef __noinline__
#define __noinline__ __attribute__((noinline))
