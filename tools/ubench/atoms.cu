// Microbenchmark: shared-memory atomic / RMW throughput on sm_100a with decorrelated lanes.
// Question: is a particle-parallel deposit with fixed-point integer shared-memory
// accumulation (native ATOMS.ADD) viable?  Reports cycles per warp-instruction per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms atoms.cu && ./atoms
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILE = 3584;      // 14 x 8 x 32 ints
constexpr int THREADS = 256;
constexpr int ITERS = 2048;
constexpr int UPD = 27;

enum Mode { DISTINCT_BANKS = 0, RANDOM_ADDR = 1, JITTERED_Z = 2, SAME_ADDR = 3, DUP2 = 4, DUP4 = 5, CONF2 = 6, CONF4 = 7 };

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }

// address of this lane for iteration it under the access pattern
template <int MODE>
__device__ __forceinline__ int base_addr(int lane, int warp, int it, unsigned &rng) {
    if (MODE == DISTINCT_BANKS) return ((it * 37 + warp * 5) & 63) * 32 + ((lane + it) & 31);   // one lane per bank
    if (MODE == RANDOM_ADDR) return lcg(rng) >> 8 & 2047;                                        // random cell
    if (MODE == JITTERED_Z) {   // lattice line along z with +-1.5 cell jitter, random (x,y) column
        const int z = (lane + (int)((lcg(rng) >> 20) % 4) - 1) & 31;
        return ((it * 13 + warp) % 60) * 32 + z;
    }
    if (MODE == DUP2) return ((it * 37 + warp * 5) & 63) * 32 + (((lane >> 1) + it) & 31);        // 16 addresses, 2 lanes each
    if (MODE == DUP4) return ((it * 37 + warp * 5) & 63) * 32 + (((lane >> 2) + it) & 31);        // 8 addresses, 4 lanes each
    if (MODE == CONF2) return (((it * 37 + warp * 5) & 31) * 2 + (lane & 1)) * 32 + (((lane >> 1) + it) & 31);   // 2 addresses per bank
    if (MODE == CONF4) return (((it * 37 + warp * 5) & 15) * 4 + (lane & 3)) * 32 + (((lane >> 2) + it) & 31);   // 4 addresses per bank
    return 17;
}

template <int MODE, typename T>
__global__ void __launch_bounds__(THREADS) k_atoms(T *out, long long *cycles) {
    __shared__ T tile[TILE + 27 * 32 + 1024];
    for (int i = threadIdx.x; i < TILE + 27 * 32 + 1024; i += THREADS) tile[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned rng = threadIdx.x * 2654435761u + blockIdx.x * 97u + 12345u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        const int b = base_addr<MODE>(lane, warp, it, rng);
        const T v = (T)(it + lane);
#pragma unroll
        for (int u = 0; u < UPD; ++u) {
            const int off = (u / 9) * 8 * 32 + ((u / 3) % 3) * 32 + (u % 3);   // window offsets in a 14 x 8 x 32 tile
            atomicAdd(&tile[b + off], v + (T)u);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    T s = 0;
    for (int i = threadIdx.x; i < TILE; i += THREADS) s += tile[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

// plain read-modify-write (LDS + add + STS), racy across warps: throughput reference only
template <int MODE>
__global__ void __launch_bounds__(THREADS) k_rmw(float *out, long long *cycles) {
    __shared__ float tile[TILE + 27 * 32 + 1024];
    for (int i = threadIdx.x; i < TILE + 27 * 32 + 1024; i += THREADS) tile[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned rng = threadIdx.x * 2654435761u + blockIdx.x * 97u + 12345u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        const int b = base_addr<MODE>(lane, warp, it, rng);
        const float v = (float)(it + lane);
#pragma unroll
        for (int u = 0; u < UPD; ++u) {
            const int off = (u / 9) * 8 * 32 + ((u / 3) % 3) * 32 + (u % 3);
            volatile float *p = &tile[b + off];
            *p = *p + v;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0;
    for (int i = threadIdx.x; i < TILE; i += THREADS) s += tile[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

// float -> fixed conversion throughput: F2I vs FFMA-magic + IADD (no memory traffic)
template <int KIND>
__global__ void __launch_bounds__(THREADS) k_cvt(int *out, long long *cycles) {
    float w = 0.001f * threadIdx.x, z = 0.37f;
    int acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UPD; ++u) {
            const float a = w + 0.01f * u;
            if (KIND == 0) acc += __float2int_rn(a * z * 1048576.f);
            else acc += __float_as_int(fmaf(a, z, 12.0f)) - 0x41400000;
        }
        w += 1e-4f;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

template <typename F>
static void run(const char *name, F launch, int ctas_per_sm, int sms) {
    long long *cyc; void *out;
    const int ctas = ctas_per_sm * sms;
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    cudaMalloc(&out, 8 * (size_t)ctas * THREADS);
    launch(ctas, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    launch(ctas, out, cyc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[4096];
    cudaMemcpy(h, cyc, sizeof(long long) * (ctas < 4096 ? ctas : 4096), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < ctas && i < 4096; ++i) mean += h[i]; mean /= (ctas < 4096 ? ctas : 4096);
    const double warp_instr_per_sm = (double)ctas_per_sm * (THREADS / 32) * ITERS * UPD;
    cudaError_t e = cudaGetLastError();
    printf("%-44s ctas/SM %d  %8.3f ms  %7.2f cycles per warp-instruction per SM (clock64)  %s\n", name, ctas_per_sm, ms,
           mean / warp_instr_per_sm, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %d kHz\n", p.name, sms, p.clockRate);
    for (int c : {1, 2, 4}) {
        run("ATOMS.ADD u32 distinct banks", [](int g, void *o, long long *c) { k_atoms<DISTINCT_BANKS, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 random cells", [](int g, void *o, long long *c) { k_atoms<RANDOM_ADDR, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 jittered z-line", [](int g, void *o, long long *c) { k_atoms<JITTERED_Z, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 one address", [](int g, void *o, long long *c) { k_atoms<SAME_ADDR, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 2 lanes per address", [](int g, void *o, long long *c) { k_atoms<DUP2, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 4 lanes per address", [](int g, void *o, long long *c) { k_atoms<DUP4, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 2 addresses per bank", [](int g, void *o, long long *c) { k_atoms<CONF2, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u32 4 addresses per bank", [](int g, void *o, long long *c) { k_atoms<CONF4, unsigned><<<g, THREADS>>>((unsigned *)o, c); }, c, sms);
        run("ATOMS.ADD u64 distinct banks", [](int g, void *o, long long *c) { k_atoms<DISTINCT_BANKS, unsigned long long><<<g, THREADS>>>((unsigned long long *)o, c); }, c, sms);
        run("ATOMS.ADD u64 random cells", [](int g, void *o, long long *c) { k_atoms<RANDOM_ADDR, unsigned long long><<<g, THREADS>>>((unsigned long long *)o, c); }, c, sms);
        run("atomicAdd f32 (CAS loop?) distinct banks", [](int g, void *o, long long *c) { k_atoms<DISTINCT_BANKS, float><<<g, THREADS>>>((float *)o, c); }, c, sms);
        run("atomicAdd f32 random cells", [](int g, void *o, long long *c) { k_atoms<RANDOM_ADDR, float><<<g, THREADS>>>((float *)o, c); }, c, sms);
        run("LDS+FADD+STS distinct banks", [](int g, void *o, long long *c) { k_rmw<DISTINCT_BANKS><<<g, THREADS>>>((float *)o, c); }, c, sms);
        run("LDS+FADD+STS random cells", [](int g, void *o, long long *c) { k_rmw<RANDOM_ADDR><<<g, THREADS>>>((float *)o, c); }, c, sms);
        run("FMUL+FMUL+F2I", [](int g, void *o, long long *c) { k_cvt<0><<<g, THREADS>>>((int *)o, c); }, c, sms);
        run("FFMA(magic)+IADD", [](int g, void *o, long long *c) { k_cvt<1><<<g, THREADS>>>((int *)o, c); }, c, sms);
    }
    return 0;
}
