// Particle -> mesh mass assignment, direct-atomic variant (APK_DEPOSIT_ATOMIC).
//
// Replaces pm.paint(pos, mass=, resampler=) as astrild calls it at
//   /root/reference/src/astrild/particles/hutils/stats_subfind.py:130-131
// (pmesh 0.1.55 window kernels: cell i centred on g = i; CIC base floor(g), TSC base
// floor(g+0.5)-1, periodic wrap of the index).  One thread per particle, 1/8/27
// RED.ADD.F32 into the fp32 mesh.  This is the small-catalogue path (10^6 halos) and the
// correctness baseline of the sorted path in deposit_sorted.cu; the index/weight arithmetic
// is done in float64 exactly like the oracle so both assign every particle to the same cells.
#include "apk_common.cuh"
#include "deposit_common.cuh"

namespace apk {

template <int S, typename PT, bool SOA>
__global__ void __launch_bounds__(256)
deposit_atomic_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2,
                      const void *__restrict__ mass, int mass_f64, long long np, DepositGeom G,
                      float *__restrict__ mesh) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        double x, y, z;
        if (SOA) { x = (double)p0[p]; y = (double)p1[p]; z = (double)p2[p]; }
        else     { x = (double)p0[3 * p]; y = (double)p0[3 * p + 1]; z = (double)p0[3 * p + 2]; }
        if (!owned_by_slab(__dmul_rn(x, G.scale), G)) continue;   // slab plans: another rank deposits it
        float m = 1.f;
        if (mass) m = mass_f64 ? (float)((const double *)mass)[p] : ((const float *)mass)[p];
        long long ix, iy, iz;
        float wx[S], wy[S], wz[S];
        window_1d<S>(grid_coord(x, G), ix, wx);
        window_1d<S>(grid_coord(y, G), iy, wy);
        window_1d<S>(grid_coord(z, G), iz, wz);
        int cy[S], cz[S];
#pragma unroll
        for (int j = 0; j < S; ++j) { cy[j] = wrap_index(iy + j, G.N); cz[j] = wrap_index(iz + j, G.N); }
#pragma unroll
        for (int jx = 0; jx < S; ++jx) {
            const int px = G.local_plane(ix + jx);
            if (px < 0) continue;   // slab plan: outside owned + ghost planes (caller error)
            const float wxm = wx[jx] * m;
#pragma unroll
            for (int jy = 0; jy < S; ++jy) {
                float *row = mesh + ((size_t)px * G.N + cy[jy]) * G.ldz;
                const float wxy = wxm * wy[jy];
#pragma unroll
                for (int jz = 0; jz < S; ++jz) atomicAdd(row + cz[jz], wxy * wz[jz]);
            }
        }
    }
}

template <int S>
static int launch_atomic(const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                         const void *mass, int mass_dtype, long long np, const DepositGeom &G,
                         float *mesh, int num_sms, cudaStream_t st) {
    const int threads = 256;
    long long want = (np + threads - 1) / threads;
    int blocks = (int)(want < (long long)num_sms * 16 ? (want > 0 ? want : 1) : (long long)num_sms * 16);
    const int mf64 = mass_dtype == APK_F64;
    if (pos_dtype == APK_F32) {
        if (layout == APK_SOA)
            deposit_atomic_kernel<S, float, true><<<blocks, threads, 0, st>>>((const float *)p0, (const float *)p1, (const float *)p2, mass, mf64, np, G, mesh);
        else
            deposit_atomic_kernel<S, float, false><<<blocks, threads, 0, st>>>((const float *)p0, nullptr, nullptr, mass, mf64, np, G, mesh);
    } else {
        if (layout == APK_SOA)
            deposit_atomic_kernel<S, double, true><<<blocks, threads, 0, st>>>((const double *)p0, (const double *)p1, (const double *)p2, mass, mf64, np, G, mesh);
        else
            deposit_atomic_kernel<S, double, false><<<blocks, threads, 0, st>>>((const double *)p0, nullptr, nullptr, mass, mf64, np, G, mesh);
    }
    APK_CUDA(cudaGetLastError());
    return 0;
}

int deposit_atomic_launch(const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                          const void *mass, int mass_dtype, long long np, int resampler,
                          const DepositGeom &G, float *mesh, int num_sms, cudaStream_t st) {
    if (np == 0) return 0;
    switch (resampler) {
        case APK_NGP: return launch_atomic<1>(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, num_sms, st);
        case APK_CIC: return launch_atomic<2>(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, num_sms, st);
        case APK_TSC: return launch_atomic<3>(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, num_sms, st);
    }
    set_error("apk_deposit: unknown resampler %d", resampler);
    return 2;
}

}  // namespace apk
