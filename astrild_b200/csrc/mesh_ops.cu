// Gridded-field helpers: the ArrayMesh side of the boundary.
//   ArrayMesh(value_map, BoxSize=...)   /root/reference/src/astrild/power_spectra/power_spectrum_3d.py:183-188, 197-212
//   value_map = paint(...).value / dx^3 /root/reference/src/astrild/particles/hutils/stats_subfind.py:131-132
// All HBM-bound elementwise/reduction kernels: grid-stride, coalesced along z.
#include "apk_common.cuh"

namespace apk {

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256) sum_kernel(const T *__restrict__ a, long long n, double *out) {
    double s = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += (double)a[i];
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

// padded mesh [rows][ldz] -> sum over the first N of every row
__global__ void __launch_bounds__(256) padded_sum_kernel(const float *__restrict__ a, long long rows, int N, int ldz, double *out) {
    double s = 0.0;
    const long long total = rows * N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / N;
        const int z = (int)(i - r * N);
        s += (double)a[r * ldz + z];
    }
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) load_mesh_kernel(const T *__restrict__ in, long long rows, int N, int ldz, double mean, float *__restrict__ mesh) {
    const long long total = rows * ldz;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / ldz;
        const int z = (int)(i - r * ldz);
        mesh[i] = z < N ? (float)((double)in[r * N + z] - mean) : 0.f;
    }
}

__global__ void __launch_bounds__(256) store_mesh_kernel(const float *__restrict__ mesh, long long rows, int N, int ldz, double scale, double *__restrict__ out) {
    const long long total = rows * N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / N;
        const int z = (int)(i - r * N);
        out[i] = (double)mesh[r * ldz + z] * scale;
    }
}

static int grid_for(long long n, int num_sms) {
    long long want = (n + 255) / 256;
    long long cap = (long long)num_sms * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

int mesh_sum_launch(apk_plan *P, const void *v, int dtype, double *sum_dev, cudaStream_t st) {
    const long long n = (long long)P->n0 * P->N * P->N;
    APK_CUDA(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
    if (dtype == APK_F64) sum_kernel<double><<<grid_for(n, P->num_sms), 256, 0, st>>>((const double *)v, n, sum_dev);
    else sum_kernel<float><<<grid_for(n, P->num_sms), 256, 0, st>>>((const float *)v, n, sum_dev);
    APK_CUDA(cudaGetLastError());
    return 0;
}

int padded_mesh_sum_launch(apk_plan *P, const float *mesh, double *sum_dev, cudaStream_t st) {
    const long long rows = (long long)P->n0 * P->N;
    APK_CUDA(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
    padded_sum_kernel<<<grid_for(rows * P->N, P->num_sms), 256, 0, st>>>(mesh, rows, P->N, P->ldz, sum_dev);
    APK_CUDA(cudaGetLastError());
    return 0;
}

int load_mesh_launch(apk_plan *P, const void *v, int dtype, double mean, float *mesh, cudaStream_t st) {
    const long long rows = (long long)P->n0 * P->N;
    const int g = grid_for(rows * P->ldz, P->num_sms);
    if (dtype == APK_F64) load_mesh_kernel<double><<<g, 256, 0, st>>>((const double *)v, rows, P->N, P->ldz, mean, mesh);
    else load_mesh_kernel<float><<<g, 256, 0, st>>>((const float *)v, rows, P->N, P->ldz, mean, mesh);
    APK_CUDA(cudaGetLastError());
    return 0;
}

int store_mesh_launch(apk_plan *P, const float *mesh, double scale, double *out, cudaStream_t st) {
    const long long rows = (long long)P->n0 * P->N;
    store_mesh_kernel<<<grid_for(rows * P->N, P->num_sms), 256, 0, st>>>(mesh, rows, P->N, P->ldz, scale, out);
    APK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace apk
