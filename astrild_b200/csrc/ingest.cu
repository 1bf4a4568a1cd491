// Ingest of gridded samples and of RAMSES / ECOSMOG output records on the device.
//
//   apk_assign_grid     value_map[(x, y, z)] = values with x = (npar * fields["x"]).astype(int)
//                       /root/reference/src/astrild/power_spectra/power_spectrum_3d.py:142-148   (PowerSpectrum3D._read_data)
//                       NGP *assignment*, not accumulation: one value per cell, truncation toward zero, negative indices
//                       wrap once (NumPy), a repeated index keeps the LAST sample's value.
//   apk_gather_records  the float64 blocks of Fortran unformatted records, picked out of the raw file bytes into
//                       contiguous columns      /root/reference/src/astrild/particles/ecosmog.py:184-230
//                       (Ecosmog.compress_snapshot: unpack("d" * ncache, content[pmin:pmax]) per field and cell octant).
// Both are HBM-bound copy / scatter kernels: coalesced reads, one pass (two for the assignment's tie rule).
#include "apk_common.cuh"

namespace apk {

// cell of sample i, or -1 if NumPy would raise IndexError (index outside [-N, N) on an axis)
template <typename CT>
__device__ __forceinline__ long long sample_cell(const CT *__restrict__ x, const CT *__restrict__ y, const CT *__restrict__ z,
                                                 long long i, int N) {
    const CT c[3] = {x[i], y[i], z[i]};
    long long cell = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const CT g = (CT)N * c[d];                      // npar * fields["x"].values, in the column's dtype
        if (!(g > (CT)-9.0e18 && g < (CT)9.0e18)) return -1;   // NaN / inf / out of int64: astype(int) is undefined there
        long long idx = (long long)g;                   // .astype(int): truncation toward zero
        if (idx < 0) idx += N;                          // NumPy: a negative index counts from the end
        if (idx < 0 || idx >= N) return -1;
        cell = cell * N + idx;
    }
    return cell;
}

// pass 1: the highest sample index wins a cell (NumPy assigns in order, so the last occurrence stays)
template <typename CT>
__global__ void __launch_bounds__(256)
assign_claim_kernel(const CT *__restrict__ x, const CT *__restrict__ y, const CT *__restrict__ z, long long n, int N,
                    unsigned int *__restrict__ winner, unsigned long long *__restrict__ bad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int nbad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long cell = sample_cell<CT>(x, y, z, i, N);
        if (cell < 0) { ++nbad; continue; }
        atomicMax(winner + cell, (unsigned int)(i + 1));
    }
    if (nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// pass 2: the winner writes its value
template <typename CT, typename VT>
__global__ void __launch_bounds__(256)
assign_write_kernel(const CT *__restrict__ x, const CT *__restrict__ y, const CT *__restrict__ z,
                    const VT *__restrict__ values, long long n, int N, const unsigned int *__restrict__ winner,
                    double *__restrict__ value_map) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long cell = sample_cell<CT>(x, y, z, i, N);
        if (cell >= 0 && winner[cell] == (unsigned int)(i + 1)) value_map[cell] = (double)values[i];
    }
}

// One CTA per piece: `count` float64 values starting at byte `src` of the raw file image (4-byte aligned only:
// Fortran record markers are 4 bytes) -> out[dst .. dst + count).
struct RecordPiece { long long src, dst, count; };

__global__ void __launch_bounds__(256)
gather_records_kernel(const unsigned char *__restrict__ raw, const RecordPiece *__restrict__ pieces, double *__restrict__ out) {
    const RecordPiece p = pieces[blockIdx.x];
    const unsigned int *w = reinterpret_cast<const unsigned int *>(raw + p.src);
    for (long long i = threadIdx.x; i < p.count; i += blockDim.x) {
        const unsigned int lo = w[2 * i], hi = w[2 * i + 1];        // little-endian halves of one double
        out[p.dst + i] = __hiloint2double((int)hi, (int)lo);
    }
}

static int ingest_grid(long long n, int num_sms) {
    long long want = (n + 255) / 256;
    long long cap = (long long)num_sms * 16;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <typename CT>
static int assign_typed(apk_plan *P, const void *x, const void *y, const void *z, const void *values, int val_dtype,
                        long long n, double *value_map, unsigned int *winner, unsigned long long *bad, cudaStream_t st) {
    const int N = P->N, g = ingest_grid(n, P->num_sms);
    assign_claim_kernel<CT><<<g, 256, 0, st>>>((const CT *)x, (const CT *)y, (const CT *)z, n, N, winner, bad);
    APK_CUDA(cudaGetLastError());
    if (val_dtype == APK_F64)
        assign_write_kernel<CT, double><<<g, 256, 0, st>>>((const CT *)x, (const CT *)y, (const CT *)z, (const double *)values, n, N, winner, value_map);
    else
        assign_write_kernel<CT, float><<<g, 256, 0, st>>>((const CT *)x, (const CT *)y, (const CT *)z, (const float *)values, n, N, winner, value_map);
    APK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace apk

using namespace apk;

extern "C" {

int apk_assign_grid(apk_plan *P, const void *x, const void *y, const void *z, int pos_dtype, const void *values,
                    int val_dtype, int64_t n, double *value_map, uint32_t *winner_scratch, uint64_t *bad_count_dev,
                    void *stream) {
    APK_REQUIRE(P && value_map && winner_scratch && bad_count_dev, "apk_assign_grid: null argument");
    APK_REQUIRE(n == 0 || (x && y && z && values), "apk_assign_grid: null sample arrays");
    APK_REQUIRE(n >= 0 && n < 0xffffffffLL, "apk_assign_grid: sample count %lld out of range", (long long)n);
    APK_REQUIRE(P->n0 == P->N, "apk_assign_grid: single-GPU plans only");
    APK_REQUIRE((pos_dtype == APK_F32 || pos_dtype == APK_F64) && (val_dtype == APK_F32 || val_dtype == APK_F64), "apk_assign_grid: bad dtype");
    DeviceGuard guard(P->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cells = (size_t)P->N * P->N * P->N;
    APK_CUDA(cudaMemsetAsync(value_map, 0, sizeof(double) * cells, st));          // np.zeros((npar, npar, npar))
    APK_CUDA(cudaMemsetAsync(winner_scratch, 0, sizeof(uint32_t) * cells, st));
    APK_CUDA(cudaMemsetAsync(bad_count_dev, 0, sizeof(uint64_t), st));
    if (n == 0) return 0;
    return pos_dtype == APK_F64
               ? assign_typed<double>(P, x, y, z, values, val_dtype, n, value_map, winner_scratch, (unsigned long long *)bad_count_dev, st)
               : assign_typed<float>(P, x, y, z, values, val_dtype, n, value_map, winner_scratch, (unsigned long long *)bad_count_dev, st);
}

int apk_gather_records(const void *raw_dev, const int64_t *pieces_dev, int64_t npieces, double *out, int device, void *stream) {
    APK_REQUIRE(npieces == 0 || (raw_dev && pieces_dev && out), "apk_gather_records: null argument");
    APK_REQUIRE(npieces >= 0 && npieces < 0x7fffffffLL, "apk_gather_records: %lld pieces", (long long)npieces);
    if (npieces == 0) return 0;
    DeviceGuard guard(device);
    gather_records_kernel<<<(unsigned int)npieces, 256, 0, (cudaStream_t)stream>>>((const unsigned char *)raw_dev,
                                                                                  (const RecordPiece *)pieces_dev, out);
    APK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
