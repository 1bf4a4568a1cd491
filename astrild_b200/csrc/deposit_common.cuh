// Window kernels and mesh geometry shared by the deposit variants (pmesh semantics, see
// deposit_atomic.cu for the reference citation).
#pragma once
#include <cuda_runtime.h>

namespace apk {

struct DepositGeom {
    int N;          // cells per side
    int ldz;        // padded floats per z-row
    double scale;   // grid units per position unit: pos_scale * N
    double shift;   // affine shift in grid units (0.5 for the interlaced twin)
    int slab;       // 0: whole periodic mesh, 1: slab with ghost planes
    int plane0;     // global index of local plane 0 (x0 - ghost_lo), may be negative
    int nplanes;    // local planes (ghost_lo + n0 + ghost_hi), == N when !slab
    float s0, s1, s2;   // scale as an unevaluated sum of three floats (error-free float32 index path)
    float t32;          // shift + (0.5 for TSC/NGP rounding) when exactly representable, else < 0 (use float64)
    int own0;       // first owned plane x0
    int nown;       // owned planes n0; slab plans deposit only particles with floor(g_x - shift) owned

    __host__ __device__ __forceinline__ int local_plane(long long ix) const;
};

// far-out-of-box indices: kept out of line so the integer division is never speculated
__host__ __device__ __noinline__ static int wrap_index_slow(long long i, int N) {
    long long r = i % N;
    return (int)(r < 0 ? r + N : r);
}

// periodic wrap of a cell index; positions within one box length of the box take the fast path
__host__ __device__ __forceinline__ int wrap_index(long long i, int N) {
    if (i < 0) i += N; else if (i >= N) i -= N;
    if ((unsigned long long)i >= (unsigned long long)N) return wrap_index_slow(i, N);
    return (int)i;
}

__host__ __device__ __forceinline__ int wrap_index32(int i, int N) {
    if (i < 0) i += N; else if (i >= N) i -= N;
    if ((unsigned)i >= (unsigned)N) return wrap_index_slow((long long)i, N);
    return i;
}

// branch-free wrap for indices within one box length of the box; `far` collects the others so the
// caller can take ONE rare slow path per particle instead of a division per axis
__device__ __forceinline__ int wrap_near(int i, int N, bool &far) {
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    far |= (unsigned)i >= (unsigned)N;
    return i;
}

__host__ __device__ __forceinline__ int DepositGeom::local_plane(long long ix) const {
    const int gx = wrap_index(ix, N);
    if (!slab) return gx;
    int rel = gx - plane0;
    if (rel < 0) rel += N;
    else if (rel >= N) rel -= N;
    return rel < nplanes ? rel : -1;
}

// g = x * scale + shift exactly as the oracle (NumPy, float64) forms it: a rounded product, then a rounded sum.  Written
// with the intrinsics because nvcc contracts `x * scale + shift` into ONE fused multiply-add (the Makefile's
// -ffp-contract=off reaches only the host compiler), which rounds differently for one position in ~10^13.
__device__ __forceinline__ double grid_coord(double x, const DepositGeom &G) {
    return __dadd_rn(__dmul_rn(x, G.scale), G.shift);
}

// slab ownership of a particle: floor of its UNSHIFTED grid coordinate lies in [own0, own0 + nown)
__device__ __forceinline__ bool owned_by_slab(double gx_unshifted, const DepositGeom &G) {
    if (!G.slab) return true;
    int rel = wrap_index((long long)floor(gx_unshifted), G.N) - G.own0;
    return rel >= 0 && rel < G.nown;
}

// base index and weights of pmesh's window of support S at grid coordinate g
template <int S>
__device__ __forceinline__ void window_1d(double g, long long &i0, float (&w)[S]) {
    if (S == 1) {
        i0 = (long long)floor(g + 0.5);
        w[0] = 1.f;
    } else if (S == 2) {
        const double f = floor(g);
        i0 = (long long)f;
        const double d = g - f;
        w[0] = (float)(1.0 - d);
        w[S - 1] = (float)d;
    } else {
        const double f = floor(g + 0.5) - 1.0;
        i0 = (long long)f;
        const double dx = g - f;            // in [0.5, 1.5)
        const double a = 1.5 - dx, b = dx - 1.0, c = dx - 0.5;
        w[0] = (float)(0.5 * a * a);
        w[S / 2] = (float)(0.75 - b * b);
        w[S - 1] = (float)(0.5 * c * c);
    }
}

}  // namespace apk
