/* C twin of oracle/pk_oracle.py  --  TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
 *
 * PARITY UNPINNED: restates pmesh 0.1.55 / nbodykit 0.3.14 semantics, which astrild pins
 * (/root/reference/poetry.lock:333-336, 478-481) but does not vendor; see pk_oracle.py.
 * Same arithmetic as the NumPy oracle (float64 throughout, scalar loops like pmesh's
 * _window_imp.c), usable at 512^3 and beyond where the NumPy formulation is too slow.
 * Anchors: src/astrild/particles/hutils/stats_subfind.py:130-131 (paint),
 *          src/astrild/power_spectra/power_spectrum_3d.py:189-195 (FFTPower binning).
 *
 * Build: make -C oracle   ->  oracle/_build/libpk_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline int64_t wrap(int64_t i, int64_t N) {
    i %= N;
    return i < 0 ? i + N : i;
}

/* per-axis window: base index and (renormalised) weights.  support 1/2/3 = NGP/CIC/TSC */
static inline void window_1d(int support, double g, int64_t *ipos, double *w) {
    int left = (support - 1) / 2;
    double shift = (support & 1) ? 0.5 : 0.0;
    int64_t i0 = (int64_t)floor(g + shift) - left;
    double dx = g - (double)i0, sum = 0.0;
    for (int j = 0; j < support; ++j) {
        double x = fabs(dx - j), v;
        if (support == 1) v = 1.0;
        else if (support == 2) v = x < 1.0 ? 1.0 - x : 0.0;
        else v = x <= 0.5 ? 0.75 - x * x : (x < 1.5 ? 0.5 * (1.5 - x) * (1.5 - x) : 0.0);
        w[j] = v;
        sum += v;
    }
    for (int j = 0; j < support; ++j) w[j] /= sum;
    *ipos = i0;
}

/* pm.paint(pos, mass, resampler): pos is AoS (Np,3) or SoA (3 pointers), f32 or f64.
 * layout 0 = AoS, 1 = SoA; pos_scale multiplies positions before *N/L is applied by the
 * caller's choice of L (box units [0,1): pass L = 1).  canvas (N^3 doubles) is accumulated into. */
void orc_paint(const void *p0, const void *p1, const void *p2, int layout, int is_f32,
               const void *mass, int mass_is_f32, int64_t Np, int N, double L, int support,
               double shift, double *canvas) {
    const double scale = (double)N / L;
    for (int64_t p = 0; p < Np; ++p) {
        double x[3];
        for (int d = 0; d < 3; ++d) {
            const void *base = layout ? (d == 0 ? p0 : d == 1 ? p1 : p2) : p0;
            int64_t idx = layout ? p : 3 * p + d;
            x[d] = is_f32 ? (double)((const float *)base)[idx] : ((const double *)base)[idx];
        }
        double m = 1.0;
        if (mass) m = mass_is_f32 ? (double)((const float *)mass)[p] : ((const double *)mass)[p];
        int64_t i0[3];
        double w[3][3];
        for (int d = 0; d < 3; ++d) window_1d(support, x[d] * scale + shift, &i0[d], w[d]);
        for (int jx = 0; jx < support; ++jx) {
            int64_t cx = wrap(i0[0] + jx, N) * (int64_t)N * N;
            for (int jy = 0; jy < support; ++jy) {
                int64_t cxy = cx + wrap(i0[1] + jy, N) * (int64_t)N;
                double wxy = m * w[0][jx] * w[1][jy];
                for (int jz = 0; jz < support; ++jz)
                    canvas[cxy + wrap(i0[2] + jz, N)] += wxy * w[2][jz];
            }
        }
    }
}

/* The same deposit for one of `nthreads` workers (pmesh's paint is single-threaded; this exists so that the CPU arm
 * of the benchmark can use all host cores).  The y axis is cut into 2*nthreads blocks of `by` rows (the last one takes
 * the remainder); in phase 0 worker `tid` deposits the particles whose HOME row lies in block 2*tid, in phase 1 those
 * of block 2*tid+1.  Blocks handled at the same time are a whole block apart (by >= support), so their windows never
 * meet, also across the periodic wrap (first block: phase 0, last block: phase 1); the caller runs phase 0 on all
 * workers, waits, then phase 1.  Deterministic for a given thread count; differs from orc_paint only in the order of
 * the float64 additions of cells next to a block boundary. */
void orc_paint_yblocks(const void *p0, const void *p1, const void *p2, int layout, int is_f32,
                       const void *mass, int mass_is_f32, int64_t Np, int N, double L, int support,
                       double shift, double *canvas, int tid, int nthreads, int phase) {
    const double scale = (double)N / L;
    const int nblocks = 2 * nthreads;
    const int by = N / nblocks;
    const int mine = 2 * tid + phase;
    for (int64_t p = 0; p < Np; ++p) {
        const void *basey = layout ? p1 : p0;
        int64_t iy = layout ? p : 3 * p + 1;
        double yy = is_f32 ? (double)((const float *)basey)[iy] : ((const double *)basey)[iy];
        /* home row = window base + left = floor(g + (support odd ? 0.5 : 0)), see window_1d: the scan over
         * the particles of other workers costs one floor each */
        int64_t hy = (int64_t)floor(yy * scale + shift + ((support & 1) ? 0.5 : 0.0));
        int blk = (int)(wrap(hy, N) / by);
        if (blk >= nblocks) blk = nblocks - 1;
        if (blk != mine) continue;
        int64_t i0y;
        double wy[3];
        window_1d(support, yy * scale + shift, &i0y, wy);
        double x[3];
        for (int d = 0; d < 3; d += 2) {
            const void *base = layout ? (d == 0 ? p0 : p2) : p0;
            int64_t idx = layout ? p : 3 * p + d;
            x[d] = is_f32 ? (double)((const float *)base)[idx] : ((const double *)base)[idx];
        }
        double m = 1.0;
        if (mass) m = mass_is_f32 ? (double)((const float *)mass)[p] : ((const double *)mass)[p];
        int64_t i0[3];
        double w[3][3];
        i0[1] = i0y;
        for (int j = 0; j < support; ++j) w[1][j] = wy[j];
        for (int d = 0; d < 3; d += 2) window_1d(support, x[d] * scale + shift, &i0[d], w[d]);
        for (int jx = 0; jx < support; ++jx) {
            int64_t cx = wrap(i0[0] + jx, N) * (int64_t)N * N;
            for (int jy = 0; jy < support; ++jy) {
                int64_t cxy = cx + wrap(i0[1] + jy, N) * (int64_t)N;
                double wxy = m * w[0][jx] * w[1][jy];
                for (int jz = 0; jz < support; ++jz)
                    canvas[cxy + wrap(i0[2] + jz, N)] += wxy * w[2][jz];
            }
        }
    }
}

/* numpy.digitize(x, bins) for increasing bins == searchsorted(bins, x, side='right') */
static inline int digitize(double x, const double *bins, int n) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x < bins[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* project_to_basis, one mu bin.  c1/c2: interleaved complex128 (N,N,Nk); c2 may be NULL
 * (auto).  kx,ky [N], kz [Nk]: host-built tables; edges2 [nedges] = kedges^2.
 * Outputs (nedges+1 long, under/overflow included) are accumulated into.
 * p = c1 conj(c2) * vol with the DC mode zeroed (nbodykit _compute_3d_power). */
void orc_bin_power(const double *c1, const double *c2, int N, const double *kx, const double *ky,
                   const double *kz, const double *edges2, int nedges, double vol, double *xsum,
                   double *ysum_re, double *ysum_im, int64_t *nsum, int ix_begin, int ix_end) {
    const int Nk = N / 2 + 1;
    if (!c2) c2 = c1;
    for (int ix = ix_begin; ix < ix_end; ++ix) {
        double kx2 = kx[ix] * kx[ix];
        for (int iy = 0; iy < N; ++iy) {
            double kxy2 = kx2 + ky[iy] * ky[iy];
            const double *a = c1 + 2 * ((int64_t)(ix * (int64_t)N + iy) * Nk);
            const double *b = c2 + 2 * ((int64_t)(ix * (int64_t)N + iy) * Nk);
            for (int iz = 0; iz < Nk; ++iz) {
                double k2 = kxy2 + kz[iz] * kz[iz];
                int dig = digitize(k2, edges2, nedges);
                int nonsing = kz[iz] > 0.0;
                double w = nonsing ? 2.0 : 1.0;
                double pre = (a[2 * iz] * b[2 * iz] + a[2 * iz + 1] * b[2 * iz + 1]) * vol;
                double pim = (a[2 * iz + 1] * b[2 * iz] - a[2 * iz] * b[2 * iz + 1]) * vol;
                if (ix == 0 && iy == 0 && iz == 0) { pre = 0.0; pim = 0.0; }
                xsum[dig] += w * sqrt(k2);
                nsum[dig] += nonsing ? 2 : 1;
                ysum_re[dig] += w * pre;
                if (!nonsing) ysum_im[dig] += pim;
            }
        }
    }
}
