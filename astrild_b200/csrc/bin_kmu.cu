// (k, mu) binning and multipoles: nbodykit's project_to_basis with Nmu > 1 and poles, fused with the interlace-combine,
// window deconvolution and c1 conj(c2) like the 1-D kernel (bin_power.cu).  SURVEY.md section 8f row N4 -- astrild hints
// at redshift-space use (/root/reference/README.md:11, src/astrild/particles/hutils/tpcf.py:12-60); the semantics are
// FFTPower(mode="2d", Nmu=, poles=, los=) of nbodykit 0.3.14 (algorithms/fftpower.py: FFTPower.run, project_to_basis):
//   mu = (k . los) / |k| (0 at k = 0),  mu bin = numpy.digitize(|mu|, linspace(0, 1, Nmu + 1)),  Hermitian weight 2 on the
//   non-singular planes,  ysum_ell += (2 ell + 1) L_ell(mu) P with the real part doubled / imaginary part dropped for
//   even ell on non-singular modes and the reverse for odd ell.
// Who decides which float lands on which side of an edge: k^2 = (ka^2 + kb^2) + kz^2, |k| = sqrt(k^2), k . los and
// mu = (k . los) / |k| are IEEE float64 operations in the oracle's order (no FMA contraction: __dmul_rn / __dadd_rn;
// sqrt and division are correctly rounded), and both digitizes are float guesses FIXED by float64 compares against the
// uploaded edges -- mode counts per (k, mu) bin are bit-identical to the oracle's.
//
// Bound: HBM for the grid reads (8 bytes per mode and grid) -- in practice the L2 atomics: a thread sends 3 + 2 nell
// fire-and-forget REDs (f64 / u64) into one of NCOPY private histograms whenever the (k, mu) bin of its walk changes;
// a last kernel folds the copies.  This mode is not on BASELINE.json's metric: built for coverage and parity.
#include "apk_common.cuh"
#include <cmath>

struct apk_kmu {
    apk_plan *plan = nullptr;
    int n_a = 0, n_b = 0, nz = 0, nedges = 0, nmu = 0, nell = 0;
    int ells[8] = {};
    double los[3] = {0, 0, 1};
    int dc_a = -1, dc_b = -1;
    bool has_comp = false, has_phase = false;
    void *tables = nullptr;
    double *ka = nullptr, *kb = nullptr, *kz = nullptr, *edges2 = nullptr, *muedges = nullptr;
    float *wz = nullptr, *ic_a = nullptr, *ic_b = nullptr, *ic_z = nullptr;
    float2 *ph_a = nullptr, *ph_b = nullptr, *ph_z = nullptr;
    double *copies = nullptr;     // [NCOPY][3 + 2 nell][nbins] (xsum, musum, nsum as u64, ysum_re[nell], ysum_im[nell])
    double kmin_guess = 0.0, inv_dk_guess = 0.0;
};

namespace apk {

constexpr int KMU_THREADS = 256, KMU_NCOPY = 32;

struct KmuArgs {
    const float2 *c1, *c1s, *c2, *c2s;
    const double *ka, *kb, *kz, *edges2, *muedges;
    const float *wz, *ic_a, *ic_b, *ic_z;
    const float2 *ph_a, *ph_b, *ph_z;
    int n_a, n_b, nz, nedges, nmu, nell;
    int ells[8];
    double los[3];
    int dc_a, dc_b;
    float kmin_f, inv_dk_f;
    double *copies;
    long long nbins;              // (nedges + 1) * (nmu + 2)
};

__device__ __forceinline__ float2 kmu_cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void kmu_red_f64(double *addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void kmu_red_u64(unsigned long long *addr, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// P_ell(mu) by Bonnet's recursion, the oracle's expression order
__device__ __forceinline__ double kmu_legendre(int ell, double mu) {
    double p0 = 1.0, p1 = mu;
    if (ell == 0) return p0;
    for (int n = 1; n < ell; ++n) {
        const double t = ((2 * n + 1) * mu * p1 - n * p0) / (n + 1);
        p0 = p1; p1 = t;
    }
    return p1;
}

// One thread per stored mode; a warp takes 32 consecutive iz (256 contiguous bytes per grid and row) of one a-row and
// walks a segment of b.  Along the walk |k| and mu move slowly, so a thread keeps the sums of its CURRENT (k, mu) bin in
// registers and sends them (3 + 2 nell REDs) only when the bin changes.  Modes outside [kedges[0], kedges[-1]) are
// neither read nor accumulated (nbodykit slices those rows away; for kmax = Nyquist they are ~48 % of the grid and would
// all hit the same few addresses).
constexpr int KMU_MAXL = 8, KMU_SEG = 32;

template <bool INTERLACED, bool CROSS, bool COMP>
__global__ void __launch_bounds__(KMU_THREADS)
bin_kmu_kernel(KmuArgs A) {
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (KMU_THREADS / 32);
    const long long warp_id = (long long)blockIdx.x * (KMU_THREADS / 32) + (threadIdx.x >> 5);
    const int n_zc = (A.nz + 31) / 32, n_sb = (A.n_b + KMU_SEG - 1) / KMU_SEG;
    const long long n_items = (long long)A.n_a * n_sb * n_zc;
    double *copy = A.copies + (size_t)(blockIdx.x % KMU_NCOPY) * (3 + 2 * A.nell) * A.nbins;
    double *g_x = copy, *g_mu = copy + A.nbins;
    unsigned long long *g_n = reinterpret_cast<unsigned long long *>(copy + 2 * A.nbins);
    double *g_re = copy + 3 * A.nbins, *g_im = g_re + (size_t)A.nell * A.nbins;
    const int nedges = A.nedges, nmu = A.nmu, nell = A.nell;
    const double e2_first = A.edges2[0], e2_last = A.edges2[nedges - 1];

    for (long long item = warp_id; item < n_items; item += warps_total) {
        const int zc = (int)(item % n_zc);
        const long long r = item / n_zc;
        const int sb = (int)(r % n_sb), ia = (int)(r / n_sb);
        const int iz = zc * 32 + lane;
        if (iz >= A.nz) continue;
        const double ka = A.ka[ia], kz = A.kz[iz];
        const double ka2 = __dmul_rn(ka, ka), kz2 = __dmul_rn(kz, kz);
        const double kal = __dmul_rn(ka, A.los[0]), kzl = __dmul_rn(kz, A.los[2]);
        const bool nonsingular = A.wz[iz] > 1.5f;
        const double w = nonsingular ? 2.0 : 1.0;
        float2 phaz = make_float2(1.f, 0.f);
        float icaz = 1.f;
        if (INTERLACED) phaz = kmu_cmul(A.ph_a[ia], A.ph_z[iz]);
        if (COMP) icaz = A.ic_a[ia] * A.ic_z[iz];
        // sums of the current bin
        long long cur = -1;
        double ax = 0.0, am = 0.0, are[KMU_MAXL], aim[KMU_MAXL];
        unsigned int an = 0;
#pragma unroll
        for (int i = 0; i < KMU_MAXL; ++i) { are[i] = 0.0; aim[i] = 0.0; }
        auto flush = [&]() {
            if (cur < 0) return;
            kmu_red_f64(g_x + cur, ax);
            kmu_red_f64(g_mu + cur, am);
            kmu_red_u64(g_n + cur, (unsigned long long)an);
#pragma unroll
            for (int i = 0; i < KMU_MAXL; ++i)
                if (i < nell) {
                    if (are[i] != 0.0) kmu_red_f64(g_re + (size_t)i * A.nbins + cur, are[i]);
                    if (aim[i] != 0.0) kmu_red_f64(g_im + (size_t)i * A.nbins + cur, aim[i]);
                    are[i] = 0.0; aim[i] = 0.0;
                }
            ax = 0.0; am = 0.0; an = 0;
        };
        const int ib1 = min((sb + 1) * KMU_SEG, A.n_b);
        for (int ib = sb * KMU_SEG; ib < ib1; ++ib) {
            const double kb = A.kb[ib];
            const double k2 = __dadd_rn(__dadd_rn(ka2, __dmul_rn(kb, kb)), kz2);
            if (k2 >= e2_last || k2 < e2_first) continue;          // outside the edges: not part of any visible bin
            // ---- k bin: numpy.digitize(k2, kedges^2) ----
            const double knorm = sqrt(k2);
            int kbin = (int)(((float)knorm - A.kmin_f) * A.inv_dk_f) + 1;
            kbin = max(1, min(kbin, nedges - 1));
            while (k2 < A.edges2[kbin - 1]) --kbin;
            while (k2 >= A.edges2[kbin]) ++kbin;
            // ---- mu and its bin: numpy.digitize(|mu|, linspace(0, 1, Nmu + 1)) -> 1 .. Nmu + 1 ----
            const double kdl = __dadd_rn(__dadd_rn(kal, __dmul_rn(kb, A.los[1])), kzl);
            const double mu = knorm > 0.0 ? kdl / knorm : 0.0;
            const double amu = fabs(mu);
            int mbin = min(max((int)(amu * nmu) + 1, 1), nmu + 1);
            while (mbin > 1 && amu < A.muedges[mbin - 1]) --mbin;
            while (mbin <= nmu && amu >= A.muedges[mbin]) ++mbin;
            const long long bin = (long long)kbin * (nmu + 2) + mbin;
            if (bin != cur) { flush(); cur = bin; }
            // ---- the mode's power ----
            const size_t idx = ((size_t)ia * A.n_b + ib) * A.nz + iz;
            float2 a = __ldcs(A.c1 + idx);
            float2 ph = make_float2(1.f, 0.f);
            if (INTERLACED) {
                ph = kmu_cmul(phaz, A.ph_b[ib]);
                const float2 s = kmu_cmul(__ldcs(A.c1s + idx), ph);
                a = make_float2(0.5f * (a.x + s.x), 0.5f * (a.y + s.y));
            }
            float pre, pim;
            if (CROSS) {
                float2 b = __ldcs(A.c2 + idx);
                if (INTERLACED) {
                    const float2 s = kmu_cmul(__ldcs(A.c2s + idx), ph);
                    b = make_float2(0.5f * (b.x + s.x), 0.5f * (b.y + s.y));
                }
                pre = a.x * b.x + a.y * b.y;
                pim = a.y * b.x - a.x * b.y;
            } else {
                pre = a.x * a.x + a.y * a.y;
                pim = 0.f;
            }
            if (COMP) { const float ic = icaz * A.ic_b[ib]; pre *= ic; pim *= ic; }
            if (iz == 0 && ia == A.dc_a && ib == A.dc_b) { pre = 0.f; pim = 0.f; }
            // ---- sums ----
            ax += w * knorm;
            am += w * amu;
            an += nonsingular ? 2u : 1u;
#pragma unroll
            for (int i = 0; i < KMU_MAXL; ++i)
                if (i < nell) {
                    const int ell = A.ells[i];
                    const double f = (2.0 * ell + 1.0) * kmu_legendre(ell, mu);
                    double re = f * (double)pre, im = f * (double)pim;
                    if (nonsingular) {
                        if (ell & 1) { re = 0.0; im *= 2.0; }
                        else { re *= 2.0; im = 0.0; }
                    }
                    are[i] += re; aim[i] += im;
                }
        }
        flush();
    }
}

// fold the private copies in a fixed order; the internal mu == 1 column (Nmu + 1) goes into the last visible one
__global__ void __launch_bounds__(256)
kmu_fold_kernel(const double *copies, int nell, long long nbins, int nmu, double *xsum, double *musum, long long *nsum,
                double *ysum_re, double *ysum_im) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const int m = (int)(b % (nmu + 2));
    const int narr = 3 + 2 * nell;
    for (int arr = 0; arr < narr; ++arr) {
        double s = 0.0;
        unsigned long long n = 0;
        for (int c = 0; c < KMU_NCOPY; ++c) {
            const double *src = copies + ((size_t)c * narr + arr) * nbins;
            for (int part = 0; part < 2; ++part) {
                if (part == 1 && m != nmu) break;                 // the last visible column also takes column Nmu + 1
                const long long q = b + part;
                if (arr == 2) n += reinterpret_cast<const unsigned long long *>(src)[q];
                else s += src[q];
            }
        }
        if (arr == 0) xsum[b] = s;
        else if (arr == 1) musum[b] = s;
        else if (arr == 2) nsum[b] = (long long)n;
        else if (arr < 3 + nell) ysum_re[(size_t)(arr - 3) * nbins + b] = s;
        else ysum_im[(size_t)(arr - 3 - nell) * nbins + b] = s;
    }
}

}  // namespace apk

using namespace apk;

extern "C" int apk_kmu_create(apk_kmu **out, apk_plan *P, int n_a, int n_b, int nz, const double *ka, const double *kb,
                              const double *kz, const double *wz, const double *kedges, int nedges, int nmu,
                              const int *ells, int nell, const double *los, const double *comp_a, const double *comp_b,
                              const double *comp_z, const double *phase_a, const double *phase_b, const double *phase_z,
                              int dc_a, int dc_b) {
    APK_REQUIRE(out && P && ka && kb && kz && wz && kedges && ells && los, "apk_kmu_create: null argument");
    APK_REQUIRE(n_a >= 1 && n_b >= 1 && nz >= 1, "apk_kmu_create: empty grid");
    APK_REQUIRE(nedges >= 2 && nedges <= 65536, "apk_kmu_create: need 2..65536 k edges, got %d", nedges);
    APK_REQUIRE(nmu >= 1 && nmu <= 4096, "apk_kmu_create: need 1..4096 mu bins, got %d", nmu);
    APK_REQUIRE(nell >= 1 && nell <= 8 && ells[0] == 0, "apk_kmu_create: 1..8 multipoles, the monopole first");
    for (int i = 0; i < nell; ++i) APK_REQUIRE(ells[i] >= 0 && ells[i] <= 16, "apk_kmu_create: multipole %d out of range", ells[i]);
    for (int i = 1; i < nedges; ++i)
        APK_REQUIRE(kedges[i] > kedges[i - 1] && kedges[0] >= 0.0, "apk_kmu_create: kedges must be non-negative and increasing");
    const bool has_comp = comp_a || comp_b || comp_z, has_phase = phase_a || phase_b || phase_z;
    APK_REQUIRE(!has_comp || (comp_a && comp_b && comp_z), "apk_kmu_create: give all three compensation tables or none");
    APK_REQUIRE(!has_phase || (phase_a && phase_b && phase_z), "apk_kmu_create: give all three phase tables or none");
    DeviceGuard guard(P->device);
    apk_kmu *K = new apk_kmu();
    K->plan = P; K->n_a = n_a; K->n_b = n_b; K->nz = nz; K->nedges = nedges; K->nmu = nmu; K->nell = nell;
    for (int i = 0; i < nell; ++i) K->ells[i] = ells[i];
    for (int d = 0; d < 3; ++d) K->los[d] = los[d];
    K->dc_a = dc_a; K->dc_b = dc_b; K->has_comp = has_comp; K->has_phase = has_phase;
    K->kmin_guess = kedges[0]; K->inv_dk_guess = 1.0 / (kedges[1] - kedges[0]);

    const size_t nd = (size_t)n_a + n_b + nz + nedges + nmu + 1;
    const size_t nc = has_phase ? (size_t)n_a + n_b + nz : 0;
    const size_t nf = (size_t)nz + (has_comp ? n_a + n_b + nz : 0);
    const size_t bytes = nd * 8 + nc * 8 + ((nf + 1) & ~(size_t)1) * 4;
    std::vector<unsigned char> host(bytes);
    double *hd = (double *)host.data();
    float2 *hc = (float2 *)(hd + nd);
    float *hf = (float *)(hc + nc);
    size_t o = 0;
    for (int i = 0; i < n_a; ++i) hd[o++] = ka[i];
    for (int i = 0; i < n_b; ++i) hd[o++] = kb[i];
    for (int i = 0; i < nz; ++i) hd[o++] = kz[i];
    for (int i = 0; i < nedges; ++i) hd[o++] = kedges[i] * kedges[i];
    // numpy.linspace(0, 1, Nmu + 1): i * step with step = 1 / Nmu, the last point set to the stop value
    for (int i = 0; i <= nmu; ++i) hd[o++] = i == nmu ? 1.0 : (double)i * (1.0 / (double)nmu);
    o = 0;
    if (has_phase) {
        const double *src[3] = {phase_a, phase_b, phase_z};
        const int cnt[3] = {n_a, n_b, nz};
        for (int d = 0; d < 3; ++d)
            for (int i = 0; i < cnt[d]; ++i) hc[o++] = make_float2((float)cos(src[d][i]), (float)sin(src[d][i]));
    }
    o = 0;
    for (int i = 0; i < nz; ++i) hf[o++] = (float)wz[i];
    if (has_comp) {
        const double *src[3] = {comp_a, comp_b, comp_z};
        const int cnt[3] = {n_a, n_b, nz};
        for (int d = 0; d < 3; ++d)
            for (int i = 0; i < cnt[d]; ++i) hf[o++] = (float)(1.0 / (src[d][i] * src[d][i]));
    }
    const size_t nbins = (size_t)(nedges + 1) * (nmu + 2);
    const size_t cbytes = sizeof(double) * KMU_NCOPY * (3 + 2 * (size_t)nell) * nbins;
    if (cudaMalloc(&K->tables, bytes) != cudaSuccess || cudaMalloc(&K->copies, cbytes) != cudaSuccess) {
        if (K->tables) cudaFree(K->tables);
        delete K; set_error("apk_kmu_create: cudaMalloc(%zu + %zu) failed", bytes, cbytes); return 1;
    }
    if (cudaMemcpy(K->tables, host.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(K->tables); cudaFree(K->copies); delete K; set_error("apk_kmu_create: table upload failed"); return 1;
    }
    double *dd = (double *)K->tables;
    K->ka = dd; K->kb = dd + n_a; K->kz = K->kb + n_b; K->edges2 = K->kz + nz; K->muedges = K->edges2 + nedges;
    float2 *dc = (float2 *)(dd + nd);
    if (has_phase) { K->ph_a = dc; K->ph_b = dc + n_a; K->ph_z = K->ph_b + n_b; }
    float *df = (float *)(dc + nc);
    K->wz = df;
    if (has_comp) { K->ic_a = df + nz; K->ic_b = K->ic_a + n_a; K->ic_z = K->ic_b + n_b; }
    *out = K;
    return 0;
}

extern "C" int apk_kmu_destroy(apk_kmu *K) {
    if (!K) return 0;
    DeviceGuard guard(K->plan->device);
    cudaFree(K->tables);
    cudaFree(K->copies);
    delete K;
    return 0;
}

extern "C" int apk_kmu_bin(apk_kmu *K, const void *c1, const void *c1s, const void *c2, const void *c2s, double *xsum,
                           double *musum, double *ysum_re, double *ysum_im, int64_t *nsum, void *stream) {
    APK_REQUIRE(K && c1 && xsum && musum && ysum_re && ysum_im && nsum, "apk_kmu_bin: null argument");
    APK_REQUIRE((c1s != nullptr) == K->has_phase, "apk_kmu_bin: an interlaced twin needs phase tables (and vice versa)");
    APK_REQUIRE(!c2s || c2, "apk_kmu_bin: c2s without c2");
    APK_REQUIRE(!c2 || ((c2s != nullptr) == (c1s != nullptr)), "apk_kmu_bin: both fields interlaced or neither");
    DeviceGuard guard(K->plan->device);
    cudaStream_t st = (cudaStream_t)stream;
    KmuArgs A;
    A.c1 = (const float2 *)c1; A.c1s = (const float2 *)c1s; A.c2 = (const float2 *)c2; A.c2s = (const float2 *)c2s;
    A.ka = K->ka; A.kb = K->kb; A.kz = K->kz; A.edges2 = K->edges2; A.muedges = K->muedges;
    A.wz = K->wz; A.ic_a = K->ic_a; A.ic_b = K->ic_b; A.ic_z = K->ic_z;
    A.ph_a = K->ph_a; A.ph_b = K->ph_b; A.ph_z = K->ph_z;
    A.n_a = K->n_a; A.n_b = K->n_b; A.nz = K->nz; A.nedges = K->nedges; A.nmu = K->nmu; A.nell = K->nell;
    for (int i = 0; i < 8; ++i) A.ells[i] = K->ells[i];
    for (int d = 0; d < 3; ++d) A.los[d] = K->los[d];
    A.dc_a = K->dc_a; A.dc_b = K->dc_b;
    A.kmin_f = (float)K->kmin_guess; A.inv_dk_f = (float)K->inv_dk_guess;
    A.copies = K->copies;
    A.nbins = (long long)(K->nedges + 1) * (K->nmu + 2);
    const size_t cbytes = sizeof(double) * KMU_NCOPY * (3 + 2 * (size_t)K->nell) * (size_t)A.nbins;
    APK_CUDA(cudaMemsetAsync(K->copies, 0, cbytes, st));
    const int blocks = K->plan->num_sms * 8;
    const bool inter = c1s != nullptr, cross = c2 != nullptr, comp = K->has_comp;
#define APK_KMU_GO(I, C, M) bin_kmu_kernel<I, C, M><<<blocks, KMU_THREADS, 0, st>>>(A)
    if (inter) { if (cross) { if (comp) APK_KMU_GO(true, true, true); else APK_KMU_GO(true, true, false); }
                 else { if (comp) APK_KMU_GO(true, false, true); else APK_KMU_GO(true, false, false); } }
    else { if (cross) { if (comp) APK_KMU_GO(false, true, true); else APK_KMU_GO(false, true, false); }
           else { if (comp) APK_KMU_GO(false, false, true); else APK_KMU_GO(false, false, false); } }
#undef APK_KMU_GO
    APK_CUDA(cudaGetLastError());
    kmu_fold_kernel<<<(unsigned)((A.nbins + 255) / 256), 256, 0, st>>>(K->copies, K->nell, A.nbins, K->nmu, xsum, musum,
                                                                         (long long *)nsum, ysum_re, ysum_im);
    APK_CUDA(cudaGetLastError());
    return 0;
}
