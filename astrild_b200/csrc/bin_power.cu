// Fused shell binning:  interlace-combine + window deconvolution + c1 conj(c2) + Hermitian
// weights + numpy.digitize(k^2, kedges^2) + per-shell sums, reading every k-grid ONCE.
//
// Replaces nbodykit's _compute_3d_power + project_to_basis (five NumPy passes over the grid)
// behind FFTPower(mode="1d"), as astrild calls it at
//   /root/reference/src/astrild/power_spectra/power_spectrum_3d.py:189-195, 216-222
//   /root/reference/src/astrild/particles/hutils/stats_subfind.py:142-148
//
// Bound: HBM.  Algorithmic bytes per launch = 8 * n_a*n_b*nz * n_grids (each complex64 once).
//
// Mapping.  The grid is complex64 [n_a][n_b][nz], z fastest.  A warp owns a work item
// (TA consecutive a-rows) x (32 consecutive iz, one per lane) x (a segment of b); it walks b,
// so every load instruction of the warp reads 256 contiguous bytes.  k^2 = (ka2+kb2)+kz2 is
// evaluated in float64 from host-built tables, and the bin comes from a float guess that is
// then FIXED against kedges^2 with float64 compares -- bit-identical to numpy.digitize.
// Along a thread's walk |k| moves by at most one fundamental per step, so the thread keeps a
// private direct-mapped window of W shells in shared memory (plain load/add/store, no atomics:
// shared-memory f64/f32 atomics are CAS loops on sm_100) and only evicts a shell when a new one
// maps to its slot.  Evictions go to a per-CTA private histogram in global memory with
// fire-and-forget RED.ADD.F64 (native at L2), and a last kernel folds the per-CTA copies in a
// fixed order.
#include "apk_common.cuh"

namespace apk {

constexpr int BIN_THREADS = 256;
constexpr int BIN_W = 8;    // window slots per thread (power of two)

struct BinArgs {
    const float2 *c1, *c1s, *c2, *c2s;
    const double *ka2, *kb2, *kz2, *edges2;
    const float *wz;
    const float *ic_a, *ic_b, *ic_z;
    const float2 *ph_a, *ph_b, *ph_z;
    int n_a, n_b, nz, nedges;
    int dc_a, dc_b;
    int seg_b;        // b-rows per work item
    int n_ga, n_sb, n_zc;
    float kmin_f, inv_dk_f;
    double *part_k, *part_p, *part_pim;   // [ctas][nedges+1]
    unsigned long long *part_n;           // [ctas][nedges+1]
    unsigned short *bins;                 // [n_a][n_b][nz] shell of every mode (MODE 1 writes it, MODE 2 reads it)
};

// shells are < 65534; the two largest codes mark modes outside the edges
constexpr unsigned short BIN_UNDER = 0xfffe, BIN_OVER = 0xffff;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ void red_add_f64(double *addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *addr, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// a-rows per work item: 4 for one grid, 2 when 2-4 grids are read per mode (register budget)
template <bool INTERLACED, bool CROSS>
struct BinRows { static constexpr int TA = (INTERLACED || CROSS) ? 2 : 4; };

// MODE 0: everything in one pass (APK_BIN_TABLE=0).  Default split: the shell of a mode, the mode
// counts and sum(w k) depend on the binning alone, so MODE 1 computes them once per binning object -- same float64
// digitize, no grid read -- and stores the shell of every mode as uint16; MODE 2, run per spectrum, reads that
// table instead of doing any float64 wavenumber arithmetic and only accumulates sum(w P).
template <bool INTERLACED, bool CROSS, bool COMP, int MODE = 0>
__global__ void __launch_bounds__(BIN_THREADS, 3)
bin_power_kernel(BinArgs A) {
    constexpr int BIN_TA = BinRows<INTERLACED, CROSS>::TA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: edges2[nedges] | ps[W][T] | ks[W][T] | meta[W][T]
    double *s_e2 = reinterpret_cast<double *>(smem_raw);
    const int e_pad = (A.nedges + 1) & ~1;
    double *s_ps = s_e2 + e_pad;
    double *s_ks = s_ps + BIN_W * BIN_THREADS;
    int2 *s_meta = reinterpret_cast<int2 *>(s_ks + BIN_W * BIN_THREADS);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int i = tid; i < A.nedges; i += BIN_THREADS) s_e2[i] = A.edges2[i];
#pragma unroll
    for (int s = 0; s < BIN_W; ++s) s_meta[s * BIN_THREADS + tid] = make_int2(-1, 0);
    __syncthreads();

    const int nb1 = A.nedges + 1;
    double *g_k = A.part_k + (size_t)blockIdx.x * nb1;
    double *g_p = A.part_p + (size_t)blockIdx.x * nb1;
    double *g_pim = A.part_pim + (size_t)blockIdx.x * nb1;
    unsigned long long *g_n = A.part_n + (size_t)blockIdx.x * nb1;

    unsigned int under_cnt = 0, over_cnt = 0;

    const long long warps_total = (long long)gridDim.x * (BIN_THREADS / 32);
    const long long warp_id = (long long)blockIdx.x * (BIN_THREADS / 32) + (tid >> 5);
    const long long n_items = (long long)A.n_ga * A.n_sb * A.n_zc;
    const int nedges = A.nedges;
    const double e2_last = s_e2[nedges - 1];
    const double e2_first = s_e2[0];

    for (long long item = warp_id; item < n_items; item += warps_total) {
        const int zc = (int)(item % A.n_zc);
        const long long r = item / A.n_zc;
        const int sb = (int)(r % A.n_sb);
        const int ga = (int)(r / A.n_sb);
        const int iz = zc * 32 + lane;
        const bool zvalid = iz < A.nz;
        const int izc = zvalid ? iz : A.nz - 1;
        const int ia0 = ga * BIN_TA;
        const int ib0 = sb * A.seg_b;
        const int ib1 = min(ib0 + A.seg_b, A.n_b);

        const double kz2 = A.kz2[izc];
        const float wz = A.wz[izc];
        const bool singular = wz < 1.5f;
        float icz = 1.f;
        float2 phz = make_float2(1.f, 0.f);
        if (COMP) icz = A.ic_z[izc];
        if (INTERLACED) phz = A.ph_z[izc];

        double ka2[BIN_TA];
        float ica[BIN_TA];
        float2 pha[BIN_TA];
        bool avalid[BIN_TA];
        size_t base[BIN_TA];
#pragma unroll
        for (int t = 0; t < BIN_TA; ++t) {
            const int ia = ia0 + t;
            avalid[t] = ia < A.n_a;
            const int iac = avalid[t] ? ia : A.n_a - 1;
            ka2[t] = A.ka2[iac];
            if (COMP) ica[t] = A.ic_a[iac] * icz;
            if (INTERLACED) pha[t] = cmul(A.ph_a[iac], phz);
            base[t] = ((size_t)iac * A.n_b) * A.nz + izc;
        }

        bool dc_row[BIN_TA];
#pragma unroll
        for (int t = 0; t < BIN_TA; ++t) dc_row[t] = iz == 0 && ia0 + t == A.dc_a;
        float2 v1[BIN_TA], v1s[BIN_TA], v2[BIN_TA], v2s[BIN_TA];
        unsigned short vb[BIN_TA];               // MODE 2: the stored shells of the prefetched row
        double n_kb2 = 0.0;                      // b-row tables travel with the prefetched row
        float n_icb = 1.f;
        float2 n_phb = make_float2(1.f, 0.f);
        // smallest kz^2 of this warp's 32 modes (kz^2 grows with iz): if even that one is beyond the last
        // edge, the whole 256-byte line is overflow and is not read at all (for kmax = Nyquist ~40 % of
        // the grid lies outside the sphere)
        const double kz2_min = __shfl_sync(0xffffffffu, kz2, 0);
        double pre_kb2 = A.kb2[min(ib0, A.n_b - 1)];     // kb2 runs one more row ahead: the skip test must not wait for it
        auto load_row = [&](int ib) {
            n_kb2 = pre_kb2;
            pre_kb2 = A.kb2[min(ib + 1, A.n_b - 1)];
            if (COMP) n_icb = A.ic_b[ib];
            if (INTERLACED) n_phb = A.ph_b[ib];
#pragma unroll
            for (int t = 0; t < BIN_TA; ++t) {
                if ((ka2[t] + n_kb2) + kz2_min >= e2_last) continue;      // warp-uniform
                const size_t idx = base[t] + (size_t)ib * A.nz;
                if (MODE == 2) vb[t] = A.bins[idx];
                if (MODE == 1) continue;
                v1[t] = __ldcs(A.c1 + idx);
                if (INTERLACED) v1s[t] = __ldcs(A.c1s + idx);
                if (CROSS) {
                    v2[t] = __ldcs(A.c2 + idx);
                    if (INTERLACED) v2s[t] = __ldcs(A.c2s + idx);
                }
            }
        };

        if (ib0 < ib1) load_row(ib0);
        for (int ib = ib0; ib < ib1; ++ib) {
            float2 c1[BIN_TA], c1s[BIN_TA], c2[BIN_TA], c2s[BIN_TA];
            unsigned short cb[BIN_TA];
#pragma unroll
            for (int t = 0; t < BIN_TA; ++t) {
                cb[t] = vb[t];
                c1[t] = v1[t];
                if (INTERLACED) c1s[t] = v1s[t];
                if (CROSS) { c2[t] = v2[t]; if (INTERLACED) c2s[t] = v2s[t]; }
            }
            const double kb2 = n_kb2;
            const float icb = n_icb;
            const float2 phb = n_phb;
            if (ib + 1 < ib1) load_row(ib + 1);   // software prefetch of the next b-row

#pragma unroll
            for (int t = 0; t < BIN_TA; ++t) {
                if (!(zvalid && avalid[t])) continue;
                const unsigned int wi = singular ? 1u : 2u;
                int bin;
                double kk = 0.0;
                if (MODE == 2) {
                    // lines beyond the last edge were skipped by load_row (cb is stale there): same test here
                    if ((ka2[t] + kb2) + kz2_min >= e2_last) continue;
                    if (cb[t] >= BIN_UNDER) continue;
                    bin = cb[t];
                } else {
                    const double k2 = (ka2[t] + kb2) + kz2;
                    const size_t idx = base[t] + (size_t)ib * A.nz;
                    if (k2 >= e2_last) { over_cnt += wi; if (MODE == 1 && (ka2[t] + kb2) + kz2_min < e2_last) A.bins[idx] = BIN_OVER; continue; }
                    if (k2 < e2_first) { under_cnt += wi; if (MODE == 1) A.bins[idx] = BIN_UNDER; continue; }
                    // --- bin: float guess, exact float64 fix-up against kedges^2 -----------------
                    // sqrt(k2): float rsqrt seed + one Newton step in float64 (rel. error ~1e-14)
                    const float k2f = (float)k2;
                    const float rs = k2f > 0.f ? rsqrtf(k2f) : 0.f;
                    const double y = (double)rs;
                    kk = k2 * y;
                    kk = fma(0.5 * y, fma(-kk, kk, k2), kk);
                    bin = (int)((k2f * rs - A.kmin_f) * A.inv_dk_f) + 1;
                    bin = max(1, min(bin, nedges - 1));
                    while (k2 < s_e2[bin - 1]) --bin;
                    while (k2 >= s_e2[bin]) ++bin;
                    if (MODE == 1) A.bins[idx] = (unsigned short)bin;
                }
                // --- power of this mode --------------------------------------------------------
                float pre = 0.f, pim = 0.f;
                if (MODE != 1) {
                    float2 a = c1[t];
                    if (INTERLACED) {
                        const float2 ph = cmul(pha[t], phb);
                        const float2 s = cmul(c1s[t], ph);
                        a = make_float2(0.5f * (a.x + s.x), 0.5f * (a.y + s.y));
                    }
                    if (CROSS) {
                        float2 b = c2[t];
                        if (INTERLACED) {
                            const float2 ph = cmul(pha[t], phb);
                            const float2 s = cmul(c2s[t], ph);
                            b = make_float2(0.5f * (b.x + s.x), 0.5f * (b.y + s.y));
                        }
                        pre = a.x * b.x + a.y * b.y;
                        pim = a.y * b.x - a.x * b.y;
                    } else {
                        pre = a.x * a.x + a.y * a.y;
                        pim = 0.f;
                    }
                    if (COMP) { const float ic = ica[t] * icb; pre *= ic; pim *= ic; }
                    if (dc_row[t] && ib == A.dc_b) { pre = 0.f; pim = 0.f; }
                }
                const double w = (double)wi;
                if (MODE != 1 && CROSS && singular && pim != 0.f) red_add_f64(g_pim + bin, (double)pim);
                // --- private window update ---------------------------------------------------
                const int slot = (bin & (BIN_W - 1)) * BIN_THREADS + tid;
                int2 m = s_meta[slot];
                double ps, ks;
                if (m.x != bin) {
                    if (m.x >= 0) {
                        if (MODE != 1) red_add_f64(g_p + m.x, s_ps[slot]);
                        if (MODE != 2) red_add_f64(g_k + m.x, s_ks[slot]);
                        if (MODE != 2) red_add_u64(g_n + m.x, (unsigned long long)m.y);
                    }
                    m = make_int2(bin, 0);
                    ps = 0.0;
                    ks = 0.0;
                } else {
                    ps = MODE != 1 ? s_ps[slot] : 0.0;
                    ks = MODE != 2 ? s_ks[slot] : 0.0;
                }
                m.y += (int)wi;
                s_meta[slot] = m;
                if (MODE != 1) s_ps[slot] = fma(w, (double)pre, ps);
                if (MODE != 2) s_ks[slot] = fma(w, kk, ks);
            }
        }
    }

    // drain the window
#pragma unroll
    for (int s = 0; s < BIN_W; ++s) {
        const int slot = s * BIN_THREADS + tid;
        const int2 m = s_meta[slot];
        if (m.x >= 0) {
            if (MODE != 1) red_add_f64(g_p + m.x, s_ps[slot]);
            if (MODE != 2) red_add_f64(g_k + m.x, s_ks[slot]);
            if (MODE != 2) red_add_u64(g_n + m.x, (unsigned long long)m.y);
        }
    }
    if (MODE == 2) return;
    // under/overflow counts: warp-reduce, one RED per warp
    for (int o = 16; o > 0; o >>= 1) {
        under_cnt += __shfl_xor_sync(0xffffffffu, under_cnt, o);
        over_cnt += __shfl_xor_sync(0xffffffffu, over_cnt, o);
    }
    if (lane == 0) {
        if (under_cnt) red_add_u64(g_n + 0, under_cnt);
        if (over_cnt) red_add_u64(g_n + nedges, over_cnt);
    }
}

// fold the per-CTA copies in a fixed order (deterministic given the copies): one warp per bin
__global__ void __launch_bounds__(128)
bin_fold_kernel(const double *part_k, const double *part_p, const double *part_pim,
                const unsigned long long *part_n, int ctas, int nb1, double *ksum,
                double *psum_re, double *psum_im, long long *nmodes) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nb1) return;
    double k = 0.0, p = 0.0, q = 0.0;
    unsigned long long n = 0;
    for (int c = lane; c < ctas; c += 32) {
        k += part_k[(size_t)c * nb1 + b];
        p += part_p[(size_t)c * nb1 + b];
        q += part_pim[(size_t)c * nb1 + b];
        n += part_n[(size_t)c * nb1 + b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k += __shfl_xor_sync(0xffffffffu, k, o);
        p += __shfl_xor_sync(0xffffffffu, p, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if (lane == 0) {
        ksum[b] = k;
        psum_re[b] = p;
        psum_im[b] = q;
        nmodes[b] = (long long)n;
    }
}

size_t bin_smem_bytes(int nedges) {
    const int e_pad = (nedges + 1) & ~1;
    return sizeof(double) * (e_pad + 2 * BIN_W * BIN_THREADS) + sizeof(int2) * BIN_W * BIN_THREADS;
}

template <bool I, bool C, bool P, int MODE>
static int launch_bin(const BinArgs &A, int ctas, size_t smem, cudaStream_t st) {
    auto kern = bin_power_kernel<I, C, P, MODE>;
    APK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, BIN_THREADS, smem, st>>>(A);
    APK_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_bin_variant(bool interlaced, bool cross, bool comp, const BinArgs &A, int ctas, size_t smem, cudaStream_t st) {
    if (interlaced) {
        if (cross) return comp ? launch_bin<true, true, true, MODE>(A, ctas, smem, st) : launch_bin<true, true, false, MODE>(A, ctas, smem, st);
        return comp ? launch_bin<true, false, true, MODE>(A, ctas, smem, st) : launch_bin<true, false, false, MODE>(A, ctas, smem, st);
    }
    if (cross) return comp ? launch_bin<false, true, true, MODE>(A, ctas, smem, st) : launch_bin<false, true, false, MODE>(A, ctas, smem, st);
    return comp ? launch_bin<false, false, true, MODE>(A, ctas, smem, st) : launch_bin<false, false, false, MODE>(A, ctas, smem, st);
}

int bin_power_launch(apk_binning *B, const void *c1, const void *c1s, const void *c2,
                     const void *c2s, double *ksum, double *psum_re, double *psum_im,
                     int64_t *nmodes, cudaStream_t st) {
    const bool interlaced = c1s != nullptr;
    const bool cross = c2 != nullptr;
    APK_REQUIRE(!interlaced || B->has_phase, "apk_bin_power: interlaced twin given but the binning has no phase tables");
    APK_REQUIRE(!cross || (interlaced == (c2s != nullptr)), "apk_bin_power: c2s must be given iff c1s is");
    BinArgs A;
    A.c1 = (const float2 *)c1; A.c1s = (const float2 *)c1s; A.c2 = (const float2 *)c2; A.c2s = (const float2 *)c2s;
    A.ka2 = B->ka2; A.kb2 = B->kb2; A.kz2 = B->kz2; A.edges2 = B->edges2; A.wz = B->wz;
    A.ic_a = B->icomp2_a; A.ic_b = B->icomp2_b; A.ic_z = B->icomp2_z;
    A.ph_a = B->ph_a; A.ph_b = B->ph_b; A.ph_z = B->ph_z;
    A.n_a = B->n_a; A.n_b = B->n_b; A.nz = B->nz; A.nedges = B->nedges;
    A.dc_a = B->dc_a; A.dc_b = B->dc_b;
    A.kmin_f = (float)B->kmin_guess; A.inv_dk_f = (float)B->inv_dk_guess;
    A.bins = B->bins;

    const int ctas = B->partial_ctas;
    const int nb1 = B->nedges + 1;
    const int ta = (interlaced || cross) ? 2 : 4;      // BinRows<>::TA
    A.n_ga = (B->n_a + ta - 1) / ta;
    A.n_zc = (B->nz + 31) / 32;
    // choose the b-segment so that there are several work items per warp
    const long long warps = (long long)ctas * (BIN_THREADS / 32);
    int seg = B->n_b;
    while (seg > 32 && (long long)A.n_ga * A.n_zc * ((B->n_b + seg - 1) / seg) < 4 * warps) seg >>= 1;
    A.seg_b = seg;
    A.n_sb = (B->n_b + seg - 1) / seg;
    A.part_k = B->partial;
    A.part_p = A.part_k + (size_t)ctas * nb1;
    A.part_pim = A.part_p + (size_t)ctas * nb1;
    A.part_n = (unsigned long long *)(A.part_pim + (size_t)ctas * nb1);
    const size_t pbytes = sizeof(double) * 4 * (size_t)ctas * nb1;
    const size_t smem = bin_smem_bytes(B->nedges);
    const bool comp = B->has_comp;

    // geometry once per binning object (the binning plan), then table-driven data passes
    const bool tabled = B->use_table;
    if (tabled && !B->bins) {
        const size_t modes = (size_t)B->n_a * B->n_b * B->nz;
        APK_CUDA(cudaMalloc(&B->bins, modes * sizeof(unsigned short)));
        APK_CUDA(cudaMalloc(&B->geo, sizeof(double) * 4 * (size_t)nb1));
        A.bins = B->bins;
        APK_CUDA(cudaMemsetAsync(B->partial, 0, pbytes, st));
        // the geometry does not depend on which grids are given: one (auto, single grid) work distribution
        BinArgs Ag = A;
        Ag.c1s = Ag.c2 = Ag.c2s = nullptr;
        Ag.n_ga = (B->n_a + 3) / 4;
        int sg = B->n_b;
        while (sg > 32 && (long long)Ag.n_ga * A.n_zc * ((B->n_b + sg - 1) / sg) < 4 * warps) sg >>= 1;
        Ag.seg_b = sg;
        Ag.n_sb = (B->n_b + sg - 1) / sg;
        if (int rc = launch_bin<false, false, false, 1>(Ag, ctas, smem, st)) return rc;
        double *g = B->geo;
        bin_fold_kernel<<<(nb1 + 3) / 4, 128, 0, st>>>(A.part_k, A.part_p, A.part_pim, A.part_n, ctas, nb1,
                                                          g, g + nb1, g + 2 * nb1, (long long *)(g + 3 * nb1));
        APK_CUDA(cudaGetLastError());
    }

    APK_CUDA(cudaMemsetAsync(B->partial, 0, pbytes, st));
    const bool timing = B->plan->timing;
    if (timing && !B->ev_ready) {
        for (auto &e : B->ev) APK_CUDA(cudaEventCreate(&e));
        B->ev_ready = true;
    }
    if (timing) APK_CUDA(cudaEventRecord(B->ev[0], st));
    int rc = tabled ? launch_bin_variant<2>(interlaced, cross, comp, A, ctas, smem, st)
                    : launch_bin_variant<0>(interlaced, cross, comp, A, ctas, smem, st);
    if (rc) return rc;
    if (timing) APK_CUDA(cudaEventRecord(B->ev[1], st));
    bin_fold_kernel<<<(nb1 + 3) / 4, 128, 0, st>>>(A.part_k, A.part_p, A.part_pim, A.part_n, ctas, nb1,
                                                      ksum, psum_re, psum_im, (long long *)nmodes);
    APK_CUDA(cudaGetLastError());
    if (tabled) {        // mode counts and sum(w k) come from the geometry pass
        APK_CUDA(cudaMemcpyAsync(ksum, B->geo, sizeof(double) * nb1, cudaMemcpyDeviceToDevice, st));
        APK_CUDA(cudaMemcpyAsync(nmodes, B->geo + 3 * (size_t)nb1, sizeof(long long) * nb1, cudaMemcpyDeviceToDevice, st));
    }
    if (timing) APK_CUDA(cudaEventRecord(B->ev[2], st));
    B->timed = timing;
    return 0;
}

}  // namespace apk
