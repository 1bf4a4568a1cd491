// Runs the SOURCE of astrild_b200/csrc/bin_kmu.cu's kernels on the CPU (tests/simt/simt.h): table set-up as in
// apk_kmu_create, then bin_kmu_kernel + kmu_fold_kernel.  Tests only.  bin_kmu_kernels.inc is produced by
// tests/simt/build_simt.py from the .cu (device part; the two inline-PTX RED helpers become plain adds).
#include "simt.h"
#include "bin_kmu_kernels.inc"

#include <cmath>
#include <vector>

using namespace apk;

extern "C" int simt_bin_kmu(const void *c1, const void *c1s, const void *c2, const void *c2s, int n_a, int n_b, int nz,
                            const double *ka, const double *kb, const double *kz, const double *wz, const double *kedges,
                            int nedges, int nmu, const int *ells, int nell, const double *los, const double *comp_a,
                            const double *comp_b, const double *comp_z, const double *phase_a, const double *phase_b,
                            const double *phase_z, int dc_a, int dc_b, int ctas, double *xsum, double *musum,
                            double *ysum_re, double *ysum_im, long long *nsum) {
    const bool has_comp = comp_a != nullptr, has_phase = phase_a != nullptr;
    const bool interlaced = c1s != nullptr, cross = c2 != nullptr;
    if (interlaced != has_phase) return 2;
    std::vector<double> e2(nedges), mue(nmu + 1);
    for (int i = 0; i < nedges; ++i) e2[i] = kedges[i] * kedges[i];
    for (int i = 0; i <= nmu; ++i) mue[i] = i == nmu ? 1.0 : (double)i * (1.0 / (double)nmu);
    std::vector<float> wzf(nz), ica, icb, icz;
    for (int i = 0; i < nz; ++i) wzf[i] = (float)wz[i];
    std::vector<float2> pa, pb, pz;
    if (has_phase) {
        for (int i = 0; i < n_a; ++i) pa.push_back(make_float2((float)std::cos(phase_a[i]), (float)std::sin(phase_a[i])));
        for (int i = 0; i < n_b; ++i) pb.push_back(make_float2((float)std::cos(phase_b[i]), (float)std::sin(phase_b[i])));
        for (int i = 0; i < nz; ++i) pz.push_back(make_float2((float)std::cos(phase_z[i]), (float)std::sin(phase_z[i])));
    }
    if (has_comp) {
        for (int i = 0; i < n_a; ++i) ica.push_back((float)(1.0 / (comp_a[i] * comp_a[i])));
        for (int i = 0; i < n_b; ++i) icb.push_back((float)(1.0 / (comp_b[i] * comp_b[i])));
        for (int i = 0; i < nz; ++i) icz.push_back((float)(1.0 / (comp_z[i] * comp_z[i])));
    }
    KmuArgs A;
    A.c1 = (const float2 *)c1; A.c1s = (const float2 *)c1s; A.c2 = (const float2 *)c2; A.c2s = (const float2 *)c2s;
    A.ka = ka; A.kb = kb; A.kz = kz; A.edges2 = e2.data(); A.muedges = mue.data(); A.wz = wzf.data();
    A.ic_a = has_comp ? ica.data() : nullptr; A.ic_b = has_comp ? icb.data() : nullptr; A.ic_z = has_comp ? icz.data() : nullptr;
    A.ph_a = has_phase ? pa.data() : nullptr; A.ph_b = has_phase ? pb.data() : nullptr; A.ph_z = has_phase ? pz.data() : nullptr;
    A.n_a = n_a; A.n_b = n_b; A.nz = nz; A.nedges = nedges; A.nmu = nmu; A.nell = nell;
    for (int i = 0; i < 8; ++i) A.ells[i] = i < nell ? ells[i] : 0;
    for (int d = 0; d < 3; ++d) A.los[d] = los[d];
    A.dc_a = dc_a; A.dc_b = dc_b;
    A.kmin_f = (float)kedges[0]; A.inv_dk_f = (float)(1.0 / (kedges[1] - kedges[0]));
    A.nbins = (long long)(nedges + 1) * (nmu + 2);
    std::vector<double> copies((size_t)KMU_NCOPY * (3 + 2 * nell) * A.nbins, 0.0);
    A.copies = copies.data();
    const bool comp = has_comp;
#define GO(I, C, M) simt::launch(ctas, KMU_THREADS, [&] { bin_kmu_kernel<I, C, M>(A); })
    if (interlaced) { if (cross) { if (comp) GO(true, true, true); else GO(true, true, false); }
                      else { if (comp) GO(true, false, true); else GO(true, false, false); } }
    else { if (cross) { if (comp) GO(false, true, true); else GO(false, true, false); }
           else { if (comp) GO(false, false, true); else GO(false, false, false); } }
#undef GO
    simt::launch((int)((A.nbins + 255) / 256), 256, [&] {
        kmu_fold_kernel(copies.data(), nell, A.nbins, nmu, xsum, musum, nsum, ysum_re, ysum_im);
    });
    return 0;
}
