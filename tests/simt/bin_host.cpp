// Runs the SOURCE of astrild_b200/csrc/bin_power.cu's kernels on the CPU (tests/simt/simt.h): table set-up as in
// apk_binning_create, launch parameters as in bin_power_launch, then bin_power_kernel + bin_fold_kernel.  Tests only.
// bin_power_kernels.inc is produced by tests/simt/build_simt.py from the .cu (device part; the two inline-PTX RED
// helpers become plain adds).
#include "simt.h"
#include "bin_power_kernels.inc"

#include <cmath>
#include <vector>

using namespace apk;

namespace {
template <bool I, bool C, bool P, int MODE>
void go(const BinArgs &A, int ctas) {
    simt::launch(ctas, BIN_THREADS, [&] { bin_power_kernel<I, C, P, MODE>(A); });
}
template <int MODE>
void go_variant(bool interlaced, bool cross, bool comp, const BinArgs &A, int ctas) {
    if (interlaced) {
        if (cross) comp ? go<true, true, true, MODE>(A, ctas) : go<true, true, false, MODE>(A, ctas);
        else comp ? go<true, false, true, MODE>(A, ctas) : go<true, false, false, MODE>(A, ctas);
    } else {
        if (cross) comp ? go<false, true, true, MODE>(A, ctas) : go<false, true, false, MODE>(A, ctas);
        else comp ? go<false, false, true, MODE>(A, ctas) : go<false, false, false, MODE>(A, ctas);
    }
}
}  // namespace

extern "C" int simt_bin_power(const void *c1, const void *c1s, const void *c2, const void *c2s, int n_a, int n_b, int nz,
                              const double *ka, const double *kb, const double *kz, const double *wz,
                              const double *kedges, int nedges, const double *comp_a, const double *comp_b,
                              const double *comp_z, const double *phase_a, const double *phase_b, const double *phase_z,
                              int dc_a, int dc_b, int ctas, double *ksum, double *psum_re, double *psum_im,
                              long long *nmodes) {
    const bool has_comp = comp_a != nullptr, has_phase = phase_a != nullptr;
    const bool interlaced = c1s != nullptr, cross = c2 != nullptr;
    if (interlaced && !has_phase) return 2;
    std::vector<double> ka2(n_a), kb2(n_b), kz2(nz), e2(nedges);
    for (int i = 0; i < n_a; ++i) ka2[i] = ka[i] * ka[i];
    for (int i = 0; i < n_b; ++i) kb2[i] = kb[i] * kb[i];
    for (int i = 0; i < nz; ++i) kz2[i] = kz[i] * kz[i];
    for (int i = 0; i < nedges; ++i) e2[i] = kedges[i] * kedges[i];
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
    BinArgs A;
    A.c1 = (const float2 *)c1; A.c1s = (const float2 *)c1s; A.c2 = (const float2 *)c2; A.c2s = (const float2 *)c2s;
    A.ka2 = ka2.data(); A.kb2 = kb2.data(); A.kz2 = kz2.data(); A.edges2 = e2.data(); A.wz = wzf.data();
    A.ic_a = has_comp ? ica.data() : nullptr; A.ic_b = has_comp ? icb.data() : nullptr; A.ic_z = has_comp ? icz.data() : nullptr;
    A.ph_a = has_phase ? pa.data() : nullptr; A.ph_b = has_phase ? pb.data() : nullptr; A.ph_z = has_phase ? pz.data() : nullptr;
    A.n_a = n_a; A.n_b = n_b; A.nz = nz; A.nedges = nedges; A.dc_a = dc_a; A.dc_b = dc_b;
    A.kmin_f = (float)kedges[0]; A.inv_dk_f = (float)(1.0 / (kedges[1] - kedges[0]));
    const int nb1 = nedges + 1;
    const int ta = (interlaced || cross) ? 2 : 4;
    A.n_ga = (n_a + ta - 1) / ta;
    A.n_zc = (nz + 31) / 32;
    const long long warps = (long long)(ctas & 0xffff) * (BIN_THREADS / 32);
    int seg = n_b;
    while (seg > 32 && (long long)A.n_ga * A.n_zc * ((n_b + seg - 1) / seg) < 4 * warps) seg >>= 1;
    A.seg_b = seg;
    A.n_sb = (n_b + seg - 1) / seg;
    const int nct = ctas & 0xffff;
    std::vector<double> partial(4 * (size_t)nct * nb1, 0.0);
    A.part_k = partial.data();
    A.part_p = A.part_k + (size_t)nct * nb1;
    A.part_pim = A.part_p + (size_t)nct * nb1;
    A.part_n = (unsigned long long *)(A.part_pim + (size_t)nct * nb1);
    const bool comp = has_comp;
    const bool tabled = (ctas >> 16) != 0;       // high half of `ctas`: experimental table-driven variant
    ctas &= 0xffff;
    std::vector<unsigned short> bins;
    std::vector<double> geo(4 * (size_t)nb1, 0.0);
    if (tabled) {
        bins.assign((size_t)n_a * n_b * nz, 0x1234);
        A.bins = bins.data();
        BinArgs Ag = A;
        Ag.c1s = Ag.c2 = Ag.c2s = nullptr;
        Ag.n_ga = (n_a + 3) / 4;
        int sg = n_b;
        while (sg > 32 && (long long)Ag.n_ga * A.n_zc * ((n_b + sg - 1) / sg) < 4 * warps) sg >>= 1;
        Ag.seg_b = sg;
        Ag.n_sb = (n_b + sg - 1) / sg;
        go<false, false, false, 1>(Ag, ctas);
        simt::launch((nb1 + 3) / 4, 128, [&] {
            bin_fold_kernel(A.part_k, A.part_p, A.part_pim, A.part_n, ctas, nb1, geo.data(), geo.data() + nb1,
                            geo.data() + 2 * nb1, (long long *)(geo.data() + 3 * nb1));
        });
        std::fill(partial.begin(), partial.end(), 0.0);
        go_variant<2>(interlaced, cross, comp, A, ctas);
    } else {
        A.bins = nullptr;
        go_variant<0>(interlaced, cross, comp, A, ctas);
    }
    simt::launch((nb1 + 3) / 4, 128, [&] {
        bin_fold_kernel(A.part_k, A.part_p, A.part_pim, A.part_n, ctas, nb1, ksum, psum_re, psum_im, nmodes);
    });
    if (tabled) {
        std::memcpy(ksum, geo.data(), sizeof(double) * nb1);
        std::memcpy(nmodes, geo.data() + 3 * (size_t)nb1, sizeof(long long) * nb1);
    }
    return 0;
}
