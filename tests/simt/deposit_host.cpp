// Runs the SOURCE of astrild_b200/csrc/deposit_sorted.cu's kernels on the CPU (tests/simt/simt.h) with the launch
// sequence of run_sorted(): count -> segment sums -> scan -> scatter -> one tile kernel per mesh.  Tests only.
// deposit_sorted_kernels.inc is produced by tests/simt/build_simt.py from the .cu (device part, unchanged).
#include "simt.h"
#include "deposit_sorted_kernels.inc"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace apk;

namespace {


DepositGeom make_geom(int N, double pos_scale, double shift, int resampler, int x0, int n0, int ghost_lo, int ghost_hi) {
    DepositGeom G;
    G.N = N; G.ldz = 2 * (N / 2 + 1); G.scale = pos_scale * (double)N; G.shift = shift;
    G.slab = n0 < N; G.plane0 = x0 - ghost_lo; G.nplanes = G.slab ? ghost_lo + n0 + ghost_hi : N;
    G.own0 = x0; G.nown = n0;
    G.s0 = (float)G.scale;
    G.s1 = (float)(G.scale - (double)G.s0);
    G.s2 = (float)(G.scale - (double)G.s0 - (double)G.s1);
    G.t32 = -1.f;
    if ((shift == 0.0 || shift == 0.5) && std::fabs(G.scale) < 1e30 && std::fabs(G.scale) > 1e-30)
        G.t32 = (float)shift + (resampler == APK_CIC ? 0.f : 0.5f);
    return G;
}

template <int S, typename PT, bool SOA, bool MASS, bool PAIR>
void run(const void *p0, const void *p1, const void *p2, const void *mass, int mass_f64, long long np,
         const DepositGeom &G, float *mesh, float *mesh1, int num_sms) {
    using VT = typename std::conditional<MASS, P4, P3>::type;
    const BrickGrid B = make_brick_grid(G, S, PAIR);
    std::vector<VT> vals((size_t)np + 1);
    std::vector<unsigned int> counts(B.nbricks + 2, 0u), start(B.nbricks + 2, 0xdeadbeefu), cursor(B.nbricks + 2, 0xdeadbeefu),
        filled(B.nbricks + 2, 0xdeadbeefu);
    const int nseg = (B.nbricks + SCAN_SEG - 1) / SCAN_SEG;
    std::vector<unsigned int> seg_total(nseg + 1, 0xdeadbeefu), seg_filled(nseg + 1, 0xdeadbeefu);
    unsigned int counter[2] = {0u, 0xdeadbeefu};

    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    const int pb = (int)std::min<long long>((np + tile - 1) / tile, (long long)num_sms * 8);
    simt::launch(pb, PART_THREADS, [&] {
        brick_count_kernel<S, PT, SOA>((const PT *)p0, (const PT *)p1, (const PT *)p2, np, G, B, counts.data());
    });
    simt::launch(nseg, 1024, [&] { brick_segsum_kernel(counts.data(), B.nbricks, seg_total.data(), seg_filled.data()); });
    simt::launch(nseg, 1024, [&] {
        brick_scan_kernel(counts.data(), B.nbricks, seg_total.data(), seg_filled.data(), start.data(), cursor.data(),
                          filled.data(), counter + 1);
    });
    simt::launch(pb, PART_THREADS, [&] {
        brick_scatter_kernel<S, PT, SOA, MASS, VT>((const PT *)p0, (const PT *)p1, (const PT *)p2, mass, mass_f64, np, G,
                                                  B, cursor.data(), vals.data());
    });
    simt::launch(B.nbricks, TILE_THREADS, [&] {
        brick_tile_kernel<S, MASS, PAIR, 0, VT>(vals.data(), start.data(), filled.data(), counter + 1, G, B, mesh);
    });
    if constexpr (PAIR)
        simt::launch(B.nbricks, TILE_THREADS, [&] {
            brick_tile_kernel<S, MASS, PAIR, 1, VT>(vals.data(), start.data(), filled.data(), counter + 1, G, B, mesh1);
        });
}

template <int S, typename PT, bool SOA>
void dispatch(const void *p0, const void *p1, const void *p2, const void *mass, int mass_f64, long long np,
              const DepositGeom &G, float *mesh, float *mesh1, int num_sms) {
    if (mesh1)
        mass ? run<S, PT, SOA, true, true>(p0, p1, p2, mass, mass_f64, np, G, mesh, mesh1, num_sms)
             : run<S, PT, SOA, false, true>(p0, p1, p2, mass, mass_f64, np, G, mesh, mesh1, num_sms);
    else
        mass ? run<S, PT, SOA, true, false>(p0, p1, p2, mass, mass_f64, np, G, mesh, nullptr, num_sms)
             : run<S, PT, SOA, false, false>(p0, p1, p2, mass, mass_f64, np, G, mesh, nullptr, num_sms);
}

}  // namespace

// resampler: 2 = CIC, 3 = TSC.  mesh (and mesh1 for the interlaced pair) are float32 [nplanes][N][2 (N/2+1)],
// accumulated into.  Slab plans: x0, n0 < N with ghost planes 1 below / 2 above.  Returns the fiber switches.
extern "C" long long simt_deposit_sorted(const void *p0, const void *p1, const void *p2, int soa, int pos_f64,
                                         const void *mass, int mass_f64, long long np, int N, double pos_scale,
                                         double shift, int resampler, int x0, int n0, float *mesh, float *mesh1,
                                         int num_sms) {
    const DepositGeom G = make_geom(N, pos_scale, shift, resampler, x0, n0, 1, 2);
    simt::switches = 0;
    if (np <= 0) return 0;
    const int S = resampler == APK_CIC ? 2 : 3;
#define APK_SIMT_GO(SS, PT, SOA) dispatch<SS, PT, SOA>(p0, p1, p2, mass, mass_f64, np, G, mesh, mesh1, num_sms)
    if (S == 2) {
        if (pos_f64) soa ? APK_SIMT_GO(2, double, true) : APK_SIMT_GO(2, double, false);
        else soa ? APK_SIMT_GO(2, float, true) : APK_SIMT_GO(2, float, false);
    } else {
        if (pos_f64) soa ? APK_SIMT_GO(3, double, true) : APK_SIMT_GO(3, double, false);
        else soa ? APK_SIMT_GO(3, float, true) : APK_SIMT_GO(3, float, false);
    }
    return simt::switches;
}


// brick_keys for every particle (float32 positions, whole-mesh plan): key and brick-local coordinates, for a direct
// check of the float-register index arithmetic at large mesh sizes.  pair: the interlaced pair's brick grid.
// brick edge along x, y as this build has them (APK_BX, APK_BY)
extern "C" void simt_brick_edges(int *bxy) { bxy[0] = BX; bxy[1] = BY; }

extern "C" void simt_brick_keys(const float *xyz, long long np, int N, double pos_scale, int resampler, int pair,
                                unsigned int *key, float *l, int *grid) {
    const DepositGeom G = make_geom(N, pos_scale, 0.0, resampler, 0, N, 1, 2);
    const int S = resampler == APK_CIC ? 2 : 3;
    const BrickGrid B = make_brick_grid(G, S, pair != 0);
    grid[0] = B.nbx; grid[1] = B.nby; grid[2] = B.nbz; grid[3] = B.zcells;
    for (long long p = 0; p < np; ++p) {
        float a[3];
        if (S == 2) brick_keys<2, float>(xyz + 3 * p, G, B, key[p], a);
        else brick_keys<3, float>(xyz + 3 * p, G, B, key[p], a);
        for (int d = 0; d < 3; ++d) l[3 * p + d] = a[d];
    }
}
