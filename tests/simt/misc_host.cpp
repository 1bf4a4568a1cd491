// Runs the SOURCE of the remaining hand-written kernels on the CPU (tests/simt/simt.h): the direct-atomic deposit,
// slab routing, the peer-store transpose, ghost-plane adds and the gridded-field helpers, each with the launch
// sequence of its host launcher.  misc_kernels.inc is produced by tests/simt/build_simt.py (device code only).
#include "simt.h"
#include "misc_kernels.inc"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace apk;

namespace {
DepositGeom make_geom(int N, double pos_scale, double shift, int x0, int n0, int ghost_lo, int ghost_hi) {
    DepositGeom G;
    G.N = N; G.ldz = 2 * (N / 2 + 1); G.scale = pos_scale * (double)N; G.shift = shift;
    G.slab = n0 < N; G.plane0 = x0 - ghost_lo; G.nplanes = G.slab ? ghost_lo + n0 + ghost_hi : N;
    G.own0 = x0; G.nown = n0;
    G.s0 = (float)G.scale;
    G.s1 = (float)(G.scale - (double)G.s0);
    G.s2 = (float)(G.scale - (double)G.s0 - (double)G.s1);
    G.t32 = -1.f;
    return G;
}

template <int S, typename PT>
void atomic_go(const void *p0, const void *p1, const void *p2, int soa, const void *mass, int mf64, long long np,
               const DepositGeom &G, float *mesh, int blocks) {
    if (soa) simt::launch(blocks, 256, [&] { deposit_atomic_kernel<S, PT, true>((const PT *)p0, (const PT *)p1, (const PT *)p2, mass, mf64, np, G, mesh); });
    else simt::launch(blocks, 256, [&] { deposit_atomic_kernel<S, PT, false>((const PT *)p0, nullptr, nullptr, mass, mf64, np, G, mesh); });
}
}  // namespace

// resampler 1 / 2 / 3 = NGP / CIC / TSC
extern "C" int simt_deposit_atomic(const void *p0, const void *p1, const void *p2, int soa, int pos_f64, const void *mass,
                                   int mass_f64, long long np, int N, double pos_scale, double shift, int resampler,
                                   int x0, int n0, float *mesh, int num_sms) {
    if (np <= 0) return 0;
    const DepositGeom G = make_geom(N, pos_scale, shift, x0, n0, 1, 2);
    long long want = (np + 255) / 256;
    const int blocks = (int)(want < (long long)num_sms * 16 ? (want > 0 ? want : 1) : (long long)num_sms * 16);
#define GO(S) (pos_f64 ? atomic_go<S, double>(p0, p1, p2, soa, mass, mass_f64, np, G, mesh, blocks) \
                        : atomic_go<S, float>(p0, p1, p2, soa, mass, mass_f64, np, G, mesh, blocks))
    if (resampler == 1) GO(1); else if (resampler == 2) GO(2); else if (resampler == 3) GO(3); else return 2;
#undef GO
    return 0;
}

// float32 SoA positions (+ optional float32 masses) of rank x0 / (N / nranks): counts[2 * nranks], out_pos[capacity * 3]
extern "C" int simt_route(const float *x, const float *y, const float *z, const float *mass, long long np, int N,
                          double pos_scale, int nranks, int x0, unsigned long long *counts, long long capacity,
                          float *out_pos, float *out_mass, int num_sms) {
    RouteGeom R;
    R.N = N; R.ppr = N / nranks; R.self = x0 / R.ppr; R.x0 = x0;
    R.scale = pos_scale * (double)N;
    R.s0 = (float)R.scale;
    R.s1 = (float)(R.scale - (double)R.s0);
    R.s2 = (float)(R.scale - (double)R.s0 - (double)R.s1);
    unsigned long long *cursor = counts + nranks;
    std::vector<unsigned char> work((size_t)capacity * 16 + 64, 0xff);
    unsigned long long *total = (unsigned long long *)work.data();
    float *stage_pos = (float *)(work.data() + 64);
    float *stage_mass = mass ? stage_pos + 3 * (size_t)capacity : nullptr;
    for (int i = 0; i < nranks; ++i) counts[i] = 0;
    if (np > 0 && capacity > 0) {
        *total = 0;
        const int blocks = (int)std::min<long long>((np + 1023) / 1024, (long long)num_sms * 8);
        simt::launch(blocks, 256, [&] {
            route_stage_kernel<float, true, float>(x, y, z, mass, np, R, counts, total, capacity, stage_pos, stage_mass);
        });
        simt::launch(1, 32, [&] { route_scan_kernel(counts, nranks, cursor); });
        const int gblocks = (int)std::min<long long>((capacity + 255) / 256, (long long)num_sms * 4);
        simt::launch(gblocks, 256, [&] {
            route_group_kernel<float, float>(stage_pos, stage_mass, total, capacity, R, cursor, out_pos, out_mass);
        });
    }
    return 0;
}

// rank `rank` of P stores its [n0][N][nz] complex64 x-slab into the receive buffers recv[s] ([N][ny][nz]) of all ranks
extern "C" int simt_transpose_p2p(const void *grid, void **recv, int rank, int P, int N, int nz, int num_sms) {
    const int n0 = N / P;
    const long long chunk_bytes = (long long)n0 * nz * 8;
    std::vector<unsigned long long> peer(P);
    for (int s = 0; s < P; ++s) peer[s] = (unsigned long long)(uintptr_t)recv[s];
    const int blocks = num_sms * 4;
    if (chunk_bytes % 16 == 0 && ((uintptr_t)grid % 16) == 0)
        simt::launch(blocks, 256, [&] { transpose_p2p_kernel<float4>((const float4 *)grid, peer.data(), 0, n0, rank * n0, P, chunk_bytes / 16, rank); });
    else
        simt::launch(blocks, 256, [&] { transpose_p2p_kernel<float2>((const float2 *)grid, peer.data(), 0, n0, rank * n0, P, chunk_bytes / 8, rank); });
    return 0;
}

extern "C" int simt_accumulate(float *dst, const float *src, long long n, int num_sms) {
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)num_sms * 8);
    simt::launch(blocks, 256, [&] { accumulate_kernel(dst, src, n); });
    return 0;
}

// gridded field (float64 [rows][N]) -> padded float32 mesh with the mean removed, its padded sum, and back
extern "C" int simt_mesh_roundtrip(const double *field, long long rows, int N, double *sum_field, float *mesh,
                                   double *sum_mesh, double scale, double *back, int num_sms) {
    const int ldz = 2 * (N / 2 + 1);
    auto grid_of = [&](long long n) { long long w = (n + 255) / 256, cap = (long long)num_sms * 8; return (int)(w < 1 ? 1 : (w < cap ? w : cap)); };
    *sum_field = 0.0;
    simt::launch(grid_of(rows * N), 256, [&] { sum_kernel<double>(field, rows * N, sum_field); });
    const double mean = *sum_field / (double)(rows * N);
    simt::launch(grid_of(rows * ldz), 256, [&] { load_mesh_kernel<double>(field, rows, N, ldz, mean, mesh); });
    *sum_mesh = 0.0;
    simt::launch(grid_of(rows * N), 256, [&] { padded_sum_kernel(mesh, rows, N, ldz, sum_mesh); });
    simt::launch(grid_of(rows * N), 256, [&] { store_mesh_kernel(mesh, rows, N, ldz, scale, back); });
    return 0;
}

// PowerSpectrum3D._read_data's NGP assignment (ingest.cu): claim pass + write pass, as apk_assign_grid launches them
extern "C" int simt_assign_grid(const void *x, const void *y, const void *z, int pos_f64, const void *values, int val_f64,
                                long long n, int N, double *value_map, unsigned int *winner, unsigned long long *bad,
                                int num_sms) {
    const size_t cells = (size_t)N * N * N;
    std::fill(value_map, value_map + cells, 0.0);
    std::fill(winner, winner + cells, 0u);
    *bad = 0;
    if (n == 0) return 0;
    const int g = ingest_grid(n, num_sms);
#define CLAIM(CT) simt::launch(g, 256, [&] { assign_claim_kernel<CT>((const CT *)x, (const CT *)y, (const CT *)z, n, N, winner, bad); })
#define WRITE(CT, VT) simt::launch(g, 256, [&] { assign_write_kernel<CT, VT>((const CT *)x, (const CT *)y, (const CT *)z, (const VT *)values, n, N, winner, value_map); })
    if (pos_f64) { CLAIM(double); if (val_f64) WRITE(double, double); else WRITE(double, float); }
    else { CLAIM(float); if (val_f64) WRITE(float, double); else WRITE(float, float); }
#undef CLAIM
#undef WRITE
    return 0;
}

// Ecosmog.compress_snapshot's record gather (ingest.cu): one CTA per piece (byte offset, destination, count)
extern "C" int simt_gather_records(const unsigned char *raw, const long long *pieces, long long npieces, double *out) {
    if (npieces == 0) return 0;
    simt::launch((int)npieces, 256, [&] { gather_records_kernel(raw, (const RecordPiece *)pieces, out); });
    return 0;
}
