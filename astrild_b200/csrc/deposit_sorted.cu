// Particle -> mesh mass assignment, sorted / shared-memory-tiled variant (APK_DEPOSIT_SORTED).
//
// Replaces pm.paint(pos, mass=, resampler=) as astrild calls it at
//   /root/reference/src/astrild/particles/hutils/stats_subfind.py:130-131
// (pmesh 0.1.55 CIC / TSC windows, see deposit_common.cuh).
//
// Bound: HBM.  Algorithmic bytes = Np*(12 + 4*[mass]) + 4*N^3; the sort
// traffic is real but not credited.
//
// Pipeline (all on one stream, no host sync):
//   1. brick_count_kernel   every particle -> key of the 12x12x32-cell brick that holds its HOME
//                           cell (float64 index arithmetic, identical to the oracle's); per-brick
//                           counts with warp-aggregated RED (__match_any_sync: one atomic per run
//                           of equal keys in a warp -- snapshot order is spatially coherent).
//   2. brick_scan_kernel    exclusive scan of the counts -> brick_start[], cursors.
//   3. brick_scatter_kernel keys are recomputed (never stored); each run of equal keys claims
//                           its slots with ONE atomicAdd on the brick's cursor and writes its
//                           payload = brick-local coordinates as 3 floats (+ mass), contiguously.
//                           One read and one write of the particles replace a multi-pass radix
//                           sort; order inside a brick is arbitrary (the deposit does not care).
//   4. brick_deposit_kernel persistent CTAs pull bricks from a counter; per brick and per chunk
//      of <= CH particles: counting-sort the chunk by home cell inside shared memory (native
//      32-bit ATOMS.ADD gives each particle its rank), then one thread per home cell sums the
//      S^3 window moments of its own particles in registers.  The moments are spread WITHOUT
//      atomics (shared-memory float atomics are CAS loops on sm_100): lanes of a warp are the 32
//      z-cells of one (x,y) column, so the z-spread is two warp shuffles, and the (x,y)-spread is
//      a plain load/add/store into the shared tile that is conflict-free because the 16 columns
//      active at a time are 3 cells apart in x and y (9 colour classes, one __syncthreads each).
//      The finished (12+S-1)^2 x (32+S-1) tile is added to the mesh with RED.ADD.F32, skipping
//      zeros; bricks are visited x-major so neighbouring tiles meet in L2.
#include "apk_common.cuh"
#include "deposit_common.cuh"
#include <algorithm>
#include <type_traits>

namespace apk {

constexpr int BX = 12, BY = 12, BZ = 32;        // brick edge in cells (x, y multiples of 3; z = warp)
constexpr int BRICK_CELLS = BX * BY * BZ;       // 4608
constexpr int DEP_THREADS = 512;                // 16 warps = 16 columns of one colour class
constexpr int CH = 5120;                        // particles per shared-memory chunk
constexpr int PPT = CH / DEP_THREADS;           // particles per thread per chunk

struct P3 { float x, y, z; };
struct P4 { float x, y, z, m; };

struct BrickGrid {
    int nbx, nby, nbz;   // bricks per axis (x counts local planes for slab plans)
    int nbricks;
};

static BrickGrid make_brick_grid(const DepositGeom &G) {
    BrickGrid B;
    B.nbx = (G.nplanes + BX - 1) / BX;
    B.nby = (G.N + BY - 1) / BY;
    B.nbz = (G.N + BZ - 1) / BZ;
    B.nbricks = B.nbx * B.nby * B.nbz;
    return B;
}

// home cell of a particle on one axis: floor(g) for CIC, floor(g + 0.5) for TSC / NGP
template <int S>
__device__ __forceinline__ double home_of(double g) { return (S == 2) ? floor(g) : floor(g + 0.5); }

// brick key and brick-local coordinates of particle p (shared by the count and scatter passes so
// both see bit-identical keys)
template <int S, typename PT, bool SOA>
__device__ __forceinline__ unsigned int brick_of(const PT *__restrict__ p0, const PT *__restrict__ p1,
                                                 const PT *__restrict__ p2, long long p, const DepositGeom &G,
                                                 const BrickGrid &B, float (&l)[3]) {
    double g[3];
    if (SOA) { g[0] = (double)p0[p]; g[1] = (double)p1[p]; g[2] = (double)p2[p]; }
    else     { g[0] = (double)p0[3 * p]; g[1] = (double)p0[3 * p + 1]; g[2] = (double)p0[3 * p + 2]; }
    int b[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        g[d] = g[d] * G.scale + G.shift;
        const double h = home_of<S>(g[d]);
        const double frac = g[d] - h;                       // [0,1) CIC, [-0.5,0.5) TSC
        int hl = (d == 0) ? G.local_plane((long long)h) : wrap_index((long long)h, G.N);
        if (hl < 0) hl = 0;                                 // slab plan, particle not routed here: caller error
        const int edge = d == 0 ? BX : (d == 1 ? BY : BZ);
        b[d] = hl / edge;
        l[d] = (float)(frac + (double)(hl - b[d] * edge));
    }
    return (unsigned int)((b[0] * B.nby + b[1]) * B.nbz + b[2]);
}

constexpr int PART_THREADS = 256;
constexpr int PART_ITEMS = 4;    // particles per thread per tile (ILP on the loads)

template <int S, typename PT, bool SOA>
__global__ void __launch_bounds__(PART_THREADS)
brick_count_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2, long long np,
                   DepositGeom G, BrickGrid B, unsigned int *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    for (long long base = (long long)blockIdx.x * tile; base < np; base += (long long)gridDim.x * tile) {
        unsigned int key[PART_ITEMS];
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const long long p = base + k * PART_THREADS + threadIdx.x;
            float l[3];
            key[k] = p < np ? brick_of<S, PT, SOA>(p0, p1, p2, p, G, B, l) : 0xffffffffu;
        }
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const unsigned int peers = __match_any_sync(0xffffffffu, key[k]);
            if (key[k] != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(counts + key[k], (unsigned int)__popc(peers));
        }
    }
}

// exclusive scan of counts[0..n) -> start[0..n], cursor[0..n) = start; one CTA (n is ~10^4..10^6)
__global__ void __launch_bounds__(1024)
brick_scan_kernel(const unsigned int *__restrict__ counts, int n, unsigned int *__restrict__ start,
                  unsigned int *__restrict__ cursor) {
    __shared__ unsigned int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int a = min(tid * per, n), b = min(a + per, n);
    unsigned int s = 0;
    for (int i = a; i < b; ++i) s += counts[i];
    unsigned int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned int w = wsum[lane];
        unsigned int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    unsigned int run = wsum[warp] + incl - s;
    for (int i = a; i < b; ++i) { start[i] = run; cursor[i] = run; run += counts[i]; }
    if (tid == 1023) start[n] = run;
}

template <int S, typename PT, bool SOA, bool MASS, typename VT>
__global__ void __launch_bounds__(PART_THREADS)
brick_scatter_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2,
                     const void *__restrict__ mass, int mass_f64, long long np, DepositGeom G, BrickGrid B,
                     unsigned int *__restrict__ cursor, VT *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    for (long long base = (long long)blockIdx.x * tile; base < np; base += (long long)gridDim.x * tile) {
        unsigned int key[PART_ITEMS];
        VT v[PART_ITEMS];
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const long long p = base + k * PART_THREADS + threadIdx.x;
            key[k] = 0xffffffffu;
            if (p < np) {
                float l[3];
                key[k] = brick_of<S, PT, SOA>(p0, p1, p2, p, G, B, l);
                v[k].x = l[0]; v[k].y = l[1]; v[k].z = l[2];
                if constexpr (MASS) v[k].m = mass_f64 ? (float)((const double *)mass)[p] : ((const float *)mass)[p];
            }
        }
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const unsigned int peers = __match_any_sync(0xffffffffu, key[k]);
            const int leader = __ffs(peers) - 1;
            unsigned int slot = 0;
            if (key[k] != 0xffffffffu && lane == leader) slot = atomicAdd(cursor + key[k], (unsigned int)__popc(peers));
            slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(peers & lt_mask);
            if (key[k] != 0xffffffffu) vals[slot] = v[k];
        }
    }
}

template <int S>
struct TileDims {
    static constexpr int TX = BX + S - 1, TY = BY + S - 1, TZ = BZ + S - 1;
    static constexpr int SIZE = TX * TY * TZ;
};

// dynamic shared memory layout of brick_deposit_kernel
template <int S, bool MASS>
struct DepSmem {
    static constexpr int tile_floats = TileDims<S>::SIZE;
    static constexpr int cnt_ints = BRICK_CELLS + 1;
    static constexpr size_t bytes = sizeof(float) * tile_floats + sizeof(int) * (cnt_ints + 64) +
                                    sizeof(float) * CH * (MASS ? 4 : 3) + 64;
};

template <int S, bool MASS, typename VT>
__global__ void __launch_bounds__(DEP_THREADS, 2)
brick_deposit_kernel(const VT *__restrict__ vals, const unsigned int *__restrict__ brick_start,
                     DepositGeom G, BrickGrid B, unsigned int *__restrict__ work_counter,
                     float *__restrict__ mesh) {
    using TD = TileDims<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);
    int *cnt = reinterpret_cast<int *>(tile + TD::SIZE);          // [BRICK_CELLS + 1]
    int *wsum = cnt + BRICK_CELLS + 1;                            // [32] scan scratch (+ pad)
    float *sx = reinterpret_cast<float *>(wsum + 64);
    float *sy = sx + CH;
    float *sz = sy + CH;
    float *sm = sz + CH;                                          // only if MASS
    __shared__ unsigned int s_brick;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int OFF = (S == 3) ? 1 : 0;   // tile origin = brick origin - OFF

    for (;;) {
        __syncthreads();
        if (tid == 0) s_brick = atomicAdd(work_counter, 1u);
        __syncthreads();
        const unsigned int brick = s_brick;
        if (brick >= (unsigned)B.nbricks) break;
        const unsigned int pbeg = brick_start[brick], pend = brick_start[brick + 1];
        if (pbeg == pend) continue;

        for (int i = tid; i < TD::SIZE; i += DEP_THREADS) tile[i] = 0.f;

        for (unsigned int c0 = pbeg; c0 < pend; c0 += CH) {
            const int nchunk = (int)min((unsigned int)CH, pend - c0);
            for (int i = tid; i <= BRICK_CELLS; i += DEP_THREADS) cnt[i] = 0;
            __syncthreads();   // also orders the tile zeroing / previous chunk's spreading

            // ---- rank my particles inside their home cell (cell | rank << 13 kept in a register,
            //      the coordinates are re-read from L1/L2 in the scatter pass to stay <= 64 regs)
            int packed[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int i = k * DEP_THREADS + tid;
                packed[k] = -1;
                if (i < nchunk) {
                    const VT v = vals[c0 + i];
                    int hx, hy, hz;
                    if (S == 2) { hx = (int)floorf(v.x); hy = (int)floorf(v.y); hz = (int)floorf(v.z); }
                    else        { hx = (int)floorf(v.x + 0.5f); hy = (int)floorf(v.y + 0.5f); hz = (int)floorf(v.z + 0.5f); }
                    hx = max(0, min(hx, BX - 1)); hy = max(0, min(hy, BY - 1)); hz = max(0, min(hz, BZ - 1));
                    const int cell = (hx * BY + hy) * BZ + hz;
                    packed[k] = cell | (atomicAdd(&cnt[cell], 1) << 13);
                }
            }
            __syncthreads();

            // ---- exclusive scan of the 4608 cell counts (9 per thread) -------------------
            {
                constexpr int PER = BRICK_CELLS / DEP_THREADS;   // 9
                int v[PER], s = 0;
#pragma unroll
                for (int k = 0; k < PER; ++k) { v[k] = cnt[tid * PER + k]; s += v[k]; }
                int incl = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) wsum[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    int w = lane < DEP_THREADS / 32 ? wsum[lane] : 0;
                    int wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, wi, o);
                        if (lane >= o) wi += t;
                    }
                    wsum[32 + lane] = wi - w;   // exclusive warp offsets
                }
                __syncthreads();
                int run = wsum[32 + warp] + incl - s;
#pragma unroll
                for (int k = 0; k < PER; ++k) { cnt[tid * PER + k] = run; run += v[k]; }
                if (tid == DEP_THREADS - 1) cnt[BRICK_CELLS] = run;
            }
            __syncthreads();

            // ---- scatter into cell order --------------------------------------------------
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (packed[k] >= 0) {
                    const VT v = vals[c0 + k * DEP_THREADS + tid];
                    const int slot = cnt[packed[k] & 8191] + (packed[k] >> 13);
                    sx[slot] = v.x; sy[slot] = v.y; sz[slot] = v.z;
                    if constexpr (MASS) sm[slot] = v.m;
                }
            }
            __syncthreads();

            // ---- moments per home cell, conflict-free spreading, 9 colour classes --------
            for (int cls = 0; cls < 9; ++cls) {
                const int cx = 3 * (warp >> 2) + cls / 3;
                const int cy = 3 * (warp & 3) + cls % 3;
                const int cell = (cx * BY + cy) * BZ + lane;
                const int beg = cnt[cell], end = cnt[cell + 1];
                float M[S][S][S];
#pragma unroll
                for (int a = 0; a < S; ++a)
#pragma unroll
                    for (int b = 0; b < S; ++b)
#pragma unroll
                        for (int c = 0; c < S; ++c) M[a][b][c] = 0.f;
                for (int p = beg; p < end; ++p) {
                    float wx[S], wy[S], wz[S];
                    const float dx = sx[p] - (float)cx, dy = sy[p] - (float)cy, dz = sz[p] - (float)lane;
                    if (S == 2) {
                        wx[0] = 1.f - dx; wx[S - 1] = dx;
                        wy[0] = 1.f - dy; wy[S - 1] = dy;
                        wz[0] = 1.f - dz; wz[S - 1] = dz;
                    } else {
                        wx[0] = 0.5f * (0.5f - dx) * (0.5f - dx); wx[S / 2] = 0.75f - dx * dx; wx[S - 1] = 0.5f * (0.5f + dx) * (0.5f + dx);
                        wy[0] = 0.5f * (0.5f - dy) * (0.5f - dy); wy[S / 2] = 0.75f - dy * dy; wy[S - 1] = 0.5f * (0.5f + dy) * (0.5f + dy);
                        wz[0] = 0.5f * (0.5f - dz) * (0.5f - dz); wz[S / 2] = 0.75f - dz * dz; wz[S - 1] = 0.5f * (0.5f + dz) * (0.5f + dz);
                    }
                    if constexpr (MASS) {
                        const float m = sm[p];
#pragma unroll
                        for (int a = 0; a < S; ++a) wx[a] *= m;
                    }
#pragma unroll
                    for (int a = 0; a < S; ++a)
#pragma unroll
                        for (int b = 0; b < S; ++b) {
                            const float wxy = wx[a] * wy[b];
#pragma unroll
                            for (int c = 0; c < S; ++c) M[a][b][c] = fmaf(wxy, wz[c], M[a][b][c]);
                        }
                }
                // z-spread by shuffles: tile z index t = lane + jz.  The lane's own target is
                // t = lane + OFF; lane 0 / lane 31 also feed the two z-halo cells.
                const bool any = __ballot_sync(0xffffffffu, end > beg) != 0u;
                if (any) {
#pragma unroll
                    for (int a = 0; a < S; ++a)
#pragma unroll
                        for (int b = 0; b < S; ++b) {
                            float own, lo_halo = 0.f, hi_halo;
                            if (S == 2) {
                                float up = __shfl_up_sync(0xffffffffu, M[a][b][S - 1], 1);
                                if (lane == 0) up = 0.f;
                                own = M[a][b][0] + up;
                                hi_halo = M[a][b][S - 1];      // lane 31 -> t = 32
                            } else {
                                float up = __shfl_up_sync(0xffffffffu, M[a][b][S - 1], 1);
                                float dn = __shfl_down_sync(0xffffffffu, M[a][b][0], 1);
                                if (lane == 0) up = 0.f;
                                if (lane == 31) dn = 0.f;
                                own = M[a][b][S / 2] + up + dn;
                                lo_halo = M[a][b][0];          // lane 0  -> t = 0
                                hi_halo = M[a][b][S - 1];      // lane 31 -> t = 33
                            }
                            float *row = tile + ((cx + a) * TD::TY + (cy + b)) * TD::TZ;
                            row[lane + OFF] += own;
                            if (S == 3 && lane == 0) row[0] += lo_halo;
                            if (lane == 31) row[TD::TZ - 1] += hi_halo;
                        }
                }
                __syncthreads();
            }
        }

        // ---- add the tile to the mesh -----------------------------------------------------
        const int bz = brick % B.nbz;
        const int by = (brick / B.nbz) % B.nby;
        const int bx = brick / (B.nbz * B.nby);
        for (int i = tid; i < TD::SIZE; i += DEP_THREADS) {
            const float v = tile[i];
            if (v == 0.f) continue;
            const int tz = i % TD::TZ;
            const int ty = (i / TD::TZ) % TD::TY;
            const int tx = i / (TD::TZ * TD::TY);
            int px = bx * BX - OFF + tx;
            if (G.slab) { if (px < 0 || px >= G.nplanes) continue; }
            else px = wrap_index(px, G.N);
            const int gy = wrap_index(by * BY - OFF + ty, G.N);
            const int gz = wrap_index(bz * BZ - OFF + tz, G.N);
            atomicAdd(mesh + ((size_t)px * G.N + gy) * G.ldz + gz, v);
        }
    }
}

static size_t max_bricks(const apk_plan *P) {
    return (size_t)((P->N + 3 + BX - 1) / BX) * ((P->N + BY - 1) / BY) * ((P->N + BZ - 1) / BZ);
}
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t deposit_sorted_workspace_bytes(const apk_plan *P, long long np, int with_mass) {
    if (np <= 0) return 0;
    const size_t vs = with_mass ? sizeof(P4) : sizeof(P3);
    return align256(vs * (size_t)np) + 3 * align256(4 * (max_bricks(P) + 2)) + 256;
}

template <int S, typename PT, bool SOA, bool MASS>
static int run_sorted(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                      int mass_dtype, long long np, const DepositGeom &G, float *mesh, cudaStream_t st) {
    using VT = typename std::conditional<MASS, P4, P3>::type;
    const BrickGrid B = make_brick_grid(G);
    const size_t need = deposit_sorted_workspace_bytes(P, np, MASS);
    APK_REQUIRE(P->workspace && P->workspace_bytes >= need,
                "apk_deposit: sorted path needs %zu workspace bytes, %zu set (apk_plan_workspace_bytes / apk_plan_set_workspace)",
                need, P->workspace_bytes);
    APK_REQUIRE(np < 0xffffffffLL, "apk_deposit: more than 2^32-1 particles on one device");
    unsigned char *w = (unsigned char *)P->workspace;
    VT *vals = (VT *)w; w += align256(sizeof(VT) * (size_t)np);
    const size_t tab = align256(4 * (max_bricks(P) + 2));
    unsigned int *counts = (unsigned int *)w; w += tab;
    unsigned int *brick_start = (unsigned int *)w; w += tab;
    unsigned int *cursor = (unsigned int *)w; w += tab;
    unsigned int *counter = (unsigned int *)w;

    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    const int pb = (int)std::min<long long>((np + tile - 1) / tile, (long long)P->num_sms * 8);
    P->mark(0, st);
    APK_CUDA(cudaMemsetAsync(counts, 0, 4 * (size_t)(B.nbricks + 1), st));
    brick_count_kernel<S, PT, SOA><<<pb, PART_THREADS, 0, st>>>((const PT *)p0, (const PT *)p1, (const PT *)p2, np, G, B, counts);
    APK_CUDA(cudaGetLastError());
    P->mark(1, st);
    brick_scan_kernel<<<1, 1024, 0, st>>>(counts, B.nbricks, brick_start, cursor);
    APK_CUDA(cudaGetLastError());
    P->mark(2, st);
    brick_scatter_kernel<S, PT, SOA, MASS, VT><<<pb, PART_THREADS, 0, st>>>(
        (const PT *)p0, (const PT *)p1, (const PT *)p2, mass, mass_dtype == APK_F64, np, G, B, cursor, vals);
    APK_CUDA(cudaGetLastError());
    APK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));

    auto kern = brick_deposit_kernel<S, MASS, VT>;
    const size_t smem = DepSmem<S, MASS>::bytes;
    APK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    APK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEP_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int ctas = std::min(P->num_sms * per_sm, B.nbricks);
    P->mark(3, st);
    kern<<<ctas, DEP_THREADS, smem, st>>>(vals, brick_start, G, B, counter, mesh);
    APK_CUDA(cudaGetLastError());
    P->mark(4, st);
    P->dep_timed = P->timing;
    P->dep_sorted = true;
    return 0;
}

template <int S, typename PT, bool SOA>
static int dispatch_mass(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                         int mass_dtype, long long np, const DepositGeom &G, float *mesh, cudaStream_t st) {
    return mass ? run_sorted<S, PT, SOA, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st)
                : run_sorted<S, PT, SOA, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st);
}

template <int S>
static int dispatch_layout(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                           const void *mass, int mass_dtype, long long np, const DepositGeom &G, float *mesh,
                           cudaStream_t st) {
    if (pos_dtype == APK_F32)
        return layout == APK_SOA ? dispatch_mass<S, float, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st)
                                 : dispatch_mass<S, float, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st);
    return layout == APK_SOA ? dispatch_mass<S, double, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st)
                             : dispatch_mass<S, double, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, st);
}

int deposit_atomic_launch(const void *, const void *, const void *, int, int, const void *, int, long long,
                          int, const DepositGeom &, float *, int, cudaStream_t);

int deposit_sorted_launch(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                          const void *mass, int mass_dtype, long long np, int resampler, const DepositGeom &G,
                          float *mesh, cudaStream_t st) {
    if (np == 0) return 0;
    if (resampler == APK_CIC) return dispatch_layout<2>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, st);
    if (resampler == APK_TSC) return dispatch_layout<3>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, st);
    // NGP has no halo and no arithmetic worth tiling: one RED per particle
    return deposit_atomic_launch(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, resampler, G, mesh, P->num_sms, st);
}

}  // namespace apk
