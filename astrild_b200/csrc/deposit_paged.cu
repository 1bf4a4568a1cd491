// Particle -> mesh mass assignment, sorted path with a ONE-PASS partition into paged brick buckets.
//
// Replaces pm.paint(pos, mass=, resampler=) as astrild calls it at
//   /root/reference/src/astrild/particles/hutils/stats_subfind.py:130-131
// (pmesh 0.1.55 CIC / TSC windows, see deposit_common.cuh).
//
// Bound: HBM.  Algorithmic bytes = Np*(12 + 4*[mass]) + 4*N^3; the partition's write and
// re-read of the particles is real traffic but not credited.
//
//   1. brick_partition_kernel  ONE pass over the particles: key of the brick (12 x 6 x 30 home cells for TSC,
//      12 x 6 x 31 for CIC) that holds the particle's home cell, one returning atomicAdd on the brick's cursor per
//      warp-run of equal keys (snapshot order is spatially coherent), payload = brick-local coordinates as 3 floats
//      (+ mass) written into the brick's PAGES.  A brick's bucket is a list of fixed-size pages (PAGE slots) taken
//      from a pool as the cursor crosses page boundaries: the particle that claims the first slot of a page allocates
//      it (one atomicAdd on the pool counter) and publishes its id in the brick's page table; the others read the id
//      from there (the allocating lane claimed its slot earlier, so the wait is short and cannot dead-lock).  No
//      count pass, no scan: the bucket sizes need not be known in advance.
//      Bricks denser than PAGES_MAX pages (8 x the mean density of one particle per cell) send the excess straight
//      to the mesh with float REDs from inside this kernel -- correct for any input, fast for any sane one.
//      PAIR mode (apk_deposit_interlaced): one partition serves both interlaced meshes -- particles whose two home
//      cells fall in different bricks are filed twice, sign bits of the payload say which mesh a copy is for.
//   2. brick_tile_kernel  one CTA per brick, one THREAD per particle: the brick's window of the mesh lives in shared
//      memory as 32-bit FIXED-POINT integers and every one of the S^3 weights goes there with a native integer
//      ATOMS.ADD (shared-memory float atomics are CAS loops on sm_100; integer adds are not).  No in-brick sort and no
//      per-cell loop: all 32 lanes work on every instruction, whatever the cell occupancy.
//        Quantum: 2^-s of the largest |mass| in the chunk (1 for unit masses), s chosen PER CHUNK of <= TILE_FLUSH
//        particles as large as 32 bits allow if every particle of the chunk put its largest possible weight (1 for
//        CIC, 0.75^3 for TSC) into one cell: s = 22 for the ~2200 particles of a brick at one particle per cell (TSC).
//        A weight is rounded to the quantum by ONE FFMA against a magic constant (1.5 * 2^(23 - s): the sum lands in a
//        binade whose ulp is the quantum, so the low mantissa bits ARE the fixed-point value).  Integer adds commute:
//        a brick's contribution to the mesh does not depend on the order in which the partition filed its particles.
//      The tile goes to the mesh as one coalesced 128-byte RED.ADD.F32 per (x,y) column, zeros skipped; bricks are
//      visited x-major so neighbouring windows meet in L2.
#include "brick_common.cuh"
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace apk {

constexpr int PAGE_SHIFT = 7, PAGE = 1 << PAGE_SHIFT;      // slots per page: 1.5 KB (unit masses) / 2 KB
constexpr int PAGES_MAX = 128;                              // page-table entries per brick: 16384 particles
constexpr unsigned int BUCKET_MAX = PAGES_MAX * PAGE;

// ---- page table access: entries are page id + 1, 0 = not allocated yet --------------------------------------
__device__ __forceinline__ unsigned int page_wait(const unsigned int *entry) {
    unsigned int id;
    while ((id = *(const volatile unsigned int *)entry) == 0u) spin_pause();
    return id - 1u;
}
__device__ __forceinline__ void page_publish(unsigned int *entry, unsigned int id) {
    *(volatile unsigned int *)entry = id + 1u;
}

// the rare particle of an over-full brick: straight to the mesh(es), float64 window arithmetic like deposit_atomic.cu
template <int S, typename PT>
__device__ __noinline__ void deposit_direct(PT x0, PT x1, PT x2, float m, const DepositGeom &G, float *__restrict__ mesh) {
    long long i0[3];
    float w[3][S];
    const double x[3] = {(double)x0, (double)x1, (double)x2};
#pragma unroll
    for (int d = 0; d < 3; ++d) window_1d<S>(x[d] * G.scale + G.shift, i0[d], w[d]);
    for (int a = 0; a < S; ++a) {
        const int px = G.local_plane(i0[0] + a);
        if (px < 0) continue;
        for (int b = 0; b < S; ++b) {
            float *row = mesh + ((size_t)px * G.N + wrap_index(i0[1] + b, G.N)) * G.ldz;
            for (int c = 0; c < S; ++c) atomicAdd(row + wrap_index(i0[2] + c, G.N), m * w[0][a] * w[1][b] * w[2][c]);
        }
    }
}

// Files this lane's payload(s) of one item: v under `slot` of brick `key` and, for the interlaced pair, the twin's own
// copy w under `slot1` of brick `key1`.  Called by ALL lanes of the warp (live / extra = this lane has such a copy).
// Two phases separated by a warp barrier: first every lane that owns the first slot of a page -- with either copy --
// allocates and publishes it, then the pages are looked up: a lane never waits for a page that a lane of its own warp
// has yet to publish (lanes of other warps claimed their slots earlier and publish without waiting for anyone).
// Returns a bit mask of the copies that did NOT fit (their brick is full): 1 = v, 2 = w.
template <bool PAIR, typename VT>
__device__ __forceinline__ unsigned int page_store(bool live, unsigned int key, unsigned int slot, const VT &v,
                                                   bool extra, unsigned int key1, unsigned int slot1, const VT &w,
                                                   unsigned int *__restrict__ table, unsigned int *__restrict__ pool_next,
                                                   VT *__restrict__ pool) {
    const bool fits = live && slot < BUCKET_MAX, fits1 = PAIR && extra && slot1 < BUCKET_MAX;
    unsigned int *entry = table + (size_t)(fits ? key : 0u) * PAGES_MAX + (fits ? slot >> PAGE_SHIFT : 0u);
    unsigned int *entry1 = table + (size_t)(fits1 ? key1 : 0u) * PAGES_MAX + (fits1 ? slot1 >> PAGE_SHIFT : 0u);
    const unsigned int r = slot & (PAGE - 1), r1 = slot1 & (PAGE - 1);
    const bool first = fits && r == 0u, first1 = fits1 && r1 == 0u;
    unsigned int id = 0u, id1 = 0u;
    if (first) { id = atomicAdd(pool_next, 1u); page_publish(entry, id); }
    if (PAIR && first1) { id1 = atomicAdd(pool_next, 1u); page_publish(entry1, id1); }
    __syncwarp();
    if (fits && !first) id = page_wait(entry);
    if (fits) pool[(size_t)id * PAGE + r] = v;
    if (PAIR) {
        if (fits1 && !first1) id1 = page_wait(entry1);
        if (fits1) pool[(size_t)id1 * PAGE + r1] = w;
    }
    return (live && !fits ? 1u : 0u) | (PAIR && extra && !fits1 ? 2u : 0u);
}

#ifndef APK_PART_CTAS
#define APK_PART_CTAS 3
#endif
template <int S, typename PT, bool SOA, bool MASS, bool PAIR, typename VT>
__global__ void __launch_bounds__(PART_THREADS, APK_PART_CTAS)
brick_partition_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2,
                       const void *__restrict__ mass, int mass_f64, long long np, DepositGeom G, DepositGeom G1,
                       BrickGrid B, unsigned int *__restrict__ cursor, unsigned int *__restrict__ table,
                       unsigned int *__restrict__ pool_next, VT *__restrict__ pool, float *__restrict__ mesh,
                       float *__restrict__ mesh1) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    const long long step = (long long)gridDim.x * tile;
    long long base = (long long)blockIdx.x * tile;
    Raw4<PT> nxt;
    if (base < np) load4<PT, SOA>(p0, p1, p2, base + threadIdx.x, PART_THREADS, np, nxt);
    for (; base < np; base += step) {
        const long long first = base + threadIdx.x;
        const Raw4<PT> cur = nxt;
        if (base + step < np) load4<PT, SOA>(p0, p1, p2, first + step, PART_THREADS, np, nxt);   // next tile in flight
        // Software pipeline of depth one: item k claims its slots (returning atomics, ~600 cycles) while item k-1,
        // whose slots have arrived meanwhile, is stored.
        VT pv = {}, pw = {};
        unsigned int pslot = 0, pslot1 = 0, pkey = 0, pkey1 = 0;
        int phead = 0, poffset = 0;
        bool plive = false, pextra = false, psplit = false;
#pragma unroll
        for (int k = 0; k <= 4; ++k) {
            VT v = {}, w = {};
            unsigned int slot = 0, slot1 = 0, key = 0xffffffffu, key1 = 0xffffffffu;
            int head = 0, offset = 0, length = 0;
            bool live = false, extra = false, split = false;
            if (k < 4) {
                const long long p = first + (long long)k * PART_THREADS;
                float l[3], l1[3];
                brick_keys<S, PT, PAIR>(cur.v + 3 * k, G, G1, B, key, l, key1, l1, split);
                if (p >= np) key = 0xffffffffu;
                live = key != 0xffffffffu;
                v.x = l[0]; v.y = l[1]; v.z = l[2];
                if constexpr (MASS) {
                    const long long pc = min(p, np - 1);
                    v.m = mass_f64 ? (float)((const double *)mass)[pc] : ((const float *)mass)[pc];
                }
                if constexpr (PAIR) {
                    v.x = fmaxf(l[0] + 1.f, 0.f); v.y = fmaxf(l[1] + 1.f, 0.f); v.z = fmaxf(l[2] + 1.f, 0.f);
                    if (split) v.y = -v.y;                           // first copy is mesh-0-only
                }
                warp_runs(key, lane, head, offset, length);
                if (live && offset == 0) slot = atomicAdd(cursor + key, (unsigned int)length);
                if constexpr (PAIR) {
                    extra = live && split;
                    if (extra) {                                     // second copy: mesh-1-only, in mesh 1's brick
                        w = v;
                        w.x = -fmaxf(l1[0] + 0.5f, 0.f); w.y = fmaxf(l1[1] + 0.5f, 0.f); w.z = fmaxf(l1[2] + 0.5f, 0.f);
                        slot1 = atomicAdd(cursor + key1, 1u);
                    }
                }
            }
            if (k > 0) {
                const int KP = k > 0 ? k - 1 : 0;                    // compile-time after unrolling: cur.v stays in registers
                const unsigned int ps = __shfl_sync(0xffffffffu, pslot, phead) + poffset;
                float pm = 1.f;
                if constexpr (MASS) pm = pv.m;
                const unsigned int full = page_store<PAIR, VT>(plive, pkey, ps, pv, pextra, pkey1, pslot1, pw, table, pool_next, pool);
                if (full) {
                    // over-full brick: the first copy serves mesh 0, and mesh 1 too unless the twin has a copy of its own
                    const PT x0 = cur.v[3 * KP], x1 = cur.v[3 * KP + 1], x2 = cur.v[3 * KP + 2];
                    if (full & 1u) {
                        deposit_direct<S, PT>(x0, x1, x2, pm, G, mesh);
                        if (PAIR && !psplit) deposit_direct<S, PT>(x0, x1, x2, pm, G1, mesh1);
                    }
                    if (PAIR && (full & 2u)) deposit_direct<S, PT>(x0, x1, x2, pm, G1, mesh1);
                }
            }
            pv = v; pw = w; pslot = slot; pslot1 = slot1; pkey = key; pkey1 = key1; phead = head; poffset = offset;
            plive = live; pextra = extra; psplit = split;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int TILE_FLUSH = 4095;                      // particles between two flushes of the tile
constexpr int TILE_THREADS = 256;
#ifndef APK_TILE_CTAS
#define APK_TILE_CTAS 6
#endif

template <int S> struct Tile {
    static constexpr int OFF = (S == 3) ? 1 : 0;      // window origin = home cell - OFF
    static constexpr int TX = BX + S - 1, TY = BY + S - 1, TZ = BZ;
    static constexpr int CELLS = TX * TY * TZ;
};

// fractional bits for a chunk of n particles: n * wmax * 2^s < 2^32 (2^31 with signed masses), s <= SMAX (the
// magic-constant trick needs wmax <= 2^(22 - s))
template <int S, bool MASS>
__device__ __forceinline__ int tile_frac_bits(int n) {
    constexpr float WMAX = (S == 3) ? 0.43f : 1.001f;          // 0.75^3 = 0.4219 / 1, padded for the rounding
    constexpr int SMAX = (S == 3) ? 23 : 22;
    const float room = ((MASS ? 2147483648.f : 4294967296.f) / WMAX) / (float)n;
    const int s = ((__float_as_int(room) >> 23) & 0xff) - 127;  // floor(log2(room))
    return min(s, SMAX);
}

// nearest brick-local home cell of coordinate l on one axis (clamped to the brick) as float and int, and the offset
// d = l - home: [-0.5, 0.5] for TSC, [0, 1] for CIC (ties land on either side; the windows are continuous there)
template <int S>
__device__ __forceinline__ void tile_home(float l, float last, float &d, int &h) {
    const float M = 12582912.f;                       // 1.5 * 2^23: adding it rounds to the nearest integer
    float t = (S == 2 ? l - 0.5f : l) + M;
    t = fminf(fmaxf(t, M), M + last);
    h = __float_as_int(t) & 0x3fffff;
    d = l - (t - M);
}

template <int S, bool MASS, typename VT>
__global__ void __launch_bounds__(TILE_THREADS, APK_TILE_CTAS)
brick_tile_kernel(const VT *__restrict__ pool, const unsigned int *__restrict__ cursor, const unsigned int *__restrict__ table,
                  DepositGeom G, BrickGrid B, float *__restrict__ mesh, int sel) {
    using T = Tile<S>;
    __shared__ unsigned int tile[T::CELLS];
    __shared__ float wmax_s[TILE_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (unsigned int brick = blockIdx.x; brick < (unsigned int)B.nbricks; brick += gridDim.x) {
        const unsigned int count = min(cursor[brick], BUCKET_MAX);
        if (count == 0u) continue;
        const unsigned int *pages = table + (size_t)brick * PAGES_MAX;
        const int bz = brick % B.nbz;
        const int by = (brick / B.nbz) % B.nby;
        const int bx = brick / (B.nbz * B.nby);

        __syncthreads();                              // (persistent grids) the previous brick's flush is complete
        for (int i = tid; i < T::CELLS; i += TILE_THREADS) tile[i] = 0u;
        __syncthreads();

        for (unsigned int c0 = 0; c0 < count; c0 += TILE_FLUSH) {
            const unsigned int c1 = min(c0 + (unsigned int)TILE_FLUSH, count);
            // ---- mass unit of the chunk: the power of two at or above its largest |mass| ----
            float unit = 1.f, inv_unit = 1.f;
            if constexpr (MASS) {
                float top = 0.f;
                for (unsigned int p = c0 + tid; p < c1; p += TILE_THREADS)
                    top = fmaxf(top, fabsf(pool[(size_t)(pages[p >> PAGE_SHIFT] - 1u) * PAGE + (p & (PAGE - 1))].m));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, o));
                if (lane == 0) wmax_s[warp] = top;
                __syncthreads();
                top = wmax_s[0];
#pragma unroll
                for (int i = 1; i < TILE_THREADS / 32; ++i) top = fmaxf(top, wmax_s[i]);
                // 2^ceil(log2(top)), clamped to normal numbers whose reciprocal is normal too
                int e = ((__float_as_int(top) >> 23) & 0xff) + ((__float_as_int(top) & 0x7fffff) ? 1 : 0);
                e = min(max(e, 2), 252);
                unit = __int_as_float(e << 23);
                inv_unit = __int_as_float((254 - e) << 23);
                if (!(top > 0.f) || !(top < 3.0e38f)) { unit = 1.f; inv_unit = 0.f; }     // all-zero, inf or NaN masses: nothing to add
            }
            const int frac_bits = tile_frac_bits<S, MASS>((int)(c1 - c0));
            const unsigned int magic_bits = ((unsigned int)(127 + 23 - frac_bits) << 23) | 0x400000u;   // 1.5 * 2^(23 - s)
            const float magic = __int_as_float((int)magic_bits);
            const float quantum = __int_as_float((127 - frac_bits) << 23) * unit;                         // 2^-s mass units

            // ---- one thread per particle; the next particle's loads are in flight while this one is deposited ----
            unsigned int p = c0 + tid;
            VT nxt = {};
            if (p < c1) nxt = pool[(size_t)(pages[p >> PAGE_SHIFT] - 1u) * PAGE + (p & (PAGE - 1))];
            for (; p < c1; p += TILE_THREADS) {
                VT v = nxt;
                const unsigned int q = p + TILE_THREADS;
                if (q < c1) nxt = pool[(size_t)(pages[q >> PAGE_SHIFT] - 1u) * PAGE + (q & (PAGE - 1))];
                if (!unpack_pair(v, sel)) continue;
                float dx, dy, dz;
                int hx, hy, hz;
                tile_home<S>(v.x, (float)(BX - 1), dx, hx);
                tile_home<S>(v.y, (float)(BY - 1), dy, hy);
                tile_home<S>(v.z, (float)(BrickZ<S>::CELLS - 1), dz, hz);
                float wx[S], wy[S], wz[S];
                if (S == 2) {
                    wx[0] = 1.f - dx; wx[S - 1] = dx; wy[0] = 1.f - dy; wy[S - 1] = dy; wz[0] = 1.f - dz; wz[S - 1] = dz;
                } else {
                    const float ax = 0.5f - dx, cx = 0.5f + dx, ay = 0.5f - dy, cy = 0.5f + dy, az = 0.5f - dz, cz = 0.5f + dz;
                    wx[0] = 0.5f * ax * ax; wx[S / 2] = fmaf(-dx, dx, 0.75f); wx[S - 1] = 0.5f * cx * cx;
                    wy[0] = 0.5f * ay * ay; wy[S / 2] = fmaf(-dy, dy, 0.75f); wy[S - 1] = 0.5f * cy * cy;
                    wz[0] = 0.5f * az * az; wz[S / 2] = fmaf(-dz, dz, 0.75f); wz[S - 1] = 0.5f * cz * cz;
                }
                if constexpr (MASS) {
                    const float m = v.m * inv_unit;                              // |m| <= 1, exact
#pragma unroll
                    for (int a = 0; a < S; ++a) wx[a] *= m;
                }
                unsigned int *cell = tile + (hx * T::TY + hy) * T::TZ + hz;     // window origin (home - OFF) in tile coordinates
#pragma unroll
                for (int a = 0; a < S; ++a)
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        const float wxy = wx[a] * wy[b];
#pragma unroll
                        for (int c = 0; c < S; ++c) {
                            const unsigned int fx = (unsigned int)__float_as_int(fmaf(wxy, wz[c], magic)) - magic_bits;
                            atomicAdd(cell + (a * T::TY + b) * T::TZ + c, fx);
                        }
                    }
            }
            __syncthreads();   // every particle of the chunk is in the tile

            // ---- tile -> mesh: one coalesced 128-byte RED per (x,y) column, zeros skipped; the tile is cleared on the way ----
            const int gz = wrap_index32(bz * BrickZ<S>::CELLS - T::OFF + lane, G.N);
            const int x0 = bx * BX - T::OFF, y0 = by * BY - T::OFF;
            for (int col = warp; col < T::TX * T::TY; col += TILE_THREADS / 32) {
                const int u = col / T::TY, w = col - u * T::TY;
                const unsigned int fx = tile[col * T::TZ + lane];
                if (c1 < count) tile[col * T::TZ + lane] = 0u;
                int px = x0 + u;
                bool ok = true;
                if (G.slab) ok = px >= 0 && px < G.nplanes;
                else px = wrap_index32(px, G.N);
                if (ok && fx != 0u) {
                    const float val = MASS ? (float)(int)fx * quantum : (float)fx * quantum;
                    atomicAdd(mesh + ((long long)px * G.N + wrap_index32(y0 + w, G.N)) * G.ldz + gz, val);
                }
            }
            __syncthreads();
        }
    }
}

static size_t paged_max_bricks(const apk_plan *P) {
    return (size_t)((P->N + 3 + BX - 1) / BX) * ((P->N + BY - 1) / BY) * ((P->N + 29) / 30);
}
static size_t align256p(size_t b) { return (b + 255) & ~(size_t)255; }

// pair != 0: room for the interlaced twins' shared partition (every particle may need two copies).
// Layout: [cursor: nbricks + 1 (+ pool counter)] [page table: nbricks * PAGES_MAX] [pool: pages * PAGE payloads]
size_t deposit_paged_workspace_bytes(const apk_plan *P, long long np, int with_mass, int pair) {
    if (np <= 0) return 0;
    const size_t vs = with_mass ? sizeof(P4) : sizeof(P3);
    const size_t nb = paged_max_bricks(P);
    const size_t pages = ((size_t)np * (pair ? 2 : 1) + PAGE - 1) / PAGE + nb;
    return align256p(4 * (nb + 8)) + align256p(4 * nb * PAGES_MAX) + align256p(vs * pages * PAGE) + 256;
}

// mesh1 != nullptr: interlaced pair -- G is the shift-0 geometry, mesh1 gets the shift-0.5 twin
template <int S, typename PT, bool SOA, bool MASS, bool PAIR>
static int run_paged(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                     int mass_dtype, long long np, const DepositGeom &G, float *mesh, float *mesh1, cudaStream_t st) {
    using VT = typename std::conditional<MASS, P4, P3>::type;
    const BrickGrid B = make_brick_grid(G, S);
    DepositGeom G1 = G;
    if (PAIR) {
        G1.shift = G.shift + 0.5;
        if (G.t32 >= 0.f) G1.t32 = G.t32 + 0.5f;
    }
    const size_t need = deposit_paged_workspace_bytes(P, np, MASS, PAIR);
    APK_REQUIRE(np * (PAIR ? 2 : 1) < 0xffffffffLL, "apk_deposit: more than 2^32-1 payload slots on one device (%lld particles%s)",
                np, PAIR ? ", interlaced pair" : "");
    // the cuFFT work areas live in the last fft_work_bytes of the same workspace and may be in use on another stream
    APK_REQUIRE(P->workspace && P->workspace_bytes >= need + P->fft_work_bytes + 256,
                "apk_deposit: sorted path needs %zu workspace bytes (+ %zu of cuFFT work area), %zu set "
                "(apk_plan_workspace_bytes / apk_plan_set_workspace)", need, P->fft_work_bytes + 256, P->workspace_bytes);
    unsigned char *w = (unsigned char *)P->workspace;
    const size_t nb = paged_max_bricks(P);
    unsigned int *cursor = (unsigned int *)w; w += align256p(4 * (nb + 8));
    unsigned int *pool_next = cursor + B.nbricks + 1;
    unsigned int *table = (unsigned int *)w; w += align256p(4 * nb * PAGES_MAX);
    VT *pool = (VT *)w;

    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    const int pb = (int)std::min<long long>((np + tile - 1) / tile, (long long)P->num_sms * 8);
    P->mark(0, st);
    P->mark(1, st);
    APK_CUDA(cudaMemsetAsync(cursor, 0, 4 * (size_t)(B.nbricks + 8), st));
    APK_CUDA(cudaMemsetAsync(table, 0, 4 * (size_t)B.nbricks * PAGES_MAX, st));
    P->mark(2, st);
    brick_partition_kernel<S, PT, SOA, MASS, PAIR, VT><<<pb, PART_THREADS, 0, st>>>(
        (const PT *)p0, (const PT *)p1, (const PT *)p2, mass, mass_dtype == APK_F64, np, G, G1, B, cursor, table, pool_next, pool,
        mesh, mesh1);
    APK_CUDA(cudaGetLastError());
    P->mark(3, st);
    auto kern = brick_tile_kernel<S, MASS, VT>;
    const int ctas = B.nbricks;          // one CTA per brick; empty bricks exit at once
    kern<<<ctas, TILE_THREADS, 0, st>>>(pool, cursor, table, G, B, mesh, PAIR ? 0 : -1);
    APK_CUDA(cudaGetLastError());
    if (PAIR) {
        if (P->first_mesh_event) APK_CUDA(cudaEventRecord(P->first_mesh_event, st));
        kern<<<ctas, TILE_THREADS, 0, st>>>(pool, cursor, table, G1, B, mesh1, 1);
        APK_CUDA(cudaGetLastError());
    }
    P->mark(4, st);
    if (P->timing) { P->dep_timed = true; P->dep_sorted = true; }
    return 0;
}

template <int S, typename PT, bool SOA>
static int paged_dispatch_mass(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                               int mass_dtype, long long np, const DepositGeom &G, float *mesh, float *mesh1, cudaStream_t st) {
    if (mesh1)
        return mass ? run_paged<S, PT, SOA, true, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                    : run_paged<S, PT, SOA, false, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
    return mass ? run_paged<S, PT, SOA, true, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, nullptr, st)
                : run_paged<S, PT, SOA, false, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, nullptr, st);
}

template <int S>
static int paged_dispatch_layout(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                                 const void *mass, int mass_dtype, long long np, const DepositGeom &G, float *mesh,
                                 float *mesh1, cudaStream_t st) {
    if (pos_dtype == APK_F32)
        return layout == APK_SOA ? paged_dispatch_mass<S, float, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                                 : paged_dispatch_mass<S, float, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
    return layout == APK_SOA ? paged_dispatch_mass<S, double, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                             : paged_dispatch_mass<S, double, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
}

int deposit_atomic_launch(const void *, const void *, const void *, int, int, const void *, int, long long,
                          int, const DepositGeom &, float *, int, cudaStream_t);

// mesh1 != nullptr: also deposit the interlaced twin (shift + 0.5) from the same partition (CIC / TSC only)
int deposit_paged_launch(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                         const void *mass, int mass_dtype, long long np, int resampler, const DepositGeom &G,
                         float *mesh, float *mesh1, cudaStream_t st) {
    if (np == 0) return 0;
    if (resampler == APK_CIC) return paged_dispatch_layout<2>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, mesh1, st);
    if (resampler == APK_TSC) return paged_dispatch_layout<3>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, mesh1, st);
    // NGP has no halo and no arithmetic worth tiling: one RED per particle
    return deposit_atomic_launch(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, resampler, G, mesh, P->num_sms, st);
}

}  // namespace apk
