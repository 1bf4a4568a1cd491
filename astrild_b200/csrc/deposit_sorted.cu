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
//   1. brick_count_kernel   every particle -> key of the brick (12 x 6 x 30 home cells for TSC, 12 x 6 x 31
//                           for CIC) that holds its HOME cell; per-brick counts with one RED per warp-run of
//                           equal keys (snapshot order is spatially coherent).  float32 positions use an
//                           error-free float32 product (no FP64 issue slots), float64 positions the oracle's
//                           float64 expression; both put every particle in the oracle's cell.
//   2. brick_scan_kernel    exclusive scan of the counts -> brick_start[], cursors.
//   3. brick_scatter_kernel keys are recomputed (never stored); each run of equal keys claims its slots with ONE
//                           atomicAdd on the brick's cursor and writes its payload = brick-local coordinates as
//                           3 floats (+ mass), contiguously.  One read and one write of the particles replace a
//                           multi-pass radix sort; order inside a brick is arbitrary (the deposit does not care).
//      PAIR mode (apk_deposit_interlaced): one partition serves both interlaced meshes -- particles whose
//      two home cells fall in different bricks are filed twice, sign bits of the payload say which
//      mesh a copy is for.
//   4. brick_deposit_kernel persistent CTAs (8 warps, 3 per SM) pull bricks from a counter; per brick and per
//      chunk of <= CH particles: counting-sort the chunk by home cell inside shared memory (native 32-bit
//      ATOMS.ADD gives each particle its rank), then one thread per home cell sums the S^3 window moments of
//      its own particles in registers (packed FFMA2).  The moments are spread WITHOUT atomics and WITHOUT
//      shared memory (shared-memory float atomics are CAS loops on sm_100): every warp owns a 3 x 3 block of
//      (x,y) columns, its 32 lanes are the 32 cells of a column along z (30 home cells + 2 halo lanes for
//      TSC), so the z-spread is two warp shuffles with no edge cases and the (x,y)-spread is an add into the
//      warp's 5 x 5 (TSC) window of registers with compile-time indices.  No barrier separates the columns:
//      warps run through their 9 columns independently.  The window then goes to the mesh as one coalesced
//      128-byte RED.ADD.F32 per (x,y) column, zeros skipped; bricks are visited x-major so neighbouring
//      windows meet in L2.
#include "brick_common.cuh"
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace apk {


// PAIR: one partition serves the interlaced twins (G: shift 0, G1: shift 0.5).  Every particle is filed
// under the brick of its mesh-0 home cell; the ~11 % whose mesh-1 home cell lies in a different brick get
// a second, mesh-1-only copy there (the first copy is then flagged mesh-0-only).
template <int S, typename PT, bool SOA, bool PAIR>
__global__ void __launch_bounds__(PART_THREADS)
brick_count_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2, long long np,
                   DepositGeom G, DepositGeom G1, BrickGrid B, unsigned int *__restrict__ counts) {
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
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float l[3], l1[3];
            unsigned int key, key1;
            bool split;
            brick_keys<S, PT, PAIR>(cur.v + 3 * k, G, G1, B, key, l, key1, l1, split);
            if (first + (long long)k * PART_THREADS >= np) key = 0xffffffffu;
            int head, offset, length;
            warp_runs(key, lane, head, offset, length);
            if (key != 0xffffffffu && offset == 0) atomicAdd(counts + key, (unsigned int)length);
            if constexpr (PAIR) {
                const bool extra = key != 0xffffffffu && split;
                if (__any_sync(0xffffffffu, extra) && extra) atomicAdd(counts + key1, 1u);
            }
        }
    }
}

// exclusive scan of counts[0..n) -> start[0..n], cursor[0..n) = start, in two launches of n / SCAN_SEG CTAs:
// per-segment totals, then every CTA sums the totals below it and scans its own segment (n is 10^4..10^7)
constexpr int SCAN_PER = 8, SCAN_SEG = 1024 * SCAN_PER;

__device__ __forceinline__ unsigned int block_exclusive_scan_1024(unsigned int s, unsigned int *wsum, unsigned int &total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
        if (lane == 31) wsum[32] = wi;
    }
    __syncthreads();
    total = wsum[32];
    const unsigned int excl = wsum[warp] + incl - s;
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(1024)
brick_segsum_kernel(const unsigned int *__restrict__ counts, int n, unsigned int *__restrict__ seg_total,
                    unsigned int *__restrict__ seg_filled) {
    __shared__ unsigned int wsum[33];
    const int a = min(blockIdx.x * SCAN_SEG + threadIdx.x * SCAN_PER, n), b = min(a + SCAN_PER, n);
    unsigned int s = 0, f = 0;
    for (int i = a; i < b; ++i) { const unsigned int c = counts[i]; s += c; f += c != 0u; }
    unsigned int total, filled;
    block_exclusive_scan_1024(s, wsum, total);
    block_exclusive_scan_1024(f, wsum, filled);
    if (threadIdx.x == 0) { seg_total[blockIdx.x] = total; seg_filled[blockIdx.x] = filled; }
}

// also lists the non-empty bricks (filled[0..nfilled), ascending) so that the deposit kernel never visits an
// empty one: slab plans, halo catalogues and the particles received from other ranks leave most bricks empty
__global__ void __launch_bounds__(1024)
brick_scan_kernel(const unsigned int *__restrict__ counts, int n, const unsigned int *__restrict__ seg_total,
                  const unsigned int *__restrict__ seg_filled, unsigned int *__restrict__ start,
                  unsigned int *__restrict__ cursor, unsigned int *__restrict__ filled, unsigned int *__restrict__ nfilled) {
    __shared__ unsigned int wsum[33];
    unsigned int below = 0, fbelow = 0, base, fbase, total, ftotal;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) { below += seg_total[i]; fbelow += seg_filled[i]; }
    block_exclusive_scan_1024(below, wsum, base);       // base = sum of the segments below this one
    block_exclusive_scan_1024(fbelow, wsum, fbase);
    const int a = min(blockIdx.x * SCAN_SEG + threadIdx.x * SCAN_PER, n), b = min(a + SCAN_PER, n);
    unsigned int v[SCAN_PER], s = 0, f = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) { v[k] = a + k < b ? counts[a + k] : 0u; s += v[k]; f += v[k] != 0u; }
    unsigned int run = base + block_exclusive_scan_1024(s, wsum, total);
    unsigned int frun = fbase + block_exclusive_scan_1024(f, wsum, ftotal);
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k)
        if (a + k < b) {
            start[a + k] = run; cursor[a + k] = run; run += v[k];
            if (v[k] != 0u) filled[frun++] = (unsigned int)(a + k);
        }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) { start[n] = base + total; *nfilled = fbase + ftotal; }
}

// PAIR payload: u + 1 per axis, u = unshifted coordinate relative to the brick origin (>= -1); the sign
// of x says "not for mesh 0", the sign of y "not for mesh 1".
#ifndef APK_SCATTER_MIN_CTAS
#define APK_SCATTER_MIN_CTAS 3
#endif
template <int S, typename PT, bool SOA, bool MASS, bool PAIR, typename VT>
__global__ void __launch_bounds__(PART_THREADS, APK_SCATTER_MIN_CTAS)
brick_scatter_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2,
                     const void *__restrict__ mass, int mass_f64, long long np, DepositGeom G, DepositGeom G1,
                     BrickGrid B, unsigned int *__restrict__ cursor, VT *__restrict__ vals) {
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
        // Software pipeline of depth one: item k claims its slots (returning atomics, ~600 cycles) and item k-1,
        // whose slots have arrived meanwhile, is stored.
        VT pv = {}, pw = {};
        unsigned int pslot = 0, pslot1 = 0;
        int phead = 0, poffset = 0;
        bool plive = false, pextra = false;
#pragma unroll
        for (int k = 0; k <= 4; ++k) {
            VT v = {}, w = {};
            unsigned int slot = 0, slot1 = 0;
            int head = 0, offset = 0, length = 0;
            bool live = false, extra = false;
            if (k < 4) {
                const long long p = first + (long long)k * PART_THREADS;
                float l[3], l1[3];
                unsigned int key, key1;
                bool split;
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
                const unsigned int ps = __shfl_sync(0xffffffffu, pslot, phead) + poffset;
                if (plive) vals[ps] = pv;
                if constexpr (PAIR) {
                    if (pextra) vals[pslot1] = pw;
                }
            }
            pv = v; pw = w; pslot = slot; pslot1 = slot1; phead = head; poffset = offset; plive = live; pextra = extra;
        }
    }
}

// dynamic shared memory layout of brick_deposit_kernel
template <bool MASS>
struct DepSmem {
    static constexpr int cnt_ints = BRICK_CELLS + 1;
    static constexpr size_t bytes = sizeof(int) * (cnt_ints + 64) + sizeof(float) * CH * (MASS ? 4 : 3) + 64;
};

// Window moments of the particles of one home cell, accumulated with packed FFMA2.
// Layout: A[a][c] = float2 over the b-pair (first, last) of the window; for TSC the middle b is kept as
// B2[a] = float2 over the c-pair (first, last) and B1[a] = the centre (b = c = middle): 27 moments in 12 packed and 3
// scalar accumulators, 15 FMA instructions per particle.  The per-axis weights are formed the same way, the two outer
// ones of an axis as one float2: 0.5 (0.5 -+ d)^2 = (s/2 -+ s d)^2 with s = sqrt(1/2).
template <int S, bool MASS>
struct Moments {
    float2 A[S][S];
    float2 B2[S];     // TSC only
    float B1[S];      // TSC only

    // outer weights of one axis as a pair (first, last), and the middle one (TSC; CIC has no middle)
    __device__ __forceinline__ static void axis(float d, float2 &wp, float &wm) {
        if (S == 2) {
            wp = make_float2(1.f - d, d);
            wm = 0.f;
        } else {
            const float s = 0.70710678118654752f;
            const float2 t = __ffma2_rn(make_float2(-s, s), make_float2(d, d), make_float2(0.5f * s, 0.5f * s));
            wp = __fmul2_rn(t, t);
            wm = fmaf(-d, d, 0.75f);
        }
    }

    // FIRST: the particle sets the moments (no zeroing pass); otherwise it is added.  !valid (FIRST only): the cell
    // is empty and everything becomes zero; its coordinates are whatever the list holds at that slot, hence the selects.
    template <bool FIRST>
    __device__ __forceinline__ void put(float dx, float dy, float dz, float m, bool valid) {
        float2 wxp, wyp, wzp;
        float wxm, wym, wzm;
        axis(FIRST && !valid ? 0.f : dx, wxp, wxm);
        axis(FIRST && !valid ? 0.f : dy, wyp, wym);
        axis(FIRST && !valid ? 0.f : dz, wzp, wzm);
        if (MASS) { wxp = __fmul2_rn(wxp, make_float2(m, m)); wxm *= m; }
        if (FIRST && !valid) { wxp = make_float2(0.f, 0.f); wxm = 0.f; }
        const float wx[3] = {wxp.x, wxm, wxp.y};              // index S - 1 of CIC is .y
        const float wz[3] = {wzp.x, wzm, wzp.y};
#pragma unroll
        for (int a = 0; a < S; ++a) {
            const float w = (a == S - 1) ? wx[2] : wx[a];
            const float2 wxy = __fmul2_rn(make_float2(w, w), wyp);
#pragma unroll
            for (int c = 0; c < S; ++c) {
                const float z = (c == S - 1) ? wz[2] : wz[c];
                A[a][c] = FIRST ? __fmul2_rn(wxy, make_float2(z, z)) : __ffma2_rn(wxy, make_float2(z, z), A[a][c]);
            }
        }
        if (S == 3) {
            const float2 mp = __fmul2_rn(wxp, make_float2(wym, wym));      // (wx0 wym, wx2 wym)
            const float mm = wxm * wym;
            const float wm[3] = {mp.x, mm, mp.y};
#pragma unroll
            for (int a = 0; a < S; ++a) {
                B2[a] = FIRST ? __fmul2_rn(make_float2(wm[a], wm[a]), wzp) : __ffma2_rn(make_float2(wm[a], wm[a]), wzp, B2[a]);
                B1[a] = FIRST ? wm[a] * wzm : fmaf(wm[a], wzm, B1[a]);
            }
        }
    }

    __device__ __forceinline__ void add(float dx, float dy, float dz, float m) { put<false>(dx, dy, dz, m, true); }
    __device__ __forceinline__ void init(float dx, float dy, float dz, float m, bool valid) { put<true>(dx, dy, dz, m, valid); }

    // moment of window offset (a, b, c)
    __device__ __forceinline__ float get(int a, int b, int c) const {
        if (b == 0) return A[a][c].x;
        if (b == S - 1) return A[a][c].y;
        if (c == 0) return B2[a].x;
        if (c == S - 1) return B2[a].y;
        return B1[a];
    }
};

// Counting sort of one chunk of <= CHUNK brick-ordered particles by home cell, inside shared memory: cnt[] becomes the
// exclusive scan of the per-cell counts (cnt[cell] .. cnt[cell + 1] = that cell's particles), sx / sy / sz (/ sm)
// the brick-local coordinates in cell order.  Called by all threads of the CTA; ends with a barrier.
template <int S, bool MASS, typename VT, int CHUNK>
__device__ __forceinline__ void sort_chunk_by_cell(const VT *__restrict__ vals, unsigned int c0, int nchunk, int sel,
                                                   int *cnt, int *wsum, float *sx, float *sy, float *sz, float *sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int OFF = (S == 3) ? 1 : 0;
    constexpr int PPT = CHUNK / DEP_THREADS;                 // particles per thread per chunk
    constexpr int NBATCH = PPT % 3 == 0 ? 3 : 2, BATCH = PPT / NBATCH;   // loads are issued BATCH at a time before first use
    static_assert(BATCH * NBATCH == PPT && PPT * DEP_THREADS == CHUNK, "chunk size must divide into the load batches");
    for (int i = tid; i <= BRICK_CELLS; i += DEP_THREADS) cnt[i] = 0;
    __syncthreads();

    // ---- rank my particles inside their home cell (cell | rank << 13 kept in a register;
    //      the coordinates are re-read from L1/L2 in the scatter pass to save registers).
    //      Loads are issued in batches before first use to overlap their latency.
    int packed[PPT];
#pragma unroll
    for (int h = 0; h < NBATCH; ++h) {
        VT v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int i = (h * BATCH + k) * DEP_THREADS + tid;
            v[k] = vals[c0 + min(i, nchunk - 1)];
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int i = (h * BATCH + k) * DEP_THREADS + tid;
            const bool keep = unpack_pair(v[k], sel);
            int hx, hy, hz;
            if (S == 2) { hx = (int)floorf(v[k].x); hy = (int)floorf(v[k].y); hz = (int)floorf(v[k].z); }
            else        { hx = (int)floorf(v[k].x + 0.5f); hy = (int)floorf(v[k].y + 0.5f); hz = (int)floorf(v[k].z + 0.5f); }
            hx = max(0, min(hx, BX - 1)); hy = max(0, min(hy, BY - 1)); hz = max(0, min(hz, BrickZ<S>::CELLS - 1));
            const int cell = (hx * BY + hy) * BZ + hz + OFF;          // z-lane = home z + OFF
            packed[h * BATCH + k] = -1;
            if (i < nchunk && keep) packed[h * BATCH + k] = cell | (atomicAdd(&cnt[cell], 1) << 13);
        }
    }
    __syncthreads();

    // ---- exclusive scan of the cell counts (9 per thread) ------------------------
    {
        constexpr int PER = BRICK_CELLS / DEP_THREADS;
        static_assert(PER * DEP_THREADS == BRICK_CELLS, "cells must divide evenly among the threads");
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
    for (int h = 0; h < NBATCH; ++h) {
        VT v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int i = (h * BATCH + k) * DEP_THREADS + tid;
            v[k] = vals[c0 + min(i, nchunk - 1)];
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int pk = packed[h * BATCH + k];
            if (pk >= 0) {
                unpack_pair(v[k], sel);
                const int slot = cnt[pk & 8191] + (pk >> 13);
                sx[slot] = v[k].x; sy[slot] = v[k].y; sz[slot] = v[k].z;
                if constexpr (MASS) sm[slot] = v[k].m;
            }
        }
    }
    __syncthreads();
}

template <int S, bool MASS, typename VT>
__global__ void __launch_bounds__(DEP_THREADS, DEP_CTAS_PER_SM)
brick_deposit_kernel(const VT *__restrict__ vals, const unsigned int *__restrict__ brick_start,
                     const unsigned int *__restrict__ filled, const unsigned int *__restrict__ nfilled_ptr,
                     DepositGeom G, BrickGrid B, float *__restrict__ mesh, int sel) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *cnt = reinterpret_cast<int *>(smem_raw);                 // [BRICK_CELLS + 1]
    int *wsum = cnt + BRICK_CELLS + 1;                            // [64] scan scratch
    float *sx = reinterpret_cast<float *>(wsum + 64);
    float *sy = sx + CH;
    float *sz = sy + CH;
    float *sm = sz + CH;                                          // only if MASS

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int OFF = (S == 3) ? 1 : 0;   // window origin = home cell - OFF
    constexpr int W = 3 + S - 1;            // edge of a warp's private window: its 3 columns + halo
    // this warp's 3 x 3 block of (x,y) columns inside the brick
    const int bi = warp / (BY / 3), bj = warp % (BY / 3);

    const unsigned int nfilled = *nfilled_ptr;
    // One CTA per non-empty brick, in list order (x-major: neighbouring windows meet in L2).  CTAs retire
    // all the time, so kernels of a higher-priority stream (the slab path's FFT / transpose of the first mesh)
    // get SMs while this one runs; a persistent grid would hold every register file until it ends.
    for (unsigned int slot = blockIdx.x; slot < nfilled; slot += gridDim.x) {
        if (slot != blockIdx.x) __syncthreads();
        const unsigned int brick = filled[slot];
        const unsigned int pbeg = brick_start[brick], pend = brick_start[brick + 1];

        const int bz = brick % B.nbz;
        const int by = (brick / B.nbz) % B.nby;
        const int bx = brick / (B.nbz * B.nby);

        for (unsigned int c0 = pbeg; c0 < pend; c0 += CH) {
            const int nchunk = (int)min((unsigned int)CH, pend - c0);
            if (c0 != pbeg) __syncthreads();   // every warp is done with the previous chunk's lists
            sort_chunk_by_cell<S, MASS, VT, CH>(vals, c0, nchunk, sel, cnt, wsum, sx, sy, sz, sm);

            // ---- moments per home cell; each warp walks its own 9 columns, no CTA barrier ----
            // lane = z-cell of the column (home z + OFF), so the z-spread is two shuffles: every lane
            // gets the middle weight of its own home cell plus the outer weights of its z-neighbours
            // (halo lanes have no home cell; CIC: lane 0 must drop its wrapped-around 'up').  The
            // (x,y)-spread is a register add into the warp's window, static indices throughout.
            //      The window (one z-cell per lane) lives in registers during this phase only, so that it
            //      does not add to the register pressure of the sort phases above.  (A rolled loop over i
            //      with a sliding window of S planes has a third of the code and measured 3 % slower.)
            float R[W][W];
#pragma unroll
            for (int u = 0; u < W; ++u)
#pragma unroll
                for (int v = 0; v < W; ++v) R[u][v] = 0.f;
            const float up_on = (S == 2 && lane == 0) ? 0.f : 1.f;
            const float fz = (float)(lane - OFF);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int cx = 3 * bi + i, cy = 3 * bj + j;
                    const int cell = (cx * BY + cy) * BZ + lane;
                    const int beg = cnt[cell], end = cnt[cell + 1];
                    if (__ballot_sync(0xffffffffu, end > beg) == 0u) continue;
                    Moments<S, MASS> M;
                    const float fx = (float)cx, fy = (float)cy;
                    {
                        const bool any = end > beg;
                        const int p = any ? beg : 0;
                        M.init(sx[p] - fx, sy[p] - fy, sz[p] - fz, MASS ? (any ? sm[p] : 0.f) : 1.f, any);
                    }
                    for (int p = beg + 1; p < end; ++p)
                        M.add(sx[p] - fx, sy[p] - fy, sz[p] - fz, MASS ? sm[p] : 1.f);
#pragma unroll
                    for (int a = 0; a < S; ++a)
#pragma unroll
                        for (int b = 0; b < S; ++b) {
                            const float up = __shfl_up_sync(0xffffffffu, M.get(a, b, S - 1), 1);
                            float own;
                            if (S == 2) {
                                own = fmaf(up, up_on, M.get(a, b, 0));
                            } else {
                                const float dn = __shfl_down_sync(0xffffffffu, M.get(a, b, 0), 1);
                                own = M.get(a, b, S / 2) + up + dn;
                            }
                            R[i + a][j + b] += own;
                        }
                }
            }

            // ---- add the warp's window to the mesh: one coalesced 128-byte RED per (x,y) column ----
            const int gz = wrap_index32(bz * BrickZ<S>::CELLS - OFF + lane, G.N);
            const int x0 = bx * BX + 3 * bi - OFF, y0 = by * BY + 3 * bj - OFF;
            int row[W];                 // offset of (y0 + v, gz) inside a plane: < N * ldz, fits 32 bits
#pragma unroll
            for (int v = 0; v < W; ++v) row[v] = wrap_index32(y0 + v, G.N) * G.ldz + gz;
#pragma unroll
            for (int u = 0; u < W; ++u) {
                int px = x0 + u;
                bool ok = true;
                if (G.slab) ok = px >= 0 && px < G.nplanes;
                else px = wrap_index32(px, G.N);
                float *plane = mesh + (long long)px * G.N * G.ldz;
#pragma unroll
                for (int v = 0; v < W; ++v)
                    if (ok && R[u][v] != 0.f) atomicAdd(plane + row[v], R[u][v]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Particle-parallel tile kernel (unit masses): one THREAD per particle, the brick's window of the mesh lives in shared
// memory as 32-bit FIXED-POINT integers and every one of the S^3 weights goes there with a native integer ATOMS.ADD
// (shared-memory float atomics are CAS loops on sm_100; integer adds are not).  No in-brick sort, no per-cell loop:
// all 32 lanes work on every instruction, whatever the cell occupancy.
//   Quantum 2^-s of a particle's mass, s chosen PER CHUNK of <= PP_FLUSH particles as large as 32 bits allow if every
//   particle of the chunk put its largest possible weight (1 for CIC, 0.75^3 for TSC) into one cell: s = 22 for the
//   ~2200 particles of a brick at one particle per cell (TSC), 20..21 for a full chunk, 22 / 23 for sparse bricks.
//   A weight is rounded to the quantum by ONE FFMA against a magic constant (1.5 * 2^(23 - s): the sum lands in a
//   binade whose ulp is the quantum, so the low mantissa bits ARE the fixed-point value).  Integer adds commute: a
//   brick's contribution to the mesh does not depend on the order in which the partition filed its particles.
//   The tile goes to the mesh as before: one coalesced 128-byte RED.ADD.F32 per (x,y) column, zeros skipped.
constexpr int PP_FLUSH = 4095;                        // particles between two flushes
constexpr int PP_THREADS = 256;
#ifndef APK_PP_CTAS
#define APK_PP_CTAS 6
#endif

// fractional bits for a chunk of n particles: n * wmax * 2^s < 2^32, s <= SMAX (the magic-constant trick needs
// wmax <= 2^(22 - s))
template <int S>
__device__ __forceinline__ int pp_frac_bits(int n) {
    constexpr float WMAX = (S == 3) ? 0.43f : 1.001f;          // 0.75^3 = 0.4219 / 1, padded for the rounding
    constexpr int SMAX = (S == 3) ? 23 : 22;
    const float room = (4294967296.f / WMAX) / (float)n;
    const int s = ((__float_as_int(room) >> 23) & 0xff) - 127;  // floor(log2(room))
    return min(s, SMAX);
}

template <int S> struct PPTile {
    static constexpr int OFF = (S == 3) ? 1 : 0;      // window origin = home cell - OFF
    static constexpr int TX = BX + S - 1, TY = BY + S - 1, TZ = BZ;
    static constexpr int CELLS = TX * TY * TZ;
};

// nearest brick-local home cell of coordinate l on one axis (clamped to the brick) as float and int, and the offset
// d = l - home: [-0.5, 0.5] for TSC, [0, 1] for CIC (ties land on either side; the windows are continuous there)
template <int S>
__device__ __forceinline__ void pp_home(float l, float last, float &d, int &h) {
    const float M = 12582912.f;                       // 1.5 * 2^23: adding it rounds to the nearest integer
    float t = (S == 2 ? l - 0.5f : l) + M;
    t = fminf(fmaxf(t, M), M + last);
    h = __float_as_int(t) & 0x3fffff;
    d = l - (t - M);
}

template <int S, typename VT>
__global__ void __launch_bounds__(PP_THREADS, APK_PP_CTAS)
brick_deposit_pp_kernel(const VT *__restrict__ vals, const unsigned int *__restrict__ brick_start,
                        const unsigned int *__restrict__ filled, const unsigned int *__restrict__ nfilled_ptr,
                        DepositGeom G, BrickGrid B, float *__restrict__ mesh, int sel) {
    using T = PPTile<S>;
    __shared__ unsigned int tile[T::CELLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int nfilled = *nfilled_ptr;

    for (unsigned int slot = blockIdx.x; slot < nfilled; slot += gridDim.x) {
        const unsigned int brick = filled[slot];
        const unsigned int pbeg = brick_start[brick], pend = brick_start[brick + 1];
        const int bz = brick % B.nbz;
        const int by = (brick / B.nbz) % B.nby;
        const int bx = brick / (B.nbz * B.nby);

        for (int i = tid; i < T::CELLS; i += PP_THREADS) tile[i] = 0u;
        __syncthreads();

        for (unsigned int c0 = pbeg; c0 < pend; c0 += PP_FLUSH) {
            const unsigned int c1 = min(c0 + (unsigned int)PP_FLUSH, pend);
            const int frac_bits = pp_frac_bits<S>((int)(c1 - c0));
            const unsigned int magic_bits = ((unsigned int)(127 + 23 - frac_bits) << 23) | 0x400000u;   // 1.5 * 2^(23 - s)
            const float magic = __int_as_float((int)magic_bits);
            const float quantum = __int_as_float((127 - frac_bits) << 23);                                // 2^-s
            // ---- one thread per particle; the next particle's loads are in flight while this one is deposited ----
            unsigned int p = c0 + tid;
            VT nxt = {};
            if (p < c1) nxt = vals[p];
            for (; p < c1; p += PP_THREADS) {
                VT v = nxt;
                if (p + PP_THREADS < c1) nxt = vals[p + PP_THREADS];
                if (!unpack_pair(v, sel)) continue;
                float dx, dy, dz;
                int hx, hy, hz;
                pp_home<S>(v.x, (float)(BX - 1), dx, hx);
                pp_home<S>(v.y, (float)(BY - 1), dy, hy);
                pp_home<S>(v.z, (float)(BrickZ<S>::CELLS - 1), dz, hz);
                float wx[S], wy[S], wz[S];
                if (S == 2) {
                    wx[0] = 1.f - dx; wx[S - 1] = dx; wy[0] = 1.f - dy; wy[S - 1] = dy; wz[0] = 1.f - dz; wz[S - 1] = dz;
                } else {
                    const float ax = 0.5f - dx, cx = 0.5f + dx, ay = 0.5f - dy, cy = 0.5f + dy, az = 0.5f - dz, cz = 0.5f + dz;
                    wx[0] = 0.5f * ax * ax; wx[S / 2] = fmaf(-dx, dx, 0.75f); wx[S - 1] = 0.5f * cx * cx;
                    wy[0] = 0.5f * ay * ay; wy[S / 2] = fmaf(-dy, dy, 0.75f); wy[S - 1] = 0.5f * cy * cy;
                    wz[0] = 0.5f * az * az; wz[S / 2] = fmaf(-dz, dz, 0.75f); wz[S - 1] = 0.5f * cz * cz;
                }
                unsigned int *cell = tile + (hx * T::TY + hy) * T::TZ + hz;     // window origin (home - OFF) in tile coordinates
#pragma unroll
                for (int a = 0; a < S; ++a)
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        const float wxy = wx[a] * wy[b];
#pragma unroll
                        for (int c = 0; c < S; ++c) {
                            const unsigned int q = (unsigned int)__float_as_int(fmaf(wxy, wz[c], magic)) - magic_bits;
                            atomicAdd(cell + (a * T::TY + b) * T::TZ + c, q);
                        }
                    }
            }
            __syncthreads();   // every particle of the chunk is in the tile

            // ---- tile -> mesh: one coalesced 128-byte RED per (x,y) column, zeros skipped; the tile is cleared on the way ----
            const int gz = wrap_index32(bz * BrickZ<S>::CELLS - T::OFF + lane, G.N);
            const int x0 = bx * BX - T::OFF, y0 = by * BY - T::OFF;
            for (int col = warp; col < T::TX * T::TY; col += PP_THREADS / 32) {
                const int u = col / T::TY, w = col - u * T::TY;
                const unsigned int q = tile[col * T::TZ + lane];
                if (c1 < pend) tile[col * T::TZ + lane] = 0u;
                int px = x0 + u;
                bool ok = true;
                if (G.slab) ok = px >= 0 && px < G.nplanes;
                else px = wrap_index32(px, G.N);
                if (ok && q != 0u)
                    atomicAdd(mesh + ((long long)px * G.N + wrap_index32(y0 + w, G.N)) * G.ldz + gz, (float)q * quantum);
            }
            __syncthreads();
        }
    }
}

static size_t max_bricks(const apk_plan *P) {
    return (size_t)((P->N + 3 + BX - 1) / BX) * ((P->N + BY - 1) / BY) * ((P->N + 29) / 30);
}
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// pair != 0: room for the interlaced twins' shared partition (every particle may need two copies)
size_t deposit_sorted_workspace_bytes(const apk_plan *P, long long np, int with_mass, int pair) {
    if (np <= 0) return 0;
    const size_t vs = with_mass ? sizeof(P4) : sizeof(P3);
    return align256(vs * (size_t)np * (pair ? 2 : 1)) + 4 * align256(4 * (max_bricks(P) + 2)) +
           2 * align256(4 * (max_bricks(P) / SCAN_SEG + 2)) + 256;
}

// mesh1 != nullptr: interlaced pair -- G is the shift-0 geometry, mesh1 gets the shift-0.5 twin
template <int S, typename PT, bool SOA, bool MASS, bool PAIR>
static int run_sorted(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                      int mass_dtype, long long np, const DepositGeom &G, float *mesh, float *mesh1, cudaStream_t st) {
    using VT = typename std::conditional<MASS, P4, P3>::type;
    const BrickGrid B = make_brick_grid(G, S);
    DepositGeom G1 = G;
    if (PAIR) {
        G1.shift = G.shift + 0.5;
        if (G.t32 >= 0.f) G1.t32 = G.t32 + 0.5f;
    }
    const size_t need = deposit_sorted_workspace_bytes(P, np, MASS, PAIR);
    APK_REQUIRE(!PAIR || 2 * np < 0xffffffffLL, "apk_deposit_interlaced: more than 2^31 particles on one device");
    APK_REQUIRE(P->workspace && P->workspace_bytes >= need,
                "apk_deposit: sorted path needs %zu workspace bytes, %zu set (apk_plan_workspace_bytes / apk_plan_set_workspace)",
                need, P->workspace_bytes);
    APK_REQUIRE(np < 0xffffffffLL, "apk_deposit: more than 2^32-1 particles on one device");
    unsigned char *w = (unsigned char *)P->workspace;
    VT *vals = (VT *)w; w += align256(sizeof(VT) * (size_t)np * (PAIR ? 2 : 1));
    const size_t tab = align256(4 * (max_bricks(P) + 2));
    unsigned int *counts = (unsigned int *)w; w += tab;
    unsigned int *brick_start = (unsigned int *)w; w += tab;
    unsigned int *cursor = (unsigned int *)w; w += tab;
    unsigned int *filled = (unsigned int *)w; w += tab;
    unsigned int *seg_total = (unsigned int *)w; w += align256(4 * (max_bricks(P) / SCAN_SEG + 2));
    unsigned int *seg_filled = (unsigned int *)w; w += align256(4 * (max_bricks(P) / SCAN_SEG + 2));
    unsigned int *counter = (unsigned int *)w;          // [1] number of non-empty bricks

    const long long tile = (long long)PART_THREADS * PART_ITEMS;
    const int pb = (int)std::min<long long>((np + tile - 1) / tile, (long long)P->num_sms * 8);
    P->mark(0, st);
    APK_CUDA(cudaMemsetAsync(counts, 0, 4 * (size_t)(B.nbricks + 1), st));
    brick_count_kernel<S, PT, SOA, PAIR><<<pb, PART_THREADS, 0, st>>>((const PT *)p0, (const PT *)p1, (const PT *)p2, np, G, G1, B, counts);
    APK_CUDA(cudaGetLastError());
    P->mark(1, st);
    const int nseg = (B.nbricks + SCAN_SEG - 1) / SCAN_SEG;
    brick_segsum_kernel<<<nseg, 1024, 0, st>>>(counts, B.nbricks, seg_total, seg_filled);
    APK_CUDA(cudaGetLastError());
    brick_scan_kernel<<<nseg, 1024, 0, st>>>(counts, B.nbricks, seg_total, seg_filled, brick_start, cursor, filled, counter + 1);
    APK_CUDA(cudaGetLastError());
    P->mark(2, st);
    brick_scatter_kernel<S, PT, SOA, MASS, PAIR, VT><<<pb, PART_THREADS, 0, st>>>(
        (const PT *)p0, (const PT *)p1, (const PT *)p2, mass, mass_dtype == APK_F64, np, G, G1, B, cursor, vals);
    APK_CUDA(cudaGetLastError());

    // unit masses: particle-parallel fixed-point tile kernel; with masses: the per-cell register-moment kernel
    static const bool force_cell = [] { const char *e = getenv("APK_TILE_KERNEL"); return e && !strcmp(e, "cell"); }();
    const int ctas = B.nbricks;     // one CTA per brick (CTAs beyond the number of non-empty bricks, which only the device knows, exit at once)
    P->mark(3, st);
    if (!MASS && !force_cell) {
        if constexpr (!MASS) {
            auto kern = brick_deposit_pp_kernel<S, VT>;
            kern<<<ctas, PP_THREADS, 0, st>>>(vals, brick_start, filled, counter + 1, G, B, mesh, PAIR ? 0 : -1);
            APK_CUDA(cudaGetLastError());
            if (PAIR) {
                if (P->first_mesh_event) APK_CUDA(cudaEventRecord(P->first_mesh_event, st));
                kern<<<ctas, PP_THREADS, 0, st>>>(vals, brick_start, filled, counter + 1, G1, B, mesh1, 1);
                APK_CUDA(cudaGetLastError());
            }
        }
    } else {
        auto kern = brick_deposit_kernel<S, MASS, VT>;
        const size_t smem = DepSmem<MASS>::bytes;
        APK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ctas, DEP_THREADS, smem, st>>>(vals, brick_start, filled, counter + 1, G, B, mesh, PAIR ? 0 : -1);
        APK_CUDA(cudaGetLastError());
        if (PAIR) {
            if (P->first_mesh_event) APK_CUDA(cudaEventRecord(P->first_mesh_event, st));
            kern<<<ctas, DEP_THREADS, smem, st>>>(vals, brick_start, filled, counter + 1, G1, B, mesh1, 1);
            APK_CUDA(cudaGetLastError());
        }
    }
    P->mark(4, st);
    if (P->timing) { P->dep_timed = true; P->dep_sorted = true; }
    return 0;
}

template <int S, typename PT, bool SOA>
static int dispatch_mass(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass,
                         int mass_dtype, long long np, const DepositGeom &G, float *mesh, float *mesh1, cudaStream_t st) {
    if (mesh1)
        return mass ? run_sorted<S, PT, SOA, true, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                    : run_sorted<S, PT, SOA, false, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
    return mass ? run_sorted<S, PT, SOA, true, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, nullptr, st)
                : run_sorted<S, PT, SOA, false, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, nullptr, st);
}

template <int S>
static int dispatch_layout(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                           const void *mass, int mass_dtype, long long np, const DepositGeom &G, float *mesh,
                           float *mesh1, cudaStream_t st) {
    if (pos_dtype == APK_F32)
        return layout == APK_SOA ? dispatch_mass<S, float, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                                 : dispatch_mass<S, float, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
    return layout == APK_SOA ? dispatch_mass<S, double, true>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st)
                             : dispatch_mass<S, double, false>(P, p0, p1, p2, mass, mass_dtype, np, G, mesh, mesh1, st);
}

int deposit_atomic_launch(const void *, const void *, const void *, int, int, const void *, int, long long,
                          int, const DepositGeom &, float *, int, cudaStream_t);

// mesh1 != nullptr: also deposit the interlaced twin (shift + 0.5) from the same partition (CIC / TSC only)
int deposit_sorted_launch(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                          const void *mass, int mass_dtype, long long np, int resampler, const DepositGeom &G,
                          float *mesh, float *mesh1, cudaStream_t st) {
    if (np == 0) return 0;
    if (resampler == APK_CIC) return dispatch_layout<2>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, mesh1, st);
    if (resampler == APK_TSC) return dispatch_layout<3>(P, p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, G, mesh, mesh1, st);
    // NGP has no halo and no arithmetic worth tiling: one RED per particle
    return deposit_atomic_launch(p0, p1, p2, layout, pos_dtype, mass, mass_dtype, np, resampler, G, mesh, P->num_sms, st);
}

}  // namespace apk
