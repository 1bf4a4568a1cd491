// Brick geometry, brick keys and payload conventions shared by the partition and tile kernels of the sorted deposit
// (deposit_sorted.cu).
#pragma once
#include "apk_common.cuh"
#include "deposit_common.cuh"
#include <algorithm>
#include <cstdint>
#include <type_traits>

namespace apk {

#ifndef APK_BX
#define APK_BX 12
#endif
#ifndef APK_BY
#define APK_BY 6
#endif
constexpr int BX = APK_BX, BY = APK_BY;         // brick edge in cells along x, y
constexpr int BZ = 32;                          // z-lanes of a column = tile cells along z (one warp)
// home cells along z: 32 - (S-1), so that a column's 32 lanes are exactly its 32 tile cells
// (30 home cells + 2 halo lanes for TSC, 31 + 1 for CIC) and the spread needs no halo special case
template <int S> struct BrickZ { static constexpr int CELLS = BZ - (S - 1); };
struct P3 { float x, y, z; };
struct P4 { float x, y, z, m; };

struct BrickGrid {
    int nbx, nby, nbz;   // bricks per axis (x counts local planes for slab plans)
    int nbricks;
};

static inline BrickGrid make_brick_grid(const DepositGeom &G, int S) {
    BrickGrid B;
    const int bzc = BZ - (S - 1);
    B.nbx = (G.nplanes + BX - 1) / BX;
    B.nby = (G.N + BY - 1) / BY;
    B.nbz = (G.N + bzc - 1) / bzc;
    B.nbricks = B.nbx * B.nby * B.nbz;
    return B;
}

// home cell of a particle on one axis: floor(g) for CIC, floor(g + 0.5) for TSC / NGP

// floor(g) as double and as int32 without 64-bit conversion instructions (quarter-rate pipe):
// adding 1.5*2^52 leaves rint(g) in the low word.  Valid for |g| < 2^31 grid units.
__device__ __forceinline__ double floor_magic(double g, int &i) {
    const double magic = 6755399441055744.0;
    const double t = g + magic;
    int r = __double2loint(t);
    double rd = t - magic;
    if (rd > g) { rd -= 1.0; r -= 1; }
    i = r;
    return rd;
}

// brick key and brick-local coordinates of particle p (shared by the count and scatter passes so
// both see bit-identical keys).  g = x * scale + shift in float64, exactly the oracle's expression.
template <int S>
__device__ __forceinline__ unsigned int brick_of(const double (&x)[3], const DepositGeom &G,
                                                 const BrickGrid &B, float (&l)[3]) {
    if (!owned_by_slab(x[0] * G.scale, G)) return 0xffffffffu;         // slab plans: not this rank's particle
    int b[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double g = x[d] * G.scale + G.shift;
        int hi;
        const double h = floor_magic(S == 2 ? g : g + 0.5, hi);         // home cell
        const float frac = (float)(g - h);                              // [0,1) CIC, [-0.5,0.5) TSC
        int hl = wrap_index32(hi, G.N);
        if (d == 0 && G.slab) {
            hl -= G.plane0;
            if (hl < 0) hl += G.N; else if (hl >= G.N) hl -= G.N;
            if (hl >= G.nplanes) hl = 0;                                // unreachable for owned particles
        }
        const int edge = d == 0 ? BX : (d == 1 ? BY : BrickZ<S>::CELLS);
        b[d] = hl / edge;
        l[d] = frac + (float)(hl - b[d] * edge);
    }
    return (unsigned int)((b[0] * B.nby + b[1]) * B.nbz + b[2]);
}

// The same for float32 positions without any float64 instruction (the partition kernels are
// issue-bound and FP64 issues at half rate).  g = x * scale is carried as p + e with
// p = fl(x*s0) and e = the exact rounding error of p plus x*(s1 + s2): an error-free product good
// to ~2^-70, so floor() and the in-cell fraction agree with the float64 expression except for
// products within ~1e-13 of an integer, where the window weights are continuous anyway.
// float32 index path in two steps so that the interlaced twins share the first one:
//   f32_base   : floor and fraction of the UNSHIFTED coordinate per axis (+ slab ownership)
//   f32_finish : home cell / brick / brick-local coordinate for one mesh (shift folded into t32)
struct AxisBase { float h[3], f[3]; bool owned; };

__device__ __forceinline__ AxisBase f32_base(const float *x, const DepositGeom &G, bool &far) {
    AxisBase a;
    a.owned = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float p = __fmul_rn(x[d], G.s0);      // intrinsics: never contracted into an FMA
        float e = fmaf(x[d], G.s0, -p);
        e = fmaf(x[d], G.s1, e);
        e = fmaf(x[d], G.s2, e);
        a.h[d] = floorf(p);
        a.f[d] = __fadd_rn(__fsub_rn(p, a.h[d]), e);   // p - h is exact; f in [e, 1 + e)
        if (d == 0 && G.slab) {                      // ownership: floor of the UNSHIFTED coordinate
            const int hu = (int)(a.h[0] + floorf(a.f[0]));
            const int rel = wrap_near(hu, G.N, far) - G.own0;
            a.owned = rel >= 0 && rel < G.nown;
        }
    }
    return a;
}

// Home cell, brick and brick-local coordinate of mesh 0 (shift folded into t32) and, if PAIR, of its interlaced
// twin half a cell further.  Everything stays in float32 registers -- cell and brick indices are small integers,
// exact in float -- so one axis costs ~25 FP32 instructions and no integer division; the brick index is
// floor((cell + 0.5) / edge), whose argument is never closer than 1/(2 edge) to an integer.  The twin's home
// cell is the same cell or the next one along each axis, so its brick and brick-local coordinate follow from
// mesh 0's with compares instead of a second floor / wrap / divide.  Needs nbricks < 2^24 (else: float64 path).
template <int S, bool PAIR>
__device__ __forceinline__ void f32_finish(const AxisBase &a, float t32, const DepositGeom &G, const BrickGrid &B,
                                           unsigned int &key0, float (&l0)[3], unsigned int &key1, float (&l1)[3],
                                           bool &far, bool &split) {
    split = false;
    if (!a.owned) { key0 = key1 = 0xffffffffu; return; }
    const float Nf = (float)G.N;
    float b0[3], b1[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float f = a.f[d] + t32;                      // + shift (+ 0.5: TSC rounds to the nearest cell)
        const float c = floorf(f);
        f -= c;                                      // [0, 1)
        float h = a.h[d] + c;                        // home cell, within one box length of the box
        h = h >= Nf ? h - Nf : h;
        h = h < 0.f ? h + Nf : h;
        far |= !(h >= 0.f && h < Nf);
        float lim = Nf;
        if (d == 0 && G.slab) {                      // local plane index of a slab plan
            h -= (float)G.plane0;
            h = h < 0.f ? h + Nf : h;
            h = h >= Nf ? h - Nf : h;
            lim = (float)G.nplanes;
            h = h >= lim ? 0.f : h;                  // unreachable for owned particles
        }
        const float edge = d == 0 ? (float)BX : (d == 1 ? (float)BY : (float)BrickZ<S>::CELLS);
        const float b = floorf(fmaf(h, 1.f / edge, 0.5f / edge));
        const float loc = fmaf(-edge, b, h);         // cell inside the brick
        b0[d] = b;
        l0[d] = ((S == 2) ? f : f - 0.5f) + loc;
        if (PAIR) {
            const bool up = f >= 0.5f;               // the twin's home cell is the next one
            const float one = up ? 1.f : 0.f;
            float loc1 = loc + one, bb = b;
            if (loc1 == edge) { loc1 = 0.f; bb = b + 1.f; }
            if (h + one >= lim) { loc1 = 0.f; bb = 0.f; split = true; }   // periodic wrap (slab planes: unreachable)
            b1[d] = bb;
            l1[d] = ((S == 2) ? f - 0.5f * one + 0.5f * (1.f - one) : f - one) + loc1;
        }
    }
    key0 = (unsigned int)(int)fmaf(fmaf(b0[0], (float)B.nby, b0[1]), (float)B.nbz, b0[2]);
    key1 = key0;
    if (PAIR) {
        key1 = (unsigned int)(int)fmaf(fmaf(b1[0], (float)B.nby, b1[1]), (float)B.nbz, b1[2]);
        split |= key1 != key0;
    }
}

// keys and brick-local coordinates of one particle for mesh 0 (G) and, if PAIR, its interlaced twin (G1).
// split: the twin needs a copy of its own -- its home cell lies in another brick, or in the SAME brick but across the
// periodic boundary (an axis covered by a single brick: N <= 12 / 6 / 30), so that its brick-local coordinate is not
// mesh 0's plus half a cell.
template <int S, typename PT, bool PAIR>
__device__ __forceinline__ void brick_keys(const PT *x, const DepositGeom &G, const DepositGeom &G1, const BrickGrid &B,
                                           unsigned int &key0, float (&l0)[3], unsigned int &key1, float (&l1)[3],
                                           bool &split) {
    if constexpr (std::is_same<PT, float>::value) {
        if (G.t32 >= 0.f && (!PAIR || G1.t32 >= 0.f) && B.nbricks < (1 << 24)) {
            bool far = false;                        // position more than a box length outside the box (rare)
            const AxisBase a = f32_base(x, G, far);
            f32_finish<S, PAIR>(a, G.t32, G, B, key0, l0, key1, l1, far, split);
            if (!far) return;
        }
    }
    const double xd[3] = {(double)x[0], (double)x[1], (double)x[2]};
    key0 = brick_of<S>(xd, G, B, l0);
    key1 = key0;
    split = false;
    if (PAIR) {
        key1 = brick_of<S>(xd, G1, B, l1);
        split = key1 != key0;
#pragma unroll
        for (int d = 0; d < 3; ++d) split |= fabsf(l1[d] - l0[d] - 0.5f) > 0.25f;
    }
}

// raw coordinates of this thread's 4 particles of a tile, v[3*k + d]: slice k of the tile is the
// PART_THREADS consecutive particles base + k*PART_THREADS + tid, so the 32 lanes of a warp always hold 32
// CONSECUTIVE particles (long runs of equal brick keys, coalesced 4-byte loads).  All 12 loads are
// issued up front; indices are clamped to np-1 and masked by the caller.
template <typename PT> struct Raw4 { PT v[12]; };

template <typename PT, bool SOA>
__device__ __forceinline__ void load4(const PT *__restrict__ p0, const PT *__restrict__ p1,
                                      const PT *__restrict__ p2, long long first, int stride, long long np,
                                      Raw4<PT> &r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long p = min(first + (long long)k * stride, np - 1);
        if (SOA) { r.v[3 * k] = __ldcs(p0 + p); r.v[3 * k + 1] = __ldcs(p1 + p); r.v[3 * k + 2] = __ldcs(p2 + p); }
        else     { r.v[3 * k] = __ldcs(p0 + 3 * p); r.v[3 * k + 1] = __ldcs(p0 + 3 * p + 1); r.v[3 * k + 2] = __ldcs(p0 + 3 * p + 2); }
    }
}

// run-length aggregation inside a warp: lanes holding the same key as their left neighbour form a
// run; returns the run's head lane and this lane's offset in it (keys of a snapshot are coherent,
// so runs are long; for random keys every lane is its own run)
__device__ __forceinline__ void warp_runs(unsigned int key, int lane, int &head, int &offset, int &length) {
    const unsigned int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned int heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    const unsigned int below = heads & ((2u << lane) - 1u);          // heads at or below this lane
    head = 31 - __clz(below);
    offset = lane - head;
    const unsigned int above = heads & ~((2u << lane) - 1u);          // heads above this lane
    const int next = above ? __ffs(above) - 1 : 32;
    length = next - head;
}

constexpr int PART_THREADS = 256;
constexpr int PART_ITEMS = 4;    // consecutive particles per thread per tile

// sel < 0: plain payload (brick-local coordinates of this mesh).  sel = 0 / 1: PAIR payload, decoded
// into the coordinates of mesh `sel`; returns false if the copy is not meant for this mesh.
template <typename VT>
__device__ __forceinline__ bool unpack_pair(VT &v, int sel) {
    if (sel < 0) return true;
    const bool skip = __float_as_int(sel == 0 ? v.x : v.y) < 0;
    const float off = sel ? -0.5f : -1.f;            // - 1 (stored offset) + 0.5 * sel (mesh shift)
    v.x = fabsf(v.x) + off; v.y = fabsf(v.y) + off; v.z = fabsf(v.z) + off;
    return !skip;
}

}  // namespace apk
