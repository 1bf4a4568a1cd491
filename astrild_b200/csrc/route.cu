// Slab routing of particles and ghost-plane accumulation for the multi-GPU path.
//
// The reference has no multi-rank path (astrild runs pmesh/nbodykit on one rank;
// /root/reference/src/astrild/particles/hutils/stats_subfind.py:130 builds the ParticleMesh
// without a communicator).  pmesh's own decomposition (domain.GridND + exchange) is what this
// replaces: every particle goes to the rank whose x-slab holds floor(g_x), g = pos*pos_scale*N,
// using the same float64 index arithmetic as the deposit so a particle never lands outside the
// ghost planes of the slab it was routed to.
//
// Bound: HBM; bytes = np * 3 * sizeof(pos) read + written (+ mass).  Snapshot order is spatially
// coherent, so per-destination counters are updated once per warp-run of equal destinations.
#include "apk_common.cuh"
#include "deposit_common.cuh"
#include <algorithm>
#include <cstdint>

namespace apk {

__device__ __forceinline__ void warp_runs_r(int key, int lane, int &head, int &offset, int &length) {
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned int heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    const unsigned int below = heads & ((2u << lane) - 1u);
    head = 31 - __clz(below);
    offset = lane - head;
    const unsigned int above = heads & ~((2u << lane) - 1u);
    length = (above ? __ffs(above) - 1 : 32) - head;
}

struct RouteGeom {
    double scale;        // grid units per position unit
    float s0, s1, s2;    // scale as three floats (error-free float32 product, see deposit_sorted.cu)
    int N, ppr, self, x0;
};

// destination rank of a particle, or -1 if it stays (the common case costs no integer division)
template <typename PT>
__device__ __forceinline__ int dest_rank(const PT x, const RouteGeom &R) {
    int cell;
    bool far = false;
    if constexpr (sizeof(PT) == 4) {
        const float pr = __fmul_rn(x, R.s0);
        // well inside the slab (the rounded product is within ~1e-4 cells of the exact one): stays, no more work
        if (pr > (float)R.x0 + 0.01f && pr < (float)(R.x0 + R.ppr) - 0.01f) return -1;
        float e = fmaf(x, R.s0, -pr);
        e = fmaf(x, R.s1, e);
        e = fmaf(x, R.s2, e);
        const float h = floorf(pr);
        const float f = __fadd_rn(__fsub_rn(pr, h), e);
        cell = wrap_near((int)(h + floorf(f)), R.N, far);
        far |= !(fabsf(pr) < 1.0e9f);
    } else {
        far = true;
    }
    if (far) cell = wrap_index((long long)floor((double)x * R.scale), R.N);
    const int rel = cell - R.x0;
    if (rel >= 0 && rel < R.ppr) return -1;
    return cell / R.ppr;
}

// pass 1 (the only pass over all particles; reads x alone): particles that leave this slab are counted per
// destination and staged in arrival order; the ones that stay are not touched (the slab deposit skips
// particles it does not own, so the caller deposits its arrays as they are)
template <typename PT, bool SOA, typename MT>
__global__ void __launch_bounds__(256)
route_stage_kernel(const PT *__restrict__ p0, const PT *__restrict__ p1, const PT *__restrict__ p2,
                   const MT *__restrict__ mass, long long np, RouteGeom R, unsigned long long *__restrict__ counts,
                   unsigned long long *__restrict__ total, long long capacity, PT *__restrict__ stage_pos,
                   MT *__restrict__ stage_mass) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long np_pad = (np + 31) & ~31LL;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np_pad; p += 4 * stride) {
        PT x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                               // 4 independent loads in flight
            const long long q = min(p + k * stride, np - 1);
            x[k] = SOA ? __ldcs(p0 + q) : __ldcs(p0 + 3 * q);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long q = p + k * stride;
            if (q >= np_pad) break;                                  // warp-uniform
            const int d = q < np ? dest_rank<PT>(x[k], R) : -1;
            const unsigned int leaving = __ballot_sync(0xffffffffu, d >= 0);
            if (leaving == 0u) continue;                             // the usual case: nobody leaves
            int head, offset, length;
            warp_runs_r(d, lane, head, offset, length);
            if (d >= 0 && offset == 0) atomicAdd(counts + d, (unsigned long long)length);
            unsigned long long slot = 0;
            if (lane == __ffs(leaving) - 1) slot = atomicAdd(total, (unsigned long long)__popc(leaving));
            slot = __shfl_sync(0xffffffffu, slot, __ffs(leaving) - 1) + __popc(leaving & ((1u << lane) - 1u));
            if (d >= 0 && (long long)slot < capacity) {
                PT y, z;
                if (SOA) { y = p1[q]; z = p2[q]; }
                else     { y = p0[3 * q + 1]; z = p0[3 * q + 2]; }
                stage_pos[3 * slot] = x[k]; stage_pos[3 * slot + 1] = y; stage_pos[3 * slot + 2] = z;
                if (mass) stage_mass[slot] = mass[q];
            }
        }
    }
}

// counts[P] -> cursor[P] = exclusive scan (P <= 1024)
__global__ void route_scan_kernel(const unsigned long long *counts, int P, unsigned long long *cursor) {
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < P; ++i) { cursor[i] = run; run += counts[i]; }
    }
}

// pass 2 (over the staged leavers only): group them by destination rank
template <typename PT, typename MT>
__global__ void __launch_bounds__(256)
route_group_kernel(const PT *__restrict__ stage_pos, const MT *__restrict__ stage_mass, const unsigned long long *__restrict__ total,
                   long long capacity, RouteGeom R, unsigned long long *__restrict__ cursor, PT *__restrict__ out_pos,
                   MT *__restrict__ out_mass) {
    // more leavers than the staging buffer holds: the caller repeats the pass with a larger one (it reads the counts);
    // grouping the staged part would write past out_pos, because the cursors come from the FULL counts
    if ((long long)*total > capacity) return;
    const long long n = (long long)*total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_pad = (n + 31) & ~31LL;                        // warp-uniform trip count
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        PT x = 0;
        int d = -1;
        if (i < n) { x = stage_pos[3 * i]; d = dest_rank<PT>(x, R); }   // d < 0 cannot happen: staged particles leave
        // one atomic per destination and warp: the leavers of a slab go to very few ranks (its two neighbours,
        // as a rule), so per-particle atomics would all hit the same two addresses
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader && d >= 0) base = atomicAdd(cursor + d, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (d < 0) continue;
        const unsigned long long slot = base + __popc(peers & ((1u << lane) - 1u));
        out_pos[3 * slot] = x; out_pos[3 * slot + 1] = stage_pos[3 * i + 1]; out_pos[3 * slot + 2] = stage_pos[3 * i + 2];
        if (stage_mass) out_mass[slot] = stage_mass[i];
    }
}

template <typename PT, bool SOA, typename MT>
static int route_typed(apk_plan *P, const void *p0, const void *p1, const void *p2, const void *mass, long long np,
                       double pos_scale, int nranks, unsigned long long *counts, long long capacity, void *out_pos,
                       void *out_mass, cudaStream_t st) {
    RouteGeom R;
    R.N = P->N; R.ppr = P->N / nranks; R.self = P->x0 / R.ppr; R.x0 = P->x0;
    R.scale = pos_scale * (double)P->N;
    R.s0 = (float)R.scale;
    R.s1 = (float)(R.scale - (double)R.s0);
    R.s2 = (float)(R.scale - (double)R.s0 - (double)R.s1);
    unsigned long long *cursor = counts + nranks;
    // staging lives in the plan workspace (the deposit that follows reuses it)
    const size_t need = (size_t)capacity * (3 * sizeof(PT) + (mass ? sizeof(MT) : 0)) + 64;
    APK_REQUIRE(capacity == 0 || (P->workspace && P->workspace_bytes >= need),
                "apk_route_particles: %zu workspace bytes needed for staging, %zu set", need, P->workspace_bytes);
    unsigned long long *total = (unsigned long long *)P->workspace;
    PT *stage_pos = (PT *)((char *)P->workspace + 64);
    MT *stage_mass = mass ? (MT *)(stage_pos + 3 * (size_t)capacity) : nullptr;
    APK_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * nranks, st));
    if (np > 0 && capacity > 0) {
        APK_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned long long), st));
        const int blocks = (int)std::min<long long>((np + 1023) / 1024, (long long)P->num_sms * 8);
        route_stage_kernel<PT, SOA, MT><<<blocks, 256, 0, st>>>((const PT *)p0, (const PT *)p1, (const PT *)p2, (const MT *)mass,
                                                               np, R, counts, total, capacity, stage_pos, stage_mass);
        APK_CUDA(cudaGetLastError());
        route_scan_kernel<<<1, 32, 0, st>>>(counts, nranks, cursor);
        APK_CUDA(cudaGetLastError());
        const int gblocks = (int)std::min<long long>((capacity + 255) / 256, (long long)P->num_sms * 4);
        route_group_kernel<PT, MT><<<gblocks, 256, 0, st>>>(stage_pos, stage_mass, total, capacity, R, cursor,
                                                            (PT *)out_pos, (MT *)out_mass);
        APK_CUDA(cudaGetLastError());
    }
    return 0;
}

int route_launch(apk_plan *P, const void *p0, const void *p1, const void *p2, int layout, int pos_dtype,
                 double pos_scale, const void *mass, int mass_dtype, long long np, int nranks,
                 unsigned long long *counts, long long capacity, void *out_pos, void *out_mass, cudaStream_t st) {
    APK_REQUIRE(nranks >= 1 && nranks <= 1024 && P->N % nranks == 0, "apk_route_particles: nmesh %d not divisible by %d ranks", P->N, nranks);
    APK_REQUIRE(P->n0 == P->N / nranks && P->x0 % P->n0 == 0, "apk_route_particles: plan slab [%d,%d) is not rank-aligned for %d ranks", P->x0, P->x0 + P->n0, nranks);
    APK_REQUIRE(!mass || mass_dtype == pos_dtype, "apk_route_particles: mass must have the dtype of the positions");
    const bool soa = layout == APK_SOA;
    if (pos_dtype == APK_F32)
        return soa ? route_typed<float, true, float>(P, p0, p1, p2, mass, np, pos_scale, nranks, counts, capacity, out_pos, out_mass, st)
                   : route_typed<float, false, float>(P, p0, p1, p2, mass, np, pos_scale, nranks, counts, capacity, out_pos, out_mass, st);
    return soa ? route_typed<double, true, double>(P, p0, p1, p2, mass, np, pos_scale, nranks, counts, capacity, out_pos, out_mass, st)
               : route_typed<double, false, double>(P, p0, p1, p2, mass, np, pos_scale, nranks, counts, capacity, out_pos, out_mass, st);
}

// ---- fused pack + peer store of the x <-> y slab transpose --------------------------------------------------
// grid is complex64 [n0][N][nz] (x-slab of this rank).  For x_local and destination rank s, the ny*nz complex of
// rows y in [s*ny, (s+1)*ny) are contiguous in the source AND in rank s's receive buffer [N][ny][nz] at
// [x0 + x_local][0][0], so the transpose is n0*P contiguous chunk copies; they are written straight into the
// peers' memory over NVLink (no pack pass, no NCCL staging).  One CTA works on one chunk at a time.
template <typename V>
__global__ void __launch_bounds__(256)
transpose_p2p_kernel(const V *__restrict__ grid, const unsigned long long *__restrict__ peer_base, long long peer_off_bytes,
                     int n0, int x0, int P, long long chunk_v /* V elements per chunk */, int rot) {
    const long long nchunks = (long long)n0 * P;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int xl = (int)(c / P);
        const int s = (int)((c % P + rot) % P);                       // rank r starts with peer r: spreads the links
        const V *src = grid + ((long long)xl * P + s) * chunk_v;
        V *dst = reinterpret_cast<V *>(peer_base[s] + peer_off_bytes) + (long long)(x0 + xl) * chunk_v;
        for (long long i = threadIdx.x; i < chunk_v; i += 4 * 256) {
            V v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k * 256 < chunk_v) v[k] = __ldcs(src + i + k * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k * 256 < chunk_v) dst[i + k * 256] = v[k];
        }
    }
}

int transpose_p2p_launch(apk_plan *P, const void *grid, const unsigned long long *peer_base, long long peer_off_bytes,
                         int nranks, cudaStream_t st) {
    APK_REQUIRE(nranks >= 1 && P->N % nranks == 0 && P->n0 == P->N / nranks, "apk_slab_transpose_p2p: plan is not a 1/%d slab", nranks);
    const long long chunk_bytes = (long long)(P->N / nranks) * P->Nk * 8;
    const int rot = P->x0 / P->n0;
    const int blocks = P->num_sms * 4;
    if (chunk_bytes % 16 == 0 && ((uintptr_t)grid % 16) == 0 && peer_off_bytes % 16 == 0)
        transpose_p2p_kernel<float4><<<blocks, 256, 0, st>>>((const float4 *)grid, peer_base, peer_off_bytes, P->n0, P->x0, nranks, chunk_bytes / 16, rot);
    else
        transpose_p2p_kernel<float2><<<blocks, 256, 0, st>>>((const float2 *)grid, peer_base, peer_off_bytes, P->n0, P->x0, nranks, chunk_bytes / 8, rot);
    APK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(256) accumulate_kernel(float *__restrict__ dst, const float *__restrict__ src, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

int accumulate_launch(apk_plan *P, float *dst, const float *src, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)P->num_sms * 8);
    accumulate_kernel<<<blocks, 256, 0, st>>>(dst, src, n);
    APK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace apk
