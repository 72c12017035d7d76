// fk.cu — RHD 21-joint forward-kinematics layer + pinhole projection, forward and backward.
// One thread per sample (59 floats in, 105 out — no cross-sample data).  A warp owns 32 consecutive samples and private
// shared-memory tiles laid out exactly like the global arrays (dense rows), so a tile moves with ONE bulk copy (TMA engine) per
// array and direction: no per-element staging loops, no block barriers; the next tile's inputs are requested as soon as this
// tile's results have been handed to the copy engine (its stores only read the result tiles).  [round 1 ncu: the cp.async / st.global staging loops of the 64-thread-block version were ~37 % of the
// kernel's instructions at 46 % issue utilisation and 14 % occupancy.]  Dense pitches cost a 4-way bank conflict on the 20
// bone-length reads and a 2-way one on the 42 uv writes of a sample; every other row pitch (3, 23, 9, 1, 63) is odd.
// The last (partial) tile of a batch and unaligned pointers take plain per-element loops.  Math in fk_math.cuh (reference
// citations there).
#include "fk_math.cuh"
#include "ptx.cuh"
#include "tc_ptx.cuh"

namespace mb {
namespace {

constexpr int FK_WARPS = 2;                      // warps per block, each with its own tiles: 5 blocks (10 warps) per SM by shared memory
constexpr int FK_THREADS = FK_WARPS * 32;
constexpr int FK_T = 32;                         // samples per tile

struct alignas(128) FkTiles {                    // every array is a multiple of 16 bytes: all of them stay 16-byte aligned
    float ra[FK_T * 3], oa[FK_T * FK_OA], bl[FK_T * FK_NODES], K[FK_T * 9], s[FK_T], root[FK_T * 3];   // 7 552 B in
    float xyz[FK_T * 63], uv[FK_T * 42];         // 13 440 B: forward results / backward upstream gradients
    unsigned long long bar;
};
constexpr uint32_t FK_IN_BYTES = FK_T * 59 * 4, FK_XYZ_BYTES = FK_T * 63 * 4, FK_UV_BYTES = FK_T * 42 * 4;

__device__ __forceinline__ void fk_bar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();          // a lost copy must not hang the GPU
}
// per-element paths (partial tile / unaligned pointers): n floats, coalesced
__device__ __forceinline__ void fk_copy_in(float* dst, const float* __restrict__ src, int n, int lane) {
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}
__device__ __forceinline__ void fk_copy_out(float* __restrict__ dst, const float* src, int n, int lane) {
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

struct FkIn { const float *ra, *oa, *bl, *K, *sc, *root; };

// request the six input arrays of a tile; `lo` = false leaves out ra / oa / bl (the backward's result tiles, requested once
// their stores have been read).  Bulk path: one lane, byte counts land on the tile's mbarrier.
__device__ __forceinline__ void fk_request_inputs(FkTiles& T, const FkIn& in, long long row0, uint32_t bar, bool lo, bool hi) {
    if (hi) {
        bulk_g2s(smem_u32(T.K), in.K + row0 * 9, FK_T * 9 * 4, bar);
        bulk_g2s(smem_u32(T.s), in.sc + row0, FK_T * 4, bar);
        bulk_g2s(smem_u32(T.root), in.root + row0 * 3, FK_T * 3 * 4, bar);
    }
    if (lo) {
        bulk_g2s(smem_u32(T.ra), in.ra + row0 * 3, FK_T * 3 * 4, bar);
        bulk_g2s(smem_u32(T.oa), in.oa + row0 * FK_OA, FK_T * FK_OA * 4, bar);
        bulk_g2s(smem_u32(T.bl), in.bl + row0 * FK_NODES, FK_T * FK_NODES * 4, bar);
    }
}
__device__ __forceinline__ void fk_copy_inputs(FkTiles& T, const FkIn& in, long long row0, int rows, int lane) {
    fk_copy_in(T.ra, in.ra + row0 * 3, rows * 3, lane);
    fk_copy_in(T.oa, in.oa + row0 * FK_OA, rows * FK_OA, lane);
    fk_copy_in(T.bl, in.bl + row0 * FK_NODES, rows * FK_NODES, lane);
    fk_copy_in(T.K, in.K + row0 * 9, rows * 9, lane);
    fk_copy_in(T.s, in.sc + row0, rows, lane);
    fk_copy_in(T.root, in.root + row0 * 3, rows * 3, lane);
}

// the two masked-mean L2 terms of the fused forward (mb_fk_loss_forward)
struct FkLoss {
    const float *gt_xyz, *gt_uv, *vis;      // [B][21][3], [B][21][2], [B][21]; a NULL ground truth switches its term off
    double* accum;                          // {S_xyz, N_xyz, S_uv, N_uv, warp ticket}, zeroed before the launch
    float* losses;                          // [2]
};

// a sample's 21 visibility flags as a bit mask: 21 independent 4-byte loads per thread (a warp's rows are one contiguous
// 2 688-byte block), issued BEFORE the wait for the tile so that their latency hides behind the bulk copies
__device__ __forceinline__ unsigned fk_vismask(const float* __restrict__ vis_row) {
    float v[21];
#pragma unroll
    for (int j = 0; j < 21; ++j) v[j] = __ldg(vis_row + j);
    unsigned m = 0u;
#pragma unroll
    for (int j = 0; j < 21; ++j) m |= (v[j] != 0.f ? 1u : 0u) << j;
    return m;
}

// LOSS: the ground-truth rows are bulk-copied INTO the result tiles; the thread that computes a keypoint takes the squared
// distance to the ground truth in its slot (visible joints; fp32 over the coordinates, fp64 above — criterions/loss.py:10-25)
// and then overwrites the slot, so the reductions cost no shared memory and no second pass.  The warp that finishes last divides
// (ticket): config 3's forward is ONE launch.
template <bool LOSS>
__global__ void __launch_bounds__(FK_THREADS)
fk_forward_kernel(FkIn in, int B, int swap, int bulk_ok, float* __restrict__ xyz, float* __restrict__ uv, FkLoss lo) {
    __shared__ FkTiles tiles[FK_WARPS];
    double acc[2] = {0.0, 0.0};
    unsigned long long nvis = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FkTiles& T = tiles[warp];
    const uint32_t bar = smem_u32(&T.bar);
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    const long long ntiles = ((long long)B + FK_T - 1) / FK_T, stride = (long long)gridDim.x * FK_WARPS;
    uint32_t parity = 0;
    const bool has_xyz = LOSS && lo.gt_xyz, has_uv = LOSS && lo.gt_uv;
    auto is_bulk = [&](long long t) { return bulk_ok && (t + 1) * FK_T <= B; };
    // inputs of tile t: asynchronously (bulk) or on the spot (loops).  LOSS: the ground truths land in the result tiles once
    // the previous tile's bulk stores have read them.
    auto load = [&](long long t) {
        const long long row0 = t * FK_T;
        if (is_bulk(t)) {
            if (lane == 0) {
                mbar_expect_tx(bar, FK_IN_BYTES + (has_xyz ? FK_XYZ_BYTES : 0u) + (has_uv ? FK_UV_BYTES : 0u));
                fk_request_inputs(T, in, row0, bar, true, true);
                if (has_xyz || has_uv) bulk_wait_read<0>();
                if (has_xyz) bulk_g2s(smem_u32(T.xyz), lo.gt_xyz + row0 * 63, FK_XYZ_BYTES, bar);
                if (has_uv) bulk_g2s(smem_u32(T.uv), lo.gt_uv + row0 * 42, FK_UV_BYTES, bar);
            }
        } else {
            const int rows = (int)(B - row0 < FK_T ? B - row0 : FK_T);
            fk_copy_inputs(T, in, row0, rows, lane);
            if (has_xyz || has_uv) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                if (has_xyz) fk_copy_in(T.xyz, lo.gt_xyz + row0 * 63, rows * 63, lane);
                if (has_uv) fk_copy_in(T.uv, lo.gt_uv + row0 * 42, rows * 42, lane);
            }
        }
    };
    long long t = (long long)blockIdx.x * FK_WARPS + warp;
    if (t < ntiles) load(t);
    for (; t < ntiles; t += stride) {
        const long long row0 = t * FK_T;
        const int rows = (int)(B - row0 < FK_T ? B - row0 : FK_T);
        unsigned vismask = 0u;
        if (LOSS && lane < rows) vismask = fk_vismask(lo.vis + (row0 + lane) * 21);
        if (is_bulk(t)) { fk_bar_wait(bar, parity); parity ^= 1u; }
        if (lane == 0) bulk_wait_read<0>();                    // the previous tile's results have left the result tiles
        __syncwarp();
        if (lane < rows) {
            fk_forward_sample<LOSS>(T.ra + lane * 3, T.oa + lane * FK_OA, T.bl + lane * FK_NODES, T.K + lane * 9, T.s[lane],
                                    T.root + lane * 3, swap, T.xyz + lane * 63, T.uv + lane * 42, vismask, has_xyz, has_uv, acc);
            if (LOSS) nvis += __popc(vismask);
        }
        fence_proxy_async();                                   // the bulk stores below read what this thread wrote
        __syncwarp();
        if (is_bulk(t)) {
            if (lane == 0) {
                bulk_s2g(xyz + row0 * 63, smem_u32(T.xyz), FK_XYZ_BYTES);
                bulk_s2g(uv + row0 * 42, smem_u32(T.uv), FK_UV_BYTES);
                bulk_commit();
            }
        } else {
            fk_copy_out(xyz + row0 * 63, T.xyz, rows * 63, lane);
            fk_copy_out(uv + row0 * 42, T.uv, rows * 42, lane);
            __syncwarp();
        }
        if (t + stride < ntiles) load(t + stride);             // the inputs are consumed: the next tile's can land
    }
    if (lane == 0) bulk_wait_all<0>();
    if (LOSS) {
        double n = (double)nvis;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
            acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane == 0) {
            if (has_xyz && acc[0] != 0.0) atomicAdd(&lo.accum[0], acc[0]);
            if (has_uv && acc[1] != 0.0) atomicAdd(&lo.accum[2], acc[1]);
            if (n != 0.0) { atomicAdd(&lo.accum[1], n); atomicAdd(&lo.accum[3], n); }
            __threadfence();
            const unsigned long long total = (unsigned long long)gridDim.x * FK_WARPS;
            if (atomicAdd(reinterpret_cast<unsigned long long*>(&lo.accum[4]), 1ULL) == total - 1) {
                __threadfence();
                const volatile double* a = lo.accum;
                lo.losses[0] = has_xyz && a[1] > 0.0 ? (float)(a[0] / a[1]) : 0.f;
                lo.losses[1] = has_uv && a[3] > 0.0 ? (float)(a[2] / a[3]) : 0.f;
            }
        }
    }
}

// LOSS: g_xyz / g_uv are the ground truths (their tiles arrive the same way) and the upstream gradients of the two L2 terms are
// formed inside fk_backward_sample<true> (visibility flags as a per-thread bit mask, loaded before the wait for the tile) —
// the gradients never exist in memory; config 3's backward is ONE launch.
template <bool LOSS>
__global__ void __launch_bounds__(FK_THREADS)
fk_backward_kernel(FkIn in, const float* __restrict__ g_xyz, const float* __restrict__ g_uv, int B, int swap, int bulk_ok,
                   float* __restrict__ g_ra, float* __restrict__ g_oa, float* __restrict__ g_bl,
                   const float* __restrict__ vis, const double* __restrict__ accum, const float* __restrict__ g_losses) {
    __shared__ FkTiles tiles[FK_WARPS];
    float kx = 0.f, ku = 0.f;
    if (LOSS) {                                                 // as head_l2_backward_kernel / masked_l2_backward_kernel
        kx = g_xyz && accum[1] > 0.0 ? (float)(2.0 * (double)g_losses[0] / accum[1]) : 0.f;
        ku = g_uv && accum[3] > 0.0 ? (float)(2.0 * (double)g_losses[1] / accum[3]) : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FkTiles& T = tiles[warp];
    const uint32_t bar = smem_u32(&T.bar);
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    const long long ntiles = ((long long)B + FK_T - 1) / FK_T, stride = (long long)gridDim.x * FK_WARPS;
    uint32_t parity = 0;
    auto is_bulk = [&](long long t) { return bulk_ok && (t + 1) * FK_T <= B; };
    // inputs + upstream gradients of tile t.  The result tiles (ra / oa / bl) are requested last, once the previous tile's
    // bulk stores have read them.
    auto load = [&](long long t) {
        const long long row0 = t * FK_T;
        if (is_bulk(t)) {
            if (lane == 0) {
                mbar_expect_tx(bar, FK_IN_BYTES + (g_xyz ? FK_XYZ_BYTES : 0u) + (g_uv ? FK_UV_BYTES : 0u));
                fk_request_inputs(T, in, row0, bar, false, true);
                if (g_xyz) bulk_g2s(smem_u32(T.xyz), g_xyz + row0 * 63, FK_XYZ_BYTES, bar);
                if (g_uv) bulk_g2s(smem_u32(T.uv), g_uv + row0 * 42, FK_UV_BYTES, bar);
                bulk_wait_read<0>();
                fk_request_inputs(T, in, row0, bar, true, false);
            }
        } else {
            const int rows = (int)(B - row0 < FK_T ? B - row0 : FK_T);
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            fk_copy_inputs(T, in, row0, rows, lane);
            if (g_xyz) fk_copy_in(T.xyz, g_xyz + row0 * 63, rows * 63, lane);
            if (g_uv) fk_copy_in(T.uv, g_uv + row0 * 42, rows * 42, lane);
        }
    };
    long long t = (long long)blockIdx.x * FK_WARPS + warp;
    if (t < ntiles) load(t);
    for (; t < ntiles; t += stride) {
        const long long row0 = t * FK_T;
        const int rows = (int)(B - row0 < FK_T ? B - row0 : FK_T);
        unsigned vismask = 0u;
        if (LOSS && lane < rows) vismask = fk_vismask(vis + (row0 + lane) * 21);
        if (is_bulk(t)) { fk_bar_wait(bar, parity); parity ^= 1u; }
        __syncwarp();
        if (lane < rows) {
            float r_gra[3], r_goa[FK_OA], r_gbl[FK_NODES];
            fk_backward_sample<LOSS>(T.ra + lane * 3, T.oa + lane * FK_OA, T.bl + lane * FK_NODES, T.K + lane * 9, T.s[lane],
                                     T.root + lane * 3, swap, g_xyz ? T.xyz + lane * 63 : nullptr, g_uv ? T.uv + lane * 42 : nullptr,
                                     r_gra, r_goa, r_gbl, vismask, kx, ku);
            // a sample's input rows become its result rows (same widths; only this thread touches them)
#pragma unroll
            for (int i = 0; i < 3; ++i) T.ra[lane * 3 + i] = r_gra[i];
#pragma unroll
            for (int i = 0; i < FK_OA; ++i) T.oa[lane * FK_OA + i] = r_goa[i];
#pragma unroll
            for (int i = 0; i < FK_NODES; ++i) T.bl[lane * FK_NODES + i] = r_gbl[i];
        }
        fence_proxy_async();
        __syncwarp();
        if (is_bulk(t)) {
            if (lane == 0) {
                bulk_s2g(g_ra + row0 * 3, smem_u32(T.ra), FK_T * 3 * 4);
                bulk_s2g(g_oa + row0 * FK_OA, smem_u32(T.oa), FK_T * FK_OA * 4);
                bulk_s2g(g_bl + row0 * FK_NODES, smem_u32(T.bl), FK_T * FK_NODES * 4);
                bulk_commit();
            }
        } else {
            fk_copy_out(g_ra + row0 * 3, T.ra, rows * 3, lane);
            fk_copy_out(g_oa + row0 * FK_OA, T.oa, rows * FK_OA, lane);
            fk_copy_out(g_bl + row0 * FK_NODES, T.bl, rows * FK_NODES, lane);
            __syncwarp();
        }
        if (t + stride < ntiles) load(t + stride);
    }
    if (lane == 0) bulk_wait_all<0>();
}

__global__ void project_forward_kernel(const float* __restrict__ xyz, const float* __restrict__ K, long long total, int N,
                                       float* __restrict__ uv) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float* k = K + (i / N) * 9;
        float u, v;
        project_point(k, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], u, v);
        uv[i * 2] = u; uv[i * 2 + 1] = v;
    }
}
__global__ void project_backward_kernel(const float* __restrict__ xyz, const float* __restrict__ K,
                                        const float* __restrict__ g_uv, long long total, int N, float* __restrict__ g_xyz) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float* k = K + (i / N) * 9;
        V3 g = project_point_bwd(k, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], g_uv[i * 2], g_uv[i * 2 + 1]);
        g_xyz[i * 3] = g.x; g_xyz[i * 3 + 1] = g.y; g_xyz[i * 3 + 2] = g.z;
    }
}

// persistent warps: at most the 5 co-resident blocks per SM, fewer when the batch has fewer tiles
inline int fk_grid(int B) {
    int nblk = (B + FK_THREADS - 1) / FK_THREADS;
    int cap = NUM_SMS * 5;
    return nblk < cap ? nblk : cap;
}
inline bool aligned16(std::initializer_list<const void*> ps) {
    uintptr_t a = 0;
    for (const void* p : ps) a |= (uintptr_t)p;
    return (a & 15) == 0;
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_fk_forward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                             const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                             int B, int swap_order, float* xyz, float* uv, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root || !xyz || !uv)
        return MB_E_NULL;
    StageTimer t(ST_FK_FWD, (cudaStream_t)stream);
    const FkIn in = {root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root};
    const int bulk_ok = aligned16({root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root, xyz, uv});
    fk_forward_kernel<false><<<fk_grid(B), FK_THREADS, 0, (cudaStream_t)stream>>>(in, B, swap_order != 0, bulk_ok, xyz, uv, FkLoss{});
    return cuda_rc();
}

extern "C" int mb_fk_backward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                              const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                              const float* g_xyz, const float* g_uv, int B, int swap_order,
                              float* g_root_angles, float* g_other_angles, float* g_bone_lengths, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root ||
        !g_root_angles || !g_other_angles || !g_bone_lengths)
        return MB_E_NULL;
    StageTimer t(ST_FK_BWD, (cudaStream_t)stream);
    const FkIn in = {root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root};
    const int bulk_ok = aligned16({root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root, g_xyz, g_uv,
                                   g_root_angles, g_other_angles, g_bone_lengths});
    fk_backward_kernel<false><<<fk_grid(B), FK_THREADS, 0, (cudaStream_t)stream>>>(in, g_xyz, g_uv, B, swap_order != 0, bulk_ok,
                                                                                    g_root_angles, g_other_angles, g_bone_lengths,
                                                                                    nullptr, nullptr, nullptr);
    return cuda_rc();
}

// ---- FK + the two L2Loss terms of the FK heads (TwoDimHandPoseWithFK / ThreeDimHandPose -> LossCalculation), one launch per
// direction.  workspace: double[8] (sums, counts, ticket) — written by the forward, read by the backward.
extern "C" size_t mb_fk_loss_workspace_bytes(int B) { (void)B; return 8 * sizeof(double); }

extern "C" int mb_fk_loss_forward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                                  const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                                  const float* gt_xyz, const float* gt_uv, const float* keypoint_vis, int B, int swap_order,
                                  int flags, float* xyz, float* uv, float* losses, void* workspace, size_t workspace_bytes,
                                  mb_stream_t stream) {
    if (B < 0 || (flags & ~(MB_HEAD_XYZ | MB_HEAD_UV))) return MB_E_RANGE;
    if (!losses || !workspace) return MB_E_NULL;
    if (workspace_bytes < mb_fk_loss_workspace_bytes(B)) return MB_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaError_t e = cudaMemsetAsync(workspace, 0, 8 * sizeof(double), s)) return (int)e;
    if (B == 0) return (int)cudaMemsetAsync(losses, 0, 2 * sizeof(float), s);
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root || !xyz || !uv)
        return MB_E_NULL;
    if (((flags & MB_HEAD_XYZ) && !gt_xyz) || ((flags & MB_HEAD_UV) && !gt_uv) || ((flags & (MB_HEAD_XYZ | MB_HEAD_UV)) && !keypoint_vis))
        return MB_E_NULL;
    StageTimer t(ST_FK_FWD, s);
    const FkIn in = {root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root};
    const int bulk_ok = aligned16({root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root, xyz, uv,
                                   (flags & MB_HEAD_XYZ) ? gt_xyz : nullptr, (flags & MB_HEAD_UV) ? gt_uv : nullptr});
    const FkLoss lo = {(flags & MB_HEAD_XYZ) ? gt_xyz : nullptr, (flags & MB_HEAD_UV) ? gt_uv : nullptr, keypoint_vis,
                       reinterpret_cast<double*>(workspace), losses};
    if (flags & (MB_HEAD_XYZ | MB_HEAD_UV))
        fk_forward_kernel<true><<<fk_grid(B), FK_THREADS, 0, s>>>(in, B, swap_order != 0, bulk_ok, xyz, uv, lo);
    else {
        if (cudaError_t e = cudaMemsetAsync(losses, 0, 2 * sizeof(float), s)) return (int)e;
        fk_forward_kernel<false><<<fk_grid(B), FK_THREADS, 0, s>>>(in, B, swap_order != 0, bulk_ok, xyz, uv, FkLoss{});
    }
    return cuda_rc();
}

extern "C" int mb_fk_loss_backward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                                   const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                                   const float* gt_xyz, const float* gt_uv, const float* keypoint_vis, int B, int swap_order,
                                   int flags, const float* g_losses, float* g_root_angles, float* g_other_angles,
                                   float* g_bone_lengths, const void* workspace, size_t workspace_bytes, mb_stream_t stream) {
    if (B < 0 || (flags & ~(MB_HEAD_XYZ | MB_HEAD_UV))) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root || !g_losses ||
        !g_root_angles || !g_other_angles || !g_bone_lengths || !workspace)
        return MB_E_NULL;
    if (((flags & MB_HEAD_XYZ) && !gt_xyz) || ((flags & MB_HEAD_UV) && !gt_uv) || ((flags & (MB_HEAD_XYZ | MB_HEAD_UV)) && !keypoint_vis))
        return MB_E_NULL;
    if (workspace_bytes < mb_fk_loss_workspace_bytes(B)) return MB_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    StageTimer t(ST_FK_BWD, s);
    const FkIn in = {root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root};
    const float* gx = (flags & MB_HEAD_XYZ) ? gt_xyz : nullptr;
    const float* gu = (flags & MB_HEAD_UV) ? gt_uv : nullptr;
    const int bulk_ok = aligned16({root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root, gx, gu,
                                   g_root_angles, g_other_angles, g_bone_lengths});
    fk_backward_kernel<true><<<fk_grid(B), FK_THREADS, 0, s>>>(in, gx, gu, B, swap_order != 0, bulk_ok, g_root_angles, g_other_angles,
                                                               g_bone_lengths, keypoint_vis, reinterpret_cast<const double*>(workspace),
                                                               g_losses);
    return cuda_rc();
}

extern "C" int mb_project_uv_forward(const float* xyz, const float* K, int B, int N, float* uv, mb_stream_t stream) {
    if (B < 0 || N < 0) return MB_E_RANGE;
    if (B == 0 || N == 0) return 0;
    if (!xyz || !K || !uv) return MB_E_NULL;
    const long long total = (long long)B * N;
    const long long blocks = (total + 255) / 256;
    project_forward_kernel<<<(unsigned)(blocks < NUM_SMS * 16 ? blocks : NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(
        xyz, K, total, N, uv);
    return cuda_rc();
}

extern "C" int mb_project_uv_backward(const float* xyz, const float* K, const float* g_uv, int B, int N,
                                      float* g_xyz, mb_stream_t stream) {
    if (B < 0 || N < 0) return MB_E_RANGE;
    if (B == 0 || N == 0) return 0;
    if (!xyz || !K || !g_uv || !g_xyz) return MB_E_NULL;
    const long long total = (long long)B * N;
    const long long blocks = (total + 255) / 256;
    project_backward_kernel<<<(unsigned)(blocks < NUM_SMS * 16 ? blocks : NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(
        xyz, K, g_uv, total, N, g_xyz);
    return cuda_rc();
}
