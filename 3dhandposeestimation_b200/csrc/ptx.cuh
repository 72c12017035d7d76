// ptx.cuh — mbarrier / bulk-copy (TMA engine) PTX wrappers shared by the tcgen05 GEMMs and the skinning kernels.
#pragma once
#include <stdint.h>

namespace mb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking test of a phase (mbarrier.test_wait never suspends the thread): for a polling scheduler
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a broken pipeline must not hang the GPU — after ~2 s worth of polls the kernel
// records the failure and every role falls through to the exit.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    for (int spin = 1;; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 255) == 0) {
            if (*abort_flag) return false;
            if (clock64() - t0 > 4000000000LL) break;          // ~2 s
        }
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// bulk copy with an L2 eviction-priority hint (createpolicy)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- thread-block clusters: rank / size of this CTA's cluster, cluster-wide barrier, multicast bulk copy ----------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one L2 read, delivered to the same shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier (same
// offset) receives the byte count
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask,
                                                   uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint "
                 "[%0], [%1], %2, [%3], %4, %5;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask), "l"(policy) : "memory");
}

// one lane of the (converged) warp, chosen by the hardware: code under it is single-threaded AND the compiler knows it, so
// bulk-copy / MMA operands stay in uniform registers without a per-instruction vote ("waterfall") loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- asynchronous global -> shared copies by the threads themselves (cp.async, LDGSTS) ----------------------------
// A staging loop `tile[i] = src[i]` is one dependent global-load round trip per element and thread; with the few
// warps per SM of the one-thread-per-hand kernels that chain was up to 40 % of their time.  cp.async requests are
// issued back to back and waited for once.
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// one warp copies n consecutive floats (16-byte requests when both sides are 16-byte aligned); the caller waits
// (cp_async_wait_all + __syncwarp) before anyone reads dst
__device__ __forceinline__ void warp_copy_async(float* dst_smem, const float* __restrict__ src, int n, int lane) {
    if ((((uintptr_t)src | (uintptr_t)smem_u32(dst_smem)) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = lane; i < n4; i += 32) cp_async16(dst_smem + 4 * i, src + 4 * i);
        for (int i = 4 * n4 + lane; i < n; i += 32) cp_async4(dst_smem + i, src + i);
    } else {
        for (int i = lane; i < n; i += 32) cp_async4(dst_smem + i, src + i);
    }
}

// rows of up to 32 samples <-> a per-warp shared tile as flat coalesced accesses (n samples of w floats, dense pitch w):
// used by the one-thread-per-sample epilogue kernels (joint_epilogue.cu, hand_trafo.cu, viewpoint.cu)
__device__ __forceinline__ void tile_load(float* tile, const float* __restrict__ src, long long base, int n, int w, int lane) {
    warp_copy_async(tile, src + base * w, n * w, lane);       // asynchronous requests, one wait
    cp_async_wait_all();
    __syncwarp();
}
__device__ __forceinline__ void tile_store(const float* tile, float* __restrict__ dst, long long base, int n, int w, int lane) {
    __syncwarp();
    float* d = dst + base * w;
    for (int i = lane; i < n * w; i += 32) d[i] = tile[i];
    __syncwarp();
}

}  // namespace mb
