// Micro-benchmark: how fast can one SM push 192-byte row pieces (rows 9 336 B apart, as verts[B][778][3]) to HBM
//   A: one cp.async.bulk.global.shared::cta per lane per segment (32 ops per warp instruction)
//   B: st.global.v2 through a transposed tile (24 lanes x 8 B per row piece), the current skin_forward path
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_rate bulk_store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NVC = 2334, WARPS = 8, P = 60, SEGS = 49;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(WARPS * 32, 1) kA(float* verts, int ngroups, int fill) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row = sm + (warp * 32 + lane) * P + ((lane & 1) ? 2 : 0);
    const int d = (8 - 2 * (lane & 3)) & 7;                  // (0, 6, 4, 2)
    for (int i = 0; i < 56; ++i) row[i] = (float)(lane + i);
    for (int g = blockIdx.x + warp * gridDim.x; g < ngroups; g += gridDim.x * WARPS) {
        float* vrow = verts + ((size_t)g * 32 + lane) * NVC;
        for (int seg = 1; seg < SEGS - 1; ++seg) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            for (int i = 0; i < fill; ++i) row[8 + ((i * 7) % 48)] = (float)(seg + i);     // stand-in for the tile writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(vrow + 48 * seg - d), "r"(smem_u32(row + 8 - d)), "r"(192) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void __launch_bounds__(WARPS * 32, 1) kB(float* verts, int ngroups, int fill) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TP = 34;
    float* tile = sm + warp * (56 * TP);
    for (int i = lane; i < 56 * TP; i += 32) tile[i] = (float)i;
    __syncwarp();
    for (int g = blockIdx.x + warp * gridDim.x; g < ngroups; g += gridDim.x * WARPS) {
        float* row0 = verts + (size_t)g * 32 * NVC;
        for (int seg = 1; seg < SEGS - 1; ++seg) {
            for (int i = 0; i < fill; ++i) tile[(8 + ((i * 7) % 48)) * TP + lane] = (float)(seg + i);
            __syncwarp();
            for (int rb = 0; rb < 8; ++rb)
#pragma unroll
                for (int it = 0; it < 3; ++it) {
                    const int i = lane + 32 * it, rr = i / 24, pp = i - rr * 24, h = rb * 4 + rr, d2 = (4 - rr) & 3;
                    const float* t = tile + (8 - 2 * d2 + 2 * pp) * TP + h;
                    const float2 v = make_float2(t[0], t[TP]);
                    *reinterpret_cast<float2*>(row0 + (size_t)h * NVC + 2 * (24 * seg - d2 + pp)) = v;
                }
            __syncwarp();
        }
    }
}
int main() {
    const int H = 1 << 20, ngroups = H / 32;
    float* verts;
    cudaMalloc(&verts, (size_t)H * NVC * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smA = WARPS * 32 * P * 4, smB = WARPS * 56 * 34 * 4;
    cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
    for (int fill : {0, 48}) {
        for (int which = 0; which < 2; ++which) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) kA<<<148, WARPS * 32, smA>>>(verts, ngroups, fill);
                else kB<<<148, WARPS * 32, smB>>>(verts, ngroups, fill);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            const double bytes = (double)H * 47 * 192;
            printf("%s fill=%d: %.3f ms  %.2f TB/s  (%s)\n", which == 0 ? "A bulk-store per lane" : "B st.v2 via tile    ", fill, best,
                   bytes / best / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
