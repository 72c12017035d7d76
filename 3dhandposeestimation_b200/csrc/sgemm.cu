// sgemm.cu — fp32 FFMA GEMM used by MB_MODE_FP32 for the two blend-shape contractions
//   forward : v_posed[B][2334] = feat[B][148]  * basis  [148][2334]   (MANOLayer.py:130-137)
//   backward: dfeat [B][148]   = dv_posed[B][2334] * basis^T[2334][148]
// Row-major everywhere; shared-memory tiled, register-blocked, double-buffered global
// loads.  This is the correctness anchor (fp32 end to end); the tensor-core path
// (blend_tc.cu) is the fast one.
#include "common.cuh"

namespace mb {
namespace {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
             float* __restrict__ C, int ldc, long long M, int N, int K) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;            // threads along N; thread tx owns columns tx + j*TXN
    __shared__ alignas(16) float As[2][BK][BM + 4];   // A tile stored transposed: As[k][m]
    __shared__ alignas(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);      // column group
    const int ty = tid / (BN / TN);      // row group
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    constexpr int A_ELEMS = BM * BK, B_ELEMS = BK * BN;
    constexpr int A_PER = (A_ELEMS + NT - 1) / NT, B_PER = (B_ELEMS + NT - 1) / NT;
    float ra[A_PER], rb[B_PER];

    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int e = tid + i * NT;
            int m = e / BK, k = e % BK;           // consecutive threads walk K (contiguous in A rows)
            float v = 0.f;
            if (e < A_ELEMS && m0 + m < M && k0 + k < K) v = A[(m0 + m) * lda + k0 + k];
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int e = tid + i * NT;
            int k = e / BN, n = e % BN;
            float v = 0.f;
            if (e < B_ELEMS && k0 + k < K && n0 + n < N) v = Bm[(size_t)(k0 + k) * ldb + n0 + n];
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int e = tid + i * NT;
            if (e < A_ELEMS) As[buf][e % BK][e / BK] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int e = tid + i * NT;
            if (e < B_ELEMS) Bs[buf][e / BN][e % BN] = rb[i];
        }
    };

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][tx + j * TXN];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        long long m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx + j * TXN;
            if (n < N) C[m * ldc + n] = acc[i][j];
        }
    }
}

}  // namespace

int launch_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                 long long M, int N, int K, cudaStream_t s) {
    if (M <= 0) return 0;
    if (N > 256) {
        constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
        dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN);
        sgemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(A, lda, Bm, ldb, C, ldc, M, N, K);
    } else {
        // narrow output (the 148-wide feature gradient): one column tile of 160
        constexpr int BM = 128, BN = 160, BK = 8, TM = 8, TN = 10;
        dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN);
        sgemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, s>>>(A, lda, Bm, ldb, C, ldc, M, N, K);
    }
    return cuda_rc();
}

}  // namespace mb
