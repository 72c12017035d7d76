// host_check.cu — TEST-ONLY harness: exposes the __host__ __device__ math of
// csrc/hand_math.cuh and csrc/fk_math.cuh to the CPU test-suite, so the exact code the GPU
// kernels execute is checked against the oracle without a GPU.  Not part of the product.
#include "../../3dhandposeestimation_b200/csrc/fk_math.cuh"

using namespace mb;

extern "C" {
__attribute__((visibility("default"))) void hc_rodrigues(const float* r, float* R) {
    M3 m = rodrigues(v3(r[0], r[1], r[2]));
    for (int i = 0; i < 9; ++i) R[i] = m.m[i];
}
__attribute__((visibility("default"))) void hc_rodrigues_bwd(const float* r, const float* dR, float* g) {
    M3 d; for (int i = 0; i < 9; ++i) d.m[i] = dR[i];
    V3 o = rodrigues_bwd(v3(r[0], r[1], r[2]), d);
    g[0] = o.x; g[1] = o.y; g[2] = o.z;
}
__attribute__((visibility("default"))) void hc_euler(const float* a, float* R) {
    M3 m = euler_xyz(a[0], a[1], a[2]);
    for (int i = 0; i < 9; ++i) R[i] = m.m[i];
}
__attribute__((visibility("default"))) void hc_euler_bwd(const float* a, const float* dR, float* g) {
    M3 d; for (int i = 0; i < 9; ++i) d.m[i] = dR[i];
    V3 o = euler_xyz_bwd(a[0], a[1], a[2], d);
    g[0] = o.x; g[1] = o.y; g[2] = o.z;
}
__attribute__((visibility("default"))) void hc_fk_forward(int B, const float* ra, const float* oa, const float* bl, const float* K,
                                                          const float* s, const float* root, int swap, float* xyz, float* uv) {
    for (int b = 0; b < B; ++b)
        fk_forward_sample(ra + b * 3, oa + b * 23, bl + b * 20, K + b * 9, s[b], root + b * 3, swap, xyz + b * 63, uv + b * 42);
}
__attribute__((visibility("default"))) void hc_fk_backward(int B, const float* ra, const float* oa, const float* bl, const float* K,
                                                           const float* s, const float* root, int swap, const float* g_xyz,
                                                           const float* g_uv, float* g_ra, float* g_oa, float* g_bl) {
    for (int b = 0; b < B; ++b)
        fk_backward_sample(ra + b * 3, oa + b * 23, bl + b * 20, K + b * 9, s[b], root + b * 3, swap,
                           g_xyz ? g_xyz + b * 63 : nullptr, g_uv ? g_uv + b * 42 : nullptr,
                           g_ra + b * 3, g_oa + b * 23, g_bl + b * 20);
}
// kind 0..3 (xyz / xy / x / y): the specialised chain step against euler_xyz + m3_mul on the same inputs
__attribute__((visibility("default"))) void hc_fk_chain(int kind, const float* A, const float* ang, float* special, float* general) {
    M3 a; for (int i = 0; i < 9; ++i) a.m[i] = A[i];
    const float x = kind == 3 ? 0.f : ang[0], y = kind == 2 ? 0.f : ang[1], z = kind == 0 ? ang[2] : 0.f;
    const M3 s = fk_chain_rot(kind, a, fk_local_rot(kind, fk_sincos(kind, x, y, z)));
    const M3 g = m3_mul(a, euler_xyz(x, y, z));
    for (int i = 0; i < 9; ++i) { special[i] = s.m[i]; general[i] = g.m[i]; }
}
__attribute__((visibility("default"))) void hc_fk_angle_grad(int kind, const float* ang, const float* dR, float* special, float* general) {
    M3 d; for (int i = 0; i < 9; ++i) d.m[i] = dR[i];
    const float x = kind == 3 ? 0.f : ang[0], y = kind == 2 ? 0.f : ang[1], z = kind == 0 ? ang[2] : 0.f;
    const V3 s = fk_angle_grad(kind, fk_sincos(kind, x, y, z), d);
    const V3 g = euler_xyz_bwd(x, y, z, d);
    special[0] = s.x; special[1] = s.y; special[2] = s.z;
    general[0] = g.x; general[1] = g.y; general[2] = g.z;
}
}
