// hand_math.cuh — small fixed-size rotation math shared by the MANO pose kernels and
// the forward-kinematics kernels.  Everything is __host__ __device__ so the same code
// is exercised on the CPU by tests/ (host harness) before it runs on the GPU.
#pragma once
#include <math.h>
#include "common.cuh"

namespace mb {

struct M3 { float m[9]; };   // row-major 3x3
struct V3 { float x, y, z; };

HD M3 m3_identity() { M3 r; r.m[0]=1.f; r.m[1]=0.f; r.m[2]=0.f; r.m[3]=0.f; r.m[4]=1.f; r.m[5]=0.f; r.m[6]=0.f; r.m[7]=0.f; r.m[8]=1.f; return r; }
HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.f; return r; }

HD M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i*3+j] = fmaf(a.m[i*3+0], b.m[0*3+j], fmaf(a.m[i*3+1], b.m[1*3+j], a.m[i*3+2] * b.m[2*3+j]));
    return r;
}
// a^T b
HD M3 m3_tmul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i*3+j] = fmaf(a.m[0*3+i], b.m[0*3+j], fmaf(a.m[1*3+i], b.m[1*3+j], a.m[2*3+i] * b.m[2*3+j]));
    return r;
}
// a b^T
HD M3 m3_mult(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i*3+j] = fmaf(a.m[i*3+0], b.m[j*3+0], fmaf(a.m[i*3+1], b.m[j*3+1], a.m[i*3+2] * b.m[j*3+2]));
    return r;
}
HD V3 m3_vec(const M3& a, const V3& v) {
    V3 r;
    r.x = fmaf(a.m[0], v.x, fmaf(a.m[1], v.y, a.m[2] * v.z));
    r.y = fmaf(a.m[3], v.x, fmaf(a.m[4], v.y, a.m[5] * v.z));
    r.z = fmaf(a.m[6], v.x, fmaf(a.m[7], v.y, a.m[8] * v.z));
    return r;
}
// a^T v
HD V3 m3_tvec(const M3& a, const V3& v) {
    V3 r;
    r.x = fmaf(a.m[0], v.x, fmaf(a.m[3], v.y, a.m[6] * v.z));
    r.y = fmaf(a.m[1], v.x, fmaf(a.m[4], v.y, a.m[7] * v.z));
    r.z = fmaf(a.m[2], v.x, fmaf(a.m[5], v.y, a.m[8] * v.z));
    return r;
}
HD void m3_add_outer(M3& acc, const V3& a, const V3& b) {   // acc += a b^T
    acc.m[0] = fmaf(a.x, b.x, acc.m[0]); acc.m[1] = fmaf(a.x, b.y, acc.m[1]); acc.m[2] = fmaf(a.x, b.z, acc.m[2]);
    acc.m[3] = fmaf(a.y, b.x, acc.m[3]); acc.m[4] = fmaf(a.y, b.y, acc.m[4]); acc.m[5] = fmaf(a.y, b.z, acc.m[5]);
    acc.m[6] = fmaf(a.z, b.x, acc.m[6]); acc.m[7] = fmaf(a.z, b.y, acc.m[7]); acc.m[8] = fmaf(a.z, b.z, acc.m[8]);
}
HD void m3_acc(M3& acc, const M3& a) { for (int i = 0; i < 9; ++i) acc.m[i] += a.m[i]; }
HD float m3_dot(const M3& a, const M3& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) s = fmaf(a.m[i], b.m[i], s);
    return s;
}
HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
HD V3 v3_add(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
HD V3 v3_sub(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
HD float v3_dot(const V3& a, const V3& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

HD void sincos_acc(float a, float* s, float* c) {
#ifdef __CUDA_ARCH__
    sincosf(a, s, c);       // accurate path (no --use_fast_math): <= 2 ulp
#else
    *s = sinf(a); *c = cosf(a);
#endif
}

// Axis-angle -> rotation.  Reference: ManoLayer.rodrigues, MANOLayer.py:82-112
// R = I + sin(t) S(n) + (1 - cos t) S(n)^2, n = r / t; Taylor form below 1e-30 (:102-110).
// Evaluated as R = I + a S(r) + b (r r^T - t^2 I), a = sin t / t, b = (1 - cos t) / t^2 — the same matrix — with the
// angle in DOUBLE: rounding t = |r| to fp32 before sin / cos costs 1 ulp of the ANGLE (3e-7 rad at t = 5), which was the
// largest error of the whole layer against the fp64 arbiter (profiles/r2: blended transforms off by 6.7e-7, verts
// by 1.3e-7 m).  t^2 is exact in double, 1 / t comes from the fp32 rsqrt refined by two Newton steps, sin / cos from the
// fp32 sincosf of the rounded angle plus the first-order term of the rounding residue; the entries are rounded to fp32
// once.  ~30 double FMAs per call (B200 runs them at half the fp32 rate), no double sqrt / divide.  rodrigues_d keeps
// the entries in double for the large-batch forward chain (mano_pose_lh.cu), whose four chained products would
// otherwise add another ~5e-7 to the blended transforms.
struct M3d { double m[9]; };  // row-major 3x3 in double: the forward chain of the large-batch pose kernel
struct V3d { double x, y, z; };
HD M3d m3d_mul(const M3d& a, const M3d& b) {
    M3d r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i*3+j] = fma(a.m[i*3+0], b.m[0*3+j], fma(a.m[i*3+1], b.m[1*3+j], a.m[i*3+2] * b.m[2*3+j]));
    return r;
}
HD V3d m3d_vec(const M3d& a, const V3d& v) {
    V3d r;
    r.x = fma(a.m[0], v.x, fma(a.m[1], v.y, a.m[2] * v.z));
    r.y = fma(a.m[3], v.x, fma(a.m[4], v.y, a.m[5] * v.z));
    r.z = fma(a.m[6], v.x, fma(a.m[7], v.y, a.m[8] * v.z));
    return r;
}
HD V3d v3d(const V3& v) { V3d r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
HD V3d v3d_add(const V3d& a, const V3d& b) { V3d r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; return r; }
HD V3d v3d_sub(const V3d& a, const V3d& b) { V3d r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; return r; }
HD M3 m3_from(const M3d& a) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = (float)a.m[i]; return r; }
HD V3 v3_from(const V3d& a) { return v3((float)a.x, (float)a.y, (float)a.z); }

HD M3d rodrigues_d(const V3& r) {
    const double x = r.x, y = r.y, z = r.z;
    const double t2 = x * x + (y * y + z * z);
    M3d R;
    if (t2 < 1e-60) {
        const double a = 1.0 - t2 / 6.0, b = 0.5 - t2 / 24.0;
        // S(r)^2 = r r^T - t2 I
        R.m[0] = 1.0 + b * (x * x - t2); R.m[1] = -a * z + b * x * y;     R.m[2] = a * y + b * x * z;
        R.m[3] = a * z + b * x * y;      R.m[4] = 1.0 + b * (y * y - t2); R.m[5] = -a * x + b * y * z;
        R.m[6] = -a * y + b * x * z;     R.m[7] = a * x + b * y * z;      R.m[8] = 1.0 + b * (z * z - t2);
        return R;
    }
    // 1 / t: fp32 estimate (t2 is rescaled into fp32's range: axis-angles below 1e-19 would underflow t2) + 2 Newton steps
    double inv;
    {
        const float f2 = (float)t2;
#ifdef __CUDA_ARCH__
        float y0 = f2 > 1e-30f ? rsqrtf(f2) : 1e15f * rsqrtf((float)(t2 * 1e30));
#else
        float y0 = f2 > 1e-30f ? 1.f / sqrtf(f2) : 1e15f / sqrtf((float)(t2 * 1e30));
#endif
        inv = (double)y0;
        inv = inv * (1.5 - 0.5 * t2 * inv * inv);
        inv = inv * (1.5 - 0.5 * t2 * inv * inv);
    }
    const double t = t2 * inv;
    const float th = (float)t;
    const float tl = (float)(t - (double)th);
    float sf, cf;
    sincos_acc(th, &sf, &cf);
    const double sd = (double)sf + (double)tl * (double)cf;      // sin(th + tl), cos(th + tl) to first order in tl (|tl| < 3e-7)
    const double cd = (double)cf - (double)tl * (double)sf;
    const double a = sd * inv, b = (1.0 - cd) * (inv * inv);
    const double bxy = b * x * y, bxz = b * x * z, byz = b * y * z;
    R.m[0] = 1.0 + b * (x * x - t2); R.m[1] = bxy - a * z;            R.m[2] = bxz + a * y;
    R.m[3] = bxy + a * z;            R.m[4] = 1.0 + b * (y * y - t2); R.m[5] = byz - a * x;
    R.m[6] = bxz - a * y;            R.m[7] = byz + a * x;            R.m[8] = 1.0 + b * (z * z - t2);
    return R;
}
HD M3 rodrigues(const V3& r) { return m3_from(rodrigues_d(r)); }

// d<dR, rodrigues(r)>/dr.  SURVEY Appendix A.2 step 6.  R = I + a S + b S^2 with
// a = sin t / t, b = (1 - cos t)/t^2 (S = skew(r)); analytic limits for small t where
// the reference's autograd yields NaN (documented deviation).
HD V3 rodrigues_bwd(const V3& r, const M3& dR) {
    float t2 = fmaf(r.x, r.x, fmaf(r.y, r.y, r.z * r.z));
    float a, b, a2, b2;
    if (t2 < 1e-2f) {        // t < 0.1: three series terms are exact to fp32 rounding
        a  = 1.f - t2 * (1.f / 6.f) + t2 * t2 * (1.f / 120.f);
        b  = 0.5f - t2 * (1.f / 24.f) + t2 * t2 * (1.f / 720.f);
        a2 = -1.f / 3.f + t2 * (1.f / 30.f) - t2 * t2 * (1.f / 840.f);
        b2 = -1.f / 12.f + t2 * (1.f / 180.f) - t2 * t2 * (1.f / 6720.f);
    } else {
        float t = sqrtf(t2), s, c;
        sincos_acc(t, &s, &c);
        a = s / t;
        // 1 - cos t = 2 sin^2(t/2): avoids cancellation for moderate t
        float sh, ch;
        sincos_acc(0.5f * t, &sh, &ch);
        float omc = 2.f * sh * sh;
        b = omc / t2;
        a2 = (c - a) / t2;             // (t cos t - sin t)/t^3
        b2 = (a - 2.f * b) / t2;       // (t sin t - 2(1 - cos t))/t^4
    }
    // <dR, S>     = x(d7 - d5) + y(d2 - d6) + z(d3 - d1)
    // <dR, S^2>   = r^T dR r - t2 tr(dR)
    // <dR, E_i S + S E_i> with S^2 = r r^T - t2 I : d(S^2)/dr_i = e_i r^T + r e_i^T - 2 r_i I
    const float* d = dR.m;
    float gS = r.x * (d[7] - d[5]) + r.y * (d[2] - d[6]) + r.z * (d[3] - d[1]);
    float tr = d[0] + d[4] + d[8];
    V3 dr = v3(fmaf(d[0], r.x, fmaf(d[1], r.y, d[2] * r.z)),
               fmaf(d[3], r.x, fmaf(d[4], r.y, d[5] * r.z)),
               fmaf(d[6], r.x, fmaf(d[7], r.y, d[8] * r.z)));        // dR r
    V3 dtr = v3(fmaf(d[0], r.x, fmaf(d[3], r.y, d[6] * r.z)),
                fmaf(d[1], r.x, fmaf(d[4], r.y, d[7] * r.z)),
                fmaf(d[2], r.x, fmaf(d[5], r.y, d[8] * r.z)));       // dR^T r
    float gS2 = v3_dot(r, dr) - t2 * tr;
    float common = a2 * gS + b2 * gS2;
    V3 g;
    g.x = common * r.x + a * (d[7] - d[5]) + b * (dr.x + dtr.x - 2.f * r.x * tr);
    g.y = common * r.y + a * (d[2] - d[6]) + b * (dr.y + dtr.y - 2.f * r.y * tr);
    g.z = common * r.z + a * (d[3] - d[1]) + b * (dr.z + dtr.z - 2.f * r.z * tr);
    return g;
}

// R = Rx(x) Ry(y) Rz(z).  Reference: get_right_hand_batch_rotation_matrix,
// forwardKinematicsLayer.py:59-96.
HD M3 euler_xyz(float x, float y, float z) {
    float sx, cx, sy, cy, sz, cz;
    sincos_acc(x, &sx, &cx); sincos_acc(y, &sy, &cy); sincos_acc(z, &sz, &cz);
    M3 R;
    R.m[0] = cy * cz;                 R.m[1] = -cy * sz;                R.m[2] = sy;
    R.m[3] = cx * sz + sx * sy * cz;  R.m[4] = cx * cz - sx * sy * sz;  R.m[5] = -sx * cy;
    R.m[6] = sx * sz - cx * sy * cz;  R.m[7] = sx * cz + cx * sy * sz;  R.m[8] = cx * cy;
    return R;
}

// (d<dR,R>/dx, /dy, /dz) for R = Rx Ry Rz.  SURVEY Appendix A.3.
HD V3 euler_xyz_bwd(float x, float y, float z, const M3& dR) {
    float sx, cx, sy, cy, sz, cz;
    sincos_acc(x, &sx, &cx); sincos_acc(y, &sy, &cy); sincos_acc(z, &sz, &cz);
    const float* d = dR.m;
    V3 g;
    // dR/dx
    g.x = d[3] * (-sx * sz + cx * sy * cz) + d[4] * (-sx * cz - cx * sy * sz) + d[5] * (-cx * cy)
        + d[6] * (cx * sz + sx * sy * cz)  + d[7] * (cx * cz - sx * sy * sz)  + d[8] * (-sx * cy);
    // dR/dy
    g.y = d[0] * (-sy * cz) + d[1] * (sy * sz) + d[2] * cy
        + d[3] * (sx * cy * cz) + d[4] * (-sx * cy * sz) + d[5] * (sx * sy)
        + d[6] * (-cx * cy * cz) + d[7] * (cx * cy * sz) + d[8] * (-cx * sy);
    // dR/dz
    g.z = d[0] * (-cy * sz) + d[1] * (-cy * cz)
        + d[3] * (cx * cz - sx * sy * sz) + d[4] * (-cx * sz - sx * sy * cz)
        + d[6] * (sx * cz + cx * sy * sz) + d[7] * (-sx * sz + cx * sy * cz);
    return g;
}

}  // namespace mb
