// fk_math.cuh — per-sample math of the RHD 21-joint forward-kinematics layer.
// __host__ __device__ so tests can run the identical code on the CPU.
//
// Reference: ForwardKinematics.forward, forwardKinematicsLayer.py:147-330:
//   R_root = Rx Ry Rz (root_angles)                                   :214-215, :59-96
//   node i (A1..E4): parent = root if i%4==0 else node i-1            :225-230
//   local angles from other_angles[23] through the DOF map            :239-274
//   R_i = R_parent R_local ; p_i = p_parent + L_i R_i[:,2]            :286-308
//   xyz = p * index_root_bone_length + kp_coord_xyz_root              :321, :333-358
//   optional per-finger order swap                                    :324-327
//   uv = (K xyz)_xy / (K xyz)_z, z==0 -> 1e-10                        utils/coordinate_trans.py:48-65
#pragma once
#include "hand_math.cuh"

namespace mb {

constexpr int FK_NODES = 20;
constexpr int FK_OA = 23;

// local Euler angles of segment `seg` (0..2) of finger f (0 = thumb)
HD void fk_local_angles(const float* oa, int f, int seg, float& x, float& y, float& z) {
    x = 0.f; y = 0.f; z = 0.f;
    if (f == 0) {
        if (seg == 0) { x = oa[0]; y = oa[1]; z = oa[2]; }
        else if (seg == 1) { x = oa[3]; y = oa[4]; z = oa[5]; }
        else { y = oa[6]; }
    } else {
        const int b = 7 + 4 * (f - 1);
        if (seg == 0) { x = oa[b]; y = oa[b + 1]; }
        else if (seg == 1) { x = oa[b + 2]; }
        else { x = oa[b + 3]; }
    }
}
// scatter (gx,gy,gz) of a segment's local angles back to other_angles
HD void fk_scatter_angles(float* g_oa, int f, int seg, const V3& g) {
    if (f == 0) {
        if (seg == 0) { g_oa[0] += g.x; g_oa[1] += g.y; g_oa[2] += g.z; }
        else if (seg == 1) { g_oa[3] += g.x; g_oa[4] += g.y; g_oa[5] += g.z; }
        else { g_oa[6] += g.y; }
    } else {
        const int b = 7 + 4 * (f - 1);
        if (seg == 0) { g_oa[b] += g.x; g_oa[b + 1] += g.y; }
        else if (seg == 1) { g_oa[b + 2] += g.x; }
        else { g_oa[b + 3] += g.x; }
    }
}
// output slot of node n (1..20) — identity, or reversed inside its finger when swapping
HD int fk_out_slot(int n, int swap) {
    if (!swap) return n;
    const int i = 1 + 4 * ((n - 1) / 4);
    return 2 * i + 3 - n;
}

HD void project_point(const float* K, float x, float y, float z, float& u, float& v) {
    const float px = fmaf(K[0], x, fmaf(K[1], y, K[2] * z));
    const float py = fmaf(K[3], x, fmaf(K[4], y, K[5] * z));
    float pz = fmaf(K[6], x, fmaf(K[7], y, K[8] * z));
    if (pz == 0.f) pz = 1e-10f;
    u = px / pz; v = py / pz;
}
// gradient of project_point w.r.t. (x,y,z); on the pz==0 branch the divisor is a constant
HD V3 project_point_bwd(const float* K, float x, float y, float z, float gu, float gv) {
    const float px = fmaf(K[0], x, fmaf(K[1], y, K[2] * z));
    const float py = fmaf(K[3], x, fmaf(K[4], y, K[5] * z));
    float pz = fmaf(K[6], x, fmaf(K[7], y, K[8] * z));
    const bool zero = (pz == 0.f);
    if (zero) pz = 1e-10f;
    const float inv = 1.f / pz;
    const float dpx = gu * inv, dpy = gv * inv;
    const float dpz = zero ? 0.f : -(gu * px * inv + gv * py * inv) * inv;
    return v3(fmaf(K[0], dpx, fmaf(K[3], dpy, K[6] * dpz)),
              fmaf(K[1], dpx, fmaf(K[4], dpy, K[7] * dpz)),
              fmaf(K[2], dpx, fmaf(K[5], dpy, K[8] * dpz)));
}

// ---- local rotations by their zero structure -----------------------------------------------------------------
// Of the 16 Euler triples only the root and the first two thumb segments use all three angles
// (forwardKinematicsLayer.py:239-274): the third thumb segment turns about y only, a finger's first segment about
// x and y, its second and third about x only.  sin 0 = 0 and cos 0 = 1 are exact and a product with an exact zero
// adds nothing in m3_mul's fmaf chains, so dropping those terms gives the SAME floats as euler_xyz + m3_mul with
// 26 sincos instead of 48 and about a third of the multiplies.  kind: 0 = xyz, 1 = xy, 2 = x, 3 = y.
HD int fk_angle_kind(int f, int seg) { return f == 0 ? (seg < 2 ? 0 : 3) : (seg == 0 ? 1 : 2); }

struct FkSC { float sx, cx, sy, cy, sz, cz; };                    // unused pairs stay (0, 1)

HD FkSC fk_sincos(int kind, float x, float y, float z) {
    FkSC q = {0.f, 1.f, 0.f, 1.f, 0.f, 1.f};
    if (kind != 3) sincos_acc(x, &q.sx, &q.cx);
    if (kind == 0 || kind == 1 || kind == 3) sincos_acc(y, &q.sy, &q.cy);
    if (kind == 0) sincos_acc(z, &q.sz, &q.cz);
    return q;
}
// the local rotation itself (the backward needs it for dRg Rl^T)
HD M3 fk_local_rot(int kind, const FkSC& q) {
    M3 R;
    if (kind == 0) {
        R.m[0] = q.cy * q.cz;                       R.m[1] = -q.cy * q.sz;                      R.m[2] = q.sy;
        R.m[3] = q.cx * q.sz + q.sx * q.sy * q.cz;  R.m[4] = q.cx * q.cz - q.sx * q.sy * q.sz;  R.m[5] = -q.sx * q.cy;
        R.m[6] = q.sx * q.sz - q.cx * q.sy * q.cz;  R.m[7] = q.sx * q.cz + q.cx * q.sy * q.sz;  R.m[8] = q.cx * q.cy;
    } else if (kind == 1) {
        R.m[0] = q.cy;           R.m[1] = 0.f;   R.m[2] = q.sy;
        R.m[3] = q.sx * q.sy;    R.m[4] = q.cx;  R.m[5] = -q.sx * q.cy;
        R.m[6] = -(q.cx * q.sy); R.m[7] = q.sx;  R.m[8] = q.cx * q.cy;
    } else if (kind == 2) {
        R.m[0] = 1.f; R.m[1] = 0.f;  R.m[2] = 0.f;
        R.m[3] = 0.f; R.m[4] = q.cx; R.m[5] = -q.sx;
        R.m[6] = 0.f; R.m[7] = q.sx; R.m[8] = q.cx;
    } else {
        R.m[0] = q.cy;  R.m[1] = 0.f; R.m[2] = q.sy;
        R.m[3] = 0.f;   R.m[4] = 1.f; R.m[5] = 0.f;
        R.m[6] = -q.sy; R.m[7] = 0.f; R.m[8] = q.cy;
    }
    return R;
}
// Rg = A Rl with the zero terms of m3_mul dropped (same association for the terms that stay)
HD M3 fk_chain_rot(int kind, const M3& A, const M3& Rl) {
    if (kind == 0) return m3_mul(A, Rl);
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a0 = A.m[i * 3], a1 = A.m[i * 3 + 1], a2 = A.m[i * 3 + 2];
        if (kind == 1) {
            r.m[i * 3]     = fmaf(a0, Rl.m[0], fmaf(a1, Rl.m[3], a2 * Rl.m[6]));
            r.m[i * 3 + 1] = fmaf(a1, Rl.m[4], a2 * Rl.m[7]);
            r.m[i * 3 + 2] = fmaf(a0, Rl.m[2], fmaf(a1, Rl.m[5], a2 * Rl.m[8]));
        } else if (kind == 2) {
            r.m[i * 3]     = a0;
            r.m[i * 3 + 1] = fmaf(a1, Rl.m[4], a2 * Rl.m[7]);
            r.m[i * 3 + 2] = fmaf(a1, Rl.m[5], a2 * Rl.m[8]);
        } else {
            r.m[i * 3]     = fmaf(a0, Rl.m[0], a2 * Rl.m[6]);
            r.m[i * 3 + 1] = a1;
            r.m[i * 3 + 2] = fmaf(a0, Rl.m[2], a2 * Rl.m[8]);
        }
    }
    return r;
}
// euler_xyz_bwd from the sines and cosines the forward sweep already has; only the angles of `kind` are meaningful
HD V3 fk_angle_grad(int kind, const FkSC& q, const M3& dR) {
    const float* d = dR.m;
    const float sx = q.sx, cx = q.cx, sy = q.sy, cy = q.cy, sz = q.sz, cz = q.cz;
    V3 g = v3(0.f, 0.f, 0.f);
    if (kind == 0) {
        g.x = d[3] * (-sx * sz + cx * sy * cz) + d[4] * (-sx * cz - cx * sy * sz) + d[5] * (-cx * cy)
            + d[6] * (cx * sz + sx * sy * cz)  + d[7] * (cx * cz - sx * sy * sz)  + d[8] * (-sx * cy);
        g.y = d[0] * (-sy * cz) + d[1] * (sy * sz) + d[2] * cy
            + d[3] * (sx * cy * cz) + d[4] * (-sx * cy * sz) + d[5] * (sx * sy)
            + d[6] * (-cx * cy * cz) + d[7] * (cx * cy * sz) + d[8] * (-cx * sy);
        g.z = d[0] * (-cy * sz) + d[1] * (-cy * cz)
            + d[3] * (cx * cz - sx * sy * sz) + d[4] * (-cx * sz - sx * sy * cz)
            + d[6] * (sx * cz + cx * sy * sz) + d[7] * (-sx * sz + cx * sy * cz);
    } else if (kind == 1) {
        g.x = d[3] * (cx * sy) + d[4] * (-sx) + d[5] * (-cx * cy) + d[6] * (sx * sy) + d[7] * cx + d[8] * (-sx * cy);
        g.y = d[0] * (-sy) + d[2] * cy + d[3] * (sx * cy) + d[5] * (sx * sy) + d[6] * (-cx * cy) + d[8] * (-cx * sy);
    } else if (kind == 2) {
        g.x = d[4] * (-sx) + d[5] * (-cx) + d[7] * cx + d[8] * (-sx);
    } else {
        g.y = d[0] * (-sy) + d[2] * cy + d[6] * (-cy) + d[8] * (-sy);
    }
    return g;
}

// keypoint o of a sample leaves: xyz / uv rows of the sample.  LOSS (the fused FK + L2Loss forward, mb_fk_loss_forward): the
// rows hold the GROUND TRUTH on entry — the squared distances of a visible joint (criterions/loss.py:10-25: fp32 sum over the
// coordinates, as masked_reduce_kernel forms it from the stored outputs) join the thread's fp64 sums acc = {S_xyz, S_uv}
// before the keypoint overwrites its slot.
template <bool LOSS>
HD void fk_emit(int o, float X, float Y, float Z, float u, float v, float* xyz, float* uv, unsigned vismask, bool has_xyz, bool has_uv,
                double* acc) {
    if (LOSS) {
        if ((vismask >> o) & 1u) {
            if (has_xyz) {
                const float a = X - xyz[o * 3], b = Y - xyz[o * 3 + 1], c = Z - xyz[o * 3 + 2];
                acc[0] += (double)(a * a + b * b + c * c);
            }
            if (has_uv) {
                const float a = u - uv[o * 2], b = v - uv[o * 2 + 1];
                acc[1] += (double)(a * a + b * b);
            }
        }
    }
    xyz[o * 3] = X; xyz[o * 3 + 1] = Y; xyz[o * 3 + 2] = Z;
    uv[o * 2] = u; uv[o * 2 + 1] = v;
}

// One sample forward.  xyz[63], uv[42] may be strided (element stride 1, caller gives row base).
// LOSS: see fk_emit; vismask bit o = joint o (output order) is visible.
template <bool LOSS = false>
HD void fk_forward_sample(const float* ra, const float* oa, const float* bl, const float* K, float s,
                          const float* root, int swap, float* xyz, float* uv, unsigned vismask = 0u, bool has_xyz = false,
                          bool has_uv = false, double* acc = nullptr) {
    const M3 Rroot = euler_xyz(ra[0], ra[1], ra[2]);
    {
        float u, v;                                                 // wrist: p = 0
        project_point(K, root[0], root[1], root[2], u, v);
        fk_emit<LOSS>(0, root[0], root[1], root[2], u, v, xyz, uv, vismask, has_xyz, has_uv, acc);
    }
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        M3 Rpar = Rroot;
        V3 P = v3(0.f, 0.f, 0.f);
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
            M3 Rg;
            if (seg < 3) {
                float x, y, z;
                fk_local_angles(oa, f, seg, x, y, z);
                const int kind = fk_angle_kind(f, seg);
                Rg = fk_chain_rot(kind, Rpar, fk_local_rot(kind, fk_sincos(kind, x, y, z)));
            } else {
                Rg = Rpar;                                          // tip: identity local rotation
            }
            const float L = bl[4 * f + seg];
            P = v3(fmaf(L, Rg.m[2], P.x), fmaf(L, Rg.m[5], P.y), fmaf(L, Rg.m[8], P.z));
            const int o = fk_out_slot(1 + 4 * f + seg, swap);
            const float X = fmaf(P.x, s, root[0]), Y = fmaf(P.y, s, root[1]), Z = fmaf(P.z, s, root[2]);
            float u, v;
            project_point(K, X, Y, Z, u, v);
            fk_emit<LOSS>(o, X, Y, Z, u, v, xyz, uv, vismask, has_xyz, has_uv, acc);
            Rpar = Rg;
        }
    }
}

// One sample backward.  g_xyz / g_uv may be NULL.
// g_ra[3], g_oa[23], g_bl[20] are overwritten.
// LOSS = true (the fused FK + L2Loss backward, mb_fk_loss_backward): g_xyz / g_uv point at the GROUND TRUTH rows instead and
// the upstream gradients of the two masked-mean L2 terms (criterions/loss.py:10-25) are formed on the spot from the recomputed
// keypoints — d/dxyz = kx [visible] (xyz - gt), d/duv = ku [visible] (uv - gt), kx = 2 g_loss_xyz / N_visible, bit o of vismask =
// joint o (output order) is visible — the same
// floats masked_l2_backward_kernel writes from the forward's stored outputs (the forward sweep here recomputes them bit for bit).
template <bool LOSS = false>
HD void fk_backward_sample(const float* ra, const float* oa, const float* bl, const float* K, float s,
                           const float* root, int swap, const float* g_xyz, const float* g_uv,
                           float* g_ra, float* g_oa, float* g_bl, unsigned vismask = 0u, float kx = 0.f, float ku = 0.f) {
    const M3 Rroot = euler_xyz(ra[0], ra[1], ra[2]);
    M3 dRroot = m3_zero();
    for (int i = 0; i < FK_OA; ++i) g_oa[i] = 0.f;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        M3 Rg[4], Rl[3];
        FkSC sc[3];
        V3 dP[4];
        {
            M3 Rpar = Rroot;
            V3 P = v3(0.f, 0.f, 0.f);
#pragma unroll
            for (int seg = 0; seg < 4; ++seg) {
                if (seg < 3) {
                    float x, y, z;
                    fk_local_angles(oa, f, seg, x, y, z);
                    const int kind = fk_angle_kind(f, seg);
                    sc[seg] = fk_sincos(kind, x, y, z);
                    Rl[seg] = fk_local_rot(kind, sc[seg]);
                    Rg[seg] = fk_chain_rot(kind, Rpar, Rl[seg]);
                } else {
                    Rg[seg] = Rpar;
                }
                const float L = bl[4 * f + seg];
                P = v3(fmaf(L, Rg[seg].m[2], P.x), fmaf(L, Rg[seg].m[5], P.y), fmaf(L, Rg[seg].m[8], P.z));
                const int o = fk_out_slot(1 + 4 * f + seg, swap);
                V3 dx = v3(0.f, 0.f, 0.f);
                if (LOSS) {
                    const float X = fmaf(P.x, s, root[0]), Y = fmaf(P.y, s, root[1]), Z = fmaf(P.z, s, root[2]);
                    const bool m = ((vismask >> o) & 1u) != 0u;
                    if (g_xyz) {
                        const float k = m ? kx : 0.f;
                        dx = v3(k * (X - g_xyz[o * 3]), k * (Y - g_xyz[o * 3 + 1]), k * (Z - g_xyz[o * 3 + 2]));
                    }
                    if (g_uv) {
                        const float k = m ? ku : 0.f;
                        float u, v;
                        project_point(K, X, Y, Z, u, v);
                        dx = v3_add(dx, project_point_bwd(K, X, Y, Z, k * (u - g_uv[o * 2]), k * (v - g_uv[o * 2 + 1])));
                    }
                } else {
                    if (g_xyz) dx = v3(g_xyz[o * 3], g_xyz[o * 3 + 1], g_xyz[o * 3 + 2]);
                    if (g_uv) {
                        const float X = fmaf(P.x, s, root[0]), Y = fmaf(P.y, s, root[1]), Z = fmaf(P.z, s, root[2]);
                        dx = v3_add(dx, project_point_bwd(K, X, Y, Z, g_uv[o * 2], g_uv[o * 2 + 1]));
                    }
                }
                dP[seg] = v3(dx.x * s, dx.y * s, dx.z * s);
                Rpar = Rg[seg];
            }
        }
        // reverse over the finger: node seg=3 (tip) .. 0
        M3 dRg = m3_zero();        // gradient of the current node's global rotation
        V3 dPacc = v3(0.f, 0.f, 0.f);
#pragma unroll
        for (int seg = 3; seg >= 0; --seg) {
            dPacc = v3_add(dPacc, dP[seg]);          // dP of this node including its descendants
            const float L = bl[4 * f + seg];
            g_bl[4 * f + seg] = fmaf(dPacc.x, Rg[seg].m[2], fmaf(dPacc.y, Rg[seg].m[5], dPacc.z * Rg[seg].m[8]));
            dRg.m[2] = fmaf(L, dPacc.x, dRg.m[2]);
            dRg.m[5] = fmaf(L, dPacc.y, dRg.m[5]);
            dRg.m[8] = fmaf(L, dPacc.z, dRg.m[8]);
            if (seg == 3) continue;                   // tip: R_tip = R_parent, dRg flows through unchanged
            const M3& Rpar = (seg == 0) ? Rroot : Rg[seg - 1];
            const M3 dRl = m3_tmul(Rpar, dRg);
            fk_scatter_angles(g_oa, f, seg, fk_angle_grad(fk_angle_kind(f, seg), sc[seg], dRl));
            dRg = m3_mult(dRg, Rl[seg]);              // becomes the parent's dRg contribution
        }
        m3_acc(dRroot, dRg);
    }
    const V3 g = euler_xyz_bwd(ra[0], ra[1], ra[2], dRroot);
    g_ra[0] = g.x; g_ra[1] = g.y; g_ra[2] = g.z;
}

}  // namespace mb
