/*
 * mano_b200.h — C ABI of libmano_b200.so: the sm_100a (B200) implementation of the
 * batched MANO layer, the RHD 21-joint forward-kinematics layer and the masked
 * joint reductions of hongrui16/3DHandPoseEstimation.
 *
 * This is the drop-in boundary.  The reference is pure Python/PyTorch with no FFI
 * of its own, so each entry point names the reference Python it replaces
 * (paths relative to the reference checkout); INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer to contiguous row-major fp32 unless
 *     stated otherwise; the library borrows it for the duration of the launch;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises;
 *   - return value 0 = success, >0 = cudaError_t of the failed launch,
 *     <0 = argument error (MB_E_*); the library never throws and never falls back
 *     to a CPU path;
 *   - B = number of hands / samples.  B == 0 is a successful no-op.
 */
#ifndef MANO_B200_H
#define MANO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 4

#if defined(__GNUC__)
#define MB_API __attribute__((visibility("default")))
#else
#define MB_API
#endif

/* argument errors */
#define MB_E_NULL      (-1)   /* required pointer is NULL                    */
#define MB_E_RANGE     (-2)   /* B < 0, nc outside [1,45], bad mode/kind     */
#define MB_E_WORKSPACE (-3)   /* workspace smaller than mb_mano_workspace_bytes */
#define MB_E_ALIGN     (-4)   /* pointer not 16-byte aligned where required  */
#define MB_E_MODEL     (-5)   /* model violates a supported-shape constraint */
#define MB_E_DEVICE    (-6)   /* device is not compute capability 10.x       */

/* blend-shape contraction precision (mb_mano_forward / mb_mano_backward `mode`) */
#define MB_MODE_FP32    0     /* FFMA, fp32 end to end (correctness anchor)             */
#define MB_MODE_F16X3   1     /* tcgen05 kind::f16, 2-way fp16 split, 3 products: fp32-accurate */
#define MB_MODE_F16     2     /* tcgen05 kind::f16, single product: fast, ~1e-5 m error  */

/* Model property, OR-ed into `mode` by the caller: the kinematic tree is MANO's wrist with five
 * chains of three joints (parents = -1,0,1,2,0,4,5,0,7,8,0,10,11,0,13,14 — every MANO pickle).
 * It lets large batches run the one-thread-per-hand pose kernels; mb_mano_model_flags computes it. */
#define MB_MODEL_CHAINS_5X3 0x100

/* Forward options, OR-ed into `mode` of mb_mano_forward:
 * MB_FWD_INFERENCE: no backward will follow — the fused forward keeps no rest-pose scratch in the workspace (a later
 *                   mb_mano_backward must then be called WITHOUT MB_BWD_WORKSPACE_VALID and recomputes it);
 * MB_FWD_FUSED / MB_FWD_UNFUSED: from 8 192 hands on (MANO tree, tensor-core modes) the forward has two implementations —
 *                   the fused blend-shape + skinning kernel with lane = vertex (csrc/vskin.cu: both contractions on the
 *                   tensor core, rest positions never leave the SM; with a backward to follow it also writes the rest-pose
 *                   scratch) and the blend-contraction + lane = hand skinning kernels.  The fused kernel is the default (the
 *                   measured faster one in both kinds of launch); MB_FWD_UNFUSED selects the two kernels (measurement /
 *                   cross-checking), MB_FWD_FUSED is accepted for symmetry. */
#define MB_FWD_INFERENCE 0x200
#define MB_FWD_FUSED     0x400
#define MB_FWD_UNFUSED   0x800

/* mb_mano_backward flags */
#define MB_BWD_WORKSPACE_VALID 1  /* workspace still holds the forward's intermediates for these inputs */

/* masked reductions */
#define MB_REDUCE_MPJPE_MM 0  /* mean ||d|| * 1000  (criterions/metrics.py:10-27) */
#define MB_REDUCE_L2       1  /* mean ||d||^2       (criterions/loss.py:10-25)    */
#define MB_VIS_F32 0
#define MB_VIS_U8  1

typedef struct CUstream_st* mb_stream_t;

MB_API int         mb_abi_version(void);
MB_API const char* mb_error_string(int code);
/* 0 if the current device can run this library (sm_100), else MB_E_DEVICE / cudaError. */
MB_API int         mb_check_device(void);

/* ------------------------------------------------------------------ MANO ---
 * Constants.  Replaces ManoLayer.__init__ (network/sub_modules/MANOLayer.py:52-80):
 * the eight fp32 constants are folded/laid out once into one blob.
 *   basis      [148][2334]  rows 0-9 shapedirs, 10-144 posedirs, 145 v_template, 146-147 zero
 *   j0 [16][3], jb [16][3][10]   J_regressor folded with v_template / shapedirs (MANOLayer.py:139)
 *   pca [nc][45] = hands_components[:nc], pose_mean [45]
 *   skin_w [778][8] / skin_b [778][8] ELL skinning weights and bone ids (zero padded)
 *   parents [16] (parents[0] = -1, parents[i] < i)
 * mb_mano_pack_constants writes `mb_mano_blob_bytes()` bytes of HOST memory; the
 * caller copies them to the device once (16-byte aligned). */
MB_API size_t mb_mano_blob_bytes(void);
MB_API int    mb_mano_pack_constants(const float* basis, const float* j0, const float* jb,
                              const float* pca, int nc, const float* pose_mean,
                              const float* skin_w, const int32_t* skin_b,
                              const int32_t* parents, void* host_blob);

/* Re-validates the skinning program inside a packed HOST blob (block partition + shared-memory slot
 * schedule) and reports stats4 = {(block, bone) entries, bone loads per sweep, blocks, max bones per block}. */
MB_API int    mb_mano_skin_program_stats(const void* host_blob, int32_t* stats4);

/* Model property bits (MB_MODEL_*) of a kinematic tree, to be OR-ed into `mode`. */
MB_API int    mb_mano_model_flags(const int32_t* parents);

/* Bytes of device scratch mb_mano_forward / mb_mano_backward need for B hands. */
MB_API size_t mb_mano_workspace_bytes(int B, int mode);

/* Replaces ManoLayer.forward == rot_pose_beta_to_mesh (MANOLayer.py:122-208, :238-240).
 *   rot[B][3] global axis-angle, coeffs[B][nc] PCA pose coefficients, betas[B][10]
 *   -> verts[B][778][3], joints[B][21][3]   (metres; joints in the reference's 21-order)
 * verts may be NULL: only the 21 joints are produced (what every head of the
 * reference consumes, resnet50MANO.py:76,87) and the 778-vertex contraction is skipped. */
MB_API int mb_mano_forward(const void* blob, int nc,
                    const float* rot, const float* coeffs, const float* betas, int B, int mode,
                    float* verts, float* joints,
                    void* workspace, size_t workspace_bytes, mb_stream_t stream);
/* Diagnostics of the fused lane = vertex forward (csrc/vskin.cu; needs MB_MODEL_CHAINS_5X3 and a tensor-core mode, any B):
 * the same outputs as mb_mano_forward through that kernel, plus — dbg (nullable, device float[4][128][16]) — the blended
 * 3x4 transforms (12) and rest positions (3) of vertices 0..127 of hands 0..3 as the epilogue sees them.  variant: bit 0
 * swaps the leading / stride fields of the MN-major operand descriptor, bits 4-6 = number of split products of the
 * transform contraction (0 = default 4). */
MB_API int mb_mano_forward_debug(const void* blob, int nc,
                    const float* rot, const float* coeffs, const float* betas, int B, int mode,
                    float* verts, float* joints, void* workspace, size_t workspace_bytes,
                    float* dbg, int variant, mb_stream_t stream);

/* Optional per-hand scale and translation of the layer's outputs — the `transl=` / `scale=` keyword extension of
 * ManoLayer.forward (the reference's layer has none, MANOLayer.py:238; its callers apply the equivalent post-ops on
 * the outputs: resnet50MANO.py:77-81, Resnet50MANO3DHandPose.py:35-60):  p' = scale[h] * p + transl[h], in place, for
 * verts[B][778][3] (nullable) and joints[B][21][3]; scale[B] / transl[B][3] may each be NULL. */
MB_API int mb_affine_forward(float* verts, float* joints, const float* scale, const float* transl, int B, mb_stream_t stream);
/* Its backward, to be called AFTER mb_mano_backward was run on the same upstream gradients: g_transl[h] = sum of the
 * hand's g, g_scale[h] = sum <g, (p' - transl) / scale> from the transformed outputs, and — the MANO backward being
 * linear in g — the parameter gradients g_rot / g_coeffs / g_betas are multiplied by scale[h] in place. */
MB_API int mb_affine_backward(const float* g_verts, const float* g_joints, const float* verts_out, const float* joints_out,
                       const float* scale, const float* transl, int B, int nc, float* g_scale, float* g_transl,
                       float* g_rot, float* g_coeffs, float* g_betas, mb_stream_t stream);

/* Replaces the autograd tape of MANOLayer.py:122-208.
 *   g_verts[B][778][3] (NULL = zero: the heads' joints-only case), g_joints[B][21][3]
 *   -> g_rot[B][3], g_coeffs[B][nc], g_betas[B][10]
 * Stateless: intermediates are recomputed from the inputs into `workspace` unless
 * MB_BWD_WORKSPACE_VALID says the forward's are still there.
 * Deviation (documented): at |axis-angle| -> 0 the analytic limit is returned where
 * the reference's autograd yields NaN (r/theta at MANOLayer.py:91). */
MB_API int mb_mano_backward(const void* blob, int nc,
                     const float* rot, const float* coeffs, const float* betas,
                     const float* g_verts, const float* g_joints, int B, int mode, int flags,
                     float* g_rot, float* g_coeffs, float* g_betas,
                     void* workspace, size_t workspace_bytes, mb_stream_t stream);

/* Stand-alone linear-blend-skinning stage (MANOLayer.py:177-205) on caller-provided
 * rest-pose vertices: v_posed[B][pitch] (pitch >= 2334 floats, multiple of 4),
 * bone[B][16][12] (row-major 3x4 [R|t] per bone, global rotation already folded in)
 * -> verts[B][778][3]; tips (verts 333,444,672,555,745) -> joints slots 4,8,12,16,20
 * when joints != NULL.  The inputs are first re-laid out into the hand-minor scratch
 * the skinning kernel consumes (workspace of mb_lbs_workspace_bytes(B) bytes); only the
 * skinning kernel itself is attributed to the lbs_fwd profiling stage. */
MB_API size_t mb_lbs_workspace_bytes(int B);
MB_API int mb_lbs_forward(const void* blob, const float* v_posed, int pitch, const float* bone, int B,
                   float* verts, float* joints, void* workspace, size_t workspace_bytes, mb_stream_t stream);

/* -------------------------------------------------------------------- FK ---
 * Replaces ForwardKinematics.forward (network/sub_modules/forwardKinematicsLayer.py:147-330)
 * including convert_rel_normalized_to_absolute (:333-358) and the projection
 * batch_project_xyz_to_uv (utils/coordinate_trans.py:29-73).
 *   root_angles[B][3], other_angles[B][23], bone_lengths[B][20], K[B][3][3],
 *   index_root_bone_length[B], kp_coord_xyz_root[B][3]
 *   -> xyz[B][21][3], uv[B][21][2]
 * swap_order != 0 applies the per-finger (i,i+3),(i+1,i+2) swap the reference does
 * when config.joint_order_switched is False (:324-327). */
MB_API int mb_fk_forward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                  const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                  int B, int swap_order, float* xyz, float* uv, mb_stream_t stream);

/* Gradient w.r.t. the three tensors that require grad in the heads.
 * g_xyz / g_uv may each be NULL (treated as zero). */
MB_API int mb_fk_backward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                   const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                   const float* g_xyz, const float* g_uv, int B, int swap_order,
                   float* g_root_angles, float* g_other_angles, float* g_bone_lengths, mb_stream_t stream);

/* The FK heads' training tail in one launch per direction: ForwardKinematics.forward (as mb_fk_forward) followed by
 * LossCalculation.compute_3d_coord_loss / compute_uv_coord_loss (criterions/loss.py:83-87 -> L2Loss, :10-25) on its two
 * outputs, as network/TwoDimHandPoseWithFK.py / ThreeDimHandPose.py + trainval.py:328-358 combine them.
 *   gt_xyz[B][21][3], gt_uv[B][21][2], keypoint_vis[B][21] (fp32, non-zero = visible); flags = MB_HEAD_XYZ | MB_HEAD_UV
 *   selects the terms (declared with the MANO heads' tail below); a term that is off is 0 and its ground truth may be NULL.
 * Forward: xyz[B][21][3], uv[B][21][2] (what the head returns) and losses = device float[2] = {loss_xyz, loss_uv}; the
 * reductions run inside the FK kernel (fp32 per-joint sums of squares, fp64 above), the last warp divides.
 * Backward: g_losses = device float[2], the upstream gradients of the two terms; the gradients w.r.t. xyz / uv are formed
 * inside the FK backward kernel from the recomputed keypoints and never stored.  The workspace
 * (mb_fk_loss_workspace_bytes: sums, visible-joint counts, ticket) must reach the backward untouched.
 * Same results as mb_fk_forward -> mb_masked_joint_reduce x2 and mb_masked_l2_backward x2 -> mb_fk_backward. */
MB_API size_t mb_fk_loss_workspace_bytes(int B);
MB_API int mb_fk_loss_forward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                       const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                       const float* gt_xyz, const float* gt_uv, const float* keypoint_vis, int B, int swap_order,
                       int flags, float* xyz, float* uv, float* losses, void* workspace, size_t workspace_bytes,
                       mb_stream_t stream);
MB_API int mb_fk_loss_backward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                        const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                        const float* gt_xyz, const float* gt_uv, const float* keypoint_vis, int B, int swap_order,
                        int flags, const float* g_losses, float* g_root_angles, float* g_other_angles,
                        float* g_bone_lengths, const void* workspace, size_t workspace_bytes, mb_stream_t stream);

/* Replaces batch_project_xyz_to_uv (utils/coordinate_trans.py:29-73) for xyz[B][N][3]:
 * p = K xyz; p_z == 0 -> 1e-10; uv = p_xy / p_z. */
MB_API int mb_project_uv_forward(const float* xyz, const float* K, int B, int N, float* uv, mb_stream_t stream);
MB_API int mb_project_uv_backward(const float* xyz, const float* K, const float* g_uv, int B, int N,
                           float* g_xyz, mb_stream_t stream);

/* Replaces match_mano_to_RHD (network/Resnet50MANO3DHandPose.py:35-60; the same body at
 * network/MANO3DHandPose.py:30-55) followed by batch_project_xyz_to_uv (Resnet50MANO3DHandPose.py:73)
 * in one pass over joints[B][21][3]:
 *   per-finger joint reversal when swap_order != 0 (the reference's `not config.joint_order_switched`),
 *   r = p - p_0, n = r / ||r_12||, x = n * index_root_bone_length + kp_coord_xyz_root, uv = project(K, x).
 * rel_normalized and uv may be NULL (not wanted); K may be NULL when uv is.  Unlike the reference the
 * input is not permuted in place.  ||r_12|| == 0 divides by zero exactly as the reference does. */
MB_API int mb_joint_epilogue_forward(const float* joints, const float* index_root_bone_length,
                              const float* kp_coord_xyz_root, const float* K, int B, int swap_order,
                              float* rel_normalized, float* xyz, float* uv, mb_stream_t stream);
/* Gradients w.r.t. joints (required), index_root_bone_length and kp_coord_xyz_root (each may be NULL);
 * g_rel / g_xyz / g_uv may each be NULL (treated as zero). */
MB_API int mb_joint_epilogue_backward(const float* joints, const float* index_root_bone_length,
                               const float* kp_coord_xyz_root, const float* K, const float* g_rel,
                               const float* g_xyz, const float* g_uv, int B, int swap_order,
                               float* g_joints, float* g_scale, float* g_root, mb_stream_t stream);

/* ------------------------------------------------ keypoint re-parameterisations ---
 * Batched, forward-only replacements of the transforms the reference applies per sample in its
 * dataloaders (dataloaderRHD.py:242-250); all tensors [B][21][3] fp32.
 * mb_bone_rel_trafo      utils/relative_trafo.py:167-216  xyz -> (length, angle_x, angle_y) per bone
 * mb_bone_rel_trafo_inv  utils/relative_trafo.py:219-270  the inverse
 * mb_canonical_trafo     utils/canonical_trafo.py:93-159  root at the origin, joint 12 on the y axis, joint 20 in
 *                        the xy-plane; total_rot_mat[B][3][3] may be NULL; cond_right[B] (bytes, may be NULL)
 *                        additionally applies flip_right_hand (:163-184) to the flagged hands
 * mb_flip_right_hand     utils/canonical_trafo.py:163-184 for xyz[B][N][3]; cond_right is [B] or, with
 *                        cond_per_joint != 0, [B][N]
 * mb_mirror_hand         the same sign flip on a chosen axis (0 = x, 1 = y, 2 = z): axis 0 with cond = "is a left
 *                        hand" is the left -> right mirroring of dataloader/RHD/dataloaderRHD.py:224-225 */
MB_API int mb_bone_rel_trafo(const float* coords_xyz, int B, float* coords_rel, mb_stream_t stream);
MB_API int mb_bone_rel_trafo_inv(const float* coords_rel, int B, float* coords_xyz, mb_stream_t stream);
MB_API int mb_canonical_trafo(const float* coords_xyz, const unsigned char* cond_right, int B, float* coords_can,
                       float* total_rot_mat, mb_stream_t stream);
MB_API int mb_flip_right_hand(const float* coords_xyz, const unsigned char* cond_right, int B, int N, int cond_per_joint,
                       float* out, mb_stream_t stream);
MB_API int mb_mirror_hand(const float* coords_xyz, const unsigned char* cond, int B, int N, int cond_per_joint, int axis,
                   float* out, mb_stream_t stream);

/* ------------------------------------------------------------ viewpoint epilogue ---
 * Replaces _get_rot_mat (utils/general.py:191-226) and its consumer in the canonical-pose heads
 * (network/Hand3DPoseNet.py:41-50, network/Hand3DPosePriorNetwork.py:38-40):
 *   theta = sqrt(ux^2 + uy^2 + uz^2 + 1e-8), n = u * (1 / theta), R = cos I + (1 - cos) n n^T + sin [n]x
 *   rel_normed[B][21][3] = can_xyz[B][21][3] @ R            (row vectors times R)
 *   xyz = rel_normed * index_root_bone_length + kp_coord_xyz_root,  uv = project(K, xyz)   (inference branch)
 * ux, uy, uz are [B] (the reference's [B,1]).  Every output may be NULL (not wanted); can_xyz may be NULL when
 * only rot_mat is wanted; uv needs xyz. */
MB_API int mb_viewpoint_forward(const float* can_xyz, const float* ux, const float* uy, const float* uz,
                         const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K, int B,
                         float* rot_mat, float* rel_normed, float* xyz, float* uv, mb_stream_t stream);
/* Gradients of (rot_mat, rel_normed) w.r.t. (can_xyz, ux, uy, uz); g_rot / g_rel may be NULL (zero);
 * g_can may be NULL; g_ux, g_uy, g_uz are all NULL or all given. */
MB_API int mb_viewpoint_backward(const float* can_xyz, const float* ux, const float* uy, const float* uz,
                          const float* g_rot, const float* g_rel, int B, float* g_can, float* g_ux, float* g_uy,
                          float* g_uz, mb_stream_t stream);

/* ------------------------------------------------------------ reductions ---
 * Replaces MPJPE.forward (criterions/metrics.py:10-27) and L2Loss.forward
 * (criterions/loss.py:10-25): global mean over the visible joints of the batch,
 * 0 when none is visible — decided on the device, no host sync.
 *   pred[n_joints][dim], gt[n_joints][dim], vis[n_joints] (fp32 non-zero = visible, or u8)
 *   dim = components per joint, 1..4: 3 for the xyz loss / MPJPE, 2 for the uv loss
 *         (LossCalculation.compute_uv_coord_loss, criterions/loss.py:86-87, sends [B,21,2] through L2Loss)
 *   accum: device double[2] scratch {sum, count} (overwritten);  out: device float[1] */
MB_API int mb_masked_joint_reduce(const float* pred, const float* gt, const void* vis, int vis_kind,
                           long long n_joints, int dim, int kind, double* accum, float* out, mb_stream_t stream);
/* d(L2)/d(pred) = g_out * 2 (pred-gt) vis / count, with `accum` as left by the forward. */
MB_API int mb_masked_l2_backward(const float* pred, const float* gt, const void* vis, int vis_kind,
                          long long n_joints, int dim, const double* accum, const float* g_out,
                          float* g_pred, mb_stream_t stream);

/* Replaces LossCalculation.compute_regularization_loss (criterions/loss.py:113-117):
 *   out = (||theta||_F + alpha_beta * ||beta||_F) / 100 over ALL n_theta / n_beta elements (the reference's
 *   torch.norm of the whole [B,nc] / [B,10] tensors; alpha_beta = 10 there).
 *   accum: device double[2] scratch {sum theta^2, sum beta^2} (overwritten; the backward reads it).
 * Backward: g_theta = g_out theta / (100 ||theta||), g_beta = g_out alpha_beta beta / (100 ||beta||), 0 at a zero norm. */
MB_API int mb_regulariser_forward(const float* theta, long long n_theta, const float* beta, long long n_beta,
                           float alpha_beta, double* accum, float* out, mb_stream_t stream);
MB_API int mb_regulariser_backward(const float* theta, long long n_theta, const float* beta, long long n_beta,
                            float alpha_beta, const double* accum, const float* g_out, float* g_theta,
                            float* g_beta, mb_stream_t stream);

/* Replaces LossCalculation.compute_hand_mask_loss (criterions/loss.py:92-111):
 *   uv -> int64 by truncation, clamped to [0, W-1] on both axes (as the reference does), samples of
 *   hand_mask[B][H][W] (fp32 or u8, MB_VIS_*) at pred_uv[B][N][2] and gt_uv[B][N][2] (u = column, v = row);
 *   out = 1 - sum(pred samples) / (sum(gt samples) + 1e-8).  No gradient exists (integer indexing).
 *   Needs W <= H (the reference would raise an IndexError for a clamped row >= H); accum: device double[2]. */
MB_API int mb_hand_mask_loss(const float* pred_uv, const float* gt_uv, const void* hand_mask, int mask_kind, int B, int N,
                      int H, int W, double* accum, float* out, mb_stream_t stream);

/* ------------------------------------------------- the MANO heads' tail in one call ---
 * MANO parameters -> 21 joints (joints-only kernels) [-> scale * p + transl, resnet50MANO.py:77-81]
 * [-> match_mano_to_RHD, Resnet50MANO3DHandPose.py:35-60] -> batch_project_xyz_to_uv (:71-73) ->
 * L2Loss on xyz and on uv (criterions/loss.py:10-25, :83-87) + compute_regularization_loss (:113-117), as the
 * training loop combines them (trainval.py:328-358).  One call enqueues the kernels of every piece back to back on
 * `stream` (no host round trip, capturable in a CUDA graph); `flags` selects the terms:
 *   MB_HEAD_XYZ / MB_HEAD_UV / MB_HEAD_REG  which of losses[0..2] = {loss_xyz, loss_uv, loss_regularization} are
 *                                           computed (the others are 0);  MB_HEAD_MATCH  apply match_mano_to_RHD
 *                                           (swap_order as in mb_joint_epilogue_forward), else xyz = the joints.
 * transl[B][3] / scale[B] may be NULL; index_root_bone_length[B] / kp_coord_xyz_root[B][3] are needed with
 * MB_HEAD_MATCH only; keypoint_vis[B][21] fp32 (non-zero = visible); gt_xyz[B][21][3], gt_uv[B][21][2].
 * Outputs: joint_xyz21[B][21][3], uv21[B][21][2] (what the head returns), losses (device float[3]).
 * The workspace (mb_mano_head_loss_workspace_bytes) and the two outputs must reach the backward untouched.
 * Backward: g_losses = device float[3], the upstream gradients of the three terms -> g_rot[B][3], g_coeffs[B][nc],
 * g_betas[B][10] and (each nullable) g_transl[B][3], g_scale[B]. */
#define MB_HEAD_XYZ   1
#define MB_HEAD_UV    2
#define MB_HEAD_REG   4
#define MB_HEAD_MATCH 8
MB_API size_t mb_mano_head_loss_workspace_bytes(int B);
MB_API int mb_mano_head_loss_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                              const float* transl, const float* scale,
                              const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K,
                              const float* gt_xyz, const float* gt_uv, const float* keypoint_vis,
                              int B, int mode, int flags, int swap_order, float alpha_beta,
                              float* joint_xyz21, float* uv21, float* losses,
                              void* workspace, size_t workspace_bytes, mb_stream_t stream);
MB_API int mb_mano_head_loss_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                               const float* transl, const float* scale,
                               const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K,
                               const float* gt_xyz, const float* gt_uv, const float* keypoint_vis,
                               int B, int mode, int flags, int swap_order, float alpha_beta,
                               const float* joint_xyz21, const float* uv21, const float* g_losses,
                               float* g_rot, float* g_coeffs, float* g_betas, float* g_transl, float* g_scale,
                               void* workspace, size_t workspace_bytes, mb_stream_t stream);

/* ----------------------------------------------------------- fitting loop ---
 * One Adam update on a flat fp32 parameter array (torch.optim.Adam semantics, no
 * weight decay, no amsgrad): used by the batched MANO fitting loop (BASELINE config 5). */
/* One whole iteration of that loop in a single kernel (SURVEY 8b `mano_fit_step`): joints-only forward, gradient of
 * L2Loss (criterions/loss.py:10-25) against target_joints[B][21][3] under keypoint_vis[B][21] (fp32, non-zero =
 * visible), joints-only backward, gradient of the regulariser (:113-117, when regularize != 0) and the Adam update
 * (step = 1, 2, ...) of params / exp_avg / exp_avg_sq, each laid out rot[B][3] | coeffs[B][nc] | betas[B][10].
 *   globals  (device, in) : {N_vis, sum theta^2, sum beta^2} over ALL ranks for the parameters as they are on entry —
 *                           they depend on the mask and the parameters only, so they come from the previous
 *                           iteration's partials (all-reduced by the caller when the batch is sharded)
 *   partials (device, out): {sum over this call's visible joints of |joint - target|^2, sum theta^2, sum beta^2 of
 *                           the UPDATED parameters} of this rank
 * `mode` must carry MB_MODEL_CHAINS_5X3 (MANO's tree), else MB_E_MODEL: use the separate calls above. */
MB_API int mb_mano_fit_step(const void* blob, int nc, float* params, float* exp_avg, float* exp_avg_sq,
                     const float* target_joints, const float* keypoint_vis, int B, int mode, const double* globals,
                     double* partials, float lr, float beta1, float beta2, float eps, int step, int regularize,
                     mb_stream_t stream);
/* Closes a fused iteration once the caller has all-reduced `partials` (mb_mano_fit_step's output): writes
 * out4 = {sum vis |d|^2, N_vis, sum theta^2, sum beta^2} and loss = out4[0] / N_vis + (sqrt(out4[2]) + 10 sqrt(out4[3])) / 100
 * for the iteration just done (norms of the parameters it started from; zero when regularize == 0), then moves the
 * updated parameters' norms partials[1..2] into globals[1..2] for the next iteration.  All pointers device doubles. */
MB_API int mb_fit_finalize(double* globals, const double* partials, int regularize, double* out4, double* loss,
                    mb_stream_t stream);
MB_API int mb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                 float lr, float beta1, float beta2, float eps, int step, mb_stream_t stream);

/* ------------------------------------------------------------ measurement ---
 * mb_launch_count: kernels launched by this library in this process so far.
 * mb_profile_enable(1): bracket every stage launch with CUDA events on its stream;
 * mb_profile_collect: synchronise on those events, write the summed milliseconds and launch
 * counts per stage (MB_N_STAGES entries each) and reset.  Stage order: pose_fwd, blend_fwd,
 * lbs_fwd, lbs_bwd, blend_bwd, pose_bwd, joints_only_fwd, joints_only_bwd, fk_fwd, fk_bwd, fused_fwd. */
#define MB_N_STAGES 11
MB_API long long mb_launch_count(void);
MB_API void      mb_profile_enable(int on);
MB_API int       mb_profile_collect(double* ms, long long* counts);

#ifdef __cplusplus
}
#endif
#endif /* MANO_B200_H */
