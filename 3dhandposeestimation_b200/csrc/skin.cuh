// skin.cuh — launchers of the register-blocked skinning kernels and the host-side skin program (skin.cu).
#pragma once
#include "common.cuh"

namespace mb {

// Builds the skin program sections of the blob and coord_map[SK_TMPL_PAD]: GEMM column (block order) ->
// original coordinate (vertex*3 + c), -1 for padding.  Returns 0 or MB_E_MODEL.
int skin_pack(const float* skin_w, const int32_t* skin_b, void* host_blob, int32_t* coord_map);

// replays the packed program's slot schedule on the host; stats[4] = entries, loads per sweep, blocks, max bones per block
int skin_program_check(const void* host_blob, int32_t* stats);

// v_posed_t [groups][SK_NCOORD][32], bone_t [groups][16][32][12] -> verts[B][778][3], tips -> joints
int launch_skin_forward(const void* blob, const float* v_posed_t, const float* bone_t, int B,
                        float* verts, float* joints, cudaStream_t s);
// exactly one of dv_t (fp32, hand-minor block order) / dvp (bf16 hi+mid UMMA tiles) is non-NULL
// dbone: [B][16][12] rows, or hand-minor [groups][192][32] when dbone_hand_minor != 0
// dparts: WorkLayout::dparts scratch, needed when skin_segments_per_unit(groups, SKB_SWEEPERS) < SK_NSEG
int launch_skin_backward(const void* blob, const float* v_posed_t, const float* bone_t, const float* g_verts,
                         const float* g_joints, int B, float* dv_t, unsigned char* dvp, float* dbone, int dbone_hand_minor,
                         float* dparts, cudaStream_t s);
// layout conversions used by the fp32 anchor mode and the stand-alone mb_lbs_forward
int launch_rows_to_t(const void* blob, const float* rows, int pitch, int B, float* t, cudaStream_t s);
int launch_t_to_rows(const void* blob, const float* t, int pitch, int B, float* rows, cudaStream_t s);
int launch_bone_rows_to_t(const float* bone, int B, float* bone_t, cudaStream_t s);

}  // namespace mb
