// blend_tc.cu — tcgen05 (5th-gen tensor core) path of the blend-shape contractions.
// Placeholder until the tensor-core kernels land: the modes report MB_E_RANGE.
#include "common.cuh"
namespace mb {
size_t blend_tc_blob_bytes() { return 0; }
void blend_tc_pack(const float*, void*) {}
int launch_blend_tc_forward(const void*, const float*, float*, int, int, cudaStream_t) { return MB_E_RANGE; }
int launch_blend_tc_backward(const void*, const float*, float*, int, int, cudaStream_t) { return MB_E_RANGE; }
}  // namespace mb
