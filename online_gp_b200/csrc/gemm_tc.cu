// fp32 tensor-core (tcgen05, kind::tf32) versions of the two real contractions: Gram and panel right-multiply.
// Placeholder until the tcgen05 kernels land: reports "unsupported" (3) so panel.cu falls through to the SIMT path.
#include "common.cuh"

namespace wiski {
int tc_gram_f32(const float*, const float*, int64_t, int64_t, int64_t, float*, float*, cudaStream_t) { return 3; }
int tc_panel_rmul_f32(const float*, int64_t, int64_t, const float*, int64_t, float*, cudaStream_t) { return 3; }
int64_t tc_gram_work_elems(int64_t, int64_t, int64_t) { return 0; }
}  // namespace wiski
