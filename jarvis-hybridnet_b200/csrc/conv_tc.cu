// Stage 2, bf16 tensor-core path (tcgen05 / TMEM / TMA).  Placeholder until the kernels land: every
// entry point fails loudly, there is no fallback to the fp32 path.
#include "v2v.cuh"

namespace jhn {
int tc_create(jhn_v2v *, const float *const *, cudaStream_t) { return fail(JHN_ERR_ARG, "bf16 tensor-core path not built yet"); }
void tc_destroy(jhn_v2v *) {}
size_t tc_workspace(const jhn_v2v *, int, int) { return 0; }
size_t tc_volume_bytes(const jhn_v2v *, int, int) { return 0; }
int tc_forward(const jhn_v2v *, const void *, int, int, int, float *, void *, size_t, cudaStream_t)
{
    return fail(JHN_ERR_ARG, "bf16 tensor-core path not built yet");
}
}  // namespace jhn
