// Batched (N >= 16) mul_mat on tcgen05 tensor cores -- placeholder until the kernel lands.
#include "ggb_internal.h"
namespace ggb {
bool gemm_supported(int, int64_t, int64_t, int64_t, int64_t, const void *) { return false; }
size_t gemm_workspace_bytes(int, int64_t, int64_t, int64_t) { return 0; }
int launch_gemm(const GemmArgs &, void *, cudaStream_t) { return set_error(GGB_E_UNSUPPORTED, "batched tensor-core path not built"); }
}
