// K1 for p > 56 — placeholder translation unit until the tiled large-p kernel lands.
#include "common.cuh"
namespace dlsa {
size_t irls_large_workspace_bytes(int, int, int64_t) { return 0; }
int irls_large_launch(const double*, int64_t, const double*, const int64_t*, int, int, int, int, const double*,
                      const int*, double*, double*, double*, int64_t, void*, size_t, cudaStream_t) {
    set_error("irls pass: p > 56 not built yet");
    return DLSA_E_UNSUPPORTED;
}
}  // namespace dlsa
