// kernels_bi_s64.cu — bi_scan_kernel<64, R, ..., SHUF = true>: the row-shuffle variants (the shapes the engine picks by default)
#include "kernels_bi_scan.cuh"

namespace tspb {

cudaError_t launch_bi_scan_s64(const BiArgs &a, int R, int grid, bool pdl, cudaStream_t st) {
    if (R == 8) return launch_bi_tr<64, 8, true>(a, grid, pdl, st);
    if (R == 4) return launch_bi_tr<64, 4, true>(a, grid, pdl, st);
    if (R == 2) return launch_bi_tr<64, 2, true>(a, grid, pdl, st);
    return cudaErrorInvalidValue;
}

}  // namespace tspb
