// kernels_bi_p64.cu — bi_scan_kernel<64, R, ..., SHUF = false>: the plain 64-thread variants (option row_shuffle = 0)
#include "kernels_bi_scan.cuh"

namespace tspb {

cudaError_t launch_bi_scan_p64(const BiArgs &a, int R, int grid, bool pdl, cudaStream_t st) {
    if (R == 8) return launch_bi_tr<64, 8, false>(a, grid, pdl, st);
    if (R == 4) return launch_bi_tr<64, 4, false>(a, grid, pdl, st);
    if (R == 2) return launch_bi_tr<64, 2, false>(a, grid, pdl, st);
    return cudaErrorInvalidValue;
}

}  // namespace tspb
