// kernels_bi_256.cu — bi_scan_kernel<256, R, ...>
#include "kernels_bi_scan.cuh"

namespace tspb {

cudaError_t launch_bi_scan_256(const BiArgs &a, int R, int grid, bool pdl, cudaStream_t st) {
    if (R == 16) return launch_bi_tr<256, 16, false>(a, grid, pdl, st);
    if (R == 8) return launch_bi_tr<256, 8, false>(a, grid, pdl, st);
    if (R == 4) return launch_bi_tr<256, 4, false>(a, grid, pdl, st);
    if (R == 2) return launch_bi_tr<256, 2, false>(a, grid, pdl, st);
    return cudaErrorInvalidValue;
}

}  // namespace tspb
