// kernels_fi.cu — first-improvement 2-opt (reference src/heuristics.c:438-502 alg_2opt).
//
// The reference sweeps node-index pairs (i<j) row-major and applies EVERY improving move at once, then
// carries on from (i, j+1) with the modified tour; sweeps repeat until one brings no gain (:492).
// Exact replay on a GPU: one launch = "find the first pair at or after the cursor, in that same order,
// whose exact delta is negative" (a grid-wide min over the linear index i*n+j), apply it, advance the
// cursor.  Blocks take rows i = cursor_row + blockIdx, +gridDim, ...; a row stops at its first hit and
// rows that start after an already published hit are skipped, so the work wasted past the hit is
// bounded by gridDim rows.  The scan is in NODE space (nrec/nds/nsucc) because the order that matters
// is the node-index order.  FP32 filter + FP64 exact check exactly as in the BI kernel.
#include "tsp_state.cuh"

namespace tspb {

constexpr int FI_THREADS = 256;

template <bool ATT>
__device__ __forceinline__ float fi_dist32(float ax, float ay, float bx, float by) {
    float dx = ax - bx, dy = ay - by;
    float s = fmaf(dy, dy, dx * dx);
    if (ATT) s *= 0.1f;
    return sqrt_approx(s);
}

// FP32_OK = false: no filter, every pair is evaluated exactly (GEO, matrix mode, oversized coordinates).
template <bool ATT, bool EXACT32, bool FP32_OK>
__global__ void __launch_bounds__(FI_THREADS) fi_search_kernel(const InstDev I, const TourDev T) {
    __shared__ int s_minj;
    __shared__ int s_last;
    __shared__ unsigned long long s_found;
    Ctl *ctl = T.ctl;
    pdl_launch_dependents();  // the apply launch may queue up behind this kernel
    pdl_wait();               // the refresh launch of the previous move is complete
    if (ctl->done) {
        // a capped run stops right after publishing a move: make sure later apply launches are no-ops
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ap_valid = 0;
        return;
    }
    const int n = T.n;
    const int tid = threadIdx.x;
    const int i0 = ctl->cur_i, j0 = ctl->cur_j;
    const float thrW = -1.0f + I.W;  // candidates: exact delta <= -1

    for (int row = i0 + (int)blockIdx.x; row < n - 1; row += (int)gridDim.x) {
        if (tid == 0) {
            s_found = *((volatile unsigned long long *)&ctl->fi_found);
            s_minj = 0x7fffffff;
        }
        __syncthreads();
        if (s_found < (unsigned long long)row * (unsigned long long)n) break;  // an earlier pair already won
        const float4 ri = T.nrec[row];
        const float dsi = T.nds[row];
        const int si = T.nsucc[row];
        const int jstart = (row == i0) ? j0 : row + 1;
        bool stop = false;
        for (int jb = jstart; jb < n && !stop; jb += FI_THREADS) {
            const int j = jb + tid;
            bool hit = false;
            if (j < n) {
                const int sj = T.nsucc[j];
                // reference heuristics.c:471: skip a1==b1 (impossible in a tour), a==b1, b==a1
                if (sj != row && si != j && si != sj) {
                    bool cand = true;
                    float4 rj;
                    float dsj = 0.f;
                    if (FP32_OK) {
                        rj = T.nrec[j];
                        dsj = T.nds[j];
                        float q = fi_dist32<ATT>(ri.x, ri.y, rj.x, rj.y) + fi_dist32<ATT>(ri.z, ri.w, rj.z, rj.w) - dsi - dsj;
                        cand = (q <= thrW);
                    }
                    if (cand) {
                        long long delta;
                        if (FP32_OK && EXACT32) {
                            delta = exact_dist(I.metric, make_double2((double)ri.x, (double)ri.y),
                                               make_double2((double)rj.x, (double)rj.y)) +
                                    exact_dist(I.metric, make_double2((double)ri.z, (double)ri.w),
                                               make_double2((double)rj.z, (double)rj.w)) -
                                    (long long)dsi - (long long)dsj;
                        } else {
                            delta = dist_nodes(I, row, j) + dist_nodes(I, si, sj) - (long long)dsi - (long long)T.nds[j];
                        }
                        hit = delta < 0;
                    }
                }
            }
            // also leave the row when somebody else published an earlier pair (polled every 8 chunks)
            bool bail = false;
            if (tid == 0 && (((jb - jstart) / FI_THREADS) & 7) == 7)
                bail = *((volatile unsigned long long *)&ctl->fi_found) < (unsigned long long)row * (unsigned long long)n;
            if (hit) atomicMin(&s_minj, j);
            if (__syncthreads_or((int)(hit || bail))) stop = true;
        }
        if (stop) {
            if (tid == 0 && s_minj != 0x7fffffff)
                atomicMin(&ctl->fi_found, (unsigned long long)row * (unsigned long long)n + (unsigned long long)s_minj);
            break;
        }
        __syncthreads();
    }

    // ---- last block: apply the winning move / close the sweep ------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned tk = atomicAdd(&ctl->ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long f = *((volatile unsigned long long *)&ctl->fi_found);
    int ci = 0, cj = 0;
    bool sweep_end = false;
    if (f != FI_NONE) {
        const int i = (int)(f / (unsigned long long)n), j = (int)(f % (unsigned long long)n);
        if (tid == 0) {
            const long long delta = move_delta_nodes(I, T, i, j);
            if (delta >= 0) ctl->error = 1;  // cannot happen: the searching thread saw delta < 0
            publish_move(T, i, j, delta);    // reference heuristics.c:476-486; applied by the next two launches
            ctl->sweep_moves += 1;
            ctl->pairs_swept += (long long)(f - ((unsigned long long)i0 * n + j0)) + 1;
        }
        ci = i;
        cj = j + 1;
        if (cj >= n) { ci = i + 1; cj = ci + 1; }
        if (ci >= n - 1) sweep_end = true;
    } else {
        sweep_end = true;
        if (tid == 0) ctl->ap_valid = 0;
        if (tid == 0) ctl->pairs_swept += (long long)((unsigned long long)(n - 1) * n - ((unsigned long long)i0 * n + j0));
    }
    if (tid == 0) {
        ctl->launches += 1;
        if (sweep_end) {
            ctl->passes += 1;
            if (ctl->sweep_moves == 0) ctl->done = 1;  // reference heuristics.c:492: the sweep brought no gain
            ctl->sweep_moves = 0;
            ci = 0;
            cj = 1;
        }
        if (ctl->max_moves >= 0 && ctl->moves >= ctl->max_moves) ctl->done = 1;
        ctl->cur_i = ci;
        ctl->cur_j = cj;
        ctl->fi_found = FI_NONE;
        ctl->ticket = 0;
    }
}

cudaError_t launch_fi_search(const InstDev &I, const TourDev &T, int grid, bool pdl, cudaStream_t st) {
    const bool att = (I.metric == M_ATT);
    const bool ex = I.exact32 != 0;
    const dim3 g(grid), b(FI_THREADS);
    if (!I.fp32_ok) return launch_maybe_pdl(fi_search_kernel<false, false, false>, g, b, 0, st, pdl, I, T);
    if (att && ex) return launch_maybe_pdl(fi_search_kernel<true, true, true>, g, b, 0, st, pdl, I, T);
    if (att) return launch_maybe_pdl(fi_search_kernel<true, false, true>, g, b, 0, st, pdl, I, T);
    if (ex) return launch_maybe_pdl(fi_search_kernel<false, true, true>, g, b, 0, st, pdl, I, T);
    return launch_maybe_pdl(fi_search_kernel<false, false, true>, g, b, 0, st, pdl, I, T);
}

}  // namespace tspb
