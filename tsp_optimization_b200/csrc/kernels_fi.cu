// kernels_fi.cu — first-improvement 2-opt (reference src/heuristics.c:438-502 alg_2opt).
//
// The reference sweeps node-index pairs (i<j) row-major and applies EVERY improving move at once, then
// carries on from (i, j+1) with the modified tour; sweeps repeat until one brings no gain (:492).
// Exact replay on a GPU: one launch = "find the first pair at or after the cursor, in that same order,
// whose exact delta is negative" (a grid-wide min over the word (i << 32) | j), apply it, advance the
// cursor.  The scan is in NODE space (nrec/nlnk, doubly linked: tsp_state.cuh) because the order that matters
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
//
// Work distribution: the sweep's pairs in row-major order, counted from the cursor, are cut into SEGMENTS of
// FI_SEG_CHUNKS x FI_CHUNK consecutive pairs (crossing row ends).  Blocks draw segments in increasing order from an atomic
// counter; a block scans its segment chunk by chunk — FI_CHUNK = 256 threads x FI_U pairs, every thread's FI_U pairs are
// loaded before the first one is evaluated so that their L2 round trips overlap — and after every chunk the block votes:
// a hit inside the chunk (the smallest offset is the first one in order) is published with atomicMin on the (row << 32 | j)
// word and ends the block; so does a hit that ANOTHER block has published in the meantime at a pair in front of this
// chunk (thread 0 polls ctl->fi_found along with its own loads).  Blocks whose next segment starts behind a published hit
// stop as well.  Every segment before the winning one has been scanned to its end without a hit, so the minimum is exactly
// the reference's "first improving pair at or after the cursor"; the work past the hit is bounded by one chunk per block
// (while moves are dense — a move every ~10^5 pairs at the start of a 100 000-node sweep — the launch is over after its
// first chunk; without the poll every block finished its 4096-pair segment, 22 us instead of ~4).
constexpr int FI_U = 4;
constexpr int FI_CHUNK = FI_THREADS * FI_U;
constexpr int FI_SEG_CHUNKS = 4;

// (row, column) of the pair with absolute index A in the row-major enumeration of all i<j; A < n(n-1)/2
__device__ __forceinline__ void fi_locate(long long A, int n, int &row, int &j) {
    const double b = 2.0 * n - 1.0;
    long long r = (long long)((b - sqrt(b * b - 8.0 * (double)A)) * 0.5);
    if (r < 0) r = 0;
    if (r > n - 2) r = n - 2;
    while (r > 0 && fi_pairs_before_row(r, n) > A) --r;
    while (r < n - 2 && fi_pairs_before_row(r + 1, n) <= A) ++r;
    row = (int)r;
    j = (int)(r + 1 + (A - fi_pairs_before_row(r, n)));
}

// Several GPUs (world > 1): the segments are dealt round-robin over the ranks (rank r scans the segments r, r + world, ...),
// every rank finds the first improving pair among ITS segments, and the minimum over the ranks — exchanged through the same
// NVLink peer slots as the best-improvement keys (xchg_min), or min-allreduced by NCCL and finished by fi_finish_kernel — is
// the reference's pair: the rank that owns the winning segment scanned all of its earlier segments to their end.
struct FiShard {
    int rank, world;
    int finish_here;  // 1: the search kernel's last block publishes the move; 0: fi_finish_kernel does (NCCL exchange)
    int late;         // 1 + parity: single GPU, hits go to ctl->fi_sel[parity] and the apply launch takes it from there (no tail)
    XchgDev xchg;
};

// The search kernel's last block (several GPUs) / fi_finish_kernel: publish the move for the apply launch, then the bookkeeping.
__device__ __forceinline__ void fi_finish(const InstDev &I, const TourDev &T, unsigned long long f, int i0, int j0) {
    Ctl *ctl = T.ctl;
    if (f != FI_NONE) {
        const int i = (int)(f >> 32), j = (int)(f & 0xffffffffull);
        const long long delta = move_delta_nodes(I, T, i, j);
        if (delta >= 0) ctl->error = 1;  // cannot happen: the searching thread saw delta < 0
        publish_move(T, i, j, delta);    // reference heuristics.c:476-486; applied by the next launch
    }
    fi_advance(T, f, i0, j0);
    ctl->fi_found = FI_NONE;
    ctl->ticket = 0;
}

template <bool ATT, bool EXACT32, bool FP32_OK>
__global__ void __launch_bounds__(FI_THREADS) fi_search_kernel(const InstDev I, const TourDev T, const FiShard S) {
    __shared__ int s_seg;
    __shared__ int s_minhit;
    __shared__ int s_last;
    __shared__ unsigned long long s_found;
    Ctl *ctl = T.ctl;
    pdl_launch_dependents();  // the apply launch may queue up behind this kernel
    pdl_wait();               // the apply launch of the previous move is complete
    // one round trip for the run state (the loads are issued together; the exit test comes after)
    const int done = *((volatile int *)&ctl->done);
    const int i0 = *((volatile int *)&ctl->cur_i), j0 = *((volatile int *)&ctl->cur_j);
    // Several GPUs: while moves are dense (a hit within the first round of segments) an exchange per move costs more than the
    // search itself; every rank then searches alone — same pair, same tour — and the ranks only share the work of the long
    // searches of the late sweeps (ctl->fi_shard, set by fi_finish from the length of the previous search).
    const bool shard = S.world > 1 && *((volatile int *)&ctl->fi_shard) != 0;
    // late selection (S.late = 1 + parity): unless this search is sharded — then its last block exchanges the winner with the
    // other ranks and publishes the move as before — the hits go to ctl->fi_sel[parity] and the kernel has no tail
    const bool late_now = S.late != 0 && !shard;
    unsigned long long *found = late_now ? &ctl->fi_sel[S.late - 1] : &ctl->fi_found;
    if (S.late && blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->fi_mode[S.late - 1] = shard ? 1 : 0;  // tells the apply launch which of the two this search was
        // the position entry the previous apply launch parked (nobody reads pos[] before this kernel's last block)
        const int pn = *((volatile int *)&ctl->fi_pend_node);
        if (pn >= 0) {
            T.pos[pn] = ctl->fi_pend_pos;
            ctl->fi_pend_node = -1;
        }
    }
    const int s_rank = shard ? S.rank : 0, s_world = shard ? S.world : 1;
    if (done) {
        // a capped run stops right after publishing a move: make sure later apply launches are no-ops
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ap_valid = 0;
        return;
    }
    const int n = T.n;
    const int tid = threadIdx.x;
    const float thrW = -1.0f + I.W;  // candidates: exact delta <= -1
    const long long total = (long long)n * (n - 1) / 2;
    const long long A0 = fi_pairs_before_row(i0, n) + (j0 - i0 - 1);  // absolute index of the cursor pair
    constexpr long long SEG = (long long)FI_SEG_CHUNKS * FI_CHUNK;

    for (bool first = true;; first = false) {
        if (tid == 0) {
            // a block's first segment is its block index (no atomic on the critical path of the launch: while moves are
            // dense the launch ends in the first round); the following ones are drawn from the counter
            s_seg = first ? (int)blockIdx.x : (int)gridDim.x + (int)atomicAdd(&ctl->fi_seg, 1u);
            s_minhit = 0x7fffffff;
            s_found = first ? FI_NONE : *((volatile unsigned long long *)found);  // one read per block: the exit below must be uniform
        }
        __syncthreads();
        const long long Abase = A0 + ((long long)s_rank + (long long)s_world * (long long)s_seg) * SEG;
        if (Abase >= total) break;  // past the end of the sweep
        int rb, jb;
        fi_locate(Abase, n, rb, jb);
        if (s_found < fi_key(rb, jb)) break;  // an earlier pair already won
        // this thread's first pair of the segment (tid pairs behind the segment's first one); its further pairs follow at a
        // stride of 256
        long long A = Abase + tid;
        int row = n, j = 0;
        if (A < total) {
            row = rb;
            long long jn = (long long)jb + tid;
            while (row < n - 1 && jn >= n) { jn -= n; row += 1; jn += row + 1; }
            j = (int)jn;
        }
        bool stop = false;
        for (int c = 0; c < FI_SEG_CHUNKS && !stop; ++c) {
            // thread 0: has another block meanwhile published a hit in front of this chunk?  (requested with the chunk's loads)
            unsigned long long f_now = FI_NONE;
            const unsigned long long chunk_first = fi_key(row, j);
            if (tid == 0) f_now = *((volatile unsigned long long *)found);
            int rows[FI_U], js[FI_U];
            float4 ri[FI_U], rj[FI_U];
            float2 li[FI_U], lj[FI_U];  // {ds, succ}
#pragma unroll
            for (int k = 0; k < FI_U; ++k) {
                rows[k] = (A + (long long)k * FI_THREADS < total) ? row : -1;
                js[k] = j;
                // next pair of this thread: 256 pairs further in the row-major order
                long long jn = (long long)j + FI_THREADS;
                while (row < n - 1 && jn >= n) { jn -= n; row += 1; jn += row + 1; }
                j = (int)jn;
            }
#pragma unroll
            for (int k = 0; k < FI_U; ++k) {
                if (rows[k] >= 0) {
                    // the row record is loaded for every pair (usually the same address for the whole warp and for all k: one
                    // broadcast line): re-using pair k-1's registers would chain this pair's loads behind that pair's round trip
                    ri[k] = T.nrec[rows[k]];
                    li[k] = *reinterpret_cast<const float2 *>(&T.nlnk[rows[k]]);
                    lj[k] = *reinterpret_cast<const float2 *>(&T.nlnk[js[k]]);
                    if (FP32_OK) rj[k] = T.nrec[js[k]];
                }
            }
            int myhit = 0x7fffffff;
#pragma unroll
            for (int k = FI_U - 1; k >= 0; --k) {
                if (rows[k] < 0) continue;
                const int r = rows[k], jj = js[k];
                const int si = __float_as_int(li[k].y), sj = __float_as_int(lj[k].y);
                const float dsi = li[k].x, dsj = lj[k].x;
                // reference heuristics.c:471: skip a1==b1 (impossible in a tour), a==b1, b==a1
                if (sj == r || si == jj || si == sj) continue;
                bool cand = true;
                if (FP32_OK) {
                    const float q = fi_dist32<ATT>(ri[k].x, ri[k].y, rj[k].x, rj[k].y) + fi_dist32<ATT>(ri[k].z, ri[k].w, rj[k].z, rj[k].w) - dsi - dsj;
                    cand = (q <= thrW);
                }
                if (cand) {
                    long long delta;
                    if (FP32_OK && EXACT32) {
                        delta = exact_dist(I.metric, make_double2((double)ri[k].x, (double)ri[k].y),
                                           make_double2((double)rj[k].x, (double)rj[k].y)) +
                                exact_dist(I.metric, make_double2((double)ri[k].z, (double)ri[k].w),
                                           make_double2((double)rj[k].z, (double)rj[k].w)) -
                                (long long)dsi - (long long)dsj;
                    } else {
                        delta = dist_nodes(I, r, jj) + dist_nodes(I, si, sj) - (long long)dsi - (long long)dsj;
                    }
                    if (delta < 0) myhit = k * FI_THREADS + tid;  // k runs downwards: the smallest offset survives
                }
            }
            const bool hit = myhit != 0x7fffffff;
            if (hit) atomicMin(&s_minhit, c * FI_CHUNK + myhit);  // offset inside the segment == row-major order
            const bool overtaken = (tid == 0) && f_now < chunk_first;
            if (__syncthreads_or((int)(hit || overtaken))) stop = true;
            A += FI_CHUNK;
        }
        if (stop) {
            if (tid == 0 && s_minhit != 0x7fffffff) {
                int hr, hj;
                fi_locate(Abase + s_minhit, n, hr, hj);
                atomicMin(found, fi_key(hr, hj));
            }
            break;
        }
        __syncthreads();  // s_seg / s_minhit are rewritten by the next round
    }

    if (late_now) return;  // the apply launch reads ctl->fi_sel[parity] itself
    // ---- last block: exchange (several GPUs), then publish the winning move / close the sweep ---------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned tk = atomicAdd(&ctl->ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long f = *((volatile unsigned long long *)&ctl->fi_found);
    if (S.world > 1 && !S.finish_here) {  // NCCL min-allreduce of ctl->fi_found follows, then fi_finish_kernel
        if (tid == 0) ctl->ticket = 0;
        return;
    }
    if (shard) {
        __shared__ unsigned long long s_x[XCHG_MAX_WORLD];
        __shared__ int s_err;
        if (tid == 0) s_err = 0;
        __syncthreads();
        const unsigned long long none62 = (1ull << 62) - 1ull;
        const unsigned long long win = xchg_min(S.xchg, S.rank, S.world, f == FI_NONE ? none62 : f, s_x, &s_err, nullptr);
        if (tid == 0) {
            f = win == none62 ? FI_NONE : win;
            if (s_err) {
                ctl->error = 2;
                ctl->done = 1;
                ctl->done_reason = DONE_OPTIMUM;
                f = FI_NONE;
            }
        }
    }
    if (tid == 0) fi_finish(I, T, f, i0, j0);
}

// NCCL variant: after ncclAllReduce(min) of ctl->fi_found every rank finishes the launch with the same pair.
__global__ void fi_finish_kernel(const InstDev I, const TourDev T) {
    Ctl *ctl = T.ctl;
    if (ctl->done) { ctl->ap_valid = 0; return; }
    fi_finish(I, T, ctl->fi_found, ctl->cur_i, ctl->cur_j);
}

// end of a late-selection run: store the position entry the last apply launch parked
__global__ void fi_flush_kernel(const TourDev T) {
    Ctl *ctl = T.ctl;
    if (ctl->fi_pend_node >= 0) {
        T.pos[ctl->fi_pend_node] = ctl->fi_pend_pos;
        ctl->fi_pend_node = -1;
    }
}

cudaError_t launch_fi_flush(const TourDev &T, cudaStream_t st) {
    fi_flush_kernel<<<1, 1, 0, st>>>(T);
    return cudaGetLastError();
}

cudaError_t launch_fi_finish(const InstDev &I, const TourDev &T, cudaStream_t st) {
    fi_finish_kernel<<<1, 1, 0, st>>>(I, T);
    return cudaGetLastError();
}

cudaError_t launch_fi_search(const InstDev &I, const TourDev &T, int rank, int world, const XchgDev *xchg, int late, int grid, bool pdl,
                             cudaStream_t st) {
    const bool att = (I.metric == M_ATT);
    const bool ex = I.exact32 != 0;
    const dim3 g(grid), b(FI_THREADS);
    FiShard S{};
    S.rank = rank;
    S.world = world;
    S.finish_here = (world == 1 || xchg != nullptr) ? 1 : 0;
    S.late = late;
    if (xchg) S.xchg = *xchg;
    if (!I.fp32_ok) return launch_maybe_pdl(fi_search_kernel<false, false, false>, g, b, 0, st, pdl, I, T, S);
    if (att && ex) return launch_maybe_pdl(fi_search_kernel<true, true, true>, g, b, 0, st, pdl, I, T, S);
    if (att) return launch_maybe_pdl(fi_search_kernel<true, false, true>, g, b, 0, st, pdl, I, T, S);
    if (ex) return launch_maybe_pdl(fi_search_kernel<false, true, true>, g, b, 0, st, pdl, I, T, S);
    return launch_maybe_pdl(fi_search_kernel<false, false, true>, g, b, 0, st, pdl, I, T, S);
}

}  // namespace tspb
