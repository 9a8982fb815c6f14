// kernels_batch.cu — batched 2-opt: one thread block drives one tour to its 2-opt local optimum with the
// whole tour state in shared memory (28 bytes per node), no host round trips.  This is the path for
// independent tour batches (GA offspring repair — reference src/genetic.c:426-443 calls alg_2opt on a
// mutated offspring; multi-start; VNS src/vns.c:143 and tabu restarts) and for small TSPLIB instances,
// where a grid-wide launch per move would be pure latency.
//
// Both modes replay the reference bit for bit:
//   FI  reference src/heuristics.c:438-502: row-major (i<j) sweep, every improving move applied at once,
//       the scan continues at (i, j+1); sweeps repeat until one brings no gain.
//   BI  reference src/tabusearch.c:107-178 (NULL tabu list): full scan, strict '<' argmin == lowest (i,j)
//       among equal deltas, apply, repeat until no negative delta.
// Evaluation is in node space (the enumeration order IS the node order); FP32 filter + FP64 exact check
// as in the grid-wide kernels.
#include "tsp_state.cuh"

namespace tspb {

constexpr int BATCH_THREADS = 256;

struct BatchSmem {
    float *sx, *sy, *dsn;  // node space: coordinates, d(k, succ k)
    float *dsp;            // position space: d(order[p], order[p+1]) — what a reversal permutes without recomputing
    int *succ, *order, *pos;
};

template <bool ATT>
__device__ __forceinline__ float bt_dist32(float ax, float ay, float bx, float by) {
    float dx = ax - bx, dy = ay - by;
    float s = fmaf(dy, dy, dx * dx);
    if (ATT) s *= 0.1f;
    return sqrt_approx(s);
}

__device__ __forceinline__ long long bt_exact(const InstDev &I, const BatchSmem &S, bool exact32, int u, int v) {
    if (I.dmat) return (long long)I.dmat[(long long)u * I.dmat_ld + v];
    if (exact32)
        return exact_dist(I.metric, make_double2((double)S.sx[u], (double)S.sy[u]), make_double2((double)S.sx[v], (double)S.sy[v]));
    return exact_dist(I.metric, I.pt64[u], I.pt64[v]);
}

// exact delta of pair (i,j) or "skip" (adjacent) -> returns false
template <bool ATT, bool EXACT32, bool FP32_OK>
__device__ __forceinline__ bool bt_eval(const InstDev &I, const BatchSmem &S, int i, int j, float xi, float yi, int si,
                                        float xsi, float ysi, float dsi, float thr, long long &delta) {
    const int sj = S.succ[j];
    if (sj == i || si == j || si == sj) return false;  // reference heuristics.c:471 / tabusearch.c:134
    const float dsj = S.dsn[j];
    if (FP32_OK) {
        float q = bt_dist32<ATT>(xi, yi, S.sx[j], S.sy[j]) + bt_dist32<ATT>(xsi, ysi, S.sx[sj], S.sy[sj]) - dsi - dsj;
        if (q > thr) return false;
    }
    delta = bt_exact(I, S, FP32_OK && EXACT32, i, j) + bt_exact(I, S, FP32_OK && EXACT32, si, sj) - (long long)dsi - (long long)dsj;
    return true;
}

// block-wide application of move (i,j) on the shared-memory tour (same orientation rule as apply_swap_range): the forward
// path a1..b is reversed in place.  Edge lengths inside the path are only PERMUTED (position-space dsp[] reversed among
// themselves); just the two new edges (a,b) and (a1,b1) are evaluated.
__device__ __forceinline__ void bt_apply(const InstDev &I, const BatchSmem &S, bool exact32, int n, int i, int j) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int pa = S.pos[i], pb = S.pos[j];
    const int a1 = S.succ[i], b1 = S.succ[j];
    __syncthreads();
    int s = pa + 1; if (s >= n) s -= n;
    int len = pb - pa; if (len < 0) len += n;
    const int e = pb;
    const int half = len >> 1, mhalf = (len - 1) >> 1;
    for (int t = tid; t < half; t += nt) {
        int A = s + t; if (A >= n) A -= n;
        int B = e - t; if (B < 0) B += n;
        const int ua = S.order[A], ub = S.order[B];
        S.order[A] = ub; S.order[B] = ua;
        S.pos[ub] = A; S.pos[ua] = B;
        if (t < mhalf) {
            int Bm = B - 1; if (Bm < 0) Bm += n;
            const float za = S.dsp[A], zc = S.dsp[Bm];
            S.dsp[A] = zc; S.dsp[Bm] = za;
        }
    }
    if (tid == 0) {
        S.dsp[pa] = (float)bt_exact(I, S, exact32, i, j);
        S.dsp[pb] = (float)bt_exact(I, S, exact32, a1, b1);
    }
    __syncthreads();
    // nodes at positions pa .. pb got a new successor
    for (int t = tid; t <= len; t += nt) {
        int P = pa + t; if (P >= n) P -= n;
        int Pn = P + 1; if (Pn >= n) Pn -= n;
        const int k = S.order[P];
        S.succ[k] = S.order[Pn];
        S.dsn[k] = S.dsp[P];
    }
    __syncthreads();
}

template <bool ATT, bool EXACT32, bool FP32_OK>
__global__ void __launch_bounds__(BATCH_THREADS) two_opt_batch_kernel(const InstDev I, int mode, int *succ_all,
                                                                      long long *obj_out, long long *counters, int batch,
                                                                      MoveRec *log, long long log_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_hits[BATCH_THREADS / 32];
    __shared__ int s_err;
    __shared__ MoveKey s_keys[BATCH_THREADS / 32];
    const int n = I.n;
    const int tid = threadIdx.x;
    BatchSmem S;
    S.sx = reinterpret_cast<float *>(smem_raw);
    S.sy = S.sx + n;
    S.dsn = S.sy + n;
    S.dsp = S.dsn + n;
    S.succ = reinterpret_cast<int *>(S.dsp + n);
    S.order = S.succ + n;
    S.pos = S.order + n;
    const bool ex32 = FP32_OK && EXACT32;
    const float W = I.W;

    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        int *gsucc = succ_all + (long long)b * n;
        for (int k = tid; k < n; k += BATCH_THREADS) {
            float2 p = I.pt32[k];
            S.sx[k] = p.x; S.sy[k] = p.y;
            S.succ[k] = gsucc[k];
        }
        if (tid == 0) s_err = 0;
        __syncthreads();
        if (tid == 0) {  // visiting order from node 0; also validates the cycle
            int at = 0;
            for (int p = 0; p < n; ++p) {
                if (at < 0 || at >= n || (p > 0 && at == 0)) { s_err = 1; break; }
                S.order[p] = at;
                S.pos[at] = p;
                at = S.succ[at];
            }
            if (at != 0) s_err = 1;
        }
        __syncthreads();
        long long moves = 0, passes = 0, evals = 0, objd = 0;
        if (!s_err && n >= 4) {
            for (int k = tid; k < n; k += BATCH_THREADS) {
                const float d = (float)bt_exact(I, S, ex32, k, S.succ[k]);
                S.dsn[k] = d;
                S.dsp[S.pos[k]] = d;
            }
            __syncthreads();
            if (mode == 1) {
                // ---------------- best improvement ----------------
                for (;;) {
                    MoveKey best = key_none();
                    float thr = W;
                    const int warp = tid >> 5, lane = tid & 31;
                    for (int i = warp; i < n - 1; i += BATCH_THREADS / 32) {
                        const float xi = S.sx[i], yi = S.sy[i], dsi = S.dsn[i];
                        const int si = S.succ[i];
                        const float xsi = S.sx[si], ysi = S.sy[si];
                        for (int j = i + 1 + lane; j < n; j += 32) {
                            long long delta;
                            if (!bt_eval<ATT, EXACT32, FP32_OK>(I, S, i, j, xi, yi, si, xsi, ysi, dsi, thr, delta)) continue;
                            if (delta < 0 && delta <= (long long)best.delta) {
                                MoveKey k;
                                k.delta = (int)delta; k.i = i; k.j = j; k.pad = 0;
                                if (key_less(k, best)) { best = k; thr = (float)k.delta + W; }
                            }
                        }
                    }
                    best = key_warp_min(best);
                    if (lane == 0) s_keys[warp] = best;
                    __syncthreads();
                    best = s_keys[0];
#pragma unroll
                    for (int w = 1; w < BATCH_THREADS / 32; ++w)
                        if (key_less(s_keys[w], best)) best = s_keys[w];
                    __syncthreads();
                    passes++;
                    if (best.delta >= 0) break;
                    if (log && tid == 0 && moves < log_cap) {
                        MoveRec mr;
                        mr.i = best.i; mr.j = best.j; mr.delta = best.delta;
                        log[moves] = mr;
                    }
                    bt_apply(I, S, ex32, n, best.i, best.j);
                    moves++;
                    objd += best.delta;
                }
                evals = passes * ((long long)n * (n - 3) / 2);
            } else {
                // ---------------- first improvement ----------------
                // The reference's sweep is a linear scan of the pairs in row-major order from a cursor.  The block scans a
                // WINDOW of that linear order per step: chunks of 32 consecutive pairs (crossing row ends) are dealt to the
                // 8 warps round-robin, K chunks per warp, each warp stopping at its first improving pair; the smallest
                // linear offset over the warps is the reference's "first improving pair at or after the cursor".  K
                // doubles while nothing is found (long quiet stretches cost few barriers) and resets after a move (tours
                // far from the optimum improve every few hundred pairs), so the work past the hit is bounded by the gap.
                constexpr int NW = BATCH_THREADS / 32;
                const int warp = tid >> 5, lane = tid & 31;
                int ci = 0, cj = 1;      // cursor (row, column)
                long long lin_prev = 1;  // statistics: linear index i*n+j of the cursor after the last move / sweep start
                long long sweep_moves = 0;
                int K = 1;
                const float thr = -1.0f + W;
                // (row, column) of the pair `k` places after the cursor; row >= n-1 means "past the end of the sweep"
                auto locate = [&](int k, int &row, int &j) {
                    row = ci;
                    long long jj = (long long)cj + k;
                    while (row < n - 1 && jj >= n) {
                        jj -= n;
                        row += 1;
                        jj += row + 1;
                    }
                    j = (int)jj;
                };
                for (;;) {
                    int myhit = 0x7fffffff;
                    // this lane's pair of the warp's first chunk; later chunks are 32*NW pairs further along the linear
                    // order, reached by walking (row, column) forward instead of locating from the cursor again
                    int row, j;
                    locate(32 * warp + lane, row, j);
                    for (int q = 0; q < K; ++q) {
                        bool hit = false;
                        const bool valid = row < n - 1;
                        if (valid) {
                            const int si = S.succ[row];
                            long long delta;
                            if (bt_eval<ATT, EXACT32, FP32_OK>(I, S, row, j, S.sx[row], S.sy[row], si, S.sx[si], S.sy[si], S.dsn[row], thr, delta))
                                hit = delta < 0;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m) {
                            myhit = 32 * (warp + q * NW) + __ffs(m) - 1;
                            break;
                        }
                        if (!__any_sync(0xffffffffu, valid)) break;  // the whole chunk lies past the end of the sweep
                        long long jj = (long long)j + 32 * NW;
                        while (row < n - 1 && jj >= n) {
                            jj -= n;
                            row += 1;
                            jj += row + 1;
                        }
                        j = (int)jj;
                    }
                    if (lane == 0) s_hits[warp] = myhit;
                    __syncthreads();
                    int first = s_hits[0];
#pragma unroll
                    for (int w = 1; w < NW; ++w) first = min(first, s_hits[w]);
                    __syncthreads();  // s_hits[] is rewritten by the next step
                    bool sweep_end = false;
                    if (first != 0x7fffffff) {
                        int fi, fj;
                        locate(first, fi, fj);
                        const int si = S.succ[fi], sj = S.succ[fj];
                        const long long fdelta = bt_exact(I, S, ex32, fi, fj) + bt_exact(I, S, ex32, si, sj) - (long long)S.dsn[fi] -
                                                 (long long)S.dsn[fj];
                        evals += (long long)fi * n + fj - lin_prev + 1;  // linear pairs swept, like the grid kernel
                        if (log && tid == 0 && moves < log_cap) {
                            MoveRec mr;
                            mr.i = fi; mr.j = fj; mr.delta = fdelta;
                            log[moves] = mr;
                        }
                        bt_apply(I, S, ex32, n, fi, fj);
                        moves++; sweep_moves++;
                        objd += fdelta;
                        ci = fi; cj = fj + 1;
                        if (cj >= n) { ci = fi + 1; cj = ci + 1; }
                        lin_prev = (long long)ci * n + cj;
                        if (ci >= n - 1) sweep_end = true;
                        K = 1;
                    } else {
                        int nr, nj;
                        locate(32 * NW * K, nr, nj);
                        ci = nr; cj = nj;
                        if (ci >= n - 1) {
                            sweep_end = true;
                            evals += (long long)(n - 1) * n - lin_prev;
                        }
                        if (K < 16) K *= 2;
                    }
                    if (sweep_end) {
                        passes++;
                        if (sweep_moves == 0) break;  // reference heuristics.c:492
                        sweep_moves = 0;
                        ci = 0; cj = 1;
                        lin_prev = 1;
                        K = 1;
                    }
                }
            }
        } else if (!s_err) {
            passes = 1;
        }
        // write back
        long long cost_part = 0;
        for (int k = tid; k < n; k += BATCH_THREADS) {
            gsucc[k] = S.succ[k];
            if (!s_err && n >= 4) cost_part += (long long)S.dsn[k];
        }
        if (mode == 1) {
            // reference tabusearch.c:168-172: cost recomputed from scratch
            if (!s_err && n < 4)
                for (int k = tid; k < n; k += BATCH_THREADS) cost_part += bt_exact(I, S, ex32, k, S.succ[k]);
            for (int m = 16; m > 0; m >>= 1) cost_part += __shfl_xor_sync(0xffffffffu, cost_part, m);
            __shared__ long long s_cost[BATCH_THREADS / 32];
            if ((tid & 31) == 0) s_cost[tid >> 5] = cost_part;
            __syncthreads();
            if (tid == 0) {
                long long c = 0;
                for (int w = 0; w < BATCH_THREADS / 32; ++w) c += s_cost[w];
                obj_out[b] = c;
            }
        } else if (tid == 0) {
            obj_out[b] = objd;
        }
        if (tid == 0) {
            counters[4 * b + 0] = moves;
            counters[4 * b + 1] = passes;
            counters[4 * b + 2] = evals;
            counters[4 * b + 3] = s_err;
        }
        __syncthreads();
    }
}

template <bool ATT, bool EXACT32, bool FP32_OK>
static cudaError_t launch_batch_t(const InstDev &I, int mode, int *succ, long long *obj, long long *counters, int batch,
                                  int num_sms, cudaStream_t st, int *launched, MoveRec *log, long long log_cap) {
    size_t smem = (size_t)I.n * 28 + 16;
    auto kern = two_opt_batch_kernel<ATT, EXACT32, FP32_OK>;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BATCH_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidValue;
    int grid = num_sms * occ;
    if (grid > batch) grid = batch;
    kern<<<grid, BATCH_THREADS, smem, st>>>(I, mode, succ, obj, counters, batch, batch == 1 ? log : nullptr, log_cap);
    if (launched) *launched = 1;
    return cudaGetLastError();
}

cudaError_t launch_two_opt_batch(const InstDev &I, int mode, int *succ, long long *obj, long long *counters, int batch,
                                 int num_sms, cudaStream_t st, int *launched, MoveRec *log, long long log_cap) {
    const bool att = (I.metric == M_ATT);
    const bool ex = I.exact32 != 0;
    if (!I.fp32_ok) return launch_batch_t<false, false, false>(I, mode, succ, obj, counters, batch, num_sms, st, launched, log, log_cap);
    if (att && ex) return launch_batch_t<true, true, true>(I, mode, succ, obj, counters, batch, num_sms, st, launched, log, log_cap);
    if (att) return launch_batch_t<true, false, true>(I, mode, succ, obj, counters, batch, num_sms, st, launched, log, log_cap);
    if (ex) return launch_batch_t<false, true, true>(I, mode, succ, obj, counters, batch, num_sms, st, launched, log, log_cap);
    return launch_batch_t<false, false, true>(I, mode, succ, obj, counters, batch, num_sms, st, launched, log, log_cap);
}

}  // namespace tspb
