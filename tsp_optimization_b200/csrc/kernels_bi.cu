// kernels_bi.cu — best-improvement 2-opt pass (reference src/tabusearch.c:107-178 alg_2opt_tabu with a
// NULL tabu list): one launch = one full scan of all n(n-3)/2 non-adjacent pairs + argmin + apply.
//
// The scan runs in POSITION space.  For positions p<q with u=node(p), v=node(q) the reference pair is
// (i,j) = (min(u,v), max(u,v)); its two removed edges are (node(p),node(p+1)) and (node(q),node(q+1))
// whichever of u,v is the smaller index, and
//     delta = D[p][q] + D[p+1][q+1] - ds[p] - ds[q],      D[p][q] = d(node(p), node(q)).
// Each D value is therefore used by two pairs, (p,q) and (p-1,q-1): a thread owns R consecutive rows
// and marches along the columns, so it needs R+1 fresh distances per R evaluated moves.
//
// Arithmetic: FP32 (2 FADD, FMUL, FFMA, MUFU.SQRT per distance) as a FILTER: Q = D1 + D2 - ds_p is
// compared with thr + ds_q where thr = (best exact delta so far) + W.  W bounds the worst FP32/rounding
// excess, so the true argmin always passes; every pair that passes is re-evaluated in
// FP64 with the reference's exact operation order and only those exact integer deltas enter the
// (delta, i, j) argmin.  Column records are staged in shared memory by TMA bulk copies (UBLKCP),
// double-buffered across tiles.
#include "tsp_state.cuh"

namespace tspb {


constexpr int BI_THREADS = 256;

template <bool ATT>
__device__ __forceinline__ float dist32(float ax, float ay, float bx, float by) {
    float dx = ax - bx;
    float dy = ay - by;
    float s = fmaf(dy, dy, dx * dx);
    if (ATT) s *= 0.1f;
    return sqrt_approx(s);
}

// Exact re-evaluation of one filtered pair (p,q): returns the reference's integer delta, or LLONG_MAX for
// pairs the reference skips.  Not inlined: it is the cold path and everything goes in/out by value so that
// the caller's running best stays in registers.
#define BI_INVALID 0x7fffffffffffffffll
template <bool EXACT32>
__device__ __noinline__ long long bi_exact_delta(const InstDev I, const float4 *rec, int n, int p, int q, float xp,
                                                 float yp, float xp1, float yp1, float c0x, float c0y, float c0z,
                                                 float c0w, float c1x, float c1y, float c1w, float ds_p) {
    if (q < p + 2 || q > n - 1 || (p == 0 && q == n - 1)) return BI_INVALID;  // reference tabusearch.c:134
    long long d1, d2;
    if (EXACT32) {
        d1 = exact_dist(I.metric, make_double2((double)xp, (double)yp), make_double2((double)c0x, (double)c0y));
        d2 = exact_dist(I.metric, make_double2((double)xp1, (double)yp1), make_double2((double)c1x, (double)c1y));
    } else {
        int u = node_of(rec[p]);
        int u1 = node_of(rec[p + 1]);
        int v = __float_as_int(c0w);
        int v1 = __float_as_int(c1w);
        d1 = exact_dist(I.metric, I.pt64[u], I.pt64[v]);
        d2 = exact_dist(I.metric, I.pt64[u1], I.pt64[v1]);
    }
    return d1 + d2 - (long long)ds_p - (long long)c0z;
}

// Inline tail of the cold path: fold an exact delta into the thread's running (delta, i, j) minimum.
#define BI_CONSIDER(r_, q_)                                                                                       \
    do {                                                                                                          \
        long long dl_ = bi_exact_delta<EXACT32>(A.inst, rec, n, p0 + (r_), (q_), xr[r_], yr[r_], xr[(r_) + 1],     \
                                                yr[(r_) + 1], c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.w, -cp[r_]); \
        if (dl_ < 0 && dl_ <= (long long)best.delta) {                                                            \
            int u_ = node_of(rec[p0 + (r_)]);                                                                     \
            int v_ = __float_as_int(c0.w);                                                                        \
            MoveKey k_;                                                                                           \
            k_.delta = (int)dl_; k_.i = min(u_, v_); k_.j = max(u_, v_); k_.pad = 0;                              \
            if (key_less(k_, best)) {                                                                             \
                best = k_;                                                                                        \
                thr = fminf(thr, (float)k_.delta + W);                                                            \
                atomicMin(&s_hint, k_.delta);                                                                     \
            }                                                                                                     \
        }                                                                                                         \
    } while (0)

template <int R, bool ATT, bool EXACT32>
__global__ void __launch_bounds__(BI_THREADS, 2) bi_scan_kernel(const BiArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ int s_hint;
    __shared__ int s_last;
    __shared__ MoveKey s_keys[BI_THREADS / 32];

    Ctl *ctl = A.tour.ctl;
    if (ctl->done) return;

    const int tid = threadIdx.x;
    const int n = A.inst.n;
    const int TJ = A.TJ;
    const int TI = BI_THREADS * R;
    const float W = A.inst.W;
    const float4 *rec = A.tour.rec;
    float4 *scols0 = reinterpret_cast<float4 *>(smem_raw);
    float4 *scols1 = scols0 + (TJ + 2);
    const unsigned col_bytes = (unsigned)(TJ + 1) * 16u;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        s_hint = *((volatile int *)&ctl->hint);
    }
    __syncthreads();

    MoveKey best = key_none();
    float thr = (float)s_hint + W;

    const int first = A.rank + A.world * (int)blockIdx.x;
    const int stride = A.world * (int)gridDim.x;

    // tile id -> (tile row I, tile column J)
    auto decode = [&](int t, int &P0, int &Q0) {
        int lo = 0, hi = A.ntr - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (A.tile_row_start[mid] <= t) lo = mid; else hi = mid - 1;
        }
        P0 = lo * TI;
        Q0 = (A.tile_row_j0[lo] + (t - A.tile_row_start[lo])) * TJ;
    };

    int t = first;
    int P0 = 0, Q0 = 0;
    if (t < A.ntiles) {
        decode(t, P0, Q0);
        if (tid == 0) {
            mbar_expect_tx(&bars[0], col_bytes);
            tma_load_1d(scols0, rec + Q0, col_bytes, &bars[0]);
        }
    }

    for (int it = 0; t < A.ntiles; ++it) {
        const int buf = it & 1;
        const unsigned parity = (unsigned)(it >> 1) & 1u;
        float4 *sc = buf ? scols1 : scols0;
        // prefetch the next tile's columns into the other buffer
        const int tn = t + stride;
        int P0n = 0, Q0n = 0;
        if (tn < A.ntiles) {
            decode(tn, P0n, Q0n);
            if (tid == 0) {
                mbar_expect_tx(&bars[buf ^ 1], col_bytes);
                tma_load_1d(buf ? scols0 : scols1, rec + Q0n, col_bytes, &bars[buf ^ 1]);
            }
        }

        // rows of this thread: p0 .. p0+R-1 (+ successor row p0+R)
        const int p0 = P0 + tid * R;
        float xr[R + 1], yr[R + 1], cp[R];
#pragma unroll
        for (int r = 0; r <= R; ++r) {
            float4 v = rec[p0 + r];
            xr[r] = v.x;
            yr[r] = v.y;
            if (r < R) cp[r] = -v.z;  // padding rows carry ds = -BIG -> cp = +BIG -> never a candidate
        }
        thr = fminf(thr, (float)(*((volatile int *)&s_hint)) + W);

        mbar_wait(&bars[buf], parity);

        // pairs with q < p+2 exist in this tile?  (mask them; they are mirrored / adjacent pairs)
        const bool diag = (Q0 < P0 + TI + 1);
        float4 c0 = sc[0];
        float U[R];
#pragma unroll
        for (int r = 0; r < R; ++r) U[r] = dist32<ATT>(xr[r], yr[r], c0.x, c0.y) + cp[r];

        if (!diag) {
#pragma unroll 2
            for (int jj = 0; jj < TJ; ++jj) {
                const float4 c1 = sc[jj + 1];
                float Dn[R + 1];
#pragma unroll
                for (int r = 0; r <= R; ++r) Dn[r] = dist32<ATT>(xr[r], yr[r], c1.x, c1.y);
                float Q[R];
                float m = TSPB_BIG;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    Q[r] = U[r] + Dn[r + 1];
                    m = fminf(m, Q[r]);
                }
                const float Tq = thr + c0.z;
                if (m <= Tq) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (Q[r] <= Tq)
                            BI_CONSIDER(r, Q0 + jj);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) U[r] = Dn[r] + cp[r];
                c0 = c1;
                if ((jj & 63) == 63) thr = fminf(thr, (float)(*((volatile int *)&s_hint)) + W);
            }
        } else {
            const int qrel0 = Q0 - p0;  // q - p0 at jj = 0
#pragma unroll 2
            for (int jj = 0; jj < TJ; ++jj) {
                const float4 c1 = sc[jj + 1];
                float Dn[R + 1];
#pragma unroll
                for (int r = 0; r <= R; ++r) Dn[r] = dist32<ATT>(xr[r], yr[r], c1.x, c1.y);
                float Q[R];
                float m = TSPB_BIG;
                const int qrel = qrel0 + jj;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float qv = U[r] + Dn[r + 1];
                    Q[r] = (qrel >= r + 2) ? qv : TSPB_BIG;
                    m = fminf(m, Q[r]);
                }
                const float Tq = thr + c0.z;
                if (m <= Tq) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (Q[r] <= Tq)
                            BI_CONSIDER(r, Q0 + jj);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) U[r] = Dn[r] + cp[r];
                c0 = c1;
                if ((jj & 63) == 63) thr = fminf(thr, (float)(*((volatile int *)&s_hint)) + W);
            }
        }

        __syncthreads();  // every thread is done with sc[] before the next prefetch overwrites it
        if (tid == 0) {
            // exchange the best exact delta with the other blocks (only ever tightens the filter)
            int h = s_hint;
            int g = atomicMin(&ctl->hint, h);
            if (g < h) s_hint = g;
        }
        t = tn;
        P0 = P0n;
        Q0 = Q0n;
    }

    // ---- block argmin -> grid argmin ("last block done") -------------------------------------------
    best = key_warp_min(best);
    if ((tid & 31) == 0) s_keys[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        MoveKey k = (tid < BI_THREADS / 32) ? s_keys[tid] : key_none();
        k = key_warp_min(k);
        if (tid == 0) {
            A.tour.block_best[blockIdx.x] = k;
            __threadfence();
            unsigned tk = atomicAdd(&ctl->ticket, 1u);
            s_last = (tk == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    MoveKey k = key_none();
    for (int b = tid; b < (int)gridDim.x; b += BI_THREADS) {
        MoveKey o = key_load_cg(&A.tour.block_best[b]);
        if (key_less(o, k)) k = o;
    }
    k = key_warp_min(k);
    if ((tid & 31) == 0) s_keys[tid >> 5] = k;
    __syncthreads();
    k = s_keys[0];
#pragma unroll
    for (int w = 1; w < BI_THREADS / 32; ++w)
        if (key_less(s_keys[w], k)) k = s_keys[w];

    if (!A.fuse_apply) {
        if (tid == 0) {
            ctl->packed = (k.delta < 0) ? key_pack(k.delta, k.i, k.j) : key_pack(0, 0x1ffff, 0x1ffff);
            ctl->last = k;
            ctl->ticket = 0;
            ctl->hint = 0;
            ctl->launches += 1;
        }
        return;
    }

    if (k.delta < 0) apply_move_block(A.inst, A.tour, k.i, k.j);
    if (tid == 0) {
        ctl->passes += 1;
        ctl->launches += 1;
        ctl->last = k;
        if (k.delta < 0) {
            ctl->moves += 1;
            ctl->obj_delta += k.delta;
            long long lc = ctl->log_count;
            if (A.tour.log && lc < A.tour.log_cap) {
                MoveRec mr;
                mr.i = k.i; mr.j = k.j; mr.delta = k.delta;
                A.tour.log[lc] = mr;
            }
            ctl->log_count = lc + 1;
        } else {
            ctl->done = 1;  // reference src/tabusearch.c:158: mindelta >= 0 -> stop
        }
        ctl->ticket = 0;
        ctl->hint = 0;
    }
}

// Applies the globally reduced key after the NCCL min-allreduce (multi-GPU): every rank applies the
// same move to its replica of the tour, so no tour data ever crosses NVLink.
__global__ void __launch_bounds__(1024) bi_apply_packed_kernel(const InstDev inst, const TourDev tour) {
    Ctl *ctl = tour.ctl;
    if (ctl->done) return;
    int delta, i, j;
    key_unpack(ctl->packed, &delta, &i, &j);
    if (delta < 0) apply_move_block(inst, tour, i, j);
    if (threadIdx.x == 0) {
        ctl->passes += 1;
        if (delta < 0) {
            ctl->moves += 1;
            ctl->obj_delta += delta;
            long long lc = ctl->log_count;
            if (tour.log && lc < tour.log_cap) {
                MoveRec mr;
                mr.i = i; mr.j = j; mr.delta = delta;
                tour.log[lc] = mr;
            }
            ctl->log_count = lc + 1;
        } else {
            ctl->done = 1;
        }
    }
}

// ---- generic exact pass (any metric, incl. GEO / matrix lookup / oversized coordinates) ------------
// Same argmin, every delta evaluated exactly (FP64 on the fly or int32 matrix gather).  One block per
// group of rows; not the throughput path.
__global__ void __launch_bounds__(256) bi_scan_exact_kernel(const InstDev inst, const TourDev tour, int rank, int world,
                                                            int fuse_apply) {
    __shared__ MoveKey s_keys[8];
    __shared__ int s_last;
    Ctl *ctl = tour.ctl;
    if (ctl->done) return;
    const int n = tour.n;
    const int tid = threadIdx.x;
    const float4 *rec = tour.rec;
    MoveKey best = key_none();
    for (int p = rank + world * (int)blockIdx.x; p < n - 2; p += world * (int)gridDim.x) {
        const float4 rp = rec[p];
        const int u = node_of(rp), u1 = node_of(rec[p + 1]);
        const long long dsp = (long long)rp.z;
        for (int q = p + 2 + tid; q < n; q += 256) {
            if (p == 0 && q == n - 1) continue;
            const float4 rq = rec[q];
            const int v = node_of(rq), v1 = node_of(rec[q + 1]);
            long long delta = dist_nodes(inst, u, v) + dist_nodes(inst, u1, v1) - dsp - (long long)rq.z;
            if (delta < 0 && delta <= best.delta) {
                MoveKey k;
                k.delta = (int)delta; k.i = min(u, v); k.j = max(u, v); k.pad = 0;
                if (key_less(k, best)) best = k;
            }
        }
    }
    best = key_warp_min(best);
    if ((tid & 31) == 0) s_keys[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        MoveKey k = (tid < 8) ? s_keys[tid] : key_none();
        k = key_warp_min(k);
        if (tid == 0) {
            tour.block_best[blockIdx.x] = k;
            __threadfence();
            unsigned tk = atomicAdd(&ctl->ticket, 1u);
            s_last = (tk == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    MoveKey k = key_none();
    for (int b = tid; b < (int)gridDim.x; b += 256) {
        MoveKey o = key_load_cg(&tour.block_best[b]);
        if (key_less(o, k)) k = o;
    }
    k = key_warp_min(k);
    if ((tid & 31) == 0) s_keys[tid >> 5] = k;
    __syncthreads();
    k = s_keys[0];
    for (int w = 1; w < 8; ++w)
        if (key_less(s_keys[w], k)) k = s_keys[w];
    if (!fuse_apply) {
        if (tid == 0) {
            ctl->packed = (k.delta < 0) ? key_pack(k.delta, k.i, k.j) : key_pack(0, 0x1ffff, 0x1ffff);
            ctl->last = k;
            ctl->ticket = 0;
            ctl->launches += 1;
        }
        return;
    }
    if (k.delta < 0) apply_move_block(inst, tour, k.i, k.j);
    if (tid == 0) {
        ctl->passes += 1;
        ctl->launches += 1;
        ctl->last = k;
        if (k.delta < 0) {
            ctl->moves += 1;
            ctl->obj_delta += k.delta;
            long long lc = ctl->log_count;
            if (tour.log && lc < tour.log_cap) {
                MoveRec mr;
                mr.i = k.i; mr.j = k.j; mr.delta = k.delta;
                tour.log[lc] = mr;
            }
            ctl->log_count = lc + 1;
        } else {
            ctl->done = 1;
        }
        ctl->ticket = 0;
    }
}

// ---- host-side launchers -----------------------------------------------------------------------------
template <int R>
static cudaError_t launch_bi_r(const BiArgs &a, int grid, cudaStream_t st) {
    size_t smem = (size_t)2 * (a.TJ + 2) * sizeof(float4);
    const bool att = (a.inst.metric == M_ATT);
    const bool ex = a.inst.exact32 != 0;
    if (att && ex) bi_scan_kernel<R, true, true><<<grid, BI_THREADS, smem, st>>>(a);
    else if (att) bi_scan_kernel<R, true, false><<<grid, BI_THREADS, smem, st>>>(a);
    else if (ex) bi_scan_kernel<R, false, true><<<grid, BI_THREADS, smem, st>>>(a);
    else bi_scan_kernel<R, false, false><<<grid, BI_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bi_scan(const BiArgs &a, int rows_per_thread, int grid, cudaStream_t st) {
    if (rows_per_thread == 8) return launch_bi_r<8>(a, grid, st);
    if (rows_per_thread == 4) return launch_bi_r<4>(a, grid, st);
    return launch_bi_r<2>(a, grid, st);
}

cudaError_t launch_bi_scan_exact(const InstDev &inst, const TourDev &tour, int rank, int world, int fuse_apply,
                                 int grid, cudaStream_t st) {
    bi_scan_exact_kernel<<<grid, 256, 0, st>>>(inst, tour, rank, world, fuse_apply);
    return cudaGetLastError();
}

cudaError_t launch_bi_apply_packed(const InstDev &inst, const TourDev &tour, cudaStream_t st) {
    bi_apply_packed_kernel<<<1, 1024, 0, st>>>(inst, tour);
    return cudaGetLastError();
}

}  // namespace tspb
