// kernels_misc.cu — instance set-up, distance matrix, tour state build/export, tour costs, and the
// nearest-neighbour construction.
#include "tsp_state.cuh"

namespace tspb {

// ---- instance tables -----------------------------------------------------------------------------------
// pt64[k] = raw point, or for GEO the (lat, lon) radians of reference src/distutil.c:51-58; pt32 = FP32 copy.
__global__ void prep_points_kernel(const double2 *raw, double2 *pt64, float2 *pt32, int n, int metric) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double2 p = raw[k];
    pt32[k] = make_float2((float)p.x, (float)p.y);
    if (metric == M_GEO) p = make_double2(geo_radians(p.x), geo_radians(p.y));
    pt64[k] = p;
}

cudaError_t launch_prep_points(const double2 *raw, double2 *pt64, float2 *pt32, int n, int metric, cudaStream_t st) {
    prep_points_kernel<<<(n + 255) / 256, 256, 0, st>>>(raw, pt64, pt32, n, metric);
    return cudaGetLastError();
}

// ---- distance matrix --------------------------------------------------------------------------------------
// out[i*ld + j] = (int32) calc_dist(i, j)  (reference src/distutil.c:73-92), ld % 4 == 0 so that every row
// starts 16-byte aligned.  One thread = 4 consecutive j of MAT_ROWS rows -> one st.global.v4.s32 per row;
// a warp writes 512 contiguous bytes.  HBM-store-bound: 4*n*ld bytes written, O(n) read.
//
// Fast kernel (EUC_2D / CEIL_2D / ATT with FP32-exact coordinates): packed FP32x2 arithmetic, MUFU.SQRT, the
// integer taken straight from the bits of (r + 1.5*2^23).  A row whose 4 entries contain a value within the
// FP32 error band of a rounding boundary (|r_fp32 - r| <= r*2^-22: two roundings of s, MUFU.SQRT <= 2^-23) is
// only FLAGGED in the hot loop; flagged rows are re-evaluated exactly after the loop and stored again, so the
// hot loop has no divergent branch.  Exact re-evaluation: integer coordinates -> the rounding decision is an
// exact FP64 compare of s = dx^2+dy^2 against the boundary's square (no sqrt: for integer s the boundary can
// never be closer than 1/(8k+4) to sqrt(s), far outside double rounding); otherwise the
// reference's operation order in FP64 (exact_dist).
constexpr int MAT_THREADS = 256;
constexpr int MAT_ROWS = 32;
constexpr int MAT_ROWS_SLOW = 16;
enum { MK_NINT = 0, MK_CEIL = 1, MK_ATT = 2 };

// exact entry for integer coordinates; k0 = FP32 estimate rint(r), correct to +-1
template <int KIND>
__device__ __forceinline__ int exact_entry_int(float xi, float yi, float xj, float yj, int k0) {
    const double dx = (double)xi - (double)xj, dy = (double)yi - (double)yj;
    const double S = dx * dx + dy * dy;  // integers < 2^53: exact whatever the contraction
    const double k = (double)k0;
    if (KIND == MK_NINT) {  // floor(sqrt(S) + 0.5)
        if (S > k * k + k) return k0 + 1;
        if (k0 >= 1 && S <= k * k - k) return k0 - 1;
        return k0;
    } else if (KIND == MK_CEIL) {  // ceil(sqrt(S))
        if (S > k * k) return k0 + 1;
        if (k0 >= 1 && S <= (k - 1.0) * (k - 1.0)) return k0 - 1;
        return k0;
    } else {  // ATT: smallest t with 10 t^2 >= S  (== t = nint(r); t < r ? t+1 : t  for r = sqrt(S/10))
        if (S > 10.0 * k * k) return k0 + 1;
        if (k0 >= 1 && S <= 10.0 * (k - 1.0) * (k - 1.0)) return k0 - 1;
        return k0;
    }
}

template <int KIND>
__global__ void __launch_bounds__(MAT_THREADS) dist_matrix_kernel(const InstDev I, int *__restrict__ out, long long ld,
                                                                  int row_begin, int row_end) {
    const int n = I.n;
    const int j4 = (blockIdx.x * MAT_THREADS + threadIdx.x) * 4;
    if (j4 >= n) return;
    const int r0 = row_begin + blockIdx.y * MAT_ROWS;
    const int nrows = min(MAT_ROWS, row_end - r0);
    const float MAGIC = 12582912.0f;  // 1.5 * 2^23: (r + MAGIC) holds rint(r) in its low mantissa bits
    const int MAGIC_BITS = 0x4B400000;
    // FP32 error of r relative to the real distance: dx, dy exact for integer coordinates (<= 3 roundings + MUFU.SQRT's
    // 2^-23: below 2^-22); FP32-exact NON-integer coordinates add the roundings of dx and dy themselves (worst case 4 u =
    // 2^-22 exactly), so their band is doubled instead of leaving the bound without margin.  ATT has one more multiply.
    const float BAND = ((KIND == MK_ATT) ? 4.76837158203125e-07f : 2.384185791015625e-07f) * (I.int_coords ? 1.0f : 2.0f);
    float cx[4], cy[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float2 p = I.pt32[min(j4 + c, n - 1)];
        cx[c] = p.x;
        cy[c] = p.y;
    }
    const f32x2 cx01 = f2pack(cx[0], cx[1]), cx23 = f2pack(cx[2], cx[3]);
    const f32x2 cy01 = f2pack(cy[0], cy[1]), cy23 = f2pack(cy[2], cy[3]);
    const f32x2 magic2 = f2pack(MAGIC, MAGIC);
    const f32x2 tenth2 = f2pack(0.1f, 0.1f);
    const float2 *__restrict__ prow = I.pt32 + r0;
    int *__restrict__ orow = out + (long long)(r0 - row_begin) * ld + j4;
    unsigned flags = 0;

    auto row = [&](int rr) {
        const float2 pi = __ldg(prow + rr);
        const f32x2 px = f2pack(pi.x, pi.x), py = f2pack(pi.y, pi.y);
        const f32x2 dx01 = f2sub(px, cx01), dx23 = f2sub(px, cx23);
        const f32x2 dy01 = f2sub(py, cy01), dy23 = f2sub(py, cy23);
        f32x2 s01 = f2fma(dy01, dy01, f2mul(dx01, dx01));
        f32x2 s23 = f2fma(dy23, dy23, f2mul(dx23, dx23));
        if (KIND == MK_ATT) {
            s01 = f2mul(s01, tenth2);
            s23 = f2mul(s23, tenth2);
        }
        const float r[4] = {sqrt_approx(f2lo(s01)), sqrt_approx(f2hi(s01)), sqrt_approx(f2lo(s23)), sqrt_approx(f2hi(s23))};
        const f32x2 r01 = f2pack(r[0], r[1]), r23 = f2pack(r[2], r[3]);
        const f32x2 t01 = f2add(r01, magic2), t23 = f2add(r23, magic2);
        const f32x2 f01 = f2sub(r01, f2sub(t01, magic2)), f23 = f2sub(r23, f2sub(t23, magic2));
        const float t[4] = {f2lo(t01), f2hi(t01), f2lo(t23), f2hi(t23)};
        const float fr[4] = {f2lo(f01), f2hi(f01), f2lo(f23), f2hi(f23)};  // r - rint(r), in [-0.5, 0.5]
        int v[4];
        float u[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[c] = __float_as_int(t[c]) - MAGIC_BITS;
            if (KIND == MK_NINT) {
                u[c] = fmaf(r[c], BAND, fabsf(fr[c]));  // unsafe iff >= 0.5: within the band of k +- 0.5
            } else {
                u[c] = fmaf(r[c], -BAND, fabsf(fr[c]));  // unsafe iff <= 0: within the band of the integer k
                v[c] += (fr[c] > 0.f) ? 1 : 0;            // ceil
            }
        }
        bool flag;
        if (KIND == MK_NINT) flag = fmaxf(fmaxf(u[0], u[1]), fmaxf(u[2], u[3])) >= 0.5f;
        else flag = fminf(fminf(u[0], u[1]), fminf(u[2], u[3])) <= 0.f;
        if (flag) flags |= 1u << rr;
        *reinterpret_cast<int4 *>(orow + (long long)rr * ld) = make_int4(v[0], v[1], v[2], v[3]);
    };

    if (nrows == MAT_ROWS) {
#pragma unroll 4
        for (int rr = 0; rr < MAT_ROWS; ++rr) row(rr);
    } else {
        for (int rr = 0; rr < nrows; ++rr) row(rr);
    }

    // exact re-evaluation of the flagged rows (rare: ~1 % of a thread's rows at 10^4-range coordinates)
    while (flags) {
        const int rr = __ffs(flags) - 1;
        flags &= flags - 1;
        const int i = r0 + rr;
        const float2 pi = __ldg(prow + rr);
        int v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (I.int_coords) {
                const float dx = pi.x - cx[c], dy = pi.y - cy[c];
                float s = fmaf(dy, dy, dx * dx);
                if (KIND == MK_ATT) s *= 0.1f;
                const int k0 = __float_as_int(sqrt_approx(s) + MAGIC) - MAGIC_BITS;
                v[c] = exact_entry_int<KIND>(pi.x, pi.y, cx[c], cy[c], k0);
            } else {
                const double2 a = I.pt64[i], b = I.pt64[min(j4 + c, n - 1)];
                v[c] = (int)(KIND == MK_ATT ? exact_att(a.x, a.y, b.x, b.y)
                                            : (KIND == MK_CEIL ? exact_ceil(a.x, a.y, b.x, b.y) : exact_euc(a.x, a.y, b.x, b.y)));
            }
        }
        *reinterpret_cast<int4 *>(orow + (long long)rr * ld) = make_int4(v[0], v[1], v[2], v[3]);
    }
}

// Any metric / any coordinates: every entry in FP64 with the reference's operation order.
__global__ void __launch_bounds__(MAT_THREADS) dist_matrix_exact_kernel(const InstDev I, int *__restrict__ out, long long ld,
                                                                        int row_begin, int row_end, unsigned long long *geo_near) {
    const int n = I.n;
    const int j4 = (blockIdx.x * MAT_THREADS + threadIdx.x) * 4;
    if (j4 >= n) return;
    const int r0 = row_begin + blockIdx.y * MAT_ROWS_SLOW;
    const int r1 = min(r0 + MAT_ROWS_SLOW, row_end);
    const int metric = I.metric;
    double2 pj[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) pj[c] = I.pt64[min(j4 + c, n - 1)];
    for (int i = r0; i < r1; ++i) {
        const double2 pi = I.pt64[i];
        int v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = (int)exact_dist(metric, pi, pj[c]);
        *reinterpret_cast<int4 *>(out + (long long)(i - row_begin) * ld + j4) = make_int4(v[0], v[1], v[2], v[3]);
        if (metric == M_GEO && geo_near) {  // entries that sit on a rounding boundary (see geo_near_boundary)
            int cnt = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (j4 + c < n && geo_near_boundary(exact_geo_len(pi.x, pi.y, pj[c].x, pj[c].y))) cnt++;
            if (cnt) atomicAdd(geo_near, (unsigned long long)cnt);
        }
    }
}

cudaError_t launch_dist_matrix(const InstDev &I, int *out, long long ld, int row_begin, int row_end, bool fast,
                               unsigned long long *geo_near, cudaStream_t st) {
    int rows = row_end - row_begin;
    if (rows <= 0) return cudaSuccess;
    const unsigned gx = (unsigned)((I.n + MAT_THREADS * 4 - 1) / (MAT_THREADS * 4));
    if (fast) {
        dim3 grid(gx, (unsigned)((rows + MAT_ROWS - 1) / MAT_ROWS));
        if (I.metric == M_ATT) dist_matrix_kernel<MK_ATT><<<grid, MAT_THREADS, 0, st>>>(I, out, ld, row_begin, row_end);
        else if (I.metric == M_CEIL_2D) dist_matrix_kernel<MK_CEIL><<<grid, MAT_THREADS, 0, st>>>(I, out, ld, row_begin, row_end);
        else dist_matrix_kernel<MK_NINT><<<grid, MAT_THREADS, 0, st>>>(I, out, ld, row_begin, row_end);
    } else {
        dim3 grid(gx, (unsigned)((rows + MAT_ROWS_SLOW - 1) / MAT_ROWS_SLOW));
        dist_matrix_exact_kernel<<<grid, MAT_THREADS, 0, st>>>(I, out, ld, row_begin, row_end, geo_near);
    }
    return cudaGetLastError();
}

// ---- tour state ------------------------------------------------------------------------------------------
// order[p] = node at position p (host walks succ[] from node 0).  Builds both views + padding.
__global__ void build_state_kernel(const InstDev I, const TourDev T, const int *order) {
    const int n = T.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < T.alloc; p += gridDim.x * blockDim.x) {
        if (p < n) {
            int u = order[p];
            int v = order[p + 1 == n ? 0 : p + 1];
            float2 c = I.pt32[u];
            float ds = (float)dist_nodes(I, u, v);
            T.rec[p] = make_float4(c.x, c.y, ds, __int_as_float(u));
            T.pos[u] = p;
        } else if (p == n) {
            int u = order[0];
            float2 c = I.pt32[u];
            T.rec[p] = make_float4(c.x, c.y, -TSPB_BIG, __int_as_float(u));
        } else {
            T.rec[p] = make_float4(0.f, 0.f, -TSPB_BIG, __int_as_float(0));
        }
    }
}

// succ[node(p)] = node(p+1); cost = sum of ds (exact integers) accumulated in a 64-bit counter.
__global__ void export_state_kernel(const TourDev T, int *succ, unsigned long long *cost) {
    const int n = T.n;
    unsigned long long local = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        float4 r = T.rec[p];
        int pn = p + 1 == n ? 0 : p + 1;
        succ[node_of(r)] = node_of(T.rec[pn]);
        local += (unsigned long long)(long long)r.z;
    }
    for (int m = 16; m > 0; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(cost, local);
}

// Node-space tables (first-improvement search) built from the position-space records: on every first-improvement run
// that follows an upload or moves applied by anything else (best-improvement passes, kicks, a restore).
__global__ void rebuild_node_space_kernel(const TourDev T) {
    const int n = T.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float4 rp = T.rec[p];
        const float4 rn = T.rec[p + 1];  // rec[n] mirrors rec[0]
        const float4 rv = T.rec[p == 0 ? n - 1 : p - 1];
        const int k = node_of(rp);
        T.nrec[k] = make_float4(rp.x, rp.y, rn.x, rn.y);
        T.nlnk[k] = make_float4(rp.z, rn.w, rv.z, rv.w);
        T.npxy[k] = make_float2(rv.x, rv.y);
    }
}

// ---- successor array -> visiting order on the device (tour upload) ------------------------------------------------------
// The reference's tours are successor arrays (edges[k].j); the position-space state needs the visiting order from node 0.
// Walking succ[] on the host is n dependent cache misses (0.35 ms at n = 100 000, a third of the fixed cost of a
// host-buffer 2-opt call); here: list ranking by pointer jumping.  The cycle is cut in front of node 0 (tail = the node whose
// successor is 0); a[i] = {next, distance to next}; every round replaces next by next's next and adds the distances, so after
// ceil(log2 n) rounds a[i] = {tail, distance to the tail} and position(i) = n-1 - distance.  Validation = what the host walk
// checked: every successor in range and hit exactly once (a permutation) and every node reaches the tail (one cycle).
__global__ void __launch_bounds__(256) rank_init_kernel(const int *succ, int n, int2 *a, int *indeg, int *err) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int s = succ[i];
    if (s < 0 || s >= n) {
        atomicOr(err, 1);
        a[i] = make_int2(i, 0);
        return;
    }
    atomicAdd(&indeg[s], 1);
    a[i] = s == 0 ? make_int2(i, 0) : make_int2(s, 1);
}

__global__ void __launch_bounds__(256) rank_jump_kernel(const int2 *a, int2 *b, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int2 me = a[i];
    const int2 nx = a[me.x];
    b[i] = make_int2(nx.x, me.y + nx.y);
}

__global__ void __launch_bounds__(256) rank_finish_kernel(const int *succ, const int2 *a, const int *indeg, int n, int *order, int *err) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int2 me = a[i];
    // the tail is its own successor in the cut list; everybody must have arrived there, with a distance that is a position
    const bool at_tail = me.x >= 0 && me.x < n && succ[me.x] == 0;
    if (indeg[i] != 1 || !at_tail || me.y < 0 || me.y > n - 1) {
        atomicOr(err, 2);
        return;
    }
    order[n - 1 - me.y] = i;
}

// order[p] = node at position p of the tour succ[] started at node 0; *err != 0 afterwards: succ[] is not one Hamiltonian cycle.
// work: 2 n int2 + n int + 1 int (device).
cudaError_t launch_succ_to_order(const int *succ, int n, void *work, int *order, int *err, cudaStream_t st) {
    int2 *a = static_cast<int2 *>(work), *b = a + n;
    int *indeg = reinterpret_cast<int *>(b + n);
    cudaError_t e = cudaMemsetAsync(indeg, 0, sizeof(int) * (size_t)n, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(err, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int grid = (n + 255) / 256;
    rank_init_kernel<<<grid, 256, 0, st>>>(succ, n, a, indeg, err);
    for (long long span = 1; span < n; span <<= 1) {
        rank_jump_kernel<<<grid, 256, 0, st>>>(a, b, n);
        int2 *t = a; a = b; b = t;
    }
    rank_finish_kernel<<<grid, 256, 0, st>>>(succ, a, indeg, n, order, err);
    return cudaGetLastError();
}

cudaError_t launch_rebuild_node_space(const TourDev &T, cudaStream_t st) {
    int grid = (T.n + 255) / 256;
    if (grid > 1184) grid = 1184;
    rebuild_node_space_kernel<<<grid, 256, 0, st>>>(T);
    return cudaGetLastError();
}

cudaError_t launch_build_state(const InstDev &I, const TourDev &T, const int *order, cudaStream_t st) {
    int grid = (T.alloc + 255) / 256;
    if (grid > 1184) grid = 1184;
    build_state_kernel<<<grid, 256, 0, st>>>(I, T, order);
    return cudaGetLastError();
}

cudaError_t launch_export_state(const TourDev &T, int *succ, unsigned long long *cost, cudaStream_t st) {
    int grid = (T.n + 255) / 256;
    if (grid > 1184) grid = 1184;
    export_state_kernel<<<grid, 256, 0, st>>>(T, succ, cost);
    return cudaGetLastError();
}

// ---- batched tour cost (reference src/genetic.c:51-60 fitness; src/tabusearch.c:168-172) -----------------
// tours[b*n + k]: as_order != 0 -> visiting order (chromosome), else successor array. One block per tour.
__global__ void __launch_bounds__(256) tour_cost_kernel(const InstDev I, const int *tours, const int *slots, int as_order,
                                                        long long *out) {
    __shared__ long long s_part[8];
    const int n = I.n;
    const int *t = tours + (long long)(slots ? slots[blockIdx.x] : (int)blockIdx.x) * n;
    long long local = 0;
    for (int k = threadIdx.x; k < n; k += 256) {
        int u, v;
        if (as_order) { u = t[k]; v = t[k + 1 == n ? 0 : k + 1]; }
        else { u = k; v = t[k]; }
        local += dist_nodes(I, u, v);
    }
    for (int m = 16; m > 0; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < 8; ++w) s += s_part[w];
        out[blockIdx.x] = s;
    }
}

cudaError_t launch_tour_cost(const InstDev &I, const int *tours, const int *slots, int as_order, long long *out, int batch,
                             cudaStream_t st) {
    tour_cost_kernel<<<batch, 256, 0, st>>>(I, tours, slots, as_order, out);
    return cudaGetLastError();
}

// ---- nearest-neighbour construction (reference src/heuristics.c:18-78 greedy) --------------------------
// n-1 dependent steps, each an argmin over the unvisited nodes with strict '<' (lowest index wins ties).
// One cooperative launch of a persistent grid would need grid-wide syncs; instead a single CLUSTER-free
// design: one block of 1024 threads per step-chain is latency-bound, so we run the whole chain inside ONE
// kernel with a grid barrier built from a monotonically increasing counter (all blocks are co-resident:
// the launcher sizes the grid to at most one block per SM).
// Distances are exact FP64 (or matrix lookups): n^2 evaluations in total are negligible next to 2-opt.

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned *)counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) nn_tour_kernel(const NnArgs A) {
    __shared__ unsigned long long s_best[8];
    const InstDev &I = A.inst;
    const int n = I.n;
    const int gsz = gridDim.x * blockDim.x;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    int cur = A.start;
    long long total = 0;
    unsigned epoch = 0;
    if (gtid == 0) A.visited[cur] = 1;
    grid_barrier(A.barrier, (++epoch) * gridDim.x);
    // ONE grid barrier per step.  The winner of step s goes through slots[s % 3]; every thread reads it after the barrier
    // and treats `nxt` as visited from its own registers, so the visited[] flag written by thread 0 only has to be
    // visible one barrier later; slots[(s+2) % 3] (last read before this barrier, next used after the following one)
    // is re-armed by thread 0 in the meantime.
    for (int step = 0; step < n - 1; ++step) {
        unsigned long long *slot = &A.slots[step % 3];
        unsigned long long best = ~0ull;
        const double2 pc = I.dmat ? make_double2(0, 0) : I.pt64[cur];
        for (int k = gtid; k < n; k += gsz) {
            if (k == cur || ((volatile unsigned char *)A.visited)[k]) continue;
            long long d = I.dmat ? (long long)I.dmat[(long long)cur * I.dmat_ld + k] : exact_dist(I.metric, pc, I.pt64[k]);
            unsigned long long key = ((unsigned long long)d << 32) | (unsigned)k;
            best = key < best ? key : best;
        }
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, best, m);
            best = o < best ? o : best;
        }
        if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) best = s_best[w] < best ? s_best[w] : best;
            if (best != ~0ull) atomicMin(slot, best);
        }
        grid_barrier(A.barrier, (++epoch) * gridDim.x);
        const unsigned long long win = *((volatile unsigned long long *)slot);
        const int nxt = (int)(win & 0xffffffffu);
        total += (long long)(win >> 32);
        if (gtid == 0) {
            A.succ[cur] = nxt;
            A.visited[nxt] = 1;
            A.slots[(step + 2) % 3] = ~0ull;
        }
        cur = nxt;
    }
    if (gtid == 0) {
        A.succ[cur] = A.start;  // closing edge, reference heuristics.c:59-62,74
        total += dist_nodes(I, cur, A.start);
        *A.cost = total;
    }
}

// The grid barrier needs every block resident at once: cooperative launch, grid <= SMs * occupancy.
int nn_max_grid(int num_sms) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nn_tour_kernel, 256, 0) != cudaSuccess || occ < 1) return num_sms;
    return num_sms;  // one block per SM keeps the barrier cheap
}

cudaError_t launch_nn_tour(const NnArgs &a, int grid, cudaStream_t st) {
    NnArgs args = a;
    void *params[] = {&args};
    return cudaLaunchCooperativeKernel((const void *)nn_tour_kernel, dim3(grid), dim3(256), params, 0, st);
}

// ---- batched nearest neighbour: one block per start node (reference src/heuristics.c:168-205 HEU_Greedy_iter = n
// independent greedy() runs; also multi-start seeds for VNS / tabu / GA) -----------------------------------------------
// Per step the block finds argmin over the unvisited nodes of the exact integer distance, lowest index among equals
// (strict '<' scan of reference heuristics.c:44-55).  FP32 is only a filter: pass 1 takes the block minimum m of the FP32
// distances; an exact integer distance differs from the real one by less than 1, so only nodes with an FP32 distance
// <= m + 1 + 2*eps can hold the minimum — pass 2 evaluates those exactly (FP64 / matrix) and reduces (distance, index).
constexpr int NNB_THREADS = 256;

__global__ void __launch_bounds__(NNB_THREADS) nn_batch_kernel(const InstDev I, const int *starts, int batch, int *succ_out,
                                                               long long *cost_out, float eps) {
    extern __shared__ __align__(16) unsigned char nnb_smem[];
    __shared__ float s_min[NNB_THREADS / 32];
    __shared__ unsigned long long s_key[NNB_THREADS / 32];
    const int n = I.n;
    float2 *sxy = reinterpret_cast<float2 *>(nnb_smem);
    unsigned char *vis = reinterpret_cast<unsigned char *>(sxy + n);
    const int tid = threadIdx.x;
    const bool filter = I.fp32_ok != 0 && I.dmat == nullptr;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const int start = starts[b];
        for (int k = tid; k < n; k += NNB_THREADS) {
            sxy[k] = I.pt32[k];
            vis[k] = (k == start) ? 1 : 0;
        }
        __syncthreads();
        int cur = start;
        long long total = 0;
        int *succ = succ_out ? succ_out + (long long)b * n : nullptr;
        for (int step = 0; step < n - 1; ++step) {
            float thr = 3.0e38f;
            if (filter) {
                const float2 pc = sxy[cur];
                float m = 3.0e38f;
                for (int k = tid; k < n; k += NNB_THREADS) {
                    if (vis[k]) continue;
                    const float dx = pc.x - sxy[k].x, dy = pc.y - sxy[k].y;
                    float s2 = fmaf(dy, dy, dx * dx);
                    if (I.metric == M_ATT) s2 *= 0.1f;
                    m = fminf(m, sqrt_approx(s2));
                }
                for (int w = 16; w > 0; w >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, w));
                if ((tid & 31) == 0) s_min[tid >> 5] = m;
                __syncthreads();
                m = s_min[0];
#pragma unroll
                for (int w = 1; w < NNB_THREADS / 32; ++w) m = fminf(m, s_min[w]);
                thr = m + 1.0f + 2.0f * eps;
            }
            unsigned long long best = ~0ull;
            {
                const float2 pc = sxy[cur];
                for (int k = tid; k < n; k += NNB_THREADS) {
                    if (vis[k]) continue;
                    if (filter) {
                        const float dx = pc.x - sxy[k].x, dy = pc.y - sxy[k].y;
                        float s2 = fmaf(dy, dy, dx * dx);
                        if (I.metric == M_ATT) s2 *= 0.1f;
                        if (sqrt_approx(s2) > thr) continue;
                    }
                    const unsigned long long key = ((unsigned long long)dist_nodes(I, cur, k) << 32) | (unsigned)k;
                    best = key < best ? key : best;
                }
            }
            for (int w = 16; w > 0; w >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, w);
                best = o < best ? o : best;
            }
            if ((tid & 31) == 0) s_key[tid >> 5] = best;
            __syncthreads();
            best = s_key[0];
#pragma unroll
            for (int w = 1; w < NNB_THREADS / 32; ++w) best = s_key[w] < best ? s_key[w] : best;
            const int nxt = (int)(best & 0xffffffffu);
            total += (long long)(best >> 32);
            if (tid == 0) {
                if (succ) succ[cur] = nxt;
                vis[nxt] = 1;
            }
            cur = nxt;
            __syncthreads();
        }
        if (tid == 0) {
            if (succ) succ[cur] = start;  // closing edge, reference heuristics.c:59-62,74
            cost_out[b] = total + dist_nodes(I, cur, start);
        }
        __syncthreads();
    }
}

cudaError_t launch_nn_batch(const InstDev &I, const int *starts, int batch, int *succ_out, long long *cost_out, float eps,
                            int num_sms, cudaStream_t st) {
    const size_t smem = (size_t)I.n * (sizeof(float2) + 1) + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(nn_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nn_batch_kernel, NNB_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidValue;
    int grid = num_sms * occ;
    if (grid > batch) grid = batch;
    nn_batch_kernel<<<grid, NNB_THREADS, smem, st>>>(I, starts, batch, succ_out, cost_out, eps);
    return cudaGetLastError();
}

}  // namespace tspb
