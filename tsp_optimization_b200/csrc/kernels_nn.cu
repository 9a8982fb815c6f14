// kernels_nn.cu — nearest-neighbour construction (reference src/heuristics.c:18-78 greedy) on a uniform cell grid.
//
// The reference scans all unvisited nodes per step with a strict '<' on the integer distance: the next node is the
// unvisited node of minimum integer distance, lowest index among equals.  For the planar metrics (EUC_2D, CEIL_2D, ATT)
// the integer distance is a non-decreasing function of the Euclidean one, so that node lies within a known radius of the
// current one and only the cells of a bucket grid around it have to be looked at:
//   * a node at real distance r has integer distance >= r - 1/2 (nint), >= r (ceil), >= r / sqrt(10) (ATT);
//   * after the (2 rho + 1)^2 window of cells around the current node's cell has been searched, every unsearched node
//     is farther than rho * h (h = cell size); so once the best integer distance found, dmin, satisfies
//     (dmin + 1) * scale <= rho * h  (scale = 1, or sqrt(10) for ATT), no unsearched node can reach dmin — not even tie it —
//     and the window's (distance, index) minimum is exactly the reference's choice.
// All distances that decide anything are exact_dist() values (FP64, reference operation order); the grid only limits
// WHICH nodes are evaluated.  n-1 strictly dependent steps: ONE warp walks the tour (cell table and the "unvisited"
// bitmask live in shared memory, the cell-sorted coordinates are read through L1 — consecutive steps look at nearly the
// same cells), ~0.3 us per step instead of the 3.6 us of a grid-wide scan + grid barrier (nn_tour_kernel, still used for
// GEO / MAN / MAX / matrix instances).  When the rings up to NN_RHO_MAX hold no acceptable node — the walker has eaten its
// neighbourhood empty, ~0.1 % of the steps of a uniform instance — the other 15 warps of the block, parked on a named
// barrier until then, join for one scan of everything still unvisited.
#include "tsp_state.cuh"

namespace tspb {

constexpr int NN_THREADS = 512;   // 128 registers per thread: the walker's loop stays spill-free
constexpr int NN_WARPS = NN_THREADS / 32;
constexpr int NN_RHO_MAX = 8;
constexpr int NN_MAX_RANGES = 2 + 2 * (2 * NN_RHO_MAX - 1);  // ranges of sorted positions in the outermost ring

struct NnGridArgs {
    InstDev inst;
    int start;
    int GX, GY;                 // cells per axis
    double xmin, ymin, h, inv_h;
    double scale;               // real distance per unit of integer distance: 1 (EUC_2D, CEIL_2D), sqrt(10) (ATT), rounded up
    int *cell_cnt;              // [GX*GY + 1] populations, then exclusive prefix sums (cell_start)
    int *cell_fill;             // [GX*GY]     scatter cursors
    double2 *spt;               // [n] points sorted by cell
    int *snode;                 // [n] node ids in the same order
    int *start_spos;            // sorted position of the start node
    int *succ;                  // out
    long long *cost;            // out
};

__device__ __forceinline__ int nn_cell_coord(double v, double vmin, double inv_h, int G) {
    const int c = __double2int_rd(__dmul_rn(__dsub_rn(v, vmin), inv_h));
    return min(max(c, 0), G - 1);
}

__global__ void __launch_bounds__(256) nn_grid_count_kernel(const NnGridArgs A) {
    const int n = A.inst.n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double2 p = A.inst.pt64[k];
        const int c = nn_cell_coord(p.y, A.ymin, A.inv_h, A.GY) * A.GX + nn_cell_coord(p.x, A.xmin, A.inv_h, A.GX);
        atomicAdd(&A.cell_cnt[c], 1);
    }
}

// exclusive prefix sums over the cells, one block: cell_cnt[c] <- number of nodes in cells < c, cell_cnt[ncell] <- n
__global__ void __launch_bounds__(1024) nn_grid_scan_kernel(const NnGridArgs A) {
    __shared__ int s_part[1024];
    const int ncell = A.GX * A.GY;
    const int tid = threadIdx.x;
    const int per = (ncell + 1023) / 1024;
    const int lo = min(tid * per, ncell), hi = min(lo + per, ncell);
    int sum = 0;
    for (int c = lo; c < hi; ++c) sum += A.cell_cnt[c];
    s_part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const int v = tid >= off ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - sum;
    for (int c = lo; c < hi; ++c) {
        const int v = A.cell_cnt[c];
        A.cell_cnt[c] = run;
        run += v;
    }
    if (tid == 1023) A.cell_cnt[ncell] = s_part[1023];
}

__global__ void __launch_bounds__(256) nn_grid_scatter_kernel(const NnGridArgs A) {
    const int n = A.inst.n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double2 p = A.inst.pt64[k];
        const int c = nn_cell_coord(p.y, A.ymin, A.inv_h, A.GY) * A.GX + nn_cell_coord(p.x, A.xmin, A.inv_h, A.GX);
        const int pos = A.cell_cnt[c] + atomicAdd(&A.cell_fill[c], 1);  // order inside a cell is irrelevant: minima are by (distance, node)
        A.spt[pos] = p;
        A.snode[pos] = k;
        if (k == A.start) *A.start_spos = pos;
    }
}

__device__ __forceinline__ void nn_bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(NN_THREADS) : "memory"); }

struct NnBest {
    unsigned long long key;  // (integer distance << 32) | node id: the reference's (distance, lowest index) order
    int spos;
    double2 p;
};

__device__ __forceinline__ void nn_consider(const NnGridArgs &A, const double2 cur, int spos, NnBest &b) {
    const double2 p = __ldg(&A.spt[spos]);
    const int node = __ldg(&A.snode[spos]);
    const unsigned long long key = ((unsigned long long)exact_dist(A.inst.metric, cur, p) << 32) | (unsigned)node;
    if (key < b.key) { b.key = key; b.spos = spos; b.p = p; }
}

// warp-wide minimum by key; every lane returns the winner.  Two REDUX.MIN (distance, then node id among the lanes that hold
// that distance) instead of five 64-bit shuffle rounds; the winner's coordinates are re-read (L1) rather than shuffled.
__device__ __forceinline__ NnBest nn_warp_min(const NnGridArgs &A, NnBest b) {
    const unsigned d = (unsigned)(b.key >> 32), node = (unsigned)b.key;
    const unsigned dmin = __reduce_min_sync(0xffffffffu, d);
    const unsigned nmin = __reduce_min_sync(0xffffffffu, d == dmin ? node : 0xffffffffu);
    NnBest r;
    r.key = ((unsigned long long)dmin << 32) | nmin;
    if (r.key == ~0ull) { r.spos = -1; r.p = make_double2(0.0, 0.0); return r; }
    const unsigned who = __ballot_sync(0xffffffffu, b.key == r.key);
    r.spos = __shfl_sync(0xffffffffu, b.spos, __ffs(who) - 1);
    r.p = __ldg(&A.spt[r.spos]);
    return r;
}

// One scan of every unvisited node by the whole block (all NN_THREADS threads call it).  Pass 1 finds an unvisited node of
// (nearly) minimum squared distance; its exact integer distance dub bounds the minimum from above, so the winner — and every
// node that could tie it — has a real distance <= (dub + 1) * scale; pass 2 evaluates exactly those and reduces (distance, node).
__device__ void nn_block_scan(const NnGridArgs &A, const unsigned *alive, int nwords, double cx, double cy, double *s_red,
                              int *s_redi, unsigned long long *s_redk, unsigned long long *out_key, int *out_spos) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 cur = make_double2(cx, cy);
    double smin = 1e300;
    int smin_pos = -1;
    for (int w = tid; w < nwords; w += NN_THREADS) {
        unsigned bits = alive[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const double2 p = __ldg(&A.spt[w * 32 + b]);
            const double dx = p.x - cx, dy = p.y - cy;
            const double s = dx * dx + dy * dy;
            if (s < smin) { smin = s; smin_pos = w * 32 + b; }
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, smin, m);
        const int op = __shfl_xor_sync(0xffffffffu, smin_pos, m);
        if (os < smin) { smin = os; smin_pos = op; }
    }
    if (lane == 0) { s_red[warp] = smin; s_redi[warp] = smin_pos; }
    __syncthreads();
    if (warp == 0) {
        smin = lane < NN_WARPS ? s_red[lane] : 1e300;
        smin_pos = lane < NN_WARPS ? s_redi[lane] : -1;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, smin, m);
            const int op = __shfl_xor_sync(0xffffffffu, smin_pos, m);
            if (os < smin) { smin = os; smin_pos = op; }
        }
        if (lane == 0) {
            double bound = 1e300;
            if (smin_pos >= 0) {
                const double r = ((double)exact_dist(A.inst.metric, cur, __ldg(&A.spt[smin_pos])) + 1.0) * A.scale;
                bound = r * r * (1.0 + 1e-9);
            }
            s_red[0] = bound;
        }
    }
    __syncthreads();
    const double bound = s_red[0];
    __syncthreads();  // s_red is reused below
    NnBest best;
    best.key = ~0ull; best.spos = -1; best.p = make_double2(0.0, 0.0);
    for (int w = tid; w < nwords; w += NN_THREADS) {
        unsigned bits = alive[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const double2 p = __ldg(&A.spt[w * 32 + b]);
            const double dx = p.x - cx, dy = p.y - cy;
            if (dx * dx + dy * dy <= bound) nn_consider(A, cur, w * 32 + b, best);
        }
    }
    best = nn_warp_min(A, best);
    if (lane == 0) { s_redk[warp] = best.key; s_redi[warp] = best.spos; }
    __syncthreads();
    if (warp == 0) {
        unsigned long long k = lane < NN_WARPS ? s_redk[lane] : ~0ull;
        int sp = lane < NN_WARPS ? s_redi[lane] : -1;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, m);
            const int op = __shfl_xor_sync(0xffffffffu, sp, m);
            if (ok < k) { k = ok; sp = op; }
        }
        if (lane == 0) { *out_key = k; *out_spos = sp; }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NN_THREADS, 1) nn_grid_walk_kernel(const NnGridArgs A) {
    extern __shared__ __align__(16) unsigned char nn_smem[];
    __shared__ double s_red[32];
    __shared__ int s_redi[32];
    __shared__ unsigned long long s_redk[32];
    __shared__ double s_cx, s_cy;
    __shared__ unsigned long long s_key;
    __shared__ int s_spos, s_cmd;
    __shared__ int s_rlo[NN_MAX_RANGES + 1], s_rpre[NN_MAX_RANGES + 1];
    const int n = A.inst.n;
    const int GX = A.GX, GY = A.GY, ncell = GX * GY;
    const int nwords = (n + 31) >> 5;
    int *cs = reinterpret_cast<int *>(nn_smem);
    unsigned *alive = reinterpret_cast<unsigned *>(cs + ncell + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i <= ncell; i += NN_THREADS) cs[i] = A.cell_cnt[i];
    for (int w = tid; w < nwords; w += NN_THREADS) {
        const int left = n - w * 32;
        alive[w] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
    }
    if (tid == 0) s_cmd = 0;
    __syncthreads();

    if (warp != 0) {
        // helpers: parked on barrier 1 until the walker needs a scan of everything unvisited (s_cmd = 1) or is done (2)
        for (;;) {
            nn_bar_all();
            if (s_cmd == 2) return;
            nn_block_scan(A, alive, nwords, s_cx, s_cy, s_red, s_redi, s_redk, &s_key, &s_spos);
        }
    }

    // ---- the walker --------------------------------------------------------------------------------------------------
    int cur_spos = *A.start_spos;
    int cur_node = A.start;
    double2 cur = A.spt[cur_spos];
    if (lane == 0) alive[cur_spos >> 5] &= ~(1u << (cur_spos & 31));
    __syncwarp();
    long long total = 0;
    const double h_safe = A.h * (1.0 - 1e-9);  // cell membership was decided in floating point: keep a margin
    for (int step = 0; step < n - 1; ++step) {
        const int cx = nn_cell_coord(cur.x, A.xmin, A.inv_h, GX), cy = nn_cell_coord(cur.y, A.ymin, A.inv_h, GY);
        NnBest best;
        best.key = ~0ull; best.spos = -1; best.p = make_double2(0.0, 0.0);
        bool accepted = false;
        for (int rho = 1; rho <= NN_RHO_MAX; ++rho) {
            const int x0 = max(cx - rho, 0), x1 = min(cx + rho, GX - 1);
            const int y0 = max(cy - rho, 0), y1 = min(cy + rho, GY - 1);
            NnBest mine;
            mine.key = ~0ull; mine.spos = -1; mine.p = make_double2(0.0, 0.0);
            if (rho == 1) {
                // the 3 x 3 window: up to three runs of consecutive sorted positions (cells of one grid row are adjacent)
                int lo[3], len[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int y = cy - 1 + r;
                    const bool in = y >= 0 && y < GY;
                    lo[r] = in ? cs[y * GX + x0] : 0;
                    len[r] = in ? cs[y * GX + x1 + 1] - lo[r] : 0;
                }
                const int t01 = len[0] + len[1], tot = t01 + len[2];
                for (int k = lane; k < tot; k += 32) {
                    const int spos = k < len[0] ? lo[0] + k : (k < t01 ? lo[1] + (k - len[0]) : lo[2] + (k - t01));
                    if ((alive[spos >> 5] >> (spos & 31)) & 1u) nn_consider(A, cur, spos, mine);
                }
            } else {
                // ring rho: the top and bottom rows over the full width, the left and right cells of the rows between
                const int nmid = min(cy + rho - 1, GY - 1) - max(cy - rho + 1, 0) + 1;
                const int nr = 2 + 2 * nmid;
                int rlo = 0, rlen = 0;
                if (lane < nr) {
                    if (lane < 2) {
                        const int y = lane == 0 ? cy - rho : cy + rho;
                        if (y >= 0 && y < GY) { rlo = cs[y * GX + x0]; rlen = cs[y * GX + x1 + 1] - rlo; }
                    } else {
                        const int y = max(cy - rho + 1, 0) + ((lane - 2) >> 1);
                        const int x = (lane & 1) ? cx + rho : cx - rho;
                        if (x >= 0 && x < GX) { rlo = cs[y * GX + x]; rlen = cs[y * GX + x + 1] - rlo; }
                    }
                }
                int pre = rlen;  // inclusive prefix sums of the run lengths
#pragma unroll
                for (int m = 1; m < 32; m <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, pre, m);
                    if (lane >= m) pre += o;
                }
                const int tot = __shfl_sync(0xffffffffu, pre, 31);
                if (lane < nr) { s_rlo[lane] = rlo; s_rpre[lane] = pre - rlen; }
                __syncwarp();
                for (int k = lane; k < tot; k += 32) {
                    int a = 0, b = nr - 1;  // last run whose exclusive prefix is <= k (empty runs share a prefix: take the last)
                    while (a < b) {
                        const int mid = (a + b + 1) >> 1;
                        if (s_rpre[mid] <= k) a = mid; else b = mid - 1;
                    }
                    const int spos = s_rlo[a] + (k - s_rpre[a]);
                    if ((alive[spos >> 5] >> (spos & 31)) & 1u) nn_consider(A, cur, spos, mine);
                }
                __syncwarp();
            }
            mine = nn_warp_min(A, mine);
            if (mine.key < best.key) best = mine;
            if (best.key != ~0ull) {
                const bool covers_all = cx - rho <= 0 && cx + rho >= GX - 1 && cy - rho <= 0 && cy + rho >= GY - 1;
                const double need = ((double)(long long)(best.key >> 32) + 1.0) * A.scale;
                if (covers_all || need <= (double)rho * h_safe) { accepted = true; break; }
            }
        }
        if (!accepted) {
            if (lane == 0) { s_cx = cur.x; s_cy = cur.y; s_cmd = 1; }
            nn_bar_all();
            nn_block_scan(A, alive, nwords, cur.x, cur.y, s_red, s_redi, s_redk, &s_key, &s_spos);
            best.key = s_key;
            best.spos = s_spos;
            best.p = __ldg(&A.spt[best.spos]);
        }
        const int nxt = (int)(best.key & 0xffffffffu);
        if (lane == 0) {
            A.succ[cur_node] = nxt;
            alive[best.spos >> 5] &= ~(1u << (best.spos & 31));
        }
        __syncwarp();
        total += (long long)(best.key >> 32);
        cur_node = nxt;
        cur_spos = best.spos;
        cur = best.p;
    }
    if (lane == 0) {
        A.succ[cur_node] = A.start;  // closing edge, reference heuristics.c:59-62,74
        total += exact_dist(A.inst.metric, cur, A.inst.pt64[A.start]);
        *A.cost = total;
        s_cmd = 2;
    }
    nn_bar_all();
}

// shared memory the walk needs for (n nodes, ncell cells)
size_t nn_grid_smem_bytes(int n, int ncell) { return sizeof(int) * ((size_t)ncell + 1) + sizeof(unsigned) * (((size_t)n + 31) / 32) + 16; }

cudaError_t launch_nn_grid(const NnGridArgs &a, int num_sms, cudaStream_t st) {
    const int n = a.inst.n, ncell = a.GX * a.GY;
    cudaError_t e = cudaMemsetAsync(a.cell_cnt, 0, sizeof(int) * ((size_t)ncell + 1), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.cell_fill, 0, sizeof(int) * (size_t)ncell, st);
    if (e != cudaSuccess) return e;
    int grid = (n + 255) / 256;
    if (grid > 4 * num_sms) grid = 4 * num_sms;
    nn_grid_count_kernel<<<grid, 256, 0, st>>>(a);
    nn_grid_scan_kernel<<<1, 1024, 0, st>>>(a);
    nn_grid_scatter_kernel<<<grid, 256, 0, st>>>(a);
    const size_t smem = nn_grid_smem_bytes(n, ncell);
    e = cudaFuncSetAttribute(nn_grid_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nn_grid_walk_kernel<<<1, NN_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tspb
