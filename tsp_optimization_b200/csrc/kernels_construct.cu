// kernels_construct.cu — extra-mileage (cheapest insertion from the farthest pair) construction, reference
// src/heuristics.c:208-314 HEU_extramileage: start from the two farthest nodes, then n-2 times insert the unvisited node c
// into the tour edge (a,b) that minimises C_ac + C_cb - C_ab, scanning nodes in index order (outer loop) and tour edges in
// the order of the reference's edges_visited[] array (inner loop) with a strict '<' — i.e. the winner is the
// lexicographic minimum of (delta, node, edge slot).  The replaced edge keeps its slot for (a,c); (c,b) is appended.
//
// One thread block, whole state in shared memory.  Instead of the reference's O(n^3) rescans every unvisited node caches
// its best (delta, slot): an insertion only changes slot j* and appends one slot, so a node whose cached slot is not j*
// just compares against those two; the others are rescanned by one warp each.  Same result, O(n^2) distance evaluations
// in the common case.  All distances are exact (FP64 in the reference's operation order, or the resident matrix).
#include "tsp_state.cuh"

namespace tspb {

constexpr int EM_THREADS = 1024;
constexpr unsigned long long EM_NONE = ~0ull;

// cache key: (delta + 2^31) in the high word, slot in the low word -> unsigned min == (lowest delta, lowest slot)
__device__ __forceinline__ unsigned long long em_key(long long delta, int slot) {
    return ((unsigned long long)(unsigned)(delta + (1ll << 31)) << 32) | (unsigned)slot;
}
__device__ __forceinline__ long long em_delta(unsigned long long k) { return (long long)(unsigned)(k >> 32) - (1ll << 31); }

__global__ void __launch_bounds__(EM_THREADS) extra_mileage_kernel(const InstDev I, int *succ_out, long long *cost_out,
                                                                   unsigned char *gwork) {
    extern __shared__ __align__(16) unsigned char em_smem[];
    __shared__ unsigned long long s_red[EM_THREADS / 32];
    __shared__ int s_qn;
    const int n = I.n;
    // the state lives in shared memory when it fits (21 bytes per node), else in a global work buffer (same code, L2 latency)
    unsigned char *state = gwork ? gwork : em_smem;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(state);  // [n] best (delta, slot) of an unvisited node
    int *ea = reinterpret_cast<int *>(key + n);                                  // [n] tour edges in edges_visited[] order
    int *eb = ea + n;
    int *queue = eb + n;                                                         // [n] nodes to rescan
    unsigned char *vis = reinterpret_cast<unsigned char *>(queue + n);           // [n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EM_THREADS / 32;

    auto block_min = [&](unsigned long long v) -> unsigned long long {
        for (int m = 16; m > 0; m >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
            v = o < v ? o : v;
        }
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        v = s_red[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) v = s_red[w] < v ? s_red[w] : v;
        __syncthreads();
        return v;
    };

    // ---- the two farthest nodes: strict '>' over the row-major (i<j) scan -> largest distance, then lowest (i,j) ----
    // two reductions: the largest distance, then the lowest (i, j) that attains it (any n)
    unsigned long long far = EM_NONE;
    for (long long p = tid; p < (long long)n * n; p += EM_THREADS) {
        const int i = (int)(p / n), j = (int)(p % n);
        if (j <= i) continue;
        const long long d = dist_nodes(I, i, j);
        if (d <= 0) continue;  // max_dist starts at 0 and the test is strict
        const unsigned long long k = (unsigned long long)((1ll << 40) - d);
        far = k < far ? k : far;
    }
    far = block_min(far);
    int nodeA = 0, nodeB = 1;  // the reference's defaults when every distance is 0
    if (far != EM_NONE) {
        const long long dmax = (1ll << 40) - (long long)far;
        unsigned long long who = EM_NONE;
        for (long long p = tid; p < (long long)n * n; p += EM_THREADS) {
            const int i = (int)(p / n), j = (int)(p % n);
            if (j <= i) continue;
            if (dist_nodes(I, i, j) == dmax) {
                const unsigned long long k = ((unsigned long long)i << 32) | (unsigned)j;
                who = k < who ? k : who;
            }
        }
        who = block_min(who);
        nodeA = (int)(who >> 32);
        nodeB = (int)(who & 0xffffffffu);
    }
    const long long dAB = dist_nodes(I, nodeA, nodeB);
    for (int k = tid; k < n; k += EM_THREADS) vis[k] = (k == nodeA || k == nodeB) ? 1 : 0;
    if (tid == 0) {
        ea[0] = nodeA; eb[0] = nodeB;
        ea[1] = nodeB; eb[1] = nodeA;
    }
    __syncthreads();
    int cnt = 2;
    long long obj = 2 * dAB;
    // initial caches: both slots
    for (int i = tid; i < n; i += EM_THREADS) {
        if (vis[i]) { key[i] = EM_NONE; continue; }
        const long long dia = dist_nodes(I, nodeA, i), dib = dist_nodes(I, i, nodeB);
        // slot 0 = (A,B): C_Ai + C_iB - C_AB ; slot 1 = (B,A): C_Bi + C_iA - C_BA  (the metrics are symmetric, evaluated as written)
        const long long d0 = dia + dib - dAB;
        const long long d1 = dist_nodes(I, nodeB, i) + dist_nodes(I, i, nodeA) - dist_nodes(I, nodeB, nodeA);
        const unsigned long long k0 = em_key(d0, 0), k1 = em_key(d1, 1);
        key[i] = k0 < k1 ? k0 : k1;
    }
    __syncthreads();

    while (cnt < n) {
        // ---- selection: lexicographic min of (delta, node, slot) ----
        unsigned long long best = EM_NONE;
        for (int i = tid; i < n; i += EM_THREADS) {
            const unsigned long long k = key[i];
            if (k == EM_NONE) continue;
            // compare (delta, i): delta is the high word of k; the slot comes from key[i*] afterwards
            const unsigned long long c = (k & 0xffffffff00000000ull) | (unsigned)i;
            best = c < best ? c : best;
        }
        best = block_min(best);
        if (best == EM_NONE) break;
        const int istar = (int)(best & 0xffffffffu);
        const unsigned long long kstar = key[istar];
        const int jstar = (int)(kstar & 0xffffffffu);
        const long long dstar = em_delta(kstar);
        const int a = ea[jstar], b = eb[jstar];
        __syncthreads();
        // ---- insertion: slot j* becomes (a, i*), (i*, b) is appended ----
        if (tid == 0) {
            eb[jstar] = istar;
            ea[cnt] = istar; eb[cnt] = b;
            vis[istar] = 1;
            key[istar] = EM_NONE;
            s_qn = 0;
        }
        obj += dstar;
        __syncthreads();
        const int newslot = cnt;
        cnt += 1;
        // ---- cache update ----
        const long long d_a_is = dist_nodes(I, a, istar), d_is_b = dist_nodes(I, istar, b);
        for (int i = tid; i < n; i += EM_THREADS) {
            const unsigned long long k = key[i];
            if (k == EM_NONE) continue;
            if ((int)(k & 0xffffffffu) == jstar) {
                queue[atomicAdd(&s_qn, 1)] = i;  // its best edge is gone: rescan
            } else {
                const long long dii = dist_nodes(I, i, istar);
                const long long de1 = dist_nodes(I, a, i) + dii - d_a_is;                     // slot j*: (a, i*)
                const long long de2 = dist_nodes(I, istar, i) + dist_nodes(I, i, b) - d_is_b; // new slot: (i*, b)
                unsigned long long kk = k;
                const unsigned long long k1 = em_key(de1, jstar), k2 = em_key(de2, newslot);
                kk = k1 < kk ? k1 : kk;
                kk = k2 < kk ? k2 : kk;
                key[i] = kk;
            }
        }
        __syncthreads();
        const int qn = s_qn;
        for (int q = warp; q < qn; q += NW) {
            const int i = queue[q];
            unsigned long long kk = EM_NONE;
            for (int j = lane; j < cnt; j += 32) {
                const int u = ea[j], v = eb[j];
                const long long d = dist_nodes(I, u, i) + dist_nodes(I, i, v) - dist_nodes(I, u, v);
                const unsigned long long k = em_key(d, j);
                kk = k < kk ? k : kk;
            }
            for (int m = 16; m > 0; m >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, kk, m);
                kk = o < kk ? o : kk;
            }
            if (lane == 0) key[i] = kk;
        }
        __syncthreads();
    }
    for (int j = tid; j < cnt; j += EM_THREADS) succ_out[ea[j]] = eb[j];
    if (tid == 0) *cost_out = obj;
}

size_t extra_mileage_state_bytes(int n) { return (size_t)n * (8 + 4 + 4 + 4 + 1) + 64; }

// gwork: extra_mileage_state_bytes(n) bytes of global memory, used when the state does not fit one block's shared memory
cudaError_t launch_extra_mileage(const InstDev &I, int *succ_out, long long *cost_out, unsigned char *gwork, bool force_global,
                                 cudaStream_t st) {
    size_t smem = extra_mileage_state_bytes(I.n);
    const bool in_smem = smem <= 200 * 1024 && !(force_global && gwork);
    if (!in_smem && !gwork) return cudaErrorInvalidValue;
    if (!in_smem) smem = 0;
    cudaError_t e = cudaFuncSetAttribute(extra_mileage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(in_smem ? smem : 0));
    if (e != cudaSuccess) return e;
    extra_mileage_kernel<<<1, EM_THREADS, smem, st>>>(I, succ_out, cost_out, in_smem ? nullptr : gwork);
    return cudaGetLastError();
}

}  // namespace tspb
