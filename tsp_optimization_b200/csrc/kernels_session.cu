// kernels_session.cu — device side of the RESIDENT sessions: the perturbation steps of the reference's meta-heuristics
// applied to the tour that already lives in HBM, so that a VNS / tabu / GA loop never moves a tour (or the n(n-1)/2-int
// tabu list) across PCIe between two 2-opt calls.
//   VNS   kick():  reference src/vns.c:11-100  — three tour indices, segment exchange a [b..c] [d..e] f -> a [d..e] [b..c] f
//   tabu  kick:    reference src/tabusearch.c:262-309 — random pair (a, b), four check_tenure() tests, 2-opt move, two
//                  tabu entries set
//   GA    chromosomes <-> successor arrays (reference src/genetic.c:34-44 from_chromosome_to_edges and :426-443)
// The random numbers stay with the caller (the reference draws them from glibc random()); only indices come down.
#include "tsp_state.cuh"

namespace tspb {

// ---- VNS kick ------------------------------------------------------------------------------------------------------
// Tour indices count from node 0 (reference vns.c:15-22 builds tour[] by walking the successors from node 0), i.e. tour
// index t lives at position (pos[0] + t) mod n.  idx1 < idx2 < idx3, idx2 - idx1 >= 2, idx3 - idx2 >= 2 (vns.c:25-50).
// a = tour[idx1], b = tour[idx1+1], c = tour[idx2], d = tour[idx2+1], e = tour[idx3], f = tour[idx3+1]; new successors
// a -> d, e -> b, c -> f (vns.c:54-62): the blocks X = tour[idx1+1..idx2] and Y = tour[idx2+1..idx3] change places, nothing
// is reversed.  For idx3 = n-1 the reference reads tour[n], one element past its calloc'ed array; here f wraps to tour[0]
// (= node 0, which closes the cycle — the only value for which the reference's result is a tour at all).
//
// Step 1 copies the records of X and Y to a scratch buffer, step 2 writes them back in the new order; every read of
// step 2 goes to the scratch buffer or to positions outside the rewritten range, so no thread reads what another writes.
__global__ void __launch_bounds__(256) vns_kick_gather_kernel(const TourDev T, int idx1, int L, float4 *scratch) {
    const int n = T.n;
    const int base = T.pos[0];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) {
        int p = base + idx1 + 1 + k;
        if (p >= n) p -= n;
        if (p >= n) p -= n;
        scratch[k] = T.rec[p];
    }
}

__global__ void __launch_bounds__(256) vns_kick_scatter_kernel(const InstDev I, const TourDev T, int idx1, int L1, int L2,
                                                               const float4 *scratch) {
    const int n = T.n;
    const int base = T.pos[0];
    const int L = L1 + L2;
    auto wrap = [n](int p) {
        if (p >= n) p -= n;
        if (p >= n) p -= n;
        return p;
    };
    auto src = [&](int k) { return k < L2 ? scratch[L1 + k] : scratch[k - L2]; };
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= L; k += gridDim.x * blockDim.x) {
        if (k == L) {  // node a: new successor d = first node of Y
            const int pa = wrap(base + idx1);
            const float4 ra = T.rec[pa];
            const float4 rd = scratch[L1];
            const float ds = (float)dist_nodes(I, node_of(ra), node_of(rd));
            T.rec[pa].z = ds;
            continue;
        }
        const float4 r = src(k);
        float4 nxt;
        float ds = r.z;
        if (k == L - 1) {  // node c: new successor f
            const int pf = wrap(base + idx1 + 1 + L);  // tour index idx3 + 1, wrapping to tour index 0
            nxt = T.rec[pf];
            ds = (float)dist_nodes(I, node_of(r), node_of(nxt));
        } else {
            nxt = src(k + 1);
            if (k == L2 - 1) ds = (float)dist_nodes(I, node_of(r), node_of(nxt));  // node e: new successor b
        }
        const int p = wrap(base + idx1 + 1 + k);
        const int u = node_of(r);
        T.rec[p] = make_float4(r.x, r.y, ds, r.w);
        T.pos[u] = p;
        if (p == 0) {  // rec[n] mirrors rec[0] (wrap-around successor of position n-1)
            *reinterpret_cast<float2 *>(&T.rec[n].x) = make_float2(r.x, r.y);
            T.rec[n].w = r.w;
        }
    }
}

cudaError_t launch_vns_kick(const InstDev &I, const TourDev &T, int idx1, int idx2, int idx3, float4 *scratch, cudaStream_t st) {
    const int L1 = idx2 - idx1, L2 = idx3 - idx2, L = L1 + L2;
    int grid = (L + 256) / 256;
    if (grid > 1184) grid = 1184;
    vns_kick_gather_kernel<<<grid, 256, 0, st>>>(T, idx1, L, scratch);
    vns_kick_scatter_kernel<<<grid, 256, 0, st>>>(I, T, idx1, L1, L2, scratch);
    return cudaGetLastError();
}

// ---- tabu kick -------------------------------------------------------------------------------------------------------
// reference src/utility.c:17-30 x_udir_pos in 64 bits
__device__ __forceinline__ long long kick_udir(int i, int j, int n) {
    if (i > j) { const int t = i; i = j; j = t; }
    return (long long)i * n + j - ((long long)(i + 1) * (i + 2)) / 2;
}
// reference src/tabusearch.c:83-92 check_tenure (lazy expiry included)
__device__ __forceinline__ bool kick_check(int *skip, long long e, int iter, int tenure) {
    if (iter < 0 || tenure < 0) return false;
    const int v = skip[e];
    if (v == 0) return false;
    if (iter - v > tenure) {
        skip[e] = 0;
        return false;
    }
    return true;
}

// One thread walks the caller's candidate pairs in order, exactly like the `while (1)` of reference tabusearch.c:264-291:
// a candidate with a == b or touching edges is skipped without any tabu test; otherwise the four edges (a,a1), (b,b1),
// (a,b), (a1,b1) are tested with && short-circuit (so the lazy expiry hits the same entries); the first candidate that
// passes is published as the move to apply and its two removed edges enter the tabu list with the current iteration
// (:293-307).  *accepted = index of that candidate, -1 when none passed, -2 on an out-of-range node.
__global__ void tabu_kick_select_kernel(const TourDev T, int *skip, const int *pairs, int count, int iter, int tenure, int *accepted) {
    Ctl *ctl = T.ctl;
    const int n = T.n;
    ctl->ap_valid = 0;
    *accepted = -1;
    for (int c = 0; c < count; ++c) {
        const int a = pairs[2 * c], b = pairs[2 * c + 1];
        if (a < 0 || b < 0 || a >= n || b >= n) { *accepted = -2; return; }
        const int pa = T.pos[a], pb = T.pos[b];
        const int a1 = node_of(T.rec[pa + 1]), b1 = node_of(T.rec[pb + 1]);  // rec[n] mirrors rec[0]
        if (a == b || a1 == b || b1 == a) continue;
        const long long e1 = kick_udir(a, a1, n), e2 = kick_udir(b, b1, n), e3 = kick_udir(a, b, n), e4 = kick_udir(a1, b1, n);
        if (!kick_check(skip, e1, iter, tenure) && !kick_check(skip, e2, iter, tenure) && !kick_check(skip, e3, iter, tenure) &&
            !kick_check(skip, e4, iter, tenure)) {
            skip[e1] = iter;
            skip[e2] = iter;
            ctl->ap_pa = pa;
            ctl->ap_pb = pb;
            ctl->ap_valid = 1;
            *accepted = c;
            return;
        }
    }
}

cudaError_t launch_tabu_kick_select(const TourDev &T, int *skip, const int *pairs, int count, int iter, int tenure, int *accepted,
                                    cudaStream_t st) {
    tabu_kick_select_kernel<<<1, 1, 0, st>>>(T, skip, pairs, count, iter, tenure, accepted);
    return cudaGetLastError();
}

// ---- populations: chromosome (visiting order) <-> successor array ------------------------------------------------------
// reference src/genetic.c:34-44 from_chromosome_to_edges: succ[chromosome[k]] = chromosome[k+1] (cyclic)
__global__ void __launch_bounds__(256) order_to_succ_kernel(const int *orders, int *pop, const int *slots, int n) {
    const int *o = orders + (long long)blockIdx.x * n;
    int *s = pop + (long long)(slots ? slots[blockIdx.x] : (int)blockIdx.x) * n;
    for (int k = threadIdx.x; k < n; k += 256) s[o[k]] = o[k + 1 == n ? 0 : k + 1];
}

__global__ void __launch_bounds__(256) copy_succ_kernel(const int *src, int *pop, const int *slots, int n) {
    const int *o = src + (long long)blockIdx.x * n;
    int *s = pop + (long long)(slots ? slots[blockIdx.x] : (int)blockIdx.x) * n;
    for (int k = threadIdx.x; k < n; k += 256) s[k] = o[k];
}

// reference src/genetic.c:437-441: the chromosome is read off the successors starting at node 0.  One warp-sized block per
// tour; the walk itself is sequential (lane 0), n dependent loads from L2.
__global__ void succ_to_order_kernel(const int *pop, const int *slots, int *orders, int n, int as_order) {
    const int *s = pop + (long long)(slots ? slots[blockIdx.x] : (int)blockIdx.x) * n;
    int *o = orders + (long long)blockIdx.x * n;
    if (!as_order) {
        for (int k = threadIdx.x; k < n; k += blockDim.x) o[k] = s[k];
        return;
    }
    if (threadIdx.x == 0) {
        int at = 0;
        for (int k = 0; k < n; ++k) {
            o[k] = at;
            at = s[at];
        }
    }
}

cudaError_t launch_population_store(const int *staged, int *pop, const int *slots, int n, int count, int as_order, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    if (as_order) order_to_succ_kernel<<<count, 256, 0, st>>>(staged, pop, slots, n);
    else copy_succ_kernel<<<count, 256, 0, st>>>(staged, pop, slots, n);
    return cudaGetLastError();
}

cudaError_t launch_population_fetch(const int *pop, const int *slots, int *out, int n, int count, int as_order, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    succ_to_order_kernel<<<count, 32, 0, st>>>(pop, slots, out, n, as_order);
    return cudaGetLastError();
}

}  // namespace tspb
