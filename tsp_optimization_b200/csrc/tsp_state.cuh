// tsp_state.cuh — device-resident instance / tour state and the block-wide move application.
//
// HBM layout.  A tour lives on the device in two views that are kept in sync:
//   position space  rec[p] = {x, y, ds, node}   p = 0..n-1 in visiting order (forward orientation),
//                   ds = exact integer d(node_p, node_{p+1}) stored as float, node id in the 4th word;
//                   rec[n] mirrors rec[0]'s coordinates with ds = -BIG (wrap-around successor of p=n-1);
//                   everything after is padding with ds = -BIG (never a candidate).
//                   pos[node] = p.
//   node space      nrec[k] = {x_k, y_k, x_succ(k), y_succ(k)}, nds[k] = d(k, succ k), nsucc[k]
//                   (only what the first-improvement search needs: it scans node-index order).
// The reference keeps edges[k].j = succ(k) and a prev[] array (reference include/utility.h:131-145,
// src/heuristics.c:444-448); succ[] is rebuilt from rec[] when a tour is downloaded.
#pragma once
#include "tsp_device.cuh"

namespace tspb {

struct InstDev {
    int n;
    int metric;             // reference weight_type value
    int exact32;            // every coordinate is exactly representable in FP32
    int fp32_ok;            // FP32 filter path usable (metric in EUC/CEIL/ATT, |coords| small enough)
    float W;                // filter window: candidate iff Q <= best + W
    float band;             // matrix kernel guard band, relative to r
    const double2 *pt64;    // node-indexed exact points (GEO: lat/lon radians)
    const float2 *pt32;     // node-indexed FP32 points
    const int *dmat;        // optional n x ld int32 distance matrix (matrix mode), else nullptr
    long long dmat_ld;
};

struct MoveRec {  // one applied move, as logged for parity tests
    int i, j;
    long long delta;
};

struct Ctl {
    int done;                       // 1 = local optimum reached (all later launches return at once)
    int error;                      // device-side consistency check failed
    unsigned ticket;                // block completion counter of the current launch
    int hint;                       // best exact delta found so far in the current BI pass (<= 0)
    long long passes;               // BI: completed scans; FI: completed sweeps
    long long moves;                // applied moves
    long long obj_delta;            // sum of applied deltas
    long long launches;             // kernel launches that did work
    long long sweep_moves;          // FI: moves applied in the current sweep
    int cur_i, cur_j;               // FI cursor (next pair to look at, node-index order)
    unsigned long long fi_found;    // FI: min linear index i*n+j of an improving pair in this launch
    unsigned long long packed;      // multi-GPU: this rank's packed best key for the NCCL min-allreduce
    long long log_count;
    long long max_moves;            // FI only: stop after this many moves (<0 = unlimited)
    long long pairs_swept;          // FI: linear pairs covered (statistics only)
    MoveKey last;                   // last selected key
};

struct TourDev {
    int n;
    int alloc;        // entries allocated in rec
    float4 *rec;
    int *pos;
    float4 *nrec;     // may be nullptr when FI is not in use
    float *nds;
    int *nsucc;
    Ctl *ctl;
    MoveKey *block_best;
    MoveRec *log;
    long long log_cap;
};

#define FI_NONE 0xffffffffffffffffull

__device__ __forceinline__ int node_of(const float4 &r) { return __float_as_int(r.w); }

// distance between two nodes by id, exact (matrix lookup when a matrix is resident).
__device__ __forceinline__ long long dist_nodes(const InstDev &I, int u, int v) {
    if (I.dmat) return (long long)I.dmat[(long long)u * I.dmat_ld + v];
    return exact_dist(I.metric, I.pt64[u], I.pt64[v]);
}

// Block-wide application of the 2-opt move (i,j), i<j node ids, exactly as the reference does it:
//   a=i, b=j, a1=succ[a], b1=succ[b]; succ[a]=b; succ[a1]=b1; reverse_path(b, a1)
// (reference src/heuristics.c:476-483, src/tabusearch.c:161-165, src/utility.c:708-722): the FORWARD path
// a1 -> ... -> b is reversed, wrap-around included, never "the shorter side", so the orientation of the
// tour — and with it the (i,j) -> (a1,b1) mapping of every later move — stays the reference's.
// In position space that is an in-place reversal of the cyclic range [pos[a]+1, pos[b]].
// Must be called by every thread of the block; ends with a __syncthreads().
__device__ __forceinline__ void apply_move_block(const InstDev &I, const TourDev &T, int i, int j) {
    const int n = T.n;
    const int tid = threadIdx.x, nt = blockDim.x;
    float4 *rec = T.rec;
    const int pa = T.pos[i];
    const int pb = T.pos[j];
    __syncthreads();  // everyone has read pos[] before anyone rewrites it
    int s = pa + 1;
    if (s >= n) s -= n;
    int len = pb - pa;
    if (len < 0) len += n;  // number of nodes on the path a1..b
    const int e = pb;
    // (x, y, node) of positions s+t <-> e-t
    const int half = len >> 1;
    for (int t = tid; t < half; t += nt) {
        int A = s + t;
        if (A >= n) A -= n;
        int B = e - t;
        if (B < 0) B += n;
        float2 xa = *reinterpret_cast<float2 *>(&rec[A].x);
        float wa = rec[A].w;
        float2 xb = *reinterpret_cast<float2 *>(&rec[B].x);
        float wb = rec[B].w;
        *reinterpret_cast<float2 *>(&rec[A].x) = xb;
        rec[A].w = wb;
        *reinterpret_cast<float2 *>(&rec[B].x) = xa;
        rec[B].w = wa;
        T.pos[__float_as_int(wb)] = A;
        T.pos[__float_as_int(wa)] = B;
    }
    // inner edge lengths: positions s .. s+len-2 are reversed among themselves
    const int m = len - 1;
    const int mhalf = m >> 1;
    for (int t = tid; t < mhalf; t += nt) {
        int A = s + t;
        if (A >= n) A -= n;
        int Cc = s + m - 1 - t;
        if (Cc >= n) Cc -= n;
        float za = rec[A].z, zc = rec[Cc].z;
        rec[A].z = zc;
        rec[Cc].z = za;
    }
    __syncthreads();
    if (tid == 0) {
        // new edges (a,b) at position pa and (a1,b1) at position pb
        int pa1 = pa + 1; if (pa1 >= n) pa1 -= n;
        int pb1 = pb + 1; if (pb1 >= n) pb1 -= n;
        int na = node_of(rec[pa]), nb = node_of(rec[pa1]);
        int na1 = node_of(rec[pb]), nb1 = node_of(rec[pb1]);
        rec[pa].z = (float)dist_nodes(I, na, nb);
        rec[pb].z = (float)dist_nodes(I, na1, nb1);
        float4 r0 = rec[0];
        rec[n] = make_float4(r0.x, r0.y, -TSPB_BIG, r0.w);
    }
    __syncthreads();
    if (T.nrec) {
        // node-space view: every node at positions pa..pb (cyclic) got a new successor
        for (int t = tid; t <= len; t += nt) {
            int P = pa + t;
            if (P >= n) P -= n;
            int Pn = P + 1;
            if (Pn >= n) Pn -= n;
            float4 rp = rec[P];
            float4 rn = rec[Pn];
            int k = node_of(rp);
            T.nrec[k] = make_float4(rp.x, rp.y, rn.x, rn.y);
            T.nds[k] = rp.z;
            T.nsucc[k] = node_of(rn);
        }
        __syncthreads();
    }
}

// Exact delta of the move (i,j) in the CURRENT tour, from node ids (used to re-derive the delta of a
// selected move and as a device-side consistency check).
__device__ __forceinline__ long long move_delta_nodes(const InstDev &I, const TourDev &T, int i, int j) {
    const int n = T.n;
    int pa = T.pos[i], pb = T.pos[j];
    int pa1 = pa + 1; if (pa1 >= n) pa1 -= n;
    int pb1 = pb + 1; if (pb1 >= n) pb1 -= n;
    int a1 = node_of(T.rec[pa1]), b1 = node_of(T.rec[pb1]);
    return dist_nodes(I, i, j) + dist_nodes(I, a1, b1) - (long long)T.rec[pa].z - (long long)T.rec[pb].z;
}

// ---- kernel argument blocks shared by the kernel translation units and engine.cu ----------------------
struct BiArgs {
    InstDev inst;
    TourDev tour;
    const int *tile_row_start;  // [ntr+1] prefix sums of tiles per tile-row
    const int *tile_row_j0;     // [ntr] first tile column of each tile-row
    int ntr;
    int ntiles;
    int TJ;          // columns per tile (even)
    int rank, world; // tiles are dealt round-robin over ranks (multi-GPU neighbourhood sharding)
    int fuse_apply;  // 1: the last block applies the move (single GPU); 0: it only publishes the key
};

struct NnArgs {
    InstDev inst;
    int start;
    int *succ;                       // out
    unsigned char *visited;          // n bytes, zeroed by the launcher
    unsigned long long *slots;       // [2] packed (dist << 32 | idx) winners, double-buffered; init to ~0
    unsigned *barrier;               // grid barrier counter, zeroed by the launcher
    long long *cost;                 // out: accumulated cost
};

}  // namespace tspb
