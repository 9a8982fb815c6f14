// kernels_bi.cu — best-improvement 2-opt pass (reference src/tabusearch.c:107-178 alg_2opt_tabu with a
// NULL tabu list): one launch = one full scan of all n(n-3)/2 non-adjacent pairs + argmin + apply.
//
// The scan runs in POSITION space.  For positions p<q with u=node(p), v=node(q) the reference pair is
// (i,j) = (min(u,v), max(u,v)); its two removed edges are (node(p),node(p+1)) and (node(q),node(q+1))
// whichever of u,v is the smaller index, and
//     delta = D[p][q] + D[p+1][q+1] - ds[p] - ds[q],      D[p][q] = d(node(p), node(q)).
// Each D value is therefore used by two pairs, (p,q) and (p-1,q-1): a thread owns R consecutive rows
// and marches along the columns, so it needs R+1 distances per R evaluated moves — R of its own and
// the first one of the thread below, which the SHUF variants fetch from the next lane with a shuffle.
//
// Arithmetic: FP32 (2 FADD, FMUL, FFMA, MUFU.SQRT per distance) as a FILTER: Q = D1 + D2 - ds_p is
// compared with thr + ds_q where thr = (best exact delta so far) + W.  W bounds the worst FP32/rounding
// excess, so the true argmin always passes; every pair that passes is re-evaluated in
// FP64 with the reference's exact operation order and only those exact integer deltas enter the
// (delta, i, j) argmin.  Column records are staged in shared memory by TMA bulk copies (UBLKCP),
// double-buffered across tiles.
#include "tsp_state.cuh"

namespace tspb {



template <bool ATT>
__device__ __forceinline__ float dist32(float ax, float ay, float bx, float by) {
    float dx = ax - bx;
    float dy = ay - by;
    float s = fmaf(dy, dy, dx * dx);
    if (ATT) s *= 0.1f;
    return sqrt_approx(s);
}


// distances from R+1 consecutive rows to one column point: rows 0..R-1 in R/2 packed pairs, row R scalar
// LAST: how D[R] — the distance from the row just below this thread's rows, i.e. the first row of the next lane — is
// obtained: 1 = computed here (scalar, the (R+1)-th square root of the column), 2 = taken from the next lane with one
// shuffle (SHUF kernels: the lanes of a warp own adjacent row groups, so that distance is lane+1's D[0]), 0 = not needed.
template <int R, bool ATT, int LAST>
__device__ __forceinline__ void column_dists(const f32x2 (&xr2)[R / 2], const f32x2 (&yr2)[R / 2], float xrl, float yrl,
                                             float cx, float cy, float (&D)[R + 1]) {
    const f32x2 cxx = f2pack(cx, cx), cyy = f2pack(cy, cy);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
        f32x2 dx = f2sub(xr2[k], cxx);
        f32x2 dy = f2sub(yr2[k], cyy);
        f32x2 s = f2fma(dy, dy, f2mul(dx, dx));
        if (ATT) s = f2mul(s, f2pack(0.1f, 0.1f));
        D[2 * k] = sqrt_approx(f2lo(s));
        D[2 * k + 1] = sqrt_approx(f2hi(s));
    }
    if (LAST == 1) D[R] = dist32<ATT>(xrl, yrl, cx, cy);
    else if (LAST == 2) D[R] = __shfl_down_sync(0xffffffffu, D[0], 1);
    else D[R] = 0.f;
}

// ---- cold path ------------------------------------------------------------------------------------------
// Exact re-evaluation of one filter hit: the R x BI_CB pairs (rows p0..p0+R-1) x (columns Q0+jj0 .. +BI_CB-1) of ONE
// thread whose FP32 block minimum passed the threshold.  WARP-COOPERATIVE: the whole warp is converged at the call (the
// hot loop has no divergent branch, the hit test is a ballot), so the 32 lanes take one pair each — column records
// from shared memory, row records from L2 — re-apply the FP32 filter, evaluate the survivors exactly (FP64 with the reference's
// operation order, reference src/tabusearch.c:150 / src/distutil.c) and reduce the best exact (delta, i, j) key with
// shuffles.  A hit costs a few hundred cycles instead of a serial walk over 32 pairs by a single lane.
constexpr int BI_CB = 4;  // columns per filter check

template <int R, bool ATT, bool EXACT32>
__device__ __noinline__ MoveKey bi_cold_warp(const InstDev I, const float4 *rec, const float4 *sc, int n, int p0, int Q0,
                                             int jj0, float thr) {
    const int lane = threadIdx.x & 31;
    MoveKey best = key_none();
#pragma unroll 1
    for (int base = 0; base < R * BI_CB; base += 32) {
        const int idx = base + lane;
        const int r = idx / BI_CB, c = idx % BI_CB;
        const int p = p0 + r, q = Q0 + jj0 + c;
        if (idx < R * BI_CB && q >= p + 2 && q <= n - 1 && !(p == 0 && q == n - 1)) {  // reference tabusearch.c:134
            const float4 rp = __ldg(&rec[p]), rp1 = __ldg(&rec[p + 1]);  // row records: one L2 round trip for the warp
            const float4 c0 = sc[jj0 + c], c1 = sc[jj0 + c + 1];
            const float qv = (dist32<ATT>(rp.x, rp.y, c0.x, c0.y) - rp.z) + dist32<ATT>(rp1.x, rp1.y, c1.x, c1.y);
            if (qv <= thr + c0.z) {
                const int u = node_of(rp), v = node_of(c0);
                long long d1, d2;
                if (EXACT32) {
                    d1 = exact_dist(I.metric, make_double2((double)rp.x, (double)rp.y), make_double2((double)c0.x, (double)c0.y));
                    d2 = exact_dist(I.metric, make_double2((double)rp1.x, (double)rp1.y), make_double2((double)c1.x, (double)c1.y));
                } else {
                    d1 = exact_dist(I.metric, I.pt64[u], I.pt64[v]);
                    d2 = exact_dist(I.metric, I.pt64[node_of(rp1)], I.pt64[node_of(c1)]);
                }
                const long long delta = d1 + d2 - (long long)rp.z - (long long)c0.z;
                if (delta < 0) {
                    MoveKey k;
                    k.delta = (int)delta; k.i = min(u, v); k.j = max(u, v); k.pad = 0;
                    if (key_less(k, best)) best = k;
                }
            }
        }
    }
    __syncwarp();
    return key_warp_min(best);
}

// BI_THREADS x R rows per tile: 256 x 8 for big instances; smaller blocks (64 / 128 threads, up to 8 per SM) give
// mid-size instances (n ~ 10^4: only ~20 k evaluations per warp and pass) enough tiles to fill 148 SMs while keeping
// R = 8 rows per thread, i.e. 1.125 sqrt per evaluated move.
//
// Shared memory per block (dynamic): two column buffers (TJ+2 records) filled by TMA bulk copies, one mbarrier per stage,
// plus the tile tables.  Row records go straight from L2 into registers: a thread's R+1 rows are 16*(R+1) contiguous bytes,
// and reading them through shared memory would put all lanes of a quarter-warp on the same banks (stride 16*R bytes).
//
// SHUF: a warp owns 32 R - 1 consecutive rows instead of 32 R: lane L holds rows p0 .. p0+R-1 with p0 = warp base + L R, the
// distance of the row below its last one is lane L+1's first distance of the same column (one SHFL instead of a square root
// and four FP32 instructions), and lane 31's last row — whose lower neighbour lives in another warp — is masked out and
// scanned again as the first row of the next warp.  R square roots per R moves (minus 1/32R): the algorithmic minimum.
#ifndef TSPB_BI_MINBLOCKS64
// Resident 64-thread blocks per SM the compiler must leave room for (register cap 65536 / (64 x this)).  Measured on B200,
// 64 x 8 x 256 row-shuffle kernel at n = 100 000 (profiles/r2_blocks_per_sm_ab.jsonl): 8 blocks (100 registers) 1288 us per
// pass, 10 blocks (94 registers, no spills) 1276 us — and 184 vs 189 us for one rank's share of eight —, 12 blocks (80
// registers, spills) 1321 us.  engine.cu (bi_blocks_per_sm) sizes the grid to match.  The pruned variants (more live state: they
// would spill under the lower cap) keep 8.
#define TSPB_BI_MINBLOCKS64 10
#endif
template <int BI_THREADS, int R, bool ATT, bool EXACT32, bool PRUNED, bool SHUF>
__global__ void __launch_bounds__(BI_THREADS, (R >= 16 ? (BI_THREADS == 256 ? 1 : 384 / BI_THREADS) : (BI_THREADS == 64 && !PRUNED ? TSPB_BI_MINBLOCKS64 : 512 / BI_THREADS))) bi_scan_kernel(const BiArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ int s_hint;
    __shared__ int s_last;
    __shared__ MoveKey s_keys[BI_THREADS / 32];
    __shared__ int s_ap[3];

    Ctl *ctl = A.tour.ctl;
    const int tid = threadIdx.x;
    // Programmatic dependent launch: let the apply kernel queue up behind us right away, and do everything that does not
    // depend on the previous kernel (barrier init, the static tile tables) before waiting for it.
    pdl_launch_dependents();
    const int n = A.inst.n;
    const int TJ = A.TJ;
    constexpr int TI = SHUF ? (BI_THREADS / 32) * (32 * R - 1) : BI_THREADS * R;  // rows (moves) of a tile
    const float W = A.inst.W;
    const float4 *rec = A.tour.rec;
    float4 *scols0 = reinterpret_cast<float4 *>(smem_raw);
    float4 *scols1 = scols0 + (TJ + 2);
    const unsigned col_bytes = (unsigned)(TJ + 1) * 16u;

    // tile tables -> shared memory (one coalesced L2 round trip instead of a dependent chain per binary-search step)
    int *s_rs = reinterpret_cast<int *>(scols1 + (TJ + 2));  // [ntr+1] prefix sums of tiles per tile-row
    int *s_rj = s_rs + (A.ntr + 1);                          // [ntr]   first tile column of each tile-row
    for (int k = tid; k <= A.ntr; k += BI_THREADS) {
        s_rs[k] = A.tile_row_start[k];
        if (k < A.ntr) s_rj[k] = A.tile_row_j0[k];
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    pdl_wait();  // the previous kernel of the stream (apply / upload) is complete and its writes are visible
    // everything the prologue needs from global memory is requested at once (ONE L2 round trip instead of a chain of four
    // dependent ones, ~0.5 us each: at n = 10 000 a pruned pass has one small tile per block and nothing to hide them behind)
    const int done = *((volatile int *)&ctl->done);
    const int hint0 = *((volatile int *)&ctl->hint);
    const unsigned live_count0 = PRUNED ? __ldcg(&ctl->live_count) : 0u;
    int2 live_first = make_int2(0, 0);
    if (PRUNED && tid == 0) live_first = __ldcg(&A.tour.live[blockIdx.x]);  // in bounds whatever live_count is (engine.cu)
    if (tid == 0) s_hint = hint0;
    if (done) {  // local optimum already reached: later launches of the same batch return at once
        if (blockIdx.x == 0 && tid == 0) ctl->ap_valid = 0;
        return;
    }
    if (A.timing && tid == 0) {
        const unsigned long long t = globaltimer_ns();
        atomicMin(&ctl->tm_scan_first, t);
        if (A.timing == 2) {
            A.dbg[2 * blockIdx.x] = t;
            if (PRUNED) {  // phase stamps: pruned variants only, so that the exhaustive (headline) kernel's code is not touched
                for (int k = 0; k < 8; ++k) A.dbg[8192 + 8 * blockIdx.x + k] = 0;
                A.dbg[8192 + 8 * blockIdx.x] = t;
            }
        }
    }
    __syncthreads();

    // This block's own winner of the previous pass is most likely still a legal move: its exact delta now is a valid
    // bound for this pass (see seed_hint_from_candidates).  One thread re-evaluates it while the first tile is in
    // flight; the result tightens s_hint / ctl->hint asynchronously.
    if (A.seed_hint && tid == BI_THREADS - 32) {
        const MoveKey c = key_load_cg(&A.tour.block_best[blockIdx.x]);
        const long long d = legal_move_delta_cg(A.inst, A.tour, c);
        if (d < 0) {
            atomicMin(&s_hint, (int)d);
            atomicMin(&ctl->hint, (int)d);
        }
    }

    MoveKey best = key_none();
    float thr = (float)s_hint + W;
    int pend_hint = 0;  // tid 0: value of ctl->hint fetched 64 columns ago

    // Tiles are DRAWN, not dealt: ctl->tile_next counts this rank's tiles handed out beyond the first wave (rank r owns
    // the tile ids r, r + world, ...; every tile costs the same — masked and padded pairs are computed too), so the blocks
    // run dry within one tile of each other whatever slows some of them down (exact-path calls, the far die's L2
    // latency).  Static dealing left the SMs idle ~10 % of a pass once a pass is only ~5 tiles deep (8 ranks).
    // [stage] = {P0, Q0, valid, columns}, written by thread 0 one tile ahead.  The 4th word (the tile width, TJ for every tile)
    // keeps a stage 16 bytes; with the 12-byte layout ptxas moved the hot loop's counter and shared-memory addresses from the
    // uniform datapath into vector registers (+5 % pass time at n = 100 000; tests/test_codegen.py watches the SASS)
    __shared__ __align__(16) int s_tile[2][4];

    // tile id -> (tile row I, tile column J)
    auto decode = [&](int t, int &P0, int &Q0) {
        int lo = 0, hi = A.ntr - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (s_rs[mid] <= t) lo = mid; else hi = mid - 1;
        }
        P0 = lo * TI;
        Q0 = (s_rj[lo] + (t - s_rs[lo])) * TJ;
    };
    // thread 0: take the next tile, start the bulk copy of its column records, publish it for stage b.  A block's first
    // tile is its block index (no atomic on the critical path of the prologue); the following ones are drawn from the
    // counter, which therefore counts from gridDim.x.  tiles_rank <= gridDim.x (one wave, n <~ 10^4) never touches it.
    // Exact tile pruning (A.pruned): the tiles come from this rank's live list (tile_filter_kernel); a live tile whose lower
    // bound meanwhile exceeds the best exact delta found in this pass is dropped at the draw (it cannot hold the argmin
    // nor a tie: ties need delta == best).
    // Exact tile pruning (PRUNED): the tiles come from this rank's live list (tile_filter_kernel); a live tile whose lower
    // bound meanwhile exceeds the best exact delta found in this pass is dropped at the draw (it cannot hold the argmin nor
    // a tie: ties need delta == best).  The exhaustive kernel keeps the loop-free draw: a loop inside this thread-0-only
    // region makes the compiler give up the uniform datapath for the hot loop's counter and shared-memory addresses
    // (+4 % pass time at n = 100 000, measured).
    const long long tiles_rank = PRUNED ? (long long)live_count0
                                        : ((long long)A.ntiles - A.rank + A.world - 1) / A.world;
    unsigned scanned = 0;  // thread 0: tiles this block really scanned (statistics of the pruned mode)
    unsigned next_raw = 0; // thread 0, pruned mode: ticket requested one draw ahead
    auto draw = [&](int b, bool first) {
        int P = 0, Q = 0, valid = 0;
        if (!PRUNED) {
            long long kl = first ? (long long)blockIdx.x : tiles_rank;
            if (!first && tiles_rank > (long long)gridDim.x) kl = (long long)gridDim.x + (long long)atomicAdd(&ctl->tile_next, 1u);
            if (kl < tiles_rank) {
                decode((int)((long long)A.rank + (long long)A.world * kl), P, Q);
                valid = 1;
            }
        } else {
            // pruned tiles are small (a few us): the ticket of the NEXT draw is requested as soon as this one is taken, so the
            // atomic's L2 round trip overlaps the tile instead of preceding it
            const bool counted = tiles_rank > (long long)gridDim.x;
            for (;;) {
                long long kl = first ? (long long)blockIdx.x : (counted ? (long long)gridDim.x + (long long)next_raw : tiles_rank);
                if (kl >= tiles_rank) break;
                const int2 ent = first ? live_first : __ldcg(&A.tour.live[kl]);
                first = false;
                if (counted) next_raw = atomicAdd(&ctl->tile_next, 1u);
                if (__int_as_float(ent.y) > (float)(*((volatile int *)&s_hint))) continue;
                decode(ent.x, P, Q);
                valid = 1;
                scanned += 1;
                break;
            }
        }
        if (valid) {
            mbar_expect_tx(&bars[b], col_bytes);
            tma_load_1d(b ? scols1 : scols0, rec + Q, col_bytes, &bars[b]);
        }
        s_tile[b][0] = P;
        s_tile[b][1] = Q;
        s_tile[b][2] = valid;
        s_tile[b][3] = TJ;
    };

    if (tid == 0) draw(0, true);
    __syncthreads();
    int P0 = s_tile[0][0], Q0 = s_tile[0][1], NCv = s_tile[0][3];
    bool have = s_tile[0][2] != 0;
    if (PRUNED && A.timing == 2 && tid == 0) A.dbg[8192 + 8 * blockIdx.x + 1] = globaltimer_ns();  // first tile drawn, its copy in flight

    for (int it = 0; have; ++it) {
        const int buf = it & 1;
        const unsigned parity = (unsigned)(it >> 1) & 1u;
        const float4 *sc = buf ? scols1 : scols0;
        // prefetch the next tile into the other stage (its slot in s_tile / scols was last read two tiles ago)
        if (tid == 0) draw(buf ^ 1, false);

        // rows of this thread: p0 .. p0+R-1 in packed pairs (+ successor row p0+R, scalar)
        const int p0 = SHUF ? P0 + (tid >> 5) * (32 * R - 1) + (tid & 31) * R : P0 + tid * R;
        f32x2 xr2[R / 2], yr2[R / 2], cp2[R / 2];
        float xrl, yrl;
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const float4 v0 = __ldg(&rec[p0 + 2 * k]), v1 = __ldg(&rec[p0 + 2 * k + 1]);
            // "+ 0" is a real FADD2 (not an identity for -0.0, so it is never folded): its 64-bit result is an aligned
            // register pair that stays live across the column loop, instead of being re-packed with MOVs per step
            const f32x2 zero2 = f2pack(0.f, 0.f);
            xr2[k] = f2add(f2pack(v0.x, v1.x), zero2);
            yr2[k] = f2add(f2pack(v0.y, v1.y), zero2);
            cp2[k] = f2sub(zero2, f2pack(v0.z, v1.z));  // padding rows carry ds = -BIG -> cp = +BIG -> never a candidate
        }
        if (!SHUF) {
            const float4 v = __ldg(&rec[p0 + R]);
            xrl = v.x;
            yrl = v.y;
        } else {
            xrl = yrl = 0.f;
            // lane 31's last row has its lower neighbour in another warp: never a candidate here (it is the next warp's first row)
            if ((tid & 31) == 31) cp2[R / 2 - 1] = f2pack(f2lo(cp2[R / 2 - 1]), TSPB_BIG);
        }
        thr = fminf(thr, (float)(*((volatile int *)&s_hint)) + W);

        mbar_wait(&bars[buf], parity);
        if (PRUNED && A.timing == 2 && tid == 0 && it == 0) A.dbg[8192 + 8 * blockIdx.x + 2] = globaltimer_ns();  // rows and columns of the first tile are here

        // pairs with q < p+2 exist in this tile?  (mask them; they are mirrored / adjacent pairs)
        const bool diag = (Q0 < P0 + TI + 1);
        const int NC = PRUNED ? NCv : TJ;  // columns of this tile
        // first column of the scan: D0 -> U2 (kept per variant so that the regular tile's shared-memory addressing stays uniform)
        float4 c0, cnext;
        f32x2 U2[R / 2];
#define BI_INIT(JB)                                                                                    \
    {                                                                                                  \
        c0 = sc[(JB)];                                                                                 \
        cnext = sc[(JB) + 1];                                                                          \
        float D0[R + 1];                                                                               \
        column_dists<R, ATT, SHUF ? 0 : 1>(xr2, yr2, xrl, yrl, c0.x, c0.y, D0);                        \
        _Pragma("unroll") for (int k = 0; k < R / 2; ++k) U2[k] = f2add(f2pack(D0[2 * k], D0[2 * k + 1]), cp2[k]); \
    }
#define UU(r_) (((r_) & 1) ? f2hi(U2[(r_) >> 1]) : f2lo(U2[(r_) >> 1]))

// one column: R+1 fresh distances -> R move deltas Q[r] = (D[p_r][q] - ds_p) + D[p_r+1][q+1]; the filter quantity
// min_r Q[r] - ds_q is folded into the running block minimum M (no branch, no per-pair state kept)
#define BI_COL(DIAG, jj_)                                                                              \
    {                                                                                                  \
        const float4 c1 = cnext;                                                                       \
        cnext = sc[(jj_) + 2]; /* prefetched one column ahead */                                       \
        float Dn[R + 1];                                                                               \
        column_dists<R, ATT, SHUF ? 2 : 1>(xr2, yr2, xrl, yrl, c1.x, c1.y, Dn);                        \
        float m = TSPB_BIG;                                                                            \
        _Pragma("unroll") for (int r = 0; r < R; ++r) {                                                \
            float qv = UU(r) + Dn[r + 1];                                                              \
            if (DIAG) qv = (qrel0 + (jj_) >= r + 2) ? qv : TSPB_BIG;                                   \
            m = fminf(m, qv);                                                                          \
        }                                                                                              \
        M = fminf(M, m - c0.z);                                                                        \
        _Pragma("unroll") for (int k = 0; k < R / 2; ++k)                                              \
            U2[k] = f2add(f2pack(Dn[2 * k], Dn[2 * k + 1]), cp2[k]);                                   \
        c0 = c1;                                                                                       \
    }

// BI_CB columns, then ONE filter check per warp (ballot); hits are resolved one lane at a time by the whole warp
#define BI_BLOCK(DIAG)                                                                                 \
    for (int jj = 0; jj < NC; jj += BI_CB) {                                                           \
        float M = TSPB_BIG;                                                                            \
        _Pragma("unroll") for (int c = 0; c < BI_CB; ++c) BI_COL(DIAG, jj + c)                         \
        unsigned hits = __ballot_sync(0xffffffffu, M <= thr);                                          \
        while (hits) {                                                                                 \
            const int L = __ffs(hits) - 1;                                                             \
            const float thrL = __shfl_sync(0xffffffffu, thr, L);                                       \
            const MoveKey nb = bi_cold_warp<R, ATT, EXACT32>(A.inst, rec, sc, n, p0 + (L - (tid & 31)) * R, Q0, jj, thrL); \
            if ((tid & 31) == L) {                                                                     \
                atomicAdd(&ctl->cold_calls, 1ull);                                                     \
                if (key_less(nb, best)) {                                                              \
                    best = nb;                                                                         \
                    atomicMin(&s_hint, nb.delta);                                                      \
                    atomicMin(&ctl->hint, nb.delta);                                                   \
                }                                                                                      \
            }                                                                                          \
            /* an exact delta of a real move bounds the minimum for every lane: tighten all, drop stale hits */ \
            if (nb.delta < 0) thr = fminf(thr, (float)nb.delta + W);                                   \
            hits &= hits - 1;                                                                          \
            hits &= __ballot_sync(0xffffffffu, M <= thr);                                              \
        }                                                                                              \
        /* every 64 columns: pick up what the other warps / blocks found (tid 0 swaps in the ctl->hint value it */ \
        /* requested 64 columns ago, so nobody waits for L2)                                                    */ \
        if ((jj & 63) == 64 - BI_CB) {                                                                 \
            if (tid == 0) {                                                                            \
                if (pend_hint < 0) atomicMin(&s_hint, pend_hint);                                      \
                pend_hint = __ldcg(&ctl->hint);                                                        \
            }                                                                                          \
            thr = fminf(thr, (float)(*((volatile int *)&s_hint)) + W);                                 \
        }                                                                                              \
    }

        const int qrel0 = Q0 - p0;  // q - p0 at jj = 0
        if (!diag) {
            BI_INIT(0)
            BI_BLOCK(false)
        } else {
            BI_INIT(0)
            BI_BLOCK(true)
        }
#undef BI_INIT
#undef BI_BLOCK
#undef BI_COL
#undef UU

        __syncthreads();  // every thread is done with this stage before the next prefetch overwrites it
        if (PRUNED && A.timing == 2 && tid == 0) {
            if (it == 0) A.dbg[8192 + 8 * blockIdx.x + 3] = globaltimer_ns();  // first tile scanned
            A.dbg[8192 + 8 * blockIdx.x + 5] = (unsigned long long)(it + 1);
        }
        // publish the best exact delta to the other blocks (fire and forget; theirs arrive through pend_hint above)
        if (tid == 0 && s_hint < 0) atomicMin(&ctl->hint, s_hint);
        P0 = s_tile[buf ^ 1][0];
        Q0 = s_tile[buf ^ 1][1];
        have = s_tile[buf ^ 1][2] != 0;
        NCv = s_tile[buf ^ 1][3];
    }

    // ---- block argmin -> grid argmin ("last block done") -------------------------------------------
    // packed_tail: every block folds its key into ctl->pass_min with ONE 64-bit atomicMin, so the last block only has to
    // read that word (instead of reducing gridDim.x keys: ~3 us of every pass at 1184 blocks).
    best = key_warp_min(best);
    if ((tid & 31) == 0) s_keys[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        MoveKey k = (tid < BI_THREADS / 32) ? s_keys[tid] : key_none();
        k = key_warp_min(k);
        if (tid == 0) {
            A.tour.block_best[blockIdx.x] = k;
            if (A.packed_tail && k.delta < 0) atomicMin(&ctl->pass_min, key_pack(k.delta, k.i, k.j));
            if (PRUNED && scanned) atomicAdd(&ctl->tiles_scanned, (unsigned long long)scanned);
            if (A.timing) {
                const unsigned long long t = globaltimer_ns();
                atomicMin(&ctl->tm_blk_end_min, t);
                if (A.timing == 2) {
                    A.dbg[2 * blockIdx.x + 1] = t;
                    if (PRUNED) A.dbg[8192 + 8 * blockIdx.x + 4] = t;  // block key folded, about to take the ticket
                }
            }
            __threadfence();
            unsigned tk = atomicAdd(&ctl->ticket, 1u);
            s_last = (tk == gridDim.x - 1);
            if (PRUNED && A.timing == 2) A.dbg[8192 + 8 * blockIdx.x + 6] = globaltimer_ns();  // ticket taken
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long t_ticket = 0;
    if (A.timing && tid == 0) {
        t_ticket = globaltimer_ns();
        const unsigned long long t_first = ctl->tm_scan_first, ap0 = ctl->tm_apply_first, ap1 = ctl->tm_apply_end;
        ctl->tm_acc[TM_SCAN] += t_ticket - t_first;
        ctl->tm_acc[TM_SPREAD] += t_ticket - ctl->tm_blk_end_min;
        if (ap1 != 0 && ap0 != ~0ull) {  // the previous pass applied a move
            ctl->tm_acc[TM_GAP] += t_first > ap1 ? t_first - ap1 : 0;
            ctl->tm_acc[TM_APPLY] += ap1 - ap0;
            ctl->tm_acc[TM_APPLY_GAP] += ap0 > ctl->tm_publish ? ap0 - ctl->tm_publish : 0;
        }
        ctl->tm_acc[TM_COUNT] += 1;
        ctl->tm_scan_first = ~0ull;
        ctl->tm_blk_end_min = ~0ull;
        ctl->tm_apply_first = ~0ull;
        ctl->tm_apply_end = 0;
    }

    MoveKey k = key_none();
    int rounds = 1;
    if (A.packed_tail) {
        if (tid == 0) {
            const unsigned long long pk = __ldcg(&ctl->pass_min);
            ctl->pass_min = KEY_PACK_NONE;
            if (pk < KEY_PACK_NONE) {
                key_unpack(pk, &k.delta, &k.i, &k.j);
                k.pad = 0;
            }
            ctl->cand[0] = k;
        }
    } else {
        MoveKey mine = key_none();
        for (int b = tid; b < (int)gridDim.x; b += BI_THREADS) {
            MoveKey o = key_load_cg(&A.tour.block_best[b]);
            if (key_less(o, mine)) mine = o;
        }
        // CTL_NCAND rounds of "block-wide minimum, then retire it": round 0 is the winner of the pass, the rest are
        // runner-ups kept as seeds for the next pass's filter (seed_hint_from_candidates)
        rounds = (A.seed_hint >= 2) ? CTL_NCAND : 1;
#pragma unroll 1
        for (int round = 0; round < rounds; ++round) {
            MoveKey m = key_warp_min(mine);
            if ((tid & 31) == 0) s_keys[tid >> 5] = m;
            __syncthreads();
            m = s_keys[0];
#pragma unroll
            for (int w = 1; w < BI_THREADS / 32; ++w)
                if (key_less(s_keys[w], m)) m = s_keys[w];
            __syncthreads();
            if (round == 0) k = m;
            if (tid == 0) ctl->cand[round] = m;
            if (m.delta < 0 && mine.i == m.i && mine.j == m.j) mine = key_none();
        }
    }

    if (tid == 0) {
        ctl->ncand = rounds;
        ctl->last = k;
        ctl->ticket = 0;
        ctl->tile_next = 0;
        ctl->live_count = 0;
        ctl->hint = 0;
        ctl->launches += 1;
        if (!A.fuse_apply) {
            // multi-GPU: this rank's key for the exchange (peer memory below, or NCCL + the decode kernel)
            ctl->packed = (k.delta < 0) ? key_pack(k.delta, k.i, k.j) : KEY_PACK_NONE;
        } else {
            ctl->passes += 1;
            publish_move(A.tour, k.i, k.j, k.delta);
            if (k.delta >= 0) { ctl->done = 1; ctl->done_reason = DONE_OPTIMUM; }  // reference src/tabusearch.c:158: mindelta >= 0 -> stop
            s_ap[0] = ctl->ap_pa;
            s_ap[1] = ctl->ap_pb;
            s_ap[2] = k.delta < 0;
            if (A.timing) {
                const unsigned long long now = globaltimer_ns();
                ctl->tm_publish = now;
                ctl->tm_acc[TM_TAIL] += now - t_ticket;
            }
        }
    }
    // Multi-GPU exchange over peer memory, fused into this kernel's tail (xchg_min): thread r stores this rank's key word
    // into rank r's slot array (NVLink peer store), then polls this rank's LOCAL slot r until rank r's word of the same
    // epoch has arrived; the block takes the minimum — every rank gets the same winner — and publishes the move for its
    // own replica of the tour.  No collective launch and no tour data on the wire.
    if (!A.fuse_apply && A.xchg.enabled) {
        __shared__ unsigned long long s_xkey[XCHG_MAX_WORLD];
        __shared__ unsigned long long s_mine;
        __shared__ int s_err;
        if (tid == 0) {
            s_mine = ctl->packed;
            s_err = 0;
        }
        __syncthreads();
        unsigned long long wait_ns = 0;
        const unsigned long long win = xchg_min(A.xchg, A.rank, A.world, s_mine, s_xkey, &s_err, A.timing ? &wait_ns : nullptr);
        if (tid == 0) {
            int delta, i, j;
            key_unpack(win, &delta, &i, &j);
            if (s_err) {
                ctl->error = 2;
                ctl->done = 1;
                ctl->done_reason = DONE_OPTIMUM;
                ctl->ap_valid = 0;
            } else {
                ctl->passes += 1;
                publish_move(A.tour, i, j, delta);
                if (delta >= 0) { ctl->done = 1; ctl->done_reason = DONE_OPTIMUM; }
            }
            if (A.timing) {
                const unsigned long long now = globaltimer_ns();
                ctl->tm_publish = now;
                ctl->tm_acc[TM_TAIL] += now - t_ticket;
                ctl->tm_acc[TM_XWAIT] += wait_ns;
            }
        }
    }
    // fuse_apply == 2: this (last) block also applies the move, saving the apply launch — every other block has
    // finished reading rec[] before it took its ticket.  Used for mid-size tours where a launch costs more than the swap.
    if (A.fuse_apply == 2) {
        __syncthreads();  // s_ap[] (thread 0 knows the winner; in packed_tail mode nobody else does)
        if (!s_ap[2]) return;
        apply_swap_range<false>(A.inst, A.tour, s_ap[0], s_ap[1], tid, BI_THREADS);
        if (tid == 0) ctl->ap_valid = 0;
        __threadfence();
        __syncthreads();
        if (A.seed_hint) seed_hint_from_candidates(A.inst, A.tour, tid);
    }
}

// NCCL variant of the multi-GPU exchange: after ncclAllReduce(min) of ctl->packed every rank decodes the same winning key
// and publishes the move for its own replica of the tour.
__global__ void bi_decode_packed_kernel(const TourDev tour) {
    Ctl *ctl = tour.ctl;
    if (ctl->done) { ctl->ap_valid = 0; return; }
    int delta, i, j;
    key_unpack(ctl->packed, &delta, &i, &j);
    ctl->passes += 1;
    publish_move(tour, i, j, delta);
    if (delta >= 0) { ctl->done = 1; ctl->done_reason = DONE_OPTIMUM; }
}

// Grid-wide application of the published move (see apply_swap_range).
// seed: number of block_best[] entries (the scan's grid size) to re-evaluate as seeds of the next pass's filter, 0 = none
// NODE: also keep the node-space view of the first-improvement search current (see apply_swap_range)
// fi_late (first improvement, searches that were not sharded): 1 + parity — nobody has published a move; every thread reads the search's winner
// ctl->fi_sel[parity] = (i << 32) | j and the two positions itself (two round trips, one more than reading a published
// move, instead of the search kernel's whole "last block" tail), global thread 0 — which applies swap 0 and therefore
// holds the four edge lengths of the move — logs it and advances the sweep.
template <bool NODE>
__global__ void __launch_bounds__(256) apply_move_kernel(const InstDev inst, const TourDev tour, int seed, int timing, int fi_late) {
    __shared__ int s_last;
    Ctl *ctl = tour.ctl;
    pdl_launch_dependents();
    pdl_wait();
    // (a sharded search — several GPUs, ctl->fi_mode — has published its move itself: the classic path below)
    if (NODE && fi_late && blockIdx.x == 0 && threadIdx.x == 0) ctl->fi_sel[(fi_late - 1) ^ 1] = FI_NONE;  // the next search's word
    if (NODE && fi_late && *((volatile int *)&ctl->fi_mode[fi_late - 1]) == 0) {
        const int gtid = blockIdx.x * 256 + threadIdx.x;
        const unsigned long long f = *((volatile unsigned long long *)&ctl->fi_sel[fi_late - 1]);
        int done = 0, i0 = 0, j0 = 0;
        if (gtid == 0) {  // only thread 0 may look at these: it is also the one that rewrites them below
            done = *((volatile int *)&ctl->done);
            i0 = *((volatile int *)&ctl->cur_i);
            j0 = *((volatile int *)&ctl->cur_j);
        }
        long long delta = 0;
        int i = 0, j = 0;
        int2 park = make_int2(-1, 0);
        if (f != FI_NONE) {  // (after `done` the searches return at once and the word stays FI_NONE)
            i = (int)(f >> 32);
            j = (int)(f & 0xffffffffull);
            // pos[i] (a, outside the reversed path) is not touched by this launch; pos[j] (b) would be, by swap 0: parked instead
            apply_swap_range<true>(inst, tour, tour.pos[i], tour.pos[j], gtid, gridDim.x * 256, &delta, &park);
        }
        if (gtid == 0 && !done) {
            ctl->fi_pend_node = park.x;
            ctl->fi_pend_pos = park.y;
            if (f != FI_NONE) {
                if (delta >= 0) ctl->error = 1;  // cannot happen: the searching thread saw delta < 0
                ctl->moves += 1;                 // reference heuristics.c:476-486
                ctl->obj_delta += delta;
                const long long lc = ctl->log_count;
                if (tour.log && lc < tour.log_cap) {
                    MoveRec mr;
                    mr.i = i; mr.j = j; mr.delta = delta;
                    tour.log[lc] = mr;
                }
                ctl->log_count = lc + 1;
            }
            ctl->ap_valid = 0;
            fi_advance(tour, f, i0, j0);
        }
        // (This launch's own word is reset by the next apply launch — also by the no-op launches that follow a finished run
        // inside a batch, or the launch after them would find this move again.)
        return;
    }
    if (!ctl->ap_valid) return;
    if (timing && threadIdx.x == 0) atomicMin(&ctl->tm_apply_first, globaltimer_ns());
    apply_swap_range<NODE>(inst, tour, ctl->ap_pa, ctl->ap_pb, blockIdx.x * 256 + threadIdx.x, gridDim.x * 256);
    if (timing) {
        __syncthreads();
        if (threadIdx.x == 0) atomicMax(&ctl->tm_apply_end, globaltimer_ns());
    }
    if (!seed) return;
    // last block done: seed the next pass's filter (BI only).  Every block winner of the pass that just ended is re-evaluated in
    // the tour as it is NOW; the best one that is still a legal move bounds the next pass's minimum, so every block of the
    // next scan starts with a tight threshold instead of re-discovering its own local winner through the exact path
    // (~700 exact-path calls per pass at n = 10 000 without it, a handful with it).
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(&ctl->apply_ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) ctl->apply_ticket = 0;
    int best = 0;
    for (int c = threadIdx.x; c < seed; c += 256) {
        const long long d = legal_move_delta_cg(inst, tour, key_load_cg(&tour.block_best[c]));
        if (d < best) best = (int)d;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, m));
    if ((threadIdx.x & 31) == 0 && best < 0) atomicMin(&ctl->hint, best);
}

// ---- exact tile pruning ---------------------------------------------------------------------------
// For a tile (tile-row I, tile-column J) every pair it holds satisfies
//     delta_exact >= 2 * scale * bd(I, J) - 1 - maxds(I) - maxds(J)
// where bd is the distance between the bounding boxes of the positions of tile-row I (TI positions + the successor of
// the last one) and of tile-column J (TJ positions + successor), maxds the largest edge length ds among them, scale = 1
// (EUC_2D / CEIL_2D; an integer distance is never more than 1/2 below the real one) or 1/sqrt(10) (ATT).  A tile whose bound
// exceeds `hint` — the exact delta of some legal move, i.e. an upper bound of the pass minimum — can hold neither the
// argmin nor a tie (a tie needs delta == minimum <= hint), so skipping it cannot change the selected move: the move log
// stays the reference's bit for bit, only the number of evaluated pairs drops.  The bound is evaluated in FP32 from the
// FP32 records; W (>= 2 + the FP32 / coordinate-rounding error of a distance, see set_instance) plus 1 is subtracted,
// far more than the roundings of the few operations below can add up to.
__global__ void __launch_bounds__(64) tile_boxes_kernel(const InstDev inst, const TourDev tour, int TI, int TJ, int ntr, int ncb,
                                                        int ncb2, int nseed) {
    __shared__ float s_r[2][5];
    Ctl *ctl = tour.ctl;
    pdl_launch_dependents();
    pdl_wait();
    if (*((volatile int *)&ctl->done)) return;
    const int n = tour.n;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    if (b >= ntr + ncb + ncb2) {
        // every block winner of the previous pass that is still a legal move bounds this pass's minimum
        const int c = (b - ntr - ncb - ncb2) * 64 + tid;
        if (c < nseed) {
            const long long d = legal_move_delta_cg(inst, tour, key_load_cg(&tour.block_best[c]));
            if (d < 0) atomicMin(&ctl->hint, (int)d);
        }
        return;
    }
    const bool is_row = b < ntr, is_coarse = b >= ntr + ncb;
    const int start = is_row ? b * TI : (is_coarse ? (b - ntr - ncb) * PRUNE_GROUP * TJ : (b - ntr) * TJ);
    const int len = is_row ? TI : (is_coarse ? PRUNE_GROUP * TJ : TJ);
    const int end = min(start + len, n);  // inclusive: the successor of the last position (position n mirrors position 0)
    float xmin = TSPB_BIG, ymin = TSPB_BIG, xmax = -TSPB_BIG, ymax = -TSPB_BIG, mds = -TSPB_BIG;
    for (int p = start + tid; p <= end; p += 64) {
        const float4 r = __ldcg(&tour.rec[p]);
        xmin = fminf(xmin, r.x); xmax = fmaxf(xmax, r.x);
        ymin = fminf(ymin, r.y); ymax = fmaxf(ymax, r.y);
        if (p < start + len && p < n) mds = fmaxf(mds, r.z);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, m));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, m));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, m));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, m));
        mds = fmaxf(mds, __shfl_xor_sync(0xffffffffu, mds, m));
    }
    if ((tid & 31) == 0) {
        float *o = s_r[tid >> 5];
        o[0] = xmin; o[1] = ymin; o[2] = xmax; o[3] = ymax; o[4] = mds;
    }
    __syncthreads();
    if (tid == 0) {
        const float4 box = make_float4(fminf(s_r[0][0], s_r[1][0]), fminf(s_r[0][1], s_r[1][1]), fmaxf(s_r[0][2], s_r[1][2]),
                                       fmaxf(s_r[0][3], s_r[1][3]));
        const float m = fmaxf(s_r[0][4], s_r[1][4]);
        if (is_row) { tour.rowbox[b] = box; tour.rowmaxds[b] = m; }
        else if (is_coarse) { tour.colbox2[b - ntr - ncb] = box; tour.colmaxds2[b - ntr - ncb] = m; }
        else { tour.colbox[b - ntr] = box; tour.colmaxds[b - ntr] = m; }
    }
}

// Two-level filter, one block per tile-row: first the coarse boxes (PRUNE_GROUP tile-columns each — a superset of their
// tiles' boxes, so a dead group has only dead tiles), then the tiles of the surviving groups.  With a few per cent of the
// tiles alive the second level touches a small fraction of the ~3*10^5 tiles of a 100 000-node tour.  Live tiles of this rank
// (tile ids rank, rank + world, ...) are appended to tour.live / live_lb.
__device__ __forceinline__ float tile_lower_bound(const float4 rb, float rmax, const float4 cb, float cmax, float scale, float W) {
    const float dx = fmaxf(0.f, fmaxf(rb.x - cb.z, cb.x - rb.z));
    const float dy = fmaxf(0.f, fmaxf(rb.y - cb.w, cb.y - rb.w));
    const float bd = sqrtf(fmaf(dy, dy, dx * dx)) * scale;
    return 2.0f * bd - rmax - cmax - W - 1.0f;
}

__global__ void __launch_bounds__(128) tile_filter_kernel(const BiArgs A) {
    extern __shared__ int s_groups[];
    __shared__ int s_ng;
    Ctl *ctl = A.tour.ctl;
    pdl_launch_dependents();
    pdl_wait();
    if (*((volatile int *)&ctl->done)) return;
    const int I = blockIdx.x;
    const int rs = A.tile_row_start[I], cnt = A.tile_row_start[I + 1] - rs, j0 = A.tile_row_j0[I];
    const float hint = (float)__ldcg(&ctl->hint);
    const float4 rb = __ldcg(&A.tour.rowbox[I]);
    const float rmax = __ldcg(&A.tour.rowmaxds[I]);
    const float scale = (A.inst.metric == M_ATT) ? 0.31622773f : 0.99999905f;  // just below 1/sqrt(10) and 1
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_ng = 0;
    __syncthreads();
    const int K0 = j0 / PRUNE_GROUP, K1 = (j0 + cnt - 1) / PRUNE_GROUP;
    for (int K = K0 + tid; K <= K1; K += 128) {
        const float lb2 = tile_lower_bound(rb, rmax, __ldcg(&A.tour.colbox2[K]), __ldcg(&A.tour.colmaxds2[K]), scale, A.inst.W);
        if (!(lb2 > hint)) s_groups[atomicAdd(&s_ng, 1)] = K;
    }
    __syncthreads();
    const int total = s_ng * PRUNE_GROUP;
    for (int w = tid; w < ((total + 31) & ~31); w += 128) {
        bool live = false;
        float lb = 0.f;
        int t = 0;
        if (w < total) {
            const int J = s_groups[w / PRUNE_GROUP] * PRUNE_GROUP + (w % PRUNE_GROUP);
            t = rs + (J - j0);
            if (J >= j0 && J < j0 + cnt && t % A.world == A.rank) {
                lb = tile_lower_bound(rb, rmax, __ldcg(&A.tour.colbox[J]), __ldcg(&A.tour.colmaxds[J]), scale, A.inst.W);
                live = !(lb > hint);
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, live);
        if (mask) {
            unsigned base = 0;
            if (lane == __ffs(mask) - 1) base = atomicAdd(&ctl->live_count, (unsigned)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
            if (live) {
                const unsigned idx = base + __popc(mask & ((1u << lane) - 1u));
                A.tour.live[idx] = make_int2(t, __float_as_int(lb));
            }
        }
    }
}

cudaError_t launch_tile_prune(const BiArgs &a, int TI, int grid_bi, bool pdl, cudaStream_t st) {
    const int ncb = (a.tour.n - 1) / a.TJ + 1;
    const int ncb2 = (ncb + PRUNE_GROUP - 1) / PRUNE_GROUP;
    const int nseed = grid_bi;
    const int gb = a.ntr + ncb + ncb2 + (nseed + 63) / 64;
    cudaError_t e = launch_maybe_pdl(tile_boxes_kernel, dim3(gb), dim3(64), 0, st, pdl, a.inst, a.tour, TI, a.TJ, a.ntr, ncb, ncb2, nseed);
    if (e != cudaSuccess) return e;
    if (a.ntr == 0) return cudaSuccess;
    return launch_maybe_pdl(tile_filter_kernel, dim3((unsigned)a.ntr), dim3(128), sizeof(int) * (size_t)(ncb2 + 2), st, pdl, a);
}

// ---- cross-rank alignment barrier (benchmarks): every rank bumps its counter in every peer's XchgMem and waits until all
// peers' counters reached the same value.  Launched between the L2 flush and the start event of a timed pass, so that the
// ranks enter the pass together whatever their flushes took.
__global__ void rank_align_kernel(const XchgDev X, int rank, int world, Ctl *ctl) {
    __shared__ unsigned s_e;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_e = *X.align_epoch + 1u;
        *X.align_epoch = s_e;
    }
    __syncthreads();
    if (tid < world) {
        const unsigned long long e = s_e;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&X.peer[tid]->align[rank]), "l"(e) : "memory");
        const unsigned long long *src = &X.peer[rank]->align[tid];
        const long long t0 = clock64();
        for (;;) {
            unsigned long long got;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
            if (got >= e) break;
            if (clock64() - t0 > 20000000000ll) {
                ctl->error = 2;
                break;
            }
        }
    }
}

cudaError_t launch_rank_align(const XchgDev &x, int rank, int world, Ctl *ctl, cudaStream_t st) {
    rank_align_kernel<<<1, 32, 0, st>>>(x, rank, world, ctl);
    return cudaGetLastError();
}

// ---- generic exact pass (any metric, incl. GEO / matrix lookup / oversized coordinates) ------------
// Same argmin, every delta evaluated exactly (FP64 on the fly or int32 matrix gather).  One block per
// group of rows; not the throughput path.
//
// MASK = true adds the tabu list of reference src/tabusearch.c:137-149: for every non-adjacent pair the four
// edges (a,b), (a,a1), (b,b1), (a,b1) are tested with check_tenure() (:83-92) under the reference's
// short-circuit order, INCLUDING its side effect — an entry whose tenure has run out is zeroed when (and only
// when) a pair actually tests it.  iter and tenure are constant during one alg_2opt_tabu call, so an entry's
// verdict never depends on whether it was already zeroed: the set of zeroed entries is the union over all
// tested pairs and does not depend on evaluation order, which is what makes the parallel scan exact.
// Zeroed indices are appended to a list so that the host can replay them on the caller's array.
struct TabuArgs {
    int *skip;                 // n(n-1)/2 ints, device copy of the caller's tabu list
    int iter, tenure;
    long long *zl;             // indices zeroed by lazy expiry
    unsigned long long *zl_count;
    long long zl_cap;
};

// reference src/utility.c:17-30 x_udir_pos (evaluated in 64 bits; the reference's int overflows for n > 46341)
__device__ __forceinline__ long long udir_pos(int i, int j, int n) {
    if (i > j) { const int t = i; i = j; j = t; }
    return (long long)i * n + j - ((long long)(i + 1) * (i + 2)) / 2;
}

// reference src/tabusearch.c:83-92 check_tenure
__device__ __forceinline__ bool tabu_check(const TabuArgs &T, long long e) {
    if (T.iter < 0 || T.tenure < 0) return false;
    const int v = *((volatile int *)&T.skip[e]);
    if (v == 0) return false;
    if (T.iter - v > T.tenure) {
        if (atomicExch(&T.skip[e], 0) != 0) {
            const unsigned long long k = atomicAdd(T.zl_count, 1ull);
            if ((long long)k < T.zl_cap) T.zl[k] = e;
        }
        return false;
    }
    return true;
}

template <bool MASK>
__global__ void __launch_bounds__(256) bi_scan_exact_kernel(const InstDev inst, const TourDev tour, int rank, int world,
                                                            int fuse_apply, const TabuArgs tabu) {
    __shared__ MoveKey s_keys[8];
    __shared__ int s_last;
    Ctl *ctl = tour.ctl;
    if (ctl->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ap_valid = 0;
        return;
    }
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
            if (MASK) {
                // a = min(u,v) as the reference enumerates the pair; a1 / b1 are the successors of a / b
                const int a = min(u, v), b = max(u, v);
                const int a1 = (u < v) ? u1 : v1, b1 = (u < v) ? v1 : u1;
                if (tabu_check(tabu, udir_pos(a, b, n)) || tabu_check(tabu, udir_pos(a, a1, n)) ||
                    tabu_check(tabu, udir_pos(b, b1, n)) || tabu_check(tabu, udir_pos(a, b1, n)))
                    continue;
            }
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
    if (tid == 0) {
        ctl->last = k;
        ctl->ticket = 0;
        ctl->launches += 1;
        if (!fuse_apply) {
            ctl->packed = (k.delta < 0) ? key_pack(k.delta, k.i, k.j) : KEY_PACK_NONE;
        } else {
            ctl->passes += 1;
            publish_move(tour, k.i, k.j, k.delta);
            if (k.delta >= 0) { ctl->done = 1; ctl->done_reason = DONE_OPTIMUM; }
        }
    }
}

// ---- host-side launchers -----------------------------------------------------------------------------
template <int T, int R, bool SHUF>
static cudaError_t launch_bi_tr(const BiArgs &a, int grid, bool pdl, cudaStream_t st) {
    const size_t smem = (size_t)2 * (a.TJ + 2) * sizeof(float4) + (size_t)(2 * a.ntr + 2) * sizeof(int);
    const bool att = (a.inst.metric == M_ATT);
    const bool ex = a.inst.exact32 != 0;
    auto go = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {  // opt in to more dynamic shared memory than the default limit (never at the supported tile widths)
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        if (!pdl) {
            kern<<<grid, T, smem, st>>>(a);
            return cudaGetLastError();
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(T);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, a);
    };
    if (a.pruned) {
        if (att && ex) return go(bi_scan_kernel<T, R, true, true, true, SHUF>);
        if (att) return go(bi_scan_kernel<T, R, true, false, true, SHUF>);
        if (ex) return go(bi_scan_kernel<T, R, false, true, true, SHUF>);
        return go(bi_scan_kernel<T, R, false, false, true, SHUF>);
    }
    if (att && ex) return go(bi_scan_kernel<T, R, true, true, false, SHUF>);
    if (att) return go(bi_scan_kernel<T, R, true, false, false, SHUF>);
    if (ex) return go(bi_scan_kernel<T, R, false, true, false, SHUF>);
    return go(bi_scan_kernel<T, R, false, false, false, SHUF>);
}

// rows (moves) of one tile for a block of `threads` threads with `rows_per_thread` rows each
int bi_tile_rows(int threads, int rows_per_thread, int row_shuffle) {
    return row_shuffle ? (threads / 32) * (32 * rows_per_thread - 1) : threads * rows_per_thread;
}
// the shuffle variant is built for the 64-thread shapes (the ones the shape model picks)
bool bi_shuffle_supported(int threads, int rows_per_thread) {
    return threads == 64 && (rows_per_thread == 2 || rows_per_thread == 4 || rows_per_thread == 8);
}

// supported (threads, rows per thread) shapes; anything else is rejected by tspb200_set_option
bool bi_shape_supported(int threads, int rows_per_thread) {
    if (threads == 256) return rows_per_thread == 2 || rows_per_thread == 4 || rows_per_thread == 8 || rows_per_thread == 16;
    if (threads == 128) return rows_per_thread == 4 || rows_per_thread == 8 || rows_per_thread == 16;
    if (threads == 64) return rows_per_thread == 2 || rows_per_thread == 4 || rows_per_thread == 8;
    return false;
}

cudaError_t launch_bi_scan(const BiArgs &a, int threads, int rows_per_thread, int grid, bool pdl, cudaStream_t st) {
    const int R = rows_per_thread;
    if (a.row_shuffle) {
        if (threads == 64 && R == 8) return launch_bi_tr<64, 8, true>(a, grid, pdl, st);
        if (threads == 64 && R == 4) return launch_bi_tr<64, 4, true>(a, grid, pdl, st);
        if (threads == 64 && R == 2) return launch_bi_tr<64, 2, true>(a, grid, pdl, st);
        return cudaErrorInvalidValue;
    }
    if (threads == 256) {
        if (R == 16) return launch_bi_tr<256, 16, false>(a, grid, pdl, st);
        if (R == 8) return launch_bi_tr<256, 8, false>(a, grid, pdl, st);
        if (R == 4) return launch_bi_tr<256, 4, false>(a, grid, pdl, st);
        if (R == 2) return launch_bi_tr<256, 2, false>(a, grid, pdl, st);
    } else if (threads == 128) {
        if (R == 16) return launch_bi_tr<128, 16, false>(a, grid, pdl, st);
        if (R == 8) return launch_bi_tr<128, 8, false>(a, grid, pdl, st);
        if (R == 4) return launch_bi_tr<128, 4, false>(a, grid, pdl, st);
    } else if (threads == 64) {
        if (R == 8) return launch_bi_tr<64, 8, false>(a, grid, pdl, st);
        if (R == 4) return launch_bi_tr<64, 4, false>(a, grid, pdl, st);
        if (R == 2) return launch_bi_tr<64, 2, false>(a, grid, pdl, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_bi_scan_exact(const InstDev &inst, const TourDev &tour, int rank, int world, int fuse_apply,
                                 int grid, cudaStream_t st) {
    bi_scan_exact_kernel<false><<<grid, 256, 0, st>>>(inst, tour, rank, world, fuse_apply, TabuArgs{});
    return cudaGetLastError();
}

cudaError_t launch_bi_scan_tabu(const InstDev &inst, const TourDev &tour, int *skip, int iter, int tenure, long long *zl,
                                unsigned long long *zl_count, long long zl_cap, int grid, cudaStream_t st) {
    TabuArgs t;
    t.skip = skip; t.iter = iter; t.tenure = tenure; t.zl = zl; t.zl_count = zl_count; t.zl_cap = zl_cap;
    bi_scan_exact_kernel<true><<<grid, 256, 0, st>>>(inst, tour, 0, 1, 1, t);
    return cudaGetLastError();
}

cudaError_t launch_bi_decode_packed(const TourDev &tour, cudaStream_t st) {
    bi_decode_packed_kernel<<<1, 1, 0, st>>>(tour);
    return cudaGetLastError();
}

// grid sized for one swap per thread (at most n/2 swaps), capped at 4 blocks per SM
// node_space: the first-improvement runs keep their node-space tables current inside the same launch
cudaError_t launch_apply_move(const InstDev &inst, const TourDev &tour, int num_sms, int seed, int timing, bool pdl,
                              bool node_space, int fi_late, cudaStream_t st) {
    int grid = (tour.n / 2 + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 4 * num_sms) grid = 4 * num_sms;
    if (node_space) return launch_maybe_pdl(apply_move_kernel<true>, dim3(grid), dim3(256), 0, st, pdl, inst, tour, seed, timing, fi_late);
    return launch_maybe_pdl(apply_move_kernel<false>, dim3(grid), dim3(256), 0, st, pdl, inst, tour, seed, timing, 0);
}

}  // namespace tspb
