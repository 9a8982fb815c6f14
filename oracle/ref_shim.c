/*
 * ref_shim.c — thin harness linked INTO oracle/_ref/libtspref.so next to the unmodified reference
 * objects.  TEST INFRASTRUCTURE ONLY.  It includes the reference's own headers, so every struct it
 * touches has the reference's real layout; it re-creates the few lines of set-up that
 * src/solver.c:264-270 (TSP_heuc) does before calling a heuristic, and exposes helpers that the
 * reference has no API for (layout report, all-pairs calc_dist, a multi-threaded partial BI scan
 * that calls the reference's own calc_dist and x_udir_pos per pair, like src/tabusearch.c:126-156, for the CPU baseline).
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utility.h"
#include "distutil.h"
#include "heuristics.h"

int alg_2opt_tabu(instance *inst, int *skip_edge, int *stored_prev, const int iter, const int tenure);

/* sizeof / offsetof of everything the drop-in boundary reads. */
int refshim_layout(int64_t *out, int cap) {
    int64_t v[] = {
        (int64_t)sizeof(instance),
        (int64_t)offsetof(instance, params.time_limit),
        (int64_t)offsetof(instance, params.integer_cost),
        (int64_t)offsetof(instance, params.perf_prof),
        (int64_t)offsetof(instance, nodes),
        (int64_t)offsetof(instance, num_nodes),
        (int64_t)offsetof(instance, weight_type),
        (int64_t)offsetof(instance, num_columns),
        (int64_t)offsetof(instance, solution.obj_best),
        (int64_t)offsetof(instance, solution.edges),
        (int64_t)sizeof(point),
        (int64_t)sizeof(edge),
        (int64_t)offsetof(instance, params.verbose),
        (int64_t)offsetof(instance, params.seed),
    };
    int k = (int)(sizeof v / sizeof v[0]);
    for (int i = 0; i < k && i < cap; i++) out[i] = v[i];
    return k;
}

static void setup_like_tsp_heuc(instance *inst) {
    /* src/solver.c:268-270 */
    inst->num_columns = (long)inst->num_nodes * (inst->num_nodes - 1) / 2;
    inst->solution.edges = (edge *)calloc((size_t)inst->num_nodes, sizeof(edge));
}

instance *refshim_new(int n, const double *xy, int wt) {
    instance *inst = (instance *)calloc(1, sizeof(instance));
    inst->params.integer_cost = 1;
    inst->params.perf_prof = 1; /* keeps plot_solution/export_tour silent */
    inst->params.time_limit = 0; /* <=0 means unlimited, heuristics.c:458 */
    inst->params.seed = -1;
    inst->num_nodes = n;
    inst->weight_type = (weight_type)wt;
    inst->nodes = (point *)calloc((size_t)n, sizeof(point));
    for (int k = 0; k < n; k++) { inst->nodes[k].x = xy[2 * k]; inst->nodes[k].y = xy[2 * k + 1]; }
    setup_like_tsp_heuc(inst);
    for (int k = 0; k < n; k++) { inst->solution.edges[k].i = k; inst->solution.edges[k].j = (k + 1) % n; }
    return inst;
}

/* Runs the reference's own TSPLIB parser (src/utility.c:351-453); the path is copied with a NUL
 * terminator, side-stepping the unterminated copy in parse_comand_line (utility.c:79-80). */
instance *refshim_parse(const char *path) {
    instance *inst = (instance *)calloc(1, sizeof(instance));
    inst->params.integer_cost = 1;
    inst->params.perf_prof = 1;
    inst->params.seed = -1;
    inst->params.file_path = strdup(path);
    parse_instance(inst);
    setup_like_tsp_heuc(inst);
    return inst;
}

void refshim_free(instance *inst) {
    if (!inst) return;
    free(inst->nodes);
    free(inst->solution.edges);
    free(inst->params.file_path);
    free(inst->name);
    free(inst);
}

int refshim_num_nodes(const instance *inst) { return inst->num_nodes; }
int refshim_weight_type(const instance *inst) { return (int)inst->weight_type; }
double refshim_obj(const instance *inst) { return inst->solution.obj_best; }
void refshim_set_obj(instance *inst, double v) { inst->solution.obj_best = v; }
void refshim_set_time_limit(instance *inst, int s) { inst->params.time_limit = s; }

void refshim_get_xy(const instance *inst, double *xy) {
    for (int k = 0; k < inst->num_nodes; k++) { xy[2 * k] = inst->nodes[k].x; xy[2 * k + 1] = inst->nodes[k].y; }
}
void refshim_get_succ(const instance *inst, int32_t *succ) {
    for (int k = 0; k < inst->num_nodes; k++) succ[k] = inst->solution.edges[k].j;
}
void refshim_set_succ(instance *inst, const int32_t *succ) {
    for (int k = 0; k < inst->num_nodes; k++) { inst->solution.edges[k].i = k; inst->solution.edges[k].j = succ[k]; }
}

void refshim_dist_matrix(instance *inst, int32_t *out) {
    int n = inst->num_nodes;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[(size_t)i * n + j] = (int32_t)calc_dist(i, j, inst);
}

/* ---- CPU baseline: rows of one best-improvement scan with the reference's calc_dist ----------- */
typedef struct {
    instance *inst; int row_begin, row_end, tid, nthreads;
    int64_t evals; double best; int bi, bj;
} ref_job;

static void *ref_scan_worker(void *arg) {
    ref_job *jb = (ref_job *)arg;
    instance *inst = jb->inst;
    int n = inst->num_nodes;
    const edge *ed = inst->solution.edges;
    int *volatile skip_edge_v = NULL;
    int *skip_edge = skip_edge_v; /* opaque to the optimiser: the index computations stay, like in the stock build */
    jb->evals = 0; jb->best = 0; jb->bi = 0; jb->bj = 0;
    for (int r0 = jb->row_begin + jb->tid * 16; r0 < jb->row_end; r0 += jb->nthreads * 16) {
        int r1 = r0 + 16 < jb->row_end ? r0 + 16 : jb->row_end;
        for (int a = r0; a < r1; a++) {
            for (int b = a + 1; b < n; b++) {
                int a1 = ed[a].j, b1 = ed[b].j;
                if (b == a1 || b1 == a) continue;
                jb->evals++;
                /* src/tabusearch.c:137-149: the stock loop computes the four tabu-list indices with the reference's own
                 * x_udir_pos() for every pair and tests `skip_edge &&` (NULL here) before the four calc_dist calls; kept
                 * so that this harness does per pair exactly what alg_2opt_tabu(inst, NULL, NULL, 1, 1) does */
                int e1 = x_udir_pos(a, b, n), e2 = x_udir_pos(a, a1, n), e3 = x_udir_pos(b, b1, n), e4 = x_udir_pos(a, b1, n);
                if (skip_edge && (skip_edge[e1] | skip_edge[e2] | skip_edge[e3] | skip_edge[e4])) continue;
                double delta = calc_dist(a, b, inst) + calc_dist(a1, b1, inst) - calc_dist(a, a1, inst) - calc_dist(b, b1, inst);
                if (delta < jb->best || (delta == jb->best && delta < 0 && (a < jb->bi || (a == jb->bi && b < jb->bj)))) {
                    jb->best = delta; jb->bi = a; jb->bj = b;
                }
            }
        }
    }
    return NULL;
}

int64_t refshim_bi_scan_rows_mt(instance *inst, int row_begin, int row_end, int threads, double *seconds,
                                int64_t *best_delta, int32_t *best_i, int32_t *best_j) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (row_end > inst->num_nodes - 1) row_end = inst->num_nodes - 1;
    pthread_t th[256];
    ref_job jobs[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        ref_job jb = {inst, row_begin, row_end, t, threads, 0, 0, 0, 0};
        jobs[t] = jb;
        pthread_create(&th[t], NULL, ref_scan_worker, &jobs[t]);
    }
    int64_t evals = 0; double bd = 0; int bi = 0, bj = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        evals += jobs[t].evals;
        ref_job *jb = &jobs[t];
        if (jb->best < bd || (jb->best == bd && bd < 0 && (jb->bi < bi || (jb->bi == bi && jb->bj < bj)))) {
            bd = jb->best; bi = jb->bi; bj = jb->bj;
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    *best_delta = (int64_t)bd; *best_i = bi; *best_j = bj;
    return evals;
}
