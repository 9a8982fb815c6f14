/*
 * tspb200_dropin.h — the reference-named entry points, exported by libtspb200_dropin.so.
 *
 * These four symbols have EXACTLY the reference's signatures (x86-64 SysV, C11) so that the reference's own
 * objects link against them unchanged (recipe in INTEGRATION.md):
 *
 *   double calc_dist(int i, int j, instance *inst);                       reference include/distutil.h:137
 *   int    alg_2opt(instance *inst);                                      reference include/heuristics.h:82
 *   int    alg_2opt_tabu(instance *inst, int *skip_edge, int *stored_prev,
 *                        const int iter, const int tenure);               reference src/tabusearch.c:107
 *   void   reverse_path(instance *inst, int start_node, int end_node,
 *                       int *prev);                                       reference include/utility.h:334
 *
 * `tspb200_ref_instance` is an ABI mirror of the reference's `instance` (include/utility.h:105-160), written
 * from its layout (sizeof 152; offsets verified against the compiled reference by tests/test_abi.py);
 * the reference's own headers are never included by the product.
 */
#ifndef TSPB200_DROPIN_H
#define TSPB200_DROPIN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double x, y; } tspb200_ref_point;  /* include/utility.h:126-129 */
typedef struct { int i, j; } tspb200_ref_edge;      /* include/utility.h:131-134: directed edge i -> j */

typedef struct {              /* sol_method, include/utility.h:105-110 */
    int id;
    int edge_type;
    char *name;
    int use_cplex;
} tspb200_ref_method;

typedef struct {              /* instance_params, include/utility.h:113-123 */
    char *file_path;
    int num_threads;
    int time_limit;           /* seconds; <= 0 means unlimited (heuristics.c:458) */
    tspb200_ref_method method;
    int verbose;
    int integer_cost;         /* 1 = integer costs (the only mode with a bit-exact contract) */
    int seed;
    int perf_prof;
    int callback_2opt;
} tspb200_ref_params;

typedef struct {              /* solution, include/utility.h:136-141 */
    double obj_best;
    tspb200_ref_edge *edges;  /* edges[k].i == k, edges[k].j == succ(k) */
    double time_to_solve;
    double *xbest;
} tspb200_ref_solution;

typedef struct {              /* instance, include/utility.h:144-160 */
    tspb200_ref_params params;
    char *name;
    char *comment;
    tspb200_ref_point *nodes;
    int num_nodes;
    int weight_type;          /* enum weight_type: EUC_2D=0 MAX_2D=1 MAN_2D=2 CEIL_2D=3 GEO=4 ATT=5 */
    long num_columns;
    int *ind;
    unsigned int *thread_seeds;
    tspb200_ref_solution solution;
} tspb200_ref_instance;

#if defined(__x86_64__) && (defined(__STDC_VERSION__) || defined(__cplusplus))
#ifdef __cplusplus
#define TSPB200_SA(c, m) static_assert(c, m)
#else
#define TSPB200_SA(c, m) _Static_assert(c, m)
#endif
TSPB200_SA(sizeof(tspb200_ref_instance) == 152, "instance size");
TSPB200_SA(offsetof(tspb200_ref_instance, params.time_limit) == 12, "time_limit");
TSPB200_SA(offsetof(tspb200_ref_instance, params.integer_cost) == 44, "integer_cost");
TSPB200_SA(offsetof(tspb200_ref_instance, params.perf_prof) == 52, "perf_prof");
TSPB200_SA(offsetof(tspb200_ref_instance, nodes) == 80, "nodes");
TSPB200_SA(offsetof(tspb200_ref_instance, num_nodes) == 88, "num_nodes");
TSPB200_SA(offsetof(tspb200_ref_instance, weight_type) == 92, "weight_type");
TSPB200_SA(offsetof(tspb200_ref_instance, num_columns) == 96, "num_columns");
TSPB200_SA(offsetof(tspb200_ref_instance, solution.obj_best) == 120, "obj_best");
TSPB200_SA(offsetof(tspb200_ref_instance, solution.edges) == 128, "edges");
#endif

double calc_dist(int i, int j, tspb200_ref_instance *inst);
int alg_2opt(tspb200_ref_instance *inst);
int alg_2opt_tabu(tspb200_ref_instance *inst, int *skip_edge, int *stored_prev, const int iter, const int tenure);
void reverse_path(tspb200_ref_instance *inst, int start_node, int end_node, int *prev);

/* Optional, same signatures as the reference (include/heuristics.h:16,32,40): the constructive callers next to the path.
 * They only take effect when the integrator weakens the reference's definitions of these symbols as well. */
int greedy(tspb200_ref_instance *inst, int starting_node);
int HEU_Greedy_iter(tspb200_ref_instance *inst);
int HEU_extramileage(tspb200_ref_instance *inst);

/* sizeof / offsets of the mirror, same order as oracle/ref_shim.c:refshim_layout (for the ABI test). */
int tspb200_dropin_layout(long long *out, int cap);
/* Releases the cached device contexts (optional; they are also released at process exit). */
void tspb200_dropin_reset(void);
/* Number of instances currently resident on the device (the drop-in keeps up to four, least recently used first to go). */
int tspb200_dropin_resident(void);

#ifdef __cplusplus
}
#endif
#endif
