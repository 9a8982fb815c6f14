/*
 * examples/two_opt_greedy.c — the reference's `-method 2OPT_GREEDY` (HEU_2opt_greedy, src/heuristics.c:572-581: nearest
 * neighbour from node 0, then alg_2opt) written against the engine's C ABI (include/tspb200.h) for a TSPLIB file with a
 * NODE_COORD_SECTION.  Plain C, no CUDA headers needed by the caller.
 *
 *   gcc -O2 -Iinclude examples/two_opt_greedy.c -Ltsp_optimization_b200/lib -ltspb200 \
 *       -Wl,-rpath,$PWD/tsp_optimization_b200/lib -o two_opt_greedy && ./two_opt_greedy data/berlin52.tsp [BI]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tspb200.h"

static int weight_type_of(const char *s) {
    if (strstr(s, "EUC_2D")) return TSPB200_EUC_2D;
    if (strstr(s, "CEIL_2D")) return TSPB200_CEIL_2D;
    if (strstr(s, "ATT")) return TSPB200_ATT;
    if (strstr(s, "GEO")) return TSPB200_GEO;
    if (strstr(s, "MAN_2D")) return TSPB200_MAN_2D;
    if (strstr(s, "MAX_2D")) return TSPB200_MAX_2D;
    return TSPB200_EUC_2D; /* the reference falls through to EUC_2D (src/distutil.c:91) */
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.tsp [BI]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    char line[512];
    int n = 0, wt = TSPB200_EUC_2D, in_coords = 0, got = 0;
    double *xy = NULL;
    while (fgets(line, sizeof line, f)) {
        if (!strncmp(line, "DIMENSION", 9)) { n = atoi(strchr(line, ':') ? strchr(line, ':') + 1 : line + 9); xy = calloc(2 * (size_t)n, sizeof *xy); }
        else if (!strncmp(line, "EDGE_WEIGHT_TYPE", 16)) wt = weight_type_of(line);
        else if (!strncmp(line, "NODE_COORD_SECTION", 18)) in_coords = 1;
        else if (!strncmp(line, "EOF", 3)) break;
        else if (in_coords && xy) {
            int id; double x, y;
            if (sscanf(line, "%d %lf %lf", &id, &x, &y) == 3 && id >= 1 && id <= n) { xy[2 * (id - 1)] = x; xy[2 * (id - 1) + 1] = y; got++; }
        }
    }
    fclose(f);
    if (n < 1 || got != n) { fprintf(stderr, "could not read %d coordinates\n", n); return 2; }

    tspb200_ctx *ctx = NULL;
    if (tspb200_create(0, &ctx)) { fprintf(stderr, "[ERROR] %s\n", tspb200_last_error(ctx)); return 1; } /* no CPU fallback */
    int rc = tspb200_set_instance(ctx, xy, n, wt);
    if (!rc && wt == TSPB200_GEO) rc = tspb200_dist_matrix_build(ctx, NULL); /* GEO 2-opt gathers from the resident matrix */
    int32_t *succ = malloc(sizeof *succ * (size_t)n);
    double cost = 0;
    if (!rc) rc = tspb200_nn_tour(ctx, 0, succ, &cost);                       /* greedy(inst, 0) */
    printf("nearest neighbour: %.0f\n", cost);
    tspb200_stats st;
    const int mode = (argc > 2 && !strcmp(argv[2], "BI")) ? TSPB200_BI : TSPB200_FI;
    if (!rc) rc = tspb200_two_opt(ctx, mode, succ, &cost, -1, &st, NULL, 0, NULL); /* alg_2opt / alg_2opt_tabu(NULL) */
    if (rc) { fprintf(stderr, "[ERROR] %s\n", tspb200_last_error(ctx)); return 1; }
    printf("2-opt (%s): %.0f after %lld moves, %lld %s, %.3f ms on the device\n", mode == TSPB200_BI ? "best improvement" : "first improvement",
           cost, (long long)st.moves, (long long)st.passes, mode == TSPB200_BI ? "passes" : "sweeps", st.gpu_ms);
    tspb200_destroy(ctx);
    free(succ);
    free(xy);
    return 0;
}
