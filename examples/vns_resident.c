/*
 * examples/vns_resident.c — the loop of the reference's HEU_VNS (src/vns.c:103-183) written against the engine's resident
 * sessions (include/tspb200.h): the tour lives in HBM across kick -> alg_2opt -> keep / restore; per iteration three
 * indices go down and one cost comes back.  The random numbers are drawn on the host exactly like the reference does
 * (URAND = random() / RAND_MAX, include/utility.h:36; rand_choice, src/utility.c:752-754; the index loops of
 * src/vns.c:25-31), so with the same seed the run visits the same tours as the reference's.  Plain C.
 *
 *   gcc -O2 -Iinclude examples/vns_resident.c -Ltsp_optimization_b200/lib -ltspb200 \
 *       -Wl,-rpath,$PWD/tsp_optimization_b200/lib -o vns_resident && ./vns_resident 2000 50 123
 */
#include <stdio.h>
#include <stdlib.h>

#include "tspb200.h"

static int rand_choice(int from, int to) { return from + (int)(((double)random() / RAND_MAX) * (to - from)); }

#define CHECK(call)                                                                  \
    do {                                                                             \
        if (call) { fprintf(stderr, "[ERROR] %s\n", tspb200_last_error(ctx)); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2000, iters = argc > 2 ? atoi(argv[2]) : 50;
    const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 123u;
    if (n < 8) { fprintf(stderr, "usage: %s n iterations seed\n", argv[0]); return 2; }
    double *xy = malloc(sizeof *xy * 2 * (size_t)n);
    srandom(n); /* some instance: uniform integer coordinates like other_codes/generate_tsp_istances.py */
    for (int k = 0; k < 2 * n; k++) xy[k] = (double)(random() % 10001);

    tspb200_ctx *ctx = NULL;
    if (tspb200_create(0, &ctx)) { fprintf(stderr, "[ERROR] %s\n", tspb200_last_error(ctx)); return 1; } /* no CPU fallback */
    CHECK(tspb200_set_instance(ctx, xy, n, TSPB200_EUC_2D));
    int32_t *succ = malloc(sizeof *succ * (size_t)n);
    double cost = 0, best = 0;
    tspb200_stats st;
    CHECK(tspb200_nn_tour(ctx, 0, succ, &cost));      /* initial solution (the reference uses HEU_2opt_greedy_iter) */
    CHECK(tspb200_tour_upload(ctx, succ, 0));
    CHECK(tspb200_fi_run(ctx, -1, &st));              /* alg_2opt */
    CHECK(tspb200_tour_cost(ctx, &best));
    CHECK(tspb200_tour_save(ctx, 0));                 /* best_sol = current */
    printf("initial 2-opt tour: %.0f\n", best);

    srandom(seed);
    for (int it = 0; it < iters; it++) {
        int i1 = rand_choice(0, n), i2 = i1, i3 = i1; /* src/vns.c:25-31 */
        while (i2 == i1 || abs(i1 - i2) <= 1) i2 = rand_choice(0, n);
        while (i3 == i1 || i3 == i2 || abs(i1 - i3) <= 1 || abs(i2 - i3) <= 1) i3 = rand_choice(0, n);
        CHECK(tspb200_vns_kick(ctx, i1, i2, i3, &cost));  /* kick(inst) */
        CHECK(tspb200_fi_run(ctx, -1, &st));              /* alg_2opt(inst) */
        cost += (double)st.obj_delta;
        if (cost < best) {
            best = cost;
            CHECK(tspb200_tour_save(ctx, 0));
            printf("iteration %d: new incumbent %.0f (%lld moves)\n", it, best, (long long)st.moves);
        } else {
            CHECK(tspb200_tour_restore(ctx, 0));
        }
    }
    CHECK(tspb200_tour_restore(ctx, 0));
    CHECK(tspb200_tour_download(ctx, succ, &cost));
    printf("best tour after %d kicks: %.0f\n", iters, cost);
    tspb200_destroy(ctx);
    free(succ);
    free(xy);
    return cost == best ? 0 : 1;
}
