/*
 * Stub <cplex.h> used ONLY to compile the unmodified reference's CPLEX-free sources into
 * oracle/_ref/libtspref.so (test infrastructure).  The reference's include/utility.h:8 includes
 * <cplex.h> and src/utility.c:662-699 names two CPLEX calls that the hot path never reaches.
 */
#ifndef ORACLE_STUB_CPLEX_H
#define ORACLE_STUB_CPLEX_H
/* the real cplex.h pulls these in; several reference sources rely on that */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
typedef struct orc_stub_cpxenv *CPXENVptr;
typedef struct orc_stub_cpxlp *CPXLPptr;
static inline int CPXwriteprob(CPXENVptr e, CPXLPptr l, const char *f, const char *t) { (void)e; (void)l; (void)f; (void)t; return 0; }
static inline int CPXsetlogfilename(CPXENVptr e, const char *f, const char *m) { (void)e; (void)f; (void)m; return 0; }
#endif
