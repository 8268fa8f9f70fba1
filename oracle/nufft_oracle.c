/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see nufft_oracle_impl.h for scope, pinning and citations).
 * Instantiates the CPU restatement for Float32 and Float64.
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared; no fast-math, no FMA contraction)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF(x) x##_f32
#include "nufft_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) x##_f64
#include "nufft_oracle_impl.h"
#undef REAL
#undef SUF

#ifdef _OPENMP
#include <omp.h>
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int orc_num_threads(void) { return 1; }
void orc_set_num_threads(int n) { (void)n; }
#endif
