/* pico_oracle.c — see pico_oracle.h. TEST INFRASTRUCTURE ONLY.
 * Build: gcc -O3 -DNDEBUG -ffp-contract=off -fopenmp -shared -fPIC (no
 * -march=native / -ffast-math: FMA contraction changes distance bits,
 * SURVEY.md §8c). */
#include "pico_oracle.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_T float
#define PO_SFX f32
#define PO_MAXVAL FLT_MAX
#include "pico_oracle_impl.inc"
#undef PO_T
#undef PO_SFX
#undef PO_MAXVAL

#define PO_T double
#define PO_SFX f64
#define PO_MAXVAL DBL_MAX
#include "pico_oracle_impl.inc"
#undef PO_T
#undef PO_SFX
#undef PO_MAXVAL

void po_free_buffer(void* p) { free(p); }

int po_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
