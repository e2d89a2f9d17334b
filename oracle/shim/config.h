/*
 * oracle/shim/config.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Build-time configuration used when the UNMODIFIED reference sources under
 * /root/reference are compiled into oracle/_ref/ (see oracle/Makefile).
 *
 * The reference selects its configuration by `#include "config.h"` from
 * matrix/matrix_type/matrix_type.h:4-8 and test/test.h:8-12; that header is
 * looked up on the include path, so putting this directory first on -I
 * replaces /root/reference/config.h without touching a reference file.
 *
 * What is configured here, and why it differs from the reference defaults:
 *   N_DIM      comes from the compiler command line (-DN_DIM=<n>U): the
 *              reference fixes the problem dimension at compile time, so the
 *              oracle is built once per shape.
 *   box bounds the reference broadcasts two scalar macros to every
 *              coordinate inside `for (unsigned i ...)` at
 *              qp_solvers/qp_solvers.c:277-280.  Here the two macros expand
 *              to an element of a driver-owned array indexed by that loop
 *              variable, so the same unmodified loop installs a
 *              per-coordinate box chosen at run time (a uniform box is the
 *              special case where every element is equal).
 */
#ifndef QPREF_SHIM_CONFIG_H
#define QPREF_SHIM_CONFIG_H

#ifndef N_DIM
#error "build with -DN_DIM=<n>U (see oracle/Makefile)"
#endif

/* static pool sizes (kmalloc/kmalloc.c:37-39).  admm() needs 8 vectors plus
 * the caller's q and x0, and 3 matrices (P, R, scratch in matrix_invert). */
#define Nx1_MAX 10
#define NxN_MAX 5

/* run-time box, owned by oracle/shim/qpref_driver.c */
extern double qpref_box_hi[N_DIM];
extern double qpref_box_lo[N_DIM];
#define ADMM_BOX_CONSTRAINT_MAX (qpref_box_hi[i])
#define ADMM_BOX_CONSTRAINT_MIN (qpref_box_lo[i])

/* knobs read only by test/test.c (compiled in so that the reference's own
 * inversion_test / scalar_prod_test can be run from the driver) */
#define NUM_INVERSION_TEST_RUNS 8
#define INVERSION_TEST_PRECISION 1e-6
#define P_RAND_ENTRY_MAX 1e3
#define P_RAND_ENTRY_MIN -1e3
#define PYTHON_COMMAND "python"

#endif
