/*
 * admm_b200.c -- drop-in replacement for the reference's admm() with the
 * reference's exact solver signature (qp_solvers/qp_solvers.h:14-16):
 *
 *     struct _matrix *admm_b200(struct _matrix *x0, unsigned iterations,
 *                               struct _quadratic_form *qf);
 *
 * This file is compiled INSIDE the reference's build (it includes the
 * reference's headers and uses its pool allocator for the returned vector, so
 * that the caller's matrix_free(x) at test/test.c:105 keeps working) and is
 * linked against libqpb200.so.  It can be handed to the reference's solver
 * function pointer unchanged:
 *
 *     test_optimizer(qf, ADMM_ITERATIONS, x0, "admm_b200", admm_b200);
 *
 * Like the reference, x0 is accepted and ignored (qp_solvers.c:256), the box
 * is the compile-time scalar pair ADMM_BOX_CONSTRAINT_{MIN,MAX}
 * (config.h:29-30) and the result is a freshly pool-allocated N x 1 matrix.
 * On failure it returns 0 like the reference's allocators do; it never exits.
 */
#include <stdio.h>

#include "matrix_ops.h"
#include "qp.h"
#include "qp_solvers.h"

#include "qp_batch.h"

struct _matrix *admm_b200(struct _matrix *x0, unsigned iterations,
			  struct _quadratic_form *qf)
{
	(void)x0;
	struct _matrix *x = matrix_alloc(Nx1);
	if (!x)
		return 0;
	/* struct qp_ref_quadratic_form / qp_ref_matrix are layout-identical to
	 * the reference structs (checked by tests/test_oracle_vs_ref.py) */
	int rc = qp_admm_qf((const struct qp_ref_quadratic_form *)qf, iterations,
			    ADMM_BOX_CONSTRAINT_MIN, ADMM_BOX_CONSTRAINT_MAX,
			    x->elements, 0, 0);
	if (rc != QP_OK) {
		fprintf(stderr, "admm_b200: qp_admm_qf failed (%d) %s\n", rc,
			qp_last_error());
		matrix_free(x);
		return 0;
	}
	return x;
}

/*
 * The same in the reference's own arithmetic: binary64 on the GPU
 * (qp_options.precision = QP_PRECISION_F64, csrc/qp_f64.cuh).  Reproduces the
 * reference's iteration count -- hence its x to ~1e-10 -- on its ill-conditioned
 * test problems, where the binary32 product path stops an iteration early or
 * late on one problem in three:
 *
 *     test_optimizer(qf, ADMM_ITERATIONS, x0, "admm_b200_f64", admm_b200_f64);
 */
struct _matrix *admm_b200_f64(struct _matrix *x0, unsigned iterations,
			      struct _quadratic_form *qf)
{
	(void)x0;
	struct _matrix *x = matrix_alloc(Nx1);
	if (!x)
		return 0;
	qp_options opt;
	qp_options_default(&opt);
	opt.precision = QP_PRECISION_F64;
	const struct qp_ref_quadratic_form *one[1] = { (const struct qp_ref_quadratic_form *)qf };
	double *xo[1] = { x->elements };
	int rc = qp_admm_qf_batch(one, 1, iterations, ADMM_BOX_CONSTRAINT_MIN,
				  ADMM_BOX_CONSTRAINT_MAX, xo, 0, 0, &opt);
	if (rc != QP_OK) {
		fprintf(stderr, "admm_b200_f64: qp_admm_qf_batch failed (%d) %s\n", rc,
			qp_last_error());
		matrix_free(x);
		return 0;
	}
	return x;
}

/*
 * Drop-in for newton_method_with_line_search() (qp_solvers/qp_solvers.h:9-12,
 * qp_solvers.c:103-144), same signature, same ownership of the result:
 *
 *     test_optimizer(qf, HESS_ITERATIONS, x0, "newton_b200", newton_b200);
 */
struct _matrix *newton_b200(struct _matrix *x0, unsigned iterations,
			    struct _quadratic_form *qf)
{
	struct _matrix *x = matrix_alloc(Nx1);
	if (!x)
		return 0;
	int rc = qp_newton_qf((const struct qp_ref_quadratic_form *)qf,
			      x0 ? x0->elements : 0, iterations, x->elements, 0, 0);
	if (rc != QP_OK) {
		fprintf(stderr, "newton_b200: qp_newton_qf failed (%d) %s\n", rc,
			qp_last_error());
		matrix_free(x);
		return 0;
	}
	return x;
}

/*
 * Drop-in for gradient_descent_with_line_search() (qp_solvers/qp_solvers.h:4-7,
 * qp_solvers.c:65-101), same signature, same ownership of the result:
 *
 *     test_optimizer(qf, GRAD_ITERATIONS, x0, "gradient_b200", gradient_b200);
 */
struct _matrix *gradient_b200(struct _matrix *x0, unsigned iterations,
			      struct _quadratic_form *qf)
{
	struct _matrix *x = matrix_alloc(Nx1);
	if (!x)
		return 0;
	int rc = qp_gradient_qf((const struct qp_ref_quadratic_form *)qf,
				x0 ? x0->elements : 0, iterations, x->elements, 0, 0);
	if (rc != QP_OK) {
		fprintf(stderr, "gradient_b200: qp_gradient_qf failed (%d) %s\n", rc,
			qp_last_error());
		matrix_free(x);
		return 0;
	}
	return x;
}
