/*
 * oracle/shim/compat_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Built with the reference's OWN config.h (N_DIM = 48, box +-1e12) and its
 * unmodified sources plus embedded_qp_solver_b200/csrc/compat/admm_b200.c into
 * oracle/_ref/libqpcompat.so.  It does what the reference's main.c:36-55 does
 * for the ADMM test -- same generators, same test_optimizer() call through the
 * solver function pointer (test/test.c:89-106) -- once with the reference's
 * admm and once with the drop-in admm_b200, and hands both results back.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "test.h"

struct _matrix *admm_b200(struct _matrix *x0, unsigned iterations,
			  struct _quadratic_form *qf);

struct _matrix *admm_b200_f64(struct _matrix *x0, unsigned iterations,
			      struct _quadratic_form *qf);
struct _matrix *newton_b200(struct _matrix *x0, unsigned iterations,
			    struct _quadratic_form *qf);
struct _matrix *gradient_b200(struct _matrix *x0, unsigned iterations,
			      struct _quadratic_form *qf);

static int ready = 0;

unsigned compat_dim(void)
{
	return N_DIM;
}

/* returns 0 on success; x_ref/x_gpu get N_DIM doubles, f_out[0..1] the
 * objective values quadratic_form_eval() gives for them */
int compat_run(unsigned seed, double *x_ref, double *x_gpu, double *f_out,
	       double *P_out, double *q_out)
{
	if (!ready) {
		kmalloc_init();
		ready = 1;
	}
	srand(seed);
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *q = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, q, 0);
	if (!p || !q || !x0 || !qf)
		return -1;
	matirx_random_pos_def(p, P_RAND_ENTRY_MIN, P_RAND_ENTRY_MAX);
	matrix_random(q, Q_RAND_ENTRY_MIN, Q_RAND_ENTRY_MAX);
	matrix_random(x0, X_RAND_ENTRY_MIN, X_RAND_ENTRY_MAX);

	/* the reference's own harness, through the solver function pointer */
	test_optimizer(qf, ADMM_ITERATIONS, x0, "admm", admm);
	test_optimizer(qf, ADMM_ITERATIONS, x0, "admm_b200", admm_b200);

	struct _matrix *a = admm(x0, ADMM_ITERATIONS, qf);
	struct _matrix *b = admm_b200(x0, ADMM_ITERATIONS, qf);
	int rc = 0;
	if (!a || !b) {
		rc = -2;
	} else {
		memcpy(x_ref, a->elements, sizeof(double) * N_DIM);
		memcpy(x_gpu, b->elements, sizeof(double) * N_DIM);
		f_out[0] = quadratic_form_eval(qf, a);
		f_out[1] = quadratic_form_eval(qf, b);
		memcpy(P_out, p->elements, sizeof(double) * N_DIM * N_DIM);
		memcpy(q_out, q->elements, sizeof(double) * N_DIM);
	}
	if (a)
		matrix_free(a);
	if (b)
		matrix_free(b);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(q);
	matrix_free(p);
	return rc;
}

/* the binary64 drop-in against the reference's admm on the reference's own problem */
int compat_run_f64(unsigned seed, double *x_ref, double *x_gpu)
{
	if (!ready) {
		kmalloc_init();
		ready = 1;
	}
	srand(seed);
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *q = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, q, 0);
	if (!p || !q || !x0 || !qf)
		return -1;
	matirx_random_pos_def(p, P_RAND_ENTRY_MIN, P_RAND_ENTRY_MAX);
	matrix_random(q, Q_RAND_ENTRY_MIN, Q_RAND_ENTRY_MAX);
	matrix_random(x0, X_RAND_ENTRY_MIN, X_RAND_ENTRY_MAX);
	test_optimizer(qf, ADMM_ITERATIONS, x0, "admm_b200_f64", admm_b200_f64);
	struct _matrix *a = admm(x0, ADMM_ITERATIONS, qf);
	struct _matrix *b = admm_b200_f64(x0, ADMM_ITERATIONS, qf);
	int rc = 0;
	if (!a || !b) {
		rc = -2;
	} else {
		memcpy(x_ref, a->elements, sizeof(double) * N_DIM);
		memcpy(x_gpu, b->elements, sizeof(double) * N_DIM);
	}
	if (a)
		matrix_free(a);
	if (b)
		matrix_free(b);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(q);
	matrix_free(p);
	return rc;
}

/* the same for the Newton solver (main.c:46-48): test_optimizer() through the function
 * pointer with newton_method_with_line_search and with the drop-in newton_b200 */
int compat_run_newton(unsigned seed, double *x_ref, double *x_gpu, double *f_out)
{
	if (!ready) {
		kmalloc_init();
		ready = 1;
	}
	srand(seed);
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *q = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, q, 0);
	if (!p || !q || !x0 || !qf)
		return -1;
	matirx_random_pos_def(p, P_RAND_ENTRY_MIN, P_RAND_ENTRY_MAX);
	matrix_random(q, Q_RAND_ENTRY_MIN, Q_RAND_ENTRY_MAX);
	matrix_random(x0, X_RAND_ENTRY_MIN, X_RAND_ENTRY_MAX);

	test_optimizer(qf, HESS_ITERATIONS, x0, "newton", newton_method_with_line_search);
	test_optimizer(qf, HESS_ITERATIONS, x0, "newton_b200", newton_b200);

	struct _matrix *a = newton_method_with_line_search(x0, HESS_ITERATIONS, qf);
	struct _matrix *b = newton_b200(x0, HESS_ITERATIONS, qf);
	int rc = 0;
	if (!a || !b) {
		rc = -2;
	} else {
		memcpy(x_ref, a->elements, sizeof(double) * N_DIM);
		memcpy(x_gpu, b->elements, sizeof(double) * N_DIM);
		f_out[0] = quadratic_form_eval(qf, a);
		f_out[1] = quadratic_form_eval(qf, b);
	}
	if (a)
		matrix_free(a);
	if (b)
		matrix_free(b);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(q);
	matrix_free(p);
	return rc;
}

/* The reference's own problem-file writer: draws a problem like main.c:36-40 and calls
 * test_reference() (test/test.c:108-130), which writes "tmp_test_file" in the current
 * directory, runs `python test/qp_ref.py tmp_test_file` and unlinks the file.  The caller
 * chdir()s to a scratch directory whose test/qp_ref.py keeps a copy of the file. */
int compat_write_reference_file(unsigned seed, double *P_out, double *q_out)
{
	if (!ready) {
		kmalloc_init();
		ready = 1;
	}
	srand(seed);
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *q = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, q, 0);
	if (!p || !q || !qf)
		return -1;
	matirx_random_pos_def(p, P_RAND_ENTRY_MIN, P_RAND_ENTRY_MAX);
	matrix_random(q, Q_RAND_ENTRY_MIN, Q_RAND_ENTRY_MAX);
	test_reference(qf);
	memcpy(P_out, p->elements, sizeof(double) * N_DIM * N_DIM);
	memcpy(q_out, q->elements, sizeof(double) * N_DIM);
	quadratic_form_free(qf);
	matrix_free(q);
	matrix_free(p);
	return 0;
}

/* and for gradient descent (main.c:42-44) */
int compat_run_gradient(unsigned seed, double *x_ref, double *x_gpu, double *f_out)
{
	if (!ready) {
		kmalloc_init();
		ready = 1;
	}
	srand(seed);
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *q = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, q, 0);
	if (!p || !q || !x0 || !qf)
		return -1;
	matirx_random_pos_def(p, P_RAND_ENTRY_MIN, P_RAND_ENTRY_MAX);
	matrix_random(q, Q_RAND_ENTRY_MIN, Q_RAND_ENTRY_MAX);
	matrix_random(x0, X_RAND_ENTRY_MIN, X_RAND_ENTRY_MAX);

	test_optimizer(qf, GRAD_ITERATIONS, x0, "gradient", gradient_descent_with_line_search);
	test_optimizer(qf, GRAD_ITERATIONS, x0, "gradient_b200", gradient_b200);

	struct _matrix *a = gradient_descent_with_line_search(x0, GRAD_ITERATIONS, qf);
	struct _matrix *b = gradient_b200(x0, GRAD_ITERATIONS, qf);
	int rc = 0;
	if (!a || !b) {
		rc = -2;
	} else {
		memcpy(x_ref, a->elements, sizeof(double) * N_DIM);
		memcpy(x_gpu, b->elements, sizeof(double) * N_DIM);
		f_out[0] = quadratic_form_eval(qf, a);
		f_out[1] = quadratic_form_eval(qf, b);
	}
	if (a)
		matrix_free(a);
	if (b)
		matrix_free(b);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(q);
	matrix_free(p);
	return rc;
}
