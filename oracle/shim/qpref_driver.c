/*
 * oracle/shim/qpref_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin C driver linked together with the UNMODIFIED reference sources into
 * oracle/_ref/libqpref_n<N>.so.  It only marshals plain arrays in and out of
 * the reference's pool-allocated `struct _matrix` objects and calls the
 * reference's public functions; no arithmetic of the path is done here.
 *
 * Reference entry points exercised (paths relative to /root/reference):
 *   admm                            qp_solvers/qp_solvers.c:255-319
 *   newton_method_with_line_search  qp_solvers/qp_solvers.c:103-144
 *   gradient_descent_with_line_search  qp_solvers/qp_solvers.c:65-101
 *   quadratic_form_eval             qp_solvers/qp.c:9-27
 *   matrix_invert                   matrix/matrix_ops/matrix_ops.c:551-630
 *   inversion_test/scalar_prod_test test/test.c:37-87
 *   kmalloc_init                    kmalloc/kmalloc.c:48-66
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "test.h" /* reference header: pulls matrix_ops.h, qp.h, qp_solvers.h */

double qpref_box_hi[N_DIM];
double qpref_box_lo[N_DIM];

static int qpref_ready = 0;

unsigned qpref_dim(void)
{
	return N_DIM;
}

void qpref_init(void)
{
	/* the reference's pools are global and must be set up exactly once
	 * per process (main.c:10) */
	if (!qpref_ready) {
		kmalloc_init();
		qpref_ready = 1;
	}
}

static void set_box(const double *lb, const double *ub)
{
	for (unsigned k = 0; k < N_DIM; k++) {
		qpref_box_lo[k] = lb[k];
		qpref_box_hi[k] = ub[k];
	}
}

/* One call of the reference's admm() on plain arrays.  Returns 0, or -1 if a
 * pool allocation failed. */
int qpref_admm(const double *P, const double *q, const double *lb,
	       const double *ub, unsigned iterations, double *x_out)
{
	qpref_init();
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *qv = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	if (!p || !qv || !x0)
		return -1;
	struct _quadratic_form *qf = quadratic_form_alloc(p, qv, 0);
	if (!qf)
		return -1;

	memcpy(p->elements, P, sizeof(double) * N_DIM * N_DIM);
	memcpy(qv->elements, q, sizeof(double) * N_DIM);
	memset(x0->elements, 0, sizeof(double) * N_DIM);
	set_box(lb, ub);

	struct _matrix *x = admm(x0, iterations, qf);
	memcpy(x_out, x->elements, sizeof(double) * N_DIM);

	matrix_free(x);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(qv);
	matrix_free(p);
	return 0;
}

/* Loop of qpref_admm over `count` packed problems; returns the seconds spent
 * inside the reference's admm() calls only (CLOCK_MONOTONIC), which is what
 * the CPU baseline reports. */
double qpref_admm_many(const double *P, const double *q, const double *lb,
		       const double *ub, unsigned long count,
		       unsigned iterations, double *x_out)
{
	qpref_init();
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *qv = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, qv, 0);
	memset(x0->elements, 0, sizeof(double) * N_DIM);

	double spent = 0;
	for (unsigned long c = 0; c < count; c++) {
		memcpy(p->elements, P + c * N_DIM * N_DIM,
		       sizeof(double) * N_DIM * N_DIM);
		memcpy(qv->elements, q + c * N_DIM, sizeof(double) * N_DIM);
		set_box(lb + c * N_DIM, ub + c * N_DIM);

		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		struct _matrix *x = admm(x0, iterations, qf);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		spent += (double)(t1.tv_sec - t0.tv_sec)
			 + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);

		if (x_out)
			memcpy(x_out + c * N_DIM, x->elements,
			       sizeof(double) * N_DIM);
		matrix_free(x);
	}

	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(qv);
	matrix_free(p);
	return spent;
}

int qpref_newton(const double *P, const double *q, const double *x0_in,
		 unsigned iterations, double *x_out)
{
	qpref_init();
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *qv = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	if (!p || !qv || !x0)
		return -1;
	struct _quadratic_form *qf = quadratic_form_alloc(p, qv, 0);
	memcpy(p->elements, P, sizeof(double) * N_DIM * N_DIM);
	memcpy(qv->elements, q, sizeof(double) * N_DIM);
	memcpy(x0->elements, x0_in, sizeof(double) * N_DIM);

	struct _matrix *x = newton_method_with_line_search(x0, iterations, qf);
	memcpy(x_out, x->elements, sizeof(double) * N_DIM);

	matrix_free(x);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(qv);
	matrix_free(p);
	return 0;
}

int qpref_gradient(const double *P, const double *q, const double *x0_in,
		   unsigned iterations, double *x_out)
{
	qpref_init();
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *qv = matrix_alloc(Nx1);
	struct _matrix *x0 = matrix_alloc(Nx1);
	if (!p || !qv || !x0)
		return -1;
	struct _quadratic_form *qf = quadratic_form_alloc(p, qv, 0);
	memcpy(p->elements, P, sizeof(double) * N_DIM * N_DIM);
	memcpy(qv->elements, q, sizeof(double) * N_DIM);
	memcpy(x0->elements, x0_in, sizeof(double) * N_DIM);

	struct _matrix *x = gradient_descent_with_line_search(x0, iterations, qf);
	memcpy(x_out, x->elements, sizeof(double) * N_DIM);

	matrix_free(x);
	quadratic_form_free(qf);
	matrix_free(x0);
	matrix_free(qv);
	matrix_free(p);
	return 0;
}

double qpref_eval(const double *P, const double *q, double r, const double *xv)
{
	qpref_init();
	struct _matrix *p = matrix_alloc(NxN);
	struct _matrix *qv = matrix_alloc(Nx1);
	struct _matrix *x = matrix_alloc(Nx1);
	struct _quadratic_form *qf = quadratic_form_alloc(p, qv, r);
	memcpy(p->elements, P, sizeof(double) * N_DIM * N_DIM);
	memcpy(qv->elements, q, sizeof(double) * N_DIM);
	memcpy(x->elements, xv, sizeof(double) * N_DIM);
	double f = quadratic_form_eval(qf, x);
	quadratic_form_free(qf);
	matrix_free(x);
	matrix_free(qv);
	matrix_free(p);
	return f;
}

/* in-place inverse of a row-major N_DIM x N_DIM array */
int qpref_invert(double *A)
{
	qpref_init();
	struct _matrix *m = matrix_alloc(NxN);
	if (!m)
		return -1;
	memcpy(m->elements, A, sizeof(double) * N_DIM * N_DIM);
	matrix_invert(m);
	memcpy(A, m->elements, sizeof(double) * N_DIM * N_DIM);
	matrix_free(m);
	return 0;
}

/* sizeof/offset facts of the reference structs, for the layout tests */
void qpref_layout(unsigned long out[6])
{
	struct _matrix m;
	struct _quadratic_form f;
	out[0] = sizeof(m);
	out[1] = (unsigned long)((char *)&m.elements - (char *)&m);
	out[2] = sizeof(f);
	out[3] = (unsigned long)((char *)&f.q - (char *)&f);
	out[4] = (unsigned long)((char *)&f.r - (char *)&f);
	m.dimensions = 0;
	MATRIX_SET_ROW((&m), 7);
	MATRIX_SET_COL((&m), 3);
	out[5] = m.dimensions;
}

/* the reference's own two self-tests; they report on stdout/stderr */
void qpref_selftests(unsigned seed)
{
	qpref_init();
	srand(seed);
	inversion_test();
	scalar_prod_test();
}
