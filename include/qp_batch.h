/*
 * qp_batch.h -- C ABI of the B200-native batched box-QP engine (libqpb200.so).
 *
 *   minimize   (1/2) x'Px + q'x      subject to   lb <= x <= ub
 *
 * solved for `count` independent problems at once with the reference's ADMM
 * iteration.  Every entry point below names the reference interface it
 * replaces; paths are relative to the reference repository
 * (YangLingyuan/Embedded-qp-solver).
 *
 *   reference, one problem per call          this library, a batch per call
 *   ---------------------------------        ------------------------------
 *   struct _matrix *admm(x0, iterations, qf) qp_solve_batch(), qp_admm_qf(),
 *     qp_solvers/qp_solvers.h:14-16            qp_admm_qf_batch()
 *     qp_solvers/qp_solvers.c:255-319
 *   quadratic_form_eval()  qp.h:20-21        qp_eval_batch()
 *   newton_method_with_line_search()         qp_newton_batch(), qp_newton_qf()
 *     qp_solvers/qp_solvers.h:9-12
 *   gradient_descent_with_line_search()      qp_gradient_batch(), qp_gradient_qf()
 *     qp_solvers/qp_solvers.h:4-7
 *   test_reference() file, test/test.c:108   qp_problem_file_write/count/read()
 *   kmalloc_init()/static pools              qp_workspace_create()/destroy()
 *     kmalloc/kmalloc.h:16-18
 *   (no counterpart: rows l <= A x <= u)     qp_solve_general_batch()
 *
 * Plain C: pointers and sizes only.  All arrays are caller-owned; the library
 * never allocates or frees memory that the caller sees.  Element order of
 * every matrix is the reference's row-major `elements[row * ncols + col]`
 * (matrix/matrix_ops/matrix_ops.c:68-78).  P must be fully populated and
 * symmetric, with P + rho*I positive definite (a convex QP, ReadMe.txt:1-3).
 * Which triangle a kernel reads is unspecified (the thread-per-QP kernels
 * read the upper one, the lane-group kernels read rows as columns); a caller
 * who holds only one triangle passes it as QP_P_UPPER.
 *
 * The arithmetic type of the batch entry points is binary32 (`float`); the
 * reference computes in binary64.  The qf entry points take the reference's
 * own `double` structs and convert.
 */
#ifndef QP_BATCH_H
#define QP_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QP_B200_ABI_VERSION 2 /* 2: qp_batch.p_layout, qp_options.schedule */

/* largest problem dimension: the box ADMM entry points (qp_solve_batch, qp_solve_batch_ws,
 * qp_admm_qf, qp_admm_qf_batch) take n <= QP_MAX_DIM_BOX; the Newton, gradient and
 * general-rows entry points take n <= QP_MAX_DIM */
#define QP_MAX_DIM 64
#define QP_MAX_DIM_BOX 128

/* Per-problem outcome, written to qp_batch_solution.status.  The reference
 * returns no status; these are defined on its state:        */
#define QP_STATUS_CONVERGED 0 /* stop test passed, qp_solvers.c:308-309          */
#define QP_STATUS_MAX_ITER  1 /* loop ran out of `iterations`, qp_solvers.c:284   */
#define QP_STATUS_SINGULAR  2 /* factorisation of P + rho*I hit a pivot <= 0; the
                                 reference prints "singular matrix" at
                                 matrix_ops.c:511-515 and carries on           */
#define QP_STATUS_NONFINITE 3 /* x contains Inf/NaN                               */

/* Return codes of the API calls (0 = success). */
#define QP_OK                 0
#define QP_ERR_INVALID       -1 /* null pointer, n == 0, bad option            */
#define QP_ERR_UNSUPPORTED   -2 /* n > QP_MAX_DIM                              */
#define QP_ERR_CUDA          -3 /* a CUDA call failed; see qp_last_error()     */
#define QP_ERR_NO_DEVICE     -4 /* no usable GPU / device id out of range      */
#define QP_ERR_IO            -5 /* problem file: cannot open / short read/write */

/* Where the arrays of a batch live. */
#define QP_MEM_HOST   0 /* host pointers: the call copies in and out, sharding
                           the batch over the GPUs in device_mask             */
#define QP_MEM_DEVICE 1 /* device pointers on qp_options.device: the kernel is
                           launched on qp_options.stream and the call returns
                           without synchronising                              */

/*
 * A batch of `count` problems of dimension n, packed back to back.
 * Replaces one `struct _quadratic_form {p, q, r}` (qp_solvers/qp.h:8-13) per
 * problem plus the compile-time box ADMM_BOX_CONSTRAINT_{MAX,MIN}
 * (config.h:29-30, broadcast at qp_solvers.c:277-280).
 */
typedef struct qp_batch {
	uint32_t n;          /* problem dimension (the reference's N_DIM)          */
	uint32_t mem;        /* QP_MEM_HOST or QP_MEM_DEVICE, for every pointer    */
	size_t count;        /* number of problems                                 */
	const float *P;      /* [count][n][n] row-major                            */
	const float *q;      /* [count][n]                                         */
	const float *lb;     /* [count][n], or NULL: lb_scalar for every entry     */
	const float *ub;     /* [count][n], or NULL: ub_scalar for every entry     */
	float lb_scalar;     /* used when lb == NULL (the reference's only mode)   */
	float ub_scalar;     /* used when ub == NULL                               */
	const float *z0;     /* [count][n] warm start of z, or NULL = 0            */
	const float *u0;     /* [count][n] warm start of u, or NULL = 0
	                        (the reference zeroes both, qp_solvers.c:263-264)  */
	const float *rho;    /* [count] penalty of each problem (> 0), or NULL:
	                        qp_options.rho for all (the reference's RHO_ADMM,
	                        qp_solvers.c:281).  A problem whose entry is not a
	                        positive finite number ends with QP_STATUS_SINGULAR */
	uint32_t p_layout;   /* QP_P_DENSE (the reference's layout) or QP_P_UPPER   */
	uint32_t p_share;    /* 0 or 1: every problem has its own P.  k > 1: problems
	                        i*k .. i*k+k-1 share P[i] (P holds ceil(count / k)
	                        matrices): one model, many right-hand sides -- an MPC
	                        host that calls admm() with the same qf->p and a new
	                        qf->q per state (struct _quadratic_form keeps them
	                        apart, qp_solvers/qp.h:8-13).  Only q, the bounds, the
	                        warm start and rho are per problem then: 64 instead of
	                        1 088 bytes per problem at n = 16.  qp_solve_batch()
	                        only (QP_ERR_UNSUPPORTED elsewhere)                    */
} qp_batch;

/* Layout of qp_batch.P. */
#define QP_P_DENSE 0 /* [count][n][n] row-major, matrix_ops.c:68-78               */
#define QP_P_UPPER 1 /* [count][qp_packed_stride(n)]: the row-major upper triangle
                        (i <= j: P00 P01 .. P0,n-1 P11 ..), n (n + 1) / 2 floats
                        padded with zeros to a multiple of 4.  P is symmetric, so
                        this is all of it: 544 instead of 1 024 bytes per problem at
                        n = 16 over PCIe and out of HBM.  n = 8, 12, 16 only
                        (QP_ERR_UNSUPPORTED otherwise); arrays 16-byte aligned   */
size_t qp_packed_stride(uint32_t n); /* floats per problem of a QP_P_UPPER batch */

/* Caller-owned outputs; any pointer may be NULL to skip that output. */
typedef struct qp_batch_solution {
	float *x;        /* [count][n] what admm() returns: the x-update iterate     */
	float *z;        /* [count][n] the clipped iterate (always inside the box)   */
	float *lambda;   /* [count][n] rho*u; >0 upper bound active, <0 lower        */
	int32_t *status; /* [count]    QP_STATUS_*                                   */
	int32_t *iters;  /* [count]    iterations executed (i+1 at the break)        */
} qp_batch_solution;

/* Solver knobs; qp_options_default() installs the reference's constants. */
typedef struct qp_options {
	float rho;            /* 1      qp_solvers.c:281                            */
	float alpha;          /* 1      over-relaxation, qp_solvers.c:282           */
	float abstol;         /* 1e-4   qp_solvers.c:17                             */
	float reltol;         /* 1e-2   qp_solvers.c:18                             */
	uint32_t max_iter;    /* 10000  the `iterations` argument, config.h:38      */
	uint32_t device_mask; /* QP_MEM_HOST: bit i = use GPU i; 0 = GPU `device`   */
	int32_t device;       /* CUDA device ordinal, -1 = the current device       */
	void *stream;         /* QP_MEM_DEVICE: cudaStream_t to launch on (NULL =
	                         the legacy default stream)                         */
	uint32_t chunk;       /* QP_MEM_HOST: problems per pipelined copy/solve
	                         chunk, 0 = choose                                  */
	uint32_t kernel;      /* QP_KERNEL_AUTO, or force one kernel family (tests,
	                         measurements); a family that cannot take the
	                         batch returns QP_ERR_UNSUPPORTED                   */
	uint32_t schedule;    /* QP_SCHEDULE_*: how problems that need many more
	                         iterations than their neighbours are handled      */
	uint32_t precision;   /* QP_PRECISION_F32 (0, the product: every entry point)
	                         or QP_PRECISION_F64: the qf entry points
	                         (qp_admm_qf_batch) solve in the reference's own
	                         binary64 -- its iteration counts on its own
	                         ill-conditioned test problems, x to ~1e-12; a
	                         correctness path, n <= QP_MAX_DIM, one GPU.
	                         QP_ERR_UNSUPPORTED in the float entry points       */
} qp_options;
#define QP_PRECISION_F32 0
#define QP_PRECISION_F64 1

/*
 * Scheduling of uneven iteration counts.  The reference's own test problems
 * (matirx_random_pos_def, matrix_ops.c:699-734) need 3 ... 10 000 iterations;
 * with one problem per lane a warp would run to its slowest problem.
 *   QP_SCHEDULE_PASSES  a problem gets a window of iterations per kernel launch
 *                       (24 in the first, doubling each time); one that has not
 *                       converged when its window closes is handed (z, u,
 *                       iteration count) to a list in device memory, which
 *                       the next launch on the same stream works through 32
 *                       at a time: stragglers share warps with stragglers and
 *                       no warp idles for longer than a window.  The list
 *                       lengths stay on the device; the library enqueues
 *                       about log2(max_iter / 24) launches and never
 *                       synchronises.  Results are bit-identical to
 *                       QP_SCHEDULE_SINGLE.  Uses a per-device scratch arena
 *                       owned by the library (allocated on first use, grown
 *                       when a larger batch arrives; 8 n + 12 bytes per
 *                       problem, + 4 n^2 for 24 <= n <= 40); solves that use
 *                       it are serialised on the device.
 *   QP_SCHEDULE_SINGLE  one launch, every problem stays with the warp it
 *                       started in; no scratch, no allocation.
 *   QP_SCHEDULE_AUTO    SINGLE, except for batches of at least 4 096 problems
 *                       with max_iter > 64 on the kernel families that have
 *                       passes (n = 4, 8, 12, 16 and n = 24, 32, 40): there
 *                       the single launch also adds up, per warp, the
 *                       iterations its problems needed and the iterations it
 *                       waited for its slowest one; the sums reach the host
 *                       asynchronously, and the NEXT such solve on the device
 *                       uses PASSES when less than half of the lane-iterations
 *                       did work and the mean iteration count is at least 16
 *                       (n <= 16) or 64 (n >= 24); it goes back to SINGLE when
 *                       fewer than 1 problem in 128 outlives the first window.
 */
#define QP_SCHEDULE_AUTO   0
#define QP_SCHEDULE_SINGLE 1
#define QP_SCHEDULE_PASSES 2

#define QP_KERNEL_AUTO 0
#define QP_KERNEL_TPQ  1 /* one QP per thread:       n = 4, 8, 12, 16              */
#define QP_KERNEL_GPQ  2 /* cyclic rows per lane:    any n <= 64                   */
#define QP_KERNEL_BPQ  3 /* four rows per lane:      n = 8, 16, 24, 32, 40         */
#define QP_KERNEL_CPQ  4 /* one CTA per QP:          any n <= 128 (auto: n > 64)   */
/* qp_solve_general_batch() only (the box entry points answer QP_ERR_UNSUPPORTED): */
#define QP_KERNEL_GEN_SMEM 5 /* operands in shared memory: n <= 64, m <= 128        */
#define QP_KERNEL_GEN_REG  6 /* operands in registers:     n <= 32, m <= 64 (auto)  */

void qp_options_default(qp_options *opt);

/*
 * Solve a batch.  Replaces a loop of
 *     x = admm(x0, iterations, qf)          qp_solvers/qp_solvers.c:255-319
 * (as called through the solver function pointer at test/test.c:97).
 * opt == NULL means qp_options_default().  Never exits the process: API
 * misuse returns QP_ERR_*, per-problem trouble goes to sol->status[].
 */
int qp_solve_batch(const qp_batch *prob, const qp_batch_solution *sol,
		   const qp_options *opt);

/* Opaque preallocated device workspace: streams, device staging buffers and
 * pinned host buffers for the QP_MEM_HOST path.  Replaces the reference's
 * static slab pools (kmalloc/kmalloc.c:37-66).  qp_solve_batch() keeps one
 * per device internally; create one explicitly to control its lifetime and
 * to keep the first call free of allocation. */
typedef struct qp_workspace qp_workspace;
int qp_workspace_create(uint32_t n, size_t max_count, uint32_t device_mask,
			qp_workspace **out);
void qp_workspace_destroy(qp_workspace *ws);
int qp_solve_batch_ws(qp_workspace *ws, const qp_batch *prob,
		      const qp_batch_solution *sol, const qp_options *opt);

/*
 * Objective values f(x) = (1/2) x'Px + q'x + r for a batch (host or device
 * pointers as prob->mem says; `x`, `f` live where prob's arrays live).
 * Replaces quadratic_form_eval(), qp_solvers/qp.c:9-27 -- the number the
 * reference's tests print (test/test.c:100-102).
 */
int qp_eval_batch(const qp_batch *prob, const float *x, float r, float *f,
		  const qp_options *opt);

/*
 * Batched unconstrained minimiser.  Replaces a loop of
 *     x = newton_method_with_line_search(x0, iterations, qf)
 *                                         qp_solvers/qp_solvers.c:103-144
 * (the second solver the reference's tests run through the same function
 * pointer, main.c:46-48): hessian_inv = P^-1 once, then Newton steps with
 * the reference's Armijo backtracking (alpha0 = 1, shrink 0.9, c2 = 1e-4,
 * :37-63) until ||P x + q|| < 0.1 (:125) or `iterations` steps.  With the
 * box inactive its answer is the ADMM answer -- the reference's own
 * cross-check (test/test.c:89-106 prints both objectives).
 * prob->lb/ub/z0/u0 are ignored.  x0 == NULL starts from 0.  x [count][n] is
 * required; status (0 gradient test passed, 1 step cap, 2 a pivot of P <= 0,
 * 3 non-finite), steps (Newton steps taken) and f (objective at x, with
 * constant term r) may be NULL.  Pointers live where prob->mem says.
 */
int qp_newton_batch(const qp_batch *prob, const float *x0, unsigned iterations,
		    float *x, int32_t *status, int32_t *steps, float *f, float r,
		    const qp_options *opt);

/*
 * The reference's third solver, same arguments and outputs:
 *     x = gradient_descent_with_line_search(x0, iterations, qf)
 *                                         qp_solvers/qp_solvers.c:65-101
 * (main.c:42-44): d = -(P x + q), the same line search, the same stop test;
 * no matrix is inverted, so status 2 does not occur.
 */
int qp_gradient_batch(const qp_batch *prob, const float *x0, unsigned iterations,
		      float *x, int32_t *status, int32_t *steps, float *f, float r,
		      const qp_options *opt);

/*
 * General inequality rows  l <= A x <= u  (SURVEY.md 8(f) rank 3).  The reference has only
 * the box; this entry point runs the SAME scaled-form ADMM (qp_solvers.c:255-319) for the
 * splitting A x = z:  (P + rho A'A) x = rho A'(z - u) - q;  z = clip(alpha A x + (1-alpha) z
 * + u, l, u);  u += alpha A x + (1-alpha) z_old - z;  stop when ||A x - z|| < sqrt(m) abstol
 * + reltol max(||A x||, ||z||) and rho ||A'(z - z_old)|| < sqrt(n) abstol + reltol rho ||A'u||.
 * With A = I it is the reference's iteration line by line (and is tested against the same
 * golden vectors); for other A there is no reference behaviour to match.
 * n <= QP_MAX_DIM, m <= QP_MAX_ROWS.  P + rho A'A must be positive definite.
 * Two kernels: operands in registers for n <= 32 and m <= 64 (csrc/qp_genr.cuh), in shared
 * memory beyond (csrc/qp_gen.cuh); qp_options.kernel = QP_KERNEL_GEN_REG / QP_KERNEL_GEN_SMEM
 * forces one (QP_ERR_UNSUPPORTED when the shape does not fit, or for a box kernel family).
 * QP_MEM_DEVICE: launched on qp_options.stream, no synchronisation; batches longer than one
 * wave draw their problems from one of 256 per-device ticket counters, zeroed on that stream.
 * QP_MEM_HOST: chunks of ~32 MB alternate between two device buffers and two streams (kept per
 * device), copy-in overlapping the solve and copy-out of the chunk before; synchronous call.
 * Pin the arrays (cudaHostRegister) for full PCIe speed: pageable arrays go through the
 * driver's staged copies.
 */
#define QP_MAX_ROWS 128
typedef struct qp_general_batch {
	uint32_t n;       /* variables                                            */
	uint32_t m;       /* rows of A                                            */
	uint32_t mem;     /* QP_MEM_HOST or QP_MEM_DEVICE, for every pointer      */
	uint32_t pad;
	size_t count;
	const float *P;   /* [count][n][n] row-major, symmetric                   */
	const float *q;   /* [count][n]                                           */
	const float *A;   /* [count][m][n] row-major                              */
	const float *l;   /* [count][m]; -INFINITY for a one-sided row            */
	const float *u;   /* [count][m]; +INFINITY for a one-sided row            */
	const float *z0;  /* [count][m] warm start of z, or NULL = 0              */
	const float *u0;  /* [count][m] warm start of the scaled dual, or NULL    */
} qp_general_batch;
typedef struct qp_general_solution {
	float *x;         /* [count][n], required                                 */
	float *z;         /* [count][m] clipped A x, or NULL                      */
	float *y;         /* [count][m] multipliers of the rows (rho u), or NULL  */
	int32_t *status;  /* [count] QP_STATUS_*, or NULL                         */
	int32_t *iters;   /* [count], or NULL                                     */
} qp_general_solution;
int qp_solve_general_batch(const qp_general_batch *prob, const qp_general_solution *sol,
			   const qp_options *opt);

/*
 * Entry points on the reference's own structs (binary64, pool-owned).  The
 * two structs below are layout-identical to `struct _matrix`
 * (matrix/matrix_type/matrix_type.h:20-23: rows in bits 31:16 and columns in
 * bits 15:0 of `dimensions`) and `struct _quadratic_form`
 * (qp_solvers/qp.h:8-13); a reference host passes its own pointers.
 */
struct qp_ref_matrix {
	unsigned dimensions;
	double *elements;
};
struct qp_ref_quadratic_form {
	struct qp_ref_matrix *p;
	struct qp_ref_matrix *q;
	double r;
};

/* One problem: what the body of a drop-in `admm_b200(x0, iterations, qf)`
 * calls.  lb/ub are the reference's scalar box.  Writes
 * n doubles to x_out; status/iters may be NULL. */
int qp_admm_qf(const struct qp_ref_quadratic_form *qf, unsigned iterations,
	       double lb, double ub, double *x_out, int32_t *status,
	       int32_t *iters);

/* Many reference structs at once (array-of-structs host): gathers P and q
 * from each qfs[i], solves them as one batch, scatters x into x_out[i]
 * (n doubles each).  All problems must share one dimension. */
int qp_admm_qf_batch(const struct qp_ref_quadratic_form *const *qfs,
		     size_t count, unsigned iterations, double lb, double ub,
		     double *const *x_out, int32_t *status, int32_t *iters,
		     const qp_options *opt);

/* One problem on the reference's structs: the body of a drop-in
 * `newton_b200(x0, iterations, qf)`.  x0 may be NULL (= 0). */
int qp_newton_qf(const struct qp_ref_quadratic_form *qf, const double *x0,
		 unsigned iterations, double *x_out, int32_t *status,
		 int32_t *steps);
int qp_gradient_qf(const struct qp_ref_quadratic_form *qf, const double *x0,
		   unsigned iterations, double *x_out, int32_t *status,
		   int32_t *steps);

/*
 * The reference's problem file: [double n][P, n*n doubles row-major][q, n doubles],
 * written by test_reference() (test/test.c:108-126) for its Python cross-check
 * (test/qp_ref.py:8-30 parses it).  Host-only calls, no GPU involved.  Files with
 * several records back to back (one dimension) are accepted as a batch; a
 * one-record file is byte-identical to what the reference writes.
 *   write: `count` records from packed double arrays, truncating or appending;
 *   count: number of records (negative = QP_ERR_*), n of the file in *n;
 *   read:  records [first, first+count) into packed double arrays.
 * `python -m embedded_qp_solver_b200.qp_ref FILE` is the drop-in for the
 * reference's `python test/qp_ref.py FILE` (test/test.c:10): it solves every
 * record on the GPU and prints the same two lines.
 */
int qp_problem_file_write(const char *path, uint32_t n, size_t count,
			  const double *P, const double *q, int append);
long long qp_problem_file_count(const char *path, uint32_t *n);
int qp_problem_file_read(const char *path, size_t first, size_t count,
			 uint32_t n, double *P, double *q);

/* 1 when the QP_MEM_HOST path packs the upper triangles of P on the host before the copy
 * (n = 8, 12, 16, one GPU fed from this host, at least 12 hardware threads,
 * QPB200_HOST_PACK != 0): P is symmetric and the kernels read only its upper triangle, so
 * 544 instead of 1 024 bytes per problem cross PCIe at n = 16.  The library times the
 * packing against the copy engine and gives it up for the rest of the process when the
 * host turns out too slow for it to pay; the answer then turns to 0. */
int qp_host_path_packs(uint32_t n);
/* Bytes the QP_MEM_HOST path has handed to the copy engine (host to device) since the
 * process started: packed and plain chunks are mixed as the host keeps up, so this, not
 * the answer above, says what crossed PCIe. */
unsigned long long qp_host_path_h2d_bytes(void);
/* The packing step of that path on its own (tests, measurements; no GPU involved): the
 * row-major upper triangles of `count` n x n matrices (n = 8, 12, 16), records
 * (n (n + 1) / 2 rounded up to a multiple of 4) floats apart, `out` 16-byte aligned.
 * wide: 0 = 128-bit moves, 1 = AVX-512 with ordinary stores, 2 = AVX-512 with non-temporal
 * stores (QP_ERR_UNSUPPORTED without AVX-512; n = 16 only), -1 = what the host path uses. */
int qp_host_pack_upper(const float *P, float *out, size_t count, uint32_t n, int wide);

/* Counters of the schedules above on `device` since the process started:
 *   stats[0] problems that outlived the first window of a PASSES solve,
 *   stats[1] problems those first passes looked at,
 *   stats[2] evictions by the later passes,
 *   stats[3] sum of iteration counts seen by AUTO's probing single launches,
 *   stats[4] sum over their warps of (largest iteration count x problems in the warp),
 *   stats[5] problems they looked at.
 * Waits for the device. */
int qp_schedule_stats(int device, unsigned long long stats[6]);

/* Pure-copy ceiling of the QP_MEM_HOST path (measurement helper): moves h2d_bytes host to
 * device and d2h_bytes device to host per repeat between pinned (pinned != 0) or pageable
 * host buffers and `device`, both directions at once on two streams, in pieces of
 * `piece_bytes` (0 = 32 MiB) the way the chunk pipeline does, and returns the seconds per
 * repeat (best of `repeats`), or a negative QP_ERR_*.  No kernel runs: this is what the
 * copy engines and the host's memory system alone allow for those byte counts. */
double qp_probe_copy_seconds(int device, size_t h2d_bytes, size_t d2h_bytes, int pinned,
			     size_t piece_bytes, int repeats);

/* Diagnostics. */
const char *qp_last_error(void);      /* text of the last QP_ERR_CUDA, per thread */
int qp_device_count(void);
const char *qp_kernel_for(uint32_t n, int aligned16); /* "tpq<16>", "gpq<40>", ... */

/* Measurement helpers used by bench.py for the roofline denominators:
 * a register-resident FP32 FMA loop (returns TFLOP/s) and a device copy
 * (returns GB/s, read+write bytes), both timed with CUDA events on `device`. */
double qp_probe_fp32_tflops(int device, int repeats);
double qp_probe_hbm_gbs(int device, size_t bytes, int repeats);

#ifdef __cplusplus
}
#endif
#endif /* QP_BATCH_H */
