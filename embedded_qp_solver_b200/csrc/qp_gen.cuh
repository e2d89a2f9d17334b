// qp_gen.cuh -- general inequality rows  l <= A x <= u  (SURVEY.md 8(f) rank 3), T lanes per
// problem (T = 4, 8, 16 or 32: n <= 2T variables, m <= 4T rows; n <= 64, m <= 128).
//
// The reference has no such path: its admm() (/root/reference/qp_solvers/qp_solvers.c:255-319)
// is the scaled-form ADMM for  min f(x) + g(z)  s.t. x = z  with g the box indicator.  This
// kernel runs the SAME iteration for the splitting  A x = z:
//
//   x-update   (P + rho A'A) x = rho A'(z - u) - q          (:146-159 has A = I: (P + rho I) x = rho (z-u) - q)
//   relaxation zh = alpha A x + (1 - alpha) z               (:161-174)
//   z-update   z = min(max(zh + u, l), u_bound)             (:176-190)
//   u-update   u += zh - z                                   (:192-203)
//   r = ||A x - z||,  s = rho ||A'(z - z_old)||              (:205-231)
//   eps_pri  = sqrt(m) abstol + reltol max(||A x||, ||z||)   (:233-243)
//   eps_dual = sqrt(n) abstol + reltol rho ||A'u||           (:245-253)
//   stop when r < eps_pri and s < eps_dual                   (:308)
//
// (Boyd et al., "Distributed optimization ... ADMM", 3.3.1 and 5.2, the text the reference
// follows.)  With A = I, m = n every line above is the reference's, so the A = I case is
// pinned by the same golden vectors as the box kernels (tests/test_general_gpu.py); for a
// general A there is no reference oracle -- PARITY UNPINNED beyond A = I -- and the checks
// are the numpy restatement in oracle/general_np.py and the KKT conditions.
//
// Correctness-first kernel: P + rho A'A, its inverse R and A live in shared memory; member g
// of a group owns coordinates i = g, g + T and rows j = g + T k (k < 4).  A'z and A'u are
// carried along, so an iteration costs one mat-vec with R, one with A and one pass over A
// for A'z and A'u together.  The groups of a warp run in lockstep until the last of them has
// stopped; a group that has stopped keeps executing but no longer commits its state.
#pragma once

#include "qp_aux_args.h"

namespace qpb {

// sum over the T lanes of a group (groups are aligned runs of T lanes)
template <int T> __device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
	for (int o = T / 2; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// smem_gauss_jordan (qp_common.cuh) for a group of T lanes: member g owns columns g, g + T
template <int T>
__device__ __forceinline__ bool smem_gauss_jordan_group(float *W, int n, int ld, float *col, int g)
{
	const int i0 = g, i1 = g + T;
	const bool h0 = i0 < n, h1 = i1 < n;
	bool bad = false;
	for (int k = 0; k < n; ++k) {
		const float piv = W[k * ld + k];
		bad |= !(piv > 0.0f);
		const float dinv = 1.0f / piv;
		__syncwarp();
		float p0 = 0.f, p1 = 0.f;
		if (h0) {
			p0 = (i0 == k) ? dinv : W[k * ld + i0] * dinv;
			W[k * ld + i0] = p0;
		}
		if (h1) {
			p1 = (i1 == k) ? dinv : W[k * ld + i1] * dinv;
			W[k * ld + i1] = p1;
		}
		if (h0) col[i0] = W[i0 * ld + k];
		if (h1) col[i1] = W[i1 * ld + k];
		__syncwarp();
		for (int i = 0; i < n; ++i) {
			if (i == k)
				continue;
			const float f = col[i];
			if (h0)
				W[i * ld + i0] = (i0 == k) ? -f * p0 : fmaf(-f, p0, W[i * ld + i0]);
			if (h1)
				W[i * ld + i1] = (i1 == k) ? -f * p1 : fmaf(-f, p1, W[i * ld + i1]);
		}
		__syncwarp();
	}
	return bad;
}

template <int T>
__global__ void __launch_bounds__(GEN_WARPS * 32) qp_gen_kernel(const GenArgs a, unsigned long long *ticket)
{
	constexpr int G = 32 / T; // problems per warp
	extern __shared__ __align__(16) float gsm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int grp = lane / T, g = lane % T;
	const int n = a.n, m = a.m, ld = gen_ld(n);
	float *R = gsm + (warp * G + grp) * gen_smem_floats_per_group(n, m, T); // P + rho A'A, then its inverse
	float *As = R + n * ld;                                                  // A, row stride ld
	float *vn0 = As + m * ld, *vn1 = vn0 + 2 * T;                     // n-vectors (broadcast buffers)
	float *vm0 = vn1 + 4 * T, *vm1 = vm0 + GEN_RPL * T;               // m-vectors
	const int i0 = g, i1 = g + T;
	const bool h0 = i0 < n, h1 = i1 < n;
	const float rho = a.rho, alpha = a.alpha;

	// first round by position, then (batches longer than one wave) G problems per ticket, the
	// next one drawn while these are solved -- see qp_genr.cuh
	const unsigned long long nw = (unsigned long long)gridDim.x * GEN_WARPS * G;
	unsigned long long base = ((unsigned long long)blockIdx.x * GEN_WARPS + warp) * G;
	for (; base < a.count;) {
		unsigned long long next = base + nw;
		if (ticket) {
			if (lane == 0)
				next = nw + atomicAdd(ticket, (unsigned long long)G);
		}
		const bool valid = base + grp < a.count;
		const unsigned long long prob = valid ? base + grp : base; // a tail group repeats the first problem
		// ---- load P, A (coalesced inside the group), q, l, u, warm start ---------------------
		const float *Pp = a.P + prob * (unsigned long long)n * n;
		const float *Ap = a.A + prob * (unsigned long long)m * n;
		for (int e = g; e < n * n; e += T) {
			const int r = e / n, c = e - r * n;
			R[r * ld + c] = __ldg(Pp + e);
		}
		for (int e = g; e < m * n; e += T) {
			const int r = e / n, c = e - r * n;
			As[r * ld + c] = __ldg(Ap + e);
		}
		const unsigned long long vo = prob * (unsigned long long)n, mo = prob * (unsigned long long)m;
		const float q0 = h0 ? __ldg(a.q + vo + i0) : 0.f, q1 = h1 ? __ldg(a.q + vo + i1) : 0.f;
		float lo[GEN_RPL], hi[GEN_RPL], z[GEN_RPL], u[GEN_RPL];
#pragma unroll
		for (int k = 0; k < GEN_RPL; ++k) {
			const int j = g + T * k;
			const bool in = j < m;
			lo[k] = in ? __ldg(a.l + mo + j) : 0.f;
			hi[k] = in ? __ldg(a.u + mo + j) : 0.f;
			z[k] = (in && a.z0) ? __ldg(a.z0 + mo + j) : 0.f;
			u[k] = (in && a.u0) ? __ldg(a.u0 + mo + j) : 0.f;
		}
		__syncwarp();

		// ---- R <- P + rho A'A (members own columns), then its inverse -------------------------
		for (int i = 0; i < n; ++i) {
			float s0 = 0.f, s1 = 0.f;
			for (int k = 0; k < m; ++k) {
				const float aki = As[k * ld + i];
				if (h0) s0 = fmaf(aki, As[k * ld + i0], s0);
				if (h1) s1 = fmaf(aki, As[k * ld + i1], s1);
			}
			if (h0) R[i * ld + i0] = fmaf(rho, s0, R[i * ld + i0]);
			if (h1) R[i * ld + i1] = fmaf(rho, s1, R[i * ld + i1]);
		}
		__syncwarp();
		const bool bad = smem_gauss_jordan_group<T>(R, n, ld, vn0, g);

		// ---- helpers ----------------------------------------------------------------------------
		// rows i0, i1 of R times an n-vector in shared memory
		auto r_times = [&](const float *v, float &o0, float &o1) {
			float s0 = 0.f, s1 = 0.f;
			for (int j = 0; j < n; ++j) {
				const float vj = v[j];
				if (h0) s0 = fmaf(R[i0 * ld + j], vj, s0);
				if (h1) s1 = fmaf(R[i1 * ld + j], vj, s1);
			}
			o0 = s0, o1 = s1;
		};
		// coordinates i0, i1 of A' v and A' w in one pass over A
		auto at_times2 = [&](const float *v, const float *w, float &a0, float &a1, float &b0,
				     float &b1) {
			float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
			for (int k = 0; k < m; ++k) {
				const float vk = v[k], wk = w[k];
				if (h0) {
					const float e = As[k * ld + i0];
					s0 = fmaf(e, vk, s0), t0 = fmaf(e, wk, t0);
				}
				if (h1) {
					const float e = As[k * ld + i1];
					s1 = fmaf(e, vk, s1), t1 = fmaf(e, wk, t1);
				}
			}
			a0 = s0, a1 = s1, b0 = t0, b1 = t1;
		};

		// A'z and A'u of the starting point
		float w1a = 0.f, w1b = 0.f, w2a = 0.f, w2b = 0.f;
		if (a.z0 || a.u0) {
#pragma unroll
			for (int k = 0; k < GEN_RPL; ++k)
				if (g + T * k < m)
					vm0[g + T * k] = z[k], vm1[g + T * k] = u[k];
			__syncwarp();
			at_times2(vm0, vm1, w1a, w1b, w2a, w2b);
			__syncwarp();
		}

		// ---- iterations: the warp's groups in lockstep ------------------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		float x0 = 0.f, x1 = 0.f;
		bool active = !bad && a.max_iter > 0; // uniform inside a group
		while (__any_sync(0xffffffffu, active)) {
			// x-update: (P + rho A'A) x = rho A'(z - u) - q
			if (h0) vn0[i0] = fmaf(rho, w1a - w2a, -q0);
			if (h1) vn0[i1] = fmaf(rho, w1b - w2b, -q1);
			__syncwarp();
			float nx0, nx1;
			r_times(vn0, nx0, nx1);
			if (h0) vn1[i0] = nx0;
			if (h1) vn1[i1] = nx1;
			__syncwarp();
			// rows of A x, relaxation, clip, dual update
			float rr = 0.f, nax = 0.f, nz = 0.f;
			float zn[GEN_RPL], un[GEN_RPL];
#pragma unroll
			for (int k = 0; k < GEN_RPL; ++k) {
				const int j = g + T * k;
				zn[k] = z[k], un[k] = u[k];
				if (j < m) {
					float ax = 0.f;
					for (int c = 0; c < n; ++c)
						ax = fmaf(As[j * ld + c], vn1[c], ax);
					const float zh = fmaf(alpha, ax, (1.0f - alpha) * z[k]);
					zn[k] = fminf(fmaxf(zh + u[k], lo[k]), hi[k]);
					un[k] = u[k] + (zh - zn[k]);
					const float dr = ax - zn[k];
					rr = fmaf(dr, dr, rr);
					nax = fmaf(ax, ax, nax);
					nz = fmaf(zn[k], zn[k], nz);
					vm0[j] = zn[k], vm1[j] = un[k];
				}
			}
			__syncwarp();
			float n1a, n1b, n2a, n2b;
			at_times2(vm0, vm1, n1a, n1b, n2a, n2b); // A'z, A'u
			const float da = n1a - w1a, db = n1b - w1b;
			float ss = fmaf(da, da, db * db), nu = fmaf(n2a, n2a, n2b * n2b);
			rr = group_sum<T>(rr), nax = group_sum<T>(nax), nz = group_sum<T>(nz);
			ss = group_sum<T>(ss), nu = group_sum<T>(nu);
			const float eps_pri = fmaf(a.reltol, sqrtf(fmaxf(nax, nz)), a.eps_abs_pri);
			const float eps_dual = fmaf(a.reltol * rho, sqrtf(nu), a.eps_abs_dual);
			__syncwarp();
			if (active) { // commit
				x0 = nx0, x1 = nx1;
#pragma unroll
				for (int k = 0; k < GEN_RPL; ++k)
					z[k] = zn[k], u[k] = un[k];
				w1a = n1a, w1b = n1b, w2a = n2a, w2b = n2b;
				++it;
				if (sqrtf(rr) < eps_pri && rho * sqrtf(ss) < eps_dual) {
					status = QP_STATUS_CONVERGED;
					active = false;
				} else if (it >= a.max_iter) {
					active = false;
				}
			}
		}

		// ---- results ----------------------------------------------------------------------------------
		const float nf = group_sum<T>(fmaf(x0, 0.0f, x1 * 0.0f));
		int iters = int(it);
		if (bad)
			status = QP_STATUS_SINGULAR, iters = 0;
		else if (!(nf == 0.0f))
			status = QP_STATUS_NONFINITE;
		if (valid) {
			if (h0) a.x[vo + i0] = x0;
			if (h1) a.x[vo + i1] = x1;
#pragma unroll
			for (int k = 0; k < GEN_RPL; ++k) {
				const int j = g + T * k;
				if (j < m) {
					if (a.z) a.z[mo + j] = z[k];
					if (a.y) a.y[mo + j] = rho * u[k];
				}
			}
			if (g == 0) {
				if (a.status) a.status[prob] = status;
				if (a.iters) a.iters[prob] = iters;
			}
		}
		__syncwarp();
		if (ticket)
			next = __shfl_sync(0xffffffffu, next, 0);
		base = next;
	}
}

} // namespace qpb
