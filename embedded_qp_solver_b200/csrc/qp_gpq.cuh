// qp_gpq.cuh -- "lane group per QP" kernel for any n <= 64 (n = 40 and 64 in BASELINE.json).
//
// What it replaces: the same reference path as qp_tpq.cuh -- admm()
// (/root/reference/qp_solvers/qp_solvers.c:255-319) with matrix_invert()
// (/root/reference/matrix/matrix_ops/matrix_ops.c:551-630) and matrix_mult()
// (:235-271) underneath.
//
// Mapping:
//   * a group of T lanes (T = 8, 16 or 32; 32/T problems per warp) owns one QP;
//     lane g of the group owns rows g, g+T, g+2T, ... of the matrix, held entirely
//     in registers (NP floats per row, NP = n rounded up to 8; padding rows and
//     columns form an identity block).
//   * Loading uses the symmetry of P: the lane that owns row r reads COLUMN r, so
//     each step of the load loop reads T consecutive floats per group.
//   * R = (P + rho I)^-1 is formed explicitly, like the reference, because the
//     per-iteration x = R y is then a register-resident mat-vec with y broadcast
//     from shared memory; triangular solves would serialise across lanes.
//   * The inverse is a BLOCK Gauss-Jordan elimination without pivoting (P + rho I
//     is symmetric positive definite, so every 4x4 pivot block is too): four pivot
//     rows are published to shared memory at once, every lane inverts the 4x4
//     pivot block redundantly in registers, and each row is updated with four
//     FMAs per element.  The elimination is written in "shifting" form -- each
//     block step drops the four pivot columns and appends the four new inverse
//     columns at the end -- so every register index is a compile-time constant
//     while the block loop stays a run-time loop.  The pivot rows are double
//     buffered: one __syncwarp per four pivots.
//   * per iteration: y to shared memory, NP FMAs per owned row, element-wise
//     updates on the owned coordinates, one butterfly all-reduce (inside the
//     group) of the five sums of squares, and a group-uniform stop test.
//   * x, z, lambda are written with lane = coordinate: coalesced per group.
//   * Arithmetic is issued as packed FP32 pairs (fma.rn.f32x2, SASS FFMA2): a row
//     lives in registers as NP/2 float2, the elimination multiplier is the scalar
//     broadcast operand of FFMA2 and the pivot rows arrive from shared memory as
//     128-bit loads that are already two aligned pairs.  FFMA2 neither raises the
//     FP32 peak nor saves issue slots (one FFMA2 takes two, measured:
//     profiles/r1_probe_ffma2_issue.txt); what it saves is instruction fetch/decode
//     and register-file reads of the shared multiplier (profiles/README.md).
#pragma once

#include "qp_common.cuh"

namespace qpb {

#ifndef QPB_GPQ_SEG
#define QPB_GPQ_SEG 4
#endif
#ifndef QPB_GPQ_MINB_BIG
#define QPB_GPQ_MINB_BIG 2
#endif
#ifndef QPB_GPQ_WARPS
#define QPB_GPQ_WARPS 4
#endif
constexpr int GPQ_WARPS = QPB_GPQ_WARPS; // warps per CTA

// lanes per QP for a padded dimension
template <int NP> struct GpqLanes {
	static constexpr int T = (NP <= 32) ? 8 : (NP <= 48 ? 16 : 32);
};

template <int NP, int T> struct GpqSmem {
	static constexpr int G = 32 / T; // QPs per warp
	// pivot rows are NP + 4 floats apart so that the four owners' 128-bit stores of
	// one column chunk fall into different bank groups; two buffers of four rows per
	// group, +4 floats so that the groups of a warp are staggered as well
	static constexpr int ROW = NP + 4;
	static constexpr int GROUP = 2 * 4 * ROW + 4;
	static constexpr int WARP = G * GROUP;
};

// CTAs per SM the register budget is capped for (rows in registers: NP * RPL floats)
template <int NP, int T> struct GpqOcc {
	static constexpr int ROWREGS = NP * ((NP + T - 1) / T);
	static constexpr int MINB = ROWREGS <= 64 ? 6 : (ROWREGS <= 96 ? 4 : QPB_GPQ_MINB_BIG);
};

// packed-pair helpers: d = a * (s, s) + c and friends (SASS FFMA2 with a scalar
// broadcast operand; no instruction is spent on building the (s, s) pair)
__device__ __forceinline__ float2 fma2_bc(float2 a, float s, float2 c)
{
	return __ffma2_rn(a, make_float2(s, s), c);
}

// one row (NH pairs) to shared memory as 128-bit stores
template <int NH> __device__ __forceinline__ void publish_row(const float2 (&row)[NH], float *dst)
{
#pragma unroll
	for (int c = 0; c < NH / 2; ++c)
		*reinterpret_cast<float4 *>(dst + 4 * c) =
			make_float4(row[2 * c].x, row[2 * c].y, row[2 * c + 1].x, row[2 * c + 1].y);
}

// row <- -e_m0.  The constants come from shared memory (ctab[0..3] = rows of -I_4,
// ctab[4] = 0): one 128-bit load fills four registers, a register move fills one
// (measured: 142 vs 138 M QP/s at n = 40; -DQPB_GPQ_MOV_FILL selects the moves)
template <int NH>
__device__ __forceinline__ void pivot_row_reset(float2 (&row)[NH], const float4 *ctab, int m0)
{
#ifndef QPB_GPQ_MOV_FILL
	const float4 e = lds128_volatile(&ctab[m0]);
	row[0] = make_float2(e.x, e.y), row[1] = make_float2(e.z, e.w);
#pragma unroll
	for (int h = 2; h < NH; h += 2) {
		const float4 zz = lds128_volatile(&ctab[4]);
		row[h] = make_float2(zz.x, zz.y), row[h + 1] = make_float2(zz.z, zz.w);
	}
#else
	(void)ctab;
	row[0] = make_float2(m0 == 0 ? -1.0f : 0.0f, m0 == 1 ? -1.0f : 0.0f);
	row[1] = make_float2(m0 == 2 ? -1.0f : 0.0f, m0 == 3 ? -1.0f : 0.0f);
#pragma unroll
	for (int h = 2; h < NH; ++h)
		row[h] = make_float2(0.0f, 0.0f);
#endif
}

// the same restricted to the first `live` 4-column chunks (product-form variant: the
// chunks behind them hold finished eta columns that must survive)
template <int NH>
__device__ __forceinline__ void publish_row_live(const float2 (&row)[NH], float *dst, int live)
{
#pragma unroll
	for (int c = 0; c < NH / 2; ++c)
		if (c < live)
			*reinterpret_cast<float4 *>(dst + 4 * c) =
				make_float4(row[2 * c].x, row[2 * c].y, row[2 * c + 1].x, row[2 * c + 1].y);
}

template <int NH>
__device__ __forceinline__ void pivot_row_reset_live(float2 (&row)[NH], const float4 *ctab, int m0,
						     int live)
{
	const float4 e = lds128_volatile(&ctab[m0]);
	row[0] = make_float2(e.x, e.y), row[1] = make_float2(e.z, e.w);
#pragma unroll
	for (int h = 2; h < NH; h += 2) {
		if (h / 2 < live) {
			const float4 zz = lds128_volatile(&ctab[4]);
			row[h] = make_float2(zz.x, zz.y), row[h + 1] = make_float2(zz.z, zz.w);
		}
	}
}

template <int NP, int T, bool PFI>
__global__ void __launch_bounds__(GPQ_WARPS * 32, GpqOcc<NP, T>::MINB)
qp_gpq_kernel(const KernelArgs p)
{
	using SM = GpqSmem<NP, T>;
	constexpr int G = SM::G;
	constexpr int RPL = (NP + T - 1) / T; // rows per lane
	constexpr int NC = NP / 4;            // 4-column chunks per row
	constexpr int NH = NP / 2;            // float2 pairs per row
	constexpr int ROW = SM::ROW;          // floats between published pivot rows
	constexpr unsigned FULL = 0xffffffffu;
	__shared__ __align__(16) float sm[GPQ_WARPS * SM::WARP];
	__shared__ float4 ctab[5]; // rows of -I_4, then a zero vector

	if (threadIdx.x < 5) {
		const int t = threadIdx.x;
		ctab[t] = make_float4(t == 0 ? -1.0f : 0.0f, t == 1 ? -1.0f : 0.0f, t == 2 ? -1.0f : 0.0f,
				      t == 3 ? -1.0f : 0.0f);
	}
	__syncthreads();

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int g = lane & (T - 1); // lane inside its group
	const int grp = lane / T;     // group inside the warp
	const int n = p.n;
	float *bc = sm + warp * SM::WARP + grp * SM::GROUP; // bc[(buf*4+m)*ROW + j]
	const float alpha = p.alpha;
	const bool relax = (alpha != 1.0f);

	const unsigned long long stride = (unsigned long long)gridDim.x * GPQ_WARPS * G;
	for (unsigned long long base = ((unsigned long long)blockIdx.x * GPQ_WARPS + warp) * G;
	     base < p.count; base += stride) {
		const unsigned long long prob = base + grp;
		const bool valid = prob < p.count;
		// penalty of this group's problem (qp_batch.rho)
		const float rho = (p.rho_v && valid) ? __ldg(p.rho_v + prob) : p.rho;
		const float rho2 = rho * rho, relrho = p.reltol * rho;

		// ---- load: the owner of row r takes column r of the symmetric P ------------------
		// S[s][h] = (row element 2h, row element 2h+1) of row g + T*s
		float2 S[RPL][NH];
		const float *Pp = p.P + (valid ? p_index(p, prob) : 0ull) * (unsigned long long)n * n;
		if (n == NP) {
			// no padding: offsets are compile-time, no per-element bounds tests
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				const int r = g + T * s;
				const bool in = valid && r < NP;
				const float *col = Pp + (in ? r : 0);
#pragma unroll
				for (int j = 0; j < NP; ++j) {
					float v = in ? __ldg(col + j * NP) : 0.0f;
					if (r == j)
						v = in ? v + rho : 1.0f;
					if (j & 1)
						S[s][j / 2].y = v;
					else
						S[s][j / 2].x = v;
				}
			}
		} else {
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				const int r = g + T * s;
#pragma unroll
				for (int j = 0; j < NP; ++j) {
					float v = 0.0f;
					if (valid && r < n && j < n)
						v = __ldg(Pp + (unsigned long long)j * n + r);
					if (r == j)
						v = (valid && r < n) ? v + rho : 1.0f; // padding: identity block
					if (j & 1)
						S[s][j / 2].y = v;
					else
						S[s][j / 2].x = v;
				}
			}
		}

		// ---- R = (P + rho I)^-1: block Gauss-Jordan, four pivots per step -----------------
		// The block loop is split by the register slot ks that holds the pivot rows, so
		// that "is this my pivot row" touches one slot's registers, not all of them.
		//
		// Look-ahead: the latency chain of a step -- publish the pivot rows, wait, load the
		// 4x4 pivot block, invert it (four dependent reciprocals) -- is taken off the
		// critical path.  Step K updates the chunk that holds the NEXT pivot columns first,
		// the next owners publish just their 4x4 block, and its inverse is computed while
		// the remaining chunks of step K are being updated.  At the end of step K the next
		// owners publish their complete rows, so step K+1 starts with everything it needs.
		bool bad = !(rho > 0.0f && rho < 3.0e38f);
		Ldl4 Bf; // L D L^T of the current pivot block
		{
			if (g < 4)
				publish_row<NH>(S[0], bc + g * ROW);
			__syncwarp();
			float Bi[4][4];
#pragma unroll
			for (int m = 0; m < 4; ++m) {
				const float4 v = *reinterpret_cast<const float4 *>(bc + m * ROW);
				Bi[m][0] = v.x, Bi[m][1] = v.y, Bi[m][2] = v.z, Bi[m][3] = v.w;
			}
			bad |= ldl4_factor(Bi, Bf);
		}
		// The run-time block loop is cut into compile-time segments: one per register slot,
		// and in the product-form variant one per four steps, so that the number of chunks
		// that are still matrix columns is known per segment up to three.
		constexpr int KPS = T / 4;                   // block steps per slot
		constexpr int SEG = PFI ? (KPS < QPB_GPQ_SEG ? KPS : QPB_GPQ_SEG) : KPS; // block steps per segment
		constexpr int NSEG = (NC + SEG - 1) / SEG;
#pragma unroll
		for (int seg = 0; seg < NSEG; ++seg) {
			const int ks = (seg * SEG) / KPS;            // slot that holds this segment's pivot rows
			const int ns = (ks + 1 < RPL) ? ks + 1 : ks; // slot after this one
			const int Kseg = (seg + 1) * SEG < NC ? (seg + 1) * SEG : NC;
			const int Kslot = (ks + 1) * KPS < NC ? (ks + 1) * KPS : NC;
			// live chunks L = NC - K in this segment: Lmin..Lmax
			const int Lmax = PFI ? NC - seg * SEG : NC, Lmin = PFI ? NC - (Kseg - 1) : NC;
			for (int K = seg * SEG; K < Kseg; ++K) {
				const int m0 = g - (4 * K - ks * T); // my pivot row inside the block, if any
				const bool owner = (m0 >= 0) && (m0 < 4);
				const float *pb = bc + (K & 1) * 4 * ROW;
				float *pn = bc + ((K + 1) & 1) * 4 * ROW;
				// next step: same slot, or the first four lanes of the next slot
				const bool same_slot = K + 1 < Kslot;
				const bool more = K + 1 < NC;
				const int m1 = same_slot ? m0 - 4 : g;
				const bool nowner = more && (m1 >= 0) && (m1 < 4);
				// product form: only the first L chunks are still matrix columns; position
				// L-1 receives this step's eta columns H and is never touched again
				const int L = PFI ? NC - K : NC;
				// my row was published at the end of the previous step: turn it into -e_m0,
				// which makes the common update below produce (row m0 of B^-1) x (pivot rows)
				if (owner) {
					if constexpr (PFI)
						pivot_row_reset_live<NH>(S[ks], ctab, m0, L);
					else
						pivot_row_reset<NH>(S[ks], ctab, m0);
				}
				// H = -(first four columns) x B^-1 by two triangular solves per row, two
				// pairs per row
				float2 H[RPL][2];
#pragma unroll
				for (int s = 0; s < RPL; ++s)
					ldl4_minus_solve(Bf, S[s][0].x, S[s][0].y, S[s][1].x, S[s][1].y, H[s][0].x,
							 H[s][0].y, H[s][1].x, H[s][1].y);
				// S(:, j-4) <- S(:, j) + H * pivot_rows(:, j), one 4-column chunk at a time
				auto update_chunk = [&](int c) {
					float2 lo[4], hi[4];
#pragma unroll
					for (int m = 0; m < 4; ++m) {
						const float4 v =
							*reinterpret_cast<const float4 *>(pb + m * ROW + 4 * c);
						lo[m] = make_float2(v.x, v.y), hi[m] = make_float2(v.z, v.w);
					}
#pragma unroll
					for (int s = 0; s < RPL; ++s) {
						float2 a = S[s][2 * c], b = S[s][2 * c + 1];
						a = fma2_bc(lo[0], H[s][0].x, a), b = fma2_bc(hi[0], H[s][0].x, b);
						a = fma2_bc(lo[1], H[s][0].y, a), b = fma2_bc(hi[1], H[s][0].y, b);
						a = fma2_bc(lo[2], H[s][1].x, a), b = fma2_bc(hi[2], H[s][1].x, b);
						a = fma2_bc(lo[3], H[s][1].y, a), b = fma2_bc(hi[3], H[s][1].y, b);
						S[s][2 * c - 2] = a, S[s][2 * c - 1] = b;
					}
				};
				auto place_eta = [&](int c) { // chunk c <- H
#pragma unroll
					for (int s = 0; s < RPL; ++s)
						S[s][2 * c] = H[s][0], S[s][2 * c + 1] = H[s][1];
				};
				if (1 < Lmin || 1 < L)
					update_chunk(1);
				else
					place_eta(0);
				// the next pivot block is complete: publish it and start on its inverse
				if (nowner) {
					const float2 a = same_slot ? S[ks][0] : S[ns][0];
					const float2 b = same_slot ? S[ks][1] : S[ns][1];
					*reinterpret_cast<float4 *>(pn + m1 * ROW) = make_float4(a.x, a.y, b.x, b.y);
				}
				__syncwarp();
				float Bn[4][4];
#pragma unroll
				for (int m = 0; m < 4; ++m) {
					const float4 v = *reinterpret_cast<const float4 *>(pn + m * ROW);
					Bn[m][0] = v.x, Bn[m][1] = v.y, Bn[m][2] = v.z, Bn[m][3] = v.w;
				}
				Ldl4 Bfn;
				const bool badn = ldl4_factor(Bn, Bfn);
#pragma unroll
				for (int c = 2; c < NC; ++c) {
					if (c < Lmin) // live in every step of this segment: no test
						update_chunk(c);
					else if (c < Lmax) {
						if (c < L)
							update_chunk(c);
						else if (c == L)
							place_eta(c - 1);
					} else if (c == Lmax && c == L)
						place_eta(c - 1);
				}
				if (Lmax == NC && L == NC)
					place_eta(NC - 1);
				// the next owners publish their complete rows
				if (nowner) {
					if constexpr (PFI) {
						if (same_slot)
							publish_row_live<NH>(S[ks], pn + m1 * ROW, L - 1);
						else
							publish_row_live<NH>(S[ns], pn + m1 * ROW, L - 1);
					} else {
						if (same_slot)
							publish_row<NH>(S[ks], pn + m1 * ROW);
						else
							publish_row<NH>(S[ns], pn + m1 * ROW);
					}
				}
				__syncwarp();
				bad |= more && badn;
				Bf = Bfn;
			}
		}
		__syncwarp();

		// ---- per-coordinate state -----------------------------------------------------------
		float q[RPL], lo[RPL], hi[RPL], z[RPL], u[RPL], x[RPL];
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = g + T * s;
			const bool in = valid && r < n;
			const unsigned long long o = prob * (unsigned long long)n + r;
			q[s] = in ? __ldg(p.q + o) : 0.0f;
			// lb/ub arrays or the scalar box; z0/u0 or zero (group-uniform choices)
			lo[s] = in ? (p.lb ? __ldg(p.lb + o) : p.lb_s) : 0.0f;
			hi[s] = in ? (p.ub ? __ldg(p.ub + o) : p.ub_s) : 0.0f;
			z[s] = (in && p.z0) ? __ldg(p.z0 + o) : 0.0f;
			u[s] = (in && p.u0) ? __ldg(p.u0 + o) : 0.0f;
			x[s] = 0.0f;
		}

		// ---- ADMM iterations (qp_solvers.c:284-310) -----------------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		const bool active = valid && !bad && p.max_iter > 0; // uniform inside a group
		// element-wise part of one iteration with x = xr: admm_update_x_hat (:161-174),
		// _z (:176-190), _u (:192-203) and the five partial sums of squares (:205-253)
		auto elementwise = [&](const float (&xr)[RPL], float &rr, float &ss, float &nx, float &nz,
				       float &nu) {
			rr = ss = nx = nz = nu = 0.f;
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				const float xs = xr[s];
				x[s] = xs;
				const float xh = relax ? fmaf(alpha, xs, (1.0f - alpha) * z[s]) : xs;
				const float zn = fminf(fmaxf(xh + u[s], lo[s]), hi[s]);
				const float dz = zn - z[s];
				ss = fmaf(dz, dz, ss);
				u[s] += xh - zn;
				const float dr = xs - zn;
				rr = fmaf(dr, dr, rr);
				nx = fmaf(xs, xs, nx);
				nz = fmaf(zn, zn, nz);
				nu = fmaf(u[s], u[s], nu);
				z[s] = zn;
			}
		};
		auto reduce5 = [&](float &rr, float &ss, float &nx, float &nz, float &nu, unsigned am) {
#pragma unroll
			for (int o = T / 2; o > 0; o >>= 1) {
				rr += __shfl_xor_sync(am, rr, o);
				ss += __shfl_xor_sync(am, ss, o);
				nx += __shfl_xor_sync(am, nx, o);
				nz += __shfl_xor_sync(am, nz, o);
				nu += __shfl_xor_sync(am, nu, o);
			}
		};
		// stop test (:233-253, :308), squared form as in qp_tpq.cuh
		auto converged = [&](float rr, float ss, float nx, float nz, float nu) {
			const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
			const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
			return (rr < eps_pri * eps_pri) && (rho2 * ss < eps_dual * eps_dual);
		};

		// Every warp-wide operation of the loops below names only the lanes of the group
		// (am), so a group simply leaves its loop when its problem is done and waits for
		// the other groups at the __syncwarp() behind it: no ballot, no flag to test.
		const unsigned am = (T == 32) ? FULL : (((1u << (T & 31)) - 1u) << (lane & ~(T - 1)));
		bool conv = false;
		if constexpr (PFI) {
		if (active) {
			do {
				// admm_update_x (:146-159): y = rho (z - u) - q ; x = R y.
				// R is held in product form, R = E_{NC-1} ... E_0: apply the block eta
				// transformations one after the other.  E_K maps the four pivot entries
				// v_P to B^-1 v_P and adds H_K v_P to every other entry; H_K sits in chunk
				// NC-1-K, and the pivot rows' H is B^-1 itself.
				float xr[RPL];
#pragma unroll
				for (int s = 0; s < RPL; ++s)
					xr[s] = fmaf(rho, z[s] - u[s], -q[s]);
#pragma unroll
				for (int K = 0; K < NC; ++K) {
					const int slot = (4 * K) / T, lane0 = (4 * K) % T;
					const int src = (lane & ~(T - 1)) + lane0;
					const float p0 = __shfl_sync(am, xr[slot], src);
					const float p1 = __shfl_sync(am, xr[slot], src + 1);
					const float p2 = __shfl_sync(am, xr[slot], src + 2);
					const float p3 = __shfl_sync(am, xr[slot], src + 3);
					const bool mine = (g >= lane0) && (g < lane0 + 4);
#pragma unroll
					for (int s = 0; s < RPL; ++s) {
						const float2 h01 = S[s][2 * (NC - 1 - K)], h23 = S[s][2 * (NC - 1 - K) + 1];
						const float b = (s == slot && mine) ? 0.0f : xr[s];
						xr[s] = fmaf(h01.x, p0, fmaf(h01.y, p1, fmaf(h23.x, p2, fmaf(h23.y, p3, b))));
					}
				}
				float rr, ss, nx, nz, nu;
				elementwise(xr, rr, ss, nx, nz, nu);
				reduce5(rr, ss, nx, nz, nu, am);
				++it;
				conv = converged(rr, ss, nx, nz, nu);
			} while (!conv && it < p.max_iter);
		}
		} else {
		// Full inverse: x = R y is a register-resident mat-vec with y broadcast from shared
		// memory.  The loop is software-pipelined: the mat-vec of the NEXT iteration is issued
		// before the five reductions and the stop decision of this one are known (it reads
		// only z and u), so the shuffle chains overlap the FMAs.  One mat-vec per problem is
		// wasted (the one after convergence).  y alternates between two buffers, so one
		// barrier per iteration is enough.
		auto publish_y = [&](float *yb) {
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				const int r = g + T * s;
				if (r < NP)
					yb[r] = fmaf(rho, z[s] - u[s], -q[s]);
			}
		};
		auto matvec = [&](const float *yb, float (&xo)[RPL]) {
			float2 acc[RPL][2];
#pragma unroll
			for (int s = 0; s < RPL; ++s)
				acc[s][0] = acc[s][1] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < NP; j += 4) {
				const float4 y = *reinterpret_cast<const float4 *>(yb + j);
				const float2 ylo = make_float2(y.x, y.y), yhi = make_float2(y.z, y.w);
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
					acc[s][0] = __ffma2_rn(S[s][j / 2], ylo, acc[s][0]);
					acc[s][1] = __ffma2_rn(S[s][j / 2 + 1], yhi, acc[s][1]);
				}
			}
#pragma unroll
			for (int s = 0; s < RPL; ++s)
				xo[s] = (acc[s][0].x + acc[s][0].y) + (acc[s][1].x + acc[s][1].y);
		};
		float xr[RPL];
#pragma unroll
		for (int s = 0; s < RPL; ++s)
			xr[s] = 0.0f;
		if (active) {
			publish_y(bc);
			__syncwarp(am);
			matvec(bc, xr);
			do {
				float rr, ss, nx, nz, nu;
				elementwise(xr, rr, ss, nx, nz, nu);
				++it;
				float *yb = bc + (it & 1u) * 4 * ROW;
				publish_y(yb);
				__syncwarp(am);
				reduce5(rr, ss, nx, nz, nu, am);
				matvec(yb, xr);
				conv = converged(rr, ss, nx, nz, nu);
			} while (!conv && it < p.max_iter);
		}
		}
		__syncwarp();
		if (conv)
			status = QP_STATUS_CONVERGED;

		// ---- results ---------------------------------------------------------------------------
		float nf = 0.0f;
#pragma unroll
		for (int s = 0; s < RPL; ++s)
			nf = fmaf(x[s], 0.0f, nf);
#pragma unroll
		for (int o = T / 2; o > 0; o >>= 1)
			nf += __shfl_xor_sync(FULL, nf, o);
		int iters = int(it);
		if (bad) {
			status = QP_STATUS_SINGULAR;
			iters = 0;
		} else if (!(nf == 0.0f)) {
			status = QP_STATUS_NONFINITE;
		}
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = g + T * s;
			if (valid && r < n) {
				const unsigned long long o = prob * (unsigned long long)n + r;
				if (p.x)
					p.x[o] = x[s];
				if (p.z)
					p.z[o] = z[s];
				if (p.lam)
					p.lam[o] = rho * u[s];
			}
		}
		if (valid && g == 0) {
			if (p.status)
				p.status[prob] = status;
			if (p.iters)
				p.iters[prob] = iters;
		}
	}
}

} // namespace qpb
