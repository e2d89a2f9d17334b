// qp_bpq.cuh -- "four rows per lane" kernel for n = 8, 16, 24, 32, 40 (n = 40 in BASELINE.json).
//
// What it replaces: the same reference path as qp_gpq.cuh -- admm()
// (/root/reference/qp_solvers/qp_solvers.c:255-319) with matrix_invert()
// (/root/reference/matrix/matrix_ops/matrix_ops.c:551-630) and matrix_mult()
// (:235-271) underneath.
//
// Why a second lane-group kernel: qp_gpq.cuh spreads rows cyclically over a power-of-two
// group (16 lanes at n = 40, three row slots of which the third is half empty: 40/48 of
// the FMAs are useful) and every shared-memory pivot value feeds RPL = 3 FMAs.  Here
//   * lane g of a group of T = n/4 lanes owns the four CONSECUTIVE rows 4g..4g+3.  At
//     n = 40 that is 10 lanes per QP and three QPs per warp (lanes 30, 31 idle): 30/32 of
//     the FMAs are useful, and every pivot value loaded from shared memory feeds four FMAs;
//   * registers hold ROW PAIRS: S[rp][j] = (row 4g+2rp, row 4g+2rp+1) at column j.  One
//     128-bit load of P[j][4g..4g+3] is then exactly the two pairs of column j, one
//     128-bit shared-memory store publishes the four pivot rows at column j, and the
//     elimination is FFMA2 with the pair of multipliers (H) in registers and the pivot
//     value as the scalar broadcast operand;
//   * a 4x4 pivot block lies in ONE lane, all four register slots, so the block loop needs
//     no per-slot split.  The owner of the next step publishes its rows chunk by chunk
//     while the current step finishes them (position-major buffer: one 128-bit store per
//     column), then turns its rows into -I_4 so that the common update produces
//     B^-1 x (pivot rows) for it.  The multipliers -C B^-1 come from two triangular
//     solves with the L D L^T of the pivot block, not from an explicit inverse;
//   * loads, stores and the per-coordinate vectors are 128-bit per lane: lane g reads
//     P[j][4g..4g+3] (column block of the symmetric P = row block), q/lb/ub/z0/u0[4g..4g+3],
//     and writes x/z/lambda[4g..4g+3];
//   * groups of 10 or 6 lanes are not a power of two: the five sums of squares are
//     all-reduced with a fold by T/2 followed by a doubling ring over the odd remainder,
//     and the stop decision is taken by lane 0 of the group and broadcast, so a group can
//     never disagree with itself.
// Same block Gauss-Jordan in shifting form and the same packed FP32 pairs (FFMA2) as
// qp_gpq.cuh.  Requires n == NP (a multiple of 8) and 16-byte aligned arrays; everything
// else goes to qp_gpq.cuh.
#pragma once

#include "qp_common.cuh"
#include "qp_gpq.cuh"

namespace qpb {

#ifndef QPB_BPQ_WARPS
#define QPB_BPQ_WARPS 4
#endif
constexpr int BPQ_WARPS = QPB_BPQ_WARPS; // warps per CTA

template <int NP> struct BpqCfg {
	static constexpr int T = NP / 4;           // lanes per QP
	static constexpr int G = 32 / T;           // QPs per warp
	static constexpr int GG = G + ((32 % T) ? 1 : 0); // + a scratch region for the idle lanes
	static constexpr int ROW = NP + 4;         // floats between published pivot rows
	static constexpr int BUF = 4 * ROW;        // one buffer: four pivot rows
	static constexpr int GROUP = 2 * BUF + 4;  // two buffers, groups staggered by 16 bytes
	static constexpr int MINB = NP <= 8 ? 6 : (NP <= 16 ? 4 : (NP <= 24 ? 3 : 2));
};

// all-reduce over the T lanes of a group (lanes base .. base+T-1, this lane is base + g).
// Power-of-two groups: xor butterfly.  T = 2 * odd: fold the two halves, then a doubling
// ring over the odd remainder.  The source lanes are computed once per kernel.  Lanes of
// one group may round differently in the ring; callers take decisions from lane 0 only.
template <int T> struct GroupRing {
	int fold, n1, n2, n4; // source lanes
	__device__ __forceinline__ GroupRing(int base, int g)
	{
		constexpr int Hh = T / 2;
		const int gh = g >= Hh ? g - Hh : g;      // position inside the half
		const int hb = base + (g >= Hh ? Hh : 0); // first lane of my half
		fold = base + (g >= Hh ? g - Hh : g + Hh);
		n1 = hb + (gh + 1) % Hh;
		n2 = hb + (gh + 2) % Hh;
		n4 = hb + (gh + 4) % Hh;
	}
	__device__ __forceinline__ float sum(float v, unsigned mask) const
	{
		if constexpr ((T & (T - 1)) == 0) {
#pragma unroll
			for (int o = T / 2; o > 0; o >>= 1)
				v += __shfl_xor_sync(mask, v, o);
			return v;
		} else {
			constexpr int Hh = T / 2; // 5 (T = 10) or 3 (T = 6)
			static_assert(T == 2 * Hh && (Hh == 3 || Hh == 5), "unsupported group size");
			// ring over Hh values w_0..w_{Hh-1} (after the fold both halves hold the same ones)
			const float w = v + __shfl_sync(mask, v, fold);
			float a = w + __shfl_sync(mask, w, n1); // w_j + w_{j+1}
			if constexpr (Hh == 3) {
				a += __shfl_sync(mask, w, n2);
			} else {
				a += __shfl_sync(mask, a, n2); // four consecutive
				a += __shfl_sync(mask, w, n4); // the fifth
			}
			return a;
		}
	}
};

// MODE (qp_common.cuh, PassArgs):
//   PASS_PLAIN  every problem of the batch to the end: the round-1 kernel
//   PASS_FIRST  the same with p.ps.window iterations per problem; what has not converged by
//               then is evicted: z, u, iteration count AND the inverse (4 n^2 bytes, so that
//               nobody eliminates twice) to scratch, the index to next_list
//   PASS_LIST   the problems of p.ps.list, G at a time: the saved inverse is reloaded (same
//               layout as P, no elimination), z / u / iteration count from scratch,
//               p.ps.window more iterations; evicts again unless next_list is null
// One instantiation per mode: as a run-time flag the two kinds of round cost ~100 spilled
// registers in the hot loops.
template <int NP, bool PFI, int MODE>
__global__ void __launch_bounds__(BPQ_WARPS * 32, BpqCfg<NP>::MINB)
qp_bpq_kernel(const KernelArgs p)
{
	using C = BpqCfg<NP>;
	constexpr int T = C::T, G = C::G;
	constexpr unsigned FULL = 0xffffffffu;
	__shared__ __align__(16) float sm[BPQ_WARPS * C::GG * C::GROUP];
	__shared__ float4 ctab[5]; // columns of -I_4, then a zero vector

	__shared__ unsigned long long cta_stat[3]; // PASS_PLAIN probe: iterations, waited-for, problems
	if (MODE == PASS_FIRST && p.ps.stats && blockIdx.x == 0 && threadIdx.x == 32)
		atomicAdd(p.ps.stats + 1, p.count);
	if (threadIdx.x < 3)
		cta_stat[threadIdx.x] = 0ull;
	if (threadIdx.x < 5) {
		const int t = threadIdx.x;
		ctab[t] = make_float4(t == 0 ? -1.0f : 0.0f, t == 1 ? -1.0f : 0.0f, t == 2 ? -1.0f : 0.0f,
				      t == 3 ? -1.0f : 0.0f);
	}
	__syncthreads();

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int grp = lane / T;     // group inside the warp (== G for the idle lanes)
	const int g = lane - grp * T; // lane inside its group
	const int gbase = grp * T;
	// the idle lanes (grp == G) fold into the last T lanes so that their shuffles stay in range
	const GroupRing<T> ring(grp < G ? gbase : 32 - T, grp < G ? g : lane - (32 - T));
	float *bc = sm + (warp * C::GG + grp) * C::GROUP; // bc[buf * BUF + 4 * j + m]
	const float alpha = p.alpha;
	const bool relax = (alpha != 1.0f);

	constexpr bool LIST = MODE == PASS_LIST;
	unsigned st_it = 0, st_wait = 0, st_n = 0; // PASS_PLAIN probe, per lane
	// problems this launch works on: the batch, or the list the previous pass left
	const unsigned long long total = LIST ? (unsigned long long)__ldcg(p.ps.list_count) : p.count;
	const unsigned long long stride = (unsigned long long)gridDim.x * BPQ_WARPS * G;
	// (work is dealt warp-major: a small batch, or the short list of a late pass, spreads over
	// the SMs one warp per SM instead of filling the first few CTAs)
	for (unsigned long long base = ((unsigned long long)warp * gridDim.x + blockIdx.x) * G;
	     base < total; base += stride) {
		const bool valid = (grp < G) && (base + grp < total);
		unsigned long long prob = base + grp;
		if (LIST)
			prob = valid ? __ldcg(p.ps.list + prob) : 0ull;
		// penalty of this group's problem (qp_batch.rho)
		const float rho = (p.rho_v && valid) ? __ldg(p.rho_v + prob) : p.rho;
		const float rho2 = rho * rho, relrho = p.reltol * rho;

		// ---- load: rows 4g..4g+3 of P + rho I, read as column blocks of the symmetric P ------
		float2 S[2][NP]; // S[rp][j] = (row 4g+2rp, row 4g+2rp+1) at column j
		{
			// (a listed problem brings its inverse: same layout, written by an earlier launch)
			const float4 *Pp = reinterpret_cast<const float4 *>(
				(LIST ? p.ps.Rs + (valid ? prob : 0ull) * (unsigned long long)(NP * NP)
				      : p.P + (valid ? p_index(p, prob) : 0ull) * (unsigned long long)(NP * NP))) + g;
#pragma unroll
			for (int j = 0; j < NP; ++j) {
				float4 v = valid ? __ldg(Pp + j * (NP / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
				if (j / 4 == g && !LIST) { // the diagonal element of row 4g + (j % 4)
					const float d = valid ? rho : 1.0f; // idle lanes / tail: identity
					if (j % 4 == 0) v.x += d;
					if (j % 4 == 1) v.y += d;
					if (j % 4 == 2) v.z += d;
					if (j % 4 == 3) v.w += d;
				}
				S[0][j] = make_float2(v.x, v.y), S[1][j] = make_float2(v.z, v.w);
			}
		}

		// ---- R = (P + rho I)^-1: block Gauss-Jordan, four pivots per step, shifting form ----
		// PFI = product form: the inverse is kept as the sequence of block eta transformations
		// instead of being multiplied out.  Only the first L = NC - K chunks are still matrix
		// columns in step K; chunk L-1 receives the step's eta columns H and is not touched
		// again.  That halves the elimination work; x = R y becomes NC rank-4 updates of a
		// vector (see the iteration loop).  The run-time block loop is cut into compile-time
		// segments of four steps so that L is known per segment up to three.
		bool bad = !(rho > 0.0f && rho < 3.0e38f);
		constexpr int NC = NP / 4;
		constexpr int SEG = PFI ? 4 : NC;
		constexpr int NSEG = (NC + SEG - 1) / SEG;
		// chunk c of my four rows -> the pivot-row buffer (position-major: dst[4 * position + m])
		auto publish_chunk = [&](float *dst, int c) {
#pragma unroll
			for (int t = 0; t < 4; ++t)
				*reinterpret_cast<float4 *>(dst + 4 * (4 * c + t)) =
					make_float4(S[0][4 * c + t].x, S[0][4 * c + t].y, S[1][4 * c + t].x, S[1][4 * c + t].y);
		};
		if constexpr (!LIST) {
		if (g == 0) {
#pragma unroll
			for (int c = 0; c < NC; ++c)
				publish_chunk(bc, c);
		}
#pragma unroll
		for (int seg = 0; seg < NSEG; ++seg) {
		const int Kseg = (seg + 1) * SEG < NC ? (seg + 1) * SEG : NC;
		const int Lmax = PFI ? NC - seg * SEG : NC, Lmin = PFI ? NC - (Kseg - 1) : NC;
		for (int K = seg * SEG; K < Kseg; ++K) {
			const int L = PFI ? NC - K : NC;
			const float *pb = bc + (K & 1) * C::BUF; // pb[4 * position + m]: pivot row m
			float *pn = bc + ((K + 1) & 1) * C::BUF; // the rows of step K + 1 are collected here
			const bool nowner = (K + 1 < NC) && (g == K + 1);
			// The owner of step K + 1 publishes its four rows as they are, chunk by chunk as
			// step K finishes them, so that the stores overlap step K's arithmetic.  (The
			// intermediate matrix of Gauss-Jordan is symmetric up to sign, so every lane could
			// publish the entries of the pivot rows at its own columns from its first four
			// columns -- four stores per lane instead of forty by one lane.  That was 11 %
			// faster at n = 40 and is NOT done: rows rebuilt from other rows' columns break the
			// row-operation invariant R A = I, and on ill-conditioned problems the left
			// residual grows a hundredfold; profiles/README.md "block pivot accuracy".)  In
			// product form only the first L chunks are still matrix columns.
			__syncwarp();
			if (g == K) {
				// my four rows become -I_4, so that the common update below produces
				// B^-1 x (pivot rows) for them.  Register moves: the kernel is bound by the
				// shared-memory pipe, and filling from a constant table in shared memory
				// (-DQPB_BPQ_LDS_FILL, faster in qp_gpq.cuh) costs 2-4 % here.
#pragma unroll
				for (int j = 0; j < NP; ++j) {
					if (j / 4 < Lmin || j / 4 < L) {
#ifdef QPB_BPQ_LDS_FILL
						const float4 c = lds128_volatile(&ctab[j < 4 ? j : 4]);
						S[0][j] = make_float2(c.x, c.y), S[1][j] = make_float2(c.z, c.w);
#else
						S[0][j] = make_float2(j == 0 ? -1.0f : 0.0f, j == 1 ? -1.0f : 0.0f);
						S[1][j] = make_float2(j == 2 ? -1.0f : 0.0f, j == 3 ? -1.0f : 0.0f);
#endif
					}
				}
			}
			// pivot block (B[m][j] = pivot row m at column j) and its L D L^T
			float B[4][4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const float4 v = *reinterpret_cast<const float4 *>(pb + 4 * j);
				B[0][j] = v.x, B[1][j] = v.y, B[2][j] = v.z, B[3][j] = v.w;
			}
			Ldl4 Bf;
			bad |= ldl4_factor(B, Bf);
			// H = -(first four columns) x B^-1 by two triangular solves; H[m][rp] is the pair
			// of multipliers of pivot row m for my row pair rp
			float2 H[4][2];
#pragma unroll
			for (int rp = 0; rp < 2; ++rp)
				ldl4_minus_solve2(Bf, S[rp][0], S[rp][1], S[rp][2], S[rp][3], H[0][rp], H[1][rp],
						  H[2][rp], H[3][rp]);
			// S(:, j-4) <- S(:, j) + H * pivot_rows(:, j)
			auto update_chunk = [&](int c) {
				float4 pr[4]; // pr[t] = the four pivot rows at position 4c + t
#pragma unroll
				for (int t = 0; t < 4; ++t)
					pr[t] = *reinterpret_cast<const float4 *>(pb + 4 * (4 * c + t));
#pragma unroll
				for (int rp = 0; rp < 2; ++rp) {
					float2 a0 = S[rp][4 * c], a1 = S[rp][4 * c + 1], a2 = S[rp][4 * c + 2],
					       a3 = S[rp][4 * c + 3];
					a0 = fma2_bc(H[0][rp], pr[0].x, a0), a1 = fma2_bc(H[0][rp], pr[1].x, a1);
					a2 = fma2_bc(H[0][rp], pr[2].x, a2), a3 = fma2_bc(H[0][rp], pr[3].x, a3);
					a0 = fma2_bc(H[1][rp], pr[0].y, a0), a1 = fma2_bc(H[1][rp], pr[1].y, a1);
					a2 = fma2_bc(H[1][rp], pr[2].y, a2), a3 = fma2_bc(H[1][rp], pr[3].y, a3);
					a0 = fma2_bc(H[2][rp], pr[0].z, a0), a1 = fma2_bc(H[2][rp], pr[1].z, a1);
					a2 = fma2_bc(H[2][rp], pr[2].z, a2), a3 = fma2_bc(H[2][rp], pr[3].z, a3);
					a0 = fma2_bc(H[3][rp], pr[0].w, a0), a1 = fma2_bc(H[3][rp], pr[1].w, a1);
					a2 = fma2_bc(H[3][rp], pr[2].w, a2), a3 = fma2_bc(H[3][rp], pr[3].w, a3);
					S[rp][4 * c - 4] = a0, S[rp][4 * c - 3] = a1, S[rp][4 * c - 2] = a2,
						      S[rp][4 * c - 1] = a3;
				}
				if (nowner)
					publish_chunk(pn, c - 1);
			};
			auto place_eta = [&](int c) { // chunk c <- H
#pragma unroll
				for (int m = 0; m < 4; ++m)
					S[0][4 * c + m] = H[m][0], S[1][4 * c + m] = H[m][1];
			};
#pragma unroll
			for (int c = 1; c < NC; ++c) {
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
			if (Lmax == NC && L == NC) {
				place_eta(NC - 1);
				if (!PFI && nowner) // full inverse: the eta columns are matrix columns again
					publish_chunk(pn, NC - 1);
			}
		}
		}
		}
		__syncwarp();
		// PASS_FIRST: the inverse (or eta file) goes to scratch now, for every problem, in uniform
		// code -- not after the window from the lanes of the evicted ones.  (Keeping S alive
		// behind the iteration loop cost the elimination loop its load scheduling: twice the
		// short-scoreboard stalls, 1.82 against 1.57 ms on 2^18 random problems that evict
		// nobody; and with the streaming pass behind a short window most problems are evicted.)
		if (MODE == PASS_FIRST || MODE == PASS_SETUP) {
			if (valid && !bad && (MODE == PASS_SETUP || p.ps.next_list)) {
				float4 *Rp = reinterpret_cast<float4 *>(p.ps.Rs + prob * (unsigned long long)(NP * NP)) + g;
#pragma unroll
				for (int j = 0; j < NP; ++j)
					Rp[j * (NP / 4)] = make_float4(S[0][j].x, S[0][j].y, S[1][j].x, S[1][j].y);
			}
		}
		if constexpr (MODE == PASS_SETUP) {
			// the launch's problems are matrices: inverse stored, verdict stored, next
			if (valid && g == 0)
				p.ps.its[prob] = bad ? 1 : 0;
			continue;
		}

		// ---- per-coordinate state: coordinates 4g..4g+3 ---------------------------------------
		float q[4], lo[4], hi[4], z[4], u[4], x[4];
		{
			const unsigned long long o4 = (valid ? prob : 0ull) * (unsigned long long)(NP / 4) + g;
			const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
			const float4 vq = valid ? __ldg(reinterpret_cast<const float4 *>(p.q) + o4) : zero4;
			const float4 vl = (valid && p.lb) ? __ldg(reinterpret_cast<const float4 *>(p.lb) + o4)
							  : make_float4(p.lb_s, p.lb_s, p.lb_s, p.lb_s);
			const float4 vh = (valid && p.ub) ? __ldg(reinterpret_cast<const float4 *>(p.ub) + o4)
							  : make_float4(p.ub_s, p.ub_s, p.ub_s, p.ub_s);
			float4 vz = (valid && p.z0) ? __ldg(reinterpret_cast<const float4 *>(p.z0) + o4) : zero4;
			float4 vu = (valid && p.u0) ? __ldg(reinterpret_cast<const float4 *>(p.u0) + o4) : zero4;
			if (LIST && valid) { // carries on from its saved state
				vz = __ldcg(reinterpret_cast<const float4 *>(p.ps.zs) + o4);
				vu = __ldcg(reinterpret_cast<const float4 *>(p.ps.us) + o4);
			}
			q[0] = vq.x, q[1] = vq.y, q[2] = vq.z, q[3] = vq.w;
			lo[0] = vl.x, lo[1] = vl.y, lo[2] = vl.z, lo[3] = vl.w;
			hi[0] = vh.x, hi[1] = vh.y, hi[2] = vh.z, hi[3] = vh.w;
			z[0] = vz.x, z[1] = vz.y, z[2] = vz.z, z[3] = vz.w;
			u[0] = vu.x, u[1] = vu.y, u[2] = vu.z, u[3] = vu.w;
			x[0] = x[1] = x[2] = x[3] = 0.0f;
		}

		// ---- ADMM iterations (qp_solvers.c:284-310) -----------------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		const bool active = valid && !bad && p.max_iter > 0; // uniform inside a group
		// iterations of this launch: to the end (PASS_PLAIN), the window (PASS_FIRST; the host
		// passes min(window, max_iter)), or the window and what the problem has left (PASS_LIST)
		unsigned done0 = 0, it_stop = 0;
		if (LIST) {
			if (valid)
				done0 = unsigned(__ldcg(p.ps.its + prob)); // < max_iter
			const unsigned left_it = p.max_iter - done0;
			it_stop = left_it < p.ps.window ? left_it : p.ps.window;
		}
#define QPB_BPQ_IT_END (MODE == PASS_PLAIN ? p.max_iter : (LIST ? it_stop : p.ps.window))
		// x = R y for my four rows, y broadcast from shared memory (full-inverse variant)
		auto publish_y = [&](float *yb) {
			*reinterpret_cast<float4 *>(yb + 4 * g) =
				make_float4(fmaf(rho, z[0] - u[0], -q[0]), fmaf(rho, z[1] - u[1], -q[1]),
					    fmaf(rho, z[2] - u[2], -q[2]), fmaf(rho, z[3] - u[3], -q[3]));
		};
		auto matvec = [&](const float *yb, float (&xo)[4]) {
			// four partial sums per row pair keep the FFMA2 chains short
			float2 acc[2][4];
#pragma unroll
			for (int rp = 0; rp < 2; ++rp)
				acc[rp][0] = acc[rp][1] = acc[rp][2] = acc[rp][3] = make_float2(0.0f, 0.0f);
#pragma unroll
			for (int j = 0; j < NP; j += 4) {
				const float4 y = *reinterpret_cast<const float4 *>(yb + j);
#pragma unroll
				for (int rp = 0; rp < 2; ++rp) {
					acc[rp][0] = fma2_bc(S[rp][j], y.x, acc[rp][0]);
					acc[rp][1] = fma2_bc(S[rp][j + 1], y.y, acc[rp][1]);
					acc[rp][2] = fma2_bc(S[rp][j + 2], y.z, acc[rp][2]);
					acc[rp][3] = fma2_bc(S[rp][j + 3], y.w, acc[rp][3]);
				}
			}
#pragma unroll
			for (int rp = 0; rp < 2; ++rp) {
				const float2 t = __fadd2_rn(__fadd2_rn(acc[rp][0], acc[rp][1]),
							    __fadd2_rn(acc[rp][2], acc[rp][3]));
				xo[2 * rp] = t.x, xo[2 * rp + 1] = t.y;
			}
		};
		// element-wise part of one iteration with x = xr: admm_update_x_hat (:161-174),
		// _z (:176-190), _u (:192-203) and the five partial sums of squares (:205-253)
		auto elementwise = [&](const float (&xr)[4], float &rr, float &ss, float &nx, float &nz,
				       float &nu) {
			rr = ss = nx = nz = nu = 0.f;
#pragma unroll
			for (int s = 0; s < 4; ++s) {
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
		// stop test (:233-253, :308), squared form as in qp_tpq.cuh; lane 0 of the group
		// decides for everybody
		// PASS_FIRST also notes whether the problem looks like a straggler: at the rate ||x - z||^2
		// shrank in the last iteration of the window (d = previous / current), it needs more than
		// 32 further iterations to get under its threshold -- rr > eps_pri^2 d^32.  Problems evicted
		// in that state go to the list the streaming pass hands out first.  (Offline check on the
		// condensed-MPC set, float64 restatement: with a window of 6 the rule flags 4.8 % of the
		// problems, among them the eight slowest (321 ... 2 783 iterations), and the longest one
		// it misses needs 315 more; windows 8 ... 24: 9 ... 24 % flagged, longest miss 213 ... 169.
		// The dual residual and plain decay rates rank them far worse.)
		float rr_prev = 0.f;
		bool slow = false;
		auto converged = [&](float rr, float ss, float nx, float nz, float nu, unsigned am) {
			const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
			const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
			const bool cp = rr < eps_pri * eps_pri, cd = rho2 * ss < eps_dual * eps_dual;
			if (MODE == PASS_FIRST) {
				float d = rr_prev * rcp_approx(rr);
				d *= d, d *= d, d *= d, d *= d, d *= d;
				slow = rr > eps_pri * eps_pri * d;
				rr_prev = rr;
			}
			int conv = cp && cd;
			if constexpr ((T & (T - 1)) != 0)
				conv = __shfl_sync(am, conv, gbase);
			return conv != 0;
		};

		// Every warp-wide operation of the loops below names only the lanes of the group
		// (am), so a group simply leaves its loop when its problem is done and waits for
		// the other groups at the __syncwarp() behind it: no ballot, no flag to test.
		const unsigned am = (((1u << T) - 1u) << gbase);
		bool conv = false;
		float xnext[4] = { 0.f, 0.f, 0.f, 0.f }; // full inverse: the x-update after the last iteration done
		if constexpr (PFI) {
		if (active) {
			do {
				// admm_update_x (:146-159): y = rho (z - u) - q ; x = R y with
				// R = E_{NC-1} ... E_0: E_K maps the four pivot entries v_P (all in lane K
				// of the group) to B^-1 v_P and adds H_K v_P to every other entry; H_K
				// sits in chunk NC-1-K, and the pivot rows' H is B^-1 itself.
				float xr[4];
				float2 v0 = make_float2(fmaf(rho, z[0] - u[0], -q[0]), fmaf(rho, z[1] - u[1], -q[1]));
				float2 v1 = make_float2(fmaf(rho, z[2] - u[2], -q[2]), fmaf(rho, z[3] - u[3], -q[3]));
#pragma unroll
				for (int K = 0; K < NC; ++K) {
					const int src = gbase + K;
					const float p0 = __shfl_sync(am, v0.x, src), p1 = __shfl_sync(am, v0.y, src);
					const float p2 = __shfl_sync(am, v1.x, src), p3 = __shfl_sync(am, v1.y, src);
					const int c = 4 * (NC - 1 - K);
					float2 a0 = (g == K) ? make_float2(0.f, 0.f) : v0;
					float2 a1 = (g == K) ? make_float2(0.f, 0.f) : v1;
					a0 = fma2_bc(S[0][c], p0, a0), a1 = fma2_bc(S[1][c], p0, a1);
					a0 = fma2_bc(S[0][c + 1], p1, a0), a1 = fma2_bc(S[1][c + 1], p1, a1);
					a0 = fma2_bc(S[0][c + 2], p2, a0), a1 = fma2_bc(S[1][c + 2], p2, a1);
					a0 = fma2_bc(S[0][c + 3], p3, a0), a1 = fma2_bc(S[1][c + 3], p3, a1);
					v0 = a0, v1 = a1;
				}
				xr[0] = v0.x, xr[1] = v0.y, xr[2] = v1.x, xr[3] = v1.y;
				float rr, ss, nx, nz, nu;
				elementwise(xr, rr, ss, nx, nz, nu);
				rr = ring.sum(rr, am), ss = ring.sum(ss, am), nx = ring.sum(nx, am);
				nz = ring.sum(nz, am), nu = ring.sum(nu, am);
				++it;
				conv = converged(rr, ss, nx, nz, nu, am);
			} while (!conv && it < QPB_BPQ_IT_END);
		}
		} else {
		// Full inverse.  The loop is software-pipelined: the mat-vec of the NEXT iteration is
		// started before the five group reductions and the stop decision of this one are
		// known (it only reads z, u, which the decision does not change), so the shuffle
		// chains overlap the FMAs instead of heading a serial chain of ~300 cycles per
		// iteration.  One mat-vec per problem is wasted (the one after convergence).
		float xr[4] = { 0.f, 0.f, 0.f, 0.f };
		if (active) {
			publish_y(bc);
			__syncwarp(am);
			matvec(bc, xr);
			do {
				float rr, ss, nx, nz, nu;
				elementwise(xr, rr, ss, nx, nz, nu);
				++it;
				float *yb = bc + (it & 1u) * C::BUF; // two buffers: no second barrier needed
				publish_y(yb);
				__syncwarp(am);
				rr = ring.sum(rr, am), ss = ring.sum(ss, am), nx = ring.sum(nx, am);
				nz = ring.sum(nz, am), nu = ring.sum(nu, am);
				matvec(yb, xr);
				conv = converged(rr, ss, nx, nz, nu, am);
			} while (!conv && it < QPB_BPQ_IT_END);
		}
		xnext[0] = xr[0], xnext[1] = xr[1], xnext[2] = xr[2], xnext[3] = xr[3];
		}
#undef QPB_BPQ_IT_END
		__syncwarp();
		if (conv)
			status = QP_STATUS_CONVERGED;

		// ---- the window closed on an unfinished problem: state and inverse to scratch, index to
		// the list of the next launch ------------------------------------------------------------
		bool evict = false;
		if (MODE != PASS_PLAIN) {
			if (p.ps.next_list)
				evict = active && !conv && done0 + it < p.max_iter;
			const unsigned em = __ballot_sync(FULL, evict && g == 0);
			if (em) {
				if (evict) {
					const unsigned long long o4 = prob * (unsigned long long)(NP / 4) + g;
					reinterpret_cast<float4 *>(p.ps.zs)[o4] = make_float4(z[0], z[1], z[2], z[3]);
					reinterpret_cast<float4 *>(p.ps.us)[o4] = make_float4(u[0], u[1], u[2], u[3]);
					// (the inverse is in scratch already: written behind the elimination)
					// the streaming pass starts from the x-update the pipelined loop has already
					// computed for the next iteration: no mat-vec in its refill block
					if (MODE == PASS_FIRST && !PFI) {
						if (p.ps.xs)
							reinterpret_cast<float4 *>(p.ps.xs)[o4] =
								make_float4(xnext[0], xnext[1], xnext[2], xnext[3]);
					}
					if (g == 0)
						p.ps.its[prob] = int(done0 + it);
				}
				// (two lists when the streaming pass follows: expected stragglers / the others;
				// the group's first lane decides, as for the stop test)
				unsigned em2 = 0;
				if (MODE == PASS_FIRST) {
					if (p.ps.next_list2)
						em2 = __ballot_sync(FULL, evict && g == 0 && !slow);
				}
				const unsigned em1 = em & ~em2;
				unsigned slot = 0, slot2 = 0;
				if (lane == 0) {
					if (em1)
						slot = atomicAdd(p.ps.next_count, unsigned(__popc(em1)));
					if (em2)
						slot2 = atomicAdd(p.ps.next_count2, unsigned(__popc(em2)));
					if (p.ps.stats)
						atomicAdd(p.ps.stats + (LIST ? 2 : 0), (unsigned long long)__popc(em));
				}
				slot = __shfl_sync(FULL, slot, 0);
				slot2 = __shfl_sync(FULL, slot2, 0);
				if (evict && g == 0) {
					const unsigned below = (1u << lane) - 1u;
					if ((em2 >> lane) & 1u)
						p.ps.next_list2[slot2 + __popc(em2 & below)] = unsigned(prob);
					else
						p.ps.next_list[slot + __popc(em1 & below)] = unsigned(prob);
				}
			}
		}

		// ---- results ---------------------------------------------------------------------------
		float nf = fmaf(x[0], 0.0f, fmaf(x[1], 0.0f, fmaf(x[2], 0.0f, x[3] * 0.0f)));
		nf = ring.sum(nf, FULL);
		int iters = int(done0 + it);
		if (bad) {
			status = QP_STATUS_SINGULAR;
			iters = 0;
		} else if (!(nf == 0.0f)) {
			status = QP_STATUS_NONFINITE;
		}
		if (MODE == PASS_PLAIN) {
			// the probe of QP_SCHEDULE_AUTO: what my group's problem needed and what its warp
			// waited for (per-lane sums in the groups' first lanes, added up at the end)
			if (p.ps.stats) {
				const bool lead = valid && g == 0;
				const unsigned mine_it = lead ? unsigned(iters) : 0u;
				const unsigned mx = __reduce_max_sync(FULL, mine_it);
				st_it += mine_it;
				st_wait += lead ? mx : 0u;
				st_n += lead ? 1u : 0u;
			}
		}
		if (valid && !evict) {
			const unsigned long long o4 = prob * (unsigned long long)(NP / 4) + g;
			if (p.x)
				reinterpret_cast<float4 *>(p.x)[o4] = make_float4(x[0], x[1], x[2], x[3]);
			if (p.z)
				reinterpret_cast<float4 *>(p.z)[o4] = make_float4(z[0], z[1], z[2], z[3]);
			if (p.lam)
				reinterpret_cast<float4 *>(p.lam)[o4] =
					make_float4(rho * u[0], rho * u[1], rho * u[2], rho * u[3]);
			if (g == 0) {
				if (p.status)
					p.status[prob] = status;
				if (p.iters)
					p.iters[prob] = iters;
			}
		}
	}
	if (MODE == PASS_PLAIN) {
		if (p.ps.stats) { // one add per warp to the CTA's sums, one per CTA to the device's
			__syncwarp();
			const unsigned a = __reduce_add_sync(FULL, st_it), b = __reduce_add_sync(FULL, st_wait),
				       c = __reduce_add_sync(FULL, st_n);
			if (lane == 0) {
				atomicAdd(&cta_stat[0], (unsigned long long)a);
				atomicAdd(&cta_stat[1], (unsigned long long)b);
				atomicAdd(&cta_stat[2], (unsigned long long)c);
			}
			__syncthreads();
			if (threadIdx.x < 3 && cta_stat[threadIdx.x])
				atomicAdd(p.ps.stats + 3 + threadIdx.x, cta_stat[threadIdx.x]);
		}
	}
}

} // namespace qpb
