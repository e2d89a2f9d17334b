// qp_tpq.cuh -- "thread per QP" kernel for small problems (n = 4, 8, 12, 16).
//
// What it replaces: one call of the reference's admm()
// (/root/reference/qp_solvers/qp_solvers.c:255-319) per problem, including the
// one-off factorisation the reference does with matrix_invert()
// (/root/reference/matrix/matrix_ops/matrix_ops.c:551-630) and the per-iteration
// helpers admm_update_x/_x_hat/_z/_u, admm_r_norm/_s_norm, admm_eps_pri/_dual
// (qp_solvers.c:146-253).
//
// Mapping (B200-first, not a translation):
//   * one lane owns one QP; a warp owns 32 consecutive QPs ("warp batch").
//     There are no shuffles or reductions on the iteration path: every dot
//     product, norm and triangular solve is a straight FMA chain in registers.
//   * P + rho*I (symmetric positive definite) is factored ONCE as L D L^T into
//     n(n+1)/2 registers (136 for n = 16).  The reference forms the explicit
//     inverse R = (P+rho I)^-1 with LU and multiplies by it every iteration;
//     here the x-update solves (P+rho I) x = rho(z-u) - q with two triangular
//     substitutions: the same 2n^2 flops per iteration as R*y, but the setup
//     costs n^3/3 flops instead of 8/3 n^3 and is better conditioned.
//   * problem data arrive by bulk asynchronous copies (cp.async.bulk, the TMA
//     engine) into a CTA-wide ring of shared-memory stages, one stage per warp
//     batch, completion signalled on an mbarrier.  Each P lands in a slot padded
//     by 16 bytes so that the per-lane 128-bit reads of the upper triangle are
//     bank-conflict free.  A warp that has emptied a stage immediately re-arms
//     it for a later batch, so loads run a whole batch ahead of compute.
//   * x, z, lambda leave through shared memory with one bulk store per array
//     per warp batch (fully coalesced 2 KB writes at n = 16); status and
//     iteration count are written as coalesced 4-byte stores.
//   * persistent grid: (#SMs x CTAs/SM) CTAs; warp batches are handed out
//     through a per-CTA ticket counter, CTA c serving batches c, c+grid, ...
#pragma once

#include "qp_common.cuh"


namespace qpb {

// warps per CTA, ring stages and CTAs per SM, chosen so that registers and the
// 227 KB of shared memory are both filled
template <int N, bool BOUNDS, bool WARM> struct TpqCfg;
template <bool B, bool W> struct TpqCfg<16, B, W> {
#ifdef QPB_TPQ16_W12
	static constexpr int WARPS = (B ? 8 : 12), STAGES = 3, MINB = 1;
#else
	static constexpr int WARPS = 8, STAGES = (B ? 3 : 4), MINB = 1;
#endif
};
template <bool B, bool W> struct TpqCfg<12, B, W> { static constexpr int WARPS = 12, STAGES = 3, MINB = 1; };
template <bool B, bool W> struct TpqCfg<8, B, W>  { static constexpr int WARPS = 8, STAGES = 4, MINB = 2; };
template <bool B, bool W> struct TpqCfg<4, B, W>  { static constexpr int WARPS = 8, STAGES = 4, MINB = 4; };

// packed upper-triangle index, i <= j
template <int N> __host__ __device__ constexpr int sidx(int i, int j)
{
	return i * N - (i * (i - 1)) / 2 + (j - i);
}

template <int N, bool BOUNDS, bool WARM> struct TpqLayout {
	using Cfg = TpqCfg<N, BOUNDS, WARM>;
	static constexpr int PAD = (N >= 8) ? 4 : 0;          // floats of padding per P slot
	static constexpr int PSLOT = N * N + PAD;             // floats per P slot
	// packed upper triangles (host path, qp_batch.cu): record and slot sizes; the slot is
	// an odd number of 16-byte units so that the lanes' 128-bit reads do not conflict
	static constexpr int PK = tpq_packed_stride(N);
	static constexpr int PKSLOT = ((PK / 4) % 2) ? PK : PK + 4;
	static_assert(PKSLOT <= PSLOT, "packed slots must fit the stage");
	static constexpr int NVEC_STAGE = 1 + (BOUNDS ? 2 : 0) + (WARM ? 2 : 0);
	static constexpr int STAGE_P = 32 * PSLOT;
	static constexpr int STAGE = STAGE_P + NVEC_STAGE * 32 * N; // floats per stage
	static constexpr int PRIV_VEC = (1 + (BOUNDS ? 2 : 0)) * 32 * N;
	static constexpr int PRIV_OUT = 3 * 32 * N;
	static constexpr int PRIV = PRIV_VEC + PRIV_OUT;      // floats per warp
	static constexpr int FLOATS = Cfg::STAGES * STAGE + Cfg::WARPS * PRIV;
	static constexpr size_t BYTES = size_t(FLOATS) * 4 + Cfg::STAGES * 8 /*mbarriers*/
					+ (1 + Cfg::STAGES) * 4 /*ticket, rounds*/ + 16;
	static_assert(BYTES * Cfg::MINB <= 232448, "shared memory budget of one SM exceeded");
};

// ---- register-resident dense routines (all indices compile-time) ------------

// In place L D L^T of the packed symmetric matrix a (upper triangle of A).
// After return a(k,k) = 1/d_k and a(k,j), j > k, holds L(j,k) (unit lower L).
// Returns true when a pivot is not strictly positive (A not positive definite).
template <int N> __device__ __forceinline__ bool ldlt_factor(float (&a)[N * (N + 1) / 2])
{
	bool bad = false;
#pragma unroll
	for (int k = 0; k < N; ++k) {
		const float piv = a[sidx<N>(k, k)];
		bad |= !(piv > 0.0f);
		const float dinv = rcp_newton(piv);
		a[sidx<N>(k, k)] = dinv;
		float l[N];
#pragma unroll
		for (int j = k + 1; j < N; ++j)
			l[j] = a[sidx<N>(k, j)] * dinv;
#pragma unroll
		for (int i = k + 1; i < N; ++i) {
#pragma unroll
			for (int j = i; j < N; ++j)
				a[sidx<N>(i, j)] = fmaf(-l[i], a[sidx<N>(k, j)], a[sidx<N>(i, j)]);
		}
#pragma unroll
		for (int j = k + 1; j < N; ++j)
			a[sidx<N>(k, j)] = l[j];
	}
	return bad;
}

// w <- (L D L^T)^-1 w, column-oriented substitutions (independent FMAs per step)
template <int N>
__device__ __forceinline__ void ldlt_solve(const float (&a)[N * (N + 1) / 2], float (&w)[N])
{
#pragma unroll
	for (int k = 0; k < N; ++k) {
#pragma unroll
		for (int i = k + 1; i < N; ++i)
			w[i] = fmaf(-a[sidx<N>(k, i)], w[k], w[i]);
	}
#pragma unroll
	for (int k = 0; k < N; ++k)
		w[k] *= a[sidx<N>(k, k)];
#pragma unroll
	for (int i = N - 1; i > 0; --i) {
#pragma unroll
		for (int k = 0; k < i; ++k)
			w[k] = fmaf(-a[sidx<N>(k, i)], w[i], w[k]);
	}
}

// ---- the kernel -----------------------------------------------------------------
//
// MODE (qp_common.cuh, PassArgs):
//   PASS_PLAIN  every problem of the batch to the end: the round-1 kernel
//   PASS_FIRST  the same with p.ps.window iterations per problem; what has not converged by
//               then is evicted (z, u, iteration count to scratch, index to next_list)
//   PASS_LIST   the problems of p.ps.list (scattered: data straight from global memory, z / u /
//               iteration count from scratch, results by per-lane stores), p.ps.window more
//               iterations each; evicts again unless next_list is null
// One instantiation per mode: a mode is a compile-time constant in the loops.

template <int N, bool BOUNDS, bool WARM, bool RELAX, bool PACKED_IN, int MODE>
__global__ void
__launch_bounds__(TpqCfg<N, BOUNDS, WARM>::WARPS * 32, TpqCfg<N, BOUNDS, WARM>::MINB)
qp_tpq_kernel(const KernelArgs p)
{
	using L = TpqLayout<N, BOUNDS, WARM>;
	using Cfg = TpqCfg<N, BOUNDS, WARM>;
	constexpr int STAGES = Cfg::STAGES;
	constexpr int NQ = N / 4; // float4 per vector
	constexpr unsigned FULL = 0xffffffffu;
	constexpr bool LIST = MODE == PASS_LIST;

	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ unsigned long long cta_stat[3]; // PASS_PLAIN probe: iterations, waited-for, problems
	float *ring = reinterpret_cast<float *>(smem_raw);
	float *priv_all = ring + STAGES * L::STAGE;
	uint64_t *full_bar = reinterpret_cast<uint64_t *>(priv_all + Cfg::WARPS * L::PRIV);
	int *ticket = reinterpret_cast<int *>(full_bar + STAGES);
	volatile int *stage_round = ticket + 1; // [STAGES] times each stage has been armed

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	float *priv = priv_all + warp * L::PRIV;
	float4 *priv_vec4 = reinterpret_cast<float4 *>(priv);
	float *priv_out = priv + L::PRIV_VEC;

	// problems this launch works on: the batch, or the list the previous pass left
	const unsigned long long total = LIST ? (unsigned long long)__ldcg(p.ps.list_count) : p.count;
	const unsigned long long nbatch = (total + 31ull) / 32ull;

	if (!LIST) {
		if (threadIdx.x == 0) {
			if (MODE == PASS_FIRST && p.ps.stats && blockIdx.x == 0)
				atomicAdd(p.ps.stats + 1, p.count);
			cta_stat[0] = cta_stat[1] = cta_stat[2] = 0ull;
			for (int s = 0; s < STAGES; ++s) {
				mbar_init(&full_bar[s], 1);
				stage_round[s] = 0;
			}
			*ticket = 0;
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
	}

	// Arm stage (t % STAGES) with the data of ticket t.  Called by a whole warp.
	auto arm = [&](unsigned long long t) {
		const unsigned long long batch = blockIdx.x + t * gridDim.x;
		if (batch >= nbatch)
			return;
		const int s = int(t % STAGES);
		const unsigned long long first = batch * 32ull;
		const unsigned long long left = p.count - first;
		const int cnt = left < 32ull ? int(left) : 32;
		float *st = ring + s * L::STAGE;
		if (lane == 0)
			mbar_expect_tx(&full_bar[s],
				       uint32_t(cnt) * ((PACKED_IN ? L::PK : N * N) + L::NVEC_STAGE * N) * 4u);
		__syncwarp();
		// (one copy per lane where the slots are padded, and where problems share their
		// matrices -- qp_batch.p_share: the lanes of a group then copy the same 1 KB out of L2)
		if (PACKED_IN) {
			if (L::PKSLOT != L::PK || p.p_share > 1u) {
				if (lane < cnt)
					bulk_g2s(st + lane * L::PKSLOT, p.P + p_index(p, first + lane) * L::PK,
						 L::PK * 4, &full_bar[s]);
			} else if (lane == 0) {
				bulk_g2s(st, p.P + first * L::PK, uint32_t(cnt) * L::PK * 4, &full_bar[s]);
			}
		} else if (L::PAD > 0 || p.p_share > 1u) {
			if (lane < cnt)
				bulk_g2s(st + lane * L::PSLOT, p.P + p_index(p, first + lane) * (N * N),
					 N * N * 4, &full_bar[s]);
		} else if (lane == 0) {
			bulk_g2s(st, p.P + first * (N * N), uint32_t(cnt) * N * N * 4,
				 &full_bar[s]);
		}
		float *sv = st + L::STAGE_P;
		const uint32_t vbytes = uint32_t(cnt) * N * 4u;
		if (lane == 1)
			bulk_g2s(sv, p.q + first * N, vbytes, &full_bar[s]);
		if (BOUNDS) {
			if (lane == 2)
				bulk_g2s(sv + 32 * N, p.lb + first * N, vbytes, &full_bar[s]);
			if (lane == 3)
				bulk_g2s(sv + 2 * 32 * N, p.ub + first * N, vbytes, &full_bar[s]);
		}
		if (WARM) {
			constexpr int o = BOUNDS ? 3 : 1;
			if (lane == 4)
				bulk_g2s(sv + o * 32 * N, p.z0 + first * N, vbytes, &full_bar[s]);
			if (lane == 5)
				bulk_g2s(sv + (o + 1) * 32 * N, p.u0 + first * N, vbytes,
					 &full_bar[s]);
		}
		__syncwarp();
		if (lane == 0) {
			__threadfence_block();
			stage_round[s] = int(t / STAGES) + 1;
		}
	};

	if (!LIST) {
		if (warp < STAGES)
			arm(warp);
	}

	bool out_pending = false; // this warp has a bulk store in flight from priv_out
	unsigned st_it = 0, st_wait = 0, st_n = 0; // PASS_PLAIN probe, per lane
	// LIST: static stride, warp-major -- a short list (the late passes hold a few hundred
	// stragglers that each run thousands of dependent iterations) spreads over the SMs one warp
	// per SM instead of filling the first few CTAs, whose warps would share issue slots
	unsigned long long lbatch = (unsigned long long)warp * gridDim.x + blockIdx.x;

	for (;;) {
		int t = 0;
		unsigned long long batch;
		if (LIST) {
			batch = lbatch;
			lbatch += (unsigned long long)gridDim.x * Cfg::WARPS;
		} else {
			if (lane == 0)
				t = atomicAdd(ticket, 1);
			t = __shfl_sync(FULL, t, 0);
			batch = blockIdx.x + (unsigned long long)t * gridDim.x;
		}
		if (batch >= nbatch)
			break;
		const int s = t % STAGES;
		const int round = t / STAGES;
		const unsigned long long first = batch * 32ull;
		const unsigned long long left = total - first;
		const int cnt = left < 32ull ? int(left) : 32;
		const bool valid = lane < cnt;
		// index of this lane's problem.  LIST: read from the list; the idle lanes of a tail
		// repeat the last entry (they compute on valid data and store nothing)
		unsigned long long gidx = first + lane;
		if (LIST)
			gidx = __ldcg(p.ps.list + first + (valid ? lane : cnt - 1));
		// penalty of this thread's problem (qp_batch.rho), requested before the wait
		const float rho = (p.rho_v && (LIST || valid)) ? __ldg(p.rho_v + gidx) : p.rho;

		float a[N * (N + 1) / 2];
		float z[N], u[N], w[N];
		unsigned done0 = 0; // LIST: iterations the problem has behind it
		// row-major n x n matrix (or its packed upper triangle) -> a[], 128 bits at a time
		auto fill_a = [&](auto ld4) {
			if (PACKED_IN) {
				// the record is the upper triangle in the order a[] keeps it
#pragma unroll
				for (int k = 0; k < L::PK / 4; ++k) {
					const float4 v = ld4(k);
					const float e[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
					for (int tt = 0; tt < 4; ++tt)
						if (4 * k + tt < N * (N + 1) / 2)
							a[4 * k + tt] = e[tt];
				}
			} else {
#pragma unroll
				for (int i = 0; i < N; ++i) {
#pragma unroll
					for (int c = i / 4; c < NQ; ++c) {
						const float4 v = ld4(i * NQ + c);
						const float e[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
						for (int k = 0; k < 4; ++k)
							if (4 * c + k >= i)
								a[sidx<N>(i, 4 * c + k)] = e[k];
					}
				}
			}
		};
		if constexpr (!LIST) {
			// wait until the stage has been armed for this round, then for its bytes
			while (stage_round[s] < round + 1)
				__nanosleep(20);
			mbar_wait(&full_bar[s], uint32_t(round & 1));

			// ---- stage -> registers / private area ------------------------------------
			const float *st = ring + s * L::STAGE;
			const float4 *slot4 =
				reinterpret_cast<const float4 *>(st + lane * (PACKED_IN ? L::PKSLOT : L::PSLOT));
			fill_a([&](int k) { return slot4[k]; });
			const float4 *sv4 = reinterpret_cast<const float4 *>(st + L::STAGE_P);
#pragma unroll
			for (int k = 0; k < NQ; ++k)
				priv_vec4[k * 32 + lane] = sv4[lane * NQ + k];
			if (BOUNDS) {
#pragma unroll
				for (int k = 0; k < NQ; ++k) {
					priv_vec4[(NQ + k) * 32 + lane] = sv4[8 * N + lane * NQ + k];
					priv_vec4[(2 * NQ + k) * 32 + lane] =
						sv4[16 * N + lane * NQ + k];
				}
			}
			if (WARM) {
				constexpr int o = BOUNDS ? 3 : 1;
#pragma unroll
				for (int k = 0; k < NQ; ++k) {
					const float4 vz = sv4[o * 8 * N + lane * NQ + k];
					const float4 vu = sv4[(o + 1) * 8 * N + lane * NQ + k];
					z[4 * k] = vz.x, z[4 * k + 1] = vz.y, z[4 * k + 2] = vz.z,
					      z[4 * k + 3] = vz.w;
					u[4 * k] = vu.x, u[4 * k + 1] = vu.y, u[4 * k + 2] = vu.z,
					      u[4 * k + 3] = vu.w;
				}
			} else {
#pragma unroll
				for (int j = 0; j < N; ++j)
					z[j] = 0.0f, u[j] = 0.0f;
			}
			__syncwarp();
			// the stage is free: order our generic-proxy reads before the async
			// writes of the refill, then start the copy for ticket t + STAGES
			fence_async_smem();
			arm((unsigned long long)t + STAGES);
		} else {
			// ---- listed problems: data straight from global memory, state from scratch ------
			const float4 *src4 = reinterpret_cast<const float4 *>(
				p.P + p_index(p, gidx) * (unsigned long long)(PACKED_IN ? L::PK : N * N));
			fill_a([&](int k) { return __ldg(src4 + k); });
			const float4 *q4 = reinterpret_cast<const float4 *>(p.q) + gidx * NQ;
#pragma unroll
			for (int k = 0; k < NQ; ++k)
				priv_vec4[k * 32 + lane] = __ldg(q4 + k);
			if (BOUNDS) {
				const float4 *l4 = reinterpret_cast<const float4 *>(p.lb) + gidx * NQ;
				const float4 *h4 = reinterpret_cast<const float4 *>(p.ub) + gidx * NQ;
#pragma unroll
				for (int k = 0; k < NQ; ++k) {
					priv_vec4[(NQ + k) * 32 + lane] = __ldg(l4 + k);
					priv_vec4[(2 * NQ + k) * 32 + lane] = __ldg(h4 + k);
				}
			}
			const float4 *zs4 = reinterpret_cast<const float4 *>(p.ps.zs) + gidx * NQ;
			const float4 *us4 = reinterpret_cast<const float4 *>(p.ps.us) + gidx * NQ;
#pragma unroll
			for (int k = 0; k < NQ; ++k) {
				const float4 vz = __ldcg(zs4 + k), vu = __ldcg(us4 + k);
				z[4 * k] = vz.x, z[4 * k + 1] = vz.y, z[4 * k + 2] = vz.z, z[4 * k + 3] = vz.w;
				u[4 * k] = vu.x, u[4 * k + 1] = vu.y, u[4 * k + 2] = vu.z, u[4 * k + 3] = vu.w;
			}
			done0 = unsigned(__ldcg(p.ps.its + gidx)); // < max_iter
			__syncwarp();
		}

		// ---- factor P + rho I  (replaces matrix_invert, matrix_ops.c:551-630) ----------
#pragma unroll
		for (int i = 0; i < N; ++i)
			a[sidx<N>(i, i)] += rho;
		const bool bad = ldlt_factor<N>(a) || !(rho > 0.0f && rho < 3.0e38f);

		// ---- ADMM iterations (qp_solvers.c:284-310) ------------------------------------
#pragma unroll
		for (int j = 0; j < N; ++j)
			w[j] = 0.0f;
		int my_iters = int(p.max_iter);
		int my_status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		const bool active = valid && !bad && p.max_iter > 0;
		const float alpha = p.alpha;
		const float rho2 = rho * rho, relrho = p.reltol * rho;
		// iterations of this launch: to the end (PASS_PLAIN), the window (PASS_FIRST; the host
		// passes min(window, max_iter)), or the window and what the problem has left (PASS_LIST)
		unsigned it_stop = 0;
		if (LIST) {
			const unsigned left_it = p.max_iter - done0;
			it_stop = left_it < p.ps.window ? left_it : p.ps.window;
		}

		// q lives in shared memory; it is fetched one iteration ahead (at the end of the
		// body) so that the first FMAs of an iteration do not wait on the load
		float4 qn[NQ];
#pragma unroll
		for (int k = 0; k < NQ; ++k)
			qn[k] = lds128(&priv_vec4[k * 32 + lane]);

		// The body holds no warp-wide operation, so lanes simply leave the loop when their
		// problem is done and wait for the others at the __syncwarp() behind it (no vote,
		// no per-lane flag to test: ~10 instructions less per iteration).
		bool conv = false;
		if (active) {
			do {
				// admm_update_x (:146-159): y = rho (z - u) - q ; x = (P + rho I)^-1 y
#pragma unroll
				for (int k = 0; k < NQ; ++k) {
					const float4 qv = qn[k];
					w[4 * k + 0] = fmaf(rho, z[4 * k + 0] - u[4 * k + 0], -qv.x);
					w[4 * k + 1] = fmaf(rho, z[4 * k + 1] - u[4 * k + 1], -qv.y);
					w[4 * k + 2] = fmaf(rho, z[4 * k + 2] - u[4 * k + 2], -qv.z);
					w[4 * k + 3] = fmaf(rho, z[4 * k + 3] - u[4 * k + 3], -qv.w);
				}
				ldlt_solve<N>(a, w);

				float rr = 0.f, ss = 0.f, nx = 0.f, nz = 0.f, nu = 0.f;
#pragma unroll
				for (int k = 0; k < NQ; ++k) {
					float lo[4], hi[4];
					if (BOUNDS) {
						const float4 l4 = priv_vec4[(NQ + k) * 32 + lane];
						const float4 h4 = priv_vec4[(2 * NQ + k) * 32 + lane];
						lo[0] = l4.x, lo[1] = l4.y, lo[2] = l4.z, lo[3] = l4.w;
						hi[0] = h4.x, hi[1] = h4.y, hi[2] = h4.z, hi[3] = h4.w;
					} else {
#pragma unroll
						for (int e = 0; e < 4; ++e)
							lo[e] = p.lb_s, hi[e] = p.ub_s;
					}
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						const int j = 4 * k + e;
						const float xj = w[j];
						// admm_update_x_hat (:161-174)
						const float xh =
							RELAX ? fmaf(alpha, xj, (1.0f - alpha) * z[j]) : xj;
						// admm_update_z (:176-190): clip(x_hat + u)
						const float zn = fminf(fmaxf(xh + u[j], lo[e]), hi[e]);
						// admm_s_norm (:218-231) uses z - z_old
						const float dz = zn - z[j];
						ss = fmaf(dz, dz, ss);
						// admm_update_u (:192-203)
						u[j] += xh - zn;
						// admm_r_norm (:205-216) uses x - z
						const float dr = xj - zn;
						rr = fmaf(dr, dr, rr);
						nx = fmaf(xj, xj, nx);
						nz = fmaf(zn, zn, nz);
						nu = fmaf(u[j], u[j], nu);
						z[j] = zn;
					}
				}
				++it;
				// (n <= 8 has the registers to keep q: loaded once before the loop)
				if (N > 8) {
#pragma unroll
					for (int k = 0; k < NQ; ++k)
						qn[k] = lds128(&priv_vec4[k * 32 + lane]);
				}
				// admm_eps_pri (:233-243), admm_eps_dual (:245-253), stop test (:308):
				//   ||x-z|| < eps_pri  and  rho ||z-z_old|| < eps_dual,
				// compared as squares so that only the two norms inside the thresholds
				// need a square root (MUFU.SQRT, no branches)
				const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
				const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
				conv = (rr < eps_pri * eps_pri) && (rho2 * ss < eps_dual * eps_dual);
			} while (!conv && it < (MODE == PASS_PLAIN ? p.max_iter : (LIST ? it_stop : p.ps.window)));
		}
		__syncwarp();
		if (conv) {
			my_status = QP_STATUS_CONVERGED;
			my_iters = int(done0 + it);
		}

		// ---- the window closed on an unfinished problem: state to scratch, index to the list ---
		bool evict = false;
		unsigned em = 0;
		if (MODE != PASS_PLAIN) {
			if (p.ps.next_list)
				evict = active && !conv && done0 + it < p.max_iter;
			em = __ballot_sync(FULL, evict);
			if (em) {
				if (evict) {
					float4 *zs4 = reinterpret_cast<float4 *>(p.ps.zs) + gidx * NQ;
					float4 *us4 = reinterpret_cast<float4 *>(p.ps.us) + gidx * NQ;
#pragma unroll
					for (int k = 0; k < NQ; ++k) {
						zs4[k] = make_float4(z[4 * k], z[4 * k + 1], z[4 * k + 2], z[4 * k + 3]);
						us4[k] = make_float4(u[4 * k], u[4 * k + 1], u[4 * k + 2], u[4 * k + 3]);
					}
					p.ps.its[gidx] = int(done0 + it);
				}
				unsigned base = 0;
				if (lane == 0) {
					base = atomicAdd(p.ps.next_count, unsigned(__popc(em)));
					if (p.ps.stats)
						atomicAdd(p.ps.stats + (LIST ? 2 : 0), (unsigned long long)__popc(em));
				}
				base = __shfl_sync(FULL, base, 0);
				if (evict)
					p.ps.next_list[base + __popc(em & ((1u << lane) - 1u))] = unsigned(gidx);
			}
		}

		// ---- results ---------------------------------------------------------------
		float nf = 0.0f;
#pragma unroll
		for (int j = 0; j < N; ++j)
			nf = fmaf(w[j], 0.0f, nf);
		if (bad) {
			my_status = QP_STATUS_SINGULAR;
			my_iters = 0;
		} else if (!(nf == 0.0f)) {
			my_status = QP_STATUS_NONFINITE;
		}
		const bool mine = valid && !evict;
		if (mine) {
			if (p.status)
				p.status[gidx] = my_status;
			if (p.iters)
				p.iters[gidx] = my_iters;
		}
		if (MODE == PASS_PLAIN) {
			// the probe of QP_SCHEDULE_AUTO: what my problem needed and what its warp waited for
			// (per-lane sums, added up once at the end of the kernel; the host keeps them from
			// overflowing 32 bits by not probing absurd max_iter x count combinations)
			if (p.ps.stats) {
				const unsigned mine_it = valid ? unsigned(my_iters) : 0u;
				const unsigned mx = __reduce_max_sync(FULL, mine_it);
				st_it += mine_it;
				st_wait += valid ? mx : 0u;
				st_n += valid ? 1u : 0u;
			}
		}

		if (!LIST && em == 0u) {
			// previous bulk stores must have finished reading priv_out
			if (out_pending) {
				if (lane == 0)
					bulk_wait_read();
				__syncwarp();
				out_pending = false;
			}
			float4 *out4 = reinterpret_cast<float4 *>(priv_out);
#pragma unroll
			for (int k = 0; k < NQ; ++k) {
				out4[lane * NQ + k] =
					make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
				out4[8 * N + lane * NQ + k] =
					make_float4(z[4 * k], z[4 * k + 1], z[4 * k + 2], z[4 * k + 3]);
				out4[16 * N + lane * NQ + k] =
					make_float4(rho * u[4 * k], rho * u[4 * k + 1], rho * u[4 * k + 2],
						    rho * u[4 * k + 3]);
			}
			fence_async_smem();
			__syncwarp();
			if (lane == 0) {
				const uint32_t ob = uint32_t(cnt) * N * 4u;
				if (p.x)
					bulk_s2g(p.x + first * N, priv_out, ob);
				if (p.z)
					bulk_s2g(p.z + first * N, priv_out + 32 * N, ob);
				if (p.lam)
					bulk_s2g(p.lam + first * N, priv_out + 2 * 32 * N, ob);
				bulk_commit();
			}
			out_pending = true;
		} else if (mine) {
			// a batch with evicted lanes, or listed (scattered) problems: per-lane stores.
			// (Nothing of an evicted problem is written here; whoever finishes it writes it.)
			float4 *x4 = reinterpret_cast<float4 *>(p.x) + gidx * NQ;
			float4 *z4 = reinterpret_cast<float4 *>(p.z) + gidx * NQ;
			float4 *l4 = reinterpret_cast<float4 *>(p.lam) + gidx * NQ;
#pragma unroll
			for (int k = 0; k < NQ; ++k) {
				if (p.x)
					x4[k] = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
				if (p.z)
					z4[k] = make_float4(z[4 * k], z[4 * k + 1], z[4 * k + 2], z[4 * k + 3]);
				if (p.lam)
					l4[k] = make_float4(rho * u[4 * k], rho * u[4 * k + 1], rho * u[4 * k + 2],
							    rho * u[4 * k + 3]);
			}
		}
	}
	// drain outstanding bulk stores before the CTA's shared memory goes away
	if (lane == 0)
		bulk_wait_read();
	__syncwarp();
	if (MODE == PASS_PLAIN) {
		if (p.ps.stats) { // one add per warp to the CTA's sums, one per CTA to the device's
			const unsigned a = __reduce_add_sync(FULL, st_it), b = __reduce_add_sync(FULL, st_wait),
				       c = __reduce_add_sync(FULL, st_n);
			if (lane == 0) {
				// (the per-lane sums fit 32 bits, their sum over 32 lanes may not: widen first)
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
