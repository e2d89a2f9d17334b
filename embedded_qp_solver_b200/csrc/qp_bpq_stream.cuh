// qp_bpq_stream.cuh -- the second launch of a pass schedule for the four-rows-per-lane kernels:
// every lane group takes its NEXT problem as soon as its current one stops.
//
// What it replaces: the iteration loop of the reference's admm()
// (/root/reference/qp_solvers/qp_solvers.c:284-310) for the problems a first pass
// (qp_bpq.cuh, PASS_FIRST) left unfinished -- the work queue SURVEY.md section 7 step 6 asks for,
// at lane-group granularity.
//
// Why: in qp_bpq.cuh the G problems of a warp (three at n = 40) run until the slowest one stops.
// On BASELINE.json's condensed-MPC set (mean 35.8 iterations, median 25, maximum 2 783) the mean
// over warps of the slowest of three is 62 iterations: 42 % of the lane-iterations wait.  A list
// pass that deals G stragglers to a warp has the same problem one level down, and a group that
// fetched its next problem straight from HBM stalled the whole warp for the round trip
// (profiles/README.md "Passes", rejected).  Here
//   * the first pass has saved each unfinished problem's inverse R = (P + rho I)^-1 (4 n^2
//     bytes), z, u and iteration count in the scratch arena (PassArgs);
//   * each lane group owns a shared-memory SLOT (R, q, lb, ub, z, u, x: 7.4 KB at n = 40) that the
//     TMA engine fills with the group's next problem (cp.async.bulk, one mbarrier per group)
//     WHILE the group iterates on its current one; when that one stops the group writes its
//     results, moves the slot into registers (40 128-bit shared-memory loads per lane) and
//     starts the copy of the problem after that -- no lane ever waits for HBM;
//   * tickets come from one device-wide counter (dynamic: the stragglers' iteration counts
//     span three orders of magnitude), requested a problem ahead, and the list entry behind a
//     ticket is read a problem ahead as well, so neither round trip is on anybody's path --
//     except behind an expected straggler, where nothing is bound ahead (see the kernel);
//   * the groups of a warp stay in ONE instruction stream: one loop trip = one ADMM iteration
//     of every group that has a problem.  The refill block (results out, slot -> registers)
//     is the only divergent code, nested inside a branch that one warp-wide vote per trip makes
//     uniform (see the kernel: convergence is not left to the compiler).  The
//     x-update a problem's first iteration here needs comes from the first pass, whose
//     software-pipelined loop had computed it when the window closed.
// The arithmetic per problem is the plain kernel's, bit for bit (tests/test_passes_gpu.py).
#pragma once

#include "qp_bpq.cuh"

namespace qpb {

template <int V> struct IC { static constexpr int value = V; }; // a compile-time int as a lambda argument

template <int NP> struct BpqStreamCfg {
	using C = BpqCfg<NP>;
	static constexpr int T = C::T, G = C::G;
	static constexpr int NGRP = BPQ_WARPS * G;                // lane groups per CTA
	static constexpr int RAW = NP * NP + 6 * NP + 2 * NP;     // R, q, lb, ub, z, u, x, two y buffers
	// consecutive groups start T 16-byte units apart (mod 128 bytes): lane l of the warp then
	// reads unit l (mod 8) of a row of banks, whatever group it is in
	static constexpr int STAG = ((T * 16) % 128) ? (T * 16) % 128 : 16;
	static constexpr int GROUP_BYTES = (RAW * 4 + 127) / 128 * 128 + STAG;
	static constexpr int BAR_BYTES = (NGRP * 8 + 127) / 128 * 128;
	static constexpr size_t BYTES = size_t(BAR_BYTES) + size_t(NGRP) * GROUP_BYTES;
	static constexpr int FIT = int((227u * 1024u) / (BYTES + 1024));
	// (one CTA per SM fewer than qp_bpq.cuh at n <= 24: the pipeline state needs the registers)
	static constexpr int WANT = NP <= 24 ? C::MINB - 1 : C::MINB;
	static constexpr int MINB = FIT < WANT ? (FIT < 1 ? 1 : FIT) : WANT;
};

// BOUNDS: lb / ub are arrays (eight registers per lane); otherwise the scalar box is read from
// the constant bank where it is used -- the loop sits at the register limit, and with those
// eight registers free the compiler schedules it noticeably better.
// FRESH: shared matrices (qp_batch.p_share).  There was no first pass over the problems, only
// one over the matrices (qp_bpq.cuh, PASS_SETUP): every problem of the batch is taken by ticket
// (no list), starts from its warm start or from zero with the inverse of its group, and the
// x-update of its first iteration is computed in the refill block.
template <int NP, bool BOUNDS, bool FRESH>
__global__ void __launch_bounds__(BPQ_WARPS * 32, BpqStreamCfg<NP>::MINB)
qp_bpq_stream_kernel(const KernelArgs p)
{
	using C = BpqCfg<NP>;
	using SC = BpqStreamCfg<NP>;
	constexpr int T = C::T, G = C::G;
	constexpr unsigned FULL = 0xffffffffu;
	constexpr unsigned NONE = 0xffffffffu; // "no problem"
	extern __shared__ __align__(128) unsigned char smem_raw[];

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int grp = lane / T;     // group inside the warp (== G for the idle lanes)
	const bool member = grp < G;
	const int g = member ? lane - grp * T : 0; // lane inside its group
	const int gbase = member ? grp * T : 0;
	const unsigned am = member ? (((1u << T) - 1u) << gbase) : 0u;
	const GroupRing<T> ring(gbase, g);

	const int cg = warp * G + (member ? grp : 0); // group inside the CTA
	float *slot = reinterpret_cast<float *>(smem_raw + SC::BAR_BYTES + size_t(cg) * SC::GROUP_BYTES);
	float *sR = slot, *sq = sR + NP * NP, *slb = sq + NP, *sub = slb + NP, *sz = sub + NP, *su = sz + NP;
	float *sx = su + NP;
	float *ybuf = sx + NP; // two buffers of NP floats
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw) + cg;
	if (member && g == 0)
		mbar_init(bar, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();

	const float alpha = p.alpha;
	const bool relax = (alpha != 1.0f);
	// two lists: the problems the first pass expects to run long (their residuals had almost
	// stopped shrinking when the window closed) are handed out first, so that the longest
	// dependent chain of the batch starts at once instead of whenever its ticket comes up
	const unsigned total1 = FRESH ? 0u : __ldcg(p.ps.list_count);
	const unsigned total = FRESH ? unsigned(p.count) : total1 + (p.ps.list2 ? __ldcg(p.ps.list_count2) : 0u);
	const unsigned ngroups = gridDim.x * BPQ_WARPS * G;
	// the first problem of a group is dealt (warp-major, as everywhere); after that tickets
	const unsigned gid = (warp * gridDim.x + blockIdx.x) * G + grp;
	// bytes of one slot fill: R, q, (lb, ub), z, u, (x) -- a fresh problem without a warm start
	// brings neither z nor u
	const uint32_t tx = uint32_t(NP * NP + NP + (BOUNDS ? 2 * NP : 0) +
				     (FRESH ? (p.z0 ? 2 * NP : 0) : 3 * NP)) * 4u;

	// start the copy of problem `prob` into the slot (the group's first lane)
	auto issue = [&](unsigned prob) {
		const unsigned long long o = (unsigned long long)prob * NP;
		mbar_expect_tx(bar, tx);
		bulk_g2s(sR, p.ps.Rs + (FRESH ? p_index(p, prob) : (unsigned long long)prob) * (NP * NP),
			 NP * NP * 4, bar);
		bulk_g2s(sq, p.q + o, NP * 4, bar);
		if (BOUNDS) {
			bulk_g2s(slb, p.lb + o, NP * 4, bar);
			bulk_g2s(sub, p.ub + o, NP * 4, bar);
		}
		if (FRESH) {
			if (p.z0) {
				bulk_g2s(sz, p.z0 + o, NP * 4, bar);
				bulk_g2s(su, p.u0 + o, NP * 4, bar);
			}
		} else {
			bulk_g2s(sz, p.ps.zs + o, NP * 4, bar);
			bulk_g2s(su, p.ps.us + o, NP * 4, bar);
			bulk_g2s(sx, p.ps.xs + o, NP * 4, bar);
		}
	};

	// The look-ahead of a group (every value is uniform inside the group unless noted):
	//   slot_prob  the problem on its way into the slot; rho_nxt, its_nxt belong to it
	//   p_nn       the problem after that (its list entry: a load in flight since the last refill)
	//   t_req      the ticket after that (an atomic in flight since the last refill; first lane)
	// A problem that is bound to a group waits for the group: behind a straggler that is the end
	// of the launch plus the follower's own iterations (and the condensed-MPC set, tiled, has
	// stragglers following stragglers).  So a group looks ahead only while its current problem
	// comes from the second list (expected short: at most a couple of hundred iterations);
	// behind a problem of the first list nothing is bound, and the refill after it fetches
	// ticket, list entry and data in turn (three round trips, once per expected straggler).
	// State is kept in few plain words (the hot loop sits at the register limit): bit 31 of a
	// problem index says "from the second list", t_req is a ticket, NO_TICKET (none requested) or
	// DRY (the lists are exhausted: no more atomics).
	constexpr unsigned SHORT = 0x80000000u, NO_TICKET = 0xffffffffu, DRY = 0xfffffffeu;
	unsigned cur = NONE, slot_prob = NONE, p_nn = NONE;
	unsigned t_req = member ? gid : DRY; // the first ticket of a group is dealt (warp-major)
	float rho_nxt = p.rho;
	int its_nxt = 0;
	uint32_t ph = 0;
	// the problem behind the next ticket (NONE when the lists are exhausted); blocks on the
	// atomic unless the ticket was requested at the last refill
	auto from_ticket = [&]() {
		if (t_req == NO_TICKET && g == 0)
			t_req = ngroups + atomicAdd(p.ps.stream_ctr, 1u);
		const unsigned t = __shfl_sync(am, t_req, gbase);
		if (t >= total) {
			t_req = DRY;
			return NONE;
		}
		t_req = NO_TICKET;
		if (FRESH)
			return t | SHORT; // no list, and nothing is known about anybody: look ahead
		return t < total1 ? __ldcg(p.ps.list + t) : (__ldcg(p.ps.list2 + (t - total1)) | SHORT);
	};
	// the next problem in line: the list entry read at the last refill, else a ticket's
	auto next_problem = [&]() {
		unsigned idx = p_nn;
		p_nn = NONE;
		if (idx == NONE)
			idx = from_ticket();
		return idx;
	};
	auto stage = [&](unsigned tagged) {
		const unsigned idx = tagged & ~SHORT;
		if (g == 0)
			issue(idx);
		rho_nxt = (!FRESH && p.rho_v) ? __ldg(p.rho_v + idx) : p.rho;
		// (FRESH: its[] holds the verdict on the group's matrix, 1 = not positive definite)
		its_nxt = __ldcg(p.ps.its + (FRESH ? p_index(p, idx) : (unsigned long long)idx));
		slot_prob = tagged;
	};

	float2 S[2][NP]; // S[rp][j] = (row 4g+2rp, row 4g+2rp+1) of R at column j
	float q[4], lo[BOUNDS ? 4 : 1], hi[BOUNDS ? 4 : 1], z[4], u[4], x[4];
	float xr[4] = { 0.f, 0.f, 0.f, 0.f };
	float rho = p.rho, rho2 = 0.f, relrho = 0.f;
	unsigned it = 0, done0 = 0;
	bool conv = false;
	unsigned st_it = 0; // PassArgs::stats[6], [7]: per group
	unsigned st_n = 0;
	bool fin = member; // "my group wants a problem"

	auto publish_y = [&](float *yb) {
		*reinterpret_cast<float4 *>(yb + 4 * g) =
			make_float4(fmaf(rho, z[0] - u[0], -q[0]), fmaf(rho, z[1] - u[1], -q[1]),
				    fmaf(rho, z[2] - u[2], -q[2]), fmaf(rho, z[3] - u[3], -q[3]));
	};
	// The mat-vec of the next iteration and the five group reductions of this one, interleaved
	// BY HAND: a shuffle stage is issued, a few column chunks of the mat-vec follow, then the
	// stage's additions and the next stage's shuffles.  Left to itself ptxas sometimes sinks the
	// whole shuffle/add network behind the 80 FFMA2 (every FADD then waits for its SHFL: one
	// straggler's iteration took 0.48 instead of 0.35 us, and the stragglers are this kernel's
	// critical path).  The arithmetic is GroupRing::sum's and matvec's, operation for operation
	// (results are compared bit for bit with the plain kernel).
	auto matvec_and_sums = [&](const float *yb, float (&xo)[4], float &rr, float &ss, float &nx,
				   float &nz, float &nu) {
		constexpr int NC = NP / 4;
		float v[5] = { rr, ss, nx, nz, nu };
		float2 acc[2][4];
#pragma unroll
		for (int rp = 0; rp < 2; ++rp)
			acc[rp][0] = acc[rp][1] = acc[rp][2] = acc[rp][3] = make_float2(0.0f, 0.0f);
		auto chunks = [&](auto c0, auto c1) {
#pragma unroll
			for (int c = decltype(c0)::value; c < decltype(c1)::value; ++c) {
				const int j = 4 * c;
				const float4 y = *reinterpret_cast<const float4 *>(yb + j);
#pragma unroll
				for (int rp = 0; rp < 2; ++rp) {
					acc[rp][0] = fma2_bc(S[rp][j], y.x, acc[rp][0]);
					acc[rp][1] = fma2_bc(S[rp][j + 1], y.y, acc[rp][1]);
					acc[rp][2] = fma2_bc(S[rp][j + 2], y.z, acc[rp][2]);
					acc[rp][3] = fma2_bc(S[rp][j + 3], y.w, acc[rp][3]);
				}
			}
		};
		if constexpr ((T & (T - 1)) == 0) {
			// xor butterfly, log2(T) stages, the chunks spread evenly between them
			constexpr int NS = T == 8 ? 3 : (T == 4 ? 2 : 1);
			float t[5];
#pragma unroll
			for (int i = 0; i < 5; ++i)
				t[i] = __shfl_xor_sync(am, v[i], T / 2);
			chunks(IC<0>{}, IC<NC / (NS + 1)>{});
#pragma unroll
			for (int i = 0; i < 5; ++i)
				v[i] += t[i];
			if constexpr (NS >= 2) {
#pragma unroll
				for (int i = 0; i < 5; ++i)
					t[i] = __shfl_xor_sync(am, v[i], T / 4);
				chunks(IC<NC / (NS + 1)>{}, IC<2 * NC / (NS + 1)>{});
#pragma unroll
				for (int i = 0; i < 5; ++i)
					v[i] += t[i];
			}
			if constexpr (NS >= 3) {
#pragma unroll
				for (int i = 0; i < 5; ++i)
					t[i] = __shfl_xor_sync(am, v[i], T / 8);
				chunks(IC<2 * NC / (NS + 1)>{}, IC<3 * NC / (NS + 1)>{});
#pragma unroll
				for (int i = 0; i < 5; ++i)
					v[i] += t[i];
			}
			chunks(IC<NS * NC / (NS + 1)>{}, IC<NC>{});
		} else {
			// T = 2 * odd: fold the halves, then the doubling ring (GroupRing::sum)
			constexpr int Hh = T / 2;
			float f[5], w[5], s1[5], sx[5], a[5];
#pragma unroll
			for (int i = 0; i < 5; ++i)
				f[i] = __shfl_sync(am, v[i], ring.fold);
			chunks(IC<0>{}, IC<NC / 4>{});
#pragma unroll
			for (int i = 0; i < 5; ++i) {
				w[i] = v[i] + f[i];
				s1[i] = __shfl_sync(am, w[i], ring.n1);
				sx[i] = __shfl_sync(am, w[i], Hh == 3 ? ring.n2 : ring.n4);
			}
			chunks(IC<NC / 4>{}, IC<NC / 2>{});
#pragma unroll
			for (int i = 0; i < 5; ++i)
				a[i] = w[i] + s1[i];
			if constexpr (Hh == 3) {
#pragma unroll
				for (int i = 0; i < 5; ++i)
					v[i] = a[i] + sx[i];
				chunks(IC<NC / 2>{}, IC<NC>{});
			} else {
				float s2[5];
#pragma unroll
				for (int i = 0; i < 5; ++i)
					s2[i] = __shfl_sync(am, a[i], ring.n2);
				chunks(IC<NC / 2>{}, IC<3 * NC / 4>{});
#pragma unroll
				for (int i = 0; i < 5; ++i) {
					a[i] += s2[i];
					v[i] = a[i] + sx[i];
				}
				chunks(IC<3 * NC / 4>{}, IC<NC>{});
			}
		}
		rr = v[0], ss = v[1], nx = v[2], nz = v[3], nu = v[4];
#pragma unroll
		for (int rp = 0; rp < 2; ++rp) {
			const float2 tt = __fadd2_rn(__fadd2_rn(acc[rp][0], acc[rp][1]),
						     __fadd2_rn(acc[rp][2], acc[rp][3]));
			xo[2 * rp] = tt.x, xo[2 * rp + 1] = tt.y;
		}
	};
	// admm_update_x_hat (:161-174), _z (:176-190), _u (:192-203) and the five partial sums of
	// squares (:205-253), as in qp_bpq.cuh
	auto elementwise = [&](const float (&xv)[4], float &rr, float &ss, float &nx, float &nz,
			       float &nu) {
		rr = ss = nx = nz = nu = 0.f;
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			const float xs = xv[s];
			x[s] = xs;
			const float xh = relax ? fmaf(alpha, xs, (1.0f - alpha) * z[s]) : xs;
			const float zn = fminf(fmaxf(xh + u[s], BOUNDS ? lo[BOUNDS ? s : 0] : p.lb_s), BOUNDS ? hi[BOUNDS ? s : 0] : p.ub_s);
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
	auto converged = [&](float rr, float ss, float nx, float nz, float nu) {
		const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
		const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
		int c = (rr < eps_pri * eps_pri) && (rho2 * ss < eps_dual * eps_dual);
		if constexpr ((T & (T - 1)) != 0)
			c = __shfl_sync(am, c, gbase);
		return c != 0;
	};

	// The groups of a warp must stay in ONE instruction stream: three streams of ten lanes cost
	// three times the issue slots.  Reconvergence behind a divergent branch is the compiler's
	// business, and it cannot be relied on here: builds in which it had threaded the exits of
	// the refill block into copies of the loop body ran 1.2 G instead of 0.7 G warp instructions
	// (5 ms instead of 1.5), and a __syncwarp() of the live lanes did not bring the groups back
	// together either.  So every trip starts with ONE warp-wide vote "does any group want a
	// refill": a vote is executed by all lanes at once by definition, and it makes the branch
	// around the refill block uniform -- the divergent part is nested inside it and has
	// reconverged when the block ends.  Groups without work (and the idle lanes) keep voting
	// until the whole warp is done.
	for (;;) {
		if (__ballot_sync(FULL, fin)) {
			if (fin) {
				// ---- results of the problem that has just stopped ------------------------------
				if (cur != NONE) {
					float nf = fmaf(x[0], 0.0f, fmaf(x[1], 0.0f, fmaf(x[2], 0.0f, x[3] * 0.0f)));
					nf = ring.sum(nf, am);
					int status = conv ? QP_STATUS_CONVERGED : QP_STATUS_MAX_ITER;
					if (!(nf == 0.0f))
						status = QP_STATUS_NONFINITE;
					const unsigned long long o4 = (unsigned long long)cur * (NP / 4) + g;
					if (p.x)
						reinterpret_cast<float4 *>(p.x)[o4] = make_float4(x[0], x[1], x[2], x[3]);
					if (p.z)
						reinterpret_cast<float4 *>(p.z)[o4] = make_float4(z[0], z[1], z[2], z[3]);
					if (p.lam)
						reinterpret_cast<float4 *>(p.lam)[o4] =
							make_float4(rho * u[0], rho * u[1], rho * u[2], rho * u[3]);
					if (g == 0) {
						if (p.status)
							p.status[cur] = status;
						if (p.iters)
							p.iters[cur] = int(done0 + it);
					}
					st_it += done0 + it, ++st_n;
				}
				// ---- the next problem: slot -> registers ---------------------------------------
				// (a loop only for FRESH: a problem whose matrix is not positive definite gets
				// its verdict here and the group takes another one)
				for (;;) {
				if (slot_prob == NONE) {
					const unsigned nx = next_problem();
					if (nx != NONE)
						stage(nx);
				}
				const bool cur_short = (slot_prob & SHORT) != 0u && slot_prob != NONE;
				cur = slot_prob == NONE ? NONE : (slot_prob & ~SHORT);
				slot_prob = NONE;
				if (cur == NONE)
					break;
				{
					mbar_wait(bar, ph);
					ph ^= 1u;
					const float4 *R4 = reinterpret_cast<const float4 *>(sR) + g;
#pragma unroll
					for (int j = 0; j < NP; ++j) {
						const float4 v = R4[j * T];
						S[0][j] = make_float2(v.x, v.y), S[1][j] = make_float2(v.z, v.w);
					}
					const float4 vq = reinterpret_cast<const float4 *>(sq)[g];
					const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
					const bool have_zu = !FRESH || p.z0 != nullptr;
					const float4 vz = have_zu ? reinterpret_cast<const float4 *>(sz)[g] : zero4;
					const float4 vu = have_zu ? reinterpret_cast<const float4 *>(su)[g] : zero4;
					if constexpr (!FRESH) {
						// the x-update of the problem's next iteration: the first pass had it
						// already (its loop is software-pipelined like the one below)
						const float4 vx = reinterpret_cast<const float4 *>(sx)[g];
						xr[0] = vx.x, xr[1] = vx.y, xr[2] = vx.z, xr[3] = vx.w;
					}
					q[0] = vq.x, q[1] = vq.y, q[2] = vq.z, q[3] = vq.w;
					if constexpr (BOUNDS) {
						const float4 vl = reinterpret_cast<const float4 *>(slb)[g];
						const float4 vh = reinterpret_cast<const float4 *>(sub)[g];
						lo[0] = vl.x, lo[1] = vl.y, lo[2] = vl.z, lo[3] = vl.w;
						hi[0] = vh.x, hi[1] = vh.y, hi[2] = vh.z, hi[3] = vh.w;
					}
					z[0] = vz.x, z[1] = vz.y, z[2] = vz.z, z[3] = vz.w;
					u[0] = vu.x, u[1] = vu.y, u[2] = vu.z, u[3] = vu.w;
					x[0] = x[1] = x[2] = x[3] = 0.0f;
					rho = rho_nxt;
					rho2 = rho * rho, relrho = p.reltol * rho;
					const bool bad_matrix = FRESH && its_nxt != 0;
					done0 = FRESH ? 0u : unsigned(its_nxt); // < max_iter
					it = 0;
					conv = false;
					// the slot is free: order the reads above before the next copy's writes
					__syncwarp(am);
					fence_async_smem();
					// ---- look ahead, behind an expected short problem only ---------------------
					if (cur_short) {
						const unsigned nx = next_problem();
						if (nx != NONE)
							stage(nx);
						if (t_req != DRY)
							p_nn = from_ticket(); // (the list entry stays in flight until the next refill)
						if (t_req != DRY) {
							if (g == 0)
								t_req = ngroups + atomicAdd(p.ps.stream_ctr, 1u);
							else
								t_req = 0u; // "requested": the first lane holds it
						}
					}
					if constexpr (FRESH) {
						if (bad_matrix || p.max_iter == 0) {
							// what the plain kernel hands back for such a problem: x = 0, the
							// starting z and u, no iteration
							const unsigned long long o4 = (unsigned long long)cur * (NP / 4) + g;
							if (p.x)
								reinterpret_cast<float4 *>(p.x)[o4] = zero4;
							if (p.z)
								reinterpret_cast<float4 *>(p.z)[o4] = vz;
							if (p.lam)
								reinterpret_cast<float4 *>(p.lam)[o4] =
									make_float4(rho * vu.x, rho * vu.y, rho * vu.z, rho * vu.w);
							if (g == 0) {
								if (p.status)
									p.status[cur] = bad_matrix ? QP_STATUS_SINGULAR : QP_STATUS_MAX_ITER;
								if (p.iters)
									p.iters[cur] = 0;
							}
							continue;
						}
						// the x-update of the first iteration (the loop below is software-pipelined)
						float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
						publish_y(ybuf);
						__syncwarp(am);
						matvec_and_sums(ybuf, xr, d0, d1, d2, d3, d4);
					}
				}
				break;
				}
				fin = false;
			}
			if (!__any_sync(FULL, cur != NONE))
				break;
		}
		if (cur != NONE)
		{
			// one iteration (qp_solvers.c:284-310), software-pipelined as in qp_bpq.cuh: xr is
			// the x-update of THIS iteration, computed at the end of the previous trip
			float rr, ss, nx, nz, nu;
			elementwise(xr, rr, ss, nx, nz, nu);
			++it;
			float *yb = ybuf + (it & 1u) * NP; // two buffers: no second barrier needed
			publish_y(yb);
			// The lanes of a group are in step here (the vote at the top of the trip saw to
			// that): their stores above precede their loads below in the warp's one instruction
			// stream, and a compiler barrier is all the ordering needed.  The barrier
			// instruction itself is only executed when the group is NOT all here -- in a lone
			// straggler's trip it was 11 % of the time (branch-resolving stalls behind it).
			if ((__activemask() & am) != am)
				__syncwarp(am);
			else
				asm volatile("" ::: "memory");
			matvec_and_sums(yb, xr, rr, ss, nx, nz, nu);
			conv = converged(rr, ss, nx, nz, nu);
			fin = conv || done0 + it >= p.max_iter;
		}
	}
	// what QP_SCHEDULE_AUTO looks at to go back to single launches
	if (p.ps.stats && member && g == 0 && st_n) {
		atomicAdd(p.ps.stats + 6, (unsigned long long)st_it);
		atomicAdd(p.ps.stats + 7, (unsigned long long)st_n);
	}
}

} // namespace qpb
