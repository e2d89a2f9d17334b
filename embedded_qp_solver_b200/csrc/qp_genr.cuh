// qp_genr.cuh -- general inequality rows  l <= A x <= u  with the operands in REGISTERS:
// n <= 32 variables, m <= 64 rows.
//
// Same iteration as qp_gen.cuh, line by line (the reference's admm(),
// /root/reference/qp_solvers/qp_solvers.c:255-319, for the splitting A x = z; every update
// cites its reference line there).  What changes is where the operands live.  qp_gen.cuh keeps
// P + rho A'A, its inverse R and A in shared memory and every FMA of an iteration waits for a
// 32-bit shared-memory load; here a group of T lanes owns one problem and lane g keeps in
// registers, inside the iteration loop,
//   Br[k][.]   rows j = g + T k (k < MR = MP/T) of B = A R      MR x NP floats
//   Ac[a][.]   columns i = g + T a (a < NR = NP/T) of A         NR x MP floats
// (128 registers at n = 16, m = 32).  The stop test (:205-253) needs A x, not x, so an
// iteration goes from the right-hand side t = rho A'(z - u) - q straight to A x = B t, and
// x = R t is formed once, from the last t, when the problem has stopped (R waits in shared
// memory).  The three vectors of an iteration -- t, z and u -- are published to shared memory
// and read by every lane as 128-bit broadcasts: per iteration and lane (NP + 2 MP) / 4 LDS.128
// feed (MR NP + 2 NR MP) / 2 packed FMAs (FFMA2), against one LDS per FMA before, and two warp
// barriers instead of four.  Setup keeps rows of P + rho A'A, then of R, and rows of A in
// registers: the elimination publishes one pivot row per step, B is built from the rows of R.
// (GenrCfg::FUSE = false, one problem per warp: R and the rows of A stay in registers instead of
// B, x = R t and A x are two steps of an iteration -- see there.)  Problems are handed out by
// ticket when a batch is longer than one wave (k_genr.cu, gen_ticket).
//   <NP, MP, T> = <8, 16, 4>   eight problems per warp
//                 <16, 32, 8>  four problems per warp (the n = 16, m = 32 shape)
//                 <32, 64, 32> one problem per warp
// Smaller n, m are padded inside the kernel: identity rows of P, zero rows and columns of
// A, free rows (-inf, +inf), which leave every sum of the iteration unchanged.  Anything
// larger (n <= 64, m <= 128) stays on qp_gen.cuh.
//
// P + rho A'A is summed over the rows of A in the order qp_gen.cuh uses, and inverted by the
// same Gauss-Jordan arithmetic (pivot row scaled by 1/pivot, fma(-f, p, w) elsewhere), so R
// has the same bits; A x = (A R) t instead of A (R t) and four partial sums per row in the
// mat-vecs, so the iterates differ from qp_gen.cuh in the last bits.  With A = I every product
// with A has one non-zero term, B = R bit for bit, and the kernel reproduces the reference's
// golden vectors under the same bars as the box kernels; beyond A = I there is no reference
// behaviour -- PARITY UNPINNED, as for qp_gen.cuh.
#pragma once

#include "qp_aux_args.h"

namespace qpb {

constexpr int GENR_WARPS = 4; // warps per CTA

template <int NP, int MP, int T> struct GenrCfg {
	static_assert(NP % T == 0 && MP % T == 0 && 32 % T == 0, "group width");
	static_assert(NP % 4 == 0 && MP % 4 == 0 && MP >= 2 * NP, "padded shape");
	static constexpr int G = 32 / T;   // problems per warp
	static constexpr int NR = NP / T;  // coordinates (rows of R, columns of A) per lane
	static constexpr int MR = MP / T;  // rows of A per lane
	static constexpr int LDP = NP + 4; // row stride of the staged [P; A]: 16-byte aligned, and
					   // lanes reading consecutive rows hit different bank quads
	static constexpr int RAW = (NP + MP) * LDP + 2 * NP + 2 * MP;
	// groups start T banks apart: neither the 128-bit broadcasts nor the lanes' 32-bit stores
	// of a warp's groups collide
	static constexpr int GS = RAW + (T % 32 + 32 - RAW % 32) % 32;
	static constexpr size_t SMEM = size_t(GENR_WARPS) * G * GS * sizeof(float);
	// MR NP + NR MP floats of matrices per lane inside the loop: 64 leave room for three CTAs per
	// SM (168 registers); 128 need ~215 registers with the loop's vectors and accumulators
	static constexpr int MINB = (MR * NP + NR * MP) <= 64 ? 3 : 2;
	// B = A R pays when a warp carries several problems.  With one problem per warp (T = 32)
	// building B costs what it saves (n = 32, m = 64: 26.9 M QP/s without, 21.8 M with), so that
	// shape keeps R and the rows of A in registers and does x = R t, A x per iteration.
	static constexpr bool FUSE = G > 1;
};

template <int T> __device__ __forceinline__ float genr_group_sum(float v)
{
#pragma unroll
	for (int o = T / 2; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ float genr_get(const float2 *row, int c)
{
	return (c & 1) ? row[c >> 1].y : row[c >> 1].x;
}

// RELAX = false: alpha == 1, the relaxation (:161-174) is the identity and is not computed
template <int NP, int MP, int T, bool RELAX>
__global__ void __launch_bounds__(GENR_WARPS * 32, GenrCfg<NP, MP, T>::MINB)
qp_genr_kernel(const GenArgs a, unsigned long long *ticket)
{
	using C = GenrCfg<NP, MP, T>;
	constexpr int G = C::G, NR = C::NR, MR = C::MR, LDP = C::LDP, NH = NP / 2, MH = MP / 2;
	constexpr bool FUSE = C::FUSE;
	constexpr unsigned FULL = 0xffffffffu;
	extern __shared__ __align__(16) float gsm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int grp = lane / T, g = lane % T;
	float *sP = gsm + (warp * G + grp) * C::GS; // staged P (padded with identity rows)
	float *sA = sP + NP * LDP;                  // staged A (padded with zeros)
	float *sT = sA + MP * LDP;                  // right-hand side of the x-update
	float *sX = sT + NP, *sZ = sX + NP, *sU = sZ + MP;
	float *sPiv = sZ;                           // the inverse publishes its pivot rows here
	const int n = a.n, m = a.m;
	const float rho = a.rho, alpha = a.alpha;
	const float inf = __int_as_float(0x7f800000);
	const bool vec = (n == NP) &&
			 (((reinterpret_cast<unsigned long long>(a.P) | reinterpret_cast<unsigned long long>(a.A)) & 15ull) == 0ull);

	// Work is handed out G problems at a time.  With a ticket counter (batches longer than one
	// wave) a warp takes its next G problems when it is done with the current ones: the warps'
	// iteration counts differ (the slowest of G problems, 25 ... 60 iterations), and with a static
	// stride the launch ends with its unluckiest warp (2^16 problems over 1 184 warps: 14 rounds
	// each, the slowest warp ~15 % behind the mean).  The first round is dealt by position, the
	// counter starts behind it, and the next ticket is drawn while the current problems are solved.
	const unsigned long long nw = (unsigned long long)gridDim.x * GENR_WARPS * G;
	unsigned long long base = ((unsigned long long)blockIdx.x * GENR_WARPS + warp) * G;
	for (; base < a.count;) {
		unsigned long long next = base + nw;
		if (ticket) {
			if (lane == 0)
				next = nw + atomicAdd(ticket, (unsigned long long)G); // the counter starts at 0
		}
		const bool valid = base + grp < a.count;
		const unsigned long long prob = valid ? base + grp : base; // a tail group repeats the first problem
		// ---- stage P and A (coalesced inside the group) ------------------------------------
		const float *Pp = a.P + prob * (unsigned long long)n * n;
		const float *Ap = a.A + prob * (unsigned long long)m * n;
		if (vec) {
			const float4 *P4 = reinterpret_cast<const float4 *>(Pp), *A4 = reinterpret_cast<const float4 *>(Ap);
#pragma unroll
			for (int e = g; e < NP * NP / 4; e += T) {
				const int r = e / (NP / 4), c = e % (NP / 4);
				*reinterpret_cast<float4 *>(sP + r * LDP + 4 * c) = __ldg(P4 + e);
			}
#pragma unroll
			for (int e = g; e < MP * NP / 4; e += T) {
				const int r = e / (NP / 4), c = e % (NP / 4);
				*reinterpret_cast<float4 *>(sA + r * LDP + 4 * c) =
					r < m ? __ldg(A4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
			}
		} else {
#pragma unroll 4
			for (int e = g; e < NP * NP; e += T) {
				const int r = e / NP, c = e % NP;
				sP[r * LDP + c] = (r < n && c < n) ? __ldg(Pp + r * n + c) : (r == c ? 1.0f : 0.0f);
			}
#pragma unroll 4
			for (int e = g; e < MP * NP; e += T) {
				const int r = e / NP, c = e % NP;
				sA[r * LDP + c] = (r < m && c < n) ? __ldg(Ap + r * n + c) : 0.0f;
			}
		}
		// ---- per-coordinate and per-row state ----------------------------------------------
		const unsigned long long vo = prob * (unsigned long long)n, mo = prob * (unsigned long long)m;
		float q[NR], x[NR], w1[NR], w2[NR]; // w1 = A'z, w2 = A'u, carried along
#pragma unroll
		for (int s = 0; s < NR; ++s) {
			const int i = g + T * s;
			q[s] = i < n ? __ldg(a.q + vo + i) : 0.f;
			x[s] = w1[s] = w2[s] = 0.f;
		}
		float lo[MR], hi[MR], z[MR], u[MR];
#pragma unroll
		for (int k = 0; k < MR; ++k) {
			const int j = g + T * k;
			const bool in = j < m;
			lo[k] = in ? __ldg(a.l + mo + j) : -inf;
			hi[k] = in ? __ldg(a.u + mo + j) : inf;
			z[k] = (in && a.z0) ? __ldg(a.z0 + mo + j) : 0.f;
			u[k] = (in && a.u0) ? __ldg(a.u0 + mo + j) : 0.f;
		}
		__syncwarp();

		// ---- my rows of A into registers -----------------------------------------------------
		float2 Ar[MR][NH];
#pragma unroll
		for (int k = 0; k < MR; ++k) {
			const float4 *row = reinterpret_cast<const float4 *>(sA + (g + T * k) * LDP);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 v = row[c];
				Ar[k][2 * c] = make_float2(v.x, v.y), Ar[k][2 * c + 1] = make_float2(v.z, v.w);
			}
		}

		// ---- my rows of P + rho A'A: sum over the rows of A in order (as qp_gen.cuh) ---------
		float2 Rr[NR][NH];
#pragma unroll
		for (int s = 0; s < NR; ++s) {
#pragma unroll
			for (int c = 0; c < NH; ++c)
				Rr[s][c] = make_float2(0.f, 0.f);
		}
#pragma unroll 2
		for (int k = 0; k < MP; ++k) {
			float aki[NR];
#pragma unroll
			for (int s = 0; s < NR; ++s)
				aki[s] = sA[k * LDP + g + T * s];
			const float4 *row = reinterpret_cast<const float4 *>(sA + k * LDP);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 v = row[c];
#pragma unroll
				for (int s = 0; s < NR; ++s) {
					Rr[s][2 * c] = __ffma2_rn(make_float2(v.x, v.y), make_float2(aki[s], aki[s]), Rr[s][2 * c]);
					Rr[s][2 * c + 1] =
						__ffma2_rn(make_float2(v.z, v.w), make_float2(aki[s], aki[s]), Rr[s][2 * c + 1]);
				}
			}
		}
#pragma unroll
		for (int s = 0; s < NR; ++s) {
			const float4 *row = reinterpret_cast<const float4 *>(sP + (g + T * s) * LDP);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 v = row[c];
				Rr[s][2 * c] = __ffma2_rn(make_float2(rho, rho), Rr[s][2 * c], make_float2(v.x, v.y));
				Rr[s][2 * c + 1] = __ffma2_rn(make_float2(rho, rho), Rr[s][2 * c + 1], make_float2(v.z, v.w));
			}
		}
		__syncwarp(); // sZ (= sPiv) is free: nobody reads the staged matrices any more

		// ---- R = (P + rho A'A)^-1: Gauss-Jordan without pivoting, rows in registers; the owner
		// of row k publishes it scaled by 1/pivot (pivot entry := 1/pivot), everybody subtracts
		// its multiple (two buffers in turn: one barrier per step) -----------------------------
		bool bad = false;
#pragma unroll
		for (int k = 0; k < NP; ++k) {
			const int sk = k / T;        // register slot of row k in its owner
			float *pv = sPiv + (k & 1) * NP;
			if (g == k % T) {
				const float dinv = 1.0f / genr_get(Rr[sk], k);
#pragma unroll
				for (int c = 0; c < NP / 4; ++c) {
					float4 v = make_float4(Rr[sk][2 * c].x * dinv, Rr[sk][2 * c].y * dinv,
							       Rr[sk][2 * c + 1].x * dinv, Rr[sk][2 * c + 1].y * dinv);
					if (k / 4 == c) {
						if (k % 4 == 0) v.x = dinv;
						if (k % 4 == 1) v.y = dinv;
						if (k % 4 == 2) v.z = dinv;
						if (k % 4 == 3) v.w = dinv;
					}
					*reinterpret_cast<float4 *>(pv + 4 * c) = v;
				}
			}
			__syncwarp();
			float2 p[NH];
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 v = *reinterpret_cast<const float4 *>(pv + 4 * c);
				p[2 * c] = make_float2(v.x, v.y), p[2 * c + 1] = make_float2(v.z, v.w);
			}
			const float dinv = genr_get(p, k);
			bad |= !(dinv > 0.0f && dinv < 3.0e38f); // pivot <= 0, NaN or vanished
#pragma unroll
			for (int s = 0; s < NR; ++s) {
				const bool mine = (s == sk) && (g == k % T);
				const float nf = -genr_get(Rr[s], k);
#pragma unroll
				for (int c = 0; c < NH; ++c) {
					float2 w = __ffma2_rn(p[c], make_float2(nf, nf), Rr[s][c]);
					if (2 * c == k) w.x = nf * dinv;
					if (2 * c + 1 == k) w.y = nf * dinv;
					Rr[s][c] = mine ? p[c] : w;
				}
			}
		}
		__syncwarp();

		// ---- R to shared memory (over the staged P), B = A R for my rows of A, then my columns of
		// A into registers -----------------------------------------------------------------------
		float *sR = sP;
		float2 Br[MR][NH]; // B = A R (FUSE), else A itself
		if constexpr (FUSE) {
#pragma unroll
		for (int s = 0; s < NR; ++s) {
#pragma unroll
			for (int c = 0; c < NP / 4; ++c)
				*reinterpret_cast<float4 *>(sR + (g + T * s) * LDP + 4 * c) =
					make_float4(Rr[s][2 * c].x, Rr[s][2 * c].y, Rr[s][2 * c + 1].x, Rr[s][2 * c + 1].y);
		}
		__syncwarp();
#pragma unroll
		for (int k = 0; k < MR; ++k) {
#pragma unroll
			for (int c = 0; c < NH; ++c)
				Br[k][c] = make_float2(0.f, 0.f);
		}
#pragma unroll
		for (int i = 0; i < NP; ++i) { // B[j][.] += A[j][i] R[i][.]
			const float4 *row = reinterpret_cast<const float4 *>(sR + i * LDP);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 v = row[c];
#pragma unroll
				for (int k = 0; k < MR; ++k) {
					const float e = genr_get(Ar[k], i);
					Br[k][2 * c] = __ffma2_rn(make_float2(v.x, v.y), make_float2(e, e), Br[k][2 * c]);
					Br[k][2 * c + 1] = __ffma2_rn(make_float2(v.z, v.w), make_float2(e, e), Br[k][2 * c + 1]);
				}
			}
		}
		} else {
#pragma unroll
		for (int k = 0; k < MR; ++k) {
#pragma unroll
			for (int c = 0; c < NH; ++c)
				Br[k][c] = Ar[k][c];
		}
		}
		float2 Ac[NR][MH];
#pragma unroll
		for (int s = 0; s < NR; ++s) {
#pragma unroll
			for (int k = 0; k < MH; ++k)
				Ac[s][k] = make_float2(sA[(2 * k) * LDP + g + T * s], sA[(2 * k + 1) * LDP + g + T * s]);
		}

		// ---- helpers: products with the register-resident matrices, vector from shared memory
		auto r_times = [&](const float *v, float (&o)[NR]) { // my rows of R times v[0..NP)
			float2 acc[NR][2];
#pragma unroll
			for (int s = 0; s < NR; ++s)
				acc[s][0] = acc[s][1] = make_float2(0.f, 0.f);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 t = *reinterpret_cast<const float4 *>(v + 4 * c);
#pragma unroll
				for (int s = 0; s < NR; ++s) {
					acc[s][0] = __ffma2_rn(Rr[s][2 * c], make_float2(t.x, t.y), acc[s][0]);
					acc[s][1] = __ffma2_rn(Rr[s][2 * c + 1], make_float2(t.z, t.w), acc[s][1]);
				}
			}
#pragma unroll
			for (int s = 0; s < NR; ++s)
				o[s] = (acc[s][0].x + acc[s][1].x) + (acc[s][0].y + acc[s][1].y);
		};
		auto b_times = [&](const float *v, float (&o)[MR]) { // my rows of B (or A) times v[0..NP)
			float2 acc[MR][2];
#pragma unroll
			for (int k = 0; k < MR; ++k)
				acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
#pragma unroll
			for (int c = 0; c < NP / 4; ++c) {
				const float4 t = *reinterpret_cast<const float4 *>(v + 4 * c);
#pragma unroll
				for (int k = 0; k < MR; ++k) {
					acc[k][0] = __ffma2_rn(Br[k][2 * c], make_float2(t.x, t.y), acc[k][0]);
					acc[k][1] = __ffma2_rn(Br[k][2 * c + 1], make_float2(t.z, t.w), acc[k][1]);
				}
			}
#pragma unroll
			for (int k = 0; k < MR; ++k)
				o[k] = (acc[k][0].x + acc[k][1].x) + (acc[k][0].y + acc[k][1].y);
		};
		// my coordinates of A'v and A'w, v and w in shared memory [0..MP)
		auto at_times2 = [&](const float *v, const float *w, float (&ov)[NR], float (&ow)[NR]) {
			float2 av[NR][2], aw[NR][2];
#pragma unroll
			for (int s = 0; s < NR; ++s)
				av[s][0] = av[s][1] = aw[s][0] = aw[s][1] = make_float2(0.f, 0.f);
#pragma unroll
			for (int c = 0; c < MP / 4; ++c) {
				const float4 tv = *reinterpret_cast<const float4 *>(v + 4 * c);
				const float4 tw = *reinterpret_cast<const float4 *>(w + 4 * c);
#pragma unroll
				for (int s = 0; s < NR; ++s) {
					av[s][0] = __ffma2_rn(Ac[s][2 * c], make_float2(tv.x, tv.y), av[s][0]);
					av[s][1] = __ffma2_rn(Ac[s][2 * c + 1], make_float2(tv.z, tv.w), av[s][1]);
					aw[s][0] = __ffma2_rn(Ac[s][2 * c], make_float2(tw.x, tw.y), aw[s][0]);
					aw[s][1] = __ffma2_rn(Ac[s][2 * c + 1], make_float2(tw.z, tw.w), aw[s][1]);
				}
			}
#pragma unroll
			for (int s = 0; s < NR; ++s) {
				ov[s] = (av[s][0].x + av[s][1].x) + (av[s][0].y + av[s][1].y);
				ow[s] = (aw[s][0].x + aw[s][1].x) + (aw[s][0].y + aw[s][1].y);
			}
		};

		// A'z and A'u of the starting point
		if (a.z0 || a.u0) {
#pragma unroll
			for (int k = 0; k < MR; ++k)
				sZ[g + T * k] = z[k], sU[g + T * k] = u[k];
			__syncwarp();
			at_times2(sZ, sU, w1, w2);
			__syncwarp();
		}

		// ---- iterations: the warp's groups in lockstep; a group that has stopped keeps
		// executing but no longer commits its state.  The loop is software-pipelined: A x of
		// the NEXT iteration (it needs only A'z and A'u) is issued before the five group sums
		// and the stop decision of this one are known, so the shuffle chains and the square
		// roots overlap a mat-vec; one mat-vec per problem is wasted.  The stop test is taken
		// on squares (as in qp_tpq.cuh) and the commit is a select: no branch in the loop.
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		bool active = !bad && a.max_iter > 0; // uniform inside a group
		const float rho2 = rho * rho, relrho = a.reltol * rho, oma = 1.0f - alpha;
		// x-update (:146-159): (P + rho A'A) x = t, t = rho A'(z - u) - q, and A x.  sT holds t
		// (published, barrier passed); v = t itself (FUSE: A x = B t, x = R t waits for the end)
		// or x = R t (then A x from the published x)
		auto ax_from_t = [&](float (&v)[NR], float (&o)[MR]) {
			if constexpr (FUSE) {
				b_times(sT, o);
			} else {
				r_times(sT, v);
#pragma unroll
				for (int s = 0; s < NR; ++s)
					sX[g + T * s] = v[s];
				__syncwarp();
				b_times(sX, o);
			}
		};
		// vc = the v of the last iteration committed, vp = the v whose A x is in ax
		float vc[NR], vp[NR], ax[MR];
#pragma unroll
		for (int s = 0; s < NR; ++s) {
			vc[s] = vp[s] = fmaf(rho, w1[s] - w2[s], -q[s]);
			sT[g + T * s] = vp[s];
		}
		__syncwarp();
		ax_from_t(vp, ax);
		while (__any_sync(FULL, active)) {
			// relaxation (:161-174), clip (:176-190), dual update (:192-203) on my rows of A x
			float rr = 0.f, nax = 0.f, nz = 0.f;
			float zn[MR], un[MR];
#pragma unroll
			for (int k = 0; k < MR; ++k) {
				const float zh = RELAX ? fmaf(alpha, ax[k], oma * z[k]) : ax[k];
				zn[k] = fminf(fmaxf(zh + u[k], lo[k]), hi[k]);
				un[k] = u[k] + (zh - zn[k]);
				const float dr = ax[k] - zn[k];
				rr = fmaf(dr, dr, rr);
				nax = fmaf(ax[k], ax[k], nax);
				nz = fmaf(zn[k], zn[k], nz);
				sZ[g + T * k] = zn[k], sU[g + T * k] = un[k];
			}
			__syncwarp();
			float n1[NR], n2[NR], vn[NR];
			at_times2(sZ, sU, n1, n2); // A'z, A'u
			float ss = 0.f, nu = 0.f;
#pragma unroll
			for (int s = 0; s < NR; ++s) {
				const float d = n1[s] - w1[s];
				ss = fmaf(d, d, ss);
				nu = fmaf(n2[s], n2[s], nu);
				vn[s] = fmaf(rho, n1[s] - n2[s], -q[s]); // the next x-update's right-hand side
				sT[g + T * s] = vn[s];              // (its last readers passed the barrier above)
			}
			__syncwarp();
			float ax2[MR];
			ax_from_t(vn, ax2);
			rr = genr_group_sum<T>(rr), nax = genr_group_sum<T>(nax), nz = genr_group_sum<T>(nz);
			ss = genr_group_sum<T>(ss), nu = genr_group_sum<T>(nu);
			// stop test (:205-253, :308): r < eps_pri and s < eps_dual, on the squares
			const float eps_pri = fmaf(a.reltol, sqrt_approx(fmaxf(nax, nz)), a.eps_abs_pri);
			const float eps_dual = fmaf(relrho, sqrt_approx(nu), a.eps_abs_dual);
			const bool conv = rr < eps_pri * eps_pri && rho2 * ss < eps_dual * eps_dual;
			// commit
#pragma unroll
			for (int s = 0; s < NR; ++s) {
				vc[s] = active ? vp[s] : vc[s];
				w1[s] = active ? n1[s] : w1[s], w2[s] = active ? n2[s] : w2[s];
				vp[s] = vn[s];
			}
#pragma unroll
			for (int k = 0; k < MR; ++k) {
				z[k] = active ? zn[k] : z[k], u[k] = active ? un[k] : u[k];
				ax[k] = ax2[k];
			}
			it += active ? 1u : 0u;
			status = (active && conv) ? QP_STATUS_CONVERGED : status;
			active = active && !conv && it < a.max_iter;
		}
		// FUSE: x = R t of the last iteration done (rows of R from shared memory)
		__syncwarp();
		if constexpr (!FUSE) {
			if (it > 0) {
#pragma unroll
				for (int s = 0; s < NR; ++s)
					x[s] = vc[s];
			}
		} else {
#pragma unroll
		for (int s = 0; s < NR; ++s)
			sT[g + T * s] = vc[s];
		__syncwarp();
		if (it > 0) {
#pragma unroll
			for (int s = 0; s < NR; ++s) {
				const float4 *row = reinterpret_cast<const float4 *>(sR + (g + T * s) * LDP);
				float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
				for (int c = 0; c < NP / 4; ++c) {
					const float4 rv = row[c];
					const float4 t = *reinterpret_cast<const float4 *>(sT + 4 * c);
					acc0 = __ffma2_rn(make_float2(rv.x, rv.y), make_float2(t.x, t.y), acc0);
					acc1 = __ffma2_rn(make_float2(rv.z, rv.w), make_float2(t.z, t.w), acc1);
				}
				x[s] = (acc0.x + acc1.x) + (acc0.y + acc1.y);
			}
		}
		}

		// ---- results -------------------------------------------------------------------------
		float nf = 0.f;
#pragma unroll
		for (int s = 0; s < NR; ++s)
			nf = fmaf(x[s], 0.0f, nf);
		nf = genr_group_sum<T>(nf);
		int iters = int(it);
		if (bad)
			status = QP_STATUS_SINGULAR, iters = 0;
		else if (!(nf == 0.0f))
			status = QP_STATUS_NONFINITE;
		if (valid) {
#pragma unroll
			for (int s = 0; s < NR; ++s)
				if (g + T * s < n)
					a.x[vo + g + T * s] = x[s];
#pragma unroll
			for (int k = 0; k < MR; ++k) {
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
		__syncwarp(); // the next problem's staging overwrites what this one's lanes may still read
		if (ticket)
			next = __shfl_sync(FULL, next, 0);
		base = next;
	}
}

} // namespace qpb
