// Host-side packing of the upper triangles of P with AVX-512 (n = 16), used by the
// QP_MEM_HOST path of qp_batch.cu when the CPU has it (checked at run time).  Kept in a
// host-only source so that the AVX-512 intrinsics never pass through nvcc's front end.
//
// The record of a problem is the row-major upper triangle (i <= j), 136 floats = 544 bytes,
// the layout qp_tpq_kernel<16, ..., PACKED> reads, written with one masked 512-bit move per
// row.  Measured on one core: 103 ns per problem against 198 ns for the 128-bit version.
// Second version: two records (1 088 bytes = 17 x 64) assembled in a cache-resident buffer
// that leave with 64-byte non-temporal stores.
#include <cstddef>
#include <cstdint>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define QPB_X86 1
#endif

namespace qpb {

bool hostpack_have_avx512()
{
#ifdef QPB_X86
	static const bool have = __builtin_cpu_supports("avx512f");
	return have;
#else
	return false;
#endif
}

#ifdef QPB_X86
__attribute__((target("avx512f"))) static inline void record16(const float *src, float *r)
{
	int k = 0;
#pragma GCC unroll 16
	for (int i = 0; i < 16; ++i) {
		const __mmask16 m = __mmask16(0xffffu >> i); // columns i..15 of row i
		_mm512_mask_storeu_ps(r + k, m, _mm512_maskz_loadu_ps(m, src + i * 16 + i));
		k += 16 - i;
	}
}

// The version the host path uses: one masked 512-bit move per row straight into the staging
// buffer with ordinary stores (part of a chunk is still in the last-level cache when the copy
// engine reads it).  End to end 64-67 M QP/s against 61-62 M with the non-temporal version
// below (n = 16, 2^20 problems, 16-vCPU host).
__attribute__((target("avx512f"))) void hostpack_upper16_avx512_cached(const float *P, float *out, size_t count)
{
	for (size_t c = 0; c < count; ++c)
		record16(P + c * 256, out + c * 136);
}

// `out` must be 16-byte aligned (it is: records are 544 bytes apart in a pinned buffer)
__attribute__((target("avx512f"))) void hostpack_upper16_avx512(const float *P, float *out, size_t count)
{
	constexpr int PK = 136;
	alignas(64) float rec[2 * PK + 16];
	size_t c = 0;
	// a record is 8.5 cache lines: an odd start leaves with 16-byte stores, then pairs
	if (count && (reinterpret_cast<uintptr_t>(out) & 63u)) {
		record16(P, rec);
		for (int v = 0; v < PK; v += 4)
			_mm_stream_ps(out + v, _mm_load_ps(rec + v));
		c = 1;
	}
	if ((reinterpret_cast<uintptr_t>(out + c * PK) & 63u) == 0) {
		for (; c + 1 < count; c += 2) {
			record16(P + c * 256, rec);
			record16(P + (c + 1) * 256, rec + PK);
			float *dst = out + c * PK;
#pragma GCC unroll 17
			for (int v = 0; v < 2 * PK; v += 16)
				_mm512_stream_ps(dst + v, _mm512_load_ps(rec + v));
		}
	}
	for (; c < count; ++c) {
		record16(P + c * 256, rec);
		float *dst = out + c * PK;
		for (int v = 0; v < PK; v += 4)
			_mm_stream_ps(dst + v, _mm_load_ps(rec + v));
	}
	_mm_sfence();
}
#else
void hostpack_upper16_avx512(const float *, float *, size_t) {}
void hostpack_upper16_avx512_cached(const float *, float *, size_t) {}
#endif

} // namespace qpb
