// qp_file.cpp -- the reference's problem file, host side of libqpb200.so (no CUDA here).
//
// Format (the reference writes one record per file):
//     [double n] [n*n doubles: P row-major] [n doubles: q]        native endianness
// written by test_reference() (/root/reference/test/test.c:108-126) and parsed by
// read_qp() (/root/reference/test/qp_ref.py:8-30).  This library also accepts several
// records back to back (all of one dimension), so that a batch can travel in one file;
// a one-record file is byte-identical to the reference's.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/qp_batch.h"

extern "C" int qp_problem_file_write(const char *path, uint32_t n, size_t count, const double *P,
				     const double *q, int append)
{
	if (!path || n == 0 || n > 65535u || (count && (!P || !q)))
		return QP_ERR_INVALID;
	FILE *fp = std::fopen(path, append ? "ab" : "wb");
	if (!fp)
		return QP_ERR_IO;
	const double dn = double(n); // the reference stores the dimension as a double (test.c:112-114)
	bool ok = true;
	for (size_t c = 0; c < count && ok; ++c) {
		ok = std::fwrite(&dn, sizeof dn, 1, fp) == 1 &&
		     std::fwrite(P + c * size_t(n) * n, sizeof(double), size_t(n) * n, fp) == size_t(n) * n &&
		     std::fwrite(q + c * size_t(n), sizeof(double), n, fp) == n;
	}
	ok = (std::fclose(fp) == 0) && ok;
	return ok ? QP_OK : QP_ERR_IO;
}

extern "C" long long qp_problem_file_count(const char *path, uint32_t *n_out)
{
	if (!path)
		return QP_ERR_INVALID;
	struct stat st;
	if (stat(path, &st) != 0)
		return QP_ERR_IO;
	FILE *fp = std::fopen(path, "rb");
	if (!fp)
		return QP_ERR_IO;
	double dn = 0.0;
	const bool got = std::fread(&dn, sizeof dn, 1, fp) == 1;
	std::fclose(fp);
	if (!got || !(dn >= 1.0) || dn > 65535.0 || dn != std::floor(dn))
		return QP_ERR_INVALID;
	const unsigned long long n = (unsigned long long)dn;
	const unsigned long long rec = (1 + n * n + n) * sizeof(double);
	if ((unsigned long long)st.st_size % rec != 0)
		return QP_ERR_INVALID;
	if (n_out)
		*n_out = uint32_t(n);
	return (long long)((unsigned long long)st.st_size / rec);
}

extern "C" int qp_problem_file_read(const char *path, size_t first, size_t count, uint32_t n,
				    double *P, double *q)
{
	if (!path || n == 0 || (count && (!P || !q)))
		return QP_ERR_INVALID;
	uint32_t fn = 0;
	const long long total = qp_problem_file_count(path, &fn);
	if (total < 0)
		return int(total);
	if (fn != n || first + count > size_t(total))
		return QP_ERR_INVALID;
	FILE *fp = std::fopen(path, "rb");
	if (!fp)
		return QP_ERR_IO;
	const size_t rec = (1 + size_t(n) * n + n) * sizeof(double);
	bool ok = std::fseek(fp, long(first * rec), SEEK_SET) == 0;
	for (size_t c = 0; c < count && ok; ++c) {
		double dn = 0.0;
		ok = std::fread(&dn, sizeof dn, 1, fp) == 1 && dn == double(n) &&
		     std::fread(P + c * size_t(n) * n, sizeof(double), size_t(n) * n, fp) == size_t(n) * n &&
		     std::fread(q + c * size_t(n), sizeof(double), n, fp) == n;
	}
	std::fclose(fp);
	return ok ? QP_OK : QP_ERR_IO;
}
