// vgb_tables.cpp -- the caller's lookup tables, built on the HOST with the C library's libm.
//
// Plain C++ (no CUDA headers) on purpose: choose_best_genotype (src/qv.cc:1804-1819) fills its tables with glibc
// pow / exp / lgamma, and confidence has to agree bit for bit, so the very same functions run here, on the box the
// job runs on.  std::pow(double, int) is pow(double, (double)int) since C++11, which is what the reference
// (compiled with -std=c++11) calls.  Built with -ffp-contract=off.
#include <cmath>

namespace vgb {

void build_call_tables(double *g, double *poisson)
{
	const double ERR_RATE = 0.01;    // src/vartype.h:13
	const double AVG_COV = 7.1;      // src/vartype.h:14
	const int MAX_COV = 63;          // src/vartype.h:27
	for (int r = 0; r <= MAX_COV; r++)
		for (int a = 0; a <= MAX_COV; a++) {
			double *c = g + (r * (MAX_COV + 1) + a) * 3;
			c[0] = std::pow(1.0 - ERR_RATE, (double)r) * std::pow(ERR_RATE, (double)a);
			c[1] = std::pow(0.5, (double)(r + a));
			c[2] = std::pow(ERR_RATE, (double)r) * std::pow(1.0 - ERR_RATE, (double)a);
		}
	const double M = std::exp(-AVG_COV);
	for (int i = 0; i <= 2 * MAX_COV; i++) poisson[i] = (M * std::pow(AVG_COV, (double)i)) / std::exp(std::lgamma(i + 1.0));
}

}  // namespace vgb
