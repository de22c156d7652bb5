// Host build of csrc/kem_math.cuh for accuracy tests (tests/test_kem_math.py).
// Reference values: long-double libm (64-bit mantissa), error reported in ulps
// of the double result.
#include "kem_math.cuh"

#include <math.h>
#include <stdint.h>

static double ulp_err(double got, long double want)
{
    if (isnan(got) || isinf(got)) return isfinite((double)want) ? 1e30 : 0.0;
    const double w = (double)want;
    int e;
    frexp(w, &e);
    const long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - want) / ulp);
}

static uint64_t next(uint64_t &s)
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return s;
}

static double uniform(uint64_t &s, double lo, double hi)
{
    return lo + (hi - lo) * ((next(s) >> 11) * (1.0 / 9007199254740992.0));
}

extern "C" {

double kem_host_exp(double x) { return kem::exp(x); }
double kem_host_div(double a, double b) { return kem::div(a, b); }
double kem_host_rcp(double b) { return kem::rcp(b); }

// max ulp error of kem::exp on n uniform samples of [lo, hi]
double kem_check_exp(long n, double lo, double hi, uint64_t seed, double *mean_out)
{
    uint64_t s = seed ? seed : 88172645463325252ull;
    double worst = 0.0, sum = 0.0;
    for (long i = 0; i < n; ++i) {
        const double x = uniform(s, lo, hi);
        const double e = ulp_err(kem::exp(x), expl((long double)x));
        worst = e > worst ? e : worst;
        sum += e;
    }
    if (mean_out) *mean_out = sum / (double)n;
    return worst;
}

// max ulp error of kem::div(a,b); a, b log-uniform in magnitude 2^[-emax, emax], random signs
double kem_check_div(long n, int emax, uint64_t seed, double *rcp_worst_out)
{
    uint64_t s = seed ? seed : 88172645463325252ull;
    double worst = 0.0, rworst = 0.0;
    for (long i = 0; i < n; ++i) {
        double a = ldexp(uniform(s, 1.0, 2.0), (int)(next(s) % (2 * emax + 1)) - emax);
        double b = ldexp(uniform(s, 1.0, 2.0), (int)(next(s) % (2 * emax + 1)) - emax);
        if (next(s) & 1) a = -a;
        if (next(s) & 1) b = -b;
        const double e = ulp_err(kem::div(a, b), (long double)a / (long double)b);
        const double er = ulp_err(kem::rcp(b), 1.0L / (long double)b);
        worst = e > worst ? e : worst;
        rworst = er > rworst ? er : rworst;
    }
    if (rcp_worst_out) *rcp_worst_out = rworst;
    return worst;
}

double kem_host_log(double x) { return kem::log(x); }
double kem_host_sqrt(double x) { return kem::sqrt(x); }
double kem_host_pow15(double x) { return kem::pow15(x); }

// which: 0 log, 1 sqrt, 2 pow15; x log-uniform in 2^[-emax, emax] (log: also a dense sweep near 1)
double kem_check_unary(int which, long n, int emax, uint64_t seed)
{
    uint64_t s = seed ? seed : 88172645463325252ull;
    double worst = 0.0;
    for (long i = 0; i < n; ++i) {
        double x = ldexp(uniform(s, 1.0, 2.0), (int)(next(s) % (2 * emax + 1)) - emax);
        if (which == 0 && (i & 3) == 0) x = uniform(s, 0.5, 2.0);
        double got;
        long double want;
        if (which == 0) { got = kem::log(x); want = logl((long double)x); }
        else if (which == 1) { got = kem::sqrt(x); want = sqrtl((long double)x); }
        else { got = kem::pow15(x); want = powl((long double)x, 1.5L); }
        const double e = ulp_err(got, want);
        worst = e > worst ? e : worst;
    }
    return worst;
}

}  // extern "C"
