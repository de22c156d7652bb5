// Host twin of csrc/kem_kernel.cuh for tests/test_generated_code_on_host.py (CPU only).
//
// A generated model translation unit (`codegen/emit.py`) includes "kem_kernel.cuh"; with this
// directory first on the include path g++ compiles the SAME emitted text -- hoist(), deriv(),
// outputs(), tonly(), tonly_dev(), the constant table -- against the host build of kem_math.cuh
// and a scalar restatement of the step kernel's prologue / RK4 loop / epilogue.  So the
// generator's rewrites (hoisting, shared exponentials, a*rcp(b), relaxed gates, literal tables)
// are checked against the oracle without a GPU.  Test infrastructure: nothing in the product
// includes this file.
#pragma once
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#include <math.h>
#include <vector>

#include "kem_math.cuh"

static inline double kem_npmod_host(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}
static inline double kem_npmod(double a, double b) { return kem_npmod_host(a, b); }

// One PDE step of scheme O1 over AoS tables (the oracle's layout), in place.  Stage times and
// update association as kem_step_kernel / build_ttab: ta = t0 + j h, tb = t0 + (j + 1/2) h,
// tc = t0 + (j + 1) h, epilogue at t0 + dt; acc = ((k1 + 2 k2) + 2 k3) + k4; y += (h/6) acc.
// `device_tonly` evaluates the time-only factors with the emitted device code (kem::exp)
// instead of the host code (libm) -- what scheme O3 does on the GPU.
template <class M>
static int twin_step(long n, double *S, double *P, double t0, double dt, int n_sub, int device_tonly,
                     const int *out_cols)
{
    constexpr int NS = M::NS, NP = M::NP, NOUT = M::NOUT, NT = M::NT;
    std::vector<double> tab((size_t)(2 * n_sub + 2) * (NT > 0 ? NT : 1), 0.0);
    const double h = dt / (double)n_sub;
    for (int k = 0; k <= 2 * n_sub + 1 && NT > 0; ++k) {
        const double t = k == 2 * n_sub + 1 ? t0 + dt
                       : (k & 1)            ? t0 + ((double)(k / 2) + 0.5) * h
                                            : t0 + (double)(k / 2) * h;
        if (device_tonly) M::tonly_dev(t, &tab[(size_t)k * NT]);
        else M::tonly(t, &tab[(size_t)k * NT]);
    }
    const double hh = 0.5 * h, h6 = h / 6.0;
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        double p[NP > 0 ? NP : 1];
        for (int c = 0; c < NP; ++c) p[c] = M::used(c) ? P[i * NP + c] : 0.0;
        typename M::H q;
        M::hoist(p, q);
        double y[NS], w[NS];
        for (int c = 0; c < NS; ++c) w[c] = y[c] = S[i * NS + c];
        for (int j = 0; j < n_sub; ++j) {
            const double *tj = tab.data() + (size_t)(2 * j) * NT;
            double acc[NS];
            for (int c = 0; c < NS; ++c) acc[c] = 0.0;
            for (int s = 0; s < 4; ++s) {
                double k[NS];
                M::deriv(w, k, q, tj + ((s + 1) >> 1) * NT);
                const double bw = (s == 0 || s == 3) ? 1.0 : 2.0;
                const double aw = (s == 2) ? h : hh;
                for (int c = 0; c < NS; ++c) {
                    acc[c] = acc[c] + bw * k[c];
                    w[c] = y[c] + aw * k[c];
                }
            }
            for (int c = 0; c < NS; ++c) {
                y[c] = y[c] + h6 * acc[c];
                w[c] = y[c];
            }
        }
        if (NOUT > 0) {
            double o[NOUT > 0 ? NOUT : 1];
            M::outputs(y, o, q, tab.data() + (size_t)(2 * n_sub + 1) * NT);
            for (int c = 0; c < NOUT; ++c) P[i * NP + out_cols[c]] = o[c];
        }
        bool finite = true;
        for (int c = 0; c < NS; ++c) {
            S[i * NS + c] = y[c];
            finite = finite && isfinite(y[c]);
        }
        bad += !finite;
    }
    return bad;
}

#define KEM_DEFINE_MODEL(M, NAME_STR, HASH_STR, OUT_COLS, USED_COLS, N_USED, N_CONST, CONST_COLS, CONST_VALS) \
    extern "C" {                                                                                        \
    void twin_dims(int *d) { d[0] = M::NS; d[1] = M::NP; d[2] = M::NOUT; d[3] = M::NT; d[4] = N_USED; d[5] = N_CONST; } \
    const char *twin_name(void) { return NAME_STR; }                                                    \
    int twin_step_rk4(long n, double *S, double *P, double t0, double dt, int n_sub, int device_tonly)  \
    {                                                                                                   \
        return twin_step<M>(n, S, P, t0, dt, n_sub, device_tonly, OUT_COLS);                            \
    }                                                                                                   \
    const int *twin_const_cols(void) { return CONST_COLS; }                                             \
    const double *twin_const_vals(void) { return CONST_VALS; }                                          \
    const int *twin_used_cols(void) { return USED_COLS; }                                               \
    }
