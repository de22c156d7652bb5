/*
 * knpemi_oracle.c -- CPU ORACLE for the membrane-ODE stage.  TEST INFRASTRUCTURE.
 *
 * This file is a plain-C restatement of the arithmetic on the reference's hot
 * path.  It is NOT part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build or call it.
 * The product (libknpemi_b200.so) never links or loads it and has no CPU
 * fallback.
 *
 * What it restates
 *   - the six model right-hand sides (`rhs_numba`, C signature
 *     void(double t, double* y, double* dy, double* p), p in/out):
 *       hh_ideal      examples/idealized_geometries/mm_hh.py:139-227
 *       hh_tissue     examples/local_astrocyte_depolarization/mm_hh.py:130-201
 *       glial_tissue  examples/local_astrocyte_depolarization/mm_glial.py:133-205
 *       glial_bench   examples/benchmark/mm_glial.py:120-204
 *       calibration   examples/calibrate_initial_conditions/mm_calibration.py:151-298
 *       hh_test       tests/mm_test_ode.py:126-169
 *     Every expression keeps the reference's association (Python precedence,
 *     left to right) and numba's lowering of integer powers (x**2 = x*x,
 *     pow(x,3) = x*(x*x), pow(x,4) = (x*x)*(x*x)); transcendental calls go to
 *     glibc libm exactly as the numba cfunc's do.
 *   - the per-row stepping loop of MembraneModel.step_lsoda
 *     (src/knpemi/odeSolver.py:106-123) with the integrator replaced by the
 *     normative fixed-step scheme O1 of SURVEY.md section 8(c): classical RK4,
 *     n_sub sub-steps, then one more RHS call at (t0+dt, y_end) so the I_ch
 *     parameter slots hold I_ch(y(t0+dt)).
 *
 * Pinning: tests/golden/make_golden.py (run in the build container, where
 * /root/reference is mounted) checks every rhs_* below BIT-FOR-BIT against the
 * reference's own compiled `rhs_numba` cfuncs and drives those cfuncs through
 * kemo_step_fn() to produce the committed trajectories in tests/golden/.
 * The integrator itself (numbalsoda LSODA, un-vendored and un-pinned in the
 * reference: pyproject.toml:14) has no golden vectors anywhere in the
 * reference -> the *scheme* is "parity unpinned"; see DESIGN.md.
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*kemo_rhs_fn)(double t, double *y, double *dy, double *p);

/* numba lowers integer powers by repeated squaring (numba/cpython/numbers.py) */
static inline double pw2(double x) { return x * x; }
static inline double pw3(double x) { double x2 = x * x; return x * x2; }
static inline double pw4(double x) { double x2 = x * x; return x2 * x2; }

/* numpy.mod on doubles: fmod, then sign fix-up towards the divisor */
static inline double npmod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

/* ------------------------------------------------------------------ hh_ideal
 * examples/idealized_geometries/mm_hh.py:139-227 (SI units) */
static void rhs_hh_ideal(double t, double *y, double *dy, double *p)
{
    const double g_Na_bar = p[0], g_K_bar = p[1], g_leak_Na = p[2], g_leak_K = p[3];
    const double m_K = p[4], m_Na = p[5], I_max = p[6], Cm = p[7], stim = p[8];
    const double K_e = p[9], K_i = p[10], Na_e = p[11], Na_i = p[12];
    const double z_K = p[19], psi = p[21];

    /* :169-170   1/psi * 1/z_K * log(.)  ==  (((1/psi)*1)/z_K)*log(.) */
    const double E_Na = ((1.0 / psi) * 1.0) / z_K * log(Na_e / Na_i);
    const double E_K = ((1.0 / psi) * 1.0) / z_K * log(K_e / K_i);

    /* :192-205 */
    const double a25 = 25. - 1.0e3 * (y[3] + 65.0e-3);
    const double alpha_m = 0.1e3 * a25 / (exp(a25 / 10.) - 1.0);
    const double beta_m = 4.e3 * exp(-1.0e3 * (y[3] + 65.0e-3) / 18.);
    dy[0] = (1.0 - y[0]) * alpha_m - y[0] * beta_m;

    const double alpha_h = 0.07e3 * exp(-1.0e3 * (y[3] + 65.0e-3) / 20.);
    const double beta_h = 1.e3 / (exp((30. - 1.0e3 * (y[3] + 65.0e-3)) / 10.) + 1.0);
    dy[1] = (1.0 - y[1]) * alpha_h - y[1] * beta_h;

    const double a10 = 10. - 1.0e3 * (y[3] + 65.0e-3);
    const double alpha_n = 0.01e3 * a10 / (exp(a10 / 10.) - 1.);
    const double beta_n = 0.125e3 * exp(-1.0e3 * (y[3] + 65.0e-3) / 80.);
    dy[2] = (1.0 - y[2]) * alpha_n - y[2] * beta_n;

    /* :208-210 */
    const double i_Stim = stim * exp(-npmod(t, 0.03) / 0.002) * (double)(t < 125e-3);
    const double i_pump = I_max / (pw2(1.0 + m_K / K_e) * pw3(1.0 + m_Na / Na_i));

    /* :213-218 */
    const double i_Na = (g_leak_Na + g_Na_bar * y[1] * pw3(y[0]) + i_Stim) * (y[3] - E_Na)
                        + 3.0 * i_pump;
    const double i_K = (g_leak_K + g_K_bar * pw4(y[2])) * (y[3] - E_K) - 2.0 * i_pump;

    p[15] = i_Na;   /* :221 */
    p[16] = i_K;    /* :223 */
    p[17] = 0.0;    /* :225 */
    dy[3] = (-i_K - i_Na) / Cm;   /* :227 */
}

/* ----------------------------------------------------------------- hh_tissue
 * examples/local_astrocyte_depolarization/mm_hh.py:130-201 (mV, ms) */
static void rhs_hh_tissue(double t, double *y, double *dy, double *p)
{
    const double g_Na_bar = p[0], g_K_bar = p[1], g_leak_Na = p[2], g_leak_K = p[3];
    const double m_K = p[4], m_Na = p[5], I_max = p[6], Cm = p[7], stim = p[8];
    const double K_e = p[9], K_i = p[10], Na_e = p[11], Na_i = p[12];
    const double z_K = p[19], psi = p[21];

    const double E_Na = ((1.0 / psi) * 1.0) / z_K * log(Na_e / Na_i);   /* :160 */
    const double E_K = ((1.0 / psi) * 1.0) / z_K * log(K_e / K_i);      /* :161 */

    /* :163-170 */
    const double alpha_m = 0.1 * (y[3] + 40.0) / (1.0 - exp(-(y[3] + 40.0) / 10.0));
    const double beta_m = 4.0 * exp(-(y[3] + 65.0) / 18.0);
    const double alpha_h = 0.07 * exp(-(y[3] + 65.0) / 20.0);
    const double beta_h = 1.0 / (1.0 + exp(-(y[3] + 35.0) / 10.0));
    const double alpha_n = 0.01 * (y[3] + 55.0) / (1.0 - exp(-(y[3] + 55.0) / 10.0));
    const double beta_n = 0.125 * exp(-(y[3] + 65.0) / 80.0);

    dy[0] = (1.0 - y[0]) * alpha_m - y[0] * beta_m;   /* :173 */
    dy[1] = (1.0 - y[1]) * alpha_h - y[1] * beta_h;   /* :176 */
    dy[2] = (1.0 - y[2]) * alpha_n - y[2] * beta_n;   /* :179 */

    const double i_Stim = stim * exp(-npmod(t, 30.0) / 2.0) * (double)(t < 125.0);  /* :182 */
    const double i_pump = I_max / (pw2(1.0 + m_K / K_e) * pw3(1.0 + m_Na / Na_i));  /* :184 */

    const double i_Na = (g_leak_Na + g_Na_bar * y[1] * pw3(y[0]) + i_Stim) * (y[3] - E_Na)
                        + 3.0 * i_pump;                                             /* :187 */
    const double i_K = (g_leak_K + g_K_bar * pw4(y[2])) * (y[3] - E_K) - 2.0 * i_pump; /* :191 */

    p[15] = i_Na;
    p[16] = i_K;
    p[17] = 0.0;
    dy[3] = (-i_K - i_Na) / Cm;   /* :201 */
}

/* -------------------------------------------------------------- glial_tissue
 * examples/local_astrocyte_depolarization/mm_glial.py:133-205 */
static void rhs_glial_tissue(double t, double *y, double *dy, double *p)
{
    (void)t;
    const double g_leak_Cl = p[0], g_leak_Na = p[1], g_leak_K = p[2], Cm = p[3];
    const double m_K = p[8], m_Na = p[9], I_max = p[10];
    const double K_e_init = p[11], K_i_init = p[12];
    const double K_e = p[13], K_i = p[14], Na_e = p[15], Na_i = p[16];
    const double Cl_e = p[17], Cl_i = p[18];
    const double z_K = p[20], z_Cl = p[21], psi = p[22];

    const double E_Na = ((1.0 / psi) * 1.0) / z_K * log(Na_e / Na_i);   /* :163 */
    const double E_K = ((1.0 / psi) * 1.0) / z_K * log(K_e / K_i);      /* :164 */
    const double E_Cl = ((1.0 / psi) * 1.0) / z_Cl * log(Cl_e / Cl_i);  /* :165 */

    const double temperature = 307e3, R = 8.315e3, F = 96500e3;         /* :168-170 */

    /* :172-174 */
    const double i_pump = I_max * (K_e / (K_e + m_K))
                          * (pow(Na_i, 1.5) / (pow(Na_i, 1.5) + pow(m_Na, 1.5)));

    /* :177-183 */
    const double E_K_init = R * temperature / F * log(K_e_init / K_i_init);
    const double dphi = y[0] - E_K;
    const double A = 1.0 + exp(18.5 / 42.4);
    const double B = 1.0 + exp(-(118.6 + E_K_init) / 44.1);
    const double C = 1.0 + exp((dphi + 18.5) / 42.4);
    const double D = 1.0 + exp(-(118.6 + y[0]) / 44.1);
    const double g_Kir = sqrt(K_e / K_e_init) * (A * B) / (C * D);

    const double i_Kir = g_leak_K * g_Kir * (y[0] - E_K);          /* :186 */
    const double i_Na = g_leak_Na * (y[0] - E_Na) + 3.0 * i_pump;  /* :189 */
    const double i_K = i_Kir - 2.0 * i_pump;                       /* :192 */
    const double i_Cl = g_leak_Cl * (y[0] - E_Cl);                 /* :195 */

    p[5] = i_Na;
    p[6] = i_K;
    p[7] = i_Cl;
    dy[0] = (-i_K - i_Na - i_Cl) / Cm;   /* :205 */
}

/* --------------------------------------------------------------- glial_bench
 * examples/benchmark/mm_glial.py:120-204 */
static void rhs_glial_bench(double t, double *y, double *dy, double *p)
{
    (void)t;
    const double psi = p[0], g_leak_Cl = p[1], g_leak_Na = p[2], g_leak_K = p[3];
    const double z_K = p[5], z_Cl = p[6], Cm = p[7];
    const double K_e = p[12], K_i = p[13], Na_e = p[14], Na_i = p[15];
    const double Cl_e = p[16], Cl_i = p[17];
    const double m_K = p[18], m_Na = p[19], I_max = p[20];

    const double E_Na = ((1.0 / psi) * 1.0) / z_K * log(Na_e / Na_i);   /* :161 */
    const double E_K = ((1.0 / psi) * 1.0) / z_K * log(K_e / K_i);      /* :162 */
    const double E_Cl = ((1.0 / psi) * 1.0) / z_Cl * log(Cl_e / Cl_i);  /* :163 */

    const double K_e_init = 3.092970607490389;   /* :165 */
    const double K_i_init = 99.3100014897692;    /* :166 */

    /* :168-170 */
    const double i_pump = I_max * (K_e / (K_e + m_K))
                          * (pow(Na_i, 1.5) / (pow(Na_i, 1.5) + pow(m_Na, 1.5)));

    /* :173-180 */
    const double E_K_init = 1.0 / psi * log(K_e_init / K_i_init);
    const double dphi = y[0] - E_K;
    const double A = 1.0 + exp(18.4 / 42.4);
    const double B = 1.0 + exp(-(0.1186e3 + E_K_init) / 0.0441e3);
    const double C = 1.0 + exp((dphi + 0.0185e3) / 0.0425e3);
    const double D = 1.0 + exp(-(0.1186e3 + y[0]) / 0.0441e3);
    const double g_Kir = sqrt(K_e / K_e_init) * (A * B) / (C * D);

    const double i_Kir = g_leak_K * g_Kir * (y[0] - E_K);          /* :183 */
    const double i_Na = g_leak_Na * (y[0] - E_Na) + 3.0 * i_pump;  /* :186 */
    const double i_K = i_Kir - 2.0 * i_pump;                       /* :189 */
    const double i_Cl = g_leak_Cl * (y[0] - E_Cl);                 /* :192 */

    p[9] = i_Na;
    p[10] = i_K;
    p[11] = i_Cl;
    dy[0] = (-i_K - i_Na - i_Cl) / Cm;   /* :204 */
}

/* --------------------------------------------------------------- calibration
 * examples/calibrate_initial_conditions/mm_calibration.py:151-298 */
static void rhs_calibration(double t, double *y, double *dy, double *p)
{
    const double temperature = 307e3, R = 8.315e3, F = 96500e3;   /* :159-161 */
    const double ICS_vol = 3.42e-11 / 2.0;                        /* :163 */
    const double ECS_vol = 7.08e-11;                              /* :164 */
    const double surface = 2.29e-6;                               /* :165 */
    const double K_e_init = 3.092970607490389;                    /* :167 */
    const double K_g_init = 99.3100014897692;                     /* :168 */

    const double K_e = y[5], K_n = y[6], K_g = y[7];
    const double Na_e = y[8], Na_n = y[9], Na_g = y[10];
    const double Cl_e = y[11], Cl_g = y[13];

    const double g_Na_bar = p[0], g_K_bar = p[1];
    const double g_leak_Na_n = p[2], g_leak_K_n = p[3];
    const double g_leak_Na_g = p[4], g_leak_K_g = p[5];
    const double Cm = p[6], stim = p[7], m_K = p[8], m_Na = p[9];
    const double I_max_n = p[10], I_max_g = p[11], g_leak_Cl_g = p[12];

    /* :196-203 (E_Cl_n is computed by the reference but never used) */
    const double E_Na_n = R * temperature / F * log(Na_e / Na_n);
    const double E_K_n = R * temperature / F * log(K_e / K_n);
    const double E_Na_g = R * temperature / F * log(Na_e / Na_g);
    const double E_K_g = R * temperature / F * log(K_e / K_g);
    const double E_Cl_g = -R * temperature / F * log(Cl_e / Cl_g);
    const double E_K_init = R * temperature / F * log(K_e_init / K_g_init);

    /* :205-212 */
    const double alpha_m = 0.1 * (y[3] + 40.0) / (1.0 - exp(-(y[3] + 40.0) / 10.0));
    const double beta_m = 4.0 * exp(-(y[3] + 65.0) / 18.0);
    const double alpha_h = 0.07 * exp(-(y[3] + 65.0) / 20.0);
    const double beta_h = 1.0 / (1.0 + exp(-(y[3] + 35.0) / 10.0));
    const double alpha_n = 0.01 * (y[3] + 55.0) / (1.0 - exp(-(y[3] + 55.0) / 10.0));
    const double beta_n = 0.125 * exp(-(y[3] + 65.0) / 80.0);

    dy[0] = (1.0 - y[0]) * alpha_m - y[0] * beta_m;   /* :215 */
    dy[1] = (1.0 - y[1]) * alpha_h - y[1] * beta_h;   /* :218 */
    dy[2] = (1.0 - y[2]) * alpha_n - y[2] * beta_n;   /* :221 */

    const double i_Stim = stim * exp(-npmod(t, 20.0) / 2.0);   /* :224 */

    /* :226-231 */
    const double i_pump_n = I_max_n / (pw2(1.0 + m_K / K_e) * pw3(1.0 + m_Na / Na_n));
    const double i_pump_g = I_max_g * (K_e / (K_e + m_K))
                            * (pow(Na_g, 1.5) / (pow(Na_g, 1.5) + pow(m_Na, 1.5)));

    /* :234-242 */
    const double dphi = y[4] - E_K_g;
    const double A = 1.0 + exp(18.4 / 42.4);
    const double B = 1.0 + exp(-(0.1186e3 + E_K_init) / 0.0441e3);
    const double C = 1.0 + exp((dphi + 0.0185e3) / 0.0425e3);
    const double D = 1.0 + exp(-(0.1186e3 + y[4]) / 0.0441e3);
    const double g_Kir = sqrt(K_e / K_e_init) * (A * B) / (C * D);
    const double I_Kir = g_leak_K_g * g_Kir * (y[4] - E_K_g);

    /* :245-262 */
    const double i_Na_n = (g_leak_Na_n + g_Na_bar * y[1] * pw3(y[0]) + i_Stim) * (y[3] - E_Na_n)
                          + 3.0 * i_pump_n;
    const double i_K_n = (g_leak_K_n + g_K_bar * pw4(y[2])) * (y[3] - E_K_n) - 2.0 * i_pump_n;
    const double i_Na_g = g_leak_Na_g * (y[4] - E_Na_g) + 3.0 * i_pump_g;
    const double i_K_g = I_Kir - 2.0 * i_pump_g;
    const double i_Cl_g = g_leak_Cl_g * (y[4] - E_Cl_g);
    const double i_Cl_n = 0.0;

    dy[3] = (-i_K_n - i_Na_n - i_Cl_n) / Cm;   /* :265 */
    dy[4] = (-i_K_g - i_Na_g - i_Cl_g) / Cm;   /* :268 */

    /* :271-298 */
    dy[5] = i_K_n * surface / (F * ECS_vol) + i_K_g * surface / (F * ECS_vol);
    dy[6] = -i_K_n * surface / (F * ICS_vol);
    dy[7] = -i_K_g * surface / (F * ICS_vol);
    dy[8] = i_Na_n * surface / (F * ECS_vol) + i_Na_g * surface / (F * ECS_vol);
    dy[9] = -i_Na_n * surface / (F * ICS_vol);
    dy[10] = -i_Na_g * surface / (F * ICS_vol);
    dy[11] = -i_Cl_n * surface / (F * ECS_vol) - i_Cl_g * surface / (F * ECS_vol);
    dy[12] = i_Cl_n * surface / (F * ICS_vol);
    dy[13] = i_Cl_g * surface / (F * ICS_vol);
}

/* ------------------------------------------------------------------- hh_test
 * tests/mm_test_ode.py:126-169 */
static void rhs_hh_test(double t, double *y, double *dy, double *p)
{
    /* :133-145 */
    const double a25 = 25. - 1.0 * (y[3] + 65.0);
    const double alpha_m = 0.1 * a25 / (exp(a25 / 10.) - 1.0);
    const double beta_m = 4. * exp(-1.0 * (y[3] + 65.0) / 18.);
    dy[0] = (1.0 - y[0]) * alpha_m - y[0] * beta_m;

    const double alpha_h = 0.07 * exp(-1.0 * (y[3] + 65.0) / 20.);
    const double beta_h = 1. / (exp((30. - 1.0 * (y[3] + 65.0)) / 10.) + 1.0);
    dy[1] = (1.0 - y[1]) * alpha_h - y[1] * beta_h;

    const double a10 = 10. - 1.0 * (y[3] + 65.0);
    const double alpha_n = 0.01 * a10 / (exp(a10 / 10.) - 1.);
    const double beta_n = 0.125 * exp(-1.0 * (y[3] + 65.0) / 80.);
    dy[2] = (1.0 - y[2]) * alpha_n - y[2] * beta_n;

    const double i_Stim = p[7] * exp(-npmod(t, 0.03) / 0.002) * (double)(t < 125.0);   /* :148 */
    const double i_pump = p[15] / (pw2(1.0 + p[13] / p[11]) * pw3(1.0 + p[14] / p[12])); /* :150 */

    const double i_Na = (p[2] + p[0] * y[1] * pw3(y[0]) + i_Stim) * (y[3] - p[4])
                        + 3.0 * i_pump;                                               /* :154 */
    const double i_K = (p[3] + p[1] * pw4(y[2])) * (y[3] - p[5]) - 2.0 * i_pump;      /* :158 */

    p[8] = i_Na;
    p[9] = i_K;
    p[10] = 0.0;
    dy[3] = (-i_K - i_Na) / p[6];   /* :169 */
}

/* ------------------------------------------------------------------ registry */
typedef struct {
    const char *name;
    int ns, np;
    kemo_rhs_fn rhs;
} kemo_model;

static const kemo_model MODELS[] = {
    {"hh_ideal", 4, 22, rhs_hh_ideal},
    {"hh_tissue", 4, 22, rhs_hh_tissue},
    {"glial_tissue", 1, 23, rhs_glial_tissue},
    {"glial_bench", 1, 21, rhs_glial_bench},
    {"calibration", 14, 13, rhs_calibration},
    {"hh_test", 4, 17, rhs_hh_test},
};
#define N_MODELS ((int)(sizeof(MODELS) / sizeof(MODELS[0])))
#define MAX_NS 64

int kemo_model_count(void) { return N_MODELS; }

int kemo_model_find(const char *name)
{
    for (int i = 0; i < N_MODELS; ++i)
        if (strcmp(name, MODELS[i].name) == 0) return i;
    return -1;
}

int kemo_model_dims(int model, int *ns, int *np)
{
    if (model < 0 || model >= N_MODELS) return -1;
    *ns = MODELS[model].ns;
    *np = MODELS[model].np;
    return 0;
}

/* one RHS evaluation (pointwise pinning against the reference cfuncs) */
int kemo_rhs(int model, double t, double *y, double *dy, double *p)
{
    if (model < 0 || model >= N_MODELS) return -1;
    MODELS[model].rhs(t, y, dy, p);
    return 0;
}

/*
 * Stage times of the normative scheme, all formed from
 * (t0, dt, n_sub) in this exact way so oracle and product agree bitwise on
 * which side of a stimulus discontinuity every evaluation falls:
 *     h    = dt / n_sub
 *     ta_j = t0 + j*h            (k1)
 *     tb_j = t0 + (j + 0.5)*h    (k2, k3)
 *     tc_j = t0 + (j + 1)*h      (k4)      [== ta_{j+1}]
 *     tend = t0 + dt             (current epilogue)
 */
static void step_row(kemo_rhs_fn rhs, int ns, double *y, double *p,
                     double t0, double dt, int n_sub)
{
    double k1[MAX_NS], k2[MAX_NS], k3[MAX_NS], k4[MAX_NS], w[MAX_NS];
    const double h = dt / (double)n_sub;
    const double hh = 0.5 * h;
    const double h6 = h / 6.0;
    for (int j = 0; j < n_sub; ++j) {
        const double ta = t0 + (double)j * h;
        const double tb = t0 + ((double)j + 0.5) * h;
        const double tc = t0 + ((double)j + 1.0) * h;
        rhs(ta, y, k1, p);
        for (int i = 0; i < ns; ++i) w[i] = y[i] + hh * k1[i];
        rhs(tb, w, k2, p);
        for (int i = 0; i < ns; ++i) w[i] = y[i] + hh * k2[i];
        rhs(tb, w, k3, p);
        for (int i = 0; i < ns; ++i) w[i] = y[i] + h * k3[i];
        rhs(tc, w, k4, p);
        for (int i = 0; i < ns; ++i)
            y[i] = y[i] + h6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
    }
    /* current epilogue: leaves I_ch(y(t0+dt)) in the output parameter slots */
    rhs(t0 + dt, y, k1, p);
}

/*
 * Advance every row of the AoS tables states[n,ns], params[n,np] by one PDE
 * step (restates the row loop odeSolver.py:107-122 with scheme O1).
 * Returns the number of rows whose end state is not finite (the reference
 * asserts `success`, odeSolver.py:121).
 */
int64_t kemo_step_fn(kemo_rhs_fn rhs, int ns, int np, int64_t n,
                     double *states, double *params,
                     double t0, double dt, int n_sub, int n_threads)
{
    int64_t bad = 0;
    if (ns > MAX_NS || n_sub < 1) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int64_t r = 0; r < n; ++r) {
        double *y = states + r * ns;
        step_row(rhs, ns, y, params + r * np, t0, dt, n_sub);
        for (int i = 0; i < ns; ++i)
            if (!isfinite(y[i])) { bad += 1; break; }
    }
    return bad;
}

int64_t kemo_step(int model, int64_t n, double *states, double *params,
                  double t0, double dt, int n_sub, int n_threads)
{
    if (model < 0 || model >= N_MODELS) return -1;
    return kemo_step_fn(MODELS[model].rhs, MODELS[model].ns, MODELS[model].np,
                        n, states, params, t0, dt, n_sub, n_threads);
}

/* ------------------------------------------------------------------ scheme O3
 * Error-controlled alternative to O1: Dormand-Prince 5(4) with
 * the step-size controller of scipy.integrate.RK45 (error norm = RMS of
 * e_i / (atol + rtol max(|y_i|, |ynew_i|)), factor = 0.9 err^(-1/5) in [0.2, 10], no growth
 * right after a rejection), restarted every PDE step like the reference restarts LSODA
 * (odeSolver.py:116-120) but warm-started from the step size the DOF last used (`hsug`,
 * one double per row, in/out; <= 0 means dt/8).  The final accepted step ends exactly at
 * t0+dt and its FSAL evaluation f(t0+dt, y_end) is the last RHS call, so the I_ch slots hold
 * I_ch(y(t0+dt)).  This is the CPU twin of csrc/kem_kernel.cuh:kem_step_dp45_kernel: same
 * formulas in the same association.
 */
#define DP_MAX_ATTEMPTS 100000

static int dp45_row(kemo_rhs_fn rhs, int ns, double *y, double *p, double t0, double dt,
                    double rtol, double atol, double *hsug, int64_t *n_acc, int64_t *n_rej)
{
    static const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 4.0 / 5.0, c5 = 8.0 / 9.0;
    static const double a21 = 1.0 / 5.0;
    static const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
    static const double a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0;
    static const double a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0,
                        a54 = -212.0 / 729.0;
    static const double a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0,
                        a64 = 49.0 / 176.0, a65 = -5103.0 / 18656.0;
    static const double b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0,
                        b5 = -2187.0 / 6784.0, b6 = 11.0 / 84.0;
    static const double e1 = 71.0 / 57600.0, e3 = -71.0 / 16695.0, e4 = 71.0 / 1920.0,
                        e5 = -17253.0 / 339200.0, e6 = 22.0 / 525.0, e7 = -1.0 / 40.0;
    double k1[MAX_NS], k2[MAX_NS], k3[MAX_NS], k4[MAX_NS], k5[MAX_NS], k6[MAX_NS], k7[MAX_NS];
    double w[MAX_NS], yn[MAX_NS];
    const double t_end = t0 + dt;
    double t = t0;
    double h_try = *hsug;
    if (!(h_try > 0.0) || !isfinite(h_try)) h_try = dt / 8.0;
    if (h_try > dt) h_try = dt;
    int rejected = 0;
    rhs(t, y, k1, p);
    for (int attempt = 0; attempt < DP_MAX_ATTEMPTS; ++attempt) {
        double h = h_try;
        int last = 0;
        if (t + h * (1.0 + 1e-9) >= t_end) {
            h = t_end - t;
            last = 1;
        }
        for (int i = 0; i < ns; ++i) w[i] = y[i] + h * (a21 * k1[i]);
        rhs(t + c2 * h, w, k2, p);
        for (int i = 0; i < ns; ++i) w[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
        rhs(t + c3 * h, w, k3, p);
        for (int i = 0; i < ns; ++i) w[i] = y[i] + h * ((a41 * k1[i] + a42 * k2[i]) + a43 * k3[i]);
        rhs(t + c4 * h, w, k4, p);
        for (int i = 0; i < ns; ++i)
            w[i] = y[i] + h * (((a51 * k1[i] + a52 * k2[i]) + a53 * k3[i]) + a54 * k4[i]);
        rhs(t + c5 * h, w, k5, p);
        for (int i = 0; i < ns; ++i)
            w[i] = y[i] + h * ((((a61 * k1[i] + a62 * k2[i]) + a63 * k3[i]) + a64 * k4[i]) + a65 * k5[i]);
        const double t_new = last ? t_end : t + h;
        rhs(t_new, w, k6, p);
        for (int i = 0; i < ns; ++i)
            yn[i] = y[i] + h * ((((b1 * k1[i] + b3 * k3[i]) + b4 * k4[i]) + b5 * k5[i]) + b6 * k6[i]);
        rhs(t_new, yn, k7, p);
        double sum = 0.0;
        for (int i = 0; i < ns; ++i) {
            const double ei = h * (((((e1 * k1[i] + e3 * k3[i]) + e4 * k4[i]) + e5 * k5[i]) + e6 * k6[i])
                                   + e7 * k7[i]);
            const double sc = atol + rtol * fmax(fabs(y[i]), fabs(yn[i]));
            const double r = ei / sc;
            sum += r * r;
        }
        const double err = sqrt(sum / (double)ns);
        if (err <= 1.0) {
            double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
            if (rejected) factor = fmin(1.0, factor);
            t = t_new;
            for (int i = 0; i < ns; ++i) { y[i] = yn[i]; k1[i] = k7[i]; }
            *n_acc += 1;
            rejected = 0;
            const double h_next = h * factor;
            h_try = last ? fmax(h_next, h_try) : h_next;
            if (last) {
                *hsug = h_try > dt ? dt : h_try;
                return 0;
            }
        } else {
            const double factor = (err == err) ? fmax(0.2, 0.9 * pow(err, -0.2)) : 0.2;
            h_try = h * factor;
            rejected = 1;
            *n_rej += 1;
            if (!(h_try > 1e-14 * fabs(dt))) break;   /* step size collapsed (or NaN) */
        }
    }
    *hsug = 0.0;
    return 1;   /* failed: the reference's `assert success` */
}

int64_t kemo_step_dp45_fn(kemo_rhs_fn rhs, int ns, int np, int64_t n, double *states, double *params,
                          double *hsug, double t0, double dt, double rtol, double atol, int n_threads,
                          int64_t *stats /* [accepted, rejected] or NULL */)
{
    int64_t bad = 0, acc = 0, rej = 0;
    if (ns > MAX_NS) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : bad, acc, rej)
    for (int64_t r = 0; r < n; ++r) {
        double *y = states + r * ns;
        int64_t a = 0, j = 0;
        int fail = dp45_row(rhs, ns, y, params + r * np, t0, dt, rtol, atol, hsug + r, &a, &j);
        acc += a;
        rej += j;
        for (int i = 0; i < ns && !fail; ++i)
            if (!isfinite(y[i])) fail = 1;
        bad += fail;
    }
    if (stats) { stats[0] = acc; stats[1] = rej; }
    return bad;
}

int64_t kemo_step_dp45(int model, int64_t n, double *states, double *params, double *hsug, double t0,
                       double dt, double rtol, double atol, int n_threads, int64_t *stats)
{
    if (model < 0 || model >= N_MODELS) return -1;
    return kemo_step_dp45_fn(MODELS[model].rhs, MODELS[model].ns, MODELS[model].np, n, states, params,
                             hsug, t0, dt, rtol, atol, n_threads, stats);
}

int kemo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
