/* edm_oracle.c — CPU oracle of the lift -> evolve -> restrict map.  TEST INFRASTRUCTURE
 * ONLY; PARITY UNPINNED (see oracle.h and the header of edm_oracle_impl.inc). */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* ---- counter-based standard normal (replaces cuRAND XORWOW, EventDrivenMap.cu:103,179;
 * same seed re-applied on every ComputeF = common random numbers) ---- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
double oracle_normal(uint64_t seed, uint64_t index) {
  uint64_t h1 = splitmix64(seed ^ splitmix64(2 * index));
  uint64_t h2 = splitmix64(seed ^ splitmix64(2 * index + 1));
  double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

void oracle_edm_cfg_default(oracle_edm_cfg* c) {
  /* parameters.hpp:1-15; float literals widened exactly */
  c->vth = (double)1.0f; c->a1 = (double)11.0f; c->a2 = (double)7.0f;
  c->b1 = (double)5.0f; c->b2 = (double)3.5f; c->I = (double)0.9f; c->L = (double)3.0f;
  c->tol = 1e-6; c->time_horizon = (double)5.0f;
  c->counter_max = 100; c->quirks = 0;
  c->beta = (double)13.0589f;  /* Driver.cu:16 */
  c->sigma = 0.0;              /* EventDrivenMap.cu:105 */
  c->seed = 42;
  c->N = 1024;                 /* EventDrivenMap.cu:70 */
  c->R = 1000;                 /* Driver.cu:19 */
  c->M = 3;                    /* parameters.hpp:12 */
  c->precision = 0;
  c->beta_ext = NULL;
}

void oracle_edm_beta(const oracle_edm_cfg* cfg, double* beta_out) {
  size_t n = (size_t)cfg->R * cfg->N;
  for (size_t i = 0; i < n; ++i)
    beta_out[i] = cfg->beta + cfg->sigma * oracle_normal(cfg->seed, (uint64_t)i);
}

#define REAL double
#define SUFFIX(name) edm_##name##_f64
#define R_EXP exp
#define R_POW pow
#define R_FABS fabs
#include "edm_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef R_EXP
#undef R_POW
#undef R_FABS

#define REAL float
#define SUFFIX(name) edm_##name##_f32
#define R_EXP expf
#define R_POW powf
#define R_FABS fabsf
#include "edm_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef R_EXP
#undef R_POW
#undef R_FABS

int oracle_edm_compute_f(const oracle_edm_cfg* cfg, const double* z, double* f_out,
                         oracle_edm_aux* aux, uint32_t r_begin, uint32_t r_end,
                         int nthreads) {
  if (!cfg || !z || !f_out || cfg->M < 1 || cfg->N < 2 || cfg->R < 1) return -1;
  return cfg->precision ? edm_compute_f_f32(cfg, z, f_out, aux, r_begin, r_end, nthreads)
                        : edm_compute_f_f64(cfg, z, f_out, aux, r_begin, r_end, nthreads);
}

int oracle_profile_compute_f(const oracle_edm_cfg* cfg, uint32_t n_coarse, const double* u, double* f_out,
                             double* restricted_out, int32_t* accept_out, int32_t* event_count_out,
                             uint32_t r_begin, uint32_t r_end, int nthreads) {
  if (!cfg || !u || !f_out || cfg->N < 2 || cfg->R < 1) return -1;
  return cfg->precision ? edm_profile_compute_f_f32(cfg, n_coarse, u, f_out, restricted_out, accept_out, event_count_out, r_begin, r_end, nthreads)
                        : edm_profile_compute_f_f64(cfg, n_coarse, u, f_out, restricted_out, accept_out, event_count_out, r_begin, r_end, nthreads);
}

/* the analytic travelling-wave lift (LiftKernel) sampled on the fine grid: a physically sensible
 * base state for the profile map */
int oracle_edm_lift(const oracle_edm_cfg* cfg, const double* z, double* v_out, double* s_out) {
  oracle_edm_aux aux;
  memset(&aux, 0, sizeof(aux));
  aux.lift_v = v_out; aux.lift_s = s_out;
  oracle_edm_cfg c = *cfg;
  c.R = 1;
  double* f = (double*)malloc(sizeof(double) * c.M);
  /* evolving one realisation is the price of reusing compute_f; cheap at test sizes */
  int rc = oracle_edm_compute_f(&c, z, f, &aux, 0, 1, 1);
  free(f);
  return rc;
}

int oracle_edm_compute_dfdu(const oracle_edm_cfg* cfg, const double* u, double eps,
                            double* jac_out, double* f0_out, int nthreads) {
  const uint32_t n = cfg->M;
  double* f0 = (double*)malloc(sizeof(double) * n);
  double* df = (double*)malloc(sizeof(double) * n);
  double* du = (double*)malloc(sizeof(double) * n);
  int rc = oracle_edm_compute_f(cfg, u, f0, NULL, 0, 0, nthreads);
  memcpy(du, u, sizeof(double) * n);
  for (uint32_t i = 0; i < n && rc == 0; ++i) {
    if (i > 0) du[i - 1] = u[i - 1];                /* NewtonSolver.cpp:184-187 */
    du[i] += eps;                                   /* :188 */
    rc = oracle_edm_compute_f(cfg, du, df, NULL, 0, 0, nthreads);
    for (uint32_t r = 0; r < n; ++r)
      jac_out[(size_t)i * n + r] = (df[r] - f0[r]) * pow(eps, -1);   /* :194 */
  }
  if (f0_out) memcpy(f0_out, f0, sizeof(double) * n);
  free(f0); free(df); free(du);
  return rc;
}
