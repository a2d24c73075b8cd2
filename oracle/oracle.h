/* oracle.h — CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.  The
 * product (armadillocudalinearinterpolation_b200/) never links, imports or calls it.
 *
 * PARITY.
 *   map (edm_*): PINNED against the real reference.  oracle/ref_build/Makefile compiles the UNMODIFIED
 *     EventDrivenMap.cu / NewtonSolver.cpp / Stability.cpp for sm_100a (two macros: `counterMax` = 100, which
 *     EventDrivenMap.cu:564 leaves undefined, and `__shfl_down` -> `__shfl_down_sync`) into oracle/_ref/; its raw
 *     outputs on a B200 are committed as tests/golden/ref_b200.npz (tools/make_ref_golden.py).  The FP32 / Q1 mode
 *     of this oracle reproduces them: every integer output exactly, floats to FP32 noise, and the reference
 *     NewtonSolver's residual history (tests/test_ref_pin.py).  The FP64 mode is the same template.
 *   interp1 / interp2: PARITY UNPINNED against Armadillo itself — an un-vendored, un-pinned dependency of the
 *     reference (Makefile:5 `-larmadillo`, Driver.o.dep:554) that is absent from this image and that the
 *     reference never calls for interpolation.  Restated from Armadillo's published algorithm
 *     (fn_interp1.hpp / fn_interp2.hpp); second opinions: numpy.interp, scipy RegularGridInterpolator.
 *
 * Compile with -O2 -ffp-contract=off (no FMA contraction) — see oracle/Makefile.
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ interp ---- */

/* arma::interp1(XG,YG,XI,YI,"*linear",extrap) — interp1_helper_linear of fn_interp1.hpp.
 * `scan` != 0 uses Armadillo's own monotone nearest-knot scan (requires sorted XI);
 * `scan` == 0 uses an order-independent binary search with the same per-query result.
 * idx_out (nullable): lower bracket index a, or -1 for extrapolated / NaN queries.
 * nthreads > 1 parallelises over queries with OpenMP (binary-search mode only). */
int oracle_interp1_f64(const double* xg, const double* yg, size_t ng, const double* xi,
                       size_t ni, double* yi, int32_t* idx_out, double extrap, int scan,
                       int nthreads);
int oracle_interp1_f32(const float* xg, const float* yg, size_t ng, const float* xi,
                       size_t ni, float* yi, int32_t* idx_out, float extrap, int scan,
                       int nthreads);

/* arma::interp2(X,Y,Z,XI,YI,ZI,"linear",extrap) — two separable passes of the interp1 rule, fn_interp2.hpp.
 * Order: along X first (whole columns of Z), then along Y — as recalled by two independent readers of
 * fn_interp2.hpp; UNVERIFIED (Armadillo is absent).  oracle_interp2_set_order(1) selects the mirrored order
 * (along Y, then X), matching the product's B200_INTERP2_ORDER_YX flag.
 * z: ny x nx column-major; zi: nyi x nxi column-major. */
void oracle_interp2_set_order(int y_first);
int oracle_interp2_grid_f64(const double* x, size_t nx, const double* y, size_t ny,
                            const double* z, const double* xi, size_t nxi,
                            const double* yi, size_t nyi, double* zi, double extrap,
                            int nthreads);
int oracle_interp2_grid_f32(const float* x, size_t nx, const float* y, size_t ny,
                            const float* z, const float* xi, size_t nxi, const float* yi,
                            size_t nyi, float* zi, float extrap, int nthreads);

/* Per-point restatement of the same two passes for scattered (xq[k], yq[k]). */
int oracle_interp2_scattered_f64(const double* x, size_t nx, const double* y, size_t ny,
                                 const double* z, const double* xq, const double* yq,
                                 size_t nq, double* zq, double extrap, int nthreads);
int oracle_interp2_scattered_f32(const float* x, size_t nx, const float* y, size_t ny,
                                 const float* z, const float* xq, const float* yq,
                                 size_t nq, float* zq, float extrap, int nthreads);

/* --------------------------------------------------------------------- map ---- */

typedef struct {
  /* parameters.hpp:1-15 */
  double vth, a1, a2, b1, b2, I, L, tol, time_horizon;
  uint32_t counter_max;      /* EventDrivenMap.cu:564 (undefined in the reference) */
  uint32_t quirks;           /* bit 0: accept[0] bias, EventDrivenMap.cu:801,817,822 */
  /* run shape */
  double beta;               /* p[0], Driver.cu:15-16 */
  double sigma;              /* mParStdDev, EventDrivenMap.cu:105,317 */
  uint64_t seed;
  uint32_t N;                /* neurons per ring = mNoThreads, EventDrivenMap.cu:70 */
  uint32_t R;                /* realisations = mNoReal, Driver.cu:19 */
  uint32_t M;                /* fronts = noSpikes, parameters.hpp:12 */
  uint32_t precision;        /* 0 = double everywhere, 1 = the reference's float device math */
  const double* beta_ext;    /* nullable [R][N]: use this ensemble instead of the generator
                                (parity tests feed the ensemble fetched from the GPU) */
} oracle_edm_cfg;

/* Optional per-run outputs (any pointer may be NULL). Shapes in comments. */
typedef struct {
  int32_t* init_index;    /* [M]      EventDrivenMap.cu:361-376 */
  double*  lift_v;        /* [N]      EventDrivenMap.cu:505-542 (identical for every realisation) */
  double*  lift_s;        /* [N] */
  int32_t* last_index;    /* [R][M]   EventDrivenMap.cu:661-668 */
  double*  last_time;     /* [R][M] */
  int32_t* crossed_index; /* [R][M] */
  double*  crossed_time;  /* [R][M] */
  int32_t* accept;        /* [R]      EventDrivenMap.cu:669-672 */
  double*  position;      /* [R][M]   EventDrivenMap.cu:769-785 */
  int32_t* event_count;   /* [R] */
  double*  mean;          /* [M]      EventDrivenMap.cu:805-824 */
  double*  beta;          /* [R][N]   per-neuron beta */
  double*  coupling;      /* [N]      EventDrivenMap.cu:111-129 */
  int32_t  init_index_clamped; /* out: Q15 soft flag (a front started outside the domain) */
  /* out: work counters over the evolved realisations (the bench's unit of work is the
   * neuron-event update = events x N) */
  uint64_t total_events, total_neuron_events, total_candidates, total_newton_its;
} oracle_edm_aux;

void oracle_edm_cfg_default(oracle_edm_cfg* cfg);

/* F(z), EventDrivenMap.cu:154-240.  z[M] = (c, T_2..T_M).  Realisations [r_begin, r_end)
 * only are evolved when r_end > r_begin (bounded sample for the CPU baseline); pass
 * 0,0 for all R.  nthreads: OpenMP threads over realisations.  Returns 0 on success. */
int oracle_edm_compute_f(const oracle_edm_cfg* cfg, const double* z, double* f_out,
                         oracle_edm_aux* aux, uint32_t r_begin, uint32_t r_end,
                         int nthreads);

/* beta ensemble: beta[r*N+j] = mean + sigma * normal(seed, r*N+j). */
void oracle_edm_beta(const oracle_edm_cfg* cfg, double* beta_out);
/* Forward-difference Jacobian exactly as NewtonSolver.cpp:164-197 / Stability.cpp:76-111
 * form it: f0 = F(u); J(:,i) = (F(u + eps e_i) - f0) * pow(eps,-1).  jac n x n col-major. */
int oracle_edm_compute_dfdu(const oracle_edm_cfg* cfg, const double* u, double eps,
                            double* jac_out, double* f0_out, int nthreads);
/* Profile map Phi_T(u) - u on coarse profiles u = (V_c[n_coarse], S_c[n_coarse]) — NEW functionality
 * (BASELINE config 5; not in the reference); definition at the head of the function in
 * edm_oracle_impl.inc.  restricted_out [nr][2 n_coarse], accept_out [nr], event_count_out [nr] nullable. */
int oracle_profile_compute_f(const oracle_edm_cfg* cfg, uint32_t n_coarse, const double* u, double* f_out,
                             double* restricted_out, int32_t* accept_out, int32_t* event_count_out,
                             uint32_t r_begin, uint32_t r_end, int nthreads);
/* LiftKernel (EventDrivenMap.cu:505-542) on the fine grid: v_out[N], s_out[N] */
int oracle_edm_lift(const oracle_edm_cfg* cfg, const double* z, double* v_out, double* s_out);
/* the counter-based standard normal used for it (exposed for the RNG parity test) */
double oracle_normal(uint64_t seed, uint64_t index);

#ifdef __cplusplus
}
#endif
#endif
