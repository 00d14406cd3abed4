/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Plain-C CPU restatement of fv2d's per-timestep finite-volume update (the hot path of
 * SURVEY.md §8a), written from the reference's algorithm with every function citing the
 * reference file:line it follows.  Pinned against the reference itself: tests/golden/
 * holds dumps produced by oracle/_ref/fv2d_ref (the unmodified reference headers +
 * vendored Kokkos-OpenMP) and tests/test_oracle_vs_reference.py requires bit-identical
 * dt sequences and states.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (fv2d_b200/) never does.
 *
 * Array layout: SoA planes over the FULL grid incl. ghosts, A[f][j][i] with
 * f in 0..3, j in 0..Nty-1, i in 0..Ntx-1, no padding (index (f*Nty + j)*Ntx + i).
 */
#ifndef FV2D_ORACLE_H_
#define FV2D_ORACLE_H_

#include "../include/fv2d_params.h"

#ifdef __cplusplus
extern "C" {
#endif

/* States.h:19-43 */
void fv2d_oracle_prim_to_cons_state(const fv2d_device_params *p, const double q[4], double u[4]);
void fv2d_oracle_cons_to_prim_state(const fv2d_device_params *p, const double u[4], double q[4]);

/* RiemannSolvers.h:7-171; solver = FV2D_HLL / FV2D_HLLC / FV2D_FSLP; states in the rotated frame */
void fv2d_oracle_riemann(const fv2d_device_params *p, int solver, const double qL[4], const double qR[4], double gdx,
                         double flux[4], double *pout);

/* Gravity.h:39-57 (float-valued, Q5) */
double fv2d_oracle_get_gravity(const fv2d_device_params *p, int i, int j, int dir);

/* SimInfo.h:576-600 (range_tot) */
void fv2d_oracle_cons_to_prim(const fv2d_device_params *p, const double *U, double *Q);
void fv2d_oracle_prim_to_cons(const fv2d_device_params *p, const double *Q, double *U);

/* SimInfo.h:602-646; counts = {negative_density, negative_pressure, nan_count} */
void fv2d_oracle_check_negatives(const fv2d_device_params *p, double eps_reset, double *Q, uint64_t counts[3]);

/* BoundaryConditions.h:82-147 */
void fv2d_oracle_fill_boundaries(const fv2d_device_params *p, double *Q);

/* ComputeDt.h:18-65; inv_dt = {hyp, tc, visc}; returns CFL / max */
double fv2d_oracle_compute_dt(const fv2d_device_params *p, const double *Q, double inv_dt[3]);

/* Update.h:59-91 */
void fv2d_oracle_compute_slopes(const fv2d_device_params *p, const double *Q, double *slopesX, double *slopesY);
/* Update.h:93-174 */
void fv2d_oracle_compute_fluxes_and_update(const fv2d_device_params *p, const double *Q, const double *slopesX,
                                           const double *slopesY, double *Unew, double dt);
/* ThermalConduction.h:36-108; returns non-zero for the unsupported TCM_B02 mode */
int fv2d_oracle_apply_thermal_conduction(const fv2d_device_params *p, const double *Q, double *Unew, double dt);
/* Viscosity.h:27-119 */
void fv2d_oracle_apply_viscosity(const fv2d_device_params *p, const double *Q, double *Unew, double dt);

/* Update.h:176-191 */
int fv2d_oracle_euler_step(const fv2d_device_params *p, double *Q, double *Unew, double dt);
/* Update.h:193-222 */
int fv2d_oracle_update(const fv2d_device_params *p, int time_stepping, double *Q, double *Unew, double dt);

/* main.cpp:62-84 without IO: runs at most max_steps steps while t + epsilon < tend.
 * dts (may be NULL) receives the dt sequence; neg_counts (may be NULL) accumulates the
 * checkNegatives counters.  Returns the number of steps done, or <0 on error. */
long fv2d_oracle_run(const fv2d_device_params *p, int time_stepping, double eps_reset, double tend, double *Q,
                     double *U, long max_steps, double *t_inout, double *dts, uint64_t neg_counts[3]);

#ifdef __cplusplus
}
#endif
#endif
