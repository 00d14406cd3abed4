/* TEST INFRASTRUCTURE — not part of the product.  See fv2d_oracle.h.
 *
 * Plain-C restatement of the reference's per-timestep update.  The expression order of
 * every formula follows the reference source so that, compiled without FMA contraction
 * (-ffp-contract=off, x86-64 baseline), results are bit-identical to the reference's
 * Kokkos-OpenMP Release build (checked in tests/test_init_and_oracle.py against the
 * dumps in tests/golden/).
 */
#include "fv2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IR FV2D_IR
#define IU FV2D_IU
#define IV FV2D_IV
#define IP FV2D_IP
#define IE FV2D_IE
#define IX FV2D_IX
#define IY FV2D_IY

typedef struct
{
  double v[4];
} State;

static inline size_t idx(const fv2d_device_params *p, int f, int i, int j)
{
  return ((size_t)f * (size_t)p->Nty + (size_t)j) * (size_t)p->Ntx + (size_t)i;
}

/* States.h:6-17 */
static inline State get_state(const fv2d_device_params *p, const double *A, int i, int j)
{
  State s;
  for (int f = 0; f < 4; ++f)
    s.v[f] = A[idx(p, f, i, j)];
  return s;
}
static inline void set_state(const fv2d_device_params *p, double *A, int i, int j, State s)
{
  for (int f = 0; f < 4; ++f)
    A[idx(p, f, i, j)] = s.v[f];
}

/* States.h:19-30 */
static inline State prim_to_cons(const fv2d_device_params *p, State q)
{
  State r;
  r.v[IR]   = q.v[IR];
  r.v[IU]   = q.v[IR] * q.v[IU];
  r.v[IV]   = q.v[IR] * q.v[IV];
  double Ek = 0.5 * (r.v[IU] * r.v[IU] + r.v[IV] * r.v[IV]) / q.v[IR];
  r.v[IE]   = (Ek + q.v[IP] / (p->gamma0 - 1.0));
  return r;
}

/* States.h:32-43 */
static inline State cons_to_prim(const fv2d_device_params *p, State u)
{
  State r;
  r.v[IR]   = u.v[IR];
  r.v[IU]   = u.v[IU] / u.v[IR];
  r.v[IV]   = u.v[IV] / u.v[IR];
  double Ek = 0.5 * r.v[IR] * (r.v[IU] * r.v[IU] + r.v[IV] * r.v[IV]);
  r.v[IP]   = (u.v[IE] - Ek) * (p->gamma0 - 1.0);
  return r;
}

/* States.h:45-46 */
static inline double speed_of_sound(const fv2d_device_params *p, State q) { return sqrt(q.v[IP] * p->gamma0 / q.v[IR]); }

/* States.h:103-110 */
static inline State swap_component(State q, int dir)
{
  if (dir == IX)
    return q;
  State r = {{q.v[IR], q.v[IV], q.v[IU], q.v[IP]}};
  return r;
}

void fv2d_oracle_prim_to_cons_state(const fv2d_device_params *p, const double q[4], double u[4])
{
  State s;
  memcpy(s.v, q, sizeof s.v);
  s = prim_to_cons(p, s);
  memcpy(u, s.v, sizeof s.v);
}
void fv2d_oracle_cons_to_prim_state(const fv2d_device_params *p, const double u[4], double q[4])
{
  State s;
  memcpy(s.v, u, sizeof s.v);
  s = cons_to_prim(p, s);
  memcpy(q, s.v, sizeof s.v);
}

/* ---------------------------------------------------------------- Riemann solvers */

/* RiemannSolvers.h:18-28 (lambda computeFlux inside hll) */
static inline State hll_phys_flux(const fv2d_device_params *p, State q)
{
  const double Ek = 0.5 * q.v[IR] * (q.v[IU] * q.v[IU] + q.v[IV] * q.v[IV]);
  const double E  = (q.v[IP] / (p->gamma0 - 1.0) + Ek);
  State f = {{q.v[IR] * q.v[IU], q.v[IR] * q.v[IU] * q.v[IU] + q.v[IP], q.v[IR] * q.v[IU] * q.v[IV],
              (q.v[IP] + E) * q.v[IU]}};
  return f;
}

/* RiemannSolvers.h:7-51 */
static void hll(const fv2d_device_params *p, State qL, State qR, State *flux, double *pout)
{
  const double aL = speed_of_sound(p, qL);
  const double aR = speed_of_sound(p, qR);

  const double sminL = qL.v[IU] - aL;
  const double smaxL = qL.v[IU] + aL;
  const double sminR = qR.v[IU] - aR;
  const double smaxR = qR.v[IU] + aR;

  const double SL = fmin(sminL, sminR);
  const double SR = fmax(smaxL, smaxR);

  State FL = hll_phys_flux(p, qL);
  State FR = hll_phys_flux(p, qR);

  if (SL >= 0.0)
  {
    *flux = FL;
    *pout = qL.v[IP];
  }
  else if (SR <= 0.0)
  {
    *flux = FR;
    *pout = qR.v[IP];
  }
  else
  {
    State uL = prim_to_cons(p, qL);
    State uR = prim_to_cons(p, qR);
    *pout    = 0.5 * (qL.v[IP] + qR.v[IP]);
    /* (SR * FL - SL * FR + SL * SR * (uR - uL)) / (SR - SL), States.h operator order */
    const double SLSR = SL * SR;
    for (int f = 0; f < 4; ++f)
      flux->v[f] = ((FL.v[f] * SR - FR.v[f] * SL) + (uR.v[f] - uL.v[f]) * SLSR) / (SR - SL);
  }
}

/* RiemannSolvers.h:53-128 */
static void hllc(const fv2d_device_params *p, State qL, State qR, State *flux, double *pout)
{
  const double rL = qL.v[IR], uL = qL.v[IU], vL = qL.v[IV], pL = qL.v[IP];
  const double rR = qR.v[IR], uR = qR.v[IU], vR = qR.v[IV], pR = qR.v[IP];

  const double entho = 1.0 / (p->gamma0 - 1.0);

  const double ekL = 0.5 * rL * (uL * uL + vL * vL);
  const double EL  = ekL + pL * entho;
  const double ekR = 0.5 * rR * (uR * uR + vR * vR);
  const double ER  = ekR + pR * entho;

  const double cfastL = speed_of_sound(p, qL);
  const double cfastR = speed_of_sound(p, qR);

  const double SL = fmin(uL, uR) - fmax(cfastL, cfastR);
  const double SR = fmax(uL, uR) + fmax(cfastL, cfastR);

  const double rcL = rL * (uL - SL);
  const double rcR = rR * (SR - uR);

  const double uS = (rcR * uR + rcL * uL + (pL - pR)) / (rcR + rcL);
  const double pS = (rcR * pL + rcL * pR + rcL * rcR * (uL - uR)) / (rcR + rcL);

  const double rSL = rL * (SL - uL) / (SL - uS);
  const double ESL = ((SL - uL) * EL - pL * uL + pS * uS) / (SL - uS);

  const double rSR = rR * (SR - uR) / (SR - uS);
  const double ESR = ((SR - uR) * ER - pR * uR + pS * uS) / (SR - uS);

  State st;
  double E;
  if (SL > 0.0)
  {
    st    = qL;
    E     = EL;
    *pout = pL;
  }
  else if (uS > 0.0)
  {
    st.v[IR] = rSL;
    st.v[IU] = uS;
    st.v[IV] = qL.v[IV];
    st.v[IP] = pS;
    E        = ESL;
    *pout    = pS;
  }
  else if (SR > 0.0)
  {
    st.v[IR] = rSR;
    st.v[IU] = uS;
    st.v[IV] = qR.v[IV];
    st.v[IP] = pS;
    E        = ESR;
    *pout    = pS;
  }
  else
  {
    st    = qR;
    E     = ER;
    *pout = pR;
  }

  flux->v[IR] = st.v[IR] * st.v[IU];
  flux->v[IU] = st.v[IR] * st.v[IU] * st.v[IU] + st.v[IP];
  flux->v[IV] = flux->v[IR] * st.v[IV];
  flux->v[IE] = (E + st.v[IP]) * st.v[IU];
}

/* RiemannSolvers.h:137-171 */
static void fslp(const fv2d_device_params *p, State qL, State qR, State *flux, double *pout, double gdx)
{
  const double rhoL = qL.v[IR], uL = qL.v[IU], pL = qL.v[IP];
  const double csL  = sqrt(p->gamma0 * pL / rhoL);
  const double rhoR = qR.v[IR], uR = qR.v[IU], pR = qR.v[IP];
  const double csR  = sqrt(p->gamma0 * pR / rhoR);

  const double a1    = rhoL * csL, a2 = rhoR * csR;
  const double ai    = p->fslp_K * (a1 < a2 ? a2 : a1);           /* Kokkos::max(a,b) = a<b ? b : a */
  const double m1    = fabs(uL) / csL, m2 = fabs(uR) / csR;
  const double mm    = (m1 < m2 ? m2 : m1);
  const double theta = (mm < 1.0 ? mm : 1.0);                     /* Kokkos::min(1.0, mm) = mm<1.0 ? mm : 1.0 */

  const double ustar = 0.5 * (uR + uL) - 0.5 / ai * (pR - pL - 0.5 * (rhoL + rhoR) * gdx);
  const double Pi    = 0.5 * (pR + pL) - theta * 0.5 * ai * (uR - uL);

  const State *qs     = (ustar > 0 ? &qL : &qR);
  const double Ekstar = 0.5 * qs->v[IR] * (qs->v[IU] * qs->v[IU] + qs->v[IV] * qs->v[IV]);
  const double Estar  = Ekstar + qs->v[IP] / (p->gamma0 - 1.0);

  flux->v[IR] = ustar * qs->v[IR];
  flux->v[IU] = ustar * qs->v[IR] * qs->v[IU] + Pi;
  flux->v[IV] = ustar * qs->v[IR] * qs->v[IV];
  flux->v[IE] = ustar * (Estar + Pi);
  *pout       = Pi;
}

/* Update.h:121-135 (lambda riemann) */
static inline void riemann(const fv2d_device_params *p, int solver, State qL, State qR, double gdx, State *flux,
                           double *pout)
{
  switch (solver)
  {
  case FV2D_HLL:
    hll(p, qL, qR, flux, pout);
    break;
  case FV2D_FSLP:
    fslp(p, qL, qR, flux, pout, gdx);
    break;
  default:
    hllc(p, qL, qR, flux, pout);
    break;
  }
}

void fv2d_oracle_riemann(const fv2d_device_params *p, int solver, const double qL[4], const double qR[4], double gdx,
                         double flux[4], double *pout)
{
  State a, b, f;
  memcpy(a.v, qL, sizeof a.v);
  memcpy(b.v, qR, sizeof b.v);
  riemann(p, solver, a, b, gdx, &f, pout);
  memcpy(flux, f.v, sizeof f.v);
}

/* ---------------------------------------------------------------- gravity */

/* SimInfo.h:494-499 */
static inline void get_pos(const fv2d_device_params *p, int i, int j, double pos[2])
{
  pos[IX] = p->xmin + (i - p->ibeg + 0.5) * p->dx;
  pos[IY] = p->ymin + (j - p->jbeg + 0.5) * p->dy;
}

/* Gravity.h:15-29: returns float (Q5) */
static inline float get_analytical_gravity(const fv2d_device_params *p, int i, int j, int dir)
{
  (void)dir;
  double pos[2];
  get_pos(p, i, j, pos);
  double g = p->hot_bubble_g0 * sin(pos[IY] * M_PI * 2.0 / p->ymax);
  return (float)g;
}

/* Gravity.h:38-57: returns float (Q5) */
static inline float get_gravity(const fv2d_device_params *p, int i, int j, int dir)
{
  double g;
  switch (p->gravity_mode)
  {
  case FV2D_GRAV_CONSTANT:
    g = (dir == IX ? p->gx : p->gy);
    break;
  case FV2D_GRAV_ANALYTICAL:
    g = get_analytical_gravity(p, i, j, dir);
    break;
  case FV2D_GRAV_NONE:
  default:
    g = 0.0;
    break;
  }
  return (float)g;
}

double fv2d_oracle_get_gravity(const fv2d_device_params *p, int i, int j, int dir) { return get_gravity(p, i, j, dir); }

/* ---------------------------------------------------------------- array-level conversions */

/* SimInfo.h:576-587 (range_tot) */
void fv2d_oracle_cons_to_prim(const fv2d_device_params *p, const double *U, double *Q)
{
#pragma omp parallel for schedule(static)
  for (int j = 0; j < p->Nty; ++j)
    for (int i = 0; i < p->Ntx; ++i)
      set_state(p, Q, i, j, cons_to_prim(p, get_state(p, U, i, j)));
}

/* SimInfo.h:589-600 (range_tot) */
void fv2d_oracle_prim_to_cons(const fv2d_device_params *p, const double *Q, double *U)
{
#pragma omp parallel for schedule(static)
  for (int j = 0; j < p->Nty; ++j)
    for (int i = 0; i < p->Ntx; ++i)
      set_state(p, U, i, j, prim_to_cons(p, get_state(p, Q, i, j)));
}

/* SimInfo.h:602-646 */
void fv2d_oracle_check_negatives(const fv2d_device_params *p, double eps_reset, double *Q, uint64_t counts[3])
{
  uint64_t nd = 0, np = 0, nn = 0;
#pragma omp parallel for schedule(static) reduction(+ : nd, np, nn)
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = p->ibeg; i < p->iend; ++i)
    {
      if (Q[idx(p, IR, i, j)] < 0)
      {
        Q[idx(p, IR, i, j)] = eps_reset;
        nd++;
      }
      if (Q[idx(p, IP, i, j)] < 0)
      {
        Q[idx(p, IP, i, j)] = eps_reset;
        np++;
      }
      for (int f = 0; f < 4; ++f)
        if (isnan(Q[idx(p, f, i, j)]))
          nn++;
    }
  counts[0] = nd;
  counts[1] = np;
  counts[2] = nn;
}

/* ---------------------------------------------------------------- boundary conditions */

/* BoundaryConditions.h:22-46 */
static inline State fill_reflecting(const fv2d_device_params *p, const double *Q, int i, int j, int iref, int jref,
                                    int dir)
{
  int isym, jsym;
  if (dir == IX)
  {
    int ipiv = (i < iref ? p->ibeg : p->iend);
    isym     = 2 * ipiv - i - 1;
    jsym     = j;
  }
  else
  {
    int jpiv = (j < jref ? p->jbeg : p->jend);
    isym     = i;
    jsym     = 2 * jpiv - j - 1;
  }
  State q = get_state(p, Q, isym, jsym);
  if (dir == IX)
    q.v[IU] *= -1.0;
  else
    q.v[IV] *= -1.0;
  return q;
}

/* BoundaryConditions.h:52-71 */
static inline State fill_periodic(const fv2d_device_params *p, const double *Q, int i, int j, int dir)
{
  if (dir == IX)
  {
    if (i < p->ibeg)
      i += p->Nx;
    else
      i -= p->Nx;
  }
  else
  {
    if (j < p->jbeg)
      j += p->Ny;
    else
      j -= p->Ny;
  }
  return get_state(p, Q, i, j);
}

static inline State bc_fill(const fv2d_device_params *p, const double *Q, int bc, int i, int j, int iref, int jref,
                            int dir)
{
  switch (bc)
  {
  default:
  case FV2D_BC_ABSORBING:
    return get_state(p, Q, iref, jref); /* BoundaryConditions.h:15-16 */
  case FV2D_BC_REFLECTING:
    return fill_reflecting(p, Q, i, j, iref, jref, dir);
  case FV2D_BC_PERIODIC:
    return fill_periodic(p, Q, i, j, dir);
  }
}

/* BoundaryConditions.h:82-147: x-pass over (i in [0,Ng), j in [jbeg,jend)), then y-pass
 * over (i in [0,Ntx), j in [0,Ng)).  Left/right (top/bottom) are written one after the
 * other for each (i,j), as in the reference lambdas. */
void fv2d_oracle_fill_boundaries(const fv2d_device_params *p, double *Q)
{
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = 0; i < p->Ng; ++i)
    {
      int ileft = i, iright = p->iend + i;
      int iref_left = p->ibeg, iref_right = p->iend - 1;
      set_state(p, Q, ileft, j, bc_fill(p, Q, p->boundary_x, ileft, j, iref_left, j, IX));
      set_state(p, Q, iright, j, bc_fill(p, Q, p->boundary_x, iright, j, iref_right, j, IX));
    }
  for (int j = 0; j < p->Ng; ++j)
    for (int i = 0; i < p->Ntx; ++i)
    {
      int jtop = j, jbot = p->jend + j;
      int jref_top = p->jbeg, jref_bot = p->jend - 1;
      set_state(p, Q, i, jtop, bc_fill(p, Q, p->boundary_y, i, jtop, i, jref_top, IY));
      set_state(p, Q, i, jbot, bc_fill(p, Q, p->boundary_y, i, jbot, i, jref_bot, IY));
    }
}

/* ---------------------------------------------------------------- time step */

/* ThermalConduction.h:8-26 (TCM_CONSTANT only; TCM_B02 is undefined behaviour in the reference, Q7b) */
static inline double compute_kappa(const fv2d_device_params *p) { return p->kappa; }
/* Viscosity.h:8-17 */
static inline double compute_mu(const fv2d_device_params *p) { return p->mu; }

/* ComputeDt.h:18-65 */
double fv2d_oracle_compute_dt(const fv2d_device_params *p, const double *Q, double inv_dt[3])
{
  double m_hyp = -1.7976931348623157e308, m_tc = -1.7976931348623157e308, m_visc = -1.7976931348623157e308;
#pragma omp parallel for schedule(static) reduction(max : m_hyp, m_tc, m_visc)
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = p->ibeg; i < p->iend; ++i)
    {
      State q   = get_state(p, Q, i, j);
      double cs = speed_of_sound(p, q);

      double hyp = (cs + fabs(q.v[IU])) / p->dx + (cs + fabs(q.v[IV])) / p->dy;

      double tc = p->epsilon;
      if (p->thermal_conductivity_active)
        tc = fmax(2.0 * compute_kappa(p) / (p->dx * p->dx), 2.0 * compute_kappa(p) / (p->dy * p->dy));

      double visc = p->epsilon;
      if (p->viscosity_active)
        visc = fmax(2.0 * compute_mu(p) / (p->dx * p->dx), 2.0 * compute_mu(p) / (p->dy * p->dy));

      m_hyp  = fmax(m_hyp, hyp);
      m_tc   = fmax(m_tc, tc);
      m_visc = fmax(m_visc, visc);
    }
  if (inv_dt)
  {
    inv_dt[0] = m_hyp;
    inv_dt[1] = m_tc;
    inv_dt[2] = m_visc;
  }
  double m = m_hyp;
  if (m < m_tc)
    m = m_tc;
  if (m < m_visc)
    m = m_visc; /* std::max({a,b,c}) */
  return p->CFL / m;
}

/* ---------------------------------------------------------------- hyperbolic update */

/* Update.h:69-85 */
static inline double minmod(double dL, double dR)
{
  if (dL * dR < 0.0)
    return 0.0;
  else if (fabs(dL) < fabs(dR))
    return dL;
  else
    return dR;
}

/* Update.h:59-91, range_slopes = dom +- 1 (SimInfo.h:562-563) */
void fv2d_oracle_compute_slopes(const fv2d_device_params *p, const double *Q, double *slopesX, double *slopesY)
{
#pragma omp parallel for schedule(static)
  for (int j = p->jbeg - 1; j < p->jend + 1; ++j)
    for (int i = p->ibeg - 1; i < p->iend + 1; ++i)
      for (int f = 0; f < 4; ++f)
      {
        double dL = Q[idx(p, f, i, j)] - Q[idx(p, f, i - 1, j)];
        double dR = Q[idx(p, f, i + 1, j)] - Q[idx(p, f, i, j)];
        double dU = Q[idx(p, f, i, j)] - Q[idx(p, f, i, j - 1)];
        double dD = Q[idx(p, f, i, j + 1)] - Q[idx(p, f, i, j)];

        slopesX[idx(p, f, i, j)] = minmod(dL, dR);
        slopesY[idx(p, f, i, j)] = minmod(dU, dD);
      }
}

/* Update.h:14-37.  PCM_WB computes a pressure extrapolation and then falls through to
 * `res = q` (missing break, Q3), so it is PCM. */
static inline State reconstruct(const fv2d_device_params *p, const double *Q, const double *slopes, int i, int j,
                                double sign, int dir)
{
  State q = get_state(p, Q, i, j);
  State res;
  switch (p->reconstruction)
  {
  case FV2D_PLM:
  {
    State slope = get_state(p, slopes, i, j);
    for (int f = 0; f < 4; ++f)
      res.v[f] = q.v[f] + slope.v[f] * sign * 0.5;
    break;
  }
  case FV2D_PCM_WB:
  default:
    res = q;
  }
  return swap_component(res, dir);
}

/* Update.h:104-169 (lambda updateAlongDir) */
static inline void update_along_dir(const fv2d_device_params *p, const double *Q, const double *slopesX,
                                    const double *slopesY, double *Unew, double dt, int i, int j, int dir)
{
  const double *slopes = (dir == IX ? slopesX : slopesY);
  int dxm = (dir == IX ? -1 : 0), dxp = (dir == IX ? 1 : 0);
  int dym = (dir == IY ? -1 : 0), dyp = (dir == IY ? 1 : 0);

  State qCL = reconstruct(p, Q, slopes, i, j, -1.0, dir);
  State qCR = reconstruct(p, Q, slopes, i, j, 1.0, dir);
  State qL  = reconstruct(p, Q, slopes, i + dxm, j + dym, 1.0, dir);
  State qR  = reconstruct(p, Q, slopes, i + dxp, j + dyp, -1.0, dir);

  const double gdx = (dir == IX ? p->gx * p->dx : p->gy * p->dy);

  State fluxL, fluxR;
  double poutL, poutR;
  riemann(p, p->riemann_solver, qL, qCL, gdx, &fluxL, &poutL);
  riemann(p, p->riemann_solver, qCR, qR, gdx, &fluxR, &poutR);

  fluxL = swap_component(fluxL, dir);
  fluxR = swap_component(fluxR, dir);

  /* Update.h:148-156 */
  if (p->well_balanced_flux_at_y_bc && (j == p->jbeg || j == p->jend - 1) && dir == IY)
  {
    double g = get_gravity(p, i, j, dir);
    if (j == p->jbeg)
    {
      State f = {{0.0, 0.0, poutR - Q[idx(p, IR, i, j)] * g * p->dy, 0.0}};
      fluxL   = f;
    }
    else
    {
      State f = {{0.0, 0.0, poutL + Q[idx(p, IR, i, j)] * g * p->dy, 0.0}};
      fluxR   = f;
    }
  }

  /* Update.h:158-159: un_loc += dt * (fluxL - fluxR) / dx  ==  ((fL - fR) * dt) / dx */
  State un            = get_state(p, Unew, i, j);
  const double delta  = (dir == IX ? p->dx : p->dy);
  for (int f = 0; f < 4; ++f)
    un.v[f] += ((fluxL.v[f] - fluxR.v[f]) * dt) / delta;

  /* Update.h:161-166: both sweeps add into IV (Q4) */
  if (p->gravity_mode != FV2D_GRAV_NONE)
  {
    double g = get_gravity(p, i, j, dir);
    un.v[IV] += dt * Q[idx(p, IR, i, j)] * g;
    un.v[IE] += dt * 0.5 * (fluxL.v[IR] + fluxR.v[IR]) * g;
  }

  set_state(p, Unew, i, j, un);
}

/* Update.h:93-174 */
void fv2d_oracle_compute_fluxes_and_update(const fv2d_device_params *p, const double *Q, const double *slopesX,
                                           const double *slopesY, double *Unew, double dt)
{
#pragma omp parallel for schedule(static)
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = p->ibeg; i < p->iend; ++i)
    {
      update_along_dir(p, Q, slopesX, slopesY, Unew, dt, i, j, IX);
      update_along_dir(p, Q, slopesX, slopesY, Unew, dt, i, j, IY);
    }
}

/* ---------------------------------------------------------------- thermal conduction */

/* ThermalConduction.h:36-108.  computeKappa is called with (x -/+ dx, y) doubles narrowed
 * to (int,int) (Q7b): irrelevant for TCM_CONSTANT, which is the only supported mode. */
int fv2d_oracle_apply_thermal_conduction(const fv2d_device_params *p, const double *Q, double *Unew, double dt)
{
  if (p->thermal_conductivity_mode != FV2D_TCM_CONSTANT)
    return 1;
  const double dx = p->dx, dy = p->dy;
#pragma omp parallel for schedule(static)
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = p->ibeg; i < p->iend; ++i)
    {
      double kappaL = 0.5 * (compute_kappa(p) + compute_kappa(p));
      double kappaR = 0.5 * (compute_kappa(p) + compute_kappa(p));
      double kappaU = 0.5 * (compute_kappa(p) + compute_kappa(p));
      double kappaD = 0.5 * (compute_kappa(p) + compute_kappa(p));

      double TC = Q[idx(p, IP, i, j)] / Q[idx(p, IR, i, j)];
      double TL = Q[idx(p, IP, i - 1, j)] / Q[idx(p, IR, i - 1, j)];
      double TR = Q[idx(p, IP, i + 1, j)] / Q[idx(p, IR, i + 1, j)];
      double TU = Q[idx(p, IP, i, j - 1)] / Q[idx(p, IR, i, j - 1)];
      double TD = Q[idx(p, IP, i, j + 1)] / Q[idx(p, IR, i, j + 1)];

      double FL = kappaL * (TC - TL) / dx;
      double FR = kappaR * (TR - TC) / dx;
      double FU = kappaU * (TC - TU) / dy;
      double FD = kappaD * (TD - TC) / dy;

      /* ThermalConduction.h:75-103: the y-boundary overrides replace FL / FR (Q7a) */
      if (j == p->jbeg && p->bctc_ymin != FV2D_BCTC_NONE)
      {
        switch (p->bctc_ymin)
        {
        case FV2D_BCTC_FIXED_TEMPERATURE:
          FL = kappaL * 2.0 * (TC - p->bctc_ymin_value) / dy;
          break;
        case FV2D_BCTC_FIXED_GRADIENT:
          FL = kappaL * p->bctc_ymin_value;
          break;
        default:
          break;
        }
      }
      if (j == p->jend - 1 && p->bctc_ymax != FV2D_BCTC_NONE)
      {
        switch (p->bctc_ymax)
        {
        case FV2D_BCTC_FIXED_TEMPERATURE:
          FR = kappaR * 2.0 * (p->bctc_ymax_value - TC) / dy;
          break;
        case FV2D_BCTC_FIXED_GRADIENT:
          FR = kappaR * p->bctc_ymax_value;
          break;
        default:
          break;
        }
      }

      Unew[idx(p, IE, i, j)] += dt / dx * (FR - FL) + dt / dy * (FD - FU);
    }
  return 0;
}

/* ---------------------------------------------------------------- viscosity */

/* Viscosity.h:27-119.  stencil[dj+1][di+1] = Q(i+di, j+dj) (Viscosity.h:45-50); the viscous
 * fluxes are NOT divided by the cell size (Q8). */
void fv2d_oracle_apply_viscosity(const fv2d_device_params *p, const double *Q, double *Unew, double dt)
{
  const double dx = p->dx, dy = p->dy;
  const double four_thirds = 4.0 / 3.0;
  const double two_thirds  = 2.0 / 3.0;
#pragma omp parallel for schedule(static)
  for (int j = p->jbeg; j < p->jend; ++j)
    for (int i = p->ibeg; i < p->iend; ++i)
    {
      State st[3][3];
      for (int di = -1; di < 2; ++di)
        for (int dj = -1; dj < 2; ++dj)
          st[dj + 1][di + 1] = get_state(p, Q, i + di, j + dj);

      const double one_over_dx = 1.0 / dx;
      const double one_over_dy = 1.0 / dy;
      const double mu          = compute_mu(p);

      State vf[2];
      for (int dir = 0; dir < 2; ++dir)
      {
        State flux = {{0.0, 0.0, 0.0, 0.0}};
        for (int side = 1; side < 3; ++side)
        {
          double sign = (side == 1 ? -1.0 : 1.0);
          if (dir == IX)
          {
            double qiU = 0.5 * (st[1][side].v[IU] + st[1][side - 1].v[IU]);
            double qiV = 0.5 * (st[1][side].v[IV] + st[1][side - 1].v[IV]);

            double dudx = one_over_dx * (st[1][side].v[IU] - st[1][side - 1].v[IU]);
            double dvdx = one_over_dx * (st[1][side].v[IV] - st[1][side - 1].v[IV]);
            double dudy = 0.25 * one_over_dy *
                          (st[2][side].v[IU] - st[0][side].v[IU] + st[2][side - 1].v[IU] - st[0][side - 1].v[IU]);
            double dvdy = 0.25 * one_over_dy *
                          (st[2][side].v[IV] - st[0][side].v[IV] + st[2][side - 1].v[IV] - st[0][side - 1].v[IV]);

            const double tau_xx = four_thirds * dudx - two_thirds * dvdy;
            const double tau_xy = dvdx + dudy;

            flux.v[IU] += sign * mu * tau_xx;
            flux.v[IV] += sign * mu * tau_xy;
            flux.v[IE] += sign * mu * (tau_xx * qiU + tau_xy * qiV);
          }
          else
          {
            double qiU = 0.5 * (st[side][1].v[IU] + st[side - 1][1].v[IU]);
            double qiV = 0.5 * (st[side][1].v[IV] + st[side - 1][1].v[IV]);

            double dudy = one_over_dy * (st[side][1].v[IU] - st[side - 1][1].v[IU]);
            double dvdy = one_over_dy * (st[side][1].v[IV] - st[side - 1][1].v[IV]);
            double dudx = 0.25 * one_over_dx *
                          (st[side][2].v[IU] - st[side][0].v[IU] + st[side - 1][2].v[IU] - st[side - 1][0].v[IU]);
            double dvdx = 0.25 * one_over_dx *
                          (st[side][2].v[IV] - st[side][0].v[IV] + st[side - 1][2].v[IV] - st[side - 1][0].v[IV]);

            const double tau_yy = four_thirds * dvdy - two_thirds * dudx;
            const double tau_xy = dvdx + dudy;

            flux.v[IU] += sign * mu * tau_xy;
            flux.v[IV] += sign * mu * tau_yy;
            flux.v[IE] += sign * mu * (tau_xy * qiU + tau_yy * qiV);
          }
        }
        vf[dir] = flux;
      }

      /* Viscosity.h:115-117: un_loc += dt * (vf_x + vf_y) */
      State un = get_state(p, Unew, i, j);
      for (int f = 0; f < 4; ++f)
        un.v[f] += (vf[0].v[f] + vf[1].v[f]) * dt;
      set_state(p, Unew, i, j, un);
    }
}

/* ---------------------------------------------------------------- step drivers */

/* Update.h:176-191 */
int fv2d_oracle_euler_step(const fv2d_device_params *p, double *Q, double *Unew, double dt)
{
  const size_t n = (size_t)4 * p->Ntx * p->Nty;
  fv2d_oracle_fill_boundaries(p, Q);

  double *sx = NULL, *sy = NULL;
  /* the reference allocates zero-filled slope arrays once (Update.h:54-55) */
  sx = (double *)calloc(n, sizeof(double));
  sy = (double *)calloc(n, sizeof(double));
  if (!sx || !sy)
  {
    free(sx);
    free(sy);
    return -1;
  }
  if (p->reconstruction == FV2D_PLM)
    fv2d_oracle_compute_slopes(p, Q, sx, sy);
  fv2d_oracle_compute_fluxes_and_update(p, Q, sx, sy, Unew, dt);
  free(sx);
  free(sy);

  int rc = 0;
  if (p->thermal_conductivity_active)
    rc = fv2d_oracle_apply_thermal_conduction(p, Q, Unew, dt);
  if (p->viscosity_active)
    fv2d_oracle_apply_viscosity(p, Q, Unew, dt);
  return rc;
}

/* Update.h:193-222 */
int fv2d_oracle_update(const fv2d_device_params *p, int time_stepping, double *Q, double *Unew, double dt)
{
  if (time_stepping == FV2D_TS_EULER)
    return fv2d_oracle_euler_step(p, Q, Unew, dt);
  if (time_stepping == FV2D_TS_RK2)
  {
    const size_t n = (size_t)4 * p->Ntx * p->Nty;
    double *U0     = (double *)malloc(n * sizeof(double));
    double *Ustar  = (double *)malloc(n * sizeof(double));
    if (!U0 || !Ustar)
    {
      free(U0);
      free(Ustar);
      return -1;
    }
    memcpy(U0, Unew, n * sizeof(double));
    memcpy(Ustar, Unew, n * sizeof(double));
    int rc = fv2d_oracle_euler_step(p, Q, Ustar, dt);

    memcpy(Unew, Ustar, n * sizeof(double));
    fv2d_oracle_cons_to_prim(p, Ustar, Q);
    rc |= fv2d_oracle_euler_step(p, Q, Unew, dt);

#pragma omp parallel for schedule(static)
    for (int j = p->jbeg; j < p->jend; ++j)
      for (int i = p->ibeg; i < p->iend; ++i)
        for (int f = 0; f < 4; ++f)
          Unew[idx(p, f, i, j)] = 0.5 * (U0[idx(p, f, i, j)] + Unew[idx(p, f, i, j)]);
    free(U0);
    free(Ustar);
    return rc;
  }
  return 0; /* the reference silently does nothing for other values */
}

/* main.cpp:62-84 without the IO */
long fv2d_oracle_run(const fv2d_device_params *p, int time_stepping, double eps_reset, double tend, double *Q,
                     double *U, long max_steps, double *t_inout, double *dts, uint64_t neg_counts[3])
{
  double t  = *t_inout;
  long step = 0;
  if (neg_counts)
    neg_counts[0] = neg_counts[1] = neg_counts[2] = 0;
  while (t + p->epsilon < tend && step < max_steps)
  {
    double dt = fv2d_oracle_compute_dt(p, Q, NULL);
    if (dts)
      dts[step] = dt;
    if (fv2d_oracle_update(p, time_stepping, Q, U, dt) != 0)
      return -1;
    fv2d_oracle_cons_to_prim(p, U, Q);
    uint64_t c[3];
    fv2d_oracle_check_negatives(p, eps_reset, Q, c);
    if (neg_counts)
    {
      neg_counts[0] += c[0];
      neg_counts[1] += c[1];
      neg_counts[2] += c[2];
    }
    t += dt;
    ++step;
  }
  *t_inout = t;
  return step;
}
