// fv2d_physics.cuh — per-cell physics as device functions, written in the SAME order of
// arithmetic as the reference's inline functions so that the operator-level kernels
// (fv2d_ops.cu, compiled with --fmad=false) reproduce the reference build bit for bit.
// The fused sweep kernel (fv2d_sweep.cu) has its own, re-associated formulation.
#pragma once

#include "fv2d_common.cuh"

namespace fv2d
{

struct State
{
  double v[4];
};

#define IR FV2D_IR
#define IU FV2D_IU
#define IV FV2D_IV
#define IP FV2D_IP
#define IE FV2D_IE

// getStateFromArray / setStateInArray (reference States.h:6-17) on the SoA layout
__device__ __forceinline__ State load_state(const double *__restrict__ A, const Layout &L, int i, int j)
{
  State s;
  const long long o = L.at(0, i, j);
#pragma unroll
  for (int f = 0; f < 4; ++f)
    s.v[f] = A[o + f * L.plane];
  return s;
}
__device__ __forceinline__ void store_state(double *__restrict__ A, const Layout &L, int i, int j, const State &s)
{
  const long long o = L.at(0, i, j);
#pragma unroll
  for (int f = 0; f < 4; ++f)
    A[o + f * L.plane] = s.v[f];
}

// reference States.h:19-30
__device__ __forceinline__ State prim_to_cons(const State &q, double gamma0)
{
  State r;
  r.v[IR]   = q.v[IR];
  r.v[IU]   = q.v[IR] * q.v[IU];
  r.v[IV]   = q.v[IR] * q.v[IV];
  double Ek = 0.5 * (r.v[IU] * r.v[IU] + r.v[IV] * r.v[IV]) / q.v[IR];
  r.v[IE]   = (Ek + q.v[IP] / (gamma0 - 1.0));
  return r;
}

// reference States.h:32-43
__device__ __forceinline__ State cons_to_prim(const State &u, double gamma0)
{
  State r;
  r.v[IR]   = u.v[IR];
  r.v[IU]   = u.v[IU] / u.v[IR];
  r.v[IV]   = u.v[IV] / u.v[IR];
  double Ek = 0.5 * r.v[IR] * (r.v[IU] * r.v[IU] + r.v[IV] * r.v[IV]);
  r.v[IP]   = (u.v[IE] - Ek) * (gamma0 - 1.0);
  return r;
}

// reference States.h:45-46
__device__ __forceinline__ double speed_of_sound(const State &q, double gamma0) { return sqrt(q.v[IP] * gamma0 / q.v[IR]); }

// reference States.h:103-110
__device__ __forceinline__ State swap_component(const State &q, int dir)
{
  if (dir == FV2D_IX)
    return q;
  State r;
  r.v[IR] = q.v[IR];
  r.v[IU] = q.v[IV];
  r.v[IV] = q.v[IU];
  r.v[IP] = q.v[IP];
  return r;
}

// reference RiemannSolvers.h:18-28
__device__ __forceinline__ State hll_phys_flux(const State &q, double gamma0)
{
  const double Ek = 0.5 * q.v[IR] * (q.v[IU] * q.v[IU] + q.v[IV] * q.v[IV]);
  const double E  = (q.v[IP] / (gamma0 - 1.0) + Ek);
  State f;
  f.v[IR] = q.v[IR] * q.v[IU];
  f.v[IU] = q.v[IR] * q.v[IU] * q.v[IU] + q.v[IP];
  f.v[IV] = q.v[IR] * q.v[IU] * q.v[IV];
  f.v[IE] = (q.v[IP] + E) * q.v[IU];
  return f;
}

// reference RiemannSolvers.h:7-51
__device__ __forceinline__ void hll(const State &qL, const State &qR, State &flux, double &pout, double gamma0)
{
  const double aL = speed_of_sound(qL, gamma0);
  const double aR = speed_of_sound(qR, gamma0);

  const double sminL = qL.v[IU] - aL;
  const double smaxL = qL.v[IU] + aL;
  const double sminR = qR.v[IU] - aR;
  const double smaxR = qR.v[IU] + aR;

  const double SL = fmin(sminL, sminR);
  const double SR = fmax(smaxL, smaxR);

  State FL = hll_phys_flux(qL, gamma0);
  State FR = hll_phys_flux(qR, gamma0);

  if (SL >= 0.0)
  {
    flux = FL;
    pout = qL.v[IP];
  }
  else if (SR <= 0.0)
  {
    flux = FR;
    pout = qR.v[IP];
  }
  else
  {
    State uL          = prim_to_cons(qL, gamma0);
    State uR          = prim_to_cons(qR, gamma0);
    pout              = 0.5 * (qL.v[IP] + qR.v[IP]);
    const double SLSR = SL * SR;
#pragma unroll
    for (int f = 0; f < 4; ++f)
      flux.v[f] = ((FL.v[f] * SR - FR.v[f] * SL) + (uR.v[f] - uL.v[f]) * SLSR) / (SR - SL);
  }
}

// reference RiemannSolvers.h:53-128
__device__ __forceinline__ void hllc(const State &qL, const State &qR, State &flux, double &pout, double gamma0)
{
  const double rL = qL.v[IR], uL = qL.v[IU], vL = qL.v[IV], pL = qL.v[IP];
  const double rR = qR.v[IR], uR = qR.v[IU], vR = qR.v[IV], pR = qR.v[IP];

  const double entho = 1.0 / (gamma0 - 1.0);

  const double ekL = 0.5 * rL * (uL * uL + vL * vL);
  const double EL  = ekL + pL * entho;
  const double ekR = 0.5 * rR * (uR * uR + vR * vR);
  const double ER  = ekR + pR * entho;

  const double cfastL = speed_of_sound(qL, gamma0);
  const double cfastR = speed_of_sound(qR, gamma0);

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
    st   = qL;
    E    = EL;
    pout = pL;
  }
  else if (uS > 0.0)
  {
    st.v[IR] = rSL;
    st.v[IU] = uS;
    st.v[IV] = qL.v[IV];
    st.v[IP] = pS;
    E        = ESL;
    pout     = pS;
  }
  else if (SR > 0.0)
  {
    st.v[IR] = rSR;
    st.v[IU] = uS;
    st.v[IV] = qR.v[IV];
    st.v[IP] = pS;
    E        = ESR;
    pout     = pS;
  }
  else
  {
    st   = qR;
    E    = ER;
    pout = pR;
  }

  flux.v[IR] = st.v[IR] * st.v[IU];
  flux.v[IU] = st.v[IR] * st.v[IU] * st.v[IU] + st.v[IP];
  flux.v[IV] = flux.v[IR] * st.v[IV];
  flux.v[IE] = (E + st.v[IP]) * st.v[IU];
}

// reference RiemannSolvers.h:137-171
__device__ __forceinline__ void fslp(const State &qL, const State &qR, State &flux, double &pout, double gdx,
                                     double gamma0, double fslp_K)
{
  const double rhoL = qL.v[IR], uL = qL.v[IU], pL = qL.v[IP];
  const double csL  = sqrt(gamma0 * pL / rhoL);
  const double rhoR = qR.v[IR], uR = qR.v[IU], pR = qR.v[IP];
  const double csR  = sqrt(gamma0 * pR / rhoR);

  const double a1 = rhoL * csL, a2 = rhoR * csR;
  const double ai = fslp_K * (a1 < a2 ? a2 : a1);
  const double m1 = fabs(uL) / csL, m2 = fabs(uR) / csR;
  const double mm = (m1 < m2 ? m2 : m1);
  const double theta = (mm < 1.0 ? mm : 1.0);

  const double ustar = 0.5 * (uR + uL) - 0.5 / ai * (pR - pL - 0.5 * (rhoL + rhoR) * gdx);
  const double Pi    = 0.5 * (pR + pL) - theta * 0.5 * ai * (uR - uL);

  const State &qs     = (ustar > 0 ? qL : qR);
  const double Ekstar = 0.5 * qs.v[IR] * (qs.v[IU] * qs.v[IU] + qs.v[IV] * qs.v[IV]);
  const double Estar  = Ekstar + qs.v[IP] / (gamma0 - 1.0);

  flux.v[IR] = ustar * qs.v[IR];
  flux.v[IU] = ustar * qs.v[IR] * qs.v[IU] + Pi;
  flux.v[IV] = ustar * qs.v[IR] * qs.v[IV];
  flux.v[IE] = ustar * (Estar + Pi);
  pout       = Pi;
}

// lambda `riemann`, reference Update.h:121-135
__device__ __forceinline__ void riemann(int solver, const State &qL, const State &qR, double gdx, State &flux,
                                        double &pout, const fv2d_device_params &p)
{
  switch (solver)
  {
  case FV2D_HLL:
    hll(qL, qR, flux, pout, p.gamma0);
    break;
  case FV2D_FSLP:
    fslp(qL, qR, flux, pout, gdx, p.gamma0, p.fslp_K);
    break;
  default:
    hllc(qL, qR, flux, pout, p.gamma0);
    break;
  }
}

// getGravity (reference Gravity.h:38-57).  The value is float-valued (Q5): gx, gy already
// are (Q1); the analytical profile is evaluated on the host with glibc sin(), narrowed to
// float, and tabulated per row (it depends on y only and ignores `dir`, Gravity.h:15-29).
__device__ __forceinline__ double get_gravity(const KParams &kp, int j, int dir)
{
  switch (kp.p.gravity_mode)
  {
  case FV2D_GRAV_CONSTANT:
    return (dir == FV2D_IX ? kp.p.gx : kp.p.gy);
  case FV2D_GRAV_ANALYTICAL:
    return kp.gtab[j];
  default:
    return 0.0;
  }
}

// minmod limiter, reference Update.h:69-85
__device__ __forceinline__ double minmod(double dL, double dR)
{
  if (dL * dR < 0.0)
    return 0.0;
  else if (fabs(dL) < fabs(dR))
    return dL;
  else
    return dR;
}

} // namespace fv2d
