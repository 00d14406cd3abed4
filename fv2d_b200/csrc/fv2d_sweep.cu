// fv2d_sweep.cu — the fused hot path: one sm_100a kernel per Runge-Kutta stage.
//
// Replaces, in ONE pass over memory, the reference's per-stage kernel chain
//   computeSlopes (Update.h:59-91) -> computeFluxesAndUpdate (Update.h:93-174)
//   -> applyThermalConduction (ThermalConduction.h:36-108) -> applyViscosity
//   (Viscosity.h:27-119) -> [RK2 correct, Update.h:214-220] -> consToPrim
//   (SimInfo.h:576-587) -> checkNegatives (SimInfo.h:602-646) -> computeDt of the new
//   state (ComputeDt.h:18-65)
// Algorithmic traffic: read Q^n + U^n, write U^{n+1} + Q^{n+1} = 128 B per cell.
//
// Work decomposition ("column sweep"): the domain is cut into strips of W = NT-4 columns
// and runs of rows; a work item is one (strip, row run).  The kernel is persistent: at most two
// CTAs of NT threads per SM pull work items from a device-wide queue until it is empty.  Thread t
// owns column i0-2+t of the item's strip (2 halo columns each side) and marches through the rows:
//   * Q rows (primitive SoA tile row + its 2-cell x halo, 4 fields) and U rows (the strip's own
//     cells) are staged into two shared-memory rings by TMA (one cp.async.bulk.tensor.3d box per
//     row and ring, completion on mbarriers), several rows ahead, so HBM latency is hidden
//     without spending registers or address arithmetic on in-flight loads;
//   * the y-direction stencil lives in registers (each thread keeps the rolling
//     q(j), q(j+1), the reconstructed +y face state and the y-face flux of its column): every
//     y-face flux is computed exactly once;
//   * the x-direction needs the neighbour columns: slopes read q(i+-1) from the ring, the
//     reconstructed +x face state (and its sound speed) and the x-face flux are exchanged
//     through small double-buffered shared arrays with ONE __syncthreads per row: every
//     x-face flux is computed exactly once (by the thread on its right);
//   * each slope, each face state, each sound speed is computed once per cell; divisions
//     and square roots are MUFU seeds + one third-order correction (no IEEE slow paths);
//   * conduction and viscosity are evaluated face by face and folded into the face fluxes;
//   * the epilogue of a row writes U^{n+1}, converts to primitives, applies the
//     negative-density/pressure reset, accumulates the CFL maximum, writes Q^{n+1}.
// The per-CTA CFL maximum goes to a device scalar with one atomicMax: the next dt never
// leaves the GPU.  The sweep is the ONLY launch of a stage: this step's dt is taken from the
// device-resident CFL maxima by every CTA at its start, the ghost cells of Q^{n+1} (boundary
// conditions, BoundaryConditions.h:82-147, or the neighbour slab's halo rows) are written by the
// epilogue of the rows they mirror, and the clock (t += dt) is advanced by the sweep's last CTA.
#include "fv2d_kernels.h"
#include "fv2d_fastmath.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace fv2d
{

// --------------------------------------------------------------------------- PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// expect_tx / TMA load (3-D tensor map: column, row, field -> shared, completion on an mbarrier)
// on precomputed shared-window addresses (keeps the producer's per-row work short)
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d_u32(uint32_t dst, const CUtensorMap *tmap, int x, int y, int z, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
               : "memory");
}

// TMA store (shared -> 3-D tensor map, bulk async-group completion) and its bookkeeping
__device__ __forceinline__ void tma_store_3d_u32(const CUtensorMap *tmap, uint32_t src, int x, int y, int z)
{
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(src), "r"(x),
               "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// makes this thread's earlier shared-memory writes visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte asynchronous copy global -> shared (a work-table entry), and its completion wait
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// orders earlier generic-proxy accesses (here: the acquire of a neighbour's halo counter) before
// later async-proxy accesses (the TMA loads of the rows it guards)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Source index of ghost index k for boundary type bc (domain [beg, end), N = end - beg):
// BoundaryConditions.h:25-38 (reflecting), :53-68 (periodic), :94-95 / :124-125 (absorbing).
__device__ __forceinline__ int bc_src(int bc, int k, int beg, int end, int N)
{
  switch (bc)
  {
  case FV2D_BC_REFLECTING:
    return 2 * (k < beg ? beg : end) - k - 1;
  case FV2D_BC_PERIODIC:
    return k < beg ? k + N : k - N;
  default:
    return k < beg ? beg : end - 1;
  }
}

// Development knobs for the A/B variants built by scripts/build_variant.sh (defaults = shipped).
#ifndef FV2D_PS_SHORT
#define FV2D_PS_SHORT 1
#endif
#ifndef FV2D_TMA_STORE_U
#define FV2D_TMA_STORE_U 1 // U^{n+1} rows leave through the ring slot U^n arrived in, by TMA store
#endif
#ifndef FV2D_HLLC_UNIFORM
#define FV2D_HLLC_UNIFORM 0 // 1: select-free HLLC tail when a whole warp takes the same branch (measured slower)
#endif

// max / min as one DSETP + two FSEL.  fmax()/fmin() lower to ~8 instructions each on sm_100a
// (NaN-propagation fix-ups); the operands here are never NaN unless the state already is, and
// then the NaN is counted by the epilogue either way.
__device__ __forceinline__ double dmax(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return (a < b) ? a : b; }

// minmod (Update.h:69-85), branch-free.  "dL*dR < 0" is evaluated as "the sign bits differ"
// on the integer pipe (one LOP3 + ISETP instead of a DMUL + DSETP on the half-rate fp64 pipe);
// the two tests disagree only when a difference is exactly zero (the result is a zero either
// way) or when the product underflows (|slope| < 1e-154).
// The zero of an extremum clears only the HIGH word of the smaller difference (one select instead
// of two): what is left is a denormal below 2^-1042, which the face-state fma (q +- s/2) rounds away
// for every q that is not itself zero or denormal - and for q = 0 it adds less than 1e-314.
#ifndef FV2D_MINMOD_HI
#define FV2D_MINMOD_HI 1
#endif
__device__ __forceinline__ double minmod_f(double dL, double dR)
{
  const double r = (fabs(dL) < fabs(dR)) ? dL : dR;
  const bool opp = (__double2hiint(dL) ^ __double2hiint(dR)) < 0;
#if FV2D_MINMOD_HI
  return __hiloint2double(opp ? 0 : __double2hiint(r), __double2loint(r));
#else
  return opp ? 0.0 : r;
#endif
}

// A face state in the frame of the face normal: n = normal velocity, t = tangential.
struct FaceState
{
  double r, n, t, p, c;
};
// A face flux in the same frame: mass, normal momentum, tangential momentum, energy.
struct FaceFlux
{
  double m, n, t, e, pout;
};

// Flux of the HLLC star state on the K side (RiemannSolvers.h:85-89 + :124-127): K = the left or
// the right face state, SK its wave speed, (uS, pS) the contact speed and pressure.  Written with
// explicit fma / mul intrinsics so that every call site compiles to the same arithmetic.
__device__ __forceinline__ FaceFlux hllc_star(const FaceState &K, double SK, double uS, double pS, double entho)
{
  const double EK = __fma_rn(__dmul_rn(0.5, K.r), __fma_rn(K.t, K.t, __dmul_rn(K.n, K.n)), __dmul_rn(K.p, entho));
  const double d  = frcp(__dsub_rn(SK, uS));
  const double w  = __dsub_rn(SK, K.n);
  const double rS = __dmul_rn(__dmul_rn(K.r, w), d);
  const double ES = __dmul_rn(__fma_rn(pS, uS, __fma_rn(w, EK, -__dmul_rn(K.p, K.n))), d);
  FaceFlux f;
  f.m    = __dmul_rn(rS, uS);
  f.n    = __fma_rn(f.m, uS, pS);
  f.t    = __dmul_rn(f.m, K.t);
  f.e    = __dmul_rn(__dadd_rn(ES, pS), uS);
  f.pout = pS;
  return f;
}

// HLLC (RiemannSolvers.h:53-128), re-associated: one reciprocal for 1/(rcL+rcR), one for
// the star state of the side that is actually taken; same branch structure:
//   SL > 0 -> left state; else uS > 0 -> left star; else SR > 0 -> right star; else right.
//   With FACEC == false the face states carry no sound speed: only max(cL, cR) enters the
//   wave speeds (RiemannSolvers.h:76-77) and cL > cR <=> pL rhoR > pR rhoL, so ONE square root
//   per face (of the faster side) replaces one per face state.
template <bool FACEC>
__device__ __forceinline__ FaceFlux hllc_f(const FaceState &L, const FaceState &R, double entho, double gamma)
{
  const bool lt = L.n < R.n; // one compare serves min and max
  double SL, SR;
  if constexpr (FACEC)
  {
    const double cmax = dmax(L.c, R.c);
    SL                = (lt ? L.n : R.n) - cmax;
    SR                = (lt ? R.n : L.n) + cmax;
  }
  else
  {
    // c = gp * x enters only as n -+ c: the product is folded into the two fmas
    const bool lfast = L.p * R.r > R.p * L.r;
    const double gp  = gamma * (lfast ? L.p : R.p);
    const double x   = csound_over_gp(gp, lfast ? L.r : R.r);
    SL               = fma(-gp, x, lt ? L.n : R.n);
    SR               = fma(gp, x, lt ? R.n : L.n);
  }

  const double rcL = L.r * (L.n - SL);
  const double rcR = R.r * (SR - R.n);
  const double inv = frcp(rcR + rcL);
  const double uS  = fma(rcR, R.n, fma(rcL, L.n, L.p - R.p)) * inv;
#if FV2D_PS_SHORT
  // p* = pL + rhoL (SL - uL)(u* - uL)  (Toro 10.36; algebraically the reference's
  // (rcR pL + rcL pR + rcL rcR (uL - uR)) / (rcL + rcR), RiemannSolvers.h:83): 2 instructions
  const double pS = fma(rcL, L.n - uS, L.p);
#else
  const double pS  = (rcR * L.p + rcL * R.p + rcL * rcR * (L.n - R.n)) * inv;
#endif

  // The three wave-speed tests of RiemannSolvers.h:93-122 read sign bits on the integer pipe
  // instead of DSETPs on the fp64 pipe.  They differ from "> 0" only for an exact +0, where
  // the two fluxes on either side of the test coincide (F*K = FK when SK = 0; both star fluxes
  // are (0, p*, 0, 0) when u* = 0).
  const bool SLpos = __double2hiint(SL) >= 0, uSpos = __double2hiint(uS) >= 0, SRpos = __double2hiint(SR) >= 0;
  const bool left = SLpos || uSpos;
  const bool star = left ? !SLpos : SRpos;

#if FV2D_HLLC_UNIFORM
  // Development variant (measured, not shipped): a select-free tail when all 32 columns of a warp
  // take the same branch.  It removes 35 FSELs per cell-update, but the branches keep the x and the
  // y Riemann problem of a row from being scheduled together: fixed-latency "wait" stalls 1.38 -> 2.29
  // per issue, eligible warps 1.20 -> 0.95, the row loop 4 % slower (profiles/README.md, round 2).
  if (__all_sync(0xffffffffu, star))
  {
    const unsigned ml = __ballot_sync(0xffffffffu, left);
    if (ml == 0xffffffffu)
      return hllc_star(L, SL, uS, pS, entho);
    if (ml == 0u)
      return hllc_star(R, SR, uS, pS, entho);
  }
#endif

  const double rK = left ? L.r : R.r;
  const double uK = left ? L.n : R.n;
  const double vK = left ? L.t : R.t;
  const double pK = left ? L.p : R.p;
  const double SK = left ? SL : SR;

  const double EK = fma(0.5 * rK, fma(uK, uK, vK * vK), pK * entho);
  const double d  = frcp(SK - uS);
  const double w  = SK - uK;
  const double rS = rK * w * d;
  const double ES = fma(pS, uS, fma(w, EK, -(pK * uK))) * d;

  const double r = star ? rS : rK;
  const double u = star ? uS : uK;
  const double p = star ? pS : pK;
  const double E = star ? ES : EK;

  FaceFlux f;
  f.m    = r * u;
  f.n    = fma(f.m, u, p);
  f.t    = f.m * vK;
  f.e    = (E + p) * u;
  f.pout = p;
  return f;
}

// HLL with Davis speeds (RiemannSolvers.h:7-51)
__device__ __forceinline__ FaceFlux hll_f(const FaceState &L, const FaceState &R, double entho)
{
  const double SL = dmin(L.n - L.c, R.n - R.c);
  const double SR = dmax(L.n + L.c, R.n + R.c);

  const double mL = L.r * L.n, mR = R.r * R.n;
  const double EL = fma(0.5 * L.r, fma(L.n, L.n, L.t * L.t), L.p * entho);
  const double ER = fma(0.5 * R.r, fma(R.n, R.n, R.t * R.t), R.p * entho);
  const double FLm = mL, FLn = fma(mL, L.n, L.p), FLt = mL * L.t, FLe = (L.p + EL) * L.n;
  const double FRm = mR, FRn = fma(mR, R.n, R.p), FRt = mR * R.t, FRe = (R.p + ER) * R.n;

  const double inv = frcp(SR - SL);
  const double ss  = SL * SR;
  FaceFlux f;
  // (SR FL - SL FR + SL SR (UR - UL)) / (SR - SL), one fma per term
  f.m    = fma(ss, R.r - L.r, fma(SR, FLm, -(SL * FRm))) * inv;
  f.n    = fma(ss, mR - mL, fma(SR, FLn, -(SL * FRn))) * inv;
  f.t    = fma(ss, fma(R.r, R.t, -(L.r * L.t)), fma(SR, FLt, -(SL * FRt))) * inv;
  f.e    = fma(ss, ER - EL, fma(SR, FLe, -(SL * FRe))) * inv;
  f.pout = 0.5 * (L.p + R.p);
  if (SL >= 0.0)
  {
    f.m = FLm, f.n = FLn, f.t = FLt, f.e = FLe, f.pout = L.p;
  }
  else if (SR <= 0.0)
  {
    f.m = FRm, f.n = FRn, f.t = FRt, f.e = FRe, f.pout = R.p;
  }
  return f;
}

// FSLP (RiemannSolvers.h:137-171)
__device__ __forceinline__ FaceFlux fslp_f(const FaceState &L, const FaceState &R, double entho, double gdx, double K)
{
  const double ai    = K * dmax(L.r * L.c, R.r * R.c);
  const double theta = dmin(1.0, dmax(fabs(L.n) * frcp(L.c), fabs(R.n) * frcp(R.c)));
  const double ustar = fma(-0.5 * frcp(ai), fma(-0.5 * (L.r + R.r), gdx, R.p - L.p), 0.5 * (R.n + L.n));
  const double Pi    = fma(-(theta * 0.5 * ai), R.n - L.n, 0.5 * (R.p + L.p));
  const bool up      = ustar > 0.0;
  const double r = up ? L.r : R.r, n = up ? L.n : R.n, t = up ? L.t : R.t, p = up ? L.p : R.p;
  const double E = fma(0.5 * r, fma(n, n, t * t), p * entho);
  FaceFlux f;
  f.m    = ustar * r;
  f.n    = fma(f.m, n, Pi);
  f.t    = f.m * t;
  f.e    = ustar * (E + Pi);
  f.pout = Pi;
  return f;
}

template <int SOLVER, bool FACEC>
__device__ __forceinline__ FaceFlux riemann_f(const FaceState &L, const FaceState &R, double entho, double gamma,
                                              double gdx, double K)
{
#ifdef FV2D_TEST_NOCOMPUTE
  FaceFlux f;
  f.m = L.r + R.r, f.n = L.n + R.n, f.t = L.t + R.t, f.e = L.p + R.p, f.pout = L.p;
  return f;
#endif
  if constexpr (SOLVER == FV2D_HLL)
    return hll_f(L, R, entho);
  else if constexpr (SOLVER == FV2D_FSLP)
    return fslp_f(L, R, entho, gdx, K);
  else
    return hllc_f<FACEC>(L, R, entho, gamma);
}

// Viscous stress flux through ONE face (Viscosity.h:63-107), in the frame of the face normal:
// n / t = velocity components normal / tangential to the face; hi / lo = the two cells across the
// face; *_p / *_m = their neighbours one cell up / down the tangential direction.  Normal
// derivatives are one-sided across the face, tangential ones the 4-point average
// (Viscosity.h:70-77).  The reference evaluates this expression twice per face (as the `side == 2`
// face of one cell and the `side == 1` face of the next, same operands in the same order); here
// every face is done once and the result is folded into the face's Riemann flux.
struct ViscFlux
{
  double n, t, e; // normal momentum, tangential momentum, energy:  tau_nn, tau_nt, tau . q  (x mu later)
};
__device__ __forceinline__ ViscFlux visc_face(double n_hi, double n_lo, double t_hi, double t_lo, double n_hi_p,
                                              double n_hi_m, double n_lo_p, double n_lo_m, double t_hi_p, double t_hi_m,
                                              double t_lo_p, double t_lo_m, double rdn, double rdt)
{
  const double c43 = 4.0 / 3.0, c23 = 2.0 / 3.0;
  const double dndn = rdn * (n_hi - n_lo);
  const double dtdn = rdn * (t_hi - t_lo);
  const double dndt = 0.25 * rdt * (n_hi_p - n_hi_m + n_lo_p - n_lo_m);
  const double dtdt = 0.25 * rdt * (t_hi_p - t_hi_m + t_lo_p - t_lo_m);
  const double tnn  = fma(c43, dndn, -(c23 * dtdt));
  const double tnt  = dtdn + dndt;
  ViscFlux f;
  f.n = tnn;
  f.t = tnt;
  f.e = fma(tnn, 0.5 * (n_hi + n_lo), tnt * (0.5 * (t_hi + t_lo)));
  return f;
}

// --------------------------------------------------------------------------- the kernel

#ifndef FV2D_NS_BIAS
#define FV2D_NS_BIAS 0 // development knob: moves ring slots from the U ring to the Q ring
#endif
#ifndef FV2D_UNROLL
#define FV2D_UNROLL 2 // rows per loop body (2 rows = 19 KB of code, inside the 32 KB L1.5 I-cache)
#endif
#ifndef FV2D_EXTRA_SMEM
#define FV2D_EXTRA_SMEM 0 // development knob: pads the CTA's shared memory to lower the occupancy
#endif
constexpr int kUnroll = FV2D_UNROLL;
#ifndef FV2D_UNROLL_DIFF
#define FV2D_UNROLL_DIFF 3 // ... for the conduction / viscosity variants (measured: C91 -2 % vs 2)
#endif
constexpr int kUnrollDiff = FV2D_UNROLL_DIFF;
// ---- shared-memory budget of a variant.  Two CTAs per SM leave 113 KB each; after the exchange
// arrays (and 384 B of barriers, work queue and broadcast scalars) the rest is cut into 8 KB ring
// slots, shared between the Q ring and the U ring so that both are requested about equally many
// rows ahead of their use.
//   exchange arrays: X2 16 KB; X1 16 KB (PLM only: a PCM face state is the cell state, read from
//   the Q ring itself); face sound speeds X1c 4 KB (all but PLM + HLLC); temperatures X1T 4 KB
//   (conduction / viscosity variants)
#ifndef FV2D_RING_CUT
#define FV2D_RING_CUT 0 // development knob: ring slots given up (what output staging for TMA stores would cost)
#endif
__host__ __device__ constexpr int ring_total(bool plm, bool facec, bool diff)
{
  return (115712 - 384 - 16384 - (plm ? 16384 : 0) - (facec ? 4096 : 0) - (diff ? 4096 : 0)) / 8192 -
         ((plm && !facec && !diff) ? FV2D_RING_CUT : 0);
}
// Newest Q row that no thread reads any more once the row barrier of iteration k is passed,
// relative to k: the viscous x-face flux reads rows k-1 .. k+1 in phase B, gravity reads rho of
// row k in the epilogue; otherwise PLM reads row k+1 last in phase C, PCM row k last in phase B.
__host__ __device__ constexpr int ring_dead(bool plm, int grav, bool diff) { return (diff || grav != 0) ? -1 : (plm ? 1 : 0); }
// Q row k+dead+NS is requested in iteration k and first read in iteration k+dead+NS-2; U row
// k-1+NU is requested in iteration k and read in iteration k-1+NU: balance the two look-aheads.
// (The conduction / viscosity variants measured 1-2 % faster with one slot moved to the U ring.)
__host__ __device__ constexpr int ring_ns(int total, int dead, bool diff) { return (total + 2 - dead) / 2 + FV2D_NS_BIAS - (diff ? 1 : 0); }

template <bool B, int N>
struct dim_if
{
  static constexpr int value = B ? N : 1;
};

#ifdef FV2D_TIMING
__device__ long long g_sweep_timing[4096]; // per CTA: cycles outside the row loops, inside them, items, total
#endif

template <int NT, int kNS, int kNU, bool PLM, bool FACEC, bool DIFF>
struct SweepSmem
{
  double ring[kNS][4][NT];  // TMA destination, Q rows: [slot][field][column incl. 2+2 halo]
  double uring[kNU][4][NT]; // TMA destination, U rows: [slot] then a dense [field][NT-4] box
  // exchange arrays, by row parity; fields are paired so that every access is one conflict-free
  // 128-bit LDS / STS (16-byte stride between neighbouring threads)
  double2 X1a[dim_if<PLM, 2>::value][dim_if<PLM, NT>::value];    // +x face state of each column: (r, n)
  double2 X1b[dim_if<PLM, 2>::value][dim_if<PLM, NT>::value];    //                                (t, p)
  double X1c[dim_if<FACEC, 2>::value][dim_if<FACEC, NT>::value]; //                                 c
  double X1T[dim_if<DIFF, 2>::value][dim_if<DIFF, NT>::value];   // temperature P / rho of each column
  double2 X2a[2][NT];  // x-face flux at the LEFT face of each column: (m, n)
  double2 X2b[2][NT];  //                                              (t, e)
  WorkItem item[2];    // the CTA's current work item (entry ci of its sequence at ci & 1) and the next one
  WorkItem item_in;    // landing slot of the table entry after those (cp.async, thread 0)
  uint64_t full[kNS];  // TMA completion barriers, one per Q ring slot
  uint64_t ufull[kNU]; // ... one per U ring slot
  double dt;           // this step's dt, broadcast by thread 0
  double inv3[3];      // the three inverse time-steps behind it {hyp, tc, visc}
  // the item after the current one as the TMA producer (thread 0) needs it: tensor-map column of its
  // Q rows, its first and last Q row (last < first: there is no next item), its U rows
  int nx_xq, nx_rbase, nx_rlast, nx_j0, nx_j1;
};

// GRAV: 0 = no gravity, 1 = gravity, 2 = gravity + well-balanced flux at the y boundary
// MODE: what kind of stage the launch is, so that its loop-invariant tests are resolved at compile
// time instead of once per row (and their code stays out of the row loop):
//   kPlain   the only stage of a forward-Euler step on a slab without neighbours (the common launch)
//   kMulti   the same on a y-slab with neighbour slabs (peer pushes, halo waits)
//   kGeneral anything: either stage of an SSP-RK2 step, with or without neighbours
//
// Persistent: the grid is min(#work items, 2 x #SMs) CTAs.  A work item is (strip, rows j0..j1-1);
// CTA b starts on item b and then pulls further items from a device-wide counter.  The two TMA
// rings do not know about items: the Q rows (j0-2 .. j1+1) and U rows (j0 .. j1-1) of the CTA's
// items form one stream each, staged kNS / kNU rows ahead of their use ACROSS item boundaries, so a
// new item starts on rows that are already in shared memory and the pipeline is filled once per
// CTA, not once per item.  Items shrink towards the end of the table (guided schedule, see
// fv2d_capi.cu: build_work_items), which evens out the tail.  A CTA knows its current item and the
// next one, no more (deeper reservations tie work to a CTA long before it can start it: measured,
// they unbalance the tail); every item has at least 8 rows, so a stream never runs past the next
// item.  The row loop contains no function call and no table walk: calls inside it made ptxas keep
// the ring bookkeeping in vector registers instead of uniform ones (+70 IMAD per row, spills).
enum : int
{
  kPlain = 0,
  kMulti = 1,
  kGeneral = 2
};
template <int NT, bool PLM, int SOLVER, int GRAV, bool DIFF, int MODE>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 4 : (NT <= 256 ? 2 : 1)))
k_sweep(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmU,
        const __grid_constant__ SweepArgs a)
{
  constexpr int W = NT - 4;
  // Do the face states carry their sound speed?  Not for PLM + HLLC (see hllc_f); PCM has one
  // sound speed per cell shared by its four faces, HLL / FSLP need both sides'.
  constexpr bool FACEC = !(PLM && SOLVER == FV2D_HLLC);
  constexpr int kDead = ring_dead(PLM, GRAV, DIFF);
  constexpr int kUnrollV = DIFF ? kUnrollDiff : kUnroll;
  constexpr int kNS   = ring_ns(ring_total(PLM, FACEC, DIFF), kDead, DIFF);
  constexpr int kNU   = ring_total(PLM, FACEC, DIFF) - kNS;
#if FV2D_RING_CUT == 0
  static_assert(kNS + kDead >= 4 && kNU >= 2, "ring too shallow");
#endif
  using Smem = SweepSmem<NT, kNS, kNU, PLM, FACEC, DIFF>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);

  const fv2d_device_params &p = a.kp.p;
  const Layout &L             = a.kp.L;
  DevScalars *const sc        = a.kp.sc;
  constexpr bool PLAIN        = (MODE == kPlain);
  const double *const U0      = (MODE == kGeneral) ? a.U0 : nullptr;
  const bool final_stage      = (MODE == kGeneral) ? (a.final_stage != 0) : true;
  double *const peer_lo       = PLAIN ? nullptr : a.peer_lo_Qout;
  double *const peer_hi       = PLAIN ? nullptr : a.peer_hi_Qout;
  const int Ng                = p.Ng;

  const int t  = threadIdx.x;
  const int tl = (t > 0 ? t - 1 : 0), tr = (t < NT - 1 ? t + 1 : NT - 1);
  const int tu = min(max(t - 2, 0), W - 1); // this thread's column inside a U box

  // ---- TMA staging (thread 0): Q rows are NT columns x 4 fields, U rows W columns x 4 fields
  constexpr uint32_t kQRowBytes = 4u * NT * sizeof(double);
  constexpr uint32_t kURowBytes = 4u * W * sizeof(double);
  const uint32_t ring0  = smem_u32(&S.ring[0][0][0]);
  const uint32_t urng0  = smem_u32(&S.uring[0][0][0]);
  const uint32_t bar0   = smem_u32(&S.full[0]);
  const uint32_t ubar0  = smem_u32(&S.ufull[0]);
  const int tma_x0      = L.lead + p.ibeg; // tensor-map column of the first domain cell
  auto stage_q = [&](int x, int r, uint32_t slot) {
    mbar_expect_tx_u32(bar0 + 8u * slot, kQRowBytes);
    tma_load_3d_u32(ring0 + kQRowBytes * slot, &tmQ, x, r, 0, bar0 + 8u * slot);
  };
  auto stage_u = [&](int x, int r, uint32_t slot) {
    mbar_expect_tx_u32(ubar0 + 8u * slot, kURowBytes);
    tma_load_3d_u32(urng0 + kQRowBytes * slot, &tmU, x, r, 0, ubar0 + 8u * slot);
  };

  // U^{n+1} of a row is written back into the ring slot its U^n arrived in (same dense [field][W] box,
  // every thread its own element) and leaves through a TMA store one iteration later; the slot is
  // re-armed for the next load one iteration after that, when the store has read it.
  auto store_u = [&](int x, int r, uint32_t slot) {
    tma_store_3d_u32(a.tm_store_u, urng0 + kQRowBytes * slot, x, r, 0);
    bulk_commit();
  };

  // ---- work queue (thread 0).  A CTA always knows its current item and the next one (the rows of
  // the next item are staged while the current one finishes).  The item after those is taken from
  // the device-wide counter as LATE as possible - 8 rows before the current item ends - because an
  // item is bound to the CTA from that moment on, and work bound early cannot be rebalanced at the
  // tail.  Nobody waits for the atomic or for the table read that follows it 5 rows later (cp.async
  // into shared memory): both have landed when the item ends.
  unsigned pend = 0; // requested table index (thread 0)
  // Ghost rows owned by a neighbour slab: its sweep of the previous stage pushed them (plain stores
  // over NVLink).  Before the first row of an item that reads such rows is staged, wait for all of
  // them, then order those writes before the async-proxy reads of TMA.  (After a final-stage sweep
  // the wait never spins: the CFL mail this sweep has already received is posted after the
  // neighbour's last push.  It does between the two stages of an RK2 step.)  Executed by ALL threads
  // of the CTA on a CTA-uniform condition: a spin loop under `if (t == 0)` anywhere inside the item
  // loop makes ptxas move the ring bookkeeping of the whole kernel from uniform to vector registers
  // (measured: +36 IMAD and 19 R2UR per row, the kMulti sweep 3.7 % slower than kPlain).
  auto wait_halo_item = [&](const WorkItem &e) {
#ifndef FV2D_X_NOWAIT
    if constexpr (!PLAIN)
#else
    if constexpr (false)
#endif
    {
      if (e.j0 < 0)
        return;
      const bool lo = e.j0 - 2 < p.jbeg && a.kp.edge_lo == EDGE_NEIGHBOUR;
      const bool hi = e.j1 + 1 >= p.jend && a.kp.edge_hi == EDGE_NEIGHBOUR;
      if (lo)
        wait_ge_sys(&sc->halo_cnt[0], a.halo_expected, sc);
      if (hi)
        wait_ge_sys(&sc->halo_cnt[1], a.halo_expected, sc);
      if (lo || hi)
        fence_proxy_async();
    }
  };
  // Q row `off` (>= 1) rows beyond the last Q row of the current item = row off-1 of the next item
  auto stage_q_next = [&](int off, uint32_t slot) {
    const int r = S.nx_rbase + off - 1;
    if (r <= S.nx_rlast)
      stage_q(S.nx_xq, r, slot);
  };
  // U row `off` (>= 0) rows beyond the last U row of the current item
  auto stage_u_next = [&](int off, uint32_t slot) {
    const int r = S.nx_j0 + off;
    if (r < S.nx_j1)
      stage_u(S.nx_xq + 2, r, slot);
  };
  auto publish_next = [&](const WorkItem &e) { // thread 0: e is the item after the current one
    S.nx_xq    = tma_x0 + e.strip * W - 2;
    S.nx_rbase = e.j0 - 2;
    S.nx_rlast = (e.j0 < 0) ? e.j0 - 3 : e.j1 + 1;
    S.nx_j0    = e.j0;
    S.nx_j1    = e.j1;
  };

  // ---- kernel prologue, part 1 (thread 0): everything that does not depend on the previous launch -
  // barriers and the CTA's first work item (the table is static).  The sweep is launched with
  // programmatic stream serialization: this part runs while the previous sweep's last CTAs finish.
  if (t == 0)
  {
#pragma unroll
    for (int s = 0; s < kNS; ++s)
      mbar_init(&S.full[s], 1);
#pragma unroll
    for (int s = 0; s < kNU; ++s)
      mbar_init(&S.ufull[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    S.item[0] = a.items[blockIdx.x]; // the first item is static: CTA b starts on item b
    S.item_in = WorkItem{0, -1, -1, 0};
  }
  // everything below reads what the previous launch wrote (Q, U, the work counter, the CFL mail)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if constexpr (!PLAIN)
    wait_halo_item(a.items[blockIdx.x]); // the CTA's first item (all threads: see wait_halo_item)

  // dt = CFL / max(inverse time-steps of the current state)   (ComputeDt.h:64): the hyperbolic
  // maximum was mailed to every rank by the last CTA of the previous final-stage sweep.  Executed by
  // thread 0 (host dt; single slab: the device's own mailbox) or by the 32 lanes of warp 1 (y-slabs).
  const bool warp_mail = !PLAIN && a.use_device_dt != 0;
  auto set_dt = [&]() {
    const int lane = t & 31;
    double dt = a.dt_host, hyp = 0.0, tc = p.epsilon, visc = p.epsilon;
    unsigned long long tg0 = 0, tg1 = 0;
    if (blockIdx.x == 0 && lane == 0)
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg0));
    if (a.use_device_dt)
    {
      bool local_mail = false;
      if constexpr (PLAIN)
      {
        // single slab: the previous launch left the maximum in this device's own mailbox (device scope)
        if (a.kp.nranks == 1 && *(volatile unsigned long long *)&sc->mail_gen[0] >= a.mail_gen)
        {
          hyp        = *(volatile double *)&sc->mail_inv[a.mail_gen & 1][0];
          local_mail = true;
        }
        else // (no step / computeDt before this launch: the wait times out into the fault flag)
          hyp = collect_cfl_mail(a.kp, a.mail_gen);
      }
      else
      {
        double m = -1.7976931348623157e308;
        if (lane < a.kp.nranks)
        {
          wait_ge_sys(&sc->mail_gen[lane], a.mail_gen, sc);
          m = ld_relaxed_sys_f64(&sc->mail_inv[a.mail_gen & 1][lane]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        hyp = m;
      }
      if (p.thermal_conductivity_active)
        tc = fmax(2.0 * p.kappa / (p.dx * p.dx), 2.0 * p.kappa / (p.dy * p.dy));
      if (p.viscosity_active)
        visc = fmax(2.0 * p.mu / (p.dx * p.dx), 2.0 * p.mu / (p.dy * p.dy));
      double m = hyp;
      if (m < tc)
        m = tc;
      if (m < visc)
        m = visc;
      dt = p.CFL / m;
      // a wait on a peer timed out: stop advancing (the host reports the fault; only a cross-GPU wait can)
      if (!local_mail && *(volatile unsigned int *)&sc->fault)
        dt = __longlong_as_double(0x7ff8000000000000LL);
    }
    if (lane == 0)
    {
      if (blockIdx.x == 0)
      {
        // how long this rank waited for the other ranks' CFL mails (0 on a single slab): the per-step
        // synchronisation cost of the y-slab decomposition, reported by bench.py
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg1));
        sc->tstamp[0] = tg0, sc->tstamp[1] = tg1, sc->tstamp[3] += tg1 - tg0;
      }
      S.dt      = dt;
      S.inv3[0] = hyp, S.inv3[1] = tc, S.inv3[2] = visc;
    }
  };

  // ---- part 2 (thread 0): the initial fill of both rings FIRST (so the rows travel while the rest is
  // set up), the second work item, and this step's dt
  if (t == 0)
  {
    const WorkItem e0 = S.item[0];
    // initial fill of both rings from the first item (>= 8 rows unless it is the CTA's only one)
    {
      const int x = tma_x0 + e0.strip * W - 2;
#pragma unroll 1
      for (int n = 0; n < kNS && n < e0.j1 - e0.j0 + 4; ++n)
        stage_q(x, e0.j0 - 2 + n, (uint32_t)n);
#pragma unroll 1
      for (int n = 0; n < kNU && n < e0.j1 - e0.j0; ++n)
        stage_u(x + 2, e0.j0 + n, (uint32_t)n);
    }
    // the second item comes from the counter; with a.persistent == 0 (degenerate grids: items of
    // fewer than 8 rows) every CTA has just its one item
    unsigned i1 = (unsigned)a.n_items;
    if (a.persistent)
      i1 = min(gridDim.x + atomicAdd(&sc->work_next, 1u), (unsigned)a.n_items);
    const WorkItem e1 = a.items[i1]; // items[n_items] is the end marker
    S.item[1] = e1;
    publish_next(e1);
    if (!warp_mail)
      set_dt(); // host dt, or a single slab's own mailbox
  }
  // y-slabs: warp 1 collects the ranks' CFL mails - one lane per rank, so the mailboxes are polled side
  // by side (a single thread walking them pays one system-scope round trip per rank) - while thread 0
  // sets up the rings and the work queue
  if constexpr (!PLAIN)
  {
    if (warp_mail && t >= 32 && t < 64)
      set_dt();
  }
  __syncthreads();

  const double dt    = S.dt;
  const double rdx   = 1.0 / p.dx;
  const double rdy   = 1.0 / p.dy;
  const double rdxy  = rdx + rdy;
  const double dtdx  = dt / p.dx;
  const double dtdy  = dt / p.dy;
  const double gamma = p.gamma0;
  const double gm1   = gamma - 1.0;
  const double entho = 1.0 / gm1;
  // conduction / viscosity, done face by face and folded into the Riemann fluxes: the reference's
  //   U += dt (vf_x + vf_y)                    (Viscosity.h:112-116, not divided by the cell size: Q8)
  //   U[IE] += dt/dx (FR - FL) + dt/dy (FD - FU)               (ThermalConduction.h:106)
  // are differences of face quantities, so  f_x -= dx mu tau_x + kappa dT/dx  etc. gives the same
  // update through dt/dx (f_left - f_right).
  // (branch-free: an inactive operator runs with a zero coefficient, which subtracts exact zeros)
  const bool tc_on = DIFF && p.thermal_conductivity_active, visc_on = DIFF && p.viscosity_active;
  const double mudx = visc_on ? p.mu * p.dx : 0.0, mudy = visc_on ? p.mu * p.dy : 0.0;
  const double kaprdx = tc_on ? p.kappa * rdx : 0.0, kaprdy = tc_on ? p.kappa * rdy : 0.0;
  const long long pitchB = (long long)L.pitch * (long long)sizeof(double);
  const long long planeB = L.plane * (long long)sizeof(double);
  const bool fold   = a.fold_ghosts != 0;
  // rows jbeg / jend-1 of the GLOBAL grid carry the well-balanced flux and the conduction boundary
  // values, whatever the boundary type (a periodic ring of slabs has them too)
  const bool glob_lo = holds_global_first_row(a.kp), glob_hi = holds_global_last_row(a.kp);

  double inv_dt_max = -1.7976931348623157e308;
#ifdef FV2D_TIMING
  long long tm_mark = clock64(), tm_gap = 0, tm_loop = 0, tm_items = 0;
  const long long tm_start = tm_mark;
#endif

  // ring bookkeeping, carried from row to row AND from item to item: (s2, ph2) = slot / phase
  // parity of the next Q row of the stream, (us, uph) = ... of the next U row
  int s2 = 0, s1 = 0, s0 = 0, sm1 = 0;
  uint32_t ph2 = 0;
  int us = 0, us_prev = 0, us_prev2 = 0;
  uint32_t uph = 0;
  auto next_q = [&]() {
    s2 = (s2 + 1 == kNS) ? 0 : s2 + 1;
    ph2 ^= (s2 == 0) ? 1u : 0u;
  };

  for (int ci = 0;; ++ci)
  {
    const WorkItem item = S.item[ci & 1];
    if (item.j0 < 0)
      break;
    if constexpr (!PLAIN)
      wait_halo_item(S.item[(ci + 1) & 1]); // the rows of the next item are staged while this one finishes
    const int j0 = item.j0, j1 = item.j1; // rows [j0, j1) are updated
    const int i0    = p.ibeg + item.strip * W; // first interior column of the strip
    const int col   = i0 - 2 + t;              // this thread's column
    const int rbase = j0 - 2;                  // first Q row of the item
    const int rlast = j1 + 1;                  // last Q row of the item
    const int nrow  = j1 - j0;
    const bool interior = (t >= 2) && (t < NT - 2) && (col < p.iend);
    const int tma_xq = tma_x0 + item.strip * W - 2;
    const int tma_xu = tma_xq + 2;
    // ghost cells of Q^{n+1} are written by the sweep that produces it: every ghost is a (sign-
    // flipped) copy of ONE domain cell (BoundaryConditions.h:82-147, x and y passes composed), so
    // the thread that owns the source column also stores the ghost columns mirrored from it ...
    unsigned xmask      = 0; // bit g: ghost column g (0..Ng-1 left of the domain, Ng..2Ng-1 right) copies this column
    const bool xstrip   = fold && (i0 < p.ibeg + Ng || (i0 + W > p.iend - Ng && i0 < p.iend));
    if (xstrip && interior && (col < p.ibeg + Ng || col >= p.iend - Ng))
    {
      for (int g = 0; g < 2 * Ng; ++g)
      {
        const int ig = (g < Ng) ? g : p.iend + g - Ng;
        if (bc_src(p.boundary_x, ig, p.ibeg, p.iend, p.Nx) == col)
          xmask |= 1u << g;
      }
    }
    // ... and the rows within Ng of the slab's edges are also stored into the ghost rows mirrored
    // from them: this slab's own (physical boundary) or the neighbour slab's (peer mapping)
    const bool yitem = (fold || peer_lo != nullptr || peer_hi != nullptr) && (j0 < p.jbeg + Ng || j1 > p.jend - Ng);

    // ---- pre-prologue: rows j0-2, j0-1, j0 of this column (the next three rows of the Q stream)
    double qn[4];  // q(row k+1)
    FaceState yp;  // +y face state of row k (frame of the y normal: n = v, t = u)
    double dyl[4]; // q(k+1) - q(k): the lower y difference of row k+1's slope, carried row to row
    double Tk = 0.0; // temperature P / rho of (col, k) (conduction)
    sm1 = s2;
    mbar_wait(&S.full[s2], ph2);
    next_q();
    s0 = s2;
    mbar_wait(&S.full[s2], ph2);
    next_q();
    s1 = s2;
    mbar_wait(&S.full[s2], ph2);
    next_q();
    {
      double qa[4], qk[4];
#pragma unroll
      for (int f = 0; f < 4; ++f)
      {
        qa[f] = S.ring[sm1][f][t];
        qk[f] = S.ring[s0][f][t];
        qn[f] = S.ring[s1][f][t];
      }
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      if constexpr (PLM)
      {
#pragma unroll
        for (int f = 0; f < 4; ++f)
          s[f] = minmod_f(qk[f] - qa[f], qn[f] - qk[f]);
      }
      yp.r = fma(0.5, s[0], qk[0]);
      yp.t = fma(0.5, s[1], qk[1]);
      yp.n = fma(0.5, s[2], qk[2]);
      yp.p = fma(0.5, s[3], qk[3]);
      yp.c = FACEC ? csound(gamma * yp.p, yp.r) : 0.0;
#pragma unroll
      for (int f = 0; f < 4; ++f)
        dyl[f] = qn[f] - qk[f];
      if constexpr (DIFF)
        Tk = qk[3] * frcp(qk[0]);
    }

    FaceFlux fy_lo;  // y-face flux below row k (becomes the low face of the next row)
    FaceState xm;    // -x face state of row k (own cell, left face), frame of the x normal
    fy_lo.m = fy_lo.n = fy_lo.t = fy_lo.e = fy_lo.pout = 0.0;
    xm.r = xm.n = xm.t = xm.p = xm.c = 0.0;

    // Global addresses = uniform per-(array, field) base + one per-thread byte offset that is
    // carried and bumped by a row pitch per iteration (two integer instructions per address).
    long long offB = (L.at(0, col, 0) + (long long)(j0 - 1) * L.pitch) * (long long)sizeof(double); // row k

    // the U ring: the warm-up iteration consumes no U row, so step back by one position (the roll
    // at the end of every iteration then lands on row j0's slot)
    uph ^= (us == 0) ? 1u : 0u;
    us = (us == 0) ? kNU - 1 : us - 1;

    if constexpr (kDead >= 0)
    {
      // the rows only the pre-prologue needed (rbase .. rbase+kDead) are dead once every thread has
      // read its column: recycle their slots before the march starts
      __syncthreads();
      if (t == 0)
      {
#pragma unroll 1
        for (int n = 0; n <= kDead; ++n)
        {
          const uint32_t slot = (uint32_t)(n == 0 ? sm1 : s0);
          const int r = rbase + kNS + n;
          if (r <= rlast)
            stage_q(tma_xq, r, slot);
          else
            stage_q_next(r - rlast, slot);
        }
      }
    }

#ifdef FV2D_TIMING
    {
      const long long now = clock64();
      tm_gap += now - tm_mark, tm_mark = now, tm_items += 1 + ((long long)nrow << 16);
    }
#endif
    // ---- march: iteration k finishes row k; k = j0-1 is the warm-up (no update)
#pragma unroll kUnrollV
    for (int k = j0 - 1; k < j1; ++k)
    {
      const int par = k & 1;
      // A. new row k+2 enters; y slopes / face states of row k+1; y-face flux at k+1/2
      mbar_wait(&S.full[s2], ph2);
      double qnn[4];
#pragma unroll
      for (int f = 0; f < 4; ++f)
        qnn[f] = S.ring[s2][f][t];

      FaceState ym, yp1;
      {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (PLM)
        {
#pragma unroll
          for (int f = 0; f < 4; ++f)
          {
            const double dup = qnn[f] - qn[f]; // also the lower difference of the next row
            s[f]             = minmod_f(dyl[f], dup);
            dyl[f]           = dup;
          }
        }
        ym.r  = fma(-0.5, s[0], qn[0]);
        ym.t  = fma(-0.5, s[1], qn[1]);
        ym.n  = fma(-0.5, s[2], qn[2]);
        ym.p  = fma(-0.5, s[3], qn[3]);
        yp1.r = fma(0.5, s[0], qn[0]);
        yp1.t = fma(0.5, s[1], qn[1]);
        yp1.n = fma(0.5, s[2], qn[2]);
        yp1.p = fma(0.5, s[3], qn[3]);
        if constexpr (!FACEC)
          ym.c = yp1.c = 0.0;
        else if constexpr (PLM)
        {
          ym.c  = csound(gamma * ym.p, ym.r);
          yp1.c = csound(gamma * yp1.p, yp1.r);
        }
        else
        {
          ym.c  = csound(gamma * ym.p, ym.r);
          yp1.c = ym.c;
        }
      }
      const double gdy = p.gy * p.dy, gdx = p.gx * p.dx;
      FaceFlux fy_hi = riemann_f<SOLVER, FACEC>(yp, ym, entho, gamma, gdy, p.fslp_K);
      // diffusive flux through the same face (between rows k and k+1), to be subtracted
      double dy_t = 0.0, dy_n = 0.0, dy_e = 0.0, Tn = 0.0;
      if constexpr (DIFF)
      {
        Tn   = qn[3] * frcp(qn[0]);
        dy_e = kaprdy * (Tn - Tk); // FD of row k = FU of row k+1 (ThermalConduction.h:66-67)
        {
          const double u_lo = S.ring[s0][1][t], v_lo = S.ring[s0][2][t];
          const ViscFlux vf = visc_face(qn[2], v_lo, qn[1], u_lo,                                        //
                                        S.ring[s1][2][tr], S.ring[s1][2][tl], S.ring[s0][2][tr], S.ring[s0][2][tl], //
                                        S.ring[s1][1][tr], S.ring[s1][1][tl], S.ring[s0][1][tr], S.ring[s0][1][tl], //
                                        rdy, rdx);
          dy_n = mudy * vf.n;
          dy_t = mudy * vf.t;
          dy_e = fma(mudy, vf.e, dy_e);
        }
      }

      // B. x-face flux at the left face of (col, k): left state from the neighbour thread
      //    (also runs, on don't-care data, in the warm-up iteration: no branch, so the x and y
      //    Riemann problems of a row can be scheduled together)
      {
        FaceState xl;
        if constexpr (PLM)
        {
          const double2 la = S.X1a[par][tl], lb = S.X1b[par][tl];
          xl.r = la.x, xl.n = la.y, xl.t = lb.x, xl.p = lb.y;
        }
        else // PCM: the +x face state of the left neighbour is its cell state, still in the ring
          xl.r = S.ring[s0][0][tl], xl.n = S.ring[s0][1][tl], xl.t = S.ring[s0][2][tl], xl.p = S.ring[s0][3][tl];
        if constexpr (FACEC)
          xl.c = S.X1c[par][tl];
        else
          xl.c = 0.0;
        FaceFlux fx = riemann_f<SOLVER, FACEC>(xl, xm, entho, gamma, gdx, p.fslp_K);
        if constexpr (DIFF)
        {
          // diffusive flux through the left x-face of (col, k): cells (col-1, k) and (col, k)
          fx.e = fma(-kaprdx, Tk - S.X1T[par][tl], fx.e); // FL (ThermalConduction.h:64)
          {
            const ViscFlux vf = visc_face(S.ring[s0][1][t], S.ring[s0][1][tl], S.ring[s0][2][t], S.ring[s0][2][tl],          //
                                          S.ring[s1][1][t], S.ring[sm1][1][t], S.ring[s1][1][tl], S.ring[sm1][1][tl], //
                                          S.ring[s1][2][t], S.ring[sm1][2][t], S.ring[s1][2][tl], S.ring[sm1][2][tl], //
                                          rdx, rdy);
            fx.n = fma(-mudx, vf.n, fx.n);
            fx.t = fma(-mudx, vf.t, fx.t);
            fx.e = fma(-mudx, vf.e, fx.e);
          }
        }
        S.X2a[par][t] = make_double2(fx.m, fx.n);
        S.X2b[par][t] = make_double2(fx.t, fx.e);
      }

      // C. x slopes / face states of row k+1 (published for the neighbour on the right)
      FaceState xm1;
      {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (PLM)
        {
#pragma unroll
          for (int f = 0; f < 4; ++f)
            s[f] = minmod_f(qn[f] - S.ring[s1][f][tl], S.ring[s1][f][tr] - qn[f]);
        }
        xm1.r = fma(-0.5, s[0], qn[0]);
        xm1.n = fma(-0.5, s[1], qn[1]);
        xm1.t = fma(-0.5, s[2], qn[2]);
        xm1.p = fma(-0.5, s[3], qn[3]);
        FaceState xp1;
        xp1.r = fma(0.5, s[0], qn[0]);
        xp1.n = fma(0.5, s[1], qn[1]);
        xp1.t = fma(0.5, s[2], qn[2]);
        xp1.p = fma(0.5, s[3], qn[3]);
        if constexpr (!FACEC)
          xm1.c = xp1.c = 0.0;
        else if constexpr (PLM)
        {
          xm1.c = csound(gamma * xm1.p, xm1.r);
          xp1.c = csound(gamma * xp1.p, xp1.r);
        }
        else
        {
          xm1.c = ym.c; // PCM: one sound speed per cell
          xp1.c = ym.c;
        }
        if constexpr (PLM)
        {
          S.X1a[par ^ 1][t] = make_double2(xp1.r, xp1.n);
          S.X1b[par ^ 1][t] = make_double2(xp1.t, xp1.p);
        }
        if constexpr (FACEC)
          S.X1c[par ^ 1][t] = xp1.c;
        if constexpr (DIFF)
          S.X1T[par ^ 1][t] = Tn;
      }

      __syncthreads();

      // E. Q rows <= k+kDead and U rows <= k-1 are dead now: refill their slots with the rows kNS /
      // kNU further down the stream.  One block under one thread test: the other warps skip it with
      // a single branch instead of issuing the (predicated-off) producer instructions.
      if (t == 0)
      {
        const uint32_t qslot = (uint32_t)(kDead == 1 ? s1 : (kDead == 0 ? s0 : sm1));
        if (k + kDead + kNS <= rlast)
          stage_q(tma_xq, k + kDead + kNS, qslot);
        else
          stage_q_next(k + kDead + kNS - rlast, qslot);
        if (k == j1 - 8)
        {
          if (a.persistent)
            pend = gridDim.x + atomicAdd(&sc->work_next, 1u);
        }
        else if (k == j1 - 3)
        {
          if (a.persistent)
            cp_async_16(smem_u32(&S.item_in), a.items + min(pend, (unsigned)a.n_items));
        }
#if FV2D_TMA_STORE_U
        if (k > j0)
        {
          store_u(tma_xu, k - 1, (uint32_t)us_prev); // U^{n+1} of row k-1: every thread wrote it before the barrier
          if (k > j0 + 1)
          {
            bulk_wait_read<1>(); // the store of row k-2 (issued one iteration ago) has read its slot
            if (k - 2 + kNU < j1)
              stage_u(tma_xu, k - 2 + kNU, (uint32_t)us_prev2);
            else
              stage_u_next(k - 2 + kNU - j1, (uint32_t)us_prev2);
          }
        }
#else
        if (k > j0)
        {
          if (k - 1 + kNU < j1)
            stage_u(tma_xu, k - 1 + kNU, (uint32_t)us_prev);
          else
            stage_u_next(k - 1 + kNU - j1, (uint32_t)us_prev);
        }
#endif
      }

      // D. finish row k
      {
        FaceFlux fxr, fxl;
        {
          const double2 ra = S.X2a[par][tr], rb = S.X2b[par][tr];
          const double2 oa = S.X2a[par][t], ob = S.X2b[par][t];
          fxr.m = ra.x, fxr.n = ra.y, fxr.t = rb.x, fxr.e = rb.y;
          fxl.m = oa.x, fxl.n = oa.y, fxl.t = ob.x, fxl.e = ob.y;
        }
        // U^n of the own cell, staged by TMA two rows ago (the slot holds garbage in the warm-up
        // iteration and for the halo threads: neither stores anything)
        double un[4];
        if (k >= j0)
          mbar_wait(&S.ufull[us], uph);
        {
          // (only the threads that own a cell of the strip read the box: the two halo threads a side
          //  would alias the edge threads' elements, which those overwrite with U^{n+1} below)
          const double *ub = &S.uring[us][0][0];
#pragma unroll
          for (int f = 0; f < 4; ++f)
            un[f] = (t >= 2 && t < NT - 2) ? ub[f * W + tu] : 0.0;
        }

        // y fluxes back in the grid frame: (m, t, n, e) -> (rho, rho u, rho v, E)
        // (fy_lo already holds hyperbolic minus diffusive flux of the face below; same for fy_hi now)
        double fyl[4] = {fy_lo.m, fy_lo.t, fy_lo.n, fy_lo.e};
        double fyh[4] = {fy_hi.m, fy_hi.t - dy_t, fy_hi.n - dy_n, fy_hi.e - dy_e};

        double gyv = 0.0, gxv = 0.0, rho_k = 0.0;
        if constexpr (GRAV != 0)
        {
          rho_k = S.ring[s0][0][t];
          // getGravity (Gravity.h:39-57); GRAV == 2 also covers "well-balanced flux, no gravity" (g = 0)
          if (p.gravity_mode == FV2D_GRAV_CONSTANT)
            gxv = p.gx, gyv = p.gy;
          else if (p.gravity_mode == FV2D_GRAV_ANALYTICAL)
            gxv = gyv = a.kp.gtab[k];
        }

        double u4[4];
        // (every multiply-add of this kernel is spelled out as an fma and the file is compiled with
        //  --fmad=false: left to the compiler, the copies of the unrolled row loop can contract the
        //  same expression differently, and a row's last bit would depend on where its work item starts)
        u4[0] = fma(fyl[0] - fyh[0], dtdy, fma(fxl.m - fxr.m, dtdx, un[0]));
        u4[1] = fma(fyl[1] - fyh[1], dtdy, fma(fxl.n - fxr.n, dtdx, un[1]));
        u4[2] = fma(fyl[2] - fyh[2], dtdy, fma(fxl.t - fxr.t, dtdx, un[2]));
        u4[3] = fma(fyl[3] - fyh[3], dtdy, fma(fxl.e - fxr.e, dtdx, un[3]));
        if constexpr (GRAV != 0)
        {
          // Update.h:161-166: both sweeps add into IV (Q4)
          const double dr = dt * rho_k, hdt = dt * 0.5;
          u4[2] += fma(dr, gxv, dr * gyv);
          u4[3] += fma(hdt * (fxl.m + fxr.m), gxv, hdt * (fyl[0] + fyh[0]) * gyv);
        }

        if constexpr (GRAV == 2)
        {
          // well-balanced flux at the global y boundary (Update.h:148-156): replaces the HYPERBOLIC
          // flux of that face by {0, 0, pout -+ rho g dy, 0}; the diffusive part of the face stays.
          // Applied as an in-place correction on the two rows it concerns (for the low face the
          // roll below already dropped the hyperbolic part of the carried flux).
          if (k == p.jbeg && glob_lo)
            u4[2] = fma(dtdy, fma(-(rho_k * gyv), p.dy, fy_hi.pout), u4[2]);
          else if (k == p.jend - 1 && glob_hi)
          {
            u4[0] = fma(dtdy, fy_hi.m, u4[0]);
            u4[1] = fma(dtdy, fy_hi.t, u4[1]);
            u4[2] = fma(dtdy, fy_hi.n - fma(rho_k * gyv, p.dy, fy_lo.pout), u4[2]);
            u4[3] += fma(dtdy, fy_hi.e, -(dt * 0.5 * fy_hi.m * gyv));
          }
        }

        if constexpr (DIFF)
        {
          // ThermalConduction.h:77-103: with a temperature boundary condition the reference replaces
          // the cell's own x-flux FL (at jbeg) / FR (at jend-1) by a y-boundary expression (Q7a) -
          // for that cell only, not for the neighbour sharing the face.  Reproduced as a correction
          // to the face-based flux on those two rows.
          const bool row_lo = (k == p.jbeg && glob_lo && p.bctc_ymin != FV2D_BCTC_NONE);
          const bool row_hi = (k == p.jend - 1 && glob_hi && p.bctc_ymax != FV2D_BCTC_NONE);
          if (tc_on && (row_lo || row_hi))
          {
            const double TC = S.ring[s0][3][t] * frcp(S.ring[s0][0][t]);
            const double TL = S.ring[s0][3][tl] * frcp(S.ring[s0][0][tl]);
            const double TR = S.ring[s0][3][tr] * frcp(S.ring[s0][0][tr]);
            const double kap = p.kappa;
            if (row_lo)
            {
              const double FL = kap * (TC - TL) * rdx;
              const double FLb =
                  (p.bctc_ymin == FV2D_BCTC_FIXED_TEMPERATURE) ? kap * 2.0 * (TC - p.bctc_ymin_value) * rdy : kap * p.bctc_ymin_value;
              u4[3] = fma(dtdx, FL - FLb, u4[3]);
            }
            if (row_hi)
            {
              const double FR = kap * (TR - TC) * rdx;
              const double FRb =
                  (p.bctc_ymax == FV2D_BCTC_FIXED_TEMPERATURE) ? kap * 2.0 * (p.bctc_ymax_value - TC) * rdy : kap * p.bctc_ymax_value;
              u4[3] = fma(dtdx, FRb - FR, u4[3]);
            }
          }
        }

        if (interior && k >= j0)
        {
#define FV2D_AT(base, f) (*reinterpret_cast<double *>(reinterpret_cast<char *>(base) + (f) * planeB + offB))
#define FV2D_CAT(base, f) (*reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + (f) * planeB + offB))
          if (U0 != nullptr) // SSP-RK2 combine (Update.h:214-220)
          {
#pragma unroll
            for (int f = 0; f < 4; ++f)
              u4[f] = 0.5 * (FV2D_CAT(U0, f) + u4[f]);
          }
#if FV2D_TMA_STORE_U
          {
            double *ub = &S.uring[us][0][0];
#pragma unroll
            for (int f = 0; f < 4; ++f)
              ub[f * W + tu] = u4[f];
            fence_proxy_async_smem();
          }
#else
#pragma unroll
          for (int f = 0; f < 4; ++f)
            FV2D_AT(a.Uout, f) = u4[f];
#endif

          // consToPrim (States.h:32-43)
          const double ir = frcp(u4[0]);
          double qo[4];
          qo[0] = u4[0];
          qo[1] = u4[1] * ir;
          qo[2] = u4[2] * ir;
          qo[3] = fma(-0.5, fma(u4[2], qo[2], u4[1] * qo[1]), u4[3]) * gm1; // E - (rho u . u) / 2
          // stores the x-ghost copies of this cell (row jt of array `base`, v component v2)
          auto put_xghosts = [&](double *base, long long plane, int jt, double v2) {
#pragma unroll 1
            for (unsigned m = xmask; m != 0; m &= m - 1)
            {
              const int g = __ffs(m) - 1;
              double *d   = base + (long long)jt * L.pitch + L.lead + ((g < Ng) ? g : p.iend + g - Ng);
              d[0] = qo[0], d[plane] = (p.boundary_x == FV2D_BC_REFLECTING) ? -qo[1] : qo[1], d[2 * plane] = v2,
              d[3 * plane] = qo[3];
            }
          };
          if (final_stage)
          {
            // checkNegatives (SimInfo.h:612-633): counted straight into the device counters on
            // the (rare) event instead of carrying three counters through the sweep; one integer
            // test of the two sign bits guards both comparisons
            if ((__double2hiint(qo[0]) | __double2hiint(qo[3])) < 0)
            {
              if (qo[0] < 0.0)
              {
                qo[0] = a.kp.eps_reset;
                atomicAdd(&sc->neg[0], 1ULL);
              }
              if (qo[3] < 0.0)
              {
                qo[3] = a.kp.eps_reset;
                atomicAdd(&sc->neg[1], 1ULL);
              }
            }
            // computeDt of the new state (ComputeDt.h:30-34); a NaN never wins the Max
            // reduction (Kokkos::Max joins with `>`)
            const double cs = csound(gamma * qo[3], qo[0]);
            const double h  = fma(cs, rdxy, fma(fabs(qo[1]), rdx, fabs(qo[2]) * rdy));
            inv_dt_max      = (h > inv_dt_max) ? h : inv_dt_max;
            // NaN count (SimInfo.h:624-631): h is NaN whenever a field is, so the per-field
            // count runs only on that (rare) path
            if (h != h)
            {
              const int n = (qo[0] != qo[0]) + (qo[1] != qo[1]) + (qo[2] != qo[2]) + (qo[3] != qo[3]);
              atomicAdd(&sc->neg[2], (unsigned long long)n);
            }
          }
#pragma unroll
          for (int f = 0; f < 4; ++f)
            FV2D_AT(a.Qout, f) = qo[f];
#undef FV2D_AT
#undef FV2D_CAT
          // ghost copies of the new row.  x: the threads next to the domain's left / right edge
          // also store the ghost columns that mirror their column (xmask).  y: the rows within Ng of
          // the slab's edges also go into the ghost rows that mirror them (push_ghost_rows).
          if (xmask != 0)
            put_xghosts(a.Qout, L.plane, k, qo[2]);
          if (yitem && (k < p.jbeg + Ng || k >= p.jend - Ng))
          {
            // target 0 .. 2Ng-1: this slab's own y-ghost rows at a physical boundary
            // (BoundaryConditions.h:112-146, with the corners the x pass defines); 2Ng / 2Ng+1: the low /
            // high neighbour slab's ghost rows, straight into its memory over NVLink
#pragma unroll 1
            for (int tg = 0; tg < 2 * Ng + 2; ++tg)
            {
              double *base    = a.Qout;
              long long plane = L.plane;
              int jt;
              bool fv = false;
              if (tg < 2 * Ng)
              {
                const int side = (tg >= Ng) ? 1 : 0;
                jt             = side ? p.jend + tg - Ng : tg;
                if (!fold || (side ? a.kp.edge_hi : a.kp.edge_lo) != EDGE_PHYSICAL ||
                    bc_src(p.boundary_y, jt, p.jbeg, p.jend, p.Ny) != k)
                  continue;
                fv = (p.boundary_y == FV2D_BC_REFLECTING);
              }
              else if (tg == 2 * Ng)
              {
                if (peer_lo == nullptr || k >= p.jbeg + Ng)
                  continue;
                base = peer_lo, plane = a.peer_lo_plane, jt = a.peer_lo_Ny + k; // its high ghost rows
              }
              else
              {
                if (peer_hi == nullptr || k < p.jend - Ng)
                  continue;
                base = peer_hi, plane = a.peer_hi_plane, jt = k - p.Ny; // its low ghost rows
              }
              const double v2 = fv ? -qo[2] : qo[2];
              double *d       = base + (long long)jt * L.pitch + L.lead + col;
              d[0] = qo[0], d[plane] = qo[1], d[2 * plane] = v2, d[3 * plane] = qo[3];
              if (xmask != 0)
                put_xghosts(base, plane, jt, v2);
            }
          }
        }
      }

      // roll the column window and the ring bookkeeping
      offB += pitchB;
      if constexpr (GRAV == 2)
      {
        // the face below row jbeg: its hyperbolic flux will be replaced by the well-balanced one, so
        // only the diffusive part is carried (w is uniform: 1 everywhere else, and 1 * x is exact)
        const double w = (k + 1 == p.jbeg && glob_lo) ? 0.0 : 1.0;
        fy_lo.m = w * fy_hi.m, fy_lo.t = fma(w, fy_hi.t, -dy_t), fy_lo.n = fma(w, fy_hi.n, -dy_n);
        fy_lo.e = fma(w, fy_hi.e, -dy_e);
      }
      else
        fy_lo.m = fy_hi.m, fy_lo.t = fy_hi.t - dy_t, fy_lo.n = fy_hi.n - dy_n, fy_lo.e = fy_hi.e - dy_e;
      fy_lo.pout = fy_hi.pout;
      Tk = Tn;
      yp = yp1;
      xm    = xm1;
#pragma unroll
      for (int f = 0; f < 4; ++f)
        qn[f] = qnn[f];
      sm1 = s0, s0 = s1, s1 = s2;
      next_q();
      us_prev2 = us_prev;
      us_prev  = us;
      us = (us + 1 == kNU) ? 0 : us + 1;
      uph ^= (us == 0) ? 1u : 0u;
    }

#ifdef FV2D_TIMING
    {
      const long long now = clock64();
      tm_loop += now - tm_mark, tm_mark = now;
    }
#endif
    // ---- end of the item.  Its last Q rows and its last U row are still in the rings: once every
    // thread is done with them their slots move on down the stream.  Multi-GPU: tell the neighbours
    // how many of their ghost rows this item has delivered.
    int n_lo = 0, n_hi = 0;
#ifndef FV2D_X_NOCOUNT
    if constexpr (!PLAIN)
#else
    if constexpr (false)
#endif
    {
      n_lo = (peer_lo != nullptr) ? max(0, min(j1, p.jbeg + Ng) - j0) : 0;
      n_hi = (peer_hi != nullptr) ? max(0, j1 - max(j0, p.jend - Ng)) : 0;
      if (n_lo + n_hi > 0)
        __threadfence_system();
    }
    if (t == 0)
    {
      // the table entry that landed in item_in is the item after next; it takes the current item's
      // slot (read by everybody when the item started, and again after the barrier below)
      cp_async_wait_all();
      S.item[ci & 1] = S.item_in;
    }
    __syncthreads();
    if (t == 0)
    {
      if constexpr (!PLAIN)
      {
        if (n_lo)
          atomicAdd_system(&a.kp.peer_sc[a.lo_rank]->halo_cnt[1], (unsigned long long)n_lo);
        if (n_hi)
          atomicAdd_system(&a.kp.peer_sc[a.hi_rank]->halo_cnt[0], (unsigned long long)n_hi);
      }
      // the item's last Q rows (slots sm1 / s0 / s1 after the last roll) and its last U row move on
      // to rows of the next item
#pragma unroll 1
      for (int n = 2 + kDead; n <= 3; ++n)
        stage_q_next(n + kNS - 3, (uint32_t)(n == 1 ? sm1 : (n == 2 ? s0 : s1)));
#if FV2D_TMA_STORE_U
      store_u(tma_xu, j1 - 1, (uint32_t)us_prev); // the item's last row
      if (nrow >= 2)
      {
        bulk_wait_read<1>();
        stage_u_next(kNU - 2, (uint32_t)us_prev2);
      }
      bulk_wait_read<0>();
      stage_u_next(kNU - 1, (uint32_t)us_prev);
#else
      stage_u_next(kNU - 1, (uint32_t)us_prev);
#endif
      // the next item becomes the current one: publish the one after it to the producer
      publish_next(S.item[ci & 1]);
    }
  }

  // ---- CTA reduction of the CFL maximum, then the sweep's bookkeeping by its last CTA
  if (final_stage)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      inv_dt_max = fmax(inv_dt_max, __shfl_xor_sync(0xffffffffu, inv_dt_max, o));
  }
  __syncthreads(); // the exchange arrays are free now: reuse as scratch
  double *red = reinterpret_cast<double *>(&S.X2a[0][0]);
  if ((t & 31) == 0)
    red[t >> 5] = inv_dt_max;
  __syncthreads();
#ifdef FV2D_TIMING
  if (t == 0 && blockIdx.x < 1024)
  {
    const long long now = clock64();
    g_sweep_timing[blockIdx.x * 4 + 0] = tm_gap + (now - tm_mark); // everything outside the row loops
    g_sweep_timing[blockIdx.x * 4 + 1] = tm_loop;
    g_sweep_timing[blockIdx.x * 4 + 2] = tm_items;
    g_sweep_timing[blockIdx.x * 4 + 3] = now - tm_start;
  }
#endif
  if (t < 32)
  {
    int last = 0;
    double hyp_slab = 0.0;
    if (t == 0)
    {
      if (final_stage)
      {
        double m = red[0];
        for (int w = 1; w < NT / 32; ++w)
          m = fmax(m, red[w]);
        atomicMax(&sc->inv_acc[1][0], encode_ordered(m));
      }
      __threadfence();
      const unsigned prev = atomicAdd(&sc->cta_done, 1u);
      if (prev == gridDim.x - 1)
      {
        last          = 1;
        sc->cta_done  = 0;
        sc->work_next = 0;
        if (final_stage && !a.partial)
          hyp_slab = decode_ordered(atomicExch(&sc->inv_acc[1][0], FV2D_ENC_NEG_MAX));
      }
    }
    last     = __shfl_sync(0xffffffffu, last, 0);
    hyp_slab = __shfl_sync(0xffffffffu, hyp_slab, 0);
    // (a partial launch - the streamed host path sweeps the slab row block by row block - leaves the
    //  maximum in the accumulator and the clock alone: fv2d_stream.cu commits the step)
    if (last && final_stage && !a.partial)
    {
      // The sweep's last CTA: the slab's maximum goes to every rank's mailbox (self included), one
      // lane per rank - the next step's dt is reduced on the device, no host round trip.  Then the
      // device-side clock (main.cpp:83).
      if (PLAIN && a.kp.nranks == 1)
      {
        if (t == 0)
          post_cfl_mail_local(a.kp, hyp_slab, a.mail_gen + 1);
      }
      else if (t < a.kp.nranks)
        post_cfl_mail_to(a.kp, t, hyp_slab, a.mail_gen + 1);
      if (t == 0)
      {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(sc->tstamp[2]));
        sc->dt                                  = dt;
        sc->dt_hist[sc->step % FV2D_DT_HISTORY] = dt;
        sc->t += dt;
        sc->step += 1;
        if (a.use_device_dt)
          sc->inv_dt_last[0] = S.inv3[0], sc->inv_dt_last[1] = S.inv3[1], sc->inv_dt_last[2] = S.inv3[2];
      }
    }
  }
}

#ifndef FV2D_SOLVER_ONLY
// --------------------------------------------------------------------------- math probe

// Evaluates the sweep's division-free primitives on arbitrary inputs so that their accuracy is a
// tested property (tests/test_gpu_properties.py): out_rcp[i] = frcp(a[i]),
// out_cs[i] = csound(a[i], b[i]) = sqrt(a[i] / b[i]).
__global__ void k_math_probe(long long n, const double *__restrict__ a, const double *__restrict__ b,
                             double *__restrict__ out_rcp, double *__restrict__ out_cs)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out_rcp[i] = frcp(a[i]);
    out_cs[i]  = csound(a[i], b[i]);
  }
}
void launch_math_probe(long long n, const double *a, const double *b, double *out_rcp, double *out_cs, cudaStream_t s)
{
  k_math_probe<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, a, b, out_rcp, out_cs);
}

// --------------------------------------------------------------------------- stand-alone ghost fill

// Ghost fill of a Q array whose ghosts were not written by a sweep (after an upload, after the
// operator-level calls, and to complete the array that fv2d_advance_host hands back): the x and
// y passes of BoundaryConditions.h:82-147 composed (see fv2d_ops.cu).  Ghost rows on a
// neighbour-slab side are left to the halo exchange: the kernel waits until the neighbour has
// pushed all of them, then applies the x boundary condition to their x-ghost columns.
__global__ void k_fill_ghosts(KParams kp, double *__restrict__ Q, unsigned long long halo_expected)
{
  const fv2d_device_params &p = kp.p;
  const long long tid         = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ng = p.Ng, Ntx = p.Ntx;
  const long long n_y = 2LL * Ng * Ntx, n_x = 2LL * Ng * p.Ny;
  if (tid >= n_y + n_x)
    return;
  int i, j;
  if (tid < n_y)
  {
    const int r = int(tid / Ntx);
    i           = int(tid - (long long)r * Ntx);
    j           = (r < Ng) ? r : p.jend + (r - Ng);
  }
  else
  {
    const long long q = tid - n_y;
    const int r       = int(q / (2 * Ng));
    const int c       = int(q - (long long)r * (2 * Ng));
    j                 = p.jbeg + r;
    i                 = (c < Ng) ? c : p.iend + (c - Ng);
  }
  int js = j, is = i;
  bool flip_u = false, flip_v = false;
  if (j < p.jbeg || j >= p.jend)
  {
    const int side = (j < p.jbeg) ? 0 : 1;
    if ((side == 0 ? kp.edge_lo : kp.edge_hi) != EDGE_PHYSICAL)
    {
      if (i >= p.ibeg && i < p.iend)
        return;
      wait_ge_sys(&kp.sc->halo_cnt[side], halo_expected, kp.sc);
    }
    else
    {
      js     = bc_src(p.boundary_y, j, p.jbeg, p.jend, p.Ny);
      flip_v = (p.boundary_y == FV2D_BC_REFLECTING);
    }
  }
  if (i < p.ibeg || i >= p.iend)
  {
    is     = bc_src(p.boundary_x, i, p.ibeg, p.iend, p.Nx);
    flip_u = (p.boundary_x == FV2D_BC_REFLECTING);
  }
  const long long os = kp.L.at(0, is, js), od = kp.L.at(0, i, j);
  const double r = Q[os], u = Q[os + kp.L.plane], v = Q[os + 2 * kp.L.plane], pr = Q[os + 3 * kp.L.plane];
  Q[od]                   = r;
  Q[od + kp.L.plane]      = flip_u ? -u : u;
  Q[od + 2 * kp.L.plane]  = flip_v ? -v : v;
  Q[od + 3 * kp.L.plane]  = pr;
}

void launch_fill_ghosts(const KParams &kp, double *Q, unsigned long long halo_expected, cudaStream_t s)
{
  const long long n = 2LL * kp.p.Ng * kp.p.Ntx + 2LL * kp.p.Ng * kp.p.Ny;
  k_fill_ghosts<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(kp, Q, halo_expected);
}

#endif // !FV2D_SOLVER_ONLY

// --------------------------------------------------------------------------- dispatch

#ifndef FV2D_NT
#define FV2D_NT 256
#endif
constexpr int kNT = FV2D_NT;

// The 108 instantiations (reconstruction x solver x gravity x diffusion x mode) are compiled in
// three translation units, one per Riemann solver (-DFV2D_SOLVER_ONLY=<solver>), so that the library
// builds in parallel; the unit without that macro holds the dispatcher and the small kernels.
#ifdef FV2D_SOLVER_ONLY

template <bool PLM, int SOLVER, int GRAV, bool DIFF, int MODE>
static cudaError_t launch_variant(const CUtensorMap &tmQ, const CUtensorMap &tmU, const SweepArgs &a, cudaStream_t s,
                                  bool configure_only)
{
  auto kern             = k_sweep<kNT, PLM, SOLVER, GRAV, DIFF, MODE>;
  constexpr bool FACEC  = !(PLM && SOLVER == FV2D_HLLC);
  constexpr int NS = ring_ns(ring_total(PLM, FACEC, DIFF), ring_dead(PLM, GRAV, DIFF), DIFF);
  constexpr int NU = ring_total(PLM, FACEC, DIFF) - NS;
  constexpr size_t smem = sizeof(SweepSmem<kNT, NS, NU, PLM, FACEC, DIFF>) + FV2D_EXTRA_SMEM;
  static_assert(kNT != 256 || smem <= 115712, "two CTAs per SM need <= 113 KB of shared memory each");
  if (configure_only)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // persistent: CTA b starts on work item b and pulls the others from the device-wide counter.
  // Programmatic stream serialization: the CTAs of this launch may become resident - and run the part
  // of their prologue in front of griddepcontrol.wait - while the last CTAs of the previous kernel in
  // the stream are still finishing.
  static const bool pdl = std::getenv("FV2D_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim          = dim3((unsigned)a.n_ctas);
  cfg.blockDim         = dim3(kNT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = s;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs    = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, tmQ, tmU, a);
}

template <bool PLM, int SOLVER, int GRAV, bool DIFF>
static cudaError_t launch_one(const CUtensorMap &tmQ, const CUtensorMap &tmU, const SweepArgs &a, cudaStream_t s,
                              bool configure_only)
{
#ifdef FV2D_ONLY_MODE // development: compile a single mode
  return launch_variant<PLM, SOLVER, GRAV, DIFF, FV2D_ONLY_MODE>(tmQ, tmU, a, s, configure_only);
#endif
  if (configure_only)
  {
    cudaError_t e = launch_variant<PLM, SOLVER, GRAV, DIFF, kPlain>(tmQ, tmU, a, s, true);
    if (e == cudaSuccess)
      e = launch_variant<PLM, SOLVER, GRAV, DIFF, kMulti>(tmQ, tmU, a, s, true);
    return e != cudaSuccess ? e : launch_variant<PLM, SOLVER, GRAV, DIFF, kGeneral>(tmQ, tmU, a, s, true);
  }
  const bool euler = a.final_stage && a.U0 == nullptr;
  static const int force_mode = std::getenv("FV2D_FORCE_MODE") ? std::atoi(std::getenv("FV2D_FORCE_MODE")) : 0; // development
  const bool alone = a.peer_lo_Qout == nullptr && a.peer_hi_Qout == nullptr && force_mode < 1;
  if (force_mode >= 2)
    return launch_variant<PLM, SOLVER, GRAV, DIFF, kGeneral>(tmQ, tmU, a, s, false);
  if (euler && alone)
    return launch_variant<PLM, SOLVER, GRAV, DIFF, kPlain>(tmQ, tmU, a, s, false);
  if (euler)
    return launch_variant<PLM, SOLVER, GRAV, DIFF, kMulti>(tmQ, tmU, a, s, false);
  return launch_variant<PLM, SOLVER, GRAV, DIFF, kGeneral>(tmQ, tmU, a, s, false);
}

template <bool PLM, int SOLVER>
static cudaError_t dispatch2(const CUtensorMap &tmQ, const CUtensorMap &tmU, const SweepArgs &a, cudaStream_t s, bool cfg,
                             int grav, bool diff)
{
#ifdef FV2D_ONLY_ONE // development: compile a single (reconstruction, gravity, diffusion) combination
  if (PLM)
    return launch_one<true, SOLVER, 0, false>(tmQ, tmU, a, s, cfg);
  return cudaErrorInvalidValue;
#endif
  switch (grav)
  {
  case 0:
    return diff ? launch_one<PLM, SOLVER, 0, true>(tmQ, tmU, a, s, cfg) : launch_one<PLM, SOLVER, 0, false>(tmQ, tmU, a, s, cfg);
  case 1:
    return diff ? launch_one<PLM, SOLVER, 1, true>(tmQ, tmU, a, s, cfg) : launch_one<PLM, SOLVER, 1, false>(tmQ, tmU, a, s, cfg);
  default:
    return diff ? launch_one<PLM, SOLVER, 2, true>(tmQ, tmU, a, s, cfg) : launch_one<PLM, SOLVER, 2, false>(tmQ, tmU, a, s, cfg);
  }
}

// entry point of this translation unit's solver
#define FV2D_CAT2(a, b) a##b
#define FV2D_ENTRY(n) FV2D_CAT2(sweep_entry_, n)
cudaError_t FV2D_ENTRY(FV2D_SOLVER_ONLY)(const CUtensorMap &tmQ, const CUtensorMap &tmU, const SweepArgs &a, cudaStream_t s,
                                         bool cfg, bool plm, int grav, bool diff)
{
  return plm ? dispatch2<true, FV2D_SOLVER_ONLY>(tmQ, tmU, a, s, cfg, grav, diff)
             : dispatch2<false, FV2D_SOLVER_ONLY>(tmQ, tmU, a, s, cfg, grav, diff);
}
#define FV2D_TIMING_READER(n) FV2D_CAT2(read_sweep_timing_, n)
int FV2D_TIMING_READER(FV2D_SOLVER_ONLY)(long long *host, int n)
{
#ifdef FV2D_TIMING
  return cudaMemcpyFromSymbol(host, g_sweep_timing, sizeof(long long) * (size_t)(n < 4096 ? n : 4096)) == cudaSuccess ? 0 : 2;
#else
  (void)host, (void)n;
  return 1;
#endif
}

#else // ---- the dispatcher unit

int sweep_strip_width() { return kNT - 4; }

// sweep_entry_<FV2D_HLL | FV2D_HLLC | FV2D_FSLP>: one translation unit each
cudaError_t sweep_entry_0(const CUtensorMap &, const CUtensorMap &, const SweepArgs &, cudaStream_t, bool, bool, int, bool);
cudaError_t sweep_entry_1(const CUtensorMap &, const CUtensorMap &, const SweepArgs &, cudaStream_t, bool, bool, int, bool);
cudaError_t sweep_entry_2(const CUtensorMap &, const CUtensorMap &, const SweepArgs &, cudaStream_t, bool, bool, int, bool);
int read_sweep_timing_0(long long *, int);
int read_sweep_timing_1(long long *, int);
int read_sweep_timing_2(long long *, int);
static_assert(FV2D_HLL == 0 && FV2D_HLLC == 1 && FV2D_FSLP == 2, "solver enum values name the translation units");

static cudaError_t dispatch_solver(const CUtensorMap &tmQ, const CUtensorMap &tmU, const SweepArgs &a, cudaStream_t s, bool cfg,
                                   int solver, bool plm, int grav, bool diff)
{
  switch (solver)
  {
  case FV2D_HLL:
    return sweep_entry_0(tmQ, tmU, a, s, cfg, plm, grav, diff);
  case FV2D_FSLP:
    return sweep_entry_2(tmQ, tmU, a, s, cfg, plm, grav, diff);
  default:
    return sweep_entry_1(tmQ, tmU, a, s, cfg, plm, grav, diff);
  }
}

cudaError_t launch_sweep(const CUtensorMap &tmapQ, const CUtensorMap &tmapU, const SweepArgs &a, cudaStream_t s)
{
  const fv2d_device_params &p = a.kp.p;
  const bool plm  = (p.reconstruction == FV2D_PLM); // PCM_WB == PCM (Q3)
  // 0: nothing; 1: gravity source terms; 2: + well-balanced flux at the y boundary (Update.h:148-156,
  // which the reference applies whatever the gravity mode)
  const int grav  = p.well_balanced_flux_at_y_bc ? 2 : (p.gravity_mode != FV2D_GRAV_NONE ? 1 : 0);
  const bool diff = p.thermal_conductivity_active || p.viscosity_active;
  return dispatch_solver(tmapQ, tmapU, a, s, false, p.riemann_solver, plm, grav, diff);
}

cudaError_t sweep_configure()
{
  CUtensorMap dummy;
  memset(&dummy, 0, sizeof dummy);
  SweepArgs a;
  memset(&a, 0, sizeof a);
  for (int plm = 0; plm < 2; ++plm)
    for (int solver = 0; solver < 3; ++solver)
      for (int grav = 0; grav < 3; ++grav)
        for (int diff = 0; diff < 2; ++diff)
        {
          cudaError_t e = dispatch_solver(dummy, dummy, a, nullptr, true, solver, plm != 0, grav, diff != 0);
          if (e != cudaSuccess)
            return e;
        }
  return cudaSuccess;
}

// Development hook (variant builds with -DFV2D_TIMING): per-CTA cycle counts of the last sweep of
// the given solver's translation unit.
int read_sweep_timing(int solver, long long *host, int n)
{
  return solver == FV2D_HLL ? read_sweep_timing_0(host, n) : (solver == FV2D_FSLP ? read_sweep_timing_2(host, n) : read_sweep_timing_1(host, n));
}

#endif // FV2D_SOLVER_ONLY

} // namespace fv2d
