// fv2d_stream.cu — device side of the streamed host path (fv2d_advance_host_stream, fv2d_capi.cu).
//
// A state that lives on the host crosses PCIe twice per time step; done one after the other
// (upload everything, step, download everything) the link runs half-duplex.  The streamed path
// moves the state in row blocks and keeps three things going at once: block b+1 coming up, block b
// being swept (partial launches of the fused sweep, fv2d_sweep.cu), block b-1 going down.  What a
// row block needs between arriving and being swept is here:
//   * its ghost cells (BoundaryConditions.h:82-147; the x and y passes composed, see fv2d_ops.cu),
//   * U = primToCons(Q) (SimInfo.h:589-600), in the operator-level arithmetic (--fmad=false:
//     bit-identical to the reference and to fv2d_prim_to_cons),
//   * its contribution to the CFL maximum of the incoming state (ComputeDt.h:30-34) in the SWEEP's
//     arithmetic: for a state that an earlier step produced, the maximum found here equals bit for
//     bit the one that step's sweep left behind, so the dt the caller passes back in can be checked,
// and the kernel that ends the step: commit if the caller's dt was the state's own, undo if not.
#include "fv2d_fastmath.cuh"
#include "fv2d_kernels.h"
#include "fv2d_physics.cuh"

namespace fv2d
{

// Source index of ghost index k (same rule as bc_source in fv2d_ops.cu)
__device__ __forceinline__ int stream_bc_src(int bc, int k, int beg, int end, int N)
{
  switch (bc)
  {
  case FV2D_BC_REFLECTING:
    return 2 * (k < beg ? beg : end) - k - 1; // BoundaryConditions.h:25-38
  case FV2D_BC_PERIODIC:
    return k < beg ? k + N : k - N; // BoundaryConditions.h:53-68
  default:
    return k < beg ? beg : end - 1; // BoundaryConditions.h:94-95, 124-125
  }
}

// One CTA row per array row of [ra, rb): a y-ghost row is filled over its full width, a domain row
// only in its 2 Ng x-ghost columns.  Every ghost is a (sign-flipped) copy of ONE domain cell; U gets
// primToCons of the ghost value (what fv2d_prim_to_cons over the whole array leaves there).
__global__ void k_fill_ghosts_rows(KParams kp, double *__restrict__ Q, double *__restrict__ U, int ra, int rb)
{
  const fv2d_device_params &p = kp.p;
  const int j = ra + blockIdx.y;
  if (j >= rb)
    return;
  const bool yghost = (j < p.jbeg || j >= p.jend);
  const int n       = yghost ? p.Ntx : 2 * p.Ng;
  int js = j;
  bool flip_v = false;
  if (yghost)
  {
    js     = stream_bc_src(p.boundary_y, j, p.jbeg, p.jend, p.Ny);
    flip_v = (p.boundary_y == FV2D_BC_REFLECTING);
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
  {
    const int i = yghost ? c : ((c < p.Ng) ? c : p.iend + (c - p.Ng));
    int is = i;
    bool flip_u = false;
    if (i < p.ibeg || i >= p.iend)
    {
      is     = stream_bc_src(p.boundary_x, i, p.ibeg, p.iend, p.Nx);
      flip_u = (p.boundary_x == FV2D_BC_REFLECTING);
    }
    State q = load_state(Q, kp.L, is, js);
    if (flip_u)
      q.v[IU] *= -1.0;
    if (flip_v)
      q.v[IV] *= -1.0;
    store_state(Q, kp.L, i, j, q);
    store_state(U, kp.L, i, j, prim_to_cons(q, p.gamma0));
  }
}

void launch_fill_ghosts_rows(const KParams &kp, double *Q, double *U, int ra, int rb, cudaStream_t s)
{
  if (rb <= ra)
    return;
  // a domain row has 2 Ng ghost cells; the (at most 2 Ng) y-ghost rows of a call loop over their width
  dim3 grid(4, (unsigned)(rb - ra));
  k_fill_ghosts_rows<<<grid, 256, 0, s>>>(kp, Q, U, ra, rb);
}

// Rows [ra, rb), all columns: U = primToCons(Q) and the CFL maximum over the domain cells among them.
// With `dense` != nullptr the rows first come out of the staging copy of the host array (dense
// [field][Nty][Ntx], where flat PCIe copies put them: padded-row 2-D copies run ~10 % slower when
// both directions of the link are busy) into the padded planes of Q.
// The inverse time step of a cell is the expression of the sweep's epilogue (fv2d_sweep.cu, "computeDt
// of the new state"), operation for operation; a NaN never wins (Kokkos::Max joins with `>`).
__global__ void k_prep_rows(KParams kp, const double *__restrict__ dense, double *__restrict__ Q, double *__restrict__ U,
                            int ra, int rb, unsigned long long *__restrict__ acc)
{
  const fv2d_device_params &p = kp.p;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = ra + blockIdx.y * blockDim.y + threadIdx.y;
  double h_max = -1.7976931348623157e308;
  if (i < p.Ntx && j < rb)
  {
    State q;
    if (dense != nullptr)
    {
      const long long o = (long long)j * p.Ntx + i, pl = (long long)p.Nty * p.Ntx;
#pragma unroll
      for (int f = 0; f < 4; ++f)
        q.v[f] = dense[o + f * pl];
      store_state(Q, kp.L, i, j, q);
    }
    else
      q = load_state(Q, kp.L, i, j);
    store_state(U, kp.L, i, j, prim_to_cons(q, p.gamma0));
    if (i >= p.ibeg && i < p.iend && j >= p.jbeg && j < p.jend)
    {
      const double rdx = 1.0 / p.dx, rdy = 1.0 / p.dy;
      const double rdxy = rdx + rdy;
      const double cs   = csound(p.gamma0 * q.v[IP], q.v[IR]);
      const double h    = fma(cs, rdxy, fma(fabs(q.v[IU]), rdx, fabs(q.v[IV]) * rdy));
      h_max             = (h > h_max) ? h : h_max;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    h_max = fmax(h_max, __shfl_xor_sync(0xffffffffu, h_max, o));
  __shared__ double smax[32];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0)
    smax[tid >> 5] = h_max;
  __syncthreads();
  if (tid == 0)
  {
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    double m     = smax[0];
    for (int w = 1; w < nw; ++w)
      m = fmax(m, smax[w]);
    atomicMax(acc, encode_ordered(m));
  }
}

void launch_prep_rows(const KParams &kp, const double *dense, double *Q, double *U, int ra, int rb, unsigned long long *acc,
                      cudaStream_t s)
{
  if (rb <= ra)
    return;
  dim3 blk(128, 2);
  dim3 grid((unsigned)((kp.p.Ntx + blk.x - 1) / blk.x), (unsigned)((rb - ra + blk.y - 1) / blk.y));
  k_prep_rows<<<grid, blk, 0, s>>>(kp, dense, Q, U, ra, rb, acc);
}

// Rows [ra, rb) of the padded planes of Q, all columns, into the dense staging copy the flat
// device-to-host copies read.
__global__ void k_pack_rows(KParams kp, const double *__restrict__ Q, double *__restrict__ dense, int ra, int rb)
{
  const fv2d_device_params &p = kp.p;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = ra + blockIdx.y * blockDim.y + threadIdx.y;
  if (i < p.Ntx && j < rb)
  {
    const State q     = load_state(Q, kp.L, i, j);
    const long long o = (long long)j * p.Ntx + i, pl = (long long)p.Nty * p.Ntx;
#pragma unroll
    for (int f = 0; f < 4; ++f)
      dense[o + f * pl] = q.v[f];
  }
}

void launch_pack_rows(const KParams &kp, const double *Q, double *dense, int ra, int rb, cudaStream_t s)
{
  if (rb <= ra)
    return;
  dim3 blk(128, 2);
  dim3 grid((unsigned)((kp.p.Ntx + blk.x - 1) / blk.x), (unsigned)((rb - ra + blk.y - 1) / blk.y));
  k_pack_rows<<<grid, blk, 0, s>>>(kp, Q, dense, ra, rb);
}

// dt = CFL / max({hyp, tc, visc}) (ComputeDt.h:36-64), as the sweep's prologue evaluates it
__device__ __forceinline__ double dt_of(const fv2d_device_params &p, double hyp, double &tc, double &visc)
{
  tc = p.epsilon, visc = p.epsilon;
  if (p.thermal_conductivity_active)
    tc = fmax(2.0 * p.kappa / (p.dx * p.dx), 2.0 * p.kappa / (p.dy * p.dy));
  if (p.viscosity_active)
    visc = fmax(2.0 * p.mu / (p.dx * p.dx), 2.0 * p.mu / (p.dy * p.dy));
  double m = hyp;
  if (m < tc)
    m = tc;
  if (m < visc)
    m = visc;
  return p.CFL / m;
}

// End of a streamed step (one thread).  inv_acc[0] holds the CFL maximum of the state that came up
// from the host, inv_acc[1] the one of the state the partial sweeps produced with dt = hint.
__global__ void k_stream_commit(KParams kp, double hint, unsigned long long mail_gen)
{
  DevScalars *const sc = kp.sc;
  double tc, visc;
  const double hyp_in  = decode_ordered(sc->inv_acc[0][0]);
  const double dt_true = dt_of(kp.p, hyp_in, tc, visc);
  const double hyp_out = decode_ordered(atomicExch(&sc->inv_acc[1][0], FV2D_ENC_NEG_MAX));
  if (dt_true == hint)
  {
    post_cfl_mail(kp, hyp_out, mail_gen);
    sc->dt                                  = hint;
    sc->dt_hist[sc->step % FV2D_DT_HISTORY] = hint;
    sc->t += hint;
    sc->step += 1;
    sc->inv_dt_last[0] = hyp_in, sc->inv_dt_last[1] = tc, sc->inv_dt_last[2] = visc;
    sc->dt_next        = dt_of(kp.p, hyp_out, tc, visc);
    sc->stream_ok      = 1;
  }
  else
  {
    // the speculative sweeps counted their negative resets into the cumulative counters: undo
    for (int k = 0; k < 4; ++k)
      sc->neg[k] = sc->neg_save[k];
    sc->stream_ok = 0;
  }
}

void launch_stream_commit(const KParams &kp, double hint, unsigned long long mail_gen, cudaStream_t s)
{
  k_stream_commit<<<1, 1, 0, s>>>(kp, hint, mail_gen);
}

} // namespace fv2d
