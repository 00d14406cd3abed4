// fv2d_kernels.h — host-callable launchers implemented in fv2d_ops.cu / fv2d_sweep.cu.
#pragma once

#include "fv2d_common.cuh"

namespace fv2d
{

// ---- operator-level kernels (fv2d_ops.cu, bit-exact reference order)
void launch_fill_boundaries(const KParams &kp, double *Q, cudaStream_t s);
void launch_fill_x_ghosts_of_halo_rows(const KParams &kp, double *Q, cudaStream_t s);
void launch_prim_to_cons(const KParams &kp, const double *Q, double *U, cudaStream_t s);
void launch_cons_to_prim(const KParams &kp, const double *U, double *Q, cudaStream_t s);
void launch_check_negatives(const KParams &kp, double *Q, unsigned long long *counts, cudaStream_t s);
void launch_compute_dt(const KParams &kp, const double *Q, unsigned long long *acc, cudaStream_t s);
void launch_finalize_dt(const KParams &kp, const unsigned long long *acc, unsigned long long mail_gen, cudaStream_t s);
void launch_compute_slopes(const KParams &kp, const double *Q, double *sX, double *sY, cudaStream_t s);
void launch_fluxes_and_update(const KParams &kp, const double *Q, const double *sX, const double *sY, double *Unew,
                              double dt, cudaStream_t s);
void launch_thermal_conduction(const KParams &kp, const double *Q, double *Unew, double dt, cudaStream_t s);
void launch_viscosity(const KParams &kp, const double *Q, double *Unew, double dt, cudaStream_t s);
void launch_rk2_correct(const KParams &kp, const double *U0, double *Unew, cudaStream_t s);
void launch_mass_energy(const KParams &kp, const double *U, double *rowsum, cudaStream_t s);
void launch_state_hash(const KParams &kp, const double *U, unsigned long long *out, cudaStream_t s);
void launch_fp64_peak(int blocks, int iters, double *out, cudaStream_t s);

// ---- streamed host path (fv2d_stream.cu): what a state arriving from the host row block by row block needs
// Ghost cells inside array rows [ra, rb) of a single slab's Q (x-ghost columns of domain rows, whole
// y-ghost rows), composed x/y boundary passes as in k_fill_boundaries, and U = primToCons of them; the
// source rows must be resident.
void launch_fill_ghosts_rows(const KParams &kp, double *Q, double *U, int ra, int rb, cudaStream_t s);
// Array rows [ra, rb), all columns: [Q = the rows of `dense`, the device copy of the host array in the
// host's own layout, if dense != nullptr;] U = primToCons(Q); the hyperbolic CFL maximum of the DOMAIN
// cells among them - in the sweep's arithmetic, so that it equals bit for bit what the sweep that
// produced the state left behind - accumulated into *acc (order-preserving encoding, atomicMax).
void launch_prep_rows(const KParams &kp, const double *dense, double *Q, double *U, int ra, int rb, unsigned long long *acc,
                      cudaStream_t s);
// Array rows [ra, rb) of Q, all columns, into the dense (host-layout) staging array.
void launch_pack_rows(const KParams &kp, const double *Q, double *dense, int ra, int rb, cudaStream_t s);
// Ends a streamed step taken with the caller's dt (`hint`): if the input state's own CFL time step
// (from inv_acc[0]) equals `hint` the step is committed (clock, dt history, CFL mail `mail_gen` of the
// new state from inv_acc[1], sc->stream_ok = 1), else the speculative bookkeeping is undone (stream_ok = 0).
void launch_stream_commit(const KParams &kp, double hint, unsigned long long mail_gen, cudaStream_t s);

// ---- fused hot path (fv2d_sweep.cu)

// Stand-alone ghost fill of Q (composed x/y passes; waits for `halo_expected` pushed rows on a
// neighbour-slab side).  Only needed when Q was not produced by a sweep: the sweep writes the
// ghosts of its own output.
void launch_fill_ghosts(const KParams &kp, double *Q, unsigned long long halo_expected, cudaStream_t s);

// One fused Runge-Kutta stage:  Uout = Uin + dt*L(Qin) [ ; Uout = 0.5*(U0 + Uout) ],
// Qout = consToPrim(Uout) incl. its ghost cells [ ; checkNegatives ; inverse-dt maxima of Qout ;
// t += dt ].
struct SweepArgs
{
  KParams kp;
  const double *Uin;
  double *Uout;
  const double *U0; // RK2 stage 2: the state at the start of the step; else nullptr
  double *Qout;
  int final_stage;   // 1: checkNegatives + dt reduction of the new state + clock advance
  int partial;       // 1: the launch covers only part of the slab's rows (streamed host path): its CFL maximum stays
                     // in the accumulator, no mail, no clock advance - fv2d_stream.cu commits the step
  int use_device_dt; // 1: dt = CFL / max(inverse time-steps mailed by generation `mail_gen`); 0: dt = dt_host
  double dt_host;
  int fold_ghosts;   // 1: the epilogue also writes the ghost cells of Qout (boundary conditions)
  int n_ctas;        // grid size = min(n_items, resident CTA slots); = n_items when !persistent
  int persistent;    // 1: CTAs pull further items from the device-wide counter (every item has >= 8 rows)
  const CUtensorMap *tm_store_u; // device copy of the TMA descriptor the stage STORES Uout rows through (columns
                                 // beyond the domain clipped)
  const WorkItem *items; // device table, n_items entries + the end marker
  int n_items;
  // multi-GPU: the neighbours' copy of Qout (peer-mapped), or nullptr at a physical edge.  The
  // stage epilogue stores its edge rows straight into the neighbour's ghost rows.
  double *peer_lo_Qout, *peer_hi_Qout;
  int lo_rank, hi_rank;
  int peer_lo_Ny;                         // rows the low neighbour owns
  long long peer_lo_plane, peer_hi_plane; // plane strides of the neighbours' arrays
  unsigned long long mail_gen;      // generation of the CFL mails this step's dt is made of; the final stage posts mail_gen + 1
  unsigned long long halo_expected; // ghost rows each neighbour must have pushed before they are read
};
// tmapQ / tmapU must describe the arrays the stage READS (Qin: boxes of strip width + 4 columns;
// Uin: boxes of strip width columns).  Returns cudaSuccess or the launch error.
cudaError_t launch_sweep(const CUtensorMap &tmapQ, const CUtensorMap &tmapU, const SweepArgs &a, cudaStream_t s);
// Opt-in dynamic shared memory etc.; call once per process before the first sweep.
cudaError_t sweep_configure();
int sweep_strip_width();
int read_sweep_timing(int solver, long long *host, int n); // development hook: 0 ok, 1 not compiled in, 2 CUDA error
// Accuracy probe of the sweep's reciprocal / sound-speed primitives (device pointers).
void launch_math_probe(long long n, const double *a, const double *b, double *out_rcp, double *out_cs, cudaStream_t s);

} // namespace fv2d
