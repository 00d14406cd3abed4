/* fv2d_b200.h — C ABI of libfv2d_b200.so: the B200-native (sm_100a) replacement for the
 * per-timestep finite-volume update of mdelorme/fv2d.
 *
 * The reference has no FFI layer; its "operator surface" is a handful of C++ functor
 * classes constructed from one Params object and driven by main.cpp.  Each entry point
 * below names the reference interface it replaces (file:line in the reference tree).
 * The C++ host mirror in fv2d_b200/host/ (UpdateFunctor, ComputeDtFunctor, ...) and the
 * Python binding in fv2d_b200/capi.py are thin layers over exactly these symbols.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; fv2d_last_error() returns a
 *     thread-local, human-readable description of the last failure;
 *   - there is NO CPU fallback: on a machine without a usable sm_100 device every compute
 *     entry point fails with FV2D_ERR_CUDA;
 *   - host arrays are fp64 SoA planes over the full grid including ghosts,
 *     A[f][j][i], f in 0..3 (rho,u,v,P | rho,rho*u,rho*v,E), j in 0..Nty-1, i in 0..Ntx-1,
 *     no padding — i.e. the reference's Q(j,i,f) (main.cpp:33-34) with the field index
 *     moved outermost;
 *   - a context owns the device copies of Q and U (reference main.cpp:33-34), the scratch
 *     arrays the reference functors own (slopes: Update.h:54-55; RK2 temporaries:
 *     Update.h:200-201) and one CUDA stream.  Calls on one context are stream-ordered;
 *     functions that return a value to the host synchronise that stream.
 */
#ifndef FV2D_B200_H_
#define FV2D_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "fv2d_params.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FV2D_OK 0
#define FV2D_ERR_ARG 1         /* bad argument / unsupported configuration */
#define FV2D_ERR_CUDA 2        /* CUDA runtime/driver failure, or no sm_100 device */
#define FV2D_ERR_IO 3          /* file could not be read / written */
#define FV2D_ERR_CONFIG 4      /* .ini error: duplicate key, bad enum string, unknown problem */

typedef struct fv2d_ctx fv2d_ctx;

const char *fv2d_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int fv2d_abi_version(void);

/* ------------------------------------------------------------------ host-side setup */

/* replaces readInifile(filename) -> Params                         (SimInfo.h:529-568)
 * `overrides` is NULL or a ';'-separated list "section.key=value;..." applied as if the
 * lines were present in the file.  Warnings about unknown sections/keys
 * (checkValidityIni, SimInfo.h:501-527) go to stderr like the reference. */
int fv2d_params_from_ini(const char *ini_path, const char *overrides, fv2d_device_params *dev, fv2d_run_params *run);

/* replaces Reader::outputValues via IOManager's ctor                (SimInfo.h:224-262,
 * IOManager.h:90-94): writes the effective configuration of `ini_path` to `out_path`. */
int fv2d_params_dump_ini(const char *ini_path, const char *overrides, const char *out_path);

/* replaces InitFunctor(params).init(Q)                              (Init.h:291-358)
 * hostQ: zero-filled on entry by this function, then domain + ghosts are written. */
int fv2d_init_problem(const fv2d_device_params *dev, const fv2d_run_params *run, double *hostQ);

/* Rows [j_first, j_first + nrows) of the array fv2d_init_problem would produce (ghost cells
 * included), layout [f][nrows][Ntx]: lets each rank of a multi-GPU job initialise only its
 * own y-slab. */
int fv2d_init_problem_rows(const fv2d_device_params *dev, const fv2d_run_params *run, int j_first, int nrows,
                           double *hostQ_rows);

/* replaces IOManager::saveSolution(Q, iteration, t) on a host copy of Q
 *                                                                   (IOManager.h:99-282)
 * Writes <output_path>/<filename_out>.h5 (+ .xmf): root attributes, vertex datasets x / y and
 * one group ite_%04d per snapshot (or, with run->multiple_outputs, one .h5/.xmf pair per
 * snapshot).  HDF5 is written by the library's own dependency-free writer.
 * *force_file_truncation is the IOManager member of the same name (IOManager.h:80), in/out:
 * the file is truncated when it is non-zero or iteration == 0, appended to otherwise. */
int fv2d_io_save_solution(const fv2d_device_params *dev, const fv2d_run_params *run, const double *hostQ,
                          int iteration, double t, int *force_file_truncation);

/* replaces IOManager::loadSnapshot(Q)                                (IOManager.h:284-398)
 * Reads run->restart_file ("file.h5" = last iteration, "file.h5:/ite_0005" = that group, or a
 * multiple-outputs snapshot file) into hostQ (zero-filled first), fills the ghost cells
 * (IOManager.h:378-379) and refuses to restart past run->tend (:381-387).  When
 * *force_file_truncation comes back non-zero the caller must re-save the loaded state
 * (IOManager.h:391-395), as the C++ IOManager does. */
int fv2d_io_load_snapshot(const fv2d_device_params *dev, const fv2d_run_params *run, double *hostQ, double *time,
                          int *iteration, int *force_file_truncation);

/* Number of CUDA devices visible to this process (what a host driver without a CUDA dependency of its
 * own needs to place the y-slabs of a multi-GPU run). */
int fv2d_device_count(int *count);

/* ------------------------------------------------------------------ context */

/* Allocates Q and U (zero-filled, like Kokkos Views: main.cpp:33-34) on CUDA device
 * `device`.  time_stepping = FV2D_TS_EULER / FV2D_TS_RK2 (Params::time_stepping,
 * SimInfo.h:471); eps_reset_negative = Params::epsilon_reset_negative (SimInfo.h:491). */
int fv2d_ctx_create(const fv2d_device_params *dev, int time_stepping, double eps_reset_negative, int device,
                    fv2d_ctx **out);

/* Multi-GPU variant: this process owns y-slab `rank` of `nranks` of the global grid described by
 * `dev`.  The Ny rows are dealt out as evenly as they go (the first Ny % nranks slabs own one row
 * more); fv2d_ctx_geometry returns the slab's rows and its offset.  Ghost rows at slab interfaces are filled by
 * fv2d_halo_* below instead of the physical y boundary condition. */
int fv2d_ctx_create_slab(const fv2d_device_params *dev, int time_stepping, double eps_reset_negative, int device,
                         int rank, int nranks, fv2d_ctx **out);
void fv2d_ctx_destroy(fv2d_ctx *ctx);

/* Optional: run all work of this context on an existing CUDA stream (cudaStream_t passed as
 * void*), e.g. the caller's torch stream.  The context never destroys a borrowed stream. */
int fv2d_ctx_set_stream(fv2d_ctx *ctx, void *cuda_stream);
int fv2d_sync(fv2d_ctx *ctx);

/* Geometry of the local slab: out[0..5] = {Ntx, Nty_local, Ny_local, j_global_offset, pitch, lead}. */
int fv2d_ctx_geometry(const fv2d_ctx *ctx, int64_t out[6]);

/* Host <-> device transfer of the (local) arrays, layout [f][Nty_local][Ntx].
 * Replace Kokkos::deep_copy to/from a host mirror (IOManager.h:113-114, 196-197). */
int fv2d_upload_Q(fv2d_ctx *ctx, const double *hostQ);
int fv2d_upload_U(fv2d_ctx *ctx, const double *hostU);
int fv2d_download_Q(fv2d_ctx *ctx, double *hostQ);
int fv2d_download_U(fv2d_ctx *ctx, double *hostU);

/* ------------------------------------------------------------------ operator-level API
 * One entry point per reference operator, same semantics, same order of arithmetic
 * (compiled without FMA contraction: results are bit-identical to the reference build). */

/* primToCons(Q, U, params) over range_tot                           (SimInfo.h:589-600) */
int fv2d_prim_to_cons(fv2d_ctx *ctx);
/* consToPrim(U, Q, params) over range_tot                           (SimInfo.h:576-587) */
int fv2d_cons_to_prim(fv2d_ctx *ctx);
/* checkNegatives(Q, params); counts = {rho<0, P<0, NaN}             (SimInfo.h:602-646) */
int fv2d_check_negatives(fv2d_ctx *ctx, uint64_t counts[3]);
/* BoundaryManager::fillBoundaries(Q)                                (BoundaryConditions.h:82-147) */
int fv2d_fill_boundaries(fv2d_ctx *ctx);
/* ComputeDtFunctor::computeDt(Q, max_dt, t, diag); inv_dt = {hyp, tc, visc} (may be NULL)
 *                                                                   (ComputeDt.h:18-65) */
int fv2d_compute_dt(fv2d_ctx *ctx, double *dt, double inv_dt[3]);
/* UpdateFunctor::computeSlopes(Q)                                   (Update.h:59-91) */
int fv2d_compute_slopes(fv2d_ctx *ctx);
/* UpdateFunctor::computeFluxesAndUpdate(Q, Unew, dt)                (Update.h:93-174) */
int fv2d_compute_fluxes_and_update(fv2d_ctx *ctx, double dt);
/* ThermalConductionFunctor::applyThermalConduction(Q, Unew, dt)     (ThermalConduction.h:36-108) */
int fv2d_apply_thermal_conduction(fv2d_ctx *ctx, double dt);
/* ViscosityFunctor::applyViscosity(Q, Unew, dt)                     (Viscosity.h:27-119) */
int fv2d_apply_viscosity(fv2d_ctx *ctx, double dt);
/* UpdateFunctor::euler_step(Q, Unew, dt)                            (Update.h:176-191) */
int fv2d_euler_step(fv2d_ctx *ctx, double dt);
/* UpdateFunctor::update(Q, Unew, dt): Euler or SSP-RK2, unfused     (Update.h:193-222) */
int fv2d_update(fv2d_ctx *ctx, double dt);

/* ------------------------------------------------------------------ fused hot path
 * One time step of the reference loop body, main.cpp:66-83:
 *     update.update(Q,U,dt); consToPrim(U,Q); checkNegatives(Q); [next computeDt]
 * as one fused sm_100a kernel per Runge-Kutta stage (+ a ghost-fill kernel).  The kernel
 * also reduces the inverse time-steps of the NEW state, so the next dt is already on the
 * device when the step ends. */

/* dt supplied by the host (the reference's calling convention). */
int fv2d_step(fv2d_ctx *ctx, double dt);
/* dt taken from the device-resident value (computed by the previous step, or by
 * fv2d_compute_dt before the first one).  No host synchronisation. */
int fv2d_step_device_dt(fv2d_ctx *ctx);
/* nsteps x fv2d_step_device_dt, no host synchronisation (enqueues only). */
int fv2d_run_steps(fv2d_ctx *ctx, int64_t nsteps);
/* Replays main.cpp:62-84 without IO: steps while t + epsilon < tend, at most max_steps.
 * Returns the number of steps done in *steps_done. */
int fv2d_run_until(fv2d_ctx *ctx, double tend, int64_t max_steps, int64_t *steps_done);

/* Device-resident clock: time, the dt the NEXT step would use, number of steps taken. */
int fv2d_get_time(fv2d_ctx *ctx, double *t, double *next_dt, int64_t *steps);
int fv2d_set_time(fv2d_ctx *ctx, double t);
/* The dt used by each of the last min(n, steps, FV2D_DT_HISTORY) steps, oldest first;
 * returns how many were written in *n_out. */
#define FV2D_DT_HISTORY 4096
int fv2d_get_dt_history(fv2d_ctx *ctx, double *dts, int64_t n, int64_t *n_out);
/* Cumulative checkNegatives counters since context creation / last reset. */
int fv2d_get_negative_counts(fv2d_ctx *ctx, uint64_t counts[3], int reset);
/* Sum over the local domain of rho*dx*dy and E*dx*dy (the reference's only conservation
 * diagnostic, python/plot_energy_evolution.py:28-47). */
int fv2d_integrate_mass_energy(fv2d_ctx *ctx, double *mass, double *energy);

/* 64-bit hash of the raw bits of the conserved state U of the local slab, keyed by GLOBAL cell
 * index and summed modulo 2^64: independent of cell order and of the slab decomposition, so the
 * wrapping sum of the slab hashes of an N-GPU run equals the hash of the same state on one GPU
 * (the multi-GPU contract is bitwise equality; main.cpp:62-84 has no counterpart). */
int fv2d_state_hash(fv2d_ctx *ctx, uint64_t *hash);

/* Measurement hooks (bench.py).  fv2d_profile_enable(ctx, 1): every sweep launch is bracketed by a
 * pair of CUDA events on the context's stream (which also keeps consecutive sweeps from overlapping
 * their prologue with the previous sweep's tail); (ctx, 2): launches are only counted; (ctx, 0): off.
 * fv2d_profile_read synchronises and returns the summed bracketed sweep time (0 unless mode 1), the
 * number of sweep launches and the number of all kernel launches since the last fv2d_profile_enable. */
int fv2d_profile_enable(fv2d_ctx *ctx, int on);
int fv2d_profile_read(fv2d_ctx *ctx, double *sweep_ms, int64_t *sweep_launches, int64_t *total_launches);

/* Test hook: the fused kernel replaces IEEE division / sqrt by MUFU seeds + one third-order
 * correction step.  Evaluates those primitives on host arrays of length n on CUDA device
 * `device`: out_rcp[i] = 1 / a[i], out_cs[i] = sqrt(a[i] / b[i]) as the sweep computes them
 * (States.h:46 speedOfSound with a = gamma0 * P, b = rho). */
int fv2d_debug_math_probe(int device, int64_t n, const double *a, const double *b, double *out_rcp, double *out_cs);

/* Measurement hook: the fp64 pipe's peak on CUDA device `device`, in thread-level DFMA
 * instructions per second (independent dependent-DFMA chains, best of 3 timed launches): the
 * denominator of the second roofline (FP64-pipe utilisation) bench.py reports. */
int fv2d_debug_fp64_peak(int device, double *dfma_per_second);

/* Measurement hook: what the y-slab decomposition costs per step in synchronisation.  Every sweep's
 * first CTA notes how long it waited for the other ranks' CFL mails before it had its dt.
 * last_wait_us: that wait in the last sweep; total_wait_us: accumulated since the last reset;
 * last_busy_us: from dt available to the last CTA done, last sweep.  (Any pointer may be NULL.) */
int fv2d_debug_sync_wait(fv2d_ctx *ctx, double *last_wait_us, double *total_wait_us, double *last_busy_us, int reset);

/* Test hook (host only, needs no GPU): the row runs [first, last) - relative to the first domain row
 * of the slab - into which the persistent sweep cuts a slab of Nx x Ny_local cells on a device with
 * num_sms SMs, in table order (every run is crossed with every strip of 252 columns to give the work
 * items).  runs receives up to max_runs (first, last) pairs, *n_runs the number of runs. */
int fv2d_debug_schedule(int Nx, int Ny_local, int num_sms, int neighbour_lo, int neighbour_hi, int32_t *runs, int max_runs,
                        int *n_runs);

/* Test hook (host only, needs no GPU): the row blocks in which fv2d_advance_host_stream moves a slab of
 * Ny rows with Ng ghost rows a side (block_rows <= 0: the default, 256, or FV2D_STREAM_ROWS).  blocks
 * receives up to max_blocks quadruples (up0, up1, sw0, sw1) of array row indices: block b uploads rows
 * [up0, up1) and then sweeps rows [sw0, sw1); *n_blocks the number of blocks. */
int fv2d_debug_stream_blocks(int Ny, int Ng, int block_rows, int32_t *blocks, int max_blocks, int *n_blocks);

/* Development hook (only in library variants built with -DFV2D_TIMING; FV2D_ERR_ARG otherwise):
 * per CTA of the last sweep, 4 values: clock cycles outside the row loops, inside them, work items
 * processed, total. */
int fv2d_debug_sweep_timing(fv2d_ctx *ctx, int64_t *out, int n);

/* Host-buffer convenience used for end-to-end measurement: upload Q (pinned or pageable
 * host memory), primToCons, computeDt, run `nsteps` fused steps, download Q; dts (may be
 * NULL) receives the dt sequence.  Equivalent to the reference main.cpp:58-84 on a state
 * that lives on the host. */
int fv2d_advance_host(fv2d_ctx *ctx, const double *hostQ_in, double *hostQ_out, int64_t nsteps, double *dts);

/* One time step on a state that lives on the host, with the transfers overlapped: the loop body of
 * main.cpp:62-84 (computeDt, UpdateFunctor::update, consToPrim, checkNegatives, t += dt) for a
 * caller who holds Q in host memory (pinned memory for the overlap; pageable memory works, slower).
 * dt_hint: the caller's idea of the time step of hostQ_in - normally *dt_next of the previous call
 *   (UpdateFunctor::update takes dt as an argument too, Update.h:193) - or <= 0 for "none".
 * With a hint on a single-slab forward-Euler context the state moves in row blocks: block b+1 is
 * uploaded while block b is swept with dt_hint and block b-1 is downloaded (full-duplex PCIe), and
 * the CFL time step of the uploaded state is evaluated on the way.  If it equals dt_hint bit for bit
 * the step stands (*streamed = 1).  If not - or without a hint - the step is (re)done with the
 * state's own time step, one transfer after the other (*streamed = 0).  Either way hostQ_out, *dt_used
 * and *dt_next (the time step of the NEW state: the next call's hint) are the same bits: the hint
 * changes how long the call takes, never what it returns.
 * The time step is CFL / max(inverse time steps) as ComputeDt.h:18-65, with the sound speed in the
 * fused sweep's arithmetic (<= 4 ulp from sqrt): *dt_used can differ from fv2d_compute_dt's in the last
 * bits, like every dt of fv2d_run_steps after the first.  dt_used / dt_next / streamed may be NULL.
 * hostQ_out may be hostQ_in (in place): the call then takes the serial route. */
int fv2d_advance_host_stream(fv2d_ctx *ctx, const double *hostQ_in, double *hostQ_out, double dt_hint, double *dt_used,
                             double *dt_next, int *streamed);

/* ------------------------------------------------------------------ multi-GPU halo exchange
 * (no reference counterpart: the reference is single-device.)  Rows are exchanged over
 * peer memory (CUDA IPC): every rank publishes a handle to its Q buffers, opens its two
 * y-neighbours' handles, and the stage epilogue's edge rows are written straight into the
 * neighbour's ghost rows. */
#define FV2D_IPC_HANDLE_BYTES 512
/* Fills `handle` (FV2D_IPC_HANDLE_BYTES bytes) describing this rank's exchange buffers. */
int fv2d_halo_export(fv2d_ctx *ctx, void *handle);
/* handles: nranks consecutive handles gathered from all ranks (e.g. torch.distributed
 * all_gather), index = rank. */
int fv2d_halo_connect(fv2d_ctx *ctx, const void *handles, int nranks);
/* (Handles from contexts of the SAME process are recognised and connected through plain
 * peer access instead of IPC.)  All ranks must stay alive, and must not destroy their
 * context, while any rank is still stepping.
 * Global dt across ranks: the last CTA of each rank's final-stage sweep writes the slab's
 * maximum inverse dt into every rank's mailbox; the next step's prologue reduces them on the
 * device.  Nothing to call per step.
 * fv2d_get_inv_dt: the three inverse-dt maxima {hyp, tc, visc} of the CURRENT state. */
int fv2d_get_inv_dt(fv2d_ctx *ctx, double inv_dt[3]);

#ifdef __cplusplus
}
#endif
#endif /* FV2D_B200_H_ */
