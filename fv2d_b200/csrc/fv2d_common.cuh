// fv2d_common.cuh — shared declarations of the CUDA side: device array layout, kernel
// parameter block, context, error helpers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/fv2d_b200.h"

namespace fv2d
{

// ---------------------------------------------------------------------------------------
// Device array layout.  One "array" (the reference's View<double***>(Nty,
// Ntx, 4), main.cpp:33-34) is 4 SoA planes.  A plane is Nty rows of `pitch` doubles; cell
// (i, j) sits at column lead + i, where lead is chosen so that the first DOMAIN cell
// (i = ibeg) starts a 128-byte line, and pitch is a multiple of 16 doubles so every row
// starts one too.  => warp-wide stores of domain cells are whole 128-byte lines, and the
// row pitch satisfies the 16-byte stride rule of TMA tensor maps.
// ---------------------------------------------------------------------------------------
struct Layout
{
  int pitch;       // doubles per row
  int lead;        // column of cell i = 0
  int rows;        // Nty of the local slab
  long long plane; // doubles per field plane (= pitch * rows)
  __host__ __device__ inline long long at(int f, int i, int j) const
  {
    return (long long)f * plane + (long long)j * pitch + lead + i;
  }
};

constexpr int kMaxRanks = 8; // y-slabs per job: the GPUs of one NVSwitch box
// One unit of work of the persistent sweep (fv2d_sweep.cu): rows [j0, j1) of strip `strip`.
// j0 < 0 marks the end of the table (entry n_items).  16 bytes: one cp.async into the CTA's
// shared-memory queue.
struct WorkItem
{
  int strip, j0, j1, pad;
};

// What lies beyond the low-j / high-j edge of the local slab.
enum : int
{
  EDGE_PHYSICAL = 0, // the global boundary: apply boundary_y
  EDGE_NEIGHBOUR = 1 // another rank's slab: ghost rows come from the halo exchange
};

// Device-resident scalars of one context (one 128-byte-aligned block).
struct DevScalars
{
  double dt;                        // dt of the step being / about to be taken
  double t;                         // simulated time
  long long step;                   // steps taken
  unsigned long long inv_acc[2][4]; // order-preserving-encoded maxima {hyp, tc, visc, -} ; [parity]
  unsigned long long neg[4];        // cumulative {rho<0, P<0, NaN, -}
  double inv_dt_last[4];            // decoded maxima behind `dt`
  double sums[2];                   // scratch for mass/energy integration
  unsigned long long hash;          // scratch for fv2d_state_hash
  unsigned long long tstamp[4];     // %globaltimer (ns) of the last sweep: CTA 0 started, CTA 0 had its dt (all CFL
                                    // mails in), last CTA done; [3] = accumulated wait for the mails since reset
  // ---- multi-GPU mailboxes: written by peers over NVLink (system-scope stores), read locally
  unsigned long long halo_cnt[2];      // ghost-row pushes received from the low-j / high-j neighbour
  unsigned long long mail_gen[kMaxRanks]; // per source rank: generation of its last CFL mail
  double mail_inv[2][kMaxRanks];       // [generation parity][source rank]: that rank's max inverse dt
  unsigned int cta_done;               // CTAs of the running sweep that have finished (local)
  unsigned int fault;                  // set if a wait on a peer timed out
  unsigned int work_next;              // work items handed out beyond the static first one per CTA (local)
  unsigned int stream_ok;              // streamed host step: 1 = the caller's dt was the state's own, step committed
  double dt_next;                      // ... and the time step of the state it produced
  unsigned long long neg_save[4];      // ... `neg` as it was when the step began (restored if not committed)
  double dt_hist[FV2D_DT_HISTORY];  // ring of dts used
};

// Parameter block handed to every kernel by value.
struct KParams
{
  fv2d_device_params p; // LOCAL slab view: Ny/Nty/jbeg/jend describe the slab
  Layout L;
  int edge_lo, edge_hi;  // EDGE_* of the slab's low-j / high-j side
  int j_global_offset;   // global row index of local row 0
  int Ny_global;
  double eps_reset;
  const double *gtab;    // per-local-row analytical gravity (float-valued), or nullptr
  DevScalars *sc;
  int rank, nranks;
  DevScalars *peer_sc[kMaxRanks]; // every rank's scalars, mapped into this device (self included)
};

// Does the slab hold the first / last row of the GLOBAL grid?  The reference ties the well-balanced
// boundary flux (Update.h:148-156) and the conduction boundary values (ThermalConduction.h:77-103) to
// the rows j == jbeg / j == jend - 1 whatever the boundary type - also with a periodic y boundary,
// where the slab's edge is a neighbour slab (EDGE_NEIGHBOUR) and not a physical edge.
__host__ __device__ inline bool holds_global_first_row(const KParams &kp) { return kp.j_global_offset == 0; }
__host__ __device__ inline bool holds_global_last_row(const KParams &kp) { return kp.j_global_offset + kp.p.Ny == kp.Ny_global; }

// Monotone map double -> uint64 so that atomicMax on the integer orders like the double.
__host__ __device__ inline unsigned long long encode_ordered(double x)
{
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
  unsigned long long b;
  memcpy(&b, &x, sizeof b);
#endif
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ inline double decode_ordered(unsigned long long e)
{
  unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, sizeof x);
  return x;
#endif
}
// encode_ordered(-DBL_MAX): identity of the Max reducers (Kokkos::Max, ComputeDt.h:50-52)
#define FV2D_ENC_NEG_MAX 0x0010000000000000ULL
constexpr int kProfMax = 8192;

#ifdef __CUDACC__
// ---- system-scope (cross-GPU, over NVLink peer mappings) flag primitives
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v)
{
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p)
{
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *p >= v.  A peer that never answers must not hang the GPU: after ~20 s the wait
// gives up and raises the context's fault flag (every host-synchronising call reports it).
__device__ __forceinline__ bool wait_ge_sys(const unsigned long long *p, unsigned long long v, DevScalars *sc)
{
  if (ld_acquire_sys(p) >= v)
    return true;
  if (*(volatile unsigned int *)&sc->fault) // an earlier wait already gave up: do not spend another 20 s
    return false;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(p) < v)
  {
    __nanosleep(64);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ULL)
    {
      sc->fault = 1;
      return false;
    }
  }
  return true;
}

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Posts this rank's maximum inverse time-step to rank q's mailbox and stamps it with generation
// `gen`: value, system-scope fence, generation (the fence-release pattern: one NVLink round trip).
// The sweep calls it from lanes 0 .. nranks-1 of one warp, so that all mailboxes are written at
// once: eight release stores issued one after the other by a single thread cost eight round trips,
// on the critical path of every rank's next step.
__device__ __forceinline__ void post_cfl_mail_to(const KParams &kp, int q, double hyp, unsigned long long gen)
{
  st_relaxed_sys_f64(&kp.peer_sc[q]->mail_inv[gen & 1][kp.rank], hyp);
  __threadfence_system();
  st_relaxed_sys_u64(&kp.peer_sc[q]->mail_gen[kp.rank], gen);
}
// Single slab: the mailbox is this device's own memory - device scope is enough for the hand-over
// between two launches on one stream (the system-scope fence costs a few microseconds per step).
__device__ __forceinline__ void post_cfl_mail_local(const KParams &kp, double hyp, unsigned long long gen)
{
  kp.sc->mail_inv[gen & 1][0] = hyp;
  __threadfence();
  *(volatile unsigned long long *)&kp.sc->mail_gen[0] = gen;
}
// ... to every rank's mailbox (self included), by one thread (the stand-alone computeDt).
__device__ __forceinline__ void post_cfl_mail(const KParams &kp, double hyp, unsigned long long gen)
{
  for (int q = 0; q < kp.nranks; ++q)
    st_relaxed_sys_f64(&kp.peer_sc[q]->mail_inv[gen & 1][kp.rank], hyp);
  __threadfence_system();
  for (int q = 0; q < kp.nranks; ++q)
    st_relaxed_sys_u64(&kp.peer_sc[q]->mail_gen[kp.rank], gen);
}
// Waits for generation `gen` of every rank's mail and returns the global maximum.
__device__ __forceinline__ double collect_cfl_mail(const KParams &kp, unsigned long long gen)
{
  double m = -1.7976931348623157e308;
  for (int q = 0; q < kp.nranks; ++q)
  {
    wait_ge_sys(&kp.sc->mail_gen[q], gen, kp.sc);
    m = fmax(m, ld_relaxed_sys_f64(&kp.sc->mail_inv[gen & 1][q]));
  }
  return m;
}
#endif

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FV2D_CUDA(call)                                                  \
  do                                                                     \
  {                                                                      \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess)                                              \
      return ::fv2d::cuda_fail(e__, #call, __FILE__, __LINE__);          \
  } while (0)

} // namespace fv2d

// The context behind the opaque C handle.
struct fv2d_ctx
{
  fv2d::KParams kp;        // local parameters + layout
  fv2d_device_params glob; // global grid
  int time_stepping;
  int device;
  int num_sms;
  int rank, nranks;
  cudaStream_t stream;
  bool own_stream;

  double *Q[2]; // ping-pong primitive arrays; Q[cur] is "the" Q
  int cur;
  double *U;      // conservative
  double *Ustar;  // RK2 stage buffer (lazy)
  double *slopesX, *slopesY; // operator-level API only (lazy)
  double *gtab;   // analytical gravity table (or nullptr)
  double *rowsum; // scratch of fv2d_integrate_mass_energy (lazy)
  fv2d::DevScalars *sc;
  fv2d::DevScalars *sc_host; // pinned mirror for readbacks
  double *stage_host;        // pinned staging row buffer for uploads/downloads (lazy)
  size_t stage_bytes;

  CUtensorMap tmapQ[2]; // TMA descriptors of Q[0], Q[1]
  CUtensorMap tmapU, tmapUstar; // ... of U and of the RK2 stage array (valid once Ustar exists)
  CUtensorMap *tmaps_dev; // device copies of the store descriptors: [0] U, [1] Ustar
  bool tmap_ok;
  // persistent sweep: work-item table (device), its length, grid size
  fv2d::WorkItem *items_dev;
  int n_items, n_ctas;
  bool persistent; // every item has >= 8 rows: CTAs take several (else one CTA per item)
  bool fold_ok;      // the sweep can write the ghost cells itself (every ghost mirrors a DOMAIN cell)
  bool ghosts_valid; // the ghosts of Q[cur] are up to date
  bool dt_valid;     // the device-resident CFL maximum describes Q[cur]

  // optional profiling: CUDA event pairs around every sweep launch
  bool profile;
  cudaEvent_t *prof_ev; // 2 * kProfMax events
  int prof_n;           // pairs recorded
  long long n_launch_sweep, n_launch_total;

  // streamed host path (fv2d_advance_host_stream): row blocks, their work tables, copy streams
  struct StreamBlock
  {
    int up0, up1;     // array rows uploaded with the block
    int sw0, sw1;     // domain rows swept once the block is resident (array row indices)
    int item_off, n_items, n_ctas, persistent;
  };
  StreamBlock *sblocks;
  int n_sblocks;
  fv2d::WorkItem *sitems_dev;
  cudaStream_t s_up, s_dn;
  double *dense_in, *dense_out; // device copies of the host arrays in the host's layout (flat PCIe copies)
  cudaEvent_t *s_ev; // 2 per block (uploaded, swept) + 2

  // multi-GPU peers (slab below / above): the neighbours' Q[0], Q[1] mapped into this device
  double *peerQ_lo[2], *peerQ_hi[2];
  int peer_lo_Ny;                          // rows the low neighbour owns (its high ghost rows start at Ng + that)
  long long peer_lo_plane, peer_hi_plane;  // plane strides of the neighbours' arrays (slabs may differ by a row)
  void *ipc_opened[24];
  int n_ipc_opened;
  bool connected;
  unsigned long long halo_gen; // sweeps whose edge rows have been pushed so far (same on all ranks)
  unsigned long long mail_gen; // CFL reductions done so far (same on all ranks)
};
