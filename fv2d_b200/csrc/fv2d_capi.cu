// fv2d_capi.cu — the C ABI (include/fv2d_b200.h): context management, host<->device
// transfers, the operator-level entry points and the fused step driver.
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

#include <unistd.h>

#include <nvtx3/nvToolsExt.h> // header-only NVTX3: no link dependency, no cost unless a tool is attached

#include "../host/Init.h"
#include "../host/SimInfo.h"
#include "../host/SnapshotIO.h"
#include "fv2d_kernels.h"

namespace fv2d
{

// NVTX range named like the reference's Kokkos kernel label (what Kokkos-tools / Nsight show for the
// reference: Update.h:67,102,215, ComputeDt.h:27, BoundaryConditions.h:89,119, ThermalConduction.h:43,
// Viscosity.h:34, SimInfo.h:580,593,611), so a timeline of this library reads like one of the reference.
struct NvtxRange
{
  explicit NvtxRange(const char *label) { nvtxRangePushA(label); }
  ~NvtxRange() { nvtxRangePop(); }
};

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
  std::ostringstream os;
  os << "CUDA error " << (int)e << " (" << cudaGetErrorString(e) << ") in " << what << " at " << file << ":" << line;
  set_error(os.str());
  return FV2D_ERR_CUDA;
}
static int arg_fail(const std::string &msg)
{
  set_error(msg);
  return FV2D_ERR_ARG;
}

static IniOverrides parse_overrides(const char *ov)
{
  IniOverrides m;
  if (!ov)
    return m;
  std::string s(ov);
  size_t pos = 0;
  while (pos < s.size())
  {
    size_t end = s.find(';', pos);
    if (end == std::string::npos)
      end = s.size();
    std::string item = s.substr(pos, end - pos);
    pos              = end + 1;
    size_t eq = item.find('='), dot = item.find('.');
    if (eq == std::string::npos || dot == std::string::npos || dot > eq)
      continue;
    auto trim = [](std::string x) {
      size_t a = x.find_first_not_of(" \t"), b = x.find_last_not_of(" \t");
      return a == std::string::npos ? std::string() : x.substr(a, b - a + 1);
    };
    m[IniFile::MakeKey(trim(item.substr(0, dot)), trim(item.substr(dot + 1, eq - dot - 1)))] = trim(item.substr(eq + 1));
  }
  return m;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda,
// so the library still loads on a machine without a driver).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap(CUtensorMap *out, double *base, const Layout &L, int box_cols, int ncols = 0)
{
  static PFN_encodeTiled fn = nullptr;
  if (!fn)
  {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FV2D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess)
    {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return FV2D_ERR_CUDA;
    }
    fn = (PFN_encodeTiled)p;
  }
  // 3-D tensor: (column, row, field), fp64, row pitch and plane stride in bytes
  // (ncols: a store descriptor is cut off after the last domain column, so that the box of the
  // last strip does not write the ghost columns and the padding behind it)
  cuuint64_t dims[3]    = {(cuuint64_t)(ncols ? ncols : L.pitch), (cuuint64_t)L.rows, 4};
  cuuint64_t strides[2] = {(cuuint64_t)L.pitch * sizeof(double), (cuuint64_t)L.plane * sizeof(double)};
  cuuint32_t box[3]     = {(cuuint32_t)box_cols, 1, 4};
  cuuint32_t estr[3]    = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
  {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return FV2D_ERR_CUDA;
  }
  return FV2D_OK;
}

static size_t array_bytes(const Layout &L) { return (size_t)4 * L.plane * sizeof(double); }

// host [f][rows][Ntx] <-> device padded planes
static int copy_h2d(fv2d_ctx *c, double *dev, const double *host)
{
  const Layout &L = c->kp.L;
  const int Ntx   = c->kp.p.Ntx;
  for (int f = 0; f < 4; ++f)
    FV2D_CUDA(cudaMemcpy2DAsync(dev + f * L.plane + L.lead, (size_t)L.pitch * sizeof(double),
                                host + (size_t)f * L.rows * Ntx, (size_t)Ntx * sizeof(double),
                                (size_t)Ntx * sizeof(double), (size_t)L.rows, cudaMemcpyHostToDevice, c->stream));
  return FV2D_OK;
}
static int copy_d2h(fv2d_ctx *c, double *host, const double *dev)
{
  const Layout &L = c->kp.L;
  const int Ntx   = c->kp.p.Ntx;
  for (int f = 0; f < 4; ++f)
    FV2D_CUDA(cudaMemcpy2DAsync(host + (size_t)f * L.rows * Ntx, (size_t)Ntx * sizeof(double),
                                dev + f * L.plane + L.lead, (size_t)L.pitch * sizeof(double),
                                (size_t)Ntx * sizeof(double), (size_t)L.rows, cudaMemcpyDeviceToHost, c->stream));
  return FV2D_OK;
}

static int check_supported(const fv2d_device_params &p)
{
  if (p.Nx < 1 || p.Ny < 1)
    return arg_fail("Nx and Ny must be positive");
  if (p.Ng < 2)
    return arg_fail("Nghosts >= 2 is required (stencil radius 2)");
  if (p.Ntx != p.Nx + 2 * p.Ng || p.Nty != p.Ny + 2 * p.Ng || p.ibeg != p.Ng || p.jbeg != p.Ng ||
      p.iend != p.Ng + p.Nx || p.jend != p.Ng + p.Ny)
    return arg_fail("inconsistent mesh extents in fv2d_device_params");
  if (p.thermal_conductivity_active && p.thermal_conductivity_mode != FV2D_TCM_CONSTANT)
    return arg_fail("thermal conductivity mode B02 is undefined behaviour in the reference (its parameters are never "
                    "read) and is not supported");
  if ((p.boundary_x == FV2D_BC_PERIODIC && p.Nx < p.Ng) || (p.boundary_y == FV2D_BC_PERIODIC && p.Ny < p.Ng))
    return arg_fail("periodic direction narrower than the ghost layer");
  return FV2D_OK;
}

// Every host-synchronising entry point goes through here: besides the stream it checks the
// context's fault flag, which a kernel raises when a wait on a peer GPU (halo rows, CFL mail) timed
// out.  From then on the sweeps run with dt = NaN, so a faulted state can not be mistaken for a result.
static int sync_ctx(fv2d_ctx *c)
{
  if (c->nranks > 1) // the flag can only be raised by a cross-GPU wait
    FV2D_CUDA(cudaMemcpyAsync(&c->sc_host->fault, &c->sc->fault, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  FV2D_CUDA(cudaStreamSynchronize(c->stream));
  if (c->sc_host->fault)
  {
    set_error("a wait on a peer GPU timed out (halo rows or CFL mail never arrived)");
    return FV2D_ERR_CUDA;
  }
  return FV2D_OK;
}

static int read_scalars(fv2d_ctx *c)
{
  FV2D_CUDA(cudaMemcpyAsync(c->sc_host, c->sc, offsetof(DevScalars, dt_hist), cudaMemcpyDeviceToHost, c->stream));
  return sync_ctx(c);
}

// The hyperbolic maximum of the CURRENT state, reduced over all slabs (latest mail generation).
// Peers post their mail from their own stream: after synchronising only the local stream a slot
// can still hold the value of two generations ago, so the scalars are re-read until every rank's
// generation has arrived (bounded: a peer that never posts is a fault).
static int current_hyp(fv2d_ctx *c, double *out)
{
  for (int tries = 0;; ++tries)
  {
    bool all = true;
    for (int q = 0; q < c->nranks; ++q)
      all = all && c->sc_host->mail_gen[q] >= c->mail_gen;
    if (all)
      break;
    if (tries > 100000)
    {
      set_error("the CFL mail of a peer GPU never arrived");
      return FV2D_ERR_CUDA;
    }
    usleep(50);
    int rc = read_scalars(c);
    if (rc)
      return rc;
  }
  double m = -1.7976931348623157e308;
  for (int q = 0; q < c->nranks; ++q)
    m = std::fmax(m, c->sc_host->mail_inv[c->mail_gen & 1][q]);
  *out = m;
  return FV2D_OK;
}

// ------------------------------------------------------------------ step drivers

static int ensure_slopes(fv2d_ctx *c)
{
  if (c->slopesX)
    return FV2D_OK;
  const size_t bytes = array_bytes(c->kp.L);
  FV2D_CUDA(cudaMalloc(&c->slopesX, bytes));
  FV2D_CUDA(cudaMalloc(&c->slopesY, bytes));
  FV2D_CUDA(cudaMemsetAsync(c->slopesX, 0, bytes, c->stream)); // zero-filled like Update.h:54-55
  FV2D_CUDA(cudaMemsetAsync(c->slopesY, 0, bytes, c->stream));
  return FV2D_OK;
}
// device copy of the store descriptor of U (slot 0) or Ustar (slot 1)
static int upload_store_tmap(fv2d_ctx *c, int slot, double *base)
{
  CUtensorMap m;
  int rc = make_tmap(&m, base, c->kp.L, sweep_strip_width(), c->kp.L.lead + c->kp.p.iend);
  if (rc)
    return rc;
  if (!c->tmaps_dev)
    FV2D_CUDA(cudaMalloc(&c->tmaps_dev, 2 * sizeof(CUtensorMap)));
  FV2D_CUDA(cudaMemcpyAsync(c->tmaps_dev + slot, &m, sizeof m, cudaMemcpyHostToDevice, c->stream));
  FV2D_CUDA(cudaStreamSynchronize(c->stream));
  return FV2D_OK;
}

static int ensure_ustar(fv2d_ctx *c)
{
  if (c->Ustar)
    return FV2D_OK;
  const size_t bytes = array_bytes(c->kp.L);
  FV2D_CUDA(cudaMalloc(&c->Ustar, bytes));
  FV2D_CUDA(cudaMemsetAsync(c->Ustar, 0, bytes, c->stream));
  int rc = upload_store_tmap(c, 1, c->Ustar);
  return rc ? rc : make_tmap(&c->tmapUstar, c->Ustar, c->kp.L, sweep_strip_width());
}

// Update.h:176-191 with the operator-level kernels
static int euler_step_ops(fv2d_ctx *c, double *Q, double *Unew, double dt)
{
  int rc;
  {
    NvtxRange r("Filling X-boundary + Filling Y-boundary");
    launch_fill_boundaries(c->kp, Q, c->stream);
  }
  if (c->kp.p.reconstruction == FV2D_PLM)
  {
    if ((rc = ensure_slopes(c)))
      return rc;
    NvtxRange r("Slopes");
    launch_compute_slopes(c->kp, Q, c->slopesX, c->slopesY, c->stream);
  }
  // (PCM never reads the slope arrays: Update.h:27-33)
  {
    NvtxRange r("Update");
    launch_fluxes_and_update(c->kp, Q, c->slopesX, c->slopesY, Unew, dt, c->stream);
  }
  if (c->kp.p.thermal_conductivity_active)
  {
    NvtxRange r("Thermal conduction");
    launch_thermal_conduction(c->kp, Q, Unew, dt, c->stream);
  }
  if (c->kp.p.viscosity_active)
  {
    NvtxRange r("Viscosity");
    launch_viscosity(c->kp, Q, Unew, dt, c->stream);
  }
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}

// Work-item table of the persistent sweep (fv2d_sweep.cu).  The slab is cut into strips of W
// columns and runs of rows; one item = (strip, run).  All strips share the same runs and the table
// is ordered run by run, strips fastest, so the CTAs working at any moment cover a band of
// neighbouring rows (halo columns are shared through L2, few DRAM pages are open).  Run heights
// follow a guided schedule made of ROUNDS: a round is as many runs as give every CTA slot one item
// (#slots / #strips), all of the same height = 1/C of the rows each slot still has to do, capped at
// `hmax`, never below `hmin`.  Equal heights inside a round keep the CTAs in step (an item is bound
// to a CTA one item ahead, so unequal neighbours in the table would not be rebalanced); halving from
// round to round lets the last CTAs finish within a few rows of each other; tall runs while there is
// plenty of work keep the per-item cost (one warm-up row + ~2 us) around 1 %.  With a neighbour
// slab the runs at that edge come first and are short: their rows travel over NVLink while the rest
// computes.
static std::vector<std::pair<int, int>> schedule_runs(int Ny, int nstrips, int slots, bool nb_lo, bool nb_hi)
{
  auto env_int = [](const char *name, int dflt) {
    const char *e = std::getenv(name);
    return (e && std::atoi(e) > 0) ? std::atoi(e) : dflt;
  };
  const int fixed   = env_int("FV2D_CHUNK_ROWS", 0);
  const int hmax = env_int("FV2D_SCHED_HMAX", 96), hmin = std::min(env_int("FV2D_SCHED_HMIN", 8), hmax);
  const double C = env_int("FV2D_SCHED_C100", 220) / 100.0;
  const int runs_per_round = std::max(1, (slots + nstrips / 2) / nstrips);
  std::vector<std::pair<int, int>> runs;
  int lo = 0, hi = Ny;
  if (std::getenv("FV2D_FORCE_EDGE_RUNS")) // development: the schedule of a slab with neighbours
    nb_lo = nb_hi = true;
  if (nb_lo && hi - lo >= 2 * hmin)
  {
    runs.emplace_back(lo, lo + hmin);
    lo += hmin;
  }
  if (nb_hi && hi - lo >= 2 * hmin)
  {
    runs.emplace_back(hi - hmin, hi);
    hi -= hmin;
  }
  while (lo < hi)
  {
    int h = fixed ? fixed : (int)((double)(hi - lo) * nstrips / (slots * C));
    h     = std::max(hmin, std::min(hmax, h));
    for (int r = 0; r < runs_per_round && lo < hi; ++r)
    {
      int hr = h;
      if (hi - lo - hr < hmin)
        hr = hi - lo; // no sliver at the end
      runs.emplace_back(lo, lo + hr);
      lo += hr;
    }
  }
  return runs;
}

static int build_work_items(fv2d_ctx *c)
{
  const int W = sweep_strip_width(), Ny = c->kp.p.Ny, jbeg = c->kp.p.jbeg;
  const int nstrips = (c->kp.p.Nx + W - 1) / W;
  const int slots   = 2 * c->num_sms;
  const std::vector<std::pair<int, int>> runs =
      schedule_runs(Ny, nstrips, slots, c->kp.edge_lo == EDGE_NEIGHBOUR, c->kp.edge_hi == EDGE_NEIGHBOUR);
  std::vector<WorkItem> items;
  items.reserve(runs.size() * nstrips + 1);
  for (const auto &r : runs)
    for (int s = 0; s < nstrips; ++s)
      items.push_back(WorkItem{s, jbeg + r.first, jbeg + r.second, 0});
  c->n_items = (int)items.size();
  // The persistent sweep stages rows of the NEXT item while it finishes the current one and never
  // looks further: that needs items of at least 8 rows (more than the deepest ring).  Degenerate
  // grids / development chunk heights below that run one CTA per item instead.
  int min_rows = Ny;
  for (const auto &r : runs)
    min_rows = std::min(min_rows, r.second - r.first);
  c->persistent = min_rows >= 8;
  c->n_ctas     = c->persistent ? std::min(c->n_items, slots) : c->n_items;
  if (const char *e = std::getenv("FV2D_MAX_CTAS")) // test / sanitizer knob: few CTAs, many items each
    if (c->persistent && std::atoi(e) > 0)
      c->n_ctas = std::min(c->n_ctas, std::atoi(e));
  items.push_back(WorkItem{0, -1, -1, 0}); // end marker
  FV2D_CUDA(cudaMalloc(&c->items_dev, items.size() * sizeof(WorkItem)));
  FV2D_CUDA(cudaMemcpyAsync(c->items_dev, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, c->stream));
  FV2D_CUDA(cudaStreamSynchronize(c->stream));
  return FV2D_OK;
}

static unsigned long long pushes_per_sweep(const fv2d_ctx *c)
{
  // every strip pushes its Ng edge rows to each neighbour once per sweep
  return (unsigned long long)c->kp.p.Ng * ((c->kp.p.Nx + sweep_strip_width() - 1) / sweep_strip_width());
}

static int compute_dt_now(fv2d_ctx *c);

// One fused time step (Euler: 1 sweep; RK2: 2 sweeps), dt either from the host or from the
// device-resident CFL maximum of the current state.  In steady state a stage is ONE launch: the
// sweep takes its dt from the device, writes the ghost cells of its output and advances the clock.
static int fused_step(fv2d_ctx *c, bool device_dt, double dt_host)
{
  if (!c->tmap_ok)
    return arg_fail("TMA descriptors unavailable");
  if (c->nranks > 1 && !c->connected)
    return arg_fail("multi-GPU context: call fv2d_halo_connect before stepping");
  int rc;
  if (device_dt && !c->dt_valid && (rc = compute_dt_now(c)))
    return rc;
  const int cur = c->cur, nxt = cur ^ 1;
  const unsigned long long pps = pushes_per_sweep(c);
  auto fill_ghosts = [&](double *Q) {
    NvtxRange r("Filling X-boundary + Filling Y-boundary");
    launch_fill_ghosts(c->kp, Q, c->halo_gen * pps, c->stream);
    c->n_launch_total++;
  };
  if (!c->ghosts_valid || !c->fold_ok)
    fill_ghosts(c->Q[cur]);
  // one range per step; the fused sweep stands for the reference's whole kernel chain
  NvtxRange step_range("Slopes + Update + Thermal conduction + Viscosity + Conservative to Primitive + Check negative "
                       "density/pressure + Computing DT (fused sweep)");
  auto prof_mark = [&](int which) {
    if (c->profile && c->prof_n < kProfMax)
      cudaEventRecord(c->prof_ev[2 * c->prof_n + which], c->stream);
    if (which == 1)
    {
      c->n_launch_sweep++;
      c->n_launch_total++;
      if (c->profile && c->prof_n < kProfMax)
        c->prof_n++;
    }
  };

  SweepArgs a;
  std::memset(&a, 0, sizeof a);
  a.kp            = c->kp;
  a.use_device_dt = device_dt ? 1 : 0;
  a.dt_host       = dt_host;
  a.fold_ghosts   = c->fold_ok ? 1 : 0;
  a.items         = c->items_dev;
  a.n_items       = c->n_items;
  a.n_ctas        = c->n_ctas;
  a.persistent    = c->persistent ? 1 : 0;
  a.lo_rank       = (c->rank + c->nranks - 1) % c->nranks;
  a.hi_rank       = (c->rank + 1) % c->nranks;
  a.mail_gen      = c->mail_gen; // the final stage of this step posts the next generation
  auto set_peers = [&](int qout) {
    a.peer_lo_Qout  = (c->kp.edge_lo == EDGE_NEIGHBOUR) ? c->peerQ_lo[qout] : nullptr;
    a.peer_hi_Qout  = (c->kp.edge_hi == EDGE_NEIGHBOUR) ? c->peerQ_hi[qout] : nullptr;
    a.halo_expected = c->halo_gen * pps;
    a.peer_lo_Ny = c->peer_lo_Ny, a.peer_lo_plane = c->peer_lo_plane, a.peer_hi_plane = c->peer_hi_plane;
  };
  cudaError_t e;
  if (c->time_stepping == FV2D_TS_RK2)
  {
    if ((rc = ensure_ustar(c)))
      return rc;
    // stage 1: U* = U + dt L(Q), Q* = consToPrim(U*)           (Update.h:204-210)
    a.Uin = c->U, a.Uout = c->Ustar, a.U0 = nullptr, a.Qout = c->Q[nxt], a.final_stage = 0;
    a.tm_store_u = c->tmaps_dev + 1;
    set_peers(nxt);
    prof_mark(0);
    e = launch_sweep(c->tmapQ[cur], c->tmapU, a, c->stream);
    prof_mark(1);
    if (e != cudaSuccess)
      return cuda_fail(e, "sweep stage 1", __FILE__, __LINE__);
    c->halo_gen++;
    // ghosts of Q* (Update.h:211 -> :179): written by stage 1 itself unless the grid is degenerate
    if (!c->fold_ok)
      fill_ghosts(c->Q[nxt]);
    // stage 2: U = 0.5 (U0 + U* + dt L(Q*)), Q = consToPrim(U)    (Update.h:211-220, main.cpp:80-81)
    a.Uin = c->Ustar, a.Uout = c->U, a.U0 = c->U, a.Qout = c->Q[cur], a.final_stage = 1;
    a.tm_store_u = c->tmaps_dev;
    set_peers(cur);
    prof_mark(0);
    e = launch_sweep(c->tmapQ[nxt], c->tmapUstar, a, c->stream);
    prof_mark(1);
    if (e != cudaSuccess)
      return cuda_fail(e, "sweep stage 2", __FILE__, __LINE__);
    // Q[cur] holds the new state again
  }
  else
  {
    a.Uin = c->U, a.Uout = c->U, a.U0 = nullptr, a.Qout = c->Q[nxt], a.final_stage = 1;
    a.tm_store_u = c->tmaps_dev;
    set_peers(nxt);
    prof_mark(0);
    e = launch_sweep(c->tmapQ[cur], c->tmapU, a, c->stream);
    prof_mark(1);
    if (e != cudaSuccess)
      return cuda_fail(e, "sweep", __FILE__, __LINE__);
    c->cur = nxt;
  }
  c->halo_gen++;
  c->mail_gen++;
  c->ghosts_valid = c->fold_ok;
  c->dt_valid     = true;
  return FV2D_OK;
}

// standalone computeDt of the current state into inv_acc[0] (the sweeps use inv_acc[1]) and sc->dt;
// posts the next generation of the CFL mail, which is what the next device-dt step reads
static int compute_dt_now(fv2d_ctx *c)
{
  if (c->nranks > 1 && !c->connected)
    return arg_fail("multi-GPU context: call fv2d_halo_connect before computing dt");
  static const unsigned long long init = FV2D_ENC_NEG_MAX;
  FV2D_CUDA(cudaMemcpyAsync(&c->sc->inv_acc[0][0], &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  NvtxRange r("Computing DT");
  launch_compute_dt(c->kp, c->Q[c->cur], &c->sc->inv_acc[0][0], c->stream);
  c->mail_gen++;
  launch_finalize_dt(c->kp, &c->sc->inv_acc[0][0], c->mail_gen, c->stream);
  c->n_launch_total += 2;
  FV2D_CUDA(cudaGetLastError());
  c->dt_valid = true;
  return FV2D_OK;
}

} // namespace fv2d

using namespace fv2d;

// ====================================================================================== C ABI

extern "C" {

const char *fv2d_last_error(void) { return g_last_error.c_str(); }
int fv2d_abi_version(void) { return 1; }

int fv2d_device_count(int *count)
{
  if (!count)
    return arg_fail("null output");
  *count = 0;
  FV2D_CUDA(cudaGetDeviceCount(count));
  return FV2D_OK;
}

int fv2d_params_from_ini(const char *ini_path, const char *overrides, fv2d_device_params *dev, fv2d_run_params *run)
{
  if (!ini_path || !dev || !run)
    return arg_fail("null argument");
  try
  {
    {
      std::ifstream probe(ini_path);
      if (!probe.good())
      {
        set_error(std::string("cannot open ini file ") + ini_path);
        return FV2D_ERR_IO;
      }
    }
    Params prm = readInifile(ini_path, parse_overrides(overrides));
    *dev       = prm.device_params;
    *run       = prm.run_pod();
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_CONFIG;
  }
  return FV2D_OK;
}

int fv2d_params_dump_ini(const char *ini_path, const char *overrides, const char *out_path)
{
  if (!ini_path || !out_path)
    return arg_fail("null argument");
  try
  {
    std::ostringstream sink;
    Params prm = readInifile(ini_path, parse_overrides(overrides), sink);
    std::ofstream out(out_path);
    if (!out.good())
    {
      set_error(std::string("cannot write ") + out_path);
      return FV2D_ERR_IO;
    }
    prm.reader.outputValues(out);
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_CONFIG;
  }
  return FV2D_OK;
}

int fv2d_init_problem(const fv2d_device_params *dev, const fv2d_run_params *run, double *hostQ)
{
  if (!dev || !run || !hostQ)
    return arg_fail("null argument");
  try
  {
    Params prm;
    static_cast<fv2d_device_params &>(prm.device_params) = *dev;
    prm.problem                                           = run->problem;
    prm.seed                                              = run->seed;
    InitFunctor init(prm);
    HostArray Q(dev->Nty, dev->Ntx);
    init.init(Q);
    std::memcpy(hostQ, Q.data.data(), Q.data.size() * sizeof(double));
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_CONFIG;
  }
  return FV2D_OK;
}

int fv2d_init_problem_rows(const fv2d_device_params *dev, const fv2d_run_params *run, int j_first, int nrows,
                           double *hostQ_rows)
{
  if (!dev || !run || !hostQ_rows)
    return arg_fail("null argument");
  if (j_first < 0 || nrows < 1 || j_first + nrows > dev->Nty)
    return arg_fail("row window outside the grid");
  try
  {
    Params prm;
    static_cast<fv2d_device_params &>(prm.device_params) = *dev;
    prm.problem                                           = run->problem;
    prm.seed                                              = run->seed;
    InitFunctor init(prm);
    HostArray Q(nrows, dev->Ntx);
    init.init_rows(Q, j_first);
    std::memcpy(hostQ_rows, Q.data.data(), Q.data.size() * sizeof(double));
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_CONFIG;
  }
  return FV2D_OK;
}

static SnapshotConfig snapshot_config(const fv2d_device_params *dev, const fv2d_run_params *run)
{
  SnapshotConfig c;
  c.device_params    = *dev;
  c.output_path      = run->output_path;
  c.filename_out     = run->filename_out;
  c.restart_file     = run->restart_file;
  c.problem          = run->problem;
  c.multiple_outputs = run->multiple_outputs != 0;
  c.tend             = run->tend;
  return c;
}

int fv2d_io_save_solution(const fv2d_device_params *dev, const fv2d_run_params *run, const double *hostQ,
                          int iteration, double t, int *force_file_truncation)
{
  if (!dev || !run || !hostQ || !force_file_truncation)
    return arg_fail("null argument");
  try
  {
    HostArray Q(dev->Nty, dev->Ntx);
    std::memcpy(Q.data.data(), hostQ, Q.data.size() * sizeof(double));
    bool force = *force_file_truncation != 0;
    saveSolutionHost(snapshot_config(dev, run), Q, iteration, t, force);
    *force_file_truncation = force ? 1 : 0;
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_IO;
  }
  return FV2D_OK;
}

int fv2d_io_load_snapshot(const fv2d_device_params *dev, const fv2d_run_params *run, double *hostQ, double *time,
                          int *iteration, int *force_file_truncation)
{
  if (!dev || !run || !hostQ || !time || !iteration || !force_file_truncation)
    return arg_fail("null argument");
  try
  {
    HostArray Q(dev->Nty, dev->Ntx);
    bool force             = *force_file_truncation != 0;
    const RestartInfo info = loadSnapshotHost(snapshot_config(dev, run), Q, force);
    std::memcpy(hostQ, Q.data.data(), Q.data.size() * sizeof(double));
    *time                  = info.time;
    *iteration             = info.iteration;
    *force_file_truncation = force ? 1 : 0;
  }
  catch (const std::exception &e)
  {
    set_error(e.what());
    return FV2D_ERR_IO;
  }
  return FV2D_OK;
}

int fv2d_ctx_create_slab(const fv2d_device_params *dev, int time_stepping, double eps_reset_negative, int device,
                         int rank, int nranks, fv2d_ctx **out)
{
  if (!dev || !out)
    return arg_fail("null argument");
  *out = nullptr;
  int rc;
  if ((rc = check_supported(*dev)))
    return rc;
  if (nranks < 1 || rank < 0 || rank >= nranks || dev->Ny < nranks)
    return arg_fail("bad slab decomposition: need 0 <= rank < nranks <= Ny");
  if (time_stepping != FV2D_TS_EULER && time_stepping != FV2D_TS_RK2)
    return arg_fail("time_stepping must be FV2D_TS_EULER or FV2D_TS_RK2");
  if (nranks > kMaxRanks)
    return arg_fail("at most 8 y-slabs (one NVSwitch box) are supported");
  if (nranks > 1 && dev->Ny / nranks < 2 * dev->Ng)
    return arg_fail("slab thinner than the ghost layer");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
  {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return FV2D_ERR_CUDA;
  }
  if (device < 0 || device >= ndev)
    return arg_fail("device index out of range");
  FV2D_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FV2D_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
  {
    set_error(std::string("device '") + prop.name + "' is not sm_100: the kernels are built for sm_100a only");
    return FV2D_ERR_CUDA;
  }

  fv2d_ctx *c = new fv2d_ctx();
  std::memset(c, 0, sizeof *c);
  c->glob          = *dev;
  c->time_stepping = time_stepping;
  c->device        = device;
  c->rank          = rank;
  c->nranks        = nranks;

  KParams &kp  = c->kp;
  kp.p         = *dev;
  // rows are dealt out as evenly as they go: the first Ny % nranks slabs get one row more
  const int Nyl = dev->Ny / nranks + (rank < dev->Ny % nranks ? 1 : 0);
  kp.p.Ny       = Nyl;
  kp.p.Nty      = Nyl + 2 * dev->Ng;
  kp.p.jend     = dev->Ng + Nyl;
  kp.Ny_global  = dev->Ny;
  kp.j_global_offset = rank * (dev->Ny / nranks) + std::min(rank, dev->Ny % nranks);
  const bool periodic_y = dev->boundary_y == FV2D_BC_PERIODIC;
  kp.edge_lo = (rank == 0 && !(periodic_y && nranks > 1)) ? EDGE_PHYSICAL : EDGE_NEIGHBOUR;
  kp.edge_hi = (rank == nranks - 1 && !(periodic_y && nranks > 1)) ? EDGE_PHYSICAL : EDGE_NEIGHBOUR;
  kp.eps_reset = eps_reset_negative;
  kp.rank      = rank;
  kp.nranks    = nranks;
  c->num_sms   = prop.multiProcessorCount;

  Layout &L = kp.L;
  L.lead    = (16 - dev->ibeg % 16) % 16;
  L.pitch   = ((L.lead + dev->Ntx + 15) / 16) * 16;
  // room for the TMA box of the last strip to stay inside the row where possible
  L.rows  = kp.p.Nty;
  L.plane = (long long)L.pitch * L.rows;

  auto fail = [&](int code) {
    fv2d_ctx_destroy(c);
    return code;
  };
#define FV2D_TRY(call)                                              \
  do                                                                \
  {                                                                 \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess)                                         \
      return fail(cuda_fail(e__, #call, __FILE__, __LINE__));       \
  } while (0)

  FV2D_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  const size_t bytes = array_bytes(L);
  FV2D_TRY(cudaMalloc(&c->Q[0], bytes));
  FV2D_TRY(cudaMalloc(&c->Q[1], bytes));
  FV2D_TRY(cudaMalloc(&c->U, bytes));
  FV2D_TRY(cudaMemsetAsync(c->Q[0], 0, bytes, c->stream));
  FV2D_TRY(cudaMemsetAsync(c->Q[1], 0, bytes, c->stream));
  FV2D_TRY(cudaMemsetAsync(c->U, 0, bytes, c->stream));
  FV2D_TRY(cudaMalloc(&c->sc, sizeof(DevScalars)));
  FV2D_TRY(cudaMemsetAsync(c->sc, 0, sizeof(DevScalars), c->stream));
  FV2D_TRY(cudaMallocHost(&c->sc_host, sizeof(DevScalars)));
  std::memset(c->sc_host, 0, sizeof(DevScalars));
  kp.sc = c->sc;
  for (int q = 0; q < kMaxRanks; ++q)
    kp.peer_sc[q] = nullptr;
  kp.peer_sc[rank] = c->sc; // a single slab mails itself

  // analytical gravity profile: glibc sin() on the host, narrowed to float (Gravity.h:15-29, Q5)
  if (dev->gravity_mode == FV2D_GRAV_ANALYTICAL)
  {
    std::vector<double> g(kp.p.Nty);
    for (int j = 0; j < kp.p.Nty; ++j)
    {
      const int jg   = j + kp.j_global_offset;
      const double y = dev->ymin + (jg - dev->jbeg + 0.5) * dev->dy;
      g[j]           = (double)(float)(dev->hot_bubble_g0 * std::sin(y * M_PI * 2.0 / dev->ymax));
    }
    FV2D_TRY(cudaMalloc(&c->gtab, g.size() * sizeof(double)));
    FV2D_TRY(cudaMemcpyAsync(c->gtab, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FV2D_TRY(cudaStreamSynchronize(c->stream));
  }
  kp.gtab = c->gtab;

  // inverse-dt accumulators start at the Max identity
  {
    DevScalars init;
    std::memset(&init, 0, sizeof init);
    for (int a = 0; a < 2; ++a)
      for (int k = 0; k < 4; ++k)
        init.inv_acc[a][k] = FV2D_ENC_NEG_MAX;
    FV2D_TRY(cudaMemcpyAsync(c->sc, &init, offsetof(DevScalars, dt_hist), cudaMemcpyHostToDevice, c->stream));
    FV2D_TRY(cudaStreamSynchronize(c->stream));
  }

  FV2D_TRY(sweep_configure());
  c->tmap_ok = (make_tmap(&c->tmapQ[0], c->Q[0], L, sweep_strip_width() + 4) == FV2D_OK) &&
               (make_tmap(&c->tmapQ[1], c->Q[1], L, sweep_strip_width() + 4) == FV2D_OK) &&
               (make_tmap(&c->tmapU, c->U, L, sweep_strip_width()) == FV2D_OK);
  if (!c->tmap_ok)
    return fail(FV2D_ERR_CUDA);
  if ((rc = upload_store_tmap(c, 0, c->U)))
    return fail(rc);
  if ((rc = build_work_items(c)))
    return fail(rc);
  // The sweep writes the ghost cells of its own output when every ghost mirrors a DOMAIN cell
  // (always, except on grids narrower than the ghost layer, where a reflecting ghost mirrors
  // another ghost: those keep the stand-alone ghost-fill launch).
  c->fold_ok = dev->Nx >= dev->Ng && kp.p.Ny >= dev->Ng;
#undef FV2D_TRY
  *out = c;
  return FV2D_OK;
}

int fv2d_ctx_create(const fv2d_device_params *dev, int time_stepping, double eps_reset_negative, int device,
                    fv2d_ctx **out)
{
  return fv2d_ctx_create_slab(dev, time_stepping, eps_reset_negative, device, 0, 1, out);
}

void fv2d_ctx_destroy(fv2d_ctx *c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  for (int k = 0; k < c->n_ipc_opened; ++k)
    cudaIpcCloseMemHandle(c->ipc_opened[k]);
  cudaFree(c->Q[0]);
  cudaFree(c->Q[1]);
  cudaFree(c->U);
  cudaFree(c->Ustar);
  cudaFree(c->slopesX);
  cudaFree(c->slopesY);
  cudaFree(c->gtab);
  cudaFree(c->items_dev);
  cudaFree(c->tmaps_dev);
  cudaFree(c->rowsum);
  cudaFree(c->sitems_dev);
  cudaFree(c->dense_in);
  cudaFree(c->dense_out);
  if (c->s_ev)
  {
    for (int k = 0; k < 2 * c->n_sblocks + 2; ++k)
      cudaEventDestroy(c->s_ev[k]);
    delete[] c->s_ev;
  }
  delete[] c->sblocks;
  if (c->s_up)
    cudaStreamDestroy(c->s_up);
  if (c->s_dn)
    cudaStreamDestroy(c->s_dn);
  cudaFree(c->sc);
  if (c->prof_ev)
  {
    for (int k = 0; k < 2 * kProfMax; ++k)
      cudaEventDestroy(c->prof_ev[k]);
    delete[] c->prof_ev;
  }
  if (c->sc_host)
    cudaFreeHost(c->sc_host);
  if (c->stage_host)
    cudaFreeHost(c->stage_host);
  if (c->own_stream && c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
}

int fv2d_ctx_set_stream(fv2d_ctx *c, void *cuda_stream)
{
  if (!c)
    return arg_fail("null context");
  FV2D_CUDA(cudaSetDevice(c->device));
  FV2D_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream)
    cudaStreamDestroy(c->stream);
  c->stream     = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return FV2D_OK;
}

int fv2d_sync(fv2d_ctx *c)
{
  if (!c)
    return arg_fail("null context");
  FV2D_CUDA(cudaSetDevice(c->device));
  return sync_ctx(c);
}

int fv2d_ctx_geometry(const fv2d_ctx *c, int64_t out[6])
{
  if (!c || !out)
    return arg_fail("null argument");
  out[0] = c->kp.p.Ntx;
  out[1] = c->kp.p.Nty;
  out[2] = c->kp.p.Ny;
  out[3] = c->kp.j_global_offset;
  out[4] = c->kp.L.pitch;
  out[5] = c->kp.L.lead;
  return FV2D_OK;
}

#define FV2D_ENTER(c)                       \
  if (!(c))                                 \
    return arg_fail("null context");        \
  FV2D_CUDA(cudaSetDevice((c)->device));

int fv2d_upload_Q(fv2d_ctx *c, const double *hostQ)
{
  FV2D_ENTER(c);
  if (!hostQ)
    return arg_fail("null host array");
  c->ghosts_valid = c->dt_valid = false;
  int rc = copy_h2d(c, c->Q[c->cur], hostQ);
  return rc ? rc : sync_ctx(c);
}
int fv2d_upload_U(fv2d_ctx *c, const double *hostU)
{
  FV2D_ENTER(c);
  if (!hostU)
    return arg_fail("null host array");
  int rc = copy_h2d(c, c->U, hostU);
  return rc ? rc : sync_ctx(c);
}
int fv2d_download_Q(fv2d_ctx *c, double *hostQ)
{
  FV2D_ENTER(c);
  if (!hostQ)
    return arg_fail("null host array");
  int rc = copy_d2h(c, hostQ, c->Q[c->cur]);
  return rc ? rc : sync_ctx(c);
}
int fv2d_download_U(fv2d_ctx *c, double *hostU)
{
  FV2D_ENTER(c);
  if (!hostU)
    return arg_fail("null host array");
  int rc = copy_d2h(c, hostU, c->U);
  return rc ? rc : sync_ctx(c);
}

int fv2d_prim_to_cons(fv2d_ctx *c)
{
  FV2D_ENTER(c);
  NvtxRange r("Primitive to Conservative");
  launch_prim_to_cons(c->kp, c->Q[c->cur], c->U, c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_cons_to_prim(fv2d_ctx *c)
{
  FV2D_ENTER(c);
  c->ghosts_valid = c->dt_valid = false; // Q is rewritten over range_tot
  NvtxRange r("Conservative to Primitive");
  launch_cons_to_prim(c->kp, c->U, c->Q[c->cur], c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_check_negatives(fv2d_ctx *c, uint64_t counts[3])
{
  FV2D_ENTER(c);
  c->ghosts_valid = c->dt_valid = false; // negative cells are reset in place
  int rc;
  if ((rc = read_scalars(c)))
    return rc;
  unsigned long long before[3] = {c->sc_host->neg[0], c->sc_host->neg[1], c->sc_host->neg[2]};
  {
    NvtxRange r("Check negative density/pressure");
    launch_check_negatives(c->kp, c->Q[c->cur], c->sc->neg, c->stream);
  }
  FV2D_CUDA(cudaGetLastError());
  if ((rc = read_scalars(c)))
    return rc;
  if (counts)
    for (int k = 0; k < 3; ++k)
      counts[k] = c->sc_host->neg[k] - before[k];
  return FV2D_OK;
}
int fv2d_fill_boundaries(fv2d_ctx *c)
{
  FV2D_ENTER(c);
  c->ghosts_valid = (c->nranks == 1); // (ghost rows on a neighbour-slab side belong to the halo exchange)
  launch_fill_boundaries(c->kp, c->Q[c->cur], c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_compute_dt(fv2d_ctx *c, double *dt, double inv_dt[3])
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = compute_dt_now(c)))
    return rc;
  if ((rc = read_scalars(c)))
    return rc;
  if (dt)
    *dt = c->sc_host->dt;
  if (inv_dt)
    for (int k = 0; k < 3; ++k)
      inv_dt[k] = c->sc_host->inv_dt_last[k];
  return FV2D_OK;
}
int fv2d_compute_slopes(fv2d_ctx *c)
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = ensure_slopes(c)))
    return rc;
  launch_compute_slopes(c->kp, c->Q[c->cur], c->slopesX, c->slopesY, c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_compute_fluxes_and_update(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = ensure_slopes(c)))
    return rc;
  launch_fluxes_and_update(c->kp, c->Q[c->cur], c->slopesX, c->slopesY, c->U, dt, c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_apply_thermal_conduction(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  launch_thermal_conduction(c->kp, c->Q[c->cur], c->U, dt, c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_apply_viscosity(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  launch_viscosity(c->kp, c->Q[c->cur], c->U, dt, c->stream);
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}
int fv2d_euler_step(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  return euler_step_ops(c, c->Q[c->cur], c->U, dt);
}
int fv2d_update(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  c->ghosts_valid = c->dt_valid = false; // (RK2 overwrites Q with consToPrim(U*): Update.h:210)
  if (c->time_stepping == FV2D_TS_EULER)
    return euler_step_ops(c, c->Q[c->cur], c->U, dt);
  // SSP-RK2, Update.h:197-221
  int rc;
  if ((rc = ensure_ustar(c)))
    return rc;
  const size_t bytes = array_bytes(c->kp.L);
  double *Q = c->Q[c->cur], *U0 = c->Q[c->cur ^ 1]; // the spare Q buffer doubles as U0
  FV2D_CUDA(cudaMemcpyAsync(U0, c->U, bytes, cudaMemcpyDeviceToDevice, c->stream));
  FV2D_CUDA(cudaMemcpyAsync(c->Ustar, c->U, bytes, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = euler_step_ops(c, Q, c->Ustar, dt)))
    return rc;
  FV2D_CUDA(cudaMemcpyAsync(c->U, c->Ustar, bytes, cudaMemcpyDeviceToDevice, c->stream));
  launch_cons_to_prim(c->kp, c->Ustar, Q, c->stream);
  if ((rc = euler_step_ops(c, Q, c->U, dt)))
    return rc;
  {
    NvtxRange r("RK2 Correct");
    launch_rk2_correct(c->kp, U0, c->U, c->stream);
  }
  FV2D_CUDA(cudaGetLastError());
  return FV2D_OK;
}

int fv2d_step(fv2d_ctx *c, double dt)
{
  FV2D_ENTER(c);
  return fused_step(c, false, dt);
}
int fv2d_step_device_dt(fv2d_ctx *c)
{
  FV2D_ENTER(c);
  return fused_step(c, true, 0.0);
}
int fv2d_run_steps(fv2d_ctx *c, int64_t nsteps)
{
  FV2D_ENTER(c);
  for (int64_t k = 0; k < nsteps; ++k)
  {
    int rc = fused_step(c, true, 0.0);
    if (rc)
      return rc;
  }
  return FV2D_OK;
}
int fv2d_run_until(fv2d_ctx *c, double tend, int64_t max_steps, int64_t *steps_done)
{
  FV2D_ENTER(c);
  int rc;
  int64_t n = 0;
  if ((rc = read_scalars(c)))
    return rc;
  double t = c->sc_host->t;
  // main.cpp:62: while (t + epsilon < tend)
  while (t + c->glob.epsilon < tend && n < max_steps)
  {
    if ((rc = fused_step(c, true, 0.0)))
      return rc;
    if ((rc = read_scalars(c)))
      return rc;
    t = c->sc_host->t;
    ++n;
  }
  if (steps_done)
    *steps_done = n;
  return FV2D_OK;
}

int fv2d_get_time(fv2d_ctx *c, double *t, double *next_dt, int64_t *steps)
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = read_scalars(c)))
    return rc;
  if (t)
    *t = c->sc_host->t;
  if (steps)
    *steps = c->sc_host->step;
  if (next_dt)
  {
    // what step_begin would compute from the current accumulator
    const fv2d_device_params &p = c->kp.p;
    double hyp = 0.0;
    if ((rc = current_hyp(c, &hyp)))
      return rc;
    double tc = p.epsilon, visc = p.epsilon;
    if (p.thermal_conductivity_active)
      tc = std::fmax(2.0 * p.kappa / (p.dx * p.dx), 2.0 * p.kappa / (p.dy * p.dy));
    if (p.viscosity_active)
      visc = std::fmax(2.0 * p.mu / (p.dx * p.dx), 2.0 * p.mu / (p.dy * p.dy));
    double m = hyp;
    if (m < tc)
      m = tc;
    if (m < visc)
      m = visc;
    *next_dt = p.CFL / m;
  }
  return FV2D_OK;
}
int fv2d_set_time(fv2d_ctx *c, double t)
{
  FV2D_ENTER(c);
  FV2D_CUDA(cudaMemcpyAsync(&c->sc->t, &t, sizeof t, cudaMemcpyHostToDevice, c->stream));
  return sync_ctx(c);
}
int fv2d_get_dt_history(fv2d_ctx *c, double *dts, int64_t n, int64_t *n_out)
{
  FV2D_ENTER(c);
  if (!dts)
    return arg_fail("null output");
  FV2D_CUDA(cudaMemcpyAsync(c->sc_host, c->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
  int rc;
  if ((rc = sync_ctx(c)))
    return rc;
  const int64_t steps = c->sc_host->step;
  int64_t m           = n;
  if (m > steps)
    m = steps;
  if (m > FV2D_DT_HISTORY)
    m = FV2D_DT_HISTORY;
  for (int64_t k = 0; k < m; ++k)
    dts[k] = c->sc_host->dt_hist[(steps - m + k) % FV2D_DT_HISTORY];
  if (n_out)
    *n_out = m;
  return FV2D_OK;
}
int fv2d_get_negative_counts(fv2d_ctx *c, uint64_t counts[3], int reset)
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = read_scalars(c)))
    return rc;
  if (counts)
    for (int k = 0; k < 3; ++k)
      counts[k] = c->sc_host->neg[k];
  if (reset)
  {
    FV2D_CUDA(cudaMemsetAsync(c->sc->neg, 0, sizeof(c->sc->neg), c->stream));
    return sync_ctx(c);
  }
  return FV2D_OK;
}
int fv2d_get_inv_dt(fv2d_ctx *c, double inv_dt[3])
{
  FV2D_ENTER(c);
  if (!inv_dt)
    return arg_fail("null output");
  int rc;
  if ((rc = read_scalars(c)))
    return rc;
  // maxima of the CURRENT state (what the next step's dt is made of), ComputeDt.h:30-52
  const fv2d_device_params &p = c->kp.p;
  if ((rc = current_hyp(c, &inv_dt[0])))
    return rc;
  inv_dt[1] = p.epsilon;
  inv_dt[2] = p.epsilon;
  if (p.thermal_conductivity_active)
    inv_dt[1] = std::fmax(2.0 * p.kappa / (p.dx * p.dx), 2.0 * p.kappa / (p.dy * p.dy));
  if (p.viscosity_active)
    inv_dt[2] = std::fmax(2.0 * p.mu / (p.dx * p.dx), 2.0 * p.mu / (p.dy * p.dy));
  return FV2D_OK;
}
int fv2d_integrate_mass_energy(fv2d_ctx *c, double *mass, double *energy)
{
  FV2D_ENTER(c);
  if (!c->rowsum)
    FV2D_CUDA(cudaMalloc(&c->rowsum, (size_t)2 * c->kp.p.Ny * sizeof(double)));
  launch_mass_energy(c->kp, c->U, c->rowsum, c->stream);
  int rc = read_scalars(c);
  if (rc)
    return rc;
  if (mass)
    *mass = c->sc_host->sums[0];
  if (energy)
    *energy = c->sc_host->sums[1];
  return FV2D_OK;
}

int fv2d_state_hash(fv2d_ctx *c, uint64_t *hash)
{
  FV2D_ENTER(c);
  if (!hash)
    return arg_fail("null output");
  FV2D_CUDA(cudaMemsetAsync(&c->sc->hash, 0, sizeof(unsigned long long), c->stream));
  launch_state_hash(c->kp, c->U, &c->sc->hash, c->stream);
  int rc = read_scalars(c);
  if (rc)
    return rc;
  *hash = c->sc_host->hash;
  return FV2D_OK;
}

int fv2d_debug_sync_wait(fv2d_ctx *c, double *last_wait_us, double *total_wait_us, double *last_busy_us, int reset)
{
  FV2D_ENTER(c);
  int rc = read_scalars(c);
  if (rc)
    return rc;
  const unsigned long long *ts = c->sc_host->tstamp;
  if (last_wait_us)
    *last_wait_us = (double)(ts[1] - ts[0]) * 1e-3;
  if (total_wait_us)
    *total_wait_us = (double)ts[3] * 1e-3;
  if (last_busy_us)
    *last_busy_us = (double)(ts[2] - ts[1]) * 1e-3;
  if (reset)
  {
    FV2D_CUDA(cudaMemsetAsync(&c->sc->tstamp[3], 0, sizeof(unsigned long long), c->stream));
    return sync_ctx(c);
  }
  return FV2D_OK;
}

int fv2d_debug_schedule(int Nx, int Ny_local, int num_sms, int neighbour_lo, int neighbour_hi, int32_t *runs, int max_runs,
                        int *n_runs)
{
  if (Nx < 1 || Ny_local < 1 || num_sms < 1 || !runs || !n_runs)
    return arg_fail("fv2d_debug_schedule: bad arguments");
  const int W = sweep_strip_width();
  const auto r = schedule_runs(Ny_local, (Nx + W - 1) / W, 2 * num_sms, neighbour_lo != 0, neighbour_hi != 0);
  *n_runs      = (int)r.size();
  for (int k = 0; k < (int)r.size() && k < max_runs; ++k)
    runs[2 * k] = r[k].first, runs[2 * k + 1] = r[k].second;
  return FV2D_OK;
}

int fv2d_debug_sweep_timing(fv2d_ctx *c, int64_t *out, int n)
{
  FV2D_ENTER(c);
  int rc = sync_ctx(c);
  if (rc)
    return rc;
  rc = read_sweep_timing(c->kp.p.riemann_solver, (long long *)out, n);
  if (rc == 1)
    return arg_fail("this build of the library has no sweep timing (rebuild with -DFV2D_TIMING)");
  if (rc)
    return cuda_fail(cudaGetLastError(), "read_sweep_timing", __FILE__, __LINE__);
  return FV2D_OK;
}

int fv2d_debug_fp64_peak(int device, double *dfma_per_second)
{
  if (!dfma_per_second)
    return arg_fail("null output");
  FV2D_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FV2D_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d = nullptr;
  FV2D_CUDA(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  FV2D_CUDA(cudaEventCreate(&e0));
  FV2D_CUDA(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) // first repetition warms up
  {
    cudaEventRecord(e0, nullptr);
    launch_fp64_peak(blocks, iters, d, nullptr);
    cudaEventRecord(e1, nullptr);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms      = 0.f;
    if (e == cudaSuccess)
      e = cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess)
    {
      cudaFree(d);
      return cuda_fail(e, "fp64 peak probe", __FILE__, __LINE__);
    }
    const double rate = 8.0 * iters * 256.0 * blocks / (ms * 1e-3); // thread-level DFMAs per second
    if (rep > 0 && rate > best)
      best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *dfma_per_second = best;
  return FV2D_OK;
}

int fv2d_debug_math_probe(int device, int64_t n, const double *a, const double *b, double *out_rcp, double *out_cs)
{
  if (n <= 0 || !a || !b || !out_rcp || !out_cs)
    return arg_fail("fv2d_debug_math_probe: bad arguments");
  FV2D_CUDA(cudaSetDevice(device));
  double *d           = nullptr;
  const size_t nbytes = (size_t)n * sizeof(double);
  FV2D_CUDA(cudaMalloc(&d, 4 * nbytes));
  cudaError_t e = cudaMemcpy(d, a, nbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d + n, b, nbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
  {
    launch_math_probe(n, d, d + n, d + 2 * n, d + 3 * n, nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpy(out_rcp, d + 2 * n, nbytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess)
    e = cudaMemcpy(out_cs, d + 3 * n, nbytes, cudaMemcpyDeviceToHost);
  cudaFree(d);
  FV2D_CUDA(e);
  return FV2D_OK;
}

int fv2d_profile_enable(fv2d_ctx *c, int on)
{
  FV2D_ENTER(c);
  if (on && !c->prof_ev)
  {
    c->prof_ev = new cudaEvent_t[2 * kProfMax];
    for (int k = 0; k < 2 * kProfMax; ++k)
      FV2D_CUDA(cudaEventCreate(&c->prof_ev[k]));
  }
  c->profile        = on == 1; // on == 2: count the launches, record no events
  c->prof_n         = 0;
  c->n_launch_sweep = 0;
  c->n_launch_total = 0;
  return FV2D_OK;
}
int fv2d_profile_read(fv2d_ctx *c, double *sweep_ms, int64_t *sweep_launches, int64_t *total_launches)
{
  FV2D_ENTER(c);
  int rc;
  if ((rc = sync_ctx(c)))
    return rc;
  double ms = 0.0;
  for (int k = 0; k < c->prof_n; ++k)
  {
    float x = 0.f;
    FV2D_CUDA(cudaEventElapsedTime(&x, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]));
    ms += x;
  }
  if (sweep_ms)
    *sweep_ms = ms;
  if (sweep_launches)
    *sweep_launches = c->profile ? c->prof_n : c->n_launch_sweep;
  if (total_launches)
    *total_launches = c->n_launch_total;
  return FV2D_OK;
}

int fv2d_advance_host(fv2d_ctx *c, const double *hostQ_in, double *hostQ_out, int64_t nsteps, double *dts)
{
  FV2D_ENTER(c);
  if (!hostQ_in || !hostQ_out)
    return arg_fail("null host array");
  if (dts && nsteps > FV2D_DT_HISTORY)
    return arg_fail("fv2d_advance_host: the dt sequence of more than FV2D_DT_HISTORY steps is not kept; pass dts = NULL "
                    "or advance in shorter calls");
  int rc;
  c->ghosts_valid = c->dt_valid = false;
  if ((rc = copy_h2d(c, c->Q[c->cur], hostQ_in)))
    return rc;
  launch_prim_to_cons(c->kp, c->Q[c->cur], c->U, c->stream);
  c->n_launch_total++;
  if ((rc = compute_dt_now(c)))
    return rc;
  for (int64_t k = 0; k < nsteps; ++k)
    if ((rc = fused_step(c, true, 0.0)))
      return rc;
  if (c->nranks > 1 && nsteps > 0)
  {
    // y-slab: the ghost rows of the new state are pushed by the neighbours' sweeps; wait for
    // them (and fill the x ghosts) so the array handed back is complete and can be fed to the
    // next call as it is
    launch_fill_ghosts(c->kp, c->Q[c->cur], c->halo_gen * pushes_per_sweep(c), c->stream);
    c->n_launch_total++;
  }
  if ((rc = copy_d2h(c, hostQ_out, c->Q[c->cur])))
    return rc;
  if (dts)
  {
    int64_t got = 0;
    return fv2d_get_dt_history(c, dts, nsteps, &got);
  }
  return sync_ctx(c);
}

// ------------------------------------------------------------------ streamed host path

// array rows [r0, r1) of the four field planes between the host array [f][Nty][Ntx] and its device
// copy in the same layout: one flat copy per plane
static int copy_rows(fv2d_ctx *c, double *dense, double *host, int r0, int r1, bool to_device, cudaStream_t s)
{
  if (r1 <= r0)
    return FV2D_OK;
  const size_t Ntx = (size_t)c->kp.p.Ntx, rows = (size_t)c->kp.L.rows;
  const size_t bytes = (size_t)(r1 - r0) * Ntx * sizeof(double);
  for (int f = 0; f < 4; ++f)
  {
    const size_t o = ((size_t)f * rows + (size_t)r0) * Ntx;
    if (to_device)
      FV2D_CUDA(cudaMemcpyAsync(dense + o, host + o, bytes, cudaMemcpyHostToDevice, s));
    else
      FV2D_CUDA(cudaMemcpyAsync(host + o, dense + o, bytes, cudaMemcpyDeviceToHost, s));
  }
  return FV2D_OK;
}

// Row blocks of the streamed path (pure host logic, also behind fv2d_debug_stream_blocks).  Block b
// brings up domain rows [up0, up1) (array row indices); the rows whose 2-row stencil is complete by
// then, [sw0, sw1), are swept: everything up to 2 rows below the block's upper edge, the last block
// up to the top of the domain (its upper neighbours are ghost rows, filled on the device).
static std::vector<fv2d_ctx::StreamBlock> stream_blocks(int Ny, int jbeg, int block_rows)
{
  const int B = std::max(16, block_rows), jend = jbeg + Ny;
  const int nb = std::max(1, Ny / B);
  std::vector<fv2d_ctx::StreamBlock> blk((size_t)nb);
  for (int b = 0; b < nb; ++b)
  {
    fv2d_ctx::StreamBlock &k = blk[b];
    k.up0 = jbeg + b * B;
    k.up1 = (b == nb - 1) ? jend : jbeg + (b + 1) * B;
    k.sw0 = (b == 0) ? jbeg : blk[b - 1].sw1;
    k.sw1 = (b == nb - 1) ? jend : k.up1 - 2;
    k.item_off = k.n_items = k.n_ctas = k.persistent = 0;
  }
  return blk;
}
static int stream_block_rows()
{
  if (const char *e = std::getenv("FV2D_STREAM_ROWS"))
    if (std::atoi(e) >= 16)
      return std::atoi(e);
  return 256;
}

// ... and the work table of each block's partial sweep
static int ensure_stream_blocks(fv2d_ctx *c)
{
  if (c->sblocks)
    return FV2D_OK;
  const fv2d_device_params &p = c->kp.p;
  const std::vector<fv2d_ctx::StreamBlock> plan = stream_blocks(p.Ny, p.jbeg, stream_block_rows());
  const int nb = (int)plan.size();
  const int W = sweep_strip_width(), nstrips = (p.Nx + W - 1) / W, slots = 2 * c->num_sms;
  std::vector<WorkItem> items;
  fv2d_ctx::StreamBlock *blk = new fv2d_ctx::StreamBlock[nb];
  for (int b = 0; b < nb; ++b)
  {
    fv2d_ctx::StreamBlock &k = blk[b];
    k = plan[b];
    const std::vector<std::pair<int, int>> runs = schedule_runs(k.sw1 - k.sw0, nstrips, slots, false, false);
    k.item_off = (int)items.size();
    int min_rows = k.sw1 - k.sw0;
    for (const auto &r : runs)
    {
      min_rows = std::min(min_rows, r.second - r.first);
      for (int s = 0; s < nstrips; ++s)
        items.push_back(WorkItem{s, k.sw0 + r.first, k.sw0 + r.second, 0});
    }
    k.n_items    = (int)items.size() - k.item_off;
    k.persistent = min_rows >= 8;
    k.n_ctas     = k.persistent ? std::min(k.n_items, slots) : k.n_items;
    items.push_back(WorkItem{0, -1, -1, 0}); // end marker of this block's table
  }
  // all or nothing: a failed allocation (the two dense copies are as large as Q) leaves the context as
  // it was, and the caller falls back to the serial route
  const size_t dense_bytes = (size_t)4 * p.Nty * p.Ntx * sizeof(double);
  WorkItem *items_dev = nullptr;
  double *din = nullptr, *dout = nullptr;
  cudaError_t e = cudaMalloc(&items_dev, items.size() * sizeof(WorkItem));
  if (e == cudaSuccess)
    e = cudaMalloc(&din, dense_bytes);
  if (e == cudaSuccess)
    e = cudaMalloc(&dout, dense_bytes);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(items_dev, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess)
  {
    cudaFree(items_dev), cudaFree(din), cudaFree(dout);
    delete[] blk;
    cudaGetLastError(); // the error is reported, not sticky
    return cuda_fail(e, "streamed path: staging buffers", __FILE__, __LINE__);
  }
  if (!c->s_up)
    FV2D_CUDA(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
  if (!c->s_dn)
    FV2D_CUDA(cudaStreamCreateWithFlags(&c->s_dn, cudaStreamNonBlocking));
  cudaEvent_t *ev = new cudaEvent_t[2 * nb + 2];
  for (int k = 0; k < 2 * nb + 2; ++k)
    cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
  c->sitems_dev = items_dev, c->dense_in = din, c->dense_out = dout;
  c->s_ev      = ev;
  c->sblocks   = blk;
  c->n_sblocks = nb;
  return FV2D_OK;
}

int fv2d_debug_stream_blocks(int Ny, int Ng, int block_rows, int32_t *blocks, int max_blocks, int *n_blocks)
{
  if (Ny < 1 || Ng < 0 || !blocks || !n_blocks)
    return arg_fail("fv2d_debug_stream_blocks: bad arguments");
  const auto b = stream_blocks(Ny, Ng, block_rows > 0 ? block_rows : stream_block_rows());
  *n_blocks    = (int)b.size();
  for (int k = 0; k < (int)b.size() && k < max_blocks; ++k)
    blocks[4 * k] = b[k].up0, blocks[4 * k + 1] = b[k].up1, blocks[4 * k + 2] = b[k].sw0, blocks[4 * k + 3] = b[k].sw1;
  return FV2D_OK;
}

static double next_dt_of(const fv2d_device_params &p, double hyp)
{
  double tc = p.epsilon, visc = p.epsilon;
  if (p.thermal_conductivity_active)
    tc = std::fmax(2.0 * p.kappa / (p.dx * p.dx), 2.0 * p.kappa / (p.dy * p.dy));
  if (p.viscosity_active)
    visc = std::fmax(2.0 * p.mu / (p.dx * p.dx), 2.0 * p.mu / (p.dy * p.dy));
  double m = hyp;
  if (m < tc)
    m = tc;
  if (m < visc)
    m = visc;
  return p.CFL / m;
}

int fv2d_advance_host_stream(fv2d_ctx *c, const double *hostQ_in, double *hostQ_out, double dt_hint, double *dt_used,
                             double *dt_next, int *streamed)
{
  FV2D_ENTER(c);
  if (!hostQ_in || !hostQ_out)
    return arg_fail("null host array");
  if (!c->tmap_ok)
    return arg_fail("TMA descriptors unavailable");
  if (c->nranks > 1 && !c->connected)
    return arg_fail("multi-GPU context: call fv2d_halo_connect before stepping");
  int rc;
  const fv2d_device_params &p = c->kp.p;
  static const unsigned long long neg_max = FV2D_ENC_NEG_MAX;
  double *hin = const_cast<double *>(hostQ_in);
  if (streamed)
    *streamed = 0;
  c->ghosts_valid = c->dt_valid = false;
  FV2D_CUDA(cudaMemcpyAsync(&c->sc->inv_acc[0][0], &neg_max, sizeof neg_max, cudaMemcpyHostToDevice, c->stream));

  // Speculation needs: a dt to speculate with, a single slab (the check of the hint is a local
  // decision), one sweep per step, and ghost cells the sweep can write itself.
  // (in place - hostQ_out == hostQ_in - the redo after a rejected hint would find its input overwritten)
  bool speculate = dt_hint > 0.0 && c->nranks == 1 && c->time_stepping != FV2D_TS_RK2 && c->fold_ok &&
                   hostQ_in != hostQ_out && !std::getenv("FV2D_STREAM_OFF");
  if (speculate && ensure_stream_blocks(c) != FV2D_OK)
    speculate = false; // no room for the staging copies: the serial route needs none
  bool need_plain_step = true;
  if (speculate)
  {
    const int cur = c->cur, nxt = cur ^ 1, nb = c->n_sblocks;
    cudaEvent_t *ev = c->s_ev;
    // the copy streams start after whatever the context's stream still has in flight
    FV2D_CUDA(cudaMemcpyAsync(c->sc->neg_save, c->sc->neg, sizeof(c->sc->neg), cudaMemcpyDeviceToDevice, c->stream));
    FV2D_CUDA(cudaEventRecord(ev[2 * nb], c->stream));
    FV2D_CUDA(cudaStreamWaitEvent(c->s_up, ev[2 * nb], 0));
    FV2D_CUDA(cudaStreamWaitEvent(c->s_dn, ev[2 * nb], 0));
    // development: where the call's time goes (FV2D_STREAM_TRACE=1 prints a line per call)
    static const bool trace = std::getenv("FV2D_STREAM_TRACE") != nullptr;
    static cudaEvent_t tev[4];
    static bool tev_ok = false;
    const auto host_t0 = std::chrono::steady_clock::now();
    if (trace)
    {
      if (!tev_ok)
        for (auto &e : tev)
          cudaEventCreate(&e);
      tev_ok = true;
      cudaEventRecord(tev[0], c->s_up);
    }
    // a periodic y boundary mirrors the LAST domain rows into the low ghost rows: bring them up first
    if (p.boundary_y == FV2D_BC_PERIODIC)
    {
      if ((rc = copy_rows(c, c->dense_in, hin, p.jend - p.Ng, p.jend, true, c->s_up)))
        return rc;
      FV2D_CUDA(cudaEventRecord(ev[2 * nb + 1], c->s_up));
      FV2D_CUDA(cudaStreamWaitEvent(c->stream, ev[2 * nb + 1], 0));
      launch_prep_rows(c->kp, c->dense_in, c->Q[cur], c->U, p.jend - p.Ng, p.jend, &c->sc->inv_acc[0][0], c->stream);
      c->n_launch_total++;
    }

    SweepArgs a;
    std::memset(&a, 0, sizeof a);
    a.kp            = c->kp;
    a.use_device_dt = 0;
    a.dt_host       = dt_hint;
    a.fold_ghosts   = 1;
    a.lo_rank = a.hi_rank = 0;
    a.mail_gen      = c->mail_gen;
    a.Uin = c->U, a.Uout = c->U, a.U0 = nullptr, a.Qout = c->Q[nxt], a.final_stage = 1, a.partial = 1;
    a.tm_store_u = c->tmaps_dev;
    NvtxRange step_range("streamed step: upload | ghost fill + primToCons + fused sweep | download, by row blocks");
    for (int b = 0; b < nb; ++b)
    {
      const fv2d_ctx::StreamBlock &k = c->sblocks[b];
      if ((rc = copy_rows(c, c->dense_in, hin, k.up0, k.up1, true, c->s_up)))
        return rc;
      FV2D_CUDA(cudaEventRecord(ev[2 * b], c->s_up));
      FV2D_CUDA(cudaStreamWaitEvent(c->stream, ev[2 * b], 0));
      // the block's rows into the padded planes, U and the CFL maximum on the way; then the ghost cells:
      // the x-ghost columns of the block's rows, the low y-ghost rows with the first block, the high
      // ones with the last (their source rows are resident by then)
      launch_prep_rows(c->kp, c->dense_in, c->Q[cur], c->U, k.up0, k.up1, &c->sc->inv_acc[0][0], c->stream);
      const int g0 = (b == 0) ? 0 : k.up0, g1 = (b == nb - 1) ? p.Nty : k.up1;
      launch_fill_ghosts_rows(c->kp, c->Q[cur], c->U, g0, g1, c->stream);
      a.items      = c->sitems_dev + k.item_off;
      a.n_items    = k.n_items;
      a.n_ctas     = k.n_ctas;
      a.persistent = k.persistent;
      cudaError_t e = launch_sweep(c->tmapQ[cur], c->tmapU, a, c->stream);
      if (e != cudaSuccess)
        return cuda_fail(e, "partial sweep", __FILE__, __LINE__);
      // the swept rows (and, with the last block, the y-ghost rows the first and last sweeps wrote) go
      // back in the host's layout
      const int d0 = k.sw0, d1 = (b == nb - 1) ? p.Nty : k.sw1;
      launch_pack_rows(c->kp, c->Q[nxt], c->dense_out, d0, d1, c->stream);
      if (b == nb - 1)
        launch_pack_rows(c->kp, c->Q[nxt], c->dense_out, 0, p.jbeg, c->stream);
      c->n_launch_sweep++;
      c->n_launch_total += 4 + (b == nb - 1);
      FV2D_CUDA(cudaEventRecord(ev[2 * b + 1], c->stream));
      FV2D_CUDA(cudaStreamWaitEvent(c->s_dn, ev[2 * b + 1], 0));
      if ((rc = copy_rows(c, c->dense_out, hostQ_out, d0, d1, false, c->s_dn)))
        return rc;
    }
    if ((rc = copy_rows(c, c->dense_out, hostQ_out, 0, p.jbeg, false, c->s_dn)))
      return rc;
    launch_stream_commit(c->kp, dt_hint, c->mail_gen + 1, c->stream);
    c->n_launch_total++;
    FV2D_CUDA(cudaGetLastError());
    const auto host_t1 = std::chrono::steady_clock::now();
    if (trace)
      cudaEventRecord(tev[1], c->s_up), cudaEventRecord(tev[2], c->stream), cudaEventRecord(tev[3], c->s_dn);
    FV2D_CUDA(cudaStreamSynchronize(c->s_dn));
    if ((rc = read_scalars(c)))
      return rc;
    if (trace)
    {
      float up = 0, comp = 0, dn = 0;
      cudaEventElapsedTime(&up, tev[0], tev[1]), cudaEventElapsedTime(&comp, tev[0], tev[2]), cudaEventElapsedTime(&dn, tev[0], tev[3]);
      fprintf(stderr, "fv2d stream trace: enqueue %.2f ms (host), uploads done at %.2f ms, kernels at %.2f, downloads at %.2f, call %.2f ms (host)\n",
              std::chrono::duration<double, std::milli>(host_t1 - host_t0).count(), up, comp, dn,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    }
    if (c->sc_host->stream_ok)
    {
      c->cur = nxt;
      c->halo_gen++;
      c->mail_gen++;
      c->ghosts_valid = true;
      c->dt_valid     = true;
      need_plain_step = false;
      if (streamed)
        *streamed = 1;
      if (dt_used)
        *dt_used = c->sc_host->dt;
      if (dt_next)
        *dt_next = c->sc_host->dt_next;
    }
    // else: dt_hint was not the state's own time step.  Q[cur] still holds the uploaded state (with
    // its ghosts), inv_acc[0] its CFL maximum; U was overwritten by the speculative sweeps.
  }
  else
  {
    if ((rc = copy_h2d(c, c->Q[c->cur], hostQ_in)))
      return rc;
    // (a y-slab's ghost rows on a neighbour side are taken from the host array, like fv2d_upload_Q)
  }
  if (need_plain_step)
  {
    // dt from the state's CFL maximum in the sweep's arithmetic (so that a call with and a call
    // without a hint give the same bits), then one ordinary fused step and the whole state back
    launch_prep_rows(c->kp, nullptr, c->Q[c->cur], c->U, 0, p.Nty, &c->sc->inv_acc[0][0], c->stream);
    c->mail_gen++;
    launch_finalize_dt(c->kp, &c->sc->inv_acc[0][0], c->mail_gen, c->stream);
    c->n_launch_total += 2;
    c->dt_valid = true;
    if ((rc = fused_step(c, true, 0.0)))
      return rc;
    if (c->nranks > 1)
    {
      launch_fill_ghosts(c->kp, c->Q[c->cur], c->halo_gen * pushes_per_sweep(c), c->stream);
      c->n_launch_total++;
    }
    if ((rc = copy_d2h(c, hostQ_out, c->Q[c->cur])))
      return rc;
    if ((rc = read_scalars(c)))
      return rc;
    if (dt_used)
      *dt_used = c->sc_host->dt;
    if (dt_next)
    {
      double hyp = 0.0;
      if ((rc = current_hyp(c, &hyp)))
        return rc;
      *dt_next = next_dt_of(p, hyp);
    }
  }
  return FV2D_OK;
}

// ------------------------------------------------------------------ multi-GPU halo exchange

// What one rank publishes about its exchange buffers (fits FV2D_IPC_HANDLE_BYTES).
struct HaloHandle
{
  uint32_t magic;
  int32_t rank, nranks, device;
  int64_t pid;
  cudaIpcMemHandle_t mem[3]; // Q[0], Q[1], scalars
  uint64_t offset[3];        // of the pointer inside its cudaMalloc allocation
  uint64_t raw[3];           // same-process shortcut: the device pointers themselves
  int64_t plane;             // doubles per field plane of this rank's arrays (slabs may differ by a row)
  int32_t ny_local;          // rows this rank owns
};
static_assert(sizeof(HaloHandle) <= FV2D_IPC_HANDLE_BYTES, "handle too large");

typedef CUresult (*PFN_getAddressRange)(CUdeviceptr *, size_t *, CUdeviceptr);

int fv2d_halo_export(fv2d_ctx *c, void *handle)
{
  FV2D_ENTER(c);
  if (!handle)
    return arg_fail("null handle");
  static PFN_getAddressRange getRange = nullptr;
  if (!getRange)
  {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FV2D_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qres));
    getRange = (PFN_getAddressRange)fp;
  }
  HaloHandle h;
  std::memset(&h, 0, sizeof h);
  h.magic  = 0x46563244u;
  h.rank   = c->rank;
  h.nranks = c->nranks;
  h.device = c->device;
  h.pid      = (int64_t)getpid();
  h.plane    = c->kp.L.plane;
  h.ny_local = c->kp.p.Ny;
  void *ptrs[3] = {c->Q[0], c->Q[1], c->sc};
  for (int k = 0; k < 3; ++k)
  {
    FV2D_CUDA(cudaIpcGetMemHandle(&h.mem[k], ptrs[k]));
    CUdeviceptr base = 0;
    size_t size      = 0;
    if (!getRange || getRange(&base, &size, (CUdeviceptr)ptrs[k]) != CUDA_SUCCESS)
      base = (CUdeviceptr)ptrs[k];
    h.offset[k] = (uint64_t)((CUdeviceptr)ptrs[k] - base);
    h.raw[k]    = (uint64_t)(uintptr_t)ptrs[k];
  }
  std::memset(handle, 0, FV2D_IPC_HANDLE_BYTES);
  std::memcpy(handle, &h, sizeof h);
  return FV2D_OK;
}

int fv2d_halo_connect(fv2d_ctx *c, const void *handles, int nranks)
{
  FV2D_ENTER(c);
  if (!handles || nranks != c->nranks)
    return arg_fail("fv2d_halo_connect: need one handle per rank of this context's decomposition");
  if (c->connected)
    return arg_fail("fv2d_halo_connect: already connected");
  const int lo = (c->rank + nranks - 1) % nranks, hi = (c->rank + 1) % nranks;
  const bool need_lo = c->kp.edge_lo == EDGE_NEIGHBOUR, need_hi = c->kp.edge_hi == EDGE_NEIGHBOUR;
  const int64_t mypid = (int64_t)getpid();
  for (int q = 0; q < nranks; ++q)
  {
    HaloHandle h;
    std::memcpy(&h, (const char *)handles + (size_t)q * FV2D_IPC_HANDLE_BYTES, sizeof h);
    if (h.magic != 0x46563244u || h.rank != q || h.nranks != nranks)
      return arg_fail("fv2d_halo_connect: malformed handle for rank " + std::to_string(q));
    if (q == c->rank)
    {
      c->kp.peer_sc[q] = c->sc;
      continue;
    }
    void *ptr[3] = {nullptr, nullptr, nullptr};
    const bool want[3] = {(q == lo && need_lo) || (q == hi && need_hi), (q == lo && need_lo) || (q == hi && need_hi), true};
    for (int k = 0; k < 3; ++k)
    {
      if (!want[k])
        continue;
      if (h.pid == mypid)
      {
        // same process (single-process multi-GPU driver / tests): plain peer access
        if (h.device != c->device)
        {
          cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
          cudaGetLastError();
        }
        ptr[k] = (void *)(uintptr_t)h.raw[k];
      }
      else
      {
        void *base = nullptr;
        FV2D_CUDA(cudaIpcOpenMemHandle(&base, h.mem[k], cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened[c->n_ipc_opened++] = base;
        ptr[k]                           = (char *)base + h.offset[k];
      }
    }
    c->kp.peer_sc[q] = (DevScalars *)ptr[2];
    if (q == lo && need_lo)
    {
      c->peerQ_lo[0]  = (double *)ptr[0];
      c->peerQ_lo[1]  = (double *)ptr[1];
      c->peer_lo_Ny    = h.ny_local;
      c->peer_lo_plane = h.plane;
    }
    if (q == hi && need_hi)
    {
      c->peerQ_hi[0]  = (double *)ptr[0];
      c->peerQ_hi[1]  = (double *)ptr[1];
      c->peer_hi_plane = h.plane;
    }
  }
  c->connected = true;
  return FV2D_OK;
}

} // extern "C"
