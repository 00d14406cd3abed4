// Operators — C++17 host mirror of the reference's operator surface over the C ABI.
//
// Same class names, constructor arguments, method names and call semantics as the reference
// (SURVEY.md §8b), so a driver written against mdelorme/fv2d's headers reads the same:
//   Array                          <- Kokkos::View<real_t***>           (SimInfo.h:17, main.cpp:33-34)
//   UpdateFunctor::update etc.     <- Update.h:40-222
//   ComputeDtFunctor::computeDt    <- ComputeDt.h:10-65
//   BoundaryManager::fillBoundaries<- BoundaryConditions.h:74-147
//   ThermalConductionFunctor       <- ThermalConduction.h:28-108
//   ViscosityFunctor               <- Viscosity.h:19-119
//   consToPrim / primToCons / checkNegatives <- SimInfo.h:576-646
// Difference that matters: the reference allocates Q and U separately; here ONE device
// context (fv2d_ctx) owns the pair, so `Array Q` and `Array U` are two views of the same
// context and every operator checks that it was handed views of one context.
// Errors of the C ABI become std::runtime_error, like the reference's configuration errors.
#pragma once

#include <algorithm>
#include <array>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fv2d_b200.h"
#include "Init.h"
#include "SimInfo.h"

namespace fv2d
{

inline void check(int rc, const char *what)
{
  if (rc != FV2D_OK)
    throw std::runtime_error(std::string(what) + ": " + fv2d_last_error());
}

// Shared owner of the device context.
struct DeviceState
{
  fv2d_ctx *ctx = nullptr;
  DeviceState(const Params &params, int device = 0)
  {
    check(fv2d_ctx_create(&params.device_params, params.time_stepping, params.epsilon_reset_negative, device, &ctx),
          "fv2d_ctx_create");
  }
  ~DeviceState() { fv2d_ctx_destroy(ctx); }
  DeviceState(const DeviceState &)            = delete;
  DeviceState &operator=(const DeviceState &) = delete;
};

// A view of the primitive (Q) or conservative (U) array of a device context.  Copies are
// shallow, like Kokkos Views.
struct Array
{
  enum Kind { PRIMITIVE, CONSERVATIVE };
  std::shared_ptr<DeviceState> state;
  Kind kind = PRIMITIVE;
  fv2d_ctx *ctx() const { return state->ctx; }

  void upload(const HostArray &h) const
  {
    check(kind == PRIMITIVE ? fv2d_upload_Q(ctx(), h.data.data()) : fv2d_upload_U(ctx(), h.data.data()), "upload");
  }
  void download(HostArray &h) const
  {
    check(kind == PRIMITIVE ? fv2d_download_Q(ctx(), h.data.data()) : fv2d_download_U(ctx(), h.data.data()), "download");
  }
};

// Allocates the Q/U pair of main.cpp:33-34 (zero-filled) on the device.
inline std::pair<Array, Array> makeArrays(const Params &params, int device = 0)
{
  auto st = std::make_shared<DeviceState>(params, device);
  return {Array{st, Array::PRIMITIVE}, Array{st, Array::CONSERVATIVE}};
}

inline void requirePair(const Array &Q, const Array &U, const char *who)
{
  if (!Q.state || Q.state != U.state || Q.kind != Array::PRIMITIVE || U.kind != Array::CONSERVATIVE)
    throw std::runtime_error(std::string(who) + ": Q and U must be the primitive/conservative views of one context");
}

class BoundaryManager
{
public:
  Params full_params;
  explicit BoundaryManager(const Params &p) : full_params(p) {}
  void fillBoundaries(Array Q) { check(fv2d_fill_boundaries(Q.ctx()), "fillBoundaries"); }
};

class ThermalConductionFunctor
{
public:
  Params full_params;
  explicit ThermalConductionFunctor(const Params &p) : full_params(p) {}
  void applyThermalConduction(Array Q, Array Unew, real_t dt)
  {
    requirePair(Q, Unew, "applyThermalConduction");
    check(fv2d_apply_thermal_conduction(Q.ctx(), dt), "applyThermalConduction");
  }
};

class ViscosityFunctor
{
public:
  Params full_params;
  explicit ViscosityFunctor(const Params &p) : full_params(p) {}
  void applyViscosity(Array Q, Array Unew, real_t dt)
  {
    requirePair(Q, Unew, "applyViscosity");
    check(fv2d_apply_viscosity(Q.ctx(), dt), "applyViscosity");
  }
};

class UpdateFunctor
{
public:
  Params full_params;
  BoundaryManager bc_manager;
  ThermalConductionFunctor tc_functor;
  ViscosityFunctor visc_functor;

  explicit UpdateFunctor(const Params &p) : full_params(p), bc_manager(p), tc_functor(p), visc_functor(p) {}

  void computeSlopes(const Array &Q) const { check(fv2d_compute_slopes(Q.ctx()), "computeSlopes"); }
  void computeFluxesAndUpdate(Array Q, Array Unew, real_t dt) const
  {
    requirePair(Q, Unew, "computeFluxesAndUpdate");
    check(fv2d_compute_fluxes_and_update(Q.ctx(), dt), "computeFluxesAndUpdate");
  }
  void euler_step(Array Q, Array Unew, real_t dt)
  {
    requirePair(Q, Unew, "euler_step");
    check(fv2d_euler_step(Q.ctx(), dt), "euler_step");
  }
  // Operator-by-operator update, bit-identical to the reference (Update.h:193-222).
  void update(Array Q, Array Unew, real_t dt)
  {
    requirePair(Q, Unew, "update");
    check(fv2d_update(Q.ctx(), dt), "update");
  }
  // The fused hot path: update + consToPrim + checkNegatives + the next computeDt in one
  // kernel per Runge-Kutta stage (main.cpp:79-81 and :66 of the next iteration).
  void fused_step(Array Q, Array Unew, real_t dt)
  {
    requirePair(Q, Unew, "fused_step");
    check(fv2d_step(Q.ctx(), dt), "fused_step");
  }
  void fused_step_device_dt(Array Q, Array Unew)
  {
    requirePair(Q, Unew, "fused_step_device_dt");
    check(fv2d_step_device_dt(Q.ctx()), "fused_step_device_dt");
  }
};

class ComputeDtFunctor
{
public:
  Params full_params;
  explicit ComputeDtFunctor(const Params &p) : full_params(p) {}

  // max_dt is accepted and ignored, like the reference (Q6).
  real_t computeDt(Array Q, real_t /*max_dt*/, real_t t, bool diag) const
  {
    double dt = 0.0, inv[3] = {0.0, 0.0, 0.0};
    check(fv2d_compute_dt(Q.ctx(), &dt, inv), "computeDt");
    if (diag)
      printDiag(std::cout, t, inv);
    return dt;
  }
  // log line of ComputeDt.h:54-62
  void printDiag(std::ostream &o, real_t t, const double inv[3]) const
  {
    const auto &params = full_params.device_params;
    o << "Computing dts at (t=" << t << ") : dt_hyp=" << 1.0 / inv[0];
    if (params.thermal_conductivity_active)
      o << "; dt_TC=" << 1.0 / inv[1];
    if (params.viscosity_active)
      o << "; dt_visc=" << 1.0 / inv[2];
    o << std::endl;
  }
};

inline void consToPrim(Array U, Array Q, const Params &)
{
  requirePair(Q, U, "consToPrim");
  check(fv2d_cons_to_prim(Q.ctx()), "consToPrim");
}
inline void primToCons(const Array &Q, const Array &U, const Params &)
{
  requirePair(Q, U, "primToCons");
  check(fv2d_prim_to_cons(Q.ctx()), "primToCons");
}
// prints like SimInfo.h:637-645
inline void printNegatives(std::ostream &o, const uint64_t c[3])
{
  if (c[0])
    o << "--> negative density: " << c[0] << std::endl;
  if (c[1])
    o << "--> negative pressure: " << c[1] << std::endl;
  if (c[2])
    o << "--> NaN detected." << std::endl;
}
inline void checkNegatives(const Array &Q, const Params &)
{
  uint64_t c[3] = {0, 0, 0};
  check(fv2d_check_negatives(Q.ctx(), c), "checkNegatives");
  printNegatives(std::cout, c);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU: the grid as y-slabs, one device context per slab, all driven by this one host
// process (no counterpart in the reference, which is single-device: main.cpp:13-101).  After
// connect() the slabs exchange their ghost rows and the global CFL maximum from inside the sweep
// kernels over peer mappings: a step is one asynchronous launch per slab, no host round trip.
class SlabSet
{
public:
  Params params;
  std::vector<fv2d_ctx *> ctx;

  // slab r runs on device first_device + (r mod ndevices)
  SlabSet(const Params &p, int nslabs, int first_device, int ndevices) : params(p)
  {
    if (nslabs < 1 || ndevices < 1)
      throw std::runtime_error("SlabSet: need at least one slab and one device");
    ctx.assign(size_t(nslabs), nullptr);
    for (int r = 0; r < nslabs; ++r)
      check(fv2d_ctx_create_slab(&params.device_params, params.time_stepping, params.epsilon_reset_negative,
                                 first_device + r % ndevices, r, nslabs, &ctx[size_t(r)]),
            "fv2d_ctx_create_slab");
    std::vector<unsigned char> handles(size_t(nslabs) * FV2D_IPC_HANDLE_BYTES);
    for (int r = 0; r < nslabs; ++r)
      check(fv2d_halo_export(ctx[size_t(r)], handles.data() + size_t(r) * FV2D_IPC_HANDLE_BYTES), "fv2d_halo_export");
    for (int r = 0; r < nslabs; ++r)
      check(fv2d_halo_connect(ctx[size_t(r)], handles.data(), nslabs), "fv2d_halo_connect");
  }
  ~SlabSet()
  {
    for (fv2d_ctx *c : ctx)
      if (c)
        fv2d_sync(c);
    for (fv2d_ctx *c : ctx)
      fv2d_ctx_destroy(c);
  }
  SlabSet(const SlabSet &)            = delete;
  SlabSet &operator=(const SlabSet &) = delete;

  // rows [first, first + n) of the global array (ghost rows included) that slab r holds
  void rowsOf(size_t r, int &first, int &n) const
  {
    int64_t g[6];
    check(fv2d_ctx_geometry(ctx[r], g), "fv2d_ctx_geometry");
    first = int(g[3]);
    n     = int(g[1]);
  }
  // global host array -> slabs (each with its ghost rows, taken from the neighbours' rows) and back
  // (the result's ghost rows are the outer slabs' own)
  void upload(const HostArray &h) const
  {
    for (size_t r = 0; r < ctx.size(); ++r)
    {
      int first, n;
      rowsOf(r, first, n);
      HostArray s(n, h.Ntx);
      for (int f = 0; f < Nfields; ++f)
        std::copy(&h(first, 0, f), &h(first, 0, f) + size_t(n) * h.Ntx, &s(0, 0, f));
      check(fv2d_upload_Q(ctx[r], s.data.data()), "upload");
    }
  }
  void download(HostArray &h) const
  {
    const int Ng = params.device_params.Ng;
    for (size_t r = 0; r < ctx.size(); ++r)
    {
      int first, n;
      rowsOf(r, first, n);
      HostArray s(n, h.Ntx);
      check(fv2d_download_Q(ctx[r], s.data.data()), "download");
      const int lo = (r == 0) ? 0 : Ng, hi = (r + 1 == ctx.size()) ? n : n - Ng; // own rows (+ the outer ghosts)
      for (int f = 0; f < Nfields; ++f)
        std::copy(&s(lo, 0, f), &s(lo, 0, f) + size_t(hi - lo) * h.Ntx, &h(first + lo, 0, f));
    }
  }
  void primToCons() const
  {
    for (fv2d_ctx *c : ctx)
      check(fv2d_prim_to_cons(c), "primToCons");
  }
  void setTime(real_t t) const
  {
    for (fv2d_ctx *c : ctx)
      check(fv2d_set_time(c, t), "set_time");
  }
  // ComputeDtFunctor::computeDt over all slabs.  fv2d_compute_dt is collective and synchronises the
  // host, so every slab gets its own host thread for the call.
  real_t computeDt(double inv[3]) const
  {
    std::vector<double> dt(ctx.size(), 0.0);
    std::vector<std::string> err(ctx.size());
    std::vector<std::array<double, 3>> iv(ctx.size());
    std::vector<std::thread> th;
    for (size_t r = 0; r < ctx.size(); ++r)
      th.emplace_back([&, r] {
        if (fv2d_compute_dt(ctx[r], &dt[r], iv[r].data()) != FV2D_OK)
          err[r] = fv2d_last_error();
      });
    for (auto &t : th)
      t.join();
    for (const auto &e : err)
      if (!e.empty())
        throw std::runtime_error("computeDt: " + e);
    for (int k = 0; k < 3; ++k)
      inv[k] = iv[0][size_t(k)];
    return dt[0];
  }
  // one fused step on every slab (asynchronous: the slabs' sweeps run side by side)
  void fusedStepDeviceDt() const
  {
    for (fv2d_ctx *c : ctx)
      check(fv2d_step_device_dt(c), "fused_step_device_dt");
  }
  void negativeCounts(uint64_t total[3]) const
  {
    total[0] = total[1] = total[2] = 0;
    for (fv2d_ctx *c : ctx)
    {
      uint64_t k[3];
      check(fv2d_get_negative_counts(c, k, 1), "negative counts");
      for (int i = 0; i < 3; ++i)
        total[i] += k[i];
    }
  }
  void getTime(real_t &t, real_t &next_dt) const { check(fv2d_get_time(ctx[0], &t, &next_dt, nullptr), "get_time"); }
  void invDt(double inv[3]) const { check(fv2d_get_inv_dt(ctx[0], inv), "get_inv_dt"); }
  void sync() const
  {
    for (fv2d_ctx *c : ctx)
      check(fv2d_sync(c), "sync");
  }
};

} // namespace fv2d
