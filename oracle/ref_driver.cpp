// TEST INFRASTRUCTURE — not part of the product.
//
// Driver for the UNMODIFIED reference headers (mdelorme/fv2d @680ff34) compiled
// against its vendored Kokkos 4.1.00 OpenMP backend.  It replays the time loop of
// the reference's main.cpp:33-84 without IOManager.h (libhdf5 is not installed),
// and dumps Q0, Q_N, U_N (domain only, [f][j][i] fp64), the dt sequence and the
// domain-integrated mass/energy, so that the C restatement in oracle/ and the CUDA
// path can be pinned against the reference's own arithmetic.
//
// Built by oracle/Makefile into oracle/_ref/fv2d_ref from the sources where they
// lie under /root/reference.  Only tests/, bench.py (cpu_baseline / --impl
// reference) and __graft_entry__.smoke() may run it.
//
// usage: fv2d_ref <file.ini> [--steps N] [--dump out.bin] [--dump-lean out.bin]
//                 [--load-q0 q0.bin] [--bench] [--warmup W] [--quiet]
// (--dump-lean: dt sequence, U_N and the mass/energy sums only - for the full-size grids, where
//  three copies of the state would not fit the box)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "SimInfo.h"

#include "ComputeDt.h"
#include "Init.h"
#include "Update.h"

using namespace fv2d;

namespace
{
void gatherDomain(const Array &A, const DeviceParams &p, std::vector<double> &out)
{
  auto h = Kokkos::create_mirror_view(A);
  Kokkos::deep_copy(h, A);
  out.resize(size_t(4) * p.Nx * p.Ny);
  for (int f = 0; f < 4; ++f)
    for (int j = 0; j < p.Ny; ++j)
      for (int i = 0; i < p.Nx; ++i)
        out[(size_t(f) * p.Ny + j) * p.Nx + i] = h(j + p.jbeg, i + p.ibeg, f);
}

void scatterDomain(Array &A, const DeviceParams &p, const std::vector<double> &in)
{
  auto h = Kokkos::create_mirror_view(A);
  Kokkos::deep_copy(h, A);
  for (int f = 0; f < 4; ++f)
    for (int j = 0; j < p.Ny; ++j)
      for (int i = 0; i < p.Nx; ++i)
        h(j + p.jbeg, i + p.ibeg, f) = in[(size_t(f) * p.Ny + j) * p.Nx + i];
  Kokkos::deep_copy(A, h);
}
} // namespace

int main(int argc, char **argv)
{
  if (argc < 2)
  {
    std::fprintf(stderr, "usage: %s <file.ini> [--steps N] [--dump f] [--load-q0 f] [--bench] [--warmup W] [--quiet]\n",
                 argv[0]);
    return 2;
  }
  std::string ini = argv[1];
  long nsteps     = -1; // -1: run to tend like the reference
  long warmup     = 0;
  std::string dump_path, q0_path;
  bool bench = false, quiet = false, lean = false;
  for (int a = 2; a < argc; ++a)
  {
    std::string s = argv[a];
    if (s == "--steps" && a + 1 < argc)
      nsteps = std::atol(argv[++a]);
    else if (s == "--warmup" && a + 1 < argc)
      warmup = std::atol(argv[++a]);
    else if (s == "--dump" && a + 1 < argc)
      dump_path = argv[++a];
    else if (s == "--dump-lean" && a + 1 < argc)
    {
      dump_path = argv[++a];
      lean      = true;
    }
    else if (s == "--load-q0" && a + 1 < argc)
      q0_path = argv[++a];
    else if (s == "--bench")
      bench = true;
    else if (s == "--quiet")
      quiet = true;
  }

  Kokkos::initialize(argc, argv);
  int rc = 0;
  {
    auto params        = readInifile(ini);
    auto device_params = params.device_params;

    Array U = Kokkos::View<real_t ***>("U", device_params.Nty, device_params.Ntx, Nfields);
    Array Q = Kokkos::View<real_t ***>("Q", device_params.Nty, device_params.Ntx, Nfields);

    real_t t = 0.0;

    InitFunctor init(params);
    UpdateFunctor update(params);
    ComputeDtFunctor computeDt(params);

    init.init(Q);
    if (!q0_path.empty())
    {
      std::vector<double> q0(size_t(4) * device_params.Nx * device_params.Ny);
      FILE *f = std::fopen(q0_path.c_str(), "rb");
      if (!f || std::fread(q0.data(), sizeof(double), q0.size(), f) != q0.size())
      {
        std::fprintf(stderr, "cannot read Q0 from %s\n", q0_path.c_str());
        return 3;
      }
      std::fclose(f);
      scatterDomain(Q, device_params, q0);
      BoundaryManager bc(params);
      bc.fillBoundaries(Q);
    }
    primToCons(Q, U, params);

    std::vector<double> Q0;
    if (!dump_path.empty() && !lean)
      gatherDomain(Q, device_params, Q0);

    if (!quiet)
    {
      std::printf("params: Nx=%d Ny=%d dx=%.17g dy=%.17g gamma0=%.17g CFL=%.17g epsilon=%.17g gx=%.17g gy=%.17g "
                  "kappa=%.17g mu=%.17g tend=%.17g threads=%d\n",
                  device_params.Nx, device_params.Ny, device_params.dx, device_params.dy, device_params.gamma0,
                  device_params.CFL, device_params.epsilon, device_params.gx, device_params.gy, device_params.kappa,
                  device_params.mu, params.tend, Kokkos::DefaultExecutionSpace().concurrency());
    }

    std::vector<double> dts;
    long step   = 0;
    auto t0     = std::chrono::steady_clock::now();
    long timed0 = 0;
    while (t + device_params.epsilon < params.tend && (nsteps < 0 || step < nsteps + warmup))
    {
      if (bench && step == warmup)
      {
        Kokkos::fence();
        t0     = std::chrono::steady_clock::now();
        timed0 = step;
      }
      real_t dt = computeDt.computeDt(Q, params.save_freq, t, false);
      dts.push_back(dt);

      update.update(Q, U, dt);
      consToPrim(U, Q, params);
      checkNegatives(Q, params);

      t += dt;
      ++step;
    }
    Kokkos::fence();
    auto t1 = std::chrono::steady_clock::now();

    if (bench)
    {
      double secs  = std::chrono::duration<double>(t1 - t0).count();
      double cells = double(device_params.Nx) * device_params.Ny * double(step - timed0);
      std::printf("bench: steps=%ld seconds=%.6f mcell_updates_per_s=%.6f threads=%d\n", step - timed0, secs,
                  cells / secs / 1e6, Kokkos::DefaultExecutionSpace().concurrency());
    }

    std::vector<double> QN, UN;
    if (!lean)
      gatherDomain(Q, device_params, QN);
    gatherDomain(U, device_params, UN);
    const size_t n = size_t(device_params.Nx) * device_params.Ny;
    double mass = 0.0, energy = 0.0;
    for (size_t k = 0; k < n; ++k)
    {
      mass += UN[k] * device_params.dx * device_params.dy;
      energy += UN[3 * n + k] * device_params.dx * device_params.dy;
    }
    if (!quiet)
    {
      for (size_t k = 0; k < dts.size() && k < 12; ++k)
        std::printf("dt[%zu]=%.17g\n", k, dts[k]);
      std::printf("steps=%ld t=%.17g mass=%.17g energy=%.17g\n", step, t, mass, energy);
    }

    if (!dump_path.empty())
    {
      FILE *f = std::fopen(dump_path.c_str(), "wb");
      if (!f)
      {
        std::fprintf(stderr, "cannot open %s\n", dump_path.c_str());
        return 4;
      }
      const char magic[8] = {'F', 'V', '2', 'D', lean ? 'L' : 'D', lean ? 'E' : 'U', lean ? 'A' : 'M', lean ? 'N' : 'P'};
      int32_t hdr[4]      = {device_params.Nx, device_params.Ny, int32_t(step), 4};
      std::fwrite(magic, 1, 8, f);
      std::fwrite(hdr, sizeof(int32_t), 4, f);
      std::fwrite(&t, sizeof(double), 1, f);
      std::fwrite(dts.data(), sizeof(double), dts.size(), f);
      if (!lean)
      {
        std::fwrite(Q0.data(), sizeof(double), Q0.size(), f);
        std::fwrite(QN.data(), sizeof(double), QN.size(), f);
      }
      std::fwrite(UN.data(), sizeof(double), UN.size(), f);
      double sums[2] = {mass, energy};
      std::fwrite(sums, sizeof(double), 2, f);
      std::fclose(f);
    }
  }
  Kokkos::finalize();
  return rc;
}
