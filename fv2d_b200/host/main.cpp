// fv2d_b200_main — host driver: the reference's main.cpp:13-101 on the B200 path.
//
//   fv2d_b200_main <file.ini> [--unfused] [--device N] [--gpus N [--share-devices]] [--max-steps N] [--quiet]
//
// Same loop, same log lines.  Default: fused hot path (one kernel per RK stage, dt resident
// on the device, read back once per step for the loop condition like the reference's
// computeDt sync).  --unfused drives the operator-level kernels one by one exactly like
// main.cpp:66-81 (bit-identical to the reference's Kokkos-OpenMP build).
// --gpus N cuts the grid into N y-slabs, one per GPU starting at --device, all driven by this
// process (fused path; results bitwise those of one GPU).  --share-devices lets slabs wrap around
// the available devices (tests on small grids: every slab's CTAs must be resident together).
#include <chrono>
#include <cstring>
#include <iostream>

#include "IOManager.h"
#include "Init.h"
#include "Operators.h"
#include "SimInfo.h"

using namespace fv2d;

// The same loop on N y-slabs (SlabSet, Operators.h): same log lines, same files; a step is one
// asynchronous sweep launch per slab, the clock and the next dt are read from slab 0.
static int run_slabs(const char *ini, int gpus, int device, bool share_devices, long max_steps, bool quiet, bool unfused)
{
  if (unfused)
    throw std::runtime_error("--unfused drives the operator-level kernels, which work on a single slab: drop --gpus");
  int ndev = 0;
  check(fv2d_device_count(&ndev), "fv2d_device_count");
  if (device < 0 || device >= ndev)
    throw std::runtime_error("--device out of range");
  int usable = ndev - device;
  if (gpus > usable && !share_devices)
    throw std::runtime_error("--gpus " + std::to_string(gpus) + " but only " + std::to_string(usable) +
                             " device(s) from --device on (slabs may share devices with --share-devices: small grids only)");
  usable = std::min(usable, gpus);

  auto params        = readInifile(ini);
  auto device_params = params.device_params;
  SlabSet slabs(params, gpus, device, usable);
  HostArray hQ(device_params.Nty, device_params.Ntx);

  real_t t         = 0.0;
  int ite          = 0;
  real_t next_save = 0.0;
  InitFunctor init(params);
  ComputeDtFunctor computeDt(params);
  IOManager ioManager(params);

  if (params.restart_file != "")
  {
    auto restart_info = ioManager.loadSnapshot(hQ);
    t                 = restart_info.time;
    ite               = restart_info.iteration;
    std::cout << "Restart at iteration " << ite << " and time " << t << std::endl;
    next_save = t + params.save_freq;
    ite++;
  }
  else
    init.init(hQ);
  slabs.upload(hQ);
  slabs.primToCons();
  slabs.setTime(t);

  int next_log = 0;
  long nstep   = 0;
  auto t0      = std::chrono::steady_clock::now();
  double inv[3];
  real_t dt = slabs.computeDt(inv); // also primes the device-resident dt of every slab
  while (t + device_params.epsilon < params.tend && (max_steps < 0 || nstep < max_steps))
  {
    bool save_needed = (t + device_params.epsilon > next_save);
    if (next_log == 0 && !quiet)
    {
      slabs.invDt(inv);
      computeDt.printDiag(std::cout, t, inv);
    }
    if (next_log == 0)
      next_log = params.log_frequency;
    else
      next_log--;
    if (save_needed)
    {
      if (!quiet)
        std::cout << " - Saving at time " << t << std::endl;
      slabs.download(hQ);
      ioManager.saveSolution(hQ, ite++, t);
      next_save += params.save_freq;
    }
    slabs.fusedStepDeviceDt();
    uint64_t c[3];
    slabs.negativeCounts(c);
    printNegatives(std::cout, c);
    slabs.getTime(t, dt);
    ++nstep;
  }
  slabs.sync();
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::cout << "Time at end is " << t << std::endl;
  slabs.download(hQ);
  ioManager.saveSolution(hQ, ite++, t);
  std::cout << nstep << " steps on " << gpus << " y-slabs, " << double(device_params.Nx) * device_params.Ny * nstep / secs / 1e6
            << " Mcell-updates/s (IO included)" << std::endl;
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 2)
  {
    std::cerr << "usage: " << argv[0]
              << " <file.ini> [--unfused] [--device N] [--gpus N [--share-devices]] [--max-steps N] [--quiet]" << std::endl;
    return 2;
  }
  bool unfused = false, quiet = false, share_devices = false;
  int device = 0, gpus = 1;
  long max_steps = -1;
  for (int a = 2; a < argc; ++a)
  {
    if (!std::strcmp(argv[a], "--unfused"))
      unfused = true;
    else if (!std::strcmp(argv[a], "--quiet"))
      quiet = true;
    else if (!std::strcmp(argv[a], "--device") && a + 1 < argc)
      device = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--gpus") && a + 1 < argc)
      gpus = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--share-devices"))
      share_devices = true;
    else if (!std::strcmp(argv[a], "--max-steps") && a + 1 < argc)
      max_steps = std::atol(argv[++a]);
  }

  try
  {
    if (gpus > 1)
      return run_slabs(argv[1], gpus, device, share_devices, max_steps, quiet, unfused);
    auto params        = readInifile(argv[1]);
    auto device_params = params.device_params;

    auto [Q, U] = makeArrays(params, device);

    real_t t         = 0.0;
    int ite          = 0;
    real_t next_save = 0.0;

    InitFunctor init(params);
    UpdateFunctor update(params);
    ComputeDtFunctor computeDt(params);
    IOManager ioManager(params);

    if (params.restart_file != "")
    {
      auto restart_info = ioManager.loadSnapshot(Q);
      t                 = restart_info.time;
      ite               = restart_info.iteration;
      std::cout << "Restart at iteration " << ite << " and time " << t << std::endl;
      next_save = t + params.save_freq;
      ite++;
    }
    else
    {
      HostArray hQ(device_params.Nty, device_params.Ntx);
      init.init(hQ);
      Q.upload(hQ);
    }
    primToCons(Q, U, params);
    check(fv2d_set_time(Q.ctx(), t), "set_time");

    int next_log = 0;
    long nstep   = 0;
    auto t0      = std::chrono::steady_clock::now();

    real_t dt = computeDt.computeDt(Q, params.save_freq, t, false); // also primes the device-resident dt
    while (t + device_params.epsilon < params.tend && (max_steps < 0 || nstep < max_steps))
    {
      bool save_needed = (t + device_params.epsilon > next_save);

      if (unfused)
        dt = computeDt.computeDt(Q, (ite == 0 ? params.save_freq : next_save - t), t, next_log == 0 && !quiet);
      else if (next_log == 0 && !quiet)
      {
        double inv[3];
        check(fv2d_get_inv_dt(Q.ctx(), inv), "get_inv_dt");
        computeDt.printDiag(std::cout, t, inv);
      }
      if (next_log == 0)
        next_log = params.log_frequency;
      else
        next_log--;

      if (save_needed)
      {
        if (!quiet)
          std::cout << " - Saving at time " << t << std::endl;
        ioManager.saveSolution(Q, ite++, t);
        next_save += params.save_freq;
      }

      if (unfused)
      {
        update.update(Q, U, dt);
        consToPrim(U, Q, params);
        checkNegatives(Q, params);
        t += dt;
      }
      else
      {
        update.fused_step_device_dt(Q, U);
        uint64_t c[3];
        check(fv2d_get_negative_counts(Q.ctx(), c, 1), "negative counts");
        printNegatives(std::cout, c);
        check(fv2d_get_time(Q.ctx(), &t, &dt, nullptr), "get_time");
      }
      ++nstep;
    }
    check(fv2d_sync(Q.ctx()), "sync");
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::cout << "Time at end is " << t << std::endl;
    ioManager.saveSolution(Q, ite++, t);
    std::cout << nstep << " steps, " << double(device_params.Nx) * device_params.Ny * nstep / secs / 1e6
              << " Mcell-updates/s (IO included)" << std::endl;
  }
  catch (const std::exception &e)
  {
    std::cerr << "fv2d_b200: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
