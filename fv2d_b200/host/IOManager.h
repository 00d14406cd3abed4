// IOManager — snapshot output / restart for the B200 host driver: the reference's class
// (reference IOManager.h:75-399) with the same constructor side effects (creates the output
// directory, dumps the effective .ini to last.ini and <path>/<name>.ini, :82-95), the same
// saveSolution(Q, iteration, t) / loadSnapshot(Q) -> RestartInfo interface and the same files
// on disk: run.h5 (+ run.xmf), or one .h5/.xmf pair per snapshot with run.multiple_outputs.
// The HDF5 container is written by H5Lite.h (libhdf5 / HighFive are not available here);
// the formats themselves live in SnapshotIO.h, on host arrays.
#pragma once

#include <filesystem>
#include <fstream>

#include "Operators.h"
#include "SnapshotIO.h"

namespace fv2d
{

inline SnapshotConfig makeSnapshotConfig(const Params &params)
{
  SnapshotConfig c;
  c.device_params    = params.device_params;
  c.output_path      = params.output_path;
  c.filename_out     = params.filename_out;
  c.restart_file     = params.restart_file;
  c.problem          = params.problem;
  c.multiple_outputs = params.multiple_outputs;
  c.tend             = params.tend;
  return c;
}

class IOManager
{
public:
  Params params;
  DeviceParams &device_params;
  bool force_file_truncation = false;

  explicit IOManager(Params &p) : params(p), device_params(params.device_params)
  {
    if (!std::filesystem::exists(params.output_path))
    {
      std::cout << "Output path does not exist, creating directory `" << params.output_path << "`." << std::endl;
      std::filesystem::create_directory(params.output_path);
    }
    std::ofstream out_ini_local("last.ini");
    params.reader.outputValues(out_ini_local);
    std::ofstream out_ini(params.output_path + "/" + params.filename_out + ".ini");
    params.reader.outputValues(out_ini);
  }

  void saveSolution(const Array &Q, int iteration, real_t t)
  {
    HostArray h(device_params.Nty, device_params.Ntx);
    Q.download(h); // Kokkos::deep_copy(Qhost, Q), IOManager.h:113-114 / :196-197
    saveSolutionHost(makeSnapshotConfig(params), h, iteration, t, force_file_truncation);
  }

  // the same on a host array (what the multi-GPU driver assembles from its slabs)
  void saveSolution(const HostArray &h, int iteration, real_t t)
  {
    saveSolutionHost(makeSnapshotConfig(params), h, iteration, t, force_file_truncation);
  }
  RestartInfo loadSnapshot(HostArray &h)
  {
    const RestartInfo info = loadSnapshotHost(makeSnapshotConfig(params), h, force_file_truncation);
    if (force_file_truncation)
      saveSolution(h, info.iteration, info.time);
    return info;
  }

  RestartInfo loadSnapshot(Array &Q)
  {
    HostArray h(device_params.Nty, device_params.Ntx);
    const RestartInfo info = loadSnapshotHost(makeSnapshotConfig(params), h, force_file_truncation);
    Q.upload(h);
    if (force_file_truncation) // restarting into another file: start it with the loaded state
      saveSolution(Q, info.iteration, info.time);
    return info;
  }
};

} // namespace fv2d
