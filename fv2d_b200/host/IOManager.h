// IOManager — snapshot output / restart for the B200 host driver.
//
// The reference writes HDF5 (+XDMF) through HighFive (reference IOManager.h:75-399).  libhdf5
// is not available in this environment, so this round keeps the reference's interface and
// bookkeeping (constructor dumps the effective .ini to last.ini and <path>/<name>.ini,
// IOManager.h:90-94; saveSolution(Q, ite, t); loadSnapshot(Q) -> RestartInfo) on a raw binary
// container with the SAME logical content and ordering as a run.h5 iteration group:
// 1-D arrays rho,u,v,prs of length Nx*Ny, j-major (IOManager.h:248-262), plus time and
// iteration.  File layout (little endian): magic "FV2DSNAP", int32 Nx, Ny, iteration, pad,
// double time, then rho[], u[], v[], prs[].  One file per snapshot: <path>/<name>_%04d.bin.
// The HDF5 writer is a "next" row of SURVEY.md §8f and is out of scope for this round.
#pragma once

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "Operators.h"

namespace fv2d
{

class IOManager
{
public:
  Params params;

  explicit IOManager(Params &p) : params(p)
  {
    std::ofstream out_ini("last.ini");
    params.reader.outputValues(out_ini);
    std::ofstream out_ini_local(params.output_path + "/" + params.filename_out + ".ini");
    params.reader.outputValues(out_ini_local);
  }

  std::string snapshotName(int iteration) const
  {
    std::ostringstream oss;
    oss << params.output_path << "/" << params.filename_out << "_" << std::setw(4) << std::setfill('0') << iteration
        << ".bin";
    return oss.str();
  }

  void saveSolution(const Array &Q, int iteration, real_t t)
  {
    const auto &d = params.device_params;
    HostArray h(d.Nty, d.Ntx);
    Q.download(h);
    FILE *f = std::fopen(snapshotName(iteration).c_str(), "wb");
    if (!f)
      throw std::runtime_error("cannot open " + snapshotName(iteration));
    const char magic[8] = {'F', 'V', '2', 'D', 'S', 'N', 'A', 'P'};
    int32_t hdr[4]      = {d.Nx, d.Ny, iteration, 0};
    std::fwrite(magic, 1, 8, f);
    std::fwrite(hdr, sizeof(int32_t), 4, f);
    std::fwrite(&t, sizeof(double), 1, f);
    std::vector<double> row(d.Nx);
    for (int fld = 0; fld < Nfields; ++fld)
      for (int j = d.jbeg; j < d.jend; ++j)
      {
        for (int i = d.ibeg; i < d.iend; ++i)
          row[i - d.ibeg] = h(j, i, fld);
        std::fwrite(row.data(), sizeof(double), row.size(), f);
      }
    std::fclose(f);
  }

  // Reads params.restart_file (a snapshot written by saveSolution), fills the ghosts
  // (IOManager.h:378-379) and refuses to restart past tend (IOManager.h:381-387).
  RestartInfo loadSnapshot(Array &Q)
  {
    const auto &d = params.device_params;
    FILE *f       = std::fopen(params.restart_file.c_str(), "rb");
    if (!f)
      throw std::runtime_error("Restart file " + params.restart_file + " cannot be opened");
    char magic[8];
    int32_t hdr[4];
    double t = 0.0;
    if (std::fread(magic, 1, 8, f) != 8 || std::fread(hdr, sizeof(int32_t), 4, f) != 4 ||
        std::fread(&t, sizeof(double), 1, f) != 1 || std::string(magic, 8) != "FV2DSNAP")
    {
      std::fclose(f);
      throw std::runtime_error("Restart file " + params.restart_file + " is not a fv2d-b200 snapshot");
    }
    if (hdr[0] * hdr[1] != d.Nx * d.Ny) // IOManager.h:343-351
    {
      std::fclose(f);
      throw std::runtime_error("Attempting to restart with a different resolution !");
    }
    HostArray h(d.Nty, d.Ntx);
    std::vector<double> row(d.Nx);
    for (int fld = 0; fld < Nfields; ++fld)
      for (int j = d.jbeg; j < d.jend; ++j)
      {
        if (std::fread(row.data(), sizeof(double), row.size(), f) != row.size())
        {
          std::fclose(f);
          throw std::runtime_error("Restart file " + params.restart_file + " is truncated");
        }
        for (int i = d.ibeg; i < d.iend; ++i)
          h(j, i, fld) = row[i - d.ibeg];
      }
    std::fclose(f);
    fillBoundariesHost(d, h);
    Q.upload(h);
    if (t + d.epsilon > params.tend)
      throw std::runtime_error("Restart time is greater than end time : restart time = " + std::to_string(t) +
                               "; tend = " + std::to_string(params.tend));
    return RestartInfo{t, hdr[2]};
  }
};

} // namespace fv2d
