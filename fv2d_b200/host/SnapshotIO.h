// SnapshotIO — the reference IOManager's file formats (reference IOManager.h:99-398) on host
// arrays: run.h5 (HDF5, through the dependency-free h5lite writer) + the XDMF sidecar, and
// restart loading.  IOManager.h wraps this with the device download/upload; the C ABI exposes
// it as fv2d_io_save_solution / fv2d_io_load_snapshot.
//
// On-disk layout, as the reference writes it:
//   unique-file mode (default, IOManager.h:191-282): <path>/<name>.h5 holds root attributes
//     Ntx, Nty, Nx, Ny, ibeg, iend, jbeg, jend (int) and problem (string), root datasets x, y
//     (vertex coordinates, length (Nx+1)(Ny+1), j-major) and one group ite_%04d per snapshot
//     with datasets rho, u, v, prs (length Nx*Ny, j-major) and attributes time (double),
//     iteration (int); <name>.xmf lists every snapshot (the footer is rewritten in place);
//   multiple-file mode (IOManager.h:107-189): <name>_%04d.h5 / .xmf per snapshot with
//     everything at the root.
#pragma once

#include <cstdio>
#include <filesystem>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

#include "H5Lite.h"
#include "Init.h"
#include "SimInfo.h"

namespace fv2d
{

constexpr int ite_nzeros              = 4;      // IOManager.h:20
constexpr const char *ite_prefix      = "ite_"; // IOManager.h:21

// What the IO layer needs from Params (IOManager.h:78-80).
struct SnapshotConfig
{
  fv2d_device_params device_params;
  std::string output_path, filename_out, restart_file, problem;
  bool multiple_outputs = false;
  real_t tend           = 0.0;
};

namespace xdmf
{
// The XDMF text of IOManager.h:26-71, produced piece by piece (same bytes).
inline std::string header(const fv2d_device_params &p, const std::string &h5_filename)
{
  std::ostringstream o;
  o << "<?xml version=\"1.0\" ?>\n"
    << "<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" [\n"
    << "<!ENTITY file \"" << h5_filename << ":\">\n"
    << "<!ENTITY fdim \"" << p.Ny << " " << p.Nx << "\">\n"
    << "<!ENTITY gdim \"" << p.Ny + 1 << " " << p.Nx + 1 << "\">\n"
    << "<!ENTITY GridEntity '\n"
    << "<Topology TopologyType=\"2DSMesh\" Dimensions=\"&gdim;\"/>\n"
    << "<Geometry GeometryType=\"X_Y\">\n"
    << "  <DataItem Dimensions=\"&gdim;\" NumberType=\"Float\" Precision=\"8\" Format=\"HDF\">&file;/x</DataItem>\n"
    << "  <DataItem Dimensions=\"&gdim;\" NumberType=\"Float\" Precision=\"8\" Format=\"HDF\">&file;/y</DataItem>\n"
    << "</Geometry>'>\n"
    << "]>\n"
    << "<Xdmf Version=\"3.0\">\n"
    << "<Domain>\n"
    << "  <Grid Name=\"TimeSeries\" GridType=\"Collection\" CollectionType=\"Temporal\">\n"
    << "    ";
  return o.str();
}
inline std::string footer() { return "\n  </Grid>\n</Domain>\n</Xdmf>"; }
inline std::string iteHeader(const std::string &name, real_t time)
{
  char tbuf[64];
  std::snprintf(tbuf, sizeof tbuf, "%lf", time);
  return "\n    <Grid Name=\"" + name + "\" GridType=\"Uniform\">\n      <Time Value=\"" + tbuf +
         "\" />\n      &GridEntity;";
}
inline std::string dataItem(const std::string &group, const std::string &field)
{
  return "<DataItem Dimensions=\"&fdim;\" NumberType=\"Float\" Precision=\"8\" Format=\"HDF\">&file;/" + group + field +
         "</DataItem>";
}
inline std::string scalarField(const std::string &group, const std::string &field)
{
  return "\n      <Attribute Name=\"" + field + "\" AttributeType=\"Scalar\" Center=\"Cell\">\n        " +
         dataItem(group, field) + "\n      </Attribute>";
}
inline std::string vectorField(const std::string &group, const std::string &name, const std::string &fx,
                               const std::string &fy)
{
  return "\n      <Attribute Name=\"" + name + "\" AttributeType=\"Vector\" Center=\"Cell\">\n" +
         "        <DataItem Dimensions=\"&fdim; 2\" ItemType=\"Function\" Function=\"JOIN($0, $1)\">\n          " +
         dataItem(group, fx) + "\n          " + dataItem(group, fy) + "\n        </DataItem>\n      </Attribute>";
}
inline std::string iteFooter() { return "\n    </Grid>\n    "; }
inline std::string iteration(const std::string &group_prefix, const std::string &name, real_t t)
{
  return iteHeader(name, t) + scalarField(group_prefix, "rho") + vectorField(group_prefix, "velocity", "u", "v") +
         scalarField(group_prefix, "prs") + iteFooter() + footer();
}
} // namespace xdmf

namespace snapshot_detail
{
inline std::string iterationName(const std::string &prefix, int iteration)
{
  std::ostringstream oss;
  oss << prefix << std::setw(ite_nzeros) << std::setfill('0') << iteration;
  return oss.str();
}
inline void writeMeshAndRootAttributes(h5lite::Object &file, const SnapshotConfig &c)
{
  const auto &d = c.device_params;
  file.createAttribute("Ntx", d.Ntx);
  file.createAttribute("Nty", d.Nty);
  file.createAttribute("Nx", d.Nx);
  file.createAttribute("Ny", d.Ny);
  file.createAttribute("ibeg", d.ibeg);
  file.createAttribute("iend", d.iend);
  file.createAttribute("jbeg", d.jbeg);
  file.createAttribute("jend", d.jend);
  file.createAttribute("problem", c.problem);
  std::vector<real_t> x, y; // vertex positions (IOManager.h:227-236)
  x.reserve(size_t(d.Nx + 1) * (d.Ny + 1));
  y.reserve(size_t(d.Nx + 1) * (d.Ny + 1));
  for (int j = d.jbeg; j <= d.jend; ++j)
    for (int i = d.ibeg; i <= d.iend; ++i)
    {
      x.push_back((i - d.ibeg) * d.dx + d.xmin);
      y.push_back((j - d.jbeg) * d.dy + d.ymin);
    }
  file.createDataSet("x", x);
  file.createDataSet("y", y);
}
inline void writeFields(h5lite::Object &where, const SnapshotConfig &c, const HostArray &Q, int iteration, real_t t)
{
  const auto &d             = c.device_params;
  const char *names[Nfields] = {"rho", "u", "v", "prs"};
  const int order[Nfields]   = {IR, IU, IV, IP};
  std::vector<real_t> table(size_t(d.Nx) * d.Ny);
  for (int f = 0; f < Nfields; ++f)
  {
    size_t lid = 0;
    for (int j = d.jbeg; j < d.jend; ++j)
      for (int i = d.ibeg; i < d.iend; ++i)
        table[lid++] = Q(j, i, order[f]);
    where.createDataSet(names[f], table);
  }
  where.createAttribute("time", t);
  where.createAttribute("iteration", iteration);
}
inline FILE *openXdmf(const std::string &path, const char *mode)
{
  FILE *fd = std::fopen(path.c_str(), mode);
  if (fd == nullptr)
    throw std::runtime_error("Failed to open XDMF file '" + path + "' with mode '" + mode + "'.");
  return fd;
}
} // namespace snapshot_detail

// IOManager::saveSolutionMultiple (IOManager.h:107-189)
inline void saveSolutionMultipleHost(const SnapshotConfig &c, const HostArray &Q, int iteration, real_t t)
{
  using namespace snapshot_detail;
  const std::string iteration_str = iterationName(c.filename_out + "_", iteration);
  const std::string h5_filename = iteration_str + ".h5", xmf_filename = iteration_str + ".xmf";
  const std::string output_path = c.output_path + "/";

  h5lite::File file(output_path + h5_filename, h5lite::File::Truncate);
  FILE *xdmf_fd = openXdmf(output_path + xmf_filename, "w+");
  writeMeshAndRootAttributes(file, c);
  writeFields(file, c, Q, iteration, t);
  file.close();

  const std::string text = xdmf::header(c.device_params, h5_filename) + xdmf::iteration("", iteration_str, t);
  std::fwrite(text.data(), 1, text.size(), xdmf_fd);
  std::fclose(xdmf_fd);
}

// IOManager::saveSolutionUnique (IOManager.h:191-282)
inline void saveSolutionUniqueHost(const SnapshotConfig &c, const HostArray &Q, int iteration, real_t t,
                                   bool &force_file_truncation)
{
  using namespace snapshot_detail;
  const std::string iteration_str = iterationName(ite_prefix, iteration);
  const std::string h5_filename = c.filename_out + ".h5", xmf_filename = c.filename_out + ".xmf";
  const std::string output_path = c.output_path + "/";

  force_file_truncation = (force_file_truncation || iteration == 0);
  const bool truncate   = force_file_truncation;

  h5lite::File file(output_path + h5_filename, truncate ? h5lite::File::Truncate : h5lite::File::ReadWrite);
  FILE *xdmf_fd = openXdmf(output_path + xmf_filename, truncate ? "w+" : "r+");
  if (truncate)
  {
    force_file_truncation = false;
    writeMeshAndRootAttributes(file, c);
    const std::string head = xdmf::header(c.device_params, h5_filename) + xdmf::footer();
    std::fwrite(head.data(), 1, head.size(), xdmf_fd);
  }
  h5lite::Object &ite_group = file.createGroup(iteration_str);
  writeFields(ite_group, c, Q, iteration, t);
  file.close();

  // the reference seeks back over sizeof(footer) — the footer AND its terminating NUL, i.e. one
  // byte more than the footer text (IOManager.h:274) — and appends the new grid + footer
  std::fseek(xdmf_fd, -(long)(xdmf::footer().size() + 1), SEEK_END);
  const std::string text = xdmf::iteration(iteration_str + "/", iteration_str, t);
  std::fwrite(text.data(), 1, text.size(), xdmf_fd);
  std::fclose(xdmf_fd);
}

// IOManager::saveSolution (IOManager.h:99-105)
inline void saveSolutionHost(const SnapshotConfig &c, const HostArray &Q, int iteration, real_t t,
                             bool &force_file_truncation)
{
  if (c.multiple_outputs)
    saveSolutionMultipleHost(c, Q, iteration, t);
  else
    saveSolutionUniqueHost(c, Q, iteration, t, force_file_truncation);
}

// IOManager::loadSnapshot (IOManager.h:284-398) up to and including the ghost fill; the caller
// uploads Q and, when force_file_truncation came back true, re-saves the loaded state
// (IOManager.h:391-395).
inline RestartInfo loadSnapshotHost(const SnapshotConfig &c, HostArray &Q, bool &force_file_truncation)
{
  const auto &d = c.device_params;
  // 'run.h5:/ite_0005' selects an iteration; 'run.h5' alone means the last one
  std::string restart_file = c.restart_file;
  std::string group        = "";
  const auto delim_multi   = restart_file.find(".h5:/");
  if (delim_multi != std::string::npos)
  {
    group = restart_file.substr(delim_multi + 5);
    restart_file.resize(delim_multi + 3);
  }
  std::error_code ec;
  if (!c.multiple_outputs &&
      std::filesystem::equivalent(restart_file, c.output_path + "/" + c.filename_out + ".h5", ec))
  {
    if (delim_multi != std::string::npos)
    {
      std::cerr << "Invalid restart file : if your restart file and output file are "
                   "the same, you can only start from the last iteration."
                << std::endl
                << std::endl;
      throw std::runtime_error("ERROR : Invalid restart_file.");
    }
  }
  else
    force_file_truncation = true;

  h5lite::File file(restart_file, h5lite::File::ReadOnly);
  real_t time;
  int iteration;
  if (file.hasAttribute("time"))
  {
    file.readAttribute("time", time);
    file.readAttribute("iteration", iteration);
  }
  else
  {
    if (group == "")
    {
      // names are indexed in increasing order: ite_0000 ... ite_NNNN, x, y
      if (file.getNumberObjects() < 3)
        throw std::runtime_error("ERROR : restart file holds no iteration.");
      group = file.getObjectName(file.getNumberObjects() - 3);
    }
    const h5lite::Object &h5_group = file.getGroup(group);
    h5_group.readAttribute("time", time);
    h5_group.readAttribute("iteration", iteration);
    group = group + "/";
  }

  const auto Nt = file.getShape(group + "rho")[0];
  if (Nt != uint64_t(d.Nx) * uint64_t(d.Ny))
  {
    std::cerr << "Attempting to restart with a different resolution ! Ncells (restart) = " << Nt
              << "; Run resolution = " << d.Nx << "x" << d.Ny << "=" << d.Nx * d.Ny << std::endl;
    throw std::runtime_error("ERROR : Trying to restart from a file with a different resolution !");
  }

  std::cout << "Loading restart data from hdf5" << std::endl;
  auto load_and_copy = [&](const std::string &var_name, int var_id) {
    const auto table = file.load(group + var_name);
    if (table.size() != size_t(d.Nx) * d.Ny)
      throw std::runtime_error("ERROR : dataset " + group + var_name + " has the wrong size.");
    size_t lid = 0;
    for (int y = 0; y < d.Ny; ++y)
      for (int x = 0; x < d.Nx; ++x)
        Q(y + d.jbeg, x + d.ibeg, var_id) = table[lid++];
  };
  load_and_copy("rho", IR);
  load_and_copy("u", IU);
  load_and_copy("v", IV);
  load_and_copy("prs", IP);
  file.close();

  fillBoundariesHost(d, Q);

  if (time + d.epsilon > c.tend)
  {
    std::cerr << "Restart time is greater than end time : " << std::endl
              << "  time: " << time << "\ttend: " << c.tend << std::endl
              << std::endl;
    throw std::runtime_error("ERROR : restart time is greater than the end time.");
  }
  std::cout << "Restart finished !" << std::endl;
  return {time, iteration};
}

} // namespace fv2d
