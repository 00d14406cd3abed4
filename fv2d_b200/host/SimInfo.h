// SimInfo — run parameters for the B200 host driver.
//
// Mirrors the reference's configuration surface (reference SimInfo.h:100-569) so that an
// .ini written for mdelorme/fv2d drives this code unchanged:
//   Reader        <- SimInfo.h:123-263   (records every queried key with provenance)
//   DeviceParams  <- SimInfo.h:266-460   (here: the C POD fv2d_device_params + the reader)
//   Params        <- SimInfo.h:463-492
//   readInifile   <- SimInfo.h:529-568
//   checkValidityIni <- SimInfo.h:501-527
// Quirks that parity depends on are reproduced and labelled.
#pragma once

#include <cmath>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "../../include/fv2d_params.h"
#include "IniFile.h"

namespace fv2d
{

using real_t          = double; // SimInfo.h:13
constexpr int Nfields = FV2D_NFIELDS;

// Same enumerator names and values as the reference (SimInfo.h:26-97).
enum IDir : uint8_t { IX = FV2D_IX, IY = FV2D_IY };
enum IVar : uint8_t { IR = FV2D_IR, IU = FV2D_IU, IV = FV2D_IV, IP = FV2D_IP, IE = FV2D_IE };
enum RiemannSolver { HLL = FV2D_HLL, HLLC = FV2D_HLLC, FSLP = FV2D_FSLP };
enum BoundaryType { BC_ABSORBING = FV2D_BC_ABSORBING, BC_REFLECTING = FV2D_BC_REFLECTING, BC_PERIODIC = FV2D_BC_PERIODIC };
enum TimeStepping { TS_EULER = FV2D_TS_EULER, TS_RK2 = FV2D_TS_RK2 };
enum ReconstructionType { PCM = FV2D_PCM, PCM_WB = FV2D_PCM_WB, PLM = FV2D_PLM };
enum ThermalConductivityMode { TCM_CONSTANT = FV2D_TCM_CONSTANT, TCM_B02 = FV2D_TCM_B02 };
enum BCTC_Mode { BCTC_NONE = FV2D_BCTC_NONE, BCTC_FIXED_TEMPERATURE = FV2D_BCTC_FIXED_TEMPERATURE, BCTC_FIXED_GRADIENT = FV2D_BCTC_FIXED_GRADIENT };
enum ViscosityMode { VSC_CONSTANT = FV2D_VSC_CONSTANT };
enum GravityMode { GRAV_NONE = FV2D_GRAV_NONE, GRAV_CONSTANT = FV2D_GRAV_CONSTANT, GRAV_ANALYTICAL = FV2D_GRAV_ANALYTICAL };
enum AnalyticalGravityMode { AGM_HOT_BUBBLE = FV2D_AGM_HOT_BUBBLE };

struct RestartInfo // SimInfo.h:20-24
{
  real_t time;
  int iteration;
};

// Optional "section.key=value" overrides applied on top of the file (used by the tests and
// the benchmark to scale Nx/Ny of a shipped configuration without editing it).  An
// override behaves exactly as if the line had been present in the file.
using IniOverrides = std::map<std::string, std::string>;

// Reader (SimInfo.h:123-263): wraps the ini file, records each queried key.
struct Reader
{
  struct value_container
  {
    std::string value;
    bool from_file        = false;
    bool is_default_value = true;
  };

  Reader() = default;
  explicit Reader(const std::string &filename, const IniOverrides &ov = {}) : reader(filename), overrides(ov) {}

  std::map<std::string, std::map<std::string, value_container>> _values;
  IniFile reader;
  IniOverrides overrides; // lower-cased "section=name" -> value

  static std::string lower(std::string s)
  {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return char(std::tolower(c)); });
    return s;
  }

  bool overridden(const std::string &section, const std::string &name, std::string &out) const
  {
    auto it = overrides.find(IniFile::MakeKey(section, name));
    if (it == overrides.end())
      return false;
    out = it->second;
    return true;
  }

  template <typename T>
  void registerValue(std::string section, std::string name, const T &value, bool is_default_value)
  {
    section = lower(section);
    name    = lower(name);
    if (_values.count(section) != 0 && _values.at(section).count(name) != 0)
      throw std::runtime_error(std::string("parameter already set : ") + name); // SimInfo.h:151-154

    std::string dummy;
    const bool in_file = (reader.HasSection(section) && reader.HasValue(section, name)) || overridden(section, name, dummy);
    value_container &slot = _values[section][name];
    if (in_file)
    {
      slot.from_file        = true;
      slot.is_default_value = is_default_value;
    }
    if constexpr (std::is_same_v<T, std::string>)
      slot.value = value;
    else if constexpr (std::is_same_v<T, bool>)
      slot.value = value ? "true" : "false";
    else if constexpr (std::is_floating_point_v<T>)
    {
      std::ostringstream os;
      os << std::scientific << std::setprecision(12) << value;
      slot.value = os.str();
    }
    else
      slot.value = std::to_string(value);
  }

  bool GetBoolean(std::string section, std::string name, bool default_value)
  {
    bool res = reader.GetBoolean(section, name, default_value);
    std::string ov;
    if (overridden(section, name, ov))
    {
      std::string s = lower(ov);
      if (s == "true" || s == "yes" || s == "on" || s == "1")
        res = true;
      else if (s == "false" || s == "no" || s == "off" || s == "0")
        res = false;
    }
    registerValue(section, name, res, res == default_value);
    return res;
  }

  int GetInteger(std::string section, std::string name, int default_value)
  {
    int res = int(reader.GetInteger(section, name, default_value));
    std::string ov;
    if (overridden(section, name, ov))
    {
      char *end = nullptr;
      long n    = std::strtol(ov.c_str(), &end, 0);
      if (end > ov.c_str())
        res = int(n);
    }
    registerValue(section, name, res, res == default_value);
    return res;
  }

  // Q1: the value (and the default) go through *float* before being widened (SimInfo.h:195-200).
  real_t GetFloat(std::string section, std::string name, real_t default_value)
  {
    real_t res = reader.GetFloat(section, name, float(default_value));
    std::string ov;
    if (overridden(section, name, ov))
    {
      char *end = nullptr;
      float x   = std::strtof(ov.c_str(), &end);
      if (end > ov.c_str())
        res = x;
    }
    registerValue(section, name, res, res == default_value);
    return res;
  }

  std::string Get(std::string section, std::string name, std::string default_value)
  {
    std::string res = reader.Get(section, name, default_value);
    std::string ov;
    if (overridden(section, name, ov))
      res = ov;
    registerValue(section, name, res, res == default_value);
    return res;
  }

  template <typename T>
  T GetMapValue(const std::map<std::string, T> &map, const std::string &section, const std::string &name,
                const std::string &default_value)
  {
    std::string tmp = Get(section, name, default_value);
    if (map.count(tmp) == 0)
    {
      tmp = "\nallowed values: ";
      for (const auto &elem : map)
        tmp += elem.first + ", ";
      throw std::runtime_error(std::string("bad parameter for ") + name + ": " + tmp); // SimInfo.h:214-220
    }
    return map.at(tmp);
  }

  // Effective configuration dump, same layout as SimInfo.h:224-262.
  void outputValues(std::ostream &o)
  {
    constexpr std::string::size_type name_width  = 26;
    constexpr std::string::size_type value_width = 20;
    auto initial_format                          = o.flags();
    std::string problem                          = _values["physics"]["problem"].value;
    o << "; Parameters used for the problem: " << problem << std::endl;
    o << std::left;
    for (const auto &p_section : _values)
    {
      if (!reader.HasSection(p_section.first))
        continue;
      o << "\n[" << p_section.first << "]" << std::endl;
      for (const auto &p_var : p_section.second)
      {
        const std::string &var_name = p_var.first;
        const value_container &val  = p_var.second;
        o << std::setw(int(std::max(var_name.length(), name_width))) << var_name << " = "
          << std::setw(int(std::max(val.value.length(), value_width))) << val.value
          << (val.from_file ? "" : " ; default ") << std::endl;
      }
    }
    o.flags(initial_format);
  }
};

// DeviceParams (SimInfo.h:266-460): the POD that is handed to the kernels, plus its reader.
struct DeviceParams : fv2d_device_params
{
  DeviceParams()
  {
    fv2d_device_params zero = {};
    static_cast<fv2d_device_params &>(*this) = zero;
    gamma0         = 5.0 / 3.0;
    boundary_x     = BC_REFLECTING;
    boundary_y     = BC_REFLECTING;
    reconstruction = PCM;
    riemann_solver = HLL;
    CFL            = 0.1;
    epsilon        = 1.0e-6;
  }

  void init_from_inifile(Reader &reader)
  {
    // Mesh (SimInfo.h:360-376)
    Nx   = reader.GetInteger("mesh", "Nx", 32);
    Ny   = reader.GetInteger("mesh", "Ny", 32);
    Ng   = reader.GetInteger("mesh", "Nghosts", 2);
    xmin = reader.GetFloat("mesh", "xmin", 0.0);
    xmax = reader.GetFloat("mesh", "xmax", 1.0);
    ymin = reader.GetFloat("mesh", "ymin", 0.0);
    ymax = reader.GetFloat("mesh", "ymax", 1.0);

    Ntx  = Nx + 2 * Ng;
    Nty  = Ny + 2 * Ng;
    ibeg = Ng;
    iend = Ng + Nx;
    jbeg = Ng;
    jend = Ng + Ny;

    dx = (xmax - xmin) / Nx;
    dy = (ymax - ymin) / Ny;

    // Solvers / run (SimInfo.h:378-388)
    CFL = reader.GetFloat("solvers", "CFL", 0.8);
    const std::map<std::string, int> bc_map{{"reflecting", BC_REFLECTING}, {"absorbing", BC_ABSORBING}, {"periodic", BC_PERIODIC}};
    boundary_x = reader.GetMapValue(bc_map, "run", "boundaries_x", "reflecting");
    boundary_y = reader.GetMapValue(bc_map, "run", "boundaries_y", "reflecting");
    const std::map<std::string, int> recons_map{{"pcm", PCM}, {"pcm_wb", PCM_WB}, {"plm", PLM}};
    reconstruction = reader.GetMapValue(recons_map, "solvers", "reconstruction", "pcm");
    const std::map<std::string, int> riemann_map{{"hll", HLL}, {"hllc", HLLC}, {"fslp", FSLP}};
    riemann_solver = reader.GetMapValue(riemann_map, "solvers", "riemann_solver", "hllc");

    // Physics (SimInfo.h:391-398)
    epsilon                    = reader.GetFloat("misc", "epsilon", 1.0e-6);
    gamma0                     = reader.GetFloat("physics", "gamma0", 5.0 / 3.0);
    m1                         = reader.GetFloat("polytrope", "m1", 1.0);
    theta1                     = reader.GetFloat("polytrope", "theta1", 10.0);
    m2                         = reader.GetFloat("polytrope", "m2", 1.0);
    theta2                     = reader.GetFloat("polytrope", "theta2", 10.0);
    well_balanced_flux_at_y_bc = reader.GetBoolean("physics", "well_balanced_flux_at_y_bc", false);
    fslp_K                     = reader.GetFloat("physics", "fslp_K", 1.1);

    // Gravity (SimInfo.h:401-410)
    const std::map<std::string, int> gravity_map{{"none", GRAV_NONE}, {"constant", GRAV_CONSTANT}, {"analytical", GRAV_ANALYTICAL}};
    gravity_mode = reader.GetMapValue(gravity_map, "gravity", "mode", "none");
    gx           = reader.GetFloat("gravity", "gx", 0.0);
    gy           = reader.GetFloat("gravity", "gy", 0.0);
    const std::map<std::string, int> analytical_gravity_map{{"hot_bubble", AGM_HOT_BUBBLE}};
    analytical_gravity_mode = reader.GetMapValue(analytical_gravity_map, "gravity", "analytical_mode", "hot_bubble");

    // Thermal conduction (SimInfo.h:413-427).  The boundary keys are bc_ymin/bc_ymax (Q2):
    // shipped files say bc_xmin/... and therefore get BCTC_NONE.
    thermal_conductivity_active = reader.GetBoolean("thermal_conduction", "active", false);
    const std::map<std::string, int> tc_map{{"constant", TCM_CONSTANT}, {"B02", TCM_B02}};
    thermal_conductivity_mode = reader.GetMapValue(tc_map, "thermal_conduction", "conductivity_mode", "constant");
    kappa                     = reader.GetFloat("thermal_conduction", "kappa", 0.0);
    const std::map<std::string, int> bctc_map{{"none", BCTC_NONE}, {"fixed_temperature", BCTC_FIXED_TEMPERATURE}, {"fixed_gradient", BCTC_FIXED_GRADIENT}};
    bctc_ymin       = reader.GetMapValue(bctc_map, "thermal_conduction", "bc_ymin", "none");
    bctc_ymax       = reader.GetMapValue(bctc_map, "thermal_conduction", "bc_ymax", "none");
    bctc_ymin_value = reader.GetFloat("thermal_conduction", "bc_ymin_value", 1.0);
    bctc_ymax_value = reader.GetFloat("thermal_conduction", "bc_ymax_value", 1.0);

    // Viscosity (SimInfo.h:430-435)
    viscosity_active = reader.GetBoolean("viscosity", "active", false);
    const std::map<std::string, int> viscosity_map{{"constant", VSC_CONSTANT}};
    viscosity_mode = reader.GetMapValue(viscosity_map, "viscosity", "viscosity_mode", "constant");
    mu             = reader.GetFloat("viscosity", "mu", 0.0);

    // Problem parameters (SimInfo.h:438-458)
    h84_pert      = reader.GetFloat("H84", "perturbation", 1.0e-4);
    c91_pert      = reader.GetFloat("C91", "perturbation", 1.0e-3);
    hot_bubble_g0 = reader.GetFloat("hot_bubble", "g0", 0.0);

    kh_a       = reader.GetFloat("kelvin_helmholtz", "a", 0.05);
    kh_amp     = reader.GetFloat("kelvin_helmholtz", "amp", 0.01);
    kh_P0      = reader.GetFloat("kelvin_helmholtz", "P0", 1.0);
    kh_rho_fac = reader.GetFloat("kelvin_helmholtz", "rho_fac", 0.0);
    kh_sigma   = reader.GetFloat("kelvin_helmholtz", "sigma", 0.2);
    // Q2: these three are looked up in the misspelt section "kelvin_helmholts".
    kh_uflow = reader.GetFloat("kelvin_helmholts", "uflow", 1.0);
    kh_y1    = reader.GetFloat("kelvin_helmholts", "y1", 0.5);
    kh_y2    = reader.GetFloat("kelvin_helmholts", "y2", 1.5);

    gresho_density = reader.GetFloat("gresho_vortex", "density", 1.0);
    gresho_Mach    = reader.GetFloat("gresho_vortex", "Mach", 0.1);
  }
};

// Index range [lower, upper) in (i, j), standing in for the reference's MDRangePolicy members.
struct ParallelRange
{
  int lower[2] = {0, 0};
  int upper[2] = {0, 0};
};

// Params (SimInfo.h:463-492)
struct Params
{
  real_t save_freq = 0.1;
  real_t tend      = 1.0;
  Reader reader;
  std::string filename_out   = "run";
  std::string output_path    = "./";
  std::string restart_file   = "";
  TimeStepping time_stepping = TS_EULER;
  bool multiple_outputs      = false;

  ParallelRange range_tot, range_dom, range_xbound, range_ybound, range_slopes;

  std::string problem;
  DeviceParams device_params;

  int seed                      = 12345;
  int log_frequency             = 10;
  real_t epsilon_reset_negative = 1.0e-8;

  fv2d_run_params run_pod() const
  {
    fv2d_run_params r = {};
    r.save_freq              = save_freq;
    r.tend                   = tend;
    r.epsilon_reset_negative = epsilon_reset_negative;
    r.time_stepping          = time_stepping;
    r.multiple_outputs       = multiple_outputs;
    r.seed                   = seed;
    r.log_frequency          = log_frequency;
    std::snprintf(r.problem, sizeof r.problem, "%s", problem.c_str());
    std::snprintf(r.filename_out, sizeof r.filename_out, "%s", filename_out.c_str());
    std::snprintf(r.output_path, sizeof r.output_path, "%s", output_path.c_str());
    std::snprintf(r.restart_file, sizeof r.restart_file, "%s", restart_file.c_str());
    return r;
  }
};

// Cell-centre position (SimInfo.h:494-499)
inline void getPos(const fv2d_device_params &params, int i, int j, real_t pos[2])
{
  pos[IX] = params.xmin + (i - params.ibeg + 0.5) * params.dx;
  pos[IY] = params.ymin + (j - params.jbeg + 0.5) * params.dy;
}

// SimInfo.h:501-527.  Raw-case section names are compared with lower-cased registrations,
// so [C91]/[H84] are reported unknown although their values are read (Q2).
inline void checkValidityIni(Params &params, std::ostream &err = std::cerr)
{
  const auto &ini_sections  = params.reader.reader.Sections();
  const auto &ini_keyvalues = params.reader.reader.Values();
  auto &valid_keyvalues     = params.reader._values;
  for (const auto &s : ini_sections)
  {
    if (valid_keyvalues.count(s) == 0)
    {
      err << "WARNING: section [" << s << "] is unknown." << std::endl;
      continue;
    }
    const std::string prefix = s + "=";
    for (const auto &kv : ini_keyvalues)
      if (kv.first.compare(0, prefix.size(), prefix) == 0)
      {
        const std::string key = kv.first.substr(prefix.size());
        if (valid_keyvalues[s].count(key) == 0)
          err << "WARNING: parameter `" << key << "` in section [" << s << "] is unknown." << std::endl;
      }
  }
}

// SimInfo.h:529-568
inline Params readInifile(std::string filename, const IniOverrides &overrides = {}, std::ostream &err = std::cerr)
{
  Params res;
  res.reader   = Reader(filename, overrides);
  auto &reader = res.reader;

  res.tend             = reader.GetFloat("run", "tend", 1.0);
  res.multiple_outputs = reader.GetBoolean("run", "multiple_outputs", false);
  res.restart_file     = reader.Get("run", "restart_file", "");
  res.save_freq        = reader.GetFloat("run", "save_freq", 1.0e-1);
  res.filename_out     = reader.Get("run", "output_filename", "run");
  res.output_path      = reader.Get("run", "output_path", "./");

  const std::map<std::string, TimeStepping> ts_map{{"euler", TS_EULER}, {"RK2", TS_RK2}};
  res.time_stepping = reader.GetMapValue(ts_map, "solvers", "time_stepping", "euler");
  res.problem       = reader.Get("physics", "problem", "blast");

  res.seed                   = reader.GetInteger("misc", "seed", 12345);
  res.log_frequency          = reader.GetInteger("misc", "log_frequency", 10);
  res.epsilon_reset_negative = reader.GetFloat("misc", "epsilon_reset_negative", 1.0e-8);

  res.device_params.init_from_inifile(res.reader);

  const auto &d    = res.device_params;
  res.range_tot    = ParallelRange{{0, 0}, {d.Ntx, d.Nty}};
  res.range_dom    = ParallelRange{{d.ibeg, d.jbeg}, {d.iend, d.jend}};
  res.range_xbound = ParallelRange{{0, d.jbeg}, {d.Ng, d.jend}};
  res.range_ybound = ParallelRange{{0, 0}, {d.Ntx, d.Ng}};
  res.range_slopes = ParallelRange{{d.ibeg - 1, d.jbeg - 1}, {d.iend + 1, d.jend + 1}};

  checkValidityIni(res, err);
  return res;
}

} // namespace fv2d
