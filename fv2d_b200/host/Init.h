// Init — problem setups, evaluated on the host in fp64 with glibc libm so that the initial
// state is bit-identical to the reference's Kokkos-OpenMP build.
//
// Mirrors the reference's InitFunctor (reference Init.h:272-359): same nine problem names,
// same formulas (Init.h:21-265), same "fill domain, then fillBoundaries" order (Init.h:319-357).
// Runs once per job; not a performance target.
#pragma once

#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "SimInfo.h"

namespace fv2d
{

// Host-side staging array, A(j, i, f) like the reference's View (main.cpp:33-34), stored as
// SoA planes [f][Nty][Ntx] — the layout fv2d_upload_Q / fv2d_download_Q speak.
// Zero-initialised: the reference relies on Kokkos zero-filling its Views (Q9).
struct HostArray
{
  int Ntx = 0, Nty = 0;
  std::vector<real_t> data;
  HostArray() = default;
  HostArray(int nty, int ntx) : Ntx(ntx), Nty(nty), data(size_t(Nfields) * nty * ntx, 0.0) {}
  real_t &operator()(int j, int i, int f) { return data[(size_t(f) * Nty + j) * Ntx + i]; }
  const real_t &operator()(int j, int i, int f) const { return data[(size_t(f) * Nty + j) * Ntx + i]; }
};

// Single-stream equivalent of Kokkos::Random_XorShift64_Pool as the reference uses it on a
// one-thread host backend (external/kokkos/algorithms/src/Kokkos_Random.hpp:745-766 generator,
// :840-850 drand, :907-942 pool seeding, :944-960 get_state/free_state): with one hardware
// thread every draw comes from pool state 0, in iteration order.  With more threads the
// reference's initial condition depends on the thread count (Q11), so this is the
// "OMP_NUM_THREADS=1" variant; parity tests for C91/H84 otherwise start from a dumped Q0.
class XorShift64Pool1
{
  uint64_t state_;
  static uint64_t step(uint64_t &s)
  {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s;
  }
  static uint32_t urand(uint64_t &s)
  {
    uint64_t tmp = step(s) * 2685821657736338717ULL;
    return uint32_t((tmp >> 16) & 0xffffffffULL);
  }
  static int rand31(uint64_t &s) { return int(urand(s) / 2); }

public:
  explicit XorShift64Pool1(uint64_t seed)
  {
    if (seed == 0)
      seed = 1318319ULL;
    uint64_t g = seed;
    for (int i = 0; i < 17; ++i)
      rand31(g);
    uint64_t n1 = uint64_t(rand31(g)), n2 = uint64_t(rand31(g)), n3 = uint64_t(rand31(g)), n4 = uint64_t(rand31(g));
    state_ = ((n1 & 0xffff) << 0) | ((n2 & 0xffff) << 16) | ((n3 & 0xffff) << 32) | ((n4 & 0xffff) << 48);
  }
  // generator.drand(start, end) on a generator checked out of, and returned to, the pool
  double drand(double start, double end)
  {
    uint64_t s = (state_ == 0 ? 1318319ULL : state_);
    uint64_t u = step(s) * 2685821657736338717ULL - 1;
    state_     = s;
    const double range = end - start;
    return range * double(u) / 18446744073709551615.0 + start;
  }
};

enum InitType { SOD_X, SOD_Y, BLAST, RAYLEIGH_TAYLOR, DIFFUSION, H84, C91, KELVIN_HELMHOLTZ, GRESHO_VORTEX }; // Init.h:271-282

// Host restatement of BoundaryManager::fillBoundaries (reference BoundaryConditions.h:82-147)
// used only at initialisation / snapshot load; the per-step fill is the CUDA ghost-fill kernel.
inline void fillBoundariesHost(const fv2d_device_params &p, HostArray &Q)
{
  auto copy = [&](int id, int jd, int is, int js, int flip) {
    for (int f = 0; f < Nfields; ++f)
      Q(jd, id, f) = (f == flip ? Q(js, is, f) * -1.0 : Q(js, is, f));
  };
  auto src = [&](int bc, int k, int beg, int end, int N) { // ghost index k -> source index
    switch (bc)
    {
    case BC_REFLECTING: return 2 * (k < beg ? beg : end) - k - 1;
    case BC_PERIODIC: return k < beg ? k + N : k - N;
    default: return k < beg ? beg : end - 1; // absorbing
    }
  };
  for (int j = p.jbeg; j < p.jend; ++j)
    for (int i = 0; i < p.Ng; ++i)
      for (int ig : {i, p.iend + i})
        copy(ig, j, src(p.boundary_x, ig, p.ibeg, p.iend, p.Nx), j, p.boundary_x == BC_REFLECTING ? IU : -1);
  for (int j = 0; j < p.Ng; ++j)
    for (int i = 0; i < p.Ntx; ++i)
      for (int jg : {j, p.jend + j})
        copy(i, jg, i, src(p.boundary_y, jg, p.jbeg, p.jend, p.Ny), p.boundary_y == BC_REFLECTING ? IV : -1);
}

struct InitFunctor
{
private:
  Params full_params;
  InitType init_type;

public:
  explicit InitFunctor(Params &params) : full_params(params)
  {
    const std::map<std::string, InitType> init_map{{"sod_x", SOD_X},
                                                   {"sod_y", SOD_Y},
                                                   {"blast", BLAST},
                                                   {"rayleigh-taylor", RAYLEIGH_TAYLOR},
                                                   {"diffusion", DIFFUSION},
                                                   {"H84", H84},
                                                   {"C91", C91},
                                                   {"kelvin_helmholtz", KELVIN_HELMHOLTZ},
                                                   {"gresho_vortex", GRESHO_VORTEX}};
    if (init_map.count(full_params.problem) == 0)
      throw std::runtime_error("Error unknown problem " + full_params.problem); // Init.h:303-304
    init_type = init_map.at(full_params.problem);
  }

  // Fills the active domain, then the ghosts (Init.h:310-358).  Q must be zero-initialised.
  void init(HostArray &Q)
  {
    const fv2d_device_params &p = full_params.device_params;
    XorShift64Pool1 pool{uint64_t(full_params.seed)};

    // The reference's one-thread host iteration order is i outer, j inner
    // (KokkosExp_MDRangePolicy.hpp:138-146, 328-345); it only matters for the RNG draws of
    // H84 / C91, which therefore run in that order on one thread.  The other problems are
    // pure functions of (i, j) and are filled row-parallel.
    const bool sequential = (init_type == H84 || init_type == C91);
    const long long ncell = (long long)p.Nx * p.Ny;
#pragma omp parallel for schedule(static) if (!sequential)
    for (long long cell = 0; cell < ncell; ++cell)
      {
        int i, j;
        if (sequential)
        {
          i = p.ibeg + int(cell / p.Ny);
          j = p.jbeg + int(cell % p.Ny);
        }
        else
        {
          j = p.jbeg + int(cell / p.Nx);
          i = p.ibeg + int(cell % p.Nx);
        }
        real_t pos[2];
        getPos(p, i, j, pos);
        const real_t x = pos[IX], y = pos[IY];
        switch (init_type)
        {
        case SOD_X: // Init.h:21-35 (IV is never written: Q9)
        case SOD_Y: // Init.h:84-98
        {
          const bool left = (init_type == SOD_X ? x : y) <= 0.5;
          Q(j, i, IR)     = left ? 1.0 : 0.125;
          Q(j, i, IP)     = left ? 1.0 : 0.1;
          Q(j, i, IU)     = 0.0;
          break;
        }
        case BLAST: // Init.h:104-131
        {
          const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
          const real_t xr = xmid - x, yr = ymid - y;
          const real_t r  = std::sqrt(xr * xr + yr * yr);
          Q(j, i, IR)     = r < 0.2 ? 1.0 : 1.2;
          Q(j, i, IU)     = 0.0;
          Q(j, i, IV)     = 0.0;
          Q(j, i, IP)     = r < 0.2 ? 10.0 : 0.1;
          break;
        }
        case DIFFUSION: // Init.h:185-207
        {
          const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
          const real_t x0 = x - xmid, y0 = y - ymid;
          const real_t r  = std::sqrt(x0 * x0 + y0 * y0);
          Q(j, i, IR)     = r < 0.2 ? 1.0 : 0.1;
          Q(j, i, IP)     = 1.0;
          Q(j, i, IU)     = 1.0;
          Q(j, i, IV)     = 1.0;
          break;
        }
        case RAYLEIGH_TAYLOR: // Init.h:212-237 (IV only written for |y| < 1/3: Q9)
        {
          const real_t ymid = 0.5 * (p.ymin + p.ymax);
          const real_t P0   = 2.5;
          Q(j, i, IR)       = y < ymid ? 1.0 : 2.0;
          Q(j, i, IU)       = 0.0;
          Q(j, i, IP)       = P0 + 0.1 * p.gy * y;
          if (y > -1.0 / 3.0 && y < 1.0 / 3.0)
            Q(j, i, IV) = 0.01 * (1.0 + std::cos(4 * M_PI * x)) * (1 + std::cos(3.0 * M_PI * y)) / 4.0;
          break;
        }
        case H84: // Init.h:136-155
        {
          const real_t rho  = std::pow(y, p.m1);
          const real_t prs  = std::pow(y, p.m1 + 1.0);
          const real_t pert = p.h84_pert * pool.drand(-0.5, 0.5);
          Q(j, i, IR)       = rho;
          Q(j, i, IU)       = 0.0;
          Q(j, i, IV)       = pert;
          Q(j, i, IP)       = prs;
          break;
        }
        case C91: // Init.h:160-181
        {
          const real_t T    = (1.0 + p.theta1 * y);
          const real_t rho  = std::pow(T, p.m1);
          real_t prs        = std::pow(T, p.m1 + 1.0);
          const real_t pert = p.c91_pert * pool.drand(-0.5, 0.5);
          prs               = prs * (1.0 + pert);
          Q(j, i, IR)       = rho;
          Q(j, i, IU)       = 0.0;
          Q(j, i, IV)       = 0.0;
          Q(j, i, IP)       = prs;
          break;
        }
        case KELVIN_HELMHOLTZ: // Init.h:246-265
        {
          const real_t q1  = std::tanh((y - p.kh_y1) / p.kh_a);
          const real_t q2  = std::tanh((y - p.kh_y2) / p.kh_a);
          const real_t s2  = p.kh_sigma * p.kh_sigma;
          const real_t dy1 = (y - p.kh_y1) * (y - p.kh_y1);
          const real_t dy2 = (y - p.kh_y2) * (y - p.kh_y2);
          const real_t rho = 1.0 + p.kh_rho_fac * 0.5 * (q1 - q2);
          const real_t u   = p.kh_uflow * (q1 - q2 - 1.0);
          const real_t v   = p.kh_amp * std::sin(2.0 * M_PI * x) * (std::exp(-dy1 / s2) + std::exp(-dy2 / s2));
          Q(j, i, IR)      = rho;
          Q(j, i, IU)      = u;
          Q(j, i, IV)      = v;
          Q(j, i, IP)      = p.kh_P0;
          break;
        }
        case GRESHO_VORTEX: // Init.h:43-78
        {
          const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
          const real_t xr = x - xmid, yr = y - ymid;
          const real_t r  = std::sqrt(xr * xr + yr * yr);
          const real_t p0 = p.gresho_density / (p.gamma0 * p.gresho_Mach * p.gresho_Mach);
          Q(j, i, IR)     = p.gresho_density;
          real_t u_phi;
          if (r < 0.2)
          {
            u_phi       = 5.0 * r;
            Q(j, i, IP) = p0 + 12.5 * r * r;
          }
          else if (r < 0.4)
          {
            u_phi       = 2.0 - 5.0 * r;
            Q(j, i, IP) = p0 + 12.5 * r * r + 4.0 * (1.0 - 5.0 * r + std::log(5.0 * r));
          }
          else
          {
            u_phi       = 0.0;
            Q(j, i, IP) = p0 - 2.0 + 4.0 * std::log(2.0);
          }
          const real_t xnr = xr / r, ynr = yr / r;
          Q(j, i, IU)      = -ynr * u_phi;
          Q(j, i, IV)      = xnr * u_phi;
          break;
        }
        }
      }

    fillBoundariesHost(p, Q);
  }
};

} // namespace fv2d
