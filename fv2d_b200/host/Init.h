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

  // One cell of the active domain: writes into q[] exactly the fields the reference's init
  // function writes (q comes in zero-filled: Q9).
  void cell(int i, int j, real_t q[4], XorShift64Pool1 &pool) const
  {
    const fv2d_device_params &p = full_params.device_params;
    real_t pos[2];
    getPos(p, i, j, pos);
    const real_t x = pos[IX], y = pos[IY];
    switch (init_type)
    {
    case SOD_X: // Init.h:21-35 (IV is never written: Q9)
    case SOD_Y: // Init.h:84-98
    {
      const bool left = (init_type == SOD_X ? x : y) <= 0.5;
      q[IR]     = left ? 1.0 : 0.125;
      q[IP]     = left ? 1.0 : 0.1;
      q[IU]     = 0.0;
      break;
    }
    case BLAST: // Init.h:104-131
    {
      const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
      const real_t xr = xmid - x, yr = ymid - y;
      const real_t r  = std::sqrt(xr * xr + yr * yr);
      q[IR]     = r < 0.2 ? 1.0 : 1.2;
      q[IU]     = 0.0;
      q[IV]     = 0.0;
      q[IP]     = r < 0.2 ? 10.0 : 0.1;
      break;
    }
    case DIFFUSION: // Init.h:185-207
    {
      const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
      const real_t x0 = x - xmid, y0 = y - ymid;
      const real_t r  = std::sqrt(x0 * x0 + y0 * y0);
      q[IR]     = r < 0.2 ? 1.0 : 0.1;
      q[IP]     = 1.0;
      q[IU]     = 1.0;
      q[IV]     = 1.0;
      break;
    }
    case RAYLEIGH_TAYLOR: // Init.h:212-237 (IV only written for |y| < 1/3: Q9)
    {
      const real_t ymid = 0.5 * (p.ymin + p.ymax);
      const real_t P0   = 2.5;
      q[IR]       = y < ymid ? 1.0 : 2.0;
      q[IU]       = 0.0;
      q[IP]       = P0 + 0.1 * p.gy * y;
      if (y > -1.0 / 3.0 && y < 1.0 / 3.0)
        q[IV] = 0.01 * (1.0 + std::cos(4 * M_PI * x)) * (1 + std::cos(3.0 * M_PI * y)) / 4.0;
      break;
    }
    case H84: // Init.h:136-155
    {
      const real_t rho  = std::pow(y, p.m1);
      const real_t prs  = std::pow(y, p.m1 + 1.0);
      const real_t pert = p.h84_pert * pool.drand(-0.5, 0.5);
      q[IR]       = rho;
      q[IU]       = 0.0;
      q[IV]       = pert;
      q[IP]       = prs;
      break;
    }
    case C91: // Init.h:160-181
    {
      const real_t T    = (1.0 + p.theta1 * y);
      const real_t rho  = std::pow(T, p.m1);
      real_t prs        = std::pow(T, p.m1 + 1.0);
      const real_t pert = p.c91_pert * pool.drand(-0.5, 0.5);
      prs               = prs * (1.0 + pert);
      q[IR]       = rho;
      q[IU]       = 0.0;
      q[IV]       = 0.0;
      q[IP]       = prs;
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
      q[IR]      = rho;
      q[IU]      = u;
      q[IV]      = v;
      q[IP]      = p.kh_P0;
      break;
    }
    case GRESHO_VORTEX: // Init.h:43-78
    {
      const real_t xmid = 0.5 * (p.xmin + p.xmax), ymid = 0.5 * (p.ymin + p.ymax);
      const real_t xr = x - xmid, yr = y - ymid;
      const real_t r  = std::sqrt(xr * xr + yr * yr);
      const real_t p0 = p.gresho_density / (p.gamma0 * p.gresho_Mach * p.gresho_Mach);
      q[IR]     = p.gresho_density;
      real_t u_phi;
      if (r < 0.2)
      {
        u_phi       = 5.0 * r;
        q[IP] = p0 + 12.5 * r * r;
      }
      else if (r < 0.4)
      {
        u_phi       = 2.0 - 5.0 * r;
        q[IP] = p0 + 12.5 * r * r + 4.0 * (1.0 - 5.0 * r + std::log(5.0 * r));
      }
      else
      {
        u_phi       = 0.0;
        q[IP] = p0 - 2.0 + 4.0 * std::log(2.0);
      }
      const real_t xnr = xr / r, ynr = yr / r;
      q[IU]      = -ynr * u_phi;
      q[IV]      = xnr * u_phi;
      break;
    }
    }
  }

  static bool isSequential(InitType t) { return t == H84 || t == C91; }

  // Fills the active domain, then the ghosts (Init.h:310-358).  Q must be zero-initialised.
  void init(HostArray &Q)
  {
    const fv2d_device_params &p = full_params.device_params;
    XorShift64Pool1 pool{uint64_t(full_params.seed)};

    // The reference's one-thread host iteration order is i outer, j inner
    // (KokkosExp_MDRangePolicy.hpp:138-146, 328-345); it only matters for the RNG draws of
    // H84 / C91, which therefore run in that order on one thread.  The other problems are
    // pure functions of (i, j) and are filled row-parallel.
    const bool sequential = isSequential(init_type);
    const long long ncell = (long long)p.Nx * p.Ny;
#pragma omp parallel for schedule(static) if (!sequential)
    for (long long cell_id = 0; cell_id < ncell; ++cell_id)
    {
      int i, j;
      if (sequential)
      {
        i = p.ibeg + int(cell_id / p.Ny);
        j = p.jbeg + int(cell_id % p.Ny);
      }
      else
      {
        j = p.jbeg + int(cell_id / p.Nx);
        i = p.ibeg + int(cell_id % p.Nx);
      }
      real_t q[4] = {Q(j, i, IR), Q(j, i, IU), Q(j, i, IV), Q(j, i, IP)};
      cell(i, j, q, pool);
      for (int f = 0; f < Nfields; ++f)
        Q(j, i, f) = q[f];
    }

    fillBoundariesHost(p, Q);
  }

  // Rows [j_first, j_first + Qrows.Nty) of what init() would produce, ghosts included, without
  // building the whole grid (each rank of a multi-GPU job only needs its own slab).  Every
  // ghost cell is a sign-flipped copy of one domain cell (BoundaryConditions.h:15-71), so it is
  // evaluated from that cell's formula directly.  The RNG-perturbed problems depend on the
  // global draw order and fall back to the full initialisation.
  void init_rows(HostArray &Qrows, int j_first)
  {
    const fv2d_device_params &p = full_params.device_params;
    if (isSequential(init_type))
    {
      HostArray full(p.Nty, p.Ntx);
      init(full);
      for (int f = 0; f < Nfields; ++f)
        for (int jl = 0; jl < Qrows.Nty; ++jl)
          for (int i = 0; i < p.Ntx; ++i)
            Qrows(jl, i, f) = full(j_first + jl, i, f);
      return;
    }
    auto src = [](int bc, int k, int beg, int end, int N) {
      if (k >= beg && k < end)
        return k;
      switch (bc)
      {
      case BC_REFLECTING: return 2 * (k < beg ? beg : end) - k - 1;
      case BC_PERIODIC: return k < beg ? k + N : k - N;
      default: return k < beg ? beg : end - 1;
      }
    };
    XorShift64Pool1 unused{uint64_t(full_params.seed)};
#pragma omp parallel for schedule(static)
    for (int jl = 0; jl < Qrows.Nty; ++jl)
    {
      const int j  = j_first + jl;
      const int js = src(p.boundary_y, j, p.jbeg, p.jend, p.Ny);
      for (int i = 0; i < p.Ntx; ++i)
      {
        const int is = src(p.boundary_x, i, p.ibeg, p.iend, p.Nx);
        real_t q[4]  = {0.0, 0.0, 0.0, 0.0};
        cell(is, js, q, unused);
        if (is != i && p.boundary_x == BC_REFLECTING)
          q[IU] *= -1.0;
        if (js != j && p.boundary_y == BC_REFLECTING)
          q[IV] *= -1.0;
        for (int f = 0; f < Nfields; ++f)
          Qrows(jl, i, f) = q[f];
      }
    }
  }
};

} // namespace fv2d
