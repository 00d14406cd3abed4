/* fv2d_params.h — plain-C POD mirrors of the reference's parameter structs.
 *
 * fv2d_device_params replaces `struct DeviceParams` (reference SimInfo.h:266-354):
 * same members, same meaning, same order; enums become int32_t, bools become int32_t,
 * and the C++ member function init_from_inifile (SimInfo.h:356-459) lives in the host
 * library instead (fv2d_params_from_ini, include/fv2d_b200.h).
 *
 * fv2d_run_params carries the host-only members of `struct Params`
 * (reference SimInfo.h:463-492) that the time loop needs.
 *
 * This header contains no computation; it is shared by the product (fv2d_b200/),
 * the C-ABI (include/fv2d_b200.h) and the test oracle (oracle/) so that all three
 * agree on the layout.
 */
#ifndef FV2D_PARAMS_H_
#define FV2D_PARAMS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV2D_NFIELDS 4 /* SimInfo.h:14 */

/* SimInfo.h:26-39 */
enum { FV2D_IX = 0, FV2D_IY = 1 };
enum { FV2D_IR = 0, FV2D_IU = 1, FV2D_IV = 2, FV2D_IP = 3, FV2D_IE = 3 };
/* SimInfo.h:41-46 */
enum { FV2D_HLL = 0, FV2D_HLLC = 1, FV2D_FSLP = 2 };
/* SimInfo.h:48-53 */
enum { FV2D_BC_ABSORBING = 0, FV2D_BC_REFLECTING = 1, FV2D_BC_PERIODIC = 2 };
/* SimInfo.h:55-59 */
enum { FV2D_TS_EULER = 0, FV2D_TS_RK2 = 1 };
/* SimInfo.h:61-66 */
enum { FV2D_PCM = 0, FV2D_PCM_WB = 1, FV2D_PLM = 2 };
/* SimInfo.h:68-72 */
enum { FV2D_TCM_CONSTANT = 0, FV2D_TCM_B02 = 1 };
/* SimInfo.h:75-80 */
enum { FV2D_BCTC_NONE = 0, FV2D_BCTC_FIXED_TEMPERATURE = 1, FV2D_BCTC_FIXED_GRADIENT = 2 };
/* SimInfo.h:82-85 */
enum { FV2D_VSC_CONSTANT = 0 };
/* SimInfo.h:87-92 */
enum { FV2D_GRAV_NONE = 0, FV2D_GRAV_CONSTANT = 1, FV2D_GRAV_ANALYTICAL = 2 };
/* SimInfo.h:94-97 */
enum { FV2D_AGM_HOT_BUBBLE = 0 };

/* Mirrors DeviceParams, SimInfo.h:266-354.  Every real-valued member holds a value
 * that went through float (strtof) exactly like the reference's reader (SimInfo.h:195-200
 * -> external/inih/INIReader.h:420-427), then widened to double. */
typedef struct fv2d_device_params
{
  /* Thermodynamics */
  double gamma0;
  /* Gravity */
  int32_t gravity_mode;
  int32_t analytical_gravity_mode;
  double gx, gy;
  int32_t well_balanced_flux_at_y_bc;
  int32_t well_balanced; /* declared by the reference, never read from the .ini */
  /* FSLP */
  double fslp_K;
  /* Thermal conductivity */
  int32_t thermal_conductivity_active;
  int32_t thermal_conductivity_mode;
  double kappa;
  int32_t bctc_ymin, bctc_ymax;
  double bctc_ymin_value, bctc_ymax_value;
  /* Viscosity */
  int32_t viscosity_active;
  int32_t viscosity_mode;
  double mu;
  /* Polytropes */
  double m1, theta1, m2, theta2;
  double h84_pert;
  double c91_pert;
  /* B02: declared by the reference (SimInfo.h:307-310) but never read from the .ini,
   * i.e. uninitialised there; zero here and TCM_B02 is rejected. */
  double b02_ymid, b02_kappa1, b02_kappa2, b02_thickness;
  double hot_bubble_g0;
  /* Kelvin-Helmholtz */
  double kh_y1, kh_y2, kh_a, kh_sigma, kh_rho_fac, kh_uflow, kh_amp, kh_P0;
  /* Gresho vortex */
  double gresho_density, gresho_Mach;
  /* Boundaries / Godunov */
  int32_t boundary_x, boundary_y;
  int32_t reconstruction, riemann_solver;
  double CFL;
  /* Mesh */
  int32_t Nx, Ny, Ng, Ntx, Nty, ibeg, iend, jbeg, jend;
  int32_t pad0_;
  double xmin, xmax, ymin, ymax, dx, dy;
  /* Misc */
  double epsilon;
} fv2d_device_params;

/* Host-only members of Params (SimInfo.h:463-492). */
typedef struct fv2d_run_params
{
  double save_freq;
  double tend;
  double epsilon_reset_negative;
  int32_t time_stepping;
  int32_t multiple_outputs;
  int32_t seed;
  int32_t log_frequency;
  char problem[64];
  char filename_out[256];
  char output_path[256];
  char restart_file[256];
} fv2d_run_params;

#ifdef __cplusplus
}
#endif
#endif /* FV2D_PARAMS_H_ */
