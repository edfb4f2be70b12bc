// mx_tables.h -- device-resident images of the module tables.  Plain structs of
// scalars and DEVICE pointers, filled by the marxb200_set_* setters (marxb200.cu).  Small per-shell /
// per-chip blocks live in one contiguous "params" buffer that the kernels stage into shared memory;
// the big tables (WFOLD 1.3 MB, grating efficiencies 0.5 MB, FEF) stay in HBM and are served by L2.
#pragma once
#include <stdint.h>

namespace mx {

constexpr int kNumShells = 4;
constexpr int kMaxChips = 6;
constexpr int kMaxContamLayers = 5;
constexpr int kMaxGauss = 18;          // gaussians per FEF row: (MAX_FEF_COLUMNS - FWHM1_COLUMN) / 3 of the reference (acis_fef.c:328-338)

struct SourceDev
{
   int source_type, spectrum_type;
   double p[3], p_normal[3];
   double distance;
   double emin, emax;
   const double *spec_energies, *spec_cum_flux;
   uint32_t spec_num;
   double mean_time;               // 1/flux/area, source.c:260-264
   double shape[3];                // extended-source parameters (marxb200_source_desc.shape)
   double rot_axis[3], rot_angle;  // LINE / IMAGE: rotation taking (-1,0,0) to p
   const float *image_cdf;         // IMAGE: cumulative image (HBM; 1 MB for 512^2, L2 resident)
   uint32_t image_size, image_nx, image_ny;
   double rad_per_xpixel, rad_per_ypixel;
};

struct DitherDev
{
   int mode;
   double ra_amp, dec_amp, roll_amp, ra_period, dec_period, roll_period, ra_phase, dec_phase, roll_phase;
   double nominal_roll, aspect_blur;
   const double *aspsol;           // mode 2: [num_aspsol][7] t, ra, dec, roll, dy, dz, dtheta (marxb200_dither_desc)
   uint32_t num_aspsol;
};

struct WfoldDev
{
   uint32_t num_arrays;
   const double *hdr;              // [num_arrays][6]: e_alpha, p_min, delta_p, p_max, norm, expon
   const uint32_t *num_theta;      // [num_arrays]
   const uint32_t *theta_offset;   // [num_arrays]
   const float *theta;
   double min_p_min;               // smallest p_min of the table: below it no array scatters (wfold.c:311-312)
};

struct HrmaShellDev
{
   double conic_p[5];              // a, b, c, xmin, xmax
   double conic_h[5];
   double to_osac_p[3], to_osac_h[3];
   double front_position, area_fraction, min_radius, max_radius;
   double p_blur, h_blur, p_scat, h_scat;
   double fwd_p[9], bwd_p[9], fwd_h[9], bwd_h[9];
   uint32_t shutter_bitmap, num_corr;
   uint32_t corr_offset;           // offset (floats) of this shell's (E, factor) pairs in HrmaDev.corr_*
   uint32_t pad;
   WfoldDev wfold_p, wfold_h;
};

struct HrmaDev
{
   HrmaShellDev shell[kNumShells];
   double vig, cap_position;
   int is_ideal, use_blur, use_wfold, use_struts, use_scale;
   uint32_t num_opt;
   const float *opt_energies, *opt_betas, *opt_deltas;   // [num_opt] each, contiguous (3*num_opt floats)
   const float *corr_energies, *corr_factors;            // concatenated over shells
   uint32_t num_corr_total;
};

struct GratingShellDev
{
   uint32_t num_orders, num_energies, num_sectors, pad;
   const int32_t *order_list;
   const float *energies;
   const float *cum_eff;           // [num_orders][num_energies]
   const double *sectors;          // [num_sectors][6]: min, max, dtheta, dtheta_blur, dpp, dpp_blur
   double dispersion_angle, period, dp_over_p, theta_blur, vig, rowland;
   double cos_dispersion, sin_dispersion;   // of dispersion_angle, evaluated on the host (diffract.c:689-700 uses libm per photon)
};

struct GratingDev
{
   int type;
   GratingShellDev shell[kNumShells];
   GratingShellDev support[2];     // LETG fine / coarse support gratings (num_orders == 0: absent)
};

struct FefDev
{
   uint32_t num_gaussians, num_energies;
   const float *energies, *channels;
   const float *gauss;             // [num_energies][num_gaussians][3] = amp, center, sigma
};

struct AcisChipDev
{
   int id;
   double x_ll[3], xhat[3], yhat[3], normal[3];
   double xlen, ylen, x_pixel_size, y_pixel_size, xpixel_offset, ypixel_offset;
   uint32_t qe_num, filter_num;
   const float *qe_energies, *qe, *filter_energies, *filter_qe;
   uint32_t contam_num_layers;
   int contam_fxy_mode;
   double contam_tau0[kMaxContamLayers], contam_tau1[kMaxContamLayers];
   uint32_t contam_num_mu[kMaxContamLayers];
   const float *contam_energies[kMaxContamLayers], *contam_mus[kMaxContamLayers], *contam_fxy[kMaxContamLayers];
   double contam_x0, contam_y0;
   uint32_t contam_blocking, pad;
   const int32_t *fef_map;         // [32][32]
};

struct AcisDev
{
   int detector_type, num_chips;
   AcisChipDev chip[kMaxChips];
   const FefDev *fefs;
   uint32_t num_fefs;
   double det_offset[3], det_matrix[9];
   int det_ideal, det_extend, dither_mode, pad;
   double focal_length, exposure_time, frame_transfer_time, frame_time;
};

struct HrcMcpDev
{
   int id; uint32_t qe_num;
   double x_ll[3], xhat[3], yhat[3], normal[3], xlen, ylen;
   const float *qe_energies, *qe;
   double u_start, v_start, u_0, v_0, cx_0, cy_0;
};
struct HesfPlateDev { double a[3], e1[3], e2[3], normal[3], len1, len2; };
struct HrcDev
{
   int detector_type, num_mcps;
   HrcMcpDev mcp[3];
   uint32_t filter_num[4];
   const float *filter_energies[4], *filter_qe[4];
   double shield_t, shield_l, shield_r, shield_x, shield_sl, shield_sr, shield_sl_gap, shield_sr_gap, shield_y_center, shield_z_center;
   double blur[13];
   double u_pixel_size, v_pixel_size;
   double det_offset[3], det_matrix[9];
   int det_ideal, det_extend, use_hesf, hesf_num_plates;
   HesfPlateDev hesf[8];
   double hesf_cr_width;
   uint32_t c_num, cr_num;
   const float *c_energies, *c_betas, *c_deltas, *cr_energies, *cr_betas, *cr_deltas;
};

}  // namespace mx
