/* marxb200.h -- C ABI of the B200-native MARX ray-trace path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch types.  The host
 * side of MARX stays C (marx/src/marx.c driver, pfile, jdfits); its module functions call the entry
 * points below.  Each entry point cites the reference interface it replaces (paths relative to the
 * MARX 5.5.3 tree).  INTEGRATION.md shows the binding a MARX maintainer would add.
 *
 * Conventions (same as the reference, marx/libsrc/marxerr.c:26-54): every function returns 0 on
 * success and -1 on error; the message is retrievable with marxb200_last_error().  There is no CPU
 * fallback: if no CUDA device is usable, marxb200_create fails.
 *
 * Data model.  The reference's array-of-structs Marx_Photon_Attr_Type (136 B, marx/libsrc/marx.h:51-100)
 * becomes a structure-of-arrays photon buffer in HBM owned by the context.  Stage calls compact the
 * live list (the reference's marx_prune_photons, marx/libsrc/photon.c:40-63) in arrival order.
 * Random draws are counter based: draw k of (ray, stage) is lane k&3 of
 * Philox4x32-10(key = seed, counter = (ray_lo, ray_hi, k>>2, stage)) mapped to [0,1] as
 * u32 * (1/4294967295.0)  (jdmath/src/random.c:151-154), so results do not depend on batch size,
 * launch geometry or GPU count.
 */
#ifndef MARXB200_H
#define MARXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MARXB200_ABI_VERSION 2
#define MARXB200_NUM_SHELLS 4          /* MARX_NUM_MIRRORS, marx/libsrc/_marx.h:46 */
#define MARXB200_MAX_CHIPS 6           /* _MARX_NUM_ACIS_S_CHIPS, _marx.h:41 */
#define MARXB200_MAX_CONTAM_LAYERS 5   /* MAX_LAYERS, marx/libsrc/aciscontam.c:76 */

typedef struct marxb200_ctx marxb200_ctx;

/* RNG sub-stream ids (counter word 3) */
enum { MARXB200_STAGE_SOURCE = 0, MARXB200_STAGE_MIRROR = 1, MARXB200_STAGE_GRATING = 2, MARXB200_STAGE_DETECTOR = 3 };

/* photon flags: identical to marx/libsrc/marx.h:65-75 */
#define MARXB200_PHOTON_UNDETECTED       0x01
#define MARXB200_PHOTON_UNREFLECTED      0x02
#define MARXB200_PHOTON_UNDIFFRACTED     0x04
#define MARXB200_PHOTON_MISSED_DETECTOR  0x08
#define MARXB200_PHOTON_MIRROR_VBLOCKED  0x10
#define MARXB200_PHOTON_DRAKE_BLOCKED    0x20
#define MARXB200_PHOTON_GRATING_VBLOCKED 0x40
#define MARXB200_BAD_PHOTON_MASK         0xFF
#define MARXB200_PHOTON_DRAKE_REFLECTED  0x100
#define MARXB200_PHOTON_ACIS_STREAKED    0x200

/* Byte-compatible image of Marx_Photon_Attr_Type (marx/libsrc/marx.h:51-100; sizeof == 136 on x86-64).
 * Used only at the host boundary (upload / download); never stored in HBM. */
typedef struct
{
   double energy;
   double x[3];
   double p[3];
   double arrival_time;
   uint32_t flags;
   float y_pixel, z_pixel, u_pixel, v_pixel;
   float dither_ra, dither_dec, dither_roll, dither_dy, dither_dz, dither_dtheta;
   float pi;
   int16_t pulse_height;
   uint32_t mirror_shell;
   int8_t ccd_num;
   int8_t detector_region;
   int8_t order;
   int8_t support_orders[4];
   uint32_t tag;
}
marxb200_photon_attr;

/* ------------------------------------------------------------------------------------------------
 * Module descriptors.  All pointers are HOST pointers; the library copies what it needs to the
 * device during the call, the caller keeps ownership.  Field meanings are those of the reference
 * statics they are filled from (cited per field group).
 * ---------------------------------------------------------------------------------------------- */

/* source + spectrum + arrival times: marx/libsrc/source.c:136-215,260-264 ; spectrum.c:138-181 */
typedef struct
{
   int32_t source_type;            /* 0 POINT (s-point.c:59-83), 1 GAUSS (s-gauss.c:78-141), 2 BETA (s-beta.c:81-139),
                                      3 DISK (s-disk.c:63-109), 4 LINE (s-line.c:62-104), 5 IMAGE (s-image.c:287-362) */
   int32_t spectrum_type;          /* 1 = FLAT, 2 = FILE (MARX_*_SPECTRUM, marx.h) */
   double p[3];                    /* unit vector FROM source TO origin (Marx_Source_Type.p) */
   double p_normal[3];
   double distance;                /* mm; <= 0: infinity */
   double emin, emax;              /* FLAT */
   const double *spec_energies;    /* FILE: inverse-CDF table (prob.c:46-58) */
   const double *spec_cum_flux;
   uint32_t spec_num;
   double total_flux;              /* photons/s/cm^2 */
   double geometric_area;          /* Marx_Mirror_Geometric_Area, cm^2 (hrma.c:736) */
   double shape[3];                /* GAUSS: sigma (rad); BETA: core radius (rad), 1/(1-alpha) with alpha = 3 beta - 1/2;
                                      DISK: theta_max (rad), x0 = (theta_min/theta_max)^2, x1 = 1 - x0;
                                      LINE: S-LineTheta (rad), cos and sin of S-LinePhi */
   /* LINE and IMAGE build the ray in a frame centred on (-1,0,0) and rotate it to p: axis and angle of that
    * rotation (JDMv_find_rotation_axis, s-line.c:77; s-image.c:303-319) */
   double rot_axis[3], rot_angle;
   const float *image_cdf;         /* IMAGE: normalised cumulative image, f32, [image_ny][image_nx] (s-image.c:138-148) */
   uint32_t image_nx, image_ny;
   double rad_per_xpixel, rad_per_ypixel;
}
marxb200_source_desc;

/* aspect dither: marx/libsrc/dither.c:50-72,167-182,508-547 (angles already in radians) */
typedef struct
{
   int32_t mode;                   /* 0 = NONE, 1 = INTERNAL, 2 = ASPSOL (DitherModel=FILE, dither.c:288-500) */
   double ra_amp, dec_amp, roll_amp;
   double ra_period, dec_period, roll_period;
   double ra_phase, dec_phase, roll_phase;
   double nominal_roll;
   double aspect_blur;
   /* ASPSOL: the states the stock reader steps through (get_single_aspsol_point, dither.c:288-359), in file order,
    * state 0 being the one init_aspsol_dither leaves behind: [num_aspsol][7] = t (s since TSTART), ra, dec, roll (rad,
    * unrolled offsets), dy, dz (mm), dtheta (rad).  A ray at time t uses the first state k >= 1 with t < t_k and its
    * predecessor; rays at or beyond the last state end the simulation (the reference stops at the end of the file). */
   const double *aspsol;
   uint32_t num_aspsol;
}
marxb200_dither_desc;

/* one WFOLD scattering table (marx/libsrc/wfold.c:42-60): arrays sorted by e_alpha */
typedef struct
{
   uint32_t num_arrays;
   const double *e_alpha, *p_min, *delta_p, *p_max, *pow_law_norm, *pow_law_expon;  /* [num_arrays] */
   const uint32_t *num_theta;      /* [num_arrays] */
   const uint32_t *theta_offset;   /* [num_arrays] offset of each array in theta_values */
   const float *theta_values;      /* concatenated */
   uint32_t total_theta;
}
marxb200_wfold_table;

/* one HRMA shell: derived fields of HRMA_Type (marx/libsrc/hrma.c:51-128,551-602,671-745,842-897) */
typedef struct
{
   uint32_t mirror_number;
   uint32_t shutter_bitmap;
   double conic_a_p, conic_b_p, conic_c_p, conic_xmin_p, conic_xmax_p;
   double conic_a_h, conic_b_h, conic_c_h, conic_xmin_h, conic_xmax_h;
   double to_osac_p[3], to_osac_h[3];
   double front_position;
   double area_fraction;           /* cumulative, normalised */
   double min_radius, max_radius;
   double p_blur, h_blur;          /* arcsec */
   double p_scat_factor, h_scat_factor;
   double fwd_matrix_p[9], bwd_matrix_p[9], fwd_matrix_h[9], bwd_matrix_h[9];
   const float *corr_energies, *corr_factors;   /* hrma/corr_<n>.dat */
   uint32_t num_corr;
   marxb200_wfold_table p_wfold, h_wfold;
}
marxb200_hrma_shell;

typedef struct
{
   marxb200_hrma_shell shells[MARXB200_NUM_SHELLS];
   double vignetting_factor;       /* HRMAVig */
   double cap_position;            /* _Marx_HRMA_Cap_Position */
   int32_t is_ideal, use_blur, use_wfold, use_struts, use_scale_factors;
   const float *opt_energies, *opt_betas, *opt_deltas;  /* hrma/iridium.dat */
   uint32_t num_opt;
}
marxb200_hrma_desc;

/* MirrorType=FLATFIELD (marx/libsrc/ffield.c:45-108): no optics; every ray starts at a uniformly drawn point of the rectangle
 * [min_y, max_y] x [min_z, max_z] in the plane x = x_pos (two draws on the MIRROR sub-stream: z first, then y), its direction
 * re-aimed for a source at finite distance; the mirror shell stays 0 */
typedef struct { double min_y, min_z, max_y, max_z, x_pos; } marxb200_flatfield_desc;

/* gratings: Grating_Type / Grating_Sector_Type (marx/libsrc/diffract.c:51-91) */
typedef struct
{
   uint32_t num_orders;
   const int32_t *order_list;      /* [num_orders] */
   uint32_t num_energies;
   const float *energies;          /* [num_energies] */
   const float *cum_eff;           /* [num_orders][num_energies], cumulative */
   double dispersion_angle;        /* radians */
   double period;                  /* um */
   double dp_over_p;
   double theta_blur;              /* radians */
   double vig;
   uint32_t num_sectors;           /* 0: statistical blur (diffract.c:771-777) */
   const double *sec_min_angle, *sec_max_angle, *sec_dtheta, *sec_dtheta_blur, *sec_dpp, *sec_dpp_blur;
}
marxb200_grating_shell;

typedef struct
{
   int32_t type;                   /* 0 none, 1 HETG, 2 LETG (MARX_GRATING_*, marx.h:364-376) */
   marxb200_grating_shell shells[MARXB200_NUM_SHELLS];
   double rowland[MARXB200_NUM_SHELLS];   /* diameter per shell (diffract.c:885-904) */
   /* LETG only: fine [0] and coarse [1] support gratings (diffract.c:917-970, :1098-1118); num_orders == 0: absent.
    * Their efficiencies are tabulated on the 1024-point grid of diffract.c:108-110. */
   marxb200_grating_shell support[2];
}
marxb200_grating_desc;

/* one FEF region function: Fef_Type (marx/libsrc/acis_fef.c:87-106) */
typedef struct
{
   uint32_t num_gaussians, num_energies;
   const float *energies, *channels;   /* [num_energies] */
   const float *gauss;                 /* [num_energies][num_gaussians][3] = (amp, center, sigma) */
}
marxb200_fef;

/* one ACIS chip: Marx_Detector_Geometry_Type (marx.h:398-425) + _Marx_Acis_Chip_Type QE (_marx.h:53-)
 * + Single_Component_Contam_Type (aciscontam.c:76-91) + FEF region map (acis_fef.c:108-115) */
typedef struct
{
   int32_t id;
   double x_ll[3], xhat[3], yhat[3], normal[3];
   double xlen, ylen, x_pixel_size, y_pixel_size, xpixel_offset, ypixel_offset;
   uint32_t qe_num; const float *qe_energies, *qe;
   uint32_t filter_num; const float *filter_energies, *filter_qe;
   uint32_t contam_num_layers;
   double contam_tau0[MARXB200_MAX_CONTAM_LAYERS], contam_tau1[MARXB200_MAX_CONTAM_LAYERS];
   uint32_t contam_num_mu[MARXB200_MAX_CONTAM_LAYERS];
   const float *contam_energies[MARXB200_MAX_CONTAM_LAYERS], *contam_mus[MARXB200_MAX_CONTAM_LAYERS];
   int32_t contam_fxy_mode;        /* 0: table fxy_vals; 1: fxy_acis_i(x0,y0); 2: fxy_acis456789 (aciscontam.c:141-187) */
   double contam_x0, contam_y0;
   uint32_t contam_blocking;       /* FXYBLK */
   const float *contam_fxy[MARXB200_MAX_CONTAM_LAYERS];   /* [(1024/blk)^2] */
   const int32_t *fef_map;         /* [32][32] -> index into marxb200_acis_desc.fefs, -1: none (i = x/32 major) */
}
marxb200_acis_chip;

typedef struct
{
   int32_t detector_type;          /* MARX_DETECTOR_ACIS_S = 2 ... ; 0 = none */
   int32_t num_chips;
   marxb200_acis_chip chips[MARXB200_MAX_CHIPS];   /* facet-list order = intersection priority */
   uint32_t num_fefs;
   const marxb200_fef *fefs;
   double det_offset[3];           /* _Marx_Det_XForm_Matrix.dx,dy,dz (detector.c:40-54) */
   double det_matrix[9];
   int32_t det_ideal, det_extend;
   double focal_length;            /* Marx_Focal_Length */
   double exposure_time, frame_transfer_time, frame_time;   /* acis-i.c:171-200 */
   int32_t dither_mode;            /* _Marx_Dither_Mode != 0: detector dither applied (detector.c:275-295) */
}
marxb200_acis_desc;

/* HRC-S (marx/libsrc/hrc-s.c, hrc_s_geom.c, hrcblur.c, hrc-i.c:66-85) with the HESF "Drake flat" (drake.c) */
typedef struct
{
   int32_t id;                     /* MCP id 1..3 */
   double x_ll[3], xhat[3], yhat[3], normal[3], xlen, ylen;
   uint32_t qe_num; const float *qe_energies, *qe;      /* MCP_QEs[Mcp_Id_Mapping[id]] (hrc-s.c:44-56) */
   double u_start, v_start, u_0, v_0, cx_0, cy_0;       /* _marx_hrc_s_compute_pixel constants (hrc_s_geom.c:344-394) */
}
marxb200_hrc_mcp;

typedef struct
{
   double a[3], e1[3], e2[3], normal[3], len1, len2;    /* Rectangle_Type, drake.c:60-69 */
}
marxb200_hesf_plate;

typedef struct
{
   int32_t detector_type;          /* MARX_DETECTOR_HRC_S */
   int32_t num_mcps;
   marxb200_hrc_mcp mcps[3];       /* facet-list order */
   uint32_t filter_num[4]; const float *filter_energies[4], *filter_qe[4];   /* UV/ion shield regions 0..3 */
   double shield_t, shield_l, shield_r, shield_x, shield_sl, shield_sr, shield_sl_gap, shield_sr_gap;
   double shield_y_center, shield_z_center;              /* hrc-s.c:66-83,431-441 */
   double blur[13];                /* Marx_HRC_Blur_Parm_Type after fixup_blur_parms (hrcblur.c:52-70) */
   double u_pixel_size, v_pixel_size;
   double det_offset[3], det_matrix[9];
   int32_t det_ideal, det_extend;
   int32_t use_hesf, hesf_num_plates;                    /* rectangles: 2 * hesf_num_plates (drake.c:233-262) */
   marxb200_hesf_plate hesf[8];
   double hesf_cr_width;
   uint32_t c_num, cr_num; const float *c_energies, *c_betas, *c_deltas, *cr_energies, *cr_betas, *cr_deltas;
}
marxb200_hrc_s_desc;

/* ------------------------------------------------------------------------------------------------ */
/* life cycle                                                                                        */
int marxb200_abi_version (void);
const char *marxb200_last_error (void);
/* device_ordinal: CUDA device; seed: RandomSeed (marx.c:846-849, JDMsrandom). */
int marxb200_create (marxb200_ctx **ctxp, int device_ordinal, uint64_t seed);
/* Optional: create the CUDA context of a device ahead of marxb200_create, e.g. from a helper thread while the host program still
 * reads its calibration files (marx.c:257-318: the *_init calls of the stock modules).  Context creation costs 0.5 - 3 s of a
 * process's life; it is the largest fixed cost of a `marx` run on the device.  Safe to call from any thread, any number of times. */
int marxb200_device_warmup (int device_ordinal);
int marxb200_destroy (marxb200_ctx *ctx);
/* use an externally owned cudaStream_t (e.g. the harness's timing stream); NULL = library stream */
int marxb200_set_stream (marxb200_ctx *ctx, void *cuda_stream);

/* on (default): every stage compacts survivors, in arrival order, into the other SoA buffer.
 * off: stages work in place and keep dead rays with their flags (parity tests compare slot by slot). */
int marxb200_set_compaction (marxb200_ctx *ctx, int on);

/* table upload: called once after the stock *_init functions have loaded the calibration files
 * (marx_mirror_init/marx_grating_init/marx_detector_init/marx_create_source, marx.h:287,354-362) */
int marxb200_set_source (marxb200_ctx *ctx, const marxb200_source_desc *d);
int marxb200_set_dither (marxb200_ctx *ctx, const marxb200_dither_desc *d);
int marxb200_set_hrma (marxb200_ctx *ctx, const marxb200_hrma_desc *d);
int marxb200_set_flatfield (marxb200_ctx *ctx, const marxb200_flatfield_desc *d);      /* instead of marxb200_set_hrma */
int marxb200_set_grating (marxb200_ctx *ctx, const marxb200_grating_desc *d);
int marxb200_set_acis (marxb200_ctx *ctx, const marxb200_acis_desc *d);
int marxb200_set_hrc_s (marxb200_ctx *ctx, const marxb200_hrc_s_desc *d);
/* convenience: read a calibration pack (tools/ + DESIGN.md "calpack") and call the setters above */
int marxb200_load_calpack (marxb200_ctx *ctx, const char *path);

/* photon buffer: marx_alloc_photon_type / marx_dealloc_photon_type (photon.c:67-106) */
int marxb200_alloc_photons (marxb200_ctx *ctx, uint64_t max_photons);

/* marx_create_photons (source.c:268-384): generate n rays with global ray indices
 * [first_ray, first_ray+n): energies, directions, Poisson arrival times (running sum continued from the
 * previous call, source.c:285,326), dither.  No energy sort is needed (draws are per-ray).
 * time_base_in < 0: continue from the context's running time; otherwise restart the sum there. */
int marxb200_create_photons (marxb200_ctx *ctx, uint64_t first_ray, uint64_t n, double time_base_in);
/* ExposureTime handling of marx_create_photons (source.c:323-334, marx.c:545-556): call directly after
 * marxb200_create_photons.  Keeps the rays up to AND INCLUDING the first one whose arrival time, counted from the
 * start of this batch, is >= exposure_left; later rays are dropped and the running end time is set to the last kept
 * ray's arrival time.  *n_kept is the reference's *num_collected. */
int marxb200_truncate_exposure (marxb200_ctx *ctx, double exposure_left, uint64_t *n_kept);
/* Multi-GPU time base: sums of the arrival-time increments of rays [first_ray, first_ray+n) per
 * super-tile of 65536 rays, in canonical order.  Ranks all-gather these (tiny) vectors and add them
 * sequentially to obtain the time_base_in of their block, which makes arrival times independent of
 * the number of GPUs (the reference's analogue: marxcat's time-ordered merge, marx/src/marxcat.c:505-535). */
int marxb200_time_sums (marxb200_ctx *ctx, uint64_t first_ray, uint64_t n, double *sums_host, uint64_t max_sums, uint64_t *n_sums);
/* marx_mirror_reflect (mirror.c:93 -> hrma.c:1161-1341) */
int marxb200_mirror_reflect (marxb200_ctx *ctx);
/* marx_grating_diffract (grating.c:87 -> diffract.c:974-1130) */
int marxb200_grating_diffract (marxb200_ctx *ctx);
/* marx_detect (detector.c:361-379 -> acis-s.c:177-248 | acis-i.c:96-166 | hrc-s.c:236-312 with drake.c:317-372) */
int marxb200_detect (marxb200_ctx *ctx);
/* The compacting stage kernels emit survivors in completion order; this puts the live list back into
 * arrival order (what marx_prune_photons preserves, photon.c:40-63).  Called implicitly by marxb200_trace
 * and by every download; exposed so that a caller timing individual stages can account for it. */
int marxb200_restore_order (marxb200_ctx *ctx);
/* all of the above for one batch, device resident (marx.c:569 + process_photons :240-273) */
int marxb200_trace (marxb200_ctx *ctx, uint64_t first_ray, uint64_t n);
/* same with an explicit time base (multi-GPU blocks, see marxb200_time_sums); time_base_in < 0: continue */
int marxb200_trace_from (marxb200_ctx *ctx, uint64_t first_ray, uint64_t n, double time_base_in);

/* Per-kernel device times (CUDA events on the launching stream), for roofline bookkeeping.  When enabled, an
 * event is recorded after every kernel launch; marxb200_get_kernel_ms synchronises, returns the accumulated
 * milliseconds and launch counts per kernel class since the last call, and resets them.  Classes:
 * 0 k0_time_sums, 1 k0_time_scan (k0_time_super + k0_time_bases + k0_time_tiles), 2 k0_source, 3 k01_source_hrma (fused), 4 k1_hrma<0>, 5 k1_hrma<1>, 6 k1_hrma<2>,
 * 7 k2_grating, 8 k3 (acis or hrc), 9 order restoration (5 kernels), 10 Level-1 transforms (marxb200_level1_transform),
 * 11 k1_hrma<3> (HRMA B1), 12 k1_hrma<4> (B2+C1), 13 k1_hrma<5> (C2): the compacting path's cut of the mirror stage behind the
 * reflectivity tests (it then runs 3 | 11 | 12 | 13 in place of 3 | 5 | 6). */
#define MARXB200_NUM_KERNEL_CLASSES 14
int marxb200_set_profiling (marxb200_ctx *ctx, int on);
int marxb200_get_kernel_ms (marxb200_ctx *ctx, double ms[MARXB200_NUM_KERNEL_CLASSES], uint64_t launches[MARXB200_NUM_KERNEL_CLASSES]);

/* number of live photons / generated rays / running time after the last stage (synchronises) */
int marxb200_get_counts (marxb200_ctx *ctx, uint64_t *n_generated, uint64_t *n_live, double *total_time);
/* per-stage live counts of the last batch: [generated, after mirror, after grating, detected]
 * (the reference's PRINT_STATS_ARRAY, marx.c:68,236-271) */
int marxb200_get_stage_counts (marxb200_ctx *ctx, uint64_t counts[4]);
/* same plus the counts between the HRMA sub-kernels: [4] after phase A, [5] after phase B ([6],[7] reserved) */
int marxb200_get_internal_counts (marxb200_ctx *ctx, uint64_t counts[8]);

/* host boundary.  download: live photons, arrival order, into AoS records (what marx_write_photons,
 * marxio.c:403-476, and the pipe/rayfile writers consume).  upload: inject photons at a stage boundary
 * (the reference's RAYFILE channel, s-rayfile.c:188-221); tags are the ray indices used for draws. */
int marxb200_download (marxb200_ctx *ctx, marxb200_photon_attr *out, uint64_t max_out, uint64_t *n_out);
int marxb200_upload (marxb200_ctx *ctx, const marxb200_photon_attr *in, uint64_t n, const uint64_t *ray_ids);
/* same for records whose arrival_time counts from the start of their batch, as in a Marx_Photon_Type filled by the
 * RAYFILE source (s-rayfile.c:188-221): start_time = pt->start_time */
int marxb200_upload_from (marxb200_ctx *ctx, const marxb200_photon_attr *in, uint64_t n, const uint64_t *ray_ids, double start_time);
/* debug/parity: download EVERY photon slot of the last stage call, dead ones included (flags say why) */
int marxb200_download_all (marxb200_ctx *ctx, marxb200_photon_attr *out, uint64_t max_out, uint64_t *n_out);

/* Event tallies: exact integer histograms of the live list, accumulated on the device over any number of batches, so
 * that per-GPU results are merged with ONE all-reduce (ncclAllReduce sum over the buffer marxb200_tally_device_ptr
 * returns) instead of moving events.  The reference has no histogram module: these are the tallies its
 * users derive from the event files (order populations, PHA / PI / energy spectra, chip and focal-plane images), and
 * the tests check them against the same binning of the downloaded events.
 * An axis bins column v as floor ((v - lo) * nbins / (hi - lo)); events outside [lo, hi) on any axis are not counted.
 * Integer-valued columns (PHA, ORDER, CCD, SHELL) bin exactly with unit-width bins, e.g. ORDER lo=-11 hi=12 nbins=23. */
enum
{
   MARXB200_TALLY_ENERGY = 0, MARXB200_TALLY_TIME, MARXB200_TALLY_PHA, MARXB200_TALLY_PI, MARXB200_TALLY_ORDER,
   MARXB200_TALLY_CCD, MARXB200_TALLY_SHELL, MARXB200_TALLY_CHIPX, MARXB200_TALLY_CHIPY, MARXB200_TALLY_YPOS,
   MARXB200_TALLY_ZPOS
};
typedef struct { int32_t column; uint32_t nbins; double lo, hi; } marxb200_tally_axis;
/* naxes = 1 or 2 (row-major: bin = bin0 * nbins1 + bin1); returns the tally id (>= 0) or -1 */
int marxb200_tally_create (marxb200_ctx *ctx, const marxb200_tally_axis *axes, int naxes);
/* bin the current live list (after whichever stage ran last) into tally `id`; asynchronous on the context's stream */
int marxb200_tally_accumulate (marxb200_ctx *ctx, int id);
int marxb200_tally_reset (marxb200_ctx *ctx, int id);
/* copy the counters (uint64, nbins0 * nbins1 of them) to the host */
int marxb200_tally_read (marxb200_ctx *ctx, int id, uint64_t *out, uint64_t max_bins);
/* the device buffer itself, for an in-place all-reduce; synchronises the context's stream first */
int marxb200_tally_device_ptr (marxb200_ctx *ctx, int id, void **dev_ptr, uint64_t *num_bins);

/* column download of the live list without the AoS detour (bulk egress).
 * Any pointer may be NULL.  Arrays must hold marxb200_get_counts().n_live entries. */
typedef struct
{
   double *energy, *time, *xpos, *ypos, *zpos, *xcos, *ycos, *zcos;
   float *chipx, *chipy, *pi;
   int16_t *pha;
   int8_t *ccd, *order, *shell;
   uint64_t *ray;
}
marxb200_columns;
int marxb200_download_columns (marxb200_ctx *ctx, const marxb200_columns *cols, uint64_t max_out, uint64_t *n_out);

/* Pipelined egress: marxb200_egress_begin restores arrival order, snapshots up to max_out live
 * photons into a device staging area and returns at once, so the next batch can be launched; marxb200_egress_end
 * waits for that snapshot, copies the requested columns to the caller's (ideally pinned) host arrays on a private
 * copy stream -- overlapping the next batch's kernels -- and blocks until they have landed. */
int marxb200_egress_begin (marxb200_ctx *ctx, uint64_t max_out);
int marxb200_egress_end (marxb200_ctx *ctx, const marxb200_columns *cols, uint64_t *n_out);

/* Bulk event-file egress: marx_write_photons (marxio.c:403-476) for the live list on the device.
 * Writes/appends the column files of the reference's output directory (energy.dat, time.dat, xpixel.dat, ... -- the
 * table at marxio.c:292-322), byte-compatible with the reference's writer so that marx2fits, marxcat and marxpileup read
 * them unchanged: 32-byte header (marxio.c:151-205) + big-endian float32/int16/int32/int8 data, row count patched at
 * offset 20 on every call.  Columns are converted and byte-swapped on the device and appended with one fwrite per file.
 *   write_mask  the columns to write: the caller's OutputVectors mask ANDed with the photon history, exactly the value
 *               marxio.c:409-414 computes (bits below = MARX_*_OK, marx.h:126-147)
 *   open_mode   1: create the files (first batch), 0: append (marx.c:585-588)
 *   total_time  the driver's accumulated time before this batch; TIME = arrival_time + total_time (marxio.c:246) */
#define MARXB200_ENERGY_OK        0x00000001
#define MARXB200_TIME_OK          0x00000002
#define MARXB200_X_VECTOR_OK      0x00000004
#define MARXB200_P_VECTOR_OK      0x00000008
#define MARXB200_TAG_OK           0x00000010
#define MARXB200_PULSEHEIGHT_OK   0x00000020
#define MARXB200_PI_OK            0x00000040
#define MARXB200_DET_PIXEL_OK     0x00000080
#define MARXB200_DET_NUM_OK       0x00000100
#define MARXB200_DET_REGION_OK    0x00000200
#define MARXB200_DET_UV_PIXEL_OK  0x00000400
#define MARXB200_MIRROR_SHELL_OK  0x00000800
#define MARXB200_SKY_DITHER_OK    0x00001000
#define MARXB200_DET_DITHER_OK    0x00002000
#define MARXB200_ORDER_OK         0x00100000
#define MARXB200_ORDER1_OK        0x00200000
#define MARXB200_ORDER2_OK        0x00400000
#define MARXB200_ORDER3_OK        0x00800000
#define MARXB200_ORDER4_OK        0x01000000
int marxb200_write_photons (marxb200_ctx *ctx, const char *dir, uint64_t write_mask, int open_mode, double total_time);
/* Background column-file writer for marxb200_write_photons.  Appending a 2^24-ray batch's event columns to their files costs a
 * single host thread ten times what the GPU needs to trace the batch (even on tmpfs).  With n_threads > 0 marxb200_write_photons
 * returns as soon as the batch has landed in one of two pinned host buffers; column file k is appended by thread k mod n_threads
 * (the order of appends to one file is kept; different files are written in parallel) while the caller traces the next batch.
 * The files are byte-identical to the synchronous writer's.  marxb200_write_flush blocks until everything queued is in the files
 * and reports the first write error; marxb200_destroy and a call with open_mode = 1 flush as well.  n_threads = 0: flush and go
 * back to writing synchronously (the default). */
int marxb200_set_async_writer (marxb200_ctx *ctx, int n_threads);
int marxb200_write_flush (marxb200_ctx *ctx);

/* Pipelined form of the same egress without the file system: _begin_packed converts the selected columns of the live list
 * to their file images (the bytes marxb200_write_photons would append, big endian) in a device staging area and returns
 * at once; _end_packed copies them into the caller's (ideally pinned) host buffer on the private copy stream --
 * overlapping the next batch's kernels -- and describes where each column landed. */
typedef struct
{
   uint32_t num_cols;
   uint64_t n_rows;
   uint64_t mask[32];              /* the MARXB200_*_OK bit of each column */
   char file[32][16];              /* its file name in a MARX output directory (marxio.c:292-322) */
   char type[32];                  /* 'E' float32, 'I' int16, 'J' int32, 'A' int8 */
   uint32_t elem_size[32];
   uint64_t offset[32];            /* byte offset of the column in the host buffer; n_rows * elem_size bytes each */
}
marxb200_packed_layout;
int marxb200_egress_begin_packed (marxb200_ctx *ctx, uint64_t write_mask, double total_time, uint64_t max_out);
int marxb200_egress_end_packed (marxb200_ctx *ctx, void *host, uint64_t host_bytes, marxb200_packed_layout *layout);

/* ------------------------------------------------------------------------------------------------------------------
 * Level-1 event transforms: the per-event part of marx2fits (marx/src/marx2fits.c:3584-3943) for the
 * device-resident event list, so that `marx` + `marx2fits` need no intermediate column files.  marx2fits reads the
 * float32/int16/int8 column files of a MARX output directory row by row and derives, per event (compute order = the order
 * of Data_Def_Table, :274-1252):
 *   EXPNO      compute_expno :3741-3763        (long) (time / TimeDel); time is the float32 of time.dat
 *   TDETX/Y    compute_tdetxy :3584-3599       marx_compute_tiled_pixel, detpix.c:151-177 (acis_geom.c:111-143,181-202,
 *                                              hrc_s_geom.c:477-515, hrc_i_geom.c:216-236)
 *   (aspect)   read_dither_value :3567-3580    the 6 dither values are only taken over when EXPNO changed (ACIS) -- every
 *                                              event of an exposure frame carries the aspect of the frame's FIRST event
 *   FLTGRADE   compute_fltgrade :3854-3866     1 uniform draw per row (ACIS)
 *   GRADE      compute_grade :3813-3818
 *   DETX/Y     compute_detxy :3676-3737        pixel adjustment (NONE / RANDOMIZE: 2 draws / EDSER: acis_subpix.c:283-321 /
 *                                              EXACT), marx_init_chip_to_mnc + marx_chip_to_mnc (pixlib.c:139-225),
 *                                              marx_mnc_to_fpc (detpix.c:182-209)
 *   X/Y        compute_xy_sky :3869-3911       marx_undither_mnc (dither.c:583-607,703-707), marx_mnc_to_ra_dec +
 *                                              marx_compute_ra_dec_offsets (pixlib.c:503-573); roll rotation without dither
 *   ENERGY     compute_acis_energy :3923-3929, PI compute_pi :3931-3941, NODE_ID :3765-3771, STATUS :3775-3790
 *   TIME       write_time :3433-3445           TimeDel * EXPNO + TSTART (ACIS), time + TSTART (HRC)
 * Rows with pha == -1 are computed (they consume their draws) but not written (:2856-2857): `keep` = 0.
 * Draw d of event row r (r counts ALL rows since marxb200_level1_reset) = lane d&3 of Philox4x32-10 (key = seed, counter =
 * (r_lo, r_hi, d>>2, MARXB200_STAGE_LEVEL1)) mapped to [0,1] like every other draw of this library.
 * The descriptor holds what the stock initialisation (main :3200-3310, get_marx_pfile_info :2371-2562, read_obspar_file
 * :2961-2985) leaves in marx2fits' statics; INTEGRATION.md shows the binding.
 * ------------------------------------------------------------------------------------------------------------------ */
#define MARXB200_STAGE_LEVEL1 4
#define MARXB200_L1_MAX_CHIPS 10
enum { MARXB200_PIXADJ_NONE = 0, MARXB200_PIXADJ_RANDOMIZE = 1, MARXB200_PIXADJ_EDSER = 2, MARXB200_PIXADJ_EXACT = 3 };   /* marx2fits.c:60-63 */

typedef struct
{
   int32_t id;                      /* Marx_Detector_Geometry_Type.id (marx.h:397-425) */
   int32_t subpix_table;            /* EDSER: 0 = front-illuminated table, 1 = back-illuminated (acis_subpix.c:262-268: CCDs 5, 7) */
   double x_ll[3], xhat[3], yhat[3];
   double x_pixel_size, y_pixel_size, xpixel_offset, ypixel_offset;
   float tdet_xoff, tdet_yoff;      /* acis.h:38-40 / hrc.h:38-40 */
}
marxb200_level1_chip;

typedef struct
{
   int32_t detector_type;           /* MARX_DETECTOR_HRC_S = 1, HRC_I = 2, ACIS_S = 3, ACIS_I = 4 (marx.h:364-367) */
   int32_t num_chips;
   marxb200_level1_chip chips[MARXB200_L1_MAX_CHIPS];
   double fp_delta_s0, fp_x0, fp_y0;   /* Marx_FP_Coord_Type (detpix.c:41-65) */
   double focal_length;             /* marx2fits Focal_Length (marx.par FocalLength) */
   double det_offset[3];            /* DetOffsetX/Y/Z */
   double time_del;                 /* TimeDel: ACIS exposure (+ frame transfer) time, 0 for HRC (:2539-2542) */
   double time_start;               /* obs.par TSTART */
   double pi_factor;                /* 1 / ACIS_eV_Per_PI (:2490-2498) */
   double nominal_roll;             /* obs.par Roll_Nom, degrees */
   int32_t used_dither;             /* DitherModel != NONE */
   int32_t pix_adjust;              /* MARXB200_PIXADJ_*; EDSER on HRC means RANDOMIZE (main :3308-3309) */
   /* EDSER sub-pixel tables (Subpix_CCD_Type, acis_subpix.c:41-60): for table t and flight grade g, subpix_npoints[t*256+g]
    * points (0: no correction) stored at subpix_data + subpix_offset[t*256+g] as energies[n], dxs[n], dys[n] */
   const int32_t *subpix_npoints;
   const uint32_t *subpix_offset;
   const float *subpix_data;
   uint64_t subpix_data_len;
}
marxb200_level1_desc;

/* Level-1 columns of the live list, in the types marx2fits computes them in (its writers narrow DETX/Y, X/Y to float32, PI and
 * STATUS widen to int32: write_float64_as_float32 / write_int16_as_int32, :3370-3431).  Any pointer may be NULL. */
typedef struct
{
   double *time, *detx, *dety, *x, *y;
   int32_t *expno, *tdetx, *tdety, *pha, *hrc_u, *hrc_v;
   float *energy;
   int16_t *ccd_id, *node_id, *chipx, *chipy, *pi, *fltgrade, *grade, *status;
   uint8_t *keep;
}
marxb200_level1_columns;

int marxb200_set_level1 (marxb200_ctx *ctx, const marxb200_level1_desc *d);
/* start of a new event file: row counter 0, no exposure seen yet (the statics of compute_expno / read_dither_value) */
int marxb200_level1_reset (marxb200_ctx *ctx);
/* Transform the live list (arrival order is restored first).  total_time as in marxb200_write_photons: the file's TIME column
 * is (float) (arrival_time + total_time).  Consecutive calls continue one event file (row counter, exposure state). */
int marxb200_level1_transform (marxb200_ctx *ctx, double total_time);
int marxb200_level1_download (marxb200_ctx *ctx, const marxb200_level1_columns *cols, uint64_t max_out, uint64_t *n_out);

/* ------------------------------------------------------------------------------------------------
 * Aspect-solution table: the row loop of marxasp, marx/src/marxasp.c:996-1027 (compute_dither :814-884,
 * compute_quaternion :886-903).  Row i holds the pointing of the INTERNAL dither model at time i * delta_time: TIME, RA, DEC, ROLL
 * (degrees), dy = dz = dtheta = 0, and the attitude quaternion.  The descriptor is what marxasp's own initialisation derives from the
 * simulation directory (marx.par / obs.par; setup_dither :387-410: amplitudes in radians, the nominal pointing and its RA / Dec
 * unit vectors); marxasp.c keeps its parameter handling and its jdfits header writing and hands the row loop to this call.
 * ------------------------------------------------------------------------------------------------ */
typedef struct
{
   double time_start, delta_time;                                  /* TSTART (s), TimeDel */
   double ra_amp, dec_amp, roll_amp;                               /* radians */
   double ra_period, dec_period, roll_period;                      /* s */
   double ra_phase, dec_phase, roll_phase;                         /* radians */
   double nominal_roll;                                            /* radians */
   double pointing[3], ra_hat[3], dec_hat[3];
}
marxb200_aspsol_desc;
#define MARXB200_ASPSOL_ROW_BYTES 76
/* Rows [first_row, first_row + n) computed on the device.  cols_host (or NULL): 8 arrays of n doubles -- time, ra, dec, roll,
 * q0..q3.  fits_rows_host (or NULL): n * 76 bytes, the binary-table rows exactly as jdfits_write_float64 / _float32 lay them out
 * (big endian: time, ra, dec, roll, dy, dz, dtheta, q_att[4]), ready for one fwrite behind the header.  device_ms (or NULL):
 * duration of the kernel from CUDA events on the context's stream. */
int marxb200_aspsol_rows (marxb200_ctx *ctx, const marxb200_aspsol_desc *desc, uint64_t first_row, uint64_t n,
                          double *cols_host, void *fits_rows_host, double *device_ms);

/* ------------------------------------------------------------------------------------------------
 * ACIS pile-up: the frame loop of marxpileup, marx/src/marxpileup.c:main :1121-1213 (process_frame :890-922:
 * store_event :754-812, collect_charge :814-845, event_detect :676-752, will_grade_migrate :668-674, write_event :622-666).
 * The input columns are the event files of a simulation directory as read_input_event :573-620 reads them (detector.dat,
 * xpixel.dat, ypixel.dat, time.dat, b_energy.dat and the six dither files sky_ra ... det_theta, in that order in dither[]; a NULL
 * dither pointer drops that column), n rows in file order.  The output columns are those write_event writes; within an exposure
 * frame the rows come out in the reference's order (reverse file order).  alpha = the Alpha parameter; frame_time = FrameTime +
 * FrameTransferTime (initialize :1083-1084).  The PHA of an island's summed energy comes from the FEF tables of the ACIS detector
 * the context was set up with (marx_map_energy_to_acis_pha, acis_fef.c:1087-1096): marxb200_set_acis / a calibration pack first.
 *
 * Random draws: the reference draws from its one global generator once per island of >= 2 photons that passed the local-maximum
 * tests, in list order.  Here draw k (0, 1, ...) of exposure frame F is lane k&3 of Philox4x32-10 (key = seed, counter =
 * (F, 0, k >> 2, 5)) mapped to [0,1] as jdmath/src/random.c:151-154, so frames are independent.
 *
 * All pointers are HOST pointers (pinned memory makes the copies run at the PCIe rate); any output pointer may be NULL.  At most
 * max_out rows are written (more rows than that: error).  Lists whose exposure frames hold up to 513 events run as ONE kernel with
 * the frames staged in shared memory; longer frames take the eight step kernels (every event walks its own frame's run of the list,
 * O(frame length) each), and a frame longer than 65536 events is refused -- a limit the reference does not have (its per-CCD pixel
 * maps make a frame O(events)), stated here because a source that bright is outside the model's validity anyway.  Lists of 2^32 - 1
 * or more events are refused.  device_ms (or NULL): duration of the kernels from CUDA events on the context's stream.
 * ------------------------------------------------------------------------------------------------ */
typedef struct
{
   const int8_t *ccd; const float *x, *y, *t, *benergy;
   const float *dither[6];
}
marxb200_pileup_in;
typedef struct
{
   int8_t *ccd; float *x, *y, *t, *benergy; int32_t *frame; int16_t *nphotons, *pha;
   float *dither[6];
}
marxb200_pileup_out;
int marxb200_pileup_run (marxb200_ctx *ctx, uint64_t n, const marxb200_pileup_in *in, double alpha, double frame_time, uint64_t seed,
                         uint64_t max_out, const marxb200_pileup_out *out, uint64_t *n_out, double *device_ms);
/* The same for the event list the detector stage left on the device (behind marxb200_detect / marxb200_trace): `marx` followed by
 * `marxpileup` without the column files in between.  The input columns are what marx_write_photons would have written for that
 * list -- chip id, chip pixels, TIME = (float) (arrival_time + total_time), the PI energy of b_energy.dat, the six dither values
 * (marxio.c:217-290) -- gathered on the device; results are identical to marxb200_pileup_run on those files' contents. */
int marxb200_pileup_events (marxb200_ctx *ctx, double total_time, double alpha, double frame_time, uint64_t seed,
                            uint64_t max_out, const marxb200_pileup_out *out, uint64_t *n_out, double *device_ms);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, the exchanges run on device buffers over NVLink / NVSwitch inside this library
 * (NCCL, opened at run time with dlopen: a single-GPU caller needs none).  The reference's analogue is N independent `marx`
 * processes whose output directories marxcat merges in time order (marx/src/marxcat.c:491-535); here the ranks trace
 * contiguous blocks of ONE simulation, so that rays, draws, arrival times and events are identical for any number of GPUs.
 *
 *   communicator   rank 0 obtains an id (marxb200_comm_get_unique_id), hands it to the other ranks by whatever the launcher
 *                  offers (MPI_Bcast, a torch.distributed broadcast, an environment variable, a file), and every rank calls
 *                  marxb200_comm_init.  marxb200_comm_init_file does the hand-over through a file for launcher-less runs.
 *   trace          marxb200_trace_sharded (collective): rays [first_ray, first_ray + n_total) are split into `world` contiguous
 *                  blocks of ceil (n_total / world) rays rounded up to a multiple of 65536 (marxb200_shard_of); rank r traces
 *                  block r.  The running arrival-time sum (source.c:326) is continued across the blocks on the device: the
 *                  ranks all-gather the per-65536-ray sums their own pre-pass produced and add them in global ray order, so
 *                  every event carries the time one GPU tracing all blocks would give it, and every rank ends with the same
 *                  running time.  time_base_in as in marxb200_trace_from (< 0: continue).
 *   merge          marxb200_merge_events_begin / _end (collective): the ranks' event lists, converted on the device to the
 *                  column-file images marx_write_photons appends (marxio.c:217-322; the columns write_mask selects), are
 *                  concatenated in rank order -- arrival order -- in a buffer of the destination GPU.  _begin packs and returns
 *                  at once, so the next block can be launched; _end moves the columns (peer writes by the copy engines into the
 *                  destination's buffer, or ncclSend / ncclRecv) on a private stream while that block is traced.  TIME of the
 *                  merged list = (float) (arrival time since the start of the simulation + total_time).
 *   tallies        marxb200_tally_allreduce: the histograms of all ranks summed in place (ncclAllReduce on the context's stream).
 * ------------------------------------------------------------------------------------------------ */
#define MARXB200_COMM_ID_BYTES 128     /* sizeof (ncclUniqueId) */
#define MARXB200_MAX_RANKS 64
int marxb200_comm_get_unique_id (void *id /* MARXB200_COMM_ID_BYTES bytes */);
int marxb200_comm_init (marxb200_ctx *ctx, const void *id, int rank, int world);
int marxb200_comm_init_file (marxb200_ctx *ctx, const char *path, int rank, int world, double timeout_s);
/* merge_transport: 0 = ncclSend / ncclRecv, 1 = peer writes into the destination's buffer (valid after the first merge) */
int marxb200_comm_info (marxb200_ctx *ctx, int *rank, int *world, int *nccl_version, int *merge_transport);
int marxb200_comm_destroy (marxb200_ctx *ctx);
/* the block of rank `rank`: *my_n may be short or 0 at the end of the range */
int marxb200_shard_of (uint64_t first_ray, uint64_t n_total, int rank, int world, uint64_t *my_first_ray, uint64_t *my_n);
int marxb200_trace_sharded (marxb200_ctx *ctx, uint64_t first_ray, uint64_t n_total, double time_base_in,
                            uint64_t *my_first_ray, uint64_t *my_n);
int marxb200_tally_allreduce (marxb200_ctx *ctx, int id);

typedef struct
{
   uint32_t num_cols, world, dst_rank, transport;
   uint64_t n_rows;                         /* rows of the merged list = sum of rows_of_rank */
   uint64_t rows_of_rank[MARXB200_MAX_RANKS];
   void *device_base;                       /* destination rank: the merged columns in HBM (NULL elsewhere); valid until the next _end */
   uint64_t device_offset[32];              /* byte offset of column j from device_base; n_rows * elem_size[j] bytes */
   uint64_t mask[32];
   char file[32][16];
   char type[32];
   uint32_t elem_size[32];
   double transfer_ms;                      /* this rank's transfers + the closing barrier (waits for the slowest rank), CUDA events on the merge stream */
   double copy_ms;                          /* this rank's transfers alone: a source rank's nvlink_bytes / copy_ms is its achieved NVLink rate */
   uint64_t nvlink_bytes;                   /* bytes this rank moved over NVLink: received (destination) or sent */
}
marxb200_merged_layout;
int marxb200_merge_events_begin (marxb200_ctx *ctx, uint64_t write_mask, double total_time, uint64_t max_rows_per_rank, int dst_rank);
int marxb200_merge_events_end (marxb200_ctx *ctx, marxb200_merged_layout *layout);
/* destination rank: copy the merged columns to the host, laid out like marxb200_egress_end_packed's buffer */
int marxb200_merge_download (marxb200_ctx *ctx, void *host, uint64_t host_bytes, marxb200_packed_layout *layout);

/* Device-to-host copy rate into pinned memory (GB/s): `reps` copies of `bytes` on a private stream, CUDA-event timed.
 * flags & 1: cudaHostAllocWriteCombined.  flags & 2 (collective, needs a communicator): the ranks start together, so that the
 * call measures the box's CONCURRENT ceiling -- the bound of the end-to-end event egress with 8 GPUs. */
int marxb200_probe_d2h (marxb200_ctx *ctx, uint64_t bytes, int reps, int flags, double *gb_per_s);

/* FP64 roofline denominator measured on this GPU: best of 5 runs of a DFMA-chain kernel (8 independent chains per
 * thread, 8 x 256-thread CTAs per SM), in TFLOP/s counting an FMA as 2 flops.  Diagnostic; leaves the photon list alone. */
int marxb200_measure_fp64_peak (marxb200_ctx *ctx, double *tflops);

/* kernel launch counter (bench.py "gpu_launches") */
int marxb200_get_launch_count (marxb200_ctx *ctx, uint64_t *n);

/* ------------------------------------------------------------------------------------------------ */
/* Work the library does AHEAD of the call that asks for it (results never depend on it):
 *  - marxb200_trace / marxb200_trace_sharded launch the arrival-time pre-pass (and, sharded, the all-gather of its sums) of the
 *    next contiguous batch -- the same number of rays right behind the batch just traced -- on a stream of their own.  A call that
 *    asks for other rays, or follows marxb200_set_source / marxb200_load_calpack, runs the pre-pass itself.
 *  - after marxb200_egress_begin_packed or marxb200_merge_events_begin, the order restoration of the next traced batch also writes
 *    the file images of the same columns (same write mask, time offset, row limit); a begin call with exactly those arguments then
 *    launches no conversion kernel, any other one converts as usual.  Costs one more staging buffer of max_out rows.
 *
 * Environment switches (read when a context / communicator is created, or at the call; all default to the fast path):
 *   MARXB200_VERBOSE=1            grids, merge transport, look-ahead statistics on stderr
 *   MARXB200_LOOKAHEAD=0          no pre-pass look-ahead              MARXB200_PREPACK=0       conversion kernel for every batch
 *   MARXB200_K01_TICKET=0         static tile stride in the fused source + HRMA-A kernel
 *   MARXB200_K1_SPLIT=0 / _K2_SPLIT=0 / _K3_SPLIT=0   the stage as fewer, larger kernels (A/B runs; identical results)
 *   MARXB200_PILEUP_FUSED=0       pile-up by the step kernels only    MARXB200_PILEUP_WINDOW=256|512|1024  first window size tried
 *   MARXB200_NCCL_LIB=path        the NCCL library to dlopen (default libnccl.so.2)
 *   MARXB200_MERGE_TRANSPORT=nccl ncclSend / ncclRecv instead of CUDA-IPC peer writes      MARXB200_NCCL_CTAS=0  NCCL's own CTA count */

#ifdef __cplusplus
}
#endif
#endif
