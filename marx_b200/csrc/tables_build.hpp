// tables_build.hpp -- turns the host descriptors of include/marxb200.h into the table images the
// kernels read (mx_tables.h / mx_kernels.cuh blobs).  Pure host C++, templated on an "uploader"
// (bytes -> pointer valid where the kernels run): marxb200.cu passes cudaMalloc+cudaMemcpy; the
// developer-only tools/hostcheck passes malloc+memcpy to step the same table code without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/marxb200.h"
#include "mx_tables.h"
#include "mx_kernels.cuh"

namespace mx {

inline size_t tb_align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

template <class Up, class T> inline int tb_up (Up &up, const T *host, size_t n, const T **out, std::string &err)
{
   const void *d = up (host, n * sizeof (T));
   if (d == nullptr) { err = "table upload failed"; return -1; }
   *out = (const T *) d;
   return 0;
}

template <class Up> int build_wfold (Up &up, const marxb200_wfold_table *w, WfoldDev *out, std::string &err)
{
   memset (out, 0, sizeof (*out));
   out->num_arrays = w->num_arrays;
   if (w->num_arrays == 0) return 0;
   std::vector<double> hdr (6 * (size_t) w->num_arrays);
   for (uint32_t i = 0; i < w->num_arrays; i++)
     {
        hdr[6 * i + 0] = w->e_alpha[i]; hdr[6 * i + 1] = w->p_min[i]; hdr[6 * i + 2] = w->delta_p[i];
        hdr[6 * i + 3] = w->p_max[i]; hdr[6 * i + 4] = w->pow_law_norm[i]; hdr[6 * i + 5] = w->pow_law_expon[i];
     }
   out->min_p_min = w->p_min[0];
   for (uint32_t i = 1; i < w->num_arrays; i++) if (w->p_min[i] < out->min_p_min) out->min_p_min = w->p_min[i];
   if (-1 == tb_up (up, hdr.data (), hdr.size (), &out->hdr, err)) return -1;
   if (-1 == tb_up (up, w->num_theta, w->num_arrays, &out->num_theta, err)) return -1;
   if (-1 == tb_up (up, w->theta_offset, w->num_arrays, &out->theta_offset, err)) return -1;
   if (-1 == tb_up (up, w->theta_values, w->total_theta, &out->theta, err)) return -1;
   return 0;
}

template <class Up> int build_hrma_blob (Up &up, const marxb200_hrma_desc *d, std::vector<unsigned char> &blob, std::string &err)
{
   size_t ncorr_total = 0;
   for (int k = 0; k < kNumShells; k++) ncorr_total += d->shells[k].num_corr;
   size_t off = tb_align16 (sizeof (K1Blob));
   const size_t off_opt_e = off; off = tb_align16 (off + 4 * (size_t) d->num_opt);
   const size_t off_opt_b = off; off = tb_align16 (off + 4 * (size_t) d->num_opt);
   const size_t off_opt_d = off; off = tb_align16 (off + 4 * (size_t) d->num_opt);
   const size_t off_corr_e = off; off = tb_align16 (off + 4 * ncorr_total);
   const size_t off_corr_f = off; off = tb_align16 (off + 4 * ncorr_total);
   uint32_t max_arrays = 0;
   if (d->use_wfold)
     for (int k = 0; k < kNumShells; k++)
       {
          if (d->shells[k].p_wfold.num_arrays > max_arrays) max_arrays = d->shells[k].p_wfold.num_arrays;
          if (d->shells[k].h_wfold.num_arrays > max_arrays) max_arrays = d->shells[k].h_wfold.num_arrays;
       }
   const size_t wkeys_stride = tb_align16 (8 * (size_t) max_arrays), wkeys_bytes = kNumShells * wkeys_stride;
   const size_t off_wkeys_p = off; off += wkeys_bytes;
   const size_t off_wkeys_h = off; off += wkeys_bytes;
   const size_t total = off;
   if (total > 200 * 1024) { err = "HRMA tables do not fit in shared memory"; return -1; }
   blob.assign (total, 0);
   K1Blob *B = reinterpret_cast<K1Blob *> (blob.data ());
   HrmaDev &H = B->H;
   H.vig = d->vignetting_factor; H.cap_position = d->cap_position;
   H.is_ideal = d->is_ideal; H.use_blur = d->use_blur; H.use_wfold = d->use_wfold;
   H.use_struts = d->use_struts; H.use_scale = d->use_scale_factors;
   H.num_opt = d->is_ideal ? 0 : d->num_opt;
   H.num_corr_total = (uint32_t) ncorr_total;
   uint32_t corr_off = 0;
   float *ce = reinterpret_cast<float *> (blob.data () + off_corr_e);
   float *cf = reinterpret_cast<float *> (blob.data () + off_corr_f);
   for (int k = 0; k < kNumShells; k++)
     {
        const marxb200_hrma_shell &s = d->shells[k];
        HrmaShellDev &h = H.shell[k];
        h.conic_p[0] = s.conic_a_p; h.conic_p[1] = s.conic_b_p; h.conic_p[2] = s.conic_c_p; h.conic_p[3] = s.conic_xmin_p; h.conic_p[4] = s.conic_xmax_p;
        h.conic_h[0] = s.conic_a_h; h.conic_h[1] = s.conic_b_h; h.conic_h[2] = s.conic_c_h; h.conic_h[3] = s.conic_xmin_h; h.conic_h[4] = s.conic_xmax_h;
        for (int i = 0; i < 3; i++) { h.to_osac_p[i] = s.to_osac_p[i]; h.to_osac_h[i] = s.to_osac_h[i]; }
        h.front_position = s.front_position; h.area_fraction = s.area_fraction;
        h.min_radius = s.min_radius; h.max_radius = s.max_radius;
        h.p_blur = s.p_blur; h.h_blur = s.h_blur; h.p_scat = s.p_scat_factor; h.h_scat = s.h_scat_factor;
        for (int i = 0; i < 9; i++)
          { h.fwd_p[i] = s.fwd_matrix_p[i]; h.bwd_p[i] = s.bwd_matrix_p[i]; h.fwd_h[i] = s.fwd_matrix_h[i]; h.bwd_h[i] = s.bwd_matrix_h[i]; }
        h.shutter_bitmap = s.shutter_bitmap;
        h.num_corr = s.num_corr; h.corr_offset = corr_off;
        if (s.num_corr)
          {
             memcpy (ce + corr_off, s.corr_energies, 4 * (size_t) s.num_corr);
             memcpy (cf + corr_off, s.corr_factors, 4 * (size_t) s.num_corr);
          }
        corr_off += s.num_corr;
        if (d->use_wfold)
          {
             if (-1 == build_wfold (up, &s.p_wfold, &h.wfold_p, err)) return -1;
             if (-1 == build_wfold (up, &s.h_wfold, &h.wfold_h, err)) return -1;
          }
     }
   if (d->num_opt)
     {
        memcpy (blob.data () + off_opt_e, d->opt_energies, 4 * (size_t) d->num_opt);
        memcpy (blob.data () + off_opt_b, d->opt_betas, 4 * (size_t) d->num_opt);
        memcpy (blob.data () + off_opt_d, d->opt_deltas, 4 * (size_t) d->num_opt);
     }
   if (!d->use_scale_factors || (ncorr_total == 0)) H.use_scale = 0;
   B->off_opt_e = (uint32_t) off_opt_e; B->off_opt_b = (uint32_t) off_opt_b; B->off_opt_d = (uint32_t) off_opt_d;
   B->off_corr_e = (uint32_t) off_corr_e; B->off_corr_f = (uint32_t) off_corr_f; B->total_bytes = (uint32_t) total;
   B->off_wkeys_p = (uint32_t) off_wkeys_p; B->off_wkeys_h = (uint32_t) off_wkeys_h;
   B->wkeys_stride = (uint32_t) wkeys_stride; B->wkeys_bytes = (uint32_t) wkeys_bytes;
   if (d->use_wfold)
     for (int k = 0; k < kNumShells; k++)
       {
          const marxb200_hrma_shell &s = d->shells[k];
          if (s.p_wfold.num_arrays) memcpy (blob.data () + off_wkeys_p + k * wkeys_stride, s.p_wfold.e_alpha, 8 * (size_t) s.p_wfold.num_arrays);
          if (s.h_wfold.num_arrays) memcpy (blob.data () + off_wkeys_h + k * wkeys_stride, s.h_wfold.e_alpha, 8 * (size_t) s.h_wfold.num_arrays);
       }
   return 0;
}

template <class Up> int build_grating_shell (Up &up, const marxb200_grating_shell &s, GratingShellDev &g, double rowland, std::string &err)
{
   if ((s.num_orders == 0) || (s.num_energies < 2)) { err = "a grating has an empty efficiency table"; return -1; }
   g.num_orders = s.num_orders; g.num_energies = s.num_energies; g.num_sectors = s.num_sectors;
   if (-1 == tb_up (up, s.order_list, s.num_orders, &g.order_list, err)) return -1;
   if (-1 == tb_up (up, s.energies, s.num_energies, &g.energies, err)) return -1;
   // transpose [order][energy] -> [energy][order] (mx_grating.cuh diffract_from_grating)
   std::vector<float> t ((size_t) s.num_orders * s.num_energies);
   for (uint32_t o = 0; o < s.num_orders; o++)
     for (uint32_t e = 0; e < s.num_energies; e++)
       t[(size_t) e * s.num_orders + o] = s.cum_eff[(size_t) o * s.num_energies + e];
   if (-1 == tb_up (up, t.data (), t.size (), &g.cum_eff, err)) return -1;
   g.dispersion_angle = s.dispersion_angle; g.period = s.period; g.dp_over_p = s.dp_over_p;
   g.cos_dispersion = cos (s.dispersion_angle); g.sin_dispersion = sin (s.dispersion_angle);
   g.theta_blur = s.theta_blur; g.vig = s.vig; g.rowland = rowland;
   g.sectors = nullptr;
   return 0;
}

template <class Up> int build_grating_blob (Up &up, const marxb200_grating_desc *d, std::vector<unsigned char> &blob, std::string &err)
{
   if ((d->type != 1) && (d->type != 2)) { err = "only HETG (1) and LETG (2) are implemented"; return -1; }
   size_t off = tb_align16 (sizeof (K2Blob));
   size_t off_sec[kNumShells];
   for (int k = 0; k < kNumShells; k++) { off_sec[k] = off; off = tb_align16 (off + 6 * 8 * (size_t) d->shells[k].num_sectors); }
   const size_t total = off;
   blob.assign (total, 0);
   K2Blob *B = reinterpret_cast<K2Blob *> (blob.data ());
   B->G.type = d->type;
   for (int k = 0; k < kNumShells; k++)
     {
        const marxb200_grating_shell &s = d->shells[k];
        GratingShellDev &g = B->G.shell[k];
        if (-1 == build_grating_shell (up, s, g, d->rowland[k], err)) return -1;
        double *sec = reinterpret_cast<double *> (blob.data () + off_sec[k]);
        for (uint32_t i = 0; i < s.num_sectors; i++)
          {
             sec[0 * s.num_sectors + i] = s.sec_min_angle[i]; sec[1 * s.num_sectors + i] = s.sec_max_angle[i];
             sec[2 * s.num_sectors + i] = s.sec_dtheta[i]; sec[3 * s.num_sectors + i] = s.sec_dtheta_blur[i];
             sec[4 * s.num_sectors + i] = s.sec_dpp[i]; sec[5 * s.num_sectors + i] = s.sec_dpp_blur[i];
          }
        B->off_sectors[k] = (uint32_t) off_sec[k];
     }
   for (int k = 0; k < 2; k++)
     {
        memset (&B->G.support[k], 0, sizeof (GratingShellDev));
        if ((d->type == 2) && (d->support[k].num_orders != 0))
          {
             if (-1 == build_grating_shell (up, d->support[k], B->G.support[k], 0.0, err)) return -1;
             B->G.support[k].num_sectors = 0;
          }
     }
   B->total_bytes = (uint32_t) total;
   return 0;
}

template <class Up> int build_hrc_blob (Up &up, const marxb200_hrc_s_desc *d, std::vector<unsigned char> &blob, std::string &err)
{
   if ((d->num_mcps < 1) || (d->num_mcps > 3)) { err = "bad MCP count"; return -1; }
   if ((d->hesf_num_plates < 0) || (d->hesf_num_plates > 4)) { err = "bad HESF plate count"; return -1; }
   const size_t total = tb_align16 (sizeof (K3HrcBlob));
   blob.assign (total, 0);
   K3HrcBlob *B = reinterpret_cast<K3HrcBlob *> (blob.data ());
   HrcDev &D = B->D;
   D.detector_type = d->detector_type; D.num_mcps = d->num_mcps;
   for (int k = 0; k < d->num_mcps; k++)
     {
        const marxb200_hrc_mcp &s = d->mcps[k];
        HrcMcpDev &g = D.mcp[k];
        g.id = s.id; g.qe_num = s.qe_num;
        for (int i = 0; i < 3; i++) { g.x_ll[i] = s.x_ll[i]; g.xhat[i] = s.xhat[i]; g.yhat[i] = s.yhat[i]; g.normal[i] = s.normal[i]; }
        g.xlen = s.xlen; g.ylen = s.ylen;
        if (s.qe_num && ((-1 == tb_up (up, s.qe_energies, s.qe_num, &g.qe_energies, err)) || (-1 == tb_up (up, s.qe, s.qe_num, &g.qe, err)))) return -1;
        g.u_start = s.u_start; g.v_start = s.v_start; g.u_0 = s.u_0; g.v_0 = s.v_0; g.cx_0 = s.cx_0; g.cy_0 = s.cy_0;
     }
   for (int r = 0; r < 4; r++)
     {
        D.filter_num[r] = d->filter_num[r];
        if (d->filter_num[r] && ((-1 == tb_up (up, d->filter_energies[r], d->filter_num[r], &D.filter_energies[r], err))
                                 || (-1 == tb_up (up, d->filter_qe[r], d->filter_num[r], &D.filter_qe[r], err)))) return -1;
     }
   D.shield_t = d->shield_t; D.shield_l = d->shield_l; D.shield_r = d->shield_r; D.shield_x = d->shield_x;
   D.shield_sl = d->shield_sl; D.shield_sr = d->shield_sr; D.shield_sl_gap = d->shield_sl_gap; D.shield_sr_gap = d->shield_sr_gap;
   D.shield_y_center = d->shield_y_center; D.shield_z_center = d->shield_z_center;
   for (int i = 0; i < 13; i++) D.blur[i] = d->blur[i];
   D.u_pixel_size = d->u_pixel_size; D.v_pixel_size = d->v_pixel_size;
   for (int i = 0; i < 3; i++) D.det_offset[i] = d->det_offset[i];
   for (int i = 0; i < 9; i++) D.det_matrix[i] = d->det_matrix[i];
   D.det_ideal = d->det_ideal; D.det_extend = d->det_extend;
   D.use_hesf = d->use_hesf; D.hesf_num_plates = d->hesf_num_plates; D.hesf_cr_width = d->hesf_cr_width;
   for (int k = 0; k < 2 * d->hesf_num_plates; k++)
     {
        const marxb200_hesf_plate &s = d->hesf[k];
        HesfPlateDev &h = D.hesf[k];
        for (int i = 0; i < 3; i++) { h.a[i] = s.a[i]; h.e1[i] = s.e1[i]; h.e2[i] = s.e2[i]; h.normal[i] = s.normal[i]; }
        h.len1 = s.len1; h.len2 = s.len2;
     }
   D.c_num = d->c_num; D.cr_num = d->cr_num;
   if (d->c_num && ((-1 == tb_up (up, d->c_energies, d->c_num, &D.c_energies, err)) || (-1 == tb_up (up, d->c_betas, d->c_num, &D.c_betas, err))
                    || (-1 == tb_up (up, d->c_deltas, d->c_num, &D.c_deltas, err)))) return -1;
   if (d->cr_num && ((-1 == tb_up (up, d->cr_energies, d->cr_num, &D.cr_energies, err)) || (-1 == tb_up (up, d->cr_betas, d->cr_num, &D.cr_betas, err))
                     || (-1 == tb_up (up, d->cr_deltas, d->cr_num, &D.cr_deltas, err)))) return -1;
   B->total_bytes = (uint32_t) total;
   return 0;
}

template <class Up> int build_acis_blob (Up &up, const marxb200_acis_desc *d, std::vector<unsigned char> &blob, std::string &err)
{
   if ((d->num_chips < 1) || (d->num_chips > kMaxChips)) { err = "bad chip count"; return -1; }
   const size_t total = tb_align16 (sizeof (K3Blob));
   blob.assign (total, 0);
   K3Blob *B = reinterpret_cast<K3Blob *> (blob.data ());
   AcisDev &A = B->A;
   A.detector_type = d->detector_type; A.num_chips = d->num_chips;
   std::vector<FefDev> fefs (d->num_fefs);
   for (uint32_t i = 0; i < d->num_fefs; i++)
     {
        const marxb200_fef &f = d->fefs[i];
        if (f.num_gaussians > (uint32_t) kMaxGauss) { err = "FEF with too many gaussians"; return -1; }
        if (f.num_energies < 2) { err = "FEF region with < 2 energies"; return -1; }
        fefs[i].num_gaussians = f.num_gaussians; fefs[i].num_energies = f.num_energies;
        if (-1 == tb_up (up, f.energies, f.num_energies, &fefs[i].energies, err)) return -1;
        if (-1 == tb_up (up, f.channels, f.num_energies, &fefs[i].channels, err)) return -1;
        if (-1 == tb_up (up, f.gauss, 3 * (size_t) f.num_energies * f.num_gaussians, &fefs[i].gauss, err)) return -1;
     }
   if (-1 == tb_up (up, fefs.data (), fefs.size (), &A.fefs, err)) return -1;
   A.num_fefs = d->num_fefs;
   for (int k = 0; k < d->num_chips; k++)
     {
        const marxb200_acis_chip &s = d->chips[k];
        AcisChipDev &g = A.chip[k];
        g.id = s.id;
        for (int i = 0; i < 3; i++) { g.x_ll[i] = s.x_ll[i]; g.xhat[i] = s.xhat[i]; g.yhat[i] = s.yhat[i]; g.normal[i] = s.normal[i]; }
        g.xlen = s.xlen; g.ylen = s.ylen; g.x_pixel_size = s.x_pixel_size; g.y_pixel_size = s.y_pixel_size;
        g.xpixel_offset = s.xpixel_offset; g.ypixel_offset = s.ypixel_offset;
        g.qe_num = s.qe_num; g.filter_num = s.filter_num;
        if (s.qe_num && ((-1 == tb_up (up, s.qe_energies, s.qe_num, &g.qe_energies, err)) || (-1 == tb_up (up, s.qe, s.qe_num, &g.qe, err)))) return -1;
        if (s.filter_num && ((-1 == tb_up (up, s.filter_energies, s.filter_num, &g.filter_energies, err)) || (-1 == tb_up (up, s.filter_qe, s.filter_num, &g.filter_qe, err)))) return -1;
        if (s.contam_num_layers > (uint32_t) kMaxContamLayers) { err = "too many contamination layers"; return -1; }
        g.contam_num_layers = s.contam_num_layers; g.contam_fxy_mode = s.contam_fxy_mode;
        g.contam_x0 = s.contam_x0; g.contam_y0 = s.contam_y0; g.contam_blocking = s.contam_blocking;
        for (uint32_t l = 0; l < s.contam_num_layers; l++)
          {
             g.contam_tau0[l] = s.contam_tau0[l]; g.contam_tau1[l] = s.contam_tau1[l]; g.contam_num_mu[l] = s.contam_num_mu[l];
             if (-1 == tb_up (up, s.contam_energies[l], s.contam_num_mu[l], &g.contam_energies[l], err)) return -1;
             if (-1 == tb_up (up, s.contam_mus[l], s.contam_num_mu[l], &g.contam_mus[l], err)) return -1;
             if (s.contam_fxy_mode == 0)
               {
                  if ((s.contam_blocking == 0) || (1024 % s.contam_blocking)) { err = "bad FXYBLK"; return -1; }
                  size_t nb = 1024 / s.contam_blocking;
                  if (-1 == tb_up (up, s.contam_fxy[l], nb * nb, &g.contam_fxy[l], err)) return -1;
               }
          }
        if (s.fef_map == nullptr) { err = "chip without FEF map"; return -1; }
        for (int i = 0; i < 1024; i++)
          if ((s.fef_map[i] < -1) || (s.fef_map[i] >= (int32_t) d->num_fefs)) { err = "FEF map entry outside [-1, num_fefs)"; return -1; }
        if (-1 == tb_up (up, s.fef_map, 1024, &g.fef_map, err)) return -1;
     }
   for (int i = 0; i < 3; i++) A.det_offset[i] = d->det_offset[i];
   for (int i = 0; i < 9; i++) A.det_matrix[i] = d->det_matrix[i];
   A.det_ideal = d->det_ideal; A.det_extend = d->det_extend; A.dither_mode = d->dither_mode;
   A.focal_length = d->focal_length; A.exposure_time = d->exposure_time;
   A.frame_transfer_time = d->frame_transfer_time; A.frame_time = d->frame_time;
   B->total_bytes = (uint32_t) total;
   return 0;
}

}  // namespace mx
