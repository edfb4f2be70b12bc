// mx_hrc.cuh -- K3 (HRC-S): HESF "Drake flat" reflection, MCP plane intersection, MCP QE, UV/ion-shield
// transmission, position blur, pixel mapping, pulse height.
// Reference: marx/libsrc/hrc-s.c:136-312 (get_filter_region, apply_hrc_qe, _marx_hrc_s_detect); drake.c:268-372;
// hrcblur.c:236-298; hrc_s_geom.c:344-394 (_marx_hrc_s_compute_pixel); hrc-i.c:66-85 (_marx_hrc_compute_pha);
// reflect.c:81-92 (marx_interp_reflectivity).
// Draw order on sub-stream MARXB200_STAGE_DETECTOR: HESF U (only if a plate is hit); MCP QE U; shield U; PHA G;
// blur: component U, radius U (retry on 0 for the gaussians), angle U.
#pragma once
#include "mx_common.cuh"
#include "mx_tables.h"
#include "mx_hrma.cuh"      // reflectivity()
#include "mx_acis.cuh"      // m3_mul_t

namespace mx {

// drake_intersection + the reflection test of _marx_drake_reflect, drake.c:268-372.
// returns 0: no plate hit, 1: reflected, -1: absorbed
MX_HD int hesf_reflect (const HrcDev &D, double energy, Vec3 &x, Vec3 &p, Rng &rng)
{
   const int imax = 2 * D.hesf_num_plates;
   for (int i = 0; i < imax; i++)
     {
        const HesfPlateDev &r = D.hesf[i];
        const Vec3 normal = v_make (r.normal[0], r.normal[1], r.normal[2]);
        const Vec3 a = v_make (r.a[0], r.a[1], r.a[2]);
        double p_dot_n = v_dot (p, normal);
        if (0.0 == p_dot_n) continue;
        double t = v_dot (v_diff (a, x), normal) / p_dot_n;
        Vec3 new_x = v_ax1_bx2 (1.0, x, t, p);
        Vec3 x_prime = v_diff (new_x, a);
        double xx = v_dot (x_prime, v_make (r.e1[0], r.e1[1], r.e1[2]));
        if ((xx < 0.0) || (xx >= r.len1)) continue;
        double yy = v_dot (x_prime, v_make (r.e2[0], r.e2[1], r.e2[2]));
        if ((yy < 0.0) || (yy >= r.len2)) continue;
        const bool use_cr = (xx < D.hesf_cr_width);
        x = new_x;
        double rfl = 1.0;
        const uint32_t n = use_cr ? D.cr_num : D.c_num;
        if (n != 0)
          {
             const float *e = use_cr ? D.cr_energies : D.c_energies;
             float beta, delta;
             interp_f2 ((float) energy, e, use_cr ? D.cr_betas : D.c_betas, use_cr ? D.cr_deltas : D.c_deltas, n, beta, delta);
             rfl = reflectivity (fabs (p_dot_n), beta, delta);
          }
        if (rfl < rng.uniform ()) return -1;
        p = v_ax1_bx2 (1.0, p, -2.0 * p_dot_n, normal);
        return 1;
     }
   return 0;
}

// get_filter_region, hrc-s.c:136-188
MX_HD int hrc_filter_region (const HrcDev &D, double y, double z)
{
   y -= D.shield_y_center;
   z -= D.shield_z_center;
   double off, sl, slg;
   if (y < 0) { y = -y; off = D.shield_l; sl = D.shield_sl; slg = D.shield_sl_gap; }
   else { off = D.shield_r; sl = D.shield_sr; slg = D.shield_sr_gap; }
   if (y < off) return 0;
   if (y < sl) return (z >= D.shield_t) ? 0 : 1;
   if (y >= slg) return (z >= D.shield_t) ? 2 : 3;
   return -1;
}

// _marx_hrc_compute_pha, hrc-i.c:66-85
MX_HD int16_t hrc_pha (double energy, Rng &rng)
{
   if (energy <= 0.5) energy = 141.582 * sqrt (energy);
   else if (energy < 2.0) energy = 107.299 * pow (energy, 0.1);
   else energy = 115.0;
   energy = energy * (1.0 + 0.424661 * rng.gaussian ());
   if (energy < 0.0) energy = 0.0;
   return (int16_t) energy;
}

// _marx_hrc_blur_position, hrcblur.c:260-298.  blur = {g1 sigma,xctr,yctr,wgt, g2 sigma,xctr,yctr,wgt, l1 hwhm,xctr,yctr,rmax,wgt}
MX_HD void hrc_blur (const HrcDev &D, double &dx, double &dy, Rng &rng)
{
   if (D.det_ideal) return;
   const double *b = D.blur;
   double r = rng.uniform (), x_0, y_0;
   if ((r < b[3]) || (r < b[3] + b[7]))
     {
        const int o = (r < b[3]) ? 0 : 4;
        double c;
        do c = rng.uniform (); while (c == 0.0);
        r = b[o] * sqrt (-2 * log (c));
        x_0 = b[o + 1]; y_0 = b[o + 2];
     }
   else
     {
        double c = rng.uniform ();
        double rmax = b[11] / b[8];
        r = b[8] * sqrt (expm1 (c * log1p (rmax * rmax)));
        x_0 = b[9]; y_0 = b[10];
     }
   double theta = (2.0 * kPI) * rng.uniform ();
   double st, ct;
   sin_cos (theta, st, ct);
   dx += x_0 + r * ct;
   dy += y_0 + r * st;
   if (D.det_extend == 0)
     {
        if (dx < 0.0) dx = 0.0;
        if (dy < 0.0) dy = 0.0;
     }
}

// _marx_hrc_s_detect for one ray (preceded by _marx_drake_reflect when HRC-HESF=yes).  Returns flags.
template <bool DET = false>
MX_HD uint32_t hrc_s_detect (const HrcDev &D, double energy, Vec3 &x, Vec3 &p, int &ccd, int &region,
                             float &ypix, float &zpix, float &upix, float &vpix, int16_t &pha, Rng &rng,
                             const DetDither &dd = DetDither {0.0, 0.0, 0.0})
{
   const uint32_t UNDETECTED = 0x01, MISSED = 0x08, DRAKE_BLOCKED = 0x20, DRAKE_REFLECTED = 0x100;
   uint32_t flags = 0;
   if (D.use_hesf)
     {
        int h = hesf_reflect (D, energy, x, p, rng);
        if (h < 0) return DRAKE_BLOCKED;
        if (h > 0) flags |= DRAKE_REFLECTED;
     }
   // _marx_dither_detector + _marx_transform_ray (hrc-s.c:264-270, detector.c:275-284): see det_dither_frame
   const double *det_off = D.det_offset, *det_mat = D.det_matrix;
   double off_l[3], mat_l[9];
   if (DET && ((dd.dy != 0) || (dd.dz != 0) || (dd.dtheta != 0)))
     {
        det_dither_frame (D.det_offset, D.det_matrix, dd, off_l, mat_l);
        det_off = off_l; det_mat = mat_l;
     }
   x.x -= det_off[0]; x.y -= det_off[1]; x.z -= det_off[2];
   x = m3_mul (det_mat, x);
   p = m3_mul (det_mat, p);

   double dx = 0, dy = 0;
   Vec3 xh = x;
   const int hit = detector_intersect (D.mcp, D.num_mcps, x, p, xh, dx, dy, D.det_extend);   // detector.c:111-168
   if (hit < 0) { ccd = -1; return flags | MISSED; }
   const HrcMcpDev &d = D.mcp[hit];
   x = xh;
   // apply_hrc_qe, hrc-s.c:192-234
   const float ef = (float) energy;
   if (d.qe_num != 0)
     {
        double qe = interp_f (ef, d.qe_energies, d.qe, d.qe_num);
        if (rng.uniform () >= qe) { ccd = -1; return flags | UNDETECTED; }
     }
   const bool hrc_i = (D.detector_type == 2);      // MARX_DETECTOR_HRC_I: one MCP, one UVIS filter, no shield regions
   if (hrc_i)
     {
        // apply_hrc_qe, hrc-i.c:88-117: Filter_QEs[mcp_id] with mcp_id == 0
        region = 0;
        if (D.filter_num[0] != 0)
          {
             double qe = interp_f (ef, D.filter_energies[0], D.filter_qe[0], D.filter_num[0]);
             if (rng.uniform () >= qe) { ccd = -1; return flags | UNDETECTED; }
          }
     }
   else
     {
        double t = (D.shield_x - x.x) / p.x;
        double y = x.y + t * p.y, z = x.z + t * p.z;
        region = hrc_filter_region (D, y, z);
        if (region < 0) { ccd = -1; return flags | UNDETECTED; }
        if (D.filter_num[region] != 0)
          {
             double qe = interp_f (ef, D.filter_energies[region], D.filter_qe[region], D.filter_num[region]);
             if (rng.uniform () >= qe) { ccd = -1; return flags | UNDETECTED; }
          }
     }
   pha = hrc_pha (energy, rng);
   hrc_blur (D, dx, dy, rng);
   ccd = d.id;
   if (hrc_i)
     {
        // _marx_hrc_i_compute_pixel, hrc_i_geom.c:146-156: LL_CXCY + d / pixel size (u_start, v_start hold LL_CXCY)
        ypix = (float) (d.u_start + dx / D.u_pixel_size);
        zpix = (float) (d.v_start + dy / D.v_pixel_size);
        upix = 0.f; vpix = 0.f;
        p = m3_mul_t (det_mat, p);
        x = m3_mul_t (det_mat, x);
        x.x += det_off[0]; x.y += det_off[1]; x.z += det_off[2];
        return flags;
     }
   // _marx_hrc_s_compute_pixel, hrc_s_geom.c:344-394
   double u = d.u_start + dx / D.u_pixel_size;
   double v = d.v_start + dy / D.v_pixel_size;
   upix = (float) u; vpix = (float) v;
   ypix = (float) (d.cx_0 + (u - d.u_0));
   zpix = (float) (d.cy_0 + (v - d.v_0));
   p = m3_mul_t (det_mat, p);
   x = m3_mul_t (det_mat, x);
   x.x += det_off[0]; x.y += det_off[1]; x.z += det_off[2];
   return flags;
}

}  // namespace mx
