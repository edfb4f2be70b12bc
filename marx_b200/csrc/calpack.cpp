// calpack.cpp -- marxb200_load_calpack: read a calibration pack (include/marxb200_calpack.h) and call
// the marxb200_set_* setters with host pointers into it.  Entry names / packed orders are those
// written by oracle/ref/calpack_*.c (the file-writing twin of the upload calls shown in INTEGRATION.md):
//
//   meta                      f64[8]  mirror, grating, detector module ids, tstart yrs, tstart secs, seed, numrays, exposure
//   source.params             f64[16] source_type, spectrum_type, p[3], p_normal[3], distance, emin, emax, total_flux, geometric_area, shape[3]
//   source.spec_energies/.spec_cum_flux  f64[n]   (FILE spectrum only)
//   source.rotation           f64[4]  LINE / IMAGE: axis[3] + angle of the rotation taking (-1,0,0) to p
//   source.image_params       f64[4]  IMAGE: nx, ny, rad per x pixel, rad per y pixel;  source.image_cdf f32[ny*nx]
//   dither.params             f64[12] mode, amp ra/dec/roll, period ra/dec/roll, phase ra/dec/roll, nominal_roll, aspect_blur
//   dither.aspsol             f64[n][7] (mode 2) t, ra, dec, roll, dy, dz, dtheta of every ASPSOL reader state
//   ffield.params             f64[5]  MirrorType=FLATFIELD instead of the hrma.* entries: min_y, min_z, max_y, max_z, x_pos
//   hrma.params               f64[7]  vig, cap_position, is_ideal, use_blur, use_wfold, use_struts, use_scale_factors
//   hrma.opt_energies/.opt_betas/.opt_deltas  f32[n]
//   hrma.shell<k>.params      f64[62] (order: see set_hrma_from_pack)
//   hrma.shell<k>.corr_energies/.corr_factors  f32[m]
//   hrma.shell<k>.wfold_{p,h}.hdr f64[n][6], .num_theta u32[n], .theta f32[sum]
//   grating.params            f64[5]  type, rowland[4]
//   grating.shell<k>.params   f64[5]  dispersion_angle, period, dp_over_p, theta_blur, vig
//   grating.shell<k>.order_list i32, .energies f32, .cum_eff f32[orders][energies], .sectors f64[n][6]
//   grating.support<k>.*      LETG fine (0) / coarse (1) support gratings, same entries as a shell without sectors
//   hrc.params f64[44]; hrc.mcp<k>.geom f64[21], .qe_energies, .qe; hrc.filter<r>.energies, .qe; hrc.hesf f64[2n][14];
//   hrc.hesf_c_{energies,betas,deltas}, hrc.hesf_cr_{...}        (HRC-S packs carry no acis.* entries)
//   acis.params               f64[21] detector_type, num_chips, det_offset[3], det_matrix[9], det_ideal, det_extend,
//                                     focal_length, exposure_time, frame_transfer_time, frame_time, dither_mode
//   acis.num_fefs u32[1]; acis.fef<j>.dims u32[2] (num_gaussians, num_energies), .energies, .channels, .gauss f32
//   acis.chip<k>.geom f64[19]; .qe_energies/.qe/.filter_energies/.filter_qe f32; .contam f64[15];
//   .contam_e<l>/.contam_mu<l>/.contam_fxy<l> f32; .fef_map i32[1024]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/marxb200.h"
#include "../../include/marxb200_calpack.h"

namespace {

struct Entry { uint32_t dtype; uint64_t count; const unsigned char *data; };
struct Pack
{
   std::vector<unsigned char> bytes;
   std::map<std::string, Entry> entries;
   std::string err;

   bool load (const char *path)
   {
      FILE *fp = fopen (path, "rb");
      if (!fp) { err = std::string ("cannot open ") + path; return false; }
      fseek (fp, 0, SEEK_END); long sz = ftell (fp); fseek (fp, 0, SEEK_SET);
      if (sz < 16) { fclose (fp); err = (sz < 0) ? "cannot determine the file size" : "file too short"; return false; }
      bytes.resize ((size_t) sz);
      if (fread (bytes.data (), 1, (size_t) sz, fp) != (size_t) sz) { fclose (fp); err = "short read"; return false; }
      fclose (fp);
      if (memcmp (bytes.data (), MARXB200_CALPACK_MAGIC, 8)) { err = "bad magic"; return false; }
      uint32_t n; memcpy (&n, bytes.data () + 8, 4);
      size_t off = 16;
      for (uint32_t i = 0; i < n; i++)
        {
           if (off + MARXB200_CALPACK_NAMELEN + 16 > bytes.size ()) { err = "truncated"; return false; }
           char name[MARXB200_CALPACK_NAMELEN + 1]; memcpy (name, bytes.data () + off, MARXB200_CALPACK_NAMELEN); name[MARXB200_CALPACK_NAMELEN] = 0;
           Entry e; memcpy (&e.dtype, bytes.data () + off + MARXB200_CALPACK_NAMELEN, 4);
           memcpy (&e.count, bytes.data () + off + MARXB200_CALPACK_NAMELEN + 8, 8);
           off += MARXB200_CALPACK_NAMELEN + 16;
           const size_t esz = mxcp_dtype_size (e.dtype);
           if ((esz == 0) || (e.count > (bytes.size () - off) / esz)) { err = "truncated data"; return false; }   // no count * size overflow
           size_t nb = (size_t) e.count * esz;
           e.data = bytes.data () + off;
           off += (nb + 7) & ~(size_t) 7;
           entries[name] = e;
        }
      return true;
   }
   const Entry *find (const std::string &name, uint32_t dtype, uint64_t min_count)
   {
      auto it = entries.find (name);
      if (it == entries.end ()) { err = "missing entry " + name; return nullptr; }
      if (it->second.dtype != dtype) { err = "wrong dtype for " + name; return nullptr; }
      if (it->second.count < min_count) { err = "entry too short: " + name; return nullptr; }
      return &it->second;
   }
   bool has (const std::string &name) { return entries.count (name) != 0; }
};

std::string nm (const char *fmt, int a, int b = 0)
{
   char buf[96];
   snprintf (buf, sizeof (buf), fmt, a, b);
   return buf;
}

#define GET(var, name, dt, minc) const Entry *var = P.find ((name), (dt), (minc)); if (!var) return -1

}  // namespace

extern "C" int marxb200_load_calpack_impl (marxb200_ctx *ctx, const char *path, char *errbuf, size_t errlen)
{
   Pack P;
   auto bail = [&] (const std::string &m) { snprintf (errbuf, errlen, "marxb200_load_calpack: %s", m.c_str ()); return -1; };
   if (!P.load (path)) return bail (P.err);
   struct Guard { Pack &P; char *b; size_t l; int &rc; ~Guard () { if (rc == -1 && !P.err.empty () && b[0] == 0) snprintf (b, l, "marxb200_load_calpack: %s", P.err.c_str ()); } };
   int rc = -1;
   errbuf[0] = 0;
   Guard guard{P, errbuf, errlen, rc};

   // ---- source ----
   {
      GET (e, "source.params", MXCP_F64, 13);
      const double *v = (const double *) e->data;
      marxb200_source_desc d; memset (&d, 0, sizeof (d));
      d.source_type = (int32_t) v[0]; d.spectrum_type = (int32_t) v[1];
      for (int i = 0; i < 3; i++) { d.p[i] = v[2 + i]; d.p_normal[i] = v[5 + i]; }
      d.distance = v[8]; d.emin = v[9]; d.emax = v[10]; d.total_flux = v[11]; d.geometric_area = v[12];
      if (e->count >= 16) for (int i = 0; i < 3; i++) d.shape[i] = v[13 + i];
      if (d.spectrum_type == 2)
        {
           GET (se, "source.spec_energies", MXCP_F64, 2);
           GET (sc, "source.spec_cum_flux", MXCP_F64, 2);
           d.spec_energies = (const double *) se->data; d.spec_cum_flux = (const double *) sc->data; d.spec_num = (uint32_t) se->count;
        }
      if (d.source_type >= 4)       // LINE, IMAGE: rotation taking (-1,0,0) to p
        {
           GET (r, "source.rotation", MXCP_F64, 4);
           const double *rv = (const double *) r->data;
           for (int i = 0; i < 3; i++) d.rot_axis[i] = rv[i];
           d.rot_angle = rv[3];
        }
      if (d.source_type == 5)
        {
           GET (ip, "source.image_params", MXCP_F64, 4);
           const double *iv = (const double *) ip->data;
           d.image_nx = (uint32_t) iv[0]; d.image_ny = (uint32_t) iv[1]; d.rad_per_xpixel = iv[2]; d.rad_per_ypixel = iv[3];
           GET (ic, "source.image_cdf", MXCP_F32, (uint64_t) d.image_nx * d.image_ny);
           d.image_cdf = (const float *) ic->data;
        }
      if (-1 == marxb200_set_source (ctx, &d)) return bail (marxb200_last_error ());
   }
   // ---- dither ----
   {
      GET (e, "dither.params", MXCP_F64, 12);
      const double *v = (const double *) e->data;
      marxb200_dither_desc d; memset (&d, 0, sizeof (d));
      d.mode = (int32_t) v[0];
      d.ra_amp = v[1]; d.dec_amp = v[2]; d.roll_amp = v[3];
      d.ra_period = v[4]; d.dec_period = v[5]; d.roll_period = v[6];
      d.ra_phase = v[7]; d.dec_phase = v[8]; d.roll_phase = v[9];
      d.nominal_roll = v[10]; d.aspect_blur = v[11];
      if (d.mode == 2)
        {
           GET (a, "dither.aspsol", MXCP_F64, 14);
           d.aspsol = (const double *) a->data; d.num_aspsol = (uint32_t) (a->count / 7);
        }
      if (-1 == marxb200_set_dither (ctx, &d)) return bail (marxb200_last_error ());
   }
   // ---- HRMA ----
   std::vector<std::vector<uint32_t>> theta_offsets;   // keep alive until the setter returns
   theta_offsets.reserve (16);
   std::vector<std::vector<double>> wf_cols;
   wf_cols.reserve (64);
   if (P.has ("ffield.params"))
     {
        GET (e, "ffield.params", MXCP_F64, 5);
        const double *v = (const double *) e->data;
        marxb200_flatfield_desc d = {v[0], v[1], v[2], v[3], v[4]};
        if (-1 == marxb200_set_flatfield (ctx, &d)) return bail (marxb200_last_error ());
     }
   else
   {
      GET (e, "hrma.params", MXCP_F64, 7);
      const double *v = (const double *) e->data;
      marxb200_hrma_desc d; memset (&d, 0, sizeof (d));
      d.vignetting_factor = v[0]; d.cap_position = v[1]; d.is_ideal = (int32_t) v[2]; d.use_blur = (int32_t) v[3];
      d.use_wfold = (int32_t) v[4]; d.use_struts = (int32_t) v[5]; d.use_scale_factors = (int32_t) v[6];
      GET (oe, "hrma.opt_energies", MXCP_F32, 0); GET (ob, "hrma.opt_betas", MXCP_F32, 0); GET (od, "hrma.opt_deltas", MXCP_F32, 0);
      d.opt_energies = (const float *) oe->data; d.opt_betas = (const float *) ob->data; d.opt_deltas = (const float *) od->data;
      d.num_opt = (uint32_t) oe->count;
      for (int k = 0; k < MARXB200_NUM_SHELLS; k++)
        {
           GET (sp, nm ("hrma.shell%d.params", k), MXCP_F64, 62);
           const double *s = (const double *) sp->data;
           marxb200_hrma_shell &h = d.shells[k];
           int n = 0;
           h.mirror_number = (uint32_t) s[n++]; h.shutter_bitmap = (uint32_t) s[n++];
           h.conic_a_p = s[n++]; h.conic_b_p = s[n++]; h.conic_c_p = s[n++]; h.conic_xmin_p = s[n++]; h.conic_xmax_p = s[n++];
           h.conic_a_h = s[n++]; h.conic_b_h = s[n++]; h.conic_c_h = s[n++]; h.conic_xmin_h = s[n++]; h.conic_xmax_h = s[n++];
           for (int i = 0; i < 3; i++) h.to_osac_p[i] = s[n++];
           for (int i = 0; i < 3; i++) h.to_osac_h[i] = s[n++];
           h.front_position = s[n++]; h.area_fraction = s[n++]; h.min_radius = s[n++]; h.max_radius = s[n++];
           h.p_blur = s[n++]; h.h_blur = s[n++]; h.p_scat_factor = s[n++]; h.h_scat_factor = s[n++];
           for (int i = 0; i < 9; i++) h.fwd_matrix_p[i] = s[n++];
           for (int i = 0; i < 9; i++) h.bwd_matrix_p[i] = s[n++];
           for (int i = 0; i < 9; i++) h.fwd_matrix_h[i] = s[n++];
           for (int i = 0; i < 9; i++) h.bwd_matrix_h[i] = s[n++];
           GET (ce, nm ("hrma.shell%d.corr_energies", k), MXCP_F32, 0);
           GET (cf, nm ("hrma.shell%d.corr_factors", k), MXCP_F32, 0);
           h.corr_energies = (const float *) ce->data; h.corr_factors = (const float *) cf->data; h.num_corr = (uint32_t) ce->count;
           for (int ph = 0; ph < 2; ph++)
             {
                const char *which = ph ? "h" : "p";
                char base[64]; snprintf (base, sizeof (base), "hrma.shell%d.wfold_%s", k, which);
                GET (wh, std::string (base) + ".hdr", MXCP_F64, 0);
                GET (wn, std::string (base) + ".num_theta", MXCP_U32, 0);
                GET (wt, std::string (base) + ".theta", MXCP_F32, 0);
                marxb200_wfold_table &w = ph ? h.h_wfold : h.p_wfold;
                uint32_t na = (uint32_t) wn->count;
                if (wh->count < 6ull * na) return bail (std::string (base) + ".hdr holds fewer than 6 values per array");
                {
                   uint64_t total = 0;
                   for (uint32_t i = 0; i < na; i++) total += ((const uint32_t *) wn->data)[i];
                   if (total > wt->count) return bail (std::string (base) + ".theta is shorter than the sum of num_theta");
                }
                const double *hdr = (const double *) wh->data;
                size_t c0 = wf_cols.size ();
                for (int col = 0; col < 6; col++)
                  {
                     wf_cols.emplace_back (na);
                     for (uint32_t i = 0; i < na; i++) wf_cols.back ()[i] = hdr[6 * i + col];
                  }
                theta_offsets.emplace_back (na);
                uint32_t acc = 0;
                const uint32_t *nt = (const uint32_t *) wn->data;
                for (uint32_t i = 0; i < na; i++) { theta_offsets.back ()[i] = acc; acc += nt[i]; }
                w.num_arrays = na;
                w.e_alpha = wf_cols[c0 + 0].data (); w.p_min = wf_cols[c0 + 1].data (); w.delta_p = wf_cols[c0 + 2].data ();
                w.p_max = wf_cols[c0 + 3].data (); w.pow_law_norm = wf_cols[c0 + 4].data (); w.pow_law_expon = wf_cols[c0 + 5].data ();
                w.num_theta = nt; w.theta_offset = theta_offsets.back ().data ();
                w.theta_values = (const float *) wt->data; w.total_theta = (uint32_t) wt->count;
             }
        }
      if (-1 == marxb200_set_hrma (ctx, &d)) return bail (marxb200_last_error ());
   }
   // ---- grating ----
   std::vector<std::vector<double>> sec_cols;
   sec_cols.reserve (32);
   {
      GET (e, "grating.params", MXCP_F64, 5);
      const double *v = (const double *) e->data;
      marxb200_grating_desc d; memset (&d, 0, sizeof (d));
      d.type = (int32_t) v[0];
      for (int k = 0; k < MARXB200_NUM_SHELLS; k++) d.rowland[k] = v[1 + k];
      if (d.type != 0)
        for (int k = 0; k < MARXB200_NUM_SHELLS; k++)
          {
             GET (gp, nm ("grating.shell%d.params", k), MXCP_F64, 5);
             GET (ol, nm ("grating.shell%d.order_list", k), MXCP_I32, 1);
             GET (en, nm ("grating.shell%d.energies", k), MXCP_F32, 2);
             GET (ce, nm ("grating.shell%d.cum_eff", k), MXCP_F32, ol->count * en->count);
             GET (sc, nm ("grating.shell%d.sectors", k), MXCP_F64, 0);
             const double *g = (const double *) gp->data;
             marxb200_grating_shell &s = d.shells[k];
             s.dispersion_angle = g[0]; s.period = g[1]; s.dp_over_p = g[2]; s.theta_blur = g[3]; s.vig = g[4];
             s.num_orders = (uint32_t) ol->count; s.order_list = (const int32_t *) ol->data;
             s.num_energies = (uint32_t) en->count; s.energies = (const float *) en->data;
             s.cum_eff = (const float *) ce->data;
             if (sc->count % 6 != 0) return bail (nm ("grating.shell%d.sectors is not a multiple of 6 values", k));
             uint32_t ns = (uint32_t) (sc->count / 6);
             const double *sec = (const double *) sc->data;
             size_t c0 = sec_cols.size ();
             for (int col = 0; col < 6; col++)
               {
                  sec_cols.emplace_back (ns);
                  for (uint32_t i = 0; i < ns; i++) sec_cols.back ()[i] = sec[6 * i + col];
               }
             s.num_sectors = ns;
             s.sec_min_angle = sec_cols[c0 + 0].data (); s.sec_max_angle = sec_cols[c0 + 1].data ();
             s.sec_dtheta = sec_cols[c0 + 2].data (); s.sec_dtheta_blur = sec_cols[c0 + 3].data ();
             s.sec_dpp = sec_cols[c0 + 4].data (); s.sec_dpp_blur = sec_cols[c0 + 5].data ();
          }
      if (d.type == 2)
        for (int k = 0; k < 2; k++)
          {
             if (!P.has (nm ("grating.support%d.params", k))) continue;
             GET (gp, nm ("grating.support%d.params", k), MXCP_F64, 5);
             GET (ol, nm ("grating.support%d.order_list", k), MXCP_I32, 0);
             GET (en, nm ("grating.support%d.energies", k), MXCP_F32, 2);
             GET (ce, nm ("grating.support%d.cum_eff", k), MXCP_F32, ol->count * en->count);
             const double *g = (const double *) gp->data;
             marxb200_grating_shell &s = d.support[k];
             s.dispersion_angle = g[0]; s.period = g[1]; s.dp_over_p = g[2]; s.theta_blur = g[3]; s.vig = g[4];
             s.num_orders = (uint32_t) ol->count; s.order_list = (const int32_t *) ol->data;
             s.num_energies = (uint32_t) en->count; s.energies = (const float *) en->data;
             s.cum_eff = (const float *) ce->data;
          }
      if (-1 == marxb200_set_grating (ctx, &d)) return bail (marxb200_last_error ());
   }
   // ---- HRC-S ----
   if (P.has ("hrc.params"))
     {
        GET (e, "hrc.params", MXCP_F64, 44);
        const double *v = (const double *) e->data;
        marxb200_hrc_s_desc d; memset (&d, 0, sizeof (d));
        int n = 0;
        d.detector_type = (int32_t) v[n++]; d.num_mcps = (int32_t) v[n++];
        for (int i = 0; i < 3; i++) d.det_offset[i] = v[n++];
        for (int i = 0; i < 9; i++) d.det_matrix[i] = v[n++];
        d.det_ideal = (int32_t) v[n++]; d.det_extend = (int32_t) v[n++];
        d.shield_t = v[n++]; d.shield_l = v[n++]; d.shield_r = v[n++]; d.shield_x = v[n++]; d.shield_sl = v[n++]; d.shield_sr = v[n++];
        d.shield_sl_gap = v[n++]; d.shield_sr_gap = v[n++]; d.shield_y_center = v[n++]; d.shield_z_center = v[n++];
        for (int i = 0; i < 13; i++) d.blur[i] = v[n++];
        d.u_pixel_size = v[n++]; d.v_pixel_size = v[n++];
        d.use_hesf = (int32_t) v[n++]; d.hesf_num_plates = (int32_t) v[n++]; d.hesf_cr_width = v[n++];
        if ((d.num_mcps > 3) || (d.hesf_num_plates > 4)) return bail ("bad HRC pack");
        for (int k = 0; k < d.num_mcps; k++)
          {
             GET (gm, nm ("hrc.mcp%d.geom", k), MXCP_F64, 21);
             const double *g = (const double *) gm->data;
             marxb200_hrc_mcp &c = d.mcps[k];
             int m = 0;
             c.id = (int32_t) g[m++];
             for (int i = 0; i < 3; i++) c.x_ll[i] = g[m++];
             for (int i = 0; i < 3; i++) c.xhat[i] = g[m++];
             for (int i = 0; i < 3; i++) c.yhat[i] = g[m++];
             for (int i = 0; i < 3; i++) c.normal[i] = g[m++];
             c.xlen = g[m++]; c.ylen = g[m++];
             c.u_start = g[m++]; c.v_start = g[m++]; c.u_0 = g[m++]; c.v_0 = g[m++]; c.cx_0 = g[m++]; c.cy_0 = g[m++];
             GET (qe_e, nm ("hrc.mcp%d.qe_energies", k), MXCP_F32, 0); GET (qe_v, nm ("hrc.mcp%d.qe", k), MXCP_F32, qe_e->count);
             c.qe_num = (uint32_t) qe_e->count; c.qe_energies = (const float *) qe_e->data; c.qe = (const float *) qe_v->data;
          }
        for (int r = 0; r < 4; r++)
          {
             GET (fe, nm ("hrc.filter%d.energies", r), MXCP_F32, 0); GET (fq, nm ("hrc.filter%d.qe", r), MXCP_F32, fe->count);
             d.filter_num[r] = (uint32_t) fe->count; d.filter_energies[r] = (const float *) fe->data; d.filter_qe[r] = (const float *) fq->data;
          }
        GET (hp, "hrc.hesf", MXCP_F64, 14ull * 2 * d.hesf_num_plates);
        const double *h = (const double *) hp->data;
        for (int k = 0; k < 2 * d.hesf_num_plates; k++)
          {
             marxb200_hesf_plate &pl = d.hesf[k];
             const double *q = h + 14 * k;
             for (int i = 0; i < 3; i++) { pl.a[i] = q[i]; pl.e1[i] = q[3 + i]; pl.e2[i] = q[6 + i]; pl.normal[i] = q[9 + i]; }
             pl.len1 = q[12]; pl.len2 = q[13];
          }
        GET (ce, "hrc.hesf_c_energies", MXCP_F32, 0); GET (cb, "hrc.hesf_c_betas", MXCP_F32, ce->count); GET (cd, "hrc.hesf_c_deltas", MXCP_F32, ce->count);
        GET (re, "hrc.hesf_cr_energies", MXCP_F32, 0); GET (rb, "hrc.hesf_cr_betas", MXCP_F32, re->count); GET (rd, "hrc.hesf_cr_deltas", MXCP_F32, re->count);
        d.c_num = (uint32_t) ce->count; d.c_energies = (const float *) ce->data; d.c_betas = (const float *) cb->data; d.c_deltas = (const float *) cd->data;
        d.cr_num = (uint32_t) re->count; d.cr_energies = (const float *) re->data; d.cr_betas = (const float *) rb->data; d.cr_deltas = (const float *) rd->data;
        if (-1 == marxb200_set_hrc_s (ctx, &d)) return bail (marxb200_last_error ());
        rc = 0;
        return 0;
     }
   // ---- ACIS ----
   std::vector<marxb200_fef> fefs;
   {
      GET (e, "acis.params", MXCP_F64, 21);
      const double *v = (const double *) e->data;
      marxb200_acis_desc d; memset (&d, 0, sizeof (d));
      int n = 0;
      d.detector_type = (int32_t) v[n++]; d.num_chips = (int32_t) v[n++];
      for (int i = 0; i < 3; i++) d.det_offset[i] = v[n++];
      for (int i = 0; i < 9; i++) d.det_matrix[i] = v[n++];
      d.det_ideal = (int32_t) v[n++]; d.det_extend = (int32_t) v[n++]; d.focal_length = v[n++];
      d.exposure_time = v[n++]; d.frame_transfer_time = v[n++]; d.frame_time = v[n++]; d.dither_mode = (int32_t) v[n++];
      if (d.num_chips > MARXB200_MAX_CHIPS) return bail ("too many chips");
      if (d.detector_type != 0)
        {
           GET (nf, "acis.num_fefs", MXCP_U32, 1);
           uint32_t num_fefs = *(const uint32_t *) nf->data;
           fefs.resize (num_fefs);
           for (uint32_t j = 0; j < num_fefs; j++)
             {
                GET (dm, nm ("acis.fef%d.dims", (int) j), MXCP_U32, 2);
                const uint32_t *dims = (const uint32_t *) dm->data;
                GET (fe, nm ("acis.fef%d.energies", (int) j), MXCP_F32, dims[1]);
                GET (fc, nm ("acis.fef%d.channels", (int) j), MXCP_F32, dims[1]);
                GET (fg, nm ("acis.fef%d.gauss", (int) j), MXCP_F32, 3ull * dims[0] * dims[1]);
                fefs[j].num_gaussians = dims[0]; fefs[j].num_energies = dims[1];
                fefs[j].energies = (const float *) fe->data; fefs[j].channels = (const float *) fc->data; fefs[j].gauss = (const float *) fg->data;
             }
           d.num_fefs = num_fefs; d.fefs = fefs.data ();
           for (int k = 0; k < d.num_chips; k++)
             {
                marxb200_acis_chip &c = d.chips[k];
                GET (gm, nm ("acis.chip%d.geom", k), MXCP_F64, 19);
                const double *g = (const double *) gm->data;
                int m = 0;
                c.id = (int32_t) g[m++];
                for (int i = 0; i < 3; i++) c.x_ll[i] = g[m++];
                for (int i = 0; i < 3; i++) c.xhat[i] = g[m++];
                for (int i = 0; i < 3; i++) c.yhat[i] = g[m++];
                for (int i = 0; i < 3; i++) c.normal[i] = g[m++];
                c.xlen = g[m++]; c.ylen = g[m++]; c.x_pixel_size = g[m++]; c.y_pixel_size = g[m++];
                c.xpixel_offset = g[m++]; c.ypixel_offset = g[m++];
                GET (qe_e, nm ("acis.chip%d.qe_energies", k), MXCP_F32, 0); GET (qe_v, nm ("acis.chip%d.qe", k), MXCP_F32, qe_e->count);
                GET (fl_e, nm ("acis.chip%d.filter_energies", k), MXCP_F32, 0); GET (fl_v, nm ("acis.chip%d.filter_qe", k), MXCP_F32, fl_e->count);
                c.qe_num = (uint32_t) qe_e->count; c.qe_energies = (const float *) qe_e->data; c.qe = (const float *) qe_v->data;
                c.filter_num = (uint32_t) fl_e->count; c.filter_energies = (const float *) fl_e->data; c.filter_qe = (const float *) fl_v->data;
                GET (ct, nm ("acis.chip%d.contam", k), MXCP_F64, 15);
                const double *cv = (const double *) ct->data;
                c.contam_num_layers = (uint32_t) cv[0]; c.contam_fxy_mode = (int32_t) cv[1];
                c.contam_x0 = cv[2]; c.contam_y0 = cv[3]; c.contam_blocking = (uint32_t) cv[4];
                if (c.contam_num_layers > MARXB200_MAX_CONTAM_LAYERS) return bail ("too many contamination layers");
                for (uint32_t l = 0; l < c.contam_num_layers; l++)
                  {
                     c.contam_tau0[l] = cv[5 + l]; c.contam_tau1[l] = cv[10 + l];
                     GET (ee, nm ("acis.chip%d.contam_e%d", k, (int) l), MXCP_F32, 2);
                     GET (mm, nm ("acis.chip%d.contam_mu%d", k, (int) l), MXCP_F32, ee->count);
                     c.contam_energies[l] = (const float *) ee->data; c.contam_mus[l] = (const float *) mm->data; c.contam_num_mu[l] = (uint32_t) ee->count;
                     if (c.contam_fxy_mode == 0)
                       {
                          uint64_t nb = c.contam_blocking ? 1024 / c.contam_blocking : 0;
                          GET (fx, nm ("acis.chip%d.contam_fxy%d", k, (int) l), MXCP_F32, nb * nb);
                          c.contam_fxy[l] = (const float *) fx->data;
                       }
                  }
                GET (fm, nm ("acis.chip%d.fef_map", k), MXCP_I32, 1024);
                c.fef_map = (const int32_t *) fm->data;
             }
        }
      if (-1 == marxb200_set_acis (ctx, &d)) return bail (marxb200_last_error ());
   }
   rc = 0;
   return 0;
}
