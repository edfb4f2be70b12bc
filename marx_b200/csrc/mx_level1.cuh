// mx_level1.cuh -- Level-1 event transforms: the per-event part of marx2fits (marx/src/marx2fits.c:3584-3943) for the
// device-resident event list (include/marxb200.h "Level-1 event transforms").
// Reference: compute_expno :3741-3763, compute_tdetxy :3584-3599 (detpix.c:151-177, acis_geom.c:111-143,181-202,
// hrc_s_geom.c:477-515, hrc_i_geom.c:216-236), read_dither_value :3567-3580, compute_fltgrade :3854-3866, compute_grade
// :3813-3818, compute_detxy :3676-3737 (pixlib.c:139-225, acis_subpix.c:283-321, detpix.c:182-209), compute_xy_sky
// :3869-3911 (dither.c:583-607, pixlib.c:503-573), compute_acis_energy/_pi :3923-3941, compute_node_id :3765,
// compute_status :3775-3790, write_time :3433-3445.
// Same operation order and the same float/double narrowing points as the reference; the inputs are narrowed exactly as
// marx_write_photons narrows them into the column files marx2fits reads (marxio.c:217-290).
#pragma once
#include "mx_common.cuh"

namespace mx {

constexpr int kL1MaxChips = 10;
constexpr uint32_t kStageLevel1 = 4;       // MARXB200_STAGE_LEVEL1
enum { L1_PIXADJ_NONE = 0, L1_PIXADJ_RANDOMIZE = 1, L1_PIXADJ_EDSER = 2, L1_PIXADJ_EXACT = 3 };   // marx2fits.c:60-63

struct Level1ChipDev
{
   int32_t id, subpix_table;
   double x_ll[3], xhat[3], yhat[3];
   double x_pixel_size, y_pixel_size, xpixel_offset, ypixel_offset;
   float tdet_xoff, tdet_yoff;
};
struct Level1Dev
{
   int32_t detector_type, num_chips;
   Level1ChipDev chips[kL1MaxChips];
   double fp_delta_s0, fp_x0, fp_y0, focal_length, det_offset[3];
   double time_del, time_start, pi_factor;
   double roll_cos, roll_sin;              // cos/sin (Nominal_Roll * PI / 180), evaluated by the host's libm at set time
   int32_t used_dither, pix_adjust;
   const int32_t *subpix_npoints;          // [2 * 256]
   const uint32_t *subpix_offset;          // [2 * 256]
   const float *subpix_data;
};
// the statics of compute_expno / read_dither_value + the row counter of the draw stream, carried from batch to batch
struct Level1State
{
   long long last_expno;
   unsigned long long rows;
   float dither[6];
   int32_t pad[2];
};
struct Level1Cols
{
   double *time, *detx, *dety, *x, *y;
   int32_t *expno, *tdetx, *tdety, *pha, *hrc_u, *hrc_v;
   float *energy;
   int16_t *ccd_id, *node_id, *chipx, *chipy, *pi, *fltgrade, *grade, *status;
   uint8_t *keep;
};

MX_HD bool l1_is_acis (const Level1Dev &D) { return (D.detector_type == 3) || (D.detector_type == 4); }
// events of one ACIS exposure frame share the aspect of the frame's first event (read_dither_value :3577-3578)
MX_HD bool l1_frames_share_aspect (const Level1Dev &D)
{
   return l1_is_acis (D) && (D.time_del > 0.0) && (D.pix_adjust != L1_PIXADJ_EXACT);
}
// (float) (arrival_time + total_time): the value marx_write_photons puts into time.dat (marxio.c:246)
MX_HD float l1_file_time (double t_abs, double start_time, double total_time) { return (float) ((t_abs - start_time) + total_time); }
MX_HD long long l1_expno (float tfile, double time_del) { return (long long) ((double) tfile / time_del); }

// one draw of row `row`: lane d&3 of Philox4x32-10 (key = seed, counter = (row_lo, row_hi, d>>2, LEVEL1))
struct Level1Draws
{
   uint32_t b[4];
   MX_HD void init (uint64_t seed, uint64_t row)
   {
      Rng r;
      r.init (seed, row, kStageLevel1);
      r.refill (0);
      b[0] = r.b0; b[1] = r.b1; b[2] = r.b2; b[3] = r.b3;
   }
   MX_HD double uniform (int d) const { return (double) b[d] * (1.0 / 4294967295.0); }
};

// Grade_Map, marx2fits.c:3793-3811 (flight grade -> ASCA grade, CALDB acisD2009-11-01gradeN0005): the flight grades of
// ASCA grades 0..6; every other flight grade is grade 7
MX_HD int l1_asca_grade (int f)
{
   switch (f)
     {
      case 0: return 0;
      case 1: case 4: case 5: case 32: case 33: case 36: case 37: case 128: case 129: case 132: case 133:
      case 160: case 161: case 164: case 165: return 1;
      case 2: case 34: case 64: case 65: case 68: case 69: case 130: case 162: return 2;
      case 8: case 12: case 136: case 140: return 3;
      case 16: case 17: case 48: case 49: return 4;
      case 3: case 6: case 9: case 13: case 20: case 21: case 35: case 38: case 40: case 44: case 52: case 53: case 96: case 97:
      case 100: case 101: case 131: case 134: case 137: case 141: case 144: case 145: case 163: case 166: case 168: case 172:
      case 176: case 177: case 192: case 193: case 196: case 197: return 5;
      case 10: case 11: case 18: case 22: case 50: case 54: case 72: case 76: case 80: case 81: case 104: case 108: case 138:
      case 139: case 208: case 209: return 6;
     }
   return 7;
}
// Flight_Grade_Table, marx2fits.c:3826-3837: the four flight grades of sub-pixel (dx, dy), row 3*dy+dx
MX_HD int l1_flight_grade (int cell, int pick)
{
   // entry 36: the word the reference reads past the table when the draw is exactly 1.0 (probability 2^-32)
   const int idx = 4 * cell + pick;
   switch (idx >> 2)
     {
      case 0: { const int t[4] = {10, 11, 138, 139}; return t[idx & 3]; }
      case 1: { const int t[4] = {2, 34, 130, 162}; return t[idx & 3]; }
      case 2: { const int t[4] = {18, 22, 50, 54}; return t[idx & 3]; }
      case 3: { const int t[4] = {8, 12, 136, 140}; return t[idx & 3]; }
      case 4: return 0;
      case 5: { const int t[4] = {16, 17, 48, 49}; return t[idx & 3]; }
      case 6: { const int t[4] = {72, 76, 104, 108}; return t[idx & 3]; }
      case 7: { const int t[4] = {64, 65, 68, 69}; return t[idx & 3]; }
      case 8: { const int t[4] = {80, 81, 208, 209}; return t[idx & 3]; }
     }
   return 0;
}

// marx_compute_acis_subpix, acis_subpix.c:283-321
MX_HD void l1_acis_subpix (const Level1Dev &D, int table, float energy, int fltgrade, float &dxp, float &dyp)
{
   const int n = D.subpix_npoints[table * 256 + fltgrade];
   if (n <= 0) { dxp = 0.0f; dyp = 0.0f; return; }
   const float *en = D.subpix_data + D.subpix_offset[table * 256 + fltgrade], *dxs = en + n, *dys = dxs + n;
   uint32_t j = bsearch_f (energy, en, (uint32_t) n);
   if (j == 0) j++;
   if (j == (uint32_t) n) j--;
   const double w1 = ((double) energy - en[j - 1]) / (en[j] - en[j - 1]);
   const double w0 = 1.0 - w1;
   dxp = (float) (w0 * dxs[j - 1] + w1 * dxs[j]);
   dyp = (float) (w0 * dys[j - 1] + w1 * dys[j]);
}

// marx_compute_tiled_pixel -> acis_i_to_tiled / acis_s_to_tiled / hrc_s_to_tiled / hrc_i_to_tiled (float arithmetic)
MX_HD void l1_tiled_pixel (const Level1Dev &D, const Level1ChipDev &g, unsigned int x, unsigned int y, unsigned int &xp, unsigned int &yp)
{
   float xf, yf;
   if (D.detector_type == 4)
     {
        if ((g.id == 0) || (g.id == 2)) { xf = g.tdet_xoff + (float) y; yf = g.tdet_yoff - (float) x; }
        else { xf = g.tdet_xoff - (float) y; yf = g.tdet_yoff + (float) x; }
     }
   else { xf = (float) x + g.tdet_xoff; yf = (float) y + g.tdet_yoff; }
   if (xf < 0.0f) xf = 0.0f;
   if (yf < 0.0f) yf = 0.0f;
   xp = (unsigned int) xf;
   yp = (unsigned int) yf;
}

// unapply_dither, dither.c:583-607
MX_HD Vec3 l1_unapply_dither (double ra, double dec, double roll, Vec3 p)
{
   double cos_ra, sin_ra, cos_dec, sin_dec, cr, sr;
   sin_cos (ra, sin_ra, cos_ra);
   sin_cos (dec, sin_dec, cos_dec);
   const double cos_theta = cos_dec * cos_ra;
   Vec3 n = v_make (0.0, -sin_dec, cos_dec * sin_ra);
   const double sin_theta = v_length (n);
   if (sin_theta >= 1e-20)
     {
        // n.x is +0 and sin_theta > 0: n.x / sin_theta is +0 again, and a zero quotient runs the FP64 division's special-case routine
        // (4.5 % of l1_transform's instructions, tools/ncu_calls.py); see div_in_branch, mx_common.cuh
        n.y /= sin_theta; n.z /= sin_theta;
        p = v_rotate_unit1 (p, n, cos_theta, sin_theta);
     }
   sin_cos (roll, sr, cr);
   return v_rotate_unit1 (p, v_make (1.0, 0.0, 0.0), cr, sr);
}

struct Level1In
{
   float tfile;                // time.dat value
   float xpixel, ypixel;       // xpixel.dat / ypixel.dat (0-based chip pixels)
   float benergy;              // b_energy.dat (ACIS PI energy)
   float upix, vpix;           // hrc_u.dat / hrc_v.dat
   int16_t pha;
   int8_t ccd;
   float dither[6];            // the aspect this row computes with (already resolved to its frame's first event)
   long long expno;
   uint64_t row;
};
struct Level1Out
{
   double time, detx, dety, x, y;
   int32_t expno, tdetx, tdety, pha, hrc_u, hrc_v;
   float energy;
   int16_t ccd_id, node_id, chipx, chipy, pi, fltgrade, grade, status;
   uint8_t keep;
   int32_t error;
};

// compute_table_values (marx2fits.c:1448-1468) for one row, Data_Def_Table order
MX_HD void level1_event (const Level1Dev &D, uint64_t seed, const Level1In &in, Level1Out &o)
{
   const bool acis = l1_is_acis (D);
   o.error = 0;
   const double time = (double) in.tfile;
   const float chipx = in.xpixel + 1.0f, chipy = in.ypixel + 1.0f;          // read_float32_add_1
   o.ccd_id = (int16_t) in.ccd;
   o.pha = (int32_t) in.pha;
   o.hrc_u = (int32_t) in.upix; o.hrc_v = (int32_t) in.vpix;
   o.expno = (int32_t) in.expno;

   int ci = -1;
   for (int k = 0; k < D.num_chips; k++) if (D.chips[k].id == (int32_t) in.ccd) { ci = k; break; }
   if (ci < 0) { o.error = 1; ci = 0; }
   const Level1ChipDev &g = D.chips[ci];

   unsigned int tx, ty;
   l1_tiled_pixel (D, g, (unsigned int) chipx, (unsigned int) chipy, tx, ty);
   o.tdetx = (int32_t) (tx + 1u); o.tdety = (int32_t) (ty + 1u);

   const double d_ra = in.dither[0], d_dec = in.dither[1], d_roll = in.dither[2];
   const double d_dy = in.dither[3], d_dz = in.dither[4], d_theta = in.dither[5];

   Level1Draws draws;
   int draw = 0;
   if (acis || (D.pix_adjust == L1_PIXADJ_RANDOMIZE)) draws.init (seed, in.row);
   int fltgrade = 0, grade = 0;
   if (acis)
     {
        const int dx = (int) (3.0 * ((double) chipx - floor ((double) chipx)));
        const int dy = (int) (3.0 * ((double) chipy - floor ((double) chipy)));
        fltgrade = l1_flight_grade (3 * dy + dx, (int) (4 * draws.uniform (draw++)));
        grade = l1_asca_grade (fltgrade & 0xFF);
     }
   o.fltgrade = (int16_t) fltgrade; o.grade = (int16_t) grade;

   // marx_init_chip_to_mnc, pixlib.c:139-197
   double cos_theta = 1.0, sin_theta = 0.0;
   if (d_theta != 0.0) sin_cos (d_theta, sin_theta, cos_theta);
   Vec3 ofs, e1, e2;
   ofs.x = D.det_offset[0] - D.focal_length;
   ofs.y = (D.det_offset[1] + d_dy) + cos_theta * g.x_ll[1] - sin_theta * g.x_ll[2];
   ofs.z = (D.det_offset[2] + d_dz) + sin_theta * g.x_ll[1] + cos_theta * g.x_ll[2];
   e1.x = g.xhat[0];
   e1.y = cos_theta * g.xhat[1] - sin_theta * g.xhat[2];
   e1.z = sin_theta * g.xhat[1] + cos_theta * g.xhat[2];
   e2.x = g.yhat[0];
   e2.y = cos_theta * g.yhat[1] - sin_theta * g.yhat[2];
   e2.z = sin_theta * g.yhat[1] + cos_theta * g.yhat[2];
   e1.x *= g.x_pixel_size; e1.y *= g.x_pixel_size; e1.z *= g.x_pixel_size;
   e2.x *= g.y_pixel_size; e2.y *= g.y_pixel_size; e2.z *= g.y_pixel_size;

   double x = (double) chipx - 1.0, y = (double) chipy - 1.0;
   switch (D.pix_adjust)
     {
      case L1_PIXADJ_EXACT: break;
      case L1_PIXADJ_NONE: x = floor (x) + 0.5; y = floor (y) + 0.5; break;
      case L1_PIXADJ_RANDOMIZE:
        x = floor (x) + draws.uniform (draw); y = floor (y) + draws.uniform (draw + 1);
        break;
      default:
        {
           float sdx, sdy;
           l1_acis_subpix (D, g.subpix_table, in.benergy, fltgrade, sdx, sdy);
           x = floor (x) + 0.5 + sdx; y = floor (y) + 0.5 + sdy;
        }
     }
   // marx_chip_to_mnc, pixlib.c:199-225
   const double xpixel = x - g.xpixel_offset, ypixel = y - g.ypixel_offset;
   Vec3 mnc;
   mnc.x = ofs.x + e1.x * xpixel + e2.x * ypixel;
   mnc.y = ofs.y + e1.y * xpixel + e2.y * ypixel;
   mnc.z = ofs.z + e1.z * xpixel + e2.z * ypixel;
   v_normalize (mnc);
   if (mnc.x == 0.0) { o.error = 2; mnc.x = -1.0; }
   // marx_mnc_to_fpc, detpix.c:182-209
   const double factor = 1.0 / (D.fp_delta_s0 * mnc.x);
   o.detx = D.fp_x0 - factor * mnc.y;
   o.dety = D.fp_y0 + factor * mnc.z;

   // compute_xy_sky :3869-3911
   if (D.used_dither == 0)
     {
        const double xx = o.detx - D.fp_x0, yy = o.dety - D.fp_y0;
        o.x = D.fp_x0 + D.roll_cos * xx + D.roll_sin * yy;
        o.y = D.fp_y0 - D.roll_sin * xx + D.roll_cos * yy;
     }
   else
     {
        mnc = l1_unapply_dither (d_ra, d_dec, d_roll, mnc);
        // marx_mnc_to_ra_dec, pixlib.c:503-527
        double perp = sqrt (mnc.x * mnc.x + mnc.y * mnc.y);
        if (perp > 1.0) perp = 1.0;
        double dec = acos (perp);
        if (!(mnc.z <= 0)) dec = -dec;
        perp = -mnc.x / perp;
        double ra = acos (perp);
        if (!(mnc.y <= 0.0)) ra = -ra;
        // marx_compute_ra_dec_offsets (0, 0, ra, dec), pixlib.c:535-573: sin (dec_0) = 0, cos (dec_0) = 1 exactly
        double s_ra, c_ra, s_dec, c_dec;
        sin_cos (ra, s_ra, c_ra);
        sin_cos (dec, s_dec, c_dec);
        const double factor2 = c_dec * (1 - c_ra);
        double sin_delta_dec = s_dec + factor2 * 0.0;
        if (fabs (sin_delta_dec) > 1.0) sin_delta_dec = (sin_delta_dec < 0) ? -1.0 : 1.0;
        const double delta_dec = asin (sin_delta_dec);
        const double num = c_dec * s_ra, den = c_dec - factor2 * 1.0;
        double delta_ra = atan (num / den);
        if (!(den >= 0)) delta_ra = (num >= 0) ? delta_ra + kPI : delta_ra - kPI;
        delta_ra = -delta_ra;
        o.x = delta_ra / D.fp_delta_s0 + D.fp_x0;
        o.y = delta_dec / D.fp_delta_s0 + D.fp_y0;
     }

   o.energy = acis ? (float) ((double) in.benergy * 1e3) : 0.0f;
   o.pi = acis ? (int16_t) ((double) in.benergy * D.pi_factor * 1e3 + 1.0) : (int16_t) 0;
   o.node_id = acis ? (int16_t) ((chipx - 1.0f) / 256.0f) : (int16_t) 0;
   o.chipx = (int16_t) chipx; o.chipy = (int16_t) chipy;                  // write_float32_as_int16
   int16_t st = 0;
   if (acis && (((chipx < 2.0f) || (chipx >= 1024.0f)) || ((chipy < 2.0f) || (chipy >= 1024.0f)))) st |= 0x0001;
   o.status = st;
   o.time = ((acis && (D.time_del > 0.0)) ? D.time_del * (double) o.expno : time) + D.time_start;
   o.keep = (uint8_t) (in.pha != -1);
}

}  // namespace mx
