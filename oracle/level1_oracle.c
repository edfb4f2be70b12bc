/* level1_oracle.c -- TEST INFRASTRUCTURE.  Plain-C restatement of the per-event transforms of the stock marx2fits
 * (marx/src/marx2fits.c) used as the parity oracle of the CUDA Level-1 kernels.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; nothing in the product path does.
 *
 * Parity pinned: tests/test_level1_oracle_vs_reference.py requires this file to reproduce, bit for bit, the EVENTS table
 * that oracle/_ref/marx2fits_replay (the UNMODIFIED marx2fits.c linked with the per-row Philox stream of
 * oracle/ref/level1_rng.c) writes for the same MARX output directory, for ACIS-S, ACIS-I, HRC-S, HRC-I and every
 * --pixadj mode, and the committed fixtures tests/golden/level1_*.npz hold such tables.
 *
 * Each function cites the reference lines it follows.  Built with -ffp-contract=off like the rest of the oracle. */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "../include/marxb200.h"

typedef struct { double x, y, z; } vec3;

/* the statics of compute_expno (marx2fits.c:3743), read_expno/read_dither_value (:3552,3567-3580) and the row counter of
 * the draw stream; zero-initialised except last_expno = -1 */
typedef struct
{
   int64_t last_expno;
   int32_t update_dither;
   float dither[6];                 /* ra, dec, roll, dy, dz, dtheta as last taken over */
   uint64_t rows;
}
level1_state;

typedef struct
{
   uint64_t n;
   const float *time, *xpixel, *ypixel, *b_energy, *hrc_u, *hrc_v;
   const float *dither[6];          /* sky_ra, sky_dec, sky_roll, det_dy, det_dz, det_theta (.dat columns); NULL: absent */
   const int16_t *pha;
   const int8_t *ccd;
}
level1_input;

/* ---- draw stream (include/marxb200.h "Level-1 event transforms") ---- */
static void philox4x32_10 (uint32_t c[4], uint32_t k0, uint32_t k1)
{
   for (int i = 0; i < 10; i++)
     {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c[0], p1 = (uint64_t) 0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t) p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
     }
}
static double row_uniform (uint64_t seed, uint64_t row, uint32_t d)    /* JDMrandom, jdmath/src/random.c:151-154 */
{
   uint32_t c[4] = { (uint32_t) row, (uint32_t) (row >> 32), d >> 2, MARXB200_STAGE_LEVEL1 };
   philox4x32_10 (c, (uint32_t) seed, (uint32_t) (seed >> 32));
   return (double) c[d & 3] * (1.0 / (double) 0xFFFFFFFFU);
}

/* ---- jdmath/src/vector.c ---- */
static double v_length (vec3 a)                                       /* JDMv_length :77-97 */
{
   double x = fabs (a.x), y = fabs (a.y), z = fabs (a.z), tmp;
   if (z < x) { tmp = z; z = x; x = tmp; }
   if (z < y) { tmp = z; z = y; y = tmp; }
   if (z == 0.0) return 0.0;
   x = x / z; y = y / z;
   return z * sqrt (1.0 + x * x + y * y);
}
static void v_normalize (vec3 *a)                                     /* JDMv_normalize :99-109 */
{
   double len = v_length (*a);
   if (len != 0.0) { a->x = a->x / len; a->y = a->y / len; a->z = a->z / len; }
}
static double v_dot (vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static vec3 v_cross (vec3 a, vec3 b)                                  /* JDMv_pcross_prod :49-57 */
{
   vec3 c;
   c.z = a.x * b.y - a.y * b.x;
   c.x = a.y * b.z - a.z * b.y;
   c.y = a.z * b.x - a.x * b.z;
   return c;
}
static vec3 v_rotate_unit1 (vec3 p, vec3 n, double c, double s)       /* JDMv_rotate_unit_vector1 :183-206 */
{
   double pn = v_dot (p, n), f = pn * (1.0 - c);
   vec3 x = v_cross (n, p), u;
   u.x = c * p.x + f * n.x + s * x.x;
   u.y = c * p.y + f * n.y + s * x.y;
   u.z = c * p.z + f * n.z + s * x.z;
   v_normalize (&u);
   return u;
}

/* unapply_dither, marx/libsrc/dither.c:583-607 (marx_undither_mnc :703-707) */
static vec3 unapply_dither (double ra, double dec, double roll, vec3 p)
{
   double cos_ra = cos (ra), sin_ra = sin (ra), cos_dec = cos (dec), sin_dec = sin (dec);
   double cos_theta = cos_dec * cos_ra, sin_theta;
   vec3 n = { 0, -sin_dec, cos_dec * sin_ra }, xaxis = { 1, 0, 0 };
   sin_theta = v_length (n);
   if (sin_theta >= 1e-20)                                            /* VERY_TINY_NUMBER, dither.c:549 */
     {
        n.x /= sin_theta; n.y /= sin_theta; n.z /= sin_theta;
        p = v_rotate_unit1 (p, n, cos_theta, sin_theta);
     }
   return v_rotate_unit1 (p, xaxis, cos (roll), sin (roll));          /* JDMv_rotate_unit_vector :208-212 */
}

/* marx_mnc_to_ra_dec, pixlib.c:503-527 */
static void mnc_to_ra_dec (vec3 mnc, double *ra, double *dec)
{
   double x = mnc.x, y = mnc.y, perp = sqrt (x * x + y * y);
   if (perp > 1.0) perp = 1.0;
   if (mnc.z <= 0) *dec = acos (perp); else *dec = -acos (perp);
   perp = -x / perp;
   if (y <= 0.0) *ra = acos (perp); else *ra = -acos (perp);
}
/* marx_compute_ra_dec_offsets, pixlib.c:535-573 */
static void ra_dec_offsets (double ra_0, double dec_0, double ra, double dec, double *delta_ra, double *delta_dec)
{
   const double PI = 3.14159265358979323846;
   double d_ra = ra - ra_0, d_dec = dec - dec_0, c = cos (dec), factor = c * (1 - cos (d_ra));
   double sin_delta_dec = sin (d_dec) + factor * sin (dec_0), num, den;
   if (fabs (sin_delta_dec) > 1.0) sin_delta_dec = (sin_delta_dec < 0) ? -1 : 1;
   *delta_dec = asin (sin_delta_dec);
   num = c * sin (d_ra);
   den = cos (d_dec) - factor * cos (dec_0);
   if (den >= 0) *delta_ra = atan (num / den);
   else if (num >= 0) *delta_ra = atan (num / den) + PI;
   else *delta_ra = atan (num / den) - PI;
}

/* JDMbinary_search_f, jdmath/src/finterpo.c:37-57 */
static unsigned int bsearch_f (float x, const float *xp, unsigned int n)
{
   unsigned int n0 = 0, n1 = n, n2;
   while (n1 > n0 + 1)
     {
        n2 = (n0 + n1) / 2;
        if (xp[n2] >= x)
          {
             if (xp[n2] == x) return n2;
             n1 = n2;
          }
        else n0 = n2;
     }
   if (x >= xp[n0]) return n1;
   return n0;
}

/* marx_compute_acis_subpix, acis_subpix.c:283-321 */
static void acis_subpix (const marxb200_level1_desc *D, int table, float energy, int fltgrade, float *dxp, float *dyp)
{
   int n = D->subpix_npoints[table * 256 + fltgrade];
   if (n <= 0) { *dxp = *dyp = 0; return; }
   const float *en = D->subpix_data + D->subpix_offset[table * 256 + fltgrade], *dxs = en + n, *dys = dxs + n;
   unsigned int j = bsearch_f (energy, en, (unsigned int) n);
   if (j == 0) j++;
   if (j == (unsigned int) n) j--;
   double w1 = ((double) energy - en[j - 1]) / (en[j] - en[j - 1]);
   double w0 = 1.0 - w1;
   *dxp = w0 * dxs[j - 1] + w1 * dxs[j];
   *dyp = w0 * dys[j - 1] + w1 * dys[j];
}

/* marx_compute_tiled_pixel, detpix.c:151-177 -> acis_i_to_tiled (acis_geom.c:111-143), acis_s_to_tiled (:181-202),
 * hrc_s_to_tiled (hrc_s_geom.c:477-515), hrc_i_to_tiled (hrc_i_geom.c:216-236).  x, y arrive as unsigned int. */
static void tiled_pixel (const marxb200_level1_desc *D, const marxb200_level1_chip *g, unsigned int x, unsigned int y,
                         unsigned int *xp, unsigned int *yp)
{
   float xf, yf;
   if (D->detector_type == 4)
     {
        if ((g->id == 0) || (g->id == 2)) { xf = g->tdet_xoff + y; yf = g->tdet_yoff - x; }
        else { xf = g->tdet_xoff - y; yf = g->tdet_yoff + x; }
     }
   else if (D->detector_type == 1) { yf = y; yf = yf + g->tdet_yoff; xf = x + g->tdet_xoff; }
   else { xf = x + g->tdet_xoff; yf = y + g->tdet_yoff; }
   if (xf < 0.0) xf = 0.0;
   if (yf < 0.0) yf = 0.0;
   *xp = (unsigned int) xf;
   *yp = (unsigned int) yf;
}

/* Grade_Map, marx2fits.c:3793-3811 (the CALDB grade table acisD2009-11-01gradeN0005: flight grade -> ASCA grade);
 * restated from its definition: grades 0..6 are the listed flight grades, everything else is 7. */
static int asca_grade (int f)
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
int level1_oracle_grade (int fltgrade) { return asca_grade (fltgrade & 0xFF); }

/* Flight_Grade_Table, marx2fits.c:3826-3837, flat (row 3*dy+dx, 4 entries each); entry 36 stands for the word the
 * reference reads past the end when the draw is exactly 1.0 (probability 2^-32) */
static const int Flight_Grades[37] =
{
   10, 11, 138, 139,   2, 34, 130, 162,   18, 22, 50, 54,
   8, 12, 136, 140,    0, 0, 0, 0,        16, 17, 48, 49,
   72, 76, 104, 108,   64, 65, 68, 69,    80, 81, 208, 209,   0
};

static const marxb200_level1_chip *find_chip (const marxb200_level1_desc *D, int id)
{
   for (int k = 0; k < D->num_chips; k++) if (D->chips[k].id == id) return &D->chips[k];
   return NULL;
}

void level1_oracle_reset (level1_state *S)
{
   S->last_expno = -1; S->update_dither = 0; S->rows = 0;
   for (int k = 0; k < 6; k++) S->dither[k] = 0;
}

/* compute_table_values (marx2fits.c:1448-1468) for every row of the input, in Data_Def_Table order.  Returns the number
 * of rows transformed, or -(row + 1) on the first row the reference would reject. */
long level1_oracle_transform (const marxb200_level1_desc *D, uint64_t seed, level1_state *S, const level1_input *in,
                              const marxb200_level1_columns *out)
{
   const int acis = (D->detector_type == 3) || (D->detector_type == 4);
   if (!acis) S->update_dither = 1;                                   /* marx2fits :2848-2849 */
   for (uint64_t i = 0; i < in->n; i++)
     {
        const uint64_t row = S->rows + i;
        uint32_t draw = 0;
        double time = (double) in->time[i];                           /* read_float32_to_float64 */
        int16_t ccdid = (int16_t) in->ccd[i];                         /* read_byte_to_int16 */
        float benergy = (acis && in->b_energy) ? in->b_energy[i] : 0.0f;
        float chipx = in->xpixel[i] + 1, chipy = in->ypixel[i] + 1;   /* read_float32_add_1 :3460-3470 */
        int32_t pha = (int32_t) in->pha[i];
        int32_t expno = 0;
        const marxb200_level1_chip *g = find_chip (D, ccdid);
        if (g == NULL) return -(long) (i + 1);

        if (out->hrc_u && in->hrc_u) out->hrc_u[i] = (int32_t) in->hrc_u[i];   /* read_float32_to_int32 */
        if (out->hrc_v && in->hrc_v) out->hrc_v[i] = (int32_t) in->hrc_v[i];

        if (acis)                                                     /* compute_expno :3741-3763 */
          {
             if (D->time_del <= 0.0)
               {
                  expno = (int32_t) S->last_expno;
                  S->last_expno++;
                  S->update_dither = 1;
               }
             else
               {
                  long e = (long) (time / D->time_del);
                  expno = (int32_t) e;
                  S->update_dither = (e != S->last_expno);
                  S->last_expno = e;
               }
          }
        unsigned int tx, ty;                                          /* compute_tdetxy :3584-3599 */
        tiled_pixel (D, g, (unsigned int) chipx, (unsigned int) chipy, &tx, &ty);

        for (int k = 0; k < 6; k++)                                   /* read_dither_value :3567-3580 */
          {
             float val = (D->used_dither && in->dither[k]) ? in->dither[k][i] : 0.0f;
             if (S->update_dither || (D->pix_adjust == MARXB200_PIXADJ_EXACT)) S->dither[k] = val;
          }
        const double d_ra = S->dither[0], d_dec = S->dither[1], d_roll = S->dither[2];
        const double d_dy = S->dither[3], d_dz = S->dither[4], d_theta = S->dither[5];

        int16_t fltgrade = 0, grade = 0;
        if (acis)                                                     /* compute_fltgrade :3854-3866, compute_grade :3813 */
          {
             int dx = (int) (3.0 * (chipx - floor (chipx)));
             int dy = (int) (3.0 * (chipy - floor (chipy)));
             fltgrade = (int16_t) Flight_Grades[4 * (3 * dy + dx) + (int) (4 * row_uniform (seed, row, draw++))];
             grade = (int16_t) asca_grade ((unsigned char) fltgrade);
          }

        /* compute_detxy :3676-3737 with marx_init_chip_to_mnc / marx_chip_to_mnc (pixlib.c:139-225) */
        double cos_theta = cos (d_theta), sin_theta = sin (d_theta);
        vec3 ofs, e1, e2;
        ofs.x = D->det_offset[0] - D->focal_length;
        ofs.y = (D->det_offset[1] + d_dy) + cos_theta * g->x_ll[1] - sin_theta * g->x_ll[2];
        ofs.z = (D->det_offset[2] + d_dz) + sin_theta * g->x_ll[1] + cos_theta * g->x_ll[2];
        e1.x = g->xhat[0];
        e1.y = cos_theta * g->xhat[1] - sin_theta * g->xhat[2];
        e1.z = sin_theta * g->xhat[1] + cos_theta * g->xhat[2];
        e2.x = g->yhat[0];
        e2.y = cos_theta * g->yhat[1] - sin_theta * g->yhat[2];
        e2.z = sin_theta * g->yhat[1] + cos_theta * g->yhat[2];
        e1.x *= g->x_pixel_size; e1.y *= g->x_pixel_size; e1.z *= g->x_pixel_size;   /* JDMv_smul */
        e2.x *= g->y_pixel_size; e2.y *= g->y_pixel_size; e2.z *= g->y_pixel_size;

        double x = chipx - 1.0, y = chipy - 1.0;
        switch (D->pix_adjust)
          {
           case MARXB200_PIXADJ_EXACT: break;
           case MARXB200_PIXADJ_NONE: x = floor (x) + 0.5; y = floor (y) + 0.5; break;
           case MARXB200_PIXADJ_RANDOMIZE:
             x = floor (x) + row_uniform (seed, row, draw++);
             y = floor (y) + row_uniform (seed, row, draw++);
             break;
           default:
             {
                float sdx, sdy;
                if ((fltgrade < 0) || (fltgrade >= 256)) return -(long) (i + 1);
                acis_subpix (D, g->subpix_table, benergy, fltgrade, &sdx, &sdy);
                x = floor (x) + 0.5 + sdx;
                y = floor (y) + 0.5 + sdy;
             }
          }
        double xpixel = x - g->xpixel_offset, ypixel = y - g->ypixel_offset;
        vec3 mnc;
        mnc.x = ofs.x + e1.x * xpixel + e2.x * ypixel;
        mnc.y = ofs.y + e1.y * xpixel + e2.y * ypixel;
        mnc.z = ofs.z + e1.z * xpixel + e2.z * ypixel;
        v_normalize (&mnc);
        if (mnc.x == 0.0) return -(long) (i + 1);                     /* marx_mnc_to_fpc, detpix.c:182-209 */
        double factor = 1.0 / (D->fp_delta_s0 * mnc.x);
        double detx = D->fp_x0 - factor * mnc.y, dety = D->fp_y0 + factor * mnc.z;

        double xsky, ysky;                                            /* compute_xy_sky :3869-3911 */
        if (D->used_dither == 0)
          {
             double theta = D->nominal_roll * 3.14159265358979323846 / 180.0;
             double c = cos (theta), s = sin (theta);
             double xx = detx - D->fp_x0, yy = dety - D->fp_y0;
             xsky = D->fp_x0 + c * xx + s * yy;
             ysky = D->fp_y0 - s * xx + c * yy;
          }
        else
          {
             double ra, dec, ora, odec;
             mnc = unapply_dither (d_ra, d_dec, d_roll, mnc);
             mnc_to_ra_dec (mnc, &ra, &dec);
             ra_dec_offsets (0, 0, ra, dec, &ora, &odec);
             ora = -ora;
             xsky = ora / D->fp_delta_s0 + D->fp_x0;
             ysky = odec / D->fp_delta_s0 + D->fp_y0;
          }

        if (out->time) out->time[i] = ((acis && (D->time_del > 0.0)) ? D->time_del * expno : time) + D->time_start;   /* write_time :3433 */
        if (out->expno) out->expno[i] = expno;
        if (out->ccd_id) out->ccd_id[i] = ccdid;
        if (out->node_id) out->node_id[i] = acis ? (int16_t) ((chipx - 1) / 256) : 0;       /* compute_node_id :3765 */
        if (out->chipx) out->chipx[i] = (int16_t) chipx;              /* write_float32_as_int16 */
        if (out->chipy) out->chipy[i] = (int16_t) chipy;
        if (out->tdetx) out->tdetx[i] = (int32_t) (tx + 1);
        if (out->tdety) out->tdety[i] = (int32_t) (ty + 1);
        if (out->detx) out->detx[i] = detx;
        if (out->dety) out->dety[i] = dety;
        if (out->x) out->x[i] = xsky;
        if (out->y) out->y[i] = ysky;
        if (out->pha) out->pha[i] = pha;
        if (out->energy) out->energy[i] = acis ? (float) (benergy * 1e3) : 0.0f;             /* compute_acis_energy :3923 */
        if (out->pi) out->pi[i] = acis ? (int16_t) (benergy * D->pi_factor * 1e3 + 1.0) : 0;  /* compute_pi :3931 */
        if (out->fltgrade) out->fltgrade[i] = fltgrade;
        if (out->grade) out->grade[i] = grade;
        if (out->status)                                              /* compute_status :3775-3790 */
          {
             int16_t st = 0;
             if (acis && (((chipx < 2.0) || (chipx >= 1024.0)) || ((chipy < 2.0) || (chipy >= 1024.0)))) st |= 0x0001;
             out->status[i] = st;
          }
        if (out->keep) out->keep[i] = (pha != -1);                    /* marx2fits :2856-2857 */
     }
   S->rows += in->n;
   return (long) in->n;
}
