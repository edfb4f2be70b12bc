/* pileup_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of marxpileup's frame loop (marx/src/marxpileup.c; SURVEY.md 8f
 * rank 4), the ACIS pile-up model applied to the event list of a simulation.  Only tests/ (and later smoke / bench's cpu_baseline)
 * may use it; the product (marx_b200/) never does.  The device implementation is marx_b200/csrc/mx_pileup.cuh + pileup_kernels.cu
 * behind marxb200_pileup_run (tests/test_gpu_zz_pileup.py compares the two).
 *
 * Parity PINNED: tests/test_pileup_oracle_vs_reference.py requires this file to reproduce the output directory of the stock program
 * run with counter-based draws (oracle/_ref/marxpileup_replay = the unmodified marxpileup.c + oracle/ref/pileup_replay.c) bit for bit
 * -- every column, every row -- and the stock program's statistics with its own RNG.
 *
 * Draw K (0, 1, ...) taken while frame F is processed = lane (K & 3) of Philox4x32-10 (key = seed, counter = (F, 0, K >> 2, 5)),
 * mapped to [0,1] as jdmath/src/random.c:151-154. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "marx_oracle.h"

#define NX 1024
#define NY 1024
#define MAX_CCDS 10

typedef struct { float benergy; unsigned int num_photons; float island_benergy; unsigned int island_num_photons; float x, y; } pixel_t;   /* :443-457 */

static void philox4x32_10 (uint32_t c[4], uint32_t k0, uint32_t k1)
{
   int i;
   for (i = 0; i < 10; i++)
     {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c[0], p1 = (uint64_t) 0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t) p1, n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t) p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
     }
}
static double frame_uniform (uint64_t seed, uint32_t frame, uint32_t k)
{
   uint32_t c[4];
   c[0] = frame; c[1] = 0; c[2] = k >> 2; c[3] = 5u;
   philox4x32_10 (c, (uint32_t) seed, (uint32_t) (seed >> 32));
   return (double) c[k & 3] * (1.0 / (double) 0xFFFFFFFFu);
}

/* columns of the input event files (read_input_event :573-620) and of the output files (write_event :622-666) */
typedef struct
{
   const int8_t *ccd; const float *x, *y, *t, *benergy;
   const float *dither[6];                  /* sky_ra, sky_dec, sky_roll, det_dy, det_dz, det_theta; NULL: simulation without dither */
}
pileup_in;
typedef struct
{
   int8_t *ccd; float *x, *y, *t, *benergy; int32_t *frame; int16_t *nphotons, *pha;
   float *dither[6];
}
pileup_out;

/* process_frame :890-922 for the events [first, first + n) of one frame; the reference's list holds them in REVERSE file order
 * (main :1148-1153 prepends).  Returns the number of output events appended at out[*n_out...], -1 on error. */
static int process_frame (oracle_t *o, pixel_t **maps, const pileup_in *in, uint64_t first, uint64_t n, uint32_t frame,
                          double alpha, double frame_time, uint64_t seed, pileup_out *out, uint64_t *n_out, uint64_t max_out)
{
   uint64_t *kept = (uint64_t *) malloc ((n ? n : 1) * sizeof (uint64_t));
   uint64_t nk = 0, a;
   uint32_t draws = 0;
   int status = 0;
   if (kept == NULL) return -1;
   /* store_event :754-812 (cent = 1, corn = side = 0: the neighbours receive + 0.0 * e) */
   for (a = 0; a < n; a++)
     {
        const uint64_t e = first + (n - 1 - a);
        const int ccd = in->ccd[e];
        const float ex = in->x[e], ey = in->y[e];
        const int center = ((ex >= 1) && (ex < NX - 1) && (ey >= 1) && (ey < NY - 1));            /* allocate_input_event :559-560 */
        pixel_t *cp;
        if ((ccd < 0) || (ccd >= MAX_CCDS)) { status = -1; break; }
        if ((center == 0) && ((ex >= NY) || (ey >= NY))) { status = -1; break; }                 /* "Corrupt file?" :562-568 */
        if (maps[ccd] == NULL)
          {
             maps[ccd] = (pixel_t *) calloc ((size_t) NX * NY, sizeof (pixel_t));
             if (maps[ccd] == NULL) { status = -1; break; }
          }
        if (!center) { kept[nk++] = e; continue; }                                               /* stays in the list, does nothing */
        cp = maps[ccd] + (unsigned int) ey * NX + (unsigned int) ex;
        {
           const int is_dup = (cp->num_photons != 0);
           const double en = in->benergy[e];                                                    /* USE_RMF_CODE 0 */
           const double cent_e = 1.0 * en;
           cp->num_photons += 1;
           cp->x = cp->x * cp->benergy + cent_e * ex;
           cp->y = cp->y * cp->benergy + cent_e * ey;
           cp->benergy += cent_e;
           cp->x /= cp->benergy;
           cp->y /= cp->benergy;
           if (!is_dup) kept[nk++] = e;                                                           /* duplicates leave the list :903-908 */
        }
     }
   /* collect_charge :814-845 */
   for (a = 0; (status == 0) && (a < nk); a++)
     {
        const uint64_t e = kept[a];
        const float ex = in->x[e], ey = in->y[e];
        pixel_t *xy0, *xy1, *xy2;
        if (!((ex >= 1) && (ex < NX - 1) && (ey >= 1) && (ey < NY - 1))) continue;
        xy0 = maps[in->ccd[e]] + ((unsigned int) ey - 1) * NX + ((unsigned int) ex - 1); xy1 = xy0 + NX; xy2 = xy1 + NX;
        xy1[1].island_benergy =
          (9.0 / 9.0) * xy0[0].benergy + (9.0 / 9.0) * xy0[1].benergy + (9.0 / 9.0) * xy0[2].benergy
          + (9.0 / 9.0) * xy1[0].benergy + (1.0) * xy1[1].benergy + (9.0 / 9.0) * xy1[2].benergy
          + (9.0 / 9.0) * xy2[0].benergy + (9.0 / 9.0) * xy2[1].benergy + (9.0 / 9.0) * xy2[2].benergy;
        xy1[1].island_num_photons =
          xy0[0].num_photons + xy0[1].num_photons + xy0[2].num_photons
          + xy1[0].num_photons + xy1[1].num_photons + xy1[2].num_photons
          + xy2[0].num_photons + xy2[1].num_photons + xy2[2].num_photons;
     }
   /* event_detect :676-752 */
   for (a = 0; (status == 0) && (a < nk); a++)
     {
        const uint64_t e = kept[a];
        const float ex = in->x[e], ey = in->y[e];
        pixel_t *xy0, *xy1, *xy2;
        unsigned int island_num_photons, i;
        double island_benergy, benergy, x, y;
        if (!((ex >= 1) && (ex < NX - 1) && (ey >= 1) && (ey < NY - 1))) continue;
        x = ex; y = ey;
        xy0 = maps[in->ccd[e]] + ((unsigned int) y - 1) * NX + ((unsigned int) x - 1); xy1 = xy0 + NX; xy2 = xy1 + NX;
        benergy = xy1[1].benergy;
        island_benergy = xy1[1].island_benergy;
        island_num_photons = xy1[1].island_num_photons;
        if (((xy0[0].benergy >= benergy) || (xy0[1].benergy >= benergy) || (xy0[2].benergy >= benergy))
            || (xy1[0].benergy > benergy) || (xy1[2].benergy >= benergy)
            || (xy2[0].benergy > benergy) || (xy2[1].benergy > benergy) || (xy2[2].benergy > benergy))
          continue;
        if (((xy0[0].island_benergy >= island_benergy) || (xy0[1].island_benergy >= island_benergy) || (xy0[2].island_benergy >= island_benergy))
            || (xy1[0].island_benergy > island_benergy) || (xy1[2].island_benergy >= island_benergy)
            || (xy2[0].island_benergy > island_benergy) || (xy2[1].island_benergy > island_benergy) || (xy2[2].island_benergy > island_benergy))
          continue;
        if (island_num_photons >= 2)
          {
             const double prob = pow (alpha, island_num_photons - 1);                             /* will_grade_migrate :668-674 */
             if (frame_uniform (seed, frame, draws++) >= prob) continue;
             x = 0; y = 0;
             for (i = 0; i < 3; i++)
               {
                  x += xy0[i].x * xy0[i].benergy; y += xy0[i].y * xy0[i].benergy;
                  x += xy1[i].x * xy1[i].benergy; y += xy1[i].y * xy1[i].benergy;
                  x += xy2[i].x * xy2[i].benergy; y += xy2[i].y * xy2[i].benergy;
               }
             x /= xy1[1].island_benergy;
             y /= xy1[1].island_benergy;
          }
        /* write_event :622-666 */
        {
           const float xpix = (float) x, ypix = (float) y, b = (float) island_benergy;
           short pha;
           int d;
           if (*n_out >= max_out) { status = -1; break; }
           if (-1 == oracle_map_energy_to_acis_pha (o, in->ccd[e], (int) xpix, (int) ypix, b, &pha)) { status = -1; break; }
           out->ccd[*n_out] = in->ccd[e]; out->x[*n_out] = xpix; out->y[*n_out] = ypix;
           out->frame[*n_out] = (int32_t) frame; out->t[*n_out] = (float) ((int32_t) frame * frame_time);
           out->nphotons[*n_out] = (int16_t) island_num_photons; out->pha[*n_out] = (int16_t) pha; out->benergy[*n_out] = b;
           for (d = 0; d < 6; d++) if (in->dither[d] && out->dither[d]) out->dither[d][*n_out] = in->dither[d][e];
           (*n_out)++;
        }
     }
   /* free_event_list :847-888: clear what the frame touched */
   for (a = 0; a < n; a++)
     {
        const uint64_t e = first + a;
        const float ex = in->x[e], ey = in->y[e];
        const int ccd = in->ccd[e];
        int i, j;
        if ((ccd < 0) || (ccd >= MAX_CCDS) || (maps[ccd] == NULL)) continue;
        if (!((ex >= 1) && (ex < NX - 1) && (ey >= 1) && (ey < NY - 1))) continue;              /* nothing was stored for it */
        for (j = -1; j <= 1; j++) for (i = -1; i <= 1; i++)
          memset (maps[ccd] + ((unsigned int) ey + j) * NX + ((unsigned int) ex + i), 0, sizeof (pixel_t));
     }
   free (kept);
   return status;
}

/* the loop of main :1121-1213.  frame_time = FrameTime + FrameTransferTime (initialize :1083-1084).  -> number of output events, -1 on error */
long long pileup_oracle_run (oracle_t *o, uint64_t n, const int8_t *ccd, const float *x, const float *y, const float *t, const float *benergy,
                             const float *const *dither, double alpha, double frame_time, uint64_t seed, uint64_t max_out,
                             int8_t *o_ccd, float *o_x, float *o_y, float *o_t, float *o_benergy, int32_t *o_frame, int16_t *o_nphotons,
                             int16_t *o_pha, float *const *o_dither)
{
   pixel_t *maps[MAX_CCDS];
   pileup_in in; pileup_out out;
   uint64_t first = 0, n_out = 0, e;
   int d, status = 0;
   memset (maps, 0, sizeof (maps));
   in.ccd = ccd; in.x = x; in.y = y; in.t = t; in.benergy = benergy;
   out.ccd = o_ccd; out.x = o_x; out.y = o_y; out.t = o_t; out.benergy = o_benergy; out.frame = o_frame; out.nphotons = o_nphotons; out.pha = o_pha;
   for (d = 0; d < 6; d++) { in.dither[d] = dither ? dither[d] : NULL; out.dither[d] = o_dither ? o_dither[d] : NULL; }
   while ((first < n) && (status == 0))
     {
        const uint32_t frame = (unsigned int) (t[first] / frame_time);                           /* read_input_event :617 */
        for (e = first + 1; e < n; e++) if ((unsigned int) (t[e] / frame_time) != frame) break;
        status = process_frame (o, maps, &in, first, e - first, frame, alpha, frame_time, seed, &out, &n_out, max_out);
        first = e;
     }
   for (d = 0; d < MAX_CCDS; d++) free (maps[d]);
   return (status == 0) ? (long long) n_out : -1;
}
