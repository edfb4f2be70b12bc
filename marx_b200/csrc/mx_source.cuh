// mx_source.cuh -- K0: source draw, Poisson arrival increment, aspect dither (per ray).
// Reference: marx/libsrc/source.c:268-384, s-point.c:59-83, spectrum.c:138-147,184-189, prob.c:46-58,
// dither.c:167-182 (get_internal_dither), :551-581 (apply_dither), :609-628 (dither_ray).
// Draw order on sub-stream MARXB200_STAGE_SOURCE: energy U; [direction draws];
// time E; dither 2 G.
#pragma once
#include "mx_common.cuh"
#include "mx_tables.h"

namespace mx {

// energy + direction.  POINT source: direction is the source vector (s-point.c:76).
MX_HD void source_draw (const SourceDev &s, Rng &rng, double &energy, Vec3 &p)
{
   if (s.spectrum_type == 2)      // MARX_FILE_SPECTRUM: inverse CDF (prob.c:55-56)
     {
        double r = rng.uniform ();
        energy = interp_d (r, s.spec_cum_flux, s.spec_energies, s.spec_num);
     }
   else                           // MARX_FLAT_SPECTRUM (spectrum.c:140-145)
     {
        double emin = s.emin;
        double de = s.emax - emin;
        energy = emin + de * rng.uniform ();
     }
   p = v_make (s.p[0], s.p[1], s.p[2]);
   if (s.source_type == 0) return;
   if (s.source_type >= 4)
     {
        // LINE / IMAGE: build the ray about (-1,0,0), then rotate it onto the source direction
        Vec3 q;
        if (s.source_type == 4)   // s-line.c:86-98
          {
             double theta = s.shape[0] * (-1.0 + 2.0 * rng.uniform ());
             double sn, cs;
             sin_cos (theta, sn, cs);
             double sin_theta = -sn;
             q = v_make (-cs, sin_theta * s.shape[1], sin_theta * s.shape[2]);
          }
        else                      // s-image.c:330-356
          {
             uint32_t ofs = bsearch_f ((float) rng.uniform (), s.image_cdf, s.image_size);
             double y = (double) (ofs / s.image_nx);
             double x = (double) (ofs % s.image_nx);
             y += -0.5 * s.image_ny + (rng.uniform () - 0.5);
             x += -0.5 * s.image_nx + (rng.uniform () - 0.5);
             y = y * s.rad_per_ypixel;
             x = x * s.rad_per_xpixel;
             double sy, cy, sx, cx;
             sin_cos (y, sy, cy);
             sin_cos (x, sx, cx);
             q = v_make (-cy * cx, cy * sx, -sy);
          }
        p = v_rotate_unit (q, v_make (s.rot_axis[0], s.rot_axis[1], s.rot_axis[2]), s.rot_angle);
        return;
     }
   // GAUSS / BETA / DISK share one construction (s-gauss.c:96-137): rotate the normal to p about p by a
   // uniform angle, then rotate p about that normal by a source-specific polar angle.  The reference keeps
   // rotating ONE normal from photon to photon (statistically a fresh uniform azimuth every time); per-ray
   // draws restart from st->p_normal, which is what the reference does at the start of every batch.
   Vec3 normal = v_make (s.p_normal[0], s.p_normal[1], s.p_normal[2]);
   normal = v_rotate_unit (normal, p, 2.0 * kPI * rng.uniform ());
   double theta;
   if (s.source_type == 3)        // DISK, s-disk.c:104
     theta = s.shape[0] * sqrt (s.shape[1] + s.shape[2] * rng.uniform ());
   else
     {
        double rnd;
        do rnd = rng.uniform (); while (rnd == 0.0);
        if (s.source_type == 1) theta = s.shape[0] * sqrt (-log (rnd));                 // GAUSS, s-gauss.c:126-136
        else theta = s.shape[0] * sqrt (pow (rnd, s.shape[1]) - 1.0);                  // BETA, s-beta.c:116-125
     }
   p = v_rotate_unit (p, normal, theta);
}

// arrival-time increment: mt * Exp(1)  (source.c:326)
MX_HD double source_time_increment (const SourceDev &s, Rng &rng)
{
   return s.mean_time * rng.expn ();
}

// apply_dither, dither.c:551-581.  p_rolled: p after the roll about the x axis (first statement of apply_dither);
// for a POINT source with DitherAmp_Roll = 0 it is the same vector for every ray, so persistent kernels compute
// it once per thread with the same operations (dither_roll_is_constant / dither_roll).
MX_HD Vec3 dither_roll (double roll, const Vec3 &p) { return v_rotate_unit (p, v_make (1, 0, 0), -roll); }
MX_HD Vec3 apply_dither_rolled (double ra, double dec, Vec3 p)
{
   double cos_ra, sin_ra, cos_dec, sin_dec;
   mx_sincos_pair (ra, dec, sin_ra, cos_ra, sin_dec, cos_dec);      // dither angles: a few 1e-5 rad
   double cos_theta = cos_dec * cos_ra;
   Vec3 n = v_make (0, sin_dec, -cos_dec * sin_ra);
   double sin_theta = v_length (n);
   if (sin_theta <= 1e-20) return p;
   // n.x is +0 and sin_theta > 0: the reference's n.x / sin_theta is +0 again (and a zero quotient is the division's slow case)
   n.y /= sin_theta; n.z /= sin_theta;
   return v_rotate_unit1 (p, n, cos_theta, sin_theta);
}
MX_HD Vec3 apply_dither (double ra, double dec, double roll, Vec3 p)
{
   return apply_dither_rolled (ra, dec, dither_roll (roll, p));
}
MX_HD bool dither_roll_is_constant (const SourceDev &s, const DitherDev &d)
{
   return (d.mode == 1) && (d.roll_amp == 0.0) && (s.source_type == 0);
}

// dither_ray + get_internal_dither.  t is the absolute time (pt->start_time + arrival_time).
// The three angles are stored through float fields before use (dither.c:173-175); that rounding is
// part of the result.
// rolled != nullptr: the caller supplies dither_roll ((float) nominal_roll, source p) (see above)
// det: optional, receives the detector dither dy, dz, dtheta (0 unless the ASPSOL model supplies them)
MX_HD void dither_ray (const DitherDev &d, Rng &rng, double t, Vec3 &p, float &f_ra, float &f_dec, float &f_roll,
                       const Vec3 *rolled = nullptr, float *det = nullptr)
{
   if (det != nullptr) { det[0] = 0.0f; det[1] = 0.0f; det[2] = 0.0f; }
   if (d.mode == 0) { f_ra = f_dec = f_roll = 0.0f; return; }
   if (d.mode == 2)
     {
        // get_aspsol_dither, dither.c:379-400: the stock reader advances `while (t >= t1)`; photon times never decrease,
        // so its position is the first state k >= 1 with t < t_k.  (t at or beyond the last state: the reference stops
        // the simulation there; marxb200_create_photons cuts the batch, the clamp only keeps the access in range.)
        const double *A = d.aspsol;
        uint32_t lo = 1, hi = d.num_aspsol - 1;
        while (lo < hi)
          {
             const uint32_t mid = lo + (hi - lo) / 2;
             if (t < A[7 * mid]) hi = mid; else lo = mid + 1;
          }
        const double *s0 = A + 7 * (lo - 1), *s1 = A + 7 * lo;
        double dt = s1[0] - s0[0];
        if (dt != 0) dt = (t - s0[0]) / dt;
        f_ra = (float) (s0[1] + dt * (s1[1] - s0[1]));
        f_dec = (float) (s0[2] + dt * (s1[2] - s0[2]));
        f_roll = (float) (s0[3] + dt * (s1[3] - s0[3]));
        if (det != nullptr)
          {
             det[0] = (float) (s0[4] + dt * (s1[4] - s0[4]));
             det[1] = (float) (s0[5] + dt * (s1[5] - s0[5]));
             det[2] = (float) (s0[6] + dt * (s1[6] - s0[6]));
          }
     }
   else
     {
        t = (2.0 * kPI) * t;
        // amp * sin(..) is exactly +-0 when the amplitude is 0 (default DitherAmp_Roll): skip the sine then
        f_ra = (d.ra_amp == 0.0) ? 0.0f : (float) (d.ra_amp * mx_sin (t / d.ra_period + d.ra_phase));
        f_dec = (d.dec_amp == 0.0) ? 0.0f : (float) (d.dec_amp * mx_sin (t / d.dec_period + d.dec_phase));
        f_roll = (d.roll_amp == 0.0) ? (float) d.nominal_roll
                                     : (float) (d.nominal_roll + d.roll_amp * mx_sin (t / d.roll_period + d.roll_phase));
     }
   double ra = f_ra, dec = f_dec, roll = f_roll;
   double delta_ra = d.aspect_blur * rng.gaussian ();
   double delta_dec = d.aspect_blur * rng.gaussian ();
   if (rolled != nullptr) p = apply_dither_rolled (ra + delta_ra, dec + delta_dec, *rolled);
   else p = apply_dither (ra + delta_ra, dec + delta_dec, roll, p);
}

}  // namespace mx
