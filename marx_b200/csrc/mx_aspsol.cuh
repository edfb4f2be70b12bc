// mx_aspsol.cuh -- per-row device functions of the aspect-solution table (marxasp, marx/src/marxasp.c):
// the ASPSOL file a user feeds to the CIAO tools together with the Level-1 events of an INTERNAL-dither simulation.
// One row per 0.256 s of exposure, each independent: pointing offsets of the dither model at the row's time -> absolute
// RA / Dec / roll in degrees -> attitude quaternion.
#pragma once
#include "mx_common.cuh"

namespace mx {

// include/marxb200.h: marxb200_aspsol_desc (same 21 doubles)
struct AspsolDev
{
   double time_start, delta_time;
   double amp[3], period[3], phase[3];      // ra, dec, roll; amplitudes in radians (setup_dither, marxasp.c:391-393)
   double nominal_roll;                     // radians
   double pointing[3], ra_hat[3], dec_hat[3];
};

// JDMv_unit_vector_to_spherical, jdmath/src/vector.c:214-250
MX_HD void unit_vector_to_spherical (const Vec3 &p, double &theta_out, double &phi_out)
{
   if (fabs (p.z) >= 1.0)
     {
        theta_out = (p.z >= 1.0) ? 0.0 : kPI;
        phi_out = 0.0;
        return;
     }
   double theta = acos (p.z);
   double sin_theta = sin (theta);
   double phi;
   if (fabs (p.x) <= fabs (p.y))
     {
        phi = acos (p.x / sin_theta);
        if (p.y < 0.0) phi = -phi;
     }
   else
     {
        phi = asin (p.y / sin_theta);
        if (p.x < 0)
          {
             if (phi >= 0) phi = kPI - phi;
             else phi = -kPI - phi;
          }
     }
   theta_out = theta;
   phi_out = phi;
}

// compute_dither, marxasp.c:814-884: t = seconds since TSTART; ra, dec, roll in degrees
MX_HD void aspsol_dither (const AspsolDev &D, double t, double &ra_out, double &dec_out, double &roll_out)
{
   const Vec3 pointing = v_make (D.pointing[0], D.pointing[1], D.pointing[2]);
   t = (2.0 * kPI) * t;
   double ra = D.amp[0] * sin (t / D.period[0] + D.phase[0]);
   double dec = D.amp[1] * sin (t / D.period[1] + D.phase[1]);
   double roll = D.amp[2] * sin (t / D.period[2] + D.phase[2]);
   double sin_ra, cos_ra, sin_dec, cos_dec;
   sin_cos (ra, sin_ra, cos_ra);
   sin_cos (dec, sin_dec, cos_dec);
   Vec3 p = v_ax1_bx2_cx3 (cos_ra * cos_dec, pointing,
                           cos_dec * sin_ra, v_make (D.ra_hat[0], D.ra_hat[1], D.ra_hat[2]),
                           sin_dec, v_make (D.dec_hat[0], D.dec_hat[1], D.dec_hat[2]));
   roll += D.nominal_roll;
   p = v_rotate_unit (p, pointing, roll);
   unit_vector_to_spherical (p, dec, ra);
   dec = kPI / 2 - dec;
   ra *= 180.0 / kPI;
   dec *= 180.0 / kPI;
   roll *= 180.0 / kPI;
   if (ra < 0) ra += 360.0;
   if (roll < 0) roll += 360.0;
   if (dec > 180) dec -= 360;
   else if (dec < -180) dec += 360;
   if (dec >= 0)
     {
        if (dec > 90) dec = 180 - dec;
     }
   else if (dec < -90) dec = -180 - dec;
   ra_out = ra; dec_out = dec; roll_out = roll;
}

// compute_quaternion, marxasp.c:886-903
MX_HD void aspsol_quaternion (double ra, double dec, double roll, double q[4])
{
   ra *= kPI / 360; dec *= kPI / 360; roll = (180.0 - roll) * kPI / 360.0;
   double cos_ra, cos_dec, cos_roll, sin_ra, sin_dec, sin_roll;
   sin_cos (ra, sin_ra, cos_ra);
   sin_cos (dec, sin_dec, cos_dec);
   sin_cos (roll, sin_roll, cos_roll);
   double q0 = cos_ra * cos_dec * cos_roll + sin_ra * sin_dec * sin_roll;
   double q1 = sin_ra * cos_dec * cos_roll - cos_ra * sin_dec * sin_roll;
   double q2 = cos_ra * sin_dec * cos_roll + sin_ra * cos_dec * sin_roll;
   double q3 = cos_ra * cos_dec * sin_roll - sin_ra * sin_dec * cos_roll;
   double len = sqrt (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
   q[0] = q0 / len; q[1] = q1 / len; q[2] = q2 / len; q[3] = q3 / len;
}

// one row of write_marxasp's loop, marxasp.c:996-1010: v[0..7] = time, ra, dec, roll, q0..q3 (dy = dz = dtheta = 0, :880-882)
MX_HD void aspsol_row (const AspsolDev &D, uint64_t row, double v[8])
{
   const double t = (double) (unsigned int) row * D.delta_time;
   aspsol_dither (D, t, v[1], v[2], v[3]);
   v[0] = t + D.time_start;
   aspsol_quaternion (v[1], v[2], v[3], v + 4);
}

constexpr int kAspsolRowWords = 19;        // FITS row: 4 x f64 + 3 x f32 + 4 x f64 = 76 bytes

void launch_aspsol_rows (const AspsolDev &D, uint64_t first_row, uint64_t n, double *cols, uint32_t *fits_rows, int num_sms, cudaStream_t s);

}  // namespace mx
