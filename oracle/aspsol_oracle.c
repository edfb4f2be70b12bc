/* aspsol_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of the row loop of marxasp (marx/src/marxasp.c), the
 * aspect-solution table of MARX's INTERNAL dither model.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may use it; the product (marx_b200/) never does.
 *
 * Parity PINNED: tests/test_aspsol_oracle_vs_reference.py requires these functions to reproduce the ASPSOL table written by the
 * stock marxasp (oracle/_ref/marxasp) bit for bit -- every column, every row -- on the committed fixtures
 * (tests/golden/aspsol_*.npz) and on fresh runs where oracle/_ref exists.
 *
 * desc[21] = time_start, delta_time, ra_amp, dec_amp, roll_amp (radians), ra/dec/roll period, ra/dec/roll phase,
 *            nominal roll (radians), nominal pointing[3], ra_hat[3], dec_hat[3]   (marxasp.c: setup_dither :387-410) */
#include <math.h>
#include <stdint.h>

#define PI 3.14159265358979323846

typedef struct { double x, y, z; } vec3;

/* JDMv_ax1_bx2_cx3, jdmath/src/vector.c:132-143 */
static vec3 ax1_bx2_cx3 (double a, vec3 x1, double b, vec3 x2, double c, vec3 x3)
{
   vec3 d;
   d.x = a * x1.x + b * x2.x + c * x3.x;
   d.y = a * x1.y + b * x2.y + c * x3.y;
   d.z = a * x1.z + b * x2.z + c * x3.z;
   return d;
}
/* JDMv_length, vector.c:77-98 (scaled by the largest component) */
static double vec_length (vec3 a)
{
   double x = fabs (a.x), y = fabs (a.y), z = fabs (a.z), tmp;
   if (z < x) { tmp = z; z = x; x = tmp; }
   if (z < y) { tmp = z; z = y; y = tmp; }
   if (z == 0.0) return 0.0;
   x = x / z; y = y / z;
   z = z * sqrt (1.0 + x * x + y * y);
   return z;
}
/* JDMv_rotate_unit_vector, vector.c:186-212 (rotate_vector1 + JDMv_normalize :99-109) */
static vec3 rotate_unit_vector (vec3 p, vec3 n, double theta)
{
   double cos_theta = cos (theta), sin_theta = sin (theta);
   double pn = p.x * n.x + p.y * n.y + p.z * n.z;
   vec3 nxp, u;
   double len;
   nxp.x = n.y * p.z - n.z * p.y;
   nxp.y = n.z * p.x - n.x * p.z;
   nxp.z = n.x * p.y - n.y * p.x;
   u = ax1_bx2_cx3 (cos_theta, p, pn * (1.0 - cos_theta), n, sin_theta, nxp);
   len = vec_length (u);
   if (len != 0.0) { u.x = u.x / len; u.y = u.y / len; u.z = u.z / len; }
   return u;
}
/* JDMv_unit_vector_to_spherical, vector.c:214-250 */
static void unit_vector_to_spherical (vec3 p, double *thetap, double *phip)
{
   double theta, phi, sin_theta;
   if (fabs (p.z) >= 1.0)
     {
        *thetap = (p.z >= 1.0) ? 0 : PI;
        *phip = 0;
        return;
     }
   theta = acos (p.z);
   sin_theta = sin (theta);
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
             if (phi >= 0) phi = PI - phi;
             else phi = -PI - phi;
          }
     }
   *thetap = theta;
   *phip = phi;
}

/* compute_dither, marxasp.c:814-884: ra, dec, roll in degrees at time t since TSTART */
static void compute_dither (const double *d, double t, double *rap, double *decp, double *rollp)
{
   double ra, dec, roll, cos_dec;
   vec3 pointing = {d[12], d[13], d[14]}, ra_hat = {d[15], d[16], d[17]}, dec_hat = {d[18], d[19], d[20]}, p;
   t = (2.0 * PI) * t;
   ra = d[2] * sin (t / d[5] + d[8]);
   dec = d[3] * sin (t / d[6] + d[9]);
   roll = d[4] * sin (t / d[7] + d[10]);
   cos_dec = cos (dec);
   p = ax1_bx2_cx3 (cos (ra) * cos_dec, pointing, cos_dec * sin (ra), ra_hat, sin (dec), dec_hat);
   roll += d[11];
   p = rotate_unit_vector (p, pointing, roll);
   unit_vector_to_spherical (p, &dec, &ra);
   dec = PI / 2 - dec;
   ra *= 180.0 / PI;
   dec *= 180.0 / PI;
   roll *= 180.0 / PI;
   if (ra < 0) ra += 360.0;
   if (roll < 0) roll += 360.0;
   if (dec > 180) dec -= 360;
   else if (dec < -180) dec += 360;
   if (dec >= 0)
     {
        if (dec > 90) dec = 180 - dec;
     }
   else if (dec < -90) dec = -180 - dec;
   *rap = ra; *decp = dec; *rollp = roll;
}

/* compute_quaternion, marxasp.c:886-903 */
static void compute_quaternion (double ra, double dec, double roll, double q[4])
{
   double cos_ra, cos_dec, cos_roll, sin_ra, sin_dec, sin_roll, q0, q1, q2, q3, len;
   ra *= PI / 360; dec *= PI / 360; roll = (180.0 - roll) * PI / 360.0;
   cos_ra = cos (ra); cos_dec = cos (dec); cos_roll = cos (roll);
   sin_ra = sin (ra); sin_dec = sin (dec); sin_roll = sin (roll);
   q0 = cos_ra * cos_dec * cos_roll + sin_ra * sin_dec * sin_roll;
   q1 = sin_ra * cos_dec * cos_roll - cos_ra * sin_dec * sin_roll;
   q2 = cos_ra * sin_dec * cos_roll + sin_ra * cos_dec * sin_roll;
   q3 = cos_ra * cos_dec * sin_roll - sin_ra * sin_dec * cos_roll;
   len = sqrt (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
   q[0] = q0 / len; q[1] = q1 / len; q[2] = q2 / len; q[3] = q3 / len;
}

/* the row loop of write_marxasp, marxasp.c:996-1027.  cols = 8 arrays of n doubles: time, ra, dec, roll, q0..q3
 * (dy, dz, dtheta are identically 0, :880-882) */
int aspsol_oracle_rows (const double *desc, uint64_t first_row, uint64_t n, double *cols)
{
   uint64_t k;
   for (k = 0; k < n; k++)
     {
        unsigned int i = (unsigned int) (first_row + k);
        double t = i * desc[1], ra, dec, roll, q[4];
        compute_dither (desc, t, &ra, &dec, &roll);
        compute_quaternion (ra, dec, roll, q);
        cols[0 * n + k] = t + desc[0];
        cols[1 * n + k] = ra; cols[2 * n + k] = dec; cols[3 * n + k] = roll;
        cols[4 * n + k] = q[0]; cols[5 * n + k] = q[1]; cols[6 * n + k] = q[2]; cols[7 * n + k] = q[3];
     }
   return 0;
}
