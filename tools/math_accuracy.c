/* math_accuracy.c -- developer check of the algorithms in marx_b200/csrc/mx_math.cuh, restated in plain C (fma from libm),
 * against the 80-bit long-double functions of glibc.  gcc -O2 -mfma tools/math_accuracy.c -o /tmp/math_accuracy -lm */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
static const double kTrig[16] = {
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
   -2.50507602534068634195e-08, 1.58969099521155010221e-10,
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
   2.08757232129817482790e-09, -1.13596475577881948265e-11,
   1.5707963267948966e+00, 6.1232339957367574e-17, 8.4784276603688985e-32, 6.36619772367581382433e-01};
static const double kLog[10] = {
   6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
   1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
   6.93147180369123816490e-01, 1.90821492927058770002e-10, 0.0};
static double trig_reduce (double x, long long *k)
{
   const double q = rint (x * kTrig[15]);
   *k = (long long) q;
   double r = fma (-q, kTrig[12], x);
   r = fma (-q, kTrig[13], r);
   return fma (-q, kTrig[14], r);
}
static double sin_kernel (double r, double z)
{
   double p = fma (z, kTrig[5], kTrig[4]);
   p = fma (z, p, kTrig[3]); p = fma (z, p, kTrig[2]); p = fma (z, p, kTrig[1]); p = fma (z, p, kTrig[0]);
   return fma (z * r, p, r);
}
static double cos_kernel (double z)
{
   double p = fma (z, kTrig[11], kTrig[10]);
   p = fma (z, p, kTrig[9]); p = fma (z, p, kTrig[8]); p = fma (z, p, kTrig[7]); p = fma (z, p, kTrig[6]);
   return fma (z, fma (z, p, -0.5), 1.0);
}
static void mx_sincos (double x, double *s, double *c)
{
   long long k;
   const double r = trig_reduce (x, &k), z = r * r;
   const double sn = sin_kernel (r, z), cs = cos_kernel (z);
   const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
   *s = (k & 2) ? -a : a;
   *c = ((k + 1) & 2) ? -b : b;
}
static void tiny_sincos (double x, double *s, double *c)
{
   const double zx = x * x;
   *s = fma (zx * x, fma (zx, 1.0 / 120.0, -1.0 / 6.0), x);
   *c = fma (zx, fma (zx, fma (zx, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
}
/* rcp.approx.ftz.f64: a reciprocal seed good to about 20 bits; modelled by a float reciprocal */
static double rcp_seed (double d) { return (double) (1.0f / (float) d); }
static double mx_log (double x)
{
   uint64_t bits; memcpy (&bits, &x, 8);
   int hx = (int) (bits >> 32); uint32_t lx = (uint32_t) bits;
   int k = (hx >> 20) - 1023;
   hx &= 0x000fffff;
   const int i = (hx + 0x95f64) & 0x100000;
   k += (i >> 20);
   bits = ((uint64_t) (uint32_t) (hx | (i ^ 0x3ff00000)) << 32) | lx;
   double m; memcpy (&m, &bits, 8);
   const double f = m - 1.0, d = 2.0 + f;
   double y = rcp_seed (d);
   double e = fma (-d, y, 1.0);
   y = fma (y, e, y);
   e = fma (-d, y, 1.0);
   y = fma (y, e, y);
   double s = f * y;
   s = fma (fma (-d, s, f), y, s);
   const double z = s * s, w = z * z;
   const double t1 = w * fma (w, fma (w, kLog[5], kLog[3]), kLog[1]);
   const double t2 = z * fma (w, fma (w, fma (w, kLog[6], kLog[4]), kLog[2]), kLog[0]);
   const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double) k;
   return fma (dk, kLog[7], f - (hfsq - fma (s, hfsq + R, dk * kLog[8])));
}
static double ulp_err (double got, long double want)
{
   double w = (double) want; int e; frexp (w, &e);
   long double ulp = ldexpl (1.0L, e - 53);
   if (w == 0.0) return (got == 0.0) ? 0.0 : 1e9;
   return (double) (fabsl ((long double) got - want) / ulp);
}
int main (int argc, char **argv)
{
   long n = (argc > 1) ? atol (argv[1]) : 4000000;
   double ranges[] = {0.8, 7.0, 700.0, 1e5, 1e7, 1e9};
   double worst = 0.0;
   srand48 (1);
   for (int r = 0; r < 6; r++)
     {
        double ms = 0, mc = 0;
        for (long i = 0; i < n; i++)
          {
             double x = (2 * drand48 () - 1) * ranges[r], s, c;
             mx_sincos (x, &s, &c);
             double es = ulp_err (s, sinl ((long double) x)), ec = ulp_err (c, cosl ((long double) x));
             if (es > ms) ms = es; if (ec > mc) mc = ec;
          }
        printf ("sincos |x| < %g: sin %.3f ulp, cos %.3f ulp\n", ranges[r], ms, mc);
        if (ms > worst) worst = ms; if (mc > worst) worst = mc;
     }
   {
      double ms = 0, mc = 0;
      for (long i = 0; i < n; i++)
        {
           double x = (2 * drand48 () - 1) * 0x1p-10, s, c;
           tiny_sincos (x, &s, &c);
           double es = ulp_err (s, sinl ((long double) x)), ec = ulp_err (c, cosl ((long double) x));
           if (es > ms) ms = es; if (ec > mc) mc = ec;
        }
      printf ("tiny   |x| < 2^-10: sin %.3f ulp, cos %.3f ulp\n", ms, mc);
      if (ms > worst) worst = ms; if (mc > worst) worst = mc;
   }
   {
      double ml = 0, m2 = 0;
      for (long i = 0; i < n; i++)
        {
           double x = (i & 1) ? drand48 () : ldexp (drand48 (), -(int) (lrand48 () % 40));
           if (x <= 0.0) continue;
           double e = ulp_err (mx_log (x), logl ((long double) x));
           if (e > ml) ml = e;
           /* the draws are u32 / 4294967295 */
           double u = (double) (uint32_t) (1 + (uint32_t) (mrand48 () & 0x7fffffff) * 2u) * (1.0 / 4294967295.0);
           e = ulp_err (mx_log (u), logl ((long double) u));
           if (e > m2) m2 = e;
        }
      printf ("log    (0, 1]: %.3f ulp, on draws %.3f ulp\n", ml, m2);
      if (ml > worst) worst = ml; if (m2 > worst) worst = m2;
   }
   printf ("worst %.3f ulp\n", worst);
   return (worst <= 1.6) ? 0 : 1;
}
