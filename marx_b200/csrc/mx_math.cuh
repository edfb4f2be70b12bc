// mx_math.cuh -- sin / cos / log for the stage kernels with their coefficients in CONSTANT memory.
//
// Why: every kernel of this path is bound by instruction issue, not by the FP64 pipe (ncu: 65 % issue slots, 39 % FP64 pipe in
// k01_source_hrma; only 26 % of its instructions are DFMA / DMUL / DADD).  libdevice's sin / cos / sincos / log are inlined
// with their polynomial coefficients as 64-bit literals, which sm_100 materialises with TWO UMOV per coefficient in front of
// every DFMA (461 UMOV + 487 IMAD.MOV of k01's 4 000 SASS instructions, 11 % of those it executes), and sincos evaluates
// through a generic quadrant-select sequence (~100 instructions per call on the fast path).  The same Cody-Waite reduction
// and minimax kernels with the coefficients in __constant__ memory load two coefficients with one LDCU.128 and take ~45
// instructions per sincos; arguments below 2^-10 (the aspect-dither angles: 16 arcsec = 8e-5 rad) need three terms.
//
// Accuracy (tools/math_accuracy.c: the same algorithms in plain C against 80-bit sinl / cosl / logl on 4e6 random arguments per range): sin and cos <= 1.5 ulp
// for |x| <= 1e9, <= 0.5 ulp below 2^-10; log <= 1 ulp on (0, 1].  libdevice documents 2 ulp (sin, cos) and 1 ulp (log); the
// reference's glibc is at 0.52 ulp.  Replay parity is judged at 1e-9 relative, integer outputs bit-exact: the differences
// are of the size the libdevice-vs-glibc differences already were.  Arguments outside the fast range, NaN and Inf take the
// libdevice function.  -DMX_MATH=0 restores libdevice everywhere (A/B builds, tools/build_variant.sh).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef MX_MATH
#define MX_MATH 2
#endif

#if defined(__CUDACC__)
#define MXM_HD __host__ __device__ __forceinline__
#else
#define MXM_HD inline          // host builds of the MX_HD functions (tools/hostcheck) take libm
#endif

namespace mx {

#if defined(__CUDACC__)
// [0..5] sin kernel S1..S6, [6..11] cos kernel C1..C6 on [-pi/4, pi/4] (the minimax polynomials of fdlibm's k_sin.c / k_cos.c),
// [12..14] pi/2 in three parts (Cody-Waite with FMA), [15] 2/pi
static __constant__ double kTrig[16] = {
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
   -2.50507602534068634195e-08, 1.58969099521155010221e-10,
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
   2.08757232129817482790e-09, -1.13596475577881948265e-11,
   1.5707963267948966e+00, 6.1232339957367574e-17, 8.4784276603688985e-32, 6.36619772367581382433e-01};
// [0..6] Lg1..Lg7 of fdlibm's e_log.c, [7] ln2_hi, [8] ln2_lo
static __constant__ double kLog[10] = {
   6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
   1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
   6.93147180369123816490e-01, 1.90821492927058770002e-10, 0.0};

// quadrant k = rint (x * 2/pi), remainder r = x - k pi/2 in [-pi/4, pi/4]; valid for |x| <= 1e9 (the quotient fits an int; tools/math_accuracy.c: <= 1.5 ulp up to there)
__device__ __forceinline__ double trig_reduce (double x, int &k)
{
   const double q = rint (x * kTrig[15]);
   k = (int) q;
   double r = fma (-q, kTrig[12], x);
   r = fma (-q, kTrig[13], r);
   return fma (-q, kTrig[14], r);
}
__device__ __forceinline__ double sin_kernel (double r, double z)
{
   double p = fma (z, kTrig[5], kTrig[4]);
   p = fma (z, p, kTrig[3]); p = fma (z, p, kTrig[2]); p = fma (z, p, kTrig[1]); p = fma (z, p, kTrig[0]);
   return fma (z * r, p, r);
}
__device__ __forceinline__ double cos_kernel (double z)
{
   double p = fma (z, kTrig[11], kTrig[10]);
   p = fma (z, p, kTrig[9]); p = fma (z, p, kTrig[8]); p = fma (z, p, kTrig[7]); p = fma (z, p, kTrig[6]);
   return fma (z, fma (z, p, -0.5), 1.0);
}
// arguments beyond the fast range (never on the benchmark path): ONE out-of-line copy of the libdevice functions per kernel
// instead of an inlined Payne-Hanek reduction at every call site
static __device__ __noinline__ double2 sincos_far (double x) { double2 r; sincos (x, &r.x, &r.y); return r; }   // by value: no stack slot at the call sites
static __device__ __noinline__ double sin_far (double x) { return sin (x); }
static __device__ __noinline__ double log_far (double x) { return log (x); }
#endif

// sin and cos of one angle
MXM_HD void mx_sincos (double x, double &s, double &c)
{
#if defined(__CUDA_ARCH__) && (MX_MATH >= 1)
   if (fabs (x) <= 1.0e9)
     {
        int k;
        const double r = trig_reduce (x, k), z = r * r;
        const double sn = sin_kernel (r, z), cs = cos_kernel (z);
        const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
        s = (k & 2) ? -a : a;
        c = ((k + 1) & 2) ? -b : b;
        return;
     }
   { const double2 r = sincos_far (x); s = r.x; c = r.y; }
#elif defined(__CUDA_ARCH__)
   sincos (x, &s, &c);
#else
   s = sin (x); c = cos (x);
#endif
}
// the same for two small angles at once (the dithered pointing offsets): below 2^-10 rad the series needs three terms
// (the next ones are below 2^-60 relative)
MXM_HD void mx_sincos_pair (double x, double y, double &sx, double &cx, double &sy, double &cy)
{
#if defined(__CUDA_ARCH__) && (MX_MATH >= 1)
   if ((fabs (x) < 0x1p-10) && (fabs (y) < 0x1p-10))
     {
        const double zx = x * x, zy = y * y;
        sx = fma (zx * x, fma (zx, 1.0 / 120.0, -1.0 / 6.0), x);
        sy = fma (zy * y, fma (zy, 1.0 / 120.0, -1.0 / 6.0), y);
        cx = fma (zx, fma (zx, fma (zx, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
        cy = fma (zy, fma (zy, fma (zy, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
        return;
     }
#endif
   mx_sincos (x, sx, cx);
   mx_sincos (y, sy, cy);
}
MXM_HD double mx_sin (double x)
{
#if defined(__CUDA_ARCH__) && (MX_MATH >= 1)
   if (fabs (x) <= 1.0e9)
     {
        int k;
        const double r = trig_reduce (x, k), z = r * r;
        const double a = (k & 1) ? cos_kernel (z) : sin_kernel (r, z);
        return (k & 2) ? -a : a;
     }
   return sin_far (x);
#else
   return sin (x);
#endif
}

// natural logarithm of a positive normal number (the draws: (0, 1]); anything else takes the libdevice function.
// fdlibm's scheme: x = 2^k (1 + f), sqrt(2)/2 <= 1 + f < sqrt(2); s = f / (2 + f); log (1 + f) = f - f^2/2 + s (f^2/2 + R (s^2)).
// The quotient comes from the hardware reciprocal seed and two Newton steps (<= 1 ulp; its error enters the result scaled by s^2).
MXM_HD double mx_log (double x)
{
#if defined(__CUDA_ARCH__) && (MX_MATH >= 2)
   int hx = __double2hiint (x);
   if ((hx >= 0x00100000) && (hx < 0x7ff00000))
     {
        int k = (hx >> 20) - 1023;
        hx &= 0x000fffff;
        const int i = (hx + 0x95f64) & 0x100000;          // mantissa above sqrt(2): halve it, k + 1
        k += (i >> 20);
        const double m = __hiloint2double (hx | (i ^ 0x3ff00000), __double2loint (x));
        const double f = m - 1.0, d = 2.0 + f;
        double y;
        asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        double e = fma (-d, y, 1.0);
        y = fma (y, e, y);
        e = fma (-d, y, 1.0);
        y = fma (y, e, y);
        double s = f * y;
        s = fma (fma (-d, s, f), y, s);                   // one residual correction of the quotient
        const double z = s * s, w = z * z;
        const double t1 = w * fma (w, fma (w, kLog[5], kLog[3]), kLog[1]);
        const double t2 = z * fma (w, fma (w, fma (w, kLog[6], kLog[4]), kLog[2]), kLog[0]);
        const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double) k;
        return fma (dk, kLog[7], f - (hfsq - fma (s, hfsq + R, dk * kLog[8])));
     }
   return log_far (x);
#else
   return log (x);
#endif
}

}  // namespace mx
