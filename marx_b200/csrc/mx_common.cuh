// mx_common.cuh -- scalar building blocks shared by every stage kernel: 3-vectors, the
// counter-based draw stream, float-table interpolation.
//
// Each routine restates the jdmath primitive the reference's inner loops call with
// the SAME operation order and the same float/double narrowing points, because replay parity is
// judged at 1e-9 relative (FP64 geometry) and bit-exact for integer outputs.  The translation unit is
// compiled with -fmad=false so that nvcc does not contract a*b+c into FMAs the reference (gcc -O2,
// x86-64 SSE2) never executes.
//
// MX_HD functions are host+device so that tools/hostcheck (a developer-only harness, never part of
// the product path) can step the very same code under a debugger; the library itself only ever calls
// them from __global__ kernels.
#pragma once
#include <stdint.h>
#include <math.h>
#include "mx_math.cuh"

#if defined(__CUDACC__)
#define MX_HD __host__ __device__ __forceinline__
// Large helpers called from several places.  ncu attributes ~25 % of k1_hrma<1>'s stall samples to "no
// instruction" (the fully inlined kernel exceeds the instruction cache), but keeping these helpers out of line
// (-DMX_OUTLINE_BIG) measured 10 % SLOWER on B200 (call/ABI overhead, lost scheduling freedom), so they are inlined.
#ifdef MX_OUTLINE_BIG
#define MX_HD_BIG __host__ __device__ __noinline__
#else
#define MX_HD_BIG __host__ __device__ __forceinline__
#endif
#else
#define MX_HD inline
#define MX_HD_BIG inline
#endif

namespace mx {

constexpr double kPI = 3.14159265358979323846;   // jdmath.h PI
constexpr double kHbarC = 1.973269631e-4;         // marx.h:477 HBAR_C (keV um, 2006 NIST)

struct Vec3 { double x, y, z; };

MX_HD Vec3 v_make (double x, double y, double z) { Vec3 a; a.x = x; a.y = y; a.z = z; return a; }
// JDMv_dot_prod, jdmath/src/vector.c:68-76
MX_HD double v_dot (const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// JDMv_cross_prod, vector.c:40-58
MX_HD Vec3 v_cross (const Vec3 &a, const Vec3 &b)
{
   Vec3 c;
   c.z = a.x * b.y - a.y * b.x;
   c.x = a.y * b.z - a.z * b.y;
   c.y = a.z * b.x - a.x * b.z;
   return c;
}
// JDMv_length, vector.c:78-97 (scaled to avoid overflow; the scaling changes rounding, so keep it)
// Device builds divide once and multiply: a/len becomes a*(1/len), which can differ from the reference's
// three divisions by one ulp (1e-16 relative, the same size as the libm differences between glibc and
// libdevice; the parity tests hold at 1e-9 with exact integer outputs) and saves four of the six FP64
// divisions of a normalisation (~5 % of the whole trace).  -DMX_STRICT_DIV restores the three divisions;
// host builds (tools/hostcheck) always use them and stay bit-identical to the reference.
#if defined(__CUDA_ARCH__) && !defined(MX_STRICT_DIV)
#define MX_RECIP_NORMALIZE 1
#endif
MX_HD double v_length (const Vec3 &a)
{
   double x = fabs (a.x), y = fabs (a.y), z = fabs (a.z), tmp;
   if (z < x) { tmp = z; z = x; x = tmp; }
   if (z < y) { tmp = z; z = y; y = tmp; }
   if (z == 0.0) return 0.0;
#ifdef MX_RECIP_NORMALIZE
   double rz = 1.0 / z;
   x = x * rz; y = y * rz;
#else
   x = x / z; y = y / z;
#endif
   z = z * sqrt (1.0 + x * x + y * y);
   return z;
}
// JDMv_normalize, vector.c:99-109
MX_HD void v_normalize (Vec3 &a)
{
   double len = v_length (a);
#ifdef MX_RECIP_NORMALIZE
   if (len != 0.0) { double r = 1.0 / len; a.x = a.x * r; a.y = a.y * r; a.z = a.z * r; }
#else
   if (len != 0.0) { a.x = a.x / len; a.y = a.y / len; a.z = a.z / len; }
#endif
}
// JDMv_ax1_bx2, vector.c:121-131
MX_HD Vec3 v_ax1_bx2 (double a, const Vec3 &x1, double b, const Vec3 &x2)
{
   Vec3 c;
   c.x = a * x1.x + b * x2.x;
   c.y = a * x1.y + b * x2.y;
   c.z = a * x1.z + b * x2.z;
   return c;
}
// JDMv_ax1_bx2_cx3, vector.c:133-145
MX_HD Vec3 v_ax1_bx2_cx3 (double a, const Vec3 &x1, double b, const Vec3 &x2, double c, const Vec3 &x3)
{
   Vec3 d;
   d.x = a * x1.x + b * x2.x + c * x3.x;
   d.y = a * x1.y + b * x2.y + c * x3.y;
   d.z = a * x1.z + b * x2.z + c * x3.z;
   return d;
}
MX_HD Vec3 v_sum (const Vec3 &a, const Vec3 &b) { return v_make (a.x + b.x, a.y + b.y, a.z + b.z); }
MX_HD Vec3 v_diff (const Vec3 &a, const Vec3 &b) { return v_make (a.x - b.x, a.y - b.y, a.z - b.z); }
// JDMv_rotate_unit_vector1, vector.c:183-202 : Rodrigues rotation of p about unit n, then renormalise
MX_HD Vec3 v_rotate_unit1 (const Vec3 &p, const Vec3 &n, double cos_theta, double sin_theta)
{
   double pn = v_dot (p, n);
   Vec3 u = v_ax1_bx2_cx3 (cos_theta, p, pn * (1.0 - cos_theta), n, sin_theta, v_cross (n, p));
   v_normalize (u);
   return u;
}
// sin and cos of one angle: one argument reduction on the device (libdevice sincos returns the values
// of its sin and cos), the two libm calls of the reference on the host
MX_HD void sin_cos (double theta, double &s, double &c)
{
   mx_sincos (theta, s, c);          // mx_math.cuh: one argument reduction, coefficients in constant memory; libm on the host
}
// JDMv_rotate_unit_vector, vector.c:204-208
MX_HD_BIG Vec3 v_rotate_unit (const Vec3 &p, const Vec3 &n, double theta)
{
   double s, c;
   sin_cos (theta, s, c);
   return v_rotate_unit1 (p, n, c, s);
}
// JDM3m_vector_mul, jdmath/src/rotate.c:121-131 (row-major 3x3)
MX_HD Vec3 m3_mul (const double *m, const Vec3 &v)
{
   Vec3 b;
   b.x = m[0] * v.x + m[1] * v.y + m[2] * v.z;
   b.y = m[3] * v.x + m[4] * v.y + m[5] * v.z;
   b.z = m[6] * v.x + m[7] * v.y + m[8] * v.z;
   return b;
}

// An FP64 quotient that is exactly zero (0 / x) leaves the fast path of the device's division and runs its ~100-instruction special-case
// routine.  Three places of this path divide a numerator that is zero for most or all rays (a vector component that is literally 0, the
// diffraction term of order 0, the scatter-angle interpolation between two zero angles); each now tests for the zero first.  Written as
// `cond ? 0 : a / b` the compiler evaluates the quotient anyway and selects afterwards, so the division that remains is pinned inside its
// branch (ncu, round 2: 3 % of k01_source_hrma's instructions, 2-3 % of k1_hrma<4|5> and k2_grating<1>).
MX_HD double div_in_branch (double a, double b)
{
#if defined(__CUDA_ARCH__)
   double q;
   asm volatile ("div.rn.f64 %0, %1, %2;" : "=d"(q) : "d"(a), "d"(b));
   return q;
#else
   return a / b;
#endif
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 draw stream (include/marxb200.h, "Random draws").  Replaces JDMrandom
// (jdmath/src/random.c:100-154) with a per-(ray, stage) counter so photons are independent.
// ------------------------------------------------------------------------------------------------
MX_HD uint32_t mulhi32 (uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
   return __umulhi (a, b);
#else
   return (uint32_t) (((uint64_t) a * b) >> 32);
#endif
}

#ifndef MX_PHILOX_UNROLL
#define MX_PHILOX_UNROLL 10
#endif
struct Rng
{
   uint32_t k0, k1, c0, c1, stage, draw;
   uint32_t b0, b1, b2, b3;
   double spare;
   int have_spare;

   MX_HD void init (uint64_t seed, uint64_t ray, uint32_t stg)
   {
      k0 = (uint32_t) seed; k1 = (uint32_t) (seed >> 32);
      c0 = (uint32_t) ray; c1 = (uint32_t) (ray >> 32);
      stage = stg; draw = 0; have_spare = 0; spare = 0.0;
      b0 = b1 = b2 = b3 = 0;
   }
   // continue a stream that another kernel left after `d` draws (sub-stage kernels of one MARX stage share a stream)
   // The block is consumed from b0: every draw returns b0 and shifts the rest down (three moves, none when the compiler knows the
   // draw index; picking word draw & 3 took three compares and three selects per draw: 6 % of k01's instructions).
   MX_HD void resume (uint32_t d, int has_spare, double spare_value)
   {
      draw = d; have_spare = has_spare; spare = spare_value;
      if ((d & 3u) != 0u)
        {
           refill (d >> 2);
#ifndef MX_PHILOX_SELECT
           for (uint32_t k = 0; k < (d & 3u); k++) { b0 = b1; b1 = b2; b2 = b3; }
#endif
        }
   }
   MX_HD void refill (uint32_t block)
   {
      uint32_t x0 = c0, x1 = c1, x2 = block, x3 = stage, ka = k0, kb = k1;
      // Developer knob (tools/build_variant.sh -DMX_PHILOX_UNROLL=k): rounds per loop trip.  Rolling the loop shrinks every inlined
      // copy, but measured on B200 only the split ACIS kernels gained (0.65 -> 0.61 ms at k = 2) while the mirror kernels lost
      // more (k = 1: k01 +5 %, k1_hrma<1> +4 %): fully unrolled stays the default.
      constexpr int kUnroll = MX_PHILOX_UNROLL;
#pragma unroll kUnroll
      for (int i = 0; i < 10; i++)
        {
           uint32_t hi0 = mulhi32 (0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
           uint32_t hi1 = mulhi32 (0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
           uint32_t n0 = hi1 ^ x1 ^ ka, n1 = lo1, n2 = hi0 ^ x3 ^ kb, n3 = lo0;
           x0 = n0; x1 = n1; x2 = n2; x3 = n3;
           ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
      b0 = x0; b1 = x1; b2 = x2; b3 = x3;
   }
   MX_HD uint32_t next_u32 ()
   {
#ifdef MX_PHILOX_SELECT            // developer A/B knob: the round-1 form (word draw & 3 of an unshifted block)
      const uint32_t lane = draw & 3u;
      if (lane == 0u) refill (draw >> 2);
      draw++;
      return lane == 0u ? b0 : (lane == 1u ? b1 : (lane == 2u ? b2 : b3));
#else
      if ((draw & 3u) == 0u) refill (draw >> 2);
      draw++;
      const uint32_t r = b0;
      b0 = b1; b1 = b2; b2 = b3;
      return r;
#endif
   }
   // JDMrandom, random.c:151-154: uniform on [0,1] inclusive
   MX_HD double uniform () { return (double) next_u32 () * (1.0 / 4294967295.0); }
   // JDMgaussian_random, jdmath/src/gaussrnd.c:30-55 (polar Box-Muller, spare cached per stream)
   MX_HD double gaussian ()
   {
      if (have_spare) { have_spare = 0; return spare; }
      double g1, g2, g;
      do
        {
           g1 = 2.0 * uniform () - 1.0;
           g2 = 2.0 * uniform () - 1.0;
           g = g1 * g1 + g2 * g2;
        }
      while ((g >= 1.0) || (g == 0.0));
      double s = sqrt (-2.0 * mx_log (g) / g);
      spare = g2 * s; have_spare = 1;
      return g1 * s;
   }
   // JDMexpn_random, gaussrnd.c:57-67
   MX_HD double expn ()
   {
      double r;
      do r = uniform (); while (r == 0.0);
      return -mx_log (r);
   }
};

// ------------------------------------------------------------------------------------------------
// float-table interpolation (jdmath/src/finterpo.c)
// ------------------------------------------------------------------------------------------------
// JDMbinary_search_f, finterpo.c:37-57: first index with xp[i] >= x (n if none), with the equality
// short-circuit that matters when the grid holds duplicate abscissae.
template <class T> MX_HD uint32_t bsearch_ref (T x, const T *xp, uint32_t n, uint32_t stride = 1)
{
   uint32_t n0 = 0, n1 = n, n2;
   while (n1 > n0 + 1)
     {
        n2 = (n0 + n1) / 2;
        T v = xp[stride * n2];
        if (v >= x)
          {
             if (v == x) return n2;
             n1 = n2;
          }
        else n0 = n2;
     }
   if (x >= xp[stride * n0]) return n1;
   return n0;
}
// Device form.  The bisection above is a chain of log2(n) DEPENDENT loads (11 for the 1426..2000-point energy grids): ncu
// attributes 10-16 % of the stall samples of k1_hrma<1> and k3_acis to waiting on them.  Unless x equals a grid value, the
// reference's result does not depend on its probe sequence: it is the lower bound lb = #{i: xp[i] < x} (for lb == 0 it returns
// 0 when x < xp[0]).  So: find lb with a branch-free binary lower bound (trip count fixed by n, one select per probe: no
// divergence, no early-exit branch) and hand the rare exact-equality (or NaN) case to the reference loop, whose answer then
// depends on which duplicate abscissa it happens to probe.  Measured on B200 (C2, 2^24 rays): k3_acis 1.21 -> 1.04 ms,
// k1_hrma<1> 1.38 -> 1.32 ms.  MX_SEARCH_ARITY = k > 1 selects a k-ary search instead (k-1 independent probes per round,
// fewer dependent latencies): 2..6 within 1 % of the binary form, 8 and 16 slower (registers) -- the latency chain was not
// the cost, the divergent loop was.  0 = the reference loop.
#ifndef MX_SEARCH_ARITY
#define MX_SEARCH_ARITY 1
#endif
template <class T> MX_HD uint32_t bsearch_fast (T x, const T *xp, uint32_t n, uint32_t stride = 1)
{
#if defined(__CUDA_ARCH__) && (MX_SEARCH_ARITY == 1)
   // branch-free binary lower bound: the trip count depends on n only (no divergence), one select per probe
   uint32_t base = 0, len = n;
   while (len > 1)
     {
        const uint32_t half = len >> 1;
        base += (xp[stride * (base + half - 1)] < x) ? half : 0u;
        len -= half;
     }
   const uint32_t lb = base + (((n > 0) && (xp[stride * base] < x)) ? 1u : 0u);
   if (!(x == x) || ((lb < n) && (xp[stride * lb] == x))) return bsearch_ref (x, xp, n, stride);
   return lb;
#elif defined(__CUDA_ARCH__) && (MX_SEARCH_ARITY > 1)
   constexpr uint32_t A = MX_SEARCH_ARITY;
   uint32_t lo = 0, hi = n;                       // xp[i] < x for i < lo, xp[i] >= x for i >= hi
   while (hi - lo > A)
     {
        const uint32_t step = (hi - lo) / A;
        uint32_t c = 0;
#pragma unroll
        for (uint32_t k = 1; k < A; k++) c += (xp[stride * (lo + step * k)] < x) ? 1u : 0u;
        const uint32_t base = lo;
        if (c > 0) lo = base + step * c + 1u;
        if (c < A - 1u) hi = base + step * (c + 1u);
     }
   uint32_t c = 0;
#pragma unroll
   for (uint32_t k = 0; k < A; k++) c += ((lo + k < hi) && (xp[stride * (lo + k)] < x)) ? 1u : 0u;
   const uint32_t lb = lo + c;
   if (!(x == x) || ((lb < n) && (xp[stride * lb] == x))) return bsearch_ref (x, xp, n, stride);
   return lb;
#else
   return bsearch_ref (x, xp, n, stride);
#endif
}
MX_HD uint32_t bsearch_f (float x, const float *xp, uint32_t n) { return bsearch_fast<float> (x, xp, n); }
// JDMbinary_search_d, jdmath/src/dinterpo.c (same algorithm on doubles)
MX_HD uint32_t bsearch_d (double x, const double *xp, uint32_t n) { return bsearch_fast<double> (x, xp, n); }
// JDMinterpolate_f, finterpo.c:59-85.  x is narrowed to float by the caller's call (the C prototype
// takes float); (yp[n1]-yp[n0]) is a float subtraction, the rest is double, the result is a float.
// The reference reads xp[n] one past the end when x exceeds the grid (finterpo.c:68); that compare
// is guarded here.
MX_HD_BIG float interp_f (float x, const float *xp, const float *yp, uint32_t n)
{
   if (n == 1) return yp[0];
   uint32_t n1 = bsearch_f (x, xp, n);
   uint32_t n0 = n1 - 1;
   if ((n1 < n) && (x == xp[n1])) return yp[n1];
   if (n1 == n) { n1--; n0--; }
   if (n1 == 0) n0 = 1;
   double x0 = xp[n0], x1 = xp[n1];
   if (x1 == x0) return yp[n1];
   float dy = yp[n1] - yp[n0];
   return (float) (yp[n0] + dy / (x1 - x0) * (x - x0));
}
// two JDMinterpolate_f calls on the SAME abscissa grid (beta and delta of hrma.c:1250-1251) share one search
MX_HD void interp_f2 (float x, const float *xp, const float *yp1, const float *yp2, uint32_t n, float &y1, float &y2)
{
   if (n == 1) { y1 = yp1[0]; y2 = yp2[0]; return; }
   uint32_t n1 = bsearch_f (x, xp, n);
   uint32_t n0 = n1 - 1;
   if ((n1 < n) && (x == xp[n1])) { y1 = yp1[n1]; y2 = yp2[n1]; return; }
   if (n1 == n) { n1--; n0--; }
   if (n1 == 0) n0 = 1;
   double x0 = xp[n0], x1 = xp[n1];
   if (x1 == x0) { y1 = yp1[n1]; y2 = yp2[n1]; return; }
   float dy1 = yp1[n1] - yp1[n0], dy2 = yp2[n1] - yp2[n0];
   y1 = (float) (yp1[n0] + dy1 / (x1 - x0) * (x - x0));
   y2 = (float) (yp2[n0] + dy2 / (x1 - x0) * (x - x0));
}
// JDMinterpolate_d, jdmath/src/dinterpo.c (all double)
MX_HD double interp_d (double x, const double *xp, const double *yp, uint32_t n)
{
   if (n == 1) return yp[0];
   uint32_t n1 = bsearch_d (x, xp, n);
   uint32_t n0 = n1 - 1;
   if ((n1 < n) && (x == xp[n1])) return yp[n1];
   if (n1 == n) { n1--; n0--; }
   if (n1 == 0) n0 = 1;
   double x0 = xp[n0], x1 = xp[n1];
   if (x1 == x0) return yp[n1];
   return yp[n0] + (yp[n1] - yp[n0]) / (x1 - x0) * (x - x0);
}

}  // namespace mx
