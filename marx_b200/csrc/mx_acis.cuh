// mx_acis.cuh -- K3: ACIS chip-plane intersection, QE x OBF x contamination, FEF pulse height, streak.
// Reference: marx/libsrc/acis-s.c:177-248 (_marx_acis_s_detect), :138-175 (_marx_acis_apply_qe_and_pha);
// detector.c:56-168 (plane intersection), trans.c:40-90; aciscontam.c:93-187; acis_fef.c:380-579,
// :910-1079 (find_fef, normalize_gaussians, mixture sampling, marx_apply_acis_rmf); acis-i.c:60-89 (streak).
// Draw order on sub-stream MARXB200_STAGE_DETECTOR: QE U; FEF { component U; 1..100 G | tail 2 U / iter }
// [+ 1 U per negative-amplitude rejection trial]; PI U; streak U (transfer window only).
#pragma once
#include "mx_common.cuh"
#include "mx_tables.h"

namespace mx {

// rotate_vector / rotate_vector_inv, trans.c:40-64
MX_HD Vec3 m3_mul_t (const double *m, const Vec3 &v)
{
   Vec3 b;
   b.x = m[0] * v.x + m[3] * v.y + m[6] * v.z;
   b.y = m[1] * v.x + m[4] * v.y + m[7] * v.z;
   b.z = m[2] * v.x + m[5] * v.y + m[8] * v.z;
   return b;
}

// Detector dither of one ray (Marx_Dither_Type dy, dz, dtheta; zero for the NONE and INTERNAL models).
// _marx_dither_detector, detector.c:275-284: the SIM offset moves by (dy, dz) and the detector matrix is rotated about
// x by dtheta (rotate_matrix :240-266) for this ray.  The reference applies and removes it on its one global matrix, ray
// after ray, so its matrix drifts by rounding (~1e-16 per ray); here every ray starts from the pristine matrix.
struct DetDither { double dy, dz, dtheta; };
MX_HD void det_dither_frame (const double *offset, const double *matrix, const DetDither &dd, double *off, double *M)
{
   off[0] = offset[0]; off[1] = offset[1] + dd.dy; off[2] = offset[2] + dd.dz;
   for (int k = 0; k < 9; k++) M[k] = matrix[k];
   double s = 0, c = 1;
   if (dd.dtheta != 0) { s = sin (dd.dtheta); c = cos (dd.dtheta); }
   const double m00 = M[4], m01 = M[5], m10 = M[7], m11 = M[8];
   M[4] = m00 * c - m01 * s; M[5] = m00 * s + m01 * c;
   M[7] = m10 * c - m11 * s; M[8] = m10 * s + m11 * c;
}

// intersect_with_detector_plane, detector.c:56-109.  must_hit: return 0 as soon as the point is off the chip;
// otherwise (second pass of DetExtendFlag=yes) report the point anyway and return whether it was on the chip.
// Plain-array form so that ACIS chips and HRC MCPs share it.
MX_HD int plane_intersect (const double *g_normal, const double *g_x_ll, const double *g_xhat, const double *g_yhat,
                           double xlen, double ylen, const Vec3 &x0, const Vec3 &p, Vec3 &x, double &dx, double &dy, bool must_hit)
{
   Vec3 normal = v_make (g_normal[0], g_normal[1], g_normal[2]);
   double pdotn = v_dot (p, normal);
   if (pdotn == 0) return -1;
   Vec3 x_ll = v_make (g_x_ll[0], g_x_ll[1], g_x_ll[2]);
   Vec3 r = v_diff (x0, x_ll);
   r = v_ax1_bx2 (1.0, r, -1.0 * v_dot (r, normal) / pdotn, p);
   int hit = 1;
   double rx = v_dot (r, v_make (g_xhat[0], g_xhat[1], g_xhat[2]));
   if ((rx < 0.0) || (rx >= xlen)) { if (must_hit) return 0; hit = 0; }
   double ry = v_dot (r, v_make (g_yhat[0], g_yhat[1], g_yhat[2]));
   if ((ry < 0.0) || (ry >= ylen)) { if (must_hit) return 0; hit = 0; }
   x = v_sum (r, x_ll);
   dx = rx; dy = ry;
   return hit;
}
MX_HD int chip_intersect (const AcisChipDev &g, const Vec3 &x0, const Vec3 &p, Vec3 &x, double &dx, double &dy, bool must_hit = true)
{
   return plane_intersect (g.normal, g.x_ll, g.xhat, g.yhat, g.xlen, g.ylen, x0, p, x, dx, dy, must_hit);
}
// _marx_intersect_with_detector, detector.c:111-168, over n facets accessed through `facet(k)`.  First facet hit wins;
// with extend_flag the ray is otherwise assigned to the facet whose CENTRE is closest to its plane intersection.
template <class Facets>
MX_HD int detector_intersect (const Facets &F, int n, const Vec3 &x0, const Vec3 &p, Vec3 &xh, double &dx, double &dy, int extend_flag)
{
   for (int k = 0; k < n; k++)
     if (1 == plane_intersect (F[k].normal, F[k].x_ll, F[k].xhat, F[k].yhat, F[k].xlen, F[k].ylen, x0, p, xh, dx, dy, true)) return k;
   if (extend_flag == 0) return -1;
   int best = -1;
   double best_r2 = -1, bdx = 0, bdy = 0;
   Vec3 bx = x0;
   for (int k = 0; k < n; k++)
     {
        Vec3 x; double ddx, ddy;
        if (-1 == plane_intersect (F[k].normal, F[k].x_ll, F[k].xhat, F[k].yhat, F[k].xlen, F[k].ylen, x0, p, x, ddx, ddy, false)) continue;
        double deltax = ddx - 0.5 * F[k].xlen, deltay = ddy - 0.5 * F[k].ylen;
        double r2 = deltax * deltax + deltay * deltay;
        if ((r2 < best_r2) || (best < 0)) { best = k; bx = x; bdx = ddx; bdy = ddy; best_r2 = r2; }
     }
   if (best < 0) return -1;
   xh = bx; dx = bdx; dy = bdy;
   return best;
}

// Code-size control for k3_acis: the kernel is ~90 KB of SASS executed by warps that sit in different phases (geometry,
// QE, contamination, FEF normalisation, sampling), and ncu showed "no instruction" (I-cache miss) as its largest stall.
// With MX_K3_OUTLINE the table interpolation and the libm bodies used by this stage exist once, as real functions.
#if defined(__CUDA_ARCH__) && defined(MX_K3_OUTLINE)
__device__ __noinline__ float acis_interp_f (float x, const float *xp, const float *yp, uint32_t n) { return interp_f (x, xp, yp, n); }
#if MX_K3_OUTLINE > 1
__device__ __noinline__ double acis_exp (double x) { return exp (x); }
__device__ __noinline__ double acis_erf (double x) { return erf (x); }
__device__ __noinline__ double acis_pow (double x, double y) { return pow (x, y); }
#else
#define acis_exp exp
#define acis_erf erf
#define acis_pow pow
#endif
#else
#define acis_interp_f interp_f
#define acis_exp exp
#define acis_erf erf
#define acis_pow pow
#endif

// compute_contamination, aciscontam.c:93-136, with the analytic f(x,y) of :141-187
MX_HD double acis_contamination (const AcisChipDev &c, double en, double cx, double cy)
{
   if (c.contam_num_layers == 0) return 1.0;
   float ef = (float) en;
   double v = 0.0;
   if (c.contam_fxy_mode != 0)
     {
        double fxy;
        if (c.contam_fxy_mode == 1)
          {
             double ddx = cx - c.contam_x0, ddy = cy - c.contam_y0;
             double r = (8.0 / 1024.0) * sqrt (ddx * ddx + ddy * ddy);
             r /= 8.07;
             fxy = 1.29 * r * r;
          }
        else
          {
             const double y_0 = 512.0;
             // aciscontam.c:141-187: two branches with different constants; one pow body serves both
             const bool low = (cy <= 512.0);
             const double den = low ? (64.0 - y_0) : (964.0 - y_0), ex = low ? 5.5 : 4.5;
             fxy = acis_pow (fabs ((cy - y_0) / den), ex);
          }
        for (uint32_t i = 0; i < c.contam_num_layers; i++)
          {
             double mu = acis_interp_f (ef, c.contam_energies[i], c.contam_mus[i], c.contam_num_mu[i]);
             v += mu * (c.contam_tau0[i] + c.contam_tau1[i] * fxy);
          }
     }
   else
     {
        if ((cx < 0) || (cx >= 1024) || (cy < 0) || (cy >= 1024)) return 0.0;
        cx /= c.contam_blocking;
        cy /= c.contam_blocking;
        uint32_t ofs = (1024 / c.contam_blocking) * (uint32_t) cy + (uint32_t) cx;
        for (uint32_t i = 0; i < c.contam_num_layers; i++)
          {
             double mu = acis_interp_f (ef, c.contam_energies[i], c.contam_mus[i], c.contam_num_mu[i]);
             double fxy = c.contam_fxy[i][ofs];
             v += mu * (c.contam_tau0[i] + c.contam_tau1[i] * fxy);
          }
     }
   return acis_exp (-v);
}


// fmod(t, T) for t >= 0, T > 0 -- EXACT, like the libm routine, without its bit-serial long division:
// with q = floor(t/T) (possibly off by one) the remainder t - q*T is a multiple of ulp(T) smaller than 2T in
// magnitude, hence representable, and a single fused multiply-add delivers it without rounding.
MX_HD double fmod_pos (double t, double T)
{
   if (!(t >= T)) return (t >= 0.0) ? t : fmod (t, T);
   if (t > 4.0e15 * T) return fmod (t, T);        // quotient beyond 2^53: leave it to libm
   double q = floor (t / T);
   double r = fma (-q, T, t);
   if (r < 0.0) r += T;
   else if (r >= T) r -= T;
   return r;
}

// One FEF row pair bracketing the photon energy: the gaussians of the event are linear interpolations between the two
// rows (acis_fef.c:1011-1032).  The reference materialises them in an array; here gaussian k is re-interpolated (same
// operations, same float roundings) wherever it is needed, so that no per-thread array -- local memory on the device,
// 320 B per ray that ncu showed spilling to DRAM -- is required.  Only the cumulative areas are kept: kMaxGauss floats
// per thread in SHARED memory (`cum`, element k at cum[k * stride]) plus a bit mask for use_tail_dist.
struct FefRows { const float *g0, *g1; double t; uint32_t ng; };
struct GaussParm { float amp, center, sigma; };   // acis_fef.c:87-96 (cum_area, use_tail_dist live in cum[] / tail_mask)

MX_HD GaussParm fef_gauss (const FefRows &R, uint32_t k)
{
   GaussParm G;
   const double t = R.t;
   float a0 = R.g0[3 * k], c0 = R.g0[3 * k + 1], s0 = R.g0[3 * k + 2];
   float a1 = R.g1[3 * k], c1 = R.g1[3 * k + 1], s1 = R.g1[3 * k + 2];
   G.center = (float) (c0 + t * (c1 - c0));
   double v = s0 + t * (s1 - s0);
   if (v <= 0.0) { G.sigma = 0.0f; G.amp = 0.0f; }
   else
     {
        G.sigma = (float) v;
        v = a0 + t * (a1 - a0);
        if ((v < 0.0) && ((a1 > 0.0f) || (a0 > 0.0f))) v = 0.0;
        G.amp = (float) v;
     }
   return G;
}

// normalize_gaussians, acis_fef.c:509-579.  gaussian_integral(0,+inf) and (-inf,0) share one erf:
// erf((+-1e37 - x0)/sigma) is exactly +-1 for every representable table value.
MX_HD int fef_normalize (const FefRows &R, float *cum, uint32_t stride, uint32_t &tail_mask)
{
   const double SQRT_2 = 1.4142135623730951, SQRT_2PI = 2.5066282746310002;
   const uint32_t num = R.ng;
   double total_pos_area = 0.0, total_neg_area = 0.0;
   int flags = 0;
   tail_mask = 0;
   for (uint32_t k = 0; k < num; k++)
     {
        const GaussParm g = fef_gauss (R, k);
        double area1 = 0.0, area2 = 0.0;
        double sigma = g.sigma * SQRT_2;
        // amp == 0 makes both areas exactly zero whatever erf returns: skip the erf (identical results)
        if ((sigma != 0.0) && (g.amp != 0.0f))
          {
             double x0 = g.center;
             // erf is exactly -1 / +1 in double beyond |z| = 5.93 (erfc (6) = 2e-17 < 2^-54): the peaks of a response function
             // sit tens of widths above zero, so almost every component skips the polynomial
             const double zarg = (0 - x0) / sigma;
             double e0;
             if (zarg <= -6.0) e0 = -1.0;
             else if (zarg >= 6.0) e0 = 1.0;
             else e0 = acis_erf (zarg);
             area1 = 0.5 * g.amp * (1.0 - e0) * (SQRT_2PI * g.sigma);
             area2 = 0.5 * g.amp * (e0 - (-1.0)) * (SQRT_2PI * g.sigma);
          }
        if (area2 > area1)
          {
             // Almost every component has area2 == 0 exactly (erf saturates) and takes the other branch; written as a plain
             // `area1 / area2` the compiler hoists the quotient above the test, and a division by zero runs the ~100-instruction
             // special-case path of the FP64 division for every (event, component): 7 % of this kernel's instructions (ncu).
             double ratio;
#if defined(__CUDA_ARCH__)
             asm volatile ("div.rn.f64 %0, %1, %2;" : "=d"(ratio) : "d"(area1), "d"(area2));
#else
             ratio = area1 / area2;
#endif
             if (ratio < 0.1) tail_mask |= (1u << k);
          }
        if (area1 >= 0) total_pos_area += area1;
        else total_neg_area -= area1;
        cum[k * stride] = (float) total_pos_area;
     }
   if (total_pos_area <= total_neg_area) flags |= 4;      // HAS_TOTAL_NEG_AREA
   if (total_neg_area != 0.0) flags |= 1;                 // HAS_NEG_AMP_GAUSSIANS
   if (total_pos_area > 0)
     for (uint32_t k = 0; k < num; k++)
       cum[k * stride] = (float) (cum[k * stride] / total_pos_area);
   return flags;
}

// compute_pha_with_pos_amps, acis_fef.c:408-468.  Same draws in the same order as the reference; the
// component search is separated from the sampling so that the lanes of a warp reconverge before the
// expensive part (in the reference's loop shape every lane would sample inside a different iteration).
// The reference loops until a value is found; the cap only bounds pathological tables.
MX_HD int fef_pha_pos (const FefRows &R, const float *cum, uint32_t stride, uint32_t tail_mask, double &phap, Rng &rng)
{
   const uint32_t num = R.ng;
   for (int guard = 0; guard < 4096; guard++)
     {
        double r = rng.uniform ();
        uint32_t k = 0;
        while ((k < num) && (cum[k * stride] <= r)) k++;
        if (k == num) continue;                  // r above every cumulative area: draw again (acis_fef.c:411-424)
        const GaussParm g = fef_gauss (R, k);
        const float center = g.center, sigma = g.sigma;
        double pha;
        if (0 == ((tail_mask >> k) & 1u))
          {
             unsigned int count = 0;
             do
               {
                  pha = center + sigma * rng.gaussian ();
                  count++;
               }
             while ((pha < 0) && (count < 100));
          }
        else
          {
             // truncated-tail sampler adapted from GSL (acis_fef.c:440-456)
             double u, v, x;
             double s = (0 - center) / sigma;      // float arithmetic, as in the reference
             do
               {
                  u = rng.uniform ();
                  do v = rng.uniform (); while (v == 0.0);
                  x = sqrt (s * s - 2 * log (v));
               }
             while (x * u > s);
             pha = center + x * sigma;
          }
        if (pha < 0) continue;                   // "Failed to find a pha value": draw a new r
        phap = pha;
        return 0;
     }
   return -1;
}

// compute_pha_with_neg_amps, acis_fef.c:471-503
MX_HD int fef_pha_neg (const FefRows &R, const float *cum, uint32_t stride, uint32_t tail_mask, double &phap, Rng &rng)
{
   int count = 0;
   while (count < 100)
     {
        double pha;
        if (-1 == fef_pha_pos (R, cum, stride, tail_mask, pha, rng)) return -1;
        double pos_sum = 0.0, sum = 0.0;
        for (uint32_t k = 0; k < R.ng; k++)
          {
             const GaussParm g = fef_gauss (R, k);
             double sigma = g.sigma, dsum = 0.0;
             if (sigma != 0.0)
               {
                  double xx = (pha - g.center) / sigma;
                  dsum = g.amp * acis_exp (-0.5 * xx * xx);
               }
             sum += dsum;
             if (dsum > 0) pos_sum += dsum;
          }
        if (rng.uniform () * pos_sum < sum) { phap = pha; return 0; }
        count++;
     }
   return -1;
}

// marx_apply_acis_rmf, acis_fef.c:967-1079 (+ find_fef :910-965).  x, y are the float chip pixels.
// cum: kMaxGauss floats of scratch for this ray, element k at cum[k * stride] (shared memory on the device).
MX_HD int acis_apply_fef (const AcisDev &A, const AcisChipDev &chip, float x, float y, double energy,
                          float &pi, int16_t &pha_out, Rng &rng, float *cum, uint32_t stride)
{
   // find_fef, acis_fef.c:910-965: off-chip pixels are an error unless DetExtendFlag=yes, which clamps them
   if ((x < 0) || (x >= 1024) || (y < 0) || (y >= 1024))
     {
        if (A.det_extend == 0) return -1;
        if (x < 0) x = 0; else if (x >= 1024) x = 1023;
        if (y < 0) y = 0; else if (y >= 1024) y = 1023;
     }
   uint32_t i = (uint32_t) (x / 32), j = (uint32_t) (y / 32);
   if ((i >= 32) || (j >= 32)) return -1;
   int fi = chip.fef_map[i * 32 + j];
   if (fi < 0) return -1;
   const FefDev &f = A.fefs[fi];
   uint32_t ng = f.num_gaussians, ne = f.num_energies;
   if (ng > (uint32_t) kMaxGauss) return -1;

   i = bsearch_f ((float) energy, f.energies, ne);
   if (i == 0) i++;
   if (i == ne) i--;
   FefRows R;
   R.t = (energy - f.energies[i - 1]) / (f.energies[i] - f.energies[i - 1]);
   R.g0 = f.gauss + (size_t) (i - 1) * ng * 3;
   R.g1 = R.g0 + (size_t) ng * 3;
   R.ng = ng;
   uint32_t tail_mask;
   int flags = fef_normalize (R, cum, stride, tail_mask);
   if (flags & 4) return -1;
   double pha;
   int status = (flags == 0) ? fef_pha_pos (R, cum, stride, tail_mask, pha, rng) : fef_pha_neg (R, cum, stride, tail_mask, pha, rng);
   if (status == -1) return -1;
   int16_t ipha = (int16_t) pha;                   // truncation, acis_fef.c:1065
   pha_out = ipha;
   pha = ipha - rng.uniform ();
   pi = acis_interp_f ((float) pha, f.channels, f.energies, ne);
   if (pi < 0) return -1;
   return 0;
}

// _marx_acis_s_detect for one ray, in two halves so that the detector stage can run as two kernels (k3_acis<DET, 1|2>: the
// 88 KB single kernel spent 43 % of its stall samples waiting for instruction fetch; each half has half the footprint, and the
// second starts from full warps).  t_abs = pt->start_time + arrival_time.  Both return flags (0 alive, possibly with
// PHOTON_ACIS_STREAKED set, which is not a "dead" bit).
// DET: the per-ray detector dither is live (ASPSOL model / uploaded records); false compiles the table-frame path only
struct AcisFrame { const double *off, *mat; double off_l[3], mat_l[9]; };
template <bool DET>
MX_HD void acis_frame (const AcisDev &A, const DetDither &dd, AcisFrame &F)
{
   // _marx_dither_detector, detector.c:275-284 (the detector dither is zero for the INTERNAL model, dither.c:177-179: the
   // frame is then the table's)
   F.off = A.det_offset; F.mat = A.det_matrix;
   if (DET && (A.dither_mode != 0) && ((dd.dy != 0) || (dd.dz != 0) || (dd.dtheta != 0)))
     {
        det_dither_frame (A.det_offset, A.det_matrix, dd, F.off_l, F.mat_l);
        F.off = F.off_l; F.mat = F.mat_l;
     }
}

// first half: _marx_transform_ray (trans.c:66-77), chip-plane intersection (detector.c:111-168), QE x filter x contamination
// (_marx_acis_apply_qe_and_pha :138-163).  Leaves x (the point on the chip) and p in the DETECTOR frame; hit = index of the chip.
template <bool DET = false>
MX_HD uint32_t acis_detect_a (const AcisDev &A, double energy, Vec3 &x, Vec3 &p, int &ccd, int &hit_out, float &chipx, float &chipy,
                              Rng &rng, const DetDither &dd = DetDither {0.0, 0.0, 0.0})
{
   const uint32_t UNDETECTED = 0x01, MISSED = 0x08;
   AcisFrame F;
   acis_frame<DET> (A, dd, F);
   x.x -= F.off[0]; x.y -= F.off[1]; x.z -= F.off[2];
   x = m3_mul (F.mat, x);
   p = m3_mul (F.mat, p);

   double dx = 0, dy = 0;
   Vec3 xh = x;
   const int hit = detector_intersect (A.chip, A.num_chips, x, p, xh, dx, dy, A.det_extend);
   hit_out = hit;
   if (hit < 0)
     {
        ccd = -1;
        return MISSED;
     }
   const AcisChipDev &d = A.chip[hit];
   x = xh;
   ccd = d.id;
   chipx = (float) (dx / d.x_pixel_size);
   chipy = (float) (dy / d.y_pixel_size);

   if (A.det_ideal == 0)
     {
        double r = rng.uniform ();
        float ef = (float) energy;
        double qe = (d.qe_num != 0) ? (double) acis_interp_f (ef, d.qe_energies, d.qe, d.qe_num) : 1.0;
        double qe_filter = (d.filter_num != 0) ? (double) acis_interp_f (ef, d.filter_energies, d.filter_qe, d.filter_num) : 1.0;
        double qe_contam = acis_contamination (d, energy, chipx, chipy);
        if (r >= qe * qe_filter * qe_contam) return UNDETECTED;
     }
   return 0;
}

// second half: FEF pulse height (marx_apply_acis_rmf), streak (_marx_acis_apply_streak, acis-i.c:60-89), and
// _marx_transform_ray_reverse (trans.c:79-90).  x, p arrive in the detector frame.
template <bool DET = false>
MX_HD uint32_t acis_detect_b (const AcisDev &A, double energy, double t_abs, Vec3 &x, Vec3 &p, int hit, float chipx, float &chipy,
                              int16_t &pha, float &pi, Rng &rng, float *fef_cum, uint32_t fef_stride,
                              const DetDither &dd = DetDither {0.0, 0.0, 0.0})
{
   const uint32_t UNDETECTED = 0x01, STREAKED = 0x200;
   uint32_t flags = 0;
   const AcisChipDev &d = A.chip[hit];
   if (-1 == acis_apply_fef (A, d, chipx, chipy, energy, pi, pha, rng, fef_cum, fef_stride))
     {
        pha = -1; pi = 0;
        return UNDETECTED;
     }
   if (A.frame_transfer_time > 0.0)
     {
        double t = fmod_pos (t_abs, A.frame_time);
        if (t > A.exposure_time)
          {
             chipy = (float) (1.0 + 1022.0 * rng.uniform ());
             double xpixel = (chipx - d.xpixel_offset) * d.x_pixel_size;
             double ypixel = (chipy - d.ypixel_offset) * d.y_pixel_size;
             Vec3 ddx = v_ax1_bx2 (xpixel, v_make (d.xhat[0], d.xhat[1], d.xhat[2]), ypixel, v_make (d.yhat[0], d.yhat[1], d.yhat[2]));
             x = v_sum (v_make (d.x_ll[0], d.x_ll[1], d.x_ll[2]), ddx);
             p = v_diff (x, v_make (A.focal_length, 0, 0));
             v_normalize (p);
             flags |= STREAKED;
          }
     }
   AcisFrame F;
   acis_frame<DET> (A, dd, F);
   p = m3_mul_t (F.mat, p);
   x = m3_mul_t (F.mat, x);
   x.x += F.off[0]; x.y += F.off[1]; x.z += F.off[2];
   return flags;
}

template <bool DET = false>
MX_HD uint32_t acis_detect (const AcisDev &A, double energy, double t_abs, Vec3 &x, Vec3 &p,
                            int &ccd, float &chipx, float &chipy, int16_t &pha, float &pi, Rng &rng,
                            float *fef_cum, uint32_t fef_stride, const DetDither &dd = DetDither {0.0, 0.0, 0.0})
{
   int hit = -1;
   const uint32_t flags = acis_detect_a<DET> (A, energy, x, p, ccd, hit, chipx, chipy, rng, dd);
   if (flags != 0) return flags;
   return acis_detect_b<DET> (A, energy, t_abs, x, p, hit, chipx, chipy, pha, pi, rng, fef_cum, fef_stride, dd);
}

}  // namespace mx
