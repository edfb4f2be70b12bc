// mx_hrma.cuh -- K1: HRMA Wolter-I shell pair, per ray.
// Reference: marx/libsrc/hrma.c:1161-1341 (_marx_hrma_mirror_reflect), :984-1050 (project_photon_to_hrma),
// :411-478 (compute_conic_intersection), :488-545 (reflect_from_conic), :1054-1093 (blur_normal),
// :928-968 (intersects_struts); reflect.c:39-77 (marx_reflectivity, with jdmath/src/complex.c);
// wfold.c:304-369 (scatter table lookup); jdmath/src/qroot.c:36-68.
// Draw order on sub-stream MARXB200_STAGE_MIRROR: vignetting U; shell U; radius U;
// azimuth U (retry on closed shutter); then per conic: blur U + G, reflect U, scatter U, sign U.
//
// Deliberate difference from the reference, invisible in its output: a ray that is blocked or absorbed
// stops here with the flag of the FIRST cause (the reference keeps tracing strut-blocked rays, hrma.c:1214,
// only to discard them; marx_write_photons never emits a flagged ray).
#pragma once
#include "mx_common.cuh"
#include "mx_tables.h"

namespace mx {

struct Cplx { double r, i; };
// JDMc_mul / JDMc_div / JDMc_abs / JDMc_sqrt, jdmath/src/complex.c:38-227
MX_HD Cplx c_mul (Cplx a, Cplx b) { Cplx z; z.r = a.r * b.r - a.i * b.i; z.i = a.r * b.i + a.i * b.r; return z; }
// The reference branches on |r2| > |i2| and the lanes of a warp split about evenly: both bodies, three FP64 divisions each, ran one
// after the other.  The two bodies are the same expressions with the roles of the real and imaginary parts exchanged (a + b == b + a
// and a * b == b * a hold bit for bit), so one copy with selected operands returns identical bits.
MX_HD Cplx c_div (Cplx z1, Cplx z2)
{
   Cplx z;
   const bool big = fabs (z2.r) > fabs (z2.i);
   const double a = big ? z2.r : z2.i, b = big ? z2.i : z2.r;       // divisor parts: the larger, the smaller
   const double u = big ? z1.r : z1.i, v = big ? z1.i : z1.r;
   const double ratio = b / a;
   const double denom = a + b * ratio;            // r2 + i2 * ratio | r2 * ratio + i2
   const double w = u * ratio;
   z.r = (u + ratio * v) / denom;                 // (r1 + ratio * i1) | (r1 * ratio + i1)
   z.i = (big ? (v - w) : (w - v)) / denom;       // (i1 - r1 * ratio) | (i1 * ratio - r1)
   return z;
}
MX_HD double c_abs (Cplx z)
{
   const double fr = fabs (z.r), fi = fabs (z.i);
   const bool big = fr > fi;
   if (!big && (fi == 0.0)) return 0.0;
   const double ratio = (big ? z.i : z.r) / (big ? z.r : z.i);
   return (big ? fr : fi) * sqrt (1.0 + ratio * ratio);
}
MX_HD Cplx c_sqrt (Cplx a)
{
   double r = c_abs (a);
   if (r == 0.0) return a;
   if (a.r >= 0.0)
     {
        a.r = sqrt (0.5 * (r + a.r));
        a.i = 0.5 * a.i / a.r;
     }
   else
     {
        r = sqrt (0.5 * (r - a.r));
        a.r = 0.5 * a.i / r;
        a.i = r;
        if (a.r < 0.0) { a.r = -a.r; a.i = -a.i; }
     }
   return a;
}
// JDMc_a_bz: a + b z ; JDMc_az1_bz2: a z1 + b z2 (complex.c:148-166)
MX_HD Cplx c_a_bz (double a, double b, Cplx z1) { Cplx z; z.r = a + b * z1.r; z.i = b * z1.i; return z; }
MX_HD Cplx c_az1_bz2 (double a, Cplx z1, double b, Cplx z2) { Cplx z; z.r = a * z1.r + b * z2.r; z.i = a * z1.i + b * z2.i; return z; }

// marx_reflectivity, reflect.c:39-77: polarisation-averaged Fresnel reflectivity, cos_theta >= 0
MX_HD_BIG double reflectivity (double cos_theta, double beta, double delta)
{
   Cplx n, root, nsqr, e_perp, e_par, num, den;
   n.r = (1.0 - delta);
   n.i = beta;
   double sin_theta = sqrt (1.0 - cos_theta * cos_theta);
   nsqr = c_mul (n, n);
   root = c_sqrt (c_a_bz (-sin_theta * sin_theta, 1.0, nsqr));
   num = c_a_bz (cos_theta, -1.0, root);
   den = c_a_bz (cos_theta, 1.0, root);
   e_perp = c_div (num, den);
   num = c_az1_bz2 (cos_theta, nsqr, -1.0, root);
   den = c_az1_bz2 (cos_theta, nsqr, 1.0, root);
   e_par = c_div (num, den);
   return 0.5 * (e_par.r * e_par.r + e_par.i * e_par.i + e_perp.r * e_perp.r + e_perp.i * e_perp.i);
}

// JDMquadratic_root, qroot.c:36-68 (a != 0 here); returns 1 real / 0 complex
MX_HD int quadratic_root (double a, double b, double c, double &rplus, double &rminus)
{
   double bsqr = b * b;
   double ac4 = a * c * 4;
   double neg_b_over_2a = -b / (2.0 * a);
   if (bsqr > ac4)
     {
        double factor = 1.0 + sqrt (1.0 - ac4 / bsqr);
        rplus = -2.0 * c / (b * factor);
        rminus = neg_b_over_2a * factor;
        return 1;
     }
   if (bsqr == ac4) { rplus = rminus = neg_b_over_2a; return 1; }
   return 0;
}

// compute_conic_intersection, hrma.c:411-478.  conic = {a, b, c, xmin, xmax}.  Line (not ray)
// intersection; of two in-range roots the one with the larger x wins.
template <bool WITH_NORMAL>
MX_HD int conic_intersection_t (const double *conic, Vec3 &x0, const Vec3 &p, Vec3 &normal)
{
   double a = conic[0], b = conic[1], c = conic[2], xmin = conic[3], xmax = conic[4];
   double t_plus, t_minus;
   t_plus = -x0.x / p.x;
   double x_y = x0.y + t_plus * p.y;
   double x_z = x0.z + t_plus * p.z;
   double alpha = a * p.x * p.x - 1.0;
   double beta = b * p.x - 2.0 * (p.y * x_y + p.z * x_z);
   double gamma = c - x_z * x_z - x_y * x_y;
   if (alpha == 0.0)
     {
        if (beta == 0.0) return -1;
        t_plus = t_minus = -gamma / beta;
     }
   else if (0 >= quadratic_root (alpha, beta, gamma, t_plus, t_minus))
     return -1;
   double x_plus = p.x * t_plus, x_minus = p.x * t_minus;
   if ((x_plus >= xmin) && (x_plus < xmax))
     {
        if ((x_minus >= xmin) && (x_minus < xmax) && (x_minus > x_plus))
          { x0.x = x_minus; x0.y = x_y + p.y * t_minus; x0.z = x_z + p.z * t_minus; }
        else
          { x0.x = x_plus; x0.y = x_y + p.y * t_plus; x0.z = x_z + p.z * t_plus; }
     }
   else if ((x_minus >= xmin) && (x_minus < xmax))
     { x0.x = x_minus; x0.y = x_y + p.y * t_minus; x0.z = x_z + p.z * t_minus; }
   else return -1;
   if (WITH_NORMAL)
     {
        normal.x = (a - 1) * x0.x + 0.5 * b;
        normal.y = -x0.y;
        normal.z = -x0.z;
        v_normalize (normal);
     }
   return 0;
}
MX_HD int conic_intersection (const double *conic, Vec3 &x0, const Vec3 &p, Vec3 &normal)
{
   return conic_intersection_t<true> (conic, x0, p, normal);
}

// blur_normal, hrma.c:1054-1093
MX_HD void blur_normal (Vec3 &n, double blur, Rng &rng)
{
   double n_y = n.y, n_z = n.z;
   double len = sqrt (n_y * n_y + n_z * n_z);
   Vec3 perp = v_make (0.0, n_z / len, -n_y / len);
   double phi = (2.0 * kPI) * rng.uniform ();
   perp = v_rotate_unit (perp, n, phi);
   phi = blur * (1.0 / 3600.0 * kPI / 180.0);
   phi = phi * rng.gaussian ();
   n = v_rotate_unit (n, perp, phi);
}

// interpolate_theta, wfold.c:304-332, for array k of table w
MX_HD double wfold_theta (const WfoldDev &w, uint32_t k, double p)
{
   const double *h = w.hdr + 6 * k;
   double p_min = h[1], delta_p = h[2], p_max = h[3];
   if (p < p_min) return 0.0;
   if (p > p_max) return pow (h[4] * (1.0 - p), h[5]);
   const float *t = w.theta + w.theta_offset[k];
   uint32_t nt = w.num_theta[k];
   double delta_i = (p - p_min) / delta_p;
   uint32_t i = (uint32_t) delta_i;
   if (i + 1 >= nt) return (double) t[nt - 1];
   delta_i -= (double) i;
   return (1.0 - delta_i) * t[i] + delta_i * t[i + 1];
}
// marx_wfold_table_interp, wfold.c:334-369
MX_HD_BIG double wfold_interp (const WfoldDev &w, double energy, double sin_alpha, double r, const double *keys = nullptr)
{
   if (w.num_arrays == 0) return 0.0;
   // (An early-out for r below every array's p_min -- 94 % of the draws -- was measured SLOWER: with 32 lanes per
   // warp some lane nearly always needs the full path, so the warp executes both.)
   if (w.num_arrays == 1) return wfold_theta (w, 0, r);
   double e_alpha = energy * sin_alpha;
   // JDMbinary_search_d over the e_alpha column: a contiguous copy staged in shared memory (`keys`, k1_hrma<1|2>) or
   // column 0 of the header rows in global memory (stride 6 doubles).
   const int ks = (keys != nullptr) ? 1 : 6;
   if (keys == nullptr) keys = w.hdr;
   const uint32_t n = w.num_arrays;
   uint32_t i = bsearch_fast<double> (e_alpha, keys, n, (uint32_t) ks);
   if (i == n) i--;
   if (i == 0) i++;
   double theta_0 = wfold_theta (w, i - 1, r);
   double theta_1 = wfold_theta (w, i, r);
   double e0 = keys[ks * (i - 1)], e1 = keys[ks * i];
   // 94 % of the draws lie below both arrays' p_min: both angles are 0, the interpolation term is 0 / (e1 - e0) = 0 and the sum theta_0
   if ((theta_1 == theta_0) && (e1 != e0)) return theta_0;
   return theta_0 + div_in_branch ((theta_1 - theta_0) * (e_alpha - e0), e1 - e0);
}

// intersects_struts, hrma.c:928-968.  struts = {xpos0, half_width0, xpos1, half_width1}
MX_HD_BIG int intersects_struts (const Vec3 &x0, const Vec3 &p0, double cap_position, const double *struts)
{
   // cos and sin of 30 degrees as the reference's libm returns them for theta = 30.0*(PI/180.0) = 0x1.0c152382d7365p-1
   // (hrma.c:930-932); spelled out so that no thread evaluates two trigonometric functions of a constant
   const double cos_theta = 0x1.bb67ae8584cabp-1, sin_theta = 0x1.fffffffffffffp-2;
   for (int s = 0; s < 2; s++)
     {
        double half_width = struts[2 * s + 1];
        double x = x0.x, y = x0.y, z = x0.z;
        double t = (struts[2 * s] + cap_position - x) / p0.x;
        y += p0.y * t; z += p0.z * t;
        for (int i = 0; i < 3; i++)
          {
             if (i != 0)
               {
                  double tmp = cos_theta * y - sin_theta * z;
                  z = sin_theta * y + cos_theta * z;
                  y = tmp;
               }
             if (((-half_width < y) && (y < half_width)) || ((-half_width < z) && (z < half_width)))
               return 1;
          }
     }
   return 0;
}

// hrma.c:908-926 (compile-time constants of the reference, not parameters)
#define MX_PRECOL_STRUTS  {1492.060, 0.5 * 0.5 * 25.4, 942.266, 0.5 * 0.5 * 25.4}
#define MX_CAP_STRUTS     {0.5 * 1.965 * 25.4, 0.5 * 0.75 * 25.4, -0.5 * 1.965 * 25.4, 0.5 * 0.75 * 25.4}
#define MX_POSTCOL_STRUTS {-1050.353, 0.5 * 0.5 * 25.4, -1271.333, 0.5 * 0.5 * 25.4}

// _marx_hrma_mirror_reflect for one ray, cut at the two points where most rays die so that the GPU can
// re-pack the survivors into full warps between the pieces (kernels k1a/k1b/k1c):
//   phase A  vignetting, shell + aperture point, precollimator struts, OSAC-P transform, P-conic intersection
//            (47.6 % of the rays of the C2 workload reach the end of A)
//   phase B  optical constants, P blur/reflect/scatter, back transform, CAP struts, OSAC-H transform, H-conic
//            intersection (27.3 % reach the end of B)
//   phase C  H blur/reflect/scatter, back transform, postcollimator struts (18.4 % survive)
// The conic normal is a pure function of the stored intersection point, so B and C recompute it from x
// bit-identically instead of carrying it.  hrma_reflect() = A;B;C is what the host-side checks step.

// inward unit normal of a conic at surface point x (tail of compute_conic_intersection, hrma.c:471-476)
MX_HD Vec3 conic_normal (const double *conic, const Vec3 &x)
{
   Vec3 n;
   n.x = (conic[0] - 1) * x.x + 0.5 * conic[1];
   n.y = -x.y;
   n.z = -x.z;
   v_normalize (n);
   return n;
}

// reflect_from_conic after the intersection test, hrma.c:499-544, in two halves (the mirror kernels can be cut between them:
// 35-45 % of the rays are absorbed at the reflectivity test, and the scatter + transforms behind it then run on full warps).
// first half: surface normal, blur, reflectivity test.  returns 0 ok, -1 absorbed; leaves the (blurred) normal in `normal`
MX_HD int reflect_test (const HrmaDev &H, const double *conic, double blur, double beta, double delta, double corr,
                        const Vec3 &x, const Vec3 &p, Vec3 &normal, Rng &rng)
{
   normal = conic_normal (conic, x);
   if (H.use_blur) blur_normal (normal, blur, rng);
   if (H.is_ideal == 0)
     {
        double p_dot_n = v_dot (p, normal);
        double r = rng.uniform ();
        double rfl = reflectivity (fabs (p_dot_n), beta, delta);
        if (r >= rfl * corr) return -1;
     }
   return 0;
}
// second half: specular reflection about `normal`, WFOLD scatter.  p_dot_n is the same dot product the first half formed.
MX_HD int reflect_scatter (const HrmaDev &H, const WfoldDev &wfold, double scat_factor, double energy, const Vec3 &normal,
                           Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   double p_dot_n = v_dot (p, normal);
   p = v_ax1_bx2 (1.0, p, -2.0 * p_dot_n, normal);
   if (H.use_wfold == 0) return 0;
   double sin_grazing = -p_dot_n;
   double r = rng.uniform ();
   double delta_grazing = wfold_interp (wfold, energy, sin_grazing, r, wfold_keys);
   delta_grazing *= scat_factor;
   if (delta_grazing > kPI / 4) return -1;
   if (rng.uniform () < 0.5) delta_grazing = -delta_grazing;
   p = v_rotate_unit (p, v_cross (p, normal), delta_grazing);
   return 0;
}
MX_HD int reflect_at_point (const HrmaDev &H, const double *conic, const WfoldDev &wfold, double scat_factor,
                            double blur, double energy, double beta, double delta, double corr,
                            const Vec3 &x, Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   Vec3 normal;
   if (0 != reflect_test (H, conic, blur, beta, delta, corr, x, p, normal, rng)) return -1;
   return reflect_scatter (H, wfold, scat_factor, energy, normal, p, rng, wfold_keys);
}

// phase A.  In: energy-independent; p from the source.  Out: x, p in the OSAC-P frame AT the P-conic
// intersection, shell.  Returns flags (0 = still alive).
MX_HD uint32_t hrma_phase_a (const HrmaDev &H, double source_distance, Vec3 &x, Vec3 &p, uint32_t &shell_out, Rng &rng)
{
   const uint32_t VBLOCKED = 0x10, UNREFLECTED = 0x02;
   // vignetting, hrma.c:1183-1192
   if (H.is_ideal == 0)
     {
        if (rng.uniform () > H.vig) return VBLOCKED;
     }
   // project_photon_to_hrma, hrma.c:984-1050
   // first shell i with r < area_fraction[i] (cumulative, non-decreasing), drawing again when r is beyond the last one
   // (hrma.c:990-1003); counted instead of scanned so that the lanes of a warp do not leave the loop one by one (k01: -1 %;
   // the same trick on the FEF component scan of k3_acis measured 5 % SLOWER and was dropped)
   uint32_t shell = 0;
   bool found = false;
   while (!found)
     {
        const double r = rng.uniform ();
        shell = 0;
#pragma unroll
        for (uint32_t i = 0; i + 1 < (uint32_t) kNumShells; i++) shell += (r < H.shell[i].area_fraction) ? 0u : 1u;
        found = (r < H.shell[kNumShells - 1].area_fraction);
     }
   const HrmaShellDev &h = H.shell[shell];
   shell_out = shell;
   {
      double radius = h.min_radius + (h.max_radius - h.min_radius) * rng.uniform ();
      double theta;
      uint32_t quad;
      do
        {
           theta = rng.uniform ();
           quad = (uint32_t) (4.0 * theta);
        }
      while (0 == (h.shutter_bitmap & (1u << quad)));
      theta = (2.0 * kPI) * (theta - 1.0 / 8.0);
      double st, ct;
      sin_cos (theta, st, ct);
      x.z = radius * ct;
      x.y = radius * st;
      x.x = h.front_position;
      x.z -= h.to_osac_p[2];
      x.y -= h.to_osac_p[1];
      if (source_distance > 0.0)
        {
           p = v_ax1_bx2 (1.0, x, source_distance, p);
           v_normalize (p);
        }
   }
   if (H.use_struts)
     {
        const double st[4] = MX_PRECOL_STRUTS;
        if (intersects_struts (x, p, H.cap_position, st)) return VBLOCKED;
     }
   // to OSAC P frame, hrma.c:1222-1235
   x = v_sum (x, v_make (h.to_osac_p[0], h.to_osac_p[1], h.to_osac_p[2]));
   x = m3_mul (h.fwd_p, x);
   p = m3_mul (h.fwd_p, p);
   Vec3 normal;          // recomputed from x by phase B (conic_normal)
   if (-1 == conic_intersection_t<false> (h.conic_p, x, p, normal)) return UNREFLECTED;
   return 0;
}

// optical constants and effective-area correction, hrma.c:1239-1262.  beta/delta/corr are float-valued
// table interpolations (finterpo.c:59-85); the square root of corr is taken at use (hrma.c:1258-1259).
MX_HD void hrma_optical_constants (const HrmaDev &H, const HrmaShellDev &h, const float *opt_e, const float *opt_b,
                                   const float *opt_d, const float *corr_e, const float *corr_f, double energy,
                                   float &beta, float &delta, float &corr)
{
   beta = 0.0f; delta = 1.0f; corr = 1.0f;
   if (H.num_opt != 0)
     {
        float ef = (float) energy;
        interp_f2 (ef, opt_e, opt_b, opt_d, H.num_opt, beta, delta);
        if (H.use_scale) corr = interp_f (ef, corr_e + h.corr_offset, corr_f + h.corr_offset, h.num_corr);
     }
}

// phase B.  In: x, p at the P intersection (OSAC-P frame).  Out: x, p at the H intersection (OSAC-H frame).
// B = B1 (normal, blur, reflectivity test at the P conic; leaves the blurred normal) ; B2 (reflection, scatter, back transform,
// CAP struts, OSAC-H transform, H-conic intersection)
MX_HD uint32_t hrma_phase_b1 (const HrmaDev &H, uint32_t shell, float beta, float delta, float corr_f32,
                              const Vec3 &x, const Vec3 &p, Vec3 &normal, Rng &rng)
{
   const uint32_t UNREFLECTED = 0x02;
   const HrmaShellDev &h = H.shell[shell];
   double corr = (H.num_opt != 0 && H.use_scale) ? sqrt ((double) corr_f32) : 1.0;
   if (0 != reflect_test (H, h.conic_p, h.p_blur, beta, delta, corr, x, p, normal, rng)) return UNREFLECTED;
   return 0;
}
MX_HD uint32_t hrma_phase_b2 (const HrmaDev &H, uint32_t shell, double energy, const Vec3 &normal_p,
                              Vec3 &x, Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   const uint32_t VBLOCKED = 0x10, UNREFLECTED = 0x02;
   const HrmaShellDev &h = H.shell[shell];
   if (0 != reflect_scatter (H, h.wfold_p, h.p_scat, energy, normal_p, p, rng, wfold_keys)) return UNREFLECTED;
   const Vec3 to_p = v_make (h.to_osac_p[0], h.to_osac_p[1], h.to_osac_p[2]);
   const Vec3 to_h = v_make (h.to_osac_h[0], h.to_osac_h[1], h.to_osac_h[2]);
   p = m3_mul (h.bwd_p, p);
   x = m3_mul (h.bwd_p, x);
   x = v_diff (x, to_p);
   if (H.use_struts)
     {
        const double st[4] = MX_CAP_STRUTS;
        if (intersects_struts (x, p, H.cap_position, st)) return VBLOCKED;
     }
   x = v_sum (x, to_h);
   p = m3_mul (h.fwd_h, p);
   x = m3_mul (h.fwd_h, x);
   Vec3 normal;          // recomputed from x by phase C
   if (-1 == conic_intersection_t<false> (h.conic_h, x, p, normal)) return UNREFLECTED;
   return 0;
}
MX_HD uint32_t hrma_phase_b (const HrmaDev &H, uint32_t shell, double energy, float beta, float delta, float corr_f32,
                             Vec3 &x, Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   Vec3 normal;
   const uint32_t flags = hrma_phase_b1 (H, shell, beta, delta, corr_f32, x, p, normal, rng);
   if (flags) return flags;
   return hrma_phase_b2 (H, shell, energy, normal, x, p, rng, wfold_keys);
}

// phase C.  In: x, p at the H intersection (OSAC-H frame).  Out: x, p in MARX coordinates behind the mirror.
// C = C1 (normal, blur, reflectivity test at the H conic; leaves the blurred normal) ; C2 (reflection, scatter, back transform,
// postcollimator struts)
MX_HD uint32_t hrma_phase_c1 (const HrmaDev &H, uint32_t shell, float beta, float delta, float corr_f32,
                              const Vec3 &x, const Vec3 &p, Vec3 &normal, Rng &rng)
{
   const uint32_t UNREFLECTED = 0x02;
   const HrmaShellDev &h = H.shell[shell];
   double corr = (H.num_opt != 0 && H.use_scale) ? sqrt ((double) corr_f32) : 1.0;
   if (0 != reflect_test (H, h.conic_h, h.h_blur, beta, delta, corr, x, p, normal, rng)) return UNREFLECTED;
   return 0;
}
MX_HD uint32_t hrma_phase_c2 (const HrmaDev &H, uint32_t shell, double energy, const Vec3 &normal_h,
                              Vec3 &x, Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   const uint32_t VBLOCKED = 0x10, UNREFLECTED = 0x02;
   const HrmaShellDev &h = H.shell[shell];
   if (0 != reflect_scatter (H, h.wfold_h, h.h_scat, energy, normal_h, p, rng, wfold_keys)) return UNREFLECTED;
   p = m3_mul (h.bwd_h, p);
   x = m3_mul (h.bwd_h, x);
   x = v_diff (x, v_make (h.to_osac_h[0], h.to_osac_h[1], h.to_osac_h[2]));
   if (H.use_struts)
     {
        const double st[4] = MX_POSTCOL_STRUTS;
        if (intersects_struts (x, p, H.cap_position, st)) return VBLOCKED;
     }
   return 0;
}
MX_HD uint32_t hrma_phase_c (const HrmaDev &H, uint32_t shell, double energy, float beta, float delta, float corr_f32,
                             Vec3 &x, Vec3 &p, Rng &rng, const double *wfold_keys = nullptr)
{
   Vec3 normal;
   const uint32_t flags = hrma_phase_c1 (H, shell, beta, delta, corr_f32, x, p, normal, rng);
   if (flags) return flags;
   return hrma_phase_c2 (H, shell, energy, normal, x, p, rng, wfold_keys);
}

// the whole stage for one ray (A;B;C) -- used by the developer host check; the kernels call the phases
MX_HD uint32_t hrma_reflect (const HrmaDev &H, const float *opt_e, const float *opt_b, const float *opt_d,
                             const float *corr_e, const float *corr_f,
                             double source_distance, double energy, Vec3 &x, Vec3 &p, uint32_t &shell_out, Rng &rng)
{
   uint32_t flags = hrma_phase_a (H, source_distance, x, p, shell_out, rng);
   if (flags) return flags;
   float beta, delta, corr;
   hrma_optical_constants (H, H.shell[shell_out], opt_e, opt_b, opt_d, corr_e, corr_f, energy, beta, delta, corr);
   flags = hrma_phase_b (H, shell_out, energy, beta, delta, corr, x, p, rng);
   if (flags) return flags;
   return hrma_phase_c (H, shell_out, energy, beta, delta, corr, x, p, rng);
}

}  // namespace mx
