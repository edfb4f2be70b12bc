// mx_grating.cuh -- K2: HETG/LETG facet diffraction, per ray.
// Reference: marx/libsrc/diffract.c:974-1130 (diffract), :622-687 (newtons_quartic, intersect),
// :689-700 (rotate), :706-823 (diffract_photon), :828-849 (diffract_photon_from_grating);
// cumulative-efficiency interpolation jdmath/src/finterpo.c:134-189 (JDMinterpolate_n_fvector).
// Draw order on sub-stream MARXB200_STAGE_GRATING: vignetting U; order U; 2 G (dtheta, dp/p).
// A ray stops at its FIRST cause of death (the reference lets torus misses draw on, diffract.c:906-911
// vs :1043-1076, but never outputs them).
#pragma once
#include "mx_common.cuh"
#include "mx_tables.h"

namespace mx {

// rotate about the x axis, diffract.c:689-700
MX_HD Vec3 rotate_x (Vec3 a, double c, double s)
{
   double ay = a.y, az = a.z;
   a.y = c * ay - s * az;
   a.z = s * ay + c * az;
   return a;
}

// newtons_quartic, diffract.c:622-652
MX_HD int newtons_quartic (double a, double b, double c, double d, double t0, double &tp)
{
   unsigned int max_it = 10;
   const double eps = 1.0e-4;
   double a2 = 2.0 * a, a3 = 3.0 * a, b2 = 2.0 * b, t;
   while (1)
     {
        double t2 = t0 * t0;
        double num = t2 * (3.0 * t2 + a2 * t0 + b) - d;
        double den = t2 * (4.0 * t0 + a3) + b2 * t0 + c;
        t = num / den;
        if (fabs (t - t0) < eps) break;
        max_it--;
        if (max_it == 0) return -1;
        t0 = t;
     }
   tp = t;
   return 0;
}

// intersect, diffract.c:654-687: Rowland torus of diameter `rowland`
MX_HD int torus_intersect (Vec3 &x0, const Vec3 &p, double rowland)
{
   double t = -x0.x / p.x;
   x0.x = 0.0;
   x0.y = x0.y + p.y * t;
   x0.z = x0.z + p.z * t;
   double pxpz_len = p.x * p.x + p.z * p.z;
   double x2 = x0.z * x0.z + x0.y * x0.y;
   double r2 = rowland * rowland;
   double pdotx = v_dot (p, x0);
   double a = 4.0 * pdotx;
   double b = 2.0 * x2 + a * pdotx - r2 * pxpz_len;
   double c = a * x2 - 2.0 * r2 * p.z * x0.z;
   double d = x2 * x2 - r2 * x0.z * x0.z;
   double t0 = -rowland * sqrt (pxpz_len);
   if (-1 == newtons_quartic (a, b, c, d, t0, t)) return -1;
   x0 = v_ax1_bx2 (1.0, x0, t, p);
   return 0;
}

// JDMv_unit_vector
MX_HD Vec3 v_unit (Vec3 a) { v_normalize (a); return a; }

// diffract_photon, diffract.c:706-823.  theta: extra facet rotation (support gratings); 0 for the
// primary grating.  Sector tables (when present) give the facet misalignment; they ARE active for the
// default HETG configuration.
MX_HD int diffract_photon (const GratingShellDev &g, double theta, double energy, const Vec3 &x, Vec3 &pio,
                           int order, bool use_sectors, Rng &rng)
{
   // order 0 (half of the rays that reach this point): 0 * c / period / energy is +0 for positive period and energy
   double n_lambda_over_d = 0.0;
   if (!((order == 0) && (g.period > 0.0) && (energy > 0.0)))
     n_lambda_over_d = div_in_branch (div_in_branch (order * (2.0 * kPI * kHbarC), g.period), energy);
   Vec3 p = pio;
   Vec3 n = v_unit (x);
   n.x = -n.x; n.y = -n.y; n.z = -n.z;
   Vec3 l = v_make (n.z, 0.0, -n.x);
   l = v_unit (l);
   Vec3 d = v_cross (n, l);
   double dtheta, dp_over_p;
   if (use_sectors)
     {
        const uint32_t ns = g.num_sectors;
        const double *min_angle = g.sectors, *max_angle = g.sectors + ns;
        double sector = atan2 (x.y, x.z);
        if (sector < 0) sector = 2 * kPI + sector;
        uint32_t sector_num = bsearch_d (sector, min_angle, ns);
        if (sector_num == 0) return -1;      // reference would index [-1]; unreachable in practice (needs x.y == 0 exactly)
        sector_num = sector_num - 1;
        if ((max_angle[sector_num] <= sector) || (min_angle[sector_num] > sector)) return -1;
        dtheta = g.sectors[2 * ns + sector_num] + g.sectors[3 * ns + sector_num] * rng.gaussian ();
        dp_over_p = g.sectors[4 * ns + sector_num] + g.sectors[5 * ns + sector_num] * rng.gaussian ();
     }
   else
     {
        dtheta = g.theta_blur * rng.gaussian ();
        dp_over_p = g.dp_over_p * rng.gaussian ();
     }
   theta -= dtheta;
   if (theta != 0.0)
     {
        Vec3 l_tmp = l, d_tmp = d;
        double c, s;
        sin_cos (theta, s, c);
        l = v_ax1_bx2 (c, l_tmp, s, d_tmp);
        d = v_ax1_bx2 (-s, l_tmp, c, d_tmp);
     }
   double p_d = n_lambda_over_d + v_dot (p, d);
   double p_l = v_dot (p, l);
   double p_n = 1.0 - p_l * p_l - p_d * p_d;
   if (p_n < 0.0) return -1;
   p_n = sqrt (p_n);
   p = v_ax1_bx2 (p_d, d, p_n, n);
   pio = v_ax1_bx2 (1.0, p, p_l, l);
   if (dp_over_p == 0) return 0;
   double factor = n_lambda_over_d * dp_over_p;
   Vec3 dp = v_ax1_bx2 (-factor, d, factor * (p_d / p_n), n);
   pio.x += dp.x; pio.y += dp.y; pio.z += dp.z;
   pio = v_unit (pio);
   return 0;
}

// bracket used by JDMinterpolate_n_fvector (finterpo.c:154-164) for a single abscissa: the first index
// c in [1, n-1] with xp[c] >= x, else n-1.  (The reference walks linearly through its energy-sorted
// batch; for one value the walk and this lower bound select the same c.)
MX_HD uint32_t nfvector_bracket (double x, const float *xp, uint32_t n)
{
   uint32_t lo = 1, hi = n - 1;       // answer in [lo, hi]
   while (lo < hi)
     {
        uint32_t mid = (lo + hi) / 2;
        if ((double) xp[mid] >= x) hi = mid; else lo = mid + 1;
     }
   return lo;
}

// order selection of one grating (diffract_photon_from_grating, diffract.c:828-849, with the per-photon row of cumulative
// efficiencies produced as in finterpo.c:166-186: double arithmetic, stored to float) for the uniform deviate r.
// cum_eff_t is TRANSPOSED on upload to [energy][order] so that the order scan reads two contiguous rows.
// Returns the index into order_list, or -1 when the photon leaves in none of the tabulated orders (absorbed).
MX_HD int select_order (const GratingShellDev &g, double energy, double r)
{
   double xe = (double) (float) energy;           // tmp_energies[] is float (diffract.c:1047)
   uint32_t c = nfvector_bracket (xe, g.energies, g.num_energies);
   double x_0 = g.energies[c - 1], x_1 = g.energies[c];
   double dx_10 = x_1 - x_0;
   const float *row0 = g.cum_eff + (size_t) (c - 1) * g.num_orders;
   const float *row1 = g.cum_eff + (size_t) c * g.num_orders;
   // The reference scans the orders linearly for the first k with r <= cum_eff[k] (diffract.c:839-847).  The
   // interpolated cumulative efficiencies are non-decreasing in k (a convex combination of two non-decreasing
   // table rows, then a monotone rounding to float), so a bisection finds the same k in log2(num_orders)
   // interpolations instead of num_orders/2 -- 7 instead of ~60 for the 121-order LETG coarse support grating.
   uint32_t lo = 0, hi = g.num_orders;
   while (lo < hi)
     {
        const uint32_t k = (lo + hi) >> 1;
        float ce;
        if (dx_10 == 0.0) ce = row0[k];
        else
          {
             double y_0 = row0[k], y_1 = row1[k];
             ce = (float) (y_0 + (y_1 - y_0) * (xe - x_0) / dx_10);
          }
        if (r <= ce) hi = k; else lo = k + 1;
     }
   return (lo == g.num_orders) ? -1 : (int) lo;
}

// order selection + diffraction from one grating
MX_HD int diffract_from_grating (const GratingShellDev &g, double theta, double energy, const Vec3 &x, Vec3 &p,
                                 int &order_out, bool use_sectors, Rng &rng)
{
   const int lo = select_order (g, energy, rng.uniform ());
   if (lo < 0) return -1;
   int order = g.order_list[lo];
   order_out = order;
   return diffract_photon (g, theta, energy, x, p, order, use_sectors, rng);
}

// diffract() for one ray of shell `shell`.  HETG: the primary grating; LETG: the primary grating followed by the
// fine support grating (facet rotated by pi/2) and three passes through the coarse support grating (pi/3,
// 2pi/3, 0), diffract.c:1098-1118.  support_orders packs the four support orders (one signed byte each).
// The stage for one ray is  vignetting test (draw 1) ; facet frame + Rowland torus ; order selection of the primary grating
// (draw 2) ; diffraction, support gratings, back rotation.  The two tests absorb half of the rays and need neither x nor p, so
// the compacting path runs them first, as a kernel of their own (grating_select), and the geometry on re-packed warps
// (grating_diffract_selected); a survivor sees the same draws either way (the torus takes none).
// Returns flags (0 alive).
MX_HD uint32_t grating_finish (const GratingDev &G, const GratingShellDev &g, double energy, Vec3 &x, Vec3 &p, int lo,
                               int &order_out, uint32_t &support_orders, Rng &rng)
{
   const uint32_t UNDIFFRACTED = 0x04;
   const int order = g.order_list[lo];
   order_out = order;
   if (-1 == diffract_photon (g, 0.0, energy, x, p, order, g.num_sectors != 0, rng)) return UNDIFFRACTED;
   support_orders = 0;
   if (G.type == 2)
     {
        const double pass_theta[4] = {kPI / 2.0, kPI / 3.0, 2.0 * kPI / 3.0, 0.0};
        for (int pass = 0; pass < 4; pass++)
          {
             const GratingShellDev &sg = G.support[pass == 0 ? 0 : 1];
             if (sg.num_orders == 0) continue;
             int so = 0;
             int rc = diffract_from_grating (sg, pass_theta[pass], energy, x, p, so, false, rng);
             support_orders |= ((uint32_t) (so & 0xFF)) << (8 * pass);
             if (rc == -1) return UNDIFFRACTED;
          }
     }
   x = rotate_x (x, g.cos_dispersion, g.sin_dispersion);
   p = rotate_x (p, g.cos_dispersion, g.sin_dispersion);
   return 0;
}
MX_HD uint32_t grating_diffract (const GratingDev &G, uint32_t shell, double energy, Vec3 &x, Vec3 &p,
                                 int &order_out, uint32_t &support_orders, Rng &rng)
{
   const uint32_t VBLOCKED = 0x10, UNDIFFRACTED = 0x04;
   const GratingShellDev &g = G.shell[shell];
   // vignetting (flagged MIRROR_VBLOCKED by the reference, diffract.c:999-1000)
   if (rng.uniform () > g.vig) return VBLOCKED;
   // rotate_photons(pt, -1), diffract.c:854-875: the angle is a per-shell constant, its cosine and sine are tabulated
   // by the host (tables_build.hpp, the reference's own libm); cos(-t) = cos(t), sin(-t) = -sin(t) exactly
   x = rotate_x (x, g.cos_dispersion, -g.sin_dispersion);
   p = rotate_x (p, g.cos_dispersion, -g.sin_dispersion);
   if (-1 == torus_intersect (x, p, g.rowland)) return UNDIFFRACTED;
   const int lo = select_order (g, energy, rng.uniform ());
   if (lo < 0) return UNDIFFRACTED;
   return grating_finish (G, g, energy, x, p, lo, order_out, support_orders, rng);
}
// first half on the compacting path: draws 1 and 2 of the GRATING sub-stream.  -> index into order_list, or -1 (vignetted / absorbed)
MX_HD int grating_select (const GratingShellDev &g, double energy, Rng &rng)
{
   if (rng.uniform () > g.vig) return -1;
   return select_order (g, energy, rng.uniform ());
}
// second half: rng resumed behind draw 2
MX_HD uint32_t grating_diffract_selected (const GratingDev &G, uint32_t shell, double energy, Vec3 &x, Vec3 &p, int lo,
                                          int &order_out, uint32_t &support_orders, Rng &rng)
{
   const uint32_t UNDIFFRACTED = 0x04;
   const GratingShellDev &g = G.shell[shell];
   x = rotate_x (x, g.cos_dispersion, -g.sin_dispersion);
   p = rotate_x (p, g.cos_dispersion, -g.sin_dispersion);
   if (-1 == torus_intersect (x, p, g.rowland)) return UNDIFFRACTED;
   return grating_finish (G, g, energy, x, p, lo, order_out, support_orders, rng);
}

}  // namespace mx
