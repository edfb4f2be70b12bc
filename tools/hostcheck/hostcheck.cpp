// hostcheck -- DEVELOPER-ONLY harness (never built into libmarxb200, never imported by the package,
// never used by tests/bench as a compute path).  It steps the MX_HD per-ray stage functions of
// marx_b200/csrc/*.cuh on the host, one ray at a time, against a replay dump written by
// oracle/_ref/marx_replay, so that physics/parity bugs can be found in this GPU-less container before
// spending GPU minutes.  The shipped library only ever runs these functions inside __global__ kernels.
//
// usage: hostcheck CALPACK REPLAY.bin [max_report]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/marxb200.h"
#include "../../marx_b200/csrc/mx_common.cuh"
#include "../../marx_b200/csrc/mx_tables.h"
#include "../../marx_b200/csrc/mx_source.cuh"
#include "../../marx_b200/csrc/mx_hrma.cuh"
#include "../../marx_b200/csrc/mx_grating.cuh"
#include "../../marx_b200/csrc/mx_acis.cuh"
#include "../../marx_b200/csrc/tables_build.hpp"

using namespace mx;

struct marxb200_ctx { int dummy; };
static SourceDev gS; static DitherDev gD;
static std::vector<unsigned char> gB1, gB2, gB3;
static int gGratingType = 0, gDetType = 0;
static char gErr[256];

struct HostUploader
{
   const void *operator() (const void *host, size_t bytes)
   {
      void *d = malloc (bytes ? bytes : 16);
      if (host && bytes) memcpy (d, host, bytes);
      return d;
   }
};

extern "C" const char *marxb200_last_error (void) { return gErr; }
extern "C" int marxb200_set_source (marxb200_ctx *, const marxb200_source_desc *d)
{
   memset (&gS, 0, sizeof (gS));
   gS.source_type = d->source_type; gS.spectrum_type = d->spectrum_type;
   for (int i = 0; i < 3; i++) { gS.p[i] = d->p[i]; gS.p_normal[i] = d->p_normal[i]; }
   gS.distance = d->distance; gS.emin = d->emin; gS.emax = d->emax;
   for (int i = 0; i < 3; i++) gS.shape[i] = d->shape[i];
   gS.spec_energies = d->spec_energies; gS.spec_cum_flux = d->spec_cum_flux; gS.spec_num = d->spec_num;
   gS.mean_time = (d->total_flux <= 0.0) ? 0.0 : 1.0 / d->total_flux / d->geometric_area;
   return 0;
}
extern "C" int marxb200_set_dither (marxb200_ctx *, const marxb200_dither_desc *d)
{
   gD.mode = d->mode; gD.ra_amp = d->ra_amp; gD.dec_amp = d->dec_amp; gD.roll_amp = d->roll_amp;
   gD.ra_period = d->ra_period; gD.dec_period = d->dec_period; gD.roll_period = d->roll_period;
   gD.ra_phase = d->ra_phase; gD.dec_phase = d->dec_phase; gD.roll_phase = d->roll_phase;
   gD.nominal_roll = d->nominal_roll; gD.aspect_blur = d->aspect_blur;
   gD.aspsol = nullptr; gD.num_aspsol = 0;
   if (d->mode == 2)
     {
        double *t = (double *) malloc (7 * sizeof (double) * d->num_aspsol);
        memcpy (t, d->aspsol, 7 * sizeof (double) * d->num_aspsol);
        gD.aspsol = t; gD.num_aspsol = d->num_aspsol;
     }
   return 0;
}
extern "C" int marxb200_set_flatfield (marxb200_ctx *, const marxb200_flatfield_desc *) { return 0; }
extern "C" int marxb200_set_hrma (marxb200_ctx *, const marxb200_hrma_desc *d)
{ HostUploader up; std::string e; int r = build_hrma_blob (up, d, gB1, e); snprintf (gErr, sizeof gErr, "%s", e.c_str ()); return r; }
extern "C" int marxb200_set_grating (marxb200_ctx *, const marxb200_grating_desc *d)
{
   gGratingType = d->type; if (d->type == 0) return 0;
   HostUploader up; std::string e; int r = build_grating_blob (up, d, gB2, e); snprintf (gErr, sizeof gErr, "%s", e.c_str ());
   if (r == 0)
     {
        K2Blob *B = (K2Blob *) gB2.data ();
        for (int k = 0; k < kNumShells; k++) B->G.shell[k].sectors = (const double *) (gB2.data () + B->off_sectors[k]);
     }
   return r;
}
extern "C" int marxb200_set_acis (marxb200_ctx *, const marxb200_acis_desc *d)
{
   gDetType = d->detector_type; if (d->detector_type == 0) return 0;
   HostUploader up; std::string e; int r = build_acis_blob (up, d, gB3, e); snprintf (gErr, sizeof gErr, "%s", e.c_str ()); return r;
}
extern "C" int marxb200_set_hrc_s (marxb200_ctx *, const marxb200_hrc_s_desc *) { snprintf (gErr, sizeof gErr, "hostcheck: HRC-S not wired"); return -1; }
extern "C" int marxb200_load_calpack_impl (marxb200_ctx *ctx, const char *path, char *errbuf, size_t errlen);

#pragma pack(push, 1)
struct Rec { marxb200_photon_attr st[4]; uint32_t draws[4]; double start; };
#pragma pack(pop)

static double relerr (double a, double b)
{
   double d = fabs (a - b), s = fmax (fabs (a), fabs (b));
   return (s > 0) ? d / s : 0.0;
}
struct Stat { double max_rel = 0; long n = 0, bad = 0; void add (double r, double tol) { n++; if (r > max_rel) max_rel = r; if (r > tol) bad++; } };

int main (int argc, char **argv)
{
   if (argc < 3) { fprintf (stderr, "usage: %s CALPACK REPLAY.bin\n", argv[0]); return 2; }
   marxb200_ctx ctx; char err[512];
   if (-1 == marxb200_load_calpack_impl (&ctx, argv[1], err, sizeof err)) { fprintf (stderr, "%s\n", err); return 1; }
   FILE *fp = fopen (argv[2], "rb");
   if (!fp) { perror (argv[2]); return 1; }
   char magic[8]; uint64_t nrays, seed, first; uint32_t hdr2[2]; int32_t mods[4];
   if (fread (magic, 1, 8, fp) != 8 || fread (&nrays, 8, 1, fp) != 1 || fread (&seed, 8, 1, fp) != 1 || fread (&first, 8, 1, fp) != 1
       || fread (hdr2, 4, 2, fp) != 2 || fread (mods, 4, 4, fp) != 4) return 1;
   printf ("replay: %llu rays seed %llu first %llu recsize %u\n", (unsigned long long) nrays, (unsigned long long) seed, (unsigned long long) first, hdr2[1]);
   int max_report = argc > 3 ? atoi (argv[3]) : 10;

   const K1Blob *B1 = (const K1Blob *) gB1.data ();
   const HrmaDev &H = B1->H;
   const float *opt_e = (const float *) (gB1.data () + B1->off_opt_e), *opt_b = (const float *) (gB1.data () + B1->off_opt_b), *opt_d = (const float *) (gB1.data () + B1->off_opt_d);
   const float *corr_e = (const float *) (gB1.data () + B1->off_corr_e), *corr_f = (const float *) (gB1.data () + B1->off_corr_f);

   Stat s_energy, s_p0, s_time, s_x1, s_p1, s_x2, s_p2, s_x3, s_p3, s_pix, s_pi, s_dith;
   long flag_mis[4] = {0, 0, 0, 0}, alive[4] = {0, 0, 0, 0}, int_mis = 0, draw_mis[4] = {0, 0, 0, 0}, reported = 0;
   double t_run = 0.0;
   const double tol = 1e-9;
   for (uint64_t i = 0; i < nrays; i++)
     {
        Rec r;
        if (fread (&r, sizeof (Rec), 1, fp) != 1) { fprintf (stderr, "short read at %llu\n", (unsigned long long) i); return 1; }
        uint64_t ray = first + i;
        // ---- K0 ----
        Rng rng; rng.init (seed, ray, 0);
        double energy; Vec3 p, x = v_make (0, 0, 0);
        source_draw (gS, rng, energy, p);
        double dt = source_time_increment (gS, rng);
        double t_abs = t_run + dt;              // reference: start_time (running) + arrival_time
        t_run = t_abs;
        float dra, ddec, droll, det[3];
        dither_ray (gD, rng, t_abs, p, dra, ddec, droll, nullptr, det);
        const marxb200_photon_attr &a0 = r.st[0];
        s_energy.add (relerr (energy, a0.energy), 0);
        s_time.add (relerr (t_abs, r.start + a0.arrival_time), 1e-12);
        for (int k = 0; k < 3; k++) s_p0.add (fabs ((&p.x)[k] - a0.p[k]), tol);
        s_dith.add (fmax (fabs (dra - a0.dither_ra), fmax (fabs (ddec - a0.dither_dec), fabs (droll - a0.dither_roll))), 0);
        s_dith.add (fmax (fabs (det[0] - a0.dither_dy), fmax (fabs (det[1] - a0.dither_dz), fabs (det[2] - a0.dither_dtheta))), 0);
        if (rng.draw != r.draws[0]) draw_mis[0]++;
        alive[0]++;
        // use the reference's time downstream so that one ulp of time never masks a stage bug
        t_abs = r.start + a0.arrival_time;
        // ---- K1 ----
        uint32_t shell = 0;
        rng.init (seed, ray, 1);
        uint32_t flags = hrma_reflect (H, opt_e, opt_b, opt_d, corr_e, corr_f, gS.distance, energy, x, p, shell, rng);
        const marxb200_photon_attr &a1 = r.st[1];
        bool ref_alive = (a1.flags & 0xFF) == 0, my_alive = (flags & 0xFF) == 0;
        if (ref_alive != my_alive || (!my_alive && ((a1.flags & flags) != flags)))
          { flag_mis[1]++; if (reported++ < max_report) printf ("ray %llu K1 flags mine %x ref %x\n", (unsigned long long) ray, flags, a1.flags); continue; }
        if (!my_alive) continue;
        alive[1]++;
        if (rng.draw != r.draws[1]) draw_mis[1]++;
        if (shell != a1.mirror_shell) int_mis++;
        for (int k = 0; k < 3; k++) { s_x1.add (relerr ((&x.x)[k], a1.x[k]), tol); s_p1.add (fabs ((&p.x)[k] - a1.p[k]), tol); }
        // ---- K2 ----
        int order = 0;
        if (gGratingType)
          {
             const GratingDev &G = ((const K2Blob *) gB2.data ())->G;
             rng.init (seed, ray, 2);
             uint32_t sorders = 0;
             flags = grating_diffract (G, shell, energy, x, p, order, sorders, rng);
             const marxb200_photon_attr &a2 = r.st[2];
             ref_alive = (a2.flags & 0xFF) == 0; my_alive = (flags & 0xFF) == 0;
             if (ref_alive != my_alive || (!my_alive && ((a2.flags & flags) != flags)))
               { flag_mis[2]++; if (reported++ < max_report) printf ("ray %llu K2 flags mine %x ref %x\n", (unsigned long long) ray, flags, a2.flags); continue; }
             if (!my_alive) continue;
             alive[2]++;
             if (rng.draw != r.draws[2]) draw_mis[2]++;
             if (order != a2.order) { int_mis++; if (reported++ < max_report) printf ("ray %llu order mine %d ref %d\n", (unsigned long long) ray, order, a2.order); }
             for (int k = 0; k < 3; k++) { s_x2.add (relerr ((&x.x)[k], a2.x[k]), tol); s_p2.add (fabs ((&p.x)[k] - a2.p[k]), tol); }
          }
        // ---- K3 ----
        if (gDetType)
          {
             const AcisDev &A = ((const K3Blob *) gB3.data ())->A;
             int ccd = -1; float chipx = 0, chipy = 0, pi = 0; int16_t pha = 0;
             rng.init (seed, ray, 3);
             { float fef_cum[mx::kMaxGauss]; DetDither dd = {a0.dither_dy, a0.dither_dz, a0.dither_dtheta};
               flags = acis_detect<true> (A, energy, t_abs, x, p, ccd, chipx, chipy, pha, pi, rng, fef_cum, 1, dd); }
             const marxb200_photon_attr &a3 = r.st[3];
             ref_alive = (a3.flags & 0xFF) == 0; my_alive = (flags & 0xFF) == 0;
             if (ref_alive != my_alive || (!my_alive && ((a3.flags & flags) != flags)) || (my_alive && flags != a3.flags))
               { flag_mis[3]++; if (reported++ < max_report) printf ("ray %llu K3 flags mine %x ref %x\n", (unsigned long long) ray, flags, a3.flags); continue; }
             if (!my_alive) continue;
             alive[3]++;
             if (rng.draw != r.draws[3]) draw_mis[3]++;
             if (ccd != a3.ccd_num || pha != a3.pulse_height)
               { int_mis++; if (reported++ < max_report) printf ("ray %llu ccd %d/%d pha %d/%d\n", (unsigned long long) ray, ccd, a3.ccd_num, pha, a3.pulse_height); }
             s_pix.add (fmax (relerr (chipx, a3.y_pixel), relerr (chipy, a3.z_pixel)), 1e-6);
             s_pi.add (relerr (pi, a3.pi), 1e-6);
             for (int k = 0; k < 3; k++) { s_x3.add (relerr ((&x.x)[k], a3.x[k]), tol); s_p3.add (fabs ((&p.x)[k] - a3.p[k]), tol); }
          }
     }
   printf ("alive: gen %ld mirror %ld grating %ld detected %ld\n", alive[0], alive[1], alive[2], alive[3]);
   printf ("flag mismatches: K1 %ld K2 %ld K3 %ld ; integer mismatches %ld ; draw-count mismatches %ld %ld %ld %ld\n",
           flag_mis[1], flag_mis[2], flag_mis[3], int_mis, draw_mis[0], draw_mis[1], draw_mis[2], draw_mis[3]);
#define SHOW(s) printf ("  %-10s n=%ld max_rel=%.3e over_tol=%ld\n", #s, s.n, s.max_rel, s.bad)
   SHOW (s_energy); SHOW (s_time); SHOW (s_p0); SHOW (s_dith); SHOW (s_x1); SHOW (s_p1); SHOW (s_x2); SHOW (s_p2); SHOW (s_x3); SHOW (s_p3); SHOW (s_pix); SHOW (s_pi);
   long bad = flag_mis[1] + flag_mis[2] + flag_mis[3] + int_mis + s_x1.bad + s_p1.bad + s_x2.bad + s_p2.bad + s_x3.bad + s_p3.bad + s_pix.bad + s_pi.bad + s_p0.bad + s_dith.bad + s_energy.bad;
   printf ("%s\n", bad ? "HOSTCHECK: MISMATCHES" : "HOSTCHECK: OK");
   return bad ? 1 : 0;
}
