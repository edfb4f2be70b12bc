// pileup_hostcheck -- DEVELOPER-ONLY harness (never built into libmarxb200, never imported by the package, never used by
// tests/bench as a compute path).  It steps the MX_HD per-event functions of marx_b200/csrc/mx_pileup.cuh on the host, one
// event at a time in the order the kernels of pileup_kernels.cu run them, on raw column files written by
// tools/pileup_hostcheck.py from the committed fixtures (tests/golden/pileup_*.npz), and writes the output columns back for
// that script to compare with the stock program's.  The shipped library only ever runs these functions inside __global__ kernels.
//
// build: g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include -x c++ tools/hostcheck/pileup_hostcheck.cpp marx_b200/csrc/calpack.cpp -o build/pileup_hostcheck
// usage: pileup_hostcheck CALPACK DIR N ALPHA FRAME_TIME SEED      (DIR holds in.<column>.bin; out.<column>.bin are written)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/marxb200.h"
#include "../../marx_b200/csrc/mx_common.cuh"
#include "../../marx_b200/csrc/mx_tables.h"
#include "../../marx_b200/csrc/mx_kernels.cuh"
#include "../../marx_b200/csrc/mx_pileup.cuh"
#include "../../marx_b200/csrc/tables_build.hpp"

using namespace mx;

struct marxb200_ctx { int dummy; };
static std::vector<unsigned char> gB3;
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
extern "C" int marxb200_set_source (marxb200_ctx *, const marxb200_source_desc *) { return 0; }
extern "C" int marxb200_set_dither (marxb200_ctx *, const marxb200_dither_desc *) { return 0; }
extern "C" int marxb200_set_hrma (marxb200_ctx *, const marxb200_hrma_desc *) { return 0; }
extern "C" int marxb200_set_flatfield (marxb200_ctx *, const marxb200_flatfield_desc *) { return 0; }
extern "C" int marxb200_set_grating (marxb200_ctx *, const marxb200_grating_desc *) { return 0; }
extern "C" int marxb200_set_hrc_s (marxb200_ctx *, const marxb200_hrc_s_desc *) { snprintf (gErr, sizeof gErr, "pile-up is an ACIS model"); return -1; }
extern "C" int marxb200_set_acis (marxb200_ctx *, const marxb200_acis_desc *d)
{
   if (d->detector_type == 0) return 0;
   HostUploader up; std::string e; int r = build_acis_blob (up, d, gB3, e); snprintf (gErr, sizeof gErr, "%s", e.c_str ()); return r;
}
extern "C" int marxb200_load_calpack_impl (marxb200_ctx *ctx, const char *path, char *errbuf, size_t errlen);

template <class T> static std::vector<T> read_col (const std::string &dir, const char *name, uint64_t n, bool optional = false)
{
   std::vector<T> v;
   FILE *fp = fopen ((dir + "/in." + name + ".bin").c_str (), "rb");
   if (fp == NULL) { if (!optional) { perror (name); exit (1); } return v; }
   v.resize (n);
   if (fread (v.data (), sizeof (T), n, fp) != n) { fprintf (stderr, "%s: short read\n", name); exit (1); }
   fclose (fp);
   return v;
}
template <class T> static void write_col (const std::string &dir, const char *name, const std::vector<T> &v, uint64_t n)
{
   FILE *fp = fopen ((dir + "/out." + name + ".bin").c_str (), "wb");
   if (fp == NULL) { perror (name); exit (1); }
   fwrite (v.data (), sizeof (T), n, fp);
   fclose (fp);
}

int main (int argc, char **argv)
{
   if (argc < 7) { fprintf (stderr, "usage: %s CALPACK DIR N ALPHA FRAME_TIME SEED\n", argv[0]); return 2; }
   marxb200_ctx ctx; char err[512];
   if (-1 == marxb200_load_calpack_impl (&ctx, argv[1], err, sizeof err)) { fprintf (stderr, "%s\n", err); return 1; }
   if (gB3.empty ()) { fprintf (stderr, "the calibration pack holds no ACIS detector\n"); return 1; }
   const std::string dir = argv[2];
   const uint64_t n = strtoull (argv[3], NULL, 10);
   static const char *dnames[6] = {"sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"};

   std::vector<int8_t> ccd = read_col<int8_t> (dir, "ccd", n);
   std::vector<float> x = read_col<float> (dir, "x", n), y = read_col<float> (dir, "y", n), t = read_col<float> (dir, "t", n),
     be = read_col<float> (dir, "benergy", n);
   std::vector<float> dith[6], odith[6];
   for (int d = 0; d < 6; d++) { dith[d] = read_col<float> (dir, dnames[d], n, true); odith[d].resize (n + 1); }

   PileupArgs a;
   memset (&a, 0, sizeof (a));
   a.ccd = ccd.data (); a.x = x.data (); a.y = y.data (); a.t = t.data (); a.benergy = be.data ();
   for (int d = 0; d < 6; d++) { a.dither[d] = dith[d].empty () ? nullptr : dith[d].data (); a.o_dither[d] = dith[d].empty () ? nullptr : odith[d].data (); }
   a.n = n; a.alpha = atof (argv[4]); a.frame_time = atof (argv[5]); a.seed = strtoull (argv[6], NULL, 10);
   for (int k = 0; k < kPuProbTable; k++) a.prob[k] = pow (a.alpha, (double) k);
   a.max_frame_events = 1u << 16;
   a.A = &((const K3Blob *) gB3.data ())->A;
   std::vector<uint32_t> frame (n + 1), key (n + 1), lo (n + 1), hi (n + 1), pn (n + 1), in (n + 1), emit (n + 1), cum (n + 1), tile_sum (n / 256 + 2);
   std::vector<float> pb (n + 1), px (n + 1), py (n + 1), ib (n + 1), sx (n + 1), sy (n + 1);
   std::vector<uint8_t> flag (n + 1); std::vector<int16_t> spha (n + 1);
   a.frame = frame.data (); a.key = key.data (); a.lo = lo.data (); a.hi = hi.data (); a.pn = pn.data (); a.in = in.data ();
   a.emit = emit.data (); a.cum = cum.data (); a.tile_sum = tile_sum.data ();
   a.pb = pb.data (); a.px = px.data (); a.py = py.data (); a.ib = ib.data (); a.sx = sx.data (); a.sy = sy.data ();
   a.flag = flag.data (); a.spha = spha.data ();
   std::vector<int8_t> o_ccd (n + 1); std::vector<float> o_x (n + 1), o_y (n + 1), o_t (n + 1), o_b (n + 1);
   std::vector<int32_t> o_frame (n + 1); std::vector<int16_t> o_np (n + 1), o_pha (n + 1);
   a.o_ccd = o_ccd.data (); a.o_x = o_x.data (); a.o_y = o_y.data (); a.o_t = o_t.data (); a.o_benergy = o_b.data ();
   a.o_frame = o_frame.data (); a.o_nphotons = o_np.data (); a.o_pha = o_pha.data ();
   a.max_out = n;
   unsigned long long n_out = 0; unsigned int error = 0;
   a.n_out = &n_out; a.error = &error;

   for (uint64_t e = 0; e < n; e++) pu_frames (a, e);
   for (uint64_t e = 0; e < n; e++) pu_store (a, e);
   for (uint64_t e = 0; e < n; e++) pu_island (a, e);
   for (uint64_t e = 0; e < n; e++) pu_detect (a, e);
   for (uint64_t e = 0; e < n; e++) pu_emit (a, e);
   // the prefix sum of pileup_kernels.cu in the layout pu_rows_through reads: inclusive inside a 256-event tile, exclusive over tiles
   {
      uint32_t total = 0;
      for (uint64_t tile = 0; tile * 256 < n; tile++)
        {
           uint32_t s = 0;
           tile_sum[tile] = total;
           for (uint64_t e = tile * 256; (e < n) && (e < (tile + 1) * 256); e++) { s += emit[e]; cum[e] = s; }
           total += s;
        }
   }
   for (uint64_t e = 0; e < n; e++) pu_scatter (a, e);
   if (error) { fprintf (stderr, "error flags %u\n", error); return 1; }

   write_col (dir, "ccd", o_ccd, n_out); write_col (dir, "x", o_x, n_out); write_col (dir, "y", o_y, n_out); write_col (dir, "t", o_t, n_out);
   write_col (dir, "benergy", o_b, n_out); write_col (dir, "frame", o_frame, n_out); write_col (dir, "nphotons", o_np, n_out);
   write_col (dir, "pha", o_pha, n_out);
   for (int d = 0; d < 6; d++) if (a.o_dither[d]) write_col (dir, dnames[d], odith[d], n_out);
   printf ("%llu events -> %llu rows\n", (unsigned long long) n, n_out);
   return 0;
}
