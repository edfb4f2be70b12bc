/* Trace-only CPU baseline: the UNMODIFIED reference stages with the reference's own RNG and its own
 * batch loop (marx/src/marx.c:545-608 without the file output): marx_create_photons ->
 * marx_mirror_reflect -> marx_grating_diffract -> marx_detect -> marx_prune_photons.
 * TEST/BENCH INFRASTRUCTURE (bench.py cpu_baseline / --impl reference).  Single-threaded like the
 * reference; bench.py runs one process per host core for the multi-core figure.
 *
 * usage: marx_trace_bench NRAYS [pfile args]     (batch size = dNumRays, seed = RandomSeed)
 * prints one JSON line: {"rays":..,"detected":..,"seconds":..,"init_seconds":..}
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <marx.h>
#include "ref_setup.h"

static double now (void)
{
   struct timespec ts;
   clock_gettime (CLOCK_MONOTONIC, &ts);
   return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main (int argc, char **argv)
{
   Ref_Setup_Type rs;
   Marx_Photon_Type *pt;
   unsigned long long nrays, done = 0, detected = 0;
   double t0, t1, t2;

   if (argc < 2) { fprintf (stderr, "usage: %s NRAYS [pfile args]\n", argv[0]); return 2; }
   nrays = strtoull (argv[1], NULL, 10);
   argv[1] = argv[0];
   t0 = now ();
   if (-1 == ref_setup (argc - 1, argv + 1, &rs)) { fprintf (stderr, "marx_trace_bench: setup failed\n"); return 1; }
   if (NULL == (pt = marx_alloc_photon_type (rs.dnum_rays))) return 1;
   t1 = now ();
   while (done < nrays)
     {
	unsigned int n, want = rs.dnum_rays;
	if (nrays - done < want) want = (unsigned int) (nrays - done);
	if (-1 == marx_create_photons (rs.source, pt, want, &n, NULL)) return 1;
	if (n == 0) break;
	if ((-1 == marx_mirror_reflect (pt, 0)) || (-1 == marx_grating_diffract (pt, 0)) || (-1 == marx_detect (pt, 0)))
	  return 1;
	marx_prune_photons (pt);
	detected += pt->num_sorted;
	done += n;
     }
   t2 = now ();
   printf ("{\"rays\": %llu, \"detected\": %llu, \"seconds\": %.6f, \"init_seconds\": %.6f}\n", done, detected, t2 - t1, t1 - t0);
   return 0;
}
