/* Replay oracle: drives the UNMODIFIED reference stages (marx_create_photons, marx_mirror_reflect,
 * marx_grating_diffract, marx_detect; marx/libsrc/marx.h:287-293,354-362) one photon per batch with
 * counter-based draws (philox_rng.c), and dumps the full FP64 Marx_Photon_Attr_Type (136 B,
 * marx.h:51-100) after every stage for ALL photons, dead ones included.  TEST INFRASTRUCTURE.
 *
 * usage: marx_replay OUTFILE NRAYS SEED FIRST_RAY  [pfile args: @@marx.par Name=Value ...]
 *
 * file layout (little endian):
 *   char magic[8] = "MRXRPLY1"; u64 nrays; u64 seed; u64 first_ray; u32 nstages(=4); u32 recsize(=136);
 *   i32 mirror_module, grating_module, detector_module, pad;
 *   then nrays * { nstages * record ; u32 draws[nstages] ; f64 start_time }
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <marx.h>
#include "ref_setup.h"
#include "philox_rng.h"

int main (int argc, char **argv)
{
   Ref_Setup_Type rs;
   Marx_Photon_Type *pt;
   FILE *fp;
   uint64_t nrays, seed, first, i;
   uint32_t u32[2];
   int32_t mods[4];

   if (argc < 5)
     {
	fprintf (stderr, "usage: %s OUTFILE NRAYS SEED FIRST_RAY [pfile args]\n", argv[0]);
	return 2;
     }
   nrays = strtoull (argv[2], NULL, 10);
   seed = strtoull (argv[3], NULL, 10);
   first = strtoull (argv[4], NULL, 10);

   argv[4] = argv[0];
   if (-1 == ref_setup (argc - 4, argv + 4, &rs))
     {
	fprintf (stderr, "marx_replay: setup failed\n");
	return 1;
     }
   replay_rng_seed (seed);

   if (NULL == (pt = marx_alloc_photon_type (1)))
     return 1;
   if (NULL == (fp = fopen (argv[1], "wb")))
     return 1;

   fwrite ("MRXRPLY1", 1, 8, fp);
   fwrite (&nrays, 8, 1, fp); fwrite (&seed, 8, 1, fp); fwrite (&first, 8, 1, fp);
   u32[0] = 4; u32[1] = (uint32_t) sizeof (Marx_Photon_Attr_Type);
   fwrite (u32, 4, 2, fp);
   mods[0] = rs.mirror_module; mods[1] = rs.grating_module; mods[2] = rs.detector_module; mods[3] = 0;
   fwrite (mods, 4, 4, fp);

   for (i = 0; i < nrays; i++)
     {
	unsigned int n;
	uint32_t draws[4];
	double start_time;
	uint64_t ray = first + i;

	replay_rng_set (ray, RNG_STAGE_SOURCE);
	if (-1 == marx_create_photons (rs.source, pt, 1, &n, NULL))
	  return 1;
	if (n == 0)
	  {
	     /* DitherModel=FILE: the ASPSOL file ended (dither.c:296-301); the stock driver stops here (marx.c:577-578).
	      * The header's ray count becomes the number of rays traced. */
	     nrays = i;
	     break;
	  }
	pt->attributes[0].tag = (unsigned int) ray;
	draws[0] = replay_rng_draws ();
	start_time = pt->start_time;
	fwrite (pt->attributes, sizeof (Marx_Photon_Attr_Type), 1, fp);

	replay_rng_set (ray, RNG_STAGE_MIRROR);
	if (-1 == marx_mirror_reflect (pt, 0)) return 1;
	draws[1] = replay_rng_draws ();
	fwrite (pt->attributes, sizeof (Marx_Photon_Attr_Type), 1, fp);

	replay_rng_set (ray, RNG_STAGE_GRATING);
	if (-1 == marx_grating_diffract (pt, 0)) return 1;
	draws[2] = replay_rng_draws ();
	fwrite (pt->attributes, sizeof (Marx_Photon_Attr_Type), 1, fp);

	replay_rng_set (ray, RNG_STAGE_DETECTOR);
	if (-1 == marx_detect (pt, 0)) return 1;
	draws[3] = replay_rng_draws ();
	fwrite (pt->attributes, sizeof (Marx_Photon_Attr_Type), 1, fp);

	fwrite (draws, 4, 4, fp);
	fwrite (&start_time, 8, 1, fp);
     }
   if (0 == fseek (fp, 8, SEEK_SET)) fwrite (&nrays, 8, 1, fp);
   fclose (fp);
   return 0;
}
