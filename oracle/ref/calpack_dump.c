/* calpack_dump: runs the reference's stock initialisation for a given marx.par (+ overrides) and
 * writes the resulting tables as a calibration pack (include/marxb200_calpack.h).  This is the
 * file-writing twin of the upload calls a MARX maintainer adds to the *_init functions
 * (INTEGRATION.md).  oracle/_ref build only; the packs it writes are committed under
 * marx_b200/caldata/.
 *
 * usage: calpack_dump OUT.calpack [pfile args: @@marx.par Name=Value ...]
 */
#include <stdio.h>
#include <stdlib.h>
#include <marx.h>
#include "ref_setup.h"
#include "calpack_io.h"

int main (int argc, char **argv)
{
   Ref_Setup_Type rs;
   mxcp_writer w;
   double meta[8];
   const char *out;

   if (argc < 2) { fprintf (stderr, "usage: %s OUT.calpack [pfile args]\n", argv[0]); return 2; }
   out = argv[1];
   argv[1] = argv[0];
   if (-1 == ref_setup (argc - 1, argv + 1, &rs)) { fprintf (stderr, "calpack_dump: setup failed\n"); return 1; }
   if (-1 == mxcp_open_write (&w, out)) { fprintf (stderr, "calpack_dump: cannot write %s\n", out); return 1; }

   meta[0] = rs.mirror_module; meta[1] = rs.grating_module; meta[2] = rs.detector_module;
   meta[3] = rs.tstart_years; meta[4] = rs.tstart_secs; meta[5] = rs.random_seed;
   meta[6] = rs.num_rays; meta[7] = rs.exposure_time;
   CP_F64 (&w, "meta", meta, 8);

   if ((0 != calpack_dump_source (&w, rs.source))      /* -1 unsupported, 1 RAYFILE (host-read photons: nothing to pack) */
       || (-1 == calpack_dump_dither (&w))
       || ((rs.mirror_module != MARX_MIRROR_HRMA) && (rs.mirror_module != MARX_MIRROR_FFIELD))
       || (-1 == ((rs.mirror_module == MARX_MIRROR_FFIELD) ? calpack_dump_ffield (&w) : calpack_dump_hrma (&w)))
       || (-1 == calpack_dump_grating (&w, rs.grating_module))
       || ((rs.grating_module != 0) && (rs.grating_module != MARX_GRATING_HETG) && (rs.grating_module != MARX_GRATING_LETG))
       || (rs.detector_module == MARX_DETECTOR_PLANE)
       || (-1 == ((rs.detector_module == MARX_DETECTOR_HRC_S) ? calpack_dump_hrc_s (&w, rs.detector_module)
                  : (rs.detector_module == MARX_DETECTOR_HRC_I) ? calpack_dump_hrc_i (&w, rs.detector_module)
                  : calpack_dump_acis_s (&w, rs.detector_module))))
     { fprintf (stderr, "calpack_dump: dump failed (only HRMA/FLATFIELD + NONE/HETG/LETG + NONE/ACIS-S/ACIS-I/HRC-S/HRC-I with POINT/GAUSS/BETA/DISK/LINE/IMAGE sources are packed)\n"); return 1; }
   mxcp_close_write (&w);
   return 0;
}
