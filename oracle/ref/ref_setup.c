/* See ref_setup.h.  TEST INFRASTRUCTURE (oracle/_ref build only). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <jdmath.h>
#include <pfile.h>
#include "ref_setup.h"

static int Num_Rays, DNum_Rays, Random_Seed = -1;
static double Exposure_Time;
static char *Data_Directory;

static Param_Table_Type Control_Parms[] =
{
   {"NumRays", PF_INTEGER_TYPE, &Num_Rays},
   {"dNumRays", PF_INTEGER_TYPE, &DNum_Rays},
   {"RandomSeed", PF_INTEGER_TYPE, &Random_Seed},
   {"DataDirectory", PF_STRING_TYPE, &Data_Directory},
   {"ExposureTime", PF_REAL_TYPE, &Exposure_Time},
   {"Verbose", PF_INTEGER_TYPE, &Marx_Verbose},
   {"FocalLength", PF_REAL_TYPE, &Marx_Focal_Length},
   {NULL, 0, NULL}
};

int ref_setup (int argc, char **argv, Ref_Setup_Type *rs)
{
   Param_File_Type *p;
   double tstart, yrs, secs_per_year = 365.25 * 24.0 * 3600.0;

   memset (rs, 0, sizeof (*rs));
   JDMATH_INIT;

   if (NULL == (p = marx_pf_parse_cmd_line ("marx.par", "r", argc, argv)))
     return -1;
   if (-1 == pf_get_parameters (p, Control_Parms))
     return -1;
   JDMsrandom ((unsigned long) (Random_Seed == -1 ? 1 : Random_Seed));
   if (-1 == marx_set_data_directory (Data_Directory))
     return -1;

   /* marx.c:setup_tstart: TStart < 2100 means years, else seconds since 1998.0 */
   if (-1 == pf_get_double (p, "TStart", &tstart))
     return -1;
   if (tstart < 2100) { yrs = tstart; tstart = (yrs - 1998.0) * secs_per_year; }
   else yrs = 1998.0 + tstart / secs_per_year;
   if (-1 == marx_set_time (yrs, tstart))
     return -1;

   if (-1 == (rs->mirror_module = marx_mirror_init (p))) return -1;
   if (-1 == (rs->grating_module = marx_grating_init (p))) return -1;
   if (-1 == (rs->detector_module = marx_detector_init (p))) return -1;

   if ((NULL == (rs->source = marx_create_source (p)))
       || (-1 == marx_open_source (rs->source)))
     return -1;

   rs->pf = p;
   rs->random_seed = Random_Seed;
   rs->num_rays = Num_Rays;
   rs->dnum_rays = DNum_Rays;
   rs->exposure_time = Exposure_Time;
   rs->tstart_years = yrs;
   rs->tstart_secs = tstart;
   return 0;
}
