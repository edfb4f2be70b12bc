/* Shared start-up for the oracle/_ref tools: performs what marx/src/marx.c:main does before its
 * photon loop (setup_parms :809-895, setup_tstart :285-377, module_init :641-656, source :437-438),
 * through the reference's PUBLIC API only.  TEST INFRASTRUCTURE. */
#ifndef ORACLE_REF_SETUP_H
#define ORACLE_REF_SETUP_H
#include <marx.h>
typedef struct
{
   Param_File_Type *pf;
   Marx_Source_Type *source;
   int mirror_module, grating_module, detector_module;
   int random_seed;
   int num_rays, dnum_rays;
   double exposure_time;
   double tstart_years, tstart_secs;
}
Ref_Setup_Type;

/* argv is passed to pfile unchanged (e.g. "@@/path/marx.par" "GratingType=NONE" ...). */
int ref_setup (int argc, char **argv, Ref_Setup_Type *rs);
#endif
