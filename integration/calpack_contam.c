/* Serialises marx/libsrc/aciscontam.c's per-CCD contamination model.  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <aciscontam.c>
#include "calpack_io.h"

int calpack_dump_contam (mxcp_writer *w, int ccd, const char *prefix)
{
   char name[MARXB200_CALPACK_NAMELEN];
   Single_Component_Contam_Type *c = Single_Component_Contam_Table + ccd;
   double v[15];
   unsigned int l, nb;

   memset (v, 0, sizeof (v));
   v[0] = c->num_layers;
   if (c->fxy == NULL) v[1] = 0;
   else if (c->fxy == fxy_acis0 || c->fxy == fxy_acis3) { v[1] = 1; v[2] = 1024.5; v[3] = 1024.5; }
   else if (c->fxy == fxy_acis1 || c->fxy == fxy_acis2) { v[1] = 1; v[2] = 0.5; v[3] = 1024.5; }
   else v[1] = 2;
   v[4] = c->blocking_factor;
   for (l = 0; l < c->num_layers; l++) { v[5 + l] = c->tau_0s[l]; v[10 + l] = c->tau_1s[l]; }
   cp_name (name, "%s.contam", prefix); CP_F64 (w, name, v, 15);
   nb = c->blocking_factor ? 1024 / c->blocking_factor : 0;
   for (l = 0; l < c->num_layers; l++)
     {
	cp_name (name, "%s.contam_e%u", prefix, l); CP_F32 (w, name, c->energies[l], c->num_mus[l]);
	cp_name (name, "%s.contam_mu%u", prefix, l); CP_F32 (w, name, c->mus[l], c->num_mus[l]);
	if (c->fxy == NULL)
	  { cp_name (name, "%s.contam_fxy%u", prefix, l); CP_F32 (w, name, c->fxy_vals[l], (uint64_t) nb * nb); }
     }
   return 0;
}
