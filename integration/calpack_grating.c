/* Serialises the post-init state of marx/libsrc/diffract.c.  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <diffract.c>
#include "calpack_io.h"

int calpack_dump_grating (mxcp_writer *w, int grating_module)
{
   char name[MARXB200_CALPACK_NAMELEN];
   double g[5];
   unsigned int k;

   g[0] = grating_module;
   if (Sim_Use_LETG) g[1] = g[2] = g[3] = g[4] = LEG_Rowland_Diameter;
   else { g[1] = g[2] = MEG_Rowland_Diameter; g[3] = g[4] = HEG_Rowland_Diameter; }
   CP_F64 (w, "grating.params", g, 5);
   if (grating_module == 0) return 0;

   for (k = 0; k < MARX_NUM_MIRRORS; k++)
     {
	Grating_Type *gt = Gratings[k];
	Grating_Sector_Type *gs = Grating_Sectors[k];
	double v[5];
	float *ce;
	unsigned int i, ns;
	double *sec;

	v[0] = gt->dispersion_angle; v[1] = gt->period; v[2] = gt->dp_over_p; v[3] = gt->theta_blur; v[4] = gt->vig;
	cp_name (name, "grating.shell%u.params", k); CP_F64 (w, name, v, 5);
	cp_name (name, "grating.shell%u.order_list", k); CP_I32 (w, name, gt->order_list, gt->num_orders);
	cp_name (name, "grating.shell%u.energies", k); CP_F32 (w, name, gt->energies, gt->num_energies);
	ce = (float *) malloc (sizeof (float) * gt->num_orders * gt->num_energies);
	for (i = 0; i < gt->num_orders; i++)
	  memcpy (ce + i * gt->num_energies, gt->cum_efficiencies[i], sizeof (float) * gt->num_energies);
	cp_name (name, "grating.shell%u.cum_eff", k); CP_F32 (w, name, ce, (uint64_t) gt->num_orders * gt->num_energies);
	free (ce);
	ns = (gs == NULL) ? 0 : gs->num_sectors;
	sec = (double *) malloc (sizeof (double) * (6 * ns + 1));
	for (i = 0; i < ns; i++)
	  {
	     sec[6*i+0] = gs->min_angle[i]; sec[6*i+1] = gs->max_angle[i];
	     sec[6*i+2] = gs->dtheta[i]; sec[6*i+3] = gs->dtheta_blur[i];
	     sec[6*i+4] = gs->dp_over_p[i]; sec[6*i+5] = gs->dp_over_p_blur[i];
	  }
	cp_name (name, "grating.shell%u.sectors", k); CP_F64 (w, name, sec, 6 * ns);
	free (sec);
     }
   if (Sim_Use_LETG)
     for (k = 0; k < 2; k++)
       {
	  /* support gratings are looked up on the global computed-efficiency grid (diffract.c:941-944) */
	  Grating_Type *gt = (k == 0) ? LEG_Fine_Grating : LEG_Coarse_Grating;
	  double v[5]; float *ce; unsigned int i;
	  if ((gt == NULL) || (gt->num_orders == 0)) continue;
	  v[0] = gt->dispersion_angle; v[1] = gt->period; v[2] = gt->dp_over_p; v[3] = gt->theta_blur; v[4] = gt->vig;
	  cp_name (name, "grating.support%u.params", k); CP_F64 (w, name, v, 5);
	  cp_name (name, "grating.support%u.order_list", k); CP_I32 (w, name, gt->order_list, gt->num_orders);
	  cp_name (name, "grating.support%u.energies", k); CP_F32 (w, name, Energies, Num_Energies);
	  ce = (float *) malloc (sizeof (float) * gt->num_orders * Num_Energies);
	  for (i = 0; i < gt->num_orders; i++)
	    memcpy (ce + i * Num_Energies, gt->cum_efficiencies[i], sizeof (float) * Num_Energies);
	  cp_name (name, "grating.support%u.cum_eff", k); CP_F32 (w, name, ce, (uint64_t) gt->num_orders * Num_Energies);
	  free (ce);
       }
   return 0;
}
