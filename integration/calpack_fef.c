/* Serialises the FEF region map of marx/libsrc/acis_fef.c (static Fef_Maps).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <acis_fef.c>
#include "calpack_io.h"

int calpack_dump_fef (mxcp_writer *w, int min_ccd, int max_ccd, int *fef_map_out)
{
   char name[MARXB200_CALPACK_NAMELEN];
   Fef_Type *seen[10 * 1024];
   unsigned int nseen = 0, i, j, k;
   int ccd;
   uint32_t nf;

   for (i = 0; i < 10 * 1024; i++) fef_map_out[i] = -1;
   for (ccd = min_ccd; ccd <= max_ccd; ccd++)
     {
	Fef_Map_Type *m = Fef_Maps[ccd];
	if (m == NULL) continue;
	for (i = 0; i < NUM_REGIONS; i++) for (j = 0; j < NUM_REGIONS; j++)
	  {
	     Fef_Type *f = m->fef_map[i][j];
	     if (f == NULL) continue;
	     for (k = 0; k < nseen; k++) if (seen[k] == f) break;
	     if (k == nseen) seen[nseen++] = f;
	     fef_map_out[ccd * 1024 + i * NUM_REGIONS + j] = (int) k;
	  }
     }
   for (k = 0; k < nseen; k++)
     {
	Fef_Type *f = seen[k];
	unsigned int ng = f->num_gaussians * f->num_energies;
	float *g = (float *) malloc (3 * ng * sizeof (float));
	uint32_t dims[2];
	for (i = 0; i < ng; i++)
	  { g[3*i] = f->gaussians[i].amp; g[3*i+1] = f->gaussians[i].center; g[3*i+2] = f->gaussians[i].sigma; }
	dims[0] = f->num_gaussians; dims[1] = f->num_energies;
	cp_name (name, "acis.fef%u.dims", k); CP_U32 (w, name, dims, 2);
	cp_name (name, "acis.fef%u.energies", k); CP_F32 (w, name, f->energies, f->num_energies);
	cp_name (name, "acis.fef%u.channels", k); CP_F32 (w, name, f->channels, f->num_energies);
	cp_name (name, "acis.fef%u.gauss", k); CP_F32 (w, name, g, 3 * ng);
	free (g);
     }
   nf = nseen;
   CP_U32 (w, "acis.num_fefs", &nf, 1);
   return 0;
}
