/* Serialises the post-init state of the reference's HRMA module (file-scope statics of
 * marx/libsrc/hrma.c) by compiling that file INTO this unit.  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <hrma.c>
#include "calpack_io.h"

int calpack_dump_hrma (mxcp_writer *w)
{
   char name[MARXB200_CALPACK_NAMELEN];
   double g[7];
   unsigned int k;

   g[0] = HRMA_Vignetting_Factor; g[1] = _Marx_HRMA_Cap_Position; g[2] = HRMA_Is_Ideal;
   g[3] = Use_Blur_Factors; g[4] = Use_Wfold_Tables; g[5] = HRMA_Use_Struts; g[6] = Use_Scale_Factors;
   CP_F64 (w, "hrma.params", g, 7);
   CP_F32 (w, "hrma.opt_energies", Energies, Num_Energies);
   CP_F32 (w, "hrma.opt_betas", Betas, Num_Energies);
   CP_F32 (w, "hrma.opt_deltas", Deltas, Num_Energies);

   for (k = 0; k < MARX_NUM_MIRRORS; k++)
     {
	HRMA_Type *h = HRMA_Mirrors + k;
	double v[62];
	unsigned int n = 0, i, j;
	v[n++] = h->mirror_number; v[n++] = h->shutter_bitmap;
	v[n++] = h->conic_a_p; v[n++] = h->conic_b_p; v[n++] = h->conic_c_p; v[n++] = h->conic_xmin_p; v[n++] = h->conic_xmax_p;
	v[n++] = h->conic_a_h; v[n++] = h->conic_b_h; v[n++] = h->conic_c_h; v[n++] = h->conic_xmin_h; v[n++] = h->conic_xmax_h;
	v[n++] = h->to_osac_p.x; v[n++] = h->to_osac_p.y; v[n++] = h->to_osac_p.z;
	v[n++] = h->to_osac_h.x; v[n++] = h->to_osac_h.y; v[n++] = h->to_osac_h.z;
	v[n++] = h->front_position; v[n++] = h->area_fraction; v[n++] = h->min_radius; v[n++] = h->max_radius;
	v[n++] = h->p_blur; v[n++] = h->h_blur; v[n++] = h->p_scat_factor; v[n++] = h->h_scat_factor;
	for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) v[n++] = h->fwd_matrix_p[i][j];
	for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) v[n++] = h->bwd_matrix_p[i][j];
	for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) v[n++] = h->fwd_matrix_h[i][j];
	for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) v[n++] = h->bwd_matrix_h[i][j];
	cp_name (name, "hrma.shell%u.params", k); CP_F64 (w, name, v, n);
	cp_name (name, "hrma.shell%u.corr_energies", k); CP_F32 (w, name, h->correction_energies, h->num_correction_factors);
	cp_name (name, "hrma.shell%u.corr_factors", k); CP_F32 (w, name, h->correction_factors, h->num_correction_factors);
	cp_name (name, "hrma.shell%u.wfold_p", k); calpack_dump_wfold (w, name, h->p_wfold_table);
	cp_name (name, "hrma.shell%u.wfold_h", k); calpack_dump_wfold (w, name, h->h_wfold_table);
     }
   return 0;
}
