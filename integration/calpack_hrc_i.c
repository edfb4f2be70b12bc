/* Serialises HRC-I detector state (marx/libsrc/hrc-i.c statics + detector.c globals) in the layout of the HRC-S pack
 * ("hrc.params", one MCP, filter 0 = the UVIS file; no shield regions, no HESF).  Reference-side binding (integration/):
 * compiled against the MARX tree, never into libmarxb200.so. */
#include <hrc-i.c>
#include "calpack_io.h"

int calpack_hrc_blur (void *bt, double *thirteen);                 /* calpack_hrcblur.c */
int calpack_hrc_i_geom (double *ll_cxcy, double *pixel_sizes);    /* calpack_hrc_i_geom.c */

int calpack_dump_hrc_i (mxcp_writer *w, int detector_module)
{
   char name[MARXB200_CALPACK_NAMELEN];
   double v[48], ll[2], pix[2], none = 0;
   float dummy = 0;
   unsigned int n = 0, k, i;
   Marx_Detector_Geometry_Type *d;

   CP_F64 (w, "hrc.hesf", &none, 0);
   CP_F32 (w, "hrc.hesf_c_energies", &dummy, 0); CP_F32 (w, "hrc.hesf_c_betas", &dummy, 0); CP_F32 (w, "hrc.hesf_c_deltas", &dummy, 0);
   CP_F32 (w, "hrc.hesf_cr_energies", &dummy, 0); CP_F32 (w, "hrc.hesf_cr_betas", &dummy, 0); CP_F32 (w, "hrc.hesf_cr_deltas", &dummy, 0);
   calpack_hrc_i_geom (ll, pix);
   v[n++] = detector_module; v[n++] = _MARX_NUM_HRC_I_CHIPS;
   v[n++] = _Marx_Det_XForm_Matrix.dx; v[n++] = _Marx_Det_XForm_Matrix.dy; v[n++] = _Marx_Det_XForm_Matrix.dz;
   for (i = 0; i < 9; i++) v[n++] = _Marx_Det_XForm_Matrix.matrix[i];
   v[n++] = _Marx_Det_Ideal_Flag; v[n++] = _Marx_Det_Extend_Flag;
   for (i = 0; i < 10; i++) v[n++] = 0.0;                          /* shield geometry: HRC-S only */
   if (-1 == calpack_hrc_blur (HRC_I_Blur_Parms, v + n)) return -1;
   n += 13;
   v[n++] = pix[0]; v[n++] = pix[1];
   v[n++] = 0; v[n++] = 0; v[n++] = 0.0;                           /* no HESF */
   CP_F64 (w, "hrc.params", v, n);

   for (k = 0, d = HRC_I_MCP; d != NULL; d = d->next, k++)
     {
	_Marx_HRC_QE_Type *q = MCP_QEs + d->id;
	double gm[21];
	n = 0;
	gm[n++] = d->id;
	gm[n++] = d->x_ll.x; gm[n++] = d->x_ll.y; gm[n++] = d->x_ll.z;
	gm[n++] = d->xhat.x; gm[n++] = d->xhat.y; gm[n++] = d->xhat.z;
	gm[n++] = d->yhat.x; gm[n++] = d->yhat.y; gm[n++] = d->yhat.z;
	gm[n++] = d->normal.x; gm[n++] = d->normal.y; gm[n++] = d->normal.z;
	gm[n++] = d->xlen; gm[n++] = d->ylen;
	gm[n++] = ll[0]; gm[n++] = ll[1];                          /* LL_CXCY (hrc_i_geom.c:146-156) in the u_start, v_start slots */
	gm[n++] = 0; gm[n++] = 0; gm[n++] = 0; gm[n++] = 0;
	cp_name (name, "hrc.mcp%u.geom", k); CP_F64 (w, name, gm, n);
	cp_name (name, "hrc.mcp%u.qe_energies", k); CP_F32 (w, name, q->energies, q->num_energies);
	cp_name (name, "hrc.mcp%u.qe", k); CP_F32 (w, name, q->eff, q->num_energies);
     }
   for (k = 0; k < 4; k++)
     {
	unsigned int ne = (k < NUM_FILTER_REGIONS) ? Filter_QEs[k].num_energies : 0;
	cp_name (name, "hrc.filter%u.energies", k); CP_F32 (w, name, ne ? Filter_QEs[k].energies : &dummy, ne);
	cp_name (name, "hrc.filter%u.qe", k); CP_F32 (w, name, ne ? Filter_QEs[k].eff : &dummy, ne);
     }
   return 0;
}
