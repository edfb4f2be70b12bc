/* Serialises HRC-S detector state (marx/libsrc/hrc-s.c statics + detector.c globals).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <hrc-s.c>
#include "calpack_io.h"

int calpack_hrc_geom (int id, double *six);           /* calpack_hrc_geom.c: pixel-mapping constants, pixel sizes */
int calpack_hrc_pixel_sizes (double *u, double *v);
int calpack_hrc_blur (void *bt, double *thirteen);   /* calpack_hrcblur.c */
int calpack_dump_hesf (mxcp_writer *w, int *n_plates, double *cr_width);   /* calpack_drake.c */

int calpack_dump_hrc_s (mxcp_writer *w, int detector_module)
{
   char name[MARXB200_CALPACK_NAMELEN];
   double v[48];
   unsigned int n = 0, k, i;
   Marx_Detector_Geometry_Type *d;
   int n_plates = 0; double cr_width = 0.0;

   if (Use_Drake_Flat && (-1 == calpack_dump_hesf (w, &n_plates, &cr_width))) return -1;
   if (!Use_Drake_Flat)
     {
	float dummy = 0; double none = 0;
	CP_F64 (w, "hrc.hesf", &none, 0);
	CP_F32 (w, "hrc.hesf_c_energies", &dummy, 0); CP_F32 (w, "hrc.hesf_c_betas", &dummy, 0); CP_F32 (w, "hrc.hesf_c_deltas", &dummy, 0);
	CP_F32 (w, "hrc.hesf_cr_energies", &dummy, 0); CP_F32 (w, "hrc.hesf_cr_betas", &dummy, 0); CP_F32 (w, "hrc.hesf_cr_deltas", &dummy, 0);
     }
   v[n++] = detector_module; v[n++] = _MARX_NUM_HRC_S_CHIPS;
   v[n++] = _Marx_Det_XForm_Matrix.dx; v[n++] = _Marx_Det_XForm_Matrix.dy; v[n++] = _Marx_Det_XForm_Matrix.dz;
   for (i = 0; i < 9; i++) v[n++] = _Marx_Det_XForm_Matrix.matrix[i];
   v[n++] = _Marx_Det_Ideal_Flag; v[n++] = _Marx_Det_Extend_Flag;
   v[n++] = Shield_OffsetT; v[n++] = Shield_OffsetL; v[n++] = Shield_OffsetR; v[n++] = Shield_OffsetX;
   v[n++] = Shield_OffsetSL; v[n++] = Shield_OffsetSR; v[n++] = Shield_OffsetSL_Gap; v[n++] = Shield_OffsetSR_Gap;
   v[n++] = Shield_Y_Center; v[n++] = Shield_Z_Center;
   if (-1 == calpack_hrc_blur (HRC_S_Blur_Parms, v + n)) return -1;
   n += 13;
   calpack_hrc_pixel_sizes (v + n, v + n + 1); n += 2;
   v[n++] = Use_Drake_Flat; v[n++] = n_plates; v[n++] = cr_width;
   CP_F64 (w, "hrc.params", v, n);

   for (k = 0, d = HRC_S_MCPs; d != NULL; d = d->next, k++)
     {
	_Marx_HRC_QE_Type *q = MCP_QEs + Mcp_Id_Mapping[d->id];
	double gm[21];
	n = 0;
	gm[n++] = d->id;
	gm[n++] = d->x_ll.x; gm[n++] = d->x_ll.y; gm[n++] = d->x_ll.z;
	gm[n++] = d->xhat.x; gm[n++] = d->xhat.y; gm[n++] = d->xhat.z;
	gm[n++] = d->yhat.x; gm[n++] = d->yhat.y; gm[n++] = d->yhat.z;
	gm[n++] = d->normal.x; gm[n++] = d->normal.y; gm[n++] = d->normal.z;
	gm[n++] = d->xlen; gm[n++] = d->ylen;
	if (-1 == calpack_hrc_geom (d->id, gm + n)) return -1;
	n += 6;
	cp_name (name, "hrc.mcp%u.geom", k); CP_F64 (w, name, gm, n);
	cp_name (name, "hrc.mcp%u.qe_energies", k); CP_F32 (w, name, q->energies, q->num_energies);
	cp_name (name, "hrc.mcp%u.qe", k); CP_F32 (w, name, q->eff, q->num_energies);
     }
   for (k = 0; k < NUM_FILTER_REGIONS; k++)
     {
	cp_name (name, "hrc.filter%u.energies", k); CP_F32 (w, name, Filter_QEs[k].energies, Filter_QEs[k].num_energies);
	cp_name (name, "hrc.filter%u.qe", k); CP_F32 (w, name, Filter_QEs[k].eff, Filter_QEs[k].num_energies);
     }
   return 0;
}
