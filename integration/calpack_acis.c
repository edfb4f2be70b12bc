/* Serialises ACIS-S detector state: chip geometry + QE (marx/libsrc/acis-s.c statics), detector
 * transform (detector.c globals), frame timing (acis-i.c globals).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <acis-s.c>
#include "calpack_io.h"

extern double Frame_Time, Exposure_Time, Frame_Transfer_Time;   /* acis-i.c:54-56 */

int calpack_dump_acis_s (mxcp_writer *w, int detector_module)
{
   char name[MARXB200_CALPACK_NAMELEN];
   double v[32];
   unsigned int n = 0, k, i;
   Marx_Detector_Geometry_Type *d;
   static int fef_map[10][1024];

   v[n++] = detector_module;
   v[n++] = (detector_module == MARX_DETECTOR_ACIS_S) ? _MARX_NUM_ACIS_S_CHIPS : 0;
   v[n++] = _Marx_Det_XForm_Matrix.dx; v[n++] = _Marx_Det_XForm_Matrix.dy; v[n++] = _Marx_Det_XForm_Matrix.dz;
   for (i = 0; i < 9; i++) v[n++] = _Marx_Det_XForm_Matrix.matrix[i];
   v[n++] = _Marx_Det_Ideal_Flag; v[n++] = _Marx_Det_Extend_Flag; v[n++] = Marx_Focal_Length;
   v[n++] = Exposure_Time; v[n++] = Frame_Transfer_Time; v[n++] = Frame_Time;
   v[n++] = _Marx_Dither_Mode;
   if (detector_module == MARX_DETECTOR_ACIS_I) return calpack_dump_acis_i (w, detector_module);
   CP_F64 (w, "acis.params", v, n);
   if (detector_module != MARX_DETECTOR_ACIS_S) return 0;

   if (-1 == calpack_dump_fef (w, 4, 9, &fef_map[0][0])) return -1;

   for (k = 0, d = ACIS_S_Chips; d != NULL; d = d->next, k++)
     {
	_Marx_Acis_Chip_Type *c = &Acis_CCDS[d->id - 4];
	double gm[19];
	n = 0;
	gm[n++] = d->id;
	gm[n++] = d->x_ll.x; gm[n++] = d->x_ll.y; gm[n++] = d->x_ll.z;
	gm[n++] = d->xhat.x; gm[n++] = d->xhat.y; gm[n++] = d->xhat.z;
	gm[n++] = d->yhat.x; gm[n++] = d->yhat.y; gm[n++] = d->yhat.z;
	gm[n++] = d->normal.x; gm[n++] = d->normal.y; gm[n++] = d->normal.z;
	gm[n++] = d->xlen; gm[n++] = d->ylen; gm[n++] = d->x_pixel_size; gm[n++] = d->y_pixel_size;
	gm[n++] = d->xpixel_offset; gm[n++] = d->ypixel_offset;
	cp_name (name, "acis.chip%u.geom", k); CP_F64 (w, name, gm, n);
	cp_name (name, "acis.chip%u.qe_energies", k); CP_F32 (w, name, c->qe_energies, c->qe_num_energies);
	cp_name (name, "acis.chip%u.qe", k); CP_F32 (w, name, c->qe, c->qe_num_energies);
	cp_name (name, "acis.chip%u.filter_energies", k); CP_F32 (w, name, c->filter_energies, c->filter_num_energies);
	cp_name (name, "acis.chip%u.filter_qe", k); CP_F32 (w, name, c->filter_qe, c->filter_num_energies);
	cp_name (name, "acis.chip%u", k);
	if (-1 == calpack_dump_contam (w, d->id, name)) return -1;
	cp_name (name, "acis.chip%u.fef_map", k); CP_I32 (w, name, fef_map[d->id], 1024);
     }
   return 0;
}
