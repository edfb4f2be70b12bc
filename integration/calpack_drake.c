/* HESF "Drake flat" plates and optical constants (marx/libsrc/drake.c statics).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <drake.c>
#include "calpack_io.h"
int calpack_dump_hesf (mxcp_writer *w, int *n_plates, double *cr_width)
{
   double v[8 * 14];
   int k, n = 2 * Drake_N_Plates;
   for (k = 0; k < n; k++)
     {
	Rectangle_Type *r = Drake_Flats + k; double *q = v + 14 * k;
	q[0] = r->a.x; q[1] = r->a.y; q[2] = r->a.z; q[3] = r->e1.x; q[4] = r->e1.y; q[5] = r->e1.z;
	q[6] = r->e2.x; q[7] = r->e2.y; q[8] = r->e2.z; q[9] = r->normal.x; q[10] = r->normal.y; q[11] = r->normal.z;
	q[12] = r->len1; q[13] = r->len2;
     }
   CP_F64 (w, "hrc.hesf", v, 14 * n);
   CP_F32 (w, "hrc.hesf_c_energies", Energies_C, Betas_C ? Num_Energies_C : 0);
   CP_F32 (w, "hrc.hesf_c_betas", Betas_C, Betas_C ? Num_Energies_C : 0);
   CP_F32 (w, "hrc.hesf_c_deltas", Deltas_C, Betas_C ? Num_Energies_C : 0);
   CP_F32 (w, "hrc.hesf_cr_energies", Energies_Cr, Betas_Cr ? Num_Energies_Cr : 0);
   CP_F32 (w, "hrc.hesf_cr_betas", Betas_Cr, Betas_Cr ? Num_Energies_Cr : 0);
   CP_F32 (w, "hrc.hesf_cr_deltas", Deltas_Cr, Betas_Cr ? Num_Energies_Cr : 0);
   *n_plates = Drake_N_Plates; *cr_width = Drake_Cr_Width;
   return 0;
}
