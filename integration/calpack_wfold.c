/* WFOLD table accessor: the table type is opaque outside marx/libsrc/wfold.c.  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <wfold.c>
#include "calpack_io.h"

int calpack_dump_wfold (mxcp_writer *w, const char *prefix, void *table)
{
   Marx_WFold_Table_Type *t = (Marx_WFold_Table_Type *) table;
   char name[MARXB200_CALPACK_NAMELEN];
   unsigned int i, n, total = 0;
   double *hdr; uint32_t *nt; float *theta;

   n = (t == NULL) ? 0 : t->num_arrays;
   for (i = 0; i < n; i++) total += t->fold_arrays[i]->num_theta_values;
   hdr = (double *) malloc ((6 * n + 1) * sizeof (double));
   nt = (uint32_t *) malloc ((n + 1) * sizeof (uint32_t));
   theta = (float *) malloc ((total + 1) * sizeof (float));
   total = 0;
   for (i = 0; i < n; i++)
     {
	Fold_Array_Type *f = t->fold_arrays[i];
	hdr[6*i+0] = t->e_alphas[i]; hdr[6*i+1] = f->p_min; hdr[6*i+2] = f->delta_p;
	hdr[6*i+3] = f->p_max; hdr[6*i+4] = f->pow_law_norm; hdr[6*i+5] = f->pow_law_expon;
	nt[i] = f->num_theta_values;
	memcpy (theta + total, f->theta_values, f->num_theta_values * sizeof (float));
	total += f->num_theta_values;
     }
   cp_name (name, "%s.hdr", prefix); CP_F64 (w, name, hdr, 6 * n);
   cp_name (name, "%s.num_theta", prefix); CP_U32 (w, name, nt, n);
   cp_name (name, "%s.theta", prefix); CP_F32 (w, name, theta, total);
   free (hdr); free (nt); free (theta);
   return 0;
}
