/* HRC blur parameters after fixup (marx/libsrc/hrcblur.c; the struct is opaque elsewhere).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <hrcblur.c>
#include "calpack_io.h"
int calpack_hrc_blur (void *p, double *b)
{
   Marx_HRC_Blur_Parm_Type *bt = (Marx_HRC_Blur_Parm_Type *) p;
   if (bt == NULL) return -1;
   b[0] = bt->gauss1_sigma; b[1] = bt->gauss1_xctr; b[2] = bt->gauss1_yctr; b[3] = bt->gauss1_wgt;
   b[4] = bt->gauss2_sigma; b[5] = bt->gauss2_xctr; b[6] = bt->gauss2_yctr; b[7] = bt->gauss2_wgt;
   b[8] = bt->lorentz1_hwhm; b[9] = bt->lorentz1_xctr; b[10] = bt->lorentz1_yctr; b[11] = bt->lorentz1_rmax; b[12] = bt->lorentz1_wgt;
   return 0;
}
