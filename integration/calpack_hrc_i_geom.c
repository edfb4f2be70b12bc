/* HRC-I pixel-mapping constants (marx/libsrc/hrc_i_geom.c statics).  Reference-side binding (integration/): compiled
 * against the MARX tree, never into libmarxb200.so. */
#include <hrc_i_geom.c>
#include "calpack_io.h"
int calpack_hrc_i_geom (double *ll_cxcy, double *pixel_sizes)
{
   ll_cxcy[0] = LL_CXCY[0]; ll_cxcy[1] = LL_CXCY[1];
   pixel_sizes[0] = X_Pixel_Size; pixel_sizes[1] = Y_Pixel_Size;
   return 0;
}
