/* GAUSS source statics (marx/libsrc/s-gauss.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-gauss.c>
#include "calpack_io.h"
int calpack_is_gauss (void *st, double *shape)
{
   if (((Marx_Source_Type *) st)->create_photons != gauss_create_photons) return 0;
   shape[0] = Sigma_Theta * (1.0 / 3600.0 * PI / 180.0); shape[1] = shape[2] = 0.0;
   return 1;
}
