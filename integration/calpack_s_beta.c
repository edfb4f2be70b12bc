/* BETA source statics (marx/libsrc/s-beta.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-beta.c>
#include "calpack_io.h"
int calpack_is_beta (void *st, double *shape)
{
   if (((Marx_Source_Type *) st)->create_photons != beta_create_photons) return 0;
   shape[0] = Core_Radius; shape[1] = 1.0 / (1.0 - Alpha); shape[2] = 0.0;
   return 1;
}
