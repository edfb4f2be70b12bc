/* POINT source (marx/libsrc/s-point.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-point.c>
#include "calpack_io.h"
int calpack_is_point (void *st) { return ((Marx_Source_Type *) st)->create_photons == point_create_photons; }
int calpack_source_shape (void *st, double *shape)
{
   shape[0] = shape[1] = shape[2] = 0.0;
   if (calpack_is_point (st)) return 0;
   if (calpack_is_gauss (st, shape)) return 1;
   if (calpack_is_beta (st, shape)) return 2;
   if (calpack_is_disk (st, shape)) return 3;
   return -1;
}
