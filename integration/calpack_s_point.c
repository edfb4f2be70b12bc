/* POINT source (marx/libsrc/s-point.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-point.c>
#include "calpack_io.h"
int calpack_is_point (void *st) { return ((Marx_Source_Type *) st)->create_photons == point_create_photons; }
/* 0 POINT, 1 GAUSS, 2 BETA, 3 DISK, 4 LINE, 5 IMAGE, -1: not covered by the CUDA path */
int calpack_source_shape (void *st, double *shape, double *rot, double *img)
{
   shape[0] = shape[1] = shape[2] = 0.0;
   rot[0] = rot[1] = rot[2] = rot[3] = 0.0;
   img[0] = img[1] = img[2] = img[3] = 0.0;
   if (calpack_is_point (st)) return 0;
   if (calpack_is_gauss (st, shape)) return 1;
   if (calpack_is_beta (st, shape)) return 2;
   if (calpack_is_disk (st, shape)) return 3;
   if (calpack_is_line (st, shape, rot)) return 4;
   if (calpack_is_image (st, rot, img)) return 5;
   return -1;
}
