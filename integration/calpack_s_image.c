/* IMAGE source statics (marx/libsrc/s-image.c: the normalised cumulative image, f32).  Reference-side binding
 * (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-image.c>
#include "calpack_io.h"
/* rot: image_create_photons' per-call rotation (s-image.c:303-319); img: nx, ny, rad per x / y pixel */
int calpack_is_image (void *stp, double *rot, double *img)
{
   Marx_Source_Type *st = (Marx_Source_Type *) stp;
   JDMVector_Type p0, normal;
   double theta;
   if (st->create_photons != image_create_photons) return 0;
   p0 = st->p;
   theta = JDMv_dot_prod (p0, JDMv_vector (-1, 0, 0));
   if (fabs (theta) > 1.0) theta = (theta < 0) ? -1.0 : 1.0;
   theta = acos (theta);
   normal = JDMv_cross_prod (JDMv_vector (-1, 0, 0), p0);
   if (theta != 0.0) JDMv_normalize (&normal);
   rot[0] = normal.x; rot[1] = normal.y; rot[2] = normal.z; rot[3] = theta;
   img[0] = X_Image_Size; img[1] = Y_Image_Size; img[2] = Rad_Per_XPixel; img[3] = Rad_Per_YPixel;
   return 1;
}
int calpack_dump_image (mxcp_writer *w)
{
   return CP_F32 (w, "source.image_cdf", Image, Image_Size);
}
