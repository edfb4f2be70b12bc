/* LINE source statics (marx/libsrc/s-line.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-line.c>
#include "calpack_io.h"
/* shape: Line_Theta (rad), cos / sin of Line_Phi (s-line.c:74-75); rot: the axis + angle taking (-1,0,0) to st->p (:77) */
int calpack_is_line (void *stp, double *shape, double *rot)
{
   Marx_Source_Type *st = (Marx_Source_Type *) stp;
   JDMVector_Type normal;
   if (st->create_photons != line_create_photons) return 0;
   shape[0] = Line_Theta; shape[1] = cos (Line_Phi); shape[2] = sin (Line_Phi);
   rot[3] = JDMv_find_rotation_axis (JDMv_vector (-1, 0, 0), st->p, &normal);
   rot[0] = normal.x; rot[1] = normal.y; rot[2] = normal.z;
   return 1;
}
