/* DISK source statics (marx/libsrc/s-disk.c).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <s-disk.c>
#include "calpack_io.h"
int calpack_is_disk (void *st, double *shape)
{
   double t1, t0, x0;
   if (((Marx_Source_Type *) st)->create_photons != disk_create_photons) return 0;
   t1 = Disk_Theta * (1.0 / 3600.0 * PI / 180.0); t0 = Disk_Min_Theta * (1.0 / 3600.0 * PI / 180.0);
   x0 = t0 / t1; x0 = x0 * x0;
   shape[0] = t1; shape[1] = x0; shape[2] = 1.0 - x0;
   return 1;
}
