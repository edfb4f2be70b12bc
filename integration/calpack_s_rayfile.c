/* RAYFILE source (marx/libsrc/s-rayfile.c): photons are read by the stock host code and injected into the device list
 * with marxb200_upload_from; this unit only recognises the source type.  Reference-side binding (integration/):
 * compiled against the MARX tree, never into libmarxb200.so. */
#include <s-rayfile.c>
#include "calpack_io.h"
int calpack_is_rayfile (void *st) { return ((Marx_Source_Type *) st)->create_photons == rayfile_create_photons; }
