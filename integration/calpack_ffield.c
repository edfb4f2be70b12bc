/* Serialises the post-init state of the reference's FLATFIELD mirror module (file-scope statics of marx/libsrc/ffield.c:45-49)
 * by compiling that file INTO this unit.  Reference-side binding (integration/): compiled against the MARX tree, never into
 * libmarxb200.so. */
#include <ffield.c>
#include "calpack_io.h"

int calpack_dump_ffield (mxcp_writer *w)
{
   double g[5];
   g[0] = FF_MinY; g[1] = FF_MinZ; g[2] = FF_MaxY; g[3] = FF_MaxZ; g[4] = FF_XPos;
   CP_F64 (w, "ffield.params", g, 5);
   return 0;
}
