/* HRC-S pixel-mapping constants (marx/libsrc/hrc_s_geom.c statics).  Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <hrc_s_geom.c>
#include "calpack_io.h"

/* u_start, v_start, u_0, v_0, cx_0, cy_0 of _marx_hrc_s_compute_pixel (hrc_s_geom.c:344-394) */
int calpack_hrc_geom (int id, double *s)
{
   switch (id)
     {
      case 3: s[0] = U_Active[0]; s[1] = V_Active[4]; s[2] = Right_UV[0]; s[3] = Right_UV[1]; s[4] = Right_CXCY[0]; s[5] = Right_CXCY[1]; break;
      case 2: s[0] = U_Active[0]; s[1] = V_Active[2]; s[2] = Middle_UV[0]; s[3] = Middle_UV[1]; s[4] = Middle_CXCY[0]; s[5] = Middle_CXCY[1]; break;
      case 1: s[0] = U_Active[0]; s[1] = V_Active[0]; s[2] = Left_UV[0]; s[3] = Left_UV[1]; s[4] = Left_CXCY[0]; s[5] = Left_CXCY[1]; break;
      default: return -1;
     }
   return 0;
}
int calpack_hrc_pixel_sizes (double *u, double *v) { *u = U_Pixel_Size; *v = V_Pixel_Size; return 0; }
