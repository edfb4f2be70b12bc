/* helpers shared by the calpack_*.c dump units.  TEST/INTEGRATION TOOLING (reference-side binding, compiled against the MARX tree). */
#ifndef ORACLE_CALPACK_IO_H
#define ORACLE_CALPACK_IO_H
#include <stdarg.h>
#include <marxb200_calpack.h>

static inline void cp_name (char *buf, const char *fmt, ...)
{
   va_list ap;
   va_start (ap, fmt);
   vsnprintf (buf, MARXB200_CALPACK_NAMELEN, fmt, ap);
   va_end (ap);
}
#define CP_F64(w,name,ptr,n) mxcp_write ((w), (name), MXCP_F64, (ptr), (n))
#define CP_F32(w,name,ptr,n) mxcp_write ((w), (name), MXCP_F32, (ptr), (n))
#define CP_I32(w,name,ptr,n) mxcp_write ((w), (name), MXCP_I32, (ptr), (n))
#define CP_U32(w,name,ptr,n) mxcp_write ((w), (name), MXCP_U32, (ptr), (n))

int calpack_dump_source (mxcp_writer *w, void *marx_source);
int calpack_dump_dither (mxcp_writer *w);
int calpack_source_shape (void *marx_source, double *shape, double *rot, double *img);
int calpack_is_line (void *st, double *shape, double *rot);
int calpack_is_image (void *st, double *rot, double *img);
int calpack_dump_image (mxcp_writer *w);
int calpack_is_gauss (void *st, double *shape);
int calpack_is_beta (void *st, double *shape);
int calpack_is_disk (void *st, double *shape);
int calpack_is_point (void *st);
int calpack_is_rayfile (void *st);
int calpack_dump_acis_i (mxcp_writer *w, int detector_module);
int calpack_dump_hrc_s (mxcp_writer *w, int detector_module);
int calpack_dump_hrc_i (mxcp_writer *w, int detector_module);
int calpack_dump_hrma (mxcp_writer *w);
int calpack_dump_ffield (mxcp_writer *w);      /* MirrorType=FLATFIELD (ffield.c) */
int calpack_dump_wfold (mxcp_writer *w, const char *prefix, void *table);
int calpack_dump_grating (mxcp_writer *w, int grating_module);
int calpack_dump_acis_s (mxcp_writer *w, int detector_module);
int calpack_dump_fef (mxcp_writer *w, int min_ccd, int max_ccd, int *fef_map_out /* [10][1024] */);
int calpack_dump_contam (mxcp_writer *w, int ccd, const char *prefix);
#endif
