/* marxb200_calpack.h -- on-disk "calibration pack": the post-init tables of the MARX modules
 * (SURVEY.md 9.2 manifest) serialised as named arrays, so the CUDA path can be exercised on a box
 * that has neither the MARX data directory nor the MARX host libraries.  In a real integration the
 * MARX *_init functions call the marxb200_set_* setters directly; the pack is what
 * those same values look like when written to a file instead.
 *
 * Layout (little endian):
 *   char magic[8] = "MXB2CAL1"; uint32 n_entries; uint32 reserved;
 *   n_entries x { char name[56]; uint32 dtype; uint32 reserved; uint64 count; data[count], padded to 8 B }
 * dtype: 0 = f64, 1 = f32, 2 = i32, 3 = u32.
 *
 * Entry names and the order of packed "params" vectors are documented next to the loader
 * (marx_b200/csrc/calpack.cpp) and next to each writer (tools/calpack/).
 */
#ifndef MARXB200_CALPACK_H
#define MARXB200_CALPACK_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MARXB200_CALPACK_MAGIC "MXB2CAL1"
#define MARXB200_CALPACK_NAMELEN 56
enum { MXCP_F64 = 0, MXCP_F32 = 1, MXCP_I32 = 2, MXCP_U32 = 3 };

static inline size_t mxcp_dtype_size (uint32_t dtype) { return dtype == MXCP_F64 ? 8 : 4; }

/* ---- writer (used by the dump tools only) ---- */
typedef struct { FILE *fp; uint32_t n_entries; } mxcp_writer;

static inline int mxcp_open_write (mxcp_writer *w, const char *path)
{
   uint32_t zero[2] = {0, 0};
   w->n_entries = 0;
   if (NULL == (w->fp = fopen (path, "wb"))) return -1;
   fwrite (MARXB200_CALPACK_MAGIC, 1, 8, w->fp);
   fwrite (zero, 4, 2, w->fp);
   return 0;
}

static inline int mxcp_write (mxcp_writer *w, const char *name, uint32_t dtype, const void *data, uint64_t count)
{
   char nm[MARXB200_CALPACK_NAMELEN];
   uint32_t hdr[2];
   size_t nbytes = (size_t) count * mxcp_dtype_size (dtype);
   static const char pad[8] = {0};
   memset (nm, 0, sizeof (nm));
   strncpy (nm, name, sizeof (nm) - 1);
   hdr[0] = dtype; hdr[1] = 0;
   fwrite (nm, 1, sizeof (nm), w->fp);
   fwrite (hdr, 4, 2, w->fp);
   fwrite (&count, 8, 1, w->fp);
   if (nbytes) fwrite (data, 1, nbytes, w->fp);
   if (nbytes % 8) fwrite (pad, 1, 8 - (nbytes % 8), w->fp);
   w->n_entries++;
   return 0;
}

static inline int mxcp_close_write (mxcp_writer *w)
{
   fseek (w->fp, 8, SEEK_SET);
   fwrite (&w->n_entries, 4, 1, w->fp);
   return fclose (w->fp);
}
#endif
