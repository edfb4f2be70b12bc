// writer_check.cpp -- host check of the background column-file writer (marx_b200/csrc/writer.cpp): many batches appended to
// several files through few threads must give the bytes a sequential writer gives, headers patched, errors reported.
// usage: writer_check DIR N_FILES N_BATCHES N_THREADS   (tests/test_writer_host.py)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../marx_b200/csrc/writer.hpp"

int main (int argc, char **argv)
{
   if (argc < 5) return 2;
   const std::string dir = argv[1];
   const int n_files = atoi (argv[2]), n_batches = atoi (argv[3]), n_threads = atoi (argv[4]);
   MxWriter *w = mxw_create (n_threads);
   std::vector<unsigned char> buf[2];
   uint32_t rows[64] = {0};
   for (int b = 0; b < n_batches; b++)
     {
        const int wb = b & 1;
        mxw_wait_buffer (w, wb);
        const size_t per = 1000 + 37 * (size_t) b;                  // rows of this batch: 4 bytes each
        buf[wb].assign (per * 4 * (size_t) n_files, 0);
        for (int f = 0; f < n_files; f++)
          for (size_t i = 0; i < per; i++)
            {
               const uint32_t v = (uint32_t) (f * 1000003u + b * 7919u + i);
               memcpy (&buf[wb][((size_t) f * per + i) * 4], &v, 4);
            }
        for (int f = 0; f < n_files; f++)
          {
             MxWriteTask t;
             char name[64]; snprintf (name, sizeof (name), "/col%02d.dat", f);
             t.path = dir + name;
             t.create = (b == 0);
             memset (t.header, 0, 32); t.header[0] = 0x83; t.header[4] = 'J';
             rows[f] += (uint32_t) per;
             t.rows_be[0] = (unsigned char) (rows[f] >> 24); t.rows_be[1] = (unsigned char) (rows[f] >> 16);
             t.rows_be[2] = (unsigned char) (rows[f] >> 8); t.rows_be[3] = (unsigned char) rows[f];
             t.data = &buf[wb][(size_t) f * per * 4]; t.bytes = per * 4; t.buffer = wb;
             mxw_submit (w, f, t);
          }
     }
   std::string err;
   if (-1 == mxw_flush (w, &err)) { fprintf (stderr, "flush: %s\n", err.c_str ()); return 1; }
   // an append to a file that does not exist must surface as an error at the next flush
   MxWriteTask bad;
   bad.path = dir + "/missing/none.dat"; bad.create = 0; memset (bad.header, 0, 32); memset (bad.rows_be, 0, 4);
   bad.data = nullptr; bad.bytes = 0; bad.buffer = 0;
   mxw_submit (w, 0, bad);
   const int rc = mxw_flush (w, &err);
   mxw_destroy (w);
   if (rc != -1) { fprintf (stderr, "the failing write was not reported\n"); return 1; }
   printf ("ok %s\n", err.c_str ());
   return 0;
}
