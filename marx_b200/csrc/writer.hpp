// writer.hpp -- background column-file writer of marxb200_write_photons (marxb200_set_async_writer).
//
// Tracing 2^24 rays takes the GPU 4.4 ms; appending that batch's 89 MB of event columns to 21 files takes one host thread
// 25-45 ms even on tmpfs (page allocation + copy), so a driver loop that writes synchronously runs 10x below the trace rate.
// The writer owns a few threads; column file k always goes to thread k mod n, so that the appends of consecutive batches to
// one file keep their order while different files are written in parallel, and the call returns as soon as the batch has
// landed in a pinned host buffer (two buffers: batch k+1 is traced and copied while batch k is being written).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

struct MxWriter;
struct MxWriteTask
{
   std::string path;
   int create;                 // 1: create the file and write the 32-byte header first (marx_create_write_dump_file, marxio.c:151-205)
   unsigned char header[32];
   const unsigned char *data; size_t bytes;
   unsigned char rows_be[4];   // cumulative row count, patched at offset 20 (marx_close_write_dump_file, marxio.c:82-126)
   int buffer;                 // which of the two host buffers `data` points into
};
MxWriter *mxw_create (int n_threads);
void mxw_destroy (MxWriter *w);                                   // flushes first
void mxw_submit (MxWriter *w, int lane, const MxWriteTask &t);
void mxw_wait_buffer (MxWriter *w, int buffer);                   // until no queued task reads this buffer any more
int mxw_flush (MxWriter *w, std::string *err);                    // until everything is written; -1 + message if any write failed
int mxw_failed (MxWriter *w, std::string *err);                   // non-blocking
