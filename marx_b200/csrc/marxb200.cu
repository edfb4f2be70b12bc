// marxb200.cu -- context, table upload and the C ABI declared in include/marxb200.h.
// Thin by design: everything per-photon happens in kernels.cu.  No CPU fallback exists: every entry
// point needs a CUDA device and fails with -1 otherwise.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/marxb200.h"
#include "../../include/marxb200_calpack.h"
#include "mx_tables.h"
#include "mx_kernels.cuh"
#include "mx_context.hpp"
#include "writer.hpp"
#include "mx_aspsol.cuh"
#include "mx_pileup.cuh"
#include "tables_build.hpp"

static_assert (sizeof (marxb200_photon_attr) == 136, "must match sizeof(Marx_Photon_Attr_Type), SURVEY.md 8a1");
static_assert (offsetof (marxb200_photon_attr, flags) == 64, "layout");
static_assert (offsetof (marxb200_photon_attr, pi) == 108, "layout");
static_assert (offsetof (marxb200_photon_attr, pulse_height) == 112, "layout");
static_assert (offsetof (marxb200_photon_attr, mirror_shell) == 116, "layout");
static_assert (offsetof (marxb200_photon_attr, ccd_num) == 120, "layout");
static_assert (offsetof (marxb200_photon_attr, tag) == 128, "layout");

using namespace mx;

static thread_local char g_err[512] = "";
int mxb_fail (const char *fmt, ...)
{
   va_list ap;
   va_start (ap, fmt);
   vsnprintf (g_err, sizeof (g_err), fmt, ap);
   va_end (ap);
   return -1;
}
#define fail mxb_fail

static cudaEvent_t prof_event (marxb200_ctx *c)
{
   cudaEvent_t e;
   if (!c->ev_pool.empty ()) { e = c->ev_pool.back (); c->ev_pool.pop_back (); return e; }
   cudaEventCreate (&e);
   return e;
}
// call before the first kernel of a sequence (begin) and after each kernel (mark)
static void prof_begin (marxb200_ctx *c)
{
   if (!c->profiling) return;
   cudaEvent_t e = prof_event (c);
   cudaEventRecord (e, c->stream);
   c->ev_marks.push_back (std::make_pair (e, -1));
}
static void prof_mark (marxb200_ctx *c, int cls)
{
   if (!c->profiling) return;
   cudaEvent_t e = prof_event (c);
   cudaEventRecord (e, c->stream);
   c->ev_marks.push_back (std::make_pair (e, cls));
}
static void prof_collect (marxb200_ctx *c)
{
   if (c->ev_marks.empty ()) return;
   cudaEventSynchronize (c->ev_marks.back ().first);
   for (size_t i = 1; i < c->ev_marks.size (); i++)
     {
        const int cls = c->ev_marks[i].second;
        if (cls < 0) continue;
        float ms = 0.f;
        cudaEventElapsedTime (&ms, c->ev_marks[i - 1].first, c->ev_marks[i].first);
        c->prof_ms[cls] += ms; c->prof_n[cls] += 1;
     }
   for (auto &m : c->ev_marks) c->ev_pool.push_back (m.first);
   c->ev_marks.clear ();
}

// ---------------------------------------------------------------------------------------------
static int dev_upload (marxb200_ctx *c, const void *host, size_t bytes, void **out)
{
   void *d = nullptr;
   if (bytes == 0) bytes = 16;
   size_t padded = (bytes + 255) & ~(size_t) 255;
   CUDA_OK (cudaMalloc (&d, padded));
   CUDA_OK (cudaMemsetAsync (d, 0, padded, c->stream));
   if (host != nullptr) CUDA_OK (cudaMemcpyAsync (d, host, bytes, cudaMemcpyHostToDevice, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->allocs.push_back (std::make_pair (d, c->alloc_tag));
   *out = d;
   return 0;
}
template <class T> static int dev_upload_t (marxb200_ctx *c, const T *host, size_t n, const T **out)
{
   void *d;
   if (-1 == dev_upload (c, host, n * sizeof (T), &d)) return -1;
   *out = (const T *) d;
   return 0;
}
// a module's setter runs again: release the tables of its previous call (the stream is drained first)
static void begin_module (marxb200_ctx *c, int tag)
{
   bool any = false;
   for (auto &p : c->allocs) if (p.second == tag) any = true;
   if (any)
     {
        cudaStreamSynchronize (c->stream);
        size_t k = 0;
        for (size_t i = 0; i < c->allocs.size (); i++)
          {
             if (c->allocs[i].second == tag) cudaFree (c->allocs[i].first);
             else c->allocs[k++] = c->allocs[i];
          }
        c->allocs.resize (k);
     }
   c->alloc_tag = tag;
}
static size_t align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

static int ensure_pinned (marxb200_ctx *c, size_t bytes)
{
   if (c->h_pinned_bytes >= bytes) return 0;
   bytes += bytes / 4 + 65536;          // event counts fluctuate from batch to batch: grow with headroom, not every time
   if (c->h_pinned) cudaFreeHost (c->h_pinned);
   c->h_pinned = nullptr; c->h_pinned_bytes = 0;
   CUDA_OK (cudaMallocHost (&c->h_pinned, bytes));
   c->h_pinned_bytes = bytes;
   return 0;
}
static int ensure_aos (marxb200_ctx *c, uint64_t n)
{
   if (c->d_aos_cap >= n) return 0;
   n += n / 4 + 1024;
   if (c->d_aos) cudaFree (c->d_aos);
   c->d_aos = nullptr; c->d_aos_cap = 0;
   CUDA_OK (cudaMalloc (&c->d_aos, (size_t) n * sizeof (marxb200_photon_attr)));
   c->d_aos_cap = n;
   return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_abi_version (void) { return MARXB200_ABI_VERSION; }
extern "C" const char *marxb200_last_error (void) { return g_err; }

extern "C" int marxb200_device_warmup (int device_ordinal)
{
   cudaError_t e = cudaSetDevice (device_ordinal);
   if (e == cudaSuccess) e = cudaFree (nullptr);
   if (e != cudaSuccess) return fail ("marxb200_device_warmup: device %d: %s", device_ordinal, cudaGetErrorString (e));
   return 0;
}

#define GUARD(idx) do { if (-1 == mxb_guard_buffer (c, (idx))) return -1; } while (0)
extern "C" int marxb200_create (marxb200_ctx **ctxp, int device_ordinal, uint64_t seed)
{
   if (ctxp == nullptr) return fail ("marxb200_create: NULL ctxp");
   *ctxp = nullptr;
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount (&ndev);
   if ((e != cudaSuccess) || (ndev == 0))
     return fail ("marxb200_create: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString (e));
   if ((device_ordinal < 0) || (device_ordinal >= ndev)) return fail ("marxb200_create: bad device ordinal %d", device_ordinal);
   CUDA_OK (cudaSetDevice (device_ordinal));
   marxb200_ctx *c = new marxb200_ctx ();
   struct Guard { marxb200_ctx *c; ~Guard () { if (c) marxb200_destroy (c); } } guard{c};     // a failing CUDA call below frees what exists
   c->device = device_ordinal;
   c->seed = seed;
   if (const char *e = getenv ("MARXB200_K3_SPLIT")) c->k3_split = atoi (e);      // developer A/B switch
   if (const char *e = getenv ("MARXB200_K01_TICKET")) c->k01_ticket = atoi (e);  // developer A/B switch
   if (const char *e = getenv ("MARXB200_K1_SPLIT")) c->k1_split = atoi (e);      // developer A/B switch
   if (const char *e = getenv ("MARXB200_K2_SPLIT")) c->k2_split = atoi (e);      // developer A/B switch
   cudaDeviceProp prop;
   CUDA_OK (cudaGetDeviceProperties (&prop, device_ordinal));
   c->num_sms = prop.multiProcessorCount;
   CUDA_OK (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
   c->own_stream = true;
   CUDA_OK (cudaMalloc (&c->d_counts, (marxb200_ctx::kNumCounts + marxb200_ctx::kNumTickets) * sizeof (unsigned long long)));
   c->d_ticket = c->d_counts + marxb200_ctx::kNumCounts;
   CUDA_OK (cudaStreamCreateWithFlags (&c->ahead_stream, cudaStreamNonBlocking));
   CUDA_OK (cudaEventCreateWithFlags (&c->ev_ahead, cudaEventDisableTiming));
   CUDA_OK (cudaEventCreateWithFlags (&c->ev_ahead_consumed, cudaEventDisableTiming));
   for (int k = 0; k < 2; k++)
     {
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_reader_go[k], cudaEventDisableTiming));
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_reader_done[k], cudaEventDisableTiming));
     }
   CUDA_OK (cudaMalloc (&c->d_snap, 2 * sizeof (unsigned long long)));
   if (const char *e = getenv ("MARXB200_LOOKAHEAD")) c->ahead_on = (atoi (e) != 0);          // developer A/B switch
   CUDA_OK (cudaMalloc (&c->d_times, 2 * sizeof (double)));
   CUDA_OK (cudaMemset (c->d_counts, 0, (marxb200_ctx::kNumCounts + marxb200_ctx::kNumTickets) * sizeof (unsigned long long)));
   CUDA_OK (cudaMemset (c->d_times, 0, 2 * sizeof (double)));
   memset (&c->S, 0, sizeof (c->S));
   memset (&c->D, 0, sizeof (c->D));
   guard.c = nullptr;
   *ctxp = c;
   return 0;
}

extern "C" int marxb200_destroy (marxb200_ctx *c)
{
   if (c == nullptr) return -1;
   cudaSetDevice (c->device);
   cudaDeviceSynchronize ();
   if (getenv ("MARXB200_VERBOSE") && (c->ahead_hits + c->ahead_misses > 0))
     fprintf (stderr, "marxb200: time pre-pass look-ahead: %llu batches found their sums ready, %llu ran the pre-pass themselves\n",
              (unsigned long long) c->ahead_hits, (unsigned long long) c->ahead_misses);
   if (c->writer) { mxw_destroy (c->writer); c->writer = nullptr; }       // waits for the queued column writes
   for (int i = 0; i < 2; i++) if (c->h_wbuf[i]) cudaFreeHost (c->h_wbuf[i]);
   mxb_comm_release (c);
   prof_collect (c);
   for (cudaEvent_t e : c->ev_pool) cudaEventDestroy (e);
   for (auto &p : c->allocs) cudaFree (p.first);
   if (c->d_upload_ids) cudaFree (c->d_upload_ids);
   for (int i = 0; i < 2; i++) if (c->slab[i]) cudaFree (c->slab[i]);
   if (c->rc_slab) cudaFree (c->rc_slab);
   cudaFree (c->d_counts); cudaFree (c->d_times);
   if (c->ahead_stream) { cudaStreamSynchronize (c->ahead_stream); cudaStreamDestroy (c->ahead_stream); }
   if (c->ev_ahead) cudaEventDestroy (c->ev_ahead);
   if (c->ev_ahead_consumed) cudaEventDestroy (c->ev_ahead_consumed);
   for (int k = 0; k < 2; k++) { if (c->ev_reader_go[k]) cudaEventDestroy (c->ev_reader_go[k]); if (c->ev_reader_done[k]) cudaEventDestroy (c->ev_reader_done[k]); }
   if (c->d_snap) cudaFree (c->d_snap);
   if (c->ahead_tile_sums) cudaFree (c->ahead_tile_sums);
   if (c->ahead_super_sums) cudaFree (c->ahead_super_sums);
   if (c->d_bitmap) cudaFree (c->d_bitmap);
   if (c->d_word_prefix) cudaFree (c->d_word_prefix);
   if (c->d_block_prefix) cudaFree (c->d_block_prefix);
   if (c->d_perm) cudaFree (c->d_perm);
   if (c->d_tile_sums) cudaFree (c->d_tile_sums);
   if (c->d_tile_base) cudaFree (c->d_tile_base);
   if (c->d_super_sums) cudaFree (c->d_super_sums);
   if (c->d_aos) cudaFree (c->d_aos);
   if (c->egress_slab) cudaFree (c->egress_slab);
   if (c->packed_slab_b) cudaFree (c->packed_slab_b);
   for (int k = 0; k < 2; k++) if (c->ev_slab_free[k]) cudaEventDestroy (c->ev_slab_free[k]);
   if (c->h_egress_count) cudaFreeHost (c->h_egress_count);
   if (c->ev_staged) cudaEventDestroy (c->ev_staged);
   if (c->ev_copied) cudaEventDestroy (c->ev_copied);
   if (c->copy_stream) cudaStreamDestroy (c->copy_stream);
   if (c->h_pinned) cudaFreeHost (c->h_pinned);
   if (c->l1_slab) cudaFree (c->l1_slab);
   if (c->d_l1_state) cudaFree (c->d_l1_state);
   if (c->own_stream && c->stream) cudaStreamDestroy (c->stream);
   delete c;
   return 0;
}

extern "C" int marxb200_set_stream (marxb200_ctx *c, void *cuda_stream)
{
   if (c == nullptr) return fail ("NULL ctx");
   GUARD (-1);
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (cuda_stream == nullptr)
     {
        if (!c->own_stream)
          {
             CUDA_OK (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
             c->own_stream = true;
          }
        return 0;
     }
   if (c->own_stream) cudaStreamDestroy (c->stream);
   c->own_stream = false;
   c->stream = (cudaStream_t) cuda_stream;
   return 0;
}

extern "C" int marxb200_set_compaction (marxb200_ctx *c, int on)
{
   if (c == nullptr) return fail ("NULL ctx");
   c->compact = on ? 1 : 0;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// table setters
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_set_source (marxb200_ctx *c, const marxb200_source_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_source: NULL argument");
   if ((d->source_type < 0) || (d->source_type > 5)) return fail ("marxb200_set_source: source type %d is not implemented (POINT, GAUSS, BETA, DISK, LINE, IMAGE are)", d->source_type);
   if ((d->spectrum_type != 1) && (d->spectrum_type != 2)) return fail ("marxb200_set_source: unknown spectrum type %d", d->spectrum_type);
   CUDA_OK (cudaSetDevice (c->device));
   if (c->comm != nullptr) CUDA_OK (cudaDeviceSynchronize ());       // a look-ahead pre-pass (comm.cu) may still read the old tables
   c->source_epoch++;
   begin_module (c, marxb200_ctx::TAG_SOURCE);
   SourceDev &S = c->S;
   memset (&S, 0, sizeof (S));
   S.source_type = d->source_type; S.spectrum_type = d->spectrum_type;
   for (int i = 0; i < 3; i++) { S.p[i] = d->p[i]; S.p_normal[i] = d->p_normal[i]; }
   S.distance = d->distance; S.emin = d->emin; S.emax = d->emax;
   for (int i = 0; i < 3; i++) S.shape[i] = d->shape[i];
   if (d->spectrum_type == 2)
     {
        if ((d->spec_num < 2) || !d->spec_energies || !d->spec_cum_flux) return fail ("marxb200_set_source: FILE spectrum needs a table");
        if (-1 == dev_upload_t (c, d->spec_energies, d->spec_num, &S.spec_energies)) return -1;
        if (-1 == dev_upload_t (c, d->spec_cum_flux, d->spec_num, &S.spec_cum_flux)) return -1;
        S.spec_num = d->spec_num;
     }
   for (int i = 0; i < 3; i++) S.rot_axis[i] = d->rot_axis[i];
   S.rot_angle = d->rot_angle;
   if (d->source_type == 5)
     {
        const uint64_t npix = (uint64_t) d->image_nx * d->image_ny;
        if ((npix == 0) || (npix > 0xFFFFFFFFull) || !d->image_cdf) return fail ("marxb200_set_source: IMAGE source needs a cumulative image");
        if (-1 == dev_upload_t (c, d->image_cdf, (size_t) npix, &S.image_cdf)) return -1;
        S.image_size = (uint32_t) npix; S.image_nx = d->image_nx; S.image_ny = d->image_ny;
        S.rad_per_xpixel = d->rad_per_xpixel; S.rad_per_ypixel = d->rad_per_ypixel;
     }
   // compute_mean_time, source.c:260-264
   S.mean_time = (d->total_flux <= 0.0) ? 0.0 : 1.0 / d->total_flux / d->geometric_area;
   c->source_distance = d->distance;
   c->have_source = true;
   return 0;
}

extern "C" int marxb200_set_dither (marxb200_ctx *c, const marxb200_dither_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_dither: NULL argument");
   if ((d->mode < 0) || (d->mode > 2)) return fail ("marxb200_set_dither: unknown dither model %d", d->mode);
   DitherDev &D = c->D;
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_DITHER);
   if ((D.mode == 2) && (d->mode != 2)) c->det_dither_dirty = true;   // the per-ray detector-dither columns hold an ASPSOL run's values
   D.aspsol = nullptr; D.num_aspsol = 0;
   if (d->mode == 2)
     {
        if ((d->num_aspsol < 2) || (d->aspsol == nullptr)) return fail ("marxb200_set_dither: the ASPSOL model needs at least two states");
        CUDA_OK (cudaSetDevice (c->device));
        if (-1 == dev_upload_t (c, d->aspsol, 7 * (size_t) d->num_aspsol, &D.aspsol)) return -1;
        D.num_aspsol = d->num_aspsol;
        c->aspsol_t_last = d->aspsol[7 * (size_t) (d->num_aspsol - 1)];
     }
   D.mode = d->mode;
   D.ra_amp = d->ra_amp; D.dec_amp = d->dec_amp; D.roll_amp = d->roll_amp;
   D.ra_period = d->ra_period; D.dec_period = d->dec_period; D.roll_period = d->roll_period;
   D.ra_phase = d->ra_phase; D.dec_phase = d->dec_phase; D.roll_phase = d->roll_phase;
   D.nominal_roll = d->nominal_roll; D.aspect_blur = d->aspect_blur;
   c->have_dither = true;
   return 0;
}

struct CudaUploader
{
   marxb200_ctx *c;
   std::string err;
   const void *operator() (const void *host, size_t bytes)
   {
      void *d = nullptr;
      if (-1 == dev_upload (c, host, bytes, &d)) { err = g_err; return nullptr; }
      return d;
   }
};

extern "C" int marxb200_set_hrma (marxb200_ctx *c, const marxb200_hrma_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_hrma: NULL argument");
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_HRMA);
   CudaUploader up{c};
   std::vector<unsigned char> blob;
   std::string err;
   if (-1 == mx::build_hrma_blob (up, d, blob, err)) return fail ("marxb200_set_hrma: %s", err.c_str ());
   if (-1 == dev_upload (c, blob.data (), blob.size (), &c->blob1)) return -1;
   c->blob1_bytes = (uint32_t) blob.size ();
   {
      // what each HRMA phase stages into shared memory (mx_kernels.cuh K1Blob): A the header; B everything up to and
      // including the P-conic WFOLD search keys; C the header plus the H-conic keys as a second segment
      const K1Blob *B = reinterpret_cast<const K1Blob *> (blob.data ());
      c->k1b_bytes = B->off_wkeys_p + B->wkeys_bytes;
      c->k1c_seg2_off = B->off_wkeys_h; c->k1c_seg2_bytes = B->wkeys_bytes;
      // the other cut (B1 | B2+C1 | C2): B1 the header and the tables, B2+C1 the header plus the P-conic keys, C2 like C
      c->k1b1_bytes = B->off_wkeys_p; c->k1b2_seg2_off = B->off_wkeys_p;
   }
   c->grid1[0] = stage_grid_size (10, c->num_sms, c->blob1_bytes);
   c->grid1[1] = stage_grid_size (11, c->num_sms, c->k1b_bytes);
   c->grid1[2] = stage_grid_size (12, c->num_sms, c->blob1_bytes, c->k1c_seg2_bytes);
   c->grid1[3] = stage_grid_size (14, c->num_sms, c->k1b1_bytes);
   c->grid1[4] = stage_grid_size (15, c->num_sms, c->blob1_bytes, c->k1c_seg2_bytes);
   c->grid1[5] = stage_grid_size (16, c->num_sms, c->blob1_bytes, c->k1c_seg2_bytes);
   if (getenv ("MARXB200_VERBOSE"))
     fprintf (stderr, "marxb200: HRMA grids A %d B %d C %d | B1 %d B2C1 %d C2 %d (x %d threads)\n", c->grid1[0], c->grid1[1], c->grid1[2],
              c->grid1[3], c->grid1[4], c->grid1[5], kStageThreads);
   c->grid01 = fused_source_grid (c->num_sms);
   c->have_hrma = true;
   c->mirror_is_flat = false;
   return 0;
}

extern "C" int marxb200_set_flatfield (marxb200_ctx *c, const marxb200_flatfield_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_flatfield: NULL argument");
   if ((d->max_y <= d->min_y) || (d->max_z <= d->min_z) || (d->x_pos < 0.0)) return fail ("marxb200_set_flatfield: FF_* parameters not physical");   // ffield.c:115-121
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_HRMA);
   c->ff[0] = d->min_y; c->ff[1] = d->min_z; c->ff[2] = d->max_y; c->ff[3] = d->max_z; c->ff[4] = d->x_pos;
   c->mirror_is_flat = true;
   c->have_hrma = true;
   return 0;
}

extern "C" int marxb200_set_grating (marxb200_ctx *c, const marxb200_grating_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_grating: NULL argument");
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_GRATING);
   c->grating_type = d->type;
   c->have_grating = true;
   if (d->type == 0) return 0;
   CudaUploader up{c};
   std::vector<unsigned char> blob;
   std::string err;
   if (-1 == mx::build_grating_blob (up, d, blob, err)) { c->have_grating = false; return fail ("marxb200_set_grating: %s", err.c_str ()); }
   if (-1 == dev_upload (c, blob.data (), blob.size (), &c->blob2)) return -1;
   c->blob2_bytes = (uint32_t) blob.size ();
   c->grid2 = stage_grid_size (2, c->num_sms, c->blob2_bytes);
   return 0;
}

extern "C" int marxb200_set_acis (marxb200_ctx *c, const marxb200_acis_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_acis: NULL argument");
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_DETECTOR);
   c->detector_type = d->detector_type;
   c->have_acis = true;
   if (d->detector_type == 0) return 0;
   CudaUploader up{c};
   std::vector<unsigned char> blob;
   std::string err;
   if (-1 == mx::build_acis_blob (up, d, blob, err)) { c->have_acis = false; return fail ("marxb200_set_acis: %s", err.c_str ()); }
   if (-1 == dev_upload (c, blob.data (), blob.size (), &c->blob3)) return -1;
   c->blob3_bytes = (uint32_t) blob.size ();
   c->grid3 = stage_grid_size (3, c->num_sms, c->blob3_bytes);
   c->detector_is_hrc = false;
   return 0;
}

extern "C" int marxb200_set_hrc_s (marxb200_ctx *c, const marxb200_hrc_s_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_hrc_s: NULL argument");
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_DETECTOR);
   c->detector_type = d->detector_type;
   c->have_acis = true;
   if (d->detector_type == 0) return 0;
   CudaUploader up{c};
   std::vector<unsigned char> blob;
   std::string err;
   if (-1 == mx::build_hrc_blob (up, d, blob, err)) { c->have_acis = false; return fail ("marxb200_set_hrc_s: %s", err.c_str ()); }
   if (-1 == dev_upload (c, blob.data (), blob.size (), &c->blob3)) return -1;
   c->blob3_bytes = (uint32_t) blob.size ();
   c->grid3 = stage_grid_size (4, c->num_sms, c->blob3_bytes);
   c->detector_is_hrc = true;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// photon buffers
// ---------------------------------------------------------------------------------------------
static size_t carve (PhotonSoA &b, unsigned char *base, uint64_t n)
{
   size_t off = 0;
   auto take = [&] (size_t elem) { size_t o = off; off = (off + elem * n + 255) & ~(size_t) 255; return base ? base + o : (unsigned char *) nullptr; };
   b.energy = (double *) take (8);
   b.x0 = (double *) take (8); b.x1 = (double *) take (8); b.x2 = (double *) take (8);
   b.p0 = (double *) take (8); b.p1 = (double *) take (8); b.p2 = (double *) take (8);
   b.time = (double *) take (8);
   b.aux = (double *) take (8);
   b.ray = (uint64_t *) take (8);
   b.slot = (uint32_t *) take (4);
   b.flags = (uint32_t *) take (4);
   b.dra = (float *) take (4); b.ddec = (float *) take (4); b.droll = (float *) take (4);
   b.ddy = (float *) take (4); b.ddz = (float *) take (4); b.ddth = (float *) take (4);
   b.chipx = (float *) take (4); b.chipy = (float *) take (4); b.pi = (float *) take (4);
   b.upix = (float *) take (4); b.vpix = (float *) take (4);
   b.sorders = (uint32_t *) take (4);
   b.pha = (int16_t *) take (2);
   b.shell = (uint8_t *) take (1);
   b.order = (int8_t *) take (1); b.ccd = (int8_t *) take (1); b.region = (int8_t *) take (1);
   return off;
}

extern "C" int marxb200_alloc_photons (marxb200_ctx *c, uint64_t max_photons)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (max_photons == 0) return fail ("marxb200_alloc_photons: max_photons must be > 0");
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (c->ahead_stream) CUDA_OK (cudaStreamSynchronize (c->ahead_stream));
   GUARD (-1);
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->ahead_valid = false;
   if (c->ahead_tile_sums) { cudaFree (c->ahead_tile_sums); c->ahead_tile_sums = nullptr; }
   if (c->ahead_super_sums) { cudaFree (c->ahead_super_sums); c->ahead_super_sums = nullptr; }
   for (int i = 0; i < 2; i++) if (c->slab[i]) { cudaFree (c->slab[i]); c->slab[i] = nullptr; }
   if (c->d_bitmap) cudaFree (c->d_bitmap);
   if (c->d_word_prefix) cudaFree (c->d_word_prefix);
   if (c->d_block_prefix) cudaFree (c->d_block_prefix);
   if (c->d_perm) cudaFree (c->d_perm);
   if (c->d_tile_sums) cudaFree (c->d_tile_sums);
   if (c->d_tile_base) cudaFree (c->d_tile_base);
   if (c->d_super_sums) cudaFree (c->d_super_sums);
   PhotonSoA tmp;
   size_t bytes = carve (tmp, nullptr, max_photons);
   for (int i = 0; i < 2; i++)
     {
        CUDA_OK (cudaMalloc (&c->slab[i], bytes));
        CUDA_OK (cudaMemsetAsync (c->slab[i], 0, bytes, c->stream));
        carve (c->buf[i], (unsigned char *) c->slab[i], max_photons);
     }
   {
      // per-ray constants: energy, time (f64), ray (u64), dither ra/dec/roll (f32), each column 256-byte aligned
      if (c->rc_slab) { cudaFree (c->rc_slab); c->rc_slab = nullptr; }
      const size_t c8 = (8 * (size_t) max_photons + 255) & ~(size_t) 255, c4 = (4 * (size_t) max_photons + 255) & ~(size_t) 255;
      CUDA_OK (cudaMalloc (&c->rc_slab, 3 * c8 + 6 * c4));
      CUDA_OK (cudaMemsetAsync (c->rc_slab, 0, 3 * c8 + 6 * c4, c->stream));
      unsigned char *b = (unsigned char *) c->rc_slab;
      c->rc.er = (double2 *) b; c->rc.time = (double *) (b + 2 * c8);          // {energy, ray id} records: 16 B per slot
      c->rc.dra = (float *) (b + 3 * c8); c->rc.ddec = (float *) (b + 3 * c8 + c4); c->rc.droll = (float *) (b + 3 * c8 + 2 * c4);
      c->rc.ddy = (float *) (b + 3 * c8 + 3 * c4); c->rc.ddz = (float *) (b + 3 * c8 + 4 * c4); c->rc.ddth = (float *) (b + 3 * c8 + 5 * c4);
   }
   uint64_t n_tiles = (max_photons + kTile - 1) / kTile + 1;
   uint64_t n_super = (n_tiles + kSuperTile - 1) / kSuperTile + 1;
   c->n_words = max_photons / 32 + 1;
   CUDA_OK (cudaMalloc (&c->d_bitmap, c->n_words * sizeof (uint32_t)));
   CUDA_OK (cudaMalloc (&c->d_word_prefix, c->n_words * sizeof (uint32_t)));
   CUDA_OK (cudaMalloc (&c->d_block_prefix, (c->n_words / 1024 + 2) * sizeof (uint32_t)));
   CUDA_OK (cudaMalloc (&c->d_perm, (max_photons + 1) * sizeof (uint32_t)));
   CUDA_OK (cudaMalloc (&c->d_tile_sums, n_tiles * sizeof (double)));
   CUDA_OK (cudaMalloc (&c->d_tile_base, n_tiles * sizeof (double)));
   CUDA_OK (cudaMalloc (&c->d_super_sums, n_super * sizeof (double)));
   CUDA_OK (cudaMalloc (&c->ahead_tile_sums, n_tiles * sizeof (double)));
   CUDA_OK (cudaMalloc (&c->ahead_super_sums, n_super * sizeof (double)));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->capacity = max_photons;
   c->cur = 0; c->stage_done = -1; c->n_generated = 0; c->ordered = true; c->prepack.done = false;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// off-stream readers of the finished list (mx_context.hpp)
// ---------------------------------------------------------------------------------------------
int mxb_guard_buffer (marxb200_ctx *c, int idx)
{
   for (int k = 0; k < 2; k++)
     if ((c->reader_buf[k] >= 0) && ((idx < 0) || (c->reader_buf[k] == idx)))
       {
          CUDA_OK (cudaStreamWaitEvent (c->stream, c->ev_reader_done[k], 0));
          c->reader_buf[k] = -1;
       }
   return 0;
}
bool mxb_prepack_matches (const marxb200_ctx *c, int owner, const mx::PackArgs &want)
{
   const marxb200_ctx::Prepack &p = c->prepack;
   return p.armed && p.done && (p.owner == owner) && (p.args.dst == want.dst) && (p.args.dev_start_time == want.dev_start_time)
          && (p.args.total_time == want.total_time) && (p.args.max_rows == want.max_rows)
          && (0 == memcmp (&p.args.plan, &want.plan, sizeof (want.plan)));
}
void mxb_prepack_arm (marxb200_ctx *c, int owner, const mx::PackArgs &next, cudaEvent_t ev_free)
{
   c->prepack.armed = (getenv ("MARXB200_PREPACK") == nullptr) || (atoi (getenv ("MARXB200_PREPACK")) != 0);      // developer A/B switch
   c->prepack.done = false; c->prepack.owner = owner; c->prepack.args = next; c->prepack.ev_free = ev_free;
}
int mxb_reader_begin (marxb200_ctx *c, int slot, cudaStream_t reader)
{
   CUDA_OK (cudaEventRecord (c->ev_reader_go[slot], c->stream));
   CUDA_OK (cudaStreamWaitEvent (reader, c->ev_reader_go[slot], 0));
   return 0;
}
int mxb_reader_end (marxb200_ctx *c, int slot, cudaStream_t reader)
{
   CUDA_OK (cudaEventRecord (c->ev_reader_done[slot], reader));
   c->reader_buf[slot] = c->cur;
   return 0;
}

static void fill_source_args (marxb200_ctx *c, SourceArgs &a, uint64_t first_ray, uint64_t n, double time_base)
{
   a.out = c->buf[0];
   a.rc = c->rc;
   a.first_ray = first_ray; a.n = n; a.seed = c->seed;
   a.S = c->S; a.D = c->D;
   a.time_base = (time_base >= 0.0) ? time_base : 0.0;
   a.use_dev_base = (time_base >= 0.0) ? 0 : 1;
   a.dev_times = c->d_times;
   a.tile_sums = c->d_tile_sums; a.tile_base = c->d_tile_base; a.supertile_sums = c->d_super_sums;
   a.n_out = c->d_counts + 0;
}

extern "C" int marxb200_create_photons (marxb200_ctx *c, uint64_t first_ray, uint64_t n, double time_base_in)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (!c->have_source) return fail ("marxb200_create_photons: no source set");
   if (n > c->capacity) return fail ("marxb200_create_photons: n=%llu exceeds the allocated capacity %llu", (unsigned long long) n, (unsigned long long) c->capacity);
   CUDA_OK (cudaSetDevice (c->device));
   SourceArgs a;
   // time_base_in < 0: continue the running sum from the device scalar (no host round trip)
   fill_source_args (c, a, first_ray, n, time_base_in);
   if (c->det_dither_dirty)
     {
        // an upload left detector-dither values in the per-ray columns: clear all slots once (the fused path never writes them)
        const size_t c4 = (4 * (size_t) c->capacity + 255) & ~(size_t) 255;
        CUDA_OK (cudaMemsetAsync (c->rc.ddy, 0, 3 * c4, c->stream));
     }
   GUARD (0);
   prof_begin (c);
   launch_time_sums (a, c->stream); prof_mark (c, 0);
   launch_time_scan (a, c->stream); prof_mark (c, 1);
   launch_source (a, c->stream); prof_mark (c, 2);
   c->launches += 5;                     // k0_time_sums, k0_time_super/_bases/_tiles, k0_source
   c->det_dither_dirty = false;          // k0_source rewrote the detector-dither columns of every slot it generated
   c->cur = 0; c->stage_done = 0; c->n_generated = n; c->ordered = true; c->first_mirror_kernel = 0; c->prepack.done = false;
   if (c->D.mode == 2)
     {
        // the stock reader ends the simulation at the first ray it cannot bracket (end of the ASPSOL file)
        launch_exposure_truncate (c->buf[0], c->d_counts + 0, c->d_times, c->aspsol_t_last, 0, c->stream);
        c->launches += 1;
        unsigned long long kept = 0;
        CUDA_OK (cudaMemcpyAsync (&kept, c->d_counts + 0, sizeof (kept), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK (cudaStreamSynchronize (c->stream));
        c->n_generated = kept;
     }
   CUDA_OK (cudaGetLastError ());
   return 0;
}

// ExposureTime handling of marx_create_photons (source.c:323-334): valid directly after marxb200_create_photons.
extern "C" int marxb200_truncate_exposure (marxb200_ctx *c, double exposure_left, uint64_t *n_kept)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (c->stage_done != 0) return fail ("marxb200_truncate_exposure: call it directly after marxb200_create_photons");
   CUDA_OK (cudaSetDevice (c->device));
   launch_exposure_truncate (c->buf[c->cur], c->d_counts + 0, c->d_times, exposure_left, 1, c->stream);
   c->launches += 1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + 0, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   CUDA_OK (cudaGetLastError ());
   c->n_generated = n;
   if (n_kept) *n_kept = n;
   return 0;
}

// super-tile sums of the arrival-time increments of rays [first_ray, first_ray+n) -- the quantity ranks
// exchange (all-gather) to give every GPU its time base (DESIGN.md "multi-GPU").
extern "C" int marxb200_time_sums (marxb200_ctx *c, uint64_t first_ray, uint64_t n, double *sums_host, uint64_t max_sums, uint64_t *n_sums)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (!c->have_source) return fail ("marxb200_time_sums: no source set");
   if (n > c->capacity) return fail ("marxb200_time_sums: n exceeds capacity");
   CUDA_OK (cudaSetDevice (c->device));
   SourceArgs a;
   fill_source_args (c, a, first_ray, n, 0.0);
   double saved[2];
   CUDA_OK (cudaMemcpyAsync (saved, c->d_times, sizeof (saved), cudaMemcpyDeviceToHost, c->stream));
   launch_time_sums (a, c->stream);
   launch_time_scan (a, c->stream);
   c->launches += 4;
   CUDA_OK (cudaStreamSynchronize (c->stream));
   CUDA_OK (cudaMemcpyAsync (c->d_times, saved, sizeof (saved), cudaMemcpyHostToDevice, c->stream));
   uint64_t n_tiles = (n + kTile - 1) / kTile, ns = (n_tiles + kSuperTile - 1) / kSuperTile;
   if (ns > max_sums) return fail ("marxb200_time_sums: need room for %llu sums", (unsigned long long) ns);
   CUDA_OK (cudaMemcpyAsync (sums_host, c->d_super_sums, ns * sizeof (double), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (n_sums) *n_sums = ns;
   return 0;
}

// The detector-dither columns (dy, dz, dtheta) are live only for the ASPSOL model or after an upload that may carry them;
// otherwise the kernels that move or export whole records see null pointers there and skip / report zeros.
static bool det_dither_live (const marxb200_ctx *c) { return (c->D.mode == 2) || c->det_dither_dirty; }
static PhotonSoA observed (const marxb200_ctx *c, const PhotonSoA &b)
{
   PhotonSoA v = b;
   if (!det_dither_live (c)) { v.ddy = nullptr; v.ddz = nullptr; v.ddth = nullptr; }
   return v;
}

static int run_stage (marxb200_ctx *c, int stage)
{
   StageArgs a;
   memset (&a, 0, sizeof (a));
   if (c->stage_done < 0) return fail ("stage %d: no photons (call marxb200_create_photons or marxb200_upload first)", stage);
   a.seed = c->seed;
   a.compact = c->compact;
   a.source_distance = c->source_distance;
   a.rc = c->rc;
   // _marx_dither_detector is a no-op when DitherModel=NONE (detector.c:277-278), whatever the records carry
   a.det_dither = (det_dither_live (c) && (c->D.mode != 0)) ? 1 : 0;
   // the mirror stage runs as three kernels (HRMA phases A, B, C, mx_hrma.cuh), each re-packing its survivors
   // ... and the ACIS detector stage as two (geometry + QE, then FEF + streak; mx_acis.cuh)
   const bool k3_two = (stage == 3) && !c->detector_is_hrc && (c->k3_split != 0);
   // the mirror stage as A | B | C (in place, or MARXB200_K1_SPLIT=0) or as A | B1 | B2+C1 | C2 (compacting)
   const bool k1_four = (stage == 1) && c->compact && (c->k1_split != 0);
   static const int phases3[3] = {0, 1, 2}, phases4[4] = {0, 3, 4, 5};
   const int *phases = k1_four ? phases4 : phases3;
   // the grating stage as k2_select | k2_grating<1> (compacting) or as one kernel
   const bool k2_two = (stage == 2) && c->compact && (c->k2_split != 0);
   const int n_kernels = (stage == 1) ? (k1_four ? 4 : 3) : ((k3_two || k2_two) ? 2 : 1);
   const int k_first = (stage == 1) ? c->first_mirror_kernel : 0;
   unsigned long long *const tickets = c->d_ticket + 4 * (stage - 1);
   if (!c->batch_zeroed)
     {
        CUDA_OK (cudaMemsetAsync (tickets, 0, 4 * sizeof (unsigned long long), c->stream));
        if (c->compact)
          {
             // output counters grow by atomics: zero them (d_counts[4], [5], [7] = after the first, second, third mirror kernel;
             // [6] = between the ACIS kernels; d_counts[stage] = stage output)
             CUDA_OK (cudaMemsetAsync (c->d_counts + stage, 0, sizeof (unsigned long long), c->stream));
             if (stage == 1) CUDA_OK (cudaMemsetAsync (c->d_counts + 4 + k_first, 0, (2 - k_first) * sizeof (unsigned long long), c->stream));
             if (k1_four) CUDA_OK (cudaMemsetAsync (c->d_counts + 7, 0, sizeof (unsigned long long), c->stream));
             if (k3_two) CUDA_OK (cudaMemsetAsync (c->d_counts + 6, 0, sizeof (unsigned long long), c->stream));
             if (k2_two) CUDA_OK (cudaMemsetAsync (c->d_counts + 8, 0, sizeof (unsigned long long), c->stream));
          }
     }
   const unsigned long long *n_in = (k_first == 1) ? c->d_counts + 4 : c->d_counts + c->stage_done;
   c->first_mirror_kernel = 0;
   if ((stage == 1) && c->mirror_is_flat)
     {
        // MirrorType=FLATFIELD: the whole stage is one kernel (ffield.c:77-108)
        a.in = c->buf[c->cur];
        a.out = c->compact ? c->buf[1 - c->cur] : c->buf[c->cur];
        GUARD (c->compact ? 1 - c->cur : c->cur);
        a.n_in = c->d_counts + c->stage_done; a.n_out = c->d_counts + 1; a.ticket = tickets; a.chunk_tiles = 4;
        for (int k = 0; k < 5; k++) a.ff[k] = c->ff[k];
        prof_begin (c);
        launch_flatfield (a, c->num_sms, c->stream);
        prof_mark (c, 4);
        c->launches += 1;
        CUDA_OK (cudaGetLastError ());
        if (c->compact) { c->cur = 1 - c->cur; c->ordered = false; }
        c->prepack.done = false;
        c->stage_done = 1;
        return 0;
     }
   for (int k = k_first; k < n_kernels; k++)
     {
        a.in = c->buf[c->cur];
        a.out = c->compact ? c->buf[1 - c->cur] : c->buf[c->cur];
        GUARD (c->compact ? 1 - c->cur : c->cur);
        a.n_in = n_in;
        a.n_out = (k == n_kernels - 1) ? c->d_counts + stage
                  : ((stage == 1) ? c->d_counts + ((k == 2) ? 7 : 4 + k) : ((stage == 2) ? c->d_counts + 8 : c->d_counts + 6));
        a.ticket = tickets + k;
        // big inputs amortise the ticket atomic over several tiles; small ones need fine-grained balancing
        a.chunk_tiles = (stage == 1 && k == 0) ? 4 : ((stage == 1 && k < n_kernels - 1) ? 2 : 1);
        prof_begin (c);
        switch (stage)
          {
           case 1:
             {
                const int ph = phases[k];
                a.blob = c->blob1; a.blob_bytes = (ph == 1) ? c->k1b_bytes : ((ph == 3) ? c->k1b1_bytes : c->blob1_bytes);
                a.seg2_off = (ph == 4) ? c->k1b2_seg2_off : c->k1c_seg2_off;
                a.seg2_bytes = (ph == 2 || ph == 4 || ph == 5) ? c->k1c_seg2_bytes : 0;
                launch_hrma (a, ph, c->grid1[ph], c->stream);
                prof_mark (c, (ph < 3) ? 4 + ph : 8 + ph);          // classes 4, 5, 6 = A, B, C; 11, 12, 13 = B1, B2+C1, C2
                break;
             }
           case 2: a.blob = c->blob2; a.blob_bytes = c->blob2_bytes;
                   // k2_select lists (row, order) of the rays that leave the primary grating in out.ray; k2_grating<1> reads the
                   // rows of the SAME input list through it (no buffer swap in between)
                   launch_grating (a, (k2_two && k == 0) ? c->num_sms * 8 : c->grid2, c->stream, k2_two ? k + 1 : 0);
                   prof_mark (c, 7);
                   break;
           case 3: a.blob = c->blob3; a.blob_bytes = c->blob3_bytes;
                   if (c->detector_is_hrc) launch_hrc (a, c->grid3, c->stream); else launch_acis (a, c->grid3, c->stream, k3_two ? k + 1 : 0);
                   prof_mark (c, 8);
                   break;
          }
        c->launches += 1;
        CUDA_OK (cudaGetLastError ());
        if (c->compact && !(k2_two && k == 0)) c->cur = 1 - c->cur;
        n_in = a.n_out;
     }
   if (c->compact) c->ordered = false;
   c->prepack.done = false;
   c->stage_done = stage;
   return 0;
}

// Bring the live list back into arrival order (the compacting stage kernels emit survivors in completion
// order).  Called lazily before the list is observed and at the end of marxb200_trace.
static int ensure_order (marxb200_ctx *c)
{
   if (c->ordered || (c->stage_done <= 0)) { c->ordered = true; return 0; }
   OrderArgs o;
   o.in = c->buf[c->cur]; o.out = observed (c, c->buf[1 - c->cur]); o.rc = c->rc;
   GUARD (1 - c->cur);
   o.n_live = c->d_counts + c->stage_done;
   o.n_slots = c->n_generated;
   o.bitmap = c->d_bitmap; o.word_prefix = c->d_word_prefix; o.block_prefix = c->d_block_prefix; o.perm = c->d_perm;
   CUDA_OK (cudaMemsetAsync (c->d_bitmap, 0, (c->n_generated / 32 + 1) * sizeof (uint32_t), c->stream));
   int nl = 0;
   const bool pack = c->prepack.armed && (c->prepack.args.dst != nullptr);
   if (pack && (c->prepack.ev_free != nullptr)) CUDA_OK (cudaStreamWaitEvent (c->stream, c->prepack.ev_free, 0));
   prof_begin (c);
   launch_restore_order (o, c->num_sms, c->stream, &nl, pack ? &c->prepack.args : nullptr);
   c->prepack.done = pack;
   prof_mark (c, 9);
   c->launches += nl;
   CUDA_OK (cudaGetLastError ());
   c->cur = 1 - c->cur;
   c->ordered = true;
   return 0;
}

extern "C" int marxb200_mirror_reflect (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (!c->have_hrma) return fail ("marxb200_mirror_reflect: no HRMA tables set");
   CUDA_OK (cudaSetDevice (c->device));
   return run_stage (c, 1);
}
extern "C" int marxb200_grating_diffract (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (!c->have_grating) return fail ("marxb200_grating_diffract: no grating set");
   CUDA_OK (cudaSetDevice (c->device));
   if (c->grating_type == 0)
     {
        // GratingType=NONE: the stage is the identity (grating.c:87-110); carry the count forward
        if (c->stage_done < 0) return fail ("marxb200_grating_diffract: no photons");
        CUDA_OK (cudaMemcpyAsync (c->d_counts + 2, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        c->stage_done = 2;
        return 0;
     }
   return run_stage (c, 2);
}
extern "C" int marxb200_detect (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (!c->have_acis) return fail ("marxb200_detect: no detector set");
   CUDA_OK (cudaSetDevice (c->device));
   if (c->detector_type == 0)
     {
        if (c->stage_done < 0) return fail ("marxb200_detect: no photons");
        CUDA_OK (cudaMemcpyAsync (c->d_counts + 3, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        c->stage_done = 3;
        return 0;
     }
   return run_stage (c, 3);
}

extern "C" int marxb200_restore_order (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   CUDA_OK (cudaSetDevice (c->device));
   return ensure_order (c);
}

// marx_create_photons + HRMA phase A in one kernel (only the compacting path; the in-place parity mode and the
// stage-by-stage API keep the separate kernels)
static int enter_mirror_after_scan (marxb200_ctx *c, const SourceArgs &a);
static int create_and_enter_mirror (marxb200_ctx *c, uint64_t first_ray, uint64_t n, double time_base_in)
{
   if (!c->have_source) return fail ("marxb200_trace: no source set");
   if (!c->have_hrma) return fail ("marxb200_trace: no HRMA tables set");
   if (n > c->capacity) return fail ("marxb200_trace: n=%llu exceeds the allocated capacity %llu", (unsigned long long) n, (unsigned long long) c->capacity);
   CUDA_OK (cudaSetDevice (c->device));
   SourceArgs a;
   fill_source_args (c, a, first_ray, n, time_base_in);
   // the pre-pass of this batch may have run behind the previous one (look_ahead below)
   const bool hit = c->ahead_valid && (c->ahead_first == first_ray) && (c->ahead_n == n) && (c->ahead_epoch == c->source_epoch);
   if (c->ahead_valid) CUDA_OK (cudaStreamWaitEvent (c->stream, c->ev_ahead, 0));
   c->ahead_valid = false;
   prof_begin (c);
   if (hit)
     {
        c->ahead_hits++;
        a.tile_sums = c->ahead_tile_sums; a.supertile_sums = c->ahead_super_sums;
        prof_mark (c, 0);
        launch_time_scan (a, c->stream, false); prof_mark (c, 1);       // k0_time_bases/_tiles; also sets d_counts[0] = n
        c->launches += 2;
     }
   else
     {
        c->ahead_misses++;
        launch_time_sums (a, c->stream); prof_mark (c, 0);
        launch_time_scan (a, c->stream); prof_mark (c, 1);    // also sets d_counts[0] = n
        c->launches += 4;                     // k0_time_sums, k0_time_super/_bases/_tiles
     }
   CUDA_OK (cudaEventRecord (c->ev_ahead_consumed, c->stream));
   a.tile_sums = c->d_tile_sums; a.supertile_sums = c->d_super_sums;
   return enter_mirror_after_scan (c, a);
}

// The arrival-time pre-pass (k0_time_sums + k0_time_super: every ray's time increment drawn and summed per tile) of the batch a
// run asks for next -- the n rays right behind this one -- on the look-ahead stream: it fills the idle SM slots at the kernel
// boundaries of the batch being traced instead of standing in front of the next one.  The sums depend on the seed, the source
// and the ray indices only; a call for other rays, or after marxb200_set_source, finds no match and runs the pre-pass itself.
static int look_ahead (marxb200_ctx *c, uint64_t first_ray, uint64_t n)
{
   if (!c->ahead_on || c->profiling || (n == 0) || (first_ray + n < first_ray) || (c->ahead_tile_sums == nullptr)) return 0;
   SourceArgs b;
   fill_source_args (c, b, first_ray, n, 0.0);
   b.tile_sums = c->ahead_tile_sums; b.supertile_sums = c->ahead_super_sums;
   CUDA_OK (cudaStreamWaitEvent (c->ahead_stream, c->ev_ahead_consumed, 0));      // the scratch is free once this batch's bases exist
   launch_time_sums (b, c->ahead_stream);
   launch_time_super (b, c->ahead_stream);
   c->launches += 2;
   CUDA_OK (cudaGetLastError ());
   CUDA_OK (cudaEventRecord (c->ev_ahead, c->ahead_stream));
   c->ahead_valid = true; c->ahead_first = first_ray; c->ahead_n = n; c->ahead_epoch = c->source_epoch;
   return 0;
}

// k01_source_hrma behind a finished arrival-time scan (tile bases, batch start and count on the device)
static int enter_mirror_after_scan (marxb200_ctx *c, const SourceArgs &a)
{
   const uint64_t n = a.n;
   StageArgs st;
   memset (&st, 0, sizeof (st));
   st.out = c->buf[1];
   st.n_out = c->d_counts + 4;
   // slot 0 of the mirror stage's tickets: the stage call that follows starts at its kernel 1.  MARXB200_K01_TICKET=0: static stride
   st.ticket = (c->batch_zeroed && c->k01_ticket) ? c->d_ticket : nullptr;
   st.seed = c->seed; st.compact = 1; st.source_distance = c->source_distance;
   st.blob = c->blob1; st.blob_bytes = c->blob1_bytes;
   if (!c->batch_zeroed) CUDA_OK (cudaMemsetAsync (c->d_counts + 4, 0, sizeof (unsigned long long), c->stream));
   GUARD (1);
   prof_begin (c);
   launch_source_hrma (a, st, c->grid01, c->stream); prof_mark (c, 3);
   c->launches += (n != 0) ? 1 : 0;      // k01_source_hrma
   CUDA_OK (cudaGetLastError ());
   c->cur = 1; c->stage_done = 0; c->n_generated = n; c->ordered = false; c->prepack.done = false;
   c->first_mirror_kernel = 1;
   return 0;
}

extern "C" int marxb200_set_profiling (marxb200_ctx *c, int on)
{
   if (c == nullptr) return fail ("NULL ctx");
   prof_collect (c);
   c->profiling = (on != 0);
   return 0;
}
extern "C" int marxb200_get_kernel_ms (marxb200_ctx *c, double ms[MARXB200_NUM_KERNEL_CLASSES], uint64_t launches[MARXB200_NUM_KERNEL_CLASSES])
{
   if (c == nullptr) return fail ("NULL ctx");
   CUDA_OK (cudaSetDevice (c->device));
   prof_collect (c);
   for (int i = 0; i < MARXB200_NUM_KERNEL_CLASSES; i++)
     {
        if (ms) ms[i] = c->prof_ms[i];
        if (launches) launches[i] = c->prof_n[i];
        c->prof_ms[i] = 0.0; c->prof_n[i] = 0;
     }
   return 0;
}

extern "C" int marxb200_trace (marxb200_ctx *c, uint64_t first_ray, uint64_t n) { return marxb200_trace_from (c, first_ray, n, -1.0); }

extern "C" int marxb200_trace_from (marxb200_ctx *c, uint64_t first_ray, uint64_t n, double time_base_in)
{
   if (c == nullptr) return fail ("NULL ctx");
   // the fused source + HRMA-A kernel serves the compacting path of the NONE / INTERNAL dither models; the ASPSOL model
   // (end-of-file cut, detector dither columns) and lists that follow an upload carrying detector dither go stage by stage
   const bool fused = c->compact && (c->D.mode != 2) && !c->det_dither_dirty && !c->mirror_is_flat;
   CUDA_OK (cudaSetDevice (c->device));
   if (c->compact && (-1 == mxb_begin_batch (c))) return -1;
   int status = 0;
   if (fused) status = create_and_enter_mirror (c, first_ray, n, time_base_in);
   else status = marxb200_create_photons (c, first_ray, n, time_base_in);
   if (status == 0) status = marxb200_mirror_reflect (c);
   if (status == 0) status = marxb200_grating_diffract (c);
   if (status == 0) status = marxb200_detect (c);
   c->batch_zeroed = false;
   if (status != 0) return -1;
   if (-1 == ensure_order (c)) return -1;
   if (fused && (first_ray + n > first_ray)) return look_ahead (c, first_ray + n, n);
   return 0;
}

// One clear for everything a traced batch counts with atomics: counts[1..] and every kernel's ticket.  counts[0] (rays generated)
// is written by the time scan; nothing between here and the end of the batch reads a count of the previous batch.
int mxb_begin_batch (marxb200_ctx *c)
{
   CUDA_OK (cudaMemsetAsync (c->d_counts + 1, 0, (marxb200_ctx::kNumCounts - 1 + marxb200_ctx::kNumTickets) * sizeof (unsigned long long), c->stream));
   c->batch_zeroed = true;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// counts and host boundary
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_get_stage_counts (marxb200_ctx *c, uint64_t counts[4])
{
   if (c == nullptr) return fail ("NULL ctx");
   CUDA_OK (cudaSetDevice (c->device));
   unsigned long long h[4];
   CUDA_OK (cudaMemcpyAsync (h, c->d_counts, sizeof (h), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   for (int i = 0; i < 4; i++) counts[i] = (i <= c->stage_done) ? h[i] : 0;
   return 0;
}

extern "C" int marxb200_get_internal_counts (marxb200_ctx *c, uint64_t counts[8])
{
   if (c == nullptr) return fail ("NULL ctx");
   CUDA_OK (cudaSetDevice (c->device));
   unsigned long long h[8];
   CUDA_OK (cudaMemcpyAsync (h, c->d_counts, sizeof (h), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   for (int i = 0; i < 8; i++) counts[i] = h[i];
   return 0;
}

extern "C" int marxb200_get_counts (marxb200_ctx *c, uint64_t *n_generated, uint64_t *n_live, double *total_time)
{
   if (c == nullptr) return fail ("NULL ctx");
   uint64_t cnt[4];
   if (-1 == marxb200_get_stage_counts (c, cnt)) return -1;
   double times[2];
   CUDA_OK (cudaMemcpyAsync (times, c->d_times, sizeof (times), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (n_generated) *n_generated = c->n_generated;
   if (n_live) *n_live = (c->stage_done >= 0) ? cnt[c->stage_done] : 0;
   if (total_time) *total_time = times[1];
   return 0;
}

static int download_impl (marxb200_ctx *c, marxb200_photon_attr *out, uint64_t max_out, uint64_t *n_out)
{
   if ((c == nullptr) || (out == nullptr)) return fail ("download: NULL argument");
   if (c->stage_done < 0) { if (n_out) *n_out = 0; return 0; }
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_order (c)) return -1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + c->stage_done, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (n > max_out) n = max_out;
   if (n_out) *n_out = n;
   if (n == 0) return 0;
   if (-1 == ensure_aos (c, n)) return -1;
   launch_soa_to_aos (observed (c, c->buf[c->cur]), c->d_counts + c->stage_done, n, c->d_aos, c->d_times, c->stream);
   c->launches += 1;
   CUDA_OK (cudaMemcpyAsync (out, c->d_aos, (size_t) n * sizeof (marxb200_photon_attr), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   return 0;
}

extern "C" int marxb200_download (marxb200_ctx *c, marxb200_photon_attr *out, uint64_t max_out, uint64_t *n_out)
{
   if (c && !c->compact) return fail ("marxb200_download: compaction is off; use marxb200_download_all");
   return download_impl (c, out, max_out, n_out);
}
extern "C" int marxb200_download_all (marxb200_ctx *c, marxb200_photon_attr *out, uint64_t max_out, uint64_t *n_out)
{
   return download_impl (c, out, max_out, n_out);
}

extern "C" int marxb200_upload (marxb200_ctx *c, const marxb200_photon_attr *in, uint64_t n, const uint64_t *ray_ids)
{
   return marxb200_upload_from (c, in, n, ray_ids, 0.0);      // arrival times are absolute
}

extern "C" int marxb200_upload_from (marxb200_ctx *c, const marxb200_photon_attr *in, uint64_t n, const uint64_t *ray_ids, double start_time)
{
   if ((c == nullptr) || (in == nullptr)) return fail ("marxb200_upload: NULL argument");
   if (n > c->capacity) return fail ("marxb200_upload: n exceeds capacity");
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_aos (c, n ? n : 1)) return -1;
   CUDA_OK (cudaMemcpyAsync (c->d_aos, in, (size_t) n * sizeof (marxb200_photon_attr), cudaMemcpyHostToDevice, c->stream));
   uint64_t *d_ids = nullptr;
   if (ray_ids)
     {
        if (c->upload_ids_cap < n)
          {
             CUDA_OK (cudaStreamSynchronize (c->stream));
             if (c->d_upload_ids) cudaFree (c->d_upload_ids);
             c->d_upload_ids = nullptr; c->upload_ids_cap = 0;
             CUDA_OK (cudaMalloc (&c->d_upload_ids, (n + n / 4 + 1024) * sizeof (uint64_t)));
             c->upload_ids_cap = n + n / 4 + 1024;
          }
        d_ids = c->d_upload_ids;
        CUDA_OK (cudaMemcpyAsync (d_ids, ray_ids, n * sizeof (uint64_t), cudaMemcpyHostToDevice, c->stream));
     }
   c->cur = 0;
   // the list keeps absolute times (pt->start_time + arrival_time); d_times[0] = the batch start the AoS download subtracts
   CUDA_OK (cudaMemcpyAsync (c->d_times, &start_time, sizeof (double), cudaMemcpyHostToDevice, c->stream));
   GUARD (0);
   launch_aos_to_soa (c->d_aos, d_ids, n, c->buf[0], c->rc, start_time, c->stream);
   c->launches += 1;
   unsigned long long nn = n;
   CUDA_OK (cudaMemcpyAsync (c->d_counts + 0, &nn, sizeof (nn), cudaMemcpyHostToDevice, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->stage_done = 0; c->n_generated = n; c->ordered = true; c->first_mirror_kernel = 0; c->prepack.done = false;
   c->det_dither_dirty = true;           // the records may carry dy, dz, dtheta (an ASPSOL run dumped to a rayfile)
   return 0;
}

extern "C" int marxb200_download_columns (marxb200_ctx *c, const marxb200_columns *cols, uint64_t max_out, uint64_t *n_out)
{
   if ((c == nullptr) || (cols == nullptr)) return fail ("marxb200_download_columns: NULL argument");
   if (c->stage_done < 0) { if (n_out) *n_out = 0; return 0; }
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_order (c)) return -1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + c->stage_done, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (n > max_out) n = max_out;
   if (n_out) *n_out = n;
   if (n == 0) return 0;
   const PhotonSoA &b = c->buf[c->cur];
#define COL(dst, src, T) if (cols->dst) CUDA_OK (cudaMemcpyAsync (cols->dst, b.src, (size_t) n * sizeof (T), cudaMemcpyDeviceToHost, c->stream))
   COL (energy, energy, double); COL (time, time, double);
   COL (xpos, x0, double); COL (ypos, x1, double); COL (zpos, x2, double);
   COL (xcos, p0, double); COL (ycos, p1, double); COL (zcos, p2, double);
   COL (chipx, chipx, float); COL (chipy, chipy, float); COL (pi, pi, float);
   COL (pha, pha, int16_t); COL (ccd, ccd, int8_t); COL (order, order, int8_t); COL (shell, shell, int8_t);
   COL (ray, ray, uint64_t);
#undef COL
   CUDA_OK (cudaStreamSynchronize (c->stream));
   return 0;
}

// ---------------------------------------------------------------------------------------------
// event tallies (include/marxb200.h): device-resident histograms of the live list
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_tally_create (marxb200_ctx *c, const marxb200_tally_axis *axes, int naxes)
{
   if ((c == nullptr) || (axes == nullptr)) return fail ("marxb200_tally_create: NULL argument");
   if ((naxes < 1) || (naxes > 2)) return fail ("marxb200_tally_create: 1 or 2 axes, not %d", naxes);
   marxb200_ctx::Tally t;
   memset (&t.plan, 0, sizeof (t.plan));
   t.plan.naxes = naxes;
   t.total = 1;
   for (int a = 0; a < naxes; a++)
     {
        if ((axes[a].column < 0) || (axes[a].column >= TALLY_NUM_COLUMNS)) return fail ("marxb200_tally_create: unknown column %d", axes[a].column);
        if ((axes[a].nbins == 0) || !(axes[a].hi > axes[a].lo)) return fail ("marxb200_tally_create: axis %d needs nbins > 0 and hi > lo", a);
        t.plan.ax[a].column = axes[a].column; t.plan.ax[a].nbins = axes[a].nbins;
        t.plan.ax[a].lo = axes[a].lo; t.plan.ax[a].scale = (double) axes[a].nbins / (axes[a].hi - axes[a].lo);
        t.total *= axes[a].nbins;
     }
   if (t.total > (1ull << 28)) return fail ("marxb200_tally_create: %llu bins is beyond the 2^28 limit", (unsigned long long) t.total);
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaMalloc (&t.bins, t.total * sizeof (unsigned long long)));
   CUDA_OK (cudaMemsetAsync (t.bins, 0, t.total * sizeof (unsigned long long), c->stream));
   c->allocs.push_back (std::make_pair ((void *) t.bins, (int) marxb200_ctx::TAG_MISC));
   c->tallies.push_back (t);
   return (int) c->tallies.size () - 1;
}
static marxb200_ctx::Tally *get_tally (marxb200_ctx *c, int id)
{
   if (c == nullptr) { fail ("NULL ctx"); return nullptr; }
   if ((id < 0) || (id >= (int) c->tallies.size ())) { fail ("unknown tally id %d", id); return nullptr; }
   return &c->tallies[id];
}
extern "C" int marxb200_tally_accumulate (marxb200_ctx *c, int id)
{
   marxb200_ctx::Tally *t = get_tally (c, id);
   if (t == nullptr) return -1;
   if (c->stage_done < 0) return fail ("marxb200_tally_accumulate: no photons");
   CUDA_OK (cudaSetDevice (c->device));
   // the compacting kernels leave energy and time in the per-ray constants until the order is restored
   if (-1 == ensure_order (c)) return -1;
   launch_tally (c->buf[c->cur], c->d_counts + c->stage_done, c->n_generated, t->plan, t->bins, c->num_sms, c->stream);
   c->launches += 1;
   CUDA_OK (cudaGetLastError ());
   return 0;
}
extern "C" int marxb200_tally_reset (marxb200_ctx *c, int id)
{
   marxb200_ctx::Tally *t = get_tally (c, id);
   if (t == nullptr) return -1;
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaMemsetAsync (t->bins, 0, t->total * sizeof (unsigned long long), c->stream));
   return 0;
}
extern "C" int marxb200_tally_read (marxb200_ctx *c, int id, uint64_t *out, uint64_t max_bins)
{
   marxb200_ctx::Tally *t = get_tally (c, id);
   if (t == nullptr) return -1;
   if ((out == nullptr) || (max_bins < t->total)) return fail ("marxb200_tally_read: the buffer must hold %llu counters", (unsigned long long) t->total);
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaMemcpyAsync (out, t->bins, t->total * sizeof (unsigned long long), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   return 0;
}
extern "C" int marxb200_tally_device_ptr (marxb200_ctx *c, int id, void **dev_ptr, uint64_t *num_bins)
{
   marxb200_ctx::Tally *t = get_tally (c, id);
   if (t == nullptr) return -1;
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   if (dev_ptr) *dev_ptr = t->bins;
   if (num_bins) *num_bins = t->total;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// marx_write_photons (marxio.c:403-476) for the device-resident live list: bulk column files, byte-compatible with
// the reference's (32-byte header marxio.c:151-205 + big-endian column data), without the AoS detour and without one
// fwrite per photon per column.
// ---------------------------------------------------------------------------------------------
typedef MxEgressCol EgressCol;
// Outfile_Info_Table, marxio.c:292-322 (same files, names, type letters)
extern const MxEgressCol kMxEgressCols[] = {
   {MARXB200_PI_OK, "b_energy.dat", "B_ENERGY", 'E', EGRESS_PI, 4},
   {MARXB200_ENERGY_OK, "energy.dat", "ENERGY", 'E', EGRESS_ENERGY, 4},
   {MARXB200_TIME_OK, "time.dat", "TIME", 'E', EGRESS_TIME, 4},
   {MARXB200_TAG_OK, "tag.dat", "TAG", 'J', EGRESS_TAG, 4},
   {MARXB200_X_VECTOR_OK, "xpos.dat", "XPOS", 'E', EGRESS_XPOS, 4},
   {MARXB200_X_VECTOR_OK, "ypos.dat", "YPOS", 'E', EGRESS_YPOS, 4},
   {MARXB200_X_VECTOR_OK, "zpos.dat", "ZPOS", 'E', EGRESS_ZPOS, 4},
   {MARXB200_P_VECTOR_OK, "xcos.dat", "COSX", 'E', EGRESS_XCOS, 4},
   {MARXB200_P_VECTOR_OK, "ycos.dat", "COSY", 'E', EGRESS_YCOS, 4},
   {MARXB200_P_VECTOR_OK, "zcos.dat", "COSZ", 'E', EGRESS_ZCOS, 4},
   {MARXB200_PULSEHEIGHT_OK, "pha.dat", "PHA", 'I', EGRESS_PHA, 2},
   {MARXB200_DET_NUM_OK, "detector.dat", "CCDID", 'A', EGRESS_CCD, 1},
   {MARXB200_DET_PIXEL_OK, "xpixel.dat", "CHIPX", 'E', EGRESS_CHIPX, 4},
   {MARXB200_DET_PIXEL_OK, "ypixel.dat", "CHIPY", 'E', EGRESS_CHIPY, 4},
   {MARXB200_DET_UV_PIXEL_OK, "hrc_u.dat", "U", 'E', EGRESS_HRC_U, 4},
   {MARXB200_DET_UV_PIXEL_OK, "hrc_v.dat", "V", 'E', EGRESS_HRC_V, 4},
   {MARXB200_MIRROR_SHELL_OK, "mirror.dat", "MIRROR", 'I', EGRESS_MIRROR, 2},
   {MARXB200_DET_REGION_OK, "hrcregion.dat", "HRCREGION", 'A', EGRESS_REGION, 1},
   {MARXB200_ORDER_OK, "order.dat", "ORDER", 'A', EGRESS_ORDER, 1},
   {MARXB200_ORDER1_OK, "ofine.dat", "FINE", 'A', EGRESS_ORDER1, 1},
   {MARXB200_ORDER2_OK, "ocoarse1.dat", "COARSE1", 'A', EGRESS_ORDER2, 1},
   {MARXB200_ORDER3_OK, "ocoarse2.dat", "COARSE2", 'A', EGRESS_ORDER3, 1},
   {MARXB200_ORDER4_OK, "ocoarse3.dat", "COARSE3", 'A', EGRESS_ORDER4, 1},
   {MARXB200_SKY_DITHER_OK, "sky_ra.dat", "RA", 'E', EGRESS_SKY_RA, 4},
   {MARXB200_SKY_DITHER_OK, "sky_dec.dat", "DEC", 'E', EGRESS_SKY_DEC, 4},
   {MARXB200_SKY_DITHER_OK, "sky_roll.dat", "ROLL", 'E', EGRESS_SKY_ROLL, 4},
   {MARXB200_DET_DITHER_OK, "det_dy.dat", "DET_DY", 'E', EGRESS_DET_DY, 4},        // 0 for INTERNAL dither (dither.c:167-182)
   {MARXB200_DET_DITHER_OK, "det_dz.dat", "DET_DZ", 'E', EGRESS_DET_DZ, 4},
   {MARXB200_DET_DITHER_OK, "det_theta.dat", "DET_THETA", 'E', EGRESS_DET_THETA, 4},
};
#define kEgressCols kMxEgressCols
constexpr int kNumEgressCols = (int) (sizeof (kMxEgressCols) / sizeof (kMxEgressCols[0]));
extern const int kMxNumEgressCols = kNumEgressCols;
static_assert (kNumEgressCols <= kMaxEgressCols, "EgressPlan too small");
namespace {

void put_be32 (unsigned char *b, uint32_t v) { b[0] = (unsigned char) (v >> 24); b[1] = (unsigned char) (v >> 16); b[2] = (unsigned char) (v >> 8); b[3] = (unsigned char) v; }
}

extern "C" int marxb200_write_photons (marxb200_ctx *c, const char *dir, uint64_t write_mask, int open_mode, double total_time)
{
   if ((c == nullptr) || (dir == nullptr)) return fail ("marxb200_write_photons: NULL argument");
   if (c->stage_done < 0) return fail ("marxb200_write_photons: no photons");
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_order (c)) return -1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + c->stage_done, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));

   EgressPlan plan;
   memset (&plan, 0, sizeof (plan));
   int which[kMaxEgressCols];
   uint64_t total = 0;
   for (int k = 0; k < kNumEgressCols; k++)
     {
        if (0 == (kEgressCols[k].mask & write_mask)) continue;
        which[plan.num_cols] = k;
        plan.kind[plan.num_cols] = kEgressCols[k].kind;
        plan.offset[plan.num_cols] = total;
        plan.num_cols++;
        total += (uint64_t) align16 ((size_t) n * kEgressCols[k].size);
     }
   // where the batch lands on the host: the context's pinned buffer, or -- with the background writer -- one of its two
   // buffers, once the writes that still read it are done
   unsigned char *h_dst = nullptr;
   int wb = 0;
   if (c->writer != nullptr)
     {
        std::string werr;
        if (open_mode && (-1 == mxw_flush (c->writer, &werr))) return fail ("marxb200_write_photons: %s", werr.c_str ());
        if (mxw_failed (c->writer, &werr)) return fail ("marxb200_write_photons: %s", werr.c_str ());
        wb = c->wbuf_next;
        c->wbuf_next ^= 1;
        mxw_wait_buffer (c->writer, wb);
        if (c->h_wbuf_bytes[wb] < total + 64)
          {
             if (c->h_wbuf[wb]) cudaFreeHost (c->h_wbuf[wb]);
             c->h_wbuf[wb] = nullptr; c->h_wbuf_bytes[wb] = 0;
             const size_t want = (size_t) total + (size_t) total / 4 + 65536;
             CUDA_OK (cudaMallocHost (&c->h_wbuf[wb], want));
             c->h_wbuf_bytes[wb] = want;
          }
        h_dst = (unsigned char *) c->h_wbuf[wb];
     }
   if (n && plan.num_cols)
     {
        // the AoS staging buffer (136 B per photon) is always large enough: at most 29 columns x 4 B = 116 B per row
        if (-1 == ensure_aos (c, n + 16)) return -1;
        if (c->writer == nullptr)
          {
             if (-1 == ensure_pinned (c, (size_t) total)) return -1;
             h_dst = (unsigned char *) c->h_pinned;
          }
        launch_egress_pack (observed (c, c->buf[c->cur]), c->d_counts + c->stage_done, n, plan, c->d_aos, c->d_times, total_time, c->stream);
        c->launches += 1;
        CUDA_OK (cudaMemcpyAsync (h_dst, c->d_aos, (size_t) total, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK (cudaStreamSynchronize (c->stream));
        CUDA_OK (cudaGetLastError ());
     }
   if (c->writer != nullptr)
     {
        for (int j = 0; j < plan.num_cols; j++)
          {
             const EgressCol &col = kEgressCols[which[j]];
             MxWriteTask t;
             t.path = std::string (dir) + "/" + col.file;
             t.create = open_mode ? 1 : 0;
             static const unsigned char magic[4] = {0x83, 0x13, 0x89, 0x8D};
             memset (t.header, 0, sizeof (t.header));
             memcpy (t.header, magic, 4);
             t.header[4] = (unsigned char) col.type;
             strncpy ((char *) t.header + 5, col.colname, 15);
             if (open_mode) c->egress_rows[which[j]] = 0;
             c->egress_rows[which[j]] += n;
             put_be32 (t.rows_be, (uint32_t) c->egress_rows[which[j]]);
             t.data = h_dst + plan.offset[j]; t.bytes = (size_t) n * col.size;
             t.buffer = wb;
             mxw_submit (c->writer, which[j], t);
          }
        return 0;
     }
   for (int j = 0; j < plan.num_cols; j++)
     {
        const EgressCol &col = kEgressCols[which[j]];
        std::string path = std::string (dir) + "/" + col.file;
        FILE *fp;
        if (open_mode)
          {
             // marx_create_write_dump_file, marxio.c:151-205: magic, type letter, column name zero-padded to offset 20,
             // num_rows, num_cols (big-endian int32), 4 reserved bytes
             unsigned char hdr[32];
             static const unsigned char magic[4] = {0x83, 0x13, 0x89, 0x8D};
             memset (hdr, 0, sizeof (hdr));
             memcpy (hdr, magic, 4);
             hdr[4] = (unsigned char) col.type;
             strncpy ((char *) hdr + 5, col.colname, 15);
             if ((nullptr == (fp = fopen (path.c_str (), "w+b"))) || (32 != fwrite (hdr, 1, 32, fp)))
               { if (fp) fclose (fp); return fail ("marxb200_write_photons: unable to create %s", path.c_str ()); }
             c->egress_rows[which[j]] = 0;
          }
        else
          {
             // "r+b" + seek to the end, as the reference does (marxio.c:336-340,378-386)
             if ((nullptr == (fp = fopen (path.c_str (), "r+b"))) || (0 != fseek (fp, 0, SEEK_END)))
               { if (fp) fclose (fp); return fail ("marxb200_write_photons: unable to open %s", path.c_str ()); }
          }
        const size_t bytes = (size_t) n * col.size;
        unsigned char rows_be[4];
        c->egress_rows[which[j]] += n;
        put_be32 (rows_be, (uint32_t) c->egress_rows[which[j]]);
        bool ok = (bytes == 0) || (bytes == fwrite ((const unsigned char *) c->h_pinned + plan.offset[j], 1, bytes, fp));
        ok = ok && (0 == fseek (fp, 20, SEEK_SET)) && (4 == fwrite (rows_be, 1, 4, fp));   // marx_close_write_dump_file :82-126
        if ((0 != fclose (fp)) || !ok) return fail ("marxb200_write_photons: write error on %s", path.c_str ());
     }
   return 0;
}

// Background writer for marxb200_write_photons: n_threads > 0 starts it (or keeps a running one), 0 flushes and stops it.
extern "C" int marxb200_set_async_writer (marxb200_ctx *c, int n_threads)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (n_threads < 0) return fail ("marxb200_set_async_writer: n_threads must be >= 0");
   if (n_threads == 0)
     {
        if (c->writer == nullptr) return 0;
        std::string werr;
        const int rc = mxw_flush (c->writer, &werr);
        mxw_destroy (c->writer);
        c->writer = nullptr;
        return (rc == 0) ? 0 : fail ("marxb200_write_photons: %s", werr.c_str ());
     }
   if (c->writer == nullptr) c->writer = mxw_create (n_threads);
   return 0;
}
// block until every queued column write has reached its file; reports the first write error
extern "C" int marxb200_write_flush (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (c->writer == nullptr) return 0;
   std::string werr;
   if (-1 == mxw_flush (c->writer, &werr)) return fail ("marxb200_write_photons: %s", werr.c_str ());
   return 0;
}

static size_t carve (PhotonSoA &b, unsigned char *base, uint64_t n);

extern "C" int marxb200_egress_begin (marxb200_ctx *c, uint64_t max_out)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (c->stage_done < 0) return fail ("marxb200_egress_begin: no photons");
   if (max_out == 0) return fail ("marxb200_egress_begin: max_out must be > 0");
   CUDA_OK (cudaSetDevice (c->device));
   if (c->copy_stream == nullptr)
     {
        CUDA_OK (cudaStreamCreateWithFlags (&c->copy_stream, cudaStreamNonBlocking));
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_staged, cudaEventDisableTiming));
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_copied, cudaEventDisableTiming));
        CUDA_OK (cudaMallocHost (&c->h_egress_count, sizeof (unsigned long long)));
     }
   if (c->egress_cap < max_out)
     {
        CUDA_OK (cudaStreamSynchronize (c->copy_stream));
        if (c->egress_slab) cudaFree (c->egress_slab);
        PhotonSoA tmp;
        size_t bytes = carve (tmp, nullptr, max_out);
        CUDA_OK (cudaMalloc (&c->egress_slab, bytes));
        carve (c->egress, (unsigned char *) c->egress_slab, max_out);
        c->egress_cap = max_out;
     }
   if (c->egress_pending) return fail ("marxb200_egress_begin: the previous egress was not ended");
   if (-1 == ensure_order (c)) return -1;
   const PhotonSoA &b = c->buf[c->cur], &e = c->egress;
   const uint64_t n = (max_out < c->capacity) ? max_out : c->capacity;
   // the staging area may still be read by the previous copy: make the main stream wait for it
   CUDA_OK (cudaStreamWaitEvent (c->stream, c->ev_copied, 0));
   if (c->prepack.owner == 0) { c->prepack.armed = false; c->prepack.done = false; }      // this call stages into the packed egress's first buffer
#define STAGE(col, T) CUDA_OK (cudaMemcpyAsync (e.col, b.col, (size_t) n * sizeof (T), cudaMemcpyDeviceToDevice, c->stream))
   STAGE (energy, double); STAGE (time, double); STAGE (x0, double); STAGE (x1, double); STAGE (x2, double);
   STAGE (p0, double); STAGE (p1, double); STAGE (p2, double); STAGE (chipx, float); STAGE (chipy, float); STAGE (pi, float);
   STAGE (pha, int16_t); STAGE (ccd, int8_t); STAGE (order, int8_t); STAGE (shell, uint8_t); STAGE (ray, uint64_t);
#undef STAGE
   CUDA_OK (cudaMemcpyAsync (c->h_egress_count, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaEventRecord (c->ev_staged, c->stream));
   c->egress_pending = true; c->egress_is_packed = false;
   return 0;
}

extern "C" int marxb200_egress_end (marxb200_ctx *c, const marxb200_columns *cols, uint64_t *n_out)
{
   if ((c == nullptr) || (cols == nullptr)) return fail ("marxb200_egress_end: NULL argument");
   if (!c->egress_pending || c->egress_is_packed) return fail ("marxb200_egress_end: no (unpacked) egress in flight");
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaEventSynchronize (c->ev_staged));
   unsigned long long n = *c->h_egress_count;
   if (n > c->egress_cap) n = c->egress_cap;
   const PhotonSoA &e = c->egress;
#define COL(dst, src, T) if (cols->dst) CUDA_OK (cudaMemcpyAsync (cols->dst, e.src, (size_t) n * sizeof (T), cudaMemcpyDeviceToHost, c->copy_stream))
   COL (energy, energy, double); COL (time, time, double);
   COL (xpos, x0, double); COL (ypos, x1, double); COL (zpos, x2, double);
   COL (xcos, p0, double); COL (ycos, p1, double); COL (zcos, p2, double);
   COL (chipx, chipx, float); COL (chipy, chipy, float); COL (pi, pi, float);
   COL (pha, pha, int16_t); COL (ccd, ccd, int8_t); COL (order, order, int8_t); COL (shell, shell, int8_t);
   COL (ray, ray, uint64_t);
#undef COL
   CUDA_OK (cudaEventRecord (c->ev_copied, c->copy_stream));
   CUDA_OK (cudaStreamSynchronize (c->copy_stream));
   c->egress_pending = false;
   if (n_out) *n_out = n;
   return 0;
}

// Pipelined PACKED egress: the column-file images marxb200_write_photons appends, produced on the device and landed in
// the caller's (pinned) host buffer while the next batch is being traced.  Staging reuses the egress slab; every column
// has a fixed region of align16 (max_out * size) bytes there, so the pack kernel needs no host knowledge of the count.
extern "C" int marxb200_egress_begin_packed (marxb200_ctx *c, uint64_t write_mask, double total_time, uint64_t max_out)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (c->stage_done < 0) return fail ("marxb200_egress_begin_packed: no photons");
   if (max_out == 0) return fail ("marxb200_egress_begin_packed: max_out must be > 0");
   if (c->egress_pending) return fail ("marxb200_egress_begin_packed: the previous egress was not ended");
   CUDA_OK (cudaSetDevice (c->device));
   if (c->copy_stream == nullptr)
     {
        CUDA_OK (cudaStreamCreateWithFlags (&c->copy_stream, cudaStreamNonBlocking));
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_staged, cudaEventDisableTiming));
        CUDA_OK (cudaEventCreateWithFlags (&c->ev_copied, cudaEventDisableTiming));
        CUDA_OK (cudaMallocHost (&c->h_egress_count, sizeof (unsigned long long)));
     }
   if (max_out > c->capacity) max_out = c->capacity;
   if ((c->egress_cap < max_out) || (c->packed_slab_b == nullptr))
     {
        CUDA_OK (cudaStreamSynchronize (c->copy_stream));
        CUDA_OK (cudaStreamSynchronize (c->stream));
        if (c->prepack.owner == 0) c->prepack.armed = false;            // its destination is about to be freed
        if (c->egress_cap < max_out)
          {
             if (c->egress_slab) cudaFree (c->egress_slab);
             c->egress_slab = nullptr; c->egress_cap = 0;
             PhotonSoA tmp;
             size_t bytes = carve (tmp, nullptr, max_out);       // 126 B per row: more than any packed row (<= 116 B)
             CUDA_OK (cudaMalloc (&c->egress_slab, bytes));
             carve (c->egress, (unsigned char *) c->egress_slab, max_out);
             c->egress_cap = max_out;
          }
        if (c->packed_slab_b) cudaFree (c->packed_slab_b);
        c->packed_slab_b = nullptr;
        PhotonSoA tmp;
        CUDA_OK (cudaMalloc (&c->packed_slab_b, carve (tmp, nullptr, c->egress_cap)));
        for (int k = 0; k < 2; k++)
          if (c->ev_slab_free[k] == nullptr) CUDA_OK (cudaEventCreateWithFlags (&c->ev_slab_free[k], cudaEventDisableTiming));
     }
   if (-1 == ensure_order (c)) return -1;
   EgressPlan &plan = c->packed_plan;
   memset (&plan, 0, sizeof (plan));
   uint64_t total = 0;
   for (int k = 0; k < kNumEgressCols; k++)
     {
        if (0 == (kEgressCols[k].mask & write_mask)) continue;
        c->packed_which[plan.num_cols] = k;
        plan.kind[plan.num_cols] = kEgressCols[k].kind;
        plan.offset[plan.num_cols] = total;
        plan.num_cols++;
        total += (uint64_t) align16 ((size_t) max_out * kEgressCols[k].size);
     }
   // Two staging buffers alternate (the next batch's images may be written while this batch's are still being copied out).
   void *slab[2] = {c->egress_slab, c->packed_slab_b};
   const int k = c->packed_k ^ 1;
   mx::PackArgs want;
   memset (&want, 0, sizeof (want));
   want.plan = plan; want.dst = (unsigned char *) slab[k]; want.dev_start_time = c->d_times; want.total_time = total_time; want.max_rows = max_out;
   if (mxb_prepack_matches (c, 0, want))
     {
        // the order restoration at the end of the trace wrote these very images (pre-pack): only the count travels
        CUDA_OK (cudaStreamWaitEvent (c->stream, c->ev_staged, 0));
        CUDA_OK (cudaMemcpyAsync (c->h_egress_count, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK (cudaEventRecord (c->ev_staged, c->stream));
     }
   else
     {
        // The conversion runs on the COPY stream (behind the previous batches' copies), so that the context's stream goes straight on
        // with the next batch; what the kernel needs from the device scalars -- the event count, the batch start -- is snapshot in
        // stream order first, because the next batch rewrites both.
        CUDA_OK (cudaStreamWaitEvent (c->stream, c->ev_staged, 0));      // the previous conversion has consumed the last snapshot
        CUDA_OK (cudaMemcpyAsync (c->d_snap, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_OK (cudaMemcpyAsync (c->d_snap + 1, c->d_times, sizeof (double), cudaMemcpyDeviceToDevice, c->stream));
        if (-1 == mxb_reader_begin (c, 0, c->copy_stream)) return -1;
        launch_egress_pack (observed (c, c->buf[c->cur]), c->d_snap, max_out, plan, slab[k], (const double *) (c->d_snap + 1), total_time, c->copy_stream);
        c->launches += 1;
        CUDA_OK (cudaMemcpyAsync (c->h_egress_count, c->d_snap, sizeof (unsigned long long), cudaMemcpyDeviceToHost, c->copy_stream));
        CUDA_OK (cudaEventRecord (c->ev_staged, c->copy_stream));
        if (-1 == mxb_reader_end (c, 0, c->copy_stream)) return -1;
     }
   CUDA_OK (cudaGetLastError ());
   c->packed_k = k;
   // a run that egresses every batch the same way: let the next batch's order restoration write the images into the other buffer
   want.dst = (unsigned char *) slab[k ^ 1];
   mxb_prepack_arm (c, 0, want, c->ev_slab_free[k ^ 1]);
   c->packed_cap = max_out;
   c->egress_pending = true; c->egress_is_packed = true;
   return 0;
}

extern "C" int marxb200_egress_end_packed (marxb200_ctx *c, void *host, uint64_t host_bytes, marxb200_packed_layout *layout)
{
   if ((c == nullptr) || (host == nullptr) || (layout == nullptr)) return fail ("marxb200_egress_end_packed: NULL argument");
   if (!c->egress_pending || !c->egress_is_packed) return fail ("marxb200_egress_end_packed: no packed egress in flight");
   CUDA_OK (cudaSetDevice (c->device));
   CUDA_OK (cudaEventSynchronize (c->ev_staged));
   unsigned long long n = *c->h_egress_count;
   if (n > c->packed_cap) n = c->packed_cap;
   const EgressPlan &plan = c->packed_plan;
   memset (layout, 0, sizeof (*layout));
   layout->num_cols = (uint32_t) plan.num_cols;
   layout->n_rows = n;
   uint64_t off = 0;
   for (int j = 0; j < plan.num_cols; j++)
     {
        const EgressCol &col = kEgressCols[c->packed_which[j]];
        layout->mask[j] = col.mask; layout->type[j] = col.type; layout->elem_size[j] = (uint32_t) col.size;
        strncpy (layout->file[j], col.file, sizeof (layout->file[j]) - 1);
        layout->offset[j] = off;
        off += (uint64_t) align16 ((size_t) n * col.size);
     }
   c->egress_pending = false; c->egress_is_packed = false;
   if (off > host_bytes) return fail ("marxb200_egress_end_packed: the host buffer holds %llu bytes, %llu are needed", (unsigned long long) host_bytes, (unsigned long long) off);
   for (int j = 0; j < plan.num_cols; j++)
     if (n) CUDA_OK (cudaMemcpyAsync ((unsigned char *) host + layout->offset[j],
                                      (const unsigned char *) ((c->packed_k == 0) ? c->egress_slab : c->packed_slab_b) + plan.offset[j],
                                      (size_t) n * layout->elem_size[j], cudaMemcpyDeviceToHost, c->copy_stream));
   CUDA_OK (cudaEventRecord (c->ev_slab_free[c->packed_k], c->copy_stream));
   CUDA_OK (cudaEventRecord (c->ev_copied, c->copy_stream));
   CUDA_OK (cudaStreamSynchronize (c->copy_stream));
   return 0;
}

// measured FP64 peak of this GPU (DFMA chains), the denominator of the FP64 roofline
extern "C" int marxb200_measure_fp64_peak (marxb200_ctx *c, double *tflops)
{
   if ((c == nullptr) || (tflops == nullptr)) return fail ("marxb200_measure_fp64_peak: NULL argument");
   CUDA_OK (cudaSetDevice (c->device));
   const int grid = c->num_sms * 8, iters = 4096;
   cudaEvent_t e0, e1;
   CUDA_OK (cudaEventCreate (&e0)); CUDA_OK (cudaEventCreate (&e1));
   double best = 0.0;
   for (int rep = 0; rep < 6; rep++)     // rep 0 warms up
     {
        CUDA_OK (cudaEventRecord (e0, c->stream));
        launch_fp64_peak (c->d_times + 1, grid, iters, c->stream);   // the sink is never written
        CUDA_OK (cudaEventRecord (e1, c->stream));
        CUDA_OK (cudaEventSynchronize (e1));
        float ms = 0.f;
        CUDA_OK (cudaEventElapsedTime (&ms, e0, e1));
        const double tf = 2.0 * 64.0 * iters * 256.0 * grid / (ms * 1e-3) * 1e-12;
        if ((rep > 0) && (tf > best)) best = tf;
     }
   cudaEventDestroy (e0); cudaEventDestroy (e1);
   CUDA_OK (cudaGetLastError ());
   *tflops = best;
   return 0;
}

extern "C" int marxb200_get_launch_count (marxb200_ctx *c, uint64_t *n)
{
   if ((c == nullptr) || (n == nullptr)) return fail ("NULL argument");
   *n = c->launches;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// Level-1 event transforms (include/marxb200.h; kernels in level1_kernels.cu)
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_level1_reset (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("marxb200_level1_reset: NULL ctx");
   if (!c->have_level1) return fail ("marxb200_level1_reset: call marxb200_set_level1 first");
   CUDA_OK (cudaSetDevice (c->device));
   Level1State st;
   memset (&st, 0, sizeof (st));
   st.last_expno = -1;                          // static long last_expno = -1 (marx2fits.c:3743)
   CUDA_OK (cudaMemcpyAsync (c->d_l1_state, &st, sizeof (st), cudaMemcpyHostToDevice, c->stream));
   CUDA_OK (cudaMemsetAsync (c->d_l1_error, 0, sizeof (unsigned int), c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->l1_rows = 0;
   return 0;
}

extern "C" int marxb200_set_level1 (marxb200_ctx *c, const marxb200_level1_desc *d)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_set_level1: NULL argument");
   if ((d->detector_type < 1) || (d->detector_type > 4)) return fail ("marxb200_set_level1: detector type %d is not HRC-S/HRC-I/ACIS-S/ACIS-I", d->detector_type);
   if ((d->num_chips < 1) || (d->num_chips > MARXB200_L1_MAX_CHIPS)) return fail ("marxb200_set_level1: %d chips", d->num_chips);
   if ((d->pix_adjust < 0) || (d->pix_adjust > 3)) return fail ("marxb200_set_level1: unknown pixel adjustment %d", d->pix_adjust);
   const bool acis = (d->detector_type >= 3);
   int pix_adjust = d->pix_adjust;
   if (!acis && (pix_adjust == MARXB200_PIXADJ_EDSER)) pix_adjust = MARXB200_PIXADJ_RANDOMIZE;     // marx2fits main :3308-3309
   if ((pix_adjust == MARXB200_PIXADJ_EDSER) && ((d->subpix_npoints == nullptr) || (d->subpix_offset == nullptr) || (d->subpix_data == nullptr)))
     return fail ("marxb200_set_level1: EDSER needs the sub-pixel tables");
   if (!(d->fp_delta_s0 > 0.0)) return fail ("marxb200_set_level1: focal-plane pixel size must be positive");
   CUDA_OK (cudaSetDevice (c->device));
   begin_module (c, marxb200_ctx::TAG_LEVEL1);
   Level1Dev &L = c->L1;
   memset (&L, 0, sizeof (L));
   L.detector_type = d->detector_type; L.num_chips = d->num_chips;
   for (int k = 0; k < d->num_chips; k++)
     {
        const marxb200_level1_chip &s = d->chips[k];
        Level1ChipDev &g = L.chips[k];
        g.id = s.id; g.subpix_table = s.subpix_table ? 1 : 0;
        for (int j = 0; j < 3; j++) { g.x_ll[j] = s.x_ll[j]; g.xhat[j] = s.xhat[j]; g.yhat[j] = s.yhat[j]; }
        g.x_pixel_size = s.x_pixel_size; g.y_pixel_size = s.y_pixel_size;
        g.xpixel_offset = s.xpixel_offset; g.ypixel_offset = s.ypixel_offset;
        g.tdet_xoff = s.tdet_xoff; g.tdet_yoff = s.tdet_yoff;
     }
   L.fp_delta_s0 = d->fp_delta_s0; L.fp_x0 = d->fp_x0; L.fp_y0 = d->fp_y0;
   L.focal_length = d->focal_length;
   for (int j = 0; j < 3; j++) L.det_offset[j] = d->det_offset[j];
   L.time_del = d->time_del; L.time_start = d->time_start; L.pi_factor = d->pi_factor;
   // compute_xy_sky :3881-3883 evaluates these per event with the host's libm; the values are constants of the run
   const double theta = d->nominal_roll * 3.14159265358979323846 / 180.0;
   L.roll_cos = cos (theta); L.roll_sin = sin (theta);
   L.used_dither = d->used_dither ? 1 : 0; L.pix_adjust = pix_adjust;
   if (pix_adjust == MARXB200_PIXADJ_EDSER)
     {
        for (int k = 0; k < 2 * 256; k++)
          {
             const int32_t np = d->subpix_npoints[k];
             if ((np > 0) && ((np == 1) || ((uint64_t) d->subpix_offset[k] + 3ull * (uint64_t) np > d->subpix_data_len)))
               return fail ("marxb200_set_level1: sub-pixel table entry %d is malformed", k);
          }
        if (-1 == dev_upload_t (c, d->subpix_npoints, 2 * 256, &L.subpix_npoints)) return -1;
        if (-1 == dev_upload_t (c, d->subpix_offset, 2 * 256, &L.subpix_offset)) return -1;
        if (-1 == dev_upload_t (c, d->subpix_data, (size_t) d->subpix_data_len, &L.subpix_data)) return -1;
     }
   if (c->d_l1_state == nullptr)
     {
        // state + the two hand-over cells + the error flag in one allocation
        void *p = nullptr;
        CUDA_OK (cudaMalloc (&p, 256));
        c->d_l1_state = (Level1State *) p;
        c->d_l1_next_dither = (float *) ((char *) p + 64);
        c->d_l1_next_expno = (long long *) ((char *) p + 96);
        c->d_l1_error = (unsigned int *) ((char *) p + 128);
     }
   c->have_level1 = true;
   return marxb200_level1_reset (c);
}

static int level1_ensure_columns (marxb200_ctx *c, uint64_t n)
{
   if (n <= c->l1_cap) return 0;
   if (c->l1_slab) { CUDA_OK (cudaStreamSynchronize (c->stream)); cudaFree (c->l1_slab); c->l1_slab = nullptr; c->l1_cap = 0; }
   const uint64_t cap = ((n + n / 4 + 1023) / 256) * 256;
   // 5 x f64, 6 x i32, 1 x f32, 8 x i16, 1 x u8, head (u32), tile heads (u32 per 256 events)
   const size_t bytes = (size_t) cap * (5 * 8 + 6 * 4 + 4 + 8 * 2 + 1 + 4) + (size_t) (cap / 256 + 2) * 4 + 64 * 256;
   CUDA_OK (cudaMalloc (&c->l1_slab, bytes));
   char *p = (char *) c->l1_slab;
   auto take = [&] (size_t elem) { char *q = p; p += ((size_t) cap * elem + 255) & ~(size_t) 255; return (void *) q; };
   Level1Cols &o = c->l1_cols;
   o.time = (double *) take (8); o.detx = (double *) take (8); o.dety = (double *) take (8); o.x = (double *) take (8); o.y = (double *) take (8);
   o.expno = (int32_t *) take (4); o.tdetx = (int32_t *) take (4); o.tdety = (int32_t *) take (4); o.pha = (int32_t *) take (4);
   o.hrc_u = (int32_t *) take (4); o.hrc_v = (int32_t *) take (4);
   o.energy = (float *) take (4);
   o.ccd_id = (int16_t *) take (2); o.node_id = (int16_t *) take (2); o.chipx = (int16_t *) take (2); o.chipy = (int16_t *) take (2);
   o.pi = (int16_t *) take (2); o.fltgrade = (int16_t *) take (2); o.grade = (int16_t *) take (2); o.status = (int16_t *) take (2);
   o.keep = (uint8_t *) take (1);
   c->d_l1_head = (uint32_t *) take (4);
   c->d_l1_tile_head = (uint32_t *) p;
   c->l1_cap = cap;
   return 0;
}

extern "C" int marxb200_level1_transform (marxb200_ctx *c, double total_time)
{
   if (c == nullptr) return fail ("marxb200_level1_transform: NULL ctx");
   if (!c->have_level1) return fail ("marxb200_level1_transform: call marxb200_set_level1 first");
   if (c->stage_done < 0) return fail ("marxb200_level1_transform: no events");
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_order (c)) return -1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + c->stage_done, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   c->l1_rows = n;
   if (n == 0) return 0;
   if (n >= 0xFFFFFFFFull) return fail ("marxb200_level1_transform: %llu events in one batch", n);
   if (-1 == level1_ensure_columns (c, n)) return -1;
   Level1Args a;
   memset (&a, 0, sizeof (a));
   a.in = observed (c, c->buf[c->cur]);
   a.n_ptr = c->d_counts + c->stage_done; a.max_n = n;
   a.dev_start_time = c->d_times; a.total_time = total_time;
   a.seed = c->seed;
   a.L = c->L1; a.state = c->d_l1_state; a.out = c->l1_cols;
   a.head = c->d_l1_head; a.tile_head = c->d_l1_tile_head;
   a.next_dither = c->d_l1_next_dither; a.next_expno = c->d_l1_next_expno; a.error_flag = c->d_l1_error;
   int nl = 0;
   prof_begin (c);
   launch_level1 (a, c->num_sms, c->stream, &nl);
   prof_mark (c, 10);
   c->launches += (uint64_t) nl;
   CUDA_OK (cudaGetLastError ());
   return 0;
}

extern "C" int marxb200_level1_download (marxb200_ctx *c, const marxb200_level1_columns *cols, uint64_t max_out, uint64_t *n_out)
{
   if ((c == nullptr) || (cols == nullptr)) return fail ("marxb200_level1_download: NULL argument");
   if (!c->have_level1) return fail ("marxb200_level1_download: call marxb200_set_level1 first");
   CUDA_OK (cudaSetDevice (c->device));
   uint64_t n = c->l1_rows;
   if (n > max_out) n = max_out;
   if (n_out) *n_out = n;
   unsigned int err = 0;
   CUDA_OK (cudaMemcpyAsync (&err, c->d_l1_error, sizeof (err), cudaMemcpyDeviceToHost, c->stream));
   const Level1Cols &b = c->l1_cols;
#define COL(f, T) if (n && cols->f) CUDA_OK (cudaMemcpyAsync (cols->f, b.f, (size_t) n * sizeof (T), cudaMemcpyDeviceToHost, c->stream))
   COL (time, double); COL (detx, double); COL (dety, double); COL (x, double); COL (y, double);
   COL (expno, int32_t); COL (tdetx, int32_t); COL (tdety, int32_t); COL (pha, int32_t); COL (hrc_u, int32_t); COL (hrc_v, int32_t);
   COL (energy, float);
   COL (ccd_id, int16_t); COL (node_id, int16_t); COL (chipx, int16_t); COL (chipy, int16_t);
   COL (pi, int16_t); COL (fltgrade, int16_t); COL (grade, int16_t); COL (status, int16_t);
   COL (keep, uint8_t);
#undef COL
   CUDA_OK (cudaStreamSynchronize (c->stream));
   // the reference stops with an error on these rows (marx_compute_tiled_pixel: "chip = %d is not appropriate for this
   // detector", detpix.c:174; marx_mnc_to_fpc: "mnc.x is 0", :195)
   if (err != 0)
     {
        // reported once: later batches of the same file are not failed by this one
        CUDA_OK (cudaMemsetAsync (c->d_l1_error, 0, sizeof (unsigned int), c->stream));
        CUDA_OK (cudaStreamSynchronize (c->stream));
     }
   if (err & 1u) return fail ("marxb200_level1: an event's chip id does not belong to this detector");
   if (err & 2u) return fail ("marxb200_level1: mnc.x is 0");
   return 0;
}

// ---------------------------------------------------------------------------------------------
// aspect-solution table (marxasp's row loop, marxasp.c:996-1027)
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_aspsol_rows (marxb200_ctx *c, const marxb200_aspsol_desc *d, uint64_t first_row, uint64_t n,
                                     double *cols_host, void *fits_rows_host, double *device_ms)
{
   if ((c == nullptr) || (d == nullptr)) return fail ("marxb200_aspsol_rows: NULL argument");
   if (!(d->delta_time > 0.0) || (d->ra_period == 0.0) || (d->dec_period == 0.0) || (d->roll_period == 0.0))
     return fail ("marxb200_aspsol_rows: delta_time must be positive and the dither periods non-zero");
   // the reference counts rows in an unsigned int (marxasp.c:907,943)
   if (first_row + n > 0xFFFFFFFFull) return fail ("marxb200_aspsol_rows: row numbers beyond 2^32");
   if (device_ms) *device_ms = 0.0;
   if (n == 0) return 0;
   CUDA_OK (cudaSetDevice (c->device));
   static_assert (sizeof (AspsolDev) == sizeof (marxb200_aspsol_desc), "descriptor layouts must agree");
   AspsolDev D;
   memcpy (&D, d, sizeof (D));
   double *d_cols = nullptr; uint32_t *d_rows = nullptr;
   if (cols_host) CUDA_OK (cudaMallocAsync ((void **) &d_cols, (size_t) n * 8 * sizeof (double), c->stream));
   if (fits_rows_host) CUDA_OK (cudaMallocAsync ((void **) &d_rows, (size_t) n * MARXB200_ASPSOL_ROW_BYTES, c->stream));
   cudaEvent_t e0 = prof_event (c), e1 = prof_event (c);
   CUDA_OK (cudaEventRecord (e0, c->stream));
   launch_aspsol_rows (D, first_row, n, d_cols, d_rows, c->num_sms, c->stream);
   CUDA_OK (cudaEventRecord (e1, c->stream));
   c->launches += 1;
   CUDA_OK (cudaGetLastError ());
   if (cols_host) CUDA_OK (cudaMemcpyAsync (cols_host, d_cols, (size_t) n * 8 * sizeof (double), cudaMemcpyDeviceToHost, c->stream));
   if (fits_rows_host) CUDA_OK (cudaMemcpyAsync (fits_rows_host, d_rows, (size_t) n * MARXB200_ASPSOL_ROW_BYTES, cudaMemcpyDeviceToHost, c->stream));
   if (d_cols) CUDA_OK (cudaFreeAsync (d_cols, c->stream));
   if (d_rows) CUDA_OK (cudaFreeAsync (d_rows, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   float ms = 0.f;
   CUDA_OK (cudaEventElapsedTime (&ms, e0, e1));
   if (device_ms) *device_ms = ms;
   c->ev_pool.push_back (e0); c->ev_pool.push_back (e1);
   return 0;
}

// ---------------------------------------------------------------------------------------------
// ACIS pile-up (marxpileup's frame loop, marxpileup.c:1121-1213; kernels in pileup_kernels.cu)
// ---------------------------------------------------------------------------------------------
// in == NULL: the input columns are gathered on the device from the context's live event list (marxb200_pileup_events)
static int pileup_impl (marxb200_ctx *c, uint64_t n, const marxb200_pileup_in *in, double total_time, double alpha, double frame_time, uint64_t seed,
                        uint64_t max_out, const marxb200_pileup_out *out, uint64_t *n_out, double *device_ms)
{
   static const marxb200_pileup_in no_host_columns = {nullptr, nullptr, nullptr, nullptr, nullptr, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
   const bool from_list = (in == nullptr);
   if (from_list) in = &no_host_columns;
   *n_out = 0;
   if (device_ms) *device_ms = 0.0;
   if (!c->have_acis || c->detector_is_hrc || (c->detector_type == 0) || (c->blob3 == nullptr))
     return fail ("marxb200_pileup_run: the context has no ACIS detector (the PHA of an island comes from its FEF tables)");
   if (!(frame_time > 0.0)) return fail ("marxb200_pileup_run: frame_time must be positive");
   if (n >= 0xFFFFFFFFull) return fail ("marxb200_pileup_run: %llu events in one call", (unsigned long long) n);
   if (n == 0) return 0;
   if (!from_list && ((in->ccd == nullptr) || (in->x == nullptr) || (in->y == nullptr) || (in->t == nullptr) || (in->benergy == nullptr)))
     return fail ("marxb200_pileup_run: the detector, pixel, time and energy columns are required");
   CUDA_OK (cudaSetDevice (c->device));
   const uint64_t cap_out = (max_out < n) ? max_out : n;
   const uint64_t n_tiles = n / 256 + 2;
   // one slab: inputs (41 B/event), scratch (59 B/event), outputs (49 B/row), every column rounded up to 256 bytes
   auto rounded = [] (uint64_t rows, size_t elem) { return ((size_t) rows * elem + 255) & ~(size_t) 255; };
   size_t bytes = rounded (n, 1) + 10 * rounded (n, 4)                                         // inputs
     + 8 * rounded (n, 4) + 6 * rounded (n, 4) + rounded (n, 1) + rounded (n, 2) + rounded (n_tiles, 4)   // scratch
     + rounded (cap_out + 1, 1) + 11 * rounded (cap_out + 1, 4) + 2 * rounded (cap_out + 1, 2)  // outputs
     + 256                                                                                     // n_out, error
     + ((mx::pileup_fused_scratch_bytes (n) + 255) & ~(size_t) 255);                           // fused kernel: ticket + tile states
   char *slab = nullptr;
   CUDA_OK (cudaMallocAsync ((void **) &slab, bytes, c->stream));
   char *p = slab;
   auto take = [&] (uint64_t rows, size_t elem) { char *q = p; p += rounded (rows, elem); return (void *) q; };
   mx::PileupArgs a;
   memset (&a, 0, sizeof (a));
   int8_t *d_ccd = (int8_t *) take (n, 1);
   float *d_in[10];
   for (int k = 0; k < 10; k++) d_in[k] = (float *) take (n, 4);
   const float *h_in[10] = {in->x, in->y, in->t, in->benergy, in->dither[0], in->dither[1], in->dither[2], in->dither[3], in->dither[4], in->dither[5]};
   a.ccd = d_ccd; a.x = d_in[0]; a.y = d_in[1]; a.t = d_in[2]; a.benergy = d_in[3];
   for (int k = 0; k < 6; k++) a.dither[k] = (from_list || (h_in[4 + k] != nullptr)) ? d_in[4 + k] : nullptr;
   a.n = n; a.alpha = alpha; a.frame_time = frame_time; a.seed = seed;
   for (int k = 0; k < mx::kPuProbTable; k++) a.prob[k] = pow (alpha, (double) k);
   a.max_frame_events = 1u << 16;
   a.A = &((const mx::K3Blob *) c->blob3)->A;
   a.frame = (uint32_t *) take (n, 4); a.key = (uint32_t *) take (n, 4); a.lo = (uint32_t *) take (n, 4); a.hi = (uint32_t *) take (n, 4);
   a.pn = (uint32_t *) take (n, 4); a.in = (uint32_t *) take (n, 4); a.emit = (uint32_t *) take (n, 4); a.cum = (uint32_t *) take (n, 4);
   a.pb = (float *) take (n, 4); a.px = (float *) take (n, 4); a.py = (float *) take (n, 4); a.ib = (float *) take (n, 4);
   a.sx = (float *) take (n, 4); a.sy = (float *) take (n, 4);
   a.flag = (uint8_t *) take (n, 1); a.spha = (int16_t *) take (n, 2); a.tile_sum = (uint32_t *) take (n_tiles, 4);
   a.o_ccd = (int8_t *) take (cap_out + 1, 1);
   a.o_x = (float *) take (cap_out + 1, 4); a.o_y = (float *) take (cap_out + 1, 4); a.o_t = (float *) take (cap_out + 1, 4);
   a.o_benergy = (float *) take (cap_out + 1, 4); a.o_frame = (int32_t *) take (cap_out + 1, 4);
   for (int k = 0; k < 6; k++) { float *q = (float *) take (cap_out + 1, 4); a.o_dither[k] = (a.dither[k] && out->dither[k]) ? q : nullptr; }
   a.o_nphotons = (int16_t *) take (cap_out + 1, 2); a.o_pha = (int16_t *) take (cap_out + 1, 2);
   a.max_out = cap_out;
   a.n_out = (unsigned long long *) p; a.error = (unsigned int *) (p + 8);
   void *fused_scratch = (void *) (p + 256);
   // the fused single-kernel form serves lists whose exposure frames fit its shared-memory window (<= 513 events guaranteed);
   // a longer frame makes it raise a flag and the step kernels run instead (MARXB200_PILEUP_FUSED=0: always the step kernels)
   bool fused = true;
   if (const char *e = getenv ("MARXB200_PILEUP_FUSED")) fused = (atoi (e) != 0);
   int status = 0;
   unsigned long long rows = 0; unsigned int err = 0;
   cudaEvent_t e0 = prof_event (c), e1 = prof_event (c);
   do
     {
#define PU_OK(expr) { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { status = fail ("%s: %s", #expr, cudaGetErrorString (e_)); break; } }
        PU_OK (cudaMemsetAsync (p, 0, 16, c->stream));
        if (from_list)
          {
             // the columns marx_write_photons would put into detector.dat, xpixel.dat, ypixel.dat, time.dat, b_energy.dat and the six
             // dither files (marxio.c:217-290: TIME = (float) (arrival_time + total_time)), straight from the list in HBM
             mx::launch_pileup_gather (mxb_observed (c, c->buf[c->cur]), c->d_counts + c->stage_done, n, c->d_times, total_time,
                                       d_ccd, d_in, c->stream);
             c->launches += 1;
          }
        else PU_OK (cudaMemcpyAsync (d_ccd, in->ccd, (size_t) n, cudaMemcpyHostToDevice, c->stream));
        bool copied = true;
        for (int k = 0; k < 10; k++)
          if (!from_list && (h_in[k] != nullptr))
            {
               cudaError_t e_ = cudaMemcpyAsync (d_in[k], h_in[k], (size_t) n * 4, cudaMemcpyHostToDevice, c->stream);
               if (e_ != cudaSuccess) { status = fail ("marxb200_pileup_run: upload: %s", cudaGetErrorString (e_)); copied = false; break; }
            }
        if (!copied) break;
        // the fused kernel with a window of 256, 512 or 1024 events (frames of up to 129 / 257 / 513 events: a longer one raises the
        // fallback flag and the next size runs), then the step kernels.  MARXB200_PILEUP_WINDOW picks the first size tried.
        int nl = 0;
        const int attempts[4] = {256, 512, 1024, 0};
        int k_first = 2;
        if (const char *e = getenv ("MARXB200_PILEUP_WINDOW")) { const int w = atoi (e); k_first = (w <= 256) ? 0 : ((w <= 512) ? 1 : 2); }
        for (int k = fused ? k_first : 3; k < 4; k++)
          {
             PU_OK (cudaMemsetAsync (p, 0, 16, c->stream));
             PU_OK (cudaEventRecord (e0, c->stream));
             if (attempts[k]) mx::launch_pileup_fused (a, fused_scratch, attempts[k], c->num_sms, c->stream, &nl);
             else mx::launch_pileup (a, c->num_sms, c->stream, &nl);
             PU_OK (cudaEventRecord (e1, c->stream));
             c->launches += (uint64_t) nl;
             PU_OK (cudaGetLastError ());
             PU_OK (cudaMemcpyAsync (&rows, a.n_out, 8, cudaMemcpyDeviceToHost, c->stream));
             PU_OK (cudaMemcpyAsync (&err, a.error, 4, cudaMemcpyDeviceToHost, c->stream));
             PU_OK (cudaStreamSynchronize (c->stream));
             if (0 == (err & mx::kPuErrFallback)) break;
          }
        if (status != 0) break;
        if (err & mx::kPuErrCcd) { status = fail ("marxb200_pileup_run: an event's CCD id is outside 0..9"); break; }
        if (err & mx::kPuErrCorrupt) { status = fail ("marxb200_pileup_run: pixel coordinate beyond the chip (corrupt file?)"); break; }
        if (err & mx::kPuErrFrameTooLong) { status = fail ("marxb200_pileup_run: an exposure frame holds more than 65536 events"); break; }
        if (err & mx::kPuErrPha) { status = fail ("marxb200_pileup_run: no FEF for an island's chip region"); break; }
        if ((err & mx::kPuErrOverflow) || (rows > cap_out)) { status = fail ("marxb200_pileup_run: more than max_out = %llu rows", (unsigned long long) max_out); break; }
        const size_t r = (size_t) rows;
        if (r > 0)
          {
             if (out->ccd) PU_OK (cudaMemcpyAsync (out->ccd, a.o_ccd, r, cudaMemcpyDeviceToHost, c->stream));
             if (out->x) PU_OK (cudaMemcpyAsync (out->x, a.o_x, r * 4, cudaMemcpyDeviceToHost, c->stream));
             if (out->y) PU_OK (cudaMemcpyAsync (out->y, a.o_y, r * 4, cudaMemcpyDeviceToHost, c->stream));
             if (out->t) PU_OK (cudaMemcpyAsync (out->t, a.o_t, r * 4, cudaMemcpyDeviceToHost, c->stream));
             if (out->benergy) PU_OK (cudaMemcpyAsync (out->benergy, a.o_benergy, r * 4, cudaMemcpyDeviceToHost, c->stream));
             if (out->frame) PU_OK (cudaMemcpyAsync (out->frame, a.o_frame, r * 4, cudaMemcpyDeviceToHost, c->stream));
             if (out->nphotons) PU_OK (cudaMemcpyAsync (out->nphotons, a.o_nphotons, r * 2, cudaMemcpyDeviceToHost, c->stream));
             if (out->pha) PU_OK (cudaMemcpyAsync (out->pha, a.o_pha, r * 2, cudaMemcpyDeviceToHost, c->stream));
             bool ok = true;
             for (int k = 0; k < 6; k++)
               if (a.o_dither[k])
                 {
                    cudaError_t e_ = cudaMemcpyAsync (out->dither[k], a.o_dither[k], r * 4, cudaMemcpyDeviceToHost, c->stream);
                    if (e_ != cudaSuccess) { status = fail ("marxb200_pileup_run: download: %s", cudaGetErrorString (e_)); ok = false; break; }
                 }
             if (!ok) break;
             PU_OK (cudaStreamSynchronize (c->stream));
          }
        float ms = 0.f;
        PU_OK (cudaEventElapsedTime (&ms, e0, e1));
        if (device_ms) *device_ms = ms;
        *n_out = rows;
#undef PU_OK
     }
   while (0);
   cudaFreeAsync (slab, c->stream);
   cudaStreamSynchronize (c->stream);
   c->ev_pool.push_back (e0); c->ev_pool.push_back (e1);
   return status;
}

extern "C" int marxb200_pileup_run (marxb200_ctx *c, uint64_t n, const marxb200_pileup_in *in, double alpha, double frame_time, uint64_t seed,
                                    uint64_t max_out, const marxb200_pileup_out *out, uint64_t *n_out, double *device_ms)
{
   if ((c == nullptr) || (in == nullptr) || (out == nullptr) || (n_out == nullptr)) return fail ("marxb200_pileup_run: NULL argument");
   return pileup_impl (c, n, in, 0.0, alpha, frame_time, seed, max_out, out, n_out, device_ms);
}

// the same for the event list the detector stage left on the device: marx -> marxpileup without the column files in between
extern "C" int marxb200_pileup_events (marxb200_ctx *c, double total_time, double alpha, double frame_time, uint64_t seed,
                                       uint64_t max_out, const marxb200_pileup_out *out, uint64_t *n_out, double *device_ms)
{
   if ((c == nullptr) || (out == nullptr) || (n_out == nullptr)) return fail ("marxb200_pileup_events: NULL argument");
   *n_out = 0;
   if (c->stage_done != 3) return fail ("marxb200_pileup_events: the live list must be behind marxb200_detect");
   CUDA_OK (cudaSetDevice (c->device));
   if (-1 == ensure_order (c)) return -1;
   unsigned long long n = 0;
   CUDA_OK (cudaMemcpyAsync (&n, c->d_counts + 3, sizeof (n), cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK (cudaStreamSynchronize (c->stream));
   return pileup_impl (c, n, nullptr, total_time, alpha, frame_time, seed, max_out, out, n_out, device_ms);
}

// ---------------------------------------------------------------------------------------------
// helpers shared with comm.cu (mx_context.hpp)
// ---------------------------------------------------------------------------------------------
void mxb_prof_begin (marxb200_ctx *c) { prof_begin (c); }
void mxb_prof_mark (marxb200_ctx *c, int cls) { prof_mark (c, cls); }
int mxb_ensure_order (marxb200_ctx *c) { return ensure_order (c); }
PhotonSoA mxb_observed (const marxb200_ctx *c, const PhotonSoA &b) { return observed (c, b); }
void mxb_fill_source_args (marxb200_ctx *c, SourceArgs &a, uint64_t first_ray, uint64_t n, double time_base) { fill_source_args (c, a, first_ray, n, time_base); }
int mxb_enter_mirror_after_scan (marxb200_ctx *c, const SourceArgs &a) { return enter_mirror_after_scan (c, a); }
int mxb_finish_trace (marxb200_ctx *c)
{
   int status = marxb200_mirror_reflect (c);
   if (status == 0) status = marxb200_grating_diffract (c);
   if (status == 0) status = marxb200_detect (c);
   c->batch_zeroed = false;
   if (status != 0) return -1;
   return ensure_order (c);
}
// every selected column in a fixed region of align16 (rows_per_col * size) bytes; returns the total size
uint64_t mxb_build_egress_plan (uint64_t write_mask, uint64_t rows_per_col, EgressPlan &plan, int *which)
{
   memset (&plan, 0, sizeof (plan));
   uint64_t total = 0;
   for (int k = 0; k < kNumEgressCols; k++)
     {
        if (0 == (kEgressCols[k].mask & write_mask)) continue;
        if (which) which[plan.num_cols] = k;
        plan.kind[plan.num_cols] = kEgressCols[k].kind;
        plan.offset[plan.num_cols] = total;
        plan.num_cols++;
        total += (uint64_t) align16 ((size_t) rows_per_col * kEgressCols[k].size);
     }
   return total;
}

extern "C" int marxb200_load_calpack_impl (marxb200_ctx *ctx, const char *path, char *errbuf, size_t errlen);
extern "C" int marxb200_load_calpack (marxb200_ctx *c, const char *path)
{
   if ((c == nullptr) || (path == nullptr)) return fail ("marxb200_load_calpack: NULL argument");
   char buf[512]; buf[0] = 0;
   if (-1 == marxb200_load_calpack_impl (c, path, buf, sizeof (buf))) return fail ("%s", buf);
   return 0;
}
