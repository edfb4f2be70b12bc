/* marx_gpu_shim.c -- reference-side binding of libmarxb200.so into the UNMODIFIED marx driver.
 *
 * The stock marx/src/marx.c is linked with
 *    -Wl,--wrap=JDMsrandom,--wrap=marx_mirror_init,--wrap=marx_grating_init,--wrap=marx_detector_init,
 *        --wrap=marx_create_photons (RAYFILE / USER / SAOSAC / SIMPUT sources keep the stock one and inject its photons),--wrap=marx_mirror_reflect,--wrap=marx_grating_diffract,--wrap=marx_detect,
 *        --wrap=marx_write_photons,--wrap=marx_prune_photons,--wrap=marx_dump_to_rayfile,--wrap=marx_dealloc_photon_type
 * so that its calls (marx.c:245,254,263,569) reach the __wrap_* functions below while pfile parameter handling,
 * the stock *_init functions (calibration file readers), obs.par and the marxio/jdfits writers stay what they are
 * (option ii).  Per-photon work happens on the GPU only: if the configuration is not covered by the
 * CUDA path the run stops with an error -- there is no CPU fallback in here.
 *
 * Data flow per iteration of the driver loop (marx.c:545-608):
 *   marx_create_photons   -> marxb200_create_photons (+ marxb200_truncate_exposure when ExposureTime > 0)
 *   marx_mirror_reflect   -> marxb200_mirror_reflect        (photons stay in HBM)
 *   marx_grating_diffract -> marxb200_grating_diffract
 *   marx_detect           -> marxb200_detect
 *   marx_write_photons    -> marxb200_write_photons: the column files of the output directory are produced from the
 *                            device-resident event list (converted + byte-swapped on the GPU, one fwrite per file),
 *                            byte-identical to the stock writer's (marxio.c:403-476).  MARXB200_EGRESS=stock selects the
 *                            stock writer instead (survivors are first copied into pt->attributes).
 *   anything else that looks at pt->attributes (the pipe and rayfile writers, marx.c:191,220,1247) first gets the
 *   survivors via marxb200_download (Marx_Photon_Attr_Type == marxb200_photon_attr, 136 B).
 * Bookkeeping of Marx_Photon_Type follows source.c:268-384 (start_time, total_time, tag_start, history).
 *
 * Table upload: on the first marx_create_photons the post-init statics of the stock modules are serialised by the
 * calpack_*.c units (the field mappings of INTEGRATION.md section 2) and handed to marxb200_load_calpack.
 * Environment: MARXB200_DEVICE (CUDA ordinal, default 0), MARXB200_WARMUP (1: create the CUDA context in a helper thread at program load), MARXB200_EGRESS (bulk [default] | stock), MARXB200_WRITER_THREADS
 * (background column-file writers, default 8; 0 = synchronous).
 * Batch size: the reference caps dNumRays at 10^6 only through the range field of its parameter file (marx/par/marx.par:9); the
 * build writes integration/_build/par/marx.par, the same file with that maximum raised to 2^28, so that `marx_gpu @@.../marx.par
 * dNumRays=16777216` runs batches the size the GPU wants (2^24 rays: 4.2 GB of HBM, 2.3 GB of host photon buffer).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <pthread.h>
#include <marx.h>
#include <marxb200.h>

extern int _Marx_Dither_Mode;          /* marx/libsrc/_marx.h:144-145 (library-private header) */
#define DITHER_MODE_NONE 0
#include "calpack_io.h"

extern void __real_JDMsrandom (unsigned long);
extern int __real_marx_create_photons (Marx_Source_Type *, Marx_Photon_Type *, unsigned int, unsigned int *, double *);
extern int __real_marx_mirror_init (Param_File_Type *);
extern int __real_marx_grating_init (Param_File_Type *);
extern int __real_marx_detector_init (Param_File_Type *);
extern int __real_marx_write_photons (Marx_Photon_Type *, unsigned long, char *, int, double);
extern void __real_marx_prune_photons (Marx_Photon_Type *);
extern int __real_marx_dump_to_rayfile (char *, int, Marx_Photon_Type *, double);
extern int __real_marx_dealloc_photon_type (Marx_Photon_Type *);

static marxb200_ctx *Ctx;
static unsigned long Seed = 1;
static int Mirror_Id = -1, Grating_Id = -1, Detector_Id = -1;
static uint64_t Next_Ray;              /* 64-bit global ray index = RNG counter; low 32 bits = the reference's tag */
static int Host_Is_Stale;              /* photons of the current batch live in HBM only */
static int Have_Support_Orders;
static int Stock_Egress;               /* MARXB200_EGRESS=stock */
static int Source_Is_Rayfile;          /* SourceType=RAYFILE, USER, SAOSAC, SIMPUT: the stock host code produces the photons, the GPU traces them */
static int Bulk_Written;               /* the current batch went to the output directory straight from the device */

/* MARXB200_TIMING=1: host wall time spent in each wrapped call, printed when the driver frees its photon buffer */
static double T_Init, T_Create, T_Stages, T_Write, T_Sync, T_Init_Create, T_Init_Dump, T_Init_Upload, T_Init_Alloc;
static double now (void)
{
   struct timespec ts;
   clock_gettime (CLOCK_MONOTONIC, &ts);
   return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* MARXB200_WARMUP=1: the CUDA context is created by a helper thread from the moment the program is loaded, while marx.c still parses
 * its parameter file and the stock *_init calls read the calibration files.  Off by default: on the B200 boxes measured here the
 * stock initialisation takes ~0.1 s against 0.5 - 1.3 s of context creation (which varies that much from run to run), so the
 * overlap was not measurable (tools/driver_startup_probe.sh). */
static pthread_t Warm_Thread;
static int Warm_Started;
static void *warm_main (void *unused)
{
   const char *dev = getenv ("MARXB200_DEVICE");
   (void) unused;
   (void) marxb200_device_warmup (dev ? atoi (dev) : 0);
   return NULL;
}
__attribute__ ((constructor)) static void warm_start (void)
{
   const char *e = getenv ("MARXB200_WARMUP");
   if ((e == NULL) || (0 == atoi (e))) return;
   if (0 == pthread_create (&Warm_Thread, NULL, warm_main, NULL)) Warm_Started = 1;
}

static int gpu_error (const char *what)
{
   marx_error ("marxb200: %s: %s", what, marxb200_last_error ());
   return -1;
}

void __wrap_JDMsrandom (unsigned long seed)       /* marx.c:846-849: RandomSeed becomes the Philox key */
{
   Seed = seed;
   __real_JDMsrandom (seed);
}
int __wrap_marx_mirror_init (Param_File_Type *p) { return Mirror_Id = __real_marx_mirror_init (p); }
int __wrap_marx_grating_init (Param_File_Type *p) { return Grating_Id = __real_marx_grating_init (p); }
int __wrap_marx_detector_init (Param_File_Type *p) { return Detector_Id = __real_marx_detector_init (p); }

static int gpu_init (Marx_Source_Type *st, Marx_Photon_Type *pt)
{
   char path[] = "/tmp/marxb200_XXXXXX";
   mxcp_writer w;
   double meta[8];
   const char *dev = getenv ("MARXB200_DEVICE");
   int fd, status = -1;

   if ((Mirror_Id != MARX_MIRROR_HRMA) && (Mirror_Id != MARX_MIRROR_FFIELD))
     { marx_error ("marxb200: MirrorType must be HRMA or FLATFIELD for the GPU path"); return -1; }
   if ((Grating_Id != 0) && (Grating_Id != MARX_GRATING_HETG) && (Grating_Id != MARX_GRATING_LETG))
     { marx_error ("marxb200: GratingType must be NONE, HETG or LETG for the GPU path"); return -1; }
   if ((Detector_Id != 0) && (Detector_Id != MARX_DETECTOR_ACIS_S) && (Detector_Id != MARX_DETECTOR_ACIS_I)
       && (Detector_Id != MARX_DETECTOR_HRC_S) && (Detector_Id != MARX_DETECTOR_HRC_I))
     { marx_error ("marxb200: DetectorType must be NONE, ACIS-S, ACIS-I, HRC-S or HRC-I for the GPU path"); return -1; }

   double t_mark = now ();
   if (Warm_Started) { (void) pthread_join (Warm_Thread, NULL); Warm_Started = 0; }
   if (-1 == marxb200_create (&Ctx, dev ? atoi (dev) : 0, (uint64_t) Seed))
     return gpu_error ("marxb200_create");
   T_Init_Create = now () - t_mark; t_mark = now ();

   if (-1 == (fd = mkstemp (path)))
     { marx_error ("marxb200: cannot create a temporary file"); return -1; }
   close (fd);
   if (-1 == mxcp_open_write (&w, path))
     { marx_error ("marxb200: cannot write %s", path); return -1; }
   memset (meta, 0, sizeof (meta));
   meta[0] = Mirror_Id; meta[1] = Grating_Id; meta[2] = Detector_Id; meta[5] = (double) Seed;
   CP_F64 (&w, "meta", meta, 8);
   if (-1 == (Source_Is_Rayfile = calpack_dump_source (&w, st)))
     marx_error ("marxb200: this SourceType has neither a device kernel nor a stock host generator");
   else if ((-1 == calpack_dump_dither (&w))
	    || (-1 == ((Mirror_Id == MARX_MIRROR_FFIELD) ? calpack_dump_ffield (&w) : calpack_dump_hrma (&w)))
	    || (-1 == calpack_dump_grating (&w, Grating_Id))
	    || (-1 == ((Detector_Id == MARX_DETECTOR_HRC_S) ? calpack_dump_hrc_s (&w, Detector_Id)
		       : (Detector_Id == MARX_DETECTOR_HRC_I) ? calpack_dump_hrc_i (&w, Detector_Id)
		       : calpack_dump_acis_s (&w, Detector_Id))))
     marx_error ("marxb200: could not serialise the module tables");
   else status = 0;
   mxcp_close_write (&w);
   T_Init_Dump = now () - t_mark; t_mark = now ();
   if ((status == 0) && (-1 == marxb200_load_calpack (Ctx, path)))
     status = gpu_error ("table upload");
   (void) unlink (path);
   if (status == -1) return -1;
   T_Init_Upload = now () - t_mark; t_mark = now ();

   if (-1 == marxb200_alloc_photons (Ctx, pt->max_n_photons))
     return gpu_error ("marxb200_alloc_photons");
   T_Init_Alloc = now () - t_mark;

   {
      /* column files are appended by background threads while the next batch is traced (marxb200_set_async_writer);
       * MARXB200_WRITER_THREADS=0 writes synchronously */
      const char *wt = getenv ("MARXB200_WRITER_THREADS");
      int n_threads = (wt != NULL) ? atoi (wt) : 8;
      if ((n_threads > 0) && (-1 == marxb200_set_async_writer (Ctx, n_threads)))
	return gpu_error ("marxb200_set_async_writer");
   }
   Have_Support_Orders = (Grating_Id == MARX_GRATING_LETG);
   Stock_Egress = ((NULL != getenv ("MARXB200_EGRESS")) && (0 == strcmp (getenv ("MARXB200_EGRESS"), "stock")));
   marx_message ("marxb200: ray trace on CUDA device %d, Philox key %lu\n", dev ? atoi (dev) : 0, Seed);
   return 0;
}

/* survivors of the current batch -> pt->attributes, arrival order (marxio.c:422-435) */
static int sync_host (Marx_Photon_Type *pt)
{
   uint64_t n_live = 0, i;
   unsigned int *idx;

   double t0 = now ();
   if (Host_Is_Stale == 0) return 0;
   if (-1 == marxb200_download (Ctx, (marxb200_photon_attr *) pt->attributes, pt->max_n_photons, &n_live))
     return gpu_error ("marxb200_download");

   if (pt->sorted_index != NULL) JDMfree_integer_vector ((int *) pt->sorted_index);
   if (NULL == (idx = (unsigned int *) JDMinteger_vector ((unsigned int) (n_live ? n_live : 1))))
     return -1;
   for (i = 0; i < n_live; i++)
     {
	idx[i] = (unsigned int) i;
	pt->sorted_energies[i] = pt->attributes[i].energy;
     }
   pt->sorted_index = idx;
   pt->n_photons = pt->num_sorted = (unsigned int) n_live;
   Host_Is_Stale = 0;
   T_Sync += now () - t0;
   return 0;
}

int __wrap_marx_create_photons (Marx_Source_Type *st, Marx_Photon_Type *pt, unsigned int num,
				unsigned int *num_collected, double *exposure_time)
{
   uint64_t n = num;
   double t_end;

   *num_collected = 0;
   if ((st == NULL) || (pt == NULL)) return -1;
   if (Ctx == NULL)
     {
	double t0 = now ();
	if (-1 == gpu_init (st, pt)) return -1;
	T_Init += now () - t0;
     }
   if (num > pt->max_n_photons)
     { marx_error ("marxb200: batch of %u rays exceeds the photon buffer", num); return -1; }

   if (Source_Is_Rayfile)
     {
	/* host-generated sources (USER s-user.c:189-236, SAOSAC s-saosac.c, SIMPUT s-simput.c) go the same way as the
	 * RAYFILE re-entry (s-rayfile.c:188-221): the stock host code reads the records, fixes up their times, tags and
	 * (record-only) dither state and sets pt->history from the file; the list is then injected into HBM and the
	 * stage wrappers skip what the history says was already done (hrma.c:1171, diffract.c:982, acis-s.c:186). */
	unsigned int got = 0;
	double t0 = now ();
	unsigned int live = 0, i;
	if (-1 == __real_marx_create_photons (st, pt, num, &got, exposure_time)) return -1;
	/* a host generator may hand over rays it has already rejected (SAOSAC: the ray-weight test of s-saosac.c:206-211 sets
	 * PHOTON_MIRROR_VBLOCKED); the stock stages skip them through marx_prune_photons (photon.c:40-63), the device list holds live
	 * rays only: drop them here, in place and in order (the records are overwritten by the next batch anyway) */
	for (i = 0; i < got; i++)
	  if (0 == (pt->attributes[i].flags & BAD_PHOTON_MASK))
	    {
	       if (live != i) pt->attributes[live] = pt->attributes[i];
	       live++;
	    }
	if (-1 == marxb200_upload_from (Ctx, (marxb200_photon_attr *) pt->attributes, live, NULL, pt->start_time))
	  return gpu_error ("marxb200_upload_from");
	*num_collected = got;
	Next_Ray += got;
	Host_Is_Stale = (got != 0);
	Bulk_Written = 0;
	T_Create += now () - t0;
	return 0;
     }

   pt->source_distance = st->distance;                    /* source.c:282-285 */
   pt->history = 0;
   pt->start_time += pt->total_time;

   double t0 = now ();
   /* time base < 0: the running sum of arrival times continues on the device (source.c:285,326) */
   if (-1 == marxb200_create_photons (Ctx, Next_Ray, n, -1.0))
     return gpu_error ("marxb200_create_photons");
   if ((exposure_time != NULL)                            /* source.c:323-334 */
       && (-1 == marxb200_truncate_exposure (Ctx, *exposure_time, &n)))
     return gpu_error ("marxb200_truncate_exposure");
   {
      /* DitherModel=FILE: the batch ends with the last ray the ASPSOL file still brackets (dither.c:296-301, 630-658;
       * source.c:355 passes the shortened count on), and marx.c:577-606 ends the run on a short or empty batch */
      uint64_t generated = n;
      if (-1 == marxb200_get_counts (Ctx, &generated, NULL, &t_end))
	return gpu_error ("marxb200_get_counts");
      if (generated < n) n = generated;
   }

   pt->history |= (MARX_ENERGY_OK | MARX_TIME_OK | MARX_X_VECTOR_OK | MARX_P_VECTOR_OK | MARX_TAG_OK);
   pt->tag_start += (unsigned int) n;                     /* source.c:346-355 */
   Next_Ray += n;
   pt->num_sorted = pt->n_photons = (unsigned int) n;
   pt->total_time = (n == 0) ? 0.0 : (t_end - pt->start_time);   /* source.c:377-381 */
   *num_collected = (unsigned int) n;
   Host_Is_Stale = 1;
   Bulk_Written = 0;
   T_Create += now () - t0;
   return 0;
}

int __wrap_marx_mirror_reflect (Marx_Photon_Type *pt, int verbose)
{
   if (pt->history & MARX_MIRROR_SHELL_OK) return 0;      /* hrma.c:1171-1173 */
   pt->history |= MARX_MIRROR_SHELL_OK;
   if (verbose > 0) marx_message ("Reflecting from HRMA [B200]\n");
   { double t0 = now (); if (-1 == marxb200_mirror_reflect (Ctx)) return gpu_error ("marxb200_mirror_reflect"); T_Stages += now () - t0; }
   return 0;
}

int __wrap_marx_grating_diffract (Marx_Photon_Type *pt, int verbose)
{
   if (Grating_Id == 0) return 0;                         /* grating.c:87-: no grating */
   if (pt->history & MARX_ORDER_OK) return 0;             /* diffract.c:982-984 */
   pt->history |= MARX_ORDER_OK;
   if (Have_Support_Orders)                               /* diffract.c:1098-1118 */
     pt->history |= (MARX_ORDER1_OK | MARX_ORDER2_OK | MARX_ORDER3_OK | MARX_ORDER4_OK);
   if (verbose > 0) marx_message ("Diffracting from %s [B200]\n", (Grating_Id == MARX_GRATING_LETG) ? "LETG" : "HETG");
   { double t0 = now (); if (-1 == marxb200_grating_diffract (Ctx)) return gpu_error ("marxb200_grating_diffract"); T_Stages += now () - t0; }
   return 0;
}

int __wrap_marx_detect (Marx_Photon_Type *pt, int verbose)
{
   if ((Detector_Id != 0) && (0 == (pt->history & MARX_DET_NUM_OK)))   /* acis-s.c:186-190, hrc-s.c:242-248 */
     {
	if (Detector_Id == MARX_DETECTOR_HRC_S)
	  pt->history |= (MARX_DET_REGION_OK | MARX_PULSEHEIGHT_OK | MARX_DET_PIXEL_OK | MARX_DET_NUM_OK | MARX_DET_UV_PIXEL_OK);
	else if (Detector_Id == MARX_DETECTOR_HRC_I)          /* hrc-i.c:125-129 */
	  pt->history |= (MARX_DET_REGION_OK | MARX_PULSEHEIGHT_OK | MARX_DET_PIXEL_OK | MARX_DET_NUM_OK);
	else
	  pt->history |= (MARX_DET_PIXEL_OK | MARX_DET_NUM_OK | MARX_PULSEHEIGHT_OK | MARX_PI_OK);
	if (verbose > 0) marx_message ("Detecting [B200]\n");
	{ double t0 = now (); if (-1 == marxb200_detect (Ctx)) return gpu_error ("marxb200_detect"); T_Stages += now () - t0; }
     }
   return 0;
}

int __wrap_marx_write_photons (Marx_Photon_Type *pt, unsigned long write_mask, char *dir, int open_mode, double total_time)
{
   if ((Ctx == NULL) || Stock_Egress || (Host_Is_Stale == 0))
     {
	if ((Ctx != NULL) && (-1 == sync_host (pt))) return -1;
	return __real_marx_write_photons (pt, write_mask, dir, open_mode, total_time);
     }
   /* marxio.c:409-414 */
   if (_Marx_Dither_Mode != DITHER_MODE_NONE)
     write_mask &= (pt->history | MARX_SKY_DITHER_OK | MARX_DET_DITHER_OK);
   else
     write_mask &= pt->history;
   {
      double t0 = now ();
      if (-1 == marxb200_write_photons (Ctx, dir, (uint64_t) write_mask, open_mode, total_time))
	return gpu_error ("marxb200_write_photons");
      T_Write += now () - t0;
   }
   Bulk_Written = 1;
   return 0;
}

void __wrap_marx_prune_photons (Marx_Photon_Type *pt)
{
   if ((Ctx != NULL) && Host_Is_Stale)
     {
	uint64_t n_live = 0;
	if (Bulk_Written && (0 == marxb200_get_counts (Ctx, NULL, &n_live, NULL)))
	  {
	     /* marx.c:593 only needs the number of survivors; they were written from the device */
	     pt->num_sorted = (unsigned int) n_live;
	     return;
	  }
	if (-1 == sync_host (pt)) return;
     }
   __real_marx_prune_photons (pt);
}

int __wrap_marx_dump_to_rayfile (char *file, int new_file, Marx_Photon_Type *pt, double total_time)
{
   if (-1 == sync_host (pt)) return -1;                   /* DumpToRayFile=yes skips process_photons (marx.c:582) */
   return __real_marx_dump_to_rayfile (file, new_file, pt, total_time);
}

int __wrap_marx_dealloc_photon_type (Marx_Photon_Type *pt)
{
   if (Ctx != NULL)
     {
	uint64_t stage[4];
	if (0 == marxb200_get_stage_counts (Ctx, stage))
	  marx_message ("marxb200: last batch: %lu generated, %lu reflected, %lu diffracted, %lu detected\n",
			(unsigned long) stage[0], (unsigned long) stage[1], (unsigned long) stage[2], (unsigned long) stage[3]);
	if (getenv ("MARXB200_TIMING") != NULL)
	  fprintf (stderr, "marxb200: host seconds in the wrapped calls: init+upload %.3f (CUDA context %.3f, table dump %.3f, table upload %.3f, "
		   "photon buffers %.3f), create_photons %.3f, stages %.3f, write_photons %.3f, download %.3f\n", T_Init, T_Init_Create,
		   T_Init_Dump, T_Init_Upload, T_Init_Alloc, T_Create, T_Stages, T_Write, T_Sync);
	/* every queued column append must be in its file before the driver closes the run (marx.c:616-618); a failed write
	 * fails the run like a failed fwrite of the stock writer (marxio.c:452-466) */
	{
	   int flushed = marxb200_write_flush (Ctx);
	   if (flushed == -1) (void) gpu_error ("marxb200_write_flush");
	   (void) marxb200_destroy (Ctx);
	   Ctx = NULL;
	   if (flushed == -1)
	     {
		(void) __real_marx_dealloc_photon_type (pt);
		return -1;
	     }
	}
     }
   return __real_marx_dealloc_photon_type (pt);
}
