/* marx2fits_gpu.c -- reference-side binding of the Level-1 transforms (marxb200_level1_*) into the UNMODIFIED marx2fits.
 *
 * marx/src/marx2fits.c is compiled into this unit where it lies under $(REF) (its per-event functions and tables are file
 * statics; `main` is renamed).  Everything host-side stays the reference's: option and parameter handling
 * (get_simulation_info, read_obspar_file), the column-file readers, the FITS header, the binary-table writer
 * (write_table_values, jdfits), the GTI extensions and the marx.par appendix.  What changes is who computes the per-event
 * columns.  The stock loop (marx2fits(), :2851-2862) calls, for every row, the ddt_compute_value hook of every column of
 * Data_Def_Table (:274-1252, compute_table_values :1448).  Here the hooks of the COMPUTED columns -- compute_expno, _tdetxy,
 * _fltgrade, _grade, _detxy, _xy_sky, _acis_energy, _pi, _node_id, _status -- are re-pointed at functions that copy row i of
 * the columns the GPU produced in one marxb200_level1_transform over the whole event list; the hooks that READ the column
 * files stay stock.  The stock marx2fits() is then called unchanged.
 *
 * There is no CPU fallback: without a CUDA device marxb200_create fails and the program stops.
 *
 *   usage: marx2fits_gpu [--pixadj=edser|none|randomize|exact] marxdir outfile
 *   environment: MARXB200_DEVICE (CUDA ordinal), L1_SEED (key of the per-row draw stream; default 1)
 */
#include "acis.h"        /* MARX_DET_FACET_PRIVATE_DATA: the tdet offsets of a facet (acis.h:38-40, same two floats in hrc.h) */
#define main marx2fits_stock_main
#include "marx2fits.c"
#undef main
/* Marx_Subpix_Table_Type is opaque outside acis_subpix.c: compile that unit in as well (its definitions then satisfy the
 * linker instead of the archive member) so that the EDSER tables can be handed to marxb200_set_level1 */
#include "acis_subpix.c"

#include <marxb200.h>

static marxb200_ctx *Ctx;
static uint64_t N_Rows, Row;
static marxb200_level1_columns L1;

#define CHECK(call) do { if (-1 == (call)) { fprintf (stderr, "marx2fits_gpu: %s\n", marxb200_last_error ()); return 1; } } while (0)

/* ---- the replacement hooks: row `Row` of the device-computed columns -> Data_Table ---- */
static int gpu_expno (Data_Def_Type *ddt)      /* first computed column of an ACIS row (Data_Def_Table order) */
{
   (void) ddt;
   Data_Table.dtt_expno = L1.expno[Row];
   return 0;
}
static int gpu_tdetxy (Data_Def_Type *ddt)
{
   (void) ddt;
   Data_Table.dtt_tdetx = L1.tdetx[Row]; Data_Table.dtt_tdety = L1.tdety[Row];
   return 0;
}
static int gpu_fltgrade (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_fltgrade = L1.fltgrade[Row]; return 0; }
static int gpu_grade (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_grade = L1.grade[Row]; return 0; }
static int gpu_detxy (Data_Def_Type *ddt)
{
   (void) ddt;
   Data_Table.dtt_detx = L1.detx[Row]; Data_Table.dtt_dety = L1.dety[Row];
   return 0;
}
static int gpu_xy_sky (Data_Def_Type *ddt)
{
   (void) ddt;
   Data_Table.dtt_xsky = L1.x[Row]; Data_Table.dtt_ysky = L1.y[Row];
   return 0;
}
static int gpu_energy (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_energy = L1.energy[Row]; return 0; }
static int gpu_pi (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_pi = L1.pi[Row]; return 0; }
static int gpu_node_id (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_node_id = L1.node_id[Row]; return 0; }
/* STATUS is the last column of Data_Def_Table: the row is complete */
static int gpu_status (Data_Def_Type *ddt) { (void) ddt; Data_Table.dtt_status = L1.status[Row]; Row++; return 0; }

static void repoint_hooks (void)
{
   Data_Def_Type *ddt;
   for (ddt = Data_Def_Table; ddt->ddt_value_ptr != NULL; ddt++)
     {
        if (ddt->ddt_compute_value == compute_expno) ddt->ddt_compute_value = gpu_expno;
        else if (ddt->ddt_compute_value == compute_tdetxy) ddt->ddt_compute_value = gpu_tdetxy;
        else if (ddt->ddt_compute_value == compute_fltgrade) ddt->ddt_compute_value = gpu_fltgrade;
        else if (ddt->ddt_compute_value == compute_grade) ddt->ddt_compute_value = gpu_grade;
        else if (ddt->ddt_compute_value == compute_detxy) ddt->ddt_compute_value = gpu_detxy;
        else if (ddt->ddt_compute_value == compute_xy_sky) ddt->ddt_compute_value = gpu_xy_sky;
        else if (ddt->ddt_compute_value == compute_acis_energy) ddt->ddt_compute_value = gpu_energy;
        else if (ddt->ddt_compute_value == compute_pi) ddt->ddt_compute_value = gpu_pi;
        else if (ddt->ddt_compute_value == compute_node_id) ddt->ddt_compute_value = gpu_node_id;
        else if (ddt->ddt_compute_value == compute_status) ddt->ddt_compute_value = gpu_status;
     }
}

/* ---- the descriptor: what the stock initialisation left in marx2fits' statics ---- */
static int32_t Subpix_Npoints[2 * 256];
static uint32_t Subpix_Offset[2 * 256];
static float *Subpix_Data;

static int fill_descriptor (marxb200_level1_desc *d)
{
   Marx_Detector_Geometry_Type *g;
   int k = 0;
   memset (d, 0, sizeof (*d));
   d->detector_type = The_Detector->detector_type;
   for (g = The_Detector->facet_list; g != NULL; g = g->next)
     {
        marxb200_level1_chip *c;
        if (k == MARXB200_L1_MAX_CHIPS) return -1;
        c = d->chips + k++;
        c->id = g->id;
        c->subpix_table = ((g->id == 5) || (g->id == 7));            /* acis_subpix.c:262-268 */
        c->x_ll[0] = g->x_ll.x; c->x_ll[1] = g->x_ll.y; c->x_ll[2] = g->x_ll.z;
        c->xhat[0] = g->xhat.x; c->xhat[1] = g->xhat.y; c->xhat[2] = g->xhat.z;
        c->yhat[0] = g->yhat.x; c->yhat[1] = g->yhat.y; c->yhat[2] = g->yhat.z;
        c->x_pixel_size = g->x_pixel_size; c->y_pixel_size = g->y_pixel_size;
        c->xpixel_offset = g->xpixel_offset; c->ypixel_offset = g->ypixel_offset;
        c->tdet_xoff = g->tdet_xoff; c->tdet_yoff = g->tdet_yoff;
     }
   d->num_chips = k;
   d->fp_delta_s0 = The_Detector->fp_coord_info->fp_delta_s0;
   d->fp_x0 = The_Detector->fp_coord_info->fp_x0; d->fp_y0 = The_Detector->fp_coord_info->fp_y0;
   d->focal_length = Focal_Length;
   d->det_offset[0] = DetOffset_X; d->det_offset[1] = DetOffset_Y; d->det_offset[2] = DetOffset_Z;
   d->time_del = TimeDel; d->time_start = Time_Start;
   d->pi_factor = Acis_PI_Factor; d->nominal_roll = Nominal_Roll;
   d->used_dither = Simulation_Used_Dither; d->pix_adjust = Pixel_Adjust;
   if (Acis_Subpixel_Object != NULL)
     {
        /* Subpix_CCD_Type (acis_subpix.c:41-60): fi / bi tables, one Subpix_Type per flight grade */
        Subpix_CCD_Type *tab[2];
        size_t total = 0, pos = 0;
        int t, f;
        tab[0] = Acis_Subpixel_Object->fi; tab[1] = Acis_Subpixel_Object->bi;
        for (t = 0; t < 2; t++) for (f = 0; f < 256; f++) if (tab[t]->s[f] != NULL) total += 3 * (size_t) tab[t]->s[f]->num_energies;
        if (NULL == (Subpix_Data = (float *) malloc ((total + 1) * sizeof (float)))) return -1;
        for (t = 0; t < 2; t++)
          for (f = 0; f < 256; f++)
            {
               Subpix_Type *s = tab[t]->s[f];
               if (s == NULL) continue;
               Subpix_Npoints[t * 256 + f] = s->num_energies;
               Subpix_Offset[t * 256 + f] = (uint32_t) pos;
               memcpy (Subpix_Data + pos, s->energies, 3 * (size_t) s->num_energies * sizeof (float));   /* energies, dxs, dys are contiguous (:64-82) */
               pos += 3 * (size_t) s->num_energies;
            }
        d->subpix_npoints = Subpix_Npoints; d->subpix_offset = Subpix_Offset; d->subpix_data = Subpix_Data;
        d->subpix_data_len = total;
     }
   return 0;
}

/* ---- the event list: the column files of the output directory -> device (through the reference's own readers) ---- */
static int read_f32_column (char *name, float *dst, uint64_t n, int required)
{
   Marx_Dump_File_Type *dft;
   char *file = make_marx_filename (name);
   if (1 != marx_file_exists (file)) return required ? -1 : 0;
   if (NULL == (dft = marx_open_read_dump_file (file))) return -1;
   if (((uint64_t) dft->num_rows != n) || (n != JDMread_float32 (dst, (unsigned int) n, dft->fp))) { marx_close_read_dump_file (dft); return -1; }
   marx_close_read_dump_file (dft);
   return 1;
}

static int upload_event_list (void)
{
   marxb200_photon_attr *ph;
   float *tmp;
   uint64_t i, n = N_Rows;
   static char *dither_files[6] = {"sky_ra.dat", "sky_dec.dat", "sky_roll.dat", "det_dy.dat", "det_dz.dat", "det_theta.dat"};
   int k;
   if ((NULL == (ph = (marxb200_photon_attr *) calloc (n + 1, sizeof (*ph)))) || (NULL == (tmp = (float *) malloc ((n + 1) * sizeof (float))))) return -1;
   if (1 != read_f32_column ("time.dat", tmp, n, 1)) return -1;
   for (i = 0; i < n; i++) ph[i].arrival_time = (double) tmp[i];
   if (1 != read_f32_column ("xpixel.dat", tmp, n, 1)) return -1;
   for (i = 0; i < n; i++) ph[i].y_pixel = tmp[i];
   if (1 != read_f32_column ("ypixel.dat", tmp, n, 1)) return -1;
   for (i = 0; i < n; i++) ph[i].z_pixel = tmp[i];
   if (-1 == (k = read_f32_column ("b_energy.dat", tmp, n, 0))) return -1;
   if (k) for (i = 0; i < n; i++) ph[i].pi = tmp[i];
   if (-1 == (k = read_f32_column ("hrc_u.dat", tmp, n, 0))) return -1;
   if (k) for (i = 0; i < n; i++) ph[i].u_pixel = tmp[i];
   if (-1 == (k = read_f32_column ("hrc_v.dat", tmp, n, 0))) return -1;
   if (k) for (i = 0; i < n; i++) ph[i].v_pixel = tmp[i];
   if (Simulation_Used_Dither)
     for (k = 0; k < 6; k++)
       {
          int got = read_f32_column (dither_files[k], tmp, n, 0);
          if (got == -1) return -1;
          if (got == 0) continue;
          for (i = 0; i < n; i++)
            switch (k)
              {
               case 0: ph[i].dither_ra = tmp[i]; break;
               case 1: ph[i].dither_dec = tmp[i]; break;
               case 2: ph[i].dither_roll = tmp[i]; break;
               case 3: ph[i].dither_dy = tmp[i]; break;
               case 4: ph[i].dither_dz = tmp[i]; break;
               default: ph[i].dither_dtheta = tmp[i]; break;
              }
       }
   {
      Marx_Dump_File_Type *dft;
      int16 *pha = (int16 *) tmp;
      if (NULL == (dft = marx_open_read_dump_file (make_marx_filename ("pha.dat")))) return -1;
      if (n != JDMread_int16 (pha, (unsigned int) n, dft->fp)) return -1;
      marx_close_read_dump_file (dft);
      for (i = 0; i < n; i++) ph[i].pulse_height = pha[i];
      if (NULL == (dft = marx_open_read_dump_file (make_marx_filename ("detector.dat")))) return -1;
      if (n != fread (tmp, 1, n, dft->fp)) return -1;
      marx_close_read_dump_file (dft);
      for (i = 0; i < n; i++) ph[i].ccd_num = ((signed char *) tmp)[i];
   }
   for (i = 0; i < n; i++) ph[i].tag = (uint32_t) i;
   if (-1 == marxb200_alloc_photons (Ctx, n + 1024)) return -1;
   if (-1 == marxb200_upload_from (Ctx, ph, n, NULL, 0.0)) return -1;
   free (ph); free (tmp);
   return 0;
}

static int alloc_columns (uint64_t n)
{
#define COL(f, T) if (NULL == (L1.f = (T *) calloc (n + 1, sizeof (T)))) return -1
   COL (time, double); COL (detx, double); COL (dety, double); COL (x, double); COL (y, double);
   COL (expno, int32_t); COL (tdetx, int32_t); COL (tdety, int32_t);
   COL (energy, float);
   COL (node_id, int16_t); COL (pi, int16_t); COL (fltgrade, int16_t); COL (grade, int16_t); COL (status, int16_t);
#undef COL
   return 0;
}

int main (int argc, char **argv)
{
   JDFits_Type *ft;
   marxb200_level1_desc desc;
   char *fits_file;
   const char *env;
   uint64_t got = 0;
   int i;

   /* options of the stock main (:3225-3283) that concern the per-event transforms */
   for (i = 1; i < argc - 2; i++)
     {
        char *arg = argv[i];
        if (0 != strncmp (arg, "--pixadj=", 9)) { fprintf (stderr, "marx2fits_gpu: unsupported option %s\n", arg); return 1; }
        arg += 9;
        if ((0 == strcmp (arg, "none")) || (0 == strcmp (arg, "NONE"))) { Pixel_Adjust = PIX_ADJ_NONE; Pix_Adj = "NONE"; }
        else if ((0 == strcmp (arg, "randomize")) || (0 == strcmp (arg, "RANDOMIZE"))) { Pixel_Adjust = PIX_ADJ_RANDOMIZE; Pix_Adj = "RANDOMIZE"; Rand_Sky = 0.5; }
        else if ((0 == strcmp (arg, "exact")) || (0 == strcmp (arg, "EXACT"))) { Pixel_Adjust = PIX_ADJ_EXACT; Pix_Adj = "EXACT"; }
        else if ((0 == strcmp (arg, "edser")) || (0 == strcmp (arg, "EDSER"))) { Pixel_Adjust = PIX_ADJ_EDSER; Pix_Adj = "EDSER"; }
        else { fprintf (stderr, "marx2fits_gpu: unsupported --pixadj option: %s\n", arg); return 1; }
     }
   if (argc < 3) { fprintf (stderr, "usage: marx2fits_gpu [--pixadj=edser|none|randomize|exact] marxdir outfile\n"); return 1; }
   Marx_Dir = argv[argc - 2];
   fits_file = argv[argc - 1];
   sprintf (Marx2fits_Pgm, "marx2fits_gpu v%s", MARX_VERSION_STRING);

   /* the stock initialisation, in the order of the stock main (:3285-3330) */
   if (-1 == get_simulation_info ()) return 1;
   if (Simulation_Used_ACIS)
     {
        if ((Pixel_Adjust == PIX_ADJ_EDSER) && (NULL == (Acis_Subpixel_Object = marx_open_acis_subpix ())))
          { fprintf (stderr, "Error opening the subpixel file\n"); return 1; }
     }
   else if (Pixel_Adjust == PIX_ADJ_EDSER) Pixel_Adjust = PIX_ADJ_RANDOMIZE;
   Obs_Par_Parms = read_obspar_file ();

   /* rows of the column files (marx2fits() walks Num_Marx_File_Rows of them; Num_Marx_Data_Values excludes pha == -1) */
   {
      Marx_Dump_File_Type *dft = marx_open_read_dump_file (make_marx_filename ("time.dat"));
      if (dft == NULL) return 1;
      N_Rows = (uint64_t) dft->num_rows;
      marx_close_read_dump_file (dft);
   }

   /* the device side: one context, the descriptor, the event list, ONE transform over the whole file */
   {
      int device = (NULL != (env = getenv ("MARXB200_DEVICE"))) ? atoi (env) : 0;
      uint64_t seed = (NULL != (env = getenv ("L1_SEED"))) ? strtoull (env, NULL, 10) : 1;
      CHECK (marxb200_create (&Ctx, device, seed));
   }
   if (-1 == fill_descriptor (&desc)) { fprintf (stderr, "marx2fits_gpu: cannot build the Level-1 descriptor\n"); return 1; }
   CHECK (marxb200_set_level1 (Ctx, &desc));
   if (N_Rows > 0)
     {
        if (-1 == upload_event_list ()) { fprintf (stderr, "marx2fits_gpu: cannot load the event list: %s\n", marxb200_last_error ()); return 1; }
        if (-1 == alloc_columns (N_Rows)) return 1;
        CHECK (marxb200_level1_transform (Ctx, 0.0));
        CHECK (marxb200_level1_download (Ctx, &L1, N_Rows, &got));
        if (got != N_Rows) { fprintf (stderr, "marx2fits_gpu: %llu of %llu rows transformed\n", (unsigned long long) got, (unsigned long long) N_Rows); return 1; }
     }
   marxb200_destroy (Ctx);

   /* the stock writer with the computed columns re-pointed at the device results */
   repoint_hooks ();
   if (-1 == open_data_def_table ()) return 1;
   if (-1 == init_data_def_write_table ()) { (void) close_data_def_table (); return 1; }
   if (NULL == (ft = jdfits_open_file (fits_file, JDFITS_WRITE_MODE))) { marx_error ("*** Unable to open output file %s\n", fits_file); return 1; }
   if ((-1 == init_marx_fits_file (ft)) || (-1 == marx2fits (ft)) || (-1 == add_goodtime_extensions (ft)) || (-1 == add_marx_par_to_file (ft)))
     { (void) jdfits_close_file (ft); (void) close_data_def_table (); return 1; }
   if (-1 == jdfits_close_file (ft)) return 1;
   if (-1 == close_data_def_table ()) return 1;
   if (Row != N_Rows) { fprintf (stderr, "marx2fits_gpu: wrote %llu of %llu rows\n", (unsigned long long) Row, (unsigned long long) N_Rows); return 1; }
   return 0;
}
