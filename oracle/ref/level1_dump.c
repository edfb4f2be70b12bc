/* level1_dump -- TEST INFRASTRUCTURE (oracle/_ref build only).
 * Runs the stock marx2fits initialisation (marx/src/marx2fits.c: main :3200-3310, get_simulation_info :2873,
 * get_marx_pfile_info :2371-2562) on a MARX output directory and prints, as "key value..." lines, every quantity its
 * per-event transforms (compute_expno/tdetxy/detxy/xy_sky/pi/..., :3584-3943) read afterwards: this is the Level-1
 * descriptor (include/marxb200.h: marxb200_level1_desc) as the reference itself derives it.  The reference source is
 * compiled into this unit where it lies (its statics are not reachable otherwise); nothing is copied.
 * usage: level1_dump [--pixadj=edser|none|randomize|exact] marxdir */
#include "acis.h"   /* MARX_DET_FACET_PRIVATE_DATA: the tdet offsets of a facet (acis.h:38-40, same two floats in hrc.h) */
#define main marx2fits_stock_main
#include "marx2fits.c"
#undef main

int main (int argc, char **argv)
{
   Marx_Detector_Geometry_Type *g;
   int i;
   for (i = 1; i < argc - 1; i++)
     {
        char *arg = argv[i];
        if (0 == strncmp (arg, "--pixadj=", 9))
          {
             arg += 9;
             if (0 == strcmp (arg, "none")) Pixel_Adjust = PIX_ADJ_NONE;
             else if (0 == strcmp (arg, "randomize")) Pixel_Adjust = PIX_ADJ_RANDOMIZE;
             else if (0 == strcmp (arg, "exact")) Pixel_Adjust = PIX_ADJ_EXACT;
             else Pixel_Adjust = PIX_ADJ_EDSER;
          }
     }
   if (argc < 2) return 2;
   Marx_Dir = argv[argc - 1];
   if (-1 == get_simulation_info ()) return 1;
   if ((0 == Simulation_Used_ACIS) && (Pixel_Adjust == PIX_ADJ_EDSER)) Pixel_Adjust = PIX_ADJ_RANDOMIZE;   /* main :3308-3309 */

   Obs_Par_Parms = read_obspar_file ();            /* Nominal_Roll, Time_Start (:2947-2985) */

   printf ("detector %s\n", DetectorType);
   printf ("detector_type %d\n", The_Detector->detector_type);
   printf ("used_acis %d\n", Simulation_Used_ACIS ? 1 : 0);
   printf ("used_dither %d\n", Simulation_Used_Dither);
   printf ("pix_adjust %d\n", Pixel_Adjust);
   printf ("time_del %.17g\n", TimeDel);
   printf ("time_start %.17g\n", Time_Start);
   printf ("pi_factor %.17g\n", Acis_PI_Factor);
   printf ("focal_length %.17g\n", Focal_Length);
   printf ("det_offset %.17g %.17g %.17g\n", DetOffset_X, DetOffset_Y, DetOffset_Z);
   printf ("nominal_roll %.17g\n", Nominal_Roll);
   printf ("fp %.17g %.17g %.17g\n", The_Detector->fp_coord_info->fp_delta_s0, The_Detector->fp_coord_info->fp_x0,
           The_Detector->fp_coord_info->fp_y0);
   printf ("facet_ids %d %d\n", The_Detector->first_facet_id, The_Detector->last_facet_id);
   printf ("grade_map");
   for (i = 0; i < 256; i++) printf (" %d", (int) Grade_Map[i]);
   printf ("\n");
   if (Simulation_Used_ACIS) printf ("subpix_file %s\n", Subpix_File);
   for (g = The_Detector->facet_list; g != NULL; g = g->next)
     printf ("chip %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", g->id,
             g->x_ll.x, g->x_ll.y, g->x_ll.z, g->xhat.x, g->xhat.y, g->xhat.z, g->yhat.x, g->yhat.y, g->yhat.z,
             g->x_pixel_size, g->y_pixel_size, g->xpixel_offset, g->ypixel_offset, (double) g->tdet_xoff, (double) g->tdet_yoff);
   return 0;
}
