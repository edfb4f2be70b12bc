/* asp_dump -- TEST INFRASTRUCTURE (oracle/_ref build only).
 * Runs the stock marxasp initialisation (marx/src/marxasp.c: marxasp_init :1076-1106, get_simulation_info :549-577,
 * setup_dither :387-410) on a MARX output directory and prints, as "key value..." lines, every quantity the row loop of
 * write_marxasp (:996-1027: compute_dither :814-884, compute_quaternion :886-903) reads afterwards: this is the aspect-solution
 * descriptor (include/marxb200.h: marxb200_aspsol_desc) as the reference itself derives it.  The reference source is compiled
 * into this unit where it lies (its statics are not reachable otherwise); nothing is copied.
 * usage: asp_dump @@marxasp.par MarxDir=DIR [TimeDel=...] */
#define main marxasp_stock_main
#include "marxasp.c"
#undef main

int main (int argc, char **argv)
{
   Param_File_Type *p = marx_pf_parse_cmd_line ("marxasp.par", NULL, argc, argv);
   if (p == NULL) return 2;
   if (-1 == marxasp_init (p)) return 1;
   pf_close_parameter_file (p);
   printf ("num_rows %u\n", (unsigned int) ((Time_Stop - Time_Start + 1.0) / Delta_Time));       /* :943 */
   printf ("time_start %.17g\n", Time_Start);
   printf ("delta_time %.17g\n", Delta_Time);
   printf ("amp %.17g %.17g %.17g\n", Ra_Amp, Dec_Amp, Roll_Amp);
   printf ("period %.17g %.17g %.17g\n", Ra_Period, Dec_Period, Roll_Period);
   printf ("phase %.17g %.17g %.17g\n", Ra_Phase, Dec_Phase, Roll_Phase);
   printf ("nominal_roll %.17g\n", Nominal_Roll_In_Radians);
   printf ("pointing %.17g %.17g %.17g\n", Nominal_Pointing.x, Nominal_Pointing.y, Nominal_Pointing.z);
   printf ("ra_hat %.17g %.17g %.17g\n", RA_Hat.x, RA_Hat.y, RA_Hat.z);
   printf ("dec_hat %.17g %.17g %.17g\n", Dec_Hat.x, Dec_Hat.y, Dec_Hat.z);
   return 0;
}
