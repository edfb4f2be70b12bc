/* Serialises source / spectrum / dither state (marx/libsrc/dither.c statics + public Marx_Source_Type).
 * Reference-side binding (integration/): compiled against the MARX tree, never into libmarxb200.so. */
#include <dither.c>
#include "calpack_io.h"

/* ASPSOL (DitherModel=FILE): the stock reader walks the file forward while the photons arrive (dither.c:288-377) and
 * converts every row on the way (spherical -> unrolled offsets, :331-350).  Here all rows are harvested through that
 * very function, so the table holds exactly the states the stock code would step through -- including the state left
 * by init_aspsol_dither, whose t1 has TSTART subtracted twice (:353 and :487).  Afterwards the reader is re-initialised
 * so that the stock obs.par code (get_aspsol_dither_mean, :420-440) still finds an open file. */
static int harvest_aspsol (mxcp_writer *w)
{
   unsigned int n = 0, cap = 4096;
   int record_only = (Get_Dither_Function == get_aspsol_dither_record_only);
   double *tab = (double *) malloc (cap * 7 * sizeof (double));
   if (tab == NULL) return -1;
   while (1)
     {
	double *r;
	if (n == cap)
	  {
	     cap *= 2;
	     if (NULL == (tab = (double *) realloc (tab, cap * 7 * sizeof (double)))) return -1;
	  }
	r = tab + 7 * n;
	r[0] = Aspsol.t1; r[1] = Aspsol.ra1; r[2] = Aspsol.dec1; r[3] = Aspsol.roll1;
	r[4] = Aspsol.dy1; r[5] = Aspsol.dz1; r[6] = Aspsol.dtheta1;
	n++;
	if (Aspsol.bt == NULL) break;                       /* cannot happen: init leaves the file open */
	if (-1 == get_single_aspsol_point ()) break;        /* end of file: the reader has closed it */
     }
   CP_F64 (w, "dither.aspsol", tab, 7 * (uint64_t) n);
   free (tab);
   memset ((char *) &Aspsol, 0, sizeof (Aspsol));
   if (-1 == init_aspsol_dither (record_only ? _MARX_DITHER_RECORD_ONLY : 0)) return -1;
   return 0;
}

int calpack_dump_dither (mxcp_writer *w)
{
   double v[12];
   v[0] = _Marx_Dither_Mode;
   v[1] = Ra_Amp; v[2] = Dec_Amp; v[3] = Roll_Amp;
   v[4] = Ra_Period; v[5] = Dec_Period; v[6] = Roll_Period;
   v[7] = Ra_Phase; v[8] = Dec_Phase; v[9] = Roll_Phase;
   v[10] = Nominal_Roll; v[11] = Aspect_Blur;
   if (Get_Dither_Function == get_zeroamp_internal_dither) v[1] = v[2] = v[3] = 0.0;
   if (-1 == CP_F64 (w, "dither.params", v, 12)) return -1;
   if (_Marx_Dither_Mode == _MARX_DITHER_MODE_ASPSOL) return harvest_aspsol (w);
   return 0;
}

int calpack_dump_source (mxcp_writer *w, void *marx_source)
{
   Marx_Source_Type *st = (Marx_Source_Type *) marx_source;
   double v[16], rot[4], img[4];
   int type;
   memset (v, 0, sizeof (v));
   int rayfile = 0;
   type = calpack_source_shape (st, v + 13, rot, img);       /* 0 POINT, 1 GAUSS, 2 BETA, 3 DISK, 4 LINE, 5 IMAGE, -1 unsupported */
   if ((type < 0) && (st->create_photons != NULL))
     {
	/* RAYFILE, and the sources that only exist as host code (USER = a dlopen'ed generator, SAOSAC = a ray file, SIMPUT = an
	 * external library): the stock host code produces the photons -- energies, directions, times, tags, dither -- and the
	 * caller injects them (marxb200_upload_from); the device source is never used and is packed as an inert POINT source.
	 * Return value 1 tells the caller. */
	type = 0; rayfile = 1;
     }
   if (type < 0) return -1;
   v[0] = type;
   v[1] = rayfile ? MARX_FLAT_SPECTRUM : st->spectrum.type;
   v[2] = st->p.x; v[3] = st->p.y; v[4] = st->p.z;
   v[5] = st->p_normal.x; v[6] = st->p_normal.y; v[7] = st->p_normal.z;
   v[8] = st->distance;
   v[9] = v[10] = 0.0;
   if (!rayfile && (st->spectrum.type == MARX_FLAT_SPECTRUM))
     { v[9] = st->spectrum.s.flat.emin; v[10] = st->spectrum.s.flat.emax; }
   v[11] = st->spectrum.total_flux;
   v[12] = Marx_Mirror_Geometric_Area;
   CP_F64 (w, "source.params", v, 16);
   if (type >= 4) CP_F64 (w, "source.rotation", rot, 4);      /* axis + angle taking (-1,0,0) to p */
   if (type == 5)
     {
	CP_F64 (w, "source.image_params", img, 4);           /* nx, ny, rad per x pixel, rad per y pixel */
	calpack_dump_image (w);
     }
   if (!rayfile && (st->spectrum.type == MARX_FILE_SPECTRUM))
     {
	CP_F64 (w, "source.spec_energies", st->spectrum.s.file.energies, st->spectrum.s.file.num);
	CP_F64 (w, "source.spec_cum_flux", st->spectrum.s.file.cum_flux, st->spectrum.s.file.num);
     }
   return rayfile;
}
