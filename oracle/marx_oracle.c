/* marx_oracle.c -- CPU restatement (plain C99) of MARX 5.5.3's per-photon ray-trace path:
 * source + arrival times + aspect dither, HRMA, HETG, ACIS-S.
 *
 * TEST INFRASTRUCTURE (see marx_oracle.h): the checker for the CUDA path, never the product.
 * Parity status: PINNED against the reference itself (oracle/_ref/marx_replay), bit for bit.
 *
 * Structure follows the reference (array of 136-byte records, one stage at a time over the batch);
 * arithmetic follows it operation by operation, including its float/double narrowing points.
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): the reference (gcc -O2, x86-64) has
 * no FMAs.  Random draws: Philox4x32-10 per (ray, stage), see oracle/ref/philox_rng.h.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "marx_oracle.h"
#include "../include/marxb200_calpack.h"

#define PI 3.14159265358979323846264338327950288      /* jdmath/src/jdmath.h:539 */
#define HBAR_C 1.973269631e-4                           /* marx/libsrc/marx.h:477 */
#define NUM_SHELLS 4
#define MAX_CHIPS 6
#define MAX_LAYERS 5
#define MAX_GAUSS 18

#define F_UNDETECTED 0x01
#define F_UNREFLECTED 0x02
#define F_UNDIFFRACTED 0x04
#define F_MISSED 0x08
#define F_VBLOCKED 0x10
#define F_STREAKED 0x200

static char Err[256];
const char *oracle_last_error (void) { return Err; }

/* ------------------------------------------------------------------------------------------- */
/* calibration pack access                                                                       */
typedef struct { char name[MARXB200_CALPACK_NAMELEN]; uint32_t dtype; uint64_t count; const unsigned char *data; } entry_t;

typedef struct
{
   uint32_t n; const double *hdr; const uint32_t *ntheta; uint32_t *offs; const float *theta;
}
wfold_t;

typedef struct
{
   const double *s;                  /* the 62 packed doubles */
   const float *corr_e, *corr_f; uint32_t ncorr;
   wfold_t wp, wh;
}
shell_t;

typedef struct
{
   const double *prm; const int32_t *orders; uint32_t norders; const float *energies; uint32_t nenergies;
   const float *cum_eff; const double *sectors; uint32_t nsectors;
}
gshell_t;

typedef struct
{
   const double *geom; const float *qe_e, *qe; uint32_t nqe; const float *fl_e, *fl; uint32_t nfl;
   const double *contam; const float *ce[MAX_LAYERS], *cmu[MAX_LAYERS], *cfxy[MAX_LAYERS]; uint32_t ncmu[MAX_LAYERS];
   const int32_t *fef_map;
}
chip_t;

typedef struct { uint32_t ng, ne; const float *energies, *channels, *gauss; } fef_t;

typedef struct { const double *geom; const float *qe_e, *qe; uint32_t nqe; } mcp_t;

struct oracle
{
   unsigned char *bytes; entry_t *entries; uint32_t nentries;
   uint64_t seed;
   const double *src, *dith, *hrma, *grat, *acis;
   const double *spec_e, *spec_c; uint32_t nspec;
   const double *src_rot, *img_prm; const float *img_cdf; uint32_t nimg;   /* LINE / IMAGE sources */
   const double *ffield;                                                                              /* MirrorType=FLATFIELD: min_y, min_z, max_y, max_z, x_pos (ffield.c:45-49) */
   uint64_t batch_first;                                                                              /* global index of the batch's first ray: the records' tags hold the low 32 bits only (marx.h:98) */
   const double *asp; uint32_t nasp, asp_pos; uint64_t last_generated;                                  /* ASPSOL states [nasp][7] (dither.c:288-359) */
   double asp_t_prev;
   double xf_off[3], xf_m[9]; int xf_init;                                 /* _Marx_Det_XForm_Matrix: dithered and restored photon after photon */
   const float *opt_e, *opt_b, *opt_d; uint32_t nopt;
   shell_t shell[NUM_SHELLS];
   gshell_t gshell[NUM_SHELLS];
   chip_t chip[MAX_CHIPS]; int nchips;
   fef_t *fefs; uint32_t nfefs;
   gshell_t support[2];                 /* LETG fine / coarse support gratings */
   const double *hrc, *hesf; mcp_t mcp[3]; int nmcps;
   const float *filt_e[4], *filt_q[4]; uint32_t nfilt[4];
   const float *hc_e, *hc_b, *hc_d, *hcr_e, *hcr_b, *hcr_d; uint32_t nhc, nhcr;
};

static const entry_t *find (oracle_t *o, const char *name)
{
   uint32_t i;
   for (i = 0; i < o->nentries; i++) if (0 == strcmp (o->entries[i].name, name)) return &o->entries[i];
   return NULL;
}
static const void *need (oracle_t *o, const char *name, uint64_t *count)
{
   const entry_t *e = find (o, name);
   if (e == NULL) { snprintf (Err, sizeof Err, "calpack: missing %s", name); if (count) *count = 0; return NULL; }
   if (count) *count = e->count;
   return e->data;
}
static const void *needf (oracle_t *o, uint64_t *count, const char *fmt, int a, int b)
{
   char nm[MARXB200_CALPACK_NAMELEN];
   snprintf (nm, sizeof nm, fmt, a, b);
   return need (o, nm, count);
}

static int load_wfold (oracle_t *o, wfold_t *w, int k, const char *which)
{
   char nm[MARXB200_CALPACK_NAMELEN]; uint64_t c; uint32_t i, acc = 0;
   snprintf (nm, sizeof nm, "hrma.shell%d.wfold_%s.hdr", k, which); w->hdr = (const double *) need (o, nm, &c);
   snprintf (nm, sizeof nm, "hrma.shell%d.wfold_%s.num_theta", k, which); w->ntheta = (const uint32_t *) need (o, nm, &c);
   w->n = (uint32_t) c;
   snprintf (nm, sizeof nm, "hrma.shell%d.wfold_%s.theta", k, which); w->theta = (const float *) need (o, nm, &c);
   if (!w->hdr || !w->ntheta || !w->theta) return -1;
   w->offs = (uint32_t *) malloc (sizeof (uint32_t) * (w->n + 1));
   for (i = 0; i < w->n; i++) { w->offs[i] = acc; acc += w->ntheta[i]; }
   return 0;
}

oracle_t *oracle_open (const char *path, uint64_t seed)
{
   FILE *fp = fopen (path, "rb");
   oracle_t *o; long sz; uint32_t i; size_t off; uint64_t c; int k;
   if (fp == NULL) { snprintf (Err, sizeof Err, "cannot open %s", path); return NULL; }
   o = (oracle_t *) calloc (1, sizeof (oracle_t));
   fseek (fp, 0, SEEK_END); sz = ftell (fp); fseek (fp, 0, SEEK_SET);
   o->bytes = (unsigned char *) malloc ((size_t) sz);
   if (fread (o->bytes, 1, (size_t) sz, fp) != (size_t) sz || memcmp (o->bytes, MARXB200_CALPACK_MAGIC, 8))
     { fclose (fp); snprintf (Err, sizeof Err, "bad calpack %s", path); oracle_close (o); return NULL; }
   fclose (fp);
   memcpy (&o->nentries, o->bytes + 8, 4);
   o->entries = (entry_t *) calloc (o->nentries, sizeof (entry_t));
   off = 16;
   for (i = 0; i < o->nentries; i++)
     {
        entry_t *e = &o->entries[i]; size_t nb;
        memcpy (e->name, o->bytes + off, MARXB200_CALPACK_NAMELEN); e->name[MARXB200_CALPACK_NAMELEN - 1] = 0;
        memcpy (&e->dtype, o->bytes + off + MARXB200_CALPACK_NAMELEN, 4);
        memcpy (&e->count, o->bytes + off + MARXB200_CALPACK_NAMELEN + 8, 8);
        off += MARXB200_CALPACK_NAMELEN + 16;
        e->data = o->bytes + off;
        nb = (size_t) e->count * mxcp_dtype_size (e->dtype);
        off += (nb + 7) & ~(size_t) 7;
     }
   o->seed = seed;
   o->src = (const double *) need (o, "source.params", NULL);
   o->dith = (const double *) need (o, "dither.params", NULL);
   o->ffield = find (o, "ffield.params") ? (const double *) need (o, "ffield.params", NULL) : NULL;
   o->hrma = o->ffield ? o->ffield : (const double *) need (o, "hrma.params", NULL);
   o->grat = (const double *) need (o, "grating.params", NULL);
   o->hrc = find (o, "hrc.params") ? (const double *) need (o, "hrc.params", NULL) : NULL;
   o->acis = o->hrc ? NULL : (const double *) need (o, "acis.params", NULL);
   if (!o->src || !o->dith || !o->hrma || !o->grat || (!o->acis && !o->hrc)) { oracle_close (o); return NULL; }
   if ((int) o->src[1] == 2)
     {
        o->spec_e = (const double *) need (o, "source.spec_energies", &c); o->nspec = (uint32_t) c;
        o->spec_c = (const double *) need (o, "source.spec_cum_flux", &c);
     }
   if ((int) o->dith[0] == 2)
     {
        o->asp = (const double *) need (o, "dither.aspsol", &c); o->nasp = (uint32_t) (c / 7);
        if (!o->asp || (o->nasp < 2)) { oracle_close (o); return NULL; }
     }
   if ((int) o->src[0] >= 4)
     {
        o->src_rot = (const double *) need (o, "source.rotation", NULL);
        if (!o->src_rot) { oracle_close (o); return NULL; }
     }
   if ((int) o->src[0] == 5)
     {
        o->img_prm = (const double *) need (o, "source.image_params", NULL);
        o->img_cdf = (const float *) need (o, "source.image_cdf", &c); o->nimg = (uint32_t) c;
        if (!o->img_prm || !o->img_cdf) { oracle_close (o); return NULL; }
     }
   if (!o->ffield)
     {
        o->opt_e = (const float *) need (o, "hrma.opt_energies", &c); o->nopt = (uint32_t) c;
        o->opt_b = (const float *) need (o, "hrma.opt_betas", &c);
        o->opt_d = (const float *) need (o, "hrma.opt_deltas", &c);
     }
   for (k = 0; (k < NUM_SHELLS) && !o->ffield; k++)
     {
        shell_t *s = &o->shell[k];
        s->s = (const double *) needf (o, &c, "hrma.shell%d.params", k, 0);
        s->corr_e = (const float *) needf (o, &c, "hrma.shell%d.corr_energies", k, 0); s->ncorr = (uint32_t) c;
        s->corr_f = (const float *) needf (o, &c, "hrma.shell%d.corr_factors", k, 0);
        if (!s->s || load_wfold (o, &s->wp, k, "p") || load_wfold (o, &s->wh, k, "h")) { oracle_close (o); return NULL; }
     }
   if ((int) o->grat[0] != 0)
     for (k = 0; k < NUM_SHELLS; k++)
       {
          gshell_t *g = &o->gshell[k];
          g->prm = (const double *) needf (o, &c, "grating.shell%d.params", k, 0);
          g->orders = (const int32_t *) needf (o, &c, "grating.shell%d.order_list", k, 0); g->norders = (uint32_t) c;
          g->energies = (const float *) needf (o, &c, "grating.shell%d.energies", k, 0); g->nenergies = (uint32_t) c;
          g->cum_eff = (const float *) needf (o, &c, "grating.shell%d.cum_eff", k, 0);
          g->sectors = (const double *) needf (o, &c, "grating.shell%d.sectors", k, 0); g->nsectors = (uint32_t) (c / 6);
          if (!g->prm || !g->orders || !g->energies || !g->cum_eff) { oracle_close (o); return NULL; }
       }
   if ((int) o->grat[0] == 2)
     for (k = 0; k < 2; k++)
       {
          gshell_t *g = &o->support[k]; char nm[MARXB200_CALPACK_NAMELEN];
          snprintf (nm, sizeof nm, "grating.support%d.params", k);
          if (!find (o, nm)) continue;
          g->prm = (const double *) needf (o, &c, "grating.support%d.params", k, 0);
          g->orders = (const int32_t *) needf (o, &c, "grating.support%d.order_list", k, 0); g->norders = (uint32_t) c;
          g->energies = (const float *) needf (o, &c, "grating.support%d.energies", k, 0); g->nenergies = (uint32_t) c;
          g->cum_eff = (const float *) needf (o, &c, "grating.support%d.cum_eff", k, 0);
          g->nsectors = 0;
       }
   if (o->hrc)
     {
        o->nmcps = (int) o->hrc[1];
        for (k = 0; k < o->nmcps; k++)
          {
             o->mcp[k].geom = (const double *) needf (o, &c, "hrc.mcp%d.geom", k, 0);
             o->mcp[k].qe_e = (const float *) needf (o, &c, "hrc.mcp%d.qe_energies", k, 0); o->mcp[k].nqe = (uint32_t) c;
             o->mcp[k].qe = (const float *) needf (o, &c, "hrc.mcp%d.qe", k, 0);
          }
        for (k = 0; k < 4; k++)
          {
             o->filt_e[k] = (const float *) needf (o, &c, "hrc.filter%d.energies", k, 0); o->nfilt[k] = (uint32_t) c;
             o->filt_q[k] = (const float *) needf (o, &c, "hrc.filter%d.qe", k, 0);
          }
        o->hesf = (const double *) need (o, "hrc.hesf", &c);
        o->hc_e = (const float *) need (o, "hrc.hesf_c_energies", &c); o->nhc = (uint32_t) c;
        o->hc_b = (const float *) need (o, "hrc.hesf_c_betas", &c); o->hc_d = (const float *) need (o, "hrc.hesf_c_deltas", &c);
        o->hcr_e = (const float *) need (o, "hrc.hesf_cr_energies", &c); o->nhcr = (uint32_t) c;
        o->hcr_b = (const float *) need (o, "hrc.hesf_cr_betas", &c); o->hcr_d = (const float *) need (o, "hrc.hesf_cr_deltas", &c);
     }
   if (o->acis && ((int) o->acis[0] != 0))
     {
        const uint32_t *nf = (const uint32_t *) need (o, "acis.num_fefs", NULL);
        uint32_t j;
        if (!nf) { oracle_close (o); return NULL; }
        o->nfefs = *nf; o->fefs = (fef_t *) calloc (o->nfefs ? o->nfefs : 1, sizeof (fef_t));
        for (j = 0; j < o->nfefs; j++)
          {
             const uint32_t *d = (const uint32_t *) needf (o, &c, "acis.fef%d.dims", (int) j, 0);
             if (!d) { oracle_close (o); return NULL; }
             o->fefs[j].ng = d[0]; o->fefs[j].ne = d[1];
             o->fefs[j].energies = (const float *) needf (o, &c, "acis.fef%d.energies", (int) j, 0);
             o->fefs[j].channels = (const float *) needf (o, &c, "acis.fef%d.channels", (int) j, 0);
             o->fefs[j].gauss = (const float *) needf (o, &c, "acis.fef%d.gauss", (int) j, 0);
          }
        o->nchips = (int) o->acis[1];
        for (k = 0; k < o->nchips; k++)
          {
             chip_t *ch = &o->chip[k]; uint32_t l;
             ch->geom = (const double *) needf (o, &c, "acis.chip%d.geom", k, 0);
             ch->qe_e = (const float *) needf (o, &c, "acis.chip%d.qe_energies", k, 0); ch->nqe = (uint32_t) c;
             ch->qe = (const float *) needf (o, &c, "acis.chip%d.qe", k, 0);
             ch->fl_e = (const float *) needf (o, &c, "acis.chip%d.filter_energies", k, 0); ch->nfl = (uint32_t) c;
             ch->fl = (const float *) needf (o, &c, "acis.chip%d.filter_qe", k, 0);
             ch->contam = (const double *) needf (o, &c, "acis.chip%d.contam", k, 0);
             ch->fef_map = (const int32_t *) needf (o, &c, "acis.chip%d.fef_map", k, 0);
             if (!ch->geom || !ch->contam || !ch->fef_map) { oracle_close (o); return NULL; }
             for (l = 0; l < (uint32_t) ch->contam[0]; l++)
               {
                  ch->ce[l] = (const float *) needf (o, &c, "acis.chip%d.contam_e%d", k, (int) l); ch->ncmu[l] = (uint32_t) c;
                  ch->cmu[l] = (const float *) needf (o, &c, "acis.chip%d.contam_mu%d", k, (int) l);
                  if ((int) ch->contam[1] == 0) ch->cfxy[l] = (const float *) needf (o, &c, "acis.chip%d.contam_fxy%d", k, (int) l);
               }
          }
     }
   return o;
}

/* rays kept by the last oracle_trace: its n, or fewer when the ASPSOL file ended inside the batch */
uint64_t oracle_last_generated (const oracle_t *o) { return o ? o->last_generated : 0; }

void oracle_close (oracle_t *o)
{
   int k;
   if (o == NULL) return;
   for (k = 0; k < NUM_SHELLS; k++) { free (o->shell[k].wp.offs); free (o->shell[k].wh.offs); }
   free (o->fefs); free (o->entries); free (o->bytes); free (o);
}

/* ------------------------------------------------------------------------------------------- */
/* draw stream: Philox4x32-10 per (ray, stage); uniform = u32/(2^32-1) (jdmath/src/random.c:151) */
typedef struct { uint64_t seed, ray; uint32_t stage, draw, blk[4]; int have_spare; double spare; } rng_t;

static void rng_set (rng_t *r, uint64_t seed, uint64_t ray, uint32_t stage)
{ r->seed = seed; r->ray = ray; r->stage = stage; r->draw = 0; r->have_spare = 0; }
/* The draws are keyed on the 64-bit global ray index; a record only carries its low 32 bits (the reference's tag,
 * marx.h:98).  Within a batch (< 2^32 rays) the full index is the batch's first ray plus the tag's distance from it. */
static uint64_t ray_of (const oracle_t *o, uint32_t tag)
{ return o->batch_first + (uint64_t) (uint32_t) (tag - (uint32_t) o->batch_first); }

static uint32_t rng_u32 (rng_t *r)
{
   if ((r->draw & 3) == 0)
     {
        uint32_t c[4], k0 = (uint32_t) r->seed, k1 = (uint32_t) (r->seed >> 32); int i;
        c[0] = (uint32_t) r->ray; c[1] = (uint32_t) (r->ray >> 32); c[2] = r->draw >> 2; c[3] = r->stage;
        for (i = 0; i < 10; i++)
          {
             uint64_t p0 = (uint64_t) 0xD2511F53u * c[0], p1 = (uint64_t) 0xCD9E8D57u * c[2];
             uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k1;
             c[1] = (uint32_t) p1; c[3] = (uint32_t) p0; c[0] = n0; c[2] = n2;
             k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
          }
        memcpy (r->blk, c, sizeof c);
     }
   return r->blk[(r->draw++) & 3];
}
static double rng_uniform (rng_t *r) { return (double) rng_u32 (r) * (1.0 / (double) 0xFFFFFFFFU); }
/* jdmath/src/gaussrnd.c:30-55 */
static double rng_gauss (rng_t *r)
{
   double g1, g2, g, s;
   if (r->have_spare) { r->have_spare = 0; return r->spare; }
   do { g1 = 2.0 * rng_uniform (r) - 1.0; g2 = 2.0 * rng_uniform (r) - 1.0; g = g1 * g1 + g2 * g2; }
   while ((g >= 1.0) || (g == 0.0));
   s = sqrt (-2.0 * log (g) / g);
   r->spare = g2 * s; r->have_spare = 1;
   return g1 * s;
}
/* gaussrnd.c:57-67 */
static double rng_expn (rng_t *r) { double u; do u = rng_uniform (r); while (u == 0.0); return -log (u); }

/* ------------------------------------------------------------------------------------------- */
/* jdmath vector / interpolation primitives                                                      */
static double dot3 (const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3 (const double *a, const double *b, double *c)          /* vector.c:40-58 */
{ double z = a[0] * b[1] - a[1] * b[0], x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2]; c[0] = x; c[1] = y; c[2] = z; }
static double len3 (const double *a)                                      /* vector.c:78-97 */
{
   double x = fabs (a[0]), y = fabs (a[1]), z = fabs (a[2]), t;
   if (z < x) { t = z; z = x; x = t; }
   if (z < y) { t = z; z = y; y = t; }
   if (z == 0.0) return 0.0;
   x = x / z; y = y / z;
   return z * sqrt (1.0 + x * x + y * y);
}
static void norm3 (double *a) { double l = len3 (a); if (l != 0.0) { a[0] = a[0] / l; a[1] = a[1] / l; a[2] = a[2] / l; } }
static void rot_unit1 (double *p, const double *n, double c, double s)   /* vector.c:173-202 */
{
   double pn = dot3 (p, n), nxp[3], f = pn * (1.0 - c), u[3];
   cross3 (n, p, nxp);
   u[0] = c * p[0] + f * n[0] + s * nxp[0];
   u[1] = c * p[1] + f * n[1] + s * nxp[1];
   u[2] = c * p[2] + f * n[2] + s * nxp[2];
   norm3 (u);
   p[0] = u[0]; p[1] = u[1]; p[2] = u[2];
}
static void rot_unit (double *p, const double *n, double theta) { rot_unit1 (p, n, cos (theta), sin (theta)); }
static void mat3 (const double *m, double *v)                            /* rotate.c:121-131 */
{
   double x = v[0], y = v[1], z = v[2];
   v[0] = m[0] * x + m[1] * y + m[2] * z;
   v[1] = m[3] * x + m[4] * y + m[5] * z;
   v[2] = m[6] * x + m[7] * y + m[8] * z;
}
static void mat3t (const double *m, double *v)                           /* trans.c:54-64 */
{
   double x = v[0], y = v[1], z = v[2];
   v[0] = m[0] * x + m[3] * y + m[6] * z;
   v[1] = m[1] * x + m[4] * y + m[7] * z;
   v[2] = m[2] * x + m[5] * y + m[8] * z;
}
/* finterpo.c:37-57 */
static unsigned int bsearch_f (float x, const float *xp, unsigned int n)
{
   unsigned int n0 = 0, n1 = n, n2;
   while (n1 > n0 + 1)
     {
        n2 = (n0 + n1) / 2;
        if (xp[n2] >= x) { if (xp[n2] == x) return n2; n1 = n2; }
        else n0 = n2;
     }
   if (x >= xp[n0]) return n1;
   return n0;
}
/* finterpo.c:59-85 (the one-past-the-end read of :68 is guarded) */
static float interp_f (float x, const float *xp, const float *yp, unsigned int n)
{
   unsigned int n0, n1; double x0, x1;
   if (n == 1) return yp[0];
   n1 = bsearch_f (x, xp, n); n0 = n1 - 1;
   if ((n1 < n) && (x == xp[n1])) return yp[n1];
   if (n1 == n) { n1--; n0--; }
   if (n1 == 0) n0 = 1;
   x0 = xp[n0]; x1 = xp[n1];
   if (x1 == x0) return yp[n1];
   return yp[n0] + (yp[n1] - yp[n0]) / (x1 - x0) * (x - x0);
}
static unsigned int bsearch_d (double x, const double *xp, unsigned int n, unsigned int stride)
{
   unsigned int n0 = 0, n1 = n, n2;
   while (n1 > n0 + 1)
     {
        n2 = (n0 + n1) / 2;
        if (xp[n2 * stride] >= x) { if (xp[n2 * stride] == x) return n2; n1 = n2; }
        else n0 = n2;
     }
   if (x >= xp[n0 * stride]) return n1;
   return n0;
}
static double interp_d (double x, const double *xp, const double *yp, unsigned int n)   /* dinterpo.c */
{
   unsigned int n0, n1; double x0, x1;
   if (n == 1) return yp[0];
   n1 = bsearch_d (x, xp, n, 1); n0 = n1 - 1;
   if ((n1 < n) && (x == xp[n1])) return yp[n1];
   if (n1 == n) { n1--; n0--; }
   if (n1 == 0) n0 = 1;
   x0 = xp[n0]; x1 = xp[n1];
   if (x1 == x0) return yp[n1];
   return yp[n0] + (yp[n1] - yp[n0]) / (x1 - x0) * (x - x0);
}

/* ------------------------------------------------------------------------------------------- */
/* stage 0: marx_create_photons (source.c:268-384) for POINT sources + dither (dither.c:551-628)  */
/* returns the number of rays kept: n, or fewer when the ASPSOL file ends inside the batch (dither.c:296-301, 361-369:
 * the first ray the reader cannot bracket is not dithered and ends the simulation, source.c:355, marx.c:577-578) */
static uint64_t stage_source (oracle_t *o, uint64_t first, uint64_t n, double *time_base, oracle_photon *ph)
{
   const double *s = o->src, *d = o->dith;
   double mt = (s[11] <= 0.0) ? 0.0 : 1.0 / s[11] / s[12];              /* source.c:260-264 */
   double t = *time_base;
   uint64_t i;
   for (i = 0; i < n; i++)
     {
        oracle_photon *at = ph + i; rng_t r;
        memset (at, 0, sizeof (*at));
        rng_set (&r, o->seed, first + i, 0);
        if ((int) s[1] == 2) at->energy = interp_d (rng_uniform (&r), o->spec_c, o->spec_e, o->nspec);   /* prob.c:55 */
        else { double emin = s[9], de = s[10] - emin; at->energy = emin + de * rng_uniform (&r); }       /* spectrum.c:140-145 */
        at->p[0] = s[2]; at->p[1] = s[3]; at->p[2] = s[4];                                               /* s-point.c:76 */
        if ((int) s[0] >= 4)
          {
             /* LINE (s-line.c:86-98) / IMAGE (s-image.c:330-356): ray about (-1,0,0), rotated onto the source direction */
             double q[3], axis[3];
             if ((int) s[0] == 4)
               {
                  double theta = s[13] * (-1.0 + 2.0 * rng_uniform (&r));
                  double sin_theta = -sin (theta);
                  q[0] = -cos (theta); q[1] = sin_theta * s[14]; q[2] = sin_theta * s[15];
               }
             else
               {
                  unsigned int nx = (unsigned int) o->img_prm[0], ny = (unsigned int) o->img_prm[1];
                  unsigned int ofs = bsearch_f ((float) rng_uniform (&r), o->img_cdf, o->nimg);
                  double y = (double) (ofs / nx), x = (double) (ofs % nx), cos_y;
                  y += -0.5 * ny + (rng_uniform (&r) - 0.5);
                  x += -0.5 * nx + (rng_uniform (&r) - 0.5);
                  y = y * o->img_prm[3];
                  x = x * o->img_prm[2];
                  cos_y = cos (y);
                  q[0] = -cos_y * cos (x); q[1] = cos_y * sin (x); q[2] = -sin (y);
               }
             axis[0] = o->src_rot[0]; axis[1] = o->src_rot[1]; axis[2] = o->src_rot[2];
             rot_unit (q, axis, o->src_rot[3]);
             at->p[0] = q[0]; at->p[1] = q[1]; at->p[2] = q[2];
          }
        else if ((int) s[0] != 0)
          {
             /* GAUSS / BETA / DISK: s-gauss.c:96-137, s-beta.c:100-125, s-disk.c:86-106 (normal restarts from
              * st->p_normal for every ray, as the reference does at the start of every batch) */
             double nrm[3], src[3], theta, rnd;
             nrm[0] = s[5]; nrm[1] = s[6]; nrm[2] = s[7]; src[0] = s[2]; src[1] = s[3]; src[2] = s[4];
             rot_unit (nrm, src, 2.0 * PI * rng_uniform (&r));
             if ((int) s[0] == 3) theta = s[13] * sqrt (s[14] + s[15] * rng_uniform (&r));
             else
               {
                  do rnd = rng_uniform (&r); while (rnd == 0.0);
                  if ((int) s[0] == 1) theta = s[13] * sqrt (-log (rnd));
                  else theta = s[13] * sqrt (pow (rnd, s[14]) - 1.0);
               }
             rot_unit (at->p, nrm, theta);
          }
        t += mt * rng_expn (&r);                                                                          /* source.c:326 */
        at->arrival_time = t;
        at->tag = (uint32_t) (first + i);
        if ((int) d[0] != 0)
          {
             /* get_internal_dither, dither.c:167-182: angles pass through float fields */
             double tt = (2.0 * PI) * t, ra, dec, roll, dra, ddec, n3[3], xax[3] = {1, 0, 0};
             double cra, sra, cdec, sdec, cth, sth;
             if ((int) d[0] == 2)
               {
                  /* get_aspsol_dither, dither.c:361-400: the forward-only reader stops at the first state k >= 1 with
                   * t < t_k (times never decrease); the seven values are interpolated between it and its predecessor
                   * and stored through the float fields of Marx_Dither_Type */
                  const double *A = o->asp, *s0, *s1; double dt;
                  while ((o->asp_pos < o->nasp) && (t >= A[7 * o->asp_pos])) o->asp_pos++;
                  if (o->asp_pos >= o->nasp)
                    {
                       memset (at, 0, sizeof (*at));
                       *time_base = (i == 0) ? *time_base : ph[i - 1].arrival_time;
                       return i;
                    }
                  s1 = A + 7 * o->asp_pos; s0 = s1 - 7;
                  dt = s1[0] - s0[0];
                  if (dt != 0) dt = (t - s0[0]) / dt;
                  at->dither[0] = (float) (s0[1] + dt * (s1[1] - s0[1]));
                  at->dither[1] = (float) (s0[2] + dt * (s1[2] - s0[2]));
                  at->dither[2] = (float) (s0[3] + dt * (s1[3] - s0[3]));
                  at->dither[3] = (float) (s0[4] + dt * (s1[4] - s0[4]));
                  at->dither[4] = (float) (s0[5] + dt * (s1[5] - s0[5]));
                  at->dither[5] = (float) (s0[6] + dt * (s1[6] - s0[6]));
               }
             else
               {
             at->dither[0] = (float) (d[1] * sin (tt / d[4] + d[7]));
             at->dither[1] = (float) (d[2] * sin (tt / d[5] + d[8]));
             at->dither[2] = (float) (d[10] + d[3] * sin (tt / d[6] + d[9]));
               }
             ra = at->dither[0]; dec = at->dither[1]; roll = at->dither[2];
             dra = d[11] * rng_gauss (&r); ddec = d[11] * rng_gauss (&r);                                /* dither.c:623-624 */
             ra += dra; dec += ddec;
             rot_unit (at->p, xax, -roll);                                                                /* apply_dither :551-581 */
             cra = cos (ra); sra = sin (ra); cdec = cos (dec); sdec = sin (dec);
             cth = cdec * cra;
             n3[0] = 0; n3[1] = sdec; n3[2] = -cdec * sra;
             sth = len3 (n3);
             if (sth > 1e-20) { n3[0] /= sth; n3[1] /= sth; n3[2] /= sth; rot_unit1 (at->p, n3, cth, sth); }
          }
     }
   *time_base = t;
   return n;
}

/* ------------------------------------------------------------------------------------------- */
/* stage 1: _marx_hrma_mirror_reflect (hrma.c:1161-1341)                                          */
typedef struct { double r, i; } cplx;
static cplx c_div (cplx a, cplx b)                                       /* complex.c:54-80 */
{
   cplx z; double ratio, denom;
   if (fabs (b.r) > fabs (b.i)) { ratio = b.i / b.r; denom = b.r + b.i * ratio; z.r = (a.r + ratio * a.i) / denom; z.i = (a.i - a.r * ratio) / denom; }
   else { ratio = b.r / b.i; denom = b.r * ratio + b.i; z.r = (a.r * ratio + a.i) / denom; z.i = (a.i * ratio - a.r) / denom; }
   return z;
}
static cplx c_sqrt (cplx a)                                              /* complex.c:168-227 */
{
   double fr = fabs (a.r), fi = fabs (a.i), r, ratio;
   if (fr > fi) { ratio = a.i / a.r; r = fr * sqrt (1.0 + ratio * ratio); }
   else if (fi == 0.0) r = 0.0;
   else { ratio = a.r / a.i; r = fi * sqrt (1.0 + ratio * ratio); }
   if (r == 0.0) return a;
   if (a.r >= 0.0) { a.r = sqrt (0.5 * (r + a.r)); a.i = 0.5 * a.i / a.r; }
   else { r = sqrt (0.5 * (r - a.r)); a.r = 0.5 * a.i / r; a.i = r; if (a.r < 0.0) { a.r = -a.r; a.i = -a.i; } }
   return a;
}
static double reflectivity (double ct, double beta, double delta)        /* reflect.c:39-77 */
{
   cplx n, nsqr, root, num, den, eperp, epar; double st;
   n.r = 1.0 - delta; n.i = beta;
   st = sqrt (1.0 - ct * ct);
   nsqr.r = n.r * n.r - n.i * n.i; nsqr.i = n.r * n.i + n.i * n.r;
   root.r = -st * st + 1.0 * nsqr.r; root.i = 1.0 * nsqr.i; root = c_sqrt (root);
   num.r = ct + -1.0 * root.r; num.i = -1.0 * root.i; den.r = ct + 1.0 * root.r; den.i = 1.0 * root.i;
   eperp = c_div (num, den);
   num.r = ct * nsqr.r + -1.0 * root.r; num.i = ct * nsqr.i + -1.0 * root.i;
   den.r = ct * nsqr.r + 1.0 * root.r; den.i = ct * nsqr.i + 1.0 * root.i;
   epar = c_div (num, den);
   return 0.5 * (epar.r * epar.r + epar.i * epar.i + eperp.r * eperp.r + eperp.i * eperp.i);
}
static int conic_hit (const double *cn, double *x0, const double *p, double *nrm)   /* hrma.c:411-478 */
{
   double a = cn[0], b = cn[1], c = cn[2], xmin = cn[3], xmax = cn[4];
   double tp, tm, xy, xz, alpha, beta, gamma, xp, xm;
   tp = -x0[0] / p[0]; xy = x0[1] + tp * p[1]; xz = x0[2] + tp * p[2];
   alpha = a * p[0] * p[0] - 1.0;
   beta = b * p[0] - 2.0 * (p[1] * xy + p[2] * xz);
   gamma = c - xz * xz - xy * xy;
   if (alpha == 0.0) { if (beta == 0.0) return -1; tp = tm = -gamma / beta; }
   else
     {                                                                   /* qroot.c:36-68 */
        double bsqr = beta * beta, ac4 = alpha * gamma * 4, nb2a = -beta / (2.0 * alpha);
        if (bsqr > ac4) { double f = 1.0 + sqrt (1.0 - ac4 / bsqr); tp = -2.0 * gamma / (beta * f); tm = nb2a * f; }
        else if (bsqr == ac4) tp = tm = nb2a;
        else return -1;
     }
   xp = p[0] * tp; xm = p[0] * tm;
   if ((xp >= xmin) && (xp < xmax))
     {
        if ((xm >= xmin) && (xm < xmax) && (xm > xp)) { x0[0] = xm; x0[1] = xy + p[1] * tm; x0[2] = xz + p[2] * tm; }
        else { x0[0] = xp; x0[1] = xy + p[1] * tp; x0[2] = xz + p[2] * tp; }
     }
   else if ((xm >= xmin) && (xm < xmax)) { x0[0] = xm; x0[1] = xy + p[1] * tm; x0[2] = xz + p[2] * tm; }
   else return -1;
   nrm[0] = (a - 1) * x0[0] + 0.5 * b; nrm[1] = -x0[1]; nrm[2] = -x0[2];
   norm3 (nrm);
   return 0;
}
static double wfold_theta (const wfold_t *w, uint32_t k, double p)        /* wfold.c:304-332 */
{
   const double *h = w->hdr + 6 * k; const float *t = w->theta + w->offs[k]; uint32_t nt = w->ntheta[k], i; double di;
   if (p < h[1]) return 0.0;
   if (p > h[3]) return pow (h[4] * (1.0 - p), h[5]);
   di = (p - h[1]) / h[2]; i = (uint32_t) di;
   if (i + 1 >= nt) return (double) t[nt - 1];
   di -= (double) i;
   return (1.0 - di) * t[i] + di * t[i + 1];
}
static double wfold_interp (const wfold_t *w, double energy, double sa, double r)   /* wfold.c:334-369 */
{
   double ea, t0, t1, e0, e1; uint32_t i;
   if (w->n == 0) return 0.0;
   if (w->n == 1) return wfold_theta (w, 0, r);
   ea = energy * sa;
   i = bsearch_d (ea, w->hdr, w->n, 6);
   if (i == w->n) i--;
   if (i == 0) i++;
   t0 = wfold_theta (w, i - 1, r); t1 = wfold_theta (w, i, r);
   e0 = w->hdr[6 * (i - 1)]; e1 = w->hdr[6 * i];
   return t0 + (t1 - t0) * (ea - e0) / (e1 - e0);
}
static int struts_hit (const double *x0, const double *p, double cap, const double *st)   /* hrma.c:928-968 */
{
   double theta = 30.0 * (PI / 180.0), ct = cos (theta), sn = sin (theta); int s, i;
   for (s = 0; s < 2; s++)
     {
        double hw = st[2 * s + 1], y = x0[1], z = x0[2], t = (st[2 * s] + cap - x0[0]) / p[0];
        y += p[1] * t; z += p[2] * t;
        for (i = 0; i < 3; i++)
          {
             if (i != 0) { double tmp = ct * y - sn * z; z = sn * y + ct * z; y = tmp; }
             if (((-hw < y) && (y < hw)) || ((-hw < z) && (z < hw))) return 1;
          }
     }
   return 0;
}
/* reflect_from_conic, hrma.c:488-545 (+ blur_normal :1054-1093).  0 ok, -1 absorbed, -2 missed */
static int conic_reflect (oracle_t *o, const double *cn, const wfold_t *w, double scat, double blur, double energy,
                          double beta, double delta, double corr, double *x, double *p, rng_t *r)
{
   double nrm[3], pdn, perp[3], len, phi, dg, ax[3];
   int ideal = (int) o->hrma[2], use_blur = (int) o->hrma[3], use_wfold = (int) o->hrma[4];
   if (-1 == conic_hit (cn, x, p, nrm)) return -2;
   if (use_blur)
     {
        len = sqrt (nrm[1] * nrm[1] + nrm[2] * nrm[2]);
        perp[0] = 0.0; perp[1] = nrm[2] / len; perp[2] = -nrm[1] / len;
        phi = (2.0 * PI) * rng_uniform (r);
        rot_unit (perp, nrm, phi);
        phi = blur * (1.0 / 3600.0 * PI / 180.0);
        phi = phi * rng_gauss (r);
        rot_unit (nrm, perp, phi);
     }
   pdn = dot3 (p, nrm);
   if (!ideal) { double u = rng_uniform (r), rfl = reflectivity (fabs (pdn), beta, delta); if (u >= rfl * corr) return -1; }
   { double f = -2.0 * pdn; p[0] = 1.0 * p[0] + f * nrm[0]; p[1] = 1.0 * p[1] + f * nrm[1]; p[2] = 1.0 * p[2] + f * nrm[2]; }
   if (!use_wfold) return 0;
   dg = wfold_interp (w, energy, -pdn, rng_uniform (r)) * scat;
   if (dg > PI / 4) return -1;
   if (rng_uniform (r) < 0.5) dg = -dg;
   cross3 (p, nrm, ax);
   rot_unit (p, ax, dg);
   return 0;
}
static void stage_mirror (oracle_t *o, uint64_t n, oracle_photon *ph)
{
   static const double precol[4] = {1492.060, 0.5 * 0.5 * 25.4, 942.266, 0.5 * 0.5 * 25.4};          /* hrma.c:908-926 */
   static const double capst[4] = {0.5 * 1.965 * 25.4, 0.5 * 0.75 * 25.4, -0.5 * 1.965 * 25.4, 0.5 * 0.75 * 25.4};
   static const double postcol[4] = {-1050.353, 0.5 * 0.5 * 25.4, -1271.333, 0.5 * 0.5 * 25.4};
   const double *H = o->hrma; double dist = o->src[8], cap = H[1];
   int ideal = (int) H[2], struts = (int) H[5], scale = (int) H[6];
   uint64_t i;
   if (o->ffield)
     {
        /* _marx_ff_mirror_reflect / project_photon, ffield.c:61-108: no optics, the ray starts on a rectangle at x = FF_XPos */
        const double *F = o->ffield;
        for (i = 0; i < n; i++)
          {
             oracle_photon *at = ph + i; rng_t r;
             if (at->flags & 0xFF) continue;
             rng_set (&r, o->seed, ray_of (o, at->tag), 1);
             at->x[2] = F[1] + rng_uniform (&r) * (F[3] - F[1]);
             at->x[1] = F[0] + rng_uniform (&r) * (F[2] - F[0]);
             at->x[0] = F[4];
             if (dist > 0.0)
               { at->p[0] = 1.0 * at->x[0] + dist * at->p[0]; at->p[1] = 1.0 * at->x[1] + dist * at->p[1]; at->p[2] = 1.0 * at->x[2] + dist * at->p[2]; norm3 (at->p); }
          }
        return;
     }
   for (i = 0; i < n; i++)
     {
        oracle_photon *at = ph + i; rng_t r; const shell_t *sh; const double *s; int k, found = 0, st;
        double radius, theta, beta = 0.0, delta = 1.0, corr = 1.0; uint32_t quad;
        if (at->flags & 0xFF) continue;
        rng_set (&r, o->seed, ray_of (o, at->tag), 1);
        if (!ideal && (rng_uniform (&r) > H[0])) { at->flags |= F_VBLOCKED; continue; }                /* hrma.c:1185-1192 */
        while (!found)                                                                                  /* :984-1050 */
          {
             double u = rng_uniform (&r);
             for (k = 0; k < NUM_SHELLS; k++) if (u < o->shell[k].s[19]) { found = 1; break; }
          }
        sh = &o->shell[k]; s = sh->s; at->mirror_shell = (uint32_t) k;
        radius = s[20] + (s[21] - s[20]) * rng_uniform (&r);
        do { theta = rng_uniform (&r); quad = (uint32_t) (4.0 * theta); } while (0 == (((uint32_t) s[1]) & (1u << quad)));
        theta = (2.0 * PI) * (theta - 1.0 / 8.0);
        at->x[2] = radius * cos (theta); at->x[1] = radius * sin (theta); at->x[0] = s[18];
        at->x[2] -= s[14]; at->x[1] -= s[13];
        if (dist > 0.0)
          { at->p[0] = 1.0 * at->x[0] + dist * at->p[0]; at->p[1] = 1.0 * at->x[1] + dist * at->p[1]; at->p[2] = 1.0 * at->x[2] + dist * at->p[2]; norm3 (at->p); }
        if (struts && struts_hit (at->x, at->p, cap, precol)) { at->flags |= F_VBLOCKED; continue; }
        at->x[0] += s[12]; at->x[1] += s[13]; at->x[2] += s[14];                                       /* :1222-1235 */
        mat3 (s + 26, at->x); mat3 (s + 26, at->p);
        if (!ideal && o->nopt)
          {
             float ef = (float) at->energy;
             beta = interp_f (ef, o->opt_e, o->opt_b, o->nopt); delta = interp_f (ef, o->opt_e, o->opt_d, o->nopt);
             if (scale && sh->ncorr) { corr = interp_f (ef, sh->corr_e, sh->corr_f, sh->ncorr); corr = sqrt (corr); }
          }
        st = conic_reflect (o, s + 2, &sh->wp, s[24], s[22], at->energy, beta, delta, corr, at->x, at->p, &r);
        if (st != 0) { at->flags |= F_UNREFLECTED; continue; }
        mat3 (s + 35, at->p); mat3 (s + 35, at->x);
        at->x[0] -= s[12]; at->x[1] -= s[13]; at->x[2] -= s[14];
        if (struts && struts_hit (at->x, at->p, cap, capst)) { at->flags |= F_VBLOCKED; continue; }
        at->x[0] += s[15]; at->x[1] += s[16]; at->x[2] += s[17];
        mat3 (s + 44, at->p); mat3 (s + 44, at->x);
        st = conic_reflect (o, s + 7, &sh->wh, s[25], s[23], at->energy, beta, delta, corr, at->x, at->p, &r);
        if (st != 0) { at->flags |= F_UNREFLECTED; continue; }
        mat3 (s + 53, at->p); mat3 (s + 53, at->x);
        at->x[0] -= s[15]; at->x[1] -= s[16]; at->x[2] -= s[17];
        if (struts && struts_hit (at->x, at->p, cap, postcol)) { at->flags |= F_VBLOCKED; continue; }
     }
}

/* ------------------------------------------------------------------------------------------- */
/* stage 2: diffract (diffract.c:974-1130), HETG                                                  */
static void rot_x (double *a, double theta)                              /* diffract.c:689-700 */
{ double c = cos (theta), s = sin (theta), ay = a[1], az = a[2]; a[1] = c * ay - s * az; a[2] = s * ay + c * az; }
static int torus_hit (double *x0, const double *p, double rowland)       /* diffract.c:622-687 */
{
   double t = -x0[0] / p[0], pxpz, x2, r2, pdx, a, b, c, d, t0, a2, a3, b2; unsigned int it = 10;
   x0[0] = 0.0; x0[1] = x0[1] + p[1] * t; x0[2] = x0[2] + p[2] * t;
   pxpz = p[0] * p[0] + p[2] * p[2]; x2 = x0[2] * x0[2] + x0[1] * x0[1]; r2 = rowland * rowland; pdx = dot3 (p, x0);
   a = 4.0 * pdx; b = 2.0 * x2 + a * pdx - r2 * pxpz; c = a * x2 - 2.0 * r2 * p[2] * x0[2]; d = x2 * x2 - r2 * x0[2] * x0[2];
   t0 = -rowland * sqrt (pxpz);
   a2 = 2.0 * a; a3 = 3.0 * a; b2 = 2.0 * b;
   while (1)
     {
        double t2 = t0 * t0, num = t2 * (3.0 * t2 + a2 * t0 + b) - d, den = t2 * (4.0 * t0 + a3) + b2 * t0 + c;
        t = num / den;
        if (fabs (t - t0) < 1.0e-4) break;
        if (--it == 0) return -1;
        t0 = t;
     }
   x0[0] = 1.0 * x0[0] + t * p[0]; x0[1] = 1.0 * x0[1] + t * p[1]; x0[2] = 1.0 * x0[2] + t * p[2];
   return 0;
}
static int facet_diffract (const gshell_t *g, double theta, oracle_photon *at, int order, rng_t *r)   /* diffract.c:706-823 */
{
   double nl = order * (2.0 * PI * HBAR_C) / g->prm[1] / at->energy;
   double n[3], l[3], d[3], dth, dpp, pd, pl, pn, p[3], f, *x = at->x;
   n[0] = x[0]; n[1] = x[1]; n[2] = x[2]; norm3 (n); n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2];
   l[0] = n[2]; l[1] = 0.0; l[2] = -n[0]; norm3 (l);
   cross3 (n, l, d);
   if (g->nsectors)
     {
        uint32_t ns = g->nsectors, k; double sector = atan2 (x[1], x[2]);
        if (sector < 0) sector = 2 * PI + sector;
        k = bsearch_d (sector, g->sectors, ns, 6);
        if (k == 0) return -1;
        k--;
        if ((g->sectors[6 * k + 1] <= sector) || (g->sectors[6 * k] > sector)) return -1;
        dth = g->sectors[6 * k + 2] + g->sectors[6 * k + 3] * rng_gauss (r);
        dpp = g->sectors[6 * k + 4] + g->sectors[6 * k + 5] * rng_gauss (r);
     }
   else { dth = g->prm[3] * rng_gauss (r); dpp = g->prm[2] * rng_gauss (r); }
   theta -= dth;
   if (theta != 0.0)
     {
        double c = cos (theta), s = sin (theta), lt[3], dt[3]; int k;
        for (k = 0; k < 3; k++) { lt[k] = l[k]; dt[k] = d[k]; }
        for (k = 0; k < 3; k++) { l[k] = c * lt[k] + s * dt[k]; d[k] = -s * lt[k] + c * dt[k]; }
     }
   pd = nl + dot3 (at->p, d); pl = dot3 (at->p, l); pn = 1.0 - pl * pl - pd * pd;
   if (pn < 0.0) return -1;
   pn = sqrt (pn);
   p[0] = pd * d[0] + pn * n[0]; p[1] = pd * d[1] + pn * n[1]; p[2] = pd * d[2] + pn * n[2];
   at->p[0] = 1.0 * p[0] + pl * l[0]; at->p[1] = 1.0 * p[1] + pl * l[1]; at->p[2] = 1.0 * p[2] + pl * l[2];
   if (dpp == 0) return 0;
   f = nl * dpp;
   { double g2 = f * (pd / pn); at->p[0] += -f * d[0] + g2 * n[0]; at->p[1] += -f * d[1] + g2 * n[1]; at->p[2] += -f * d[2] + g2 * n[2]; }
   norm3 (at->p);
   return 0;
}
/* diffract_photon_from_grating with the per-photon row of JDMinterpolate_n_fvector (finterpo.c:134-189) */
static int order_and_diffract (const gshell_t *g, double theta, oracle_photon *at, int8_t *order_out, rng_t *r, int use_sectors)
{
   double u = rng_uniform (r), xe = (double) (float) at->energy, x0, x1, dx; uint32_t c = 1, k;
   gshell_t tmp = *g;
   if (!use_sectors) tmp.nsectors = 0;
   while ((c < g->nenergies - 1) && (xe > g->energies[c])) c++;
   x0 = g->energies[c - 1]; x1 = g->energies[c]; dx = x1 - x0;
   for (k = 0; k < g->norders; k++)
     {
        float ce;
        if (dx == 0.0) ce = g->cum_eff[(size_t) k * g->nenergies + c - 1];
        else
          {
             double y0 = g->cum_eff[(size_t) k * g->nenergies + c - 1], y1 = g->cum_eff[(size_t) k * g->nenergies + c];
             ce = (float) (y0 + (y1 - y0) * (xe - x0) / dx);
          }
        if (u <= ce) { *order_out = (int8_t) g->orders[k]; return facet_diffract (&tmp, theta, at, g->orders[k], r); }
     }
   return -1;
}

static void stage_grating (oracle_t *o, uint64_t n, oracle_photon *ph)
{
   uint64_t i;
   if ((int) o->grat[0] == 0) return;
   for (i = 0; i < n; i++)
     {
        oracle_photon *at = ph + i; rng_t r; const gshell_t *g; int rc = -1;
        if (at->flags & 0xFF) continue;
        g = &o->gshell[at->mirror_shell];
        rng_set (&r, o->seed, ray_of (o, at->tag), 2);
        if (rng_uniform (&r) > g->prm[4]) { at->flags |= F_VBLOCKED; continue; }                        /* :994-1001 */
        rot_x (at->x, -1 * g->prm[0]); rot_x (at->p, -1 * g->prm[0]);                                   /* :1013 */
        if (-1 == torus_hit (at->x, at->p, o->grat[1 + at->mirror_shell])) { at->flags |= F_UNDIFFRACTED; continue; }
        rc = order_and_diffract (g, 0.0, at, &at->order, &r, 1);                                         /* :1024-1079 */
        if (rc == -1) { at->flags |= F_UNDIFFRACTED; continue; }
        if ((int) o->grat[0] == 2)                                                                       /* LETG support gratings, :1083-1123 */
          {
             static const double pass_theta[4] = {PI / 2.0, PI / 3.0, 2.0 * PI / 3.0, 0.0};
             int pass;
             for (pass = 0; pass < 4; pass++)
               {
                  const gshell_t *sg = &o->support[pass == 0 ? 0 : 1];
                  if (sg->norders == 0) continue;
                  if (-1 == order_and_diffract (sg, pass_theta[pass], at, &at->support_orders[pass], &r, 0))
                    { at->flags |= F_UNDIFFRACTED; break; }
               }
             if (at->flags & 0xFF) continue;
          }
        rot_x (at->x, 1 * g->prm[0]); rot_x (at->p, 1 * g->prm[0]);                                     /* :1127 */
     }
}

/* ------------------------------------------------------------------------------------------- */
/* stage 3: _marx_acis_s_detect (acis-s.c:177-248)                                                */
typedef struct { float amp, center, sigma, cum_area; int tail; } gauss_t;                                /* acis_fef.c:87-96 */

static double contamination (const chip_t *ch, double en, double cx, double cy)                          /* aciscontam.c:93-187 */
{
   const double *c = ch->contam; uint32_t nl = (uint32_t) c[0], i; int mode = (int) c[1]; double v = 0.0, fxy;
   if (nl == 0) return 1.0;
   if (mode != 0)
     {
        if (mode == 1) { double dx = cx - c[2], dy = cy - c[3], rr = (8.0 / 1024.0) * sqrt (dx * dx + dy * dy); rr /= 8.07; fxy = 1.29 * rr * rr; }
        else if (cy <= 512.0) fxy = pow (fabs ((cy - 512.0) / (64.0 - 512.0)), 5.5);
        else fxy = pow (fabs ((cy - 512.0) / (964.0 - 512.0)), 4.5);
        for (i = 0; i < nl; i++) { double mu = interp_f (en, ch->ce[i], ch->cmu[i], ch->ncmu[i]); v += mu * (c[5 + i] + c[10 + i] * fxy); }
     }
   else
     {
        uint32_t blk = (uint32_t) c[4], ofs;
        if ((cx < 0) || (cx >= 1024) || (cy < 0) || (cy >= 1024)) return 0.0;
        cx /= blk; cy /= blk;
        ofs = (1024 / blk) * (unsigned int) cy + (unsigned int) cx;
        for (i = 0; i < nl; i++) { double mu = interp_f (en, ch->ce[i], ch->cmu[i], ch->ncmu[i]); fxy = ch->cfxy[i][ofs]; v += mu * (c[5 + i] + c[10 + i] * fxy); }
     }
   return exp (-v);
}
static double gauss_integral (double xmin, double xmax, const gauss_t *g)                                /* acis_fef.c:380-392 */
{
   double sigma = g->sigma * 1.4142135623730951, x0;
   if (sigma == 0.0) return 0.0;
   x0 = g->center;
   return 0.5 * g->amp * (erf ((xmax - x0) / sigma) - erf ((xmin - x0) / sigma)) * (2.5066282746310002 * g->sigma);
}
static int normalize_gaussians (gauss_t *g, uint32_t num)                                                /* acis_fef.c:509-579 */
{
   double pos = 0.0, neg = 0.0; uint32_t k; int flags = 0;
   for (k = 0; k < num; k++)
     {
        double a1 = gauss_integral (0, 1e37, g + k), a2 = gauss_integral (-1e37, 0, g + k);
        g[k].tail = 0;
        if (a2 > a1) { double ratio = a1 / a2; if (ratio < 0.1) g[k].tail = 1; }
        if (a1 >= 0) pos += a1; else neg -= a1;
        g[k].cum_area = pos;
     }
   if (pos <= neg) flags |= 4;
   if (neg != 0.0) flags |= 1;
   if (pos > 0) for (k = 0; k < num; k++) g[k].cum_area /= pos;
   return flags;
}
static int pha_pos (const gauss_t *g, uint32_t num, double *phap, rng_t *r)                              /* acis_fef.c:408-468 */
{
   int guard;
   for (guard = 0; guard < 4096; guard++)
     {
        double u = rng_uniform (r); uint32_t k;
        for (k = 0; k < num; k++)
          {
             double pha;
             if (g[k].cum_area <= u) continue;
             if (g[k].tail == 0)
               { unsigned int count = 0; do { pha = g[k].center + g[k].sigma * rng_gauss (r); count++; } while ((pha < 0) && (count < 100)); }
             else
               {
                  double uu, v, x, s = (0 - g[k].center) / g[k].sigma;
                  do { uu = rng_uniform (r); do v = rng_uniform (r); while (v == 0.0); x = sqrt (s * s - 2 * log (v)); } while (x * uu > s);
                  pha = g[k].center + x * g[k].sigma;
               }
             if (pha < 0) break;
             *phap = pha;
             return 0;
          }
     }
   return -1;
}
static int pha_neg (const gauss_t *g, uint32_t num, double *phap, rng_t *r)                              /* acis_fef.c:471-503 */
{
   int count = 0;
   while (count < 100)
     {
        double pha, pos = 0.0, sum = 0.0; uint32_t k;
        if (-1 == pha_pos (g, num, &pha, r)) return -1;
        for (k = 0; k < num; k++)
          {
             double sigma = g[k].sigma, ds = 0.0;
             if (sigma != 0.0) { double xx = (pha - g[k].center) / sigma; ds = g[k].amp * exp (-0.5 * xx * xx); }
             sum += ds; if (ds > 0) pos += ds;
          }
        if (rng_uniform (r) * pos < sum) { *phap = pha; return 0; }
        count++;
     }
   return -1;
}
static int apply_fef (oracle_t *o, const chip_t *ch, float x, float y, double energy, float *pip, int16_t *phap, rng_t *r)   /* acis_fef.c:910-1079 */
{
   unsigned int i, j, k; const fef_t *f; int fi, flags, status; double t, pha; const float *g0, *g1; gauss_t G[MAX_GAUSS];
   if ((x < 0) || (x >= 1024) || (y < 0) || (y >= 1024))                 /* find_fef :917-941: clamp with DetExtendFlag=yes */
     {
        if (0 == (int) o->acis[15]) return -1;
        if (x < 0) x = 0; else if (x >= 1024) x = 1023;
        if (y < 0) y = 0; else if (y >= 1024) y = 1023;
     }
   i = (unsigned int) (x / 32); j = (unsigned int) (y / 32);
   if ((i >= 32) || (j >= 32)) return -1;
   fi = ch->fef_map[i * 32 + j];
   if (fi < 0) return -1;
   f = &o->fefs[fi];
   if (f->ng > MAX_GAUSS) return -1;
   i = bsearch_f (energy, f->energies, f->ne);
   if (i == 0) i++;
   if (i == f->ne) i--;
   t = (energy - f->energies[i - 1]) / (f->energies[i] - f->energies[i - 1]);
   g0 = f->gauss + (size_t) (i - 1) * f->ng * 3; g1 = g0 + (size_t) f->ng * 3;
   for (k = 0; k < f->ng; k++)
     {
        float a0 = g0[3 * k], c0 = g0[3 * k + 1], s0 = g0[3 * k + 2], a1 = g1[3 * k], c1 = g1[3 * k + 1], s1 = g1[3 * k + 2]; double v;
        G[k].center = c0 + t * (c1 - c0);
        v = s0 + t * (s1 - s0);
        if (v <= 0.0) { G[k].sigma = 0.0; G[k].amp = 0.0; }
        else { G[k].sigma = v; v = a0 + t * (a1 - a0); if ((v < 0.0) && ((a1 > 0.0) || (a0 > 0.0))) v = 0.0; G[k].amp = v; }
     }
   flags = normalize_gaussians (G, f->ng);
   if (flags & 4) return -1;
   status = (flags == 0) ? pha_pos (G, f->ng, &pha, r) : pha_neg (G, f->ng, &pha, r);
   if (status == -1) return -1;
   *phap = (short) pha;
   pha = *phap - rng_uniform (r);
   *pip = interp_f (pha, f->channels, f->energies, f->ne);
   if (*pip < 0) return -1;
   return 0;
}
/* marx_map_energy_to_acis_pha, acis_fef.c:1087-1096 (find_fef :910-966 + JDMinterpolate_f): the deterministic energy -> PHA map
 * marxpileup uses for its summed island energies (oracle/pileup_oracle.c).  x, y arrive as ints (the prototype truncates). */
int oracle_map_energy_to_acis_pha (oracle_t *o, int ccd_id, int xi, int yi, double energy, short *phap)
{
   float x = (float) xi, y = (float) yi; unsigned int i, j; int k, fi; const chip_t *ch = NULL; const fef_t *f;
   if ((x < 0) || (x >= 1024) || (y < 0) || (y >= 1024))
     {
        if (0 == (int) o->acis[15]) return -1;
        if (x < 0) x = 0; else if (x >= 1024) x = 1023;
        if (y < 0) y = 0; else if (y >= 1024) y = 1023;
     }
   i = (unsigned int) (x / 32); j = (unsigned int) (y / 32);
   if ((i >= 32) || (j >= 32)) return -1;
   for (k = 0; k < o->nchips; k++) if ((int) o->chip[k].geom[0] == ccd_id) ch = &o->chip[k];
   if (ch == NULL) return -1;
   fi = ch->fef_map[i * 32 + j];
   if (fi < 0) return -1;
   f = &o->fefs[fi];
   *phap = (short) interp_f ((float) energy, f->energies, f->channels, f->ne);
   return 0;
}
/* _marx_dither_detector / _marx_undither_detector, detector.c:240-295: the ONE global transform is modified and restored
 * for every photon that reaches the detector, so it drifts by rounding exactly as the reference's does */
static void xf_rotate (double *m, double theta)
{
   double m00, m01, m10, m11, c, s;
   if (theta == 0) { s = 0; c = 1; } else { s = sin (theta); c = cos (theta); }
   m00 = m[4]; m01 = m[5]; m10 = m[7]; m11 = m[8];
   m[4] = m00 * c - m01 * s; m[5] = m00 * s + m01 * c;
   m[7] = m10 * c - m11 * s; m[8] = m10 * s + m11 * c;
}
static void xf_begin (oracle_t *o, const double *off, const double *M)
{
   int k;
   if (o->xf_init) return;
   for (k = 0; k < 3; k++) o->xf_off[k] = off[k];
   for (k = 0; k < 9; k++) o->xf_m[k] = M[k];
   o->xf_init = 1;
}
static void xf_dither (oracle_t *o, const oracle_photon *at)
{
   if ((int) o->dith[0] == 0) return;
   o->xf_off[1] += at->dither[3]; o->xf_off[2] += at->dither[4];
   xf_rotate (o->xf_m, at->dither[5]);
}
static void xf_undither (oracle_t *o, const oracle_photon *at)
{
   if ((int) o->dith[0] == 0) return;
   o->xf_off[1] -= at->dither[3]; o->xf_off[2] -= at->dither[4];
   xf_rotate (o->xf_m, -at->dither[5]);
}
static int plane_hit (const double *g, const double *x0, const double *p, double *x, double *dx, double *dy, int must_hit)   /* detector.c:56-109 */
{
   const double *xll = g + 1, *xhat = g + 4, *yhat = g + 7, *nrm = g + 10; double pdn = dot3 (p, nrm), r[3], f, rx, ry; int hit = 1;
   if (pdn == 0) return -1;
   r[0] = x0[0] - xll[0]; r[1] = x0[1] - xll[1]; r[2] = x0[2] - xll[2];
   f = -1.0 * dot3 (r, nrm) / pdn;
   r[0] = 1.0 * r[0] + f * p[0]; r[1] = 1.0 * r[1] + f * p[1]; r[2] = 1.0 * r[2] + f * p[2];
   rx = dot3 (r, xhat); if ((rx < 0.0) || (rx >= g[13])) { if (must_hit) return 0; hit = 0; }
   ry = dot3 (r, yhat); if ((ry < 0.0) || (ry >= g[14])) { if (must_hit) return 0; hit = 0; }
   x[0] = r[0] + xll[0]; x[1] = r[1] + xll[1]; x[2] = r[2] + xll[2];
   *dx = rx; *dy = ry;
   return hit;
}
/* _marx_intersect_with_detector, detector.c:111-168; geom[k] = the packed facet k (stride in doubles via accessor) */
static int detector_hit (const double *const *geom, int n, const double *x0, const double *p, double *xh, double *dx, double *dy, int extend)
{
   int k, best = -1; double best_r2 = -1, bx[3] = {0, 0, 0}, bdx = 0, bdy = 0;
   for (k = 0; k < n; k++) if (1 == plane_hit (geom[k], x0, p, xh, dx, dy, 1)) return k;
   if (!extend) return -1;
   for (k = 0; k < n; k++)
     {
        double x[3], ddx, ddy, deltax, deltay, r2;
        if (-1 == plane_hit (geom[k], x0, p, x, &ddx, &ddy, 0)) continue;
        deltax = ddx - 0.5 * geom[k][13]; deltay = ddy - 0.5 * geom[k][14];
        r2 = deltax * deltax + deltay * deltay;
        if ((r2 < best_r2) || (best < 0)) { best = k; bx[0] = x[0]; bx[1] = x[1]; bx[2] = x[2]; bdx = ddx; bdy = ddy; best_r2 = r2; }
     }
   if (best < 0) return -1;
   xh[0] = bx[0]; xh[1] = bx[1]; xh[2] = bx[2]; *dx = bdx; *dy = bdy;
   return best;
}
/* stage 3 (HRC-S / HRC-I): _marx_drake_reflect (drake.c:317-372) + _marx_hrc_s_detect (hrc-s.c:236-312) / _marx_hrc_i_detect (hrc-i.c:119-186) */
static void stage_detect_hrc (oracle_t *o, uint64_t n, oracle_photon *ph)
{
   const double *H = o->hrc; const double *off = o->xf_off, *M = o->xf_m, *S = H + 16, *B = H + 26;
   int ideal = (int) H[14], extend = (int) H[15], use_hesf = (int) H[41], nplates = (int) H[42];
   int hrc_i = ((int) H[0] == 2);                                       /* MARX_DETECTOR_HRC_I (hrc-i.c:119-186) */
   double upix = H[39], vpix = H[40], crw = H[43];
   uint64_t i;
   xf_begin (o, H + 2, H + 5);
   for (i = 0; i < n; i++)
     {
        oracle_photon *at = ph + i; rng_t r; int k, hit = -1; double dx = 0, dy = 0, xh[3]; const double *g = NULL; const mcp_t *m;
        double t, y, z, u, v; int region;
        if (at->flags & 0xFF) continue;
        rng_set (&r, o->seed, ray_of (o, at->tag), 3);
        if (use_hesf)
          for (k = 0; k < 2 * nplates; k++)                                                              /* drake.c:268-313 */
            {
               const double *q = o->hesf + 14 * k; double pdn = dot3 (at->p, q + 9), dd[3], nx[3], xp[3], xx, yy, rfl = 1.0; int cr;
               if (0.0 == pdn) continue;
               dd[0] = q[0] - at->x[0]; dd[1] = q[1] - at->x[1]; dd[2] = q[2] - at->x[2];
               t = dot3 (dd, q + 9) / pdn;
               nx[0] = 1.0 * at->x[0] + t * at->p[0]; nx[1] = 1.0 * at->x[1] + t * at->p[1]; nx[2] = 1.0 * at->x[2] + t * at->p[2];
               xp[0] = nx[0] - q[0]; xp[1] = nx[1] - q[1]; xp[2] = nx[2] - q[2];
               xx = dot3 (xp, q + 3); if ((xx < 0.0) || (xx >= q[12])) continue;
               yy = dot3 (xp, q + 6); if ((yy < 0.0) || (yy >= q[13])) continue;
               cr = (xx < crw);
               at->x[0] = nx[0]; at->x[1] = nx[1]; at->x[2] = nx[2];
               if (cr ? o->nhcr : o->nhc)                                                                /* reflect.c:81-92 */
                 {
                    float b = interp_f (at->energy, cr ? o->hcr_e : o->hc_e, cr ? o->hcr_b : o->hc_b, cr ? o->nhcr : o->nhc);
                    float d = interp_f (at->energy, cr ? o->hcr_e : o->hc_e, cr ? o->hcr_d : o->hc_d, cr ? o->nhcr : o->nhc);
                    rfl = reflectivity (fabs (pdn), b, d);
                 }
               if (rfl < rng_uniform (&r)) { at->flags |= 0x20; break; }
               { double f = -2.0 * pdn; at->p[0] = 1.0 * at->p[0] + f * q[9]; at->p[1] = 1.0 * at->p[1] + f * q[10]; at->p[2] = 1.0 * at->p[2] + f * q[11]; }
               at->flags |= 0x100;
               break;
            }
        if (at->flags & 0xFF) continue;
        xf_dither (o, at);                                                                               /* hrc-s.c:264-270 */
        at->x[0] -= off[0]; at->x[1] -= off[1]; at->x[2] -= off[2];
        mat3 (M, at->x); mat3 (M, at->p);
        { const double *gg[3]; for (k = 0; k < o->nmcps; k++) gg[k] = o->mcp[k].geom; hit = detector_hit (gg, o->nmcps, at->x, at->p, xh, &dx, &dy, extend); }
        if (hit < 0) { at->flags |= F_MISSED; at->ccd_num = -1; goto undither; }
        m = &o->mcp[hit]; g = m->geom;
        at->x[0] = xh[0]; at->x[1] = xh[1]; at->x[2] = xh[2];
        /* apply_hrc_qe, hrc-s.c:192-234 */
        if (m->nqe && (rng_uniform (&r) >= interp_f ((float) at->energy, m->qe_e, m->qe, m->nqe))) { at->flags |= F_UNDETECTED; at->ccd_num = -1; goto undither; }
        if (hrc_i) region = 0;                                                                           /* hrc-i.c:88-117: one UVIS filter */
        else
        {
        t = (S[3] - at->x[0]) / at->p[0]; y = at->x[1] + t * at->p[1]; z = at->x[2] + t * at->p[2];
        y -= S[8]; z -= S[9];                                                                            /* get_filter_region :136-188 */
        {
           double o1, sl, slg;
           if (y < 0) { y = -y; o1 = S[1]; sl = S[4]; slg = S[6]; } else { o1 = S[2]; sl = S[5]; slg = S[7]; }
           if (y < o1) region = 0;
           else if (y < sl) region = (z >= S[0]) ? 0 : 1;
           else if (y >= slg) region = (z >= S[0]) ? 2 : 3;
           else region = -1;
        }
        }
        if (region < 0) { at->flags |= F_UNDETECTED; at->ccd_num = -1; goto undither; }
        if (o->nfilt[region] && (rng_uniform (&r) >= interp_f ((float) at->energy, o->filt_e[region], o->filt_q[region], o->nfilt[region])))
          { at->flags |= F_UNDETECTED; at->ccd_num = -1; goto undither; }
        at->detector_region = (int8_t) region;
        {                                                                                                /* hrc-i.c:66-85 */
           double e = at->energy;
           if (e <= 0.5) e = 141.582 * sqrt (e); else if (e < 2.0) e = 107.299 * pow (e, 0.1); else e = 115.0;
           e = e * (1.0 + 0.424661 * rng_gauss (&r));
           if (e < 0.0) e = 0.0;
           at->pulse_height = (short) e;
        }
        if (!ideal)                                                                                      /* hrcblur.c:260-298 */
          {
             double rr = rng_uniform (&r), x_0, y_0, th;
             if (rr < B[3]) { double c; do c = rng_uniform (&r); while (c == 0.0); rr = B[0] * sqrt (-2 * log (c)); x_0 = B[1]; y_0 = B[2]; }
             else if (rr < B[3] + B[7]) { double c; do c = rng_uniform (&r); while (c == 0.0); rr = B[4] * sqrt (-2 * log (c)); x_0 = B[5]; y_0 = B[6]; }
             else { double c = rng_uniform (&r), rmax = B[11] / B[8]; rr = B[8] * sqrt (expm1 (c * log1p (rmax * rmax))); x_0 = B[9]; y_0 = B[10]; }
             th = (2.0 * PI) * rng_uniform (&r);
             dx += x_0 + rr * cos (th); dy += y_0 + rr * sin (th);
             if (!extend) { if (dx < 0.0) dx = 0.0; if (dy < 0.0) dy = 0.0; }
          }
        at->ccd_num = (int8_t) g[0];
        if (hrc_i)                                                                                       /* hrc_i_geom.c:146-156 */
          { double px = dx / upix, py = dy / vpix; at->y_pixel = g[15] + px; at->z_pixel = g[16] + py; }
        else
          {
        u = g[15] + dx / upix; v = g[16] + dy / vpix;                                                    /* hrc_s_geom.c:344-394 */
        at->u_pixel = u; at->v_pixel = v;
        at->y_pixel = g[19] + (u - g[17]); at->z_pixel = g[20] + (v - g[18]);
          }
        mat3t (M, at->p); mat3t (M, at->x);
        at->x[0] += off[0]; at->x[1] += off[1]; at->x[2] += off[2];
      undither:
        xf_undither (o, at);
     }
}

static void stage_detect (oracle_t *o, uint64_t n, oracle_photon *ph)
{
   const double *A = o->acis; const double *off = o->xf_off, *M = o->xf_m;
   int ideal = (int) A[14], extend = (int) A[15]; double focal = A[16], texp = A[17], tft = A[18], tframe = A[19];
   uint64_t i;
   if ((int) A[0] == 0) return;
   xf_begin (o, A + 2, A + 5);
   for (i = 0; i < n; i++)
     {
        oracle_photon *at = ph + i; rng_t r; int k, hit = -1; double dx = 0, dy = 0, xh[3]; const chip_t *ch; const double *g;
        if (at->flags & 0xFF) continue;
        rng_set (&r, o->seed, ray_of (o, at->tag), 3);
        xf_dither (o, at);                                                                               /* acis-s.c:209-211 */
        at->x[0] -= off[0]; at->x[1] -= off[1]; at->x[2] -= off[2];                                      /* trans.c:66-77 */
        mat3 (M, at->x); mat3 (M, at->p);
        { const double *gg[MAX_CHIPS]; for (k = 0; k < o->nchips; k++) gg[k] = o->chip[k].geom; hit = detector_hit (gg, o->nchips, at->x, at->p, xh, &dx, &dy, extend); }
        if (hit < 0) { at->flags |= F_MISSED; at->ccd_num = -1; goto undither; }
        ch = &o->chip[hit]; g = ch->geom;
        at->x[0] = xh[0]; at->x[1] = xh[1]; at->x[2] = xh[2];
        at->ccd_num = (int8_t) g[0];
        at->y_pixel = dx / g[15]; at->z_pixel = dy / g[16];                                              /* acis-s.c:229-230 */
        if (!ideal)
          {
             double u = rng_uniform (&r), qe, qf, qc;
             qe = ch->nqe ? interp_f (at->energy, ch->qe_e, ch->qe, ch->nqe) : 1.0;
             qf = ch->nfl ? interp_f (at->energy, ch->fl_e, ch->fl, ch->nfl) : 1.0;
             qc = contamination (ch, at->energy, at->y_pixel, at->z_pixel);
             if (u >= qe * qf * qc) { at->flags |= F_UNDETECTED; goto undither; }
          }
        if (-1 == apply_fef (o, ch, at->y_pixel, at->z_pixel, at->energy, &at->pi, &at->pulse_height, &r))
          { at->pulse_height = -1; at->pi = 0; at->flags |= F_UNDETECTED; goto undither; }
        if (tft > 0.0)                                                                                   /* acis-i.c:60-89 */
          {
             double t = fmod (at->arrival_time, tframe);
             if (t > texp)
               {
                  double xp, yp;
                  at->z_pixel = 1.0 + 1022.0 * rng_uniform (&r);
                  xp = (at->y_pixel - g[17]) * g[15]; yp = (at->z_pixel - g[18]) * g[16];
                  at->x[0] = g[1] + (xp * g[4] + yp * g[7]); at->x[1] = g[2] + (xp * g[5] + yp * g[8]); at->x[2] = g[3] + (xp * g[6] + yp * g[9]);
                  at->p[0] = at->x[0] - focal; at->p[1] = at->x[1] - 0; at->p[2] = at->x[2] - 0;
                  norm3 (at->p);
                  at->flags |= F_STREAKED;
               }
          }
        mat3t (M, at->p); mat3t (M, at->x);                                                              /* trans.c:79-90 */
        at->x[0] += off[0]; at->x[1] += off[1]; at->x[2] += off[2];
      undither:
        xf_undither (o, at);
     }
}

/* ------------------------------------------------------------------------------------------- */
long oracle_trace (oracle_t *o, uint64_t first_ray, uint64_t n, double *time_base,
                   oracle_photon *st0, oracle_photon *st1, oracle_photon *st2, oracle_photon *st3)
{
   oracle_photon *work; uint64_t i; long detected = 0; double tb = time_base ? *time_base : 0.0;
   if ((o == NULL) || (n == 0)) return -1;
   work = (oracle_photon *) malloc (n * sizeof (oracle_photon));
   if (work == NULL) return -1;
   {
      /* the ASPSOL reader only moves forward (dither.c:361-369); a trace that restarts the clock rewinds it */
      uint64_t kept;
      if (o->asp && ((o->asp_pos < 1) || (tb < o->asp_t_prev))) o->asp_pos = 1;
      o->batch_first = first_ray;
      kept = stage_source (o, first_ray, n, &tb, work);
      o->asp_t_prev = tb;
      o->last_generated = kept;
      if (kept < n)
        {
           /* the slots behind the cut stay zero in every stage record */
           memset (work + kept, 0, (n - kept) * sizeof (oracle_photon));
           if (st0) memset (st0 + kept, 0, (n - kept) * sizeof (oracle_photon));
           if (st1) memset (st1 + kept, 0, (n - kept) * sizeof (oracle_photon));
           if (st2) memset (st2 + kept, 0, (n - kept) * sizeof (oracle_photon));
           if (st3) memset (st3 + kept, 0, (n - kept) * sizeof (oracle_photon));
           n = kept;
        }
   }
   if (time_base) *time_base = tb;
   if (n == 0) { free (work); return 0; }
   if (st0) memcpy (st0, work, n * sizeof (oracle_photon));
   stage_mirror (o, n, work);
   if (st1) memcpy (st1, work, n * sizeof (oracle_photon));
   stage_grating (o, n, work);
   if (st2) memcpy (st2, work, n * sizeof (oracle_photon));
   if (o->hrc) stage_detect_hrc (o, n, work); else stage_detect (o, n, work);
   if (st3) memcpy (st3, work, n * sizeof (oracle_photon));
   for (i = 0; i < n; i++) if ((work[i].flags & 0xFF) == 0) detected++;
   free (work);
   return detected;
}
