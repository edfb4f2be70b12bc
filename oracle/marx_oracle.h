/* marx_oracle.h -- CPU restatement (plain C) of the reference's per-photon ray-trace path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link, load or call it; the product (libmarxb200.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the restatement is
 * pinned against outputs of the reference itself: oracle/_ref/marx_replay drives the unmodified MARX
 * 5.5.3 stage functions with the same counter-based draws, and tests/test_oracle_vs_reference.py
 * requires this file to reproduce its per-stage FP64 photon records bit for bit on the committed
 * fixtures (tests/golden/) and, where oracle/_ref exists, on freshly generated replays.
 *
 * Each function cites the reference file:line it follows (paths relative to the MARX tree).
 */
#ifndef MARX_ORACLE_H
#define MARX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as Marx_Photon_Attr_Type, marx/libsrc/marx.h:51-100 (136 bytes) */
typedef struct
{
   double energy;
   double x[3];
   double p[3];
   double arrival_time;
   uint32_t flags;
   float y_pixel, z_pixel, u_pixel, v_pixel;
   float dither[6];
   float pi;
   int16_t pulse_height;
   uint32_t mirror_shell;
   int8_t ccd_num, detector_region, order, support_orders[4];
   uint32_t tag;
}
oracle_photon;

typedef struct oracle oracle_t;

/* load the calibration pack (the post-init tables of the reference modules) */
oracle_t *oracle_open (const char *calpack_path, uint64_t seed);
void oracle_close (oracle_t *o);
/* marx_map_energy_to_acis_pha, acis_fef.c:1087-1096 */
int oracle_map_energy_to_acis_pha (oracle_t *o, int ccd_id, int x, int y, double energy, short *phap);
const char *oracle_last_error (void);

/* Trace rays [first_ray, first_ray + n).  st[s] (s = 0..3, each n records or NULL) receives the photon
 * records after source / mirror / grating / detector, dead rays included (flags say why, first cause).
 * arrival_time is ABSOLUTE (reference: pt->start_time + arrival_time); *time_base is the running time,
 * updated to the time of the last ray.  Returns the number of detected photons, or -1. */
long oracle_trace (oracle_t *o, uint64_t first_ray, uint64_t n, double *time_base,
                   oracle_photon *st0, oracle_photon *st1, oracle_photon *st2, oracle_photon *st3);

/* DitherModel=FILE only: the number of rays the last oracle_trace kept -- its n, or fewer when the ASPSOL file ended
 * inside the batch (dither.c:296-301; the reference ends the simulation there).  Slots behind the cut are zero. */
uint64_t oracle_last_generated (const oracle_t *o);

#ifdef __cplusplus
}
#endif
#endif
