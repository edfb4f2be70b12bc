/* Counter-based replacement for jdmath's global RNG inside the stock marx2fits (TEST INFRASTRUCTURE, oracle/_ref build
 * only; linked AHEAD of libjdmath.a like philox_rng.c).  marx2fits draws per event row, in this order:
 * compute_fltgrade 1 U (ACIS only, marx2fits.c:3864), compute_detxy 2 U (--pixadj=randomize only, :3707-3708), so the
 * number of draws per row is a constant of the run (L1_NDRAW in the environment).  Call number K of JDMrandom is therefore
 * draw K % NDRAW of row K / NDRAW = lane (d & 3) of Philox4x32-10 (key = seed, counter = (row_lo, row_hi, d >> 2, 4)):
 * the stream the CUDA Level-1 kernel and oracle/level1_oracle.c use (stage 4 = MARXB200_STAGE_LEVEL1). */
#include <stdlib.h>
#include <stdint.h>
#include <jdmath.h>

static uint64_t Seed, Calls;
static int Ndraw = -1;

static void philox4x32_10 (uint32_t c[4], uint32_t k0, uint32_t k1)
{
   int i;
   for (i = 0; i < 10; i++)
     {
	uint64_t p0 = (uint64_t) 0xD2511F53u * c[0];
	uint64_t p1 = (uint64_t) 0xCD9E8D57u * c[2];
	uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k0;
	uint32_t n1 = (uint32_t) p1;
	uint32_t n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k1;
	uint32_t n3 = (uint32_t) p0;
	c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
	k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
     }
}

static uint32_t next_u32 (void)
{
   uint32_t c[4];
   uint64_t row, d;
   if (Ndraw < 0)
     {
	const char *s = getenv ("L1_NDRAW"), *sd = getenv ("L1_SEED");
	Ndraw = (s != NULL) ? atoi (s) : 1;
	if (Ndraw < 1) Ndraw = 1;
	if (sd != NULL) Seed = strtoull (sd, NULL, 10);
     }
   row = Calls / (uint64_t) Ndraw; d = Calls % (uint64_t) Ndraw;
   Calls++;
   c[0] = (uint32_t) row; c[1] = (uint32_t) (row >> 32); c[2] = (uint32_t) (d >> 2); c[3] = 4u;
   philox4x32_10 (c, (uint32_t) Seed, (uint32_t) (Seed >> 32));
   return c[d & 3];
}

struct _JDMRandom_Type { int unused; };
uint32 JDMgenerate_uint32_random (JDMRandom_Type *rt) { (void) rt; return next_u32 (); }
double JDMgenerate_random (JDMRandom_Type *rt) { (void) rt; return (double) next_u32 () * (1.0 / (double) (uint32) 0xFFFFFFFFU); }
uint32 JDMuint32_random (void) { return next_u32 (); }
double JDMrandom (void) { return (double) next_u32 () * (1.0 / (double) (uint32) 0xFFFFFFFFU); }
int JDMseed_random (JDMRandom_Type *rt, unsigned long s) { (void) rt; (void) s; return 0; }
int JDMsrandom (unsigned long s) { (void) s; return 0; }
JDMRandom_Type *JDMcreate_random (void) { return (JDMRandom_Type *) calloc (1, sizeof (JDMRandom_Type)); }
void JDMfree_random (JDMRandom_Type *r) { free (r); }
uint32 JDMfast_uint32_random (void) { return next_u32 (); }
void JDMseed_fast_random (unsigned long s) { (void) s; }
double JDMfast_random (void) { return JDMrandom (); }
double JDMgaussian_random (void) { abort (); return 0; }   /* never drawn by marx2fits */
double JDMexpn_random (void) { abort (); return 0; }
