/* See philox_rng.h.  TEST INFRASTRUCTURE (oracle/_ref build only). */
#include <math.h>
#include <stdlib.h>
#include <jdmath.h>
#include "philox_rng.h"

static uint64_t Seed;
static uint64_t Ray;
static uint32_t Stage;
static uint32_t Draw;
static uint32_t Block[4];
static uint32_t Block_Index = 0xFFFFFFFFu;
static int Have_Spare;
static double Spare;

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

void replay_rng_seed (uint64_t seed) { Seed = seed; }
void replay_rng_set (uint64_t ray, uint32_t stage)
{
   Ray = ray; Stage = stage; Draw = 0; Block_Index = 0xFFFFFFFFu; Have_Spare = 0;
}
uint32_t replay_rng_draws (void) { return Draw; }

static uint32_t next_u32 (void)
{
   uint32_t b = Draw >> 2;
   if (b != Block_Index)
     {
	Block[0] = (uint32_t) Ray; Block[1] = (uint32_t) (Ray >> 32);
	Block[2] = b; Block[3] = Stage;
	philox4x32_10 (Block, (uint32_t) Seed, (uint32_t) (Seed >> 32));
	Block_Index = b;
     }
   return Block[(Draw++) & 3];
}

/* --- the jdmath entry points (jdmath/src/random.c, gaussrnd.c) --- */
struct _JDMRandom_Type { int unused; };
uint32 JDMgenerate_uint32_random (JDMRandom_Type *rt) { (void) rt; return next_u32 (); }
double JDMgenerate_random (JDMRandom_Type *rt)
{ (void) rt; return (double) next_u32 () * (1.0 / (double) (uint32) 0xFFFFFFFFU); }
uint32 JDMuint32_random (void) { return next_u32 (); }
double JDMrandom (void) { return (double) next_u32 () * (1.0 / (double) (uint32) 0xFFFFFFFFU); }
int JDMseed_random (JDMRandom_Type *rt, unsigned long s) { (void) rt; Seed = s; return 0; }
int JDMsrandom (unsigned long s) { Seed = s; return 0; }
JDMRandom_Type *JDMcreate_random (void) { return (JDMRandom_Type *) calloc (1, sizeof (JDMRandom_Type)); }
void JDMfree_random (JDMRandom_Type *r) { free (r); }
uint32 JDMfast_uint32_random (void) { return next_u32 (); }
void JDMseed_fast_random (unsigned long s) { (void) s; }
double JDMfast_random (void) { return JDMrandom (); }

double JDMgaussian_random (void)
{
   double g1, g2, g, s;
   if (Have_Spare) { Have_Spare = 0; return Spare; }
   do
     {
	g1 = 2.0 * JDMrandom () - 1.0;
	g2 = 2.0 * JDMrandom () - 1.0;
	g = g1 * g1 + g2 * g2;
     }
   while ((g >= 1.0) || (g == 0.0));
   s = sqrt (-2.0 * log (g) / g);
   Spare = g2 * s; Have_Spare = 1;
   return g1 * s;
}

double JDMexpn_random (void)
{
   double r;
   do r = JDMrandom (); while (r == 0.0);
   return -log (r);
}
