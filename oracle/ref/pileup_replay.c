/* marxpileup_replay -- TEST INFRASTRUCTURE (oracle/_ref build only).
 * The UNMODIFIED marxpileup.c (compiled into this unit where it lies; nothing is copied) with a counter-based replacement for
 * jdmath's global RNG, so that its only draw -- will_grade_migrate (marxpileup.c:668-674), once per candidate island of >= 2
 * photons -- can be reproduced by a parallel implementation: the K-th call of JDMrandom while frame F is processed
 * (process_frame :890-922; frames arrive in increasing order, each once) is lane (K & 3) of
 * Philox4x32-10 (key = seed, counter = (F, 0, K >> 2, 5)), mapped to [0,1] as random.c:151-154 does.
 * The driver below is the loop of marxpileup.c:main (:1121-1213) with one added line that tells the stream which frame is being
 * processed.  Seed: PILEUP_SEED in the environment.  usage: exactly like marxpileup. */
#include <stdlib.h>
#include <stdint.h>

static uint64_t Replay_Seed;
static unsigned int Replay_Frame, Replay_K;

#define main marxpileup_stock_main
#include "marxpileup.c"
#undef main

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
   unsigned int k = Replay_K++;
   c[0] = Replay_Frame; c[1] = 0; c[2] = k >> 2; c[3] = 5u;
   philox4x32_10 (c, (uint32_t) Replay_Seed, (uint32_t) (Replay_Seed >> 32));
   return c[k & 3];
}
/* the whole of jdmath/src/random.c's interface, so that the archive member is never pulled in */
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

int main (int argc, char **argv)
{
   Input_Event_Type *event_list = NULL;
   unsigned int frame, last_frame = 0;
   const char *sd = getenv ("PILEUP_SEED");
   if (sd != NULL) Replay_Seed = strtoull (sd, NULL, 10);

   if (-1 == initialize (argc, argv)) return 1;
   if (-1 == open_marx_input_files ()) return 1;
   if (-1 == open_marx_output_files ()) { close_marx_input_files (); return 2; }
   while (1)
     {
	Input_Event_Type *evt;
	int status = read_input_event (&evt, &frame);
	if (status == -1) return 1;
	if ((status == 1) && (frame == last_frame))
	  {
	     evt->next = event_list;
	     event_list = evt;
	     continue;
	  }
	if ((event_list != NULL) || (status == 0))
	  {
	     Replay_Frame = last_frame; Replay_K = 0;          /* the added line */
	     if (-1 == process_frame (event_list, last_frame)) return 1;
	     free_event_list (event_list);
	     if (status == 0) break;
	  }
	event_list = evt;
	evt->next = NULL;
	last_frame = frame;
     }
   close_marx_input_files ();
   close_marx_output_files ();
   deallocate_buffers ();
   if (-1 == copy_files ()) return 1;
   fprintf (stdout, "Total Number Input: %u\nTotal Number Detected: %u\n", Num_Input, Num_Detected);
   return 0;
}
