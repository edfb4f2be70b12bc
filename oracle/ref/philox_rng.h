/* Counter-based replacement for jdmath's global RNG, linked AHEAD of libjdmath.a so that
 * random.o / gaussrnd.o are never pulled from the archive.  TEST INFRASTRUCTURE.
 *
 * Draw k of (ray, stage) = lane (k & 3) of Philox4x32-10 with
 *    key     = (seed_lo, seed_hi)
 *    counter = (ray_lo, ray_hi, k >> 2, stage)
 * mapped to a double exactly as jdmath/src/random.c:151-154 does: u32 * (1/4294967295.0),
 * i.e. uniform on [0,1] INCLUSIVE.  JDMgaussian_random / JDMexpn_random keep the algorithms of
 * jdmath/src/gaussrnd.c:30-67; the cached Box-Muller spare is reset whenever (ray, stage) is set.
 * The same stream definition is used by oracle/marx_oracle.c and by the CUDA kernels.
 */
#ifndef ORACLE_PHILOX_RNG_H
#define ORACLE_PHILOX_RNG_H
#include <stdint.h>
enum { RNG_STAGE_SOURCE = 0, RNG_STAGE_MIRROR = 1, RNG_STAGE_GRATING = 2, RNG_STAGE_DETECTOR = 3 };
void replay_rng_seed (uint64_t seed);
void replay_rng_set (uint64_t ray, uint32_t stage);   /* also clears draw counter + gaussian spare */
uint32_t replay_rng_draws (void);                     /* uniforms consumed since last _set */
#endif
