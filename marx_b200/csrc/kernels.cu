// kernels.cu -- the sm_100a stage kernels over the HBM photon SoA.
//
//   K0  source + arrival times + aspect dither      (k0_time_sums, k0_time_scan, k0_source)
//   K1  HRMA shell pair                             (k1_hrma)
//   K2  HETG facet diffraction                      (k2_grating)
//   K3  ACIS-S detection                            (k3_acis)
//
// K1..K3 are persistent kernels: a grid of (SM count x resident CTAs) blocks pulls 256-ray tiles from a
// ticket counter, stages the stage's small tables into shared memory with one TMA bulk copy
// (cp.async.bulk + mbarrier), traces one ray per thread, and compacts survivors into the output SoA in
// arrival order with warp ballots + a block prefix + a decoupled look-back over tile aggregates (the
// GPU form of marx_prune_photons, marx/libsrc/photon.c:40-63).  All loads/stores of the SoA columns are
// unit-stride across the warp.
//
// No tensor cores: nothing on this path is a dense contraction (BASELINE.json north_star).
// Compiled with -fmad=false: see mx_common.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mx_common.cuh"
#include "mx_tables.h"
#include "mx_source.cuh"
#include "mx_hrma.cuh"
#include "mx_grating.cuh"
#include "mx_acis.cuh"
#include "mx_kernels.cuh"
#include "../../include/marxb200.h"

namespace mx {

// ---------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D TMA bulk copy (global -> shared), relaxed gpu-scope ld/st
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (unsigned long long *bar, uint32_t count)
{
   asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
   asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx (unsigned long long *bar, uint32_t bytes)
{
   asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s (void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
   asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32 (dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait (unsigned long long *bar, uint32_t parity)
{
   asm volatile (
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed (const unsigned long long *p)
{
   unsigned long long v;
   asm volatile ("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_relaxed (unsigned long long *p, unsigned long long v)
{
   asm volatile ("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Stage a table blob into dynamic shared memory with ONE TMA bulk copy; every thread waits on the
// mbarrier.  blob_bytes is a multiple of 16 and both addresses are 16-byte aligned.
__device__ __forceinline__ void stage_blob (unsigned char *smem, const void *blob, uint32_t blob_bytes, unsigned long long *bar)
{
   if (threadIdx.x == 0) mbar_init (bar, 1);
   __syncthreads ();
   if (threadIdx.x == 0)
     {
        mbar_expect_tx (bar, blob_bytes);
        tma_bulk_g2s (smem, blob, blob_bytes, bar);
     }
   mbar_wait (bar, 0);
}

// ---------------------------------------------------------------------------------------------
// tile bookkeeping: ticket, block prefix, decoupled look-back
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ULL << 62, kFlagPrefix = 2ULL << 62, kValueMask = (1ULL << 62) - 1;

struct TileShared
{
   unsigned long long tile;
   unsigned long long excl;
   uint32_t warp_count[kTile / 32];
};

// Exclusive rank of this thread's survivor within the tile and the tile aggregate.
__device__ __forceinline__ uint32_t block_rank (bool alive, TileShared &ts, uint32_t &aggregate)
{
   const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   uint32_t ballot = __ballot_sync (0xffffffffu, alive);
   uint32_t rank = __popc (ballot & ((1u << lane) - 1u));
   if (lane == 0) ts.warp_count[warp] = __popc (ballot);
   __syncthreads ();
   uint32_t off = 0, total = 0;
#pragma unroll
   for (int w = 0; w < kTile / 32; w++)
     {
        uint32_t c = ts.warp_count[w];
        if (w < (int) warp) off += c;
        total += c;
     }
   aggregate = total;
   return off + rank;
}

// Publish this tile's aggregate and return the number of survivors in all earlier tiles
// (warp 0 walks back 32 tiles at a time; Merrill & Garland decoupled look-back).
__device__ __forceinline__ unsigned long long tile_exclusive_prefix (unsigned long long *status, unsigned long long tile,
                                                                     uint32_t aggregate, TileShared &ts)
{
   if (threadIdx.x < 32)
     {
        const uint32_t lane = threadIdx.x;
        unsigned long long excl = 0;
        if (tile == 0)
          {
             if (lane == 0) st_relaxed (status, kFlagPrefix | aggregate);
          }
        else
          {
             if (lane == 0) st_relaxed (status + tile, kFlagAgg | aggregate);
             long long j = (long long) tile - 1 - lane;          // this lane inspects tile j
             while (true)
               {
                  unsigned long long w = kFlagPrefix;             // tiles before 0 behave like an empty prefix
                  if (j >= 0)
                    {
                       do w = ld_relaxed (status + j); while ((w >> 62) == 0);
                    }
                  uint32_t is_prefix = __ballot_sync (0xffffffffu, (w >> 62) == 2);
                  unsigned long long v = w & kValueMask;
                  if (is_prefix)
                    {
                       // add aggregates of lanes closer than the first prefix, plus that prefix
                       uint32_t first = __ffs (is_prefix) - 1;
                       if (lane > first) v = 0;
#pragma unroll
                       for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
                       excl += v;
                       break;
                    }
#pragma unroll
                  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
                  excl += v;
                  j -= 32;
               }
             if (lane == 0) st_relaxed (status + tile, kFlagPrefix | (excl + aggregate));
          }
        if (lane == 0) ts.excl = excl;
     }
   __syncthreads ();
   return ts.excl;
}

// ---------------------------------------------------------------------------------------------
// K0: source, arrival times, dither
// ---------------------------------------------------------------------------------------------
// deterministic inclusive scan of one double per thread over the 256-thread tile; returns the tile total
__device__ __forceinline__ double tile_inclusive_scan (double v, double &total)
{
   __shared__ double warp_tot[kTile / 32];
   const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int o = 1; o < 32; o <<= 1)
     {
        double u = __shfl_up_sync (0xffffffffu, v, o);
        if (lane >= (uint32_t) o) v += u;
     }
   if (lane == 31) warp_tot[warp] = v;
   __syncthreads ();
   double off = 0.0, tot = 0.0;
#pragma unroll
   for (int w = 0; w < kTile / 32; w++)
     {
        double c = warp_tot[w];
        if (w < (int) warp) off += c;
        tot += c;
     }
   __syncthreads ();
   total = tot;
   return off + v;
}

// draws of one ray on the SOURCE sub-stream up to and including the arrival-time increment
__device__ __forceinline__ void k0_draw (const SourceArgs &a, uint64_t i, Rng &rng, double &energy, Vec3 &p, double &dt)
{
   rng.init (a.seed, a.first_ray + i, MARXB200_STAGE_SOURCE);
   source_draw (a.S, rng, energy, p);
   dt = source_time_increment (a.S, rng);
}

// pass 1: per-tile sums of the arrival-time increments
__global__ void __launch_bounds__ (kTile) k0_time_sums (const __grid_constant__ SourceArgs a)
{
   const uint64_t i = (uint64_t) blockIdx.x * kTile + threadIdx.x;
   double dt = 0.0;
   if (i < a.n)
     {
        Rng rng; double e; Vec3 p;
        k0_draw (a, i, rng, e, p, dt);
     }
   double total;
   (void) tile_inclusive_scan (dt, total);
   if (threadIdx.x == 0) a.tile_sums[blockIdx.x] = total;
}

// pass 2 (one CTA): canonical order = sequential over tiles inside a super-tile, sequential over
// super-tiles; results do not depend on how a run is split into batches or GPUs as long as the
// splits are super-tile aligned (DESIGN.md "arrival times").
__global__ void __launch_bounds__ (kTile) k0_time_scan (const __grid_constant__ SourceArgs a)
{
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const uint64_t n_super = (n_tiles + kSuperTile - 1) / kSuperTile;
   for (uint64_t s = threadIdx.x; s < n_super; s += blockDim.x)
     {
        double acc = 0.0;
        uint64_t t0 = s * kSuperTile, t1 = min (t0 + (uint64_t) kSuperTile, n_tiles);
        for (uint64_t t = t0; t < t1; t++) acc += a.tile_sums[t];
        a.supertile_sums[s] = acc;
     }
   __syncthreads ();
   if (threadIdx.x == 0)
     {
        double acc = a.use_dev_base ? a.dev_times[1] : a.time_base;
        a.dev_times[0] = acc;
        for (uint64_t s = 0; s < n_super; s++)
          {
             double v = a.supertile_sums[s];
             a.tile_base[s * kSuperTile] = acc;          // base of the first tile of the super-tile
             acc += v;
          }
        a.dev_times[1] = acc;
     }
   __syncthreads ();
   for (uint64_t s = threadIdx.x; s < n_super; s += blockDim.x)
     {
        uint64_t t0 = s * kSuperTile, t1 = min (t0 + (uint64_t) kSuperTile, n_tiles);
        double acc = a.tile_base[t0];
        for (uint64_t t = t0; t < t1; t++)
          {
             double v = a.tile_sums[t];
             a.tile_base[t] = acc;
             acc += v;
          }
     }
   if (threadIdx.x == 0) *a.n_out = a.n;
}

// pass 3: marx_create_photons for one ray per thread
__global__ void __launch_bounds__ (kTile) k0_source (const __grid_constant__ SourceArgs a)
{
   const uint64_t i = (uint64_t) blockIdx.x * kTile + threadIdx.x;
   const bool valid = i < a.n;
   Rng rng; double energy = 0.0, dt = 0.0; Vec3 p = v_make (0, 0, 0);
   if (valid) k0_draw (a, i, rng, energy, p, dt);
   double total;
   double t = tile_inclusive_scan (dt, total) + a.tile_base[blockIdx.x];
   if (!valid) return;
   float dra, ddec, droll;
   dither_ray (a.D, rng, t, p, dra, ddec, droll);
   const PhotonSoA &o = a.out;
   o.energy[i] = energy;
   o.p0[i] = p.x; o.p1[i] = p.y; o.p2[i] = p.z;
   o.time[i] = t;
   o.ray[i] = a.first_ray + i;
   o.flags[i] = 0;
   o.dra[i] = dra; o.ddec[i] = ddec; o.droll[i] = droll;
}

// ---------------------------------------------------------------------------------------------
// shared skeleton of the persistent stage kernels
// ---------------------------------------------------------------------------------------------
struct Carry { double time; float dra, ddec, droll; uint8_t shell; int8_t order; };

template <class Body>
__device__ __forceinline__ void stage_loop (const StageArgs &a, TileShared &ts, Body body)
{
   const unsigned long long n_in = *a.n_in;
   while (true)
     {
        __syncthreads ();
        if (threadIdx.x == 0) ts.tile = atomicAdd (a.ticket, 1ULL);
        __syncthreads ();
        const unsigned long long tile = ts.tile;
        const unsigned long long base = tile * kTile;
        if (base >= n_in)
          {
             if ((tile == 0) && (threadIdx.x == 0)) *a.n_out = 0;
             break;
          }
        const unsigned long long i = base + threadIdx.x;
        const bool valid = i < n_in;
        body (tile, i, valid, (base + kTile >= n_in));
     }
}

// K1 ------------------------------------------------------------------------------------------
// The mirror stage is three kernels (HRMA phases A, B, C of mx_hrma.cuh), each with its own fused
// compaction, so that every phase starts with full warps although 52 % / 43 % / 33 % of its rays die.
// State handed from phase to phase through otherwise unused SoA columns:
//   pha  (i16)  draws consumed so far on the MIRROR sub-stream | 0x4000 if a Box-Muller spare is cached
//   aux  (f64)  the cached spare
//   chipx, chipy, pi (f32)  beta, delta, effective-area correction (float-valued table lookups)
template <int PHASE>
__global__ void __launch_bounds__ (kTile) k1_hrma (const __grid_constant__ StageArgs a)
{
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   __shared__ TileShared ts;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   const K1Blob &B = *reinterpret_cast<const K1Blob *> (smem);
   const HrmaDev &H = B.H;

   stage_loop (a, ts, [&] (unsigned long long tile, unsigned long long i, bool valid, bool last_tile)
     {
        const PhotonSoA &in = a.in, &out = a.out;
        uint32_t flags = 0xFFu;
        double energy = 0.0; Vec3 x = v_make (0, 0, 0), p = v_make (0, 0, 0);
        uint32_t shell = 0; uint64_t ray = 0;
        float beta = 0.f, delta = 1.f, corr = 1.f;
        Rng rng;
        bool active = valid;
        if (valid && !a.compact) active = ((in.flags[i] & 0xFFu) == 0);
        if (active)
          {
             ray = in.ray[i];
             p = v_make (in.p0[i], in.p1[i], in.p2[i]);
             rng.init (a.seed, ray, MARXB200_STAGE_MIRROR);
             if (PHASE == 0)
               flags = hrma_phase_a (H, a.source_distance, x, p, shell, rng);
             else
               {
                  x = v_make (in.x0[i], in.x1[i], in.x2[i]);
                  energy = in.energy[i];
                  shell = in.shell[i];
                  const int st = in.pha[i];
                  rng.resume ((uint32_t) (st & 0x3FFF), (st & 0x4000) ? 1 : 0, (st & 0x4000) ? in.aux[i] : 0.0);
                  if (PHASE == 1)
                    {
                       hrma_optical_constants (H, H.shell[shell],
                                               reinterpret_cast<const float *> (smem + B.off_opt_e),
                                               reinterpret_cast<const float *> (smem + B.off_opt_b),
                                               reinterpret_cast<const float *> (smem + B.off_opt_d),
                                               reinterpret_cast<const float *> (smem + B.off_corr_e),
                                               reinterpret_cast<const float *> (smem + B.off_corr_f),
                                               energy, beta, delta, corr);
                       flags = hrma_phase_b (H, shell, energy, beta, delta, corr, x, p, rng);
                    }
                  else
                    {
                       beta = in.chipx[i]; delta = in.chipy[i]; corr = in.pi[i];
                       flags = hrma_phase_c (H, shell, energy, beta, delta, corr, x, p, rng);
                    }
               }
          }
        const int16_t rng_state = (int16_t) ((rng.draw & 0x3FFFu) | (rng.have_spare ? 0x4000u : 0u));
        unsigned long long j = i;
        bool write = active;
        if (a.compact)
          {
             const bool alive = active && (flags == 0);
             uint32_t aggregate;
             uint32_t rank = block_rank (alive, ts, aggregate);
             unsigned long long excl = tile_exclusive_prefix (a.tile_status, tile, aggregate, ts);
             j = excl + rank;
             write = alive;
             if (alive)
               {
                  // payload carried through the stage untouched
                  out.time[j] = in.time[i];
                  out.ray[j] = ray;
                  out.dra[j] = in.dra[i]; out.ddec[j] = in.ddec[i]; out.droll[j] = in.droll[i];
                  out.energy[j] = (PHASE == 0) ? in.energy[i] : energy;
                  out.shell[j] = (uint8_t) shell;
               }
             if (last_tile && (threadIdx.x == 0)) *a.n_out = excl + aggregate;
          }
        else if (last_tile && (threadIdx.x == 0)) *a.n_out = *a.n_in;
        if (write)
          {
             out.x0[j] = x.x; out.x1[j] = x.y; out.x2[j] = x.z;
             out.p0[j] = p.x; out.p1[j] = p.y; out.p2[j] = p.z;
             out.flags[j] = flags;
             if ((PHASE == 0) && !a.compact) out.shell[j] = (uint8_t) shell;
             if (PHASE < 2)
               {
                  out.pha[j] = rng_state;
                  if (rng.have_spare) out.aux[j] = rng.spare;
               }
             if (PHASE == 1) { out.chipx[j] = beta; out.chipy[j] = delta; out.pi[j] = corr; }
          }
     });
}

// K2 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (kTile) k2_grating (const __grid_constant__ StageArgs a)
{
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   __shared__ TileShared ts;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   // the shell descriptors hold global pointers for the big tables; sector tables are re-pointed to smem
   K2Blob &B = *reinterpret_cast<K2Blob *> (smem);
   if (threadIdx.x < kNumShells)
     B.G.shell[threadIdx.x].sectors = reinterpret_cast<const double *> (smem + B.off_sectors[threadIdx.x]);
   __syncthreads ();
   const GratingDev &G = B.G;

   stage_loop (a, ts, [&] (unsigned long long tile, unsigned long long i, bool valid, bool last_tile)
     {
        const PhotonSoA &in = a.in, &out = a.out;
        uint32_t flags = 0xFFu;
        double energy = 0.0; Vec3 x = v_make (0, 0, 0), p = v_make (0, 0, 0);
        uint32_t shell = 0; uint64_t ray = 0; int order = 0;
        bool active = valid;
        if (valid && !a.compact) active = ((in.flags[i] & 0xFFu) == 0);
        if (active)
          {
             energy = in.energy[i];
             x = v_make (in.x0[i], in.x1[i], in.x2[i]);
             p = v_make (in.p0[i], in.p1[i], in.p2[i]);
             shell = in.shell[i];
             ray = in.ray[i];
             Rng rng; rng.init (a.seed, ray, MARXB200_STAGE_GRATING);
             flags = grating_diffract (G, shell, energy, x, p, order, rng);
          }
        if (a.compact)
          {
             const bool alive = active && (flags == 0);
             uint32_t aggregate;
             uint32_t rank = block_rank (alive, ts, aggregate);
             unsigned long long excl = tile_exclusive_prefix (a.tile_status, tile, aggregate, ts);
             if (alive)
               {
                  const unsigned long long j = excl + rank;
                  out.energy[j] = energy;
                  out.x0[j] = x.x; out.x1[j] = x.y; out.x2[j] = x.z;
                  out.p0[j] = p.x; out.p1[j] = p.y; out.p2[j] = p.z;
                  out.time[j] = in.time[i];
                  out.ray[j] = ray;
                  out.flags[j] = 0;
                  out.dra[j] = in.dra[i]; out.ddec[j] = in.ddec[i]; out.droll[j] = in.droll[i];
                  out.shell[j] = (uint8_t) shell;
                  out.order[j] = (int8_t) order;
               }
             if (last_tile && (threadIdx.x == 0)) *a.n_out = excl + aggregate;
          }
        else if (active)
          {
             out.x0[i] = x.x; out.x1[i] = x.y; out.x2[i] = x.z;
             out.p0[i] = p.x; out.p1[i] = p.y; out.p2[i] = p.z;
             out.flags[i] = flags;
             out.order[i] = (int8_t) order;
          }
        if (!a.compact && last_tile && (threadIdx.x == 0)) *a.n_out = *a.n_in;
     });
}

// K3 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (kTile) k3_acis (const __grid_constant__ StageArgs a)
{
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   __shared__ TileShared ts;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   const AcisDev &A = reinterpret_cast<const K3Blob *> (smem)->A;

   stage_loop (a, ts, [&] (unsigned long long tile, unsigned long long i, bool valid, bool last_tile)
     {
        const PhotonSoA &in = a.in, &out = a.out;
        uint32_t flags = 0xFFu;
        double energy = 0.0, t_abs = 0.0; Vec3 x = v_make (0, 0, 0), p = v_make (0, 0, 0);
        uint64_t ray = 0; int ccd = -1; float chipx = 0, chipy = 0, pi = 0; int16_t pha = 0;
        bool active = valid;
        if (valid && !a.compact) active = ((in.flags[i] & 0xFFu) == 0);
        if (active)
          {
             energy = in.energy[i];
             x = v_make (in.x0[i], in.x1[i], in.x2[i]);
             p = v_make (in.p0[i], in.p1[i], in.p2[i]);
             t_abs = in.time[i];
             ray = in.ray[i];
             Rng rng; rng.init (a.seed, ray, MARXB200_STAGE_DETECTOR);
             flags = acis_detect (A, energy, t_abs, x, p, ccd, chipx, chipy, pha, pi, rng);
          }
        if (a.compact)
          {
             const bool alive = active && ((flags & 0xFFu) == 0);
             uint32_t aggregate;
             uint32_t rank = block_rank (alive, ts, aggregate);
             unsigned long long excl = tile_exclusive_prefix (a.tile_status, tile, aggregate, ts);
             if (alive)
               {
                  const unsigned long long j = excl + rank;
                  out.energy[j] = energy;
                  out.x0[j] = x.x; out.x1[j] = x.y; out.x2[j] = x.z;
                  out.p0[j] = p.x; out.p1[j] = p.y; out.p2[j] = p.z;
                  out.time[j] = t_abs;
                  out.ray[j] = ray;
                  out.flags[j] = flags;
                  out.dra[j] = in.dra[i]; out.ddec[j] = in.ddec[i]; out.droll[j] = in.droll[i];
                  out.shell[j] = in.shell[i];
                  out.order[j] = in.order[i];
                  out.ccd[j] = (int8_t) ccd;
                  out.chipx[j] = chipx; out.chipy[j] = chipy;
                  out.pha[j] = pha; out.pi[j] = pi;
               }
             if (last_tile && (threadIdx.x == 0)) *a.n_out = excl + aggregate;
          }
        else if (active)
          {
             out.x0[i] = x.x; out.x1[i] = x.y; out.x2[i] = x.z;
             out.p0[i] = p.x; out.p1[i] = p.y; out.p2[i] = p.z;
             out.flags[i] = flags;
             out.ccd[i] = (int8_t) ccd;
             out.chipx[i] = chipx; out.chipy[i] = chipy;
             out.pha[i] = pha; out.pi[i] = pi;
          }
        if (!a.compact && last_tile && (threadIdx.x == 0)) *a.n_out = *a.n_in;
     });
}

// ---------------------------------------------------------------------------------------------
// host boundary: SoA <-> 136-byte AoS records (marxb200_photon_attr == Marx_Photon_Attr_Type)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256) soa_to_aos (PhotonSoA in, const unsigned long long *n_ptr, uint64_t max_n,
                                                    marxb200_photon_attr *aos, const double *dev_start_time)
{
   const uint64_t n = min ((uint64_t) *n_ptr, max_n);
   const double start_time = *dev_start_time;
   for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
     {
        marxb200_photon_attr r;
        r.energy = in.energy[i];
        r.x[0] = in.x0[i]; r.x[1] = in.x1[i]; r.x[2] = in.x2[i];
        r.p[0] = in.p0[i]; r.p[1] = in.p1[i]; r.p[2] = in.p2[i];
        r.arrival_time = in.time[i] - start_time;
        r.flags = in.flags[i];
        r.y_pixel = in.chipx[i]; r.z_pixel = in.chipy[i]; r.u_pixel = 0.f; r.v_pixel = 0.f;
        r.dither_ra = in.dra[i]; r.dither_dec = in.ddec[i]; r.dither_roll = in.droll[i];
        r.dither_dy = 0.f; r.dither_dz = 0.f; r.dither_dtheta = 0.f;
        r.pi = in.pi[i];
        r.pulse_height = in.pha[i];
        r.mirror_shell = in.shell[i];
        r.ccd_num = in.ccd[i];
        r.detector_region = 0;
        r.order = in.order[i];
        r.support_orders[0] = r.support_orders[1] = r.support_orders[2] = r.support_orders[3] = 0;
        r.tag = (uint32_t) in.ray[i];
        // 17 aligned 8-byte stores per record
        const uint64_t *src = reinterpret_cast<const uint64_t *> (&r);
        uint64_t *dst = reinterpret_cast<uint64_t *> (aos + i);
#pragma unroll
        for (int k = 0; k < (int) (sizeof (marxb200_photon_attr) / 8); k++) dst[k] = src[k];
     }
}

__global__ void __launch_bounds__ (256) aos_to_soa (const marxb200_photon_attr *aos, const uint64_t *ray_ids, uint64_t n,
                                                    PhotonSoA out, double start_time)
{
   for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
     {
        marxb200_photon_attr r;
        const uint64_t *src = reinterpret_cast<const uint64_t *> (aos + i);
        uint64_t *dst = reinterpret_cast<uint64_t *> (&r);
#pragma unroll
        for (int k = 0; k < (int) (sizeof (marxb200_photon_attr) / 8); k++) dst[k] = src[k];
        out.energy[i] = r.energy;
        out.x0[i] = r.x[0]; out.x1[i] = r.x[1]; out.x2[i] = r.x[2];
        out.p0[i] = r.p[0]; out.p1[i] = r.p[1]; out.p2[i] = r.p[2];
        out.time[i] = r.arrival_time + start_time;
        out.ray[i] = ray_ids ? ray_ids[i] : (uint64_t) r.tag;
        out.flags[i] = r.flags;
        out.dra[i] = r.dither_ra; out.ddec[i] = r.dither_dec; out.droll[i] = r.dither_roll;
        out.chipx[i] = r.y_pixel; out.chipy[i] = r.z_pixel; out.pi[i] = r.pi;
        out.pha[i] = r.pulse_height;
        out.shell[i] = (uint8_t) r.mirror_shell;
        out.order[i] = r.order;
        out.ccd[i] = r.ccd_num;
     }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned int n_tiles_of (uint64_t n) { return (unsigned int) ((n + kTile - 1) / kTile); }

void launch_time_sums (const SourceArgs &a, cudaStream_t s)
{
   if (a.n == 0) return;
   k0_time_sums<<<n_tiles_of (a.n), kTile, 0, s>>> (a);
}
void launch_time_scan (const SourceArgs &a, cudaStream_t s)
{
   k0_time_scan<<<1, kTile, 0, s>>> (a);
}
void launch_source (const SourceArgs &a, cudaStream_t s)
{
   if (a.n == 0) return;
   k0_source<<<n_tiles_of (a.n), kTile, 0, s>>> (a);
}

template <class K>
static int occupancy_grid (K kernel, int num_sms, uint32_t smem_bytes)
{
   int per_sm = 1;
   cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
   cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, kTile, smem_bytes);
   if (per_sm < 1) per_sm = 1;
   return per_sm * num_sms;
}
int stage_grid_size (int stage, int num_sms, uint32_t blob_bytes)
{
   switch (stage)
     {
      case 10: return occupancy_grid (k1_hrma<0>, num_sms, blob_bytes);
      case 11: return occupancy_grid (k1_hrma<1>, num_sms, blob_bytes);
      case 12: return occupancy_grid (k1_hrma<2>, num_sms, blob_bytes);
      case 2: return occupancy_grid (k2_grating, num_sms, blob_bytes);
      case 3: return occupancy_grid (k3_acis, num_sms, blob_bytes);
     }
   return num_sms;
}
void launch_hrma (const StageArgs &a, int phase, int grid, cudaStream_t s)
{
   switch (phase)
     {
      case 0: k1_hrma<0><<<grid, kTile, a.blob_bytes, s>>> (a); break;
      case 1: k1_hrma<1><<<grid, kTile, a.blob_bytes, s>>> (a); break;
      default: k1_hrma<2><<<grid, kTile, a.blob_bytes, s>>> (a); break;
     }
}
void launch_grating (const StageArgs &a, int grid, cudaStream_t s) { k2_grating<<<grid, kTile, a.blob_bytes, s>>> (a); }
void launch_acis (const StageArgs &a, int grid, cudaStream_t s) { k3_acis<<<grid, kTile, a.blob_bytes, s>>> (a); }

void launch_soa_to_aos (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, void *aos,
                        const double *dev_start_time, cudaStream_t s)
{
   if (max_n == 0) return;
   unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (max_n + 255) / 256);
   soa_to_aos<<<grid, 256, 0, s>>> (in, n, max_n, (marxb200_photon_attr *) aos, dev_start_time);
}
void launch_aos_to_soa (const void *aos, const uint64_t *ray_ids, uint64_t n, const PhotonSoA &out, double start_time,
                        cudaStream_t s)
{
   if (n == 0) return;
   unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (n + 255) / 256);
   aos_to_soa<<<grid, 256, 0, s>>> ((const marxb200_photon_attr *) aos, ray_ids, n, out, start_time);
}

}  // namespace mx
