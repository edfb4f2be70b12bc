// kernels.cu -- the sm_100a stage kernels over the HBM photon SoA.
//
//   K0   source + arrival times + aspect dither      k0_time_sums, k0_time_super/_bases/_tiles, k0_source
//   K01  K0 fused with HRMA phase A                  k01_source_hrma      (marxb200_trace)
//   K1   HRMA shell pair, phases A / B / C           k1_hrma<0|1|2>
//   K2   HETG / LETG facet diffraction               k2_grating
//   K3   ACIS-S / ACIS-I, HRC-S / HRC-I detection     k3_acis, k3_hrc
//        arrival-order restoration                   order_mark, order_scan_words, order_scan_blocks, order_rank, order_gather
//        host boundary                               soa_to_aos, aos_to_soa, egress_pack, exposure_truncate
//
// K1..K3 are persistent kernels (grid = SM count x resident CTAs).  Every WARP pulls chunks of 32-ray tiles from a
// ticket counter, stages the stage's small tables into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier),
// traces one ray per lane, and re-packs survivors through a warp-private queue in shared memory into full, coalesced
// rows of the output list (ballot + popc ranks; the GPU form of marx_prune_photons, marx/libsrc/photon.c:40-63).
// Per-ray constants (energy, time, ray id, dither angles) are written once at the ray's batch slot (RayConst) and
// reached through the 4-byte slot key, so a compaction moves only what the stage produced.  The list is put back
// into arrival order once per batch (order_*).  All loads/stores of list columns are unit-stride across the warp.
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
#include "mx_hrc.cuh"
#include "mx_kernels.cuh"
#include "../../include/marxb200.h"

// resident CTAs per SM the register allocation is sized for (developer knobs: tools/build_variant.sh)
#ifndef MX_K1_MINBLOCKS
#define MX_K1_MINBLOCKS 3
#else
#define MX_K1_MINBLOCKS_SET 1
#endif
#ifndef MX_K2_MINBLOCKS
#define MX_K2_MINBLOCKS 3
#endif
// k01: 4 since its tiles come from a ticket (64 registers, 32 warps per SM: 1.224 -> 1.205 ms; with the static stride 3 was better)
#ifndef MX_K01_MINBLOCKS
#define MX_K01_MINBLOCKS 4
#endif
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
__device__ __forceinline__ void stage_blob (unsigned char *smem, const void *blob, uint32_t blob_bytes, unsigned long long *bar,
                                            uint32_t seg2_off = 0, uint32_t seg2_bytes = 0)
{
   if (threadIdx.x == 0) mbar_init (bar, 1);
   __syncthreads ();
   if (threadIdx.x == 0)
     {
        mbar_expect_tx (bar, blob_bytes + seg2_bytes);
        tma_bulk_g2s (smem, blob, blob_bytes, bar);
        // optional second byte range of the blob, placed behind the first (128-byte aligned)
        if (seg2_bytes) tma_bulk_g2s (smem + ((blob_bytes + 127u) & ~127u), (const unsigned char *) blob + seg2_off, seg2_bytes, bar);
     }
   mbar_wait (bar, 0);
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

// the same scan, carrying one 64-bit word from thread 0 to every thread across its first barrier (k01's next tile index)
__device__ __forceinline__ double tile_inclusive_scan_bcast (double v, double &total, unsigned long long &word)
{
   __shared__ double warp_tot[kTile / 32];
   __shared__ unsigned long long bcast;
   const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int o = 1; o < 32; o <<= 1)
     {
        double u = __shfl_up_sync (0xffffffffu, v, o);
        if (lane >= (uint32_t) o) v += u;
     }
   if (lane == 31) warp_tot[warp] = v;
   if (threadIdx.x == 0) bcast = word;
   __syncthreads ();
   double off = 0.0, tot = 0.0;
#pragma unroll
   for (int w = 0; w < kTile / 32; w++)
     {
        double c = warp_tot[w];
        if (w < (int) warp) off += c;
        tot += c;
     }
   word = bcast;
   __syncthreads ();
   total = tot;
   return off + v;
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

// pass 2: canonical order = sequential over tiles inside a super-tile, sequential over super-tiles; results do
// not depend on how a run is split into batches or GPUs as long as the splits are super-tile aligned (DESIGN.md
// "arrival times").  Three small kernels perform exactly the additions the original single-CTA kernel did (which spent
// 0.13 ms per 2^24-ray batch on uncoalesced, latency-bound loops): the sequential chains run out of shared memory, one
// chain per CTA, all super-tiles in parallel.
//   k0_time_super: super-tile s -> sum of its tile sums, left to right from 0
//   k0_time_bases: (one CTA) running base of every super-tile, left to right from the batch's time base
//   k0_time_tiles: super-tile s -> base of each of its tiles, left to right from the super-tile base
__global__ void __launch_bounds__ (kSuperTile) k0_time_super (const __grid_constant__ SourceArgs a)
{
   __shared__ double v[kSuperTile];
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const uint64_t t0 = (uint64_t) blockIdx.x * kSuperTile, t = t0 + threadIdx.x;
   v[threadIdx.x] = (t < n_tiles) ? a.tile_sums[t] : 0.0;
   __syncthreads ();
   if (threadIdx.x == 0)
     {
        const int cnt = (int) min ((uint64_t) kSuperTile, n_tiles - t0);
        double acc = 0.0;
        for (int k = 0; k < cnt; k++) acc += v[k];
        a.supertile_sums[blockIdx.x] = acc;
     }
}
__global__ void __launch_bounds__ (kSuperTile) k0_time_bases (const __grid_constant__ SourceArgs a)
{
   __shared__ double v[kSuperTile];
   __shared__ double carry;
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const uint64_t n_super = (n_tiles + kSuperTile - 1) / kSuperTile;
   if (threadIdx.x == 0)
     {
        carry = a.use_dev_base ? a.dev_times[1] : a.time_base;
        a.dev_times[0] = carry;
        *a.n_out = a.n;
     }
   for (uint64_t s0 = 0; s0 < n_super; s0 += kSuperTile)
     {
        const uint64_t sidx = s0 + threadIdx.x;
        __syncthreads ();
        v[threadIdx.x] = (sidx < n_super) ? a.supertile_sums[sidx] : 0.0;
        __syncthreads ();
        if (threadIdx.x == 0)
          {
             const int cnt = (int) min ((uint64_t) kSuperTile, n_super - s0);
             double acc = carry;
             for (int k = 0; k < cnt; k++) { const double x = v[k]; v[k] = acc; acc += x; }
             carry = acc;
          }
        __syncthreads ();
        if (sidx < n_super) a.tile_base[sidx * kSuperTile] = v[threadIdx.x];      // base of the first tile of the super-tile
     }
   __syncthreads ();
   if (threadIdx.x == 0) a.dev_times[1] = carry;
}
// Multi-GPU form of k0_time_bases (marxb200_trace_sharded, comm.cu): the rays of one collective step are split into
// `world` contiguous blocks of ns_blk super-tiles, rank r traces block r.  all_sums = [world][ns_blk] holds the super-tile
// sums of EVERY rank's block (ncclAllGather of the vectors k0_time_super produced; short or empty blocks are zero padded,
// and acc + 0.0 == acc).  One thread adds them in global ray order from the running end time of the previous step --
// exactly the additions k0_time_bases performs when one GPU traces the same blocks one after the other -- and notes the
// bases of this rank's super-tiles on the way.  Every rank ends with the same running end time: no host round trip, no
// second pass over the increments.
__global__ void __launch_bounds__ (kSuperTile) k0_time_bases_sharded (const __grid_constant__ SourceArgs a, const double *all_sums,
                                                                      int rank, int world, uint32_t ns_blk)
{
   __shared__ double v[kSuperTile];
   __shared__ double carry;
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const uint64_t n_super = (n_tiles + kSuperTile - 1) / kSuperTile;      // super-tiles of this rank's block (<= ns_blk)
   if (threadIdx.x == 0) carry = a.use_dev_base ? a.dev_times[1] : a.time_base;
   for (int r = 0; r < world; r++)
     {
        if ((r == rank) && (threadIdx.x == 0))
          {
             a.dev_times[0] = carry;               // `carry` is only ever written by this thread
             *a.n_out = a.n;
          }
        for (uint32_t s0 = 0; s0 < ns_blk; s0 += kSuperTile)
          {
             const uint32_t sidx = s0 + threadIdx.x;
             __syncthreads ();
             v[threadIdx.x] = (sidx < ns_blk) ? all_sums[(size_t) r * ns_blk + sidx] : 0.0;
             __syncthreads ();
             if (threadIdx.x == 0)
               {
                  const int cnt = (int) min ((uint32_t) kSuperTile, ns_blk - s0);
                  double acc = carry;
                  for (int k = 0; k < cnt; k++) { const double x = v[k]; v[k] = acc; acc += x; }
                  carry = acc;
               }
             __syncthreads ();
             if ((r == rank) && (sidx < n_super)) a.tile_base[(uint64_t) sidx * kSuperTile] = v[threadIdx.x];
          }
     }
   __syncthreads ();
   if (threadIdx.x == 0) a.dev_times[1] = carry;
}
__global__ void __launch_bounds__ (kSuperTile) k0_time_tiles (const __grid_constant__ SourceArgs a)
{
   __shared__ double v[kSuperTile];
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const uint64_t t0 = (uint64_t) blockIdx.x * kSuperTile, t = t0 + threadIdx.x;
   v[threadIdx.x] = (t < n_tiles) ? a.tile_sums[t] : 0.0;
   __syncthreads ();
   if (threadIdx.x == 0)
     {
        const int cnt = (int) min ((uint64_t) kSuperTile, n_tiles - t0);
        double acc = a.tile_base[t0];
        for (int k = 0; k < cnt; k++) { const double x = v[k]; v[k] = acc; acc += x; }
     }
   __syncthreads ();
   if (t < n_tiles) a.tile_base[t] = v[threadIdx.x];
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
   float dra, ddec, droll, det[3];
   dither_ray (a.D, rng, t, p, dra, ddec, droll, nullptr, det);
   const PhotonSoA &o = a.out;
   o.energy[i] = energy;
   o.p0[i] = p.x; o.p1[i] = p.y; o.p2[i] = p.z;
   o.time[i] = t;
   o.ray[i] = a.first_ray + i;
   o.slot[i] = (uint32_t) i;
   o.flags[i] = 0;
   o.order[i] = 0; o.sorders[i] = 0;      // stay 0 when GratingType=NONE (memset of source.c:287)
   o.dra[i] = dra; o.ddec[i] = ddec; o.droll[i] = droll;
   o.ddy[i] = det[0]; o.ddz[i] = det[1]; o.ddth[i] = det[2];
   const RayConst &rc = a.rc;           // slot == i
   rc_store (rc, i, energy, a.first_ray + i); rc.time[i] = t;
   rc.dra[i] = dra; rc.ddec[i] = ddec; rc.droll[i] = droll;
   rc.ddy[i] = det[0]; rc.ddz[i] = det[1]; rc.ddth[i] = det[2];
}

// ---------------------------------------------------------------------------------------------
// skeleton of the persistent stage kernels: autonomous warps + warp-private re-packing queues
// ---------------------------------------------------------------------------------------------
// Every warp pulls chunks of a.chunk_tiles 32-ray tiles from a ticket counter and traces one ray per
// lane.  Survivors are ranked with a ballot and appended to the warp's PRIVATE ring buffer in shared
// memory; whenever 32 or more are queued the warp reserves 32 output slots with one atomicAdd and writes a
// FULL, coalesced row of the output SoA (the carried columns are gathered from the input SoA through
// the staged source index).  There is no block barrier and no waiting on other tiles anywhere in the
// loop: rays die at very different depths, and in the first version (ordered decoupled look-back + block
// barrier per tile) 20-35 % of all stall samples sat on those two waits (profiles/README.md).
// The price is that a stage's output is no longer in arrival order; the order is restored once, by
// restore_order below, when the list is observed (download) or at the end of marxb200_trace.
template <int ND, int NU> struct WarpQueue { double d[ND][kQueueCap]; uint32_t u[NU][kQueueCap]; };

template <int ND, int NU, class Trace, class FlushEntry, class InPlace>
__device__ __forceinline__ void run_stage (const StageArgs &a, WarpQueue<ND, NU> &q, Trace trace, FlushEntry flush_entry,
                                           InPlace store_in_place)
{
   const uint32_t lane = threadIdx.x & 31;
   const unsigned long long n_in = *a.n_in;
   uint32_t head = 0, count = 0;

   auto flush = [&] (uint32_t n_flush)
     {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd (a.n_out, (unsigned long long) n_flush);
        base = __shfl_sync (0xffffffffu, base, 0);
        if (lane < n_flush) flush_entry ((head + lane) & (kQueueCap - 1), base + lane);
        head = (head + n_flush) & (kQueueCap - 1);
        count -= n_flush;
        __syncwarp ();
     };

   // (measured without effect or slower: drawing the NEXT chunk's ticket before tracing the current one -- the extra
   // live registers spill, -2 %; prefetch.global.L2 of the next tile's columns -- 0 %; 4-tile chunks for every stage --
   // load imbalance in the detector kernel, -2 %)
   unsigned long long chunk = 0;
   if (lane == 0) chunk = atomicAdd (a.ticket, 1ULL);
   chunk = __shfl_sync (0xffffffffu, chunk, 0);
   while (true)
     {
        const unsigned long long base0 = chunk * (unsigned long long) (a.chunk_tiles * kWarpTile);
        if (base0 >= n_in)
          {
             if (!a.compact && (chunk == 0) && (lane == 0)) *a.n_out = n_in;     // n_in == 0
             break;
          }
        if (!a.compact && (chunk == 0) && (lane == 0)) *a.n_out = n_in;
#pragma unroll 1
        for (int t = 0; t < a.chunk_tiles; t++)
          {
             const unsigned long long base = base0 + (unsigned long long) t * kWarpTile;
             if (base >= n_in) break;
             const unsigned long long i = base + lane;
             bool active = i < n_in;
             if (active && !a.compact) active = ((a.in.flags[i] & 0xFFu) == 0);
             double d[ND]; uint32_t u[NU];
             uint32_t flags = 0xFFu;
             if (active) flags = trace (i, d, u);
             if (a.compact)
               {
                  const bool alive = active && ((flags & 0xFFu) == 0);
                  const uint32_t ballot = __ballot_sync (0xffffffffu, alive);
                  if (alive)
                    {
                       const uint32_t pos = (head + count + __popc (ballot & ((1u << lane) - 1u))) & (kQueueCap - 1);
#pragma unroll
                       for (int k = 0; k < ND; k++) q.d[k][pos] = d[k];
#pragma unroll
                       for (int k = 0; k < NU; k++) q.u[k][pos] = u[k];
                    }
                  count += __popc (ballot);
                  __syncwarp ();
                  if (count >= 32) flush (32);
               }
             else if (active) store_in_place (i, d, u, flags);
          }
        unsigned long long next_chunk = 0;
        if (lane == 0) next_chunk = atomicAdd (a.ticket, 1ULL);
        chunk = __shfl_sync (0xffffffffu, next_chunk, 0);
     }
   if (a.compact && (count > 0)) flush (count);
}

template <int ND, int NU>
__device__ __forceinline__ WarpQueue<ND, NU> &my_queue (unsigned char *smem, uint32_t blob_bytes)
{
   return reinterpret_cast<WarpQueue<ND, NU> *> (smem + ((blob_bytes + 127u) & ~127u))[threadIdx.x >> 5];
}

// K1 ------------------------------------------------------------------------------------------
// The mirror stage is a chain of kernels (HRMA phases of mx_hrma.cuh) so that every phase starts with
// full warps although most of its rays die.  Two cuts of the same per-ray sequence exist:
//   PHASE 0 | 1 | 2   = A | B | C: cut at the two conic intersections (in-place parity mode, MARXB200_K1_SPLIT=0)
//   PHASE 0 | 3 | 4 | 5 = A | B1 | B2+C1 | C2: additionally cut behind the two reflectivity tests, where 35-40 % of the rays
//                       of B and C are absorbed (the scatter, the frame transforms and the next intersection then run on
//                       re-packed warps); the compacting path
// State handed from phase to phase through otherwise unused SoA columns:
//   pha  (i16)  draws consumed so far on the MIRROR sub-stream | 0x4000 if a Box-Muller spare is cached
//   aux  (f64)  the cached spare
//   chipx, chipy, pi (f32)  beta, delta, effective-area correction (float-valued table lookups)
//   energy, time, ray (f64 bit patterns; PHASE 3 -> 4 -> 5 only)  the blurred surface normal that passed the reflectivity test.
//               These list columns are dead between the source and the order restoration: the per-ray constants live in RayConst.
template <int PHASE> struct K1Shape
{
   static constexpr bool kNormalOut = (PHASE == 3) || (PHASE == 4), kNormalIn = (PHASE == 4) || (PHASE == 5);
   static constexpr bool kSpareOut = (PHASE == 1) || kNormalOut, kOptOut = (PHASE == 1) || (PHASE == 3), kLast = (PHASE == 2) || (PHASE == 5);
   // B1 leaves x and p as they are: its queue carries the index of the input row instead (x, p and the slot key are copied from
   // there at flush; 52 instead of 100 bytes per entry keeps three CTAs per SM next to the optical-constant tables)
   static constexpr bool kCarryXP = (PHASE == 3);
   static constexpr int XO = kCarryXP ? 0 : 6;     // queue columns [0, XO) = x, p
   static constexpr int ND = XO + (kNormalOut ? 3 : 0) + (kSpareOut ? 1 : 0), NU = 2 + (kOptOut ? 3 : 0);
};

// Resident CTAs per SM by phase: B1 and C2 (phases 3, 5) run faster with four (64 registers, 32 warps: 0.705 -> 0.652 ms and 0.217 ->
// 0.199 ms per C2 batch), B2+C1 (phase 4: reflection, scatter, two transforms, the H intersection and the second reflectivity test in
// one kernel) loses 4 % to the spills and keeps three.  -DMX_K1_MINBLOCKS=k forces one value for all phases (A/B builds).
#ifdef MX_K1_MINBLOCKS_SET
constexpr int k1_min_blocks (int) { return MX_K1_MINBLOCKS; }
#else
constexpr int k1_min_blocks (int phase) { return ((phase == 3) || (phase == 5)) ? 4 : MX_K1_MINBLOCKS; }
#endif
template <int PHASE>
__global__ void __launch_bounds__ (kStageThreads, k1_min_blocks (PHASE)) k1_hrma (const __grid_constant__ StageArgs a)
{
   using Shape = K1Shape<PHASE>;
   constexpr int ND = Shape::ND, NU = Shape::NU;
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   // only the phases that look up the optical constants (B, B1) need the tables behind the header; phases that scatter stage
   // the WFOLD search keys of their conic (B: contiguous with the tables; C, B2+C1, C2: as a second segment behind the header)
   const uint32_t seg1 = (PHASE == 1 || PHASE == 3) ? a.blob_bytes : (uint32_t) sizeof (K1Blob);
   const uint32_t seg2 = (PHASE == 2 || PHASE == 4 || PHASE == 5) ? a.seg2_bytes : 0u;
   stage_blob (smem, a.blob, seg1, &bar, a.seg2_off, seg2);
   const K1Blob &B = *reinterpret_cast<const K1Blob *> (smem);
   const HrmaDev &H = B.H;
   const uint32_t staged = seg2 ? (((seg1 + 127u) & ~127u) + seg2) : seg1;
#ifdef MX_NO_WFOLD_KEYS_SMEM
   const unsigned char *wkeys = nullptr;
#else
   const unsigned char *wkeys = (B.wkeys_bytes == 0) ? nullptr
                                : ((PHASE == 1) ? smem + B.off_wkeys_p : ((PHASE != 3 && seg2) ? smem + ((seg1 + 127u) & ~127u) : nullptr));
#endif
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, staged);
   const PhotonSoA &in = a.in, &out = a.out;

   auto trace = [&] (unsigned long long i, double *d, uint32_t *u) -> uint32_t
     {
        Vec3 x = v_make (0, 0, 0), p = v_make (in.p0[i], in.p1[i], in.p2[i]), normal = v_make (0, 0, 0);
        uint32_t shell = 0, flags;
        float beta = 0.f, delta = 1.f, corr = 1.f;
        Rng rng;
        const uint32_t slot = in.slot[i];
        double energy; uint64_t ray;
        rc_load (a.rc, slot, energy, ray);
        rng.init (a.seed, ray, MARXB200_STAGE_MIRROR);
        if (PHASE == 0)
          flags = hrma_phase_a (H, a.source_distance, x, p, shell, rng);
        else
          {
             x = v_make (in.x0[i], in.x1[i], in.x2[i]);
             shell = in.shell[i];
             const int st = in.pha[i];
             rng.resume ((uint32_t) (st & 0x3FFF), (st & 0x4000) ? 1 : 0, (st & 0x4000) ? in.aux[i] : 0.0);
             const double *wk = wkeys ? reinterpret_cast<const double *> (wkeys + shell * B.wkeys_stride) : nullptr;
             if (Shape::kNormalIn)
               normal = v_make (in.energy[i], in.time[i], __longlong_as_double ((long long) in.ray[i]));
             if (PHASE == 1 || PHASE == 3)
               hrma_optical_constants (H, H.shell[shell],
                                       reinterpret_cast<const float *> (smem + B.off_opt_e),
                                       reinterpret_cast<const float *> (smem + B.off_opt_b),
                                       reinterpret_cast<const float *> (smem + B.off_opt_d),
                                       reinterpret_cast<const float *> (smem + B.off_corr_e),
                                       reinterpret_cast<const float *> (smem + B.off_corr_f),
                                       energy, beta, delta, corr);
             else if (PHASE == 2 || PHASE == 4)
               { beta = in.chipx[i]; delta = in.chipy[i]; corr = in.pi[i]; }
             if (PHASE == 1) flags = hrma_phase_b (H, shell, energy, beta, delta, corr, x, p, rng, wk);
             else if (PHASE == 2) flags = hrma_phase_c (H, shell, energy, beta, delta, corr, x, p, rng, wk);
             else if (PHASE == 3) flags = hrma_phase_b1 (H, shell, beta, delta, corr, x, p, normal, rng);
             else if (PHASE == 4)
               {
                  flags = hrma_phase_b2 (H, shell, energy, normal, x, p, rng, wk);
                  if (flags == 0) flags = hrma_phase_c1 (H, shell, beta, delta, corr, x, p, normal, rng);
               }
             else flags = hrma_phase_c2 (H, shell, energy, normal, x, p, rng, wk);
          }
        if (!Shape::kCarryXP) { d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = p.x; d[4] = p.y; d[5] = p.z; }
        u[0] = Shape::kCarryXP ? (uint32_t) i : slot;
        u[1] = shell | (((rng.draw & 0x3FFFu) | (rng.have_spare ? 0x4000u : 0u)) << 8);
        if (Shape::kNormalOut) { d[Shape::XO] = normal.x; d[Shape::XO + 1] = normal.y; d[Shape::XO + 2] = normal.z; }
        if (Shape::kSpareOut) d[ND - 1] = rng.spare;
        if (Shape::kOptOut)
          { u[NU - 3] = __float_as_uint (beta); u[NU - 2] = __float_as_uint (delta); u[NU - 1] = __float_as_uint (corr); }
        return flags;
     };
   auto write_row = [&] (unsigned long long j, const double *d, const uint32_t *u, uint32_t flags)
     {
        if (!Shape::kCarryXP)
          {
             out.x0[j] = d[0]; out.x1[j] = d[1]; out.x2[j] = d[2];
             out.p0[j] = d[3]; out.p1[j] = d[4]; out.p2[j] = d[5];
          }
        out.flags[j] = flags;
        out.shell[j] = (uint8_t) (u[1] & 0xFFu);
        if (!Shape::kLast) out.pha[j] = (int16_t) (u[1] >> 8);
        else
          {
             // the last mirror kernel hands the scratch columns back zeroed, as the reference's memset of the batch
             // (source.c:287) leaves them when no detector follows (DetectorType=NONE)
             out.pha[j] = 0; out.chipx[j] = 0.f; out.chipy[j] = 0.f; out.pi[j] = 0.f;
          }
        if (Shape::kNormalOut)
          { out.energy[j] = d[Shape::XO]; out.time[j] = d[Shape::XO + 1]; out.ray[j] = (uint64_t) __double_as_longlong (d[Shape::XO + 2]); }
        if (Shape::kSpareOut) out.aux[j] = d[ND - 1];
        if (Shape::kOptOut)
          { out.chipx[j] = __uint_as_float (u[NU - 3]); out.chipy[j] = __uint_as_float (u[NU - 2]); out.pi[j] = __uint_as_float (u[NU - 1]); }
     };
   auto flush_entry = [&] (uint32_t pos, unsigned long long j)
     {
        double d[ND]; uint32_t u[NU];
#pragma unroll
        for (int k = 0; k < ND; k++) d[k] = q.d[k][pos];
#pragma unroll
        for (int k = 0; k < NU; k++) u[k] = q.u[k][pos];
        write_row (j, d, u, 0);
        if (Shape::kCarryXP)
          {
             const uint32_t src = u[0];      // a row this warp read a moment ago: L1/L2 hits
             out.x0[j] = in.x0[src]; out.x1[j] = in.x1[src]; out.x2[j] = in.x2[src];
             out.p0[j] = in.p0[src]; out.p1[j] = in.p1[src]; out.p2[j] = in.p2[src];
             out.slot[j] = in.slot[src];
          }
        else out.slot[j] = u[0];
     };
   auto in_place = [&] (unsigned long long i, const double *d, const uint32_t *u, uint32_t flags) { write_row (i, d, u, flags); };
   run_stage<ND, NU> (a, q, trace, flush_entry, in_place);
}

// K1 (MirrorType=FLATFIELD) ---------------------------------------------------------------------
// _marx_ff_mirror_reflect, marx/libsrc/ffield.c:61-108: no optics.  Draw order on the MIRROR sub-stream: z, then y.
__global__ void __launch_bounds__ (kStageThreads) k1_flatfield (const __grid_constant__ StageArgs a)
{
   constexpr int ND = 6, NU = 1;
   extern __shared__ __align__ (128) unsigned char smem[];
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, 0);
   const PhotonSoA &in = a.in, &out = a.out;
   auto trace = [&] (unsigned long long i, double *d, uint32_t *u) -> uint32_t
     {
        Vec3 p = v_make (in.p0[i], in.p1[i], in.p2[i]), x;
        const uint32_t slot = in.slot[i];
        Rng rng;
        double energy; uint64_t ray;
        rc_load (a.rc, slot, energy, ray);
        rng.init (a.seed, ray, MARXB200_STAGE_MIRROR);
        x.z = a.ff[1] + rng.uniform () * (a.ff[3] - a.ff[1]);
        x.y = a.ff[0] + rng.uniform () * (a.ff[2] - a.ff[0]);
        x.x = a.ff[4];
        if (a.source_distance > 0.0)
          {
             p = v_ax1_bx2 (1.0, x, a.source_distance, p);
             v_normalize (p);
          }
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = p.x; d[4] = p.y; d[5] = p.z;
        u[0] = slot;
        return 0u;
     };
   auto write_row = [&] (unsigned long long j, const double *d)
     {
        out.x0[j] = d[0]; out.x1[j] = d[1]; out.x2[j] = d[2];
        out.p0[j] = d[3]; out.p1[j] = d[4]; out.p2[j] = d[5];
        out.flags[j] = 0; out.shell[j] = 0;
        out.pha[j] = 0; out.chipx[j] = 0.f; out.chipy[j] = 0.f; out.pi[j] = 0.f;
     };
   auto flush_entry = [&] (uint32_t pos, unsigned long long j)
     {
        double d[ND];
#pragma unroll
        for (int k = 0; k < ND; k++) d[k] = q.d[k][pos];
        write_row (j, d);
        out.slot[j] = q.u[0][pos];
        out.order[j] = 0; out.sorders[j] = 0;
     };
   auto in_place = [&] (unsigned long long i, const double *d, const uint32_t *, uint32_t) { write_row (i, d); };
   run_stage<ND, NU> (a, q, trace, flush_entry, in_place);
}

// K2 ------------------------------------------------------------------------------------------
// The compacting path runs the grating stage as two kernels (mx_grating.cuh: grating_select, grating_diffract_selected).
// k2_select needs neither x nor p: it reads a row's slot key and shell, draws the vignetting and order-selection deviates and
// appends (row index | order index << 32) of the ~half of the rays that leave the primary grating in an order to the `ray` column
// of the OUTPUT buffer (dead until the order restoration, and not written by k2_grating).  k2_grating<1> then walks that list:
// the facet rotation, the Rowland-torus intersection, the diffraction and the support gratings run on full warps.  Warps are
// uniform here (one table search + five interpolations per ray), so a grid-stride loop with warp-aggregated appends suffices.
// (measured without effect on B200: 32 registers per thread for 64 resident warps per SM, and two tiles per warp with the two table
// searches advanced in lockstep -- 0.279 / 0.280 / 0.279 ms for the stage.  The kernel is not short of warps or of independent
// loads: its 97 B of DRAM traffic per ray are the 32-byte sectors of the two per-ray constants gathered through the slot key.)
// (Round 2, measured without gain: the four shells' energy grids -- 1408 floats each for the HETG -- copied into shared memory once per
// CTA, so that the ten probes of the energy bracket stay on the SM: 0.273 against 0.266 ms for the stage.  The bracket's probes hit
// L1 / L2 lines shared by the whole warp; the kernel waits for the sector-granular gathers through the slot key, see above.)
__global__ void __launch_bounds__ (256) k2_select (const __grid_constant__ StageArgs a)
{
   const K2Blob *B = reinterpret_cast<const K2Blob *> (a.blob);          // header fields through L1; the tables live in L2
   const unsigned long long n_in = *a.n_in;
   const uint32_t lane = threadIdx.x & 31;
   const unsigned long long n_warps = ((unsigned long long) gridDim.x * blockDim.x) >> 5;
   for (unsigned long long tile = ((unsigned long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile * 32 < n_in; tile += n_warps)
     {
        const unsigned long long i = tile * 32 + lane;
        int lo = -1;
        if (i < n_in)
          {
             const uint32_t slot = a.in.slot[i], shell = a.in.shell[i];
             Rng rng;
             double energy; uint64_t ray;
             rc_load (a.rc, slot, energy, ray);
             rng.init (a.seed, ray, MARXB200_STAGE_GRATING);
             lo = grating_select (B->G.shell[shell], energy, rng);
          }
        const uint32_t ballot = __ballot_sync (0xffffffffu, lo >= 0);
        unsigned long long base = 0;
        if ((lane == 0) && ballot) base = atomicAdd (a.n_out, (unsigned long long) __popc (ballot));
        base = __shfl_sync (0xffffffffu, base, 0);
        if (lo >= 0)
          a.out.ray[base + __popc (ballot & ((1u << lane) - 1u))] = (unsigned long long) i | ((unsigned long long) (uint32_t) lo << 32);
     }
}

// PHASE 0: the whole stage for the rows of `in`; PHASE 1: the second half for the rows k2_select listed in out.ray
template <int PHASE>
__global__ void __launch_bounds__ (kStageThreads, MX_K2_MINBLOCKS) k2_grating (const __grid_constant__ StageArgs a)
{
   constexpr int ND = 6, NU = 3;
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   // the shell descriptors hold global pointers for the big tables; sector tables are re-pointed to smem
   K2Blob &B = *reinterpret_cast<K2Blob *> (smem);
   if (threadIdx.x < kNumShells)
     B.G.shell[threadIdx.x].sectors = reinterpret_cast<const double *> (smem + B.off_sectors[threadIdx.x]);
   __syncthreads ();
   const GratingDev &G = B.G;
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, a.blob_bytes);
   const PhotonSoA &in = a.in, &out = a.out;

   auto trace = [&] (unsigned long long i, double *d, uint32_t *u) -> uint32_t
     {
        unsigned long long src = i;
        int lo = 0;
        if (PHASE == 1)
          {
             const unsigned long long packed = out.ray[i];
             src = packed & 0xFFFFFFFFull; lo = (int) (packed >> 32);
          }
        Vec3 x = v_make (in.x0[src], in.x1[src], in.x2[src]), p = v_make (in.p0[src], in.p1[src], in.p2[src]);
        int order = 0;
        uint32_t sorders = 0;
        Rng rng;
        const uint32_t slot = in.slot[src], shell = in.shell[src];
        double energy; uint64_t ray;
        rc_load (a.rc, slot, energy, ray);
        rng.init (a.seed, ray, MARXB200_STAGE_GRATING);
        uint32_t flags;
        if (PHASE == 1)
          {
             rng.resume (2, 0, 0.0);            // behind the vignetting and order-selection draws
             flags = grating_diffract_selected (G, shell, energy, x, p, lo, order, sorders, rng);
          }
        else flags = grating_diffract (G, shell, energy, x, p, order, sorders, rng);
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = p.x; d[4] = p.y; d[5] = p.z;
        u[0] = slot;
        u[1] = (uint32_t) (order & 0xFF) | (shell << 8);
        u[2] = sorders;
        return flags;
     };
   auto write_row = [&] (unsigned long long j, const double *d, const uint32_t *u, uint32_t flags)
     {
        out.x0[j] = d[0]; out.x1[j] = d[1]; out.x2[j] = d[2];
        out.p0[j] = d[3]; out.p1[j] = d[4]; out.p2[j] = d[5];
        out.flags[j] = flags;
        out.order[j] = (int8_t) (u[1] & 0xFFu);
        out.sorders[j] = u[2];
     };
   auto flush_entry = [&] (uint32_t pos, unsigned long long j)
     {
        double d[ND]; uint32_t u[NU];
#pragma unroll
        for (int k = 0; k < ND; k++) d[k] = q.d[k][pos];
#pragma unroll
        for (int k = 0; k < NU; k++) u[k] = q.u[k][pos];
        write_row (j, d, u, 0);
        out.slot[j] = u[0];
        out.shell[j] = (uint8_t) (u[1] >> 8);
     };
   auto in_place = [&] (unsigned long long i, const double *d, const uint32_t *u, uint32_t flags) { write_row (i, d, u, flags); };
   run_stage<ND, NU> (a, q, trace, flush_entry, in_place);
}

// K3 ------------------------------------------------------------------------------------------
#ifndef MX_K3_MINBLOCKS
#define MX_K3_MINBLOCKS 3
#endif
// PHASE 0: the whole detector stage in one kernel; 1, 2: its two halves (acis_detect_a / _b, mx_acis.cuh) as two kernels with a
// re-packed list in between.  The chip index found by the first half travels in the pha column (not yet used), the chip
// pixels in theirs; the QE test is the only draw of the first half, so the second resumes the DETECTOR sub-stream at draw 0 or 1.
#ifndef MX_K3A_MINBLOCKS
#define MX_K3A_MINBLOCKS MX_K3_MINBLOCKS
#endif
#ifndef MX_K3B_MINBLOCKS
#define MX_K3B_MINBLOCKS MX_K3_MINBLOCKS
#endif
constexpr int k3_min_blocks (int phase) { return (phase == 1) ? MX_K3A_MINBLOCKS : ((phase == 2) ? MX_K3B_MINBLOCKS : MX_K3_MINBLOCKS); }
template <bool DET, int PHASE>
__global__ void __launch_bounds__ (kStageThreads, k3_min_blocks (PHASE)) k3_acis (const __grid_constant__ StageArgs a)
{
   constexpr int ND = 6, NU = 7;
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   const AcisDev &A = reinterpret_cast<const K3Blob *> (smem)->A;
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, a.blob_bytes);
   // FEF scratch: the cumulative gaussian areas of the ray a thread is tracing, [kMaxGauss][kStageThreads] floats
   // behind the warp queues (element k of thread t at [k * kStageThreads + t]: conflict-free)
   float *fef_cum = reinterpret_cast<float *> (smem + ((a.blob_bytes + 127u) & ~127u) + (kStageThreads / 32) * sizeof (WarpQueue<ND, NU>)) + threadIdx.x;
   const PhotonSoA &in = a.in, &out = a.out;

   auto trace = [&] (unsigned long long i, double *d, uint32_t *u) -> uint32_t
     {
        Vec3 x = v_make (in.x0[i], in.x1[i], in.x2[i]), p = v_make (in.p0[i], in.p1[i], in.p2[i]);
        int ccd = -1; float chipx = 0, chipy = 0, pi = 0; int16_t pha = 0;
        Rng rng;
        const uint32_t slot = in.slot[i];
        double energy; uint64_t ray;
        rc_load (a.rc, slot, energy, ray);
        rng.init (a.seed, ray, MARXB200_STAGE_DETECTOR);
        DetDither dd = {0.0, 0.0, 0.0};
        if (DET) { dd.dy = a.rc.ddy[slot]; dd.dz = a.rc.ddz[slot]; dd.dtheta = a.rc.ddth[slot]; }
        uint32_t flags;
        if (PHASE == 0)
          flags = acis_detect<DET> (A, energy, a.rc.time[slot], x, p, ccd, chipx, chipy, pha, pi, rng, fef_cum, kStageThreads, dd);
        else if (PHASE == 1)
          {
             int hit = -1;
             flags = acis_detect_a<DET> (A, energy, x, p, ccd, hit, chipx, chipy, rng, dd);
             pha = ((flags & 0xFFu) == 0) ? (int16_t) hit : (int16_t) 0;
          }
        else
          {
             const int hit = (int) in.pha[i];
             ccd = (int) in.ccd[i]; chipx = in.chipx[i]; chipy = in.chipy[i];
             rng.resume (A.det_ideal ? 0u : 1u, 0, 0.0);
             flags = acis_detect_b<DET> (A, energy, a.rc.time[slot], x, p, hit, chipx, chipy, pha, pi, rng, fef_cum, kStageThreads, dd);
          }
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = p.x; d[4] = p.y; d[5] = p.z;
        // ids produced by the earlier stages travel through the queue (coalesced loads here instead of dependent
        // gathers when a row is flushed): flags use bits 0..9, shell and order ride in the upper half
        u[0] = slot;
        u[1] = flags | ((uint32_t) in.shell[i] << 16) | (((uint32_t) (uint8_t) in.order[i]) << 24);
        u[2] = ((uint32_t) (ccd & 0xFF)) | (((uint32_t) (uint16_t) pha) << 8);
        u[3] = __float_as_uint (chipx); u[4] = __float_as_uint (chipy); u[5] = __float_as_uint (pi);
        u[6] = in.sorders[i];
        return flags;
     };
   auto write_row = [&] (unsigned long long j, const double *d, const uint32_t *u)
     {
        out.x0[j] = d[0]; out.x1[j] = d[1]; out.x2[j] = d[2];
        out.p0[j] = d[3]; out.p1[j] = d[4]; out.p2[j] = d[5];
        out.flags[j] = u[1] & 0xFFFFu;
        out.ccd[j] = (int8_t) (u[2] & 0xFFu);
        out.pha[j] = (int16_t) (uint16_t) (u[2] >> 8);
        out.region[j] = 0; out.upix[j] = 0.f; out.vpix[j] = 0.f;
        out.chipx[j] = __uint_as_float (u[3]); out.chipy[j] = __uint_as_float (u[4]); out.pi[j] = __uint_as_float (u[5]);
     };
   auto flush_entry = [&] (uint32_t pos, unsigned long long j)
     {
        double d[ND]; uint32_t u[NU];
#pragma unroll
        for (int k = 0; k < ND; k++) d[k] = q.d[k][pos];
#pragma unroll
        for (int k = 0; k < NU; k++) u[k] = q.u[k][pos];
        write_row (j, d, u);
        out.slot[j] = u[0];
        out.shell[j] = (uint8_t) ((u[1] >> 16) & 0xFFu);
        out.order[j] = (int8_t) (u[1] >> 24);
        out.sorders[j] = u[NU - 1];
     };
   auto in_place = [&] (unsigned long long i, const double *d, const uint32_t *u, uint32_t) { write_row (i, d, u); };
   run_stage<ND, NU> (a, q, trace, flush_entry, in_place);
}

// K3 (HRC-S) -----------------------------------------------------------------------------------
template <bool DET>
__global__ void __launch_bounds__ (kStageThreads) k3_hrc (const __grid_constant__ StageArgs a)
{
   constexpr int ND = 6, NU = 8;
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   stage_blob (smem, a.blob, a.blob_bytes, &bar);
   const HrcDev &D = reinterpret_cast<const K3HrcBlob *> (smem)->D;
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, a.blob_bytes);
   const PhotonSoA &in = a.in, &out = a.out;

   auto trace = [&] (unsigned long long i, double *d, uint32_t *u) -> uint32_t
     {
        Vec3 x = v_make (in.x0[i], in.x1[i], in.x2[i]), p = v_make (in.p0[i], in.p1[i], in.p2[i]);
        int ccd = -1, region = 0; float ypix = 0, zpix = 0, upix = 0, vpix = 0; int16_t pha = 0;
        Rng rng;
        const uint32_t slot = in.slot[i];
        double energy; uint64_t ray;
        rc_load (a.rc, slot, energy, ray);
        rng.init (a.seed, ray, MARXB200_STAGE_DETECTOR);
        DetDither dd = {0.0, 0.0, 0.0};
        if (DET) { dd.dy = a.rc.ddy[slot]; dd.dz = a.rc.ddz[slot]; dd.dtheta = a.rc.ddth[slot]; }
        uint32_t flags = hrc_s_detect<DET> (D, energy, x, p, ccd, region, ypix, zpix, upix, vpix, pha, rng, dd);
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = p.x; d[4] = p.y; d[5] = p.z;
        u[0] = slot;
        u[1] = flags | ((uint32_t) in.shell[i] << 16) | (((uint32_t) (uint8_t) in.order[i]) << 24);
        u[2] = ((uint32_t) (ccd & 0xFF)) | (((uint32_t) (uint16_t) pha) << 8) | (((uint32_t) (region & 0xFF)) << 24);
        u[3] = __float_as_uint (ypix); u[4] = __float_as_uint (zpix);
        u[5] = __float_as_uint (upix); u[6] = __float_as_uint (vpix);
        u[7] = in.sorders[i];
        return flags;
     };
   auto write_row = [&] (unsigned long long j, const double *d, const uint32_t *u)
     {
        out.x0[j] = d[0]; out.x1[j] = d[1]; out.x2[j] = d[2];
        out.p0[j] = d[3]; out.p1[j] = d[4]; out.p2[j] = d[5];
        out.flags[j] = u[1] & 0xFFFFu;
        out.ccd[j] = (int8_t) (u[2] & 0xFFu);
        out.pha[j] = (int16_t) (uint16_t) ((u[2] >> 8) & 0xFFFFu);
        out.region[j] = (int8_t) (u[2] >> 24);
        out.chipx[j] = __uint_as_float (u[3]); out.chipy[j] = __uint_as_float (u[4]);
        out.upix[j] = __uint_as_float (u[5]); out.vpix[j] = __uint_as_float (u[6]);
        out.pi[j] = 0.f;
     };
   auto flush_entry = [&] (uint32_t pos, unsigned long long j)
     {
        double d[ND]; uint32_t u[NU];
#pragma unroll
        for (int k = 0; k < ND; k++) d[k] = q.d[k][pos];
#pragma unroll
        for (int k = 0; k < NU; k++) u[k] = q.u[k][pos];
        write_row (j, d, u);
        out.slot[j] = u[0];
        out.shell[j] = (uint8_t) ((u[1] >> 16) & 0xFFu);
        out.order[j] = (int8_t) (u[1] >> 24);
        out.sorders[j] = u[NU - 1];
     };
   auto in_place = [&] (unsigned long long i, const double *d, const uint32_t *u, uint32_t) { write_row (i, d, u); };
   run_stage<ND, NU> (a, q, trace, flush_entry, in_place);
}

// K0 + K1a fused -------------------------------------------------------------------------------
// marxb200_trace's first kernel: source draw, arrival time, dither AND HRMA phase A for one ray per thread, so
// that the 52 % of the rays that miss the paraboloid never touch HBM (the separate k0_source / k1_hrma<0>
// pair writes 77 B and re-reads 36 B for every generated ray).  Persistent 256-thread CTAs walk the 256-ray
// tiles of the canonical time sum (k0_time_sums / k0_time_scan must have run); survivors go through the
// warp-private re-packing queue straight into the list that k1_hrma<1> consumes.
__global__ void __launch_bounds__ (kTile, MX_K01_MINBLOCKS) k01_source_hrma (const __grid_constant__ SourceArgs a, const __grid_constant__ StageArgs st)
{
   constexpr int ND = 6, NU = 2;
   extern __shared__ __align__ (128) unsigned char smem[];
   __shared__ __align__ (8) unsigned long long bar;
   stage_blob (smem, st.blob, (uint32_t) sizeof (K1Blob), &bar);
   const HrmaDev &H = reinterpret_cast<const K1Blob *> (smem)->H;
   WarpQueue<ND, NU> &q = my_queue<ND, NU> (smem, (uint32_t) sizeof (K1Blob));
   const PhotonSoA &out = st.out;
   const uint32_t lane = threadIdx.x & 31;
   uint32_t head = 0, count = 0;

   auto flush = [&] (uint32_t n_flush)
     {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd (st.n_out, (unsigned long long) n_flush);
        base = __shfl_sync (0xffffffffu, base, 0);
        if (lane < n_flush)
          {
             const uint32_t pos = (head + lane) & (kQueueCap - 1);
             const unsigned long long j = base + lane;
             out.x0[j] = q.d[0][pos]; out.x1[j] = q.d[1][pos]; out.x2[j] = q.d[2][pos];
             out.p0[j] = q.d[3][pos]; out.p1[j] = q.d[4][pos]; out.p2[j] = q.d[5][pos];
             const uint32_t slot = q.u[0][pos], pk = q.u[1][pos];
             out.slot[j] = slot;
             out.shell[j] = (uint8_t) (pk & 0xFFu); out.pha[j] = (int16_t) (pk >> 8);
             out.flags[j] = 0; out.order[j] = 0; out.sorders[j] = 0;
          }
        head = (head + n_flush) & (kQueueCap - 1);
        count -= n_flush;
        __syncwarp ();
     };

   // POINT source, no roll dither: the rolled source direction is one vector for the whole run
   const bool const_roll = dither_roll_is_constant (a.S, a.D);
   Vec3 rolled = v_make (0, 0, 0);
   if (const_roll) rolled = dither_roll ((double) (float) a.D.nominal_roll, v_make (a.S.p[0], a.S.p[1], a.S.p[2]));

   // Tiles are handed out by a ticket counter (st.ticket, or a static stride when it is null): a CTA that becomes resident late --
   // an NCCL kernel of the merge stream held its slot, or another stream's kernel did -- finds the work gone instead of running a
   // full 1/grid share after everybody else has finished.  The next ticket is drawn before the tile is traced, so that the
   // atomic's round trip is not on the tile's critical path.
   __shared__ unsigned long long s_tile;
   const uint64_t n_tiles = (a.n + kTile - 1) / kTile;
   const bool dynamic = (st.ticket != nullptr);
   // tile: what every thread traces now; upcoming: the tile after it -- known to thread 0 only until the scan's first barrier
   // hands it to everybody (no barrier of its own)
   unsigned long long tile = blockIdx.x, upcoming = 0;
   if (dynamic)
     {
        if (threadIdx.x == 0) s_tile = atomicAdd (st.ticket, 1ull);
        __syncthreads ();
        tile = s_tile;
        if (threadIdx.x == 0) upcoming = atomicAdd (st.ticket, 1ull);
     }
   while (tile < n_tiles)
     {
        const uint64_t i = tile * kTile + threadIdx.x;
        const bool valid = i < a.n;
        Rng rng; double energy = 0.0, dt = 0.0; Vec3 p = v_make (0, 0, 0);
        if (valid) k0_draw (a, i, rng, energy, p, dt);
        double total;
        double t;
        unsigned long long after = 0;
        if (dynamic)
          {
             t = tile_inclusive_scan_bcast (dt, total, upcoming) + a.tile_base[tile];
             if (threadIdx.x == 0) after = atomicAdd (st.ticket, 1ull);      // consumed one whole tile from now
          }
        else
          {
             t = tile_inclusive_scan (dt, total) + a.tile_base[tile];
             upcoming = tile + gridDim.x;
          }
        tile = upcoming; upcoming = after;
        bool alive = false;
        Vec3 x = v_make (0, 0, 0); uint32_t shell = 0; float dra = 0.f, ddec = 0.f, droll = 0.f;
        if (valid)
          {
             dither_ray (a.D, rng, t, p, dra, ddec, droll, const_roll ? &rolled : nullptr);
             rng.init (a.seed, a.first_ray + i, MARXB200_STAGE_MIRROR);
             alive = (0 == hrma_phase_a (H, st.source_distance, x, p, shell, rng));
             // the per-ray constants go straight to their slot (slot == i).  They are stored for EVERY ray, dead ones
             // included: full, coalesced 256-byte rows per column (36 B per generated ray), whereas storing only the
             // survivors' made every sector a partial write that the L2 had to fill from DRAM first (ncu: 539 MB of
             // reads in a kernel that reads nothing)
             rc_store (a.rc, i, energy, a.first_ray + i); a.rc.time[i] = t;
             a.rc.dra[i] = dra; a.rc.ddec[i] = ddec; a.rc.droll[i] = droll;
          }
        const uint32_t ballot = __ballot_sync (0xffffffffu, alive);
        if (alive)
          {
             const uint32_t pos = (head + count + __popc (ballot & ((1u << lane) - 1u))) & (kQueueCap - 1);
             q.d[0][pos] = x.x; q.d[1][pos] = x.y; q.d[2][pos] = x.z;
             q.d[3][pos] = p.x; q.d[4][pos] = p.y; q.d[5][pos] = p.z;
             q.u[0][pos] = (uint32_t) i;
             q.u[1][pos] = shell | ((rng.draw & 0x3FFFu) << 8);

          }
        count += __popc (ballot);
        __syncwarp ();
        if (count >= 32) flush (32);
     }
   if (count > 0) flush (count);
}

// ---------------------------------------------------------------------------------------------
// arrival-order restoration: the live list's `slot` keys are distinct indices into the batch, so the rank
// of a photon is the number of set bits below its slot in a bitmap of the live slots.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256) order_mark (OrderArgs a)
{
   const unsigned long long n = *a.n_live;
   for (unsigned long long s = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (unsigned long long) gridDim.x * blockDim.x)
     {
        const uint32_t key = a.in.slot[s];
        atomicOr (a.bitmap + (key >> 5), 1u << (key & 31u));
     }
}
// one block per 1024 bitmap words: exclusive popcount prefix inside the block + the block total
__global__ void __launch_bounds__ (256) order_scan_words (OrderArgs a, uint32_t n_words)
{
   __shared__ uint32_t warp_tot[8];
   const uint32_t w0 = blockIdx.x * 1024u + threadIdx.x * 4u;
   uint32_t c[4], sum = 0;
#pragma unroll
   for (int k = 0; k < 4; k++) { c[k] = (w0 + k < n_words) ? __popc (a.bitmap[w0 + k]) : 0u; sum += c[k]; }
   const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   uint32_t incl = sum;
#pragma unroll
   for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync (0xffffffffu, incl, o); if (lane >= (uint32_t) o) incl += v; }
   if (lane == 31) warp_tot[warp] = incl;
   __syncthreads ();
   uint32_t off = 0, tot = 0;
#pragma unroll
   for (int w = 0; w < 8; w++) { if (w < (int) warp) off += warp_tot[w]; tot += warp_tot[w]; }
   uint32_t run = off + incl - sum;
#pragma unroll
   for (int k = 0; k < 4; k++) { if (w0 + k < n_words) a.word_prefix[w0 + k] = run; run += c[k]; }
   if (threadIdx.x == 0) a.block_prefix[blockIdx.x] = tot;
}
// one block: exclusive scan of the block totals (sequential chunks per thread, then a warp/block scan)
__global__ void __launch_bounds__ (1024) order_scan_blocks (OrderArgs a, uint32_t n_blocks)
{
   __shared__ uint32_t part[1024];
   const uint32_t per = (n_blocks + 1023u) / 1024u;
   const uint32_t b0 = threadIdx.x * per;
   uint32_t sum = 0;
   for (uint32_t k = 0; k < per; k++) if (b0 + k < n_blocks) sum += a.block_prefix[b0 + k];
   part[threadIdx.x] = sum;
   __syncthreads ();
   if (threadIdx.x == 0) { uint32_t acc = 0; for (int t = 0; t < 1024; t++) { uint32_t v = part[t]; part[t] = acc; acc += v; } }
   __syncthreads ();
   uint32_t acc = part[threadIdx.x];
   for (uint32_t k = 0; k < per; k++) if (b0 + k < n_blocks) { uint32_t v = a.block_prefix[b0 + k]; a.block_prefix[b0 + k] = acc; acc += v; }
}
// rank of every live photon -> inverse permutation perm[rank] = position in the unordered list (4-byte scattered writes)
__global__ void __launch_bounds__ (256) order_rank (OrderArgs a)
{
   const unsigned long long n = *a.n_live;
   for (unsigned long long s = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (unsigned long long) gridDim.x * blockDim.x)
     {
        const uint32_t key = a.in.slot[s], w = key >> 5;
        const unsigned long long j = (unsigned long long) a.block_prefix[w >> 10] + a.word_prefix[w]
                                     + __popc (a.bitmap[w] & ((1u << (key & 31u)) - 1u));
        a.perm[j] = (uint32_t) s;
     }
}
// gather through the permutation: every column store of a warp is a full, coalesced row; the loads are a local
// permutation of the unordered list (a stage kernel emits survivors in completion order, so neighbours in arrival
// order sit within a few hundred entries of each other) and are absorbed by L1/L2.  The first version scattered the
// 23 columns instead: partial-sector writes made it run at 1.1 TB/s (1.15 ms for the 5.1e6 events of a C1 batch).
//
// PACK: the kernel also emits the file images of the row's columns (what egress_pack would produce from the list it writes) into
// a staging buffer, straight from the registers that hold the row: the packed egress / the multi-GPU merge of a run that asks for
// the same columns batch after batch then needs no conversion pass of its own (marxb200.cu "pre-pack").
__device__ __forceinline__ uint32_t bswap32 (uint32_t v) { return __byte_perm (v, 0u, 0x0123); }
__device__ __forceinline__ uint16_t bswap16 (uint16_t v) { return (uint16_t) ((v << 8) | (v >> 8)); }
struct EventRow
{
   double energy, x0, x1, x2, p0, p1, p2, time;
   uint64_t ray;
   float chipx, chipy, pi, upix, vpix, dra, ddec, droll, ddy, ddz, ddth;
   uint32_t sorders;
   int16_t pha; uint8_t shell, region; int8_t order, ccd;
};
// the file image of one column of a row: its bits, right-aligned (marxio.c:217-322: the casts of the MAKE_WRITE_*_FUNC macros, big endian)
__device__ __forceinline__ uint32_t egress_from_row (const EventRow &r, int kind, double start_time, double total_time)
{
   float f = 0.0f;
   switch (kind)
     {
      case EGRESS_PI: f = r.pi; break;
      case EGRESS_ENERGY: f = (float) r.energy; break;
      case EGRESS_TIME: f = (float) ((r.time - start_time) + total_time); break;
      case EGRESS_XPOS: f = (float) r.x0; break;
      case EGRESS_YPOS: f = (float) r.x1; break;
      case EGRESS_ZPOS: f = (float) r.x2; break;
      case EGRESS_XCOS: f = (float) r.p0; break;
      case EGRESS_YCOS: f = (float) r.p1; break;
      case EGRESS_ZCOS: f = (float) r.p2; break;
      case EGRESS_CHIPX: f = r.chipx; break;
      case EGRESS_CHIPY: f = r.chipy; break;
      case EGRESS_HRC_U: f = r.upix; break;
      case EGRESS_HRC_V: f = r.vpix; break;
      case EGRESS_SKY_RA: f = r.dra; break;
      case EGRESS_SKY_DEC: f = r.ddec; break;
      case EGRESS_SKY_ROLL: f = r.droll; break;
      case EGRESS_DET_DY: f = r.ddy; break;
      case EGRESS_DET_DZ: f = r.ddz; break;
      case EGRESS_DET_THETA: f = r.ddth; break;
      case EGRESS_TAG: return bswap32 ((uint32_t) r.ray);
      case EGRESS_PHA: return bswap16 ((uint16_t) r.pha);
      case EGRESS_MIRROR: return bswap16 ((uint16_t) r.shell);
      case EGRESS_CCD: return (unsigned char) r.ccd;
      case EGRESS_REGION: return (unsigned char) r.region;
      case EGRESS_ORDER: return (unsigned char) r.order;
      case EGRESS_ORDER1: return r.sorders & 0xFFu;
      case EGRESS_ORDER2: return (r.sorders >> 8) & 0xFFu;
      case EGRESS_ORDER3: return (r.sorders >> 16) & 0xFFu;
      case EGRESS_ORDER4: return r.sorders >> 24;
      default: return 0;
     }
   return bswap32 (__float_as_uint (f));
}
__device__ __forceinline__ void egress_store (unsigned char *base, int kind, uint64_t i, uint32_t v)
{
   switch (kind)
     {
      case EGRESS_PHA: case EGRESS_MIRROR:
        reinterpret_cast<uint16_t *> (base)[i] = (uint16_t) v; break;
      case EGRESS_CCD: case EGRESS_REGION: case EGRESS_ORDER: case EGRESS_ORDER1: case EGRESS_ORDER2: case EGRESS_ORDER3: case EGRESS_ORDER4:
        base[i] = (unsigned char) v; break;
      default:
        reinterpret_cast<uint32_t *> (base)[i] = v; break;
     }
}
template <bool PACK>
__global__ void __launch_bounds__ (256) order_gather (OrderArgs a, const __grid_constant__ PackArgs pk)
{
   const unsigned long long n = *a.n_live;
   const PhotonSoA &in = a.in, &out = a.out;
   double start_time = 0.0;
   if (PACK && (pk.dev_start_time != nullptr)) start_time = *pk.dev_start_time;
   for (unsigned long long j = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (unsigned long long) gridDim.x * blockDim.x)
     {
        const uint32_t s = a.perm[j], key = in.slot[s];
        const RayConst &rc = a.rc;
        EventRow r;
        rc_load (rc, key, r.energy, r.ray);
        r.x0 = in.x0[s]; r.x1 = in.x1[s]; r.x2 = in.x2[s]; r.p0 = in.p0[s]; r.p1 = in.p1[s]; r.p2 = in.p2[s];
        r.time = rc.time[key];
        r.dra = rc.dra[key]; r.ddec = rc.ddec[key]; r.droll = rc.droll[key];
        r.ddy = 0.0f; r.ddz = 0.0f; r.ddth = 0.0f;
        if (out.ddy != nullptr) { r.ddy = rc.ddy[key]; r.ddz = rc.ddz[key]; r.ddth = rc.ddth[key]; }   // null: detector dither not live
        r.chipx = in.chipx[s]; r.chipy = in.chipy[s]; r.pi = in.pi[s];
        r.pha = in.pha[s]; r.shell = in.shell[s]; r.order = in.order[s]; r.ccd = in.ccd[s];
        r.upix = in.upix[s]; r.vpix = in.vpix[s]; r.sorders = in.sorders[s]; r.region = in.region[s];
        out.energy[j] = r.energy;
        out.x0[j] = r.x0; out.x1[j] = r.x1; out.x2[j] = r.x2;
        out.p0[j] = r.p0; out.p1[j] = r.p1; out.p2[j] = r.p2;
        out.time[j] = r.time; out.aux[j] = in.aux[s];
        out.ray[j] = r.ray; out.slot[j] = key; out.flags[j] = in.flags[s];
        out.dra[j] = r.dra; out.ddec[j] = r.ddec; out.droll[j] = r.droll;
        if (out.ddy != nullptr) { out.ddy[j] = r.ddy; out.ddz[j] = r.ddz; out.ddth[j] = r.ddth; }
        out.chipx[j] = r.chipx; out.chipy[j] = r.chipy; out.pi[j] = r.pi;
        out.pha[j] = r.pha; out.shell[j] = r.shell; out.order[j] = r.order; out.ccd[j] = r.ccd;
        out.upix[j] = r.upix; out.vpix[j] = r.vpix; out.sorders[j] = r.sorders; out.region[j] = r.region;
        if (PACK && (j < pk.max_rows))
          for (int c = 0; c < pk.plan.num_cols; c++)
            egress_store (pk.dst + pk.plan.offset[c], pk.plan.kind[c], j, egress_from_row (r, pk.plan.kind[c], start_time, pk.total_time));
     }
}
void launch_restore_order (const OrderArgs &a, int num_sms, cudaStream_t s, int *n_launches, const PackArgs *pack)
{
   const uint32_t n_words = (uint32_t) (a.n_slots / 32 + 1), n_blocks = (n_words + 1023u) / 1024u;
   const int grid = num_sms * 8;
   order_mark<<<grid, 256, 0, s>>> (a);
   order_scan_words<<<n_blocks, 256, 0, s>>> (a, n_words);
   order_scan_blocks<<<1, 1024, 0, s>>> (a, n_blocks);
   order_rank<<<grid, 256, 0, s>>> (a);
   if (pack != nullptr) order_gather<true><<<grid, 256, 0, s>>> (a, *pack);
   else
     {
        PackArgs none;
        memset (&none, 0, sizeof (none));
        order_gather<false><<<grid, 256, 0, s>>> (a, none);
     }
   if (n_launches) *n_launches = 5;
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
        r.y_pixel = in.chipx[i]; r.z_pixel = in.chipy[i]; r.u_pixel = in.upix[i]; r.v_pixel = in.vpix[i];
        r.dither_ra = in.dra[i]; r.dither_dec = in.ddec[i]; r.dither_roll = in.droll[i];
        const bool det = (in.ddy != nullptr);          // null: no detector dither in this run (NONE / INTERNAL models)
        r.dither_dy = det ? in.ddy[i] : 0.f; r.dither_dz = det ? in.ddz[i] : 0.f; r.dither_dtheta = det ? in.ddth[i] : 0.f;
        r.pi = in.pi[i];
        r.pulse_height = in.pha[i];
        r.mirror_shell = in.shell[i];
        r.ccd_num = in.ccd[i];
        r.detector_region = in.region[i];
        r.order = in.order[i];
        const uint32_t so = in.sorders[i];
        r.support_orders[0] = (int8_t) (so & 0xFFu); r.support_orders[1] = (int8_t) ((so >> 8) & 0xFFu);
        r.support_orders[2] = (int8_t) ((so >> 16) & 0xFFu); r.support_orders[3] = (int8_t) (so >> 24);
        r.tag = (uint32_t) in.ray[i];
        // 17 aligned 8-byte stores per record
        const uint64_t *src = reinterpret_cast<const uint64_t *> (&r);
        uint64_t *dst = reinterpret_cast<uint64_t *> (aos + i);
#pragma unroll
        for (int k = 0; k < (int) (sizeof (marxb200_photon_attr) / 8); k++) dst[k] = src[k];
     }
}

__global__ void __launch_bounds__ (256) aos_to_soa (const marxb200_photon_attr *aos, const uint64_t *ray_ids, uint64_t n,
                                                    PhotonSoA out, RayConst rc, double start_time)
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
        out.slot[i] = (uint32_t) i;
        out.flags[i] = r.flags;
        out.dra[i] = r.dither_ra; out.ddec[i] = r.dither_dec; out.droll[i] = r.dither_roll;
        rc_store (rc, i, r.energy, ray_ids ? ray_ids[i] : (uint64_t) r.tag); rc.time[i] = r.arrival_time + start_time;
        rc.dra[i] = r.dither_ra; rc.ddec[i] = r.dither_dec; rc.droll[i] = r.dither_roll;
        out.ddy[i] = r.dither_dy; out.ddz[i] = r.dither_dz; out.ddth[i] = r.dither_dtheta;
        rc.ddy[i] = r.dither_dy; rc.ddz[i] = r.dither_dz; rc.ddth[i] = r.dither_dtheta;
        out.chipx[i] = r.y_pixel; out.chipy[i] = r.z_pixel; out.pi[i] = r.pi;
        out.pha[i] = r.pulse_height;
        out.shell[i] = (uint8_t) r.mirror_shell;
        out.order[i] = r.order;
        out.ccd[i] = r.ccd_num;
        out.region[i] = r.detector_region;
        out.upix[i] = r.u_pixel; out.vpix[i] = r.v_pixel;
        out.sorders[i] = ((uint32_t) (uint8_t) r.support_orders[0]) | (((uint32_t) (uint8_t) r.support_orders[1]) << 8)
                         | (((uint32_t) (uint8_t) r.support_orders[2]) << 16) | (((uint32_t) (uint8_t) r.support_orders[3]) << 24);
     }
}

// ---------------------------------------------------------------------------------------------
// ExposureTime truncation (source.c:323-334): keep rays up to and including the first one whose arrival time,
// counted from the batch start, reaches the exposure left.  Arrival times are a running sum, hence monotone: one
// thread bisects.  Updates the generated count and the running end time (pt->total_time, source.c:377-381).
// ---------------------------------------------------------------------------------------------
// inclusive = 1: the ExposureTime cut (keep the crossing ray; times counted from the batch start);
// inclusive = 0: the end of an ASPSOL file (dither.c:296-301, 361-369: the first ray at or beyond the last state is not
// dithered and ends the batch; `limit` is an absolute time)
__global__ void exposure_truncate (const double *time, unsigned long long *n_ptr, double *dev_times, double limit, int inclusive)
{
   if ((blockIdx.x != 0) || (threadIdx.x != 0)) return;
   const unsigned long long n = *n_ptr;
   if (n == 0) return;
   const double start = inclusive ? dev_times[0] : 0.0;
   unsigned long long lo = 0, hi = n;            // first index with time - start >= limit, or n
   while (lo < hi)
     {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (time[mid] - start >= limit) hi = mid; else lo = mid + 1;
     }
   const unsigned long long keep = inclusive ? ((lo < n) ? lo + 1 : n) : lo;
   *n_ptr = keep;
   if (keep < n) dev_times[1] = (keep > 0) ? time[keep - 1] : dev_times[0];
}
void launch_exposure_truncate (const PhotonSoA &buf, unsigned long long *n, double *dev_times, double limit, int inclusive, cudaStream_t s)
{
   exposure_truncate<<<1, 32, 0, s>>> (buf.time, n, dev_times, limit, inclusive);
}

// ---------------------------------------------------------------------------------------------
// Bulk event egress in the reference's on-disk column format (marxio.c:217-322): each selected column of the live
// list is converted to its file type (float32 / int16 / int32 / int8, the casts of the MAKE_WRITE_*_FUNC macros) and
// byte-swapped to big endian (JDMwrite_float32/int16/int32) on the device, so that the host only appends the packed
// images to the column files.  One thread per (photon, column) pair would waste the SoA locality; instead each
// thread handles one photon and walks the (uniform) column list, so that every store instruction of a warp hits
// consecutive elements of one packed column.
// ---------------------------------------------------------------------------------------------
// the file image of one (photon, column) element: its bits, right-aligned (4, 2 or 1 bytes wide by kind)
__device__ __forceinline__ uint32_t egress_load (const PhotonSoA &in, int kind, uint64_t i, double start_time, double total_time)
{
   float f = 0.0f;
   switch (kind)
     {
      case EGRESS_PI: f = in.pi[i]; break;
      case EGRESS_ENERGY: f = (float) in.energy[i]; break;
      // write_time: (float) (at->arrival_time + total_time) with arrival_time = time - batch start (soa_to_aos)
      case EGRESS_TIME: f = (float) ((in.time[i] - start_time) + total_time); break;
      case EGRESS_XPOS: f = (float) in.x0[i]; break;
      case EGRESS_YPOS: f = (float) in.x1[i]; break;
      case EGRESS_ZPOS: f = (float) in.x2[i]; break;
      case EGRESS_XCOS: f = (float) in.p0[i]; break;
      case EGRESS_YCOS: f = (float) in.p1[i]; break;
      case EGRESS_ZCOS: f = (float) in.p2[i]; break;
      case EGRESS_CHIPX: f = in.chipx[i]; break;
      case EGRESS_CHIPY: f = in.chipy[i]; break;
      case EGRESS_HRC_U: f = in.upix[i]; break;
      case EGRESS_HRC_V: f = in.vpix[i]; break;
      case EGRESS_SKY_RA: f = in.dra[i]; break;
      case EGRESS_SKY_DEC: f = in.ddec[i]; break;
      case EGRESS_SKY_ROLL: f = in.droll[i]; break;
      case EGRESS_DET_DY: f = in.ddy ? in.ddy[i] : 0.0f; break;        // null: 0 for the NONE / INTERNAL models (dither.c:177-179)
      case EGRESS_DET_DZ: f = in.ddz ? in.ddz[i] : 0.0f; break;
      case EGRESS_DET_THETA: f = in.ddth ? in.ddth[i] : 0.0f; break;
      case EGRESS_TAG: return bswap32 ((uint32_t) in.ray[i]);
      case EGRESS_PHA: return bswap16 ((uint16_t) in.pha[i]);
      case EGRESS_MIRROR: return bswap16 ((uint16_t) in.shell[i]);
      case EGRESS_CCD: return (unsigned char) in.ccd[i];
      case EGRESS_REGION: return (unsigned char) in.region[i];
      case EGRESS_ORDER: return (unsigned char) in.order[i];
      case EGRESS_ORDER1: return in.sorders[i] & 0xFFu;
      case EGRESS_ORDER2: return (in.sorders[i] >> 8) & 0xFFu;
      case EGRESS_ORDER3: return (in.sorders[i] >> 16) & 0xFFu;
      case EGRESS_ORDER4: return in.sorders[i] >> 24;
      default: return 0;
     }
   return bswap32 (__float_as_uint (f));
}
// The column walk is cut into groups of kEgressGroup columns: all loads of a group are issued before its first store (a plain
// load-store-load-store walk leaves one request in flight per thread and ran at a third of the HBM rate: 0.10 ms per 1.2e6 events).
constexpr int kEgressGroup = 8;
__global__ void __launch_bounds__ (256) egress_pack (PhotonSoA in, const __grid_constant__ EgressPlan plan, const unsigned long long *n_ptr, uint64_t max_n,
                                                     unsigned char *dst, const double *dev_start_time, double total_time)
{
   const uint64_t n = min ((uint64_t) *n_ptr, max_n);
   const double start_time = (dev_start_time != nullptr) ? *dev_start_time : 0.0;     // null: TIME = absolute time + total_time
   for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
     {
#pragma unroll
        for (int g = 0; g < kMaxEgressCols; g += kEgressGroup)
          {
             if (g >= plan.num_cols) break;
             uint32_t v[kEgressGroup];
#pragma unroll
             for (int k = 0; k < kEgressGroup; k++)
               v[k] = (g + k < plan.num_cols) ? egress_load (in, plan.kind[g + k], i, start_time, total_time) : 0u;
#pragma unroll
             for (int k = 0; k < kEgressGroup; k++)
               if (g + k < plan.num_cols) egress_store (dst + plan.offset[g + k], plan.kind[g + k], i, v[k]);
          }
     }
}
void launch_egress_pack (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, const EgressPlan &plan, void *dst,
                         const double *dev_start_time, double total_time, cudaStream_t s)
{
   if ((max_n == 0) || (plan.num_cols == 0)) return;
   unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (max_n + 255) / 256);
   egress_pack<<<grid, 256, 0, s>>> (in, plan, n, max_n, (unsigned char *) dst, dev_start_time, total_time);
}

// ---------------------------------------------------------------------------------------------
// Event tallies: exact histograms of the live list (order populations, PHA / PI / energy spectra, chip images ...) kept
// on the device so that G GPUs merge them with one all-reduce instead of moving events.  Counts are
// integers: the result does not depend on the order of the list, the grid or the GPU count.  Small histograms are
// accumulated per CTA in shared memory (u32) and flushed with one 64-bit atomic per non-empty bin; large ones (images)
// go straight to global atomics, which L2 resolves.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kTallySmemBins = 8192;
__device__ __forceinline__ bool tally_value (const PhotonSoA &in, int column, unsigned long long i, double &v)
{
   switch (column)
     {
      case TALLY_ENERGY: v = in.energy[i]; return true;
      case TALLY_TIME: v = in.time[i]; return true;
      case TALLY_PHA: v = (double) in.pha[i]; return true;
      case TALLY_PI: v = (double) in.pi[i]; return true;
      case TALLY_ORDER: v = (double) in.order[i]; return true;
      case TALLY_CCD: v = (double) in.ccd[i]; return true;
      case TALLY_SHELL: v = (double) in.shell[i]; return true;
      case TALLY_CHIPX: v = (double) in.chipx[i]; return true;
      case TALLY_CHIPY: v = (double) in.chipy[i]; return true;
      case TALLY_YPOS: v = in.x1[i]; return true;
      case TALLY_ZPOS: v = in.x2[i]; return true;
     }
   return false;
}
__global__ void __launch_bounds__ (256) tally_events (PhotonSoA in, const unsigned long long *n_ptr, uint64_t max_n, TallyPlan plan,
                                                      unsigned long long *bins, uint32_t total_bins, int use_smem)
{
   extern __shared__ uint32_t local[];
   unsigned long long n = *n_ptr;
   if (n > max_n) n = max_n;
   if (use_smem)
     {
        for (uint32_t b = threadIdx.x; b < total_bins; b += blockDim.x) local[b] = 0;
        __syncthreads ();
     }
   for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x)
     {
        if ((in.flags[i] & 0xFFu) != 0) continue;            // in-place lists keep their dead rays
        uint32_t bin = 0;
        bool ok = true;
        for (int a = 0; a < plan.naxes; a++)
          {
             double v;
             ok = ok && tally_value (in, plan.ax[a].column, i, v);
             const double f = floor ((v - plan.ax[a].lo) * plan.ax[a].scale);
             ok = ok && (f >= 0.0) && (f < (double) plan.ax[a].nbins);      // NaN fails both
             if (!ok) break;
             bin = bin * plan.ax[a].nbins + (uint32_t) f;
          }
        if (!ok) continue;
        if (use_smem) atomicAdd (&local[bin], 1u);
        else atomicAdd (&bins[bin], 1ULL);
     }
   if (use_smem)
     {
        __syncthreads ();
        for (uint32_t b = threadIdx.x; b < total_bins; b += blockDim.x)
          if (local[b] != 0) atomicAdd (&bins[b], (unsigned long long) local[b]);
     }
}
void launch_tally (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, const TallyPlan &plan,
                   unsigned long long *bins, int num_sms, cudaStream_t s)
{
   if (max_n == 0) return;
   uint32_t total = plan.ax[0].nbins * ((plan.naxes > 1) ? plan.ax[1].nbins : 1u);
   const int use_smem = (total <= kTallySmemBins) ? 1 : 0;
   unsigned int grid = (unsigned int) min ((uint64_t) num_sms * 4, (max_n + 255) / 256);
   tally_events<<<grid, 256, use_smem ? total * sizeof (uint32_t) : 0, s>>> (in, n, max_n, plan, bins, total, use_smem);
}

// ---------------------------------------------------------------------------------------------
// FP64 roofline denominator: a DFMA-chain microbenchmark (SURVEY.md 8d asks for the FP64 peak to be MEASURED on the
// box; MEASURED_PEAKS.json only holds HBM and bf16).  8 independent chains per thread hide the DFMA latency.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256) fp64_peak_kernel (double *sink, int iters, double a, double b)
{
   double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
   for (int i = 0; i < iters; i++)
     {
#pragma unroll
        for (int k = 0; k < 8; k++)
          {
             x0 = __fma_rn (x0, a, b); x1 = __fma_rn (x1, a, b); x2 = __fma_rn (x2, a, b); x3 = __fma_rn (x3, a, b);
             x4 = __fma_rn (x4, a, b); x5 = __fma_rn (x5, a, b); x6 = __fma_rn (x6, a, b); x7 = __fma_rn (x7, a, b);
          }
     }
   const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
   if (r == 123.456) sink[0] = r;        // never true: keeps the chains alive
}
void launch_fp64_peak (double *sink, int grid, int iters, cudaStream_t s) { fp64_peak_kernel<<<grid, 256, 0, s>>> (sink, iters, 0.999999, 1e-9); }

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned int n_tiles_of (uint64_t n) { return (unsigned int) ((n + kTile - 1) / kTile); }

void launch_time_sums (const SourceArgs &a, cudaStream_t s)
{
   if (a.n == 0) return;
   k0_time_sums<<<n_tiles_of (a.n), kTile, 0, s>>> (a);
}
void launch_time_scan (const SourceArgs &a, cudaStream_t s, bool with_super)
{
   const unsigned int n_super = (unsigned int) ((n_tiles_of (a.n) + kSuperTile - 1) / kSuperTile);
   if (n_super && with_super) k0_time_super<<<n_super, kSuperTile, 0, s>>> (a);
   k0_time_bases<<<1, kSuperTile, 0, s>>> (a);
   if (n_super) k0_time_tiles<<<n_super, kSuperTile, 0, s>>> (a);
}
// the sharded scan in two halves around the all-gather of the super-tile sums (comm.cu)
void launch_time_super (const SourceArgs &a, cudaStream_t s)
{
   const unsigned int n_super = (unsigned int) ((n_tiles_of (a.n) + kSuperTile - 1) / kSuperTile);
   if (n_super) k0_time_super<<<n_super, kSuperTile, 0, s>>> (a);
}
void launch_time_bases_sharded (const SourceArgs &a, const double *all_sums, int rank, int world, uint32_t ns_blk, cudaStream_t s)
{
   const unsigned int n_super = (unsigned int) ((n_tiles_of (a.n) + kSuperTile - 1) / kSuperTile);
   k0_time_bases_sharded<<<1, kSuperTile, 0, s>>> (a, all_sums, rank, world, ns_blk);
   if (n_super) k0_time_tiles<<<n_super, kSuperTile, 0, s>>> (a);
}
void launch_source (const SourceArgs &a, cudaStream_t s)
{
   if (a.n == 0) return;
   k0_source<<<n_tiles_of (a.n), kTile, 0, s>>> (a);
}

uint32_t stage_smem_bytes (int stage, uint32_t blob_bytes, uint32_t seg2_bytes)
{
   const uint32_t warps = kStageThreads / 32, base = (blob_bytes + 127u) & ~127u;
   const uint32_t hdr = ((uint32_t) sizeof (K1Blob) + 127u) & ~127u;
   switch (stage)
     {
      case 10: return hdr + warps * (uint32_t) sizeof (WarpQueue<K1Shape<0>::ND, K1Shape<0>::NU>);
      case 11: return base + warps * (uint32_t) sizeof (WarpQueue<K1Shape<1>::ND, K1Shape<1>::NU>);
      case 12: return hdr + ((seg2_bytes + 127u) & ~127u) + warps * (uint32_t) sizeof (WarpQueue<K1Shape<2>::ND, K1Shape<2>::NU>);
      case 13: return hdr + (kTile / 32) * (uint32_t) sizeof (WarpQueue<6, 2>);
      case 14: return base + warps * (uint32_t) sizeof (WarpQueue<K1Shape<3>::ND, K1Shape<3>::NU>);
      case 15: return hdr + ((seg2_bytes + 127u) & ~127u) + warps * (uint32_t) sizeof (WarpQueue<K1Shape<4>::ND, K1Shape<4>::NU>);
      case 16: return hdr + ((seg2_bytes + 127u) & ~127u) + warps * (uint32_t) sizeof (WarpQueue<K1Shape<5>::ND, K1Shape<5>::NU>);
      case 2: return base + warps * (uint32_t) sizeof (WarpQueue<6, 3>);
      case 3: return base + warps * (uint32_t) sizeof (WarpQueue<6, 7>) + (uint32_t) (kMaxGauss * kStageThreads * sizeof (float));
      case 4: return base + warps * (uint32_t) sizeof (WarpQueue<6, 8>);
     }
   return base;
}
template <class K>
static int occupancy_grid (K kernel, int num_sms, uint32_t smem_bytes)
{
   int per_sm = 1;
   cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
   cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, kStageThreads, smem_bytes);
   if (per_sm < 1) per_sm = 1;
   return per_sm * num_sms;
}
int stage_grid_size (int stage, int num_sms, uint32_t blob_bytes, uint32_t seg2_bytes)
{
   const uint32_t smem = stage_smem_bytes (stage, blob_bytes, seg2_bytes);
   switch (stage)
     {
      case 10: return occupancy_grid (k1_hrma<0>, num_sms, smem);
      case 11: return occupancy_grid (k1_hrma<1>, num_sms, smem);
      case 12: return occupancy_grid (k1_hrma<2>, num_sms, smem);
      case 14: return occupancy_grid (k1_hrma<3>, num_sms, smem);
      case 15: return occupancy_grid (k1_hrma<4>, num_sms, smem);
      case 16: return occupancy_grid (k1_hrma<5>, num_sms, smem);
      case 2: cudaFuncSetAttribute (k2_grating<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              return occupancy_grid (k2_grating<0>, num_sms, smem);
      // the detector-dither variants run on the same grid (ticket-driven persistent kernels: any grid size is correct)
      case 3: cudaFuncSetAttribute (k3_acis<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              cudaFuncSetAttribute (k3_acis<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              cudaFuncSetAttribute (k3_acis<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              cudaFuncSetAttribute (k3_acis<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              cudaFuncSetAttribute (k3_acis<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              return occupancy_grid (k3_acis<false, 0>, num_sms, smem);
      case 4: cudaFuncSetAttribute (k3_hrc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
              return occupancy_grid (k3_hrc<false>, num_sms, smem);
     }
   return num_sms;
}
int fused_source_grid (int num_sms)
{
   int per_sm = 1;
   const uint32_t smem = stage_smem_bytes (13, 0);
   cudaFuncSetAttribute (k01_source_hrma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
   cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k01_source_hrma, kTile, smem);
   return (per_sm < 1 ? 1 : per_sm) * num_sms;
}
void launch_source_hrma (const SourceArgs &a, const StageArgs &st, int grid, cudaStream_t s)
{
   if (a.n == 0) return;
   k01_source_hrma<<<grid, kTile, stage_smem_bytes (13, 0), s>>> (a, st);
}
void launch_hrma (const StageArgs &a, int phase, int grid, cudaStream_t s)
{
   // phases 0, 1, 2 = A, B, C; 3, 4, 5 = B1, B2+C1, C2 (the cut behind the reflectivity tests)
   const bool two_seg = (phase == 2) || (phase == 4) || (phase == 5);
   const uint32_t smem = stage_smem_bytes ((phase < 3) ? 10 + phase : 11 + phase, a.blob_bytes, two_seg ? a.seg2_bytes : 0u);
   switch (phase)
     {
      case 0: k1_hrma<0><<<grid, kStageThreads, smem, s>>> (a); break;
      case 1: k1_hrma<1><<<grid, kStageThreads, smem, s>>> (a); break;
      case 2: k1_hrma<2><<<grid, kStageThreads, smem, s>>> (a); break;
      case 3: k1_hrma<3><<<grid, kStageThreads, smem, s>>> (a); break;
      case 4: k1_hrma<4><<<grid, kStageThreads, smem, s>>> (a); break;
      default: k1_hrma<5><<<grid, kStageThreads, smem, s>>> (a); break;
     }
}
void launch_flatfield (const StageArgs &a, int num_sms, cudaStream_t s)
{
   const uint32_t smem = (kStageThreads / 32) * (uint32_t) sizeof (WarpQueue<6, 1>);
   static int per_sm = 0;
   if (per_sm == 0)
     {
        cudaFuncSetAttribute (k1_flatfield, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k1_flatfield, kStageThreads, smem);
        if (per_sm < 1) per_sm = 1;
     }
   k1_flatfield<<<per_sm * num_sms, kStageThreads, smem, s>>> (a);
}
void launch_grating (const StageArgs &a, int grid, cudaStream_t s, int phase)
{
   if (phase == 1) k2_select<<<grid, 256, 0, s>>> (a);                   // grid: any size (grid-stride)
   else if (phase == 2) k2_grating<1><<<grid, kStageThreads, stage_smem_bytes (2, a.blob_bytes), s>>> (a);
   else k2_grating<0><<<grid, kStageThreads, stage_smem_bytes (2, a.blob_bytes), s>>> (a);
}
void launch_acis (const StageArgs &a, int grid, cudaStream_t s, int phase)
{
   const uint32_t smem = stage_smem_bytes (3, a.blob_bytes);
   if (a.det_dither)
     switch (phase)
       {
        case 1: k3_acis<true, 1><<<grid, kStageThreads, smem, s>>> (a); break;
        case 2: k3_acis<true, 2><<<grid, kStageThreads, smem, s>>> (a); break;
        default: k3_acis<true, 0><<<grid, kStageThreads, smem, s>>> (a); break;
       }
   else
     switch (phase)
       {
        case 1: k3_acis<false, 1><<<grid, kStageThreads, smem, s>>> (a); break;
        case 2: k3_acis<false, 2><<<grid, kStageThreads, smem, s>>> (a); break;
        default: k3_acis<false, 0><<<grid, kStageThreads, smem, s>>> (a); break;
       }
}
void launch_hrc (const StageArgs &a, int grid, cudaStream_t s)
{
   if (a.det_dither) k3_hrc<true><<<grid, kStageThreads, stage_smem_bytes (4, a.blob_bytes), s>>> (a);
   else k3_hrc<false><<<grid, kStageThreads, stage_smem_bytes (4, a.blob_bytes), s>>> (a);
}

void launch_soa_to_aos (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, void *aos,
                        const double *dev_start_time, cudaStream_t s)
{
   if (max_n == 0) return;
   unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (max_n + 255) / 256);
   soa_to_aos<<<grid, 256, 0, s>>> (in, n, max_n, (marxb200_photon_attr *) aos, dev_start_time);
}
void launch_aos_to_soa (const void *aos, const uint64_t *ray_ids, uint64_t n, const PhotonSoA &out, const RayConst &rc,
                        double start_time, cudaStream_t s)
{
   if (n == 0) return;
   unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (n + 255) / 256);
   aos_to_soa<<<grid, 256, 0, s>>> ((const marxb200_photon_attr *) aos, ray_ids, n, out, rc, start_time);
}

}  // namespace mx
