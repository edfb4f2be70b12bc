// pileup_kernels.cu -- sm_100a kernels of the ACIS pile-up model (mx_pileup.cuh; marx/src/marxpileup.c).
//
// One event per thread in every step; a step reads what the previous one wrote for the OTHER events of the same exposure frame,
// so the steps are separate launches on one stream (pu_island needs every pixel's sum, pu_detect every island's sum, pu_emit
// every island's verdict).  The frame of an event is a contiguous run of the list: neighbouring threads walk the same few
// dozen rows (unit-stride, L1-resident).  The output row of an emitted event comes from a prefix sum over the emit flags
// (pu_scan_tiles: inclusive inside 256-event tiles; pu_scan_totals: one CTA, exclusive over the tile totals) -- the
// reference's output order is frame by frame, REVERSE file order inside a frame, which pu_scatter derives from three prefix values.
#include <cuda_runtime.h>
#include "mx_pileup.cuh"
#include "mx_kernels.cuh"

namespace mx {

constexpr int kPuTile = 256;            // pu_rows_through (mx_pileup.cuh) shifts by 8

template <int STEP>
__global__ void __launch_bounds__ (kPuTile) pu_step (const __grid_constant__ PileupArgs a)
{
   for (uint64_t e = (uint64_t) blockIdx.x * kPuTile + threadIdx.x; e < a.n; e += (uint64_t) gridDim.x * kPuTile)
     {
        if (STEP == 0) pu_frames (a, e);
        if (STEP == 1) pu_store (a, e);
        if (STEP == 2) pu_island (a, e);
        if (STEP == 3) pu_detect (a, e);
        if (STEP == 4) pu_emit (a, e);
        if (STEP == 5) pu_scatter (a, e);
     }
}

// inclusive prefix sum of the emit flags inside each 256-event tile -> cum; the tile's total -> tile_sum
__global__ void __launch_bounds__ (kPuTile) pu_scan_tiles (const __grid_constant__ PileupArgs a)
{
   __shared__ uint32_t warp_sum[kPuTile / 32];
   const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
   for (uint64_t tile = blockIdx.x; tile * kPuTile < a.n; tile += gridDim.x)
     {
        const uint64_t i = tile * kPuTile + threadIdx.x;
        uint32_t s = (i < a.n) ? a.emit[i] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
          {
             const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, s, d);
             if (lane >= (uint32_t) d) s += o;
          }
        if (lane == 31u) warp_sum[w] = s;
        __syncthreads ();
        uint32_t before = 0;
        for (uint32_t k = 0; k < w; k++) before += warp_sum[k];
        s += before;
        if (i < a.n) a.cum[i] = s;
        if (threadIdx.x == kPuTile - 1) a.tile_sum[tile] = s;
        __syncthreads ();
     }
}

// exclusive prefix sum over the tile totals, in place (one CTA, sequential over chunks of 1024 tiles)
__global__ void __launch_bounds__ (1024) pu_scan_totals (const __grid_constant__ PileupArgs a)
{
   __shared__ uint32_t warp_sum[32];
   __shared__ uint32_t carry_s;
   const uint64_t n_tiles = (a.n + kPuTile - 1) / kPuTile;
   const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
   if (threadIdx.x == 0) carry_s = 0;
   __syncthreads ();
   for (uint64_t base = 0; base < n_tiles; base += 1024)
     {
        const uint64_t t = base + threadIdx.x;
        const uint32_t own = (t < n_tiles) ? a.tile_sum[t] : 0u;
        uint32_t s = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
          {
             const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, s, d);
             if (lane >= (uint32_t) d) s += o;
          }
        if (lane == 31u) warp_sum[w] = s;
        __syncthreads ();
        uint32_t before = carry_s;
        for (uint32_t k = 0; k < w; k++) before += warp_sum[k];
        const uint32_t incl = s + before;
        if (t < n_tiles) a.tile_sum[t] = incl - own;
        __syncthreads ();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads ();
     }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fused form: ONE persistent kernel.  A CTA takes tiles of kFuTile consecutive events from a ticket counter, stages the input
// columns of a window of up to kFuCap events (the tile plus what the tile's last frame needs beyond it) in shared memory and
// runs the per-event functions of mx_pileup.cuh on the window -- no per-event scratch goes to HBM.  A tile OWNS the frames
// whose first event lies in it, so every frame is processed exactly once.
//
// What the step kernels do by walking an event's whole frame run (O(frame length) per event and step) is done here with
//   * block scans over the window: the start / end of every event's frame (segmented max / min of the frame-head positions),
//     the number of drawing islands behind an event, the output row of an event;
//   * a hash table in shared memory keyed by (frame, pixel): one pass inserts every event and keeps the LAST event of a
//     pixel (atomicMax on the stored index = the pixel's representative, marxpileup.c:903-908) and marks pixels hit more than once;
//     the 3 x 3 neighbourhood of a representative is nine probes.  Only pixels hit more than once replay their events.
// Output rows (reference order: frames ascending, reverse file order inside a frame) are placed by the step path's three small
// kernels (prefix sum over the emit flags + scatter) from the rows this kernel stages at the events' own slots.  A frame that does not fit
// the window (more than kFuCap - kFuTile + 1 events is the guaranteed size) raises kPuErrFallback: the host then runs the step
// kernels above.  Same arithmetic, same order of operations: the rows are bit-identical to the step kernels' (tests).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kFuEmpty = 0xFFFFFFFFu;

// W = window capacity in events (tile = W / 2 events, W / 2 threads, 2 W hash slots: load <= 0.5): <256> keeps eight 128-thread CTAs per SM in
// flight -- the phases below are short and separated by block barriers, so it is the number of independent CTAs that hides the
// barrier waits (ncu: with one 512-thread CTA pair per SM 33 % of the stall samples sat on barriers) -- and serves frames of up to
// 129 events; <1024> serves frames of up to 513 events.
template <int W> struct FuSmem
{
   static constexpr int kFuCap = W, kFuTile = W / 2, kFuThreads = W / 2, kFuHash = 2 * W;
   uint32_t frame[kFuCap], key[kFuCap], lo[kFuCap], hi[kFuCap], pn[kFuCap], in[kFuCap], emit[kFuCap], cum[kFuCap];
   float x[kFuCap], y[kFuCap], t[kFuCap], benergy[kFuCap], pb[kFuCap], px[kFuCap], py[kFuCap], ib[kFuCap], sx[kFuCap], sy[kFuCap];
   uint32_t table[kFuHash];                 // (frame, pixel) -> index of the pixel's last event
   uint8_t multi[kFuHash];                  // the pixel was hit more than once: its events are replayed in order
   uint32_t occupied[kFuHash / 4];          // one bit per second-hash value of an inserted (frame, pixel): 8 bits per window slot
   uint32_t draws[kFuCap];                  // inclusive prefix of (flag == 2)
   uint16_t slot[kFuCap];                   // where an event's pixel sits in the table
   int16_t spha[kFuCap];
   int8_t ccd[kFuCap];
   uint8_t flag[kFuCap];
   uint32_t warp_sum[kFuThreads / 32];
   PileupArgs args;                         // the argument block with its per-event columns re-pointed at this window
   unsigned long long tile;
   uint32_t own_lo, own_hi, total;
   unsigned long long base;
};

// block-wide inclusive scan of two consecutive window entries per thread (v0 at 2 tid, v1 at 2 tid + 1) with operator OP;
// returns the scanned values in v0, v1 and leaves the block total in S.total
template <int W, class OP>
__device__ __forceinline__ void fu_scan2 (FuSmem<W> &S, uint32_t &v0, uint32_t &v1, uint32_t identity, OP op)
{
   const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
   v1 = op (v0, v1);
   uint32_t s = v1;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1)
     {
        const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, s, d);
        if (lane >= (uint32_t) d) s = op (o, s);
     }
   __syncthreads ();                        // the previous user of warp_sum is done
   if (lane == 31u) S.warp_sum[warp] = s;
   __syncthreads ();
   // every warp scans the warp totals itself, with shuffles (a loop over S.warp_sum[0 .. warp) was a chain of dependent
   // shared-memory loads, longest for the last warp: 9 % of the kernel's stall samples)
   constexpr uint32_t kWarps = (uint32_t) FuSmem<W>::kFuThreads / 32u;
   uint32_t ws = (lane < kWarps) ? S.warp_sum[lane] : identity;
#pragma unroll
   for (uint32_t d = 1; d < kWarps; d <<= 1)
     {
        const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, ws, d);
        if (lane >= d) ws = op (o, ws);
     }
   uint32_t before = __shfl_sync (0xFFFFFFFFu, ws, (warp + 31u) & 31u);
   if (warp == 0u) before = identity;
   const uint32_t excl = op (before, __shfl_up_sync (0xFFFFFFFFu, s, 1));
   const uint32_t left = (lane == 0u) ? before : excl;          // everything before this thread's pair
   v0 = op (left, v0);
   v1 = op (left, v1);
   if (tid == (uint32_t) FuSmem<W>::kFuThreads - 1u) S.total = v1;
}

// key = 1024 y + x: the pixels above and below differ by 1024, and a product with an odd constant keeps the low 10 bits of such
// keys EQUAL -- taking the slot from the low bits of one product put every column of an event cluster into two slots and made the
// probe chains as long as the cluster is tall (ncu: 26 % of the kernel's instructions in the probe loop at 3 active lanes).  Both
// hashes therefore fold the high half into the low one between two multiplications.
__device__ __forceinline__ uint32_t fu_mix (uint32_t h)
{
   h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
   return h;
}
template <int W> __device__ __forceinline__ uint32_t fu_hash (uint32_t key, uint32_t frame_id)
{
   return fu_mix (key * 0x9E3779B1u + frame_id) & (uint32_t) (FuSmem<W>::kFuHash - 1);
}
// second, independent hash into the occupancy bitmap (8 * kFuHash bits)
template <int W> __device__ __forceinline__ uint32_t fu_hash2 (uint32_t key, uint32_t frame_id)
{
   return fu_mix (key + frame_id * 0xC2B2AE35u + 0x27D4EB2Fu) & (uint32_t) (8 * FuSmem<W>::kFuHash - 1);
}
// the table slot of pixel (frame_id, key), or kFuHash if the pixel is empty
template <int W> __device__ __forceinline__ uint32_t fu_find (const FuSmem<W> &S, uint32_t key, uint32_t frame_id)
{
   constexpr int kFuHash = FuSmem<W>::kFuHash;
   // an empty pixel is recognised by one bit test in most cases (<= 1024 pixels in 16384 bits: 6 % false positives), without
   // entering the probe loop, whose trip counts differ from lane to lane
   const uint32_t b = fu_hash2<W> (key, frame_id);
   if (0u == (S.occupied[b >> 5] & (1u << (b & 31u)))) return (uint32_t) kFuHash;
   uint32_t s = fu_hash<W> (key, frame_id);
   while (true)
     {
        const uint32_t j = S.table[s];
        if (j == kFuEmpty) return (uint32_t) kFuHash;
        if ((S.key[j] == key) && (S.lo[j] == frame_id)) return s;
        s = (s + 1u) & (uint32_t) (kFuHash - 1);
     }
}
// the 3 x 3 neighbourhood of representative e by eight probes: slot [r][c] = last event on pixel (y - 1 + r, x - 1 + c) of e's frame.
// (Advancing the eight probe sequences of a thread together, one slot each per round, measured SLOWER -- 1.23 against 1.04 ms for
// the whole call: the probe state of eight sequences costs more than the divergence it removes.)
template <int W> __device__ __forceinline__ void fu_hood (const FuSmem<W> &S, uint32_t e, PuHood &h)
{
   constexpr int kFuHash = FuSmem<W>::kFuHash;
   const uint32_t key = S.key[e], fid = S.lo[e];
#pragma unroll
   for (int r = 0; r < 3; r++)
#pragma unroll
     for (int c = 0; c < 3; c++)
       {
          int32_t j = (int32_t) e;
          if ((r != 1) || (c != 1))
            {
               const uint32_t nk = key + (uint32_t) ((r - 1) * 1024 + (c - 1));        // x, y of a keyed event lie in [1, 1022]
               const uint32_t s = fu_find (S, nk, fid);
               j = (s == (uint32_t) kFuHash) ? -1 : (int32_t) S.table[s];
            }
          h.at[r][c] = j;
       }
}

template <int W>
__global__ void __launch_bounds__ (W / 2, (W <= 256) ? 8 : ((W <= 512) ? 4 : 2)) pu_fused (const __grid_constant__ PileupArgs g, unsigned long long *ticket)
{
   constexpr int kFuCap = FuSmem<W>::kFuCap, kFuTile = FuSmem<W>::kFuTile, kFuThreads = FuSmem<W>::kFuThreads, kFuHash = FuSmem<W>::kFuHash;
   extern __shared__ __align__ (16) unsigned char fu_raw[];
   FuSmem<W> &S = *reinterpret_cast<FuSmem<W> *> (fu_raw);
   const uint32_t tid = threadIdx.x;
   // the per-event functions see the window through the same argument block, its columns pointing into shared memory
   PileupArgs &a = S.args;
   if (tid == 0)
     {
        a = g;
        a.ccd = S.ccd; a.x = S.x; a.y = S.y; a.t = S.t; a.benergy = S.benergy;
        a.frame = S.frame; a.key = S.key; a.lo = S.lo; a.hi = S.hi; a.pn = S.pn; a.in = S.in; a.emit = S.emit; a.cum = S.cum;
        a.pb = S.pb; a.px = S.px; a.py = S.py; a.ib = S.ib; a.sx = S.sx; a.sy = S.sy; a.spha = S.spha; a.flag = S.flag;
     }
   const uint64_t n_tiles = (g.n + kFuTile - 1) / kFuTile;
   auto op_max = [] (uint32_t p, uint32_t q) { return max (p, q); };
   auto op_add = [] (uint32_t p, uint32_t q) { return p + q; };
   while (true)
     {
        __syncthreads ();
        if (tid == 0) S.tile = atomicAdd (ticket, 1ull);
        __syncthreads ();
        const uint64_t tile = S.tile;
        if (tile >= n_tiles) break;
        const uint64_t t0 = tile * kFuTile;
        const uint32_t cnt = (uint32_t) min ((uint64_t) kFuCap, g.n - t0);
        const uint32_t in_tile = min ((uint32_t) kFuTile, cnt);
        for (uint32_t e = tid; e < cnt; e += kFuThreads)
          {
             S.ccd[e] = g.ccd[t0 + e]; S.x[e] = g.x[t0 + e]; S.y[e] = g.y[t0 + e]; S.t[e] = g.t[t0 + e]; S.benergy[e] = g.benergy[t0 + e];
          }
        for (uint32_t k = tid; k < (uint32_t) kFuHash; k += kFuThreads) { S.table[k] = kFuEmpty; S.multi[k] = 0; }
        for (uint32_t k = tid; k < (uint32_t) kFuHash / 4u; k += kFuThreads) S.occupied[k] = 0;
        if (tid == 0) { S.own_lo = cnt; S.own_hi = cnt; a.n = cnt; }
        __syncthreads ();
        for (uint32_t e = tid; e < cnt; e += kFuThreads) pu_frames (a, e);
        __syncthreads ();
        // the frames this tile owns: from the first frame head inside the tile to the end of the frame of the tile's last event
        const uint32_t prev_frame = (t0 == 0) ? 0u : (unsigned int) ((double) g.t[t0 - 1] / g.frame_time);
        const uint32_t i0 = 2u * tid, i1 = i0 + 1u;
        auto is_head = [&] (uint32_t e) -> bool
          {
             if (e >= cnt) return false;
             return (e == 0) ? ((t0 == 0) || (S.frame[0] != prev_frame)) : (S.frame[e] != S.frame[e - 1]);
          };
        const bool h0 = is_head (i0), h1 = is_head (i1);
        if (h0 && (i0 < in_tile)) atomicMin (&S.own_lo, i0);
        if (h1 && (i1 < in_tile)) atomicMin (&S.own_lo, i1);
        if ((i0 >= in_tile) && (i0 < cnt) && (S.frame[i0] != S.frame[in_tile - 1])) atomicMin (&S.own_hi, i0);
        if ((i1 >= in_tile) && (i1 < cnt) && (S.frame[i1] != S.frame[in_tile - 1])) atomicMin (&S.own_hi, i1);
        __syncthreads ();
        uint32_t own_lo = S.own_lo, own_hi = S.own_hi;
        if (own_lo >= in_tile) { own_lo = 0; own_hi = 0; }          // no frame starts here: an earlier tile owns all of it
        else if ((own_hi == cnt) && (t0 + cnt < g.n) && (S.frame[cnt - 1] == S.frame[in_tile - 1]))
          {
             // the last owned frame runs past the window: the step kernels take over (host side)
             if (tid == 0) atomicOr (g.error, kPuErrFallback);
             own_hi = own_lo;
          }
        // frame start of every event: running maximum of the head positions (+ 1, so that 0 can be the identity)
        {
           uint32_t v0 = h0 ? i0 + 1u : 0u, v1 = h1 ? i1 + 1u : 0u;
           fu_scan2 (S, v0, v1, 0u, op_max);
           S.lo[i0] = (v0 > 0u) ? v0 - 1u : 0u;
           S.lo[i1] = (v1 > 0u) ? v1 - 1u : 0u;
        }
        __syncthreads ();
        // frame end: the first head position above the event.  Scan the window backwards: scan index i stands for event
        // kFuCap - 1 - i, a head at event q contributes kFuCap - q (0: no head), the running maximum is the NEAREST head at or
        // above the event
        {
           const uint32_t q0 = (uint32_t) kFuCap - 1u - i0, q1 = (uint32_t) kFuCap - 1u - i1;
           uint32_t v0 = is_head (q0) ? (uint32_t) kFuCap - q0 : 0u, v1 = is_head (q1) ? (uint32_t) kFuCap - q1 : 0u;
           fu_scan2 (S, v0, v1, 0u, op_max);
           S.hi[q0] = v0; S.hi[q1] = v1;                                // "first head at a position >= the event", encoded
        }
        __syncthreads ();
        // the value at e + 1 is "first head at a position > e": decode it (no head: the window end)
        for (uint32_t e = tid; e < cnt; e += kFuThreads)
          {
             const uint32_t v = (e + 1u < (uint32_t) kFuCap) ? S.hi[e + 1u] : 0u;
             S.cum[e] = (v > 0u) ? (uint32_t) kFuCap - v : cnt;
          }
        __syncthreads ();
        for (uint32_t e = tid; e < cnt; e += kFuThreads) S.hi[e] = min (S.cum[e], max (own_hi, e + 1u));
        if (tid == 0) a.n = own_hi;
        __syncthreads ();
        // one pass over the owned events: (frame, pixel) -> last event of the pixel, and the pixel's event count
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             const uint32_t key = S.key[e];
             if (key == kPuNoKey) continue;
             const uint32_t fid = S.lo[e];
             uint32_t s = fu_hash<W> (key, fid);
             while (true)
               {
                  const uint32_t old = atomicCAS (&S.table[s], kFuEmpty, e);
                  if (old == kFuEmpty)
                    {
                       const uint32_t b = fu_hash2<W> (key, fid);
                       atomicOr (&S.occupied[b >> 5], 1u << (b & 31u));
                       break;
                    }
                  // the slot's pixel is identified by any of its events: key and frame start
                  if ((S.key[old] == key) && (S.lo[old] == fid)) { atomicMax (&S.table[s], e); S.multi[s] = 1; break; }
                  s = (s + 1u) & (uint32_t) (kFuHash - 1);
               }
             S.slot[e] = (uint16_t) s;
          }
        __syncthreads ();
        // pixel state at the representatives (store_event): a pixel hit once needs no walk
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             if ((S.key[e] == kPuNoKey) || (S.table[S.slot[e]] != e)) continue;
             pu_pixel_state (a, e, S.multi[S.slot[e]] ? S.lo[e] : e);
          }
        __syncthreads ();
        // neighbourhoods by nine probes, island sums (collect_charge)
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             if (S.pn[e] == 0) continue;
             PuHood h;
             fu_hood (S, e, h);
             pu_island_from (a, e, h);
          }
        __syncthreads ();
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             if (S.pn[e] == 0) continue;
             PuHood h;
             fu_hood (S, e, h);
             pu_detect_from (a, e, h);
          }
        __syncthreads ();
        // draw index of an island = number of drawing islands of its frame behind it in file order
        {
           uint32_t v0 = ((i0 >= own_lo) && (i0 < own_hi) && (S.flag[i0] == 2)) ? 1u : 0u;
           uint32_t v1 = ((i1 >= own_lo) && (i1 < own_hi) && (S.flag[i1] == 2)) ? 1u : 0u;
           fu_scan2 (S, v0, v1, 0u, op_add);
           S.draws[i0] = v0; S.draws[i1] = v1;
        }
        __syncthreads ();
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             if (S.flag[e] == 0) continue;
             PuHood h;
             if (S.flag[e] == 2) fu_hood (S, e, h);
             else
               {
#pragma unroll
                  for (int q = 0; q < 9; q++) h.at[q / 3][q % 3] = -1;
               }
             pu_emit_from (a, e, S.draws[S.hi[e] - 1u] - S.draws[e], h);
          }
        __syncthreads ();
        // hand the rows to the row-placement kernels (pu_scan_tiles, pu_scan_totals, pu_step<5>): every owned event's emit flag and, for
        // the emitted ones, the staged row at the event's own slot of the per-event scratch -- frame bounds as GLOBAL indices.  (An
        // ordered decoupled look-back inside this kernel made every tile wait for the slowest tile in flight: the tiles of a wave
        // finished together, and a third of the kernel's time -- later, with the whole CTA looking back, 45 % of its instructions --
        // went into that wait.  The placement needs one 4-byte flag per event and ~30 bytes per row instead.)
        for (uint32_t e = own_lo + tid; e < own_hi; e += kFuThreads)
          {
             const uint64_t ge = t0 + e;
             const uint32_t em = S.emit[e];
             g.emit[ge] = em;
             if (em == 0) continue;
             g.lo[ge] = (uint32_t) (t0 + S.lo[e]); g.hi[ge] = (uint32_t) (t0 + S.hi[e]);
             g.frame[ge] = S.frame[e]; g.in[ge] = S.in[e]; g.ib[ge] = S.ib[e];
             g.sx[ge] = S.sx[e]; g.sy[ge] = S.sy[e]; g.spha[ge] = S.spha[e];
          }
     }
}

// scratch: the ticket counter (zeroed here)
size_t pileup_fused_scratch_bytes (uint64_t) { return 2 * sizeof (unsigned long long); }
template <int W>
static void launch_pileup_fused_w (const PileupArgs &a, void *scratch, int num_sms, cudaStream_t s)
{
   static int per_sm = 0;
   if (per_sm == 0)
     {
        cudaFuncSetAttribute (pu_fused<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof (FuSmem<W>));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, pu_fused<W>, W / 2, sizeof (FuSmem<W>));
        if (per_sm < 1) per_sm = 1;
     }
   cudaMemsetAsync (scratch, 0, pileup_fused_scratch_bytes (a.n), s);
   const uint64_t tiles = (a.n + W / 2 - 1) / (W / 2);
   const unsigned int grid = (unsigned int) min (tiles, (uint64_t) num_sms * per_sm);
   unsigned long long *w = (unsigned long long *) scratch;
   pu_fused<W><<<grid, W / 2, sizeof (FuSmem<W>), s>>> (a, w);
}
// window: 256 (frames of up to 129 events), 512 (up to 257) or 1024 (up to 513)
void launch_pileup_fused (const PileupArgs &a, void *scratch, int window, int num_sms, cudaStream_t s, int *n_launches)
{
   *n_launches = 0;
   if (a.n == 0) return;
   if (window <= 256) launch_pileup_fused_w<256> (a, scratch, num_sms, s);
   else if (window <= 512) launch_pileup_fused_w<512> (a, scratch, num_sms, s);
   else launch_pileup_fused_w<1024> (a, scratch, num_sms, s);
   // row placement: prefix sum over the emit flags, then the rows go to their places in the reference's order
   const unsigned int tiles = (unsigned int) ((a.n + kPuTile - 1) / kPuTile);
   const unsigned int grid = min (tiles, (unsigned int) num_sms * 8u);
   pu_scan_tiles<<<grid, kPuTile, 0, s>>> (a);
   pu_scan_totals<<<1, 1024, 0, s>>> (a);
   pu_step<5><<<grid, kPuTile, 0, s>>> (a);
   *n_launches = 4;
}

// input columns of the pile-up model from the live event list in HBM: what marx_write_photons would have put into the column files
// marxpileup reads (marxio.c:217-290: chip id, chip pixels, (float) (arrival time + total_time), PI energy, the six dither values)
__global__ void __launch_bounds__ (256) pu_gather (PhotonSoA in, const unsigned long long *n_ptr, uint64_t max_n, const double *dev_start_time,
                                                  double total_time, int8_t *ccd, float *x, float *y, float *t, float *b,
                                                  float *d0, float *d1, float *d2, float *d3, float *d4, float *d5)
{
   const uint64_t n = min ((uint64_t) *n_ptr, max_n);
   const double start = *dev_start_time;
   for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
     {
        ccd[i] = in.ccd[i]; x[i] = in.chipx[i]; y[i] = in.chipy[i];
        t[i] = (float) ((in.time[i] - start) + total_time);
        b[i] = in.pi[i];
        d0[i] = in.dra[i]; d1[i] = in.ddec[i]; d2[i] = in.droll[i];
        d3[i] = in.ddy ? in.ddy[i] : 0.0f; d4[i] = in.ddz ? in.ddz[i] : 0.0f; d5[i] = in.ddth ? in.ddth[i] : 0.0f;
     }
}
void launch_pileup_gather (const PhotonSoA &in, const unsigned long long *n_ptr, uint64_t max_n, const double *dev_start_time, double total_time,
                           int8_t *ccd, float *const cols[10], cudaStream_t s)
{
   if (max_n == 0) return;
   const unsigned int grid = (unsigned int) min ((uint64_t) 148 * 8, (max_n + 255) / 256);
   pu_gather<<<grid, 256, 0, s>>> (in, n_ptr, max_n, dev_start_time, total_time, ccd, cols[0], cols[1], cols[2], cols[3],
                                   cols[4], cols[5], cols[6], cols[7], cols[8], cols[9]);
}

void launch_pileup (const PileupArgs &a, int num_sms, cudaStream_t s, int *n_launches)
{
   *n_launches = 0;
   if (a.n == 0) return;
   const unsigned int tiles = (unsigned int) ((a.n + kPuTile - 1) / kPuTile);
   const unsigned int grid = min (tiles, (unsigned int) num_sms * 8u);
   pu_step<0><<<grid, kPuTile, 0, s>>> (a);
   pu_step<1><<<grid, kPuTile, 0, s>>> (a);
   pu_step<2><<<grid, kPuTile, 0, s>>> (a);
   pu_step<3><<<grid, kPuTile, 0, s>>> (a);
   pu_step<4><<<grid, kPuTile, 0, s>>> (a);
   pu_scan_tiles<<<grid, kPuTile, 0, s>>> (a);
   pu_scan_totals<<<1, 1024, 0, s>>> (a);
   pu_step<5><<<grid, kPuTile, 0, s>>> (a);
   *n_launches = 8;
}

}  // namespace mx
