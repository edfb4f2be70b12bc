// level1_kernels.cu -- sm_100a kernels of the Level-1 event transforms (mx_level1.cuh).
//
//   l1_frame_heads    per 256-event tile: exposure number of every event, "first event of its exposure frame" flags,
//                     inclusive max-scan of the flagged positions -> for every event the position of the latest frame
//                     head inside its tile (0: none yet), and the tile's last head
//   l1_scan_tiles     one CTA: running maximum over the tile aggregates (exclusive, in place)
//   l1_transform      one event per thread: resolve the event's aspect (its frame's first event, possibly in an earlier
//                     tile or -- through Level1State -- in an earlier batch), run level1_event, store 20 columns
//   l1_update_state   one thread: carry last exposure number / held aspect / row count into the next batch
//
// The reference walks the rows sequentially because "take the aspect over only when EXPNO changed" (marx2fits.c:3567-3580)
// is a running state; on the device that state is the segmented broadcast above.  HRC data and --pixadj=exact take every
// row's own aspect and skip the two scan kernels.  All loads and stores are unit-stride over the SoA columns.
#define MX_MATH 0      // Level-1 keeps libdevice sin / cos: its sky-coordinate checks are pinned at the libdevice-vs-glibc level (tests/test_gpu_level1.py)
#include <cuda_runtime.h>
#include "mx_level1.cuh"
#include "mx_kernels.cuh"

namespace mx {

constexpr int kL1Tile = 256;

__global__ void __launch_bounds__ (kL1Tile) l1_frame_heads (const __grid_constant__ Level1Args a)
{
   __shared__ uint32_t warp_max[kL1Tile / 32];
   const unsigned long long n = min ((unsigned long long) *a.n_ptr, (unsigned long long) a.max_n);
   const double start = *a.dev_start_time;
   const long long carried = a.state->last_expno;
   const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
   for (unsigned long long tile = blockIdx.x; tile * kL1Tile < n; tile += gridDim.x)
     {
        const unsigned long long i = tile * kL1Tile + threadIdx.x;
        uint32_t h = 0;
        if (i < n)
          {
             const long long e = l1_expno (l1_file_time (a.in.time[i], start, a.total_time), a.L.time_del);
             const long long prev = (i == 0) ? carried : l1_expno (l1_file_time (a.in.time[i - 1], start, a.total_time), a.L.time_del);
             if (e != prev) h = (uint32_t) i + 1u;
          }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
          {
             const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, h, d);
             if (lane >= (uint32_t) d) h = max (h, o);
          }
        if (lane == 31u) warp_max[w] = h;
        __syncthreads ();
        uint32_t before = 0;
        for (uint32_t k = 0; k < w; k++) before = max (before, warp_max[k]);
        h = max (h, before);
        if (i < n) a.head[i] = h;
        if (threadIdx.x == kL1Tile - 1) a.tile_head[tile] = h;
        __syncthreads ();
     }
}

// exclusive running maximum over the tile aggregates, in place (n_tiles <= a few 10^5: one CTA, sequential over chunks)
__global__ void __launch_bounds__ (1024) l1_scan_tiles (const __grid_constant__ Level1Args a)
{
   __shared__ uint32_t warp_max[32];
   __shared__ uint32_t carry_s;
   const unsigned long long n = min ((unsigned long long) *a.n_ptr, (unsigned long long) a.max_n);
   const unsigned long long n_tiles = (n + kL1Tile - 1) / kL1Tile;
   const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
   if (threadIdx.x == 0) carry_s = 0;
   __syncthreads ();
   for (unsigned long long base = 0; base < n_tiles; base += 1024)
     {
        const unsigned long long t = base + threadIdx.x;
        const uint32_t own = (t < n_tiles) ? a.tile_head[t] : 0u;
        uint32_t h = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
          {
             const uint32_t o = __shfl_up_sync (0xFFFFFFFFu, h, d);
             if (lane >= (uint32_t) d) h = max (h, o);
          }
        if (lane == 31u) warp_max[w] = h;
        __syncthreads ();
        uint32_t before = carry_s;
        for (uint32_t k = 0; k < w; k++) before = max (before, warp_max[k]);
        const uint32_t incl = max (h, before);
        // exclusive value: everything before this tile
        uint32_t excl = __shfl_up_sync (0xFFFFFFFFu, h, 1);
        excl = (lane == 0u) ? before : max (excl, before);
        if (t < n_tiles) a.tile_head[t] = excl;
        __syncthreads ();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads ();
     }
}

template <bool FRAMES>
__global__ void __launch_bounds__ (256) l1_transform (const __grid_constant__ Level1Args a)
{
   const unsigned long long n = min ((unsigned long long) *a.n_ptr, (unsigned long long) a.max_n);
   const double start = *a.dev_start_time;
   const Level1State st = *a.state;
   const bool acis = l1_is_acis (a.L);
   for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x)
     {
        Level1In in;
        in.tfile = l1_file_time (a.in.time[i], start, a.total_time);
        in.xpixel = a.in.chipx[i]; in.ypixel = a.in.chipy[i];
        in.benergy = a.in.pi[i];
        in.upix = a.in.upix[i]; in.vpix = a.in.vpix[i];
        in.pha = a.in.pha[i];
        in.ccd = a.in.ccd[i];
        in.row = st.rows + i;
        in.expno = 0;
        if (acis) in.expno = (a.L.time_del > 0.0) ? l1_expno (in.tfile, a.L.time_del) : st.last_expno + (long long) i;   // compute_expno :3748-3755
        // aspect: the row's own, or that of the first event of its exposure frame
        unsigned long long src = i;
        bool from_state = false;
        if (FRAMES)
          {
             const uint32_t h = max (a.head[i], a.tile_head[i / kL1Tile]);
             if (h == 0u) from_state = true; else src = (unsigned long long) h - 1ull;
          }
        if (from_state)
          {
#pragma unroll
             for (int k = 0; k < 6; k++) in.dither[k] = st.dither[k];
          }
        else if (a.L.used_dither)
          {
             in.dither[0] = a.in.dra[src]; in.dither[1] = a.in.ddec[src]; in.dither[2] = a.in.droll[src];
             in.dither[3] = a.in.ddy ? a.in.ddy[src] : 0.0f;
             in.dither[4] = a.in.ddz ? a.in.ddz[src] : 0.0f;
             in.dither[5] = a.in.ddth ? a.in.ddth[src] : 0.0f;
          }
        else
          {
#pragma unroll
             for (int k = 0; k < 6; k++) in.dither[k] = 0.0f;
          }
        Level1Out o;
        level1_event (a.L, a.seed, in, o);
        if (o.error) atomicOr (a.error_flag, (unsigned int) o.error);
        a.out.time[i] = o.time; a.out.detx[i] = o.detx; a.out.dety[i] = o.dety; a.out.x[i] = o.x; a.out.y[i] = o.y;
        a.out.expno[i] = o.expno; a.out.tdetx[i] = o.tdetx; a.out.tdety[i] = o.tdety; a.out.pha[i] = o.pha;
        a.out.hrc_u[i] = o.hrc_u; a.out.hrc_v[i] = o.hrc_v;
        a.out.energy[i] = o.energy;
        a.out.ccd_id[i] = o.ccd_id; a.out.node_id[i] = o.node_id; a.out.chipx[i] = o.chipx; a.out.chipy[i] = o.chipy;
        a.out.pi[i] = o.pi; a.out.fltgrade[i] = o.fltgrade; a.out.grade[i] = o.grade; a.out.status[i] = o.status;
        a.out.keep[i] = o.keep;
        // the last row leaves the aspect it computed with for the next batch (read_dither_value's held values)
        if (i + 1 == n)
          {
#pragma unroll
             for (int k = 0; k < 6; k++) a.next_dither[k] = in.dither[k];
             a.next_expno[0] = acis ? in.expno : st.last_expno;
          }
     }
}

__global__ void l1_update_state (const __grid_constant__ Level1Args a)
{
   const unsigned long long n = min ((unsigned long long) *a.n_ptr, (unsigned long long) a.max_n);
   if (n == 0) return;
   Level1State *s = a.state;
   const bool counter_mode = l1_is_acis (a.L) && !(a.L.time_del > 0.0);       // compute_expno :3748-3753: last_expno++ per row
   s->last_expno = counter_mode ? s->last_expno + (long long) n : a.next_expno[0];
   for (int k = 0; k < 6; k++) s->dither[k] = a.next_dither[k];
   s->rows += n;
}

void launch_level1 (const Level1Args &a, int num_sms, cudaStream_t s, int *n_launches)
{
   *n_launches = 0;
   if (a.max_n == 0) return;
   const unsigned int tiles = (unsigned int) ((a.max_n + kL1Tile - 1) / kL1Tile);
   const unsigned int grid = min (tiles, (unsigned int) num_sms * 8u);
   if (l1_frames_share_aspect (a.L))
     {
        l1_frame_heads<<<grid, kL1Tile, 0, s>>> (a);
        l1_scan_tiles<<<1, 1024, 0, s>>> (a);
        l1_transform<true><<<grid, 256, 0, s>>> (a);
        *n_launches += 3;
     }
   else
     {
        l1_transform<false><<<grid, 256, 0, s>>> (a);
        *n_launches += 1;
     }
   l1_update_state<<<1, 1, 0, s>>> (a);
   *n_launches += 1;
}

}  // namespace mx
