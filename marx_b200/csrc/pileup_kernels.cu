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
