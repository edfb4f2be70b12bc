// aspsol_kernels.cu -- sm_100a kernel of the aspect-solution table (mx_aspsol.cuh; marxasp).
//
//   asp_rows    one table row per thread (rows are independent: 13 transcendental calls, no input but the row number).
//               Output (a) eight unit-stride f64 columns [8][n] and / or (b) the FITS binary-table image of the rows exactly as
//               jdfits would write them (marxasp.c:1012-1019: 76 big-endian bytes per row), transposed through shared memory so
//               that the 19 words of 256 rows leave the CTA as one contiguous, coalesced store.
//   The kernel is compute bound by construction (76 B written per ~1.2e3 FP64 operations); it is sized like the stage kernels
//   (grid = SMs x resident CTAs, grid-stride over 256-row tiles).
#define MX_MATH 0      // the aspect-solution rows keep libdevice sin / cos (tests/test_gpu_marxasp.py pins them at 3e-13 deg)
#include <cuda_runtime.h>
#include "mx_aspsol.cuh"

namespace mx {

constexpr int kAspTile = 256;

__device__ __forceinline__ uint32_t bswap32 (uint32_t v) { return __byte_perm (v, 0, 0x0123); }

__global__ void __launch_bounds__ (kAspTile) asp_rows (const __grid_constant__ AspsolDev D, uint64_t first_row, uint64_t n,
                                                      double *__restrict__ cols, uint32_t *__restrict__ fits_rows)
{
   __shared__ uint32_t img[kAspTile * kAspsolRowWords];
   const uint64_t n_tiles = (n + kAspTile - 1) / kAspTile;
   for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
     {
        const uint64_t k = tile * kAspTile + threadIdx.x;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (k < n)
          {
             aspsol_row (D, first_row + k, v);
             if (cols != nullptr)
               {
#pragma unroll
                  for (int c = 0; c < 8; c++) cols[(uint64_t) c * n + k] = v[c];
               }
          }
        if (fits_rows != nullptr)
          {
             // row image: time, ra, dec, roll | dy, dz, dtheta (float32 zeros) | q_att[4], every value big endian
             uint32_t *r = img + threadIdx.x * kAspsolRowWords;
#pragma unroll
             for (int c = 0; c < 8; c++)
               {
                  const unsigned long long b = (unsigned long long) __double_as_longlong (v[c]);
                  const int w = (c < 4) ? 2 * c : 2 * c + 3;
                  r[w] = bswap32 ((uint32_t) (b >> 32)); r[w + 1] = bswap32 ((uint32_t) b);
               }
             r[8] = 0u; r[9] = 0u; r[10] = 0u;
             __syncthreads ();
             const uint64_t base = tile * kAspTile * kAspsolRowWords;
             const uint64_t words = min ((uint64_t) kAspTile, n - tile * kAspTile) * kAspsolRowWords;
             for (uint32_t w = threadIdx.x; w < words; w += kAspTile) fits_rows[base + w] = img[w];
             __syncthreads ();
          }
     }
}

void launch_aspsol_rows (const AspsolDev &D, uint64_t first_row, uint64_t n, double *cols, uint32_t *fits_rows, int num_sms, cudaStream_t s)
{
   if (n == 0) return;
   int per_sm = 1;
   cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, asp_rows, kAspTile, 0);
   const uint64_t n_tiles = (n + kAspTile - 1) / kAspTile;
   const unsigned int grid = (unsigned int) min ((uint64_t) num_sms * (uint64_t) (per_sm < 1 ? 1 : per_sm), n_tiles);
   asp_rows<<<grid, kAspTile, 0, s>>> (D, first_row, n, cols, fits_rows);
}

}  // namespace mx
