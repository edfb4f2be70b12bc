// mx_kernels.cuh -- launch-side view of the stage kernels (kernels.cu) used by the C ABI (marxb200.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "mx_tables.h"
#include "mx_level1.cuh"

namespace mx {

constexpr int kTile = 256;               // K0: rays per tile == threads per CTA (also the unit of the canonical time sum)
#ifndef MX_STAGE_THREADS
#define MX_STAGE_THREADS 256
#endif
constexpr int kStageThreads = MX_STAGE_THREADS;   // K1..K3: threads per persistent CTA (warps are autonomous)
constexpr int kWarpTile = 32;            // rays per warp tile
constexpr int kQueueCap = 64;            // entries of a warp's staging queue
constexpr int kSuperTile = 256;          // tiles per super-tile of the canonical arrival-time sum

// Structure-of-arrays photon buffer in HBM (replaces Marx_Photon_Attr_Type[], marx.h:51-100).
struct PhotonSoA
{
   double *energy;
   double *x0, *x1, *x2;
   double *p0, *p1, *p2;
   double *time;                         // absolute: pt->start_time + arrival_time
   double *aux;                          // scratch between the HRMA sub-kernels (cached Box-Muller spare)
   uint64_t *ray;                        // global ray index (RNG counter; low 32 bits = tag)
   uint32_t *slot;                       // index of the ray inside its batch: the key that restores arrival order
   uint32_t *flags;
   float *dra, *ddec, *droll;            // Marx_Dither_Type ra/dec/roll
   float *ddy, *ddz, *ddth;              // Marx_Dither_Type dy/dz/dtheta: detector dither (0 unless DitherModel=FILE)
   float *chipx, *chipy, *pi;
   float *upix, *vpix;                   // HRC u/v pixels
   uint32_t *sorders;                    // LETG support-grating orders, one signed byte per pass
   int16_t *pha;
   uint8_t *shell;
   int8_t *order, *ccd, *region;
};

// Per-ray quantities that no stage changes after the source has set them.  They are written ONCE, at the ray's batch
// slot, and read through the `slot` key of a list entry; the compacting stage kernels therefore move only what they
// produce (x, p, flags, ids) plus the 4-byte slot.  (The first version carried these 36 bytes from list to list in every
// stage: 7 dependent gathers + 7 stores per survivor per stage, 13 % of k1_hrma<1>'s stall samples.)  The arrival-order
// restoration materialises them into the list's own columns, which is what the host boundary reads.
// energy and ray id share ONE 16-byte record per slot: every stage kernel needs both (the ray id keys its draws), and a gather
// through the slot key costs a 32-byte DRAM sector per touched array once the list is sparse -- one sector instead of two
// (k2_select: 102 -> ~70 B of DRAM traffic per input ray).
struct RayConst
{
   double2 *er;                          // x = energy, y = the bit pattern of the 64-bit global ray index
   double *time;
   float *dra, *ddec, *droll;
   float *ddy, *ddz, *ddth;
};
#if defined(__CUDACC__)
__device__ __forceinline__ void rc_load (const RayConst &rc, uint32_t slot, double &energy, uint64_t &ray)
{
   const double2 v = rc.er[slot];
   energy = v.x; ray = (uint64_t) __double_as_longlong (v.y);
}
__device__ __forceinline__ void rc_store (const RayConst &rc, uint64_t slot, double energy, uint64_t ray)
{
   rc.er[slot] = make_double2 (energy, __longlong_as_double ((long long) ray));
}
#endif

// Blob staged into shared memory by K1 with one TMA bulk copy.
struct alignas (16) K1Blob
{
   HrmaDev H;
   uint32_t off_opt_e, off_opt_b, off_opt_d, off_corr_e, off_corr_f, total_bytes, pad0, pad1;
   // contiguous copies of the WFOLD e_alpha search keys (column 0 of the header rows), one block of wkeys_stride bytes per
   // shell: [.. off_wkeys_p ..] directly behind the correction tables (staged by k1_hrma<1> together with them),
   // [.. off_wkeys_h ..] behind that (staged by k1_hrma<2> as a second segment).  0 bytes when WFOLD is off.
   uint32_t off_wkeys_p, off_wkeys_h, wkeys_stride, wkeys_bytes;
};
struct alignas (16) K2Blob
{
   GratingDev G;
   uint32_t off_sectors[kNumShells];     // byte offsets of each shell's [6][num_sectors] doubles
   uint32_t total_bytes, pad0, pad1, pad2;
};
struct alignas (16) K3Blob
{
   AcisDev A;
   uint32_t total_bytes, pad0, pad1, pad2;
};
struct alignas (16) K3HrcBlob
{
   HrcDev D;
   uint32_t total_bytes, pad0, pad1, pad2;
};

static_assert (sizeof (K1Blob) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

struct StageArgs
{
   PhotonSoA in, out;
   RayConst rc;                          // indexed by in.slot[i]
   const unsigned long long *n_in;       // device: number of input slots
   unsigned long long *n_out;            // device: number of output photons; compact: zeroed before launch, grown by atomics
   unsigned long long *ticket;           // device: chunk ticket counter (zeroed before launch)
   uint64_t seed;
   int chunk_tiles;                      // 32-ray warp tiles handed out per ticket
   int compact;                          // 1: order-preserving compaction into `out`; 0: in place, dead rays kept
   double source_distance;
   const void *blob;                     // K1Blob / K2Blob / K3Blob in global memory
   uint32_t blob_bytes;                  // bytes staged into shared memory from the start of the blob
   uint32_t seg2_off, seg2_bytes;        // optional second staged segment (byte range of the blob), placed behind the first
   int det_dither;                       // detector stage: read the per-ray detector dither (ASPSOL model or uploaded photons)
   double ff[5];                         // MirrorType=FLATFIELD: min_y, min_z, max_y, max_z, x_pos (k1_flatfield)
};

struct SourceArgs
{
   PhotonSoA out;
   RayConst rc;                          // written at the batch slot of every generated (k0_source) / surviving (k01) ray
   uint64_t first_ray, n;
   uint64_t seed;
   SourceDev S;
   DitherDev D;
   double time_base;                     // absolute time of the batch start (used when use_dev_base == 0)
   int use_dev_base;                     // 1: continue from dev_times[1] (end of the previous batch), no host round trip
   double *dev_times;                    // device: [0] = start time of the current batch, [1] = running end time
   double *tile_sums;                    // [n_tiles]
   double *tile_base;                    // [n_tiles]
   double *supertile_sums;               // [n_supertiles]
   unsigned long long *n_out;            // device: count[0] = n
};

void launch_time_sums (const SourceArgs &a, cudaStream_t s);
void launch_time_scan (const SourceArgs &a, cudaStream_t s, bool with_super = true);   // false: the super-tile sums exist already
void launch_time_super (const SourceArgs &a, cudaStream_t s);                 // multi-GPU scan, first half: k0_time_super
// second half, behind the all-gather of the super-tile sums: k0_time_bases_sharded + k0_time_tiles
void launch_time_bases_sharded (const SourceArgs &a, const double *all_sums, int rank, int world, uint32_t ns_blk, cudaStream_t s);
void launch_source (const SourceArgs &a, cudaStream_t s);
void launch_exposure_truncate (const PhotonSoA &buf, unsigned long long *n, double *dev_times, double limit, int inclusive, cudaStream_t s);
void launch_hrma (const StageArgs &a, int phase, int grid, cudaStream_t s);
void launch_flatfield (const StageArgs &a, int num_sms, cudaStream_t s);       // MirrorType=FLATFIELD: the whole mirror stage
int fused_source_grid (int num_sms);
void launch_source_hrma (const SourceArgs &a, const StageArgs &st, int grid, cudaStream_t s);   // k0_source + k1_hrma<0> in one kernel   // phase 0,1,2 = k1a,k1b,k1c
void launch_grating (const StageArgs &a, int grid, cudaStream_t s, int phase = 0);   // phase 0: one kernel; 1: k2_select, 2: k2_grating<1>
void launch_acis (const StageArgs &a, int grid, cudaStream_t s, int phase = 0);   // phase 0: one kernel; 1, 2: the two halves
void launch_hrc (const StageArgs &a, int grid, cudaStream_t s);
int stage_grid_size (int stage, int num_sms, uint32_t blob_bytes, uint32_t seg2_bytes = 0);
uint32_t stage_smem_bytes (int stage, uint32_t blob_bytes, uint32_t seg2_bytes = 0);

// arrival-order restoration of an unordered live list (DESIGN.md "compaction"): bitmap over batch slots ->
// popcount prefix -> rank (inverse permutation) -> gather
struct OrderArgs
{
   PhotonSoA in, out;
   RayConst rc;
   const unsigned long long *n_live;
   uint64_t n_slots;                     // slots of the batch (= generated rays)
   uint32_t *bitmap;                     // [n_slots/32 + 1], zeroed before launch
   uint32_t *word_prefix;                // [n_words]
   uint32_t *block_prefix;               // [n_words/1024 + 1]
   uint32_t *perm;                       // [n_slots]: position in the unordered list of the photon with arrival rank j
};
struct PackArgs;
void launch_restore_order (const OrderArgs &a, int num_sms, cudaStream_t s, int *n_launches, const PackArgs *pack = nullptr);

// host boundary helpers (AoS <-> SoA); `aos` is a device buffer of 136-byte records
void launch_soa_to_aos (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, void *aos,
                        const double *dev_start_time, cudaStream_t s);
void launch_aos_to_soa (const void *aos, const uint64_t *ray_ids, uint64_t n, const PhotonSoA &out, const RayConst &rc,
                        double start_time, cudaStream_t s);

// bulk egress in the reference's column-file format (marxio.c:292-322): which packed columns to produce
enum EgressKind
{
   EGRESS_PI = 0, EGRESS_ENERGY, EGRESS_TIME, EGRESS_TAG, EGRESS_XPOS, EGRESS_YPOS, EGRESS_ZPOS, EGRESS_XCOS, EGRESS_YCOS,
   EGRESS_ZCOS, EGRESS_PHA, EGRESS_CCD, EGRESS_CHIPX, EGRESS_CHIPY, EGRESS_HRC_U, EGRESS_HRC_V, EGRESS_MIRROR, EGRESS_REGION,
   EGRESS_ORDER, EGRESS_ORDER1, EGRESS_ORDER2, EGRESS_ORDER3, EGRESS_ORDER4, EGRESS_SKY_RA, EGRESS_SKY_DEC, EGRESS_SKY_ROLL,
   EGRESS_DET_DY, EGRESS_DET_DZ, EGRESS_DET_THETA, EGRESS_NUM_KINDS
};
// Event tallies (marxb200_tally_*): exact integer histograms of the live list, 1 or 2 axes
enum TallyColumn
{
   TALLY_ENERGY = 0, TALLY_TIME, TALLY_PHA, TALLY_PI, TALLY_ORDER, TALLY_CCD, TALLY_SHELL, TALLY_CHIPX, TALLY_CHIPY,
   TALLY_YPOS, TALLY_ZPOS, TALLY_NUM_COLUMNS
};
struct TallyAxis { int column; uint32_t nbins; double lo, scale; };      // bin = floor ((v - lo) * scale), dropped unless 0 <= bin < nbins
struct TallyPlan { int naxes; TallyAxis ax[2]; };
void launch_tally (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, const TallyPlan &plan,
                   unsigned long long *bins, int num_sms, cudaStream_t s);

constexpr int kMaxEgressCols = 32;
struct EgressPlan
{
   int num_cols;
   int kind[kMaxEgressCols];
   uint64_t offset[kMaxEgressCols];      // byte offset of each packed column in the staging buffer (4-byte aligned)
};
// file images emitted by the order restoration itself (order_gather<true>): the columns of `plan` for rows [0, max_rows) into dst
struct PackArgs
{
   EgressPlan plan;
   unsigned char *dst;
   const double *dev_start_time;         // null: TIME = absolute time + total_time
   double total_time;
   uint64_t max_rows;
};
void launch_fp64_peak (double *sink, int grid, int iters, cudaStream_t s);   // 64 DFMA per thread per iteration
void launch_egress_pack (const PhotonSoA &in, const unsigned long long *n, uint64_t max_n, const EgressPlan &plan, void *dst,
                         const double *dev_start_time, double total_time, cudaStream_t s);

// Level-1 event transforms (level1_kernels.cu)
struct Level1Args
{
   PhotonSoA in;                         // the live list in arrival order
   const unsigned long long *n_ptr;      // device: number of events
   uint64_t max_n;
   const double *dev_start_time;         // device: start time of the batch (time column is absolute)
   double total_time;                    // TIME of the file = (float) ((time - start) + total_time)
   uint64_t seed;
   Level1Dev L;
   Level1State *state;                   // device
   Level1Cols out;
   uint32_t *head, *tile_head;           // [max_n], [max_n / 256 + 1]: frame-head scan scratch
   float *next_dither;                   // device [6]
   long long *next_expno;                // device [1]
   unsigned int *error_flag;             // device
};
void launch_level1 (const Level1Args &a, int num_sms, cudaStream_t s, int *n_launches);

}  // namespace mx
