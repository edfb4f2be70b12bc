// mx_pileup.cuh -- ACIS pile-up model applied to the event list of a simulation: the frame loop of marxpileup
// (marx/src/marxpileup.c:main :1121-1213 -> process_frame :890-922).
//
// The reference keeps one 1024 x 1024 pixel map per CCD, stores the events of one exposure frame into it in list order
// (store_event :754-812), forms 3 x 3 island sums (collect_charge :814-845), keeps the local maxima (event_detect :676-752),
// lets islands of >= 2 photons survive grade migration with probability alpha^(n-1) (will_grade_migrate :668-674) and writes
// one event per surviving island (write_event :622-666), then clears what the frame touched.  Nothing crosses a frame, and a
// frame is a contiguous run of the (arrival-ordered) event list: here every step is EVENT-parallel, each event resolving its
// pixel and its 3 x 3 neighbourhood by walking its own frame's run of the list (tens of events, L1/L2-resident), so there is no
// pixel map at all.  "Representative" of a pixel = the last event in file order that fell on it: the reference's list holds
// the frame in REVERSE file order (main :1148-1153 prepends) and drops the later duplicates (:903-908).
//
//   pu_frames    frame number, pixel key, error checks (read_input_event :573-620, allocate_input_event :559-568)
//   pu_store     frame bounds; per-pixel photon count, summed energy and the order-dependent energy-weighted mean position
//   pu_island    3 x 3 sums of energy and photon count around every occupied pixel
//   pu_detect    the two local-maximum tests (asymmetric >= / > tie rules); which islands need a draw
//   pu_emit      draw k of frame F = lane k&3 of Philox4x32-10 (key = seed, counter = (F, 0, k>>2, 5)), k = number of islands
//                that drew before this one in list order; centroid; PHA of the summed energy (marx_map_energy_to_acis_pha,
//                acis_fef.c:1087-1096); stages the output row at the event's slot
//   (prefix sum over the emit flags, pileup_kernels.cu)
//   pu_scatter   output row = rows of earlier frames + rows of this frame behind it in file order (reverse order in a frame)
//
// All functions are MX_HD: the device kernels (pileup_kernels.cu) call them one event per thread; tools/hostcheck steps the
// same functions on the host against the committed fixtures (developer harness, not a compute path).
#ifndef MX_PILEUP_CUH
#define MX_PILEUP_CUH
#include "mx_common.cuh"
#include "mx_tables.h"

namespace mx {

constexpr uint32_t kPuNoKey = 0xFFFFFFFFu;
constexpr int kPuProbTable = 64;
constexpr uint32_t kPuErrCcd = 1u, kPuErrCorrupt = 2u, kPuErrFrameTooLong = 4u, kPuErrPha = 8u, kPuErrOverflow = 16u, kPuErrFallback = 0x100u;

struct PileupArgs
{
   // input columns (read_input_event :573-620): detector.dat, xpixel.dat, ypixel.dat, time.dat, b_energy.dat + the dither files
   const int8_t *ccd; const float *x, *y, *t, *benergy; const float *dither[6];
   uint64_t n;
   double alpha, frame_time; uint64_t seed;
   double prob[kPuProbTable];                  // pow (alpha, k) from the host's libm (will_grade_migrate :668-674)
   uint32_t max_frame_events;
   const AcisDev *A;
   // per-event scratch
   uint32_t *frame, *key, *lo, *hi;
   float *pb, *px, *py; uint32_t *pn;          // pixel state at its representative (pn = 0: not a representative)
   float *ib; uint32_t *in;                    // island sums
   uint8_t *flag;                              // 0 rejected, 1 accepted without a draw, 2 draws
   float *sx, *sy; int16_t *spha; uint32_t *emit, *cum, *tile_sum;
   // output columns (write_event :622-666)
   int8_t *o_ccd; float *o_x, *o_y, *o_t, *o_benergy; int32_t *o_frame; int16_t *o_nphotons, *o_pha; float *o_dither[6];
   uint64_t max_out;
   unsigned long long *n_out; unsigned int *error;
};

MX_HD void pu_error (const PileupArgs &a, unsigned int bits)
{
#if defined(__CUDA_ARCH__)
   atomicOr (a.error, bits);
#else
   *a.error |= bits;
#endif
}

MX_HD void pu_frames (const PileupArgs &a, uint64_t e)
{
   const float ex = a.x[e], ey = a.y[e];
   const int ccd = a.ccd[e];
   const bool center = (ex >= 1) && (ex < 1024 - 1) && (ey >= 1) && (ey < 1024 - 1);                  // :559-560
   a.frame[e] = (unsigned int) ((double) a.t[e] / a.frame_time);                                     // :617
   uint32_t key = kPuNoKey;
   if ((ccd < 0) || (ccd >= 10)) pu_error (a, kPuErrCcd);
   else if (center) key = ((uint32_t) ccd << 20) | ((uint32_t) ey << 10) | (uint32_t) ex;
   else if ((ex >= 1024) || (ey >= 1024)) pu_error (a, kPuErrCorrupt);                                // "Corrupt file?" :562-568
   a.key[e] = key;
   a.pn[e] = 0; a.flag[e] = 0; a.emit[e] = 0;
}

// store_event :754-812 for the pixel whose representative is e (the LAST event of the frame on that pixel), replayed in list order
// (= reverse file order: e first, then the earlier events of the frame on the same pixel); cent = 1: the neighbours receive + 0.0 * e
MX_HD void pu_pixel_state (const PileupArgs &a, uint64_t e, uint64_t lo)
{
   const uint32_t key = a.key[e];
   float cb = 0.0f, cx = 0.0f, cy = 0.0f; uint32_t np = 0;
   for (uint64_t j = e + 1; j-- > lo; )
     {
        if (a.key[j] != key) continue;
        const float ex = a.x[j], ey = a.y[j];
        const double cent_e = 1.0 * (double) a.benergy[j];
        np += 1;
        cx = (float) ((double) (cx * cb) + cent_e * (double) ex);
        cy = (float) ((double) (cy * cb) + cent_e * (double) ey);
        cb = (float) ((double) cb + cent_e);
        cx = cx / cb;
        cy = cy / cb;
     }
   a.pb[e] = cb; a.px[e] = cx; a.py[e] = cy; a.pn[e] = np;
}

MX_HD void pu_store (const PileupArgs &a, uint64_t e)
{
   const uint32_t f = a.frame[e];
   uint64_t lo = e, hi = e + 1;
   uint32_t steps = 0;
   while ((lo > 0) && (a.frame[lo - 1] == f) && (steps < a.max_frame_events)) { lo--; steps++; }
   while ((hi < a.n) && (a.frame[hi] == f) && (steps < a.max_frame_events)) { hi++; steps++; }
   if (steps >= a.max_frame_events) { pu_error (a, kPuErrFrameTooLong); lo = e; hi = e + 1; }
   a.lo[e] = (uint32_t) lo; a.hi[e] = (uint32_t) hi;
   const uint32_t key = a.key[e];
   if (key == kPuNoKey) return;
   // a later event of the frame on the same pixel owns it: e left the list (:903-908)
   for (uint64_t j = e + 1; j < hi; j++) if (a.key[j] == key) return;
   pu_pixel_state (a, e, lo);
}

// the occupied pixels of e's 3 x 3 neighbourhood: slot [r][c] = representative event of pixel (y - 1 + r, x - 1 + c), or -1
struct PuHood { int32_t at[3][3]; };           // event indices stay below 2^32 - 1 (marxb200_pileup_run refuses longer lists)
MX_HD void pu_hood (const PileupArgs &a, uint64_t e, PuHood &h)
{
   const uint32_t key = a.key[e];
   const int ccd = (int) (key >> 20), iy = (int) ((key >> 10) & 1023u), ix = (int) (key & 1023u);
   for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) h.at[r][c] = -1;
   const uint64_t lo = a.lo[e], hi = a.hi[e];
   for (uint64_t j = lo; j < hi; j++)
     {
        if (a.pn[j] == 0) continue;
        const uint32_t k = a.key[j];
        if ((int) (k >> 20) != ccd) continue;
        const int dy = (int) ((k >> 10) & 1023u) - iy, dx = (int) (k & 1023u) - ix;
        if ((dy < -1) || (dy > 1) || (dx < -1) || (dx > 1)) continue;
        const int slot = (dy + 1) * 3 + (dx + 1);              // compile-time slots: the table stays in registers on the device
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 9; q++) if (q == slot) h.at[q / 3][q % 3] = (int32_t) j;
     }
}

// collect_charge :814-845, for the neighbourhood h of representative e
MX_HD void pu_island_from (const PileupArgs &a, uint64_t e, const PuHood &h)
{
   double s = 0.0; uint32_t np = 0;
   bool first = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
   for (int r = 0; r < 3; r++)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
     for (int c = 0; c < 3; c++)
       {
          const int32_t j = h.at[r][c];
          const float b = (j < 0) ? 0.0f : a.pb[j];
          const double term = (((r == 1) && (c == 1)) ? (1.0) : (9.0 / 9.0)) * (double) b;
          s = first ? term : s + term;
          first = false;
          np += (j < 0) ? 0u : a.pn[j];
       }
   a.ib[e] = (float) s; a.in[e] = np;
}
MX_HD void pu_island (const PileupArgs &a, uint64_t e)
{
   if (a.pn[e] == 0) return;
   PuHood h; pu_hood (a, e, h);
   pu_island_from (a, e, h);
}

// event_detect :676-752, the two local-maximum tests
MX_HD void pu_detect_from (const PileupArgs &a, uint64_t e, const PuHood &h)
{
   float pb[3][3], ib[3][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
   for (int r = 0; r < 3; r++)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
     for (int c = 0; c < 3; c++)
       {
          const int32_t j = h.at[r][c];
          pb[r][c] = (j < 0) ? 0.0f : a.pb[j];
          ib[r][c] = (j < 0) ? 0.0f : a.ib[j];
       }
   const double benergy = pb[1][1], island_benergy = ib[1][1];
   if (((pb[0][0] >= benergy) || (pb[0][1] >= benergy) || (pb[0][2] >= benergy))
       || (pb[1][0] > benergy) || (pb[1][2] >= benergy)
       || (pb[2][0] > benergy) || (pb[2][1] > benergy) || (pb[2][2] > benergy))
     return;
   if (((ib[0][0] >= island_benergy) || (ib[0][1] >= island_benergy) || (ib[0][2] >= island_benergy))
       || (ib[1][0] > island_benergy) || (ib[1][2] >= island_benergy)
       || (ib[2][0] > island_benergy) || (ib[2][1] > island_benergy) || (ib[2][2] > island_benergy))
     return;
   a.flag[e] = (a.in[e] >= 2) ? 2 : 1;
}
MX_HD void pu_detect (const PileupArgs &a, uint64_t e)
{
   if (a.pn[e] == 0) return;
   PuHood h; pu_hood (a, e, h);
   pu_detect_from (a, e, h);
}

// marx_map_energy_to_acis_pha, acis_fef.c:1087-1096 (find_fef :910-966 + JDMinterpolate_f); x, y arrive as ints
MX_HD int pu_energy_to_pha (const AcisDev &A, int ccd_id, int xi, int yi, double energy, int16_t &pha)
{
   float x = (float) xi, y = (float) yi;
   if ((x < 0) || (x >= 1024) || (y < 0) || (y >= 1024))
     {
        if (A.det_extend == 0) return -1;
        if (x < 0) x = 0; else if (x >= 1024) x = 1023;
        if (y < 0) y = 0; else if (y >= 1024) y = 1023;
     }
   const uint32_t i = (uint32_t) (x / 32), j = (uint32_t) (y / 32);
   if ((i >= 32) || (j >= 32)) return -1;
   const AcisChipDev *ch = nullptr;
   for (int k = 0; k < A.num_chips; k++) if (A.chip[k].id == ccd_id) ch = &A.chip[k];
   if (ch == nullptr) return -1;
   const int fi = ch->fef_map[i * 32 + j];
   if (fi < 0) return -1;
   const FefDev &f = A.fefs[fi];
   pha = (int16_t) interp_f ((float) energy, f.energies, f.channels, f.num_energies);
   return 0;
}

// the rest of event_detect (grade migration draw, centroid) and write_event's PHA.  k: number of islands of the frame that drew before
// this one (list order = reverse file order); h: the neighbourhood (read for flag == 2 only)
MX_HD void pu_emit_from (const PileupArgs &a, uint64_t e, uint32_t k, const PuHood &h)
{
   const uint32_t flag = a.flag[e];
   double x = a.x[e], y = a.y[e];
   const uint32_t np = a.in[e];
   if (flag == 2)
     {
        const uint32_t m = np - 1u;
        const double prob = (m < (uint32_t) kPuProbTable) ? a.prob[m] : pow (a.alpha, (double) m);
        Rng rng; rng.init (a.seed, (uint64_t) a.frame[e], 5u); rng.resume (k, 0, 0.0);
        if (rng.uniform () >= prob) return;
        x = 0; y = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < 3; c++)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int r = 0; r < 3; r++)
            {
               const int32_t j = h.at[r][c];
               const float qx = (j < 0) ? 0.0f : a.px[j], qy = (j < 0) ? 0.0f : a.py[j], qb = (j < 0) ? 0.0f : a.pb[j];
               x += (double) (qx * qb); y += (double) (qy * qb);
            }
        x /= (double) a.ib[e];
        y /= (double) a.ib[e];
     }
   const float xpix = (float) x, ypix = (float) y;
   int16_t pha = 0;
   if (-1 == pu_energy_to_pha (*a.A, (int) (a.key[e] >> 20), (int) xpix, (int) ypix, (double) a.ib[e], pha)) { pu_error (a, kPuErrPha); return; }
   a.sx[e] = xpix; a.sy[e] = ypix; a.spha[e] = pha; a.emit[e] = 1;
}
MX_HD void pu_emit (const PileupArgs &a, uint64_t e)
{
   const uint32_t flag = a.flag[e];
   if (flag == 0) return;
   uint32_t k = 0;
   PuHood h;
   for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) h.at[r][c] = -1;
   if (flag == 2)
     {
        const uint64_t hi = a.hi[e];
        for (uint64_t j = e + 1; j < hi; j++) k += (a.flag[j] == 2) ? 1u : 0u;
        pu_hood (a, e, h);
     }
   pu_emit_from (a, e, k, h);
}

// inclusive count of emitted rows up to and including event i (cum: inclusive inside the 256-event tile; tile_sum: exclusive over tiles)
MX_HD uint32_t pu_rows_through (const PileupArgs &a, uint64_t i) { return a.cum[i] + a.tile_sum[i >> 8]; }

MX_HD void pu_scatter (const PileupArgs &a, uint64_t e)
{
   if (e + 1 == a.n) *a.n_out = pu_rows_through (a, e);
   if (a.emit[e] == 0) return;
   const uint64_t lo = a.lo[e], hi = a.hi[e];
   const uint64_t before = (lo == 0) ? 0 : pu_rows_through (a, lo - 1);
   const uint64_t pos = before + (pu_rows_through (a, hi - 1) - pu_rows_through (a, e));
   if (pos >= a.max_out) { pu_error (a, kPuErrOverflow); return; }
   const uint32_t f = a.frame[e];
   a.o_ccd[pos] = a.ccd[e]; a.o_x[pos] = a.sx[e]; a.o_y[pos] = a.sy[e];
   a.o_frame[pos] = (int32_t) f; a.o_t[pos] = (float) ((int32_t) f * a.frame_time);
   a.o_nphotons[pos] = (int16_t) a.in[e]; a.o_pha[pos] = a.spha[e]; a.o_benergy[pos] = a.ib[e];
   for (int d = 0; d < 6; d++) if (a.dither[d] && a.o_dither[d]) a.o_dither[d][pos] = a.dither[d][e];
}

#if defined(__CUDACC__)
// pileup_kernels.cu: the eight launches on stream s (5 steps, 2 scan kernels, scatter)
void launch_pileup (const PileupArgs &a, int num_sms, cudaStream_t s, int *n_launches);
// the fused single-kernel form (frames staged in shared memory); sets kPuErrFallback in *a.error when a frame does not fit
size_t pileup_fused_scratch_bytes (uint64_t n);
struct PhotonSoA;
// cols: x, y, t, benergy, then the six dither columns
void launch_pileup_gather (const PhotonSoA &in, const unsigned long long *n_ptr, uint64_t max_n, const double *dev_start_time, double total_time,
                           int8_t *ccd, float *const cols[10], cudaStream_t s);
void launch_pileup_fused (const PileupArgs &a, void *scratch, int window, int num_sms, cudaStream_t s, int *n_launches);
#endif

}  // namespace mx
#endif
