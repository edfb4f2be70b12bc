// mx_context.hpp -- the context behind marxb200_ctx and the helpers shared by the translation units that implement the
// C ABI (marxb200.cu: life cycle, tables, stages, host boundary; comm.cu: the multi-GPU exchanges).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <utility>
#include "../../include/marxb200.h"
#include "mx_tables.h"
#include "mx_kernels.cuh"

struct MxComm;                                  // comm.cu
struct MxWriter;                                // writer.cpp
using mx::PhotonSoA; using mx::RayConst; using mx::SourceDev; using mx::DitherDev; using mx::TallyPlan; using mx::EgressPlan;
using mx::Level1Dev; using mx::Level1State; using mx::Level1Cols; using mx::kMaxEgressCols;

struct marxb200_ctx
{
   int device = 0;
   int num_sms = 0;
   uint64_t seed = 0;
   cudaStream_t stream = nullptr;
   bool own_stream = false;
   int compact = 1;

   // photon buffers
   uint64_t capacity = 0;
   void *slab[2] = {nullptr, nullptr};
   PhotonSoA buf[2];
   void *rc_slab = nullptr;
   RayConst rc;                                  // per-ray constants, indexed by batch slot (mx_kernels.cuh)
   int cur = 0;

   // device scalars: counts[0..3] + ticket + total_time
   // one allocation: counts[12] (generated, after mirror, after grating, detected, after k1a, after k1b, ...) followed by
   // ticket[12] (stage s, kernel k of its call: slot 4 (s - 1) + k), so that a traced batch clears counts[1..] and all tickets at once
   unsigned long long *d_counts = nullptr;
   unsigned long long *d_ticket = nullptr;      // = d_counts + kNumCounts
   static constexpr int kNumCounts = 12, kNumTickets = 12;
   int k01_ticket = 1;
   // Readers of the finished list on OTHER streams (slot 0: the packed egress on the copy stream, slot 1: the multi-GPU merge on
   // the merge stream): they convert the list to file images while the next batch starts.  reader_buf = the list buffer being
   // read (-1: none); the first kernel of the context's stream that writes into that buffer waits for ev_reader_done.
   cudaEvent_t ev_reader_go[2] = {nullptr, nullptr}, ev_reader_done[2] = {nullptr, nullptr};
   int reader_buf[2] = {-1, -1};
   // Pre-pack: a run that converts every batch to the same file images (marxb200_egress_begin_packed or marxb200_merge_events_begin
   // with the same columns, time offset and row limit as for the previous batch) gets them from the order restoration at the end of
   // the trace (order_gather<true>), written straight from the registers that hold the row; the begin call then finds them done and
   // launches no conversion kernel.  Two staging buffers alternate: the next batch's images are written while the previous batch's
   // are still being copied out.  owner: 0 = packed egress, 1 = merge.
   struct Prepack
   {
      bool armed = false, done = false;
      int owner = -1;
      mx::PackArgs args;
      cudaEvent_t ev_free = nullptr;           // recorded behind the copies that last read args.dst (may be null / never recorded)
   } prepack;
   void *packed_slab_b = nullptr;               // second staging buffer of the packed egress (the first is egress_slab)
   int packed_k = 0;                            // staging buffer of the batch being egressed
   cudaEvent_t ev_slab_free[2] = {nullptr, nullptr};
   unsigned long long *d_snap = nullptr;        // [0] event count, [1] batch start time (f64 bits): what the egress pack reads, snapshot in stream order
   bool batch_zeroed = false;                   // inside marxb200_trace(_sharded): the stage calls skip their own clears
   // time pre-pass of the NEXT contiguous batch, run on its own stream while this one is traced (marxb200_trace)
   cudaStream_t ahead_stream = nullptr;
   cudaEvent_t ev_ahead = nullptr, ev_ahead_consumed = nullptr;
   double *ahead_tile_sums = nullptr, *ahead_super_sums = nullptr;
   bool ahead_on = true, ahead_valid = false;
   uint64_t ahead_first = 0, ahead_n = 0, ahead_epoch = 0, ahead_hits = 0, ahead_misses = 0;
   uint32_t *d_bitmap = nullptr, *d_word_prefix = nullptr, *d_block_prefix = nullptr, *d_perm = nullptr;   // order restoration scratch
   uint64_t n_words = 0;
   bool ordered = true;                          // live list is in arrival order
   double *d_tile_sums = nullptr, *d_tile_base = nullptr, *d_super_sums = nullptr, *d_times = nullptr;   // d_times: [batch start, running end]
   int stage_done = -1;                          // index into d_counts of the latest valid count
   uint64_t n_generated = 0;

   // tables
   // every cudaMalloc'd table with the module that owns it (freed in destroy, or when the module's setter runs again)
   enum { TAG_MISC = 0, TAG_SOURCE, TAG_DITHER, TAG_HRMA, TAG_GRATING, TAG_DETECTOR, TAG_LEVEL1 };
   std::vector<std::pair<void *, int>> allocs;
   int alloc_tag = TAG_MISC;
   uint64_t *d_upload_ids = nullptr; uint64_t upload_ids_cap = 0;      // ray ids of marxb200_upload
   SourceDev S; DitherDev D;
   bool have_source = false, have_dither = false, have_hrma = false, have_grating = false, have_acis = false;
   int grating_type = 0, detector_type = 0;
   double source_distance = 0.0;
   void *blob1 = nullptr, *blob2 = nullptr, *blob3 = nullptr;
   uint32_t blob1_bytes = 0, blob2_bytes = 0, blob3_bytes = 0;
   uint32_t k1b_bytes = 0, k1c_seg2_off = 0, k1c_seg2_bytes = 0, k1b1_bytes = 0, k1b2_seg2_off = 0;
   struct Tally { TallyPlan plan; unsigned long long *bins; uint64_t total; };
   std::vector<Tally> tallies;
   bool det_dither_dirty = false;                // uploaded photons may carry detector dither: the per-ray columns are live
   double aspsol_t_last = 0.0;                   // ASPSOL dither: time of the last state (rays at or beyond it end the run)
   int grid1[6] = {0, 0, 0, 0, 0, 0}, grid2 = 0, grid3 = 0, grid01 = 0;
   bool detector_is_hrc = false;
   bool mirror_is_flat = false; double ff[5] = {0, 0, 0, 0, 0};      // MirrorType=FLATFIELD (marxb200_set_flatfield)
   int first_mirror_kernel = 0;                  // 1: phase A already ran fused with the source (marxb200_trace)
   int k3_split = 1;                             // ACIS detector stage as two kernels (MARXB200_K3_SPLIT=0: one kernel)
   int k2_split = 1;                             // compacting grating stage as k2_select + k2_grating<1> (MARXB200_K2_SPLIT=0: one kernel)
   int k1_split = 1;                             // compacting mirror stage cut behind the reflectivity tests: A | B1 | B2+C1 | C2
                                                 // (MARXB200_K1_SPLIT=0: A | B | C, which the in-place parity mode always runs)

   // host boundary staging
   void *d_aos = nullptr; uint64_t d_aos_cap = 0;
   void *h_pinned = nullptr; size_t h_pinned_bytes = 0;

   uint64_t launches = 0;
   uint64_t egress_rows[32] = {0};               // cumulative rows per column file (marxio.c File_Pointers[].num_rows)

   // pipelined egress
   cudaStream_t copy_stream = nullptr;
   cudaEvent_t ev_staged = nullptr, ev_copied = nullptr;
   void *egress_slab = nullptr; uint64_t egress_cap = 0; PhotonSoA egress;
   unsigned long long *h_egress_count = nullptr;      // pinned
   bool egress_pending = false, egress_is_packed = false;
   EgressPlan packed_plan; int packed_which[kMaxEgressCols]; uint64_t packed_cap = 0;

   // Level-1 event transforms (marxb200_level1_*)
   bool have_level1 = false;
   Level1Dev L1;
   Level1State *d_l1_state = nullptr;
   void *l1_slab = nullptr; uint64_t l1_cap = 0; Level1Cols l1_cols;
   uint32_t *d_l1_head = nullptr, *d_l1_tile_head = nullptr;
   float *d_l1_next_dither = nullptr; long long *d_l1_next_expno = nullptr; unsigned int *d_l1_error = nullptr;
   uint64_t l1_rows = 0;                         // rows of the last transform (what marxb200_level1_download returns)

   // optional per-kernel timing
   bool profiling = false;
   uint64_t source_epoch = 0;                 // bumped by marxb200_set_source: invalidates work done ahead for the old source
   cudaEvent_t ev_prev = nullptr;
   std::vector<std::pair<cudaEvent_t, int>> ev_marks;   // (event recorded after a kernel, class)
   std::vector<cudaEvent_t> ev_pool;
   double prof_ms[MARXB200_NUM_KERNEL_CLASSES] = {0};
   uint64_t prof_n[MARXB200_NUM_KERNEL_CLASSES] = {0};

   // background column-file writer (marxb200_set_async_writer): two pinned batch buffers, file k on thread k mod n
   MxWriter *writer = nullptr;
   void *h_wbuf[2] = {nullptr, nullptr}; size_t h_wbuf_bytes[2] = {0, 0}; int wbuf_next = 0;

   // multi-GPU exchanges (comm.cu): NCCL communicators, time-base all-gather scratch, event-merge buffers
   MxComm *comm = nullptr;
};

// error convention of the ABI (marxerr.c:26-54): set the thread's message, return -1
int mxb_fail (const char *fmt, ...);
#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return mxb_fail ("%s: %s", #expr, cudaGetErrorString (e_)); } while (0)

// per-kernel CUDA-event marks (marxb200_set_profiling)
void mxb_prof_begin (marxb200_ctx *c);
void mxb_prof_mark (marxb200_ctx *c, int cls);
// arrival-order restoration of the live list, and the list as the host boundary may see it (detector-dither columns nulled
// when they are not live)
int mxb_ensure_order (marxb200_ctx *c);
mx::PhotonSoA mxb_observed (const marxb200_ctx *c, const mx::PhotonSoA &b);
// source arguments of one batch (time_base < 0: continue the running sum kept on the device)
void mxb_fill_source_args (marxb200_ctx *c, mx::SourceArgs &a, uint64_t first_ray, uint64_t n, double time_base);
// the fused source + HRMA phase A entry behind an already computed time scan; then the remaining stages of one batch
int mxb_enter_mirror_after_scan (marxb200_ctx *c, const mx::SourceArgs &a);
int mxb_finish_trace (marxb200_ctx *c);
int mxb_begin_batch (marxb200_ctx *c);
bool mxb_prepack_matches (const marxb200_ctx *c, int owner, const mx::PackArgs &want);   // the current list's images exist in want.dst
void mxb_prepack_arm (marxb200_ctx *c, int owner, const mx::PackArgs &next, cudaEvent_t ev_free);
int mxb_guard_buffer (marxb200_ctx *c, int idx);   // before a kernel of the context's stream writes list buffer idx (-1: any)
int mxb_reader_begin (marxb200_ctx *c, int slot, cudaStream_t reader);   // reader waits for the list; call mxb_reader_end after its launches
int mxb_reader_end (marxb200_ctx *c, int slot, cudaStream_t reader);          // clears the batch's counters and tickets with one memset
// the column files of marx_write_photons selected by write_mask, each in a fixed region of rows_per_col rows
struct MxEgressCol { uint64_t mask; const char *file, *colname; char type; int kind; int size; };
extern const MxEgressCol kMxEgressCols[];
extern const int kMxNumEgressCols;
uint64_t mxb_build_egress_plan (uint64_t write_mask, uint64_t rows_per_col, mx::EgressPlan &plan, int *which);
// comm.cu: release the communicators and buffers of a context (called by marxb200_destroy)
void mxb_comm_release (marxb200_ctx *c);
