// comm.cu -- the multi-GPU exchanges of the ray-trace path behind the C ABI (include/marxb200.h, "multi-GPU").
//
// One process per GPU.  The path shards by contiguous blocks of global ray indices (counter-based draw streams: every ray is
// the same ray on any GPU), so three exchanges remain, all on device buffers over NVLink / NVSwitch:
//
//   1. arrival times are ONE running sum over all rays (source.c:326): the ranks all-gather the super-tile sums the trace's
//      own pre-pass produces (ncclAllGather, n/65536 doubles per rank) and k0_time_bases_sharded adds them in global ray
//      order on every GPU -- marxb200_trace_sharded.  No host round trip, no second pass over the increments.
//   2. the per-GPU event lists are merged on one GPU in arrival order, which for contiguous blocks is rank order (the
//      time-ordered merge of marxcat, marx/src/marxcat.c:505-535, degenerates to a concatenation): every rank converts its
//      list to the column-file images of marx_write_photons on the device, the counts are all-gathered, and each column
//      goes straight into its place in the destination GPU's merged column -- by copy-engine peer writes into the
//      destination's buffer (mapped into every rank with CUDA IPC; no SM is taken from the next batch's kernels), or by
//      grouped ncclSend / ncclRecv when peer mapping is unavailable -- marxb200_merge_events_begin / _end.
//   3. tallies (integer histograms on the device) are summed in place -- marxb200_tally_allreduce (ncclAllReduce).
//
// NCCL is opened at run time (dlopen "libnccl.so.2": the library a host process already carries, e.g. the one bundled with
// PyTorch, or the system's): a single-GPU user of libmarxb200.so needs no NCCL at all.  The merge runs on its own stream
// and its own communicator (ncclCommSplit), so that it overlaps the next batch's trace, whose time-base all-gather uses
// the context's stream and the first communicator.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <mutex>
#include <string>
#include "mx_context.hpp"

using namespace mx;
#define fail mxb_fail

// ---------------------------------------------------------------------------------------------
// the handful of NCCL entry points this file uses, declared as in nccl.h (2.18 or newer: ncclCommSplit)
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *nccl_comm_t;
typedef struct { char internal[MARXB200_COMM_ID_BYTES]; } nccl_unique_id;
enum { NCCL_CHAR = 0, NCCL_INT32 = 2, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };     // ncclDataType_t
enum { NCCL_SUM = 0, NCCL_MIN = 3 };                                            // ncclRedOp_t

// ncclConfig_t as of NCCL 2.18 (the first release with ncclCommSplit): newer libraries accept an older, shorter struct (they read
// `size` bytes and default the rest).  Used to cap the collectives of this path -- a few kilobytes each -- at ONE CTA: an NCCL
// kernel spins while it waits for its peers, and every CTA it holds is a slot the persistent trace kernels cannot use.
struct nccl_config_v21800
{
   size_t size; unsigned int magic; unsigned int version;
   int blocking, cgaClusterSize, minCTAs, maxCTAs;
   const char *netName;
   int splitShare;
};
const int kNcclUndefInt = (int) 0x80000000;      // NCCL_CONFIG_UNDEF_INT (INT_MIN)
nccl_config_v21800 one_cta_config ()
{
   nccl_config_v21800 c;
   c.size = sizeof (c); c.magic = 0xcafebeef; c.version = 21800;
   c.blocking = kNcclUndefInt; c.cgaClusterSize = kNcclUndefInt; c.minCTAs = 1; c.maxCTAs = 1;
   c.netName = nullptr; c.splitShare = kNcclUndefInt;
   return c;
}

struct NcclApi
{
   void *handle = nullptr;
   int version = 0;
   int (*GetVersion) (int *) = nullptr;
   int (*GetUniqueId) (nccl_unique_id *) = nullptr;
   int (*CommInitRank) (nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
   int (*CommInitRankConfig) (nccl_comm_t *, int, nccl_unique_id, int, void *) = nullptr;
   int (*CommSplit) (nccl_comm_t, int, int, nccl_comm_t *, void *) = nullptr;
   int (*CommDestroy) (nccl_comm_t) = nullptr;
   const char *(*GetErrorString) (int) = nullptr;
   int (*AllGather) (const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
   int (*AllReduce) (const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
   int (*Broadcast) (const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
   int (*Send) (const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
   int (*Recv) (void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
   int (*GroupStart) (void) = nullptr;
   int (*GroupEnd) (void) = nullptr;
   std::string error;
};

NcclApi *nccl_api ()
{
   static NcclApi api;
   static std::once_flag once;
   std::call_once (once, [] ()
     {
        const char *names[] = {getenv ("MARXB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
          {
             if ((n == nullptr) || (*n == 0)) continue;
             api.handle = dlopen (n, RTLD_NOW | RTLD_GLOBAL);
             if (api.handle) break;
             api.error = dlerror ();
          }
        if (api.handle == nullptr) return;
        bool ok = true;
#define SYM(field, name) do { *(void **) (&api.field) = dlsym (api.handle, name); if (api.field == nullptr) { ok = false; api.error = std::string ("missing symbol ") + name; } } while (0)
        SYM (GetVersion, "ncclGetVersion"); SYM (GetUniqueId, "ncclGetUniqueId"); SYM (CommInitRank, "ncclCommInitRank"); SYM (CommInitRankConfig, "ncclCommInitRankConfig");
        SYM (CommSplit, "ncclCommSplit"); SYM (CommDestroy, "ncclCommDestroy"); SYM (GetErrorString, "ncclGetErrorString");
        SYM (AllGather, "ncclAllGather"); SYM (AllReduce, "ncclAllReduce"); SYM (Broadcast, "ncclBroadcast");
        SYM (Send, "ncclSend"); SYM (Recv, "ncclRecv"); SYM (GroupStart, "ncclGroupStart"); SYM (GroupEnd, "ncclGroupEnd");
#undef SYM
        if (!ok) { dlclose (api.handle); api.handle = nullptr; return; }
        api.GetVersion (&api.version);
     });
   return (api.handle != nullptr) ? &api : nullptr;
}
int no_nccl ()
{
   return fail ("NCCL is not available (dlopen libnccl.so.2 / $MARXB200_NCCL_LIB failed); the multi-GPU entry points need it");
}
}  // namespace

#define NCCL_OK(expr) do { int r_ = (expr); if (r_ != 0) return fail ("%s: %s", #expr, N->GetErrorString (r_)); } while (0)

enum { MERGE_NCCL = 0, MERGE_PEER = 1 };

struct MxComm
{
   nccl_comm_t comm = nullptr, comm_merge = nullptr;
   int rank = 0, world = 1;
   // time-base exchange
   double *d_all_sums = nullptr; uint64_t all_sums_cap = 0;        // [world][ns_blk]
   // time-base look-ahead: the pre-pass and the all-gather of the NEXT contiguous batch run on their own stream while this
   // batch is traced, so that neither the collective's latency nor the wait for the slowest rank is on the critical path
   cudaStream_t ahead_stream = nullptr;
   cudaEvent_t ev_ahead = nullptr, ev_consumed = nullptr;
   double *ahead_tile_sums = nullptr, *ahead_super_sums = nullptr, *ahead_all_sums = nullptr;
   uint64_t ahead_capacity = 0;                                     // rays the look-ahead scratch was sized for
   bool ahead_on = true, ahead_valid = false;
   uint64_t ahead_first = 0, ahead_total = 0, ahead_epoch = 0;
   uint64_t ahead_hits = 0, ahead_misses = 0;
   // event merge
   cudaStream_t merge_stream = nullptr;
   cudaEvent_t ev_packed = nullptr, ev_counts = nullptr, ev_pushed = nullptr, ev_t0 = nullptr, ev_tc = nullptr, ev_t1 = nullptr;
   unsigned long long *d_all_counts = nullptr, *h_all_counts = nullptr;     // [world]; pinned host copy
   int *d_flag = nullptr;                                                   // barrier / agreement scratch
   int dst = -1; uint64_t max_rows = 0, mask = 0;
   EgressPlan plan; int which[kMaxEgressCols]; uint64_t stage_bytes = 0;
   void *stage = nullptr;                     // this rank's packed columns of the batch being merged: column c at plan.offset[c], max_rows rows each
   void *stage_bufs[2] = {nullptr, nullptr};  // two staging areas alternate (pre-pack, mx_context.hpp); stage == stage_bufs[stage_k]
   int stage_k = 0;
   cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};      // behind the transfers that last read stage_bufs[k]
   void *merged = nullptr;                    // dst: merged columns, column c at merged_off[c], world * max_rows rows each
   void *peer_merged = nullptr;               // other ranks, MERGE_PEER: the destination's buffer mapped with CUDA IPC
   uint64_t merged_off[kMaxEgressCols]; uint64_t merged_bytes = 0;
   int transport = MERGE_NCCL;
   bool pending = false;
   uint64_t total_rows = 0;                   // of the last finished merge
   uint64_t counts[64];
   float last_ms = 0.f, last_copy_ms = 0.f; uint64_t last_bytes = 0;
};

// ---------------------------------------------------------------------------------------------
// communicator life cycle
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_comm_get_unique_id (void *id)
{
   if (id == nullptr) return fail ("marxb200_comm_get_unique_id: NULL argument");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   static_assert (sizeof (nccl_unique_id) == MARXB200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
   nccl_unique_id u;
   NCCL_OK (N->GetUniqueId (&u));
   memcpy (id, &u, sizeof (u));
   return 0;
}

static void merge_release (marxb200_ctx *c);

void mxb_comm_release (marxb200_ctx *c)
{
   MxComm *m = c->comm;
   if (m == nullptr) return;
   NcclApi *N = nccl_api ();
   if (getenv ("MARXB200_VERBOSE"))
     fprintf (stderr, "marxb200: rank %d time-base look-ahead: %llu batches found their sums exchanged ahead, %llu ran the pre-pass themselves\n",
              m->rank, (unsigned long long) m->ahead_hits, (unsigned long long) m->ahead_misses);
   merge_release (c);
   if (m->merge_stream) cudaStreamDestroy (m->merge_stream);
   if (m->ahead_stream) { cudaStreamSynchronize (m->ahead_stream); cudaStreamDestroy (m->ahead_stream); }
   for (cudaEvent_t e : {m->ev_packed, m->ev_counts, m->ev_pushed, m->ev_t0, m->ev_tc, m->ev_t1, m->ev_ahead, m->ev_consumed}) if (e) cudaEventDestroy (e);
   for (double *p : {m->ahead_tile_sums, m->ahead_super_sums, m->ahead_all_sums}) if (p) cudaFree (p);
   for (int k = 0; k < 2; k++) if (m->ev_stage_free[k]) cudaEventDestroy (m->ev_stage_free[k]);
   if (m->d_all_sums) cudaFree (m->d_all_sums);
   if (m->d_all_counts) cudaFree (m->d_all_counts);
   if (m->h_all_counts) cudaFreeHost (m->h_all_counts);
   if (m->d_flag) cudaFree (m->d_flag);
   if (N)
     {
        if (m->comm_merge) N->CommDestroy (m->comm_merge);
        if (m->comm) N->CommDestroy (m->comm);
     }
   delete m;
   c->comm = nullptr;
}

extern "C" int marxb200_comm_destroy (marxb200_ctx *c)
{
   if (c == nullptr) return fail ("NULL ctx");
   cudaSetDevice (c->device);
   cudaDeviceSynchronize ();
   mxb_comm_release (c);
   return 0;
}

extern "C" int marxb200_comm_init (marxb200_ctx *c, const void *id, int rank, int world)
{
   if ((c == nullptr) || (id == nullptr)) return fail ("marxb200_comm_init: NULL argument");
   if ((world < 1) || (world > 64) || (rank < 0) || (rank >= world)) return fail ("marxb200_comm_init: rank %d of %d", rank, world);
   if (c->comm != nullptr) return fail ("marxb200_comm_init: the context already has a communicator");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   CUDA_OK (cudaSetDevice (c->device));
   MxComm *m = new MxComm ();
   m->rank = rank; m->world = world;
   c->comm = m;
   nccl_unique_id u;
   memcpy (&u, id, sizeof (u));
   int status = 0;
   do
     {
#define INIT_OK(expr, what) { if (0 != (expr)) { status = fail ("marxb200_comm_init: %s failed: %s", what, cudaGetErrorString (cudaGetLastError ())); break; } }
        // MARXB200_NCCL_CTAS=0 leaves NCCL's own choice of CTAs per collective (A/B runs); the default caps it at one
        const char *ctas = getenv ("MARXB200_NCCL_CTAS");
        const bool cap = (ctas == nullptr) || (atoi (ctas) != 0);
        nccl_config_v21800 cfg = one_cta_config ();
        int r = cap ? N->CommInitRankConfig (&m->comm, world, u, rank, &cfg) : N->CommInitRank (&m->comm, world, u, rank);
        if (r != 0) { status = fail ("marxb200_comm_init: ncclCommInitRank: %s", N->GetErrorString (r)); break; }
        // a second communicator for the merge stream: collectives of one communicator must be issued in one order on all ranks,
        // and the merge of batch k is interleaved with the trace of batch k+1.  With the peer-write transport it only carries the
        // counts and the closing barrier (one CTA); the ncclSend / ncclRecv transport moves the columns with it and keeps NCCL's CTAs.
        const char *force = getenv ("MARXB200_MERGE_TRANSPORT");
        const bool merge_moves_data = (force != nullptr) && (0 == strcmp (force, "nccl"));
        nccl_config_v21800 cfg2 = one_cta_config ();
        r = N->CommSplit (m->comm, 0, rank, &m->comm_merge, (cap && !merge_moves_data) ? (void *) &cfg2 : nullptr);
        if (r != 0) { status = fail ("marxb200_comm_init: ncclCommSplit: %s", N->GetErrorString (r)); break; }
        INIT_OK (cudaStreamCreateWithFlags (&m->merge_stream, cudaStreamNonBlocking), "stream");
        INIT_OK (cudaEventCreateWithFlags (&m->ev_packed, cudaEventDisableTiming), "event");
        INIT_OK (cudaEventCreateWithFlags (&m->ev_counts, cudaEventDisableTiming), "event");
        INIT_OK (cudaEventCreateWithFlags (&m->ev_pushed, cudaEventDisableTiming), "event");
        INIT_OK (cudaStreamCreateWithFlags (&m->ahead_stream, cudaStreamNonBlocking), "stream");
        INIT_OK (cudaEventCreateWithFlags (&m->ev_ahead, cudaEventDisableTiming), "event");
        INIT_OK (cudaEventCreateWithFlags (&m->ev_consumed, cudaEventDisableTiming), "event");
        if (const char *e = getenv ("MARXB200_LOOKAHEAD")) m->ahead_on = (atoi (e) != 0);      // A/B switch
        INIT_OK (cudaEventCreate (&m->ev_t0), "event");
        INIT_OK (cudaEventCreate (&m->ev_tc), "event");
        INIT_OK (cudaEventCreate (&m->ev_t1), "event");
        INIT_OK (cudaMalloc (&m->d_all_counts, 65 * sizeof (unsigned long long)), "cudaMalloc");      // [world] + this rank's own snapshot
        INIT_OK (cudaMallocHost (&m->h_all_counts, 64 * sizeof (unsigned long long)), "cudaMallocHost");
        INIT_OK (cudaMalloc (&m->d_flag, 64 * sizeof (int)), "cudaMalloc");
        INIT_OK (cudaMemset (m->d_flag, 0, 64 * sizeof (int)), "cudaMemset");
#undef INIT_OK
     }
   while (0);
   if (status != 0) { mxb_comm_release (c); return status; }
   return 0;
}

// Rendezvous through a file both sides can see (N `marx` processes on one box have no launcher to carry the id): rank 0
// writes the 128-byte id to <path>.tmp and renames it to <path>; the others poll for it.  The path must be fresh for every
// job; rank 0 removes the file once every rank has joined (ncclCommInitRank returns only then).
extern "C" int marxb200_comm_init_file (marxb200_ctx *c, const char *path, int rank, int world, double timeout_s)
{
   if ((c == nullptr) || (path == nullptr)) return fail ("marxb200_comm_init_file: NULL argument");
   unsigned char id[MARXB200_COMM_ID_BYTES];
   if (rank == 0)
     {
        if (-1 == marxb200_comm_get_unique_id (id)) return -1;
        const std::string tmp = std::string (path) + ".tmp";
        FILE *fp = fopen (tmp.c_str (), "wb");
        if ((fp == nullptr) || (sizeof (id) != fwrite (id, 1, sizeof (id), fp)) || (0 != fclose (fp)) || (0 != rename (tmp.c_str (), path)))
          return fail ("marxb200_comm_init_file: unable to write %s", path);
     }
   else
     {
        const double t_end = (double) time (nullptr) + ((timeout_s > 0.0) ? timeout_s : 60.0);
        bool got = false;
        while (!got)
          {
             FILE *fp = fopen (path, "rb");
             if (fp != nullptr)
               {
                  got = (sizeof (id) == fread (id, 1, sizeof (id), fp));
                  fclose (fp);
               }
             if (got) break;
             if ((double) time (nullptr) > t_end) return fail ("marxb200_comm_init_file: no id in %s after %.0f s", path, timeout_s);
             usleep (5000);
          }
     }
   const int status = marxb200_comm_init (c, id, rank, world);
   if (rank == 0) unlink (path);
   return status;
}

extern "C" int marxb200_comm_info (marxb200_ctx *c, int *rank, int *world, int *nccl_version, int *merge_transport)
{
   if (c == nullptr) return fail ("NULL ctx");
   if (c->comm == nullptr) return fail ("marxb200_comm_info: no communicator (marxb200_comm_init)");
   NcclApi *N = nccl_api ();
   if (rank) *rank = c->comm->rank;
   if (world) *world = c->comm->world;
   if (nccl_version) *nccl_version = N ? N->version : 0;
   if (merge_transport) *merge_transport = c->comm->transport;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// 1. sharded trace: rays [first_ray, first_ray + n_total) in `world` contiguous blocks, rank r traces block r
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_shard_of (uint64_t first_ray, uint64_t n_total, int rank, int world, uint64_t *my_first_ray, uint64_t *my_n)
{
   if ((world < 1) || (rank < 0) || (rank >= world)) return fail ("marxb200_shard_of: rank %d of %d", rank, world);
   const uint64_t super = (uint64_t) kTile * kSuperTile;                 // 65536 rays: the unit of the canonical time sum
   uint64_t blk = (n_total + (uint64_t) world - 1) / (uint64_t) world;
   blk = ((blk + super - 1) / super) * super;
   const uint64_t lo = (uint64_t) rank * blk;
   const uint64_t n = (lo >= n_total) ? 0 : ((n_total - lo < blk) ? n_total - lo : blk);
   if (my_first_ray) *my_first_ray = first_ray + lo;
   if (my_n) *my_n = n;
   return 0;
}

// The pre-pass of the batch a run asks for next -- rays [first_ray, first_ray + n_total) right behind the one just launched -- and the
// all-gather of its sums, on the look-ahead stream.  The sums depend on the seed, the source and the ray indices only; a call that
// asks for something else (or follows a marxb200_set_source) finds no match and runs the pre-pass itself.
static int look_ahead (marxb200_ctx *c, NcclApi *N, uint64_t first_ray, uint64_t n_total, uint32_t ns_blk)
{
   MxComm *m = c->comm;
   if (!m->ahead_on || c->profiling || (first_ray + n_total < first_ray)) return 0;
   const uint64_t super = (uint64_t) kTile * kSuperTile;
   if (m->ahead_capacity != c->capacity)
     {
        CUDA_OK (cudaStreamSynchronize (m->ahead_stream));
        CUDA_OK (cudaStreamSynchronize (c->stream));
        for (double **p : {&m->ahead_tile_sums, &m->ahead_super_sums, &m->ahead_all_sums}) { if (*p) cudaFree (*p); *p = nullptr; }
        m->ahead_capacity = 0;
        const uint64_t n_super = (c->capacity + super - 1) / super + 2;
        CUDA_OK (cudaMalloc (&m->ahead_tile_sums, n_super * kSuperTile * sizeof (double)));
        CUDA_OK (cudaMalloc (&m->ahead_super_sums, n_super * sizeof (double)));
        CUDA_OK (cudaMalloc (&m->ahead_all_sums, n_super * m->world * sizeof (double)));
        m->ahead_capacity = c->capacity;
     }
   uint64_t first = 0, n = 0;
   marxb200_shard_of (first_ray, n_total, m->rank, m->world, &first, &n);
   SourceArgs b;
   mxb_fill_source_args (c, b, first, n, 0.0);
   b.tile_sums = m->ahead_tile_sums; b.supertile_sums = m->ahead_super_sums;
   // the scratch is free once the bases of the current batch have been formed from it
   CUDA_OK (cudaStreamWaitEvent (m->ahead_stream, m->ev_consumed, 0));
   CUDA_OK (cudaMemsetAsync (m->ahead_super_sums, 0, (size_t) ns_blk * sizeof (double), m->ahead_stream));
   launch_time_sums (b, m->ahead_stream);
   launch_time_super (b, m->ahead_stream);
   NCCL_OK (N->AllGather (m->ahead_super_sums, m->ahead_all_sums, ns_blk, NCCL_FLOAT64, m->comm, m->ahead_stream));
   c->launches += (n != 0) ? 2 : 0;
   CUDA_OK (cudaGetLastError ());
   CUDA_OK (cudaEventRecord (m->ev_ahead, m->ahead_stream));
   m->ahead_valid = true; m->ahead_first = first_ray; m->ahead_total = n_total; m->ahead_epoch = c->source_epoch;
   return 0;
}

extern "C" int marxb200_trace_sharded (marxb200_ctx *c, uint64_t first_ray, uint64_t n_total, double time_base_in,
                                       uint64_t *my_first_ray, uint64_t *my_n)
{
   if (c == nullptr) return fail ("NULL ctx");
   MxComm *m = c->comm;
   if (m == nullptr) return fail ("marxb200_trace_sharded: no communicator (marxb200_comm_init)");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   if (!c->have_source) return fail ("marxb200_trace_sharded: no source set");
   if (!c->have_hrma) return fail ("marxb200_trace_sharded: no HRMA tables set");
   if (!c->compact) return fail ("marxb200_trace_sharded: compaction must be on");
   // the ASPSOL model ends a run at the end of the aspect file (a cut inside one rank's block) and the ExposureTime cut
   // likewise: those runs go block by block through marxb200_create_photons / marxb200_truncate_exposure
   if ((c->D.mode == 2) || c->det_dither_dirty) return fail ("marxb200_trace_sharded: the ASPSOL dither model is not sharded");
   if (c->mirror_is_flat) return fail ("marxb200_trace_sharded: MirrorType=FLATFIELD is not sharded");
   uint64_t first = 0, n = 0;
   marxb200_shard_of (first_ray, n_total, m->rank, m->world, &first, &n);
   if (n > c->capacity) return fail ("marxb200_trace_sharded: this rank's block of %llu rays exceeds the allocated capacity %llu",
                                     (unsigned long long) n, (unsigned long long) c->capacity);
   CUDA_OK (cudaSetDevice (c->device));
   const uint64_t super = (uint64_t) kTile * kSuperTile;
   uint64_t blk = (n_total + (uint64_t) m->world - 1) / (uint64_t) m->world;
   const uint32_t ns_blk = (uint32_t) ((blk + super - 1) / super);
   if (ns_blk == 0)
     {
        if (my_first_ray) *my_first_ray = first;
        if (my_n) *my_n = 0;
        return fail ("marxb200_trace_sharded: n_total must be > 0");
     }
   if (m->all_sums_cap < (uint64_t) ns_blk * m->world)
     {
        CUDA_OK (cudaStreamSynchronize (c->stream));
        if (m->d_all_sums) cudaFree (m->d_all_sums);
        m->d_all_sums = nullptr; m->all_sums_cap = 0;
        // room for the largest block this context can trace
        uint64_t cap = (c->capacity + super - 1) / super + 1;
        if (cap < ns_blk) cap = ns_blk;
        CUDA_OK (cudaMalloc (&m->d_all_sums, cap * m->world * sizeof (double)));
        m->all_sums_cap = cap * m->world;
     }
   // Was this batch's pre-pass already run (and its sums exchanged) behind the previous call?  Every rank sees the same arguments,
   // hence takes the same branch.  Either way the context's stream waits for the look-ahead: it used the same communicator.
   if (-1 == mxb_begin_batch (c)) return -1;
   const bool hit = m->ahead_valid && (m->ahead_first == first_ray) && (m->ahead_total == n_total) && (m->ahead_epoch == c->source_epoch);
   if (m->ahead_valid) CUDA_OK (cudaStreamWaitEvent (c->stream, m->ev_ahead, 0));
   m->ahead_valid = false;
   SourceArgs a;
   mxb_fill_source_args (c, a, first, n, time_base_in);
   mxb_prof_begin (c);
   if (hit)
     {
        m->ahead_hits++;
        SourceArgs b = a;
        b.tile_sums = m->ahead_tile_sums;
        mxb_prof_mark (c, 0);
        launch_time_bases_sharded (b, m->ahead_all_sums, m->rank, m->world, ns_blk, c->stream); mxb_prof_mark (c, 1);
        c->launches += (n != 0) ? 2 : 1;
     }
   else
     {
        m->ahead_misses++;
        // the scratch holds capacity / 65536 + 2 sums, ns_blk <= capacity / 65536 + 1: zero the padding of a short block
        CUDA_OK (cudaMemsetAsync (c->d_super_sums, 0, (size_t) ns_blk * sizeof (double), c->stream));
        launch_time_sums (a, c->stream); mxb_prof_mark (c, 0);
        launch_time_super (a, c->stream);
        NCCL_OK (N->AllGather (c->d_super_sums, m->d_all_sums, ns_blk, NCCL_FLOAT64, m->comm, c->stream));
        launch_time_bases_sharded (a, m->d_all_sums, m->rank, m->world, ns_blk, c->stream); mxb_prof_mark (c, 1);
        c->launches += (n != 0) ? 4 : 1;
     }
   CUDA_OK (cudaGetLastError ());
   CUDA_OK (cudaEventRecord (m->ev_consumed, c->stream));
   if (-1 == mxb_enter_mirror_after_scan (c, a)) return -1;
   if (-1 == mxb_finish_trace (c)) return -1;
   if ((first_ray + n_total > first_ray) && (-1 == look_ahead (c, N, first_ray + n_total, n_total, ns_blk))) return -1;
   if (my_first_ray) *my_first_ray = first;
   if (my_n) *my_n = n;
   return 0;
}

// ---------------------------------------------------------------------------------------------
// 3. tallies: one in-place all-reduce on the context's stream (ordered with the accumulate / read calls around it)
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_tally_allreduce (marxb200_ctx *c, int id)
{
   if (c == nullptr) return fail ("NULL ctx");
   MxComm *m = c->comm;
   if (m == nullptr) return fail ("marxb200_tally_allreduce: no communicator (marxb200_comm_init)");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   if ((id < 0) || (id >= (int) c->tallies.size ())) return fail ("unknown tally id %d", id);
   CUDA_OK (cudaSetDevice (c->device));
   marxb200_ctx::Tally &t = c->tallies[id];
   if (m->ahead_valid) CUDA_OK (cudaStreamWaitEvent (c->stream, m->ev_ahead, 0));      // the look-ahead's all-gather uses the same communicator
   NCCL_OK (N->AllReduce (t.bins, t.bins, (size_t) t.total, NCCL_UINT64, NCCL_SUM, m->comm, c->stream));
   return 0;
}

// ---------------------------------------------------------------------------------------------
// 2. event-list merge
// ---------------------------------------------------------------------------------------------
static void merge_release (marxb200_ctx *c)
{
   MxComm *m = c->comm;
   if (m == nullptr) return;
   if (m->merge_stream) cudaStreamSynchronize (m->merge_stream);
   if (m->peer_merged) cudaIpcCloseMemHandle (m->peer_merged);
   if (m->merged) cudaFree (m->merged);
   if (c->prepack.owner == 1) { c->prepack.armed = false; c->prepack.done = false; }
   for (int k = 0; k < 2; k++) { if (m->stage_bufs[k]) cudaFree (m->stage_bufs[k]); m->stage_bufs[k] = nullptr; }
   m->peer_merged = m->merged = m->stage = nullptr;
   m->dst = -1; m->max_rows = 0; m->mask = 0; m->pending = false;
}

// (re)build the merge buffers for (mask, max_rows, dst).  Collective: the IPC handle of the destination's buffer is
// broadcast, every rank maps it, and the ranks agree on the transport (peer writes only if EVERY rank could map it).
static int merge_setup (marxb200_ctx *c, uint64_t write_mask, uint64_t max_rows, int dst)
{
   MxComm *m = c->comm;
   NcclApi *N = nccl_api ();
   CUDA_OK (cudaStreamSynchronize (c->stream));
   merge_release (c);
   m->stage_bytes = mxb_build_egress_plan (write_mask, max_rows, m->plan, m->which);
   if (m->plan.num_cols == 0) return fail ("marxb200_merge_events_begin: the write mask selects no column");
   for (int k = 0; k < 2; k++)
     {
        CUDA_OK (cudaMalloc (&m->stage_bufs[k], (size_t) m->stage_bytes + 256));
        if (m->ev_stage_free[k] == nullptr) CUDA_OK (cudaEventCreateWithFlags (&m->ev_stage_free[k], cudaEventDisableTiming));
     }
   m->stage_k = 0; m->stage = m->stage_bufs[0];
   uint64_t off = 0;
   for (int j = 0; j < m->plan.num_cols; j++)
     {
        m->merged_off[j] = off;
        off += (((uint64_t) max_rows * m->world * kMxEgressCols[m->which[j]].size) + 255) & ~(uint64_t) 255;
     }
   m->merged_bytes = off;
   if (m->rank == dst) CUDA_OK (cudaMalloc (&m->merged, (size_t) off + 256));
   // peer mapping of the destination's buffer
   cudaIpcMemHandle_t handle;
   memset (&handle, 0, sizeof (handle));
   int ok = 1;
   const char *force = getenv ("MARXB200_MERGE_TRANSPORT");
   if ((force != nullptr) && (0 == strcmp (force, "nccl"))) ok = 0;
   if (m->world == 1) ok = 0;
   if ((m->rank == dst) && ok && (cudaSuccess != cudaIpcGetMemHandle (&handle, m->merged))) { ok = 0; cudaGetLastError (); }
   static_assert (sizeof (cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
   void *d_handle = (void *) (m->d_flag + 16);         // 64 bytes of the scratch
   CUDA_OK (cudaMemcpyAsync (d_handle, &handle, sizeof (handle), cudaMemcpyHostToDevice, m->merge_stream));
   NCCL_OK (N->Broadcast (d_handle, d_handle, sizeof (handle), NCCL_CHAR, dst, m->comm_merge, m->merge_stream));
   CUDA_OK (cudaMemcpyAsync (&handle, d_handle, sizeof (handle), cudaMemcpyDeviceToHost, m->merge_stream));
   CUDA_OK (cudaStreamSynchronize (m->merge_stream));
   if ((m->rank != dst) && ok)
     {
        bool zero = true;
        for (size_t k = 0; k < sizeof (handle); k++) if (((const unsigned char *) &handle)[k] != 0) zero = false;
        if (zero || (cudaSuccess != cudaIpcOpenMemHandle (&m->peer_merged, handle, cudaIpcMemLazyEnablePeerAccess)))
          { ok = 0; m->peer_merged = nullptr; cudaGetLastError (); }
     }
   CUDA_OK (cudaMemcpyAsync (m->d_flag, &ok, sizeof (int), cudaMemcpyHostToDevice, m->merge_stream));
   NCCL_OK (N->AllReduce (m->d_flag, m->d_flag, 1, NCCL_INT32, NCCL_MIN, m->comm_merge, m->merge_stream));
   int all_ok = 0;
   CUDA_OK (cudaMemcpyAsync (&all_ok, m->d_flag, sizeof (int), cudaMemcpyDeviceToHost, m->merge_stream));
   CUDA_OK (cudaStreamSynchronize (m->merge_stream));
   m->transport = all_ok ? MERGE_PEER : MERGE_NCCL;
   if (!all_ok && m->peer_merged) { cudaIpcCloseMemHandle (m->peer_merged); m->peer_merged = nullptr; }
   if (getenv ("MARXB200_VERBOSE"))
     fprintf (stderr, "marxb200: rank %d merge buffers: %d columns, %llu rows per rank, dst %d, transport %s\n", m->rank, m->plan.num_cols,
              (unsigned long long) max_rows, dst, all_ok ? "peer writes (CUDA IPC + copy engines)" : "ncclSend/ncclRecv");
   m->dst = dst; m->max_rows = max_rows; m->mask = write_mask;
   return 0;
}

extern "C" int marxb200_merge_events_begin (marxb200_ctx *c, uint64_t write_mask, double total_time, uint64_t max_rows_per_rank, int dst_rank)
{
   if (c == nullptr) return fail ("NULL ctx");
   MxComm *m = c->comm;
   if (m == nullptr) return fail ("marxb200_merge_events_begin: no communicator (marxb200_comm_init)");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   if (c->stage_done < 0) return fail ("marxb200_merge_events_begin: no photons");
   if ((dst_rank < 0) || (dst_rank >= m->world)) return fail ("marxb200_merge_events_begin: destination rank %d of %d", dst_rank, m->world);
   if (max_rows_per_rank == 0) return fail ("marxb200_merge_events_begin: max_rows_per_rank must be > 0");
   if (m->pending) return fail ("marxb200_merge_events_begin: the previous merge was not ended");
   CUDA_OK (cudaSetDevice (c->device));
   if ((m->stage == nullptr) || (m->mask != write_mask) || (m->max_rows != max_rows_per_rank) || (m->dst != dst_rank))
     if (-1 == merge_setup (c, write_mask, max_rows_per_rank, dst_rank)) return -1;
   if (-1 == mxb_ensure_order (c)) return -1;
   // The count is snapshot in stream order (the next batch clears the context's counters); the conversion to file images then runs on
   // the MERGE stream -- behind the previous merge's transfers, which read the same staging area -- while the context's stream goes
   // straight on with the next batch (its first kernel that writes into the list buffer being read waits, mxb_guard_buffer).
   // TIME of the merged list = (float) (absolute arrival time + total_time): in a sharded run the list's times already count
   // from the start of the simulation on every rank
   CUDA_OK (cudaStreamWaitEvent (c->stream, m->ev_packed, 0));         // the previous conversion has consumed the last snapshot
   CUDA_OK (cudaMemcpyAsync (m->d_all_counts + 64, c->d_counts + c->stage_done, sizeof (unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
   const int k = m->stage_k ^ 1;
   PackArgs want;
   memset (&want, 0, sizeof (want));
   want.plan = m->plan; want.dst = (unsigned char *) m->stage_bufs[k]; want.dev_start_time = nullptr; want.total_time = total_time; want.max_rows = m->max_rows;
   if (-1 == mxb_reader_begin (c, 1, m->merge_stream)) return -1;            // the merge stream continues behind the list (and the snapshot)
   if (!mxb_prepack_matches (c, 1, want))
     {
        launch_egress_pack (mxb_observed (c, c->buf[c->cur]), m->d_all_counts + 64, m->max_rows, m->plan, m->stage_bufs[k], nullptr, total_time, m->merge_stream);
        c->launches += 1;
        CUDA_OK (cudaGetLastError ());
        if (-1 == mxb_reader_end (c, 1, m->merge_stream)) return -1;
     }
   // else: the order restoration at the end of the trace wrote these very images into stage_bufs[k] (pre-pack): nothing to convert
   CUDA_OK (cudaEventRecord (m->ev_packed, m->merge_stream));
   m->stage_k = k; m->stage = m->stage_bufs[k];
   // a run that merges every batch the same way: let the next batch's order restoration write the images into the other buffer
   want.dst = (unsigned char *) m->stage_bufs[k ^ 1];
   mxb_prepack_arm (c, 1, want, m->ev_stage_free[k ^ 1]);
   // counts of all ranks, on the merge stream (the context's stream goes on with the next batch)
   NCCL_OK (N->AllGather (m->d_all_counts + 64, m->d_all_counts, 1, NCCL_UINT64, m->comm_merge, m->merge_stream));
   CUDA_OK (cudaMemcpyAsync (m->h_all_counts, m->d_all_counts, m->world * sizeof (unsigned long long), cudaMemcpyDeviceToHost, m->merge_stream));
   CUDA_OK (cudaEventRecord (m->ev_counts, m->merge_stream));
   m->pending = true;
   return 0;
}

extern "C" int marxb200_merge_events_end (marxb200_ctx *c, marxb200_merged_layout *layout)
{
   if (c == nullptr) return fail ("NULL ctx");
   MxComm *m = c->comm;
   if ((m == nullptr) || !m->pending) return fail ("marxb200_merge_events_end: no merge in flight");
   NcclApi *N = nccl_api ();
   if (N == nullptr) return no_nccl ();
   CUDA_OK (cudaSetDevice (c->device));
   m->pending = false;
   CUDA_OK (cudaEventSynchronize (m->ev_counts));
   uint64_t row0[65];
   row0[0] = 0;
   bool overflow = false;
   for (int r = 0; r < m->world; r++)
     {
        m->counts[r] = m->h_all_counts[r];
        if (m->counts[r] > m->max_rows) overflow = true;          // every rank sees the same counts: the same verdict everywhere
        row0[r + 1] = row0[r] + m->counts[r];
     }
   if (overflow)
     {
        CUDA_OK (cudaEventRecord (m->ev_pushed, m->merge_stream));
        return fail ("marxb200_merge_events_end: a rank holds more than max_rows_per_rank = %llu events", (unsigned long long) m->max_rows);
     }
   const uint64_t n_mine = m->counts[m->rank];
   const int dst = m->dst;
   CUDA_OK (cudaEventRecord (m->ev_t0, m->merge_stream));
   unsigned char *target = (unsigned char *) ((m->rank == dst) ? m->merged : m->peer_merged);
   if (m->transport == MERGE_PEER)
     {
        // every rank writes its rows of every column into the destination's buffer: copy engines over NVLink, no SM involved
        for (int j = 0; (j < m->plan.num_cols) && (n_mine > 0); j++)
          {
             const uint64_t sz = (uint64_t) kMxEgressCols[m->which[j]].size;
             CUDA_OK (cudaMemcpyAsync (target + m->merged_off[j] + row0[m->rank] * sz, (const unsigned char *) m->stage + m->plan.offset[j],
                                       (size_t) (n_mine * sz), cudaMemcpyDefault, m->merge_stream));
          }
        CUDA_OK (cudaEventRecord (m->ev_tc, m->merge_stream));
        // stream-ordered barrier: the destination's all-reduce completes only after every rank's, which follow their copies
        NCCL_OK (N->AllReduce (m->d_flag + 1, m->d_flag + 1, 1, NCCL_INT32, NCCL_MIN, m->comm_merge, m->merge_stream));
     }
   else
     {
        NCCL_OK (N->GroupStart ());
        for (int j = 0; j < m->plan.num_cols; j++)
          {
             const uint64_t sz = (uint64_t) kMxEgressCols[m->which[j]].size;
             if (m->rank != dst)
               {
                  if (n_mine > 0)
                    { const int r_ = N->Send ((const unsigned char *) m->stage + m->plan.offset[j], (size_t) (n_mine * sz), NCCL_CHAR, dst, m->comm_merge, m->merge_stream);
                      if (r_ != 0) { N->GroupEnd (); return fail ("ncclSend: %s", N->GetErrorString (r_)); } }
               }
             else
               for (int r = 0; r < m->world; r++)
                 {
                    if ((r == dst) || (m->counts[r] == 0)) continue;
                    const int r_ = N->Recv (target + m->merged_off[j] + row0[r] * sz, (size_t) (m->counts[r] * sz), NCCL_CHAR, r, m->comm_merge, m->merge_stream);
                    if (r_ != 0) { N->GroupEnd (); return fail ("ncclRecv: %s", N->GetErrorString (r_)); }
                 }
          }
        NCCL_OK (N->GroupEnd ());
        if (m->rank == dst)
          for (int j = 0; (j < m->plan.num_cols) && (n_mine > 0); j++)
            {
               const uint64_t sz = (uint64_t) kMxEgressCols[m->which[j]].size;
               CUDA_OK (cudaMemcpyAsync (target + m->merged_off[j] + row0[dst] * sz, (const unsigned char *) m->stage + m->plan.offset[j],
                                         (size_t) (n_mine * sz), cudaMemcpyDeviceToDevice, m->merge_stream));
            }
        CUDA_OK (cudaEventRecord (m->ev_tc, m->merge_stream));
     }
   CUDA_OK (cudaEventRecord (m->ev_t1, m->merge_stream));
   CUDA_OK (cudaEventRecord (m->ev_pushed, m->merge_stream));
   CUDA_OK (cudaEventRecord (m->ev_stage_free[m->stage_k], m->merge_stream));
   // this rank's part is done when its transfers are; the next batch's kernels are already queued on the context's stream
   CUDA_OK (cudaStreamSynchronize (m->merge_stream));
   CUDA_OK (cudaGetLastError ());
   CUDA_OK (cudaEventElapsedTime (&m->last_ms, m->ev_t0, m->ev_t1));
   CUDA_OK (cudaEventElapsedTime (&m->last_copy_ms, m->ev_t0, m->ev_tc));
   m->total_rows = row0[m->world];
   uint64_t row_bytes = 0;
   for (int j = 0; j < m->plan.num_cols; j++) row_bytes += (uint64_t) kMxEgressCols[m->which[j]].size;
   // bytes that crossed NVLink for this rank: received by the destination, sent by the others
   m->last_bytes = ((m->rank == dst) ? (m->total_rows - n_mine) : n_mine) * row_bytes;
   if (layout != nullptr)
     {
        memset (layout, 0, sizeof (*layout));
        layout->num_cols = (uint32_t) m->plan.num_cols;
        layout->n_rows = m->total_rows;
        layout->world = (uint32_t) m->world; layout->dst_rank = (uint32_t) dst;
        layout->transport = (uint32_t) m->transport;
        for (int r = 0; r < m->world; r++) layout->rows_of_rank[r] = m->counts[r];
        layout->device_base = (m->rank == dst) ? m->merged : nullptr;
        for (int j = 0; j < m->plan.num_cols; j++)
          {
             const MxEgressCol &col = kMxEgressCols[m->which[j]];
             layout->mask[j] = col.mask; layout->type[j] = col.type; layout->elem_size[j] = (uint32_t) col.size;
             strncpy (layout->file[j], col.file, sizeof (layout->file[j]) - 1);
             layout->device_offset[j] = m->merged_off[j];
          }
        layout->transfer_ms = m->last_ms; layout->copy_ms = m->last_copy_ms; layout->nvlink_bytes = m->last_bytes;
     }
   return 0;
}

// destination rank: the merged columns to the host, packed like marxb200_egress_end_packed lays them out
extern "C" int marxb200_merge_download (marxb200_ctx *c, void *host, uint64_t host_bytes, marxb200_packed_layout *layout)
{
   if ((c == nullptr) || (host == nullptr) || (layout == nullptr)) return fail ("marxb200_merge_download: NULL argument");
   MxComm *m = c->comm;
   if ((m == nullptr) || (m->merged == nullptr) || (m->rank != m->dst)) return fail ("marxb200_merge_download: this rank holds no merged list");
   if (m->pending) return fail ("marxb200_merge_download: a merge is in flight (marxb200_merge_events_end first)");
   CUDA_OK (cudaSetDevice (c->device));
   memset (layout, 0, sizeof (*layout));
   layout->num_cols = (uint32_t) m->plan.num_cols;
   layout->n_rows = m->total_rows;
   uint64_t off = 0;
   for (int j = 0; j < m->plan.num_cols; j++)
     {
        const MxEgressCol &col = kMxEgressCols[m->which[j]];
        layout->mask[j] = col.mask; layout->type[j] = col.type; layout->elem_size[j] = (uint32_t) col.size;
        strncpy (layout->file[j], col.file, sizeof (layout->file[j]) - 1);
        layout->offset[j] = off;
        off += ((m->total_rows * (uint64_t) col.size) + 15) & ~(uint64_t) 15;
     }
   if (off > host_bytes) return fail ("marxb200_merge_download: the host buffer holds %llu bytes, %llu are needed", (unsigned long long) host_bytes, (unsigned long long) off);
   for (int j = 0; (j < m->plan.num_cols) && (m->total_rows > 0); j++)
     CUDA_OK (cudaMemcpyAsync ((unsigned char *) host + layout->offset[j], (const unsigned char *) m->merged + m->merged_off[j],
                               (size_t) (m->total_rows * layout->elem_size[j]), cudaMemcpyDeviceToHost, m->merge_stream));
   CUDA_OK (cudaStreamSynchronize (m->merge_stream));
   return 0;
}

// ---------------------------------------------------------------------------------------------
// Host-copy ceiling probe: `reps` device-to-host copies of `bytes` into a pinned buffer the library allocates with
// cudaHostAlloc, timed with CUDA events on a private stream.  flags & 2 (collective, needs a communicator): the ranks start
// together (all-reduce as a barrier), so that the call measures the box's CONCURRENT D2H ceiling -- the bound of the
// end-to-end event egress at 8 GPUs (bench.py "d2h_ceiling").  flags & 1: write-combined host memory.
// ---------------------------------------------------------------------------------------------
extern "C" int marxb200_probe_d2h (marxb200_ctx *c, uint64_t bytes, int reps, int flags, double *gb_per_s)
{
   if ((c == nullptr) || (gb_per_s == nullptr)) return fail ("marxb200_probe_d2h: NULL argument");
   if ((bytes == 0) || (reps < 1)) return fail ("marxb200_probe_d2h: bytes and reps must be positive");
   *gb_per_s = 0.0;
   CUDA_OK (cudaSetDevice (c->device));
   void *d = nullptr, *h = nullptr;
   cudaStream_t s = nullptr;
   cudaEvent_t e0 = nullptr, e1 = nullptr;
   int status = 0;
   do
     {
#define P_OK(expr) { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { status = fail ("%s: %s", #expr, cudaGetErrorString (e_)); break; } }
        P_OK (cudaMalloc (&d, (size_t) bytes));
        P_OK (cudaHostAlloc (&h, (size_t) bytes, (flags & 1) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
        memset (h, 0, (size_t) bytes);                         // first touch by the calling (NUMA-bound) thread
        P_OK (cudaMemset (d, 1, (size_t) bytes));
        P_OK (cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking));
        P_OK (cudaEventCreate (&e0)); P_OK (cudaEventCreate (&e1));
        P_OK (cudaMemcpyAsync (h, d, (size_t) bytes, cudaMemcpyDeviceToHost, s));        // warm-up
        if ((c->comm != nullptr) && (flags & 2))
          {
             NcclApi *N = nccl_api ();
             if (N) N->AllReduce (c->comm->d_flag + 2, c->comm->d_flag + 2, 1, NCCL_INT32, NCCL_MIN, c->comm->comm_merge, s);
          }
        P_OK (cudaEventRecord (e0, s));
        cudaError_t ec = cudaSuccess;
        for (int k = 0; (k < reps) && (ec == cudaSuccess); k++) ec = cudaMemcpyAsync (h, d, (size_t) bytes, cudaMemcpyDeviceToHost, s);
        P_OK (ec);
        P_OK (cudaEventRecord (e1, s));
        P_OK (cudaStreamSynchronize (s));
        float ms = 0.f;
        P_OK (cudaEventElapsedTime (&ms, e0, e1));
        *gb_per_s = (double) bytes * reps / (ms * 1e-3) * 1e-9;
#undef P_OK
     }
   while (0);
   if (e0) cudaEventDestroy (e0);
   if (e1) cudaEventDestroy (e1);
   if (s) cudaStreamDestroy (s);
   if (h) cudaFreeHost (h);
   if (d) cudaFree (d);
   return status;
}
