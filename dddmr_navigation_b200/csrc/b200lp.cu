// b200lp.cu — C ABI (include/b200lp.h) over the sm_100a kernels in lp_kernels.cuh.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC,-pthread,-ffp-contract=off
// -fmad=false is part of the numeric contract (lp_device.cuh); -ffp-contract=off keeps the host code that plans a
// single robot's velocity samples (plan_samples) on the arithmetic the kernels use. There is no CPU path in this file.
// -DB200LP_COUNT=1 / -DB200LP_CHECKS=1 give the counting and the checking build (never timed).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost a null-pointer test unless a profiler injects its library

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200lp.h"
#include "lp_hostpack.h"
#include "lp_kernels.cuh"
#include "lp_observe.cuh"

using namespace lp;

namespace {

thread_local std::string g_create_error;

// One NVTX range per C-ABI call that reaches the device (SURVEY.md §5 "Tracing"): nsys / ncu --nvtx show the upload, the grid
// build, the plan cycle and the observation producer as named spans on the calling thread.
struct TraceRange {
  explicit TraceRange(const char* name) { nvtxRangePushA(name); }
  ~TraceRange() { nvtxRangePop(); }
  TraceRange(const TraceRange&) = delete;
  TraceRange& operator=(const TraceRange&) = delete;
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max(n, (size_t)16);
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max(n, (size_t)16);
    cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct b200lp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t cev[3] = {nullptr, nullptr, nullptr};  // set_cloud: start, cloud arrived, grid built
  bool cloud_timing_pending = false;
  std::string err;
  Consts C{};
  b200lp_grid_config gcfg{};
  int64_t launches = 0;

  // cloud / grid
  GridDev grid{};
  size_t raw_stride = 0;
  size_t last_h2d_bytes = 0;         // what the last set_cloud* copied host -> device
  DevBuf<char> d_raw;
  DevBuf<float4> d_pts;
  DevBuf<uint32_t> d_cell_start, d_rank, d_sat;  // d_rank: every point's rank inside its cell (hist_kernel -> scatter_kernel)
  DevBuf<unsigned long long> d_scan_status;      // scan_kernel: per block {epoch, status, value}
  DevBuf<unsigned> d_scan_ticket;
  unsigned scan_epoch = 0;
  DevBuf<float4> d_packed;  // the raw cloud as 16-byte records (x, y, z, original index)
  cudaStream_t copy_stream = nullptr;
  cudaStream_t prep_stream = nullptr;  // prep_kernel runs here while the grid of a cloud just handed over is still being built
  bool cycle_overlapped = false, have_cycle_event = false;
  PackPool* pack_pool = nullptr;     // host threads of the packing upload (created by the first large host cloud)
  PinBuf<float> h_stage;             // pinned staging buffer: x,y,z of every point, 12 bytes each
  int pack_threads_used = 0;         // threads that packed the last cloud (0: it was copied as is)
  cudaEvent_t chunk_ev[16] = {};
  DevBuf<BoundsDev> d_bounds;
  DevBuf<uint32_t> d_total;
  PinBuf<BoundsDev> h_bounds;
  bool have_cloud = false;
  float cell_xy_used = 0.f, cell_z_used = 0.f;

  // plan (single-robot path) — host copy, uploaded with every query
  std::vector<double> plan_host;
  // ... unless b200lp_prune_plan produced it on the device (then d_plan7 already holds plan_n_device poses)
  bool plan_on_device = false;
  int plan_n_device = 0;
  DevBuf<double> d_gplan7;  // Local_Planner::global_plan_
  size_t n_gplan = 0;
  DevBuf<float4> d_prune_pcl;  // pcl_prune_plan_ (x, y, z, intensity tag)
  DevBuf<PruneMeta> d_prune_meta;
  PinBuf<PruneMeta> h_prune_meta;
  DevBuf<int> d_blocked;
  PinBuf<int> h_blocked;
  bool have_prune = false;

  // observation producer
  DevBuf<char> d_scan;                          // the uploaded lidar scan
  DevBuf<float4> d_obs_a, d_obs_b;              // sort ping-pong: (x, y, z in base_link, voxel key)
  DevBuf<uint32_t> d_obs_hist, d_obs_sums, d_obs_heads, d_obs_counts;
  PinBuf<uint32_t> h_obs_counts;
  DevBuf<float4> d_obs_out[B200LP_MAX_SENSORS];  // Sensor::sensor_current_observation_ per sensor, pcl::PointXYZ layout
  size_t n_obs_out[B200LP_MAX_SENSORS] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool have_obs[B200LP_MAX_SENSORS] = {false, false, false, false, false, false, false, false};
  cudaEvent_t oev[3] = {nullptr, nullptr, nullptr};  // observation: start, end, scan uploaded

  // peer-memory argmin exchange of sample-sharded cycles
  DevBuf<char> d_peer_block;              // [2 x kMaxPeers slots the peers write into | CloudHeader | shared-cloud rows]
  size_t peer_cloud_cap = 0;              // points the shared-cloud row buffer holds (b200lp_peer_reserve_cloud)
  PeerTable peer_table{};                 // every rank's slot array, mapped into this process
  PeerBlocks peer_blocks{};               // ... and the base of every rank's block
  unsigned long long cloud_seq = 0;       // shared-cloud cycles so far
  cudaStream_t push_stream = nullptr;     // root: pushes of the upload pieces to the peers
  cudaEvent_t push_done = nullptr;        // ... as of the last piece (the next upload may overwrite d_raw after it)
  bool push_pending = false;
  DevBuf<unsigned> d_push_ticket;
  DevBuf<int> d_share_err;
  PinBuf<CloudHeader> h_share_hdr;        // peer: the root's header, written by share_wait_header_kernel
  PinBuf<int> h_share_err;
  PeerSlot* peer_slots() const { return reinterpret_cast<PeerSlot*>(d_peer_block.p); }
  CloudHeader* my_header() const { return reinterpret_cast<CloudHeader*>(d_peer_block.p + kPeerHeaderOff); }
  char* my_rows() const { return d_peer_block.p + kPeerRowsOff; }
  void* peer_opened[kMaxPeers] = {nullptr};
  int peer_rank = -1, peer_world = 0;
  unsigned long long peer_seq = 0;

  // per-cycle state
  size_t n_robots = 0;
  int t_cap = 0;
  int shard_rank = 0, shard_count = 1;
  float axes_host[3][kMaxAxis];           // plan_samples: the VelocityIterator chains of the window below
  float axes_window[6] = {0, 0, 0, 0, 0, 0};
  int axes_n[3] = {0, 0, 0};
  bool axes_valid = false;
  float row_work[kMaxAxis];               // plan_samples: inclusive prefix of the estimated work per linear-speed row (shard cuts)
  bool row_work_valid = false;
  const double* plan_src = nullptr;       // fleet calls whose plan table already sits in pinned host memory: uploaded from there
  PrepPlan prep_plan{};                   // host-planned sample layout of the single-robot launch being issued (plan_samples)
  unsigned long long sample_cuts_hash = 0; // hash of the sample cuts that layout used: all ranks of an exchange cycle must agree
  ShardCuts cuts{};                       // where sample-sharded launches cut the estimated-work axis (n == 0: equal shares)
  DevBuf<RobotIn> d_robots;
  DevBuf<RobotMeta> d_meta;
  DevBuf<double> d_plan7;
  DevBuf<float4> d_plan_pts;
  DevBuf<float4> d_rec_vel;
  DevBuf<int> d_rec_steps, d_rec_sample, d_first_hit;
  DevBuf<double> d_rec_dt, d_cost, d_scores;
  DevBuf<long long> d_rec_pose_off;      // row of every trajectory in d_poses
  DevBuf<float4> d_poses;                // forward-simulated (x, y, th) of every pose of the cycle, robot frame
  DevBuf<double2> d_rec_pp;              // pure-pursuit (distance, yaw) of every trajectory's last pose
  long long pose_stride = 0;             // rows of d_poses reserved per robot
  DevBuf<BlockBest> d_partial;           // argmin_kernel CTA partials
  DevBuf<unsigned> d_tickets2;           // argmin_kernel last-CTA tickets, one per robot (self-resetting)
  DevBuf<unsigned> d_tickets;            // prep_kernel chunk tickets, one per robot (self-resetting)
  DevBuf<PrepAgg> d_aggs;                // prep_kernel look-back aggregates
  DevBuf<unsigned long long> d_work;     // plan_kernel work counter (reset by argmin_kernel)
  DevBuf<uint32_t> d_surv;               // cull_kernel: per trajectory, the mask of the poses that survive the float pre-cull
  DevBuf<unsigned> d_hist;               // SM clocks every velocity sample's trajectory took in the previous cycle (0: unknown)
  DevBuf<int> d_order;                   // ... and the work lists of the cycle, one per cost class (kCostClasses x T entries)
  DevBuf<unsigned> d_class_counts;       // entries per list (zeroed by prep_kernel)
  DevBuf<unsigned long long> d_tstart;   // globaltimer at the start of the cycle (prep_kernel's first CTA)
  bool plan_uploaded = false;            // d_plan7 already holds the host plan (b200lp_set_plan uploads it)
  cudaEvent_t fleet_fork_ev = nullptr, fleet_plan_ev = nullptr;  // fleet calls: the plan table's upload beside prep_kernel
  cudaEvent_t plan_ev = nullptr;         // ... as of this event on the main stream
  bool plan_ev_pending = false;          // a prep kernel on the second stream has not yet been ordered behind it
  bool adaptive_cuts = true;             // sample-sharded exchange cycles move their cuts with the ranks' device times
  uint32_t last_cycle_ns = 0, last_peer_ns[kMaxPeers] = {0};
  unsigned epoch = 0;                    // launch number, the "published" flag value of d_aggs
  int plan_ctas_per_sm = 0, sm_count = 0;
  DevBuf<b200lp_result> d_results;
  DevBuf<unsigned long long> d_count;
  DevBuf<char> d_scratch;
  PinBuf<RobotIn> h_robots;
  PinBuf<double> h_plan7;
  PinBuf<b200lp_result> h_results;
  PinBuf<RobotMeta> h_meta;
  PinBuf<DirectOut> h_direct;            // single-robot cycles: result block the kernel writes into host memory
  PinBuf<unsigned long long> h_fleet_seq; // fleet cycles: raised by argmin_kernel when every robot's result is in h_results / h_meta
  DevBuf<unsigned> d_fleet_done;         // robots whose result has been published (self-resetting)
  unsigned long long fleet_seq = 0;
  unsigned long long direct_seq = 0;
  bool cycle_timing_pending = false, cycle_timing_direct = false;
  PinBuf<unsigned long long> h_count;
  std::vector<RobotMeta> meta_host;
  bool have_cycle = false;

  // timing of the last call
  float ms_upload = 0.f, ms_grid = 0.f, ms_plan = 0.f, ms_readback = 0.f;
  float ms_k_prep = 0.f, ms_k_cull = 0.f, ms_k_plan = 0.f, ms_k_argmin = 0.f;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return ctx->fail(B200LP_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

namespace {

// B200LP_HOST_TRACE=1 (tools only): host clock at the stations of a single-robot cycle, printed to stderr
inline long long host_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

int grid_blocks(size_t n, int threads, int sm_count) {
  const size_t want = (n + threads - 1) / threads;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sm_count * 16));
}

int sm_count_of(int device) {
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

// largest num_steps a parameter set can produce (see traj_precheck)
double max_steps_bound(const b200lp_limits& L, const b200lp_params& P) {
  const double eps = 1e-4;
  if (P.theory == B200LP_THEORY_DD_ROTATE_INPLACE) return std::ceil(6.28 / P.angular_sim_granularity) + 1;
  double vmax;
  if (P.theory == B200LP_THEORY_DD_SIMPLE) vmax = L.max_vel_x;
  else vmax = L.max_vel_trans;
  if (vmax < 0) return INFINITY;  // the reference disables the speed check; unbounded
  const double a = (vmax + eps) * P.sim_time / P.sim_granularity;
  const double b = (L.max_vel_theta + eps) * P.sim_time / P.angular_sim_granularity;
  return std::ceil(std::max(a, b)) + 1;
}

int axis_count(double sample) { return std::max(2, (int)sample) + 1; }

int traj_cap(const b200lp_params& P) {
  if (!(P.linear_x_sample * P.angular_z_sample > 0)) return 1;
  if (P.theory == B200LP_THEORY_DD_ROTATE_INPLACE) return 2;
  long long n = (long long)axis_count(P.linear_x_sample) * axis_count(P.angular_z_sample);
  if (P.theory == B200LP_THEORY_OMNI_SIMPLE) n *= axis_count(P.linear_y_sample);
  return (int)std::min<long long>(n, 1ll << 30);
}

#ifndef B200LP_UPLOAD_CHUNKS
#define B200LP_UPLOAD_CHUNKS 4
#endif
constexpr int kUploadChunks = B200LP_UPLOAD_CHUNKS;  // <= 8 (chunk_ev)

// Grid geometry from the bounds of the finite points: cell sizes (coarsened until the dense grid fits max_cells), dims,
// origin. Returns the number of cells.
size_t size_grid(b200lp_ctx* ctx, const float mn[3], const float mx[3], size_t n_finite) {
  GridDev& g = ctx->grid;
  float cxy = ctx->gcfg.cell_xy > 0.f ? ctx->gcfg.cell_xy : 0.2f;
  float cz = ctx->gcfg.cell_z > 0.f ? ctx->gcfg.cell_z : 0.4f;
  const uint32_t max_cells = ctx->gcfg.max_cells ? ctx->gcfg.max_cells : (1u << 26);
  size_t n_cells = 1;
  if (n_finite) {
    for (;;) {
      const double nx = std::floor(((double)mx[0] - mn[0]) / cxy) + 1, ny = std::floor(((double)mx[1] - mn[1]) / cxy) + 1,
                   nz = std::floor(((double)mx[2] - mn[2]) / cz) + 1;
      if (nx * ny * nz <= (double)max_cells && nx < 2e9 && ny < 2e9 && nz < 2e9) {
        g.nx = (int)nx;
        g.ny = (int)ny;
        g.nz = (int)nz;
        break;
      }
      cxy *= 1.25f;
      cz *= 1.25f;
    }
    g.org[0] = mn[0];
    g.org[1] = mn[1];
    g.org[2] = mn[2];
    g.inv_xy = 1.0f / cxy;
    g.inv_z = 1.0f / cz;
    float cm = 1.f;
    for (int a = 0; a < 3; ++a) cm = std::max(cm, std::max(std::fabs(mn[a]), std::fabs(mx[a])));
    g.cmax = cm;
    g.n_kept = (uint32_t)n_finite;
    n_cells = (size_t)g.nx * g.ny * g.nz;
  }
  ctx->cell_xy_used = cxy;
  ctx->cell_z_used = cz;
  return n_cells;
}

// Everything of the grid build behind the bounds: tables, histogram (per upload piece when `piece_bound` names the pieces,
// each one behind its copy event — or, for a rank that RECEIVES the cloud from a peer, behind a kernel that waits for the
// piece's flag in `wait_hdr`), scan, scatter, summed-volume table.
int grid_tail(b200lp_ctx* ctx, const char* rec, size_t rec_stride, size_t n, size_t n_cells, const size_t* piece_bound,
              const CloudHeader* wait_hdr, unsigned long long wait_seq) {
  const int sms = sm_count_of(ctx->device);
  GridDev& g = ctx->grid;
  const int nb = (int)((n_cells + kScanItems - 1) / kScanItems);
  CK(ctx->d_cell_start.reserve(n_cells + 1));
  CK(ctx->d_rank.reserve(std::max<size_t>(n, 1)));
  if (ctx->d_scan_status.cap < (size_t)nb) {  // (new words carry epoch 0: never the current one)
    CK(ctx->d_scan_status.reserve(nb));
    CK(cudaMemsetAsync(ctx->d_scan_status.p, 0, ctx->d_scan_status.cap * sizeof(unsigned long long), ctx->stream));
  }
  if (!ctx->d_scan_ticket.p) {
    CK(ctx->d_scan_ticket.reserve(1));
    CK(cudaMemsetAsync(ctx->d_scan_ticket.p, 0, sizeof(unsigned), ctx->stream));
  }
  if (++ctx->scan_epoch >= (1u << 30)) ctx->scan_epoch = 1u;
  CK(ctx->d_pts.reserve(std::max<size_t>(g.n_kept, 1)));
  const size_t n_sat = ((size_t)g.nx + 1) * ((size_t)g.ny + 1) * ((size_t)g.nz + 1);
  CK(ctx->d_sat.reserve(n_sat));
  CK(cudaMemsetAsync(ctx->d_cell_start.p, 0, (n_cells + 1) * sizeof(uint32_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_sat.p, 0, n_sat * sizeof(uint32_t), ctx->stream));
  g.pts = ctx->d_pts.p;
  g.cell_start = ctx->d_cell_start.p;
  g.sat = ctx->d_sat.p;
  // scatter_kernel writes 16-byte records to random places: half a 32-byte sector each, i.e. a fetch of the other half
  // from DRAM when the sector is not in L2. Clearing the output first — while the cloud is still on its way, for the
  // uploads that come in pieces — leaves its sectors resident and dirty in the 126 MB L2, so the records land there:
  // grid tail at 2 M points 0.070 -> 0.052 ms, upload unchanged (B200LP_PREZERO_PTS=0 switches it off; gpurun r4v / r4w).
  // Not for outputs that do not fit L2 beside the rows they are sorted from (8 M points: 128 MB).
  static const bool prezero = [] { const char* e = getenv("B200LP_PREZERO_PTS"); return !e || atoi(e) != 0; }();
  if (prezero && g.n_kept && (size_t)g.n_kept * sizeof(float4) <= (size_t)64 << 20)
    CK(cudaMemsetAsync(ctx->d_pts.p, 0, (size_t)g.n_kept * sizeof(float4), ctx->stream));
  if (piece_bound) {  // every piece is counted as soon as it has landed; only the last one is not hidden by the upload
    for (int c = 0; c < kPackChunks; ++c) {
      const size_t i0 = piece_bound[c], i1 = piece_bound[c + 1];
      if (wait_hdr) {
        share_wait_piece_kernel<<<1, 32, 0, ctx->stream>>>(wait_hdr, wait_seq, c, (long long)4e9, ctx->d_share_err.p);
        ++ctx->launches;
      } else if (i1 > i0) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
      }
      if (c == kPackChunks - 1) CK(cudaEventRecord(ctx->cev[1], ctx->stream));  // everything has arrived
      if (i1 > i0 && g.n_kept) {
        hist_kernel<<<grid_blocks(i1 - i0, 256, sms), 256, 0, ctx->stream>>>(rec, rec_stride, i0, i1, g, ctx->d_cell_start.p,
                                                                              ctx->d_rank.p);
        ++ctx->launches;
      }
    }
  } else if (g.n_kept) {
    hist_kernel<<<grid_blocks(n, 256, sms), 256, 0, ctx->stream>>>(rec, rec_stride, 0, n, g, ctx->d_cell_start.p, ctx->d_rank.p);
    ++ctx->launches;
  }
  if (g.n_kept) {
    // B200LP_TAIL_TRACE=1 (tools only): CUDA events between the kernels of the tail, printed after a synchronise
    static const bool tail_trace = getenv("B200LP_TAIL_TRACE") != nullptr;
    cudaEvent_t te[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    auto mark = [&](int k) {
      if (!tail_trace) return;
      cudaEventCreate(&te[k]);
      cudaEventRecord(te[k], ctx->stream);
    };
    mark(0);
    scan_kernel<<<nb, 256, 0, ctx->stream>>>(ctx->d_cell_start.p, n_cells, ctx->d_scan_status.p, ctx->d_scan_ticket.p,
                                             ctx->scan_epoch, ctx->d_total.p);
    mark(1);
    // (the summed-volume passes need the scan only, like the scatter; running them beside it on a second stream was
    // measured twice and changes nothing at 2 M points: 0.055-0.057 against 0.052 ms for the tail, gpurun r4w)
    scatter_kernel<<<grid_blocks(n, 256, sms), 256, 0, ctx->stream>>>(rec, rec_stride, n, g, ctx->d_rank.p, ctx->d_pts.p);
    mark(2);
    const size_t ny_threads = ((size_t)g.nx + 1) * (size_t)g.nz, nz_threads = ((size_t)g.nx + 1) * ((size_t)g.ny + 1);
    sat_y_kernel<<<(unsigned)((ny_threads * 32 + 255) / 256), 256, 0, ctx->stream>>>(g, ctx->d_sat.p);
    mark(3);
    sat_z_kernel<<<(unsigned)((nz_threads + 255) / 256), 256, 0, ctx->stream>>>(g, ctx->d_sat.p);
    mark(4);
    if (tail_trace) {
      cudaStreamSynchronize(ctx->stream);
      float ms[4] = {0, 0, 0, 0}, since_arrival = 0.f;
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], te[k], te[k + 1]);
      cudaEventElapsedTime(&since_arrival, ctx->cev[1], te[0]);
      fprintf(stderr, "grid tail: last piece arrived -> scan starts %.1f us, scan %.1f, scatter %.1f, sat_y %.1f, sat_z %.1f us\n",
              since_arrival * 1e3f, ms[0] * 1e3f, ms[1] * 1e3f, ms[2] * 1e3f, ms[3] * 1e3f);
      for (auto& e : te) cudaEventDestroy(e);
    }
    ctx->launches += 4;
  }
  CK(cudaGetLastError());
  ctx->have_cloud = true;
  return B200LP_OK;
}

// Brings the cloud to the device and builds the voxel grid + summed-volume table.
//   src == nullptr : the raw cloud is already in d_raw (or n == 0)
//   otherwise      : `src` (host, or device when on_device) is copied in pieces on the copy stream, overlapped with the
//                    first pass over the pieces that have arrived on the main stream. Two ways:
//     packing upload (large padded host clouds): host threads pack x,y,z into pinned staging AND reduce the bounds, so the
//       grid is sized on the host without a device round trip, hist_kernel runs per piece under the rest of the upload,
//       and the call returns when the caller's buffer has been read — nothing waits for the device;
//     plain upload: bounds_pack_kernel per piece (bounds + 16-byte records on the device), one host round trip for the
//       bounds, then the histogram over the whole cloud.
int build_grid(b200lp_ctx* ctx, const void* src, size_t n, size_t stride, bool on_device, int share_root = -1) {
  TraceRange trace_range("b200lp cloud upload + grid build");
  const int sms = sm_count_of(ctx->device);
  GridDev& g = ctx->grid;
  g.n_raw = (uint32_t)n;
  g.n_kept = 0;
  g.nx = g.ny = g.nz = 1;
  g.org[0] = g.org[1] = g.org[2] = 0.f;
  g.inv_xy = g.inv_z = 1.f;
  g.cmax = 1.f;
  CK(ctx->d_total.reserve(1));
  CK(cudaEventRecord(ctx->cev[0], ctx->stream));
  ctx->pack_threads_used = 0;
  // large host clouds with padding between the points: pack on the host, upload 12 bytes per point (see PackPool)
  bool packing = false;
  const bool share = share_root >= 0;  // this rank is the root of a shared cloud: always the packing upload (12-byte rows)
  if (src && !on_device && ((stride >= 16 && n * stride >= kPackMinBytes) || share)) {
    if (!ctx->pack_pool) {
      const int t = share ? std::max(1, pack_threads_wanted()) : pack_threads_wanted();
      if (t > 0) ctx->pack_pool = new (std::nothrow) PackPool(t);
      if (ctx->pack_pool && ctx->pack_pool->threads() == 0) {  // no thread could be started: plain copies from now on
        delete ctx->pack_pool;
        ctx->pack_pool = nullptr;
      }
    }
    // the staging buffer may still be feeding the copies of the previous cloud
    packing = ctx->pack_pool != nullptr && cudaStreamSynchronize(ctx->copy_stream) == cudaSuccess &&
              ctx->h_stage.reserve(n * 3) == cudaSuccess;
  }
  if (share && n && !packing) return ctx->fail(B200LP_E_NOMEM, "set_cloud_shared: the packing upload is not available (threads / pinned memory)");
  const char* rec = nullptr;  // what the histogram / scatter passes read, and its stride
  size_t rec_stride = 0;
  size_t n_cells = 1;
  bool hist_per_chunk = false;
  size_t chunk_bound[kPackChunks + 1] = {0};
  if (packing) {
    PackPool& pool = *ctx->pack_pool;
    static const bool up_trace = getenv("B200LP_HOST_TRACE") != nullptr;
    const long long up_t0 = up_trace ? host_ns() : 0;
    long long up_t1 = 0;
    pool.start((const char*)src, stride, ctx->h_stage.p, n);
    ctx->pack_threads_used = pool.threads();
    ctx->raw_stride = 12;  // what d_raw holds from here on
    cudaError_t pe = cudaStreamWaitEvent(ctx->copy_stream, ctx->cev[0], 0);  // the copies may not overtake earlier work on d_raw
    if (pe == cudaSuccess && ctx->push_pending) {  // ... nor the pushes of the previous shared cloud, which read d_raw
      pe = cudaStreamWaitEvent(ctx->copy_stream, ctx->push_done, 0);
      ctx->push_pending = false;
    }
    if (share && pe == cudaSuccess) {  // the peers must be through with the rows of the previous cycle
      share_wait_acks_kernel<<<1, 32, 0, ctx->push_stream>>>(ctx->my_header(), ctx->peer_world, ctx->peer_rank, ctx->cloud_seq - 1,
                                                             (long long)4e9, ctx->d_share_err.p);
      ++ctx->launches;
    }
    for (int c = 0; c < kPackChunks; ++c) {
      const size_t i0 = pool.bound(c), i1 = pool.bound(c + 1);
      chunk_bound[c] = i0;
      chunk_bound[c + 1] = i1;
      pool.wait_chunk(c);  // also on the error path: the workers read the caller's buffer until the last chunk is packed
      if (up_trace && c == 0) up_t1 = host_ns();
      if (up_trace && c == kPackChunks - 1)
        fprintf(stderr, "upload trace: %d threads, first piece (%zu points) packed %.1f us after start(), all %zu points %.1f us\n",
                pool.threads(), pool.bound(1), (up_t1 - up_t0) / 1e3, n, (host_ns() - up_t0) / 1e3);
      if (pe != cudaSuccess || i1 == i0) continue;
      pe = cudaMemcpyAsync(ctx->d_raw.p + i0 * 12, ctx->h_stage.p + i0 * 3, (i1 - i0) * 12, cudaMemcpyHostToDevice, ctx->copy_stream);
      if (pe == cudaSuccess) pe = cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream);
      if (share && pe == cudaSuccess) {  // piece c goes on to every peer as soon as it has landed here
        pe = cudaStreamWaitEvent(ctx->push_stream, ctx->chunk_ev[c], 0);
        const size_t b0 = i0 * 12, b1 = i1 * 12;  // (piece bounds are multiples of 4 points: b0 is 16-byte aligned)
        share_push_kernel<<<grid_blocks((b1 - b0) / 16 + 1, 256, sm_count_of(ctx->device)), 256, 0, ctx->push_stream>>>(
            ctx->d_raw.p, b0, b1, ctx->peer_blocks, ctx->peer_world, ctx->peer_rank, c, ctx->cloud_seq, ctx->d_push_ticket.p);
        ++ctx->launches;
      }
    }
    if (share && pe == cudaSuccess) {
      for (int c = 0; c < kPackChunks; ++c)  // empty pieces still raise their flag
        if (pool.bound(c + 1) == pool.bound(c)) {
          share_push_kernel<<<1, 256, 0, ctx->push_stream>>>(ctx->d_raw.p, 0, 0, ctx->peer_blocks, ctx->peer_world, ctx->peer_rank, c,
                                                             ctx->cloud_seq, ctx->d_push_ticket.p);
          ++ctx->launches;
        }
      pe = cudaEventRecord(ctx->push_done, ctx->push_stream);
      ctx->push_pending = true;
    }
    if (pe != cudaSuccess) return ctx->fail(B200LP_E_CUDA, "set_cloud (packing upload): %s", cudaGetErrorString(pe));
    // the caller's buffer is free from here on, and the bounds are already on the host
    const HostBounds hb = pool.bounds();
    if (share) {  // size, bounds and piece boundaries for the peers: a one-warp kernel on the idle prep stream
      CloudHeader H{};
      H.n = n;
      H.n_finite = hb.n_finite;
      for (int a = 0; a < 3; ++a) { H.mn[a] = hb.mn[a]; H.mx[a] = hb.mx[a]; }
      for (int c = 0; c <= kPackChunks; ++c) H.bound[c] = pool.bound(c);
      H.seq = ctx->cloud_seq;
      share_header_kernel<<<1, 32, 0, ctx->prep_stream>>>(H, ctx->peer_blocks, ctx->peer_world, ctx->peer_rank);
      ++ctx->launches;
    }
    n_cells = size_grid(ctx, hb.mn, hb.mx, hb.n_finite);
    rec = ctx->d_raw.p;
    rec_stride = 12;
    hist_per_chunk = true;
  } else {
    CK(ctx->d_bounds.reserve(1));
    CK(ctx->h_bounds.reserve(1));
    CK(ctx->d_packed.reserve(std::max<size_t>(n, 1)));
    BoundsDev init;
    for (int a = 0; a < 3; ++a) {
      init.mn[a] = 0xffffffffu;
      init.mx[a] = 0u;
    }
    init.n_finite = 0;
    init.pad = 0;
    *ctx->h_bounds.p = init;
    CK(cudaMemcpyAsync(ctx->d_bounds.p, ctx->h_bounds.p, sizeof(BoundsDev), cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
      const int chunks = (src && n >= (size_t)kUploadChunks * 65536) ? kUploadChunks : 1;
      // one piece: plain copy on the main stream; several: on the copy stream, each piece handed over by an event
      cudaStream_t cs = chunks > 1 ? ctx->copy_stream : ctx->stream;
      if (src && chunks > 1) CK(cudaStreamWaitEvent(cs, ctx->cev[0], 0));  // the copy may not overtake earlier work on d_raw
      for (int c = 0; c < chunks; ++c) {
        const size_t i0 = n * c / chunks, i1 = n * (c + 1) / chunks;
        if (src) {
          CK(cudaMemcpyAsync(ctx->d_raw.p + i0 * stride, (const char*)src + i0 * stride, (i1 - i0) * stride,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs));
          if (chunks > 1) {
            CK(cudaEventRecord(ctx->chunk_ev[c], cs));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
          }
        }
        if (c == chunks - 1) CK(cudaEventRecord(ctx->cev[1], ctx->stream));  // everything has arrived
        bounds_pack_kernel<<<grid_blocks(i1 - i0, 256, sms), 256, 0, ctx->stream>>>(ctx->d_raw.p, i0, i1, stride, ctx->d_packed.p,
                                                                                    ctx->d_bounds.p);
        ++ctx->launches;
      }
    } else {
      CK(cudaEventRecord(ctx->cev[1], ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->h_bounds.p, ctx->d_bounds.p, sizeof(BoundsDev), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // the host needs the bounds to size the grid; the caller's buffer is free from here on
    const BoundsDev b = *ctx->h_bounds.p;
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
    if (b.n_finite)
      for (int a = 0; a < 3; ++a) {
        mn[a] = ord2f(b.mn[a]);
        mx[a] = ord2f(b.mx[a]);
      }
    n_cells = size_grid(ctx, mn, mx, b.n_finite);
    rec = (const char*)ctx->d_packed.p;
    rec_stride = 16;
  }
  return grid_tail(ctx, rec, rec_stride, n, n_cells, hist_per_chunk ? chunk_bound : nullptr, nullptr, 0ull);
}

// set_cloud returns as soon as the host buffer has been consumed; its device timeline is read back lazily
void resolve_cloud_timing(b200lp_ctx* ctx) {
  if (!ctx->cloud_timing_pending) return;
  if (cudaEventSynchronize(ctx->cev[2]) == cudaSuccess) {
    cudaEventElapsedTime(&ctx->ms_upload, ctx->cev[0], ctx->cev[1]);
    cudaEventElapsedTime(&ctx->ms_grid, ctx->cev[1], ctx->cev[2]);
  }
  ctx->cloud_timing_pending = false;
}

void resolve_cycle_timing(b200lp_ctx* ctx) {
  resolve_cloud_timing(ctx);
  if (!ctx->cycle_timing_pending) return;
  if (cudaEventSynchronize(ctx->ev[2]) == cudaSuccess) {
    ctx->ms_readback = 0.f;
    if (!ctx->cycle_timing_direct) cudaEventElapsedTime(&ctx->ms_readback, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->ms_k_prep, ctx->ev[1], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->ms_k_cull, ctx->ev[ctx->cycle_overlapped ? 6 : 4], ctx->ev[7]);
    cudaEventElapsedTime(&ctx->ms_k_plan, ctx->ev[7], ctx->ev[5]);
    cudaEventElapsedTime(&ctx->ms_k_argmin, ctx->ev[5], ctx->ev[2]);
    if (ctx->cycle_overlapped) ctx->ms_plan = ctx->ms_k_prep + ctx->ms_k_cull + ctx->ms_k_plan + ctx->ms_k_argmin;  // prep ran under the grid build
    else cudaEventElapsedTime(&ctx->ms_plan, ctx->ev[1], ctx->ev[2]);
  }
  ctx->cycle_timing_pending = false;
}

// The sample layout of a single-robot launch, planned on the host (PrepPlan, lp_kernels.cuh): velocity window and the three
// VelocityIterator axes (the code every CTA would run, IEEE double on both sides), the shard [lo, hi) of a sample-sharded
// launch and the chunk layout that makes prep_kernel one wave. Returns the number of chunks to launch.
//
// Where a shard is cut: the sample grid is ordered by rising linear speed and a trajectory's cost grows with its pose count
// ceil(max(|v| T / g, |w| T / g_a)), so equal sample counts leave the last rank with about twice the poses of the first
// (C4 on 8 GPUs: 0.23 vs 0.34 ms). Per linear-speed row the expected pose count has a closed form (|w| taken as uniform
// over the angular axis, plus a fixed per-trajectory term); the rows' prefix sum is cut at the shares ctx->cuts names
// (k / count by default; adapt_cuts moves them with the device times of earlier cycles, which every rank sees through the
// exchange slots). Every rank runs this code on the same query and the same shares, so all ranks cut alike; the hash of
// the cuts travels in the exchange slots and a mismatch is reported (error bit 64).
int plan_samples(b200lp_ctx* ctx, const RobotIn& q, int rank, int count) {
  const b200lp_limits& L = ctx->C.lim;
  const b200lp_params& P = ctx->C.par;
  PrepPlan& pp = ctx->prep_plan;
  float (*ax)[kMaxAxis] = ctx->axes_host;  // the chains of the last window (kept: a robot at steady speed keeps its window)
  int n[3] = {0, 0, 0};
  const bool sampling_on = P.linear_x_sample * P.angular_z_sample > 0;
  const bool axes = sampling_on && P.theory != B200LP_THEORY_DD_ROTATE_INPLACE;
  long long n_raw = 0;
  pp.planned = 1;
  if (axes) {
    float win[6];
    velocity_window(L, P, q, win, win + 3);
    const float* mn = win;
    const float* mx = win + 3;
    if (ctx->axes_valid && memcmp(win, ctx->axes_window, sizeof(win)) == 0) {
      for (int a = 0; a < 3; ++a) n[a] = ctx->axes_n[a];  // same window, same chains, same closed form: nothing to redo
    } else {
      const int want[3] = {(int)P.linear_x_sample, (int)P.linear_y_sample, (int)P.angular_z_sample};
      n[0] = velocity_iterator_dev((double)mn[0], (double)mx[0], want[0], ax[0]);
      if (P.theory == B200LP_THEORY_OMNI_SIMPLE) n[1] = velocity_iterator_dev((double)mn[1], (double)mx[1], want[1], ax[1]);
      else { ax[1][0] = 0.f; n[1] = 1; }
      n[2] = velocity_iterator_dev((double)mn[2], (double)mx[2], want[2], ax[2]);
      // B200LP_AXIS_DEBUG (tests only): "0" leaves the chains to the kernel, "skew" hands it a closed form that is slightly
      // off, so that entries really differ and travel as exceptions (or, beyond kAxisExceptions, force the fallback)
      const char* dbg = getenv("B200LP_AXIS_DEBUG");
      pp.host_axes = (dbg && !strcmp(dbg, "0")) ? 0 : 1;
      pp.n_exc = 0;
      for (int a = 0; a < 3 && pp.host_axes; ++a) {
        AxisPlan& A = pp.ax[a];
        A.mn = (double)mn[a];
        A.step = 0.0;
        A.n_out = n[a];
        A.zero_at = -1;
        A.pad = 0;
        A.last = ax[a][n[a] - 1];
        if (a == 1 && P.theory != B200LP_THEORY_OMNI_SIMPLE) continue;  // the single 0.0f of a differential drive
        if (mn[a] != mx[a]) A.step = ((double)mx[a] - (double)mn[a]) / (double)(std::max(2, want[a]) - 1);  // as velocity_iterator_dev
        if (dbg && !strcmp(dbg, "skew")) A.step = nextafter(A.step * (1.0 + 1e-9), 1e300);  // (tests: a closed form that is off)
        if (n[a] == std::max(2, want[a]) + 1)  // one entry more than asked for: the inserted zero
          for (int i = 1; i < n[a] - 1; ++i)
            if (ax[a][i] == 0.0f && ax[a][i - 1] < 0.0f) { A.zero_at = i; break; }
        for (int i = 0; i < n[a]; ++i) {
          const float v = axis_value(A, i);
          if (memcmp(&v, &ax[a][i], sizeof(float)) == 0) continue;
          if (pp.n_exc == kAxisExceptions) { pp.host_axes = 0; break; }  // the kernel runs the chains itself
          pp.exc_at[pp.n_exc] = (a << 24) | i;
          pp.exc_val[pp.n_exc++] = ax[a][i];
        }
      }
      memcpy(ctx->axes_window, win, sizeof(win));
      for (int a = 0; a < 3; ++a) ctx->axes_n[a] = n[a];
      ctx->axes_valid = true;
      ctx->row_work_valid = false;
    }
    n_raw = (long long)n[0] * n[1] * n[2];
  } else {
    pp.host_axes = 0;
    pp.n_exc = 0;
    ctx->axes_valid = false;
    ctx->row_work_valid = false;
    if (sampling_on) n_raw = 2;
  }
  pp.n_raw = n_raw;
#if B200LP_CHECKS  // (checking build only) B200LP_CHECK_SELFTEST: hand the kernel a wrong sample count, so that a test can see a check fire
  if (getenv("B200LP_CHECK_SELFTEST")) pp.n_raw = n_raw + 1;
#endif
  // ---- the W + 1 sample cuts ----
  long long cut[B200LP_MAX_PEERS + 1];
  const int W = std::max(1, std::min(count, (int)B200LP_MAX_PEERS));
  for (int k = 0; k <= W; ++k) cut[k] = n_raw * k / W;
  if (W > 1 && axes && n_raw > 0) {
    float* w = ctx->row_work;  // inclusive prefix of the rows' estimated work: a function of the window, kept with the axes
    const int nx = n[0], nths = n[2];
    const long long row = (long long)n[1] * nths;
    if (!ctx->row_work_valid) {
      const float ta = (float)(P.sim_time / P.sim_granularity), tb = (float)(P.sim_time / P.angular_sim_granularity);
      const float bw = fmaxf(fabsf(ax[2][0]), fabsf(ax[2][nths - 1])) * tb;  // largest angular step count of the axis
      float acc = 0.f;
      for (int ix = 0; ix < nx; ++ix) {
        const float a = fabsf(ax[0][ix]) * ta;
        const float mean_steps = (a < bw) ? a + (bw - a) * (bw - a) / (2.0f * bw) : a;  // E[max(a, U(0, bw))]
        acc += mean_steps + 10.0f;  // + the fixed part of a trajectory (work fetch, critic epilogue, prep)
        w[ix] = acc;
      }
      ctx->row_work_valid = true;
    }
    const float total = w[nx - 1];
    for (int k = 1; k < W; ++k) {
      const float share = ctx->cuts.n == W ? ctx->cuts.frac[k] : (float)k / (float)W;
      const float target = total * share;
      int a = 0, b = nx - 1;  // smallest row whose inclusive prefix reaches the target
      while (a < b) {
        const int m = (a + b) >> 1;
        if (w[m] >= target) b = m; else a = m + 1;
      }
      const float prev = a ? w[a - 1] : 0.f;
      const float span = w[a] - prev;
      float frac = span > 0.f ? (target - prev) / span : 0.f;
      frac = fminf(fmaxf(frac, 0.f), 1.f);
      long long c = (long long)a * row + (long long)(frac * (float)row);
      c = std::max((long long)a * row, std::min(c, (long long)(a + 1) * row));
      cut[k] = std::max(c, cut[k - 1]);
    }
  }
  const int r = std::max(0, std::min(rank, W - 1));
  pp.lo = cut[r];
  pp.hi = cut[r + 1];
  unsigned long long h = 1469598103934665603ull;  // FNV-1a over the cuts
  for (int k = 0; k <= W; ++k)
    for (int b = 0; b < 8; ++b) {
      h ^= ((unsigned long long)cut[k] >> (8 * b)) & 0xffull;
      h *= 1099511628211ull;
    }
  ctx->sample_cuts_hash = h;
  // ---- chunk layout: the shard in kPrepThreads-sample chunks, everything else in about two count chunks per SM ----
  const long long outside = n_raw - (pp.hi - pp.lo);
  const long long target = 2ll * std::max(1, ctx->sm_count);
  const long long per = std::max(1ll, (outside + target * kPrepThreads - 1) / (target * kPrepThreads));
  pp.count_span = (int)std::min<long long>(per, 1 << 20) * kPrepThreads;
  pp.nb = (int)((pp.lo + pp.count_span - 1) / pp.count_span);
  pp.ns = (int)std::max<long long>(1, (pp.hi - pp.lo + kPrepThreads - 1) / kPrepThreads);
  const int na = (int)((n_raw - pp.hi + pp.count_span - 1) / pp.count_span);
  return pp.nb + pp.ns + na;
}

// Moves the shard cuts of the NEXT sample-sharded cycle with the device times the ranks needed for this one. Every rank
// ends an exchange cycle with the same W durations (they travel in the exchange slots) and the same current cuts, and
// runs the same arithmetic: all ranks arrive at the same new cuts without another exchange. Model: the time of rank r is
// spread evenly over its share [f_r, f_r+1) of the estimated-work axis; the new cuts sit at equal shares of the summed
// time, approached half way per cycle.
void adapt_cuts(b200lp_ctx* ctx, const uint32_t* ns, int W) {
  if (W < 2 || W > B200LP_MAX_PEERS) return;
  double T[B200LP_MAX_PEERS], total = 0.0, mx = 0.0;
  for (int r = 0; r < W; ++r) {
    if (ns[r] == 0u) return;
    T[r] = (double)ns[r];
    total += T[r];
    mx = std::max(mx, T[r]);
  }
  if (mx <= 1.03 * total / W) return;  // balanced to 3 %: leave the cuts alone
  double f[B200LP_MAX_PEERS + 1];
  for (int k = 0; k <= W; ++k) f[k] = ctx->cuts.n == W ? (double)ctx->cuts.frac[k] : (double)k / W;
  double nf[B200LP_MAX_PEERS + 1];
  nf[0] = 0.0;
  nf[W] = 1.0;
  double cum = 0.0;
  int r = 0;
  for (int k = 1; k < W; ++k) {
    const double target = total * k / W;
    while (r < W - 1 && cum + T[r] < target) cum += T[r++];
    const double want = f[r] + (target - cum) / T[r] * (f[r + 1] - f[r]);
    nf[k] = f[k] + 0.5 * (want - f[k]);
  }
  for (int k = 1; k <= W; ++k) nf[k] = std::max(nf[k], nf[k - 1]);
  ctx->cuts.n = W;
  for (int k = 0; k <= W; ++k) ctx->cuts.frac[k] = (float)std::min(1.0, std::max(0.0, nf[k]));
  ctx->cuts.frac[0] = 0.f;
  ctx->cuts.frac[W] = 1.f;
}

int run_cycle(b200lp_ctx* ctx, size_t n_robots, int rank, int count, b200lp_result* outs, bool plan_resident = false,
              bool exchange = false) {
  TraceRange trace_range("b200lp plan cycle");
  static const bool host_trace = getenv("B200LP_HOST_TRACE") != nullptr;
  long long ht[5] = {0, 0, 0, 0, 0};
  if (host_trace) ht[0] = host_ns();
  // inputs are already staged in h_robots / h_plan7 (pinned); total plan poses in plan_total
  const int t_cap = traj_cap(ctx->C.par);
  const size_t T = n_robots * (size_t)t_cap;
  size_t plan_total = 0;
  for (size_t i = 0; i < n_robots; ++i) plan_total = std::max<size_t>(plan_total, ctx->h_robots.p[i].plan_off + ctx->h_robots.p[i].plan_n);
  const int nc = std::max(1, ctx->C.n_critics);
  const int n_chunks_cap = (t_cap + kPrepSamples - 1) / kPrepSamples + 4;  // (a planned layout never needs more: three regions, each rounded up once)
  // Upper bound on the trajectories one launch scores per robot. A sample shard's cuts sit at equal shares of the estimated
  // pose count, not of the sample count (prep_kernel), so a shard of slow trajectories can hold far more than t_cap / count
  // of them: the only bound that always holds is t_cap. plan_kernel takes the real count from prep_kernel's meta.
  const int cap_local = t_cap;
  CK(ctx->d_robots.reserve(n_robots));
  CK(ctx->d_meta.reserve(n_robots));
  if (!plan_resident) CK(ctx->d_plan7.reserve(std::max<size_t>(plan_total * 7, 7)));  // (a resident plan must not move)
  CK(ctx->d_plan_pts.reserve(std::max<size_t>(plan_total, 1)));
  CK(ctx->d_rec_vel.reserve(T));
  CK(ctx->d_rec_steps.reserve(T));
  CK(ctx->d_rec_sample.reserve(T));
  CK(ctx->d_first_hit.reserve(T));
  CK(ctx->d_rec_dt.reserve(T));
  CK(ctx->d_cost.reserve(T));
  CK(ctx->d_scores.reserve(T * nc));
  CK(ctx->d_results.reserve(n_robots));
  CK(ctx->h_results.reserve(n_robots));
  CK(ctx->h_meta.reserve(n_robots));
  // the forward simulation of every scored trajectory is materialised once per cycle (16 B per pose)
  // (a multiple of 256: cull_kernel's CTAs and the words of its survivor bits never straddle two robots)
  const long long pose_stride = ((long long)cap_local * (long long)max_steps_bound(ctx->C.lim, ctx->C.par) + 255) / 256 * 256;
  if ((double)pose_stride * (double)n_robots * 16.0 > 64e9)
    return ctx->fail(B200LP_E_NOMEM, "plan: %zu robots x %d trajectories x %d poses need more than 64 GB of pose storage",
                     n_robots, cap_local, (int)max_steps_bound(ctx->C.lim, ctx->C.par));
  CK(ctx->d_poses.reserve((size_t)pose_stride * n_robots));
  CK(ctx->d_rec_pose_off.reserve(T));
  if (T > 0x7fffffffull) return ctx->fail(B200LP_E_INVALID, "plan: %zu robots x %d trajectories exceed the work-list index range", n_robots, t_cap);
  CK(ctx->d_surv.reserve(T * (size_t)(((int)max_steps_bound(ctx->C.lim, ctx->C.par) + 31) / 32)));
  CK(ctx->d_order.reserve(T * (size_t)kCostClasses));
  if (ctx->d_hist.cap < T) {
    CK(ctx->d_hist.reserve(T));
    CK(cudaMemsetAsync(ctx->d_hist.p, 0, ctx->d_hist.cap * sizeof(unsigned), ctx->stream));  // (cull_kernel runs on this stream)
  }
  int want_pp = 0;
  for (int k = 0; k < ctx->C.n_critics; ++k) want_pp |= ctx->C.critics[k].kind == B200LP_CRITIC_PURE_PURSUIT;
  CK(ctx->d_rec_pp.reserve(want_pp ? T : 1));
  ctx->pose_stride = pose_stride;
  const int argmin_ctas = (int)std::max(1, std::min(kArgminMaxCtas, (cap_local + 2047) / 2048));
  CK(ctx->d_partial.reserve(std::max(n_robots * (size_t)kArgminMaxCtas, (size_t)16384)));  // >= the persistent plan grid (SMs x resident CTAs)
  bool init_on_main = false;  // first use / growth: counters are zeroed on the main stream, prep_kernel must stay behind that
  if (ctx->d_tickets2.cap < n_robots) {
    init_on_main = true;
    CK(ctx->d_tickets2.reserve(n_robots));
    CK(cudaMemsetAsync(ctx->d_tickets2.p, 0, ctx->d_tickets2.cap * sizeof(unsigned), ctx->stream));
  }
  if (ctx->d_tickets.cap < n_robots) {
    init_on_main = true;
    CK(ctx->d_tickets.reserve(n_robots));
    CK(cudaMemsetAsync(ctx->d_tickets.p, 0, ctx->d_tickets.cap * sizeof(unsigned), ctx->stream));
  }
  if (ctx->d_aggs.cap < n_robots * (size_t)n_chunks_cap) {
    init_on_main = true;
    CK(ctx->d_aggs.reserve(n_robots * (size_t)n_chunks_cap));
    CK(cudaMemsetAsync(ctx->d_aggs.p, 0, ctx->d_aggs.cap * sizeof(PrepAgg), ctx->stream));
    ctx->epoch = 0;
  }
  if (!ctx->d_work.p) {
    init_on_main = true;
    CK(ctx->d_work.reserve(1));
    CK(cudaMemsetAsync(ctx->d_work.p, 0, sizeof(unsigned long long), ctx->stream));
    CK(ctx->d_tstart.reserve(1));
    CK(cudaMemsetAsync(ctx->d_tstart.p, 0, sizeof(unsigned long long), ctx->stream));
    CK(ctx->d_class_counts.reserve(kCostClasses));
    CK(cudaMemsetAsync(ctx->d_class_counts.p, 0, kCostClasses * sizeof(unsigned), ctx->stream));
  }
  if (!ctx->plan_ctas_per_sm) {
    ctx->sm_count = sm_count_of(ctx->device);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->plan_ctas_per_sm, plan_kernel, kThreads, 0));
    if (ctx->plan_ctas_per_sm < 1) return ctx->fail(B200LP_E_CUDA, "plan_kernel does not fit on an SM");
  }
  if (!ctx->have_cloud) {  // no cloud yet: an empty one (collision critics return 0.0, size() < 5)
    int rc = build_grid(ctx, nullptr, 0, 16, false);
    if (rc) return rc;
  }
  ctx->n_robots = n_robots;
  ctx->t_cap = t_cap;
  ctx->shard_rank = rank;
  ctx->shard_count = count;
  if (++ctx->epoch == 0u) ctx->epoch = 1u;
  const bool direct = n_robots == 1;
  if (!direct) {
    if (!ctx->h_fleet_seq.p) {
      CK(ctx->h_fleet_seq.reserve(1));
      *ctx->h_fleet_seq.p = 0ull;
      CK(ctx->d_fleet_done.reserve(1));
      CK(cudaMemsetAsync(ctx->d_fleet_done.p, 0, sizeof(unsigned), ctx->stream));
    }
    ++ctx->fleet_seq;
  }
  if (direct) {
    if (!ctx->h_direct.p) {
      CK(ctx->h_direct.reserve(1));
      memset(ctx->h_direct.p, 0, sizeof(DirectOut));
    }
    ++ctx->direct_seq;
  }
  ctx->cycle_timing_pending = false;  // nobody asked for the previous cycle's timeline; its events are re-recorded now

  const unsigned long long work = (unsigned long long)n_robots * (unsigned long long)cap_local;
  const unsigned plan_grid = (unsigned)std::max<unsigned long long>(
      1ull, std::min<unsigned long long>((work + kWarpsPerCta - 1) / kWarpsPerCta,
                                         (unsigned long long)ctx->sm_count * ctx->plan_ctas_per_sm));

  // A cloud handed over just before this call may still be on its way (set_cloud returns when the caller's buffer has been
  // read): the query upload and prep_kernel do not touch the grid, so they run on a second stream underneath the rest of the
  // upload / grid build, and only plan_kernel waits for both. Everything the prep stream overwrites was last used by the
  // previous cycle, whose end on the main stream is ev[2] (as recorded then; it is re-recorded further down).
  bool overlap = false;
  if (ctx->cloud_timing_pending) {
    if (cudaEventQuery(ctx->cev[2]) == cudaErrorNotReady) overlap = !init_on_main;
    else resolve_cloud_timing(ctx);  // the grid is ready: read its timeline now and stop asking
  }
  cudaStream_t ps = overlap ? ctx->prep_stream : ctx->stream;
  if (overlap && ctx->have_cycle_event) CK(cudaStreamWaitEvent(ps, ctx->ev[2], 0));
  if (overlap && ctx->plan_ev_pending) CK(cudaStreamWaitEvent(ps, ctx->plan_ev, 0));
  ctx->plan_ev_pending = false;  // (on the main stream the upload is ordered before the kernels anyway)
  ctx->cycle_overlapped = overlap;
  // A fleet's plan table (1.7 MB for 512 robots) goes up on the copy stream BESIDE prep_kernel: the robots' records carry the
  // one pose of each plan prep_kernel needs (the goal), and the float copy of the plan positions is made by cull_kernel,
  // which waits for the table like plan_kernel behind it. Ordered behind everything the main stream was given so far
  // (the previous cycle's readers of d_plan7 / d_plan_pts).
  int defer_pts = 0;
  if (n_robots > 1) CK(cudaMemcpyAsync(ctx->d_robots.p, ctx->h_robots.p, n_robots * sizeof(RobotIn), cudaMemcpyHostToDevice, ps));
  if (plan_total && !plan_resident && !ctx->plan_uploaded) {
    const double* src = ctx->plan_src ? ctx->plan_src : ctx->h_plan7.p;
    if (n_robots > 1 && !overlap) {
      if (!ctx->fleet_fork_ev) CK(cudaEventCreateWithFlags(&ctx->fleet_fork_ev, cudaEventDisableTiming));
      if (!ctx->fleet_plan_ev) CK(cudaEventCreateWithFlags(&ctx->fleet_plan_ev, cudaEventDisableTiming));
      CK(cudaEventRecord(ctx->fleet_fork_ev, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->fleet_fork_ev, 0));
      CK(cudaMemcpyAsync(ctx->d_plan7.p, src, plan_total * 7 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
      CK(cudaEventRecord(ctx->fleet_plan_ev, ctx->copy_stream));
      defer_pts = 1;
    } else {
      CK(cudaMemcpyAsync(ctx->d_plan7.p, src, plan_total * 7 * sizeof(double), cudaMemcpyHostToDevice, ps));
    }
  }
  CK(cudaEventRecord(ctx->ev[1], ps));
  // single-robot cycles hand their query to the kernels as an argument; the device copy (read by the read-back kernels
  // only) follows the launches
  const int by_value = n_robots == 1 ? 1 : 0;
  const RobotIn q0 = ctx->h_robots.p[0];
  int n_chunks = n_chunks_cap - 4;
  if (host_trace) ht[1] = host_ns();
  if (by_value) n_chunks = plan_samples(ctx, q0, rank, count);
  else ctx->prep_plan.planned = 0;
  if (n_chunks > n_chunks_cap) return ctx->fail(B200LP_E_STATE, "plan: %d sample chunks planned, %d provided for", n_chunks, n_chunks_cap);
  // the longest axis the parameter set can produce (VelocityIterator: max(2, samples) entries + the inserted zero) + the slot
  // the iterator writes ahead of its count; b200lp_create keeps this below kMaxAxis
  const int axis_cap = std::max(axis_count(ctx->C.par.linear_x_sample),
                                std::max(axis_count(ctx->C.par.linear_y_sample), axis_count(ctx->C.par.angular_z_sample))) + 1;
  prep_kernel<<<dim3((unsigned)n_chunks, (unsigned)n_robots), kPrepThreads, 3 * (size_t)axis_cap * sizeof(float), ps>>>(
      ctx->C, ctx->d_robots.p, q0, by_value, ctx->d_tstart.p, t_cap, ctx->prep_plan, ctx->epoch, ctx->d_tickets.p,
      ctx->d_aggs.p, ctx->d_rec_vel.p, ctx->d_rec_steps.p, ctx->d_rec_dt.p, ctx->d_rec_sample.p, ctx->d_meta.p, ctx->d_plan7.p,
      ctx->d_plan_pts.p, ctx->d_rec_pose_off.p, ctx->d_poses.p, pose_stride, ctx->d_rec_pp.p, want_pp, ctx->d_class_counts.p, axis_cap,
      defer_pts);
  if (host_trace) ht[2] = host_ns();
  CK(cudaEventRecord(ctx->ev[4], ps));
  if (overlap) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[4], 0));
    CK(cudaEventRecord(ctx->ev[6], ctx->stream));  // cull_kernel's start on the main stream (after the grid build)
  }
  // the float pre-cull of every pose (needs the grid) and the work lists plan_kernel drains
  const int mask_stride = ((int)max_steps_bound(ctx->C.lim, ctx->C.par) + 31) / 32;
  if (defer_pts) CK(cudaStreamWaitEvent(ctx->stream, ctx->fleet_plan_ev, 0));  // cull_kernel converts the plan positions
  // (a planned launch lists at most ns x kPrepThreads trajectories: a sample shard does not pay for the whole grid's CTAs)
  const int cull_traj = by_value ? std::min<long long>(t_cap, (long long)ctx->prep_plan.ns * kPrepThreads) : t_cap;
  cull_kernel<<<dim3((unsigned)((cull_traj + kCullTraj - 1) / kCullTraj), (unsigned)n_robots), kCullThreads, 0, ctx->stream>>>(
      ctx->C, ctx->grid, ctx->d_robots.p, q0, by_value, ctx->d_meta.p, t_cap, ctx->d_rec_steps.p, ctx->d_rec_pose_off.p,
      ctx->d_poses.p, ctx->d_surv.p, mask_stride, ctx->d_order.p, T, ctx->d_class_counts.p, ctx->d_hist.p, ctx->d_plan7.p,
      ctx->d_plan_pts.p, defer_pts);
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  PeerExchange px{};
  px.t_start = ctx->d_tstart.p;
  if (exchange) {  // the cross-GPU argmin through peer memory, by the last CTA of plan_kernel
    px.world = ctx->peer_world;
    px.rank = ctx->peer_rank;
    px.seq = ctx->peer_seq;  // (advanced by the caller before anything could fail)
    px.mine = ctx->peer_slots();
    px.timeout_cycles = (long long)4e9;  // ~2 s of SM clocks
    px.cuts_hash = ctx->sample_cuts_hash;
    px.peers = ctx->peer_table;
  }
  plan_kernel<<<plan_grid, kThreads, 0, ctx->stream>>>(
      ctx->C, ctx->grid, ctx->d_robots.p, q0, by_value, ctx->d_meta.p, (int)n_robots, t_cap, ctx->d_rec_vel.p,
      ctx->d_rec_steps.p, ctx->d_rec_pose_off.p, ctx->d_poses.p, ctx->d_rec_pp.p, ctx->d_plan_pts.p, ctx->d_cost.p,
      ctx->d_scores.p, ctx->d_first_hit.p, ctx->d_work.p, ctx->d_partial.p, ctx->d_tickets2.p, ctx->d_results.p,
      direct ? ctx->h_direct.p : nullptr, ctx->direct_seq, px, ctx->d_surv.p, mask_stride, ctx->d_order.p, T,
      ctx->d_class_counts.p, ctx->d_hist.p);
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  if (n_robots > 1) {  // a single robot's argmin is folded into plan_kernel
    argmin_kernel<<<dim3((unsigned)argmin_ctas, (unsigned)n_robots), kArgminThreads, 0, ctx->stream>>>(
        ctx->C, ctx->d_meta.p, t_cap, ctx->d_rec_vel.p, ctx->d_cost.p, ctx->d_first_hit.p, ctx->d_partial.p,
        ctx->d_tickets2.p, ctx->d_results.p, ctx->d_work.p, ctx->h_results.p, ctx->h_meta.p, ctx->d_fleet_done.p,
        ctx->h_fleet_seq.p, ctx->fleet_seq);
    ++ctx->launches;
  }
  ctx->launches += 3;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->have_cycle_event = true;
  if (n_robots == 1)  // for the read-back kernels (poses_kernel, count_radius_kernel); off the cycle's critical path
    CK(cudaMemcpyAsync(ctx->d_robots.p, ctx->h_robots.p, sizeof(RobotIn), cudaMemcpyHostToDevice, ctx->stream));
  if (host_trace) ht[3] = host_ns();
  if (direct) {
    // the kernel's last CTA writes the result block into pinned host memory and raises seq: no copies, no stream sync
    CK(cudaGetLastError());
    volatile unsigned long long* flag = &ctx->h_direct.p->seq;
    unsigned spins = 0;
    while (*flag != ctx->direct_seq) {
      __builtin_ia32_pause();
      if ((++spins & 0xfffu) == 0u) {  // every few microseconds: is the stream dead or done without raising the flag?
        const cudaError_t qe = cudaStreamQuery(ctx->stream);
        if (qe == cudaErrorNotReady) continue;
        if (qe != cudaSuccess) return ctx->fail(B200LP_E_CUDA, "plan: %s", cudaGetErrorString(qe));
        if (*flag != ctx->direct_seq) return ctx->fail(B200LP_E_CUDA, "plan: the kernels finished without publishing a result");
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    ctx->h_results.p[0] = ctx->h_direct.p->r;
    ctx->h_meta.p[0] = ctx->h_direct.p->m;
    ctx->last_cycle_ns = ctx->h_direct.p->cycle_ns;
    if (host_trace) {
      ht[4] = host_ns();
      fprintf(stderr, "host trace: entry -> launches %.1f us (reserves, events), plan_samples + prep launch %.1f us, cull + plan launches %.1f us, "
                      "wait %.1f us; call %.1f us, device cycle %.1f us\n", (ht[1] - ht[0]) / 1e3, (ht[2] - ht[1]) / 1e3, (ht[3] - ht[2]) / 1e3,
              (ht[4] - ht[3]) / 1e3, (ht[4] - ht[0]) / 1e3, ctx->last_cycle_ns / 1e3);
    }
    if (exchange) memcpy(ctx->last_peer_ns, ctx->h_direct.p->peer_ns, sizeof(ctx->last_peer_ns));
  } else {
    // argmin_kernel writes every robot's result and meta block into pinned host memory and the last robot raises the
    // sequence word: no read-back copies, no stream synchronisation (a fleet of 512: -0.02 ms per call)
    CK(cudaGetLastError());
    volatile unsigned long long* flag = ctx->h_fleet_seq.p;
    unsigned spins = 0;
    while (*flag != ctx->fleet_seq) {
      __builtin_ia32_pause();
      if ((++spins & 0xfffu) == 0u) {  // every few microseconds: is the stream dead or done without raising the flag?
        const cudaError_t qe = cudaStreamQuery(ctx->stream);
        if (qe == cudaErrorNotReady) continue;
        if (qe != cudaSuccess) return ctx->fail(B200LP_E_CUDA, "plan_batch: %s", cudaGetErrorString(qe));
        if (*flag != ctx->fleet_seq) return ctx->fail(B200LP_E_CUDA, "plan_batch: the kernels finished without publishing the results");
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (host_trace) {
      ht[4] = host_ns();
      fprintf(stderr, "host trace (fleet of %zu): entry -> launches %.1f us, prep launch %.1f us, other launches %.1f us, wait %.1f us; call %.1f us\n",
              n_robots, (ht[1] - ht[0]) / 1e3, (ht[2] - ht[1]) / 1e3, (ht[3] - ht[2]) / 1e3, (ht[4] - ht[3]) / 1e3, (ht[4] - ht[0]) / 1e3);
    }
  }
  ctx->meta_host.assign(ctx->h_meta.p, ctx->h_meta.p + n_robots);
  if (exchange && (ctx->meta_host[0].error & 8))
    return ctx->fail(B200LP_E_STATE, "plan_shard_exchange: a peer did not deliver its result within the time limit "
                                     "(call b200lp_peer_resync on every rank before the next cycle)");
  if (exchange && (ctx->meta_host[0].error & 64))
    return ctx->fail(B200LP_E_STATE, "plan_shard_exchange: the ranks cut the sample grid differently "
                                     "(call b200lp_peer_resync on every rank before the next cycle)");
  if (exchange && ctx->adaptive_cuts && !ctx->meta_host[0].error) adapt_cuts(ctx, ctx->last_peer_ns, ctx->peer_world);
  for (size_t i = 0; i < n_robots; ++i) {
    if (ctx->meta_host[i].error & 16)
      return ctx->fail(B200LP_E_STATE, "robot %zu: plan_kernel scored %d trajectories / %lld poses where prep_kernel listed %d / %lld",
                       i, ctx->h_results.p[i].n_traj, (long long)ctx->h_results.p[i].n_poses,
                       ctx->meta_host[i].t_end - ctx->meta_host[i].t_begin, (long long)ctx->meta_host[i].n_poses);
    if (ctx->meta_host[i].error)
      return ctx->fail(B200LP_E_INVALID, "robot %zu: a trajectory exceeds B200LP_MAX_STEPS=%d poses or the trajectory / pose list overflowed",
                       i, B200LP_MAX_STEPS);
    outs[i] = ctx->h_results.p[i];
  }
  ctx->have_cycle = true;
  ctx->cycle_timing_pending = true;  // the device timeline is read back when somebody asks for it
  ctx->cycle_timing_direct = true;  // (no read-back copies on either path: results reach the host through mapped memory)
  return B200LP_OK;
}

void fill_robot(RobotIn* r, const b200lp_query* q, int64_t plan_off, int32_t plan_n, const double* goal7 /* may be null */) {
  memcpy(r->pose, q->pose, sizeof(r->pose));
  memcpy(r->twist, q->twist, sizeof(r->twist));
  r->max_speed_override = q->max_speed_override;
  r->heading_deviation = q->heading_deviation;
  r->plan_off = plan_off;
  r->plan_n = plan_n;
  r->goal_valid = (goal7 && plan_n > 0) ? 1 : 0;
  if (r->goal_valid) memcpy(r->goal, goal7, sizeof(r->goal));
  else memset(r->goal, 0, sizeof(r->goal));
}

}  // namespace

extern "C" {

int b200lp_abi_version(void) { return B200LP_ABI_VERSION; }

const char* b200lp_last_error(const b200lp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b200lp_create(b200lp_ctx** out, int device, const b200lp_limits* limits, const b200lp_params* params,
                  const float* cuboid_xyz, const b200lp_critic* critics, int n_critics, const b200lp_grid_config* grid) {
  auto bad = [&](const char* msg) {
    g_create_error = msg;
    return B200LP_E_INVALID;
  };
  if (!out || !limits || !params || !cuboid_xyz) return bad("null argument");
  if (n_critics < 0 || n_critics > B200LP_MAX_CRITICS || (n_critics && !critics)) return bad("bad critic list");
  if (params->theory < 0 || params->theory > B200LP_THEORY_DD_ROTATE_INPLACE) return bad("unknown theory");
  for (int k = 0; k < n_critics; ++k)
    if (critics[k].kind < 0 || critics[k].kind > B200LP_CRITIC_TWIRLING) return bad("unknown critic kind");
  if (!(params->controller_frequency > 0) || !(params->sim_time > 0) || !(params->sim_granularity > 0) ||
      !(params->angular_sim_granularity > 0))
    return bad("controller_frequency, sim_time and the granularities must be positive");
  if (params->theory == B200LP_THEORY_DD_ROTATE_INPLACE && !(std::fabs(limits->rotation_speed) > 0))
    return bad("rotation_speed must be non-zero");
  if (!(max_steps_bound(*limits, *params) <= (double)B200LP_MAX_STEPS))
    return bad("parameter set can produce more than B200LP_MAX_STEPS poses per trajectory");
  if (!(limits->max_vel_theta * params->sim_time < 100.0)) return bad("max_vel_theta*sim_time must stay below 100 rad");
  if ((int)params->linear_x_sample > kMaxAxis - 2 || (int)params->angular_z_sample > kMaxAxis - 2 ||
      (params->theory == B200LP_THEORY_OMNI_SIMPLE && (int)params->linear_y_sample > kMaxAxis - 2))
    return bad("too many samples per velocity axis");
  if ((long long)axis_count(params->linear_x_sample) * axis_count(params->angular_z_sample) *
          (params->theory == B200LP_THEORY_OMNI_SIMPLE ? axis_count(params->linear_y_sample) : 1) > (1ll << 24))
    return bad("more than 2^24 velocity samples");

  for (int k = 0; k < 24; ++k)
    if (!std::isfinite(cuboid_xyz[k])) return bad("cuboid vertices must be finite");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    return B200LP_E_CUDA;
  }
  if (device < 0 || device >= ndev) return bad("device index out of range");
  b200lp_ctx* ctx = new (std::nothrow) b200lp_ctx();
  if (!ctx) return B200LP_E_NOMEM;
  ctx->device = device;
  auto cuda_fail = [&](cudaError_t err, const char* what) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    delete ctx;
    return B200LP_E_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  if ((e = cudaStreamCreateWithFlags(&ctx->prep_stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");

  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  for (auto& ev : ctx->chunk_ev)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  for (auto& ev : ctx->cev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  for (auto& ev : ctx->oev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  ctx->C.lim = *limits;
  ctx->C.par = *params;
  memcpy(ctx->C.cuboid, cuboid_xyz, sizeof(ctx->C.cuboid));
  ctx->C.n_critics = n_critics;
  for (int k = 0; k < n_critics; ++k) {
    ctx->C.critics[k].kind = critics[k].kind;
    ctx->C.critics[k].pad = 0;
    ctx->C.critics[k].weight = critics[k].weight;
    ctx->C.critics[k].tw = critics[k].translation_weight;
    ctx->C.critics[k].ow = critics[k].orientation_weight;
  }
  {  // robot-frame bounding box of the cuboid: what plan_kernel's float pre-cull rotates instead of the 8 vertices
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) {
      float mn = ctx->C.cuboid[0][a], mx = mn;
      for (int k = 1; k < 8; ++k) {
        mn = std::min(mn, ctx->C.cuboid[k][a]);
        mx = std::max(mx, ctx->C.cuboid[k][a]);
      }
      ctx->C.box_c[a] = 0.5f * (mn + mx);
      ctx->C.box_h[a] = std::nextafter(0.5f * (mx - mn), INFINITY) + 1e-6f * std::max(std::fabs(mn), std::fabs(mx));
      ext += std::max(std::fabs(mn), std::fabs(mx));
    }
    // float rounding of loose_box: ~1.5e-6 relative error of the rotation entries times the cuboid's reach, plus a floor
    ctx->C.box_slack = 2e-5f * (1.0f + ext);
    ctx->C.pad2 = 0.f;
  }
  if (grid) ctx->gcfg = *grid;
  *out = ctx;
  return B200LP_OK;
}

static void peer_detach(b200lp_ctx* ctx);

void b200lp_destroy(b200lp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->d_raw.release(); ctx->d_pts.release(); ctx->d_cell_start.release(); ctx->d_rank.release(); ctx->d_scan_status.release(); ctx->d_scan_ticket.release();
  ctx->d_packed.release(); ctx->d_sat.release(); ctx->d_bounds.release(); ctx->d_total.release();
  ctx->h_bounds.release(); ctx->d_robots.release(); ctx->d_meta.release(); ctx->d_plan7.release();
  ctx->d_plan_pts.release(); ctx->d_rec_vel.release(); ctx->d_rec_steps.release(); ctx->d_rec_sample.release();
  ctx->d_first_hit.release(); ctx->d_rec_dt.release(); ctx->d_cost.release(); ctx->d_scores.release();
  ctx->d_rec_pose_off.release(); ctx->d_poses.release(); ctx->d_rec_pp.release(); ctx->d_gplan7.release(); ctx->d_prune_pcl.release(); ctx->d_prune_meta.release(); ctx->h_prune_meta.release(); ctx->d_blocked.release(); ctx->h_blocked.release(); ctx->d_partial.release(); ctx->d_tickets2.release(); ctx->d_tickets.release(); ctx->d_aggs.release(); ctx->d_work.release(); ctx->d_tstart.release(); ctx->d_results.release(); ctx->d_count.release(); ctx->d_scratch.release();
  ctx->h_robots.release(); ctx->h_plan7.release(); ctx->h_results.release(); ctx->h_meta.release();
  ctx->h_count.release(); ctx->h_direct.release(); ctx->h_fleet_seq.release(); ctx->d_fleet_done.release();
  ctx->d_surv.release(); ctx->d_order.release(); ctx->d_hist.release(); ctx->d_class_counts.release();
  ctx->d_scan.release(); ctx->d_obs_a.release(); ctx->d_obs_b.release(); ctx->d_obs_hist.release(); ctx->d_obs_sums.release();
  ctx->d_obs_heads.release(); ctx->d_obs_counts.release(); ctx->h_obs_counts.release();
  for (auto& b : ctx->d_obs_out) b.release();
  for (auto& ev : ctx->oev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->chunk_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->cev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->plan_ev) cudaEventDestroy(ctx->plan_ev);
  if (ctx->fleet_fork_ev) cudaEventDestroy(ctx->fleet_fork_ev);
  if (ctx->fleet_plan_ev) cudaEventDestroy(ctx->fleet_plan_ev);
  peer_detach(ctx);
  ctx->d_peer_block.release();
  ctx->d_push_ticket.release(); ctx->d_share_err.release(); ctx->h_share_hdr.release(); ctx->h_share_err.release();
  if (ctx->push_done) cudaEventDestroy(ctx->push_done);
  if (ctx->push_stream) { cudaStreamSynchronize(ctx->push_stream); cudaStreamDestroy(ctx->push_stream); }
  if (ctx->prep_stream) { cudaStreamSynchronize(ctx->prep_stream); cudaStreamDestroy(ctx->prep_stream); }

  delete ctx->pack_pool;
  ctx->h_stage.release();
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int set_cloud_common(b200lp_ctx* ctx, const void* pts, size_t n, size_t stride, bool on_device) {
  if (!ctx) return B200LP_E_INVALID;
  if ((n && !pts) || stride < 12 || (stride & 3)) return ctx->fail(B200LP_E_INVALID, "set_cloud: bad pointer or stride");
  if (n > 0xfffffff0ull) return ctx->fail(B200LP_E_INVALID, "set_cloud: more than 2^32 points");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_raw.reserve(std::max<size_t>(n * stride, 16)));
  ctx->raw_stride = stride;
  ctx->cloud_timing_pending = false;
  int rc = build_grid(ctx, n ? pts : nullptr, n, stride, on_device);  // records cev[0] (start) and cev[1] (cloud arrived)
  if (rc) return rc;
  ctx->last_h2d_bytes = on_device ? 0 : n * ctx->raw_stride;  // raw_stride is 12 when the cloud was packed on the host
  CK(cudaEventRecord(ctx->cev[2], ctx->stream));
  // No synchronisation here: the caller's buffer was consumed before build_grid's host round trip for the bounds, and
  // the histogram / scan / scatter / summed-volume kernels still in flight are ordered before everything later on this
  // stream. ms_upload (start -> last piece on the device, bounds_pack_kernel of the earlier pieces underneath) and
  // ms_grid_build (what the grid adds after that) are read back when somebody asks (b200lp_last_timing) or the next
  // cycle has drained the stream anyway.
  ctx->cloud_timing_pending = true;
  ctx->have_cycle = false;
  return B200LP_OK;
}

int b200lp_set_cloud(b200lp_ctx* ctx, const void* pts, size_t n, size_t stride_bytes) {
  return set_cloud_common(ctx, pts, n, stride_bytes, false);
}
int b200lp_set_cloud_device(b200lp_ctx* ctx, const void* dev_pts, size_t n, size_t stride_bytes) {
  return set_cloud_common(ctx, dev_pts, n, stride_bytes, true);
}

int b200lp_peer_reserve_cloud(b200lp_ctx* ctx, size_t max_points) {
  if (!ctx) return B200LP_E_INVALID;
  if (ctx->d_peer_block.p) return ctx->fail(B200LP_E_STATE, "peer_reserve_cloud: call it before b200lp_peer_export");
  if (max_points > 0xfffffff0ull) return ctx->fail(B200LP_E_INVALID, "peer_reserve_cloud: more than 2^32 points");
  ctx->peer_cloud_cap = max_points;
  return B200LP_OK;
}

int b200lp_set_cloud_shared(b200lp_ctx* ctx, int root, const void* pts, size_t n, size_t stride) {
  if (!ctx) return B200LP_E_INVALID;
  if (ctx->peer_world < 1) return ctx->fail(B200LP_E_STATE, "set_cloud_shared: call b200lp_peer_attach first");
  if (root < 0 || root >= ctx->peer_world) return ctx->fail(B200LP_E_INVALID, "set_cloud_shared: bad root");
  const bool is_root = root == ctx->peer_rank;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->push_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->push_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->push_done, cudaEventDisableTiming));
    CK(ctx->d_push_ticket.reserve(1));
    CK(ctx->d_share_err.reserve(1));
    CK(ctx->h_share_hdr.reserve(1));
    CK(ctx->h_share_err.reserve(1));
    memset(ctx->h_share_hdr.p, 0, sizeof(CloudHeader));
    CK(cudaMemsetAsync(ctx->d_push_ticket.p, 0, sizeof(unsigned), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_share_err.p, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  ++ctx->cloud_seq;  // before anything can fail: the ranks stay in step
  ctx->cloud_timing_pending = false;
  if (is_root) {
    if ((n && !pts) || stride < 12 || (stride & 3)) return ctx->fail(B200LP_E_INVALID, "set_cloud_shared: bad pointer or stride");
    if (n > ctx->peer_cloud_cap)
      return ctx->fail(B200LP_E_INVALID, "set_cloud_shared: %zu points exceed the %zu reserved with b200lp_peer_reserve_cloud", n, ctx->peer_cloud_cap);
    if (ctx->peer_world == 1) return set_cloud_common(ctx, pts, n, stride, false);  // nobody to share with
    CK(ctx->d_raw.reserve(std::max<size_t>(n * 12, 16)));
    ctx->raw_stride = 12;
    if (n == 0) {  // an empty cloud: header and piece flags only
      CloudHeader H{};
      H.seq = ctx->cloud_seq;
      share_wait_acks_kernel<<<1, 32, 0, ctx->push_stream>>>(ctx->my_header(), ctx->peer_world, ctx->peer_rank, ctx->cloud_seq - 1,
                                                             (long long)4e9, ctx->d_share_err.p);
      share_header_kernel<<<1, 32, 0, ctx->push_stream>>>(H, ctx->peer_blocks, ctx->peer_world, ctx->peer_rank);
      for (int c = 0; c < kPackChunks; ++c)
        share_push_kernel<<<1, 256, 0, ctx->push_stream>>>(ctx->d_raw.p, 0, 0, ctx->peer_blocks, ctx->peer_world, ctx->peer_rank, c,
                                                           ctx->cloud_seq, ctx->d_push_ticket.p);
      ctx->launches += 2 + kPackChunks;
    }
    int rc = build_grid(ctx, n ? pts : nullptr, n, stride, false, n ? root : -1);
    if (rc) return rc;
    ctx->last_h2d_bytes = n * 12;
  } else {
    // receive: the root's header first (the host sizes the grid with it), then the pieces, each behind its flag
    share_wait_header_kernel<<<1, 32, 0, ctx->stream>>>(ctx->my_header(), ctx->cloud_seq, ctx->h_share_hdr.p, (long long)4e9);
    ++ctx->launches;
    CK(cudaGetLastError());
    volatile unsigned long long* flag = &ctx->h_share_hdr.p->seq;
    unsigned spins = 0;
    while (*flag != ctx->cloud_seq) {
      __builtin_ia32_pause();
      if ((++spins & 0xfffu) == 0u) {
        const cudaError_t qe = cudaStreamQuery(ctx->stream);
        if (qe == cudaErrorNotReady) continue;
        if (qe != cudaSuccess) return ctx->fail(B200LP_E_CUDA, "set_cloud_shared: %s", cudaGetErrorString(qe));
        if (*flag != ctx->cloud_seq) return ctx->fail(B200LP_E_CUDA, "set_cloud_shared: the header kernel finished without publishing");
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const CloudHeader H = *ctx->h_share_hdr.p;
    if (H.n == ~0ull)
      return ctx->fail(B200LP_E_STATE, "set_cloud_shared: the root did not deliver the cloud within the time limit "
                                       "(call b200lp_peer_resync on every rank before the next cycle)");
    if (H.n > ctx->peer_cloud_cap)
      return ctx->fail(B200LP_E_INVALID, "set_cloud_shared: the root's %llu points exceed the %zu reserved here", H.n, ctx->peer_cloud_cap);
    GridDev& g = ctx->grid;
    g.n_raw = (uint32_t)H.n;
    g.n_kept = 0;
    g.nx = g.ny = g.nz = 1;
    g.org[0] = g.org[1] = g.org[2] = 0.f;
    g.inv_xy = g.inv_z = 1.f;
    g.cmax = 1.f;
    CK(ctx->d_total.reserve(1));
    CK(cudaEventRecord(ctx->cev[0], ctx->stream));
    ctx->pack_threads_used = 0;
    const size_t n_cells = size_grid(ctx, H.mn, H.mx, (size_t)H.n_finite);
    size_t bound[kPackChunks + 1];
    for (int c = 0; c <= kPackChunks; ++c) bound[c] = (size_t)H.bound[c];
    int rc = grid_tail(ctx, ctx->my_rows(), 12, (size_t)H.n, n_cells, bound, ctx->my_header(), ctx->cloud_seq);
    if (rc) return rc;
    share_ack_kernel<<<1, 32, 0, ctx->stream>>>(ctx->peer_blocks.base[root], ctx->peer_rank, ctx->cloud_seq);
    ++ctx->launches;
    ctx->last_h2d_bytes = 0;
    ctx->raw_stride = 12;
  }
  CK(cudaEventRecord(ctx->cev[2], ctx->stream));
  ctx->cloud_timing_pending = true;
  ctx->have_cycle = false;
  return B200LP_OK;
}

int b200lp_last_upload(const b200lp_ctx* ctx, size_t* h2d_bytes, int32_t* pack_threads) {
  if (!ctx || !ctx->have_cloud) return B200LP_E_STATE;
  if (h2d_bytes) *h2d_bytes = ctx->last_h2d_bytes;
  if (pack_threads) *pack_threads = ctx->pack_threads_used;
  return B200LP_OK;
}

int b200lp_set_plan(b200lp_ctx* ctx, const double* p, size_t n) {
  if (!ctx) return B200LP_E_INVALID;
  if (n && !p) return ctx->fail(B200LP_E_INVALID, "set_plan: null plan");
  if (n > B200LP_MAX_PLAN) return ctx->fail(B200LP_E_INVALID, "set_plan: more than B200LP_MAX_PLAN poses");
  ctx->plan_host.assign(p, p + n * 7);
  ctx->plan_on_device = false;
  // upload now, so that the cycle itself starts with its first kernel: the copy is ordered on the main stream, a prep
  // kernel on the second stream waits for plan_ev
  ctx->plan_uploaded = false;
  if (n) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->plan_ev) CK(cudaEventSynchronize(ctx->plan_ev));  // the staging buffer may still feed the previous upload
    CK(ctx->h_plan7.reserve(n * 7));
    CK(ctx->d_plan7.reserve(std::max<size_t>(n * 7, (size_t)B200LP_MAX_PLAN * 7)));
    memcpy(ctx->h_plan7.p, p, n * 7 * sizeof(double));
    CK(cudaMemcpyAsync(ctx->d_plan7.p, ctx->h_plan7.p, n * 7 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->plan_ev) CK(cudaEventCreateWithFlags(&ctx->plan_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->plan_ev, ctx->stream));
    ctx->plan_uploaded = true;
    ctx->plan_ev_pending = true;
  }
  return B200LP_OK;
}

static int plan_shard_common(b200lp_ctx* ctx, const b200lp_query* q, int rank, int count, b200lp_result* out, bool exchange) {
  if (!ctx) return B200LP_E_INVALID;
  if (!q || !out || count < 1 || rank < 0 || rank >= count) return ctx->fail(B200LP_E_INVALID, "plan: bad argument");
  CK(cudaSetDevice(ctx->device));
  const bool resident = ctx->plan_on_device;
  const size_t np = resident ? (size_t)ctx->plan_n_device : ctx->plan_host.size() / 7;
  CK(ctx->h_robots.reserve(1));
  CK(ctx->h_plan7.reserve(std::max<size_t>(np * 7, 7)));
  fill_robot(ctx->h_robots.p, q, 0, (int32_t)np, (np && !resident) ? ctx->plan_host.data() + (np - 1) * 7 : nullptr);
  if (np && !resident && !ctx->plan_uploaded) memcpy(ctx->h_plan7.p, ctx->plan_host.data(), np * 7 * sizeof(double));
  return run_cycle(ctx, 1, rank, count, out, resident, exchange);
}

int b200lp_plan_shard(b200lp_ctx* ctx, const b200lp_query* q, int rank, int count, b200lp_result* out) {
  return plan_shard_common(ctx, q, rank, count, out, false);
}

int b200lp_peer_export(b200lp_ctx* ctx, uint8_t handle[B200LP_PEER_HANDLE_BYTES]) {
  if (!ctx) return B200LP_E_INVALID;
  if (!handle) return ctx->fail(B200LP_E_INVALID, "peer_export: null handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == B200LP_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  static_assert(kMaxPeers == B200LP_MAX_PEERS, "peer table size");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->d_peer_block.p) {
    const size_t bytes = kPeerRowsOff + ((ctx->peer_cloud_cap * 12 + 255) & ~(size_t)255);
    CK(ctx->d_peer_block.reserve(bytes));
    CK(cudaMemsetAsync(ctx->d_peer_block.p, 0, kPeerRowsOff, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->d_peer_block.p));
  memcpy(handle, &h, sizeof(h));
  return B200LP_OK;
}

static void peer_detach(b200lp_ctx* ctx) {
  for (int r = 0; r < kMaxPeers; ++r) {
    if (ctx->peer_opened[r]) cudaIpcCloseMemHandle(ctx->peer_opened[r]);
    ctx->peer_opened[r] = nullptr;
    ctx->peer_table.slots[r] = nullptr;
    ctx->peer_blocks.base[r] = nullptr;
  }
  ctx->peer_rank = -1;
  ctx->peer_world = 0;
}

int b200lp_peer_attach(b200lp_ctx* ctx, int rank, int world, const uint8_t* handles) {
  if (!ctx) return B200LP_E_INVALID;
  if (!handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return ctx->fail(B200LP_E_INVALID, "peer_attach: bad rank / world (at most %d peers)", kMaxPeers);
  if (!ctx->d_peer_block.p) return ctx->fail(B200LP_E_STATE, "peer_attach: call b200lp_peer_export first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  peer_detach(ctx);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ctx->peer_table.slots[r] = ctx->peer_slots();
      ctx->peer_blocks.base[r] = ctx->d_peer_block.p;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * B200LP_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      peer_detach(ctx);
      (void)cudaGetLastError();
      return ctx->fail(B200LP_E_CUDA, "peer_attach: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
    }
    ctx->peer_opened[r] = p;
    ctx->peer_table.slots[r] = (PeerSlot*)p;
    ctx->peer_blocks.base[r] = (char*)p;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  ctx->cuts = ShardCuts{};
  // The slots and the cloud header were zeroed when the block was allocated (b200lp_peer_export), i.e. before any peer
  // could know its handle: a FIRST attach must not clear them again — a faster peer may already have stored its first
  // result. Only a ctx that has exchanged before starts over (then the application needs a barrier between this call
  // and the first exchange, as for b200lp_peer_resync).
  if (ctx->peer_seq != 0 || ctx->cloud_seq != 0) {
    ctx->peer_seq = 0;
    ctx->cloud_seq = 0;
    CK(cudaMemsetAsync(ctx->d_peer_block.p, 0, kPeerRowsOff, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return B200LP_OK;
}

int b200lp_plan_shard_exchange(b200lp_ctx* ctx, const b200lp_query* q, b200lp_result* out) {
  if (!ctx) return B200LP_E_INVALID;
  if (ctx->peer_world < 1) return ctx->fail(B200LP_E_STATE, "plan_shard_exchange: call b200lp_peer_attach first");
  ++ctx->peer_seq;  // before anything can fail: a rank whose cycle errors out early stays in step with its peers
  return plan_shard_common(ctx, q, ctx->peer_rank, ctx->peer_world, out, true);
}

int b200lp_peer_resync(b200lp_ctx* ctx) {
  if (!ctx) return B200LP_E_INVALID;
  if (ctx->peer_world < 1) return ctx->fail(B200LP_E_STATE, "peer_resync: call b200lp_peer_attach first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->peer_seq = 0;
  ctx->cuts = ShardCuts{};
  ctx->cloud_seq = 0;
  if (ctx->push_stream) CK(cudaStreamSynchronize(ctx->push_stream));
  CK(cudaMemsetAsync(ctx->d_peer_block.p, 0, kPeerRowsOff, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200LP_OK;
}

int b200lp_set_shard_cuts(b200lp_ctx* ctx, const float* shares, int count) {
  if (!ctx) return B200LP_E_INVALID;
  if (count < 0 || count > B200LP_MAX_PEERS || (count && !shares)) return ctx->fail(B200LP_E_INVALID, "set_shard_cuts: bad count");
  ShardCuts c{};
  c.n = count;
  for (int k = 0; k <= count && count; ++k) {
    if (!(shares[k] >= 0.f && shares[k] <= 1.f) || (k && shares[k] < shares[k - 1]))
      return ctx->fail(B200LP_E_INVALID, "set_shard_cuts: shares must rise from 0 to 1");
    c.frac[k] = shares[k];
  }
  if (count && (c.frac[0] != 0.f || c.frac[count] != 1.f)) return ctx->fail(B200LP_E_INVALID, "set_shard_cuts: shares must rise from 0 to 1");
  ctx->cuts = c;
  return B200LP_OK;
}

int b200lp_get_shard_cuts(const b200lp_ctx* ctx, float* shares, int* count) {
  if (!ctx || !shares || !count) return B200LP_E_INVALID;
  *count = ctx->cuts.n;
  for (int k = 0; k <= ctx->cuts.n; ++k) shares[k] = ctx->cuts.frac[k];
  return B200LP_OK;
}

int b200lp_set_adaptive_cuts(b200lp_ctx* ctx, int on) {
  if (!ctx) return B200LP_E_INVALID;
  ctx->adaptive_cuts = on != 0;
  return B200LP_OK;
}

int b200lp_last_cycle_ns(const b200lp_ctx* ctx, uint32_t* cycle_ns, uint32_t* peer_ns) {
  if (!ctx || !ctx->have_cycle) return B200LP_E_STATE;
  if (cycle_ns) *cycle_ns = ctx->last_cycle_ns;
  if (peer_ns) memcpy(peer_ns, ctx->last_peer_ns, sizeof(ctx->last_peer_ns));
  return B200LP_OK;
}

int b200lp_plan(b200lp_ctx* ctx, const b200lp_query* q, b200lp_result* out) { return b200lp_plan_shard(ctx, q, 0, 1, out); }

int b200lp_plan_batch(b200lp_ctx* ctx, const b200lp_query* qs, size_t n_robots, const double* plans,
                      const int64_t* plan_offsets, b200lp_result* outs) {
  if (!ctx) return B200LP_E_INVALID;
  if (!qs || !outs || !n_robots || !plan_offsets) return ctx->fail(B200LP_E_INVALID, "plan_batch: bad argument");
  if (n_robots > 65535) return ctx->fail(B200LP_E_INVALID, "plan_batch: more than 65535 robots per call");
  CK(cudaSetDevice(ctx->device));
  const int64_t total = plan_offsets[n_robots];
  if (total < 0 || (total && !plans)) return ctx->fail(B200LP_E_INVALID, "plan_batch: bad plan table");
  CK(ctx->h_robots.reserve(n_robots));
  CK(ctx->h_plan7.reserve(std::max<size_t>((size_t)total * 7, 7)));
  for (size_t i = 0; i < n_robots; ++i) {
    const int64_t a = plan_offsets[i], b = plan_offsets[i + 1];
    if (a < 0 || b < a || b > total || b - a > B200LP_MAX_PLAN) return ctx->fail(B200LP_E_INVALID, "plan_batch: bad plan offsets for robot %zu", i);
    fill_robot(ctx->h_robots.p + i, qs + i, a, (int32_t)(b - a), b > a ? plans + (size_t)(b - 1) * 7 : nullptr);
  }
  // A plan table that already sits in page-locked memory is uploaded from where it is (the call returns after the cycle, so
  // the caller's buffer outlives the copy); anything else goes through the pinned staging buffer first — for 512 robots x 60
  // poses that memcpy alone is ~0.1 ms of a 1.1 ms step.
  ctx->plan_src = nullptr;
  if (total) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, plans) == cudaSuccess && at.type == cudaMemoryTypeHost) ctx->plan_src = plans;
    else cudaGetLastError();  // (an unregistered pointer is not an error of this call)
    if (!ctx->plan_src) memcpy(ctx->h_plan7.p, plans, (size_t)total * 7 * sizeof(double));
  }
  ctx->plan_on_device = false;  // the fleet call brings its own plans; a device-side prune plan does not survive it
  ctx->plan_uploaded = false;   // ... and neither does the uploaded copy of the single-robot plan (plan_host does)
  const int rc = run_cycle(ctx, n_robots, 0, 1, outs);
  ctx->plan_src = nullptr;
  return rc;
}

int b200lp_traj_count(const b200lp_ctx* ctx, size_t robot, int32_t* n_traj_global, int32_t* t_begin, int32_t* t_end) {
  if (!ctx || !ctx->have_cycle || robot >= ctx->n_robots) return B200LP_E_STATE;
  const RobotMeta& m = ctx->meta_host[robot];
  if (n_traj_global) *n_traj_global = m.n_traj;
  if (t_begin) *t_begin = m.t_begin;
  if (t_end) *t_end = m.t_end;
  return B200LP_OK;
}

int b200lp_read_trajectories(b200lp_ctx* ctx, size_t robot, const b200lp_traj_view* v) {
  if (!ctx) return B200LP_E_INVALID;
  if (!v) return ctx->fail(B200LP_E_INVALID, "read_trajectories: null view");
  if (!ctx->have_cycle || robot >= ctx->n_robots) return ctx->fail(B200LP_E_STATE, "read_trajectories: no plan result for that robot");
  CK(cudaSetDevice(ctx->device));
  const RobotMeta& m = ctx->meta_host[robot];
  const size_t n = (size_t)m.n_traj, off = robot * (size_t)ctx->t_cap;
  const int nc = ctx->C.n_critics;
  if (!n) return B200LP_OK;
  std::vector<float4> vel;
  if (v->vel) {
    vel.resize(n);
    CK(cudaMemcpyAsync(vel.data(), ctx->d_rec_vel.p + off, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (v->sample_index) CK(cudaMemcpyAsync(v->sample_index, ctx->d_rec_sample.p + off, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->num_steps) CK(cudaMemcpyAsync(v->num_steps, ctx->d_rec_steps.p + off, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->time_delta) CK(cudaMemcpyAsync(v->time_delta, ctx->d_rec_dt.p + off, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->cost) CK(cudaMemcpyAsync(v->cost, ctx->d_cost.p + off, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->first_hit_pose) CK(cudaMemcpyAsync(v->first_hit_pose, ctx->d_first_hit.p + off, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->critic_scores && nc) CK(cudaMemcpyAsync(v->critic_scores, ctx->d_scores.p + off * nc, n * nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (v->vel)
    for (size_t i = 0; i < n; ++i) {
      v->vel[i * 3] = vel[i].x;
      v->vel[i * 3 + 1] = vel[i].y;
      v->vel[i * 3 + 2] = vel[i].z;
    }
  // trajectories outside a sample shard were not scored in this launch: mark them
  if (m.t_begin > 0 || m.t_end < m.n_traj) {
    const double nan = std::nan("");
    for (size_t i = 0; i < n; ++i) {
      if ((int)i >= m.t_begin && (int)i < m.t_end) continue;
      if (v->cost) v->cost[i] = nan;
      if (v->first_hit_pose) v->first_hit_pose[i] = -1;
      if (v->critic_scores)
        for (int k = 0; k < nc; ++k) v->critic_scores[i * nc + k] = nan;
    }
  }
  return B200LP_OK;
}

int b200lp_read_pose_batch(b200lp_ctx* ctx, size_t robot, int32_t t0, int32_t t1, int64_t* pose_offsets,
                           const b200lp_pose_view* v, size_t capacity_poses) {
  if (!ctx) return B200LP_E_INVALID;
  if (!v) return ctx->fail(B200LP_E_INVALID, "read_pose_batch: null view");
  if (!ctx->have_cycle || robot >= ctx->n_robots) return ctx->fail(B200LP_E_STATE, "read_pose_batch: no plan result for that robot");
  const RobotMeta& m = ctx->meta_host[robot];
  if (t0 < 0 || t1 < t0 || t1 > m.n_traj) return ctx->fail(B200LP_E_INVALID, "read_pose_batch: trajectory range out of bounds");
  const size_t nt = (size_t)(t1 - t0);
  if (pose_offsets) pose_offsets[0] = 0;
  if (!nt) return B200LP_OK;
  CK(cudaSetDevice(ctx->device));
  // exclusive prefix sum of num_steps -> output row of every trajectory
  std::vector<int> steps(nt);
  CK(cudaMemcpyAsync(steps.data(), ctx->d_rec_steps.p + robot * (size_t)ctx->t_cap + t0, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<long long> off(nt + 1);
  off[0] = 0;
  for (size_t i = 0; i < nt; ++i) off[i + 1] = off[i] + steps[i];
  const size_t n = (size_t)off[nt];
  if (pose_offsets)
    for (size_t i = 0; i <= nt; ++i) pose_offsets[i] = off[i];
  if (n > capacity_poses) return ctx->fail(B200LP_E_INVALID, "read_pose_batch: %zu poses exceed the view capacity %zu", n, capacity_poses);
  if (!n) return B200LP_OK;
  // scratch layout in one allocation (8-byte fields first)
  const size_t b_off = 0, b_pose = b_off + (nt + 1) * 8, b_pcl = b_pose + (v->pose ? n * 56 : 0),
               b_cub = b_pcl + (v->pcl_pose ? n * 12 : 0), b_aabb = b_cub + (v->cuboid ? n * 96 : 0),
               b_nr1 = b_aabb + (v->aabb ? n * 24 : 0), b_col = b_nr1 + (v->n_r1 ? n * 4 : 0),
               total = b_col + (v->collide ? n : 0);
  CK(ctx->d_scratch.reserve(total));
  char* d = ctx->d_scratch.p;
  CK(cudaMemcpyAsync(d + b_off, off.data(), (nt + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  poses_kernel<<<(unsigned)nt, 32, 0, ctx->stream>>>(
      ctx->C, ctx->grid, ctx->d_robots.p, (int)robot, ctx->t_cap, t0, (const long long*)(d + b_off), ctx->d_rec_vel.p,
      ctx->d_rec_steps.p, ctx->d_rec_dt.p, v->pose ? (double*)(d + b_pose) : nullptr, v->pcl_pose ? (float*)(d + b_pcl) : nullptr,
      v->cuboid ? (float*)(d + b_cub) : nullptr, v->aabb ? (float*)(d + b_aabb) : nullptr,
      v->collide ? (unsigned char*)(d + b_col) : nullptr, v->n_r1 ? (int*)(d + b_nr1) : nullptr);
  ++ctx->launches;
  if (v->pose) CK(cudaMemcpyAsync(v->pose, d + b_pose, n * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->pcl_pose) CK(cudaMemcpyAsync(v->pcl_pose, d + b_pcl, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->cuboid) CK(cudaMemcpyAsync(v->cuboid, d + b_cub, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->aabb) CK(cudaMemcpyAsync(v->aabb, d + b_aabb, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->n_r1) CK(cudaMemcpyAsync(v->n_r1, d + b_nr1, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (v->collide) CK(cudaMemcpyAsync(v->collide, d + b_col, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return B200LP_OK;
}

int b200lp_read_poses(b200lp_ctx* ctx, size_t robot, int32_t id, const b200lp_pose_view* v) {
  if (!ctx) return B200LP_E_INVALID;
  if (!ctx->have_cycle || robot >= ctx->n_robots) return ctx->fail(B200LP_E_STATE, "read_poses: no plan result for that robot");
  if (id < 0 || id >= ctx->meta_host[robot].n_traj) return ctx->fail(B200LP_E_INVALID, "read_poses: trajectory id out of range");
  return b200lp_read_pose_batch(ctx, robot, id, id + 1, nullptr, v, (size_t)B200LP_MAX_STEPS);
}

int b200lp_set_global_plan(b200lp_ctx* ctx, const double* p, size_t n) {
  if (!ctx) return B200LP_E_INVALID;
  if (!p) return ctx->fail(B200LP_E_INVALID, "set_global_plan: null plan");
  if (n < 3) return ctx->fail(B200LP_E_INVALID, "set_global_plan: size of global plan is smaller than 3");  // local_planner.cpp:324-327
  if (n > 0x7fffffffull / 7) return ctx->fail(B200LP_E_INVALID, "set_global_plan: plan too long");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_gplan7.reserve(n * 7));
  CK(cudaMemcpyAsync(ctx->d_gplan7.p, p, n * 7 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the caller may free its buffer on return
  ctx->n_gplan = n;
  return B200LP_OK;
}

int b200lp_prune_plan(b200lp_ctx* ctx, const double robot_xyz[3], double forward_distance, double backward_distance,
                      b200lp_prune_info* out) {
  TraceRange trace_range("b200lp prune plan");
  if (!ctx) return B200LP_E_INVALID;
  if (!robot_xyz || !out) return ctx->fail(B200LP_E_INVALID, "prune_plan: null argument");
  if (ctx->n_gplan < 3) {  // prunePlan's first early return (:376-377): nothing changes
    out->status = 1; out->nearest_index = -1;
    out->n_prune = ctx->plan_on_device ? ctx->plan_n_device : (int32_t)(ctx->plan_host.size() / 7);
    out->n_backward = 0;
    return B200LP_OK;
  }
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_plan7.reserve((size_t)B200LP_MAX_PLAN * 7));
  CK(ctx->d_prune_pcl.reserve(B200LP_MAX_PLAN));
  CK(ctx->d_prune_meta.reserve(1));
  CK(ctx->h_prune_meta.reserve(1));
  prune_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_gplan7.p, (int)ctx->n_gplan, robot_xyz[0], robot_xyz[1], robot_xyz[2],
                                           forward_distance, backward_distance, B200LP_MAX_PLAN, ctx->d_plan7.p,
                                           ctx->d_prune_pcl.p, ctx->d_prune_meta.p);
  ++ctx->launches;
  CK(cudaMemcpyAsync(ctx->h_prune_meta.p, ctx->d_prune_meta.p, sizeof(PruneMeta), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  const PruneMeta m = *ctx->h_prune_meta.p;
  if (m.overflow) return ctx->fail(B200LP_E_INVALID, "prune_plan: %d poses exceed B200LP_MAX_PLAN=%d", m.info.n_prune, B200LP_MAX_PLAN);
  *out = m.info;
  if (m.info.status == 2) out->n_prune = 0;
  ctx->plan_on_device = true;
  ctx->plan_n_device = out->n_prune;
  ctx->have_prune = true;
  return B200LP_OK;
}

int b200lp_read_prune_plan(b200lp_ctx* ctx, double* poses7, float* pcl_xyzi, size_t capacity) {
  if (!ctx) return B200LP_E_INVALID;
  if (!ctx->have_prune || !ctx->plan_on_device) return ctx->fail(B200LP_E_STATE, "read_prune_plan: no device-side prune plan");
  const size_t n = (size_t)ctx->plan_n_device;
  if (n > capacity) return ctx->fail(B200LP_E_INVALID, "read_prune_plan: %zu poses exceed the capacity %zu", n, capacity);
  if (!n) return B200LP_OK;
  CK(cudaSetDevice(ctx->device));
  if (poses7) CK(cudaMemcpyAsync(poses7, ctx->d_plan7.p, n * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (pcl_xyzi) CK(cudaMemcpyAsync(pcl_xyzi, ctx->d_prune_pcl.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200LP_OK;
}

int b200lp_path_blocked(b200lp_ctx* ctx, double check_radius, b200lp_blocked* out) {
  TraceRange trace_range("b200lp path blocked");
  if (!ctx) return B200LP_E_INVALID;
  if (!out) return ctx->fail(B200LP_E_INVALID, "path_blocked: null argument");
  if (!ctx->have_prune || !ctx->plan_on_device) return ctx->fail(B200LP_E_STATE, "path_blocked: no device-side prune plan");
  if (!ctx->have_cloud) return ctx->fail(B200LP_E_STATE, "path_blocked: no cloud");
  b200lp_blocked B{};
  B.n_total = ctx->plan_n_device;
  // selfMark's guard: `points.size() <= 5 || pcl_prune_plan_.points.size() <= 0` => ratio 0 (path_blocked_strategy.cpp:62-64)
  if (!(ctx->grid.n_raw <= 5 || B.n_total <= 0)) {
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_blocked.reserve(2));
    CK(ctx->h_blocked.reserve(2));
    CK(cudaMemsetAsync(ctx->d_blocked.p, 0, 2 * sizeof(int), ctx->stream));
    const float r2 = (float)(check_radius * check_radius);  // what pcl::KdTreeFLANN::radiusSearch hands to FLANN
    blocked_kernel<<<(unsigned)((B.n_total * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->grid, ctx->d_prune_pcl.p, B.n_total, (float)std::fabs(check_radius), r2, ctx->d_blocked.p, ctx->d_blocked.p + 1);
    ++ctx->launches;
    CK(cudaMemcpyAsync(ctx->h_blocked.p, ctx->d_blocked.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    B.n_blocked = ctx->h_blocked.p[0];
    B.n_checked = ctx->h_blocked.p[1];
    const float orig = (float)B.n_total, blocked = (float)B.n_blocked;  // :91-93: float division, then * 100.0 in double
    B.ratio = (blocked) / (orig) * 100.0;
  }
  B.opinion = B.ratio > 0.0 ? 1 : 0;
  *out = B;
  return B200LP_OK;
}

// tf2::transformToEigen(TransformStamped) = Translation3d * Quaterniond(w, x, y, z) as a row-major 3x4 (no normalisation,
// Eigen's toRotationMatrix operation order; the same expression as lp::quat_to_matrix on the device and the oracle)
static void transform_to_rows(const double p[7], double m[12]) {
  const double x = p[3], y = p[4], z = p[5], w = p[6];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  m[0] = 1.0 - (tyy + tzz); m[1] = txy - twz;         m[2] = txz + twy;          m[3] = p[0];
  m[4] = txy + twz;         m[5] = 1.0 - (txx + tzz); m[6] = tyz - twx;          m[7] = p[1];
  m[8] = txz - twy;         m[9] = tyz + twx;         m[10] = 1.0 - (txx + tyy); m[11] = p[2];
}

int b200lp_sensor_observation(b200lp_ctx* ctx, int sensor, const void* scan, size_t n, size_t stride,
                              const double base_from_sensor[7], const double global_from_base[7],
                              const b200lp_sensor_params* sp, b200lp_observation_info* info) {
  TraceRange trace_range("b200lp sensor observation");
  if (!ctx) return B200LP_E_INVALID;
  if (sensor < 0 || sensor >= B200LP_MAX_SENSORS) return ctx->fail(B200LP_E_INVALID, "sensor_observation: sensor index out of range");
  if (!sp || !base_from_sensor || !global_from_base || (n && !scan) || stride < 12 || (stride & 3))
    return ctx->fail(B200LP_E_INVALID, "sensor_observation: bad argument");
  if (n > ((size_t)1 << 26)) return ctx->fail(B200LP_E_INVALID, "sensor_observation: more than 2^26 points in one scan");
  ObsDev P{};
  transform_to_rows(base_from_sensor, P.m1);
  transform_to_rows(global_from_base, P.m2);
  P.apply_m2 = sp->is_local_planner ? 1 : 0;
  const float leaf = sp->leaf_size > 0.f ? sp->leaf_size : 0.1f;
  P.inv_leaf = 1.0f / leaf;  // pcl::VoxelGrid::setLeafSize: inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  // pcl::PassThrough::setFilterLimits(const float&, const float&): the doubles are narrowed at the call (:242, :249)
  P.lo[0] = P.lo[1] = (float)(-sp->perception_window_size);
  P.hi[0] = P.hi[1] = (float)(sp->perception_window_size);
  P.lo[2] = 0.0f;
  P.hi[2] = (float)(sp->marking_height);
  double cells = 1.0;
  uint32_t ext[3];
  for (int a = 0; a < 3; ++a) {
    const float flo = std::floor(P.lo[a] * P.inv_leaf), fhi = std::floor(P.hi[a] * P.inv_leaf);
    if (!(std::fabs(flo) < 8388608.f) || !(std::fabs(fhi) < 8388608.f))
      return ctx->fail(B200LP_E_INVALID, "sensor_observation: window / leaf out of range");
    P.lb[a] = (int)flo;
    ext[a] = fhi >= flo ? (uint32_t)((int)fhi - (int)flo + 1) : 1u;
    cells *= (double)ext[a];
  }
  if (cells > (double)(1u << 27))
    return ctx->fail(B200LP_E_INVALID, "sensor_observation: the pass-through window holds %.0f voxels of %.3f m (limit 2^27)", cells, leaf);
  P.d0 = ext[0];
  P.d1 = ext[1];
  int key_bits = 1;
  while (((uint64_t)1 << key_bits) < (uint64_t)cells) ++key_bits;
  const int passes = (key_bits + kObsMaxBits - 1) / kObsMaxBits;
  const int bits = (key_bits + passes - 1) / passes;

  CK(cudaSetDevice(ctx->device));
  ctx->have_obs[sensor] = false;
  ctx->n_obs_out[sensor] = 0;
  b200lp_observation_info I{};
  I.n_scan = (int64_t)n;
  if (n) {
    const bool large = n >= kObsLargeScan;
    const size_t tile = (size_t)kObsThreads * (large ? kObsItemsLarge : kObsItemsSmall);
    const unsigned nb = (unsigned)((n + tile - 1) / tile);
    const size_t hist_n = ((size_t)1 << bits) * nb;
    const unsigned nbk = (unsigned)((hist_n + kScanItems - 1) / kScanItems);
    const unsigned nbh = (unsigned)((n + kObsHeadTile - 1) / kObsHeadTile);
    CK(ctx->d_scan.reserve(n * stride));
    CK(ctx->d_obs_a.reserve(n));
    CK(ctx->d_obs_b.reserve(n));
    CK(ctx->d_obs_hist.reserve(hist_n + 1));
    if (ctx->d_scan_status.cap < (size_t)nbk) {  // (new words carry epoch 0: never the current one)
      CK(ctx->d_scan_status.reserve(nbk));
      CK(cudaMemsetAsync(ctx->d_scan_status.p, 0, ctx->d_scan_status.cap * sizeof(unsigned long long), ctx->stream));
    }
    if (!ctx->d_scan_ticket.p) {
      CK(ctx->d_scan_ticket.reserve(1));
      CK(cudaMemsetAsync(ctx->d_scan_ticket.p, 0, sizeof(unsigned), ctx->stream));
    }
    CK(ctx->d_obs_heads.reserve(nbh));
    CK(ctx->d_obs_counts.reserve(2));
    CK(ctx->h_obs_counts.reserve(2));
    CK(ctx->d_obs_out[sensor].reserve(n));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->oev[0], st));
    CK(cudaMemcpyAsync(ctx->d_scan.p, scan, n * stride, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->oev[2], st));
    if (large) obs_key_kernel<kObsItemsLarge><<<nb, kObsThreads, 0, st>>>(ctx->d_scan.p, n, stride, P, bits, ctx->d_obs_a.p, ctx->d_obs_hist.p);
    else obs_key_kernel<kObsItemsSmall><<<nb, kObsThreads, 0, st>>>(ctx->d_scan.p, n, stride, P, bits, ctx->d_obs_a.p, ctx->d_obs_hist.p);
    float4 *src = ctx->d_obs_a.p, *dst = ctx->d_obs_b.p;
    int launches = 1;
    for (int p = 0; p < passes; ++p) {
      if (p) {
        if (large) obs_hist_kernel<kObsItemsLarge><<<nb, kObsThreads, 0, st>>>(src, ctx->d_obs_counts.p, p * bits, bits, ctx->d_obs_hist.p);
        else obs_hist_kernel<kObsItemsSmall><<<nb, kObsThreads, 0, st>>>(src, ctx->d_obs_counts.p, p * bits, bits, ctx->d_obs_hist.p);
        ++launches;
      }
      if (++ctx->scan_epoch >= (1u << 30)) ctx->scan_epoch = 1u;
      scan_kernel<<<nbk, 256, 0, st>>>(ctx->d_obs_hist.p, hist_n, ctx->d_scan_status.p, ctx->d_scan_ticket.p, ctx->scan_epoch,
                                       ctx->d_obs_counts.p);  // total = points in the window
      if (large)
        obs_scatter_kernel<kObsItemsLarge><<<nb, kObsThreads, 0, st>>>(src, n, ctx->d_obs_counts.p, p == 0 ? 1 : 0, p * bits, bits,
                                                                       ctx->d_obs_hist.p, dst);
      else
        obs_scatter_kernel<kObsItemsSmall><<<nb, kObsThreads, 0, st>>>(src, n, ctx->d_obs_counts.p, p == 0 ? 1 : 0, p * bits, bits,
                                                                       ctx->d_obs_hist.p, dst);
      launches += 2;
      std::swap(src, dst);
    }
    obs_heads_kernel<<<nbh, kObsThreads, 0, st>>>(src, ctx->d_obs_counts.p, ctx->d_obs_heads.p);
    obs_centroid_kernel<<<nbh, kObsThreads, 0, st>>>(src, ctx->d_obs_counts.p, ctx->d_obs_heads.p, P, ctx->d_obs_out[sensor].p);
    launches += 2;
    ctx->launches += launches;
    I.n_launches = launches;
    CK(cudaMemcpyAsync(ctx->h_obs_counts.p, ctx->d_obs_counts.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->oev[1], st));
    CK(cudaStreamSynchronize(st));  // the caller may free the scan on return, and needs the point count
    CK(cudaGetLastError());
    I.n_window = ctx->h_obs_counts.p[0];
    I.n_points = ctx->h_obs_counts.p[1];
    cudaEventElapsedTime(&I.ms_device, ctx->oev[0], ctx->oev[1]);
    cudaEventElapsedTime(&I.ms_upload, ctx->oev[0], ctx->oev[2]);
  }
  ctx->n_obs_out[sensor] = (size_t)I.n_points;
  ctx->have_obs[sensor] = true;
  if (info) *info = I;
  return B200LP_OK;
}

int b200lp_read_observation(b200lp_ctx* ctx, int sensor, void* out, size_t capacity, size_t stride, size_t* n_points) {
  if (!ctx) return B200LP_E_INVALID;
  if (sensor < 0 || sensor >= B200LP_MAX_SENSORS) return ctx->fail(B200LP_E_INVALID, "read_observation: sensor index out of range");
  if (!ctx->have_obs[sensor]) return ctx->fail(B200LP_E_STATE, "read_observation: sensor %d has no observation", sensor);
  if (stride != 16 && stride != 32) return ctx->fail(B200LP_E_INVALID, "read_observation: stride must be 16 (PointXYZ) or 32 (PointXYZI)");
  const size_t n = ctx->n_obs_out[sensor];
  if (n_points) *n_points = n;
  if (n > capacity) return ctx->fail(B200LP_E_INVALID, "read_observation: %zu points exceed the capacity %zu", n, capacity);
  if (!n) return B200LP_OK;
  if (!out) return ctx->fail(B200LP_E_INVALID, "read_observation: null buffer");
  CK(cudaSetDevice(ctx->device));
  if (stride == 16) {
    CK(cudaMemcpyAsync(out, ctx->d_obs_out[sensor].p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    CK(ctx->d_scratch.reserve(n * 32));
    obs_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_obs_out[sensor].p, n, (float4*)ctx->d_scratch.p);
    ++ctx->launches;
    CK(cudaMemcpyAsync(out, ctx->d_scratch.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return B200LP_OK;
}

int b200lp_aggregate_observations(b200lp_ctx* ctx, const int32_t* sensors, int n_sensors, size_t* n_total) {
  TraceRange trace_range("b200lp aggregate observations");
  if (!ctx) return B200LP_E_INVALID;
  if (n_sensors < 0 || n_sensors > B200LP_MAX_SENSORS || (n_sensors && !sensors))
    return ctx->fail(B200LP_E_INVALID, "aggregate_observations: bad sensor list");
  size_t total = 0;
  for (int k = 0; k < n_sensors; ++k) {
    if (sensors[k] < 0 || sensors[k] >= B200LP_MAX_SENSORS) return ctx->fail(B200LP_E_INVALID, "aggregate_observations: sensor index out of range");
    if (!ctx->have_obs[sensors[k]]) return ctx->fail(B200LP_E_STATE, "aggregate_observations: sensor %d has no observation", sensors[k]);
    total += ctx->n_obs_out[sensors[k]];
  }
  if (total > 0xfffffff0ull) return ctx->fail(B200LP_E_INVALID, "aggregate_observations: more than 2^32 points");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_raw.reserve(std::max<size_t>(total * 16, 16)));
  ctx->raw_stride = 16;
  ctx->cloud_timing_pending = false;
  size_t off = 0;
  for (int k = 0; k < n_sensors; ++k) {  // `*aggregate += *plugin->getObservation()` in plugin order
    const size_t m = ctx->n_obs_out[sensors[k]];
    if (m) CK(cudaMemcpyAsync(ctx->d_raw.p + off * 16, ctx->d_obs_out[sensors[k]].p, m * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    off += m;
  }
  int rc = build_grid(ctx, nullptr, total, 16, false);
  if (rc) return rc;
  ctx->last_h2d_bytes = 0;
  CK(cudaEventRecord(ctx->cev[2], ctx->stream));
  ctx->cloud_timing_pending = true;
  ctx->have_cycle = false;
  if (n_total) *n_total = total;
  return B200LP_OK;
}

int b200lp_count_radius(b200lp_ctx* ctx, int64_t* sum_n_r1, int64_t* n_poses) {
  if (!ctx) return B200LP_E_INVALID;
  if (!ctx->have_cycle) return ctx->fail(B200LP_E_STATE, "count_radius: no plan result");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_count.reserve(2));
  CK(ctx->h_count.reserve(2));
  CK(cudaMemsetAsync(ctx->d_count.p, 0, 16, ctx->stream));
  const int gx = (ctx->t_cap + kWarpsPerCta - 1) / kWarpsPerCta;
  count_radius_kernel<<<dim3(gx, (unsigned)ctx->n_robots), kThreads, 0, ctx->stream>>>(
      ctx->C, ctx->grid, ctx->d_robots.p, ctx->d_meta.p, ctx->t_cap, ctx->d_rec_vel.p, ctx->d_rec_steps.p, ctx->d_rec_dt.p,
      ctx->d_count.p);
  ++ctx->launches;
  CK(cudaMemcpyAsync(ctx->h_count.p, ctx->d_count.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  if (sum_n_r1) *sum_n_r1 = (int64_t)ctx->h_count.p[0];
  if (n_poses) *n_poses = (int64_t)ctx->h_count.p[1];
  return B200LP_OK;
}

int b200lp_last_timing(const b200lp_ctx* ctx, float* ms_upload, float* ms_grid_build, float* ms_plan_kernels,
                       float* ms_readback) {
  if (!ctx) return B200LP_E_INVALID;
  resolve_cycle_timing(const_cast<b200lp_ctx*>(ctx));
  if (ms_upload) *ms_upload = ctx->ms_upload;
  if (ms_grid_build) *ms_grid_build = ctx->ms_grid;
  if (ms_plan_kernels) *ms_plan_kernels = ctx->ms_plan;
  if (ms_readback) *ms_readback = ctx->ms_readback;
  return B200LP_OK;
}

int b200lp_last_kernel_ms(const b200lp_ctx* ctx, float* ms_prep_kernel, float* ms_plan_kernel, float* ms_argmin_kernel) {
  if (!ctx) return B200LP_E_INVALID;
  resolve_cycle_timing(const_cast<b200lp_ctx*>(ctx));
  if (ms_prep_kernel) *ms_prep_kernel = ctx->ms_k_prep;
  if (ms_plan_kernel) *ms_plan_kernel = ctx->ms_k_plan;
  if (ms_argmin_kernel) *ms_argmin_kernel = ctx->ms_k_argmin;
  return B200LP_OK;
}

int b200lp_last_kernel_times(const b200lp_ctx* ctx, float* ms, int n) {
  if (!ctx || !ctx->have_cycle || !ms || n < 1) return B200LP_E_STATE;
  resolve_cycle_timing(const_cast<b200lp_ctx*>(ctx));
  const float v[4] = {ctx->ms_k_prep, ctx->ms_k_cull, ctx->ms_k_plan, ctx->ms_k_argmin};
  for (int k = 0; k < n && k < 4; ++k) ms[k] = v[k];
  return B200LP_OK;
}

int b200lp_work_counters(b200lp_ctx* ctx, uint64_t out[4], int reset) {
  if (!ctx) return B200LP_E_INVALID;
#if B200LP_COUNT
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  unsigned long long v[4] = {0ull, 0ull, 0ull, 0ull};
  if (out) {
    CK(cudaMemcpyFromSymbol(v, g_counters, sizeof(v)));
    for (int k = 0; k < 4; ++k) out[k] = v[k];
  }
  if (reset) {
    const unsigned long long z[4] = {0ull, 0ull, 0ull, 0ull};
    CK(cudaMemcpyToSymbol(g_counters, z, sizeof(z)));
  }
  return B200LP_OK;
#else
  (void)out;
  (void)reset;
  return ctx->fail(B200LP_E_STATE, "work_counters: this is not the counting build (libb200lp_count.so, -DB200LP_COUNT=1)");
#endif
}

int64_t b200lp_launch_count(const b200lp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200lp_grid_info(const b200lp_ctx* ctx, int32_t dims[3], float origin[3], float cell[2], int64_t* n_points_kept) {
  if (!ctx || !ctx->have_cloud) return B200LP_E_STATE;
  if (dims) { dims[0] = ctx->grid.nx; dims[1] = ctx->grid.ny; dims[2] = ctx->grid.nz; }
  if (origin) { origin[0] = ctx->grid.org[0]; origin[1] = ctx->grid.org[1]; origin[2] = ctx->grid.org[2]; }
  if (cell) { cell[0] = ctx->cell_xy_used; cell[1] = ctx->cell_z_used; }
  if (n_points_kept) *n_points_kept = ctx->grid.n_kept;
  return B200LP_OK;
}

void* b200lp_stream(const b200lp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
