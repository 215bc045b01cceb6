// lp_kernels.cuh — the sm_100a kernels of the rollout-and-score path.
//
//   grid build : bounds_pack_kernel (clouds not packed on the host) -> hist_kernel (per upload piece; leaves every point's
//                rank inside its cell) -> scan_kernel (one launch, decoupled look-back) -> scatter_kernel -> sat_y / sat_z
//   plan cycle : prep_kernel (velocity samples, trajectory list in the reference's order, forward simulation, pure-pursuit
//                terms) -> cull_kernel (float pre-cull of every pose against the summed-volume table, work lists) ->
//                plan_kernel (persistent warps: cuboid geometry, obstacle sweep, critics, ordered sum; for one robot also
//                the argmin and — for sample shards — the cross-GPU exchange) -> argmin_kernel (fleets)
//   shared map : share_* kernels (one upload for all ranks of a peer group, rows pushed over NVLink)
//   read-back  : poses_kernel, count_radius_kernel (diagnostics / parity / roofline accounting)
//   either side: prune_kernel, blocked_kernel here; the lidar observation producer in lp_observe.cuh
#pragma once
#include "lp_device.cuh"

namespace lp {

__device__ __forceinline__ bool better(unsigned long long ca, int ia, unsigned long long cb, int ib) {
  // getBestTrajectory: `cost <= minimum_cost` while scanning in id order => min cost, ties -> largest id
  return ca < cb || (ca == cb && ia > ib);
}

__device__ __forceinline__ unsigned __smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// =============================================================================================
// grid build
// =============================================================================================
__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving float -> uint
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

struct BoundsDev {
  unsigned mn[3], mx[3];  // ordered-uint encoded
  unsigned n_finite;
  unsigned pad;
};

__device__ __forceinline__ float3 load_xyz(const char* raw, size_t i, size_t stride) {
  if ((stride & 15) == 0) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(raw + i * stride));
    return make_float3(v.x, v.y, v.z);
  }
  const float* p = reinterpret_cast<const float*>(raw + i * stride);
  return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

__device__ __forceinline__ bool finite3(float3 v) {
  return isfinite(v.x) && isfinite(v.y) && isfinite(v.z);
}

// Pass 1 over points [i0, i1) of the raw cloud (launched per upload chunk, so it overlaps the rest of the upload):
// pack x,y,z + the original index into 16-byte float4 records — every later pass reads these, fully coalesced, instead
// of the 32-byte-stride PointXYZI input — and reduce the bounds of the finite points.
__global__ void __launch_bounds__(256) bounds_pack_kernel(const char* __restrict__ raw, size_t i0, size_t i1, size_t stride,
                                                          float4* __restrict__ packed, BoundsDev* __restrict__ b) {
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  unsigned cnt = 0;
  for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (size_t)gridDim.x * blockDim.x) {
    const float3 v = load_xyz(raw, i, stride);
    packed[i] = make_float4(v.x, v.y, v.z, __uint_as_float((uint32_t)i));
    if (finite3(v)) {
      ++cnt;
      mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
      mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
      mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], o));
    }
    cnt += __shfl_xor_sync(kFull, cnt, o);
  }
  // one set of global atomics per CTA (same-address atomics serialise in L2)
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ unsigned s_cnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_mn[warp][a] = mn[a]; s_mx[warp][a] = mx[a]; }
    s_cnt[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned total = 0;
    for (int w = 0; w < 8; ++w) {
      total += s_cnt[w];
#pragma unroll
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], s_mn[w][a]); mx[a] = fmaxf(mx[a], s_mx[w][a]); }
    }
    if (total) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        atomicMin(&b->mn[a], f2ord(mn[a]));
        atomicMax(&b->mx[a], f2ord(mx[a]));
      }
      atomicAdd(&b->n_finite, total);
    }
  }
}

__device__ __forceinline__ uint32_t cell_key(const GridDev& g, float3 v) {
  const int cx = cell_clamped(v.x, g.org[0], g.inv_xy, g.nx);
  const int cy = cell_clamped(v.y, g.org[1], g.inv_xy, g.ny);
  const int cz = cell_clamped(v.z, g.org[2], g.inv_z, g.nz);
  return (uint32_t)(((size_t)cz * g.ny + cy) * (size_t)g.nx + cx);
}

// histogram of points per cell (the key is three subtract-multiply-floors: cheaper to recompute in the scatter pass than
// to store and re-read)
// Records [i0, i1) of `raw` at `stride` bytes: the 16-byte packed records of bounds_pack_kernel, or the 12-byte x,y,z rows
// of a cloud that was packed on the host (launched per upload chunk then, so it overlaps the rest of the upload).
__global__ void __launch_bounds__(256) hist_kernel(const char* __restrict__ raw, size_t stride, size_t i0, size_t i1, GridDev g,
                                                   uint32_t* __restrict__ counts, uint32_t* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (size_t w0 = i0 + (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); w0 < i1; w0 += (size_t)gridDim.x * blockDim.x) {
    const size_t i = w0 + lane;
    uint32_t key = 0xffffffffu;
    if (i < i1) {
      const float3 v = load_xyz(raw, i, stride);
      if (finite3(v)) key = cell_key(g, v);
    }
    // Neighbouring points of a cloud mostly fall into the same cell (obstacle surfaces sampled at centimetres, cells of
    // decimetres), and same-address atomics serialise in L2: the lanes of a warp that share a cell are found with
    // __match_any_sync and the lowest of them counts all of them with ONE atomic. Its answer — how many points the cell
    // held before — makes every point's RANK inside its cell; the scatter pass then needs no atomics at all
    // (slot = cell_start[cell] + rank), and this pass runs per upload piece underneath the rest of the upload.
    const unsigned peers = __match_any_sync(kFull, key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0u;
    LP_CHECK(key == 0xffffffffu || key < (uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz);
    if (lane == leader && key != 0xffffffffu) base = atomicAdd(counts + key, (uint32_t)__popc(peers));
    base = __shfl_sync(kFull, base, leader);
    if (i < i1) rank[i] = base + (uint32_t)__popc(peers & lt);
  }
}

// exclusive scan of counts[0..n) in place, ONE pass (decoupled look-back, Merrill & Garland): blocks of 2048 items take
// their place in the order by ticket (so a block only ever waits for blocks that are running), publish their aggregate,
// walk back over their predecessors' published words 32 at a time until one carries an inclusive prefix, and publish
// their own. A word is {build epoch : 30, status : 2, value : 32}; the epoch makes last build's words invisible, so
// nothing is cleared between builds.
constexpr int kScanItems = 2048;
constexpr unsigned long long kScanAggregate = 1ull, kScanPrefix = 2ull;
__global__ void __launch_bounds__(256) scan_kernel(uint32_t* __restrict__ data, size_t n, unsigned long long* status,
                                                   unsigned* __restrict__ ticket, unsigned epoch, uint32_t* __restrict__ total_out) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned s_block;
  if (threadIdx.x == 0) s_block = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned block = s_block;
  const size_t base = (size_t)block * kScanItems + (size_t)threadIdx.x * 8;
  uint32_t v[8];
  uint32_t sum = 0;
  if (base + 8 <= n) {
    const uint4 a = *reinterpret_cast<const uint4*>(data + base), b = *reinterpret_cast<const uint4*>(data + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (base + k < n) ? data[base + k] : 0u;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += v[k];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t woff = 0, block_sum = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < warp) woff += s_warp[w];
    block_sum += s_warp[w];
  }
  const unsigned long long tag = (unsigned long long)epoch << 34;
  volatile unsigned long long* st = status;
  if (threadIdx.x == 0) st[block] = tag | ((block == 0 ? kScanPrefix : kScanAggregate) << 32) | block_sum;
  // look back with the whole block: thread t takes predecessor block - 1 - t (256 per trip; up to 256 blocks, i.e. half a
  // million cells, need ONE trip), every warp sums its lanes down to the nearest inclusive prefix, thread 0 joins the warps
  __shared__ uint32_t s_part[8];
  __shared__ int s_found[8];
  uint32_t prefix = 0;
  for (int hi = (int)block - 1; hi >= 0; hi -= 256) {
    const int j = hi - (int)threadIdx.x;
    unsigned long long w = 0ull;
    if (j >= 0) {
      do { w = st[j]; } while ((w >> 34) != (unsigned long long)epoch);  // (its block holds a ticket: it is running)
    }
    const unsigned is_prefix = __ballot_sync(kFull, j >= 0 && ((w >> 32) & 3ull) == kScanPrefix);
    const int stop = is_prefix ? __ffs(is_prefix) - 1 : 32;  // nearest predecessor of this warp with an inclusive prefix
    uint32_t part = (j >= 0 && lane <= stop) ? (uint32_t)w : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
    __syncthreads();  // (s_part / s_found of the previous trip have been read)
    if (lane == 0) { s_part[warp] = part; s_found[warp] = is_prefix ? 1 : 0; }
    __syncthreads();
    bool found = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // (every thread: the same answer, no broadcast needed)
      if (!found) prefix += s_part[k];
      found = found || s_found[k];
    }
    if (found) break;
  }
  if (threadIdx.x == 0) {
    if (block != 0) {
      __threadfence();
      st[block] = tag | (kScanPrefix << 32) | (unsigned long long)(prefix + block_sum);
    }
    if ((size_t)(block + 1) * kScanItems >= n) {  // the last block: cell_start[n_cells], and the ticket for the next build
      *total_out = prefix + block_sum;
      data[n] = prefix + block_sum;
      *ticket = 0u;
    }
  }
  uint32_t run = prefix + woff + incl - sum;
  uint32_t o8[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    o8[k] = run;
    run += v[k];
  }
  if (base + 8 <= n) {
    *reinterpret_cast<uint4*>(data + base) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
    *reinterpret_cast<uint4*>(data + base + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (base + k < n) data[base + k] = o8[k];
  }
}

// counting-sort scatter: slot = cell_start[cell] + the point's rank inside its cell (hist_kernel). The order of points
// inside a cell is whatever order the histogram's atomics were served in; every consumer is an any-hit or a count.
// (A cloud arrives in no particular order, so every 16-byte record lands in a sector of its own: the pass is bound by the
// latency of dependent random accesses, 2 M points take ~33 us. Eight points per thread in block-contiguous tiles were
// slower, 39 us.)
__global__ void __launch_bounds__(256) scatter_kernel(const char* __restrict__ raw, size_t stride, size_t n, GridDev g,
                                                      const uint32_t* __restrict__ rank, float4* __restrict__ out) {
  constexpr int kU = 4;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kU * step) {
    float3 v[kU];
    uint32_t slot[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {  // independent loads first
      const size_t i = i0 + u * step;
      ok[u] = false;
      if (i < n) {
        v[u] = load_xyz(raw, i, stride);
        slot[u] = __ldg(rank + i);
        ok[u] = finite3(v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (ok[u]) slot[u] += __ldg(g.cell_start + cell_key(g, v[u]));
#pragma unroll
    for (int u = 0; u < kU; ++u) LP_CHECK(!ok[u] || slot[u] < g.n_kept);
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (ok[u]) out[slot[u]] = make_float4(v[u].x, v[u].y, v[u].z, __uint_as_float((uint32_t)(i0 + u * step)));  // w = original index
  }
}

// Summed-volume table of the per-cell point counts: sat[Z][Y][X] = points with cell (x,y,z) < (X,Y,Z)
// componentwise, dims (nz+1)(ny+1)(nx+1), zero-initialised by the caller. The x prefix comes straight from
// cell_start; sat_y_kernel accumulates along y, sat_z_kernel along z. Consecutive threads walk consecutive X.
// one warp per (X, z): the lanes scan 32 consecutive y at a time. (Both passes are bound by L2 transactions, not by latency:
// the lanes of a warp read one 32-byte sector each — issuing all loads of a warp before the first scan changed nothing,
// 0.0520 vs 0.0522 ms for the whole tail at C2, and neither did a tiled version with the lanes along x and coalesced
// 128-byte reads (0.0541 vs 0.0542 ms): at this size the tail is five dependent launches of ~10 us each whatever they do.
// B200LP_TAIL_TRACE=1 times the kernels of the tail one by one.)
__global__ void __launch_bounds__(256) sat_y_kernel(GridDev g, uint32_t* __restrict__ sat) {
  const size_t nxp = (size_t)g.nx + 1;
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= nxp * (size_t)g.nz) return;
  const int X = (int)(wid % nxp), z = (int)(wid / nxp);
  uint32_t carry = 0u;
  for (int y0 = 0; y0 < g.ny; y0 += 32) {
    const int y = y0 + lane;
    uint32_t v = 0u;
    if (y < g.ny) {
      const size_t row = ((size_t)z * g.ny + y) * (size_t)g.nx;
      v = g.cell_start[row + X] - g.cell_start[row];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, v, o);
      if (lane >= o) v += t;
    }
    if (y < g.ny) sat[((size_t)(z + 1) * (g.ny + 1) + (y + 1)) * nxp + X] = carry + v;
    carry += __shfl_sync(kFull, v, 31);
  }
}
__global__ void __launch_bounds__(256) sat_z_kernel(GridDev g, uint32_t* __restrict__ sat) {
  const size_t plane = ((size_t)g.nx + 1) * ((size_t)g.ny + 1);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  uint32_t run = 0u;
  for (int Z = 1; Z <= g.nz; ++Z) {
    run += sat[(size_t)Z * plane + i];
    sat[(size_t)Z * plane + i] = run;
  }
}

// =============================================================================================
// prep: velocity sampling (initialise()) and the generated-trajectory list (generateTrajectory's
// early rejects, num_steps, dt), one CTA per robot.
// =============================================================================================
// trajectory_generators::VelocityIterator (velocity_iterator.h:44-69); values are rounded to float where
// the reference stores them into an Eigen::Vector3f (dd_simple…cpp:282-286).
// One lane runs it — or the host, for single-robot launches (plan_samples in b200lp.cu): the only serial dependency is
// the `next += step` chain (b200lp_create bounds n so that every index below stays inside out[kMaxAxis]). Plain IEEE
// double adds and compares: host and device produce the same bits.
__host__ __device__ inline int velocity_iterator_dev(double mn, double mx, int n, float* out) {
  if (mn == mx) {
    out[0] = (float)mn;
    return 1;
  }
  n = n < 2 ? 2 : n;
  const double step = (mx - mn) / (double)(n - 1);
  double next = mn;
  int cnt = 0;
  for (int j = 0; j < n - 1; ++j) {
    const double cur = next;
    next += step;
    out[cnt] = (float)cur;
    out[cnt + 1] = 0.0f;  // kept only when the window crosses zero here, else overwritten by the next value
    cnt += (cur < 0 && next > 0) ? 2 : 1;
  }
  out[cnt] = (float)mx;
  return cnt + 1 < kMaxAxis ? cnt + 1 : kMaxAxis;
}

// The velocity window of a cycle (dd_simple…cpp:255-277; omni…cpp:277-303); Vector3f stores round to float.
// mn / mx: linear x, linear y, angular z. IEEE double arithmetic only, so the host evaluates it bit for bit like a CTA.
__host__ __device__ inline void velocity_window(const b200lp_limits& L, const b200lp_params& P, const RobotIn& q, float* mn,
                                                float* mx) {
  const double max_vel_th = L.max_vel_theta, min_vel_th = -1.0 * max_vel_th;
  const float acc0 = (float)L.acc_lim_x, acc1 = (float)L.acc_lim_y, acc2 = (float)L.acc_lim_theta;
  const double sim_period = 1.0 / P.controller_frequency;
  const double tx = q.twist[0], ty = q.twist[1], tw = q.twist[2];
  double min_vel_x = L.min_vel_x, max_vel_x = L.max_vel_x;
  float mx0, mn0, mx1 = 0.f, mn1 = 0.f, mx2, mn2;
  if (P.theory == B200LP_THEORY_DD_SIMPLE) {
    if (q.max_speed_override > 0.0) max_vel_x = fmin(max_vel_x, q.max_speed_override);
    mx0 = (float)fmin(max_vel_x, tx + (double)acc0 * sim_period);
    mx2 = (float)fmin(max_vel_th, tw + (double)acc2 * sim_period);
    mn0 = (float)fmax(min_vel_x, tx / L.deceleration_ratio);
    mn2 = (float)fmax(min_vel_th, tw - (double)acc2 * sim_period);
    if (mx0 < mn0) {
      mn0 = (float)(tx / L.deceleration_ratio);
      mx0 = (float)(tx / L.deceleration_ratio);
    }
  } else {
    const double min_vel_y = L.min_vel_y, max_vel_y = L.max_vel_y;
    mx0 = (float)fmin(max_vel_x, tx + (double)acc0 * sim_period);
    mx1 = (float)fmin(max_vel_y, ty + (double)acc1 * sim_period);
    mx2 = (float)fmin(max_vel_th, tw + (double)acc2 * sim_period);
    mn0 = (float)fmax(min_vel_x, tx - (double)acc0 * sim_period);
    mn1 = (float)fmax(min_vel_y, ty - (double)acc1 * sim_period);
    mn2 = (float)fmax(min_vel_th, tw - (double)acc2 * sim_period);
    if (tx >= max_vel_x / L.deceleration_ratio) mn0 = (float)fmax(min_vel_x, tx / L.deceleration_ratio);
    else if (tx <= min_vel_x / L.deceleration_ratio) mx0 = (float)fmin(max_vel_x, tx / L.deceleration_ratio);
    if (ty >= max_vel_y / L.deceleration_ratio) mn1 = (float)fmax(min_vel_y, ty / L.deceleration_ratio);
    else if (ty <= min_vel_y / L.deceleration_ratio) mx1 = (float)fmin(max_vel_y, ty / L.deceleration_ratio);
  }
  mn[0] = mn0; mn[1] = mn1; mn[2] = mn2;
  mx[0] = mx0; mx[1] = mx1; mx[2] = mx2;
}

__device__ __forceinline__ bool motor_ok(const b200lp_limits& L, float v0, float v2) {
  const double vr = (double)v0 + L.robot_radius * (double)v2;
  const double vl = (double)v0 - L.robot_radius * (double)v2;
  const double rpm_r = vr * L.gear_ratio * 60. / 3.1415926 / L.wheel_diameter;
  const double rpm_l = vl * L.gear_ratio * 60. / 3.1415926 / L.wheel_diameter;
  return !(lpm::dabs(rpm_r) >= L.max_motor_shaft_rpm || lpm::dabs(rpm_l) >= L.max_motor_shaft_rpm);
}

// generateTrajectory prologue (dd_simple…cpp:355-388, omni…cpp:386-424, dd_rotate_inplace…cpp:329-351).
// returns false when the reference returns false before simulating.
__device__ __forceinline__ bool traj_precheck(const Consts& C, const RobotIn& q, float v0, float v1, float v2,
                                              int* num_steps, double* dt, int* err) {
  const b200lp_limits& L = C.lim;
  const b200lp_params& P = C.par;
  const double eps = 1e-4;
  double vmag, sim_time = P.sim_time;
  const double aw = (double)fabsf(v2);
  if (P.theory == B200LP_THEORY_DD_SIMPLE) {
    vmag = (double)fabsf(v0);
    if ((L.min_vel_x >= 0 && vmag + eps < L.min_vel_x) && (L.min_vel_theta >= 0 && aw + eps < L.min_vel_theta)) return false;
    if (L.max_vel_x >= 0 && vmag - eps > L.max_vel_x) return false;
  } else if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
    vmag = (double)(float)lpm::dsqrt((double)v0 * (double)v0 + (double)v1 * (double)v1);  // hypotf
    if ((L.min_vel_trans >= 0 && vmag + eps < L.min_vel_trans) && (L.min_vel_theta >= 0 && aw + eps < L.min_vel_theta))
      return false;
    if (L.max_vel_trans >= 0 && vmag - eps > L.max_vel_trans) return false;
    if (q.max_speed_override > 0.0 && vmag - eps > q.max_speed_override) return false;
  } else {
    vmag = (double)fabsf(v0);
    sim_time = 6.28 / aw;
  }
  const double sim_time_distance = vmag * sim_time;
  const double sim_time_angle = aw * sim_time;
  const double a = sim_time_distance / P.sim_granularity, b = sim_time_angle / P.angular_sim_granularity;
  const double steps_d = ceil((a < b) ? b : a);  // std::max(a,b)
  if (!(steps_d <= (double)B200LP_MAX_STEPS)) {  // also catches NaN/inf
    *err = 1;
    return false;
  }
  const int n = (int)steps_d;
  if (n == 0) return false;
  *num_steps = n;
  *dt = sim_time / n;
  return true;
}

// Work lists of a cycle: cull_kernel sorts the trajectories into kCostClasses lists and plan_kernel drains them in order,
// so that the tail of the persistent kernel is made of cheap trajectories (those in free space only pay the path critics).
//  * Classes 0 and 1: the trajectories that took longest in the PREVIOUS cycle (plan_kernel leaves every trajectory's
//    SM-clock duration in `hist`, keyed by trajectory id (a dependent load of the sample index at the end of every
//    work item cost 5 % of the kernel); consecutive cycles see nearly the same map from nearly
//    the same pose). Nothing else predicts the expensive case — a trajectory that brushes along an obstacle for all of
//    its poses without ever hitting it takes 40-50 us on C4, the average collider (it stops at its first hit) 18 us —
//    and one of those starting half way through the launch is what the last tenth of a C4 shard's kernel waited for
//    (26 of 2960 warps busy).
//  * Classes 2..9: by the number of poses that survive the pre-cull (an upper bound on the obstacle-query work), most
//    first. This is also where a sample without history goes (first cycle, samples new to a shard).
// Only the top decile is ordered by history: draining ALL trajectories longest-first was measured and lost — the SMs then
// run nothing but long obstacle sweeps for the first half of the launch and every trajectory takes 6 % longer (C4 shard:
// 94.2 instead of 88.9 us of work per resident warp, as much as the shorter tail won). The number of cloud points inside
// the poses' candidate boxes, which the summed-volume table would hand the pre-cull for free, ranks the colliders first
// and did not help either. The order only decides who is scored when; results do not depend on it.
constexpr int kCostClasses = 10;
constexpr unsigned kHistVeryLong = 76000u, kHistLong = 53000u;  // SM clocks: ~40 us and ~28 us
__host__ __device__ constexpr int cost_class(unsigned hist_clocks, int survivors) {
#ifdef B200LP_ONE_CLASS  // A/B builds: one list in (roughly) reverse id order
  return 0;
#endif
#ifndef B200LP_NO_HISTORY  // A/B builds
  if (hist_clocks >= kHistVeryLong) return 0;
  if (hist_clocks >= kHistLong) return 1;
#endif
  return survivors == 0 ? 9 : survivors <= 4 ? 8 : survivors <= 8 ? 7 : survivors <= 16 ? 6 : survivors <= 24 ? 5
         : survivors <= 32 ? 4 : survivors <= 48 ? 3 : 2;
}

// Where a sample-sharded launch cuts the estimated-work axis: frac[k] = share of the work below cut k (frac[0] = 0,
// frac[n] = 1). n == 0: equal shares. Kernel argument, identical on every rank of a cycle.
struct ShardCuts {
  int n;
  float frac[B200LP_MAX_PEERS + 1];
};

// Host-planned sample layout of a SINGLE-ROBOT launch (plan_samples in b200lp.cu). The host owns the query, so it runs the
// three serial VelocityIterator chains once (every CTA used to repeat them: 17 k cycles per CTA at C4's 361-sample axes)
// and, for sample shards, cuts the sample grid itself and lays the chunks out so that the launch is ONE wave:
//   chunks [0, nb)            count the valid samples below the shard, `count_span` samples per CTA, nothing written;
//   chunks [nb, nb + ns)      the shard [lo, hi): one sample per thread, records + forward simulation as ever;
//   chunks [nb + ns, n_chunks) count the samples above the shard (n_samples / n_traj stay global).
// Fleet launches (planned == 0) keep one chunk per kPrepThreads samples and build the axes in the kernel: their windows
// differ per robot and live on the device.
// The axes travel as a closed form, not as arrays (a 12 KB kernel parameter block cost ~10 us of launch latency): entry j
// of a VelocityIterator chain is (float)(mn + j * step) unless the chain's accumulated rounding moved it across a float
// rounding boundary. That is not rare: minimum and maximum are floats, so the entries at simple fractions of the axis
// (1/2, 1/5, 7/30 ... of the way) sit EXACTLY half way between two floats in exact arithmetic and the last bits of the
// double decide — over random windows one window in nine has such an entry, a 361-entry axis typically ten
// (tests/cpp/axis_check.cu). The host therefore runs the real chain, compares it entry by entry with the closed form
// (the same axis_value() the kernel evaluates, DMUL + DADD on both sides) and lists the entries that differ as
// exceptions; with more than kAxisExceptions of them (3 windows in 10 000) the kernel runs the chains itself.
constexpr int kAxisExceptions = 32;
struct AxisPlan {
  double mn, step;  // the chain: next += step from mn
  float last;       // the axis' final entry, (float)max
  int n_out;        // entries of the axis (with the inserted zero)
  int zero_at;      // index of the 0.0f inserted where the chain crosses zero, or -1
  int pad;
};
__host__ __device__ inline float axis_value(const AxisPlan& a, int i) {
  if (i == a.n_out - 1) return a.last;
  if (i == a.zero_at) return 0.0f;
  const int j = (a.zero_at >= 0 && i > a.zero_at) ? i - 1 : i;
  return (float)(a.mn + (double)j * a.step);
}
struct PrepPlan {
  int planned;    // 0: fleet layout, nothing below is read
  int host_axes;  // 1: ax / exceptions are valid
  AxisPlan ax[3];
  int n_exc;
  int exc_at[kAxisExceptions];  // axis << 24 | entry
  float exc_val[kAxisExceptions];
  int nb, ns, count_span;
  long long n_raw, lo, hi;
};

// Per-(robot, chunk) aggregate of the decoupled look-back that orders the trajectory list across CTAs: ONE 16-byte word
// {kept samples, valid trajectories, pose count | error bits << 24, launch epoch}, written with a single 16-byte store
// and polled with a single 16-byte volatile load — payload and flag travel together, so neither side needs a fence (the
// fenced two-word version spent 7-18 us per CTA in MEMBAR.GPU when several hundred CTAs looked back at once). Which of
// a chunk's valid samples lie below the shard's first / end sample follows from the chunk's position in the layout.
struct alignas(16) PrepAgg {
  unsigned keep, valid;  // samples kept by the motor constraint / trajectories that passed the prologue
  unsigned poses_err;    // sum of num_steps of the chunk's shard trajectories (< 2^24) | error bits << 24
  unsigned epoch;        // == launch epoch once the chunk has published
};
static_assert(sizeof(PrepAgg) == 16, "PrepAgg layout");
__device__ __forceinline__ void agg_store(PrepAgg* p, unsigned keep, unsigned valid, unsigned poses_err, unsigned epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(keep), "r"(valid), "r"(poses_err), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint4 agg_load(const PrepAgg* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

#ifndef B200LP_PREP_THREADS
#define B200LP_PREP_THREADS 128
#endif
#ifndef B200LP_ROLL_BLOCK
#define B200LP_ROLL_BLOCK 4  // rollout steps per block (A/B builds)
#endif
#ifndef B200LP_ROLL_PIPE
#define B200LP_ROLL_PIPE 0  // 1: three-stage software pipeline of the rollout loop (A/B builds; measured slower, see the loop)
#endif
constexpr int kPrepThreads = B200LP_PREP_THREADS;  // samples per chunk: 16.5 k samples (C2) make 129 CTAs, one per SM — the
                                                    // forward simulation is bound by the XU / FP64 pipes of the SMs it runs on
constexpr int kPrepWarps = kPrepThreads / 32;
constexpr int kPrepSamples = kPrepThreads;      // one velocity sample per thread
constexpr long long kSpinLimit = 4000000000ll;  // ~2 s of SM clocks: a look-back that waits this long reports an error
// serial per-CTA jobs (three velocity axes, two pose matrices) are spread over different warps where possible
__device__ __forceinline__ bool prep_job(int job, int tid) { return tid == (job % kPrepWarps) * 32 + job / kPrepWarps; }

#ifdef B200LP_PREP_TRACE  // tools only: phase timestamps of thread 0 of the first / last chunk, printed by the kernel
#define PREP_T(k) do { if (tid == 0) trace_t[k] = clock64(); } while (0)
#else
#define PREP_T(k) do { } while (0)
#endif
// grid = (n_chunks, robots). Chunk ids are handed out by a per-robot ticket, so a CTA only ever waits for
// chunks that are already running; every CTA publishes its aggregate BEFORE it looks back.
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(Consts C, const RobotIn* __restrict__ robots, RobotIn q0,
                                                             int by_value, unsigned long long* __restrict__ t_start, int t_cap,
                                                             const __grid_constant__ PrepPlan PP, unsigned epoch,
                                                             unsigned* __restrict__ tickets, PrepAgg* aggs,
                                                             float4* __restrict__ rec_vel, int* __restrict__ rec_steps,
                                                             double* __restrict__ rec_dt, int* __restrict__ rec_sample,
                                                             RobotMeta* __restrict__ meta,
                                                             const double* __restrict__ plan7,
                                                             float4* __restrict__ plan_pts,
                                                             long long* __restrict__ rec_pose_off,
                                                             float4* __restrict__ pose_rows, long long pose_stride,
                                                             double2* __restrict__ rec_pp, int want_pp,
                                                             unsigned* __restrict__ class_counts, int axis_cap, int defer_pts) {
  // the three velocity axes: dynamic shared memory, `axis_cap` floats each (what the parameter set can produce plus the slot
  // the iterator writes ahead; 24 KB of static arrays for kMaxAxis entries cost a resident CTA per SM)
  extern __shared__ float s_axes[];
  float* const s_x = s_axes;
  float* const s_y = s_axes + axis_cap;
  float* const s_th = s_axes + 2 * axis_cap;
  __shared__ double s_R0[9], s_t0[3], s_gL[9], s_gt[3];
  __shared__ unsigned long long s_wposes[kPrepThreads / 32];
  __shared__ int s_n[3];
  __shared__ int s_wkeep[kPrepThreads / 32], s_wvalid[kPrepThreads / 32];
  __shared__ int s_chunk;
  __shared__ int s_red[6];
  __shared__ unsigned long long s_poses;
  const int robot = blockIdx.y;
  const int n_chunks = gridDim.x;
  // single-robot launches carry their query as a kernel argument: no host -> device copy in front of the cycle
  const RobotIn q = by_value ? q0 : robots[robot];
  const b200lp_limits& L = C.lim;
  const b200lp_params& P = C.par;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef B200LP_PREP_TRACE
  long long trace_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const unsigned long long trace_g0 = globaltimer_ns();
#endif
  PREP_T(0);

  if (tid == 0) {
    s_chunk = (int)atomicAdd(tickets + robot, 1u);
    s_n[0] = s_n[1] = s_n[2] = 0;
    s_red[0] = s_red[1] = s_red[2] = s_red[3] = s_red[4] = s_red[5] = 0;
    s_poses = 0ull;
  }
  __syncthreads();
  const int chunk = s_chunk;
  if (chunk == 0 && robot == 0 && tid == 0) *t_start = globaltimer_ns();  // the cycle's first CTA: start of its device timeline
  // the work lists of this cycle (cull_kernel fills them, plan_kernel drains them) start empty; the previous cycle's
  // plan_kernel is behind us in stream order
  if (chunk == 0 && robot == 0 && tid < kCostClasses) class_counts[tid] = 0u;

  if (prep_job(3, tid)) {
    // tf2::transformToEigen(robot_pose_) (dd_simple…cpp:355)
    quat_to_matrix(q.pose[3], q.pose[4], q.pose[5], q.pose[6], s_R0);
    s_t0[0] = q.pose[0]; s_t0[1] = q.pose[1]; s_t0[2] = q.pose[2];
  }
  if (prep_job(4, tid) && q.plan_n > 0) {
    const double* e = q.goal_valid ? q.goal : plan7 + (q.plan_off + q.plan_n - 1) * 7;  // prune_plan_.poses.back()
    quat_to_matrix(e[3], e[4], e[5], e[6], s_gL);
    s_gt[0] = e[0]; s_gt[1] = e[1]; s_gt[2] = e[2];
  }
  // pcl_prune_plan_: float-cast plan positions (model_shared_data.h:83-91); left to cull_kernel when the plan table is
  // still on its way (defer_pts: fleet calls)
  if (chunk == 0 && !defer_pts)
    for (int i = tid; i < q.plan_n; i += blockDim.x) {
      const double* p = plan7 + (q.plan_off + i) * 7;
      plan_pts[q.plan_off + i] = make_float4((float)p[0], (float)p[1], (float)p[2], 0.f);
    }

  const bool sampling_on = P.linear_x_sample * P.angular_z_sample > 0;
  if (PP.planned && PP.host_axes) {
    // the host ran the VelocityIterator chains and checked this closed form against them, entry by entry
    for (int i = tid; i < PP.ax[0].n_out; i += kPrepThreads) s_x[i] = axis_value(PP.ax[0], i);
    for (int i = tid; i < PP.ax[1].n_out; i += kPrepThreads) s_y[i] = axis_value(PP.ax[1], i);
    for (int i = tid; i < PP.ax[2].n_out; i += kPrepThreads) s_th[i] = axis_value(PP.ax[2], i);
    if (tid < 3) s_n[tid] = PP.ax[tid].n_out;
    __syncthreads();
    if (tid < PP.n_exc) {
      float* dst = (PP.exc_at[tid] >> 24) == 0 ? s_x : (PP.exc_at[tid] >> 24) == 1 ? s_y : s_th;
      dst[PP.exc_at[tid] & 0xffffff] = PP.exc_val[tid];
    }
  } else if (sampling_on && P.theory != B200LP_THEORY_DD_ROTATE_INPLACE && (prep_job(0, tid) || prep_job(1, tid) || prep_job(2, tid))) {
    float mn[3], mx[3];
    velocity_window(L, P, q, mn, mx);
    if (prep_job(0, tid)) s_n[0] = velocity_iterator_dev((double)mn[0], (double)mx[0], (int)P.linear_x_sample, s_x);
    if (prep_job(1, tid)) {
      if (P.theory == B200LP_THEORY_OMNI_SIMPLE) s_n[1] = velocity_iterator_dev((double)mn[1], (double)mx[1], (int)P.linear_y_sample, s_y);
      else { s_y[0] = 0.f; s_n[1] = 1; }
    }
    if (prep_job(2, tid)) s_n[2] = velocity_iterator_dev((double)mn[2], (double)mx[2], (int)P.angular_z_sample, s_th);
  }
  __syncthreads();
  PREP_T(1);  // window + velocity axes done

  int n_raw;
  const int nys = s_n[1], nths = s_n[2];
  if (!sampling_on) n_raw = 0;
  else if (P.theory == B200LP_THEORY_DD_ROTATE_INPLACE) n_raw = 2;
  else n_raw = s_n[0] * nys * nths;
  // ---- the samples of this chunk (PrepPlan) ----
  long long lo = 0, hi = n_raw;         // the launch's sample shard
  long long s0 = (long long)chunk * kPrepThreads, s1 = s0 + kPrepThreads;
  bool count_chunk = false;             // only counts its samples (they belong to another rank's shard)
  if (PP.planned) {
    lo = PP.lo; hi = PP.hi;
    if (chunk < PP.nb) {
      count_chunk = true;
      s0 = (long long)chunk * PP.count_span;
      s1 = s0 + PP.count_span < lo ? s0 + PP.count_span : lo;
    } else if (chunk < PP.nb + PP.ns) {
      s0 = lo + (long long)(chunk - PP.nb) * kPrepThreads;
      s1 = s0 + kPrepThreads < hi ? s0 + kPrepThreads : hi;
    } else {
      count_chunk = true;
      s0 = hi + (long long)(chunk - PP.nb - PP.ns) * PP.count_span;
      s1 = s0 + PP.count_span;
    }
  }
  if (s1 > n_raw) s1 = n_raw;
  LP_CHECK(!PP.planned || (PP.lo <= PP.hi && PP.hi <= n_raw && PP.n_raw == n_raw));  // the host planned the samples the kernel sees
  LP_CHECK(chunk < n_chunks);

  // ---- this chunk's samples: one per thread (count chunks: a strided run per thread, counted only) ----
  auto sample = [&](int si, float& a0, float& a1, float& a2, int& st, double& d, int& e, bool& k, bool& v) {
    a0 = a1 = a2 = 0.f;
    if (P.theory == B200LP_THEORY_DD_ROTATE_INPLACE) {
      a2 = (si == 0) ? (float)L.rotation_speed : (float)(-1.0 * L.rotation_speed);
      k = motor_ok(L, a0, a2);
    } else {
      const int ith = si % nths, iy = (si / nths) % nys, ix = si / (nths * nys);
      a0 = s_x[ix]; a1 = s_y[iy]; a2 = s_th[ith];
      k = (P.theory == B200LP_THEORY_OMNI_SIMPLE) || !L.use_motor_constraint || motor_ok(L, a0, a2);
    }
    v = k && traj_precheck(C, q, a0, a1, a2, &st, &d, &e);
  };
  const long long s = s0 + tid;
  bool keep = false, valid = false;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  int steps = 0, err = 0;
  double dt = 0.0;
  int ck = 0, cv = 0;
  if (count_chunk) {
    for (long long ss = s; ss < s1; ss += kPrepThreads) {
      bool k, v;
      int st = 0;
      double d;
      sample((int)ss, v0, v1, v2, st, d, err, k, v);
      ck += k ? 1 : 0;
      cv += v ? 1 : 0;
    }
  } else if (s < s1) {
    sample((int)s, v0, v1, v2, steps, dt, err, keep, valid);
    ck = keep ? 1 : 0;
    cv = valid ? 1 : 0;
  }
  const unsigned mk = __ballot_sync(kFull, keep), mv = __ballot_sync(kFull, valid);
  ck = __reduce_add_sync(kFull, ck);
  cv = __reduce_add_sync(kFull, cv);
  if (lane == 0) {
    s_wkeep[warp] = ck;
    s_wvalid[warp] = cv;
  }
  const bool in_shard = valid && s >= lo && s < hi;
  const unsigned long long my_poses = in_shard ? (unsigned long long)steps : 0ull;
  unsigned long long poses = my_poses;
  // inclusive warp scan of the shard's pose counts -> row of every trajectory in the pose array
  unsigned long long pscan = my_poses;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long tv = __shfl_up_sync(kFull, pscan, o);
    if (lane >= o) pscan += tv;
  }
  if (lane == 31) s_wposes[warp] = pscan;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    err |= __shfl_xor_sync(kFull, err, o);
    poses += __shfl_xor_sync(kFull, poses, o);
  }
  if (lane == 0) {
    atomicOr(&s_red[2], err);
    atomicAdd(&s_poses, poses);
  }
  __syncthreads();
  const int own_err = s_red[2];
  const long long own_poses = (long long)s_poses;
  int offk = 0, offv = 0, totk = 0, totv = 0;
  unsigned long long offp = 0ull;
#pragma unroll
  for (int w = 0; w < kPrepThreads / 32; ++w) {
    const int a = s_wkeep[w], b = s_wvalid[w];
    if (w < warp) { offk += a; offv += b; offp += s_wposes[w]; }
    totk += a; totv += b;
  }
  PrepAgg* my_aggs = aggs + (size_t)robot * n_chunks;
  PREP_T(2);  // samples checked, CTA-level counts known
  // (pose counts: a shard chunk holds at most kPrepThreads x B200LP_MAX_STEPS < 2^24 poses, count chunks none)
  if (tid == 0) agg_store(my_aggs + chunk, (unsigned)totk, (unsigned)totv, (unsigned)own_poses | ((unsigned)own_err << 24), epoch);
  // ---- look back over every earlier chunk ----
  // In a planned layout the chunks below the shard lie entirely below `lo`, the shard's chunks between `lo` and `hi`;
  // fleet launches have lo = 0 and hi = n_raw.
  const int c_lo = PP.planned ? PP.nb : 0, c_hi = PP.planned ? PP.nb + PP.ns : n_chunks;
  int pk = 0, pv = 0, plo = 0, phi = 0, perr = 0;
  long long pposes = 0;
  for (int c = tid; c < chunk; c += kPrepThreads) {
    const PrepAgg* a = my_aggs + c;
    // Every chunk polled here holds a ticket, i.e. its CTA is running; the time limit only guards against a device fault
    // in that CTA, and turns a hang into an error bit the host reports.
    uint4 w = agg_load(a);
    if (w.w != epoch) {
      const long long t_start_poll = clock64();
      do {
        w = agg_load(a);
        if (clock64() - t_start_poll > kSpinLimit) { perr |= 32; w = make_uint4(0u, 0u, 0u, epoch); }
      } while (w.w != epoch);
    }
    pk += (int)w.x; pv += (int)w.y;
    if (c < c_lo) plo += (int)w.y;
    if (c < c_hi) phi += (int)w.y;
    pposes += (long long)(w.z & 0xffffffu);
    perr |= (int)(w.z >> 24);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pk += __shfl_xor_sync(kFull, pk, o); pv += __shfl_xor_sync(kFull, pv, o);
    plo += __shfl_xor_sync(kFull, plo, o); phi += __shfl_xor_sync(kFull, phi, o);
    pposes += __shfl_xor_sync(kFull, pposes, o);
    perr |= __shfl_xor_sync(kFull, perr, o);
  }
  __syncthreads();  // s_red / s_poses were consumed by thread 0 above; reuse them for the prefix
  if (tid == 0) { s_red[0] = s_red[1] = s_red[3] = s_red[4] = 0; s_poses = 0ull; }
  __syncthreads();
  if (lane == 0 && (chunk > 0 || perr)) {
    atomicAdd(&s_red[0], pk); atomicAdd(&s_red[1], pv);
    atomicAdd(&s_red[3], plo); atomicAdd(&s_red[4], phi);
    atomicOr(&s_red[5], perr);
    atomicAdd(&s_poses, (unsigned long long)pposes);
  }
  __syncthreads();
  const int base_keep = s_red[0], base_valid = s_red[1];
  PREP_T(3);  // look-back done

  const unsigned lt = (1u << lane) - 1u;
  const long long pose_row = (long long)robot * pose_stride + (long long)(s_poses + offp + pscan - my_poses);
  size_t rec = 0;
  bool roll = false;
  if (valid) {
    const int sample_index = base_keep + offk + __popc(mk & lt);
    const int id = base_valid + offv + __popc(mv & lt);
    if (id < t_cap) {
      rec = (size_t)robot * t_cap + id;
      rec_vel[rec] = make_float4(v0, v1, v2, 0.f);
      rec_steps[rec] = steps;
      rec_dt[rec] = dt;
      rec_sample[rec] = sample_index;
      rec_pose_off[rec] = pose_row;
      roll = in_shard && (long long)(s_poses + offp + pscan) <= pose_stride;
    }
  }
  // ---- forward simulation of this thread's trajectory (computeNewPositions, dd_simple…cpp:457-464,
  // omni_simple…cpp:498-505): x,y,th in the robot frame after every step, then the pure-pursuit terms
  // of the last pose ----
  if (roll) {
    float x = 0.f, y = 0.f, th = 0.f;
    const double wdt = (double)v2 * dt;  // loop invariant of th' = (float)(th + w*dt)
    float4* out = pose_rows + pose_row;
    LP_CHECK(steps >= 1 && steps <= B200LP_MAX_STEPS);
    LP_CHECK(pose_row >= (long long)robot * pose_stride && pose_row + steps <= (long long)(robot + 1) * pose_stride);
    // Blocks of 4 steps: the heading chain th' = (float)(th + w*dt) is the only dependency the expensive sin/cos
    // evaluations have, so it runs ahead and the four evaluations overlap in the FP64 pipe.
    // B200LP_ROLL_PIPE=1 is the same arithmetic as a three-stage software pipeline (heading chain of block i + 1, sin/cos
    // of block i, position chains of block i - 1 in one loop body, independent of each other). Measured twice and slower
    // both times — round 1 with 256-trajectory CTAs (39 vs 36 us at C2) and again with one warp per SM sub-partition
    // (28.6 vs 26.8 us at C2, 29.5 vs 27.0 at C1, +4 us on a C4 shard; same box, gpurun r4s): the loop is bound by the
    // conversion (XU) and FP64 pipes of the sub-partition that holds the warp, not by the dependency chains.
    constexpr int kU = B200LP_ROLL_BLOCK;
#if B200LP_ROLL_PIPE
    float tho_c[kU], thn_c[kU], thn_p[kU];
    double ex_p[kU], ey_p[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {  // stage A of block 0
      tho_c[u] = th;
      th = (float)((double)th + wdt);
      thn_c[u] = th;
      thn_p[u] = 0.f; ex_p[u] = 0.0; ey_p[u] = 0.0;
    }
    const int n_blocks = (steps + kU - 1) / kU;
    for (int i = 0; i <= n_blocks; ++i) {
      float tho_n[kU], thn_n[kU];
      double ex_n[kU], ey_n[kU];
      // stage C of block i - 1 (nothing to store for i == 0: the guard is on the step index)
      const int kc = (i - 1) * kU;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        x = (float)((double)x + ex_p[u]);
        y = (float)((double)y + ey_p[u]);
        if (i > 0 && kc + u < steps) out[kc + u] = make_float4(x, y, thn_p[u], 0.f);
      }
      // stage B of block i (one block past the end is evaluated and dropped)
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        float sn, cs;
        lpm::sincosf(tho_c[u], &sn, &cs);
        if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
          const double a = 1.57079632679489661923 + (double)tho_c[u];  // M_PI_2 + pos[2]
          ex_n[u] = ((double)(v0 * cs) + (double)v1 * lpm::cos(a)) * dt;
          ey_n[u] = ((double)(v0 * sn) + (double)v1 * lpm::sin(a)) * dt;
        } else {
          ex_n[u] = (double)(v0 * cs) * dt;
          ey_n[u] = (double)(v0 * sn) * dt;
        }
      }
      // stage A of block i + 1
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        tho_n[u] = th;
        th = (float)((double)th + wdt);
        thn_n[u] = th;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        ex_p[u] = ex_n[u]; ey_p[u] = ey_n[u]; thn_p[u] = thn_c[u];
        tho_c[u] = tho_n[u]; thn_c[u] = thn_n[u];
      }
    }
#else
    for (int k0 = 0; k0 < steps; k0 += kU) {
      float tho[kU], thn[kU];
      double ex[kU], ey[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        tho[u] = th;
        th = (float)((double)th + wdt);
        thn[u] = th;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        float sn, cs;
        lpm::sincosf(tho[u], &sn, &cs);
        if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
          const double a = 1.57079632679489661923 + (double)tho[u];  // M_PI_2 + pos[2]
          ex[u] = ((double)(v0 * cs) + (double)v1 * lpm::cos(a)) * dt;
          ey[u] = ((double)(v0 * sn) + (double)v1 * lpm::sin(a)) * dt;
        } else {
          ex[u] = (double)(v0 * cs) * dt;
          ey[u] = (double)(v0 * sn) * dt;
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        x = (float)((double)x + ex[u]);
        y = (float)((double)y + ey[u]);
        if (k0 + u < steps) out[k0 + u] = make_float4(x, y, thn[u], 0.f);
      }
    }
#endif
    PREP_T(4);  // rollout done
    // the block loop may run past the last step: restore the state after step `steps - 1`
    {
      const float4 last = out[steps - 1];
      x = last.x; y = last.y; th = last.z;
    }
    if (want_pp && q.plan_n > 0 && steps >= 2) {
      double Lm[9], tv[3], dist, yaw;
      pose_affine(s_R0, s_t0, x, y, th, Lm, tv);
      pure_pursuit_terms(Lm, tv, s_gL, s_gt, &dist, &yaw);
      rec_pp[rec] = make_double2(dist, yaw);
    }
  }
  PREP_T(5);  // pure-pursuit terms done
#ifdef B200LP_PREP_TRACE
  if (tid == 0 && (chunk % 100 == 0 || chunk == n_chunks - 1) && robot == 0)
    printf("prep trace chunk %d/%d on sm %u: started %lld ns after the first CTA, steps %d: axes %lld, precheck %lld, look-back %lld, rollout %lld, pure pursuit %lld cycles (roll=%d), ends at %lld ns\n", chunk,
           n_chunks, (unsigned)__smid(), (long long)(trace_g0 - *t_start), steps, trace_t[1] - trace_t[0], trace_t[2] - trace_t[1], trace_t[3] - trace_t[2],
           trace_t[4] ? trace_t[4] - trace_t[3] : 0ll, trace_t[4] ? trace_t[5] - trace_t[4] : 0ll, (int)roll, (long long)(globaltimer_ns() - *t_start));
#endif
  if (chunk == n_chunks - 1 && tid == 0) {  // the last ticket: every chunk of this robot has started
    RobotMeta m;
    m.n_samples = base_keep + totk;
    m.n_traj = base_valid + totv;
    m.t_begin = s_red[3] + (chunk < c_lo ? totv : 0);
    m.t_end = s_red[4] + (chunk < c_hi ? totv : 0);
    m.error = (s_red[5] | own_err) | (m.n_traj > t_cap ? 2 : 0) | ((long long)s_poses + own_poses > pose_stride ? 4 : 0);
    m.pad = 0;
    m.n_poses = (long long)s_poses + own_poses;
    meta[robot] = m;
    tickets[robot] = 0u;  // ready for the next launch
  }
}

// =============================================================================================
// Sample-sharded cycles: the cross-GPU argmin through peer device memory.
// Every rank owns 2 x kMaxPeers slots (double-buffered by the cycle's sequence number) that its peers can write over
// NVLink (CUDA IPC mapping). The last CTA of plan_kernel — every other CTA has left — stores this rank's local best into
// slot [rank] of EVERY peer (payload, system-wide fence, then the sequence word), polls its own slots until all `world` of
// them carry this cycle's sequence number, reduces them with the reference's rule (min cost, ties -> largest id:
// local_planner.cpp:460) and hands the GLOBAL best to the spinning host thread like a single-robot cycle does. No
// host-launched collective, no extra launch, no stream synchronisation: the exchange costs one NVLink store round plus the
// skew between the ranks. Double buffering is enough: a peer can only be one cycle ahead (its cycle k+1 exchange needs our
// cycle k+1 store, which is stream-ordered after our cycle k reads).
// Every slot also carries the device time its rank needed for the cycle (prep_kernel's first CTA -> here), so all ranks
// end a cycle with the same W durations: the host moves the shard cuts of the next cycle with them (b200lp.cu).
// =============================================================================================
constexpr int kMaxPeers = 16;
struct alignas(16) PeerSlot {  // 64 bytes
  unsigned long long cost_bits;  // ~0 = this rank has no feasible trajectory
  int32_t id;                    // GLOBAL trajectory id
  uint32_t cycle_ns;             // device time of this rank's cycle up to the exchange
  double xv, yv, thetav;         // the winner's velocities travel with it: only its owner has them
  unsigned long long cuts_hash;  // which shard cuts this rank scored with: all ranks of a cycle must agree
  unsigned long long seq;        // written last
};
static_assert(sizeof(PeerSlot) == 64, "PeerSlot layout");
struct PeerTable {
  PeerSlot* slots[kMaxPeers];  // base of every rank's slot array in THIS process's address space
};
struct PeerExchange {
  int world, rank;             // world == 0: no exchange in this launch
  unsigned long long seq;
  PeerSlot* mine;
  long long timeout_cycles;
  unsigned long long* t_start; // prep_kernel's start-of-cycle timestamp (device word; valid in every launch)
  unsigned long long cuts_hash;
  PeerTable peers;
};

__device__ __forceinline__ void peer_exchange(const PeerExchange& px, int lane, const b200lp_result& r, const RobotMeta& meta0,
                                              DirectOut* direct, unsigned long long direct_seq) {
  const int buf = (int)(px.seq & 1ull);
  const unsigned my_ns = (unsigned)min(globaltimer_ns() - *px.t_start, 0xffffffffull);
  if (lane < px.world) {
    PeerSlot* dst = px.peers.slots[lane] + buf * kMaxPeers + px.rank;
    dst->cost_bits = (r.best_id < 0) ? ~0ull : lpm::d2u(r.best_cost);
    dst->id = r.best_id;
    dst->cycle_ns = my_ns;
    dst->xv = r.xv; dst->yv = r.yv; dst->thetav = r.thetav;
    dst->cuts_hash = px.cuts_hash;
    __threadfence_system();
    *(volatile unsigned long long*)&dst->seq = px.seq;
  }
  unsigned long long cb = ~0ull;
  int id = -1;
  unsigned ns = 0u;
  double xv = 0.0, yv = 0.0, thetav = 0.0;
  bool timed_out = false, cuts_differ = false;
  if (lane < px.world) {
    const volatile PeerSlot* src = px.mine + buf * kMaxPeers + lane;
    const long long t0 = clock64();
    while (src->seq != px.seq) {
      if (clock64() - t0 > px.timeout_cycles) { timed_out = true; break; }
    }
    __threadfence_system();
    if (!timed_out) {
      cb = src->cost_bits; id = src->id; ns = src->cycle_ns;
      xv = src->xv; yv = src->yv; thetav = src->thetav;
      cuts_differ = src->cuts_hash != px.cuts_hash;
    }
  }
  if (lane < kMaxPeers) direct->peer_ns[lane] = lane < px.world ? ns : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ocb = __shfl_xor_sync(kFull, cb, o);
    const int oid = __shfl_xor_sync(kFull, id, o);
    const double oxv = __shfl_xor_sync(kFull, xv, o), oyv = __shfl_xor_sync(kFull, yv, o), oth = __shfl_xor_sync(kFull, thetav, o);
    if (better(ocb, oid, cb, id)) { cb = ocb; id = oid; xv = oxv; yv = oyv; thetav = oth; }
  }
  const unsigned any_timeout = __ballot_sync(kFull, timed_out), any_differ = __ballot_sync(kFull, cuts_differ);
  __syncwarp();
  if (lane == 0) {
    b200lp_result g = r;  // n_samples / n_traj / n_collided / n_poses stay this shard's
    g.best_id = (cb == ~0ull) ? -1 : id;
    g.best_cost = (cb == ~0ull) ? -1.0 : lpm::u2d(cb);
    g.xv = (cb == ~0ull) ? 0.0 : xv;
    g.yv = (cb == ~0ull) ? 0.0 : yv;
    g.thetav = (cb == ~0ull) ? 0.0 : thetav;
    RobotMeta m = meta0;
    if (any_timeout) m.error |= 8;  // a peer did not deliver in time
    if (any_differ) m.error |= 64;  // a peer cut the sample grid differently: the shards do not tile it
    direct->r = g;
    direct->m = m;
    direct->cycle_ns = my_ns;
    __threadfence_system();
    *(volatile unsigned long long*)&direct->seq = direct_seq;
  }
}

// =============================================================================================
// One map for every rank of a peer group (replicated-map configurations: fleets, sample shards): ONE rank — the root —
// brings the cloud from its host, packed to 12-byte rows, and pushes every upload piece into the row buffer of every
// peer over NVLink as soon as the piece has landed in its own memory; the peers count the piece into their histogram
// as soon as its flag is up and build their grid locally. The host link is crossed once instead of once per rank.
// Block layout of every rank's peer allocation: [2 x kMaxPeers PeerSlot | CloudHeader | rows].
// =============================================================================================
constexpr int kSharePieces = kPackChunks;  // the pieces of the packing upload (lp_hostpack.h)
struct alignas(16) CloudHeader {
  unsigned long long n;                         // points of the cloud (rows in the buffer)
  unsigned long long n_finite;
  float mn[4], mx[4];                           // bounds of the finite points (root's packing threads)
  unsigned long long bound[kSharePieces + 1];   // piece c holds points [bound[c], bound[c + 1])
  unsigned long long seq;                       // == cycle number once the fields above are visible
  unsigned long long piece_seq[kSharePieces];   // == cycle number once piece c has landed in the rows
  unsigned long long ack[kMaxPeers];            // (root's copy) rank r has finished reading the rows of cycle ack[r]
};
constexpr size_t kPeerHeaderOff = 2 * kMaxPeers * sizeof(PeerSlot);
constexpr size_t kPeerRowsOff = 4096;
static_assert(kPeerHeaderOff + sizeof(CloudHeader) <= kPeerRowsOff, "peer block layout");
struct PeerBlocks {
  char* base[kMaxPeers];  // every rank's peer allocation in THIS process's address space
};
__device__ __forceinline__ CloudHeader* peer_header(char* base) { return reinterpret_cast<CloudHeader*>(base + kPeerHeaderOff); }

// root, before the first piece of a cycle: every peer must have finished reading the rows of the previous cycle
__global__ void __launch_bounds__(32) share_wait_acks_kernel(const CloudHeader* mine, int world, int root, unsigned long long prev_seq,
                                                             long long timeout_cycles, int* __restrict__ err) {
  const int r = threadIdx.x;
  if (r >= world || r == root || prev_seq == 0ull) return;
  const volatile unsigned long long* a = &mine->ack[r];
  const long long t0 = clock64();
  while (*a < prev_seq)
    if (clock64() - t0 > timeout_cycles) { atomicOr(err, 1); break; }
}

// root: copy bytes [b0, b1) of the rows (16-byte aligned start) into the same place of every peer's row buffer; the last
// CTA to finish raises the piece's flag on every peer
__global__ void __launch_bounds__(256) share_push_kernel(const char* __restrict__ rows, size_t b0, size_t b1, PeerBlocks peers, int world,
                                                         int root, int piece, unsigned long long seq, unsigned* __restrict__ ticket) {
  const size_t n16 = (b1 - b0) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(rows + b0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(src + i);
    for (int r = 0; r < world; ++r)
      if (r != root) reinterpret_cast<uint4*>(peers.base[r] + kPeerRowsOff + b0)[i] = v;
  }
  if (blockIdx.x == 0) {  // the (at most 12) bytes behind the last 16-byte word
    const size_t t0 = b0 + n16 * 16;
    for (size_t b = t0 + (size_t)threadIdx.x * 4; b < b1; b += 4 * blockDim.x) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(rows + b);
      for (int r = 0; r < world; ++r)
        if (r != root) *reinterpret_cast<uint32_t*>(peers.base[r] + kPeerRowsOff + b) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < world && (int)threadIdx.x != root)
    *(volatile unsigned long long*)&peer_header(peers.base[threadIdx.x])->piece_seq[piece] = seq;
  if (threadIdx.x == 0) *ticket = 0u;
}

// root: the cloud's size, bounds and piece boundaries for every peer (known once the packing threads are through)
__global__ void __launch_bounds__(32) share_header_kernel(CloudHeader h, PeerBlocks peers, int world, int root) {
  const int r = threadIdx.x;
  if (r >= world || r == root) return;
  CloudHeader* d = peer_header(peers.base[r]);
  d->n = h.n;
  d->n_finite = h.n_finite;
#pragma unroll
  for (int a = 0; a < 4; ++a) { d->mn[a] = h.mn[a]; d->mx[a] = h.mx[a]; }
#pragma unroll
  for (int c = 0; c <= kSharePieces; ++c) d->bound[c] = h.bound[c];
  __threadfence_system();
  *(volatile unsigned long long*)&d->seq = h.seq;
}

// peer: wait for the root's header and hand it to the host (mapped pinned memory; the host spins on host_out->seq)
__global__ void __launch_bounds__(32) share_wait_header_kernel(const CloudHeader* mine, unsigned long long seq, CloudHeader* host_out,
                                                               long long timeout_cycles) {
  if (threadIdx.x != 0) return;
  const volatile unsigned long long* s = &mine->seq;
  const long long t0 = clock64();
  bool ok = true;
  while (*s != seq)
    if (clock64() - t0 > timeout_cycles) { ok = false; break; }
  __threadfence_system();
  const volatile CloudHeader* m = mine;
  host_out->n = ok ? m->n : ~0ull;  // ~0: timed out
  host_out->n_finite = m->n_finite;
  for (int a = 0; a < 4; ++a) { host_out->mn[a] = m->mn[a]; host_out->mx[a] = m->mx[a]; }
  for (int c = 0; c <= kSharePieces; ++c) host_out->bound[c] = m->bound[c];
  __threadfence_system();
  *(volatile unsigned long long*)&host_out->seq = seq;
}

// peer: hold the stream until piece `piece` of cycle `seq` has landed
__global__ void __launch_bounds__(32) share_wait_piece_kernel(const CloudHeader* mine, unsigned long long seq, int piece,
                                                              long long timeout_cycles, int* __restrict__ err) {
  if (threadIdx.x != 0) return;
  const volatile unsigned long long* s = &mine->piece_seq[piece];
  const long long t0 = clock64();
  while (*s != seq)
    if (clock64() - t0 > timeout_cycles) { atomicOr(err, 2); break; }
  __threadfence_system();
}

// peer, after its grid is built: the rows may be overwritten by the next cycle
__global__ void __launch_bounds__(32) share_ack_kernel(char* root_base, int rank, unsigned long long seq) {
  if (threadIdx.x != 0) return;
  __threadfence_system();
  *(volatile unsigned long long*)&peer_header(root_base)->ack[rank] = seq;
}

// =============================================================================================
// cull: the float pre-cull of every pose and the work lists of the cycle, in front of plan_kernel.
// A CTA takes kCullTraj consecutive trajectories of one robot — their pose rows are one contiguous range of the cycle's pose
// array — and gives every row a thread: a conservative box around the pose's candidate box (loose_box) and one look at the
// summed-volume table; poses whose box holds no cloud point cannot collide (about 6 of 10 at C2) and never reach the
// double-precision geometry. The survivor bits of the range are collected in shared memory; one thread per trajectory then
// cuts its bits out (`surv`: mask_stride words per trajectory, bit k = pose k survives), counts them and appends the
// trajectory to the work list of its cost class. Trajectories are visited from the END of the list (longest rollouts first),
// so every class list starts with its longest members.
// =============================================================================================
#ifndef B200LP_CULL_THREADS
#define B200LP_CULL_THREADS 256
#endif
constexpr int kCullThreads = B200LP_CULL_THREADS;
constexpr int kCullTraj = 32;
#ifndef B200LP_PRECULL
#define B200LP_PRECULL 1  // 0: every pose goes through the double-precision geometry (A/B builds, tools/time_variants.py)
#endif
// cells_with_points() without the cell range: 32-bit index arithmetic (the table has < 2^29 entries: b200lp_grid_config
// caps the cells at 2^26)
__device__ __forceinline__ uint32_t box_point_count(const GridDev& g, const float* lo, const float* hi) {
  const float fx0 = cell_f(lo[0], g.org[0], g.inv_xy), fx1 = cell_f(hi[0], g.org[0], g.inv_xy);
  const float fy0 = cell_f(lo[1], g.org[1], g.inv_xy), fy1 = cell_f(hi[1], g.org[1], g.inv_xy);
  const float fz0 = cell_f(lo[2], g.org[2], g.inv_z), fz1 = cell_f(hi[2], g.org[2], g.inv_z);
  const float nxm = (float)(g.nx - 1), nym = (float)(g.ny - 1), nzm = (float)(g.nz - 1);
  const bool empty = !(lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) || fx1 < 0.f || fy1 < 0.f || fz1 < 0.f ||
                     fx0 > nxm || fy0 > nym || fz0 > nzm || g.n_kept == 0;
  if (empty) return 0u;
  const unsigned x0 = (unsigned)fmaxf(fx0, 0.f), x1 = (unsigned)fminf(fx1, nxm) + 1u;
  const unsigned y0 = (unsigned)fmaxf(fy0, 0.f), y1 = (unsigned)fminf(fy1, nym) + 1u;
  const unsigned z0 = (unsigned)fmaxf(fz0, 0.f), z1 = (unsigned)fminf(fz1, nzm) + 1u;
  const unsigned sx = (unsigned)(g.nx + 1), sy = (unsigned)(g.ny + 1) * sx;
  const uint32_t* s00 = g.sat + (z0 * sy + y0 * sx);
  const uint32_t* s01 = g.sat + (z0 * sy + y1 * sx);
  const uint32_t* s10 = g.sat + (z1 * sy + y0 * sx);
  const uint32_t* s11 = g.sat + (z1 * sy + y1 * sx);
  const uint32_t up = (__ldg(s11 + x1) - __ldg(s11 + x0)) - (__ldg(s10 + x1) - __ldg(s10 + x0));
  const uint32_t dn = (__ldg(s01 + x1) - __ldg(s01 + x0)) - (__ldg(s00 + x1) - __ldg(s00 + x0));
  return up - dn;  // (modular: the differences of a summed-volume table)
}

__global__ void __launch_bounds__(kCullThreads)
cull_kernel(Consts C, GridDev g, const RobotIn* __restrict__ robots, RobotIn q0, int by_value, const RobotMeta* __restrict__ meta,
            int t_cap, const int* __restrict__ rec_steps, const long long* __restrict__ rec_pose_off,
            const float4* __restrict__ poses, uint32_t* __restrict__ surv, int mask_stride, int* __restrict__ order,
            size_t order_stride, unsigned* __restrict__ class_counts, const unsigned* __restrict__ hist,
            const double* __restrict__ plan7, float4* __restrict__ plan_pts, int defer_pts) {
  constexpr int kBitWords = kCullTraj * (B200LP_MAX_STEPS / 32) + 1;
  __shared__ float s_R0f[9], s_t0f[3];
  __shared__ long long s_off[kCullTraj];
  __shared__ int s_n[kCullTraj];
  __shared__ uint32_t s_bits[kBitWords];
  __shared__ unsigned s_cnt[kCostClasses], s_base[kCostClasses];
  const int robot = blockIdx.y;
  if (defer_pts && blockIdx.x == 0) {  // pcl_prune_plan_ (model_shared_data.h:83-91) of this robot: see prep_kernel
    const long long off = by_value ? q0.plan_off : robots[robot].plan_off;
    const int pn = by_value ? q0.plan_n : robots[robot].plan_n;
    for (int i = threadIdx.x; i < pn; i += kCullThreads) {
      const double* p = plan7 + (off + i) * 7;
      plan_pts[off + i] = make_float4((float)p[0], (float)p[1], (float)p[2], 0.f);
    }
  }
  const RobotMeta m = meta[robot];
  const int n_local = min(m.t_end, t_cap) - m.t_begin;
  const int first = (int)blockIdx.x * kCullTraj;
  if (first >= n_local) return;  // (whole CTA)
  const int cnt = min(kCullTraj, n_local - first);
  LP_CHECK(cnt >= 1 && m.t_begin >= 0 && m.t_end <= m.n_traj);
  const int id_lo = m.t_begin + (n_local - first - cnt);  // the CTA's trajectories, ascending ids id_lo .. id_lo + cnt - 1
  const int tid = threadIdx.x;
  if (tid < cnt) {
    const size_t rec = (size_t)robot * t_cap + id_lo + tid;
    s_off[tid] = rec_pose_off[rec];
    s_n[tid] = rec_steps[rec];
  }
  if (tid == 32) {  // (another warp than the one that loads the offsets)
    const RobotIn& q = by_value ? q0 : robots[robot];
    double R0[9];
    quat_to_matrix(q.pose[3], q.pose[4], q.pose[5], q.pose[6], R0);
#pragma unroll
    for (int a = 0; a < 9; ++a) s_R0f[a] = (float)R0[a];
#pragma unroll
    for (int a = 0; a < 3; ++a) s_t0f[a] = (float)q.pose[a];
  }
  if (tid < kCostClasses) s_cnt[tid] = 0u;
  __syncthreads();
  // the rows of the cycle's pose array these trajectories occupy (prep_kernel lays trajectories out back to back)
  const long long row0 = s_off[0];
  const int n_rows = (int)min((long long)(kBitWords - 1) * 32, s_off[cnt - 1] + s_n[cnt - 1] - row0);
  for (int r = tid; r < ((n_rows + 31) & ~31); r += kCullThreads) {
    bool keep = false;
    if (r < n_rows) {
#if B200LP_PRECULL
      const float4 pz = __ldg(poses + row0 + r);
      float lo[3], hi[3];
      loose_box(C, g, s_R0f, s_t0f, pz.x, pz.y, pz.z, lo, hi);
      keep = box_point_count(g, lo, hi) != 0u;
#else
      keep = true;
#endif
    }
    const unsigned mk = __ballot_sync(kFull, keep);
    if ((tid & 31) == 0) s_bits[r >> 5] = mk;
  }
  if (tid == 0) s_bits[(n_rows + 31) >> 5] = 0u;  // (the word behind the last one is read by the funnel shifts below)
  __syncthreads();
  int cls = -1, rec_i = 0;
  unsigned pos = 0u;
  if (tid < cnt) {
    const size_t rec = (size_t)robot * t_cap + id_lo + tid;
    const int b0 = (int)(s_off[tid] - row0), n = s_n[tid];
    int total = 0;
    for (int w = 0; w * 32 < n; ++w) {
      const int b = b0 + w * 32;
      unsigned mk = 0u;
      LP_CHECK(b >= 0 && (b >= n_rows || (b >> 5) + 1 < kBitWords));
      if (b < n_rows) mk = __funnelshift_r(s_bits[b >> 5], s_bits[(b >> 5) + 1], (unsigned)(b & 31));
      LP_CHECK(w < mask_stride);
      if (n - w * 32 < 32) mk &= (1u << (n - w * 32)) - 1u;  // (the rows behind belong to the next trajectory)
      surv[rec * (size_t)mask_stride + w] = mk;
      total += __popc(mk);
    }
    cls = cost_class(__ldg(hist + rec), total);
    rec_i = (int)rec;
    pos = atomicAdd(&s_cnt[cls], 1u);
  }
  __syncthreads();
  if (tid < kCostClasses && s_cnt[tid]) s_base[tid] = atomicAdd(class_counts + tid, s_cnt[tid]);
  __syncthreads();
  LP_CHECK(cls < 0 || (size_t)s_base[cls] + pos < order_stride);
  if (cls >= 0) order[(size_t)cls * order_stride + s_base[cls] + pos] = rec_i;
}

// =============================================================================================
// plan: persistent warps, one trajectory per work item (robot, local trajectory index), handed out by a
// global counter; grid = min(work, SMs x resident CTAs). Each warp rolls the trajectory out 32 poses at a
// time, queries the grid, evaluates the critic stack and writes cost / per-critic scores / first-hit pose.
// argmin_kernel then picks the best trajectory per robot.
// =============================================================================================

// Per-warp state that lives in shared memory, not in registers: the work loop of plan_kernel is register-bound (96 per
// thread for 5 CTAs per SM), and everything here is read a few times per trajectory at most.
struct WarpCtx {
  double R0[9], t0[3];
  double heading_deviation;
  const float4* plan;    // the robot's prune plan (shared-memory copy for single-robot launches)
  int plan_n;
  // running best of the trajectories this warp scored (single-robot launches): (cost bits, id), collisions
  unsigned long long best_cb;
  int best_id, n_coll;
  unsigned t_item;       // SM clock at the start of the current work item
  int gbox[32 / kGroup][6];  // united cell box of each group of kGroup consecutive stash columns
};

struct Best {
  unsigned long long cb;  // cost bits, ~0 = none
  int bi, ncoll;
  int nscored = 0;         // trajectories scored / their poses: what the launch really did, checked against prep_kernel's list
  long long nposes = 0;
  __device__ __forceinline__ void take(unsigned long long ocb, int obi) {
    if (ocb != ~0ull && (cb == ~0ull || better(ocb, obi, cb, bi))) { cb = ocb; bi = obi; }
  }
  __device__ __forceinline__ void reset() { cb = ~0ull; bi = -1; ncoll = 0; nscored = 0; nposes = 0; }
  __device__ __forceinline__ void store(BlockBest* d) const {
    d->cost_bits = cb; d->id = bi; d->n_collided = ncoll; d->n_scored = nscored; d->pad = 0; d->poses_scored = nposes;
  }
  __device__ __forceinline__ void merge(const BlockBest& o) {
    ncoll += o.n_collided; nscored += o.n_scored; nposes += o.poses_scored;
    take(o.cost_bits, o.id);
  }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ocb = __shfl_xor_sync(kFull, cb, o);
      const int obi = __shfl_xor_sync(kFull, bi, o);
      ncoll += __shfl_xor_sync(kFull, ncoll, o);
      nscored += __shfl_xor_sync(kFull, nscored, o);
      nposes += __shfl_xor_sync(kFull, nposes, o);
      take(ocb, obi);
    }
  }
};


struct CtaShared {
  float stash[kWarpsPerCta][F_COUNT * 32];
  float4 pre[kWarpsPerCta][32 * kPreStride];
  WarpCtx wc[kWarpsPerCta];
  float4 plan_xy[kWarpsPerCta][kPlanWarp / 2];  // every warp's copy of its robot's prune plan, two points per entry
  float2 plan_z[kWarpsPerCta][kPlanWarp / 2];
  unsigned short list[kWarpsPerCta][64];  // poses that survived the pre-cull and wait for the exact geometry, ascending
  int scored[kWarpsPerCta];               // trajectories / poses each warp scored (single-robot launches)
  long long poses_scored[kWarpsPerCta];
  unsigned cls_first[kCostClasses + 1];   // first work index of every cost class; [kCostClasses] = number of work items
#if B200LP_STREAM == 2
  SweepRing ring[kWarpsPerCta];           // bulk-copy staging of the candidate stream (lp_device.cuh)
#endif
};

// lanes that head a group of kGroup consecutive stash columns
__host__ __device__ constexpr unsigned group_heads() {
  unsigned m = 0u;
  for (int i = 0; i < 32; i += kGroup) m |= 1u << i;
  return m;
}

#ifndef B200LP_PLAN_MIN_CTAS
#define B200LP_PLAN_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(kThreads, B200LP_PLAN_MIN_CTAS)
plan_kernel(Consts C, GridDev g, const RobotIn* __restrict__ robots, RobotIn q0, int by_value, RobotMeta* meta, int n_robots,
            int t_cap, const float4* __restrict__ rec_vel, const int* __restrict__ rec_steps,
            const long long* __restrict__ rec_pose_off, const float4* __restrict__ poses,
            const double2* __restrict__ rec_pp, const float4* __restrict__ plan_pts, double* __restrict__ out_cost,
            double* __restrict__ out_scores, int* __restrict__ out_first_hit,
            unsigned long long* __restrict__ work_counter, BlockBest* partial, unsigned* __restrict__ tickets,
            b200lp_result* __restrict__ results, DirectOut* direct, unsigned long long direct_seq, PeerExchange px,
            const uint32_t* __restrict__ surv, int mask_stride, const int* __restrict__ order, size_t order_stride,
            const unsigned* __restrict__ class_counts, unsigned* __restrict__ hist) {
  __shared__ CtaShared S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* stash = S.stash[warp];
  float4* pre = S.pre[warp];
  WarpCtx& W = S.wc[warp];
  unsigned ring_state = 0u;
#if B200LP_STREAM == 2
  SweepRing* ring = &S.ring[warp];
  ring_init(ring, lane);
#else
  SweepRing* ring = nullptr;
#endif
  // Work space: the class lists cull_kernel filled, most expensive class first; position w of the concatenation is entry
  // w - cls_first[k] of list k.
  if (threadIdx.x == 0) {
    unsigned acc = 0u;
#pragma unroll
    for (int k = 0; k < kCostClasses; ++k) {
      S.cls_first[k] = acc;
      acc += class_counts[k];
    }
    S.cls_first[kCostClasses] = acc;
  }
  const int nc = C.n_critics;

  __syncthreads();  // (cls_first)

  // Single-robot launches fold Local_Planner::getBestTrajectory into this kernel (warp -> CTA -> last CTA by ticket);
  // fleets run argmin_kernel afterwards. The (cost bits, id) order is total, so the merge tree does not matter.
  const int fused_argmin = n_robots == 1 ? 1 : 0;
  if (lane == 0) {
    S.scored[warp] = 0; S.poses_scored[warp] = 0;
    W.best_cb = ~0ull; W.best_id = -1; W.n_coll = 0;
  }
  int cur_robot = -1;
  const unsigned total = S.cls_first[kCostClasses];  // (written before the __syncthreads above)
  // Work items: position w of the concatenated class lists, claimed with an atomic on the global counter. A claimed item
  // is a trajectory nobody else can take, and trajectories differ a lot in cost (C2: mean 17 us, p99 55 us, max 90 us), so a
  // warp claims its next item LATE — when the obstacle query of the current one is over and only the path critics and the
  // scoring are left — and fetches the list entry when those are done: both round trips overlap work, and no item waits
  // behind a long one while other warps leave (claiming one or two items ahead at the top of the loop left the last 25 %
  // of the launch with < 2 % of the warps busy).
  auto list_entry = [&](unsigned w) -> int {  // (robot * t_cap + id) of work item w, -1 past the end
    if (w >= total) return -1;
    int k = 0;
#pragma unroll
    for (int c = 1; c < kCostClasses; ++c) k += w >= S.cls_first[c] ? 1 : 0;
    return __ldg(order + (size_t)k * order_stride + (w - S.cls_first[k]));
  };
  unsigned ticket = 0u;  // lane 0
  int next_rec = -1;     // lane 0: list entry of the next item (-1: past the end)
  if (lane == 0) {
    ticket = (unsigned)atomicAdd(work_counter, 1ull);
    next_rec = list_entry(ticket);
  }
  for (;;) {
    const int rec_i = __shfl_sync(kFull, next_rec, 0);
    if (rec_i < 0) break;
    const int robot = rec_i / t_cap;
    const int id = rec_i - robot * t_cap;
    LP_CHECK(robot >= 0 && robot < n_robots && id >= meta[robot].t_begin && id < meta[robot].t_end);
    if (robot != cur_robot) {
      cur_robot = robot;
      const RobotIn& q = by_value ? q0 : robots[robot];
      __syncwarp();
      if (lane == 0) {
        W.plan_n = q.plan_n;
        W.heading_deviation = q.heading_deviation;
        W.plan = plan_pts + q.plan_off;
        // tf2::transformToEigen(robot_pose_) (dd_simple…cpp:355)
        quat_to_matrix(q.pose[3], q.pose[4], q.pose[5], q.pose[6], W.R0);
        W.t0[0] = q.pose[0]; W.t0[1] = q.pose[1]; W.t0[2] = q.pose[2];
      }
      // the warp's copy of the robot's prune plan (pcl_prune_plan_), two points per entry, for the path critics' scans
      if (q.plan_n <= kPlanWarp) {
        const float4* gp = plan_pts + q.plan_off;
        for (int k = lane; 2 * k < q.plan_n; k += 32) {
          const float4 a = gp[2 * k], b = gp[min(2 * k + 1, q.plan_n - 1)];
          S.plan_xy[warp][k] = make_float4(a.x, a.y, b.x, b.y);
          S.plan_z[warp][k] = make_float2(a.z, b.z);
        }
      }
      __syncwarp();
    }
    const size_t rec = (size_t)rec_i;
#ifdef B200LP_TRAJ_TRACE  // tools only (tools/traj_trace.py): when every trajectory started and how long it took
    const unsigned long long trace_t0 = globaltimer_ns();
#endif
#ifndef B200LP_NO_HIST_RECORD
    if (lane == 0) W.t_item = (unsigned)clock64();  // (its duration becomes the trajectory's cost estimate of the next cycle)
#endif
    const int n = rec_steps[rec];
    const long long pose_row = rec_pose_off[rec];
    const float4* traj_poses = poses + pose_row;  // prep_kernel's forward simulation

    // ---- critic stack analysis (warp-uniform) ------------------------------------------------
    // A critic's value may be known before the rollout (collision_model.cpp:53-55, stick_path_model.cpp:53-55,
    // pure_pursuit_model.cpp:62-84, shortest_angle_model.cpp:51-69, twirling_model.cpp:51-55).
    auto upfront = [&](const CriticDev& cr, double& v) -> bool {
      const int plan_n = W.plan_n;
      switch (cr.kind) {
        case B200LP_CRITIC_COLLISION:
        case B200LP_CRITIC_COLLISION_MIN_MAX:
          if (g.n_raw < 5) { v = 0.0; return true; }
          return false;
        case B200LP_CRITIC_STICK_PATH:
        case B200LP_CRITIC_TOWARD_GLOBAL_PLAN:
          if (plan_n < 3) { v = 10.0; return true; }
          return false;
        case B200LP_CRITIC_PURE_PURSUIT:
          if (plan_n == 0 || n < 2) { v = -4.0; return true; }
          return false;
        case B200LP_CRITIC_SHORTEST_ANGLE: {
          const double thetav = (double)rec_vel[rec].z;
          if (W.heading_deviation >= 0) v = (thetav >= 0) ? cr.weight : cr.weight * 2;
          else v = (thetav >= 0) ? cr.weight * 2 : cr.weight;
          return true;
        }
        case B200LP_CRITIC_TWIRLING:
          v = lpm::dabs((double)rec_vel[rec].z) * cr.weight;
          return true;
      }
      return false;
    };
    // (kept as bits of ONE general register, not as bools: long-lived predicates starve the sweep's inner loop, whose
    // twelve compares then serialise through the one predicate that is left — measured +18 % on the whole kernel)
    enum : unsigned { kNeedBox = 1u, kNeedMM = 2u, kNeedStick = 4u, kNeedLastNN = 8u, kNeedPP = 16u, kEarlyOk = 32u, kRejected = 64u };
    unsigned fl = kEarlyOk;    // kEarlyOk: may a collision hit end the rollout?
    int first_coll_kind = -1;  // kind of the stack's first live collision critic: 0 box, 1 min-max
    {
      bool rollout_dep_seen = false;
#pragma unroll 1
      for (int k = 0; k < nc; ++k) {
        const CriticDev& cr = C.critics[k];
        double v;
        if (upfront(cr, v)) {
          if (v < 0) break;  // known-negative up front: later critics are never reached
          continue;
        }
        switch (cr.kind) {
          case B200LP_CRITIC_COLLISION:
          case B200LP_CRITIC_COLLISION_MIN_MAX: {
            const bool mm = cr.kind == B200LP_CRITIC_COLLISION_MIN_MAX;
            if (rollout_dep_seen) fl &= ~kEarlyOk;
            fl |= mm ? kNeedMM : kNeedBox;
            if (first_coll_kind < 0) first_coll_kind = mm ? 1 : 0;
            break;
          }
          case B200LP_CRITIC_STICK_PATH: fl |= kNeedStick; rollout_dep_seen = true; break;
          case B200LP_CRITIC_TOWARD_GLOBAL_PLAN: fl |= kNeedLastNN; rollout_dep_seen = true; break;
          case B200LP_CRITIC_PURE_PURSUIT: fl |= kNeedPP; rollout_dep_seen = true; break;
          default: break;
        }
      }
    }
    // ---- pass 1: obstacle query ------------------------------------------------------------------------
    // cull_kernel has already looked at every pose in float: `surv` holds, per trajectory, the mask of the poses whose
    // conservative candidate box contains cloud points at all (about 4 of 10 at C2); the others cannot collide. The
    // survivors are queued in ascending pose order; as soon as 32 wait — or the trajectory is exhausted — their cuboids are
    // built with the reference's arithmetic (double precision), and groups of kGroup consecutive survivors are swept
    // against their united candidate stream, lowest pose first, so the first colliding pose found is the reference's.
    int hit_box = -1, hit_mm = -1;  // first colliding pose per collision-critic kind
    // kRejected: the stack's first collision critic hit and ends the evaluation
    if (fl & (kNeedBox | kNeedMM)) {
      unsigned short* list = S.list[warp];
      const unsigned lt = (1u << lane) - 1u;
      int cnt = 0, base = 0;  // survivors waiting / next pose to look at (warp-uniform)
      for (;;) {
        while (cnt < 32 && base < n) {
          const unsigned mk = __ldg(surv + rec * (size_t)mask_stride + (base >> 5));  // (warp-uniform address)
          LP_CHECK(cnt + __popc(mk) <= 64 && (base >> 5) < mask_stride);
          if ((mk >> lane) & 1u) list[cnt + __popc(mk & lt)] = (unsigned short)(base + lane);
          cnt += __popc(mk);
          base += 32;
        }
        if (cnt == 0) break;
        __syncwarp();
        const int m_here = min(cnt, 32);
        const bool live = lane < m_here;
        LP_CHECK(!live || list[lane] < n);
        const float4 pz = live ? __ldg(traj_poses + list[lane]) : make_float4(0.f, 0.f, 0.f, 0.f);
        double L[9], t[3];
        pose_affine(W.R0, W.t0, pz.x, pz.y, pz.z, L, t);
        CellBox cbx;
        pose_geometry(C, g, L, t, stash, pre, &cbx, lane, live, nullptr);
        group_union(cbx);
        if ((lane & (kGroup - 1)) == 0) {  // the group's box goes to shared memory: six registers less across the sweeps
          int* gb = W.gbox[lane / kGroup];
          gb[0] = cbx.x0; gb[1] = cbx.x1; gb[2] = cbx.y0; gb[3] = cbx.y1; gb[4] = cbx.z0; gb[5] = cbx.z1;
        }
        // groups of kGroup consecutive survivors whose united box holds points, once per collision-critic kind still undecided
        const unsigned heads = __ballot_sync(kFull, cbx.x0 <= cbx.x1) & group_heads();
        __syncwarp();
#pragma unroll 1
        for (int kind = 0; kind < 2; ++kind) {
          if (kind == 0 ? !((fl & kNeedBox) && hit_box < 0) : !((fl & kNeedMM) && hit_mm < 0)) continue;
          unsigned todo = heads;
#pragma unroll 1
          while (todo) {
            const int col0 = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int* gb = W.gbox[col0 / kGroup];
            const CellBox ub = {gb[0], gb[1], gb[2], gb[3], gb[4], gb[5]};
            const SweepGrid sg = {g.pts, g.cell_start, g.nx, g.ny, g.cmax};
            const unsigned h = kind ? sweep_points<true>(sg, stash, pre, col0, lane, ub, ring, ring_state)
                                    : sweep_points<false>(sg, stash, pre, col0, lane, ub, ring, ring_state);
            if (h) {
              const int hp = (int)list[col0 + (__ffs(h) - 1)];
              if (kind) hit_mm = hp; else hit_box = hp;
              break;
            }
          }
        }
        // A hit of the FIRST collision critic of the stack ends the trajectory (the reference returns -1
        // there) when nothing that precedes it in the stack depends on the rollout. A later collision
        // critic hitting first only retires that critic: the earlier one must still run to the end.
        if ((fl & kEarlyOk) && first_coll_kind >= 0 && (first_coll_kind ? hit_mm : hit_box) >= 0) {
          fl |= kRejected;
          break;
        }
        if (!((fl & kNeedBox) && hit_box < 0) && !((fl & kNeedMM) && hit_mm < 0)) break;  // every collision critic is decided
        // drop the chunk from the queue
        const int rest = cnt - m_here;  // < 32
        const unsigned short carry = lane < rest ? list[m_here + lane] : (unsigned short)0;
        __syncwarp();
        if (lane < rest) list[lane] = carry;
        cnt = rest;
        if (cnt == 0 && base >= n) break;
        __syncwarp();
      }
      __syncwarp();
    }

    if (lane == 0) ticket = (unsigned)atomicAdd(work_counter, 1ull);  // the next item, claimed late (see the top of the loop)
    // ---- pass 2: path critics, only for trajectories the collision critic did not already reject ----
    double stick_sum = 0.0;
    float last_nn = 0.f;
    if (!(fl & kRejected) && (fl & (kNeedStick | kNeedLastNN))) {
      for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const int n_here = min(32, n - base);
        float d2 = 0.f;
        if (k < n) {
          const float4 pz = __ldg(traj_poses + k);
          // Trajectory::getPCLPoint: the translation of pos_af3 * [Rz(th), (x,y,0)] (same expression as pose_affine)
          float pw[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) pw[a] = (float)((W.R0[a * 3] * (double)pz.x + W.R0[a * 3 + 1] * (double)pz.y) + W.t0[a]);
          d2 = W.plan_n <= kPlanWarp ? plan_nn_d2_pairs(S.plan_xy[warp], S.plan_z[warp], (W.plan_n + 1) >> 1, pw[0], pw[1], pw[2])
                                     : plan_nn_d2(W.plan, W.plan_n, pw[0], pw[1], pw[2]);
        }
        const float sq = lpm::fsqrt(d2);
        if (fl & kNeedStick) {
          // normalized_distance += sqrt(d2), in pose order, in double (stick_path_model.cpp:68)
          for (int j = 0; j < n_here; ++j) stick_sum += (double)__shfl_sync(kFull, sq, j);
        }
        if (n - 1 - base < 32 && n - 1 >= base) last_nn = __shfl_sync(kFull, sq, n - 1 - base);
      }
    }
    double pp_dist = 0.0, pp_yaw = 0.0;
    if (fl & kNeedPP) {
      const double2 pp = rec_pp[rec];
      pp_dist = pp.x;
      pp_yaw = pp.y;
    }

    if (lane == 0) next_rec = list_entry(ticket);
    // ---- StackedScoringModel::scoreTrajectory (stacked_scoring_model.cpp:75-93) -----------------
    double cost = 0.0;
    int first_hit = -1;
    bool stopped = false;
#pragma unroll 1
    for (int k = 0; k < nc; ++k) {
      const CriticDev& cr = C.critics[k];
      double v = lpm::u2d(0x7ff8000000000000ull);  // NaN = not evaluated (an earlier critic rejected the trajectory)
      if (!stopped) {
        if (!upfront(cr, v)) {
          switch (cr.kind) {
            case B200LP_CRITIC_COLLISION:
            case B200LP_CRITIC_COLLISION_MIN_MAX: {
              const int hp = cr.kind == B200LP_CRITIC_COLLISION ? hit_box : hit_mm;
              v = (hp >= 0) ? -1.0 : 0.0;
              if (hp >= 0 && first_hit < 0) first_hit = hp;
              break;
            }
            case B200LP_CRITIC_STICK_PATH: v = stick_sum / (double)W.plan_n; break;
            case B200LP_CRITIC_TOWARD_GLOBAL_PLAN: v = (double)last_nn * cr.weight; break;
            case B200LP_CRITIC_PURE_PURSUIT: v = cr.tw * pp_dist + cr.ow * pp_yaw; break;
            default: break;
          }
        }
        if (v < 0) {
          cost = v;
          stopped = true;
        } else {
          cost += v;
        }
      }
      if (lane == 0) out_scores[rec * nc + k] = v;
    }
    if (lane == 0) {
      out_cost[rec] = cost;
      out_first_hit[rec] = first_hit;
#ifndef B200LP_NO_HIST_RECORD
      hist[rec] = max(1u, (unsigned)clock64() - W.t_item);
#endif
#ifdef B200LP_TRAJ_TRACE
      out_cost[rec] = (double)(trace_t0 - *px.t_start);             // start, ns after the cycle's first CTA
      out_first_hit[rec] = (int)(globaltimer_ns() - trace_t0);      // duration, ns
      out_scores[rec * nc] = (double)(first_hit >= 0 ? first_hit : -1);
#endif
    }
    if (fused_argmin && lane == 0) {  // running best and counters of this warp (kept out of the registers of the work loop)
      S.scored[warp] += 1;
      S.poses_scored[warp] += n;
      W.n_coll += first_hit >= 0 ? 1 : 0;
      if (cost >= 0.0 && cost <= 9999999.0) {  // local_planner.cpp:450,460
        const unsigned long long cbits = lpm::d2u(cost);
        if (W.best_cb == ~0ull || better(cbits, id, W.best_cb, W.best_id)) { W.best_cb = cbits; W.best_id = id; }
      }
    }
    __syncwarp();
  }
  if (!fused_argmin) return;

  // ---- getBestTrajectory for the single robot of this launch --------------------------------------------------
  __shared__ BlockBest s_best[kWarpsPerCta];
  __shared__ int s_last;
  __shared__ b200lp_result s_result;
  __shared__ RobotMeta s_meta;
  if (lane == 0) {
    Best wb = {W.best_cb, W.best_id, W.n_coll};
    wb.nscored = S.scored[warp];
    wb.nposes = S.poses_scored[warp];
    wb.store(&s_best[warp]);
  }
  __syncthreads();
  Best b = {~0ull, -1, 0};
  if (warp == 0) {
    if (lane < kWarpsPerCta) b.merge(s_best[lane]);
    b.warp_reduce();
    if (lane == 0) {
      b.store(partial + blockIdx.x);
      __threadfence();
      s_last = atomicAdd(tickets, 1u) == gridDim.x - 1 ? 1 : 0;
    }
  }
  __syncthreads();
  if (!s_last) return;
  // the last CTA: every other CTA has stored its partial and left the work loop; all its threads merge the partials
  __threadfence();
  b.reset();
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) {
    const uint4 r0 = __ldcg(reinterpret_cast<const uint4*>(partial + i));      // cost bits (lo, hi), id, n_collided
    const uint4 r1 = __ldcg(reinterpret_cast<const uint4*>(partial + i) + 1);  // n_scored, pad, poses (lo, hi)
    BlockBest pb;
    pb.cost_bits = ((unsigned long long)r0.y << 32) | (unsigned long long)r0.x;
    pb.id = (int)r0.z; pb.n_collided = (int)r0.w; pb.n_scored = (int)r1.x; pb.pad = 0;
    pb.poses_scored = (long long)(((unsigned long long)r1.w << 32) | (unsigned long long)r1.z);
    b.merge(pb);
  }
  b.warp_reduce();
  if (lane == 0) b.store(&s_best[warp]);
  __syncthreads();
  if (threadIdx.x == 0) {
    b.reset();
    for (int w = 0; w < kWarpsPerCta; ++w) b.merge(s_best[w]);
    RobotMeta m = meta[0];
    // what this launch really scored must be what prep_kernel listed for it: anything else means trajectories were
    // skipped (work space too small, pose rows overflowed) and the argmin cannot be trusted
    if (b.nscored != m.t_end - m.t_begin || b.nposes != m.n_poses) m.error |= 16;
    b200lp_result r;
    r.best_id = (b.cb == ~0ull) ? -1 : b.bi;
    r.n_samples = m.n_samples;
    r.n_traj = b.nscored;
    r.n_collided = b.ncoll;
    r.n_poses = b.nposes;
    r.best_cost = (b.cb == ~0ull) ? -1.0 : lpm::u2d(b.cb);
    r.xv = r.yv = r.thetav = 0.0;
    if (b.cb != ~0ull) {
      const float4 v = rec_vel[b.bi];
      r.xv = (double)v.x;
      r.yv = (C.par.theory == B200LP_THEORY_OMNI_SIMPLE) ? (double)v.y : 0.0;
      r.thetav = (double)v.z;
    }
    results[0] = r;
    *tickets = 0u;          // ready for the next launch
    *work_counter = 0ull;
    meta[0] = m;
    s_result = r;
    s_meta = m;
  }
  if (px.world > 0) {
    // sample-sharded cycle: the cross-GPU argmin through peer memory, by the first warp of this last CTA (see below)
    __syncthreads();
    if (warp == 0) peer_exchange(px, lane, s_result, s_meta, direct, direct_seq);
    return;
  }
  if (threadIdx.x == 0 && direct) {  // hand the result to the spinning host thread
    direct->r = s_result;
    direct->m = s_meta;
    direct->cycle_ns = (unsigned)min(globaltimer_ns() - *px.t_start, 0xffffffffull);
    __threadfence_system();
    *(volatile unsigned long long*)&direct->seq = direct_seq;
  }
}

// Local_Planner::getBestTrajectory (local_planner.cpp:447-480) over the costs plan_kernel wrote (fleet launches).
// grid = (B, robots): B CTAs split a robot's trajectory range, the last one to finish (ticket) merges the B
// partials and writes b200lp_result. The (cost bits, id) order is total, so any merge tree gives the reference's
// sequential `<=` scan result.
constexpr int kArgminThreads = 256;
constexpr int kArgminMaxCtas = 64;

__global__ void __launch_bounds__(kArgminThreads)
argmin_kernel(Consts C, const RobotMeta* __restrict__ meta, int t_cap, const float4* __restrict__ rec_vel,
              const double* __restrict__ cost, const int* __restrict__ first_hit, BlockBest* partial,
              unsigned* __restrict__ tickets, b200lp_result* __restrict__ results,
              unsigned long long* __restrict__ work_counter, b200lp_result* host_results, RobotMeta* host_meta,
              unsigned* __restrict__ robots_done, unsigned long long* host_seq, unsigned long long seq) {
  __shared__ BlockBest s_best[kArgminThreads / 32];
  __shared__ int s_last;
  const int robot = blockIdx.y, B = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RobotMeta m = meta[robot];
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *work_counter = 0ull;  // ready for the next plan_kernel
  const int n_local = m.t_end - m.t_begin;
  const int per = (n_local + B - 1) / B;
  const int lo = m.t_begin + (int)blockIdx.x * per, hi = min(m.t_end, lo + per);
  Best b = {~0ull, -1, 0};
  for (int id0 = lo + (int)threadIdx.x; id0 < hi; id0 += kArgminThreads * 4) {
    double c[4];
    int fh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // independent loads first
      const int id = id0 + u * kArgminThreads;
      const size_t rec = (size_t)robot * t_cap + id;
      c[u] = id < hi ? cost[rec] : -1.0;
      fh[u] = id < hi ? first_hit[rec] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      b.ncoll += fh[u] >= 0 ? 1 : 0;
      if (c[u] >= 0.0 && c[u] <= 9999999.0) b.take(lpm::d2u(c[u]), id0 + u * kArgminThreads);  // local_planner.cpp:450,460
    }
  }
  b.warp_reduce();
  if (lane == 0) b.store(&s_best[warp]);
  __syncthreads();
  if (warp == 0) {
    b.reset();
    if (lane < kArgminThreads / 32) b.merge(s_best[lane]);
    b.warp_reduce();
    if (lane == 0) {
      s_last = 1;
      if (B > 1) {
        b.store(&partial[(size_t)robot * B + blockIdx.x]);
        __threadfence();
        s_last = atomicAdd(tickets + robot, 1u) == (unsigned)(B - 1) ? 1 : 0;
      }
    }
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  if (B > 1) {  // merge the CTA partials
    __threadfence();
    b.reset();
    for (int i = lane; i < B; i += 32) {
      const volatile BlockBest* pb = &partial[(size_t)robot * B + i];
      b.ncoll += pb->n_collided;
      b.take(pb->cost_bits, pb->id);
    }
    b.warp_reduce();
  }
  if (lane == 0) {
    b200lp_result r;
    r.best_id = (b.cb == ~0ull) ? -1 : b.bi;
    r.n_samples = m.n_samples;
    r.n_traj = n_local;
    r.n_collided = b.ncoll;
    r.n_poses = m.n_poses;
    r.best_cost = (b.cb == ~0ull) ? -1.0 : lpm::u2d(b.cb);
    r.xv = r.yv = r.thetav = 0.0;
    if (b.cb != ~0ull) {
      const float4 v = rec_vel[(size_t)robot * t_cap + b.bi];
      r.xv = (double)v.x;
      r.yv = (C.par.theory == B200LP_THEORY_OMNI_SIMPLE) ? (double)v.y : 0.0;
      r.thetav = (double)v.z;
    }
    results[robot] = r;
    if (B > 1) tickets[robot] = 0u;  // ready for the next launch
    // Fleet results go straight into mapped pinned host memory, like a single robot's (DirectOut): the robot that
    // finishes last raises the sequence word the host thread spins on — no read-back copies, no stream synchronisation.
    if (host_results) {
      host_results[robot] = r;
      host_meta[robot] = m;
      __threadfence_system();  // this robot's block is visible to the host before it counts as done
      if (atomicAdd(robots_done, 1u) == gridDim.y - 1u) {
        *robots_done = 0u;  // ready for the next launch
        __threadfence_system();
        *(volatile unsigned long long*)host_seq = seq;
      }
    }
  }
}

// =============================================================================================
// the steps either side of the path
// =============================================================================================
struct PruneMeta {
  b200lp_prune_info info;
  int32_t overflow, pad;
};

// Local_Planner::prunePlan (local_planner.cpp:374-445), one CTA. The prune plan is two contiguous index ranges of the
// global plan: [nn - nb + 1, nn] (the backward walk, reversed at :418) followed by [nn, nn + nf - 1].
__global__ void __launch_bounds__(256) prune_kernel(const double* __restrict__ g7, int n, double rx, double ry, double rz,
                                                    double forward_distance, double backward_distance, int cap,
                                                    double* __restrict__ plan7, float4* __restrict__ pcl,
                                                    PruneMeta* __restrict__ meta) {
  __shared__ float s_d[8];
  __shared__ int s_i[8];
  __shared__ int s_nn, s_nb, s_nf, s_status;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // nearestKSearch(robot_pose, 1): float positions, L2_Simple, lowest index among equal distances
  const float qx = (float)rx, qy = (float)ry, qz = (float)rz;
  float best = 3.402823466e+38f;
  int bi = 0x7fffffff;
  for (int i = tid; i < n; i += 256) {
    const float d = l2_simple(qx, qy, qz, (float)g7[i * 7], (float)g7[i * 7 + 1], (float)g7[i * 7 + 2]);
    if (d < best) { best = d; bi = i; }  // ascending i per thread: strict < keeps the lowest index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(kFull, best, o);
    const int oi = __shfl_xor_sync(kFull, bi, o);
    if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
  }
  if (lane == 0) { s_d[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_d[w] < best || (s_d[w] == best && s_i[w] < bi)) { best = s_d[w]; bi = s_i[w]; }
    int status = 0, nb = 0, nf = 0;
    if ((double)lpm::fsqrt(best) > 1.0) {
      status = 2;  // deviated from the plan: the prune plan stays cleared
    } else {
      const int nn = bi;
      auto dist = [&](int a, int b) {  // getDistanceBTWPoseStamp
        const double dx = g7[a * 7] - g7[b * 7], dy = g7[a * 7 + 1] - g7[b * 7 + 1], dz = g7[a * 7 + 2] - g7[b * 7 + 2];
        return lpm::dsqrt((dx * dx + dy * dy) + dz * dz);
      };
      int last = nn;
      for (int i = nn; i >= 0; --i) {
        ++nb;
        if (i < nn) backward_distance -= dist(last, i);
        last = i;
        if (backward_distance < 0) break;
      }
      for (int i = nn; i < n; ++i) {
        ++nf;
        if (i > nn) forward_distance -= dist(last, i);
        last = i;
        if (forward_distance < 0) break;
      }
    }
    s_nn = bi; s_nb = nb; s_nf = nf; s_status = status;
    PruneMeta m;
    m.info.status = status;
    m.info.nearest_index = bi;
    m.info.n_prune = nb + nf;
    m.info.n_backward = nb;
    m.overflow = (nb + nf > cap) ? 1 : 0;
    m.pad = 0;
    *meta = m;
  }
  __syncthreads();
  const int nn = s_nn, nb = s_nb, nf = s_nf;
  if (s_status != 0 || nb + nf > cap) return;
  for (int k = tid; k < nb + nf; k += 256) {
    // prune_plan_.poses order
    const int src = k < nb ? nn - nb + 1 + k : nn + (k - nb);
#pragma unroll
    for (int a = 0; a < 7; ++a) plan7[k * 7 + a] = g7[src * 7 + a];
    // pcl_prune_plan_ order: the backward walk as pushed (nn, nn-1, ...), then the forward walk; intensity tags :407,:424-429
    const int src2 = k < nb ? nn - k : nn + (k - nb);
    const float tag = k < nb ? -1.f : (src2 == 0 ? 0.f : 1.f);
    pcl[k] = make_float4((float)g7[src2 * 7], (float)g7[src2 * 7 + 1], (float)g7[src2 * 7 + 2], tag);
  }
}

// PathBlockedStrategy::selfMark (path_blocked_strategy.cpp:56-100): one warp per prune-plan point; the lanes scan the
// cell rows overlapping the check ball for ANY point with float d^2 < r^2.
__global__ void __launch_bounds__(256) blocked_kernel(GridDev g, const float4* __restrict__ pcl, int n, float r, float r2,
                                                      int* __restrict__ n_blocked, int* __restrict__ n_checked) {
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= n) return;
  const float4 q = pcl[i];
  if (q.w < 0.f) return;  // backward poses are skipped
  if (lane == 0) atomicAdd(n_checked, 1);
  if (g.n_kept == 0) return;
  // cells that can hold a point with d^2 < r^2: |p - q| < r per axis, padded for the float rounding of d^2
  const float rr = r * 1.0001f + 1e-3f + 1e-5f * g.cmax;
  const float fx0 = cell_f(q.x - rr, g.org[0], g.inv_xy), fx1 = cell_f(q.x + rr, g.org[0], g.inv_xy);
  const float fy0 = cell_f(q.y - rr, g.org[1], g.inv_xy), fy1 = cell_f(q.y + rr, g.org[1], g.inv_xy);
  const float fz0 = cell_f(q.z - rr, g.org[2], g.inv_z), fz1 = cell_f(q.z + rr, g.org[2], g.inv_z);
  if (fx1 < 0.f || fy1 < 0.f || fz1 < 0.f || fx0 > (float)(g.nx - 1) || fy0 > (float)(g.ny - 1) || fz0 > (float)(g.nz - 1)) return;
  const int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, (float)(g.nx - 1));
  const int y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, (float)(g.ny - 1));
  const int z0 = (int)fmaxf(fz0, 0.f), z1 = (int)fminf(fz1, (float)(g.nz - 1));
  bool hit = false;
  for (int iz = z0; iz <= z1 && !hit; ++iz)
    for (int iy = y0; iy <= y1 && !hit; ++iy) {
      const size_t base = ((size_t)iz * g.ny + iy) * (size_t)g.nx;
      const uint32_t b = g.cell_start[base + x0], e = g.cell_start[base + x1 + 1];
      for (uint32_t j0 = b; j0 < e && !hit; j0 += 32) {
        const uint32_t j = j0 + lane;
        bool in = false;
        if (j < e) {
          const float4 p = __ldg(g.pts + j);
          in = l2_simple(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
        }
        hit = __any_sync(kFull, in);
      }
    }
  if (hit && lane == 0) atomicAdd(n_blocked, 1);
}

// =============================================================================================
// read-back / diagnostics
// =============================================================================================
// lane-per-pose scan of every cell overlapping the 1 m ball: n_r1 and (optionally) the exact collision flag
__device__ __forceinline__ void ball_scan(const GridDev& g, const float* stash, int col, int mode /*-1 none,0 box,1 minmax*/,
                                          int* n_r1, int* collide) {
  int cnt = 0, hit = 0;
  const float qx = stash[F_PX * 32 + col], qy = stash[F_PY * 32 + col], qz = stash[F_PZ * 32 + col];
  const float r = 1.0f + 1e-3f + 1e-5f * g.cmax;
  const float fx0 = cell_f(qx - r, g.org[0], g.inv_xy), fx1 = cell_f(qx + r, g.org[0], g.inv_xy);
  const float fy0 = cell_f(qy - r, g.org[1], g.inv_xy), fy1 = cell_f(qy + r, g.org[1], g.inv_xy);
  const float fz0 = cell_f(qz - r, g.org[2], g.inv_z), fz1 = cell_f(qz + r, g.org[2], g.inv_z);
  if (!(fx1 < 0.f || fy1 < 0.f || fz1 < 0.f || fx0 > (float)(g.nx - 1) || fy0 > (float)(g.ny - 1) ||
        fz0 > (float)(g.nz - 1)) && g.n_kept > 0) {
    const int ix0 = (int)fmaxf(fx0, 0.f), ix1 = (int)fminf(fx1, (float)(g.nx - 1));
    const int iy0 = (int)fmaxf(fy0, 0.f), iy1 = (int)fminf(fy1, (float)(g.ny - 1));
    const int iz0 = (int)fmaxf(fz0, 0.f), iz1 = (int)fminf(fz1, (float)(g.nz - 1));
    for (int iz = iz0; iz <= iz1; ++iz)
      for (int iy = iy0; iy <= iy1; ++iy) {
        const size_t base = ((size_t)iz * g.ny + iy) * (size_t)g.nx;
        const uint32_t b = g.cell_start[base + ix0], e = g.cell_start[base + ix1 + 1];
        for (uint32_t j = b; j < e; ++j) {
          const float4 p = g.pts[j];
          if (l2_simple(qx, qy, qz, p.x, p.y, p.z) < 1.0f) {
            ++cnt;
            if (mode == 0 && exact_in_box(stash, col, p.x, p.y, p.z)) hit = 1;
            if (mode == 1 && exact_in_aabb(stash, col, p.x, p.y, p.z)) hit = 1;
          }
        }
      }
  }
  *n_r1 = cnt;
  *collide = hit;
}

// one warp (= one CTA) recomputes one trajectory and writes every per-pose quantity the reference's
// Trajectory holds; CTA b handles trajectory id0 + b and writes its poses at row pose_off[b]
__global__ void __launch_bounds__(32) poses_kernel(Consts C, GridDev g, const RobotIn* __restrict__ robots, int robot,
                                                   int t_cap, int id0, const long long* __restrict__ pose_off,
                                                   const float4* __restrict__ rec_vel,
                                                   const int* __restrict__ rec_steps, const double* __restrict__ rec_dt,
                                                   double* __restrict__ o_pose, float* __restrict__ o_pcl,
                                                   float* __restrict__ o_cuboid, float* __restrict__ o_aabb,
                                                   unsigned char* __restrict__ o_collide, int* __restrict__ o_nr1) {
  __shared__ float stash[F_COUNT * 32];
  __shared__ double R0[9], t0[3];
  const int lane = threadIdx.x;
  const int id = id0 + blockIdx.x;
  const long long row0 = pose_off[blockIdx.x];
  const RobotIn& q = robots[robot];
  if (lane == 0) {
    quat_to_matrix(q.pose[3], q.pose[4], q.pose[5], q.pose[6], R0);
    t0[0] = q.pose[0]; t0[1] = q.pose[1]; t0[2] = q.pose[2];
  }
  __syncwarp();
  const size_t rec = (size_t)robot * t_cap + id;
  const float4 vel = rec_vel[rec];
  const int n = rec_steps[rec];
  const double dt = rec_dt[rec];
  int mode = -1;
  for (int k = C.n_critics - 1; k >= 0; --k) {
    if (C.critics[k].kind == B200LP_CRITIC_COLLISION) mode = 0;
    if (C.critics[k].kind == B200LP_CRITIC_COLLISION_MIN_MAX) mode = 1;
  }
  Carry carry = {0.f, 0.f, 0.f};
  for (int base = 0; base < n; base += 32) {
    float px, py, pth;
    rollout32(carry, lane, C.par.theory, vel.x, vel.y, vel.z, dt, px, py, pth);
    const bool live = base + lane < n;
    const long long k = row0 + base + lane;
    double L[9], t[3];
    pose_affine(R0, t0, px, py, pth, L, t);
    float verts[24];
    pose_geometry(C, g, L, t, stash, nullptr, nullptr, lane, live, verts);
    __syncwarp();
    if (live) {
      if (o_pose) {
        double qd[4];
        matrix_to_quat(L, qd);  // tf2::eigenToTransform (dd_simple…cpp:434)
        o_pose[k * 7 + 0] = t[0]; o_pose[k * 7 + 1] = t[1]; o_pose[k * 7 + 2] = t[2];
        o_pose[k * 7 + 3] = qd[0]; o_pose[k * 7 + 4] = qd[1]; o_pose[k * 7 + 5] = qd[2]; o_pose[k * 7 + 6] = qd[3];
      }
      for (int a = 0; a < 3; ++a) {
        if (o_pcl) o_pcl[k * 3 + a] = stash[(F_PX + a) * 32 + lane];
        if (o_aabb) {
          o_aabb[k * 6 + a] = stash[(F_MNX + a) * 32 + lane];
          o_aabb[k * 6 + 3 + a] = stash[(F_MXX + a) * 32 + lane];
        }
      }
      if (o_cuboid)
        for (int j = 0; j < 24; ++j) o_cuboid[k * 24 + j] = verts[j];
      if (o_nr1 || o_collide) {
        int nr1 = 0, col = 0;
        if (g.n_raw >= 5) ball_scan(g, stash, lane, mode, &nr1, &col);
        if (o_nr1) o_nr1[k] = nr1;
        if (o_collide) o_collide[k] = (unsigned char)col;
      }
    }
    __syncwarp();
  }
}

// sum over all poses of the launch's trajectory range of |radiusSearch(pose, 1.0)|
__global__ void __launch_bounds__(kThreads) count_radius_kernel(Consts C, GridDev g, const RobotIn* __restrict__ robots,
                                                                const RobotMeta* __restrict__ meta, int t_cap,
                                                                const float4* __restrict__ rec_vel,
                                                                const int* __restrict__ rec_steps,
                                                                const double* __restrict__ rec_dt,
                                                                unsigned long long* __restrict__ out /* [2]: sum, poses */) {
  __shared__ float s_stash[kWarpsPerCta][F_COUNT * 32];
  __shared__ double R0[9], t0[3];
  const int robot = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RobotMeta m = meta[robot];
  const RobotIn& q = robots[robot];
  if (threadIdx.x == 0) {
    quat_to_matrix(q.pose[3], q.pose[4], q.pose[5], q.pose[6], R0);
    t0[0] = q.pose[0]; t0[1] = q.pose[1]; t0[2] = q.pose[2];
  }
  __syncthreads();
  const int local = blockIdx.x * kWarpsPerCta + warp;
  if (local >= m.t_end - m.t_begin) return;
  const size_t rec = (size_t)robot * t_cap + m.t_begin + local;
  const float4 vel = rec_vel[rec];
  const int n = rec_steps[rec];
  const double dt = rec_dt[rec];
  float* stash = s_stash[warp];
  Carry carry = {0.f, 0.f, 0.f};
  unsigned long long sum = 0ull;
  for (int base = 0; base < n; base += 32) {
    float px, py, pth;
    rollout32(carry, lane, C.par.theory, vel.x, vel.y, vel.z, dt, px, py, pth);
    double L[9], t[3];
    pose_affine(R0, t0, px, py, pth, L, t);
    stash[F_PX * 32 + lane] = (float)t[0];
    stash[F_PY * 32 + lane] = (float)t[1];
    stash[F_PZ * 32 + lane] = (float)t[2];
    __syncwarp();
    if (base + lane < n && g.n_raw >= 5) {
      int nr1, col;
      ball_scan(g, stash, lane, -1, &nr1, &col);
      sum += (unsigned long long)nr1;
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  if (lane == 0) {
    atomicAdd(out, sum);
    atomicAdd(out + 1, (unsigned long long)n);
  }
}

}  // namespace lp
