// lp_observe.cuh — the observation producer in front of the path: what
// perception_3d::MultiLayerSpinningLidar::cbSensor does to one lidar scan before the local planner sees it
// (dddmr_perception_3d/plugins/multilayer_spinning_lidar.cpp:232-269):
//
//   pcl::transformPointCloud(scan, base_link<-sensor)           :232-233
//   pcl::PassThrough x, y in [-window, window], z in [0, marking_height]   :240-251
//   pcl::VoxelGrid leaf 0.1 (centroid per voxel, voxels in ascending index order)   :253-256
//   pcl::transformPointCloud(map<-base_link) when is_local_planner           :264-268
//
// Device formulation: one pass turns every scan point into a 16-byte record (x, y, z in base_link, voxel key) and counts
// the first radix digit; a STABLE least-significant-digit radix sort (10-bit digits, warp match-any ranking) orders the
// surviving records by voxel key and, inside a voxel, by scan index; one thread per voxel head then adds the voxel's
// points in that order (pcl::CentroidPoint's float accumulation), divides, applies the second transform and writes the
// observation as pcl::PointXYZ-layout float4 (x, y, z, 1). Nothing here needs the data's bounding box: the key is built on
// the pass-through window (known from the parameters), which orders and groups voxels exactly like PCL's
// data-dependent `ijk0 + ijk1*div_b[0] + ijk2*div_b[0]*div_b[1]`.
//
// This translation unit is compiled with -fmad=false (see lp_device.cuh): the double transform and the float sums
// round like the CPU oracle's.
#pragma once
#include "lp_kernels.cuh"

namespace lp {

constexpr int kObsThreads = 256;
// records per thread in the sort passes (template parameter kItems): 16 (4096 records per CTA) for scans of millions of
// points, 4 (1024 per CTA) below that so that a lidar revolution of a few hundred thousand points still covers every SM
constexpr int kObsItemsLarge = 16, kObsItemsSmall = 4;
constexpr size_t kObsLargeScan = (size_t)1 << 21;
constexpr int kObsMaxBits = 10;                       // radix digit width (<= 10: 8 warps x 1024 counters = 32 KB smem)
constexpr int kObsHeadTile = 1024;                    // records per CTA in the head / centroid passes
constexpr uint32_t kObsInvalid = 0xffffffffu;         // key of a point the pass-through filters dropped

struct ObsDev {
  double m1[12];  // base_link <- sensor, row-major 3x4 (tf2::transformToEigen)
  double m2[12];  // map <- base_link
  float lo[3], hi[3];  // pass-through limits (inclusive), as the floats pcl::PassThrough::setFilterLimits stores
  float inv_leaf;      // 1.0f / leaf
  int lb[3];           // floor(lo * inv_leaf): voxel coordinate of the window's lower corner
  uint32_t d0, d1;     // voxel-coordinate extents of the window in x and y
  int apply_m2;        // is_local_planner_
};

// pcl::transformPointCloud(cloud, cloud, Affine3d) for one point: double arithmetic left to right, rounded to float.
__device__ __forceinline__ float3 obs_transform(const double* m, float3 p) {
  const double x = p.x, y = p.y, z = p.z;
  float3 r;
  r.x = (float)(m[0] * x + m[1] * y + m[2] * z + m[3]);
  r.y = (float)(m[4] * x + m[5] * y + m[6] * z + m[7]);
  r.z = (float)(m[8] * x + m[9] * y + m[10] * z + m[11]);
  return r;
}

// Pass 1: transform, pass-through, voxel key; per-CTA histogram of the first radix digit.
// hist layout: [digit][cta] (digit-major), so that one exclusive scan yields every (digit, cta) output offset.
template <int kItems>
__global__ void __launch_bounds__(kObsThreads) obs_key_kernel(const char* __restrict__ raw, size_t n, size_t stride, ObsDev P,
                                                              int bits, float4* __restrict__ rec, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[1 << kObsMaxBits];
  const int nbins = 1 << bits;
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) s_h[b] = 0u;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * (kObsThreads * kItems);
  const uint32_t mask = (uint32_t)nbins - 1u;
#pragma unroll 4
  for (int k = 0; k < kItems; ++k) {
    const size_t i = base + (size_t)k * kObsThreads + threadIdx.x;
    if (i >= n) break;
    const float3 v = obs_transform(P.m1, load_xyz(raw, i, stride));
    uint32_t key = kObsInvalid;
    // pcl::PassThrough: non-finite points and field values outside [min, max] are removed
    if (finite3(v) && !(v.x < P.lo[0] || v.x > P.hi[0]) && !(v.y < P.lo[1] || v.y > P.hi[1]) && !(v.z < P.lo[2] || v.z > P.hi[2])) {
      const int c0 = (int)floorf(v.x * P.inv_leaf) - P.lb[0];
      const int c1 = (int)floorf(v.y * P.inv_leaf) - P.lb[1];
      const int c2 = (int)floorf(v.z * P.inv_leaf) - P.lb[2];
      key = ((uint32_t)c2 * P.d1 + (uint32_t)c1) * P.d0 + (uint32_t)c0;
      atomicAdd(&s_h[key & mask], 1u);
    }
    rec[i] = make_float4(v.x, v.y, v.z, __uint_as_float(key));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) hist[(size_t)b * gridDim.x + blockIdx.x] = s_h[b];
}

// Per-CTA digit histogram of a later pass (records are compact by then: n_kept of them, count read from the device).
template <int kItems>
__global__ void __launch_bounds__(kObsThreads) obs_hist_kernel(const float4* __restrict__ rec, const uint32_t* __restrict__ n_ptr,
                                                               int shift, int bits, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[1 << kObsMaxBits];
  const int nbins = 1 << bits;
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) s_h[b] = 0u;
  __syncthreads();
  const size_t n = *n_ptr, base = (size_t)blockIdx.x * (kObsThreads * kItems);
  const uint32_t mask = (uint32_t)nbins - 1u;
  if (base < n) {
#pragma unroll 4
    for (int k = 0; k < kItems; ++k) {
      const size_t i = base + (size_t)k * kObsThreads + threadIdx.x;
      if (i >= n) break;
      atomicAdd(&s_h[(__float_as_uint(__ldg(rec + i).w) >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) hist[(size_t)b * gridDim.x + blockIdx.x] = s_h[b];
}

// Stable scatter of one radix pass. `hist` holds the block-local exclusive scan of the [digit][cta] counts and
// exclusively scanned in place by scan_kernel, so the first output slot of (digit, cta) is hist[i]. Warp w of a CTA owns the contiguous sub-tile
// [w * 32 * kItems, (w+1) * 32 * kItems) of the CTA's tile and walks it 32 records at a time: a record's rank among
// the records of its digit is (same digit in earlier warps) + (same digit earlier in this warp) + (same digit in lower
// lanes of this row) — the three terms come from a cross-warp scan of per-warp counters, the counter value when the row
// is processed, and __match_any_sync. Input order is preserved inside every digit: the sort is stable.
// first_pass: n is the scan size and dropped records (key == kObsInvalid) are skipped; otherwise n comes from n_ptr.
template <int kItems>
__global__ void __launch_bounds__(kObsThreads) obs_scatter_kernel(const float4* __restrict__ in, size_t n_static,
                                                                  const uint32_t* __restrict__ n_ptr, int first_pass, int shift,
                                                                  int bits, const uint32_t* __restrict__ hist,
                                                                  float4* __restrict__ out) {
  __shared__ uint32_t s_wh[kObsThreads / 32][1 << kObsMaxBits];
  const size_t n = first_pass ? n_static : (size_t)*n_ptr;
  const size_t tile = (size_t)blockIdx.x * (kObsThreads * kItems);
  if (tile >= n) return;
  const int nbins = 1 << bits, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t mask = (uint32_t)nbins - 1u;
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) {
#pragma unroll
    for (int w = 0; w < kObsThreads / 32; ++w) s_wh[w][b] = 0u;
  }
  __syncthreads();
  const size_t wbase = tile + (size_t)warp * 32 * kItems;
  uint32_t packed[kItems];  // digit << 16 | rank inside the warp's sub-tile; ~0 = not a live record
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const size_t i = wbase + (size_t)r * 32 + lane;
    uint32_t key = kObsInvalid;
    if (i < n) key = __float_as_uint(__ldg(in + i).w);
    const bool live = i < n && (!first_pass || key != kObsInvalid);
    const uint32_t d = live ? ((key >> shift) & mask) : (uint32_t)nbins;  // dead lanes share a digit nobody counts
    const unsigned peers = __match_any_sync(kFull, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0u;
    if (live && lane == leader) {
      old = s_wh[warp][d];
      s_wh[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(kFull, old, leader);
    packed[r] = live ? ((d << 16) | (old + (uint32_t)__popc(peers & ((1u << lane) - 1u)))) : 0xffffffffu;
    __syncwarp();
  }
  __syncthreads();
  // cross-warp exclusive scan per digit, seeded with the (digit, cta) global offset
  for (int b = threadIdx.x; b < nbins; b += kObsThreads) {
    const size_t hi = (size_t)b * gridDim.x + blockIdx.x;
    uint32_t run = hist[hi];
#pragma unroll
    for (int w = 0; w < kObsThreads / 32; ++w) {
      const uint32_t c = s_wh[w][b];
      s_wh[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if (packed[r] == 0xffffffffu) continue;
    const size_t i = wbase + (size_t)r * 32 + lane;
    out[s_wh[warp][packed[r] >> 16] + (packed[r] & 0xffffu)] = __ldg(in + i);
  }
}

__device__ __forceinline__ bool obs_is_head(const float4* __restrict__ rec, size_t i) {
  return i == 0 || __float_as_uint(__ldg(rec + i).w) != __float_as_uint(__ldg(rec + i - 1).w);
}

// Voxel heads (first record of every run of equal keys) per tile of kObsHeadTile sorted records.
__global__ void __launch_bounds__(kObsThreads) obs_heads_kernel(const float4* __restrict__ rec, const uint32_t* __restrict__ n_ptr,
                                                                uint32_t* __restrict__ tile_heads) {
  __shared__ uint32_t s_w[kObsThreads / 32];
  const size_t n = *n_ptr, base = (size_t)blockIdx.x * kObsHeadTile + (size_t)threadIdx.x * 4;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n) c += obs_is_head(rec, base + k) ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kObsThreads / 32; ++w) t += s_w[w];
    tile_heads[blockIdx.x] = t;
  }
}

// One thread per voxel head: pcl::CentroidPoint (AccumulatorXYZ: Eigen::Vector3f sum in record order, then / n),
// the map<-base_link transform, output slot = number of heads before this one (voxels leave in ascending key order,
// like VoxelGrid's sorted index vector). counts[0] = n_kept (in), counts[1] = number of voxels (out).
__global__ void __launch_bounds__(kObsThreads) obs_centroid_kernel(const float4* __restrict__ rec, uint32_t* __restrict__ counts,
                                                                   const uint32_t* __restrict__ tile_heads, ObsDev P,
                                                                   float4* __restrict__ out) {
  __shared__ uint32_t s_w[kObsThreads / 32];
  __shared__ uint32_t s_prefix;
  const size_t n = counts[0];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // heads in earlier tiles
  uint32_t pre = 0;
  for (unsigned j = threadIdx.x; j < blockIdx.x; j += kObsThreads) pre += tile_heads[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(kFull, pre, o);
  if (lane == 0) s_w[warp] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kObsThreads / 32; ++w) t += s_w[w];
    s_prefix = t;
  }
  __syncthreads();
  const uint32_t prefix = s_prefix;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kObsHeadTile + (size_t)threadIdx.x * 4;
  bool head[4];
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    head[k] = base + k < n && obs_is_head(rec, base + k);
    c += head[k] ? 1u : 0u;
  }
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_w[w];
  uint32_t slot = prefix + woff + incl - c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!head[k]) continue;
    const size_t i0 = base + k;
    const uint32_t key = __float_as_uint(__ldg(rec + i0).w);
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    uint32_t cnt = 0;
    for (size_t j = i0; j < n; ++j) {
      const float4 p = __ldg(rec + j);
      if (__float_as_uint(p.w) != key) break;
      sx += p.x;
      sy += p.y;
      sz += p.z;
      ++cnt;
    }
    const float fn = (float)cnt;
    float3 v = make_float3(sx / fn, sy / fn, sz / fn);
    if (P.apply_m2) v = obs_transform(P.m2, v);
    out[slot++] = make_float4(v.x, v.y, v.z, 1.0f);
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kObsThreads - 1) {
    uint32_t t = 0;
    for (int w = 0; w < kObsThreads / 32; ++w) t += s_w[w];
    counts[1] = prefix + t;
  }
}

// pcl::PointXYZ (16 B) -> pcl::PointXYZI (32 B: x, y, z, 1, intensity 0, padding 0) for read-back in the layout of
// Sensor::sensor_current_observation_ / SharedData::aggregate_observation_.
__global__ void __launch_bounds__(256) obs_expand_kernel(const float4* __restrict__ in, size_t n, float4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[2 * i] = __ldg(in + i);
  out[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace lp
