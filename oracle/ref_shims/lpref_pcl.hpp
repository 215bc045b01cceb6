// lpref_pcl.hpp — the slice of PCL 1.15 the reference's path uses (TEST INFRASTRUCTURE, see lpref_eigen.hpp):
// point types with PCL's memory layout, PointCloud, transformPointCloud(Affine3d) and getMinMax3D in PCL's scalar
// arithmetic, and KdTreeFLANN's radiusSearch / nearestKSearch answered exactly (FLANN L2_Simple float
// accumulation, strict `<` radius, results sorted by distance) by the reference's own vendored nanoflann kd-tree.
#pragma once
#include <math.h>

#include <algorithm>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "lpref_eigen.hpp"
#include "nanoflann.hpp"  // /root/reference/src/dddmr_global_planner/include/global_planner (1.5.1, vendored upstream)

namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) PointXYZI {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  float intensity = 0.f, pad2_[3] = {0.f, 0.f, 0.f};
};
using pointXYZ = PointXYZ;
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32, "pcl point layout");

struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };

template <class PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
};

// pcl::transformPointCloud(in, out, Affine3d) with SSE/AVX disabled (the reference's PCL build): per point
// out.c = (float)(M(c,0)*x + M(c,1)*y + M(c,2)*z + M(c,3)), double arithmetic, left to right
template <class PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3d& T) {
  if (&in != &out) {
    out.header = in.header;
    out.is_dense = in.is_dense;
    out.points.assign(in.points.begin(), in.points.end());
    out.width = in.width;
    out.height = in.height;
  }
  for (std::size_t i = 0; i < out.points.size(); ++i) {
    const double x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    PointT& o = out.points[i];
    o.x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
    o.y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
    o.z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
  }
}

template <class PointT>
void getMinMax3D(const PointCloud<PointT>& cloud, PointT& min_pt, PointT& max_pt) {
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (const auto& p : cloud.points) {
    if (!cloud.is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  min_pt.x = mn[0]; min_pt.y = mn[1]; min_pt.z = mn[2];
  max_pt.x = mx[0]; max_pt.y = mx[1]; max_pt.z = mx[2];
}

template <class PointT>
class KdTreeFLANN {
  struct Adaptor {
    const std::vector<PointT>* pts = nullptr;
    std::size_t kdtree_get_point_count() const { return pts->size(); }
    float kdtree_get_pt(const std::size_t i, const std::size_t d) const {
      const PointT& p = (*pts)[i];
      return d == 0 ? p.x : (d == 1 ? p.y : p.z);
    }
    template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
  };
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Adaptor>, Adaptor, 3>;

 public:
  using Ptr = std::shared_ptr<KdTreeFLANN<PointT>>;
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) {
    cloud_ = cloud;
    adaptor_.pts = &cloud_->points;
    tree_.reset(new Tree(3, adaptor_, nanoflann::KDTreeSingleIndexAdaptorParams(15)));
  }
  void setInputCloud(const typename PointCloud<PointT>::Ptr& cloud) { setInputCloud(typename PointCloud<PointT>::ConstPtr(cloud)); }
  // PCL hands FLANN radius*radius as a float; FLANN admits dist < radius^2 (strict) and sorts by distance
  int radiusSearch(const PointT& q, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances,
                   unsigned int max_nn = 0) const {
    (void)max_nn;
    const float query[3] = {q.x, q.y, q.z};
    std::vector<nanoflann::ResultItem<uint32_t, float>> res;
    tree_->radiusSearch(query, static_cast<float>(radius * radius), res, nanoflann::SearchParameters(0.f, true));
    k_indices.resize(res.size());
    k_sqr_distances.resize(res.size());
    for (std::size_t i = 0; i < res.size(); ++i) { k_indices[i] = (int)res[i].first; k_sqr_distances[i] = res[i].second; }
    return (int)res.size();
  }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    const float query[3] = {q.x, q.y, q.z};
    std::vector<uint32_t> idx((std::size_t)k);
    std::vector<float> d((std::size_t)k);
    const std::size_t n = tree_->knnSearch(query, (std::size_t)k, idx.data(), d.data());
    k_indices.resize(n);
    k_sqr_distances.resize(n);
    for (std::size_t i = 0; i < n; ++i) { k_indices[i] = (int)idx[i]; k_sqr_distances[i] = d[i]; }
    return (int)n;
  }

 private:
  typename PointCloud<PointT>::ConstPtr cloud_;
  Adaptor adaptor_;
  std::unique_ptr<Tree> tree_;
};
}  // namespace pcl
