// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_kdtree.hpp — CPU restatement of util::KdTree (/root/reference/src/util/PointCloudUtils.h:346-423)
// = nanoflann 1.7.1 KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<float>, ..., 3>, leaf_max_size 10,
// single-threaded build (/root/reference/thirdparty/nanoflann/nanoflann.hpp):
//   KNNResultSet::addPoint / worstDist  :234-268      L2_Simple evalMetric / accum_dist  :638-656
//   divideTree                          :1149-1212    middleSplit_ / planeSplit          :1320-1427
//   computeInitialDistances             :1429-1452    findNeighbors / knnSearch          :1708-1756
//   init_vind / computeBoundingBox      :1836-1873    searchLevel                        :1885-1960
// Cross-checked against the vendored header by oracle/ref_check (tests/test_oracle_pins.py).
#pragma once
#include <cstdint>
#include <limits>
#include <vector>
#include "orc_voxel.hpp"

namespace orc {

class KdTree {
 public:
  void setInputCloud(const std::vector<P3>& cloud) {  // PointCloudUtils.h:379-393
    pts_ = cloud;
    nodes_.clear(); vacc_.clear(); root_ = -1;
    if (pts_.empty()) return;
    vacc_.resize(pts_.size());
    for (uint32_t i = 0; i < (uint32_t)pts_.size(); ++i) vacc_[i] = i;
    for (int d = 0; d < 3; ++d) root_bbox_[d].low = root_bbox_[d].high = get(vacc_[0], d);
    for (size_t k = 1; k < pts_.size(); ++k)
      for (int d = 0; d < 3; ++d) { float v = get(vacc_[k], d); if (v < root_bbox_[d].low) root_bbox_[d].low = v; if (v > root_bbox_[d].high) root_bbox_[d].high = v; }
    nodes_.reserve(pts_.size() / 4 + 8);
    root_ = divideTree(0, pts_.size(), root_bbox_);
  }
  bool built() const { return root_ >= 0; }
  size_t size() const { return pts_.size(); }
  const std::vector<uint32_t>& vacc() const { return vacc_; }

  // knnSearch (:1743-1751): returns number of results; indices/dists ascending by f32 squared L2.
  size_t knnSearch(const float* q, size_t k, uint32_t* out_idx, float* out_d) const {
    if (root_ < 0 || k == 0) return 0;
    Result rs{out_idx, out_d, k, 0};
    if (k) out_d[k - 1] = std::numeric_limits<float>::max();  // KNNResultSet::init
    float dists[3] = {0, 0, 0};
    float dist = 0;
    for (int i = 0; i < 3; ++i) {
      if (q[i] < root_bbox_[i].low) { dists[i] = (q[i] - root_bbox_[i].low) * (q[i] - root_bbox_[i].low); dist += dists[i]; }
      if (q[i] > root_bbox_[i].high) { dists[i] = (q[i] - root_bbox_[i].high) * (q[i] - root_bbox_[i].high); dist += dists[i]; }
    }
    searchLevel(rs, q, root_, dist, dists);
    return rs.count;
  }

 private:
  struct Interval { float low, high; };
  struct Node { int child1 = -1, child2 = -1; size_t left = 0, right = 0; int divfeat = 0; float divlow = 0, divhigh = 0; };
  struct Result {
    uint32_t* idx; float* d; size_t cap, count;
    float worst() const { return (count < cap || !count) ? std::numeric_limits<float>::max() : d[count - 1]; }
    void add(float dist, uint32_t index) {
      size_t i;
      for (i = count; i > 0; --i) {
        if (d[i - 1] > dist) { if (i < cap) { d[i] = d[i - 1]; idx[i] = idx[i - 1]; } }
        else break;
      }
      if (i < cap) { d[i] = dist; idx[i] = index; }
      if (count < cap) count++;
    }
  };
  std::vector<P3> pts_;
  std::vector<uint32_t> vacc_;
  std::vector<Node> nodes_;
  Interval root_bbox_[3];
  int root_ = -1;

  float get(uint32_t i, int d) const { return d == 0 ? pts_[i].x : (d == 1 ? pts_[i].y : pts_[i].z); }

  int divideTree(size_t left, size_t right, Interval* bbox) {
    int me = (int)nodes_.size();
    nodes_.emplace_back();
    if ((right - left) <= 10) {
      nodes_[me].left = left; nodes_[me].right = right;
      for (int i = 0; i < 3; ++i) bbox[i].low = bbox[i].high = get(vacc_[left], i);
      for (size_t k = left + 1; k < right; ++k)
        for (int i = 0; i < 3; ++i) { float v = get(vacc_[k], i); if (bbox[i].low > v) bbox[i].low = v; if (bbox[i].high < v) bbox[i].high = v; }
    } else {
      size_t idx; int cutfeat; float cutval;
      middleSplit(left, right - left, idx, cutfeat, cutval, bbox);
      nodes_[me].divfeat = cutfeat;
      Interval lb[3] = {bbox[0], bbox[1], bbox[2]};
      lb[cutfeat].high = cutval;
      int c1 = divideTree(left, left + idx, lb);
      Interval rb[3] = {bbox[0], bbox[1], bbox[2]};
      rb[cutfeat].low = cutval;
      int c2 = divideTree(left + idx, right, rb);
      nodes_[me].child1 = c1; nodes_[me].child2 = c2;
      nodes_[me].divlow = lb[cutfeat].high;
      nodes_[me].divhigh = rb[cutfeat].low;
      for (int i = 0; i < 3; ++i) { bbox[i].low = std::min(lb[i].low, rb[i].low); bbox[i].high = std::max(lb[i].high, rb[i].high); }
    }
    return me;
  }
  void middleSplit(size_t ind, size_t count, size_t& index, int& cutfeat, float& cutval, const Interval* bbox) {
    const float EPS = 0.00001f;
    float max_span = bbox[0].high - bbox[0].low;
    for (int i = 1; i < 3; ++i) { float span = bbox[i].high - bbox[i].low; if (span > max_span) max_span = span; }
    float max_spread = -1;
    cutfeat = 0;
    float min_elem = 0, max_elem = 0;
    for (int i = 0; i < 3; ++i) {
      float span = bbox[i].high - bbox[i].low;
      if (span >= (1 - EPS) * max_span) {
        float mn = get(vacc_[ind], i), mx = mn;
        for (size_t k = 1; k < count; ++k) { float v = get(vacc_[ind + k], i); if (v < mn) mn = v; if (v > mx) mx = v; }
        float spread = mx - mn;
        if (spread > max_spread) { cutfeat = i; max_spread = spread; min_elem = mn; max_elem = mx; }
      }
    }
    float split_val = (bbox[cutfeat].low + bbox[cutfeat].high) / 2;
    if (split_val < min_elem) cutval = min_elem;
    else if (split_val > max_elem) cutval = max_elem;
    else cutval = split_val;
    size_t lim1, lim2;
    planeSplit(ind, count, cutfeat, cutval, lim1, lim2);
    if (lim1 > count / 2) index = lim1;
    else if (lim2 < count / 2) index = lim2;
    else index = count / 2;
  }
  void planeSplit(size_t ind, size_t count, int cutfeat, float cutval, size_t& lim1, size_t& lim2) {
    size_t left = 0, right = count - 1;
    for (;;) {
      while (left <= right && get(vacc_[ind + left], cutfeat) < cutval) ++left;
      while (right && left <= right && get(vacc_[ind + right], cutfeat) >= cutval) --right;
      if (left > right || !right) break;
      std::swap(vacc_[ind + left], vacc_[ind + right]);
      ++left; --right;
    }
    lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && get(vacc_[ind + left], cutfeat) <= cutval) ++left;
      while (right && left <= right && get(vacc_[ind + right], cutfeat) > cutval) --right;
      if (left > right || !right) break;
      std::swap(vacc_[ind + left], vacc_[ind + right]);
      ++left; --right;
    }
    lim2 = left;
  }
  void searchLevel(Result& rs, const float* vec, int ni, float mindist, float* dists) const {
    const Node& node = nodes_[ni];
    if (node.child1 < 0 && node.child2 < 0) {
      float worst = rs.worst();
      for (size_t i = node.left; i < node.right; ++i) {
        uint32_t a = vacc_[i];
        float dist = 0;
        for (int d = 0; d < 3; ++d) { const float diff = vec[d] - get(a, d); dist += diff * diff; }
        if (dist < worst) rs.add(dist, a);
      }
      return;
    }
    int idx = node.divfeat;
    float val = vec[idx];
    float diff1 = val - node.divlow, diff2 = val - node.divhigh;
    int best, other; float cut_dist;
    if ((diff1 + diff2) < 0) { best = node.child1; other = node.child2; cut_dist = (val - node.divhigh) * (val - node.divhigh); }
    else { best = node.child2; other = node.child1; cut_dist = (val - node.divlow) * (val - node.divlow); }
    searchLevel(rs, vec, best, mindist, dists);
    float dst = dists[idx];
    mindist = mindist + cut_dist - dst;
    dists[idx] = cut_dist;
    if (mindist * 1.0f <= rs.worst()) searchLevel(rs, vec, other, mindist, dists);
    dists[idx] = dst;
  }
};

}  // namespace orc
