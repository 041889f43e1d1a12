// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_icp.hpp — CPU restatement of optimization::IterativeClosestPointOptimizer (scan-to-map path)
//   optimize                      /root/reference/src/optimization/IterativeClosestPointOptimizer.cpp:255-463
//   find_correspondences          :587-645   (surfel, O(1) lookup)
//   find_correspondences_kdtree   :647-767   (5-NN + per-query plane fit)
//   is_collinear                  :785-792
//   optimize_loop                 :40-251    (loop-closure ICP between two keyframes, 100 GN iterations, 1-NN inlier ratio)
//   find_correspondences_loop     :465-585
//   ICPConfig / OptimizationStats IterativeClosestPointOptimizer.h:55-76, :203-210
// The LidarFrame argument of the reference is replaced by the local (sensor-frame) feature cloud;
// frame->set_pose(T)/get_pose() inside the loop (:284,:606) is the identity round trip on T.
#pragma once
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>
#include "orc_eigen.hpp"
#include "orc_kdtree.hpp"
#include "orc_pko.hpp"
#include "orc_se3.hpp"
#include "orc_voxel.hpp"

namespace orc {

struct ICPConfig {  // IterativeClosestPointOptimizer.h:55-76 with the values Estimator.cpp:62-70 wires in
  int max_iterations = 4;
  double translation_tolerance = 0.005;
  double rotation_tolerance = 0.005;
  double max_correspondence_distance = 1.0;
  int min_correspondence_points = 10;
  bool use_robust_loss = true;
  double robust_loss_delta = 0.1;
  bool use_surfel_correspondence = true;
};

struct Correspondences {  // DualFrameCorrespondences, IterativeClosestPointOptimizer.h:128-144
  std::vector<std::array<double, 3>> points_last, points_curr, normals_last;
  std::vector<double> residuals;
  std::vector<int> query_index;  // (oracle extra) index of the query point of each correspondence
  void clear() { points_last.clear(); points_curr.clear(); normals_last.clear(); residuals.clear(); query_index.clear(); }
  size_t size() const { return points_last.size(); }
};

struct IterTrace {      // one Gauss-Newton iteration, for parity tests
  int n_corr = 0;
  double scale = 1.0;   // residual_normalization_scale in effect
  double delta = 0.0;   // robust kernel delta (PKO alpha) in effect
  float H[36] = {0};    // faithful sequential-f32 accumulation (as the reference)
  float g[6] = {0};
  float cost = 0;
  double H64[36] = {0}; // same terms accumulated in f64 (for tolerance-based comparison)
  double g64[6] = {0};
  double cost64 = 0;
  float dx[6] = {0};
  float T_in[16] = {0}; // pose at the start of the iteration (row-major 4x4)
  float T_out[16] = {0};
  int em_iters = 0, kmeans_iters = 0;
};

struct OptimizationStats { size_t num_correspondences = 0, num_iterations = 0; double initial_cost = 0, final_cost = 0, optimization_time_ms = 0; bool converged = false; };

class ICPOptimizer {
 public:
  ICPConfig cfg;
  std::shared_ptr<AdaptiveMEstimator> ame;
  OptimizationStats last_stats;
  std::vector<IterTrace> trace;   // filled when keep_trace
  bool keep_trace = false;

  explicit ICPOptimizer(const ICPConfig& c, std::shared_ptr<AdaptiveMEstimator> a = nullptr) : cfg(c), ame(std::move(a)) {}

  // find_correspondences (:587-645)
  static size_t find_correspondences(const VoxelMap* map, const P3* local, size_t m, const SE3f& pose, double max_dist, Correspondences& out) {
    out.clear();
    if (!map || map->empty()) return 0;
    if (!local || m == 0) return 0;
    float T[16]; pose.Matrix(T);
    for (size_t idx = 0; idx < m; ++idx) {
      float w[3];
      transform_point_4x4(T, local[idx].x, local[idx].y, local[idx].z, w);
      float nf[3], cf[3];
      if (!map->GetSurfelAtPoint(w, nf, cf)) continue;
      double n[3] = {nf[0], nf[1], nf[2]}, c[3] = {cf[0], cf[1], cf[2]};
      double d[3] = {(double)w[0] - c[0], (double)w[1] - c[1], (double)w[2] - c[2]};
      double residual = std::abs(dot3<double>(n, d));
      if (residual > max_dist) continue;
      out.points_last.push_back({c[0], c[1], c[2]});
      out.points_curr.push_back({(double)local[idx].x, (double)local[idx].y, (double)local[idx].z});
      out.normals_last.push_back({n[0], n[1], n[2]});
      out.residuals.push_back(residual);
      out.query_index.push_back((int)idx);
    }
    return out.size();
  }

  static bool is_collinear(const double* p1, const double* p2, const double* p3, double threshold) {  // :785-792
    double a[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, b[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    double na = sqnorm3<double>(a), nb = sqnorm3<double>(b);
    if (na > 0) { double s = std::sqrt(na); a[0] /= s; a[1] /= s; a[2] /= s; }   // Eigen normalized(): only if squaredNorm > 0
    if (nb > 0) { double s = std::sqrt(nb); b[0] /= s; b[1] /= s; b[2] /= s; }
    double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    return norm3<double>(c) < threshold;
  }

  // find_correspondences_kdtree (:647-767).  kd: tree over map_cloud (== VoxelMap::GetPointCloud()).
  static size_t find_correspondences_kdtree(const KdTree* kd, const std::vector<P3>& map_cloud, const P3* local, size_t m,
                                            const SE3f& pose, double max_dist, Correspondences& out,
                                            std::vector<int>* knn_dump = nullptr) {
    out.clear();
    if (!kd || !kd->built() || map_cloud.empty() || !local || m == 0) return 0;
    float T[16]; pose.Matrix(T);
    const int K = 5;
    if (knn_dump) knn_dump->assign(m * K, -1);
    for (size_t idx = 0; idx < m; ++idx) {
      float w[3];
      transform_point_4x4(T, local[idx].x, local[idx].y, local[idx].z, w);
      uint32_t nn[K]; float nd[K];
      int found = (int)kd->knnSearch(w, K, nn, nd);
      if (knn_dump) for (int k = 0; k < found; ++k) (*knn_dump)[idx * K + k] = (int)nn[k];
      if (found < 5) continue;
      double sel[K][3];
      for (int k = 0; k < K; ++k) { sel[k][0] = map_cloud[nn[k]].x; sel[k][1] = map_cloud[nn[k]].y; sel[k][2] = map_cloud[nn[k]].z; }
      if (is_collinear(sel[0], sel[1], sel[2], 0.5)) continue;
      double cen[3] = {0, 0, 0};
      for (int k = 0; k < K; ++k) { cen[0] += sel[k][0]; cen[1] += sel[k][1]; cen[2] += sel[k][2]; }
      cen[0] /= (double)K; cen[1] /= (double)K; cen[2] /= (double)K;
      double A[K * 3];
      for (int k = 0; k < K; ++k) for (int a = 0; a < 3; ++a) A[k * 3 + a] = sel[k][a] - cen[a];
      double nrm[3];
      smallest_right_singular_vec_nx3(A, K, nrm);
      double plane_d = -dot3<double>(nrm, cen);
      double wd[3] = {w[0], w[1], w[2]};
      double distance = std::abs(dot3<double>(nrm, wd) + plane_d);
      if (distance > max_dist) continue;
      out.points_last.push_back({cen[0], cen[1], cen[2]});
      out.points_curr.push_back({(double)local[idx].x, (double)local[idx].y, (double)local[idx].z});
      out.normals_last.push_back({nrm[0], nrm[1], nrm[2]});
      out.residuals.push_back(distance);
      out.query_index.push_back((int)idx);
    }
    return out.size();
  }

  // find_correspondences_loop (:465-585): query the CURRENT keyframe's points (at pose_curr) in the kd-tree of the matched
  // keyframe's cloud (already in world coordinates).  No distance gate.  points_last = the NEAREST neighbour taken back into the
  // matched keyframe's local frame through T_lw_last (f32 matrix, f64 product), not the plane centroid.
  // T_lw_last = Matrix4f::inverse() in the reference (a general 4x4 inverse from Eigen, not vendored): restated here as the rigid
  // inverse [R^T | -(R^T t)] evaluated in f32 - PARITY UNPINNED at this one step, the two agree to f32 rounding.
  static void rigid_inverse_f32(const SE3f& T, float* Ri, float* ti) {
    mat3_transpose<float>(T.R.m, Ri);
    float rt[3]; mat3_mul_vec<float>(Ri, T.t, rt);
    ti[0] = -rt[0]; ti[1] = -rt[1]; ti[2] = -rt[2];
  }
  static size_t find_correspondences_loop(const KdTree* kd, const std::vector<P3>& map_last_world, const SE3f& pose_last, const P3* local, size_t m,
                                          const SE3f& pose_curr, Correspondences& out) {
    out.clear();
    if (!kd || !kd->built() || map_last_world.empty() || !local || m == 0) return 0;
    float T[16]; pose_curr.Matrix(T);
    float Ri[9], ti[3];
    rigid_inverse_f32(pose_last, Ri, ti);
    const int K = 5;
    for (size_t idx = 0; idx < m; ++idx) {
      float w[3];
      transform_point_4x4(T, local[idx].x, local[idx].y, local[idx].z, w);
      uint32_t nn[K]; float nd[K];
      int found = (int)kd->knnSearch(w, K, nn, nd);
      if (found < 5) continue;
      double sel[K][3];
      for (int k = 0; k < K; ++k) { sel[k][0] = map_last_world[nn[k]].x; sel[k][1] = map_last_world[nn[k]].y; sel[k][2] = map_last_world[nn[k]].z; }
      if (is_collinear(sel[0], sel[1], sel[2], 0.5)) continue;
      double cen[3] = {0, 0, 0};
      for (int k = 0; k < K; ++k) { cen[0] += sel[k][0]; cen[1] += sel[k][1]; cen[2] += sel[k][2]; }
      cen[0] /= (double)K; cen[1] /= (double)K; cen[2] /= (double)K;
      double A[K * 3];
      for (int k = 0; k < K; ++k) for (int a = 0; a < 3; ++a) A[k * 3 + a] = sel[k][a] - cen[a];
      double nrm[3];
      smallest_right_singular_vec_nx3(A, K, nrm);
      double plane_d = -dot3<double>(nrm, cen);
      double wd[3] = {w[0], w[1], w[2]};
      double distance = std::abs(dot3<double>(nrm, wd) + plane_d);
      double Rid[9], pl[3];
      for (int a = 0; a < 9; ++a) Rid[a] = (double)Ri[a];
      mat3_mul_vec<double>(Rid, sel[0], pl);   // T_lw_last.block<3,3>.cast<double>() * pt_world + T_lw_last.block<3,1>.cast<double>()
      pl[0] += (double)ti[0]; pl[1] += (double)ti[1]; pl[2] += (double)ti[2];
      out.points_last.push_back({pl[0], pl[1], pl[2]});
      out.points_curr.push_back({(double)local[idx].x, (double)local[idx].y, (double)local[idx].z});
      out.normals_last.push_back({nrm[0], nrm[1], nrm[2]});
      out.residuals.push_back(distance);
      out.query_index.push_back((int)idx);
    }
    return out.size();
  }

  // optimize_loop (:40-251).  curr/matched: the keyframes' local feature clouds and world poses.  Returns success; on success
  // relative = curr_pose^-1 * optimised_curr_pose and the 1-NN (< 1 m) inlier ratio.
  bool optimize_loop(const P3* curr_local, size_t m_curr, const SE3f& curr_pose, const P3* matched_local, size_t m_matched, const SE3f& matched_pose,
                     SE3f& relative, float& inlier_ratio, int* iterations_out = nullptr) {
    last_stats = OptimizationStats();
    trace.clear();
    SE3f cur = curr_pose;
    if (ame) ame->reset();
    bool success = false;
    std::vector<P3> target(m_matched);   // transform_point_cloud (PointCloudUtils.cpp:102-125)
    {
      float Tm[16]; matched_pose.Matrix(Tm);
      for (size_t i = 0; i < m_matched; ++i) { float w[3]; transform_point_4x4(Tm, matched_local[i].x, matched_local[i].y, matched_local[i].z, w); target[i] = P3{w[0], w[1], w[2]}; }
    }
    KdTree kd; kd.setInputCloud(target);
    double scale = 1.0;
    std::string loss_type = "huber";
    if (ame) loss_type = ame->cfg.loss_type;
    int iters = 0;
    for (int it = 0; it < 100; ++it) {
      Correspondences corr;
      find_correspondences_loop(&kd, target, matched_pose, curr_local, m_curr, cur, corr);
      if (corr.size() < (size_t)cfg.min_correspondence_points) break;
      if (it == 0 && !corr.residuals.empty()) {  // :62-74
        std::vector<double> r = corr.residuals;
        std::sort(r.begin(), r.end());
        double mean = std::accumulate(r.begin(), r.end(), 0.0) / r.size();
        double var = 0.0;
        for (double v : r) var += (v - mean) * (v - mean);
        var /= r.size();
        scale = std::sqrt(var) / 6.0;
      }
      double adaptive_delta = cfg.robust_loss_delta;
      if (ame && ame->cfg.use_adaptive_m_estimator) {
        std::vector<double> nr;
        nr.reserve(corr.residuals.size());
        for (double r : corr.residuals) nr.push_back(r / std::max(scale, 1e-6));
        if (!nr.empty()) adaptive_delta = ame->calculate_scale_factor(nr);
      }
      const float* R = cur.R.m;
      const float* t = cur.t;
      const float* Rm = matched_pose.R.m;
      const float* tm = matched_pose.t;
      float H[36] = {0}, g[6] = {0};
      double H64[36] = {0}, g64[6] = {0};
      for (size_t i = 0; i < corr.size(); ++i) {  // :133-184
        float pm[3] = {(float)corr.points_last[i][0], (float)corr.points_last[i][1], (float)corr.points_last[i][2]};
        float p[3] = {(float)corr.points_curr[i][0], (float)corr.points_curr[i][1], (float)corr.points_curr[i][2]};
        float n[3] = {(float)corr.normals_last[i][0], (float)corr.normals_last[i][1], (float)corr.normals_last[i][2]};
        float Rq[3]; mat3_mul_vec<float>(Rm, pm, Rq);
        float q[3] = {Rq[0] + tm[0], Rq[1] + tm[1], Rq[2] + tm[2]};
        float Rp[3]; mat3_mul_vec<float>(R, p, Rp);
        float pw[3] = {Rp[0] + t[0], Rp[1] + t[1], Rp[2] + t[2]};
        float d[3] = {pw[0] - q[0], pw[1] - q[1], pw[2] - q[2]};
        float residual = dot3<float>(n, d);
        float normalized = (float)(corr.residuals[i] / std::max(scale, 1e-6));
        float J[6];
        for (int j = 0; j < 3; ++j) J[j] = sum3<float>(n[0] * R[0 * 3 + j], n[1] * R[1 * 3 + j], n[2] * R[2 * 3 + j]);
        float mn[3] = {-n[0], -n[1], -n[2]}, u[3];
        for (int j = 0; j < 3; ++j) u[j] = sum3<float>(mn[0] * R[0 * 3 + j], mn[1] * R[1 * 3 + j], mn[2] * R[2 * 3 + j]);
        const float ps[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
        for (int j = 0; j < 3; ++j) J[3 + j] = sum3<float>(u[0] * ps[0 * 3 + j], u[1] * ps[1 * 3 + j], u[2] * ps[2 * 3 + j]);
        float weight = 1.0f;
        if (cfg.use_robust_loss) {
          float an = std::abs(normalized);
          float delta = (float)adaptive_delta;
          if (loss_type == "cauchy") { float ratio = an / delta; weight = 1.0f / (1.0f + ratio * ratio); }
          else if (an > delta) weight = delta / an;
        }
        for (int a = 0; a < 6; ++a) {
          float wJ = weight * J[a];
          for (int b = 0; b < 6; ++b) { H[a * 6 + b] += wJ * J[b]; H64[a * 6 + b] += (double)wJ * (double)J[b]; }
        }
        float wr = weight * residual;
        for (int a = 0; a < 6; ++a) { g[a] += wr * J[a]; g64[a] += (double)wr * (double)J[a]; }
      }
      float mg[6], dx[6];
      for (int a = 0; a < 6; ++a) mg[a] = -g[a];
      ldlt6_solve(H, mg, dx);
      float dt[3] = {dx[0], dx[1], dx[2]}, dw[3] = {dx[3], dx[4], dx[5]};
      SE3f delta_T = (norm3<float>(dw) < 1e-10f) ? SE3f(SO3f::Identity(), dt) : SE3f(SO3f::Exp(dw), dt);
      IterTrace tr;
      if (keep_trace) {
        tr.n_corr = (int)corr.size(); tr.scale = scale; tr.delta = adaptive_delta;
        std::memcpy(tr.H, H, sizeof H); std::memcpy(tr.g, g, sizeof g);
        std::memcpy(tr.H64, H64, sizeof H64); std::memcpy(tr.g64, g64, sizeof g64);
        std::memcpy(tr.dx, dx, sizeof dx);
        cur.Matrix(tr.T_in);
        if (ame) { tr.em_iters = ame->last_em_iters; tr.kmeans_iters = ame->last_kmeans_iters; }
      }
      cur = cur * delta_T;
      if (keep_trace) { cur.Matrix(tr.T_out); trace.push_back(tr); }
      ++iters;
      last_stats.num_correspondences = corr.size();
      if (norm3<float>(dt) < cfg.translation_tolerance && norm3<float>(dw) < cfg.rotation_tolerance) {
        relative = curr_pose.Inverse() * cur;
        success = true;
        break;
      }
    }
    last_stats.num_iterations = iters;
    if (iterations_out) *iterations_out = iters;
    if (success) {  // :214-247
      int inl = 0, tot = 0;
      float Tc[16]; cur.Matrix(Tc);
      for (size_t i = 0; i < m_curr; ++i) {
        // Matrix().block<3,1>(0,3) + Matrix().block<3,3>(0,0) * p : translation FIRST, then the product
        float Rp[3]; mat3_mul_vec<float>(cur.R.m, &curr_local[i].x, Rp);
        float w[3] = {cur.t[0] + Rp[0], cur.t[1] + Rp[1], cur.t[2] + Rp[2]};
        uint32_t nn[1] = {0}; float nd[1] = {0.0f};
        kd.knnSearch(w, 1, nn, nd);
        if (std::sqrt(nd[0]) < 1.0f) inl++;
        tot++;
      }
      inlier_ratio = (float)inl / (float)tot;
      if (inlier_ratio < 0.5f) success = false;
    }
    last_stats.converged = success;
    return success;
  }

  // optimize (:255-463).  kd/map_cloud only used when !use_surfel_correspondence.
  bool optimize(const VoxelMap* map, const P3* local, size_t m, const SE3f& initial, SE3f& optimized,
                const KdTree* kd = nullptr, const std::vector<P3>* map_cloud = nullptr) {
    auto t0 = std::chrono::high_resolution_clock::now();
    last_stats = OptimizationStats();
    trace.clear();
    SE3f cur = initial;
    optimized = cur;
    if (ame) ame->reset();
    double total_initial_cost = 0, total_final_cost = 0;
    int total_iterations = 0;
    double scale = 1.0;
    for (int it = 0; it < cfg.max_iterations; ++it) {
      Correspondences corr;
      size_t nc;
      if (cfg.use_surfel_correspondence) nc = find_correspondences(map, local, m, cur, cfg.max_correspondence_distance, corr);
      else {
        std::vector<P3> cloud_copy;  // the reference re-materialises the L0 cloud every iteration (:679)
        if (map_cloud) cloud_copy = *map_cloud;
        nc = (!map || map->empty()) ? 0 : find_correspondences_kdtree(kd, cloud_copy, local, m, cur, cfg.max_correspondence_distance, corr);
      }
      if (nc < (size_t)cfg.min_correspondence_points) return false;
      if (it == 0 && !corr.residuals.empty()) {  // :304-316
        std::vector<double> r = corr.residuals;
        std::sort(r.begin(), r.end());
        double mean = std::accumulate(r.begin(), r.end(), 0.0) / r.size();
        double var = 0.0;
        for (double v : r) var += (v - mean) * (v - mean);
        var /= r.size();
        scale = std::sqrt(var) / 6.0;
      }
      double adaptive_delta = cfg.robust_loss_delta;  // :319-332
      if (ame && ame->cfg.use_adaptive_m_estimator) {
        std::vector<double> nr;
        nr.reserve(corr.residuals.size());
        for (double r : corr.residuals) nr.push_back(r / std::max(scale, 1e-6));
        if (!nr.empty()) adaptive_delta = ame->calculate_scale_factor(nr);
      }
      const float* R = cur.R.m;
      const float* t = cur.t;
      float H[36] = {0}, g[6] = {0}, total_cost = 0.0f;
      double H64[36] = {0}, g64[6] = {0}, cost64 = 0;
      std::string loss_type = "huber";
      if (ame) loss_type = ame->cfg.loss_type;
      for (size_t i = 0; i < corr.size(); ++i) {  // :359-410
        float p[3] = {(float)corr.points_curr[i][0], (float)corr.points_curr[i][1], (float)corr.points_curr[i][2]};
        float q[3] = {(float)corr.points_last[i][0], (float)corr.points_last[i][1], (float)corr.points_last[i][2]};
        float n[3] = {(float)corr.normals_last[i][0], (float)corr.normals_last[i][1], (float)corr.normals_last[i][2]};
        float Rp[3]; mat3_mul_vec<float>(R, p, Rp);
        float pw[3] = {Rp[0] + t[0], Rp[1] + t[1], Rp[2] + t[2]};
        float d[3] = {pw[0] - q[0], pw[1] - q[1], pw[2] - q[2]};
        float residual = dot3<float>(n, d);
        float normalized = (float)(corr.residuals[i] / std::max(scale, 1e-6));
        float J[6];
        for (int j = 0; j < 3; ++j) J[j] = sum3<float>(n[0] * R[0 * 3 + j], n[1] * R[1 * 3 + j], n[2] * R[2 * 3 + j]);  // n^T R
        float mn[3] = {-n[0], -n[1], -n[2]}, u[3];
        for (int j = 0; j < 3; ++j) u[j] = sum3<float>(mn[0] * R[0 * 3 + j], mn[1] * R[1 * 3 + j], mn[2] * R[2 * 3 + j]);  // (-n)^T R
        const float ps[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
        for (int j = 0; j < 3; ++j) J[3 + j] = sum3<float>(u[0] * ps[0 * 3 + j], u[1] * ps[1 * 3 + j], u[2] * ps[2 * 3 + j]);
        float weight = 1.0f;
        if (cfg.use_robust_loss) {
          float an = std::abs(normalized);
          float delta = (float)adaptive_delta;
          if (loss_type == "cauchy") { float ratio = an / delta; weight = 1.0f / (1.0f + ratio * ratio); }
          else if (an > delta) weight = delta / an;
        }
        for (int a = 0; a < 6; ++a) {
          float wJ = weight * J[a];
          for (int b = 0; b < 6; ++b) { H[a * 6 + b] += wJ * J[b]; H64[a * 6 + b] += (double)wJ * (double)J[b]; }
        }
        float wr = weight * residual;
        for (int a = 0; a < 6; ++a) { g[a] += wr * J[a]; g64[a] += (double)wr * (double)J[a]; }
        total_cost += wr * residual;
        cost64 += (double)wr * (double)residual;
      }
      if (it == 0) total_initial_cost = total_cost;
      total_final_cost = total_cost;
      float mg[6], dx[6];
      for (int a = 0; a < 6; ++a) mg[a] = -g[a];
      ldlt6_solve(H, mg, dx);  // :418
      float dt[3] = {dx[0], dx[1], dx[2]}, dw[3] = {dx[3], dx[4], dx[5]};
      SE3f delta_T = (norm3<float>(dw) < 1e-10f) ? SE3f(SO3f::Identity(), dt) : SE3f(SO3f::Exp(dw), dt);  // :426-431
      IterTrace tr;
      if (keep_trace) {
        tr.n_corr = (int)nc; tr.scale = scale; tr.delta = adaptive_delta;
        std::memcpy(tr.H, H, sizeof H); std::memcpy(tr.g, g, sizeof g); tr.cost = total_cost;
        std::memcpy(tr.H64, H64, sizeof H64); std::memcpy(tr.g64, g64, sizeof g64); tr.cost64 = cost64;
        std::memcpy(tr.dx, dx, sizeof dx);
        cur.Matrix(tr.T_in);
        if (ame) { tr.em_iters = ame->last_em_iters; tr.kmeans_iters = ame->last_kmeans_iters; }
      }
      cur = cur * delta_T;  // :434
      if (keep_trace) { cur.Matrix(tr.T_out); trace.push_back(tr); }
      float translation_delta = norm3<float>(dt), rotation_delta = norm3<float>(dw);
      total_iterations++;
      last_stats.num_correspondences = nc;
      bool converged = (translation_delta < cfg.translation_tolerance) && (rotation_delta < cfg.rotation_tolerance);
      if (converged) break;
    }
    optimized = cur;
    last_stats.num_iterations = total_iterations;
    last_stats.initial_cost = total_initial_cost;
    last_stats.final_cost = total_final_cost;
    last_stats.converged = true;
    auto t1 = std::chrono::high_resolution_clock::now();
    last_stats.optimization_time_ms = (double)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
    return true;
  }
};

}  // namespace orc
