// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over oracle/include/orc_*.hpp (see orc_capi.h).
#include "orc_capi.h"
#include <cstring>
#include "orc_pipeline.hpp"

using namespace orc;

static ICPConfig to_icp(const orc_icp_cfg* c) {
  ICPConfig o;
  o.max_iterations = c->max_iterations; o.translation_tolerance = c->translation_tolerance; o.rotation_tolerance = c->rotation_tolerance;
  o.max_correspondence_distance = c->max_correspondence_distance; o.min_correspondence_points = c->min_correspondence_points;
  o.use_robust_loss = c->use_robust_loss != 0; o.robust_loss_delta = c->robust_loss_delta;
  o.use_surfel_correspondence = c->use_surfel_correspondence != 0;
  return o;
}
static PkoConfig to_pko(const orc_icp_cfg* c) {
  PkoConfig o;
  o.use_adaptive_m_estimator = c->use_adaptive_m_estimator != 0;
  o.loss_type = c->loss_type == 1 ? "cauchy" : "huber";
  o.min_scale_factor = c->min_scale_factor; o.max_scale_factor = c->max_scale_factor;
  o.num_alpha_segments = c->num_alpha_segments; o.truncated_threshold = c->truncated_threshold;
  o.gmm_components = c->gmm_components; o.gmm_sample_size = c->gmm_sample_size;
  o.pko_kernel_type = c->pko_kernel_type == 1 ? "cauchy" : "huber";
  return o;
}
static void copy_trace(const std::vector<IterTrace>& tr, orc_iter_trace* out, int cap, int* n) {
  int k = 0;
  for (; k < (int)tr.size() && k < cap; ++k) {
    const IterTrace& s = tr[k]; orc_iter_trace& d = out[k];
    d.n_corr = s.n_corr; d.scale = s.scale; d.delta = s.delta;
    std::memcpy(d.H, s.H, sizeof d.H); std::memcpy(d.g, s.g, sizeof d.g); d.cost = s.cost;
    std::memcpy(d.H64, s.H64, sizeof d.H64); std::memcpy(d.g64, s.g64, sizeof d.g64); d.cost64 = s.cost64;
    std::memcpy(d.dx, s.dx, sizeof d.dx); std::memcpy(d.T_in, s.T_in, sizeof d.T_in); std::memcpy(d.T_out, s.T_out, sizeof d.T_out);
    d.em_iters = s.em_iters; d.kmeans_iters = s.kmeans_iters;
  }
  if (n) *n = k;
}

static SE3f raw_se3(const float* T) { SE3f p; for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) p.R.m[i * 3 + j] = T[i * 4 + j]; p.t[i] = T[i * 4 + 3]; } return p; }

extern "C" {

void orc_default_icp_cfg(orc_icp_cfg* c) {
  c->max_iterations = 4; c->translation_tolerance = 0.005; c->rotation_tolerance = 0.005; c->max_correspondence_distance = 1.0;
  c->min_correspondence_points = 10; c->use_robust_loss = 1; c->robust_loss_delta = 0.1; c->use_surfel_correspondence = 1;
  c->use_adaptive_m_estimator = 1; c->loss_type = 0; c->min_scale_factor = 0.1; c->max_scale_factor = 10.0; c->num_alpha_segments = 100;
  c->truncated_threshold = 10.0; c->gmm_components = 3; c->gmm_sample_size = 100; c->pko_kernel_type = 0;
}
void orc_default_pipe_cfg(orc_pipe_cfg* c, int mid360) {
  c->voxel_size = mid360 ? 0.4f : 0.5f; c->point_stride = mid360 ? 4 : 8; c->map_voxel_size = mid360 ? 0.4f : 0.5f; c->max_range = 100.0;
  c->surfel_planarity_threshold = 0.1f; c->keyframe_distance_threshold = 1.0; c->keyframe_rotation_threshold = 0.3;
  orc_default_icp_cfg(&c->icp);
  if (mid360) c->icp.use_surfel_correspondence = 0;
}

uint64_t orc_filter_morton_key(float x, float y, float z, float voxel) { return filter_morton_key(x, y, z, 1.0f / voxel); }
uint64_t orc_voxel_key_hash(int x, int y, int z) { return voxel_key_morton(VoxelKey{x, y, z}); }
void orc_point_to_key(const float* p, float voxel, int factor, int level, int* key) {
  VoxelMap m(voxel); m.SetHierarchyFactor(factor);
  VoxelKey k = m.PointToVoxelKey(p, level); key[0] = k.x; key[1] = k.y; key[2] = k.z;
}
void orc_parent_key(const int* key, int factor, int* parent) {
  VoxelMap m(0.5f); m.SetHierarchyFactor(factor);
  VoxelKey k = m.GetParentKey(VoxelKey{key[0], key[1], key[2]}); parent[0] = k.x; parent[1] = k.y; parent[2] = k.z;
}

void orc_filter(const float* xyz, size_t n, int stride, float voxel, float* out_xyz, uint64_t* out_keys, size_t* m) {
  FastVoxelFilter f(voxel);
  std::vector<P3> out; std::vector<uint64_t> keys;
  f.filter(reinterpret_cast<const P3*>(xyz), n, out, stride, &keys);
  if (out_xyz) std::memcpy(out_xyz, out.data(), out.size() * sizeof(P3));
  if (out_keys) std::memcpy(out_keys, keys.data(), keys.size() * sizeof(uint64_t));
  *m = out.size();
}

void* orc_map_create(float voxel, int factor, float planarity, int compute_surfels) {
  VoxelMap* m = new VoxelMap(voxel);
  m->SetHierarchyFactor(factor); m->SetPlanarityThreshold(planarity); m->SetComputeSurfels(compute_surfels != 0);
  return m;
}
void orc_map_destroy(void* h) { delete static_cast<VoxelMap*>(h); }
void orc_map_clear(void* h) { static_cast<VoxelMap*>(h)->Clear(); }
void orc_map_update(void* h, const float* xyz, size_t n, const double* sensor, double max_distance) {
  static_cast<VoxelMap*>(h)->UpdateVoxelMap(reinterpret_cast<const P3*>(xyz), n, sensor, max_distance, true);
}
void orc_map_counts(void* h, size_t* l0, size_t* l1, size_t* surfels) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  if (l0) *l0 = m->GetVoxelCount();
  if (l1) *l1 = m->GetL1VoxelCount();
  if (surfels) *surfels = m->GetSurfelCount();
}
void orc_map_export_l0(void* h, int* keys, float* cent, int* counts) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  size_t i = 0;
  for (const auto& kv : m->l0().values) {
    if (keys) { keys[i * 3] = kv.first.x; keys[i * 3 + 1] = kv.first.y; keys[i * 3 + 2] = kv.first.z; }
    if (cent) { cent[i * 3] = kv.second.c[0]; cent[i * 3 + 1] = kv.second.c[1]; cent[i * 3 + 2] = kv.second.c[2]; }
    if (counts) counts[i] = kv.second.point_count;
    ++i;
  }
}
void orc_map_export_l1(void* h, int* keys, int* nchild, int* children, int* has_surfel, float* normal, float* centroid,
                       float* planarity, int* last_child_count) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  size_t i = 0;
  for (const auto& kv : m->l1().values) {
    const auto& nd = kv.second;
    if (keys) { keys[i * 3] = kv.first.x; keys[i * 3 + 1] = kv.first.y; keys[i * 3 + 2] = kv.first.z; }
    if (nchild) nchild[i] = (int)nd.children.size();
    if (children) for (size_t c = 0; c < nd.children.size() && c < 27; ++c) {
      children[(i * 27 + c) * 3] = nd.children.at(c).x; children[(i * 27 + c) * 3 + 1] = nd.children.at(c).y; children[(i * 27 + c) * 3 + 2] = nd.children.at(c).z;
    }
    if (has_surfel) has_surfel[i] = nd.has_surfel ? 1 : 0;
    if (normal) for (int a = 0; a < 3; ++a) normal[i * 3 + a] = nd.normal[a];
    if (centroid) for (int a = 0; a < 3; ++a) centroid[i * 3 + a] = nd.centroid[a];
    if (planarity) planarity[i] = nd.planarity;
    if (last_child_count) last_child_count[i] = nd.last_child_count;
    ++i;
  }
}
int orc_map_lookup(void* h, const float* p, float* n, float* c) { return static_cast<VoxelMap*>(h)->GetSurfelAtPoint(p, n, c) ? 1 : 0; }
void orc_map_transform_rehash(void* h, const float* T16) { static_cast<VoxelMap*>(h)->ApplyTransformAndRehash(T16); }

size_t orc_icp_correspondences(void* map, const float* local_xyz, size_t m, const float* T16, double max_dist, int* state, int* l1key,
                               uint64_t* morton, float* normal, float* centroid, double* residual, float* world) {
  VoxelMap* vm = static_cast<VoxelMap*>(map);
  size_t acc = 0;
  for (size_t i = 0; i < m; ++i) {
    float w[3];
    transform_point_4x4(T16, local_xyz[i * 3], local_xyz[i * 3 + 1], local_xyz[i * 3 + 2], w);
    if (world) { world[i * 3] = w[0]; world[i * 3 + 1] = w[1]; world[i * 3 + 2] = w[2]; }
    VoxelKey k = vm->PointToVoxelKey(w, 1);
    if (l1key) { l1key[i * 3] = k.x; l1key[i * 3 + 1] = k.y; l1key[i * 3 + 2] = k.z; }
    if (morton) morton[i] = voxel_key_morton(k);
    float nf[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
    int st = 0; double res = 0;
    if (vm->GetSurfelAtPoint(w, nf, cf)) {
      double n[3] = {nf[0], nf[1], nf[2]};
      double d[3] = {(double)w[0] - (double)cf[0], (double)w[1] - (double)cf[1], (double)w[2] - (double)cf[2]};
      res = std::abs(dot3<double>(n, d));
      st = (res > max_dist) ? 1 : 2;
      if (st == 2) acc++;
    }
    if (state) state[i] = st;
    if (normal) for (int a = 0; a < 3; ++a) normal[i * 3 + a] = nf[a];
    if (centroid) for (int a = 0; a < 3; ++a) centroid[i * 3 + a] = cf[a];
    if (residual) residual[i] = res;
  }
  return acc;
}

int orc_icp_optimize(void* map, const float* local_xyz, size_t m, const float* T_init16, const orc_icp_cfg* cfg, float* T_out16,
                     orc_iter_trace* trace, int trace_cap, int* n_trace) {
  auto ame = std::make_shared<AdaptiveMEstimator>(to_pko(cfg));
  ICPOptimizer icp(to_icp(cfg), ame);
  icp.keep_trace = trace != nullptr;
  SE3f init = raw_se3(T_init16), out;  // optimize() takes an SE3f whose rotation is used verbatim (ICP.cpp:265)
  bool ok = icp.optimize(static_cast<VoxelMap*>(map), reinterpret_cast<const P3*>(local_xyz), m, init, out);
  out.Matrix(T_out16);
  if (trace) copy_trace(icp.trace, trace, trace_cap, n_trace);
  return ok ? 1 : 0;
}

int orc_icp_optimize_kdtree(const float* map_xyz, size_t nmap, const float* local_xyz, size_t m, const float* T_init16,
                            const orc_icp_cfg* cfg, float* T_out16, orc_iter_trace* trace, int trace_cap, int* n_trace) {
  std::vector<P3> cloud(nmap);
  std::memcpy(cloud.data(), map_xyz, nmap * sizeof(P3));
  KdTree kd; kd.setInputCloud(cloud);
  // a non-empty VoxelMap stand-in is only needed for the `empty()` guard (:653)
  VoxelMap vm(0.5f);
  if (nmap) { double s[3] = {0, 0, 0}; vm.SetComputeSurfels(false); vm.UpdateVoxelMap(cloud.data(), 1, s, 1e30, true); }
  auto ame = std::make_shared<AdaptiveMEstimator>(to_pko(cfg));
  ICPConfig ic = to_icp(cfg); ic.use_surfel_correspondence = false;
  ICPOptimizer icp(ic, ame);
  icp.keep_trace = trace != nullptr;
  SE3f init = raw_se3(T_init16), out;
  bool ok = icp.optimize(&vm, reinterpret_cast<const P3*>(local_xyz), m, init, out, &kd, &cloud);
  out.Matrix(T_out16);
  if (trace) copy_trace(icp.trace, trace, trace_cap, n_trace);
  return ok ? 1 : 0;
}

// optimize_loop (ICP.cpp:40-251): loop-closure ICP of the current keyframe against a matched keyframe (local feature clouds + world poses)
int orc_icp_optimize_loop(const float* curr_xyz, size_t m_curr, const float* T_curr16, const float* matched_xyz, size_t m_matched,
                          const float* T_matched16, const orc_icp_cfg* cfg, float* T_rel16, float* inlier_ratio, int* iterations,
                          orc_iter_trace* trace, int trace_cap, int* n_trace) {
  auto ame = std::make_shared<AdaptiveMEstimator>(to_pko(cfg));
  ICPOptimizer icp(to_icp(cfg), ame);
  icp.keep_trace = trace != nullptr;
  SE3f rel;
  float ratio = 0.0f;
  int iters = 0;
  bool ok = icp.optimize_loop(reinterpret_cast<const P3*>(curr_xyz), m_curr, raw_se3(T_curr16), reinterpret_cast<const P3*>(matched_xyz), m_matched,
                              raw_se3(T_matched16), rel, ratio, &iters);
  if (T_rel16) rel.Matrix(T_rel16);
  if (inlier_ratio) *inlier_ratio = ratio;
  if (iterations) *iterations = iters;
  if (trace) copy_trace(icp.trace, trace, trace_cap, n_trace);
  return ok ? 1 : 0;
}

size_t orc_kdtree_correspondences(const float* map_xyz, size_t nmap, const float* local_xyz, size_t m, const float* T16, double max_dist,
                                  int* knn, int* state, float* normal, float* centroid, double* residual) {
  std::vector<P3> cloud(nmap);
  std::memcpy(cloud.data(), map_xyz, nmap * sizeof(P3));
  KdTree kd; kd.setInputCloud(cloud);
  SE3f pose; // use the matrix verbatim (no re-projection): teacher-forced pose
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) pose.R.m[i * 3 + j] = T16[i * 4 + j]; pose.t[i] = T16[i * 4 + 3]; }
  Correspondences corr; std::vector<int> dump;
  size_t nc = ICPOptimizer::find_correspondences_kdtree(&kd, cloud, reinterpret_cast<const P3*>(local_xyz), m, pose, max_dist, corr, &dump);
  if (knn) std::memcpy(knn, dump.data(), dump.size() * sizeof(int));
  if (state) for (size_t i = 0; i < m; ++i) state[i] = 0;
  for (size_t c = 0; c < corr.size(); ++c) {
    int q = corr.query_index[c];
    if (state) state[q] = 2;
    if (normal) for (int a = 0; a < 3; ++a) normal[q * 3 + a] = (float)corr.normals_last[c][a];
    if (centroid) for (int a = 0; a < 3; ++a) centroid[q * 3 + a] = (float)corr.points_last[c][a];
    if (residual) residual[q] = corr.residuals[c];
  }
  return nc;
}

void orc_knn(const float* map_xyz, size_t nmap, const float* q_xyz, size_t m, int k, int* idx, float* d2, int* found) {
  std::vector<P3> cloud(nmap);
  std::memcpy(cloud.data(), map_xyz, nmap * sizeof(P3));
  KdTree kd; kd.setInputCloud(cloud);
  std::vector<uint32_t> ii(k); std::vector<float> dd(k);
  for (size_t i = 0; i < m; ++i) {
    size_t f = kd.knnSearch(q_xyz + i * 3, (size_t)k, ii.data(), dd.data());
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < (int)f ? (int)ii[j] : -1; d2[i * k + j] = j < (int)f ? dd[j] : 0.0f; }
    if (found) found[i] = (int)f;
  }
}

double orc_pko_scale(const double* residuals, size_t n, const orc_icp_cfg* cfg, double* sample, int* n_sample, double* means, double* vars,
                     double* weights, int* em_iters, double* js) {
  AdaptiveMEstimator a(to_pko(cfg));
  std::vector<double> r(residuals, residuals + n);
  double alpha = a.calculate_scale_factor(r);
  if (sample) std::memcpy(sample, a.last_sample.data(), a.last_sample.size() * sizeof(double));
  if (n_sample) *n_sample = (int)a.last_sample.size();
  for (size_t j = 0; j < a.gmm_means.size(); ++j) { if (means) means[j] = a.gmm_means[j]; if (vars) vars[j] = a.gmm_variances[j]; if (weights) weights[j] = a.gmm_weights[j]; }
  if (em_iters) *em_iters = a.last_em_iters;
  if (js) std::memcpy(js, a.last_js.data(), a.last_js.size() * sizeof(double));
  return alpha;
}
void orc_shuffle_head(int n, int head, int* out) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::mt19937 g(42);
  std::shuffle(idx.begin(), idx.end(), g);
  for (int i = 0; i < head && i < n; ++i) out[i] = idx[i];
}

void orc_svd3f(const float* A, float* U, float* S, float* V) { jacobi_svd3<float>(A, U, S, V); }
void orc_so3_normalize(const float* R, float* out) { SO3f r = SO3f::FromMatrix(R); std::memcpy(out, r.m, sizeof r.m); }
void orc_so3_exp(const float* w, float* out) { SO3f r = SO3f::Exp(w); std::memcpy(out, r.m, sizeof r.m); }
void orc_ldlt6_solve(const float* H, const float* b, float* x) { ldlt6_solve(H, b, x); }
void orc_se3_mul(const float* A16, const float* B16, float* C16) { (raw_se3(A16) * raw_se3(B16)).Matrix(C16); }
void orc_se3_inv(const float* A16, float* C16) { raw_se3(A16).Inverse().Matrix(C16); }
void orc_fit_plane(const float* cents, int n, float* mu, float* normal, float* planarity) {
  std::vector<std::array<float, 3>> c(n);
  for (int i = 0; i < n; ++i) c[i] = {cents[i * 3], cents[i * 3 + 1], cents[i * 3 + 2]};
  VoxelMap::FitPlane(c, mu, normal, *planarity);
}

void* orc_pipe_create(const orc_pipe_cfg* c) {
  PipelineConfig p;
  p.voxel_size = c->voxel_size; p.point_stride = c->point_stride; p.map_voxel_size = c->map_voxel_size; p.max_range = c->max_range;
  p.surfel_planarity_threshold = c->surfel_planarity_threshold; p.keyframe_distance_threshold = c->keyframe_distance_threshold;
  p.keyframe_rotation_threshold = c->keyframe_rotation_threshold; p.icp = to_icp(&c->icp); p.pko = to_pko(&c->icp);
  return new Pipeline(p);
}
void orc_pipe_destroy(void* h) { delete static_cast<Pipeline*>(h); }
int orc_pipe_process(void* h, const float* xyz, size_t n, size_t stride_floats, float* pose16, int* flags, double* times_ms,
                     int* n_features, int* n_corr, int* n_iters) {
  Pipeline* p = static_cast<Pipeline*>(h);
  std::vector<P3> raw(n);
  for (size_t i = 0; i < n; ++i) raw[i] = P3{xyz[i * stride_floats], xyz[i * stride_floats + 1], xyz[i * stride_floats + 2]};
  bool ok = p->process_scan(raw.data(), n);
  if (pose16) p->pose.Matrix(pose16);
  if (flags) *flags = (p->last_was_keyframe ? 1 : 0) | (p->last_icp_ok ? 2 : 0);
  if (times_ms) { times_ms[0] = p->last_times.preprocess_ms; times_ms[1] = p->last_times.icp_ms; times_ms[2] = p->last_times.map_update_ms; times_ms[3] = p->last_times.total_ms; }
  if (n_features) *n_features = (int)p->feature_cloud.size();
  if (n_corr) *n_corr = (int)p->icp.last_stats.num_correspondences;
  if (n_iters) *n_iters = (int)p->icp.last_stats.num_iterations;
  return ok ? 1 : 0;
}
void* orc_pipe_map(void* h) { return &static_cast<Pipeline*>(h)->map; }
size_t orc_pipe_features(void* h, float* xyz, size_t cap) {
  Pipeline* p = static_cast<Pipeline*>(h);
  size_t k = std::min(cap, p->feature_cloud.size());
  if (xyz) std::memcpy(xyz, p->feature_cloud.data(), k * sizeof(P3));
  return p->feature_cloud.size();
}

// ops: (op, key) pairs; op 0 = operator[] (insert if absent), 1 = erase.  Iteration order of the restated DenseMap
// (pinned against the real ankerl::unordered_dense in tests/test_oracle_pins.py)
size_t orc_dense_order(const int64_t* ops, size_t n_ops, uint64_t* out_keys) {
  struct H { uint64_t operator()(uint64_t k) const { return k; } };
  orc::DenseMap<uint64_t, int, H> m;
  for (size_t i = 0; i < n_ops; ++i) {
    uint64_t k = (uint64_t)ops[i * 2 + 1];
    if (ops[i * 2] == 0) m[k] += 1; else m.erase(k);
  }
  size_t j = 0;
  for (const auto& kv : m.values) out_keys[j++] = kv.first;
  return j;
}

}  // extern "C"

// ---- scan loaders on file images (orc_ingest.hpp) ---------------------------------------------------------
#include "orc_ingest.hpp"
static int emit_cloud(const std::vector<float>& v, float* out_xyz, size_t cap, size_t* n) {
  *n = v.size() / 3;
  if (*n > cap) return -1;
  if (!v.empty()) std::memcpy(out_xyz, v.data(), v.size() * sizeof(float));
  return 0;
}
extern "C" int orc_ply_load(const void* image, size_t len, float* out_xyz, size_t cap, size_t* n) {
  return emit_cloud(orc::ply_load(std::string(static_cast<const char*>(image), len)), out_xyz, cap, n);
}
extern "C" int orc_kitti_load(const void* image, size_t len, float* out_xyz, size_t cap, size_t* n) {
  return emit_cloud(orc::kitti_load(std::string(static_cast<const char*>(image), len)), out_xyz, cap, n);
}

// ---- util::VoxelGrid (orc_voxelgrid.hpp) ------------------------------------------------------------------
#include "orc_voxelgrid.hpp"
extern "C" int orc_voxel_grid_filter(const float* xyz, size_t n, float leaf, float* out_xyz, size_t cap, size_t* m) {
  return emit_cloud(orc::voxel_grid_filter(xyz, n, leaf), out_xyz, cap, m);
}
