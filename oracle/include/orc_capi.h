/* ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement (oracle/include/orc_*.hpp)
 * for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference). */
#ifndef ORC_CAPI_H
#define ORC_CAPI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int max_iterations;
  double translation_tolerance, rotation_tolerance, max_correspondence_distance;
  int min_correspondence_points;
  int use_robust_loss;
  double robust_loss_delta;
  int use_surfel_correspondence;
  /* PKO (AdaptiveMEstimator) */
  int use_adaptive_m_estimator;
  int loss_type;        /* 0 huber, 1 cauchy */
  double min_scale_factor, max_scale_factor;
  int num_alpha_segments;
  double truncated_threshold;
  int gmm_components, gmm_sample_size;
  int pko_kernel_type;  /* 0 huber, 1 cauchy */
} orc_icp_cfg;

typedef struct {
  int n_corr;
  double scale, delta;
  float H[36], g[6], cost;
  double H64[36], g64[6], cost64;
  float dx[6];
  float T_in[16], T_out[16];
  int em_iters, kmeans_iters;
} orc_iter_trace;

typedef struct {
  float voxel_size; int point_stride; float map_voxel_size; double max_range;
  float surfel_planarity_threshold; double keyframe_distance_threshold, keyframe_rotation_threshold;
  orc_icp_cfg icp;
} orc_pipe_cfg;

void orc_default_icp_cfg(orc_icp_cfg* c);
void orc_default_pipe_cfg(orc_pipe_cfg* c, int mid360);

/* keys */
uint64_t orc_filter_morton_key(float x, float y, float z, float voxel);
uint64_t orc_voxel_key_hash(int x, int y, int z);
void orc_point_to_key(const float* p, float voxel, int factor, int level, int* key);
void orc_parent_key(const int* key, int factor, int* parent);

/* FastVoxelFilter */
void orc_filter(const float* xyz, size_t n, int stride, float voxel, float* out_xyz, uint64_t* out_keys, size_t* m);

/* scan loaders on file images (orc_ingest.hpp): out_xyz capacity in points; return 0 ok / -1 buffer too small */
int orc_ply_load(const void* image, size_t len, float* out_xyz, size_t cap, size_t* n);
int orc_kitti_load(const void* image, size_t len, float* out_xyz, size_t cap, size_t* n);

/* util::VoxelGrid::filter (orc_voxelgrid.hpp) */
int orc_voxel_grid_filter(const float* xyz, size_t n, float leaf, float* out_xyz, size_t cap, size_t* m);

/* VoxelMap */
void* orc_map_create(float voxel, int factor, float planarity, int compute_surfels);
void orc_map_destroy(void* h);
void orc_map_clear(void* h);
void orc_map_update(void* h, const float* xyz, size_t n, const double* sensor, double max_distance);
void orc_map_counts(void* h, size_t* l0, size_t* l1, size_t* surfels);
void orc_map_export_l0(void* h, int* keys, float* cent, int* counts);
void orc_map_export_l1(void* h, int* keys, int* nchild, int* children /*27*3 per L1*/, int* has_surfel, float* normal,
                       float* centroid, float* planarity, int* last_child_count);
int orc_map_lookup(void* h, const float* p, float* n, float* c);
void orc_map_transform_rehash(void* h, const float* T16);

/* ICP */
/* per-query outputs: state 0 = no surfel, 1 = surfel found but gated out, 2 = accepted */
size_t orc_icp_correspondences(void* map, const float* local_xyz, size_t m, const float* T16, double max_dist, int* state,
                               int* l1key, uint64_t* morton, float* normal, float* centroid, double* residual, float* world);
int orc_icp_optimize(void* map, const float* local_xyz, size_t m, const float* T_init16, const orc_icp_cfg* cfg, float* T_out16,
                     orc_iter_trace* trace, int trace_cap, int* n_trace);
/* KDTree mode: map_xyz = VoxelMap::GetPointCloud() order */
int orc_icp_optimize_kdtree(const float* map_xyz, size_t nmap, const float* local_xyz, size_t m, const float* T_init16,
                            const orc_icp_cfg* cfg, float* T_out16, orc_iter_trace* trace, int trace_cap, int* n_trace);
size_t orc_kdtree_correspondences(const float* map_xyz, size_t nmap, const float* local_xyz, size_t m, const float* T16, double max_dist,
                                  int* knn /*m*5*/, int* state, float* normal /*f32 cast*/, float* centroid, double* residual);
void orc_knn(const float* map_xyz, size_t nmap, const float* q_xyz, size_t m, int k, int* idx, float* d2, int* found);

/* PKO */
/* optimize_loop (ICP.cpp:40-251): returns 1 on success (converged within 100 iterations and inlier ratio >= 0.5) */
int orc_icp_optimize_loop(const float* curr_xyz, size_t m_curr, const float* T_curr16, const float* matched_xyz, size_t m_matched,
                          const float* T_matched16, const orc_icp_cfg* cfg, float* T_rel16, float* inlier_ratio, int* iterations,
                          orc_iter_trace* trace, int trace_cap, int* n_trace);
double orc_pko_scale(const double* residuals, size_t n, const orc_icp_cfg* cfg, double* sample /*<=gmm_sample_size*/, int* n_sample,
                     double* means, double* vars, double* weights, int* em_iters, double* js /*num_alpha_segments+1*/);
void orc_shuffle_head(int n, int head, int* out);  /* first `head` entries of std::shuffle(iota(n), mt19937(42)) */

/* numerics */
void orc_svd3f(const float* A, float* U, float* S, float* V);
void orc_so3_normalize(const float* R, float* out);
void orc_so3_exp(const float* w, float* out);
void orc_ldlt6_solve(const float* H, const float* b, float* x);
void orc_se3_mul(const float* A16, const float* B16, float* C16);
void orc_se3_inv(const float* A16, float* C16);
void orc_fit_plane(const float* cents, int n, float* mu, float* normal, float* planarity);

/* container semantics probe: iteration order of the restated dense map under (op, key) pairs, op 0 = operator[], 1 = erase */
size_t orc_dense_order(const int64_t* ops, size_t n_ops, uint64_t* out_keys);

/* pipeline (Estimator-lite) */
void* orc_pipe_create(const orc_pipe_cfg* cfg);
void orc_pipe_destroy(void* h);
/* returns 1 if processed; flags bit0 = keyframe created, bit1 = ICP ok; times_ms[4] = preprocess, icp, map_update, total */
int orc_pipe_process(void* h, const float* xyz, size_t n, size_t stride_floats, float* pose16, int* flags, double* times_ms,
                     int* n_features, int* n_corr, int* n_iters);
void* orc_pipe_map(void* h);
size_t orc_pipe_features(void* h, float* xyz, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
