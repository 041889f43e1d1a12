/* b2lo.h — C ABI of the B200-native scan-to-map registration engine.
 *
 * This is the drop-in boundary underneath three C++ classes of SiarheiHerasiuta/lidar_odometry
 * (the reference has no FFI layer of its own; every entry point cites the class member it replaces):
 *
 *   lidar_slam::map::FastVoxelFilter                      src/database/VoxelMap.h:53-143
 *   lidar_slam::map::VoxelMap                             src/database/VoxelMap.h:188-332, VoxelMap.cpp
 *   lidar_slam::optimization::IterativeClosestPointOptimizer::optimize
 *                                                         src/optimization/IterativeClosestPointOptimizer.h:158-225, .cpp:255-463
 *   lidar_slam::optimization::AdaptiveMEstimator (PKO)    src/optimization/AdaptiveMEstimator.h:58-143  (runs on the device inside optimize)
 *
 * Conventions: plain pointers and sizes, POD structs, no C++/torch types.  Host pointers unless the
 * name ends in _dev.  Point clouds are AoS float32 xyz with a caller-given stride in floats (3 for
 * util::Point3D, 4 for KITTI .bin xyzI).  Poses are row-major 4x4 float32 (util::SE3::Matrix()).
 * Every function returns an int: 0 = ok, < 0 = error (B2LO_E_*), > 0 = soft outcome (B2LO_S_*).
 * No exception crosses this boundary.  All work of a context is issued on its own CUDA stream; a
 * context and the maps created from it must be used from one thread at a time, except the const
 * readers b2lo_map_counts / b2lo_map_export_surfels which take the map's internal mutex
 * (VoxelMap::m_mutex, VoxelMap.h:331).  There is NO CPU fallback: without a CUDA device every
 * call fails with B2LO_E_CUDA.
 */
#ifndef B2LO_H
#define B2LO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2LO_OK 0
#define B2LO_S_INSUFFICIENT 1 /* optimize(): correspondences < min_correspondence_points (ICP.cpp:298-302) -> reference returns false */
#define B2LO_S_EMPTY 2        /* empty input / empty map: nothing done */
#define B2LO_E_CUDA (-1)
#define B2LO_E_ARG (-2)
#define B2LO_E_RANGE (-3)     /* a voxel key left the 21-bit-per-axis Z-order domain (|coord| >= 2^20 voxels) */
#define B2LO_E_CAPACITY (-4)
#define B2LO_E_NOMEM (-5)

typedef struct b2lo_ctx b2lo_ctx;
typedef struct b2lo_map b2lo_map;
typedef struct b2lo_odom b2lo_odom;

/* optimization::ICPConfig (ICP.h:55-76) + optimization::AdaptiveMEstimatorConfig (AdaptiveMEstimator.h:28-45),
 * with the values Estimator.cpp:49-70 wires in as defaults (b2lo_default_icp_cfg). */
typedef struct {
  int max_iterations;            /* >= 1, no upper bound (ICPConfig's default is 50); b2lo_icp_stats::it traces the first B2LO_MAX_ITERS */
  double translation_tolerance, rotation_tolerance;
  double max_correspondence_distance;
  int min_correspondence_points;
  int use_robust_loss;
  double robust_loss_delta;
  int use_surfel_correspondence; /* 1: O(1) surfel lookup, 0: exact 5-NN + plane fit ("KDTree" mode) */
  int use_adaptive_m_estimator;
  int loss_type;                 /* 0 huber, 1 cauchy (GN weight, ICP.cpp:394-403) */
  double min_scale_factor, max_scale_factor;
  int num_alpha_segments;        /* <= 128 */
  double truncated_threshold;
  int gmm_components;            /* == 3 */
  int gmm_sample_size;           /* 1..128 */
  int pko_kernel_type;           /* 0 huber, 1 cauchy */
} b2lo_icp_cfg;

/* OptimizationStats (ICP.h:203-210) + per-iteration taps used by the parity tests. */
#define B2LO_MAX_ITERS 16
typedef struct {
  int n_corr;
  double scale, delta;
  double H[36], g[6], cost; /* f64 block-tree accumulation of the f32 terms (row-major 6x6; full symmetric) */
  float dx[6];
  float T_in[16], T_out[16];
  int em_iters, kmeans_iters;
} b2lo_iter_trace;
typedef struct {
  int status;               /* B2LO_OK or B2LO_S_INSUFFICIENT */
  int num_iterations;
  int num_correspondences;  /* of the last iteration */
  int converged;            /* tolerance met (the reference reports true whenever it returns true) */
  double initial_cost, final_cost;
  float device_ms;          /* CUDA-event time of the optimize call on the context stream */
  b2lo_iter_trace it[B2LO_MAX_ITERS];
} b2lo_icp_stats;

void b2lo_default_icp_cfg(b2lo_icp_cfg* cfg);
const char* b2lo_version(void);
const char* b2lo_last_error(void);
/* Opt-in for the throughput mode (many independent sequences on one GPU, b2lo_odom_process_batch_dev): their streams overlap only if
 * they map to different hardware work queues, and the driver's default is 8.  Sets CUDA_DEVICE_MAX_CONNECTIONS=<connections> (1..32)
 * for THIS process unless the variable is already set; it takes effect only when called before the process's first CUDA call.  The
 * library never touches the environment on its own. */
int b2lo_process_env_for_batches(int connections);
void b2lo_struct_sizes(size_t out[5]); /* sizeof icp_cfg, iter_trace, icp_stats, odom_cfg, odom_result: lets a binding verify its mirrors */

/* ---- context ------------------------------------------------------------------------------------ */
int b2lo_ctx_create(int device, b2lo_ctx** out);
int b2lo_ctx_destroy(b2lo_ctx* ctx);
int b2lo_ctx_sync(b2lo_ctx* ctx);
void* b2lo_ctx_stream(b2lo_ctx* ctx);          /* cudaStream_t */
long long b2lo_ctx_launch_count(b2lo_ctx* ctx); /* kernels launched by this library on ctx since creation */
int b2lo_ctx_io_bytes(b2lo_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h); /* PCIe bytes moved by this library on ctx */

/* optional per-kernel CUDA-event timing on the context stream (adds event records; keep it off when timing whole scans).
 * slots: 0 K1 downsample (5 kernels), 1 K2 surfel correspondence, 2 K4 PKO fit, 3 K4 PKO arg-min, 4 K5 normal equations + solve,
 * 5 K6 map update (all kernels but the cull scan), 6 cloud transform, 7 K3 kNN + plane fit (3 kernels), 8 K6 radius-cull scan */
int b2lo_ctx_host_us(b2lo_ctx* ctx, double out[8], int reset); /* host wall-clock split of b2lo_odom_process: 0 gather+H2D enqueue, 1 K1+ICP enqueue, 2 wait for the pose, 3 host pose algebra, 4 K6 enqueue + wait (debug aid) */
int b2lo_ctx_debug_clocks(b2lo_ctx* ctx, long long out[32]); /* SM-clock stamps of the single-CTA phases of the last optimize (debug aid) */
int b2lo_ctx_profile(b2lo_ctx* ctx, int enable);
int b2lo_ctx_profile_read(b2lo_ctx* ctx, int slot, double* total_ms, long long* launches);

/* ---- FastVoxelFilter::filter (VoxelMap.h:73-104) ------------------------------------------------ */
/* out_xyz: capacity >= ceil(n/stride) points (packed xyz); *m receives the voxel count (getVoxelCount). */
int b2lo_filter(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, int stride, float voxel_size, float* out_xyz,
                uint64_t* out_keys /* nullable: Morton key of every output voxel */, size_t* m);
/* device-resident variant: raw scan already in HBM, result stays in the context's feature buffer */
int b2lo_filter_dev(b2lo_ctx* ctx, const float* xyz_dev, size_t n, size_t stride_floats, int stride, float voxel_size);
int b2lo_ctx_features(b2lo_ctx* ctx, float* out_xyz, size_t cap, size_t* m); /* copy the feature buffer to the host */

/* ---- scan ingest: K1 reads the dataset's own records in place ------------------------------
 * A scan file image is a stream of fixed-size records holding three IEEE f32 coordinates at fixed byte offsets:
 *   KITTI .bin (load_kitti_binary, src/util/PointCloudUtils.cpp:19-65): 16-byte records x,y,z,intensity -> {16, 0, 4, 8};
 *   binary PLY (PLYPlayer::load_ply_point_cloud, app/player/ply_player.cpp:267-343): the vertex record described by the header,
 *   x/y/z copied bytewise from their property offsets whatever their declared type, no byte swap (as the reference does).
 * The device reads only every stride-th record (one or two 32 B sectors each); no host repack into util::PointCloud, no staging copy
 * for page-locked or device-resident file images.  Offsets need no alignment (e.g. 15-byte xyz+rgb vertices). */
typedef struct { uint32_t record_bytes, off_x, off_y, off_z; } b2lo_record_fmt;
void b2lo_kitti_record_fmt(b2lo_record_fmt* fmt);
/* PLYPlayer::parse_ply_header (ply_player.cpp:373-461) on a file image: element/property bookkeeping exactly as the reference does it
 * (every `property` line of every element adds its size to the vertex record; `list` and unknown types count 4 bytes; exact-match
 * `ply` / `end_header` lines).  Outputs: the record format, the header's vertex count, the byte offset of the first record and
 * whether the body is binary.  *n_records = min(vertex_count, whole records present) for binary bodies (a truncated last record is
 * dropped, :321-324), vertex_count for ASCII.  Returns B2LO_E_ARG for the files the reference rejects (no x/y/z, no vertices, no header). */
int b2lo_ply_parse_header(const void* file, size_t len, b2lo_record_fmt* fmt, size_t* vertex_count, size_t* data_offset, int* is_binary,
                          size_t* n_records);
/* ASCII PLY body (ply_player.cpp:344-364) on the host: one vertex per line, every whitespace-separated float of the line is read; lines
 * with fewer values than header properties are skipped.  out_xyz: capacity >= vertex_count points; *n receives the points kept. */
int b2lo_ply_read_ascii(const void* file, size_t len, float* out_xyz, size_t cap, size_t* n);
/* FastVoxelFilter::filter over a record stream (host image / device-resident image); results as b2lo_filter / b2lo_filter_dev */
int b2lo_filter_records(b2lo_ctx* ctx, const void* records, size_t n_records, const b2lo_record_fmt* fmt, int stride, float voxel_size,
                        float* out_xyz, uint64_t* out_keys /*nullable*/, size_t* m);
int b2lo_filter_records_dev(b2lo_ctx* ctx, const void* records_dev, size_t n_records, const b2lo_record_fmt* fmt, int stride, float voxel_size);

/* ---- VoxelMap ------------------------------------------------------------------------------------ */
int b2lo_map_create(b2lo_ctx* ctx, float voxel_size, int hierarchy_factor, float planarity_threshold, int compute_surfels,
                    size_t l0_capacity_hint, b2lo_map** out);                         /* VoxelMap(), Set* (VoxelMap.h:195-209) */
int b2lo_map_destroy(b2lo_map* map);
int b2lo_map_clear(b2lo_map* map);                                                    /* Clear (VoxelMap.cpp:122-126) */
int b2lo_map_set_planarity_threshold(b2lo_map* map, float thr);
int b2lo_map_set_compute_surfels(b2lo_map* map, int on);
/* UpdateVoxelMap(cloud, sensor_position, max_distance, is_keyframe=true) (VoxelMap.cpp:128-262) */
int b2lo_map_update(b2lo_map* map, const float* world_xyz, size_t m, size_t stride_floats, const double sensor[3], double max_distance);
int b2lo_map_counts(b2lo_map* map, size_t* l0, size_t* l1, size_t* surfels);          /* GetVoxelCount/GetL1VoxelCount/GetSurfelCount */
int b2lo_map_lookup(b2lo_map* map, const float p[3], float n[3], float c[3]);         /* GetSurfelAtPoint: 1 found, 0 not */
/* GetPointCloud (VoxelMap.cpp:388-403): L0 centroids in the reference's dense (insertion/swap-erase) order */
int b2lo_map_export_l0(b2lo_map* map, float* xyz, int* keys /*nullable, 3 per voxel*/, int* counts /*nullable*/, size_t cap, size_t* n);
/* GetL1Surfels (VoxelMap.cpp:405-418); order unspecified (the reference's L1 order is unobservable elsewhere) */
int b2lo_map_export_surfels(b2lo_map* map, float* centroid, float* normal, float* planarity, int* l1keys /*nullable*/, size_t cap, size_t* n);
/* full L1 dump for parity tests: children as L0 keys in the reference's child-set order */
int b2lo_map_export_l1(b2lo_map* map, int* keys, int* nchild, int* children /*27*3 per L1*/, int* has_surfel, float* normal,
                       float* centroid, float* planarity, int* last_child_count, size_t cap, size_t* n);
int b2lo_map_rebuild_knn(b2lo_map* map);                                              /* RebuildKdTree (VoxelMap.cpp:420-438) */
int b2lo_map_has_knn(b2lo_map* map);                                                  /* HasKdTree (VoxelMap.h:268) */
int b2lo_map_transform_rehash(b2lo_map* map, const float T16[16]);                    /* ApplyTransformAndRehash (:264-302) */

/* ---- final-map export ----------------------------------------------------------------------
 * util::VoxelGrid::filter (src/util/PointCloudUtils.h:462-557) as Estimator::save_map_to_ply applies it to the accumulated keyframe clouds
 * (src/processing/Estimator.cpp:1248-1305): one running-average centroid per leaf-sized voxel, emitted in std::map<VoxelKey> order
 * (x, then y, then z).  out_xyz: capacity `cap` points; *m receives the voxel count (also on B2LO_E_CAPACITY).  Points with a non-finite
 * coordinate or beyond +-2^20 leaves are dropped.  Overwrites the context's feature buffer. */
int b2lo_voxel_grid_filter(b2lo_ctx* ctx, const float* xyz, size_t n, size_t stride_floats, float leaf_size, float* out_xyz, size_t cap, size_t* m);

/* ---- IterativeClosestPointOptimizer ------------------------------------------------------------- */
/* find_correspondences (ICP.cpp:587-645) at a fixed pose; per-query taps for bit-exact parity:
 * state 0 no surfel / 1 gated out / 2 accepted; l1key 3 ints; morton = VoxelKeyHash; residual f64. */
int b2lo_icp_correspondences(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T16[16],
                             double max_distance, int* state, int* l1key, uint64_t* morton, float* normal, float* centroid,
                             double* residual, size_t* n_accepted);
/* find_correspondences_kdtree (ICP.cpp:647-767) at a fixed pose: knn = m*5 indices into the b2lo_map_export_l0 order
 * (-1 when fewer were found), d2 = m*5 f32 squared distances, found per query, state 0 (<5 / collinear) / 1 gated / 2 accepted;
 * normal/centroid are the f32 casts the Gauss-Newton loop consumes; *n_scanned = queries that needed the exact full scan. */
int b2lo_icp_correspondences_knn(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T16[16],
                                 double max_distance, int* knn, float* d2, int* found, int* state, float* normal, float* centroid,
                                 double* residual, size_t* n_accepted, size_t* n_scanned);
/* optimize (ICP.cpp:255-463).  T_init is used verbatim (SE3f initial_transform); on B2LO_S_INSUFFICIENT
 * T_out = T_init, as the reference leaves optimized_transform. */
int b2lo_icp_optimize(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16],
                      const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats /*nullable*/);
/* parity tap: ONE Gauss-Newton iteration (the loop body ICP.cpp:280-448: correspondences, PKO delta, normal equations, solve, SE(3)
 * update) entered at pose T_in.  scale > 0: the residual normalisation scale an earlier iteration fixed (ICP.cpp:304-316 computes it at
 * iteration 0 only and reuses it); scale <= 0: computed from this pose's residuals.  Lets a test feed the reference's pose of iteration k
 * (teacher forcing) and compare C, alpha, H, g and the updated pose of every iteration, not only of iteration 0. */
int b2lo_icp_iterate(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_in[16], double scale,
                     const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats /*nullable*/);
/* same, query cloud = the context's feature buffer left by b2lo_filter / b2lo_filter_dev */
int b2lo_icp_optimize_features(b2lo_map* map, const float T_init[16], const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats);

/* ---- point-sharded scan-to-map ICP for dense scans: this rank holds a contiguous slice of the query cloud, the map
 * is replicated.  One Gauss-Newton iteration = shard_corr -> [all-gather 3 doubles: C_r, sum r, sum r^2] -> shard_sample ->
 * [all-reduce gmm_sample_size doubles] -> shard_accumulate -> [all-reduce 28 doubles: 21 lower-triangle H, 6 g, cost] ->
 * shard_finish (every rank solves the identical 6x6 system).  The *_dev buffers are DEVICE pointers owned by the caller (the
 * tensors handed to NCCL); every call enqueues on the context stream, only shard_finish synchronises.  scale = the residual
 * normalisation sqrt(var)/6 of iteration 0 (ICP.cpp:304-316) computed by the caller from the gathered moments. */
int b2lo_icp_shard_begin(b2lo_map* map, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16], const b2lo_icp_cfg* cfg);
int b2lo_icp_shard_corr(b2lo_map* map, const b2lo_icp_cfg* cfg, double* stats3_dev);
int b2lo_icp_shard_sample(b2lo_map* map, const b2lo_icp_cfg* cfg, long long offset, long long c_total, double scale, double* sample_dev /*128 doubles*/);
int b2lo_icp_shard_accumulate(b2lo_map* map, const b2lo_icp_cfg* cfg, long long c_total, double scale, const double* sample_dev, double* acc28_dev);
int b2lo_icp_shard_finish(b2lo_map* map, const b2lo_icp_cfg* cfg, const double* acc28_dev, float T_out[16], int* done /*0 go on, 1 finished, 2 failed*/,
                          b2lo_icp_stats* stats /*nullable*/);

/* The same loop in ONE call with device-ordered exchanges: the all-gather of the counts / moments, the all-reduce of the global GMM
 * sample and the all-reduce of the 28 normal-equation sums are NCCL calls enqueued on the context stream between the kernels (no host
 * synchronisation inside the Gauss-Newton loop; the counts, the C < min test and the residual scale are evaluated on the device from the
 * gathered values).  NCCL is the one already loaded in the process (dlopen "libnccl.so.2"); the communicator is created from a unique id
 * that rank 0 makes with b2lo_shard_unique_id and the host distributes (128 bytes).  world = 1 needs no NCCL (local copies).  Every rank
 * gets the same pose.  collective_ms (nullable): CUDA-event time of the exchanges of the last iteration issued. */
typedef struct b2lo_shard_comm b2lo_shard_comm;
int b2lo_shard_unique_id(void* out, size_t bytes /* >= 128 */);
int b2lo_shard_comm_create(b2lo_ctx* ctx, int world, int rank, const void* unique_id, size_t bytes, b2lo_shard_comm** out);
int b2lo_shard_comm_destroy(b2lo_shard_comm* comm);
/* Peer-memory exchange instead of NCCL (one process per GPU on one NVLink / NVSwitch node, <= 8 ranks): every rank owns a mailbox in
 * its HBM; b2lo_shard_comm_ipc_handle returns its cudaIpc handle (64 bytes), the host hands all handles to every rank (rank order) and
 * b2lo_shard_comm_open_peers maps them.  From then on each exchange of b2lo_icp_shard_optimize is ONE small kernel per rank that stores
 * its payload straight into every peer's mailbox over NVLink, publishes an epoch and adds the peers' payloads in rank order (identical
 * bits on every rank) - no NCCL launch.  b2lo_shard_comm_create may then be given unique_id = NULL (no NCCL at all).  A rank that never
 * arrives makes the others give up after ~3 s with B2LO_E_CUDA instead of hanging the GPU. */
int b2lo_shard_comm_ipc_handle(b2lo_shard_comm* comm, void* out, size_t bytes /* >= 64 */);
int b2lo_shard_comm_open_peers(b2lo_shard_comm* comm, const void* handles /* world x bytes_each, rank order */, size_t bytes_each);
int b2lo_icp_shard_optimize(b2lo_map* map, b2lo_shard_comm* comm, const float* local_xyz, size_t m, size_t stride_floats, const float T_init[16],
                            const b2lo_icp_cfg* cfg, float T_out[16], b2lo_icp_stats* stats /*nullable*/, float* collective_ms /*nullable*/);

/* ---- loop-closure ICP ----------------------------------------------------------------------
 * optimize_loop (IterativeClosestPointOptimizer.cpp:40-251; correspondences :465-585): registers the CURRENT keyframe's local
 * feature cloud (world pose T_curr) against a MATCHED keyframe's local feature cloud (world pose T_matched): exact 5-NN in the
 * matched cloud moved to world coordinates, plane through the 5 neighbours, no distance gate, residual anchored at the nearest
 * neighbour, PKO-weighted Gauss-Newton for at most 100 iterations, then the 1-NN (< 1 m) inlier ratio.
 * Returns B2LO_OK when the reference returns true (converged AND inlier ratio >= 0.5): T_rel = T_curr^-1 * optimised pose.
 * B2LO_S_INSUFFICIENT otherwise (T_rel is still written once the loop converged, as the reference does; identity before that).
 * The reference runs this on a background thread next to optimize(): give it its own b2lo_ctx (own stream and scratch). */
int b2lo_icp_optimize_loop(b2lo_ctx* ctx, const float* curr_xyz, size_t m_curr, size_t curr_stride_floats, const float T_curr[16],
                           const float* matched_xyz, size_t m_matched, size_t matched_stride_floats, const float T_matched[16],
                           const b2lo_icp_cfg* cfg, float T_rel[16], float* inlier_ratio /*nullable*/, b2lo_icp_stats* stats /*nullable*/);

/* ---- pose algebra used at the boundary (util::SE3 / SO3, MathUtils.h:57-168) --------------------- */
void b2lo_se3_mul(const float A16[16], const float B16[16], float C16[16]);  /* SE3::operator*, re-projects the rotation */
void b2lo_se3_inv(const float A16[16], float C16[16]);                       /* SE3::Inverse */
void b2lo_se3_from_rt(const float T16_in[16], float T16_out[16]);            /* SE3(Matrix3f, Vector3f) ctor re-projection */
void b2lo_so3_log(const float T16[16], float w[3]);                          /* SO3::Log of the rotation block */
void b2lo_so3_exp(const float w[3], float R9[9]);                            /* SO3::Exp (MathUtils.cpp:23-39), row-major 3x3 */
/* the engine's small dense numerics, exposed so parity tests can pin them on the host (same inline code the kernels run) */
void b2lo_svd3(const float A9[9], float U9[9], float S3[3], float V9[9]);    /* JacobiSVD<Matrix3f> (VoxelMap.cpp:239, MathUtils.cpp:88) */
void b2lo_ldlt6_solve(const float H36[36], const float b6[6], float x6[6]);  /* Matrix<float,6,6>::ldlt().solve (ICP.cpp:418) */
void b2lo_fit_plane(const float* pts, int n, float mu[3], float normal[3], float* planarity); /* surfel PCA (VoxelMap.cpp:223-242) */
uint64_t b2lo_voxel_key_hash(int x, int y, int z);                           /* VoxelKeyHash (VoxelMap.h:166-183) */

/* ---- per-scan driver kept on the device (Estimator.cpp:116-233, 349-368, 449-470) -- */
typedef struct {
  float voxel_size; int point_stride; float map_voxel_size; double max_range;
  float surfel_planarity_threshold; double keyframe_distance_threshold, keyframe_rotation_threshold;
  b2lo_icp_cfg icp;
} b2lo_odom_cfg;
typedef struct {
  float pose[16];
  int keyframe, icp_status, n_features, n_corr, n_iters;
  float device_ms;       /* whole scan, CUDA events */
  size_t l0, l1;
} b2lo_odom_result;
void b2lo_default_odom_cfg(b2lo_odom_cfg* cfg, int mid360);
int b2lo_odom_create(b2lo_ctx* ctx, const b2lo_odom_cfg* cfg, b2lo_odom** out);
int b2lo_odom_destroy(b2lo_odom* od);
int b2lo_odom_reset(b2lo_odom* od);   /* clear the map and the pose state (new sequence) */
/* steady-state scans replay one captured CUDA graph (K1 -> ICP -> pose/keyframe decision -> gated K6 -> read-backs); how often it was
 * replayed / rebuilt (buffers grew) and how many kernels one replay holds.  B2LO_NO_GRAPH=1 in the environment disables the capture. */
int b2lo_odom_graph_stats(b2lo_odom* od, long long* replays, long long* builds, long long* kernels_per_replay);
b2lo_map* b2lo_odom_map(b2lo_odom* od);
/* process_frame on a host scan (H2D inside) or on a scan already resident in HBM */
int b2lo_odom_process(b2lo_odom* od, const float* xyz, size_t n, size_t stride_floats, b2lo_odom_result* res);
int b2lo_odom_process_dev(b2lo_odom* od, const float* xyz_dev, size_t n, size_t stride_floats, b2lo_odom_result* res);
/* Look-ahead for recorded sequences - an EXTENSION of this engine, not a call pattern of the reference: its players load a scan and then
 * process it, strictly one after the other (app/player/kitti_player.cpp:109-123; there is no prefetch thread).  A player that has the
 * next scan in memory anyway (a recorded dataset) may announce, BEFORE processing scan i, the buffer that the following
 * b2lo_odom_process{,_dev} call will be given.
 * Its voxel downsample (K1, preprocess_frame, Estimator.cpp:561-589) then runs on a side stream into a second feature buffer while
 * scan i registers, and the next call finds its features ready.  Results are bit-identical with and without it; K1 does not depend on
 * the pose or the map.  The buffer must stay valid and unchanged until that next call returns.  on_device = 0: host memory, which
 * must be page-locked and mapped (cudaMallocHost / cudaHostRegister; K1 reads it in place) - for pageable memory the announcement is
 * ignored and B2LO_S_EMPTY is returned.  A following call with another buffer / size simply runs its own K1. */
int b2lo_odom_lookahead(b2lo_odom* od, const float* xyz_next, size_t n, size_t stride_floats, int on_device);
/* b2lo_odom_lookahead(next host scan) + b2lo_odom_process(this host scan) in one call; next_xyz = NULL: no announcement */
int b2lo_odom_process_la(b2lo_odom* od, const float* xyz, size_t n, size_t stride_floats, const float* next_xyz, size_t next_n,
                         size_t next_stride_floats, b2lo_odom_result* res);
/* Record-stream input for the per-scan driver: after this call b2lo_odom_process / _process_dev / _lookahead take `xyz` as the
 * address of the first RECORD of a file image in the given format and `n` as the record count (stride_floats is ignored; pass 3).
 * fmt = NULL returns to float-stride clouds.  Takes effect with the next scan. */
int b2lo_odom_set_record_fmt(b2lo_odom* od, const b2lo_record_fmt* fmt);
/* Throughput mode for batches of recorded sequences on one GPU: one scan on each of `count` INDEPENDENT sequences (own b2lo_odom, own
 * b2lo_ctx / stream each; the caller created them), all scans resident in HBM.  Every sequence's launch sequence (one CUDA graph replay) is
 * enqueued before the first result is waited for, so the sequences overlap on the device from a single host thread: one sequence keeps
 * about one SM busy (the PKO fit and the Gauss-Newton finish are one-CTA latency chains), the other 147 are free.  next_xyz_dev / next_n
 * (nullable, per sequence) announce the scans of the NEXT call like b2lo_odom_lookahead(..., on_device = 1).  Results are those of
 * b2lo_odom_process_dev on each sequence alone, bit for bit.  Returns the first error, else the last soft outcome, else B2LO_OK; res[i]
 * is filled per sequence (res[i].device_ms = that sequence's own CUDA-event time, overlapping with the others). */
int b2lo_odom_process_batch_dev(b2lo_odom* const* ods, const float* const* xyz_dev, const size_t* n, const float* const* next_xyz_dev,
                                const size_t* next_n, size_t stride_floats, int count, b2lo_odom_result* res);

/* Lock-step batches: S independent sequences on one GPU advance by ONE scan per call, and every kernel of the scan is started once per
 * call for all of them (grid.y = sequence; per-sequence arguments in a device array) inside one replayed CUDA graph.  The single-CTA
 * latency chains of a scan (PKO fit, Gauss-Newton finish, update close) run for S sequences at once on S SMs, and the GPU's front end
 * dispatches ~30 kernels per STEP instead of ~30 per scan (what caps b2lo_odom_process_batch_dev).  Results per sequence are those of the
 * sequence processed alone, bit for bit.  The sequences (own b2lo_odom and own b2lo_ctx each, surfel mode, same device) stay usable on
 * their own between calls; first frames, empty scans and record-stream input fall back to the per-sequence path inside the call.
 * A batch of 32 or more sequences runs as forked branches of the step's graph (one per 16 sequences, at most eight), so that the
 * single-SM tails of one branch overlap the kernels of the others: ~100 k scans/s for 128 sequences, ~125 k for 384 on one B200.
 * res[i].device_ms = CUDA-event time of the whole step (all sequences share it). */
typedef struct b2lo_lockstep b2lo_lockstep;
int b2lo_lockstep_create(b2lo_odom* const* ods, int count, b2lo_lockstep** out);
int b2lo_lockstep_destroy(b2lo_lockstep* ls);
int b2lo_lockstep_process_dev(b2lo_lockstep* ls, const float* const* xyz_dev, const size_t* n, size_t stride_floats, b2lo_odom_result* res,
                              float* device_ms /*nullable*/);
int b2lo_lockstep_stats(b2lo_lockstep* ls, long long* kernels_per_step, long long* replays, long long* builds, long long* fallbacks);

#ifdef __cplusplus
}
#endif
#endif /* B2LO_H */
