// ORACLE - TEST INFRASTRUCTURE ONLY.  Stand-in for /root/reference/src/processing/Estimator.h, found first on the include path when
// oracle/Makefile compiles the UNMODIFIED app/player/ply_player.cpp for its PLY reader (parse_ply_header / load_ply_point_cloud,
// ply_player.cpp:267-461).  The real header pulls the loop detector and the pose graph (OpenCV, third-party solvers: absent here); the
// player's run loop only needs these members to exist - none of them is called by the functions the tests use.
#pragma once
#include <cstddef>
#include <memory>
#include "database/LidarFrame.h"
#include "util/ConfigUtils.h"
#include "util/MathUtils.h"
#include "util/PointCloudUtils.h"

namespace lidar_slam {
namespace processing {
class Estimator {
 public:
  template <class... A> explicit Estimator(A&&...) {}
  template <class... A> bool process_frame(A&&...) { return false; }
  util::SE3f get_current_pose() const { return util::SE3f(); }
  util::PointCloudConstPtr get_local_map() const { return nullptr; }
  std::size_t get_keyframe_count() const { return 0; }
  std::shared_ptr<database::LidarFrame> get_keyframe(std::size_t) const { return nullptr; }
  const void* get_voxel_map() const { return nullptr; }
  template <class... A> void get_debug_clouds(A&&...) const {}
  template <class... A> void get_optimization_statistics(A&&...) const {}
  template <class... A> bool save_map_to_ply(A&&...) const { return false; }
  template <class... A> void enable_loop_closure(A&&...) {}
  template <class... A> void print_timing_statistics(A&&...) const {}
};
}  // namespace processing
}  // namespace lidar_slam
