// ORACLE - TEST INFRASTRUCTURE ONLY.  Stand-in for /root/reference/src/viewer/PangolinViewer.h (Pangolin / OpenGL: absent here); see
// oracle/ref_stubs/player/processing/Estimator.h.  Every member the player touches exists and does nothing.
#pragma once
namespace lidar_slam {
namespace viewer {
class PangolinViewer {
 public:
  template <class... A> bool initialize(A&&...) { return false; }
  bool is_ready() const { return true; }
  bool should_close() const { return true; }
  void shutdown() {}
  template <class... A> void update_current_frame(A&&...) {}
  template <class... A> void add_trajectory_frame(A&&...) {}
  template <class... A> void update_map_points(A&&...) {}
  template <class... A> void add_keyframe(A&&...) {}
  template <class... A> void update_last_keyframe(A&&...) {}
  template <class... A> void update_voxel_map(A&&...) {}
  template <class... A> void update_icp_debug_clouds(A&&...) {}
  template <class... A> void process_keyboard_input(A&&...) {}
  template <class... A> void set_frame_info(A&&...) {}
  template <class... A> void update_statistics(A&&...) {}
};
}  // namespace viewer
}  // namespace lidar_slam
