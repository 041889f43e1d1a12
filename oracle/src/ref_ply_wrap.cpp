// ORACLE - TEST INFRASTRUCTURE ONLY.  C entry point over the REAL PLY reader of the reference: the UNMODIFIED
// /root/reference/app/player/ply_player.cpp (PLYPlayer::parse_ply_header / load_ply_point_cloud, :267-461) compiled by oracle/Makefile
// into oracle/_ref/libref_ply.so.  That translation unit also holds the player's run loop, which names processing::Estimator and
// viewer::PangolinViewer; their headers (OpenCV, Pangolin, the pose-graph solver) are replaced by the do-nothing stand-ins of
// oracle/ref_stubs/player/ found first on the include path, and Eigen::Quaternionf (trajectory writer only) comes from the force-included
// oracle/ref_stubs/eigen_quaternion.h.  Nothing of that is executed here: the wrapper only calls the two reader functions
// (private members: this file alone is compiled with -fno-access-control) so that tests/test_reference_core.py can compare
// oracle/include/orc_ingest.hpp with them on the same files.
#include <cstring>
#include <string>
#include "ply_player.h"
#include "util/ConfigUtils.h"
#include "util/LogUtils.h"
#include "util/PointCloudUtils.h"

// the two symbols of src/util/ConfigUtils.cpp (yaml-cpp) the player's run loop references; never called here
lidar_slam::util::SystemConfig lidar_slam::util::ConfigManager::create_default_config() { return lidar_slam::util::SystemConfig(); }
bool lidar_slam::util::ConfigManager::load_from_file(const std::string&) { return false; }

static const bool g_quiet = [] { lidar_slam::Logger::level = static_cast<lidar_slam::LogLevel>(4); return true; }();

extern "C" {
// returns 1 and the points when the reference loads the file, 0 when it rejects it (nullptr or empty cloud), -1 if cap is too small
int ref_ply_load_file(const char* path, float* out_xyz, size_t cap, size_t* n) {
  lidar_slam::app::PLYPlayer player;
  auto cloud = player.load_ply_point_cloud(path);
  *n = 0;
  if (!cloud) return 0;
  *n = cloud->size();
  if (cloud->size() > cap) return -1;
  for (size_t i = 0; i < cloud->size(); ++i) { out_xyz[i * 3] = (*cloud)[i].x; out_xyz[i * 3 + 1] = (*cloud)[i].y; out_xyz[i * 3 + 2] = (*cloud)[i].z; }
  return cloud->empty() ? 0 : 1;
}
// header facts: returns parse_ply_header's bool; *n_props = number of vertex properties, stride = sum of their byte sizes
int ref_ply_parse_header(const char* path, size_t* vertex_count, int* is_binary, int* n_props, size_t* stride) {
  lidar_slam::app::PLYPlayer player;
  size_t vc = 0; bool bin = false;
  std::vector<lidar_slam::app::PLYPlayer::PLYPropertyInfo> props;
  const bool ok = player.parse_ply_header(path, vc, props, bin);
  *vertex_count = vc; *is_binary = bin ? 1 : 0; *n_props = (int)props.size();
  size_t s = 0;
  for (const auto& p : props) s += p.byte_size;
  *stride = s;
  return ok ? 1 : 0;
}
}
