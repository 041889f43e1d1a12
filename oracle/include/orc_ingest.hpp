// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's scan loaders, on file images held in memory:
//   util::load_kitti_binary          /root/reference/src/util/PointCloudUtils.cpp:19-65
//   PLYPlayer::parse_ply_header      /root/reference/app/player/ply_player.cpp:373-461
//   PLYPlayer::load_ply_point_cloud  /root/reference/app/player/ply_player.cpp:267-371
// The reference reads through std::ifstream; the same iostream extraction rules are kept here by running std::istringstream over
// the image (getline for lines, operator>> for tokens, size_t and float).  Parity unpinned: the reference holds no fixtures for its
// loaders and ply_player.cpp cannot be compiled here (it pulls Estimator, the viewer and Eigen).
#pragma once
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace orc {

struct PlyProp { std::string name; size_t bytes; };
struct PlyInfo { size_t vertices = 0; std::vector<PlyProp> props; bool binary = false; int ix = -1, iy = -1, iz = -1; };

inline size_t ply_type_bytes(const std::string& t) {  // :389-395
  for (const char* s : {"char", "uchar", "int8", "uint8"}) if (t == s) return 1;
  for (const char* s : {"short", "ushort", "int16", "uint16"}) if (t == s) return 2;
  for (const char* s : {"int", "uint", "float", "int32", "uint32", "float32"}) if (t == s) return 4;
  for (const char* s : {"double", "float64"}) if (t == s) return 8;
  return 4;
}

// :373-461 — false when x/y/z are missing or the vertex count is 0
inline bool ply_header(const std::string& image, PlyInfo& info) {
  std::istringstream in(image);
  std::string line;
  bool started = false;
  while (std::getline(in, line)) {
    if (line == "ply") { started = true; continue; }
    if (!started) continue;
    if (line == "end_header") break;
    std::istringstream ls(line);
    std::string kw;
    ls >> kw;
    if (kw == "format") {
      std::string f; ls >> f;
      info.binary = f == "binary_little_endian" || f == "binary_big_endian";
    } else if (kw == "element") {
      std::string what; ls >> what;
      if (what == "vertex") ls >> info.vertices;
    } else if (kw == "property") {
      std::string type, name; ls >> type >> name;
      info.props.push_back({name, ply_type_bytes(type)});
      const int at = (int)info.props.size() - 1;
      if (name == "x") info.ix = at; else if (name == "y") info.iy = at; else if (name == "z") info.iz = at;
    }
  }
  return info.ix >= 0 && info.iy >= 0 && info.iz >= 0 && info.vertices != 0;
}

// :267-371 — xyz of every vertex the reference would push_back
inline std::vector<float> ply_load(const std::string& image) {
  std::vector<float> out;
  PlyInfo info;
  if (!ply_header(image, info)) return out;
  std::istringstream in(image);
  std::string line;
  while (std::getline(in, line)) if (line == "end_header") break;
  if (info.binary) {
    size_t rec = 0;
    for (const auto& p : info.props) rec += p.bytes;
    std::vector<char> buf(rec);
    for (size_t i = 0; i < info.vertices; ++i) {
      in.read(buf.data(), (std::streamsize)rec);
      if (!in.good()) break;
      float v[3] = {0, 0, 0};
      size_t off = 0;
      for (size_t k = 0; k < info.props.size(); ++k) {
        if ((int)k == info.ix) std::memcpy(&v[0], buf.data() + off, 4);
        else if ((int)k == info.iy) std::memcpy(&v[1], buf.data() + off, 4);
        else if ((int)k == info.iz) std::memcpy(&v[2], buf.data() + off, 4);
        off += info.props[k].bytes;
      }
      out.insert(out.end(), v, v + 3);
    }
  } else {
    for (size_t i = 0; i < info.vertices; ++i) {
      if (!std::getline(in, line)) break;
      std::istringstream ls(line);
      std::vector<float> vals;
      float f;
      while (ls >> f) vals.push_back(f);
      if (vals.size() >= info.props.size()) { out.push_back(vals[info.ix]); out.push_back(vals[info.iy]); out.push_back(vals[info.iz]); }
    }
  }
  return out;
}

// PointCloudUtils.cpp:19-65 — whole 16-byte records only, intensity dropped
inline std::vector<float> kitti_load(const std::string& image) {
  std::vector<float> out;
  const size_t n = image.size() / 16;
  out.resize(3 * n);
  for (size_t i = 0; i < n; ++i) std::memcpy(&out[3 * i], image.data() + 16 * i, 12);
  return out;
}

}  // namespace orc
