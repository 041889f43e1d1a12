// b2lo_ingest.cu — scan ingest: the dataset's own records are what K1 reads.
//
// Reference loaders this replaces on the way into the hot path:
//   util::load_kitti_binary            /root/reference/src/util/PointCloudUtils.cpp:19-65   (16 B records x,y,z,intensity; xyz kept)
//   PLYPlayer::parse_ply_header        /root/reference/app/player/ply_player.cpp:373-461
//   PLYPlayer::load_ply_point_cloud    /root/reference/app/player/ply_player.cpp:267-371    (binary: bytewise x/y/z at property offsets)
// Both build a util::PointCloud (AoS 12 B) on the host that FastVoxelFilter::filter then strides over.  Here the file image itself
// (page-locked, or already in HBM) is the K1 input: a b2lo_record_fmt says where the three floats sit in a record, and k_flt_insert
// reads every stride-th record in place.  Only the header parse (a few hundred bytes of text) and ASCII bodies stay on the host.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "b2lo_internal.h"

using namespace b2;

extern "C" void b2lo_kitti_record_fmt(b2lo_record_fmt* fmt) {
  if (!fmt) return;
  fmt->record_bytes = 16; fmt->off_x = 0; fmt->off_y = 4; fmt->off_z = 8;
}

namespace {

// std::getline over a memory image: [*pos, next '\n') without the newline; false at the end of the image
bool next_line(const char* p, size_t len, size_t* pos, const char** line, size_t* n) {
  if (*pos >= len) return false;
  const char* s = p + *pos;
  const void* nl = std::memchr(s, '\n', len - *pos);
  size_t l = nl ? (size_t)(static_cast<const char*>(nl) - s) : len - *pos;
  *line = s; *n = l;
  *pos += l + (nl ? 1 : 0);
  return true;
}
bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
// operator>>(istream&, string&): skip whitespace, take the run of non-whitespace
std::string next_token(const char* s, size_t n, size_t* at) {
  size_t i = *at;
  while (i < n && is_space(s[i])) ++i;
  size_t b = i;
  while (i < n && !is_space(s[i])) ++i;
  *at = i;
  return std::string(s + b, i - b);
}
size_t type_size(const std::string& t) {  // get_type_size lambda, ply_player.cpp:389-395
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 4;
}
bool line_is(const char* s, size_t n, const char* lit) { size_t l = std::strlen(lit); return n == l && std::memcmp(s, lit, l) == 0; }

struct PlyHeader {
  size_t vertex_count = 0, n_props = 0, record_bytes = 0, data_offset = 0;
  long x = -1, y = -1, z = -1;
  size_t off[3] = {0, 0, 0};
  bool binary = false;
};

// the header bookkeeping of parse_ply_header + the offsets load_ply_point_cloud derives from it (:287-337)
bool parse_header(const char* p, size_t len, PlyHeader* h) {
  size_t pos = 0, n;
  const char* line;
  bool in_header = false, ended = false;
  std::vector<size_t> sizes;
  while (next_line(p, len, &pos, &line, &n)) {
    if (line_is(line, n, "ply")) { in_header = true; continue; }
    if (!in_header) continue;
    if (line_is(line, n, "end_header")) { ended = true; break; }
    size_t at = 0;
    const std::string tok = next_token(line, n, &at);
    if (tok == "format") {
      const std::string f = next_token(line, n, &at);
      h->binary = (f == "binary_little_endian" || f == "binary_big_endian");
    } else if (tok == "element") {
      const std::string e = next_token(line, n, &at);
      if (e == "vertex") {
        // iss >> size_t: digits only (an optional '+'), 0 on failure
        const std::string v = next_token(line, n, &at);
        size_t i = 0, val = 0;
        bool any = false;
        if (i < v.size() && v[i] == '+') ++i;
        for (; i < v.size() && v[i] >= '0' && v[i] <= '9'; ++i) { val = val * 10 + (size_t)(v[i] - '0'); any = true; }
        h->vertex_count = any ? val : 0;
      }
    } else if (tok == "property") {
      const std::string type = next_token(line, n, &at);
      const std::string name = next_token(line, n, &at);
      const long idx = (long)sizes.size();
      sizes.push_back(type_size(type));
      if (name == "x") h->x = idx; else if (name == "y") h->y = idx; else if (name == "z") h->z = idx;
    }
  }
  h->data_offset = ended ? pos : len;   // the body reader skips lines up to end_header; without one nothing is left to read
  h->n_props = sizes.size();
  if (h->x < 0 || h->y < 0 || h->z < 0) return false;
  if (h->vertex_count == 0) return false;
  size_t o = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    if ((long)i == h->x) h->off[0] = o;
    if ((long)i == h->y) h->off[1] = o;
    if ((long)i == h->z) h->off[2] = o;
    o += sizes[i];
  }
  h->record_bytes = o;
  return true;
}

// operator>>(istream&, float&) over a line: leading whitespace skipped, decimal floats only (num_get accepts neither inf/nan nor hex)
bool next_float(const char* s, size_t n, size_t* at, float* out) {
  size_t i = *at;
  while (i < n && is_space(s[i])) ++i;
  if (i >= n) return false;
  size_t j = i;
  if (s[j] == '+' || s[j] == '-') ++j;
  if (j >= n || !((s[j] >= '0' && s[j] <= '9') || s[j] == '.')) return false;
  // the longest decimal-float prefix: digits [. digits] [e[+-]digits]
  size_t k = j;
  bool digits = false;
  while (k < n && s[k] >= '0' && s[k] <= '9') { ++k; digits = true; }
  if (k < n && s[k] == '.') { ++k; while (k < n && s[k] >= '0' && s[k] <= '9') { ++k; digits = true; } }
  if (!digits) return false;
  if (k < n && (s[k] == 'e' || s[k] == 'E')) {
    size_t e = k + 1;
    if (e < n && (s[e] == '+' || s[e] == '-')) ++e;
    if (e < n && s[e] >= '0' && s[e] <= '9') { while (e < n && s[e] >= '0' && s[e] <= '9') ++e; k = e; }
    else return false;   // "1e" / "1e+": num_get consumes the exponent marker and fails
  }
  const std::string num(s + i, k - i);
  *out = std::strtof(num.c_str(), nullptr);
  *at = k;
  return true;
}

}  // namespace

extern "C" int b2lo_ply_parse_header(const void* file, size_t len, b2lo_record_fmt* fmt, size_t* vertex_count, size_t* data_offset, int* is_binary,
                                     size_t* n_records) {
  if (!file || !fmt) return B2LO_E_ARG;
  PlyHeader h;
  const bool ok = parse_header(static_cast<const char*>(file), len, &h);
  if (vertex_count) *vertex_count = h.vertex_count;
  if (data_offset) *data_offset = h.data_offset;
  if (is_binary) *is_binary = h.binary ? 1 : 0;
  if (n_records) *n_records = 0;
  if (!ok) { set_error("ply: header rejected (missing x/y/z, no vertices, or no header)"); return B2LO_E_ARG; }
  if (h.record_bytes > 0xffffffffull) { set_error("ply: vertex record too large"); return B2LO_E_ARG; }
  fmt->record_bytes = (uint32_t)h.record_bytes; fmt->off_x = (uint32_t)h.off[0]; fmt->off_y = (uint32_t)h.off[1]; fmt->off_z = (uint32_t)h.off[2];
  if (n_records) {
    if (h.binary) {
      const size_t whole = h.record_bytes ? (len - h.data_offset) / h.record_bytes : 0;
      *n_records = whole < h.vertex_count ? whole : h.vertex_count;
    } else *n_records = h.vertex_count;
  }
  return B2LO_OK;
}

extern "C" int b2lo_ply_read_ascii(const void* file, size_t len, float* out_xyz, size_t cap, size_t* n_out) {
  if (!file || !n_out) return B2LO_E_ARG;
  *n_out = 0;
  PlyHeader h;
  const char* p = static_cast<const char*>(file);
  if (!parse_header(p, len, &h)) { set_error("ply: header rejected (missing x/y/z, no vertices, or no header)"); return B2LO_E_ARG; }
  if (h.binary) { set_error("ply: body is binary, use b2lo_filter_records"); return B2LO_E_ARG; }
  size_t pos = h.data_offset, n, kept = 0;
  const char* line;
  std::vector<float> vals;
  for (size_t i = 0; i < h.vertex_count; ++i) {
    if (!next_line(p, len, &pos, &line, &n)) break;
    vals.clear();
    size_t at = 0;
    float v;
    while (next_float(line, n, &at, &v)) vals.push_back(v);
    if (vals.size() >= h.n_props) {
      if (kept >= cap || !out_xyz) { set_error("ply: output buffer too small"); return B2LO_E_CAPACITY; }
      out_xyz[3 * kept] = vals[(size_t)h.x]; out_xyz[3 * kept + 1] = vals[(size_t)h.y]; out_xyz[3 * kept + 2] = vals[(size_t)h.z];
      ++kept;
    }
  }
  *n_out = kept;
  return B2LO_OK;
}

static int check_fmt(const b2lo_record_fmt* fmt) {
  if (!fmt || fmt->record_bytes == 0) { set_error("records: no format"); return B2LO_E_ARG; }
  for (int a = 0; a < 3; ++a)
    if ((size_t)(&fmt->off_x)[a] + 4 > fmt->record_bytes) { set_error("records: coordinate offset outside the record"); return B2LO_E_ARG; }
  return B2LO_OK;
}

namespace b2 {
// pageable host image: only the sampled records are touched; their coordinates go through the pinned staging area as packed xyz
int ctx_stage_records_h2d(b2lo_ctx* ctx, const void* bytes, size_t n_records, const b2lo_record_fmt* fmt, size_t take_every) {
  if (take_every < 1) take_every = 1;
  const size_t nt = (n_records + take_every - 1) / take_every;
  int rc = ctx_reserve_points(ctx, nt);
  if (rc) return rc;
  if (ctx->stage_busy) { B2_CUDA(cudaEventSynchronize(ctx->ev_stage)); ctx->stage_busy = false; }
  float* h = ctx->h_stage;
  const unsigned char* b = static_cast<const unsigned char*>(bytes);
  const size_t step = (size_t)fmt->record_bytes * take_every;
  for (size_t j = 0; j < nt; ++j) {
    const unsigned char* r = b + j * step;
    std::memcpy(h + 3 * j, r + fmt->off_x, 4); std::memcpy(h + 3 * j + 1, r + fmt->off_y, 4); std::memcpy(h + 3 * j + 2, r + fmt->off_z, 4);
  }
  if (nt) B2_CUDA(cudaMemcpyAsync(ctx->d_stage, h, nt * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaEventRecord(ctx->ev_stage, ctx->stream));
  ctx->stage_busy = true;
  ctx->h2d_bytes += nt * 3 * sizeof(float);
  return B2LO_OK;
}
}  // namespace b2

extern "C" int b2lo_filter_records_dev(b2lo_ctx* ctx, const void* records_dev, size_t n_records, const b2lo_record_fmt* fmt, int stride, float voxel_size) {
  if (!ctx) return B2LO_E_ARG;
  if (stride < 1 || !(voxel_size > 0.0f)) { set_error("filter: bad stride / voxel size"); return B2LO_E_ARG; }
  int rc = check_fmt(fmt);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  if (!records_dev || n_records == 0) { B2_CUDA(cudaMemsetAsync(ctx->d_nfeat, 0, sizeof(int), ctx->stream)); return B2LO_S_EMPTY; }
  const size_t ns = (n_records + (size_t)stride - 1) / (size_t)stride;
  return filter_run(ctx, static_cast<const float*>(records_dev), ns, (size_t)fmt->record_bytes * (size_t)stride, voxel_size, 0, nullptr, fmt);
}

extern "C" int b2lo_filter_records(b2lo_ctx* ctx, const void* records, size_t n_records, const b2lo_record_fmt* fmt, int stride, float voxel_size,
                                   float* out_xyz, uint64_t* out_keys, size_t* m) {
  if (!ctx || !m) return B2LO_E_ARG;
  *m = 0;
  if (stride < 1 || !(voxel_size > 0.0f)) { set_error("filter: bad stride / voxel size"); return B2LO_E_ARG; }
  int rc = check_fmt(fmt);
  if (rc) return rc;
  if (!records || n_records == 0) return B2LO_S_EMPTY;
  const size_t ns = (n_records + (size_t)stride - 1) / (size_t)stride;
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    // page-locked image: K1 reads the sampled records in place over PCIe; pageable: gather the sampled coordinates, one H2D
    cudaPointerAttributes attr;
    const bool pinned = (cudaPointerGetAttributes(&attr, records) == cudaSuccess) && attr.type == cudaMemoryTypeHost && attr.devicePointer;
    if (!pinned) cudaGetLastError();
    if (pinned && !getenv("B2LO_NO_ZERO_COPY")) {
      ctx->h2d_bytes += ns * 32;
      rc = filter_run(ctx, static_cast<const float*>(attr.devicePointer), ns, (size_t)fmt->record_bytes * (size_t)stride, voxel_size, 0, nullptr, fmt);
    } else {
      rc = ctx_stage_records_h2d(ctx, records, n_records, fmt, (size_t)stride);
      if (!rc) rc = filter_run(ctx, ctx->d_stage, ns, 3, voxel_size);
    }
    if (rc) return rc;
    if (out_keys) {
      std::vector<unsigned long long> keys(ns);
      B2_CUDA(cudaMemcpyAsync(keys.data(), ctx->d_feat_key, ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      B2_CUDA(cudaStreamSynchronize(ctx->stream));
      std::memcpy(out_keys, keys.data(), ns * sizeof(unsigned long long));   // the first *m entries are meaningful
      ctx->d2h_bytes += ns * sizeof(unsigned long long);
    }
  }
  return b2lo_ctx_features(ctx, out_xyz, ns, m);
}
