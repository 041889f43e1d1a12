// ORACLE — TEST INFRASTRUCTURE ONLY.  Probes of the REAL vendored containers, compiled in place from
// /root/reference/thirdparty/{unordered_dense,nanoflann} by oracle/Makefile into oracle/_ref/libref_cont.so.
// Used to pin oracle/include/orc_dense_map.hpp (iteration order under insert/erase) and
// oracle/include/orc_kdtree.hpp (kNN indices, order and distances).
#include <cstdint>
#include <cstddef>
#include <vector>
#include "unordered_dense.h"
#include "nanoflann.hpp"

// ops: (op, key) pairs; op 0 = operator[] (insert if absent), 1 = erase.  Writes the final iteration order.
extern "C" size_t ref_dense_order(const int64_t* ops, size_t n_ops, uint64_t* out_keys) {
  ankerl::unordered_dense::map<uint64_t, int> m;
  for (size_t i = 0; i < n_ops; ++i) {
    uint64_t k = (uint64_t)ops[i * 2 + 1];
    if (ops[i * 2] == 0) m[k] += 1; else m.erase(k);
  }
  size_t j = 0;
  for (const auto& kv : m) out_keys[j++] = kv.first;
  return j;
}

struct Cloud {
  const float* p; size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return p[idx * 3 + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
extern "C" void ref_knn(const float* map_xyz, size_t nmap, const float* q, size_t m, int k, int* idx, float* d2, int* found) {
  Cloud c{map_xyz, nmap};
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Cloud>, Cloud, 3>;
  Tree tree(3, c, nanoflann::KDTreeSingleIndexAdaptorParams(10));
  tree.buildIndex();
  std::vector<uint32_t> ii(k); std::vector<float> dd(k);
  for (size_t i = 0; i < m; ++i) {
    size_t f = tree.knnSearch(q + i * 3, (size_t)k, ii.data(), dd.data());
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < (int)f ? (int)ii[j] : -1; d2[i * k + j] = j < (int)f ? dd[j] : 0.0f; }
    if (found) found[i] = (int)f;
  }
}
