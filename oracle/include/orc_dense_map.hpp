// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_dense_map.hpp — restatement of the ORDERING semantics of ankerl::unordered_dense 4.8.1
// (/root/reference/thirdparty/unordered_dense/unordered_dense.h):
//   * values live in a vector in insertion order; iteration walks that vector  (:1448-1453)
//   * operator[] / try_emplace / insert append at the back                      (:1220-1242)
//   * erase(key) moves the LAST element into the erased slot                    (:1147-1167)
//   * clear() empties the vector                                                (:1492-1495)
// Bucket placement (wyhash mixing, Robin-Hood distances, :918,:980-1002) never influences iteration
// order and is not part of the contract; the index here is a plain linear-probing table.
// Cross-checked against the vendored header by oracle/ref_check (tests/test_oracle_pins.py).
#pragma once
#include <cstdint>
#include <cstddef>
#include <utility>
#include <vector>

namespace orc {

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

template <class K, class V, class Hash>
class DenseMap {
 public:
  using value_type = std::pair<K, V>;
  std::vector<value_type> values;  // insertion order (the contract)

  DenseMap() { rebuild(16); }
  size_t size() const { return values.size(); }
  bool empty() const { return values.empty(); }
  void clear() { values.clear(); std::fill(tab_.begin(), tab_.end(), kEmpty); }
  void reserve(size_t n) { values.reserve(n); if (n * 2 > tab_.size()) { size_t c = tab_.size(); while (c < n * 2) c <<= 1; rebuild(c); } }

  // returns index into values or -1
  int64_t find(const K& k) const {
    size_t mask = tab_.size() - 1;
    size_t h = mix64(Hash{}(k)) & mask;
    for (;;) {
      uint32_t e = tab_[h];
      if (e == kEmpty) return -1;
      if (values[e].first == k) return (int64_t)e;
      h = (h + 1) & mask;
    }
  }
  bool contains(const K& k) const { return find(k) >= 0; }

  // operator[] semantics: find or append default-constructed value; returns (index, inserted)
  std::pair<size_t, bool> try_emplace(const K& k) {
    int64_t f = find(k);
    if (f >= 0) return {(size_t)f, false};
    if ((values.size() + 1) * 2 > tab_.size()) rebuild(tab_.size() * 2);
    values.emplace_back(k, V{});
    place((uint32_t)(values.size() - 1));
    return {values.size() - 1, true};
  }
  V& operator[](const K& k) { return values[try_emplace(k).first].second; }

  // erase(key): swap-with-last.  returns 1 if erased.
  size_t erase(const K& k) {
    int64_t f = find(k);
    if (f < 0) return 0;
    erase_at((size_t)f);
    return 1;
  }
  void erase_at(size_t idx) {
    remove_from_table(values[idx].first);
    size_t last = values.size() - 1;
    if (idx != last) {
      // re-point the moved element's bucket
      size_t mask = tab_.size() - 1;
      size_t h = mix64(Hash{}(values[last].first)) & mask;
      while (tab_[h] != (uint32_t)last) h = (h + 1) & mask;
      tab_[h] = (uint32_t)idx;
      values[idx] = std::move(values[last]);
    }
    values.pop_back();
  }

 private:
  static constexpr uint32_t kEmpty = 0xffffffffu;
  std::vector<uint32_t> tab_;

  void rebuild(size_t cap) {
    tab_.assign(cap, kEmpty);
    for (uint32_t i = 0; i < values.size(); ++i) place(i);
  }
  void place(uint32_t i) {
    size_t mask = tab_.size() - 1;
    size_t h = mix64(Hash{}(values[i].first)) & mask;
    while (tab_[h] != kEmpty) h = (h + 1) & mask;
    tab_[h] = i;
  }
  // backward-shift deletion for linear probing
  void remove_from_table(const K& k) {
    size_t mask = tab_.size() - 1;
    size_t h = mix64(Hash{}(k)) & mask;
    while (!(values[tab_[h]].first == k)) h = (h + 1) & mask;
    size_t hole = h;
    size_t j = (hole + 1) & mask;
    while (tab_[j] != kEmpty) {
      size_t home = mix64(Hash{}(values[tab_[j]].first)) & mask;
      // can element at j move to hole?  yes iff home is cyclically outside (hole, j]
      bool movable = ((j > hole) ? (home <= hole || home > j) : (home <= hole && home > j));
      if (movable) { tab_[hole] = tab_[j]; hole = j; }
      j = (j + 1) & mask;
    }
    tab_[hole] = kEmpty;
  }
};

struct Empty {};
template <class K, class Hash>
class DenseSet {
 public:
  DenseMap<K, Empty, Hash> m;
  size_t size() const { return m.size(); }
  bool empty() const { return m.empty(); }
  bool insert(const K& k) { return m.try_emplace(k).second; }
  size_t erase(const K& k) { return m.erase(k); }
  const K& at(size_t i) const { return m.values[i].first; }
};

}  // namespace orc
