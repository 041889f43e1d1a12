// b2lo_pko_host.cpp — host-side constant tables of the device PKO kernel (b2lo_icp.cu, k_icp_pko).
//
// optimization::AdaptiveMEstimator draws its GMM sample with libstdc++:
//     std::iota(idx); std::shuffle(idx.begin(), idx.end(), std::mt19937(42)); sample = residuals[idx[0..100)]
// (/root/reference/src/optimization/AdaptiveMEstimator.cpp:319-328) and its k-means seeds with
//     std::uniform_int_distribution<>(0, ns-1)(std::mt19937(42))            (:336-345).
// Both are pure functions of n (the correspondence count).  std::shuffle is a forward Fisher-Yates,
//     for i = 1..n-1: swap(a[i], a[r_i]),  r_i uniform in [0, i]
// whose r_i sequence depends on n only through (a) the parity of n (libstdc++ draws two positions per
// engine call and does the odd one out first) and (b) whether n*n fits the engine range (n <= 65535).
// So three r-sequences (even n, odd n, n >= 65536) describe every n.  The first 100 entries of the
// shuffled array then follow from a backward trace: element j < 100 is the LARGEST i >= 100 with
// r_i == j if there is one, else a trace through the first 99 swaps.  We record r_i by running the
// REAL std::shuffle over a token type whose ADL swap logs the two positions, so the tables are
// libstdc++'s by construction (the reference links the same library), and keep only
//     head_r[mode][i]  (i < 128)   and   hits[mode][j] = ascending list of i >= HEAD with r_i == j.
// About 100 * ln(N/100) entries per mode instead of an O(n) host shuffle per Gauss-Newton iteration.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>
#include "b2lo_internal.h"

namespace b2 {

constexpr int HEAD = 128;                 // head positions tracked (gmm_sample_size <= 128)
constexpr int N_PAIRED_MAX = 65535;       // largest n using the two-draws-per-call path
constexpr int N_LARGE_MAX = 1 << 22;      // largest n covered by the tables

namespace {
struct Tok { int v; };
struct SwapLog { Tok* base; std::vector<std::pair<int, int>>* log; };
thread_local SwapLog g_log{nullptr, nullptr};
inline void swap(Tok& a, Tok& b) {  // found by ADL from std::iter_swap
  if (g_log.log) g_log.log->emplace_back((int)(&a - g_log.base), (int)(&b - g_log.base));
  Tok t = a; a = b; b = t;
}
// r[i] for i in [1, n): the partner position of element i in std::shuffle(iota(n), mt19937(42))
std::vector<int> shuffle_partners(int n) {
  std::vector<Tok> v(n);
  for (int i = 0; i < n; ++i) v[i].v = i;
  std::vector<std::pair<int, int>> log;
  log.reserve(n);
  g_log = SwapLog{v.data(), &log};
  std::mt19937 g(42);
  std::shuffle(v.begin(), v.end(), g);
  g_log = SwapLog{nullptr, nullptr};
  std::vector<int> r(n, -1);
  r[0] = 0;
  for (auto& pr : log) {
    int hi = std::max(pr.first, pr.second), lo = std::min(pr.first, pr.second);
    r[hi] = lo;  // iter_swap(i, first + d) with d <= i
  }
  // positions that swapped with themselves are logged as (i,i) too; any -1 left would be a libstdc++ that
  // skips self-swaps: treat as self.
  for (int i = 1; i < n; ++i) if (r[i] < 0) r[i] = i;
  return r;
}
double kernel_w(int type, double r, double delta) {  // AdaptiveMEstimator::pko_kernel_weight (:128-156), huber / cauchy
  if (type == 0) { double a = std::abs(r); return a <= delta ? 1.0 : delta / a; }
  double e2 = r * r, d2 = delta * delta;
  return d2 / (d2 + e2);
}
}  // namespace

void pko_build_host(const b2lo_icp_cfg* cfg, PkoTables* t, std::vector<int>* hits) {
  std::memset(t, 0, sizeof(*t));
  t->min_sf = cfg->min_scale_factor; t->max_sf = cfg->max_scale_factor; t->trunc = cfg->truncated_threshold;
  t->sample_size = cfg->gmm_sample_size; t->kernel_type = cfg->pko_kernel_type;
  int S = cfg->num_alpha_segments;
  t->n_alpha = S + 1;
  auto Zf = [&](double alpha) {  // calculate_partition_function_integration (:692-708)
    double integral = 0.0;
    for (double x = 0.0; x <= cfg->truncated_threshold; x += 0.01) integral += kernel_w(cfg->pko_kernel_type, x, alpha) * 0.01;
    return std::max(integral, 1e-10);
  };
  for (int j = 0; j < 32; ++j) { t->exp2_t1[j] = std::exp2((double)j / 32.0); t->exp2_t2[j] = std::exp2((double)j / 1024.0); }
  t->alpha[0] = cfg->min_scale_factor;
  t->Z[0] = Zf(t->alpha[0]);
  for (int i = 1; i <= S; ++i) {  // initialize_pko (:218-241)
    double tt = (double)i / (double)S;
    double ls = (std::pow(100.0, tt) - 1.0) / 99.0;
    double alpha = cfg->min_scale_factor + (cfg->max_scale_factor - cfg->min_scale_factor) * ls;
    t->alpha[i] = alpha;
    t->Z[i] = Zf(alpha);
  }
  for (int ns = 1; ns <= 128; ++ns) {  // k-means seeds (:336-345): two draws from one generator
    std::mt19937 gen(42);
    std::uniform_int_distribution<> dis(0, ns - 1);
    t->kmeans_seed[ns][0] = dis(gen);
    t->kmeans_seed[ns][1] = dis(gen);
  }
  hits->clear();
  const int sizes[3] = {N_PAIRED_MAX - 1, N_PAIRED_MAX, N_LARGE_MAX};
  for (int mode = 0; mode < 3; ++mode) {
    std::vector<int> r = shuffle_partners(sizes[mode]);
    for (int i = 0; i < HEAD; ++i) t->head_r[mode][i] = r[i];
    for (int j = 0; j < HEAD; ++j) {   // backward trace of position j through swaps HEAD-1 .. 1
      int pos = j;
      for (int i = HEAD - 1; i >= 1; --i) { int ri = r[i]; pos = (pos == i) ? ri : ((pos == ri) ? i : pos); }
      t->head_pos[mode][j] = pos;
    }
    std::vector<std::vector<int>> per(HEAD);
    for (int i = HEAD; i < sizes[mode]; ++i) if (r[i] < HEAD) per[r[i]].push_back(i);
    for (int j = 0; j < HEAD; ++j) {
      t->hit_off[mode][j] = (int)hits->size();
      hits->insert(hits->end(), per[j].begin(), per[j].end());
    }
    t->hit_off[mode][HEAD] = (int)hits->size();
  }
  t->hit_n = (int)hits->size();
}

}  // namespace b2
