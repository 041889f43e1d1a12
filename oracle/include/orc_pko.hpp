// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_eigen.hpp header note).
//
// orc_pko.hpp — CPU restatement of optimization::AdaptiveMEstimator (PKO scale selection)
//   /root/reference/src/optimization/AdaptiveMEstimator.cpp:30-95 (ctor/reset), :128-156 (kernels),
//   :218-241 (initialize_pko), :243-291 (calculate_pko_scale_factor), :294-485 (fit_gmm),
//   :675-708 (gaussian_pdf, partition function), :710-787 (calculate_js_divergence)
// Pinned against the reference's own AdaptiveMEstimator.cpp (Eigen-free, compiled in place by
// oracle/Makefile into oracle/_ref/libref_pko.so) in tests/test_oracle_pins.py.
// libstdc++'s std::shuffle / uniform_int_distribution are used directly, as the reference does.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <random>
#include <string>
#include <vector>

namespace orc {

struct PkoConfig {
  bool use_adaptive_m_estimator = true;
  std::string loss_type = "huber";
  double min_scale_factor = 0.1;
  double max_scale_factor = 10.0;
  int num_alpha_segments = 100;
  double truncated_threshold = 10.0;
  int gmm_components = 3;
  int gmm_sample_size = 100;
  std::string pko_kernel_type = "huber";
};

class AdaptiveMEstimator {
 public:
  explicit AdaptiveMEstimator(const PkoConfig& c = PkoConfig()) : cfg(c) {}
  PkoConfig cfg;
  std::vector<double> alpha_candidates, partition_functions;
  std::vector<double> gmm_means, gmm_variances, gmm_weights;
  std::vector<double> last_sample;  // trace for parity tests
  int last_em_iters = 0, last_kmeans_iters = 0;

  void reset() {}  // :92-95 only resets two scalars that never feed the result

  double kernel(double r, double delta) const {  // :128-156
    const std::string& k = cfg.pko_kernel_type;
    if (k == "huber") { double a = std::abs(r); return a <= delta ? 1.0 : delta / a; }
    if (k == "tukey") { double a = std::abs(r); if (a < delta) { double x = a / delta, x2 = x * x; return (1 - x2) * (1 - x2); } return 0.0; }
    if (k == "welsch") { double e2 = r * r, d2 = delta * delta; return std::exp(-e2 / d2 / 2.0); }
    if (k == "gemanMcClure") { double e2 = r * r, d2 = delta * delta; return r * d2 / (d2 + e2) / (d2 + e2); }
    if (k == "pseudoHuber") { double d2 = delta * delta; return d2 / std::pow(d2 + r * r, 1.5); }
    double e2 = r * r, d2 = delta * delta;  // cauchy and default
    return d2 / (d2 + e2);
  }
  static double gaussian_pdf(double x, double mean, double variance) {  // :675-685
    if (variance <= 0.0) return 0.0;
    double diff = x - mean;
    double exponent = -0.5 * (diff * diff) / variance;
    double normalization = 1.0 / std::sqrt(2.0 * M_PI * variance);
    return normalization * std::exp(exponent);
  }
  double partition_function(double alpha) const {  // :692-708
    const double bound = cfg.truncated_threshold, step = 0.01;
    double integral = 0.0;
    for (double x = 0.0; x <= bound; x += step) integral += kernel(x, alpha) * step;
    return std::max(integral, 1e-10);
  }
  void initialize_pko() {  // :218-241
    int S = cfg.num_alpha_segments;
    alpha_candidates.assign(S + 1, 0.0);
    partition_functions.assign(S + 1, 0.0);
    alpha_candidates[0] = cfg.min_scale_factor;
    partition_functions[0] = partition_function(cfg.min_scale_factor);
    for (int i = 1; i <= S; ++i) {
      double t = (double)i / (double)S;
      double ls = (std::pow(100.0, t) - 1.0) / 99.0;
      double alpha = cfg.min_scale_factor + (cfg.max_scale_factor - cfg.min_scale_factor) * ls;
      alpha_candidates[i] = alpha;
      partition_functions[i] = partition_function(alpha);
    }
  }

  void fit_gmm(const std::vector<double>& residuals) {  // :294-485
    if (residuals.empty()) return;
    int n = (int)residuals.size();
    int sample_size;
    if (cfg.gmm_sample_size > 0) sample_size = cfg.gmm_sample_size;
    else { sample_size = std::max(100, (int)(n * 0.1)); sample_size = std::min(sample_size, 10000); }
    if (sample_size > n) sample_size = n;
    std::vector<int> indices(n);
    std::iota(indices.begin(), indices.end(), 0);
    std::mt19937 g(42);
    std::shuffle(indices.begin(), indices.end(), g);
    std::vector<double> s(sample_size);
    for (int i = 0; i < sample_size; ++i) s[i] = residuals[indices[i]];
    last_sample = s;
    n = sample_size;
    const int K = cfg.gmm_components;
    std::mt19937 gen(42);
    std::uniform_int_distribution<> dis(0, (int)s.size() - 1);
    gmm_means.resize(K);
    gmm_means[0] = 0.0;
    for (int i = 1; i < K; ++i) gmm_means[i] = s[dis(gen)];
    std::vector<int> clusters(s.size());
    std::vector<double> new_means(K);
    last_kmeans_iters = 0;
    while (true) {
      ++last_kmeans_iters;
      for (size_t i = 0; i < s.size(); ++i) {
        double min_dist = std::numeric_limits<double>::max();
        int ci = 0;
        for (int j = 0; j < K; ++j) { double d = std::abs(s[i] - gmm_means[j]); if (d < min_dist) { min_dist = d; ci = j; } }
        clusters[i] = ci;
      }
      std::fill(new_means.begin(), new_means.end(), 0.0);
      std::vector<int> counts(K, 0);
      for (size_t i = 0; i < s.size(); ++i) { new_means[clusters[i]] += s[i]; counts[clusters[i]]++; }
      for (int j = 0; j < K; ++j) { if (j == 0) new_means[j] = 0.0; else if (counts[j] > 0) new_means[j] /= (double)counts[j]; }
      if (gmm_means == new_means) break;
      new_means[0] = 0.0;
      gmm_means = new_means;
    }
    double mean_of_data = std::accumulate(s.begin(), s.end(), 0.0) / s.size();
    double initial_variance = 0.0;
    for (double x : s) initial_variance += std::pow(x - mean_of_data, 2);
    initial_variance /= s.size();
    gmm_variances.assign(K, initial_variance);
    std::vector<int> cc(K, 0);
    for (size_t i = 0; i < s.size(); ++i) cc[clusters[i]]++;
    gmm_weights.resize(K);
    for (int j = 0; j < K; ++j) gmm_weights[j] = (double)cc[j] / (double)s.size();
    const int max_iterations = 100;
    const double convergence_threshold = 1e-6;
    std::vector<std::vector<double>> resp(n, std::vector<double>(K));
    last_em_iters = 0;
    for (int iter = 0; iter < max_iterations; ++iter) {
      ++last_em_iters;
      std::vector<double> sum_resp(n, 0.0);
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < K; ++j) { resp[i][j] = gmm_weights[j] * gaussian_pdf(s[i], gmm_means[j], gmm_variances[j]); sum_resp[i] += resp[i][j]; }
        for (int j = 0; j < K; ++j) resp[i][j] /= sum_resp[i];
      }
      std::vector<double> Nk(K, 0.0);
      for (int j = 0; j < K; ++j) for (int i = 0; i < n; ++i) Nk[j] += resp[i][j];
      std::vector<double> nw(K), nm(K, 0.0), nv(K, 0.0);
      for (int j = 0; j < K; ++j) {
        nw[j] = Nk[j] / (double)n;
        if (j == 0) nm[j] = 0.0;
        else { for (int i = 0; i < n; ++i) nm[j] += resp[i][j] * s[i]; nm[j] /= Nk[j]; }
        for (int i = 0; i < n; ++i) { double diff = s[i] - nm[j]; nv[j] += resp[i][j] * diff * diff; }
        nv[j] /= Nk[j];
        nv[j] = std::max(nv[j], 1e-6);
      }
      double change = 0.0;
      for (int j = 1; j < K; ++j) change += std::abs(nm[j] - gmm_means[j]);
      gmm_weights = nw;
      nm[0] = 0.0;
      gmm_means = nm;
      gmm_variances = nv;
      if (change < convergence_threshold) break;
    }
  }

  double js_divergence(double alpha) const {  // :710-787
    int num_segments = 100;
    double dr = cfg.truncated_threshold / (double)num_segments;
    double pf = 0.0;
    for (size_t j = 0; j < alpha_candidates.size(); ++j)
      if (std::abs(alpha_candidates[j] - alpha) < 1e-10) { pf = partition_functions[j]; break; }
    if (pf == 0.0) pf = partition_function(alpha);
    if (pf < 1e-10) return std::numeric_limits<double>::max();
    double cost = 0.0, cnt = 0.0;
    for (int i = 0; i < num_segments; ++i) {
      double r = dr * (1 + (double)i);
      double Pr = 0.0;
      if (!gmm_weights.empty() && !gmm_means.empty() && !gmm_variances.empty())
        for (int m = 0; m < cfg.gmm_components && m < (int)gmm_weights.size(); ++m) Pr += gmm_weights[m] * gaussian_pdf(r, gmm_means[m], gmm_variances[m]);
      Pr += 1e-10;
      double Q = kernel(r, alpha) / (pf + 1e-10) + 1e-10;
      double M = 0.5 * (Pr + Q);
      double jsd = 0.5 * (Pr * std::log(Pr / M) + Q * std::log(Q / M));
      if (std::isnan(jsd)) continue;
      cost += jsd;
      cnt += 1.0;
    }
    if (cnt == 0) return std::numeric_limits<double>::max();
    return cost / cnt;
  }

  double calculate_scale_factor(const std::vector<double>& residuals) {  // :63-79 -> :243-291
    if (residuals.empty()) return 1.0;
    if (alpha_candidates.empty()) initialize_pko();
    fit_gmm(residuals);
    double best_alpha = cfg.min_scale_factor;
    double best_cost = std::numeric_limits<double>::max();
    last_js.assign(alpha_candidates.size(), 0.0);
    for (size_t i = 1; i < alpha_candidates.size(); ++i) {
      double js = js_divergence(alpha_candidates[i]);
      last_js[i] = js;
      if (js < best_cost) { best_cost = js; best_alpha = alpha_candidates[i]; }
    }
    return best_alpha;  // log_residual_histogram (:795-914) has no effect on results; skipped
  }
  std::vector<double> last_js;
};

}  // namespace orc
