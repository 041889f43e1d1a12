// ORACLE — TEST INFRASTRUCTURE ONLY.  C wrapper around the REAL reference class
// lidar_slam::optimization::AdaptiveMEstimator, compiled in place from
// /root/reference/src/optimization/AdaptiveMEstimator.cpp (Eigen-free) by oracle/Makefile into
// oracle/_ref/libref_pko.so.  Used to pin oracle/include/orc_pko.hpp and the CUDA PKO kernel.
#include "optimization/AdaptiveMEstimator.h"
#include <vector>
#include <string>

extern "C" double ref_pko_scale(const double* residuals, size_t n, int use_adaptive, int loss_type, double min_sf, double max_sf,
                                int segments, double trunc, int comps, int sample_size, int kernel_type) {
  using lidar_slam::optimization::AdaptiveMEstimator;
  AdaptiveMEstimator a(use_adaptive != 0, loss_type == 1 ? "cauchy" : "huber", min_sf, max_sf, segments, trunc, comps, sample_size,
                       kernel_type == 1 ? "cauchy" : "huber");
  a.reset();
  std::vector<double> r(residuals, residuals + n);
  return a.calculate_scale_factor(r);
}
