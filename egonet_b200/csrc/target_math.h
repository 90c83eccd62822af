// Gaussian heat-map targets, arithmetic shared by the CUDA kernel (loss.cu) and the host harness
// (tests/native/target_host.cpp).
// upstream: generate_target libs/common/img_proc.py:347-409 (target_type 'gaussian').
//   target = zeros((K, heatmap_size[0], heatmap_size[1]))  -- rows = heatmap_size[0], cols = heatmap_size[1]
//   feat_stride = input_size / heatmap_size;  mu = int(joint / feat_stride + 0.5)   (int() truncates towards 0)
//   ul = int(mu - 3 sigma), br = int(mu + 3 sigma + 1); a dot entirely outside the map clears the joint's weight
//   g = exp(-((x - x0)^2 + (y - y0)^2) / (2 sigma^2)) in float32, x0 = (2 * 3 sigma + 1) // 2, copied where it overlaps
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define EGN_THD __host__ __device__ __forceinline__
#else
#define EGN_THD inline
#endif

namespace egn {

struct TargetDot {
  int ul_x, ul_y, br_x, br_y;   // patch corners on the map (br exclusive)
  float x0;                     // patch centre inside the patch (same for x and y)
  bool visible;                 // the dot is drawn
  float weight;                 // target_weight of the joint: joints_vis, cleared when a visible dot misses the map
};

// hs0 / hs1: heatmap_size[0] / [1] (rows / columns of the target as upstream allocates it); in0 / in1: input_size.
EGN_THD TargetDot target_dot(double jx, double jy, float vis, double in0, double in1, int hs0, int hs1, double sigma) {
  TargetDot d;
  d.visible = vis > 0.5f;
  const double tmp = sigma * 3.0;
  const int mu_x = (int)(jx / (in0 / (double)hs0) + 0.5);
  const int mu_y = (int)(jy / (in1 / (double)hs1) + 0.5);
  d.ul_x = (int)((double)mu_x - tmp);
  d.ul_y = (int)((double)mu_y - tmp);
  d.br_x = (int)((double)mu_x + tmp + 1.0);
  d.br_y = (int)((double)mu_y + tmp + 1.0);
  d.weight = vis;
  if (d.visible && (d.ul_x >= hs1 || d.ul_y >= hs0 || d.br_x < 0 || d.br_y < 0)) {
    d.visible = false;
    d.weight = 0.f;
  }
  d.x0 = (float)floor((2.0 * tmp + 1.0) / 2.0);
  return d;
}

// value of map pixel (row, col); 0 outside the dot
EGN_THD float target_value(const TargetDot& d, int row, int col, int hs0, int hs1, double sigma) {
  if (!d.visible) return 0.f;
  const int x_lo = d.ul_x > 0 ? d.ul_x : 0, x_hi = d.br_x < hs1 ? d.br_x : hs1;
  const int y_lo = d.ul_y > 0 ? d.ul_y : 0, y_hi = d.br_y < hs0 ? d.br_y : hs0;
  if (col < x_lo || col >= x_hi || row < y_lo || row >= y_hi) return 0.f;
  const float dx = (float)(col - d.ul_x) - d.x0, dy = (float)(row - d.ul_y) - d.x0;
  const float q = (dx * dx + dy * dy) / (float)(2.0 * sigma * sigma);
  return expf(-q);
}

}  // namespace egn
