// Host-compiled harness around egonet_b200/csrc/target_math.h (TEST ONLY).
#include "target_math.h"

extern "C" void host_generate_target(const double* joints, const float* vis, int N, int K, double in0, double in1,
                                     int hs0, int hs1, double sigma, float* target, float* weight) {
  for (int m = 0; m < N * K; ++m) {
    const egn::TargetDot d = egn::target_dot(joints[3 * m], joints[3 * m + 1], vis[m], in0, in1, hs0, hs1, sigma);
    weight[m] = d.weight;
    for (int r = 0; r < hs0; ++r)
      for (int c = 0; c < hs1; ++c) target[((size_t)m * hs0 + r) * hs1 + c] = egn::target_value(d, r, c, hs0, hs1, sigma);
  }
}
