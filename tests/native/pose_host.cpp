// Host-compiled harness around egonet_b200/csrc/pose_math.h (TEST ONLY).
// The header's functions are __host__ __device__; compiling them with g++ lets
// tests/test_native_host.py check the exact kernel source against the oracle
// and the golden vectors on a machine without a GPU.  Not part of the product.
#include "pose_math.h"

extern "C" {

void host_pose_solve(const double* kpts_3d, int N, int P, const double* kpts_2d, int stride_2d,
                     double fx, double cx, int alpha_mode, double* pose_out, double* rot_out) {
  for (int n = 0; n < N; ++n)
    egn::pose_solve_one(kpts_3d + (size_t)n * P * 3, P, kpts_2d ? kpts_2d[(size_t)n * stride_2d] : 0.0,
                        fx, cx, alpha_mode, pose_out + (size_t)n * 7,
                        rot_out ? rot_out + (size_t)n * 9 : nullptr);
}

void host_local_to_screen(const float* coords, const double* center, const double* scale,
                          const double* rot, int N, int K, int res_w, int res_h, double* screen) {
  for (int n = 0; n < N; ++n) {
    double M[6];
    egn::inverse_crop_affine(center[2 * n], center[2 * n + 1], scale[2 * n], rot ? rot[n] : 0.0,
                             res_w, res_h, M);
    for (int k = 0; k < K; ++k) {
      const int t = n * K + k;
      const float lx = (float)((double)coords[2 * t] * (double)res_w);
      const float ly = (float)((double)coords[2 * t + 1] * (double)res_h);
      screen[2 * t + 0] = M[0] * (double)lx + M[1] * (double)ly + M[2];
      screen[2 * t + 1] = M[3] * (double)lx + M[4] * (double)ly + M[5];
    }
  }
}

void host_kabsch(const double* H, double* R) {
  double h[3][3], r[3][3];
  for (int i = 0; i < 9; ++i) h[i / 3][i % 3] = H[i];
  egn::kabsch_rotation(h, r);
  for (int i = 0; i < 9; ++i) R[i] = r[i / 3][i % 3];
}

void host_rigid_transform(const double* X, const double* Y, const double* W, int w_mode, int N, int P, double* R,
                          double* t, double* aligned) {
  const size_t ws = w_mode == 2 ? (size_t)P * P : (size_t)P;
  for (int n = 0; n < N; ++n)
    egn::rigid_transform_one(X + (size_t)n * P * 3, Y + (size_t)n * P * 3, W ? W + n * ws : nullptr, w_mode, P,
                             R + (size_t)n * 9, t + (size_t)n * 3, aligned ? aligned + (size_t)n * P * 3 : nullptr);
}

void host_similarity_transform(const double* X, const double* Y, int N, int P, int scale, double* d, double* b,
                               double* Z, double* T, double* c) {
  for (int n = 0; n < N; ++n)
    egn::similarity_transform_one(X + (size_t)n * P * 3, Y + (size_t)n * P * 3, P, scale, d + n, b + n,
                                  Z + (size_t)n * P * 3, T + (size_t)n * 9, c + (size_t)n * 3);
}
}
