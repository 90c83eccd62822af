// fp64 geometry kernels around the networks: local->screen affine and the
// per-instance pose solve (template cuboid -> Kabsch -> Euler -> alpha).
//
// Both are latency-trivial (a few kFLOP per instance); they exist so that the
// whole per-crop path stays on the device between the HC forward and the single
// D2H copy of the [N,7] pose records.  One thread per (instance[, key-point]).
#include "common.h"
#include "pose_math.h"

namespace egn {

__global__ void local_to_screen_kernel(const float* __restrict__ coords,
                                       const double* __restrict__ center,
                                       const double* __restrict__ scale,
                                       const double* __restrict__ rot, int N, int K, int res_w,
                                       int res_h, double* __restrict__ screen) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * K) return;
  const int n = t / K;
  double M[6];
  inverse_crop_affine(center[2 * n], center[2 * n + 1], scale[2 * n], rot ? rot[n] : 0.0, res_w,
                      res_h, M);
  // local_coord *= resolution: float32 array times int array, result stored as float32
  const float lx = (float)((double)coords[2 * t] * (double)res_w);
  const float ly = (float)((double)coords[2 * t + 1] * (double)res_h);
  screen[2 * t + 0] = M[0] * (double)lx + M[1] * (double)ly + M[2];
  screen[2 * t + 1] = M[3] * (double)lx + M[4] * (double)ly + M[5];
}

__global__ void pose_solve_kernel(const double* __restrict__ kpts_3d, int N, int P,
                                  const double* __restrict__ kpts_2d, int stride_2d, double fx,
                                  double cx, int alpha_mode, double* __restrict__ pose_out,
                                  double* __restrict__ rot_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double x0 = kpts_2d ? kpts_2d[(size_t)n * stride_2d] : 0.0;
  pose_solve_one(kpts_3d + (size_t)n * P * 3, P, x0, fx, cx, alpha_mode, pose_out + (size_t)n * 7,
                 rot_out ? rot_out + (size_t)n * 9 : nullptr);
}

// alpha = ry - atan2(-z, x) - pi/2 wrapped into [-pi, pi] (egonet.py:203-236)
__global__ void observation_angle_kernel(const double* __restrict__ ry, const double* __restrict__ x3d,
                                         const double* __restrict__ z3d, int stride_x, int stride_z,
                                         double x_offset, int N, double* __restrict__ alpha) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  alpha[n] = observation_angle(ry[n], x3d[(size_t)n * stride_x] - x_offset, z3d[(size_t)n * stride_z]);
}

// one thread per instance: compute_rigid_transform / procrustes_transform (transformation.py:99-141)
__global__ void rigid_transform_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                       const double* __restrict__ W, int w_mode, int N, int P,
                                       double* __restrict__ R, double* __restrict__ t, double* __restrict__ aligned) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t wstride = w_mode == 2 ? (size_t)P * P : (size_t)P;
  rigid_transform_one(X + (size_t)n * P * 3, Y + (size_t)n * P * 3, W ? W + n * wstride : nullptr, w_mode, P,
                      R ? R + (size_t)n * 9 : nullptr, t ? t + (size_t)n * 3 : nullptr,
                      aligned ? aligned + (size_t)n * P * 3 : nullptr);
}

// one thread per instance: compute_similarity_transform (transformation.py:48-97)
__global__ void similarity_transform_kernel(const double* __restrict__ X, const double* __restrict__ Y, int N, int P,
                                            int optimal_scale, double* __restrict__ d, double* __restrict__ b,
                                            double* __restrict__ Z, double* __restrict__ T, double* __restrict__ c) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  similarity_transform_one(X + (size_t)n * P * 3, Y + (size_t)n * P * 3, P, optimal_scale, d ? d + n : nullptr,
                           b ? b + n : nullptr, Z ? Z + (size_t)n * P * 3 : nullptr, T ? T + (size_t)n * 9 : nullptr,
                           c ? c + (size_t)n * 3 : nullptr);
}

}  // namespace egn

extern "C" {

int egn_rigid_transform(const double* X, const double* Y, const double* W, int w_mode, int N, int P, double* R_out,
                        double* t_out, double* aligned_out, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0 && P >= 1, "egn_rigid_transform: bad shape");
  EGN_REQUIRE(w_mode >= 0 && w_mode <= 2 && (w_mode == 0) == (W == nullptr), "egn_rigid_transform: w_mode 0 (W null), 1 ([N,P]) or 2 ([N,P,P])");
  EGN_REQUIRE(N == 0 || (X && Y), "egn_rigid_transform: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  rigid_transform_kernel<<<ceil_div(N, 64), 64, 0, as_stream(stream)>>>(X, Y, W, w_mode, N, P, R_out, t_out, aligned_out);
  EGN_LAUNCH_CHECK("rigid_transform_kernel");
  return EGN_OK;
}

int egn_similarity_transform(const double* X, const double* Y, int N, int P, int compute_optimal_scale, double* d_out,
                             double* b_out, double* Z_out, double* T_out, double* c_out, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0 && P >= 1, "egn_similarity_transform: bad shape");
  EGN_REQUIRE(N == 0 || (X && Y), "egn_similarity_transform: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  similarity_transform_kernel<<<ceil_div(N, 64), 64, 0, as_stream(stream)>>>(X, Y, N, P, compute_optimal_scale ? 1 : 0,
                                                                             d_out, b_out, Z_out, T_out, c_out);
  EGN_LAUNCH_CHECK("similarity_transform_kernel");
  return EGN_OK;
}

int egn_observation_angle(const double* ry, const double* x3d, int stride_x, const double* z3d,
                          int stride_z, double x_offset, int N, double* alpha, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0 && stride_x >= 0 && stride_z >= 0, "egn_observation_angle: bad shape");
  EGN_REQUIRE(N == 0 || (ry && x3d && z3d && alpha), "egn_observation_angle: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  observation_angle_kernel<<<ceil_div(N, 128), 128, 0, as_stream(stream)>>>(ry, x3d, z3d, stride_x, stride_z,
                                                                             x_offset, N, alpha);
  EGN_LAUNCH_CHECK("observation_angle_kernel");
  return EGN_OK;
}

int egn_local_to_screen(const float* coords, const double* center, const double* scale,
                        const double* rot, int N, int K, int res_w, int res_h, double* screen,
                        void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0 && K > 0 && res_w > 0 && res_h > 0, "egn_local_to_screen: bad shape");
  EGN_REQUIRE(N == 0 || (coords && center && scale && screen), "egn_local_to_screen: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  const int threads = 128;
  local_to_screen_kernel<<<ceil_div(N * K, threads), threads, 0, as_stream(stream)>>>(
      coords, center, scale, rot, N, K, res_w, res_h, screen);
  EGN_LAUNCH_CHECK("local_to_screen_kernel");
  return EGN_OK;
}

int egn_pose_solve(const double* kpts_3d, int N, int P, const double* kpts_2d, int stride_2d,
                   double fx, double cx, int alpha_mode, double* pose_out, double* rot_out,
                   void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0, "egn_pose_solve: negative N");
  EGN_REQUIRE(N == 0 || (kpts_3d && pose_out), "egn_pose_solve: null pointer");
  EGN_REQUIRE(P == 8 || P == 32, "egn_pose_solve: P must be 8 or 32 (got %d)", P);
  EGN_REQUIRE(alpha_mode == EGN_ALPHA_TRANS || alpha_mode == EGN_ALPHA_PROJ,
              "egn_pose_solve: unknown alpha_mode %d", alpha_mode);
  EGN_REQUIRE(alpha_mode != EGN_ALPHA_PROJ || N == 0 || (kpts_2d && stride_2d > 0),
              "egn_pose_solve: proj mode needs kpts_2d");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  const int threads = 64;
  pose_solve_kernel<<<ceil_div(N, threads), threads, 0, as_stream(stream)>>>(
      kpts_3d, N, P, kpts_2d, stride_2d, fx, cx, alpha_mode, pose_out, rot_out);
  EGN_LAUNCH_CHECK("pose_solve_kernel");
  return EGN_OK;
}

}  // extern "C"
