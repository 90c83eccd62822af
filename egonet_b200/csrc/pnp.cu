// Batched reprojection refinement: one thread per instance runs the whole of
// cv2.solvePnP(SOLVEPNP_ITERATIVE) + Rodrigues + transform (pnp_math.h) in fp64.
// upstream: pnp_refine libs/common/transformation.py:143-157 (one cv2 call per instance on the host;
// callers: tools/inference_legacy.py:518-547, libs/trainer/trainer.py:355-381).
// Latency-bound (a 12x12 and a few 6x6 Jacobi eigen-decompositions per instance, ~0.3 MFLOP);
// it keeps the refined boxes on the device next to the pose records.
#include "common.h"
#include "pnp_math.h"

namespace egn {

__global__ void __launch_bounds__(64) pnp_refine_kernel(const double* __restrict__ kpts_3d,
                                                        const double* __restrict__ kpts_2d, int N, int P,
                                                        PnpCamera cam, int max_iter, double eps,
                                                        double* __restrict__ refined, double* __restrict__ pose6,
                                                        double* __restrict__ info, int32_t* __restrict__ status) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int st = pnp_refine_one(kpts_3d + (size_t)n * P * 3, kpts_2d + (size_t)n * P * 2, P, cam, max_iter, eps,
                                refined + (size_t)n * P * 3, pose6 ? pose6 + (size_t)n * 6 : nullptr,
                                info ? info + (size_t)n * 2 : nullptr);
  if (status) status[n] = st;
}

// refine_with_predicted_bbox (tools/inference_legacy.py:518-547): the prediction's points 1.. are relative to
// point 0; make them absolute, refine by pnp_refine, discard the result when the refined root moved further
// than `threshold` from the predicted one.  ok[n] = 1 kept / 0 discarded (refined then holds the absolute,
// unrefined box).
__global__ void __launch_bounds__(64) refine_bbox_kernel(const double* __restrict__ pred_rel,
                                                         const double* __restrict__ kpts_2d, int N, int P,
                                                         PnpCamera cam, int max_iter, double eps, double threshold,
                                                         double* __restrict__ refined, int32_t* __restrict__ ok,
                                                         int32_t* __restrict__ status) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double* pr = pred_rel + (size_t)n * P * 3;
  double box[kPnpMaxPoints * 3];
  for (int i = 0; i < P; ++i)
    for (int a = 0; a < 3; ++a) box[3 * i + a] = pr[3 * i + a] + (i > 0 ? pr[a] : 0.0);
  double* out = refined + (size_t)n * P * 3;
  const int st = pnp_refine_one(box, kpts_2d + (size_t)n * P * 2, P, cam, max_iter, eps, out, nullptr, nullptr);
  const double dx = out[0] - box[0], dy = out[1] - box[1], dz = out[2] - box[2];
  const bool keep = !(sqrt(dx * dx + dy * dy + dz * dz) > threshold);
  if (!keep)
    for (int i = 0; i < 3 * P; ++i) out[i] = box[i];
  ok[n] = keep ? 1 : 0;
  if (status) status[n] = st;
}

}  // namespace egn

extern "C" int egn_refine_with_bbox(const double* pred_rel, const double* kpts_2d, int N, int P, double fx, double fy,
                                    double cx, double cy, double threshold, int max_iter, double* refined, int32_t* ok,
                                    int32_t* status, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0, "egn_refine_with_bbox: negative N");
  EGN_REQUIRE(P >= 6 && P <= kPnpMaxPoints, "egn_refine_with_bbox: P must be in [6, %d] (got %d)", kPnpMaxPoints, P);
  EGN_REQUIRE(fx != 0.0 && fy != 0.0 && max_iter >= 0, "egn_refine_with_bbox: bad camera / max_iter");
  EGN_REQUIRE(N == 0 || (pred_rel && kpts_2d && refined && ok), "egn_refine_with_bbox: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  PnpCamera cam{fx, fy, cx, cy};
  refine_bbox_kernel<<<ceil_div(N, 64), 64, 0, as_stream(stream)>>>(pred_rel, kpts_2d, N, P, cam,
                                                                    max_iter > 0 ? max_iter : 20, 1.1920928955078125e-07,
                                                                    threshold, refined, ok, status);
  EGN_LAUNCH_CHECK("refine_bbox_kernel");
  return EGN_OK;
}

extern "C" int egn_pnp_refine(const double* kpts_3d, const double* kpts_2d, int N, int P, double fx, double fy,
                              double cx, double cy, int max_iter, double* refined, double* pose6, double* info,
                              int32_t* status, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0, "egn_pnp_refine: negative N");
  EGN_REQUIRE(P >= 6 && P <= kPnpMaxPoints, "egn_pnp_refine: P must be in [6, %d] (got %d; the DLT needs 6 points)",
              kPnpMaxPoints, P);
  EGN_REQUIRE(fx != 0.0 && fy != 0.0, "egn_pnp_refine: zero focal length");
  EGN_REQUIRE(max_iter >= 0, "egn_pnp_refine: negative max_iter");
  EGN_REQUIRE(N == 0 || (kpts_3d && kpts_2d && refined), "egn_pnp_refine: null pointer");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  PnpCamera cam{fx, fy, cx, cy};
  const int threads = 64;
  pnp_refine_kernel<<<ceil_div(N, threads), threads, 0, as_stream(stream)>>>(
      kpts_3d, kpts_2d, N, P, cam, max_iter > 0 ? max_iter : 20, 1.1920928955078125e-07, refined, pose6, info, status);
  EGN_LAUNCH_CHECK("pnp_refine_kernel");
  return EGN_OK;
}
