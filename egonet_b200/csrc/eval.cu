// Batched overlaps of the KITTI object evaluation: every (detection, ground truth) pair of a
// frame in one launch, fp64, one thread per pair (latency trivial: ~200 flops per pair).
// upstream: tools/kitti-eval/evaluate_object_3d_offline.cpp:224-344 (see eval_math.h).
#include "common.h"
#include "eval_math.h"

namespace egn {

__global__ void box_overlap_kernel(const double* __restrict__ det, const double* __restrict__ gt, int D, int G, int criterion,
                                   double* __restrict__ ground, double* __restrict__ box3d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * G) return;
  const double* a = det + (size_t)(i / G) * 7;
  const double* b = gt + (size_t)(i % G) * 7;
  const EvalBox d{a[0], a[1], a[2], a[3], a[4], a[5], a[6]}, g{b[0], b[1], b[2], b[3], b[4], b[5], b[6]};
  if (ground) ground[i] = ground_box_overlap(d, g, criterion);
  if (box3d) box3d[i] = box3d_overlap(d, g, criterion);
}

__global__ void image_overlap_kernel(const double* __restrict__ det, const double* __restrict__ gt, int D, int G,
                                     int criterion, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * G) return;
  out[i] = image_box_overlap(det + (size_t)(i / G) * 4, gt + (size_t)(i % G) * 4, criterion);
}

}  // namespace egn

extern "C" int egn_box_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* ground_out,
                                double* box3d_out, void* stream) {
  using namespace egn;
  EGN_REQUIRE(D >= 0 && G >= 0, "egn_box_overlaps: negative count");
  EGN_REQUIRE(criterion >= -1 && criterion <= 1, "egn_box_overlaps: criterion -1 (union), 0 (detection) or 1 (ground truth)");
  EGN_REQUIRE(D * G == 0 || (det && gt && (ground_out || box3d_out)), "egn_box_overlaps: null pointer");
  if (int rc = require_device()) return rc;
  if (D * G == 0) return EGN_OK;
  box_overlap_kernel<<<ceil_div(D * G, 128), 128, 0, as_stream(stream)>>>(det, gt, D, G, criterion, ground_out, box3d_out);
  EGN_LAUNCH_CHECK("box_overlap_kernel");
  return EGN_OK;
}

extern "C" int egn_image_box_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* out,
                                      void* stream) {
  using namespace egn;
  EGN_REQUIRE(D >= 0 && G >= 0, "egn_image_box_overlaps: negative count");
  EGN_REQUIRE(criterion >= -1 && criterion <= 1, "egn_image_box_overlaps: criterion -1, 0 or 1");
  EGN_REQUIRE(D * G == 0 || (det && gt && out), "egn_image_box_overlaps: null pointer");
  if (int rc = require_device()) return rc;
  if (D * G == 0) return EGN_OK;
  image_overlap_kernel<<<ceil_div(D * G, 128), 128, 0, as_stream(stream)>>>(det, gt, D, G, criterion, out);
  EGN_LAUNCH_CHECK("image_overlap_kernel");
  return EGN_OK;
}
