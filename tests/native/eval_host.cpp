// Host-compiled harness around egonet_b200/csrc/eval_math.h (TEST ONLY): the exact kernel source checked on CPU.
#include "eval_math.h"

extern "C" {
void host_box_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* ground, double* box3d) {
  for (int i = 0; i < D * G; ++i) {
    const double* a = det + (size_t)(i / G) * 7;
    const double* b = gt + (size_t)(i % G) * 7;
    const egn::EvalBox d{a[0], a[1], a[2], a[3], a[4], a[5], a[6]}, g{b[0], b[1], b[2], b[3], b[4], b[5], b[6]};
    ground[i] = egn::ground_box_overlap(d, g, criterion);
    box3d[i] = egn::box3d_overlap(d, g, criterion);
  }
}
void host_image_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* out) {
  for (int i = 0; i < D * G; ++i) out[i] = egn::image_box_overlap(det + (size_t)(i / G) * 4, gt + (size_t)(i % G) * 4, criterion);
}
}
