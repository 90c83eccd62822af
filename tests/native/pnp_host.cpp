// Host-compiled harness around egonet_b200/csrc/pnp_math.h (TEST ONLY): the exact kernel source of the
// reprojection refinement checked against cv2.solvePnP goldens on a machine without a GPU.
#include "pnp_math.h"

extern "C" {

void host_pnp_refine(const double* kpts_3d, const double* kpts_2d, int N, int P, double fx, double fy, double cx,
                     double cy, int max_iter, double* refined, double* pose6, double* info, int* status) {
  egn::PnpCamera cam{fx, fy, cx, cy};
  for (int n = 0; n < N; ++n)
    status[n] = egn::pnp_refine_one(kpts_3d + (size_t)n * P * 3, kpts_2d + (size_t)n * P * 2, P, cam,
                                    max_iter > 0 ? max_iter : 20, 1.1920928955078125e-07,
                                    refined + (size_t)n * P * 3, pose6 + (size_t)n * 6, info + (size_t)n * 2);
}

void host_so3(const double* r, double* R, double* r_back, double* Jl) {
  egn::so3_exp(r, R);
  egn::so3_log(R, r_back);
  egn::so3_left_jacobian(r, Jl);
}

void host_eigh12(const double* A, double* w, double* V) {
  double a[144];
  for (int i = 0; i < 144; ++i) a[i] = A[i];
  egn::jacobi_eigh<12>(a, V);
  for (int i = 0; i < 12; ++i) w[i] = a[i * 12 + i];
}
}
