// Host-compiled harness around egonet_b200/csrc/crop_math.h (TEST ONLY).
// The header's functions are __host__ __device__; compiling them with g++ lets
// tests/test_native_host.py check the exact kernel arithmetic bit for bit against
// the golden crops (cv2.warpAffine + torchvision) on a machine without a GPU.
#include "crop_math.h"

extern "C" {

// Same loop nest as crop_warp_kernel (crop.cu), one crop at a time.
void host_crop_instances(const uint8_t* img, int img_h, int img_w, int pitch, const double* center,
                         const double* scale, int N, int res_w, int res_h, const float* mean,
                         const float* stdv, float* out_nchw, uint8_t* out_u8) {
  const size_t plane = (size_t)res_w * res_h;
  for (int n = 0; n < N; ++n) {
    double M[6], Mi[6];
    egn::forward_crop_affine(center[2 * n], center[2 * n + 1], scale[2 * n], res_w, res_h, M);
    egn::cv_invert_affine(M, Mi);
    for (int y = 0; y < res_h; ++y)
      for (int x = 0; x < res_w; ++x) {
        uint8_t px[3];
        egn::warp_pixel_u8<3>(img, img_h, img_w, pitch, egn::warp_position(Mi, x, y), px);
        for (int c = 0; c < 3; ++c) {
          if (out_nchw)
            out_nchw[(size_t)n * 3 * plane + c * plane + (size_t)y * res_w + x] =
                egn::normalize_px(px[c], mean ? mean[c] : 0.f, stdv ? stdv[c] : 1.f);
          if (out_u8) out_u8[((size_t)n * plane + (size_t)y * res_w + x) * 3 + c] = px[c];
        }
      }
  }
}

void host_forward_crop_affine(double cx, double cy, double scale0, int res_w, int res_h, double* M) {
  egn::forward_crop_affine(cx, cy, scale0, res_w, res_h, M);
}

void host_bilinear_weights(int fx, int fy, int* w) { egn::bilinear_weights(fx, fy, w); }
}
