// Integer / fp32 arithmetic of the crop front-end, shared by the CUDA kernel (crop.cu) and the
// host-compiled unit harness (tests/native/crop_host.cpp): every function is __host__ __device__ so
// the exact kernel source is checked bit for bit on a CPU before it runs on the GPU.
//
// upstream: EgoNet.crop_single_instance egonet.py:68-95
//   get_affine_transform(c, s, 0, (h, w))        img_proc.py:26-64 (float32 points -> cv2.getAffineTransform)
//   cv2.warpAffine(img, trans, (w, h), INTER_LINEAR)   egonet.py:85-89  (third party: OpenCV imgwarp.cpp,
//       fixed-point path: 10-bit affine offsets, 5-bit sub-pixel positions, 15-bit bilinear weights,
//       BORDER_CONSTANT 0)
//   transforms.ToTensor() + Normalize(mean, std)       car_instance.py:522-531
#pragma once

#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define EGN_CHD __host__ __device__ __forceinline__
#else
#define EGN_CHD inline
#endif

namespace egn {

// products / sums that must round exactly like separate C double operations (no FMA contraction)
EGN_CHD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
EGN_CHD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}

// cv::saturate_cast<int>(double) == cvRound: round half to even, saturating
EGN_CHD int cv_round(double v) {
  const double r = rint(v);
  if (r >= 2147483647.0) return 2147483647;
  if (r <= -2147483648.0) return (int)(-2147483647 - 1);
  return (int)r;
}

// Forward crop affine (source image -> crop), get_affine_transform(..., inv=0) with rot = 0:
// three float32 points per side exactly as upstream builds them, then the exact 3-point affine
// in double (cv2.getAffineTransform).  Only scale[0] is used (img_proc.py:41-42).
EGN_CHD void forward_crop_affine(double cx, double cy, double scale0, int res_w, int res_h, double M[6]) {
  const double src_w = scale0 * 200.0;
  const double p1 = src_w * -0.5;                   // src_dir = (0, -src_w/2) for rot = 0
  const float dst_dir_y = (float)((double)res_w * -0.5);
  const float s0x = (float)cx, s0y = (float)cy;
  const float s1x = (float)(cx + 0.0), s1y = (float)(cy + p1);
  const float d0x = (float)(res_w * 0.5), d0y = (float)(res_h * 0.5);
  const float d1x = (float)(res_w * 0.5 + 0.0), d1y = (float)(res_h * 0.5 + (double)dst_dir_y);
  const float s2x = s1x + (-(s0y - s1y)), s2y = s1y + (s0x - s1x);
  const float d2x = d1x + (-(d0y - d1y)), d2y = d1y + (d0x - d1x);
  // affine mapping src_i -> dst_i
  const double px0 = s0x, py0 = s0y, px1 = s1x, py1 = s1y, px2 = s2x, py2 = s2y;
  const double ax = px1 - px0, ay = py1 - py0, bx = px2 - px0, by = py2 - py0;
  const double det = ax * by - bx * ay;
  const double q[2][3] = {{(double)d0x, (double)d1x, (double)d2x}, {(double)d0y, (double)d1y, (double)d2y}};
  for (int r = 0; r < 2; ++r) {
    const double u = q[r][1] - q[r][0], v = q[r][2] - q[r][0];
    const double a = (u * by - v * ay) / det;
    const double b = (v * ax - u * bx) / det;
    M[3 * r + 0] = a;
    M[3 * r + 1] = b;
    M[3 * r + 2] = q[r][0] - a * px0 - b * py0;
  }
}

// cv::warpAffine without WARP_INVERSE_MAP inverts the 2x3 matrix in double first.
EGN_CHD void cv_invert_affine(const double M[6], double Mi[6]) {
  double D = dadd(dmul(M[0], M[4]), -dmul(M[1], M[3]));
  D = D != 0.0 ? 1.0 / D : 0.0;
  const double A11 = dmul(M[4], D), A22 = dmul(M[0], D);
  Mi[0] = A11;
  Mi[1] = dmul(M[1], -D);
  Mi[3] = dmul(M[3], -D);
  Mi[4] = A22;
  Mi[2] = dadd(dmul(-Mi[0], M[2]), -dmul(Mi[1], M[5]));
  Mi[5] = dadd(dmul(-Mi[3], M[2]), -dmul(Mi[4], M[5]));
}

// Source position of output pixel (x, y) in 1/32-pixel fixed point: integer part (sx, sy) and the
// 5-bit fractions (fx, fy).  AB_BITS = 10, INTER_BITS = 5, round_delta = 1024 / 32 / 2 = 16.
struct WarpPos {
  int sx, sy, fx, fy;
};
EGN_CHD WarpPos warp_position(const double Mi[6], int x, int y) {
  const int adelta = cv_round(dmul(dmul(Mi[0], (double)x), 1024.0));
  const int bdelta = cv_round(dmul(dmul(Mi[3], (double)x), 1024.0));
  const int X0 = cv_round(dmul(dadd(dmul(Mi[1], (double)y), Mi[2]), 1024.0)) + 16;
  const int Y0 = cv_round(dmul(dadd(dmul(Mi[4], (double)y), Mi[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  WarpPos p;
  p.sx = X >> 5;
  p.sy = Y >> 5;
  if (p.sx < -32768) p.sx = -32768;          // saturate_cast<short>
  if (p.sx > 32767) p.sx = 32767;
  if (p.sy < -32768) p.sy = -32768;
  if (p.sy > 32767) p.sy = 32767;
  p.fx = X & 31;
  p.fy = Y & 31;
  return p;
}

// 15-bit bilinear weights of OpenCV's BilinearTab_i[fy*32+fx] = {w00, w01, w10, w11}.
// (1-fy/32)(1-fx/32)*32768 etc. are exact integers; the only entry whose short saturates is
// (0,0): 32768 -> 32767, and the table's sum fix-up then adds the missing 1 to w11.
EGN_CHD void bilinear_weights(int fx, int fy, int w[4]) {
  w[0] = (32 - fy) * (32 - fx) * 32;
  w[1] = (32 - fy) * fx * 32;
  w[2] = fy * (32 - fx) * 32;
  w[3] = fy * fx * 32;
  if ((fx | fy) == 0) {
    w[0] = 32767;
    w[3] = 1;
  }
}

// One output pixel of cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) on an 8-bit, C-channel image.
template <int C>
EGN_CHD void warp_pixel_u8(const uint8_t* img, int img_h, int img_w, int pitch, WarpPos p, uint8_t out[C]) {
  int w[4];
  bilinear_weights(p.fx, p.fy, w);
  int acc[C];
  for (int c = 0; c < C; ++c) acc[c] = 1 << 14;
  for (int dy = 0; dy < 2; ++dy) {
    const int yy = p.sy + dy;
    if (yy < 0 || yy >= img_h) continue;
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = p.sx + dx;
      if (xx < 0 || xx >= img_w) continue;
      const uint8_t* s = img + (size_t)yy * pitch + (size_t)xx * C;
      const int wk = w[2 * dy + dx];
      for (int c = 0; c < C; ++c) acc[c] += wk * (int)s[c];
    }
  }
  for (int c = 0; c < C; ++c) {
    int v = acc[c] >> 15;
    out[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

// ToTensor (uint8 -> float32, / 255) then Normalize ((x - mean) / std), all float32 IEEE ops.
EGN_CHD float normalize_px(uint8_t v, float mean, float stdv) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), stdv);
#else
  volatile float t = (float)v / 255.0f;
  volatile float u = t - mean;
  return u / stdv;
#endif
}

}  // namespace egn
