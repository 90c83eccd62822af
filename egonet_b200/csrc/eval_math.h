// Rotated-box overlaps of the KITTI object benchmark, shared by the CUDA kernel (eval.cu) and the host-compiled
// unit harness (tests/native/eval_host.cpp): every function is __host__ __device__.
//
// upstream: tools/kitti-eval/evaluate_object_3d_offline.cpp
//   toPolygon          :266-290   oriented ground-plane rectangle of a box (ry, l, w, t1 = x, t3 = z)
//   groundBoxOverlap   :293-314   bird's-eye-view overlap, criterion -1 union / 0 detection / 1 ground truth
//   box3DOverlap       :317-344   3-D overlap = ground intersection x height overlap
//   imageBoxOverlap    :224-262   axis-aligned image-plane boxes
// The reference builds Boost.Geometry polygons and calls intersection() / union_(); both operands are convex
// (rectangles), so the intersection area is computed here by Sutherland-Hodgman clipping + the shoelace formula,
// and area(union) = area(a) + area(b) - area(intersection) (what union_ yields for overlapping rectangles; for
// disjoint ones the reference's ratio is 0 as well because the intersection is empty).
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define EGN_HD __host__ __device__ __forceinline__
#else
#define EGN_HD inline
#endif

namespace egn {

// box parameters in the order the evaluator's structs use: ry, h, w, l, t1 (x), t2 (y, bottom), t3 (z)
struct EvalBox {
  double ry, h, w, l, t1, t2, t3;
};

// corners of the ground-plane rectangle in the reference's order: R(ry) * (+-l/2, +-w/2) + (t1, t3)
EGN_HD void ground_polygon(const EvalBox& g, double px[4], double pz[4]) {
  const double c = cos(g.ry), s = sin(g.ry);
  const double lx[4] = {g.l / 2, g.l / 2, -g.l / 2, -g.l / 2};
  const double lz[4] = {g.w / 2, -g.w / 2, -g.w / 2, g.w / 2};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    px[i] = c * lx[i] + s * lz[i] + g.t1;
    pz[i] = -s * lx[i] + c * lz[i] + g.t3;
  }
}

EGN_HD double polygon_area(const double* x, const double* y, int n) {
  double a = 0.0;
  for (int i = 0; i < n; ++i) {
    const int j = i + 1 == n ? 0 : i + 1;
    a += x[i] * y[j] - x[j] * y[i];
  }
  return fabs(a) * 0.5;
}

// area of the intersection of two convex quadrilaterals (any orientation)
EGN_HD double convex_quad_intersection_area(const double ax[4], const double ay[4], const double bx[4], const double by[4]) {
  double sx[16], sy[16], tx[16], ty[16];
  int n = 4;
  for (int i = 0; i < 4; ++i) {
    sx[i] = ax[i];
    sy[i] = ay[i];
  }
  // orientation of the clip polygon: a point is inside an edge when it lies on the polygon's side of it
  double orient = 0.0;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    orient += bx[i] * by[j] - bx[j] * by[i];
  }
  const double sgn = orient >= 0.0 ? 1.0 : -1.0;
  for (int e = 0; e < 4 && n > 0; ++e) {
    const int e2 = (e + 1) & 3;
    const double ex = bx[e2] - bx[e], ey = by[e2] - by[e];
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int j = i + 1 == n ? 0 : i + 1;
      const double di = sgn * (ex * (sy[i] - by[e]) - ey * (sx[i] - bx[e]));
      const double dj = sgn * (ex * (sy[j] - by[e]) - ey * (sx[j] - bx[e]));
      if (di >= 0.0) {
        tx[m] = sx[i];
        ty[m] = sy[i];
        ++m;
      }
      if ((di >= 0.0) != (dj >= 0.0)) {
        const double t = di / (di - dj);
        tx[m] = sx[i] + t * (sx[j] - sx[i]);
        ty[m] = sy[i] + t * (sy[j] - sy[i]);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) {
      sx[i] = tx[i];
      sy[i] = ty[i];
    }
  }
  return n >= 3 ? polygon_area(sx, sy, n) : 0.0;
}

EGN_HD double overlap_ratio(double inter, double a, double b, int criterion) {
  const double den = criterion == -1 ? a + b - inter : (criterion == 0 ? a : b);
  return den > 0.0 ? inter / den : 0.0;
}

// groundBoxOverlap(d, g, criterion): bird's-eye-view overlap
EGN_HD double ground_box_overlap(const EvalBox& d, const EvalBox& g, int criterion) {
  double dx[4], dz[4], gx[4], gz[4];
  ground_polygon(d, dx, dz);
  ground_polygon(g, gx, gz);
  const double inter = convex_quad_intersection_area(gx, gz, dx, dz);
  return overlap_ratio(inter, polygon_area(dx, dz, 4), polygon_area(gx, gz, 4), criterion);
}

// box3DOverlap(d, g, criterion): ground intersection x overlap of the vertical extents [t2 - h, t2]
EGN_HD double box3d_overlap(const EvalBox& d, const EvalBox& g, int criterion) {
  double dx[4], dz[4], gx[4], gz[4];
  ground_polygon(d, dx, dz);
  ground_polygon(g, gx, gz);
  const double inter_area = convex_quad_intersection_area(gx, gz, dx, dz);
  const double ymax = fmin(d.t2, g.t2), ymin = fmax(d.t2 - d.h, g.t2 - g.h);
  const double inter_vol = inter_area * fmax(0.0, ymax - ymin);
  return overlap_ratio(inter_vol, d.h * d.l * d.w, g.h * g.l * g.w, criterion);
}

// imageBoxOverlap(a, b, criterion) on [x1, y1, x2, y2]
EGN_HD double image_box_overlap(const double a[4], const double b[4], int criterion) {
  const double w = fmin(a[2], b[2]) - fmax(a[0], b[0]), h = fmin(a[3], b[3]) - fmax(a[1], b[1]);
  if (w <= 0 || h <= 0) return 0.0;
  return overlap_ratio(w * h, (a[2] - a[0]) * (a[3] - a[1]), (b[2] - b[0]) * (b[3] - b[1]), criterion);
}

}  // namespace egn
