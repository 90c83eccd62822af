// fp64 geometry shared by the CUDA kernels (pose.cu) and the host-compiled unit
// harness (tests/native/pose_host.cpp): every function is __host__ __device__ so
// the exact same source is validated on CPU against the oracle before it is
// run on the GPU.
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define EGN_HD __host__ __device__ __forceinline__
#else
#define EGN_HD inline
#endif

namespace egn {

// ---------------------------------------------------------------------------
// local (0..1 crop coordinates) -> screen coordinates
// upstream: EgoNet.get_keypoints egonet.py:436-453; get_affine_transform
// img_proc.py:26-64 (float32 point construction, then cv2.getAffineTransform =
// exact 3-point affine solved in double); affine_transform_modified :71-78.
// ---------------------------------------------------------------------------
EGN_HD void inverse_crop_affine(double cx, double cy, double scale0, double rot_deg,
                                    int res_w, int res_h, double M[6]) {
  const double src_w = scale0 * 200.0;  // SIZE = 200, only scale[0] is used (img_proc.py:41-42)
  const double rot_rad = 3.141592653589793 * rot_deg / 180.0;
  const double sn = sin(rot_rad), cs = cos(rot_rad);
  const double p1 = src_w * -0.5;
  const double sdx = 0.0 * cs - p1 * sn, sdy = 0.0 * sn + p1 * cs;
  const float dst_dir_y = (float)((double)res_w * -0.5);
  // float32 point arrays exactly as upstream builds them
  float s0x = (float)cx, s0y = (float)cy;
  float s1x = (float)(cx + sdx), s1y = (float)(cy + sdy);
  float d0x = (float)(res_w * 0.5), d0y = (float)(res_h * 0.5);
  float d1x = (float)(res_w * 0.5 + 0.0), d1y = (float)(res_h * 0.5 + (double)dst_dir_y);
  // get_3rd_point(a, b) = b + (-(a-b).y, (a-b).x), all float32
  float s2x = s1x + (-(s0y - s1y)), s2y = s1y + (s0x - s1x);
  float d2x = d1x + (-(d0y - d1y)), d2y = d1y + (d0x - d1x);
  // inv=1: the affine that maps dst_i -> src_i
  const double px0 = d0x, py0 = d0y, px1 = d1x, py1 = d1y, px2 = d2x, py2 = d2y;
  const double ax = px1 - px0, ay = py1 - py0, bx = px2 - px0, by = py2 - py0;
  const double det = ax * by - bx * ay;
  const double q[2][3] = {{(double)s0x, (double)s1x, (double)s2x}, {(double)s0y, (double)s1y, (double)s2y}};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double u = q[r][1] - q[r][0], v = q[r][2] - q[r][0];
    const double a = (u * by - v * ay) / det;
    const double b = (v * ax - u * bx) / det;
    M[3 * r + 0] = a;
    M[3 * r + 1] = b;
    M[3 * r + 2] = q[r][0] - a * px0 - b * py0;
  }
}

// ---------------------------------------------------------------------------
// pose solve
// ---------------------------------------------------------------------------
// interp_dict['bbox12'] (car_instance.py:63-70), already 0-based

EGN_HD double det3(const double m[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
         m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix: on return A = U*diag(s), V
// orthogonal, H = U diag(s) V^T, singular values sorted descending.
EGN_HD void svd3(double A[3][3], double V[3][3], double s[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double al = 0, be = 0, ga = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        al += A[i][p] * A[i][p];
        be += A[i][q] * A[i][q];
        ga += A[i][p] * A[i][q];
      }
      if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
      rotated = true;
      const double zeta = (be - al) / (2.0 * ga);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double x = A[i][p], y = A[i][q];
        A[i][p] = c * x - sn * y;
        A[i][q] = sn * x + c * y;
        x = V[i][p];
        y = V[i][q];
        V[i][p] = c * x - sn * y;
        V[i][q] = sn * x + c * y;
      }
    }
    if (!rotated) break;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) s[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  // sort columns by descending singular value (3-element network)
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const int a = pass == 1 ? 1 : 0, b = pass == 0 ? 1 : 2;
    if (s[a] < s[b]) {
      double t = s[a]; s[a] = s[b]; s[b] = t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        t = A[i][a]; A[i][a] = A[i][b]; A[i][b] = t;
        t = V[i][a]; V[i][a] = V[i][b]; V[i][b] = t;
      }
    }
  }
}

// R = argmin ||R X + t - Y|| with the reflection fix of transformation.py:125-132.
// s_out (optional): singular values of H, descending; d_out (optional): the sign applied to the
// smallest one (-1 when the unconstrained optimum is a reflection).
EGN_HD void kabsch_rotation(const double H_in[3][3], double R[3][3], double* s_out = nullptr,
                            double* d_out = nullptr) {
  double A[3][3], V[3][3], s[3], U[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = H_in[i][j];
  svd3(A, V, s);
  const double tiny = 1e-13 * (s[0] > 0 ? s[0] : 1.0);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double inv = s[j] > tiny ? 1.0 / s[j] : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][j] = A[i][j] * inv;
  }
  if (!(s[1] > tiny)) {
    // rank <= 1 (collinear points): any unit vector orthogonal to u0
    const int k = fabs(U[0][0]) < fabs(U[1][0]) ? (fabs(U[0][0]) < fabs(U[2][0]) ? 0 : 2)
                                                 : (fabs(U[1][0]) < fabs(U[2][0]) ? 1 : 2);
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    const double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
    double n2 = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      U[i][1] = e[i] - d * U[i][0];
      n2 += U[i][1] * U[i][1];
    }
    const double inv = 1.0 / sqrt(n2);
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][1] *= inv;
  }
  if (s[2] > tiny) {
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][2] = A[i][2] / s[2];
  } else {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  // numpy: H = U S Vt, R = Vt.T U.T; if det(R) < 0 flip the last row of Vt (smallest sigma)
  const double d = det3(V) * det3(U) < 0 ? -1.0 : 1.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + d * V[i][2] * U[j][2];
  if (s_out) {
    s_out[0] = s[0]; s_out[1] = s[1]; s_out[2] = s[2];
  }
  if (d_out) *d_out = d;
}

// ---------------------------------------------------------------------------
// General point-set alignment (transformation.py:48-141), points stored [P,3] row-major.
// ---------------------------------------------------------------------------
// compute_rigid_transform(X, Y, W) transformation.py:99-134: least-squares R, t with Y ~ R X + t.
// Centroids are the UNWEIGHTED means (as upstream); W is null, a [P] diagonal (w_mode 1) or a full
// [P,P] matrix (w_mode 2): H = Xm W Ym^T.  aligned (optional, [P,3]) = R X + t, i.e. procrustes_transform
// (transformation.py:136-141).
EGN_HD void rigid_transform_one(const double* X, const double* Y, const double* W, int w_mode, int P,
                                double* R_out, double* t_out, double* aligned) {
  double cX[3] = {0, 0, 0}, cY[3] = {0, 0, 0};
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      cX[a] += X[3 * i + a];
      cY[a] += Y[3 * i + a];
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cX[a] /= P;
    cY[a] /= P;
  }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < P; ++i) {
    // row i of (Xm W): for the diagonal / unweighted forms only column i of W is non-zero
    if (w_mode == 2) {
      for (int j = 0; j < P; ++j) {
        const double wij = W[(size_t)i * P + j];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) H[a][b] += (X[3 * i + a] - cX[a]) * wij * (Y[3 * j + b] - cY[b]);
      }
    } else {
      const double wi = w_mode == 1 ? W[i] : 1.0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) H[a][b] += (X[3 * i + a] - cX[a]) * wi * (Y[3 * i + b] - cY[b]);
    }
  }
  double R[3][3];
  kabsch_rotation(H, R);
  double t[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) t[a] = -(R[a][0] * cX[0] + R[a][1] * cX[1] + R[a][2] * cX[2]) + cY[a];
  if (R_out) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R_out[3 * i + j] = R[i][j];
  }
  if (t_out) {
    t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2];
  }
  if (aligned) {
    for (int i = 0; i < P; ++i)
#pragma unroll
      for (int a = 0; a < 3; ++a)
        aligned[3 * i + a] = R[a][0] * X[3 * i] + R[a][1] * X[3 * i + 1] + R[a][2] * X[3 * i + 2] + t[a];
  }
}

// compute_similarity_transform(X, Y, compute_optimal_scale) transformation.py:48-97 (MATLAB procrustes):
// X = targets, Y = inputs, both [P,3].  out5 = {d, b}; Z [P,3] transformed Y; T [9] rotation (row-major,
// applied on the right: Z = b * Y T + c); c [3].
EGN_HD void similarity_transform_one(const double* X, const double* Y, int P, int optimal_scale, double* d_out,
                                     double* b_out, double* Z, double* T_out, double* c_out) {
  double muX[3] = {0, 0, 0}, muY[3] = {0, 0, 0};
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      muX[a] += X[3 * i + a];
      muY[a] += Y[3 * i + a];
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    muX[a] /= P;
    muY[a] /= P;
  }
  double ssX = 0, ssY = 0;
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double x = X[3 * i + a] - muX[a], y = Y[3 * i + a] - muY[a];
      ssX += x * x;
      ssY += y * y;
    }
  const double normX = sqrt(ssX), normY = sqrt(ssY);
  // A = X0^T Y0 = U S Vt;  T = V U^T with the reflection fix on V's last column
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) A[a][b] += ((X[3 * i + a] - muX[a]) / normX) * ((Y[3 * i + b] - muY[b]) / normY);
  double T[3][3], s[3], d;
  kabsch_rotation(A, T, s, &d);
  const double traceTA = s[0] + s[1] + d * s[2];
  double b, dd;
  if (optimal_scale) {
    b = traceTA * normX / normY;
    dd = 1.0 - traceTA * traceTA;
  } else {
    b = 1.0;
    dd = 1.0 + ssY / ssX - 2.0 * traceTA * normY / normX;
  }
  const double zs = optimal_scale ? normX * traceTA : normY;     // Z = zs * (Y0 T) + muX
  if (Z) {
    for (int i = 0; i < P; ++i) {
      const double y0 = (Y[3 * i] - muY[0]) / normY, y1 = (Y[3 * i + 1] - muY[1]) / normY,
                   y2 = (Y[3 * i + 2] - muY[2]) / normY;
#pragma unroll
      for (int a = 0; a < 3; ++a) Z[3 * i + a] = zs * (y0 * T[0][a] + y1 * T[1][a] + y2 * T[2][a]) + muX[a];
    }
  }
  if (T_out) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) T_out[3 * i + j] = T[i][j];
  }
  if (c_out) {
#pragma unroll
    for (int a = 0; a < 3; ++a) c_out[a] = muX[a] - b * (muY[0] * T[0][a] + muY[1] * T[1][a] + muY[2] * T[2][a]);
  }
  if (d_out) *d_out = dd;
  if (b_out) *b_out = b;
}

EGN_HD double observation_angle(double ry, double x3d, double z3d) {
  const double PI = 3.141592653589793;
  double alpha = ry - atan2(-z3d, x3d) - 0.5 * PI;
  while (alpha > PI) alpha -= PI * 2;
  while (alpha < -PI) alpha += PI * 2;
  return alpha;
}

// One instance: pr = [P,3] predicted cuboid (P = 8 or 32), kpt_x0 = screen x of
// the first 2D key-point (proj mode), o = [7] euler xyz | translation | alpha,
// rot = [9] or nullptr.
EGN_HD void pose_solve_one(const double* pr, int P, double kpt_x0, double fx, double cx,
                           int alpha_mode, double* o, double* rot) {
  // interp_dict['bbox12'] (car_instance.py:63-70), 0-based
  const int kPar[12] = {0, 2, 4, 6, 0, 1, 2, 3, 0, 1, 4, 5};
  const int kChi[12] = {1, 3, 5, 7, 4, 5, 6, 7, 2, 3, 6, 7};
  // --- template cuboid (egonet.py:238-263) ---
  double len[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) {
    const double dx = pr[3 * kPar[e]] - pr[3 * kChi[e]];
    const double dy = pr[3 * kPar[e] + 1] - pr[3 * kChi[e] + 1];
    const double dz = pr[3 * kPar[e] + 2] - pr[3 * kChi[e] + 2];
    len[e] = sqrt(dx * dx + dy * dy + dz * dz);
  }
  const double h = (((len[0] + len[1]) + len[2]) + len[3]) / 4;
  const double l = (((len[4] + len[5]) + len[6]) + len[7]) / 4;
  const double w = (((len[8] + len[9]) + len[10]) + len[11]) / 4;
  // float32-rounded offsets, as upstream (np.float32(l) / 2 etc., egonet.py:252-255)
  const double ox = (double)((float)l / 2.0f), oy = (double)(float)h, oz = (double)((float)w / 2.0f);
  double T[32][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    T[i][0] = (i < 4 ? l : 0.0) - ox;
    T[i][1] = ((i & 1) ? h : 0.0) - oy;
    T[i][2] = (((i >> 1) & 1) ? 0.0 : w) - oz;
  }
  if (P == 32) {
    for (int e = 0; e < 12; ++e) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double pa = T[kPar[e]][a], seg = T[kChi[e]][a] - pa;
        T[8 + e][a] = pa + 0.332 * seg;
        T[20 + e][a] = pa + 0.667 * seg;
      }
    }
  }
  // --- Kabsch (transformation.py:99-134): X = template, Y = prediction ---
  double cX[3] = {0, 0, 0}, cY[3] = {0, 0, 0};
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      cX[a] += T[i][a];
      cY[a] += pr[3 * i + a];
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cX[a] /= P;
    cY[a] /= P;
  }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < P; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) H[a][b] += (T[i][a] - cX[a]) * (pr[3 * i + b] - cY[b]);
  double R[3][3];
  kabsch_rotation(H, R);
  // --- Euler angles of the extrinsic 'yxz' sequence, reordered to [x, y, z] (egonet.py:274-276)
  const double r21 = fmin(1.0, fmax(-1.0, R[2][1]));
  const double ex = asin(r21);
  const double ey = atan2(-R[2][0], R[2][2]);
  const double ez = atan2(-R[0][1], R[1][1]);
  // --- translation = first predicted point (egonet.py:294) ---
  const double tx = pr[0], ty = pr[1], tz = pr[2];
  // --- observation angle (egonet.py:203-236) ---
  double x3d, z3d;
  if (alpha_mode == 1 /* EGN_ALPHA_PROJ */) {
    x3d = kpt_x0 - cx;
    z3d = fx;
  } else {
    x3d = tx;
    z3d = tz;
  }
  const double alpha = observation_angle(ey, x3d, z3d);
  o[0] = ex; o[1] = ey; o[2] = ez; o[3] = tx; o[4] = ty; o[5] = tz; o[6] = alpha;
  if (rot) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) rot[3 * i + j] = R[i][j];
  }
}

}  // namespace egn
