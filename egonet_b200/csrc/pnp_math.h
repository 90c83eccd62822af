// fp64 arithmetic of the reprojection refinement, shared by the CUDA kernel
// (pnp.cu) and the host-compiled unit harness (tests/native/pnp_host.cpp).
//
// upstream: pnp_refine libs/common/transformation.py:143-157
//     (success, R, T) = cv2.solvePnP(prediction, observation, intrinsics, dist_coeffs, flags=SOLVEPNP_ITERATIVE)
//     refined = cv2.Rodrigues(R)[0] @ prediction.T + T
// cv2.solvePnP is third-party arithmetic that is not under the reference tree (OpenCV 3.4.2 pinned upstream,
// 4.13.0 in the build image).  Its published algorithm for SOLVEPNP_ITERATIVE without an extrinsic guess
// (calib3d, cvFindExtrinsicCameraParams2 + CvLevMarq) is restated here step by step:
//   1. image points are normalised with the intrinsics (no distortion);
//   2. planarity test on the 3x3 scatter matrix of the object points (W[2] / W[1] < 1e-3 -> planar);
//   3. non-planar: DLT -- the 12-vector of the smallest singular value of L^T L, sign fixed by det > 0,
//      rotation = polar factor of its 3x3 part, translation rescaled by |R|_F / |RR|_F;
//   4. Levenberg-Marquardt on [rvec, tvec] with pixel residuals: lambda = 10^k, k0 = -3, the diagonal of
//      J^T J multiplied by (1 + lambda), step rejected (k += 1, up to 16) while the residual norm grows,
//      k -= 1 on acceptance, at most 20 accepted steps, stop when |dp| / |p| < FLT_EPSILON.
// The planar branch of OpenCV (homography initialisation) is not restated: planar inputs are reported with
// status EGN_PNP_PLANAR and left unrefined (upstream keeps the prediction when solvePnP fails).
#pragma once

#include <cmath>

#include "pose_math.h"

namespace egn {

constexpr int kPnpMaxPoints = 64;

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (row-major).  On return the diagonal
// of A holds the eigenvalues and the COLUMNS of V the eigenvectors.
template <int N>
EGN_HD void jacobi_eigh(double* A, double* V) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double app = A[p * N + p], aqq = A[q * N + q];
        if (fabs(apq) <= 1e-300) continue;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(1.0 + theta * theta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < N; ++k) {   // columns p, q
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {   // rows p, q
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
}

// Rodrigues vector -> rotation matrix (row-major 3x3)
EGN_HD void so3_exp(const double r[3], double R[9]) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 2.220446049250313e-16) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = cos(th), s = sin(th), c1 = 1.0 - c;
  const double x = r[0] / th, y = r[1] / th, z = r[2] / th;
  R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

// rotation matrix -> Rodrigues vector (the matrix -> vector branch of cvRodrigues2)
EGN_HD void so3_log(const double R[9], double r[3]) {
  r[0] = R[7] - R[5];
  r[1] = R[2] - R[6];
  r[2] = R[3] - R[1];
  const double s = sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      r[0] = r[1] = r[2] = 0.0;
      return;
    }
    double t = (R[0] + 1.0) * 0.5;
    double rx = sqrt(t > 0 ? t : 0.0);
    t = (R[4] + 1.0) * 0.5;
    double ry = sqrt(t > 0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1.0) * 0.5;
    double rz = sqrt(t > 0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    const double k = th / sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * k;
    r[1] = ry * k;
    r[2] = rz * k;
    return;
  }
  const double k = 0.5 / s * th;
  r[0] *= k;
  r[1] *= k;
  r[2] *= k;
}

// Left Jacobian of SO(3): exp(r + dr) = exp(Jl(r) dr) exp(r) to first order, so
// d(R(r) X)/dr = -[R X]x Jl(r) -- the analytic dR/dr of cvRodrigues2 contracted with X.
EGN_HD void so3_left_jacobian(const double r[3], double J[9]) {
  const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double th = sqrt(th2);
  double a, b;
  if (th < 1e-8) {
    a = 0.5;
    b = 1.0 / 6.0;
  } else {
    a = (1.0 - cos(th)) / th2;
    b = (th - sin(th)) / (th2 * th);
  }
  const double x = r[0], y = r[1], z = r[2];
  // I + a [r]x + b [r]x^2
  J[0] = 1.0 - b * (y * y + z * z); J[1] = -a * z + b * x * y;        J[2] = a * y + b * x * z;
  J[3] = a * z + b * x * y;         J[4] = 1.0 - b * (x * x + z * z); J[5] = -a * x + b * y * z;
  J[6] = -a * y + b * x * z;        J[7] = a * x + b * y * z;         J[8] = 1.0 - b * (x * x + y * y);
}

struct PnpCamera {
  double fx, fy, cx, cy;
};

enum { EGN_PNP_OK = 0, EGN_PNP_PLANAR = 1, EGN_PNP_DEGENERATE = 2 };

// Residual norm |project(X; r, t) - uv|_2 in pixels; optionally accumulates J^T J (6x6) and J^T e.
EGN_HD double pnp_residual(const double* X, const double* uv, int P, const PnpCamera& cam, const double p[6],
                           double* JtJ, double* JtE) {
  double R[9], Jl[9];
  so3_exp(p, R);
  if (JtJ) {
    so3_left_jacobian(p, Jl);
    for (int i = 0; i < 36; ++i) JtJ[i] = 0.0;
    for (int i = 0; i < 6; ++i) JtE[i] = 0.0;
  }
  double sq = 0.0;
  for (int i = 0; i < P; ++i) {
    const double X0 = X[3 * i], X1 = X[3 * i + 1], X2 = X[3 * i + 2];
    const double Y0 = R[0] * X0 + R[1] * X1 + R[2] * X2;
    const double Y1 = R[3] * X0 + R[4] * X1 + R[5] * X2;
    const double Y2 = R[6] * X0 + R[7] * X1 + R[8] * X2;
    const double xc = Y0 + p[3], yc = Y1 + p[4], zc = Y2 + p[5];
    const double iz = zc != 0.0 ? 1.0 / zc : 1.0;
    const double x = xc * iz, y = yc * iz;
    const double e0 = cam.fx * x + cam.cx - uv[2 * i], e1 = cam.fy * y + cam.cy - uv[2 * i + 1];
    sq += e0 * e0 + e1 * e1;
    if (!JtJ) continue;
    // d(u,v)/d(Xc)
    const double d[2][3] = {{cam.fx * iz, 0.0, -cam.fx * x * iz}, {0.0, cam.fy * iz, -cam.fy * y * iz}};
    // -[Y]x
    const double H[3][3] = {{0.0, Y2, -Y1}, {-Y2, 0.0, Y0}, {Y1, -Y0, 0.0}};
    double J[2][6];
    for (int a = 0; a < 2; ++a) {
      double w[3];
      for (int k = 0; k < 3; ++k) w[k] = d[a][0] * H[0][k] + d[a][1] * H[1][k] + d[a][2] * H[2][k];
      for (int k = 0; k < 3; ++k) J[a][k] = w[0] * Jl[k] + w[1] * Jl[3 + k] + w[2] * Jl[6 + k];
      for (int k = 0; k < 3; ++k) J[a][3 + k] = d[a][k];
    }
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b) JtJ[a * 6 + b] += J[0][a] * J[0][b] + J[1][a] * J[1][b];
      JtE[a] += J[0][a] * e0 + J[1][a] * e1;
    }
  }
  return sqrt(sq);
}

// x = pinv(A) b for a symmetric 6x6 A (cv::solve(..., DECOMP_SVD): singular values below
// 2 * DBL_EPSILON * sum(w) are dropped).
EGN_HD void solve_sym6(const double* A_in, const double* b, double* x) {
  // Well-conditioned (the damped normal equations almost always are): Cholesky, the same solution as the SVD route up
  // to rounding at a tenth of its cost.  A pivot below 1e-10 of the largest diagonal entry falls through to the
  // eigen-decomposition, which reproduces the truncated pseudo-inverse of cv::solve(DECOMP_SVD).
  {
    double L[36];
    double dmax = 0.0;
    for (int i = 0; i < 6; ++i) dmax = fmax(dmax, fabs(A_in[i * 6 + i]));
    bool ok = dmax > 0.0;
    for (int j = 0; j < 6 && ok; ++j) {
      double d = A_in[j * 6 + j];
      for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
      if (!(d > 1e-10 * dmax)) {
        ok = false;
        break;
      }
      const double ljj = sqrt(d);
      L[j * 6 + j] = ljj;
      for (int i = j + 1; i < 6; ++i) {
        double v = A_in[i * 6 + j];
        for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
        L[i * 6 + j] = v / ljj;
      }
    }
    if (ok) {
      double y[6];
      for (int i = 0; i < 6; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[i * 6 + k] * y[k];
        y[i] = v / L[i * 6 + i];
      }
      for (int i = 5; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < 6; ++k) v -= L[k * 6 + i] * x[k];
        x[i] = v / L[i * 6 + i];
      }
      return;
    }
  }
  double A[36], V[36];
  for (int i = 0; i < 36; ++i) A[i] = A_in[i];
  jacobi_eigh<6>(A, V);
  double sum = 0.0;
  for (int i = 0; i < 6; ++i) sum += fabs(A[i * 6 + i]);
  const double thr = sum * 2.0 * 2.220446049250313e-16;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int k = 0; k < 6; ++k) {
    const double w = A[k * 6 + k];
    if (fabs(w) <= thr) continue;
    double proj = 0.0;
    for (int i = 0; i < 6; ++i) proj += V[i * 6 + k] * b[i];
    proj /= w;
    for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + k] * proj;
  }
}

// Steps 1-3: DLT initialisation.  p = [rvec, tvec].
EGN_HD int pnp_dlt_init(const double* X, const double* uv, int P, const PnpCamera& cam, double p[6]) {
  // planarity test: singular values of the scatter matrix of the object points
  double Mc[3] = {0, 0, 0};
  for (int i = 0; i < P; ++i)
    for (int a = 0; a < 3; ++a) Mc[a] += X[3 * i + a];
  for (int a = 0; a < 3; ++a) Mc[a] /= P;
  double MM[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, MV[9];
  for (int i = 0; i < P; ++i)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) MM[a * 3 + b] += (X[3 * i + a] - Mc[a]) * (X[3 * i + b] - Mc[b]);
  jacobi_eigh<3>(MM, MV);
  double w0 = MM[0], w1 = MM[4], w2 = MM[8], tmp;
  if (w0 < w1) { tmp = w0; w0 = w1; w1 = tmp; }
  if (w1 < w2) { tmp = w1; w1 = w2; w2 = tmp; }
  if (w0 < w1) { tmp = w0; w0 = w1; w1 = tmp; }
  if (!(w1 > 0.0)) return EGN_PNP_DEGENERATE;
  if (w2 / w1 < 1e-3) return EGN_PNP_PLANAR;
  // L^T L accumulated row by row (two rows of L per point)
  double LL[144], LV[144];
  for (int i = 0; i < 144; ++i) LL[i] = 0.0;
  const double ifx = 1.0 / cam.fx, ify = 1.0 / cam.fy;
  for (int i = 0; i < P; ++i) {
    const double x = -((uv[2 * i] - cam.cx) * ifx), y = -((uv[2 * i + 1] - cam.cy) * ify);
    const double M0 = X[3 * i], M1 = X[3 * i + 1], M2 = X[3 * i + 2];
    const double ra[12] = {M0, M1, M2, 1.0, 0, 0, 0, 0, x * M0, x * M1, x * M2, x};
    const double rb[12] = {0, 0, 0, 0, M0, M1, M2, 1.0, y * M0, y * M1, y * M2, y};
    for (int a = 0; a < 12; ++a)
      for (int b = 0; b < 12; ++b) LL[a * 12 + b] += ra[a] * ra[b] + rb[a] * rb[b];
  }
  jacobi_eigh<12>(LL, LV);
  int kmin = 0;
  for (int k = 1; k < 12; ++k)
    if (LL[k * 12 + k] < LL[kmin * 12 + kmin]) kmin = k;
  double v[12];
  for (int i = 0; i < 12; ++i) v[i] = LV[i * 12 + kmin];
  double RR[3][3] = {{v[0], v[1], v[2]}, {v[4], v[5], v[6]}, {v[8], v[9], v[10]}};
  double tt[3] = {v[3], v[7], v[11]};
  if (det3(RR) < 0) {
    for (int a = 0; a < 3; ++a) {
      tt[a] = -tt[a];
      for (int b = 0; b < 3; ++b) RR[a][b] = -RR[a][b];
    }
  }
  double sc = 0.0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) sc += RR[a][b] * RR[a][b];
  sc = sqrt(sc);
  if (!(sc > 2.220446049250313e-16)) return EGN_PNP_DEGENERATE;
  // polar factor: RR = U S V^T -> R = U V^T
  double V3[3][3], s3[3];
  svd3(RR, V3, s3);   // RR <- U diag(s)
  if (!(s3[2] > 1e-14 * s3[0])) return EGN_PNP_DEGENERATE;
  double R[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      R[a * 3 + b] = RR[a][0] / s3[0] * V3[b][0] + RR[a][1] / s3[1] * V3[b][1] + RR[a][2] / s3[2] * V3[b][2];
  so3_log(R, p);
  const double k = sqrt(3.0) / sc;   // |R|_F / |RR|_F
  for (int a = 0; a < 3; ++a) p[3 + a] = tt[a] * k;
  return EGN_PNP_OK;
}

// Step 4: CvLevMarq driven exactly as cvFindExtrinsicCameraParams2 drives it.  Returns the number of
// accepted iterations; *err_out = final residual norm (pixels).
EGN_HD int pnp_levmarq(const double* X, const double* uv, int P, const PnpCamera& cam, double p[6],
                       int max_iter, double eps, double* err_out) {
  int lambda_lg10 = -3, iters = 0;
  double JtJ[36], JtE[6], prev[6], A[36], dp[6];
  double prev_err = 0.0, err = 0.0;
  for (;;) {
    const double e_here = pnp_residual(X, uv, P, cam, p, JtJ, JtE);   // state CALC_J
    if (iters == 0) prev_err = e_here;
    for (int i = 0; i < 6; ++i) prev[i] = p[i];
    for (;;) {                                                        // step() + state CHECK_ERR
      const double lambda = exp(lambda_lg10 * 2.302585092994046);
      for (int i = 0; i < 36; ++i) A[i] = JtJ[i];
      for (int i = 0; i < 6; ++i) A[i * 6 + i] *= 1.0 + lambda;
      solve_sym6(A, JtE, dp);
      for (int i = 0; i < 6; ++i) p[i] = prev[i] - dp[i];
      err = pnp_residual(X, uv, P, cam, p, nullptr, nullptr);
      if (err > prev_err && ++lambda_lg10 <= 16) continue;
      break;
    }
    lambda_lg10 = lambda_lg10 - 1 > -16 ? lambda_lg10 - 1 : -16;
    double dn = 0.0, pn = 0.0;
    for (int i = 0; i < 6; ++i) {
      dn += (p[i] - prev[i]) * (p[i] - prev[i]);
      pn += prev[i] * prev[i];
    }
    if (++iters >= max_iter || sqrt(dn) < eps * sqrt(pn)) break;   // CV_RELATIVE_L2: |p - prev| / |prev|
    prev_err = err;
  }
  *err_out = err;
  return iters;
}

// One instance of pnp_refine.  X [P,3] prediction (camera frame), uv [P,2] observation (pixels).
// refined [P,3] = R X + T; pose6 = rvec | tvec; info = {accepted LM iterations, final residual norm (px)}.
EGN_HD int pnp_refine_one(const double* X, const double* uv, int P, const PnpCamera& cam, int max_iter,
                          double eps, double* refined, double* pose6, double* info) {
  double p[6] = {0, 0, 0, 0, 0, 0};
  int status = pnp_dlt_init(X, uv, P, cam, p);
  double err = 0.0;
  int iters = 0;
  if (status == EGN_PNP_OK) {
    iters = pnp_levmarq(X, uv, P, cam, p, max_iter, eps, &err);
    for (int i = 0; i < 6; ++i)
      if (!(fabs(p[i]) < 1e300)) status = EGN_PNP_DEGENERATE;   // NaN / inf
  }
  if (status != EGN_PNP_OK) {
    for (int i = 0; i < 6; ++i) p[i] = 0.0;    // upstream returns the prediction unchanged on failure
    err = pnp_residual(X, uv, P, cam, p, nullptr, nullptr);
  }
  double R[9];
  so3_exp(p, R);
  for (int i = 0; i < P; ++i)
    for (int a = 0; a < 3; ++a)
      refined[3 * i + a] = R[3 * a] * X[3 * i] + R[3 * a + 1] * X[3 * i + 1] + R[3 * a + 2] * X[3 * i + 2] + p[3 + a];
  if (pose6)
    for (int i = 0; i < 6; ++i) pose6[i] = p[i];
  if (info) {
    info[0] = (double)iters;
    info[1] = err;
  }
  return status;
}

}  // namespace egn
