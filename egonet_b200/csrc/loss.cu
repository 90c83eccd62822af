// Heat-map MSE loss, forward + gradient in one pass (the loss of the reference's IGR training loop).
//
// upstream: JointsMSELoss.forward libs/loss/function.py:28-46 and JointsCompositeLoss.calc_hm_loss
// libs/loss/function.py:95-111:  loss = (1/K) * sum_k 0.5 * mean_{b,h,w} (w_bk * (pred - gt))^2,
// with w = target_weight[b,k] when use_target_weight (else 1).  d loss / d pred = w^2 (pred - gt) / (K*B*H*W).
// Back-propagation through HC continues in hrnet_train.cu.  Below it: the coordinate and
// cross-ratio terms of JointsCompositeLoss (function.py:113-202).
//
// HBM-bound elementwise + reduction: 8 B read (+4 B written with the gradient) per element; per-thread
// fp64 partial sums, shuffle + shared reduction, one fp64 atomicAdd per CTA.
#include "common.h"
#include "target_math.h"

namespace egn {

// One CTA per (sample, joint) map: Gaussian dot targets of generate_target (img_proc.py:347-409), written
// once, coalesced (64 x 64 x 4 B per map).
__global__ void __launch_bounds__(256)
generate_target_kernel(const double* __restrict__ joints, const float* __restrict__ vis, double in0, double in1,
                       int hs0, int hs1, double sigma, float* __restrict__ target, float* __restrict__ weight) {
  const int m = blockIdx.x;
  const float v = vis ? vis[m] : (float)joints[3 * m + 2];
  const TargetDot d = target_dot(joints[3 * m], joints[3 * m + 1], v, in0, in1, hs0, hs1, sigma);
  if (threadIdx.x == 0 && weight) weight[m] = d.weight;
  float* out = target + (size_t)m * hs0 * hs1;
  for (int e = threadIdx.x; e < hs0 * hs1; e += blockDim.x) out[e] = target_value(d, e / hs1, e % hs1, hs0, hs1, sigma);
}

__global__ void __launch_bounds__(256)
mse_hm_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ weight,
              int64_t n, int hw, double scale, float* __restrict__ grad, double* __restrict__ acc) {
  double s = 0.0;
  const float gscale = (float)(2.0 * scale);   // d/dp of scale * d^2
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float w = weight ? __ldg(weight + i / hw) : 1.0f;
    const float d = w * __ldg(pred + i) - w * __ldg(gt + i);   // upstream multiplies both operands by w
    s += (double)d * (double)d;
    if (grad) grad[i] = gscale * w * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(acc, s * scale);
  }
}

__global__ void mse_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)*acc; }

// ---------------------------------------------------------------------------
// Coordinate + cross-ratio terms of JointsCompositeLoss (function.py:113-202), forward and gradient, ONE CTA
// (B*K*2 <= ~20k elements, B*L <= ~3k lines: latency bound).
//   coor: criterion(pred, gt / img_size) with mean reduction over B*K*2 (calc_coor_loss :159-168)
//   cr  : per (sample, line of 4 points) v = (|AC|^2 |BD|^2) / (|BC|^2 |AD|^2) / target_cr^2 (appro_cr,
//         img_proc.py:709-720), line loss = criterion(v, 1); lines whose smallest non-zero pairwise distance is
//         <= threshold are masked out (get_cr_mask :140-153); loss = sum(mask * line loss) / sum(mask)  (:113-138)
// criterion kind: 0 = mse, 1 = smooth L1 (beta 1), 2 = L1 (loss_dict :16-19).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float crit_value(int kind, float d) {
  const float a = fabsf(d);
  return kind == 0 ? d * d : (kind == 1 ? (a < 1.f ? 0.5f * d * d : a - 0.5f) : a);
}
__device__ __forceinline__ float crit_grad(int kind, float d) {
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return kind == 0 ? 2.f * d : (kind == 1 ? (fabsf(d) < 1.f ? d : sgn) : sgn);
}

__device__ double block_sum(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
  return t;
}

__global__ void __launch_bounds__(256)
coord_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt_px, int B, int K, float img_w, float img_h,
                  int coor_kind, float coor_weight, const int32_t* __restrict__ cr_idx, int L, int cr_kind, float cr_weight,
                  float target_cr, float cr_threshold, float* __restrict__ loss_out, float* __restrict__ grad) {
  __shared__ double sm[8];
  const int n = B * K * 2;
  // ---- coordinate term
  double s = 0.0;
  if (coor_weight != 0.f) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const float g = gt_px[e] / ((e & 1) ? img_h : img_w);
      const float d = pred[e] - g;
      s += (double)crit_value(coor_kind, d);
      if (grad) grad[e] = coor_weight * crit_grad(coor_kind, d) / (float)n;
    }
  } else if (grad) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) grad[e] = 0.f;
  }
  const double coor = block_sum(s, sm) / (double)n;
  // ---- cross-ratio term
  double cr = 0.0;
  if (cr_idx && L > 0 && cr_weight != 0.f) {
    const int lines = B * L;
    // pass 1: mask count
    double cnt = 0.0;
    for (int e = threadIdx.x; e < lines; e += blockDim.x) {
      const int b = e / L, l = e - b * L;
      float mind = 3.4e38f;
      for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
          const float* pi = pred + ((size_t)b * K + cr_idx[4 * l + i]) * 2;
          const float* pj = pred + ((size_t)b * K + cr_idx[4 * l + j]) * 2;
          // scipy.spatial.distance_matrix on the float32 coordinates promoted to float64
          const double dx = (double)pi[0] - (double)pj[0], dy = (double)pi[1] - (double)pj[1];
          const float dist = (float)sqrt(dx * dx + dy * dy);
          if (dist != 0.f && dist < mind) mind = dist;
        }
      if (mind < 3.0e38f && mind > cr_threshold) cnt += 1.0;
    }
    const double total = block_sum(cnt, sm);
    __syncthreads();                                       // grad[] of the coordinate term complete before the atomics
    if (total > 0.0) {
      double acc = 0.0;
      for (int e = threadIdx.x; e < lines; e += blockDim.x) {
        const int b = e / L, l = e - b * L;
        float P[4][2];
        float mind = 3.4e38f;
        for (int i = 0; i < 4; ++i) {
          const float* pi = pred + ((size_t)b * K + cr_idx[4 * l + i]) * 2;
          P[i][0] = pi[0];
          P[i][1] = pi[1];
        }
        for (int i = 0; i < 4; ++i)
          for (int j = i + 1; j < 4; ++j) {
            const double dx = (double)P[i][0] - (double)P[j][0], dy = (double)P[i][1] - (double)P[j][1];
            const float dist = (float)sqrt(dx * dx + dy * dy);
            if (dist != 0.f && dist < mind) mind = dist;
          }
        if (!(mind < 3.0e38f && mind > cr_threshold)) continue;
        const float ACx = P[2][0] - P[0][0], ACy = P[2][1] - P[0][1];
        const float BDx = P[3][0] - P[1][0], BDy = P[3][1] - P[1][1];
        const float BCx = P[2][0] - P[1][0], BCy = P[2][1] - P[1][1];
        const float ADx = P[3][0] - P[0][0], ADy = P[3][1] - P[0][1];
        const float a = ACx * ACx + ACy * ACy, bb = BDx * BDx + BDy * BDy;
        const float c = BCx * BCx + BCy * BCy, d = ADx * ADx + ADy * ADy;
        const float v = (a * bb) / (c * d) / (target_cr * target_cr);
        acc += (double)crit_value(cr_kind, v - 1.f);
        if (grad) {
          const float k = cr_weight * crit_grad(cr_kind, v - 1.f) * v / (float)total;
          // dv/dA = v (-2 AC / a + 2 AD / d), dv/dB = v (-2 BD / b + 2 BC / c), dv/dC = v (2 AC / a - 2 BC / c),
          // dv/dD = v (2 BD / b - 2 AD / d)
          const float g[4][2] = {{k * (-2.f * ACx / a + 2.f * ADx / d), k * (-2.f * ACy / a + 2.f * ADy / d)},
                                 {k * (-2.f * BDx / bb + 2.f * BCx / c), k * (-2.f * BDy / bb + 2.f * BCy / c)},
                                 {k * (2.f * ACx / a - 2.f * BCx / c), k * (2.f * ACy / a - 2.f * BCy / c)},
                                 {k * (2.f * BDx / bb - 2.f * ADx / d), k * (2.f * BDy / bb - 2.f * ADy / d)}};
          for (int i = 0; i < 4; ++i) {
            float* gp = grad + ((size_t)b * K + cr_idx[4 * l + i]) * 2;
            atomicAdd(gp, g[i][0]);
            atomicAdd(gp + 1, g[i][1]);
          }
        }
      }
      cr = block_sum(acc, sm) / total;
    }
  }
  if (threadIdx.x == 0) {
    loss_out[0] = (float)((double)coor_weight * coor + (double)cr_weight * cr);
    loss_out[1] = (float)coor;
    loss_out[2] = (float)cr;
  }
}

}  // namespace egn

extern "C" int egn_mse_hm_fwd_bwd(const float* pred, const float* target, const float* target_weight, int B, int K,
                                  int H, int W, float* loss_out, float* grad_out, void* workspace8, void* stream) {
  using namespace egn;
  EGN_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0, "egn_mse_hm_fwd_bwd: bad shape");
  EGN_REQUIRE(pred && target && loss_out && workspace8, "egn_mse_hm_fwd_bwd: null pointer");
  if (int rc = require_device()) return rc;
  cudaStream_t st = as_stream(stream);
  double* acc = static_cast<double*>(workspace8);
  EGN_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const int64_t n = (int64_t)B * K * H * W;
  const double scale = 0.5 / (double)n;      // 0.5 * mean over B*H*W, then mean over K
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  mse_hm_kernel<<<blocks, 256, 0, st>>>(pred, target, target_weight, n, H * W, scale, grad_out, acc);
  EGN_LAUNCH_CHECK("mse_hm_kernel");
  mse_finalize_kernel<<<1, 1, 0, st>>>(acc, loss_out);
  EGN_LAUNCH_CHECK("mse_finalize_kernel");
  return EGN_OK;
}

extern "C" int egn_generate_target(const double* joints, const float* joints_vis, int N, int K, int input_size0,
                                   int input_size1, int heatmap_size0, int heatmap_size1, double sigma, float* target,
                                   float* target_weight, void* stream) {
  using namespace egn;
  EGN_REQUIRE(N >= 0 && K > 0 && input_size0 > 0 && input_size1 > 0 && heatmap_size0 > 0 && heatmap_size1 > 0 && sigma > 0,
              "egn_generate_target: bad shape / sigma");
  EGN_REQUIRE(N == 0 || (joints && target), "egn_generate_target: null pointer");
  EGN_REQUIRE((int64_t)N * K <= 0x7fffffff, "egn_generate_target: too many maps");
  if (int rc = require_device()) return rc;
  if (N == 0) return EGN_OK;
  generate_target_kernel<<<N * K, 256, 0, as_stream(stream)>>>(joints, joints_vis, (double)input_size0, (double)input_size1,
                                                               heatmap_size0, heatmap_size1, sigma, target, target_weight);
  EGN_LAUNCH_CHECK("generate_target_kernel");
  return EGN_OK;
}

extern "C" int egn_coord_loss_fwd_bwd(const float* coords_pred, const float* coords_gt_px, int B, int K, float img_w,
                                      float img_h, int coor_kind, float coor_weight, const int32_t* cr_indices, int L,
                                      int cr_kind, float cr_weight, float target_cr, float cr_threshold, float* loss_out,
                                      float* grad_out, void* stream) {
  using namespace egn;
  EGN_REQUIRE(B > 0 && K > 0 && img_w > 0 && img_h > 0, "egn_coord_loss_fwd_bwd: bad shape");
  EGN_REQUIRE(coords_pred && coords_gt_px && loss_out, "egn_coord_loss_fwd_bwd: null pointer");
  EGN_REQUIRE(coor_kind >= 0 && coor_kind <= 2 && cr_kind >= 0 && cr_kind <= 2, "egn_coord_loss_fwd_bwd: criterion 0 (mse), 1 (sl1) or 2 (l1)");
  EGN_REQUIRE(L >= 0 && (L == 0 || cr_indices || cr_weight == 0.f), "egn_coord_loss_fwd_bwd: cr_indices missing");
  EGN_REQUIRE(cr_weight == 0.f || target_cr != 0.f, "egn_coord_loss_fwd_bwd: target_cr must be non-zero");
  if (int rc = require_device()) return rc;
  coord_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(coords_pred, coords_gt_px, B, K, img_w, img_h, coor_kind, coor_weight,
                                                      cr_indices, L, cr_kind, cr_weight, target_cr, cr_threshold, loss_out,
                                                      grad_out);
  EGN_LAUNCH_CHECK("coord_loss_kernel");
  return EGN_OK;
}
