// Heat-map MSE loss, forward + gradient in one pass (the loss of the reference's IGR training loop).
//
// upstream: JointsMSELoss.forward libs/loss/function.py:28-46 and JointsCompositeLoss.calc_hm_loss
// libs/loss/function.py:95-111:  loss = (1/K) * sum_k 0.5 * mean_{b,h,w} (w_bk * (pred - gt))^2,
// with w = target_weight[b,k] when use_target_weight (else 1).  d loss / d pred = w^2 (pred - gt) / (K*B*H*W).
// This is only the loss end of SURVEY.md section 8a row a12; back-propagation through HC (train-mode
// BatchNorm, dgrad / wgrad kernels) is not built.
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
