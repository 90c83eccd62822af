// Training step of the HC network (BASELINE configs[3]): train-mode forward
// (convolution -> BatchNorm with batch statistics -> ReLU / residual add / fuse sums), backward (BatchNorm,
// ReLU, fuse, conv data- and weight-gradients) and the optimiser update, all on the device.
//
// upstream: libs/trainer/trainer.py:183-198 (zero_grad -> forward -> loss -> backward -> step),
//           libs/model/heatmapModel/hrnet.py:63-133, 282-300, 563-614 in train mode (nn.BatchNorm2d momentum 0.1),
//           libs/loss/function.py:28-46 (heat-map MSE, egn_mse_hm_fwd_bwd in loss.cu),
//           libs/optimizer/optimizer.py:9-41 (Adam / SGD).
//
// Arithmetic: fp32 storage and fp32 CUDA-core kernels (the reference trains in fp32); per-channel BatchNorm
// sums are accumulated in fp64.  Activations are NHWC with channels padded to 16 (pad lanes zero).  Parameters
// and gradients live in ONE flat fp32 device buffer each, in state_dict order, owned by the caller (the Python
// mirror points its nn.Parameters into it), so any optimiser -- torch's or egn_adam_step / egn_sgd_step below --
// updates them in place.
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "hrnet_graph.h"
#include "kernels.h"
#include "simt_gemm.cuh"

namespace egn {

// ---------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------
// fp32 NCHW [B,C,H,W] -> NHWC [B,H,W,Cp], pad lanes zero
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W,
                                        int Cp) {
  const int64_t total = (int64_t)B * H * W * Cp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cp);
    int64_t r = e / Cp;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    out[e] = c < C ? x[(((int64_t)b * C + c) * H + h) * W + w] : 0.f;
  }
}

// ---------------------------------------------------------------------------
// per-channel reductions over the pixels of an NHWC tensor (BatchNorm statistics and its backward sums)
//   MODE 0: acc[c] += sum v,  acc[Cp + c] += sum v^2                         (v = a)
//   MODE 1: g = a * (mask > 0 or no mask);  acc[c] += sum g,  acc[Cp + c] += sum g * xhat,
//           xhat = (y - mean) * invstd                                        (a = dout, mask = out)
//   MODE 2: acc[c] += sum a                                                   (bias gradient)
// One thread owns 4 channels and a strided subset of the block's pixels; fp64 partial sums, shared-memory
// then global double atomics.
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const float* __restrict__ a, const float* __restrict__ mask,
                                                             const float* __restrict__ y, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int64_t npix, int Cp,
                                                             int pix_per_block, double* __restrict__ acc) {
  extern __shared__ double sh[];   // [2 * Cp]
  const int quads = Cp >> 2;
  const int lanes = 256 / quads;
  for (int i = threadIdx.x; i < 2 * Cp; i += 256) sh[i] = 0.0;
  __syncthreads();
  const int q = threadIdx.x % quads, lane = threadIdx.x / quads;
  if (lane < lanes) {
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float4 mu = make_float4(0, 0, 0, 0), is = mu;
    if (MODE == 1) {
      mu = *reinterpret_cast<const float4*>(mean + 4 * q);
      is = *reinterpret_cast<const float4*>(invstd + 4 * q);
    }
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    const int64_t p1 = min(npix, p0 + pix_per_block);
    for (int64_t p = p0 + lane; p < p1; p += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(a + p * Cp + 4 * q);
      float g[4] = {v.x, v.y, v.z, v.w};
      if (MODE == 1) {
        if (mask) {
          const float4 m = *reinterpret_cast<const float4*>(mask + p * Cp + 4 * q);
          g[0] = m.x > 0.f ? g[0] : 0.f;
          g[1] = m.y > 0.f ? g[1] : 0.f;
          g[2] = m.z > 0.f ? g[2] : 0.f;
          g[3] = m.w > 0.f ? g[3] : 0.f;
        }
        const float4 yy = *reinterpret_cast<const float4*>(y + p * Cp + 4 * q);
        const float xh[4] = {(yy.x - mu.x) * is.x, (yy.y - mu.y) * is.y, (yy.z - mu.z) * is.z, (yy.w - mu.w) * is.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s0[i] += (double)g[i];
          s1[i] += (double)g[i] * (double)xh[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s0[i] += (double)g[i];
          if (MODE == 0) s1[i] += (double)g[i] * (double)g[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&sh[4 * q + i], s0[i]);
      if (MODE != 2) atomicAdd(&sh[Cp + 4 * q + i], s1[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (MODE == 2 ? Cp : 2 * Cp); i += 256) atomicAdd(&acc[i], sh[i]);
}

// BatchNorm statistics from the sums: batch mean / biased variance (normalisation), running statistics with
// momentum and the unbiased variance (nn.BatchNorm2d, hrnet.py:17 BN_MOMENTUM = 0.1)
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, int Cp, double n, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  if (c >= C) {
    mean[c] = 0.f;
    invstd[c] = 0.f;
    return;
  }
  const double m = acc[c] / n;
  double var = acc[Cp + c] / n - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
    const double unbiased = n > 1 ? var * n / (n - 1.0) : var;
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

// out = [relu]( gamma * (y - mean) * invstd + beta [+ res] ), pad lanes zero
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ res,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int64_t npix, int C, int Cp, int relu, float* __restrict__ out) {
  const int quads = Cp >> 2;
  const int64_t total = npix * quads;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % quads) * 4;
    const int64_t o = (e / quads) * Cp + c;
    const float4 v = *reinterpret_cast<const float4*>(y + o);
    const float in[4] = {v.x, v.y, v.z, v.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (res) {
      const float4 rr = *reinterpret_cast<const float4*>(res + o);
      r[0] = rr.x; r[1] = rr.y; r[2] = rr.z; r[3] = rr.w;
    }
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c + i < C) {
        z[i] = (in[i] - mean[c + i]) * invstd[c + i] * gamma[c + i] + beta[c + i] + r[i];
        if (relu) z[i] = fmaxf(z[i], 0.f);
      } else {
        z[i] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(out + o) = make_float4(z[0], z[1], z[2], z[3]);
  }
}

// backward sums -> parameter gradients (dgamma = sum g xhat, dbeta = sum g) and float copies for the apply pass
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ acc, int C, int Cp, float* __restrict__ sums /*[2*Cp]*/,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  sums[c] = c < C ? (float)acc[c] : 0.f;
  sums[Cp + c] = c < C ? (float)acc[Cp + c] : 0.f;
  if (c < C) {
    dbeta[c] = (float)acc[c];
    dgamma[c] = (float)acc[Cp + c];
  }
}

// dy = gamma * invstd * (g - (sum_g + xhat * sum_gx) / n),  g = dout * (out > 0);  residual branch: dres (+)= g
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ out_mask,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, int64_t npix, int C, int Cp,
                                                           float inv_n, float* __restrict__ dy, float* __restrict__ dres,
                                                           int dres_accumulate) {
  const int quads = Cp >> 2;
  const int64_t total = npix * quads;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % quads) * 4;
    const int64_t o = (e / quads) * Cp + c;
    const float4 dv = *reinterpret_cast<const float4*>(dout + o);
    float g[4] = {dv.x, dv.y, dv.z, dv.w};
    if (out_mask) {
      const float4 m = *reinterpret_cast<const float4*>(out_mask + o);
      g[0] = m.x > 0.f ? g[0] : 0.f;
      g[1] = m.y > 0.f ? g[1] : 0.f;
      g[2] = m.z > 0.f ? g[2] : 0.f;
      g[3] = m.w > 0.f ? g[3] : 0.f;
    }
    const float4 yv = *reinterpret_cast<const float4*>(y + o);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    float d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c + i < C) {
        const float xh = (yy[i] - mean[c + i]) * invstd[c + i];
        d[i] = gamma[c + i] * invstd[c + i] * (g[i] - (sums[c + i] + xh * sums[Cp + c + i]) * inv_n);
      } else {
        d[i] = 0.f;
        g[i] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(dy + o) = make_float4(d[0], d[1], d[2], d[3]);
    if (dres) {
      float4 r = make_float4(g[0], g[1], g[2], g[3]);
      if (dres_accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dres + o);
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      *reinterpret_cast<float4*>(dres + o) = r;
    }
  }
}

// ---------------------------------------------------------------------------
// convolution data gradient: dx[b,ih,iw,ci] (+)= sum_{r,s,co} dy[b,oh,ow,co] * w[co,ci,r,s] with
// oh * stride + r - pad = ih (same for columns).  Implicit GEMM on the register-tiled FFMA core (simt_gemm.cuh):
// M = input pixels (128 per CTA), N = Cin (64 / 128 per CTA), K = taps x Cout; weights in the "dgrad" layout
// [tap][Cout_p][Cin_p].
// ---------------------------------------------------------------------------
struct DgradArgs {
  const float* dy;    // [B, OH, OW, Cout_p]
  const float* w;     // [taps][Cout_p][Cin_p]
  float* dx;          // [B, H, W, Cin_p]
  int B, H, W, Cin_p, OH, OW, Cout_p, ksize, stride, pad, accumulate;
};

template <int GROUPS>
__global__ void __launch_bounds__(SG_THREADS, 2) conv_dgrad_kernel(DgradArgs p) {
  constexpr int BN = 64 * GROUPS;
  __shared__ __align__(16) float As[SG_BK][SG_APITCH];
  __shared__ __align__(16) float Bs[SG_BK][BN];
  const int t = threadIdx.x;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * BN;
  const int lp = t >> 2, lq = t & 3;
  int lb[2], lih[2], liw[2];
  bool lvalid[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t lm = m0 + lp + 64 * h;
    lvalid[h] = lm < M;
    lb[h] = lih[h] = liw[h] = 0;
    if (lvalid[h]) {
      lb[h] = (int)(lm / ((int64_t)p.H * p.W));
      const int r = (int)(lm - (int64_t)lb[h] * p.H * p.W);
      lih[h] = r / p.W;
      liw[h] = r - lih[h] * p.W;
    }
  }
  const int tx = t & 15, ty = t >> 4;
  float acc[8][4 * GROUPS];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * GROUPS; ++j) acc[i][j] = 0.f;
  const int taps = p.ksize * p.ksize;
  const int kslabs = p.Cout_p / SG_BK;
  const int nslab = taps * kslabs;
  float4 a_reg[2], b_reg[GROUPS];
  auto fetch = [&](int slab) {
    const int tap = slab / kslabs, c0 = (slab - tap * kslabs) * SG_BK;
    const int r = tap / p.ksize, q = tap - r * p.ksize;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int nh = lih[h] + p.pad - r, nw = liw[h] + p.pad - q;
      const int oh = nh / p.stride, ow = nw / p.stride;
      const bool pv = lvalid[h] && nh >= 0 && nw >= 0 && oh * p.stride == nh && ow * p.stride == nw && oh < p.OH && ow < p.OW;
      a_reg[h] = pv ? *reinterpret_cast<const float4*>(p.dy + (((int64_t)lb[h] * p.OH + oh) * p.OW + ow) * p.Cout_p + c0 + lq * 4)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* wt = p.w + ((size_t)tap * p.Cout_p + c0) * p.Cin_p;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int idx = t + g * SG_THREADS;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      b_reg[g] = n0 + col < p.Cin_p ? __ldg(reinterpret_cast<const float4*>(wt + (size_t)row * p.Cin_p + n0 + col))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[lq * 4 + 0][lp + 64 * h] = a_reg[h].x;
      As[lq * 4 + 1][lp + 64 * h] = a_reg[h].y;
      As[lq * 4 + 2][lp + 64 * h] = a_reg[h].z;
      As[lq * 4 + 3][lp + 64 * h] = a_reg[h].w;
    }
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int idx = t + g * SG_THREADS;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[row][col]) = b_reg[g];
    }
  };
  fetch(0);
  stage();
  __syncthreads();
  for (int slab = 0; slab < nslab; ++slab) {
    if (slab + 1 < nslab) fetch(slab + 1);
    sg_slab_fma<GROUPS, 4>(As, Bs, tx, ty, acc);
    __syncthreads();
    if (slab + 1 < nslab) {
      stage();
      __syncthreads();
    }
  }
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int n = n0 + sg_col<4>(tx, g, 0);
    if (n >= p.Cin_p) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + sg_row(ty, i);
      if (m >= M) continue;
      float4 v = make_float4(acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
      float* dst = p.dx + m * p.Cin_p + n;
      if (p.accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dst);
        v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
      }
      *reinterpret_cast<float4*>(dst) = v;
    }
  }
}

// ---------------------------------------------------------------------------
// convolution weight gradient: dw[tap][ci][co] += sum_{b,oh,ow} x[b, oh*stride + r - pad, ow*stride + s - pad, ci] * dy[b,oh,ow,co]
// GEMM with K = output pixels on the same FFMA core.  The taps are folded into M (row m' = tap * Cin_p + ci: the
// A operand of a row is the input pixel shifted by its tap), so 48-channel layers fill 432 of 512 tile rows
// instead of 48 of 128 per tap; N = Cout (64 / 128 per CTA); split-K over blockIdx.z, partial tiles combined with
// fp32 atomics into a zeroed [taps * Cin_p][Cout_p] buffer.
// ---------------------------------------------------------------------------
struct WgradArgs {
  const float* x;     // [B, H, W, Cin_p]
  const float* dy;    // [B, OH, OW, Cout_p]
  float* dw;          // [taps][Cin_p][Cout_p], zeroed
  int B, H, W, Cin_p, OH, OW, Cout_p, ksize, stride, pad;
  int pix_per_split;
};

template <int GROUPS>
__global__ void __launch_bounds__(SG_THREADS, 2) conv_wgrad_kernel(WgradArgs p) {
  constexpr int BN = 64 * GROUPS;
  __shared__ __align__(16) float As[SG_BK][SG_APITCH];   // [pixel][tap * Cin_p + ci]
  __shared__ __align__(16) float Bs[SG_BK][BN];          // [pixel][co]
  const int t = threadIdx.x;
  const int taps = p.ksize * p.ksize;
  const int Mrows = taps * p.Cin_p;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * BN;
  const int64_t Mpix = (int64_t)p.B * p.OH * p.OW;
  const int64_t k_begin = (int64_t)blockIdx.z * p.pix_per_split;
  const int64_t k_end = min(Mpix, k_begin + p.pix_per_split);
  // A-load role: slab pixel t / 32 (+ 8), 4 consecutive rows at (t % 32) * 4; the rows' tap / channel are fixed
  const int apx = t >> 5, am = m0 + (t & 31) * 4;
  const bool arow = am < Mrows;
  const int atap = arow ? am / p.Cin_p : 0, aci = arow ? am - atap * p.Cin_p : 0;
  const int ar = atap / p.ksize, as_ = atap - ar * p.ksize;
  const int tx = t & 15, ty = t >> 4;
  float acc[8][4 * GROUPS];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * GROUPS; ++j) acc[i][j] = 0.f;
  float4 a_reg[2], b_reg[GROUPS];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t m = k0 + apx + 8 * h;
      a_reg[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (arow && m < k_end) {
        const int bi = (int)(m / ((int64_t)p.OH * p.OW));
        const int rem = (int)(m - (int64_t)bi * p.OH * p.OW);
        const int oh = rem / p.OW, ow = rem - oh * p.OW;
        const int ih = oh * p.stride + ar - p.pad, iw = ow * p.stride + as_ - p.pad;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
          a_reg[h] = *reinterpret_cast<const float4*>(p.x + (((int64_t)bi * p.H + ih) * p.W + iw) * p.Cin_p + aci);
      }
    }
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int idx = t + g * SG_THREADS;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      const int64_t m = k0 + row;
      b_reg[g] = (m < k_end && n0 + col < p.Cout_p) ? *reinterpret_cast<const float4*>(p.dy + m * p.Cout_p + n0 + col)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&As[apx + 8 * h][(t & 31) * 4]) = a_reg[h];
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int idx = t + g * SG_THREADS;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[row][col]) = b_reg[g];
    }
  };
  if (k_begin < k_end) {
    fetch(k_begin);
    stage();
    __syncthreads();
    for (int64_t k0 = k_begin; k0 < k_end; k0 += SG_BK) {
      const bool more = k0 + SG_BK < k_end;
      if (more) fetch(k0 + SG_BK);
      sg_slab_fma<GROUPS, 4>(As, Bs, tx, ty, acc);
      __syncthreads();
      if (more) {
        stage();
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + sg_row(ty, i);
    if (m >= Mrows) continue;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = n0 + sg_col<4>(tx, g, j);
        if (co < p.Cout_p) atomicAdd(p.dw + (size_t)m * p.Cout_p + co, acc[i][g * 4 + j]);
      }
  }
}

// ---------------------------------------------------------------------------
// fuse backward: out = relu(sum_j up(term_j)); dterm_j[b,h',w',c] (+)= sum over the 2^sh x 2^sh block of dout * (out > 0)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fuse_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                       float* __restrict__ dterm, int B, int H, int W, int Cp, int shift,
                                                       int accumulate) {
  const int quads = Cp >> 2;
  const int Hs = H >> shift, Ws = W >> shift, span = 1 << shift;
  const int64_t total = (int64_t)B * Hs * Ws * quads;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % quads) * 4;
    int64_t pix = e / quads;
    const int ws = (int)(pix % Ws);
    pix /= Ws;
    const int hs = (int)(pix % Hs);
    const int b = (int)(pix / Hs);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dh = 0; dh < span; ++dh)
      for (int dw = 0; dw < span; ++dw) {
        const int64_t o = ((((int64_t)b * H + (hs << shift) + dh) * W) + (ws << shift) + dw) * Cp + c;
        const float4 g = *reinterpret_cast<const float4*>(dout + o);
        const float4 m = *reinterpret_cast<const float4*>(out + o);
        sum.x += m.x > 0.f ? g.x : 0.f;
        sum.y += m.y > 0.f ? g.y : 0.f;
        sum.z += m.z > 0.f ? g.z : 0.f;
        sum.w += m.w > 0.f ? g.w : 0.f;
      }
    float* dst = dterm + (((int64_t)b * Hs + hs) * Ws + ws) * Cp + c;
    if (accumulate) {
      const float4 old = *reinterpret_cast<const float4*>(dst);
      sum.x += old.x; sum.y += old.y; sum.z += old.z; sum.w += old.w;
    }
    *reinterpret_cast<float4*>(dst) = sum;
  }
}

// ---------------------------------------------------------------------------
// weight repacking between the state_dict layout (OIHW) and the kernels' layouts
// ---------------------------------------------------------------------------
// OIHW -> forward [tap][Cin_p][Cout_p] and dgrad [tap][Cout_p][Cin_p] (pad entries stay zero)
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int Cin_p, int Cout_p,
                                    float* __restrict__ fwd, float* __restrict__ dgrad) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int tp = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int co = (int)(e / ((int64_t)taps * Cin));
    const float v = w[e];
    fwd[((size_t)tp * Cin_p + ci) * Cout_p + co] = v;
    dgrad[((size_t)tp * Cout_p + co) * Cin_p + ci] = v;
  }
}
// [tap][Cin_p][Cout_p] -> OIHW
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, int Cout, int Cin, int taps, int Cin_p, int Cout_p,
                                    float* __restrict__ out) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int tp = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int co = (int)(e / ((int64_t)taps * Cin));
    out[e] = dw[((size_t)tp * Cin_p + ci) * Cout_p + co];
  }
}
__global__ void copy_bias_kernel(const float* __restrict__ b, int C, int Cp, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < Cp) out[c] = c < C ? b[c] : 0.f;
}
__global__ void bias_grad_kernel(const double* __restrict__ acc, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}

// ---------------------------------------------------------------------------
// coordinate head tail (hrnet.py:457-458, 607-608): a valid kh x kw conv over the whole kh x kw map (= a linear
// layer over L = kh*kw*Cp inputs) + bias + sigmoid.  Forward is head_tail_kernel (conv_simt.cu); backward here.
// ---------------------------------------------------------------------------
// OIHW [Cout][Cin][kh*kw] -> [Cout][tap][Cp] (pad lanes stay zero)
__global__ void pack_tail_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int Cp, float* __restrict__ out) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int tp = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int co = (int)(e / ((int64_t)taps * Cin));
    out[((size_t)co * taps + tp) * Cp + ci] = w[e];
  }
}
// dlogit = dcoords * c * (1 - c)
__global__ void tail_dlogit_kernel(const float* __restrict__ dcoords, const float* __restrict__ coords, int n,
                                   float* __restrict__ dlogit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dlogit[i] = dcoords[i] * coords[i] * (1.f - coords[i]);
}
// dW[j][tap][ci] (written straight to OIHW), dbias[j]
__global__ void tail_wgrad_kernel(const float* __restrict__ dlogit, const float* __restrict__ xin, int B, int Cout, int Cin,
                                  int taps, int Cp, float* __restrict__ dw_oihw, float* __restrict__ dbias) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int tp = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int j = (int)(e / ((int64_t)taps * Cin));
    const int L = taps * Cp;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dlogit[(size_t)b * Cout + j], xin[(size_t)b * L + tp * Cp + ci], s);
    dw_oihw[e] = s;
    if (tp == 0 && ci == 0) {
      float sb = 0.f;
      for (int b = 0; b < B; ++b) sb += dlogit[(size_t)b * Cout + j];
      dbias[j] = sb;
    }
  }
}
// dxin[b][e] (+)= sum_j dlogit[b][j] * W[j][e]
__global__ void tail_dgrad_kernel(const float* __restrict__ dlogit, const float* __restrict__ w, int B, int Cout, int L,
                                  float* __restrict__ dxin, int accumulate) {
  const int64_t total = (int64_t)B * L;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / L), i = (int)(e % L);
    float s = 0.f;
    for (int j = 0; j < Cout; ++j) s = fmaf(dlogit[(size_t)b * Cout + j], w[(size_t)j * L + i], s);
    dxin[e] = accumulate ? dxin[e] + s : s;
  }
}

// ---------------------------------------------------------------------------
// optimisers over flat buffers (libs/optimizer/optimizer.py:9-41: torch.optim.Adam / SGD semantics)
// ---------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const uint8_t* __restrict__ trainable, int64_t n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bc1, float bc2_sqrt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (trainable && !trainable[i]) continue;
    float gi = g[i];
    if (weight_decay != 0.f) gi += weight_decay * p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    // torch.optim.Adam: step_size = lr / bias_correction1; denom = sqrt(v) / sqrt(bias_correction2) + eps
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                           const uint8_t* __restrict__ trainable, int64_t n, float lr, float momentum, float weight_decay,
                           int nesterov, int first_step) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (trainable && !trainable[i]) continue;
    float gi = g[i];
    if (weight_decay != 0.f) gi += weight_decay * p[i];
    if (momentum != 0.f) {
      const float b = first_step ? gi : momentum * buf[i] + gi;
      buf[i] = b;
      gi = nesterov ? gi + momentum * b : b;
    }
    p[i] -= lr * gi;
  }
}

static int grid_for(int64_t work_items, int threads = 256) {
  return (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div64(work_items, threads)), 148 * 8);
}

}  // namespace egn

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
struct egn_hrnet_train {
  egn_hrnet* g = nullptr;                  // op graph + state_dict inventory (never finalized: no folded weights)
  std::vector<int64_t> key_offset;         // per state_dict entry: element offset in the flat buffer, -1 = not a float entry
  std::vector<uint8_t> key_trainable;      // 1 = parameter (has a gradient), 0 = buffer (running statistics)
  int64_t flat_size = 0;
  struct ConvT {
    int64_t w_off = -1, bias_off = -1, gamma_off = -1, beta_off = -1, rm_off = -1, rv_off = -1;
    int Cin = 0, Cout = 0, Cin_p = 0, Cout_p = 0, taps = 1, ksize = 1;
    float *w_fwd = nullptr, *w_dgrad = nullptr;    // packed every forward from the flat buffer
    float *bias_p = nullptr;                        // [Cout_p] zero (BN layers) or the padded conv bias
    float *mean = nullptr, *invstd = nullptr;      // [Cout_p] batch statistics of the last forward
    float *sums = nullptr;                         // [2*Cout_p] backward sums (float copies)
    double* acc = nullptr;                          // [2*Cout_p] fp64 reduction scratch
  };
  std::vector<ConvT> convs;                // indexed by op index (unused entries for FUSE ops)
  float* dw_scratch = nullptr;             // largest [taps][Cin_p][Cout_p]
  size_t dw_scratch_elems = 0;
  // coordinate head (head_type 'coordinates'): linspace coordinate maps, the tail's packed weights, per-batch
  // coords / logits / dlogit scratch (sized at the first forward with a larger batch)
  float *d_xs = nullptr, *d_ys = nullptr, *w_tail = nullptr, *tail_io = nullptr;
  int tail_io_batch = 0;
  int head1_op = -1, tail_op = -1;
  // workspace plan (per-crop element offsets; multiplied by the batch at run time)
  std::vector<int64_t> out_off, grad_off;  // per tensor
  std::vector<int64_t> y_off;              // per op (pre-BN conv output), -1 when the op has none
  int64_t x16_off = 0, dy_off = 0, ws_per_crop = 0;
  int in_cp = 16;
  int last_batch = 0;
};

namespace egn {

static int64_t flat_index(const egn_hrnet_train* t, const std::string& key) {
  const auto& keys = t->g->keys;
  auto it = std::find(keys.begin(), keys.end(), key);
  return it == keys.end() ? -1 : t->key_offset[it - keys.begin()];
}

static void free_train(egn_hrnet_train* t) {
  for (auto& c : t->convs) {
    cudaFree(c.w_fwd); cudaFree(c.w_dgrad); cudaFree(c.bias_p); cudaFree(c.mean); cudaFree(c.invstd);
    cudaFree(c.sums); cudaFree(c.acc);
  }
  cudaFree(t->dw_scratch);
  cudaFree(t->d_xs); cudaFree(t->d_ys); cudaFree(t->w_tail); cudaFree(t->tail_io);
  if (t->g) egn_hrnet_destroy(t->g);
}

// device buffers that do not depend on the batch size (allocated lazily on the first forward: needs a device)
static int ensure_device_buffers(egn_hrnet_train* t) {
  if (t->dw_scratch) return EGN_OK;
  size_t biggest = 0;
  for (size_t i = 0; i < t->g->ops.size(); ++i) {
    const Op& op = t->g->ops[i];
    if (op.kind != Op::CONV && op.kind != Op::STEM) continue;
    egn_hrnet_train::ConvT& c = t->convs[i];
    const size_t wn = (size_t)c.taps * c.Cin_p * c.Cout_p;
    biggest = std::max(biggest, wn);
    EGN_CUDA_CHECK(cudaMalloc(&c.w_fwd, wn * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.w_dgrad, wn * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemset(c.w_fwd, 0, wn * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemset(c.w_dgrad, 0, wn * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.bias_p, c.Cout_p * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemset(c.bias_p, 0, c.Cout_p * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.mean, c.Cout_p * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.invstd, c.Cout_p * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.sums, 2 * c.Cout_p * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&c.acc, 2 * c.Cout_p * sizeof(double)));
  }
  t->dw_scratch_elems = biggest;
  EGN_CUDA_CHECK(cudaMalloc(&t->dw_scratch, biggest * sizeof(float)));
  if (t->tail_op >= 0) {
    const egn_hrnet_cfg& c = t->g->cfg;
    std::vector<float> xs(c.heatmap_w), ys(c.heatmap_h);     // numpy.linspace(0, 1, n).astype(float32), hrnet.py:461-466
    for (int i = 0; i < c.heatmap_w; ++i) xs[i] = c.heatmap_w > 1 ? (float)((double)i / (double)(c.heatmap_w - 1)) : 0.f;
    for (int i = 0; i < c.heatmap_h; ++i) ys[i] = c.heatmap_h > 1 ? (float)((double)i / (double)(c.heatmap_h - 1)) : 0.f;
    if (c.heatmap_w > 1) xs.back() = 1.f;
    if (c.heatmap_h > 1) ys.back() = 1.f;
    EGN_CUDA_CHECK(cudaMalloc(&t->d_xs, xs.size() * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&t->d_ys, ys.size() * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemcpy(t->d_xs, xs.data(), xs.size() * sizeof(float), cudaMemcpyHostToDevice));
    EGN_CUDA_CHECK(cudaMemcpy(t->d_ys, ys.data(), ys.size() * sizeof(float), cudaMemcpyHostToDevice));
    const egn_hrnet_train::ConvT& ct = t->convs[t->tail_op];
    const size_t wn = (size_t)ct.Cout * ct.taps * ct.Cin_p;
    EGN_CUDA_CHECK(cudaMalloc(&t->w_tail, wn * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemset(t->w_tail, 0, wn * sizeof(float)));
  }
  return EGN_OK;
}

static char* train_ws_base(void* workspace) {
  uintptr_t p = reinterpret_cast<uintptr_t>(workspace);
  return reinterpret_cast<char*>((p + 1023) & ~uintptr_t(1023));
}

}  // namespace egn

extern "C" {

int egn_hrnet_train_create(const egn_hrnet_cfg* cfg, egn_hrnet_train** out) {
  using namespace egn;
  EGN_REQUIRE(cfg && out, "egn_hrnet_train_create: null argument");
  egn_hrnet_cfg c = *cfg;
  c.precision = EGN_PREC_FP32;
  c.conv_impl = EGN_CONV_SIMT;
  c.keep_taps = 1;
  egn_hrnet* g = nullptr;
  if (int rc = egn_hrnet_create(&c, &g)) return rc;
  egn_hrnet_train* t = new egn_hrnet_train();
  t->g = g;
  // flat layout: every float state_dict entry in order (num_batches_tracked is an int64 scalar kept by the caller)
  int64_t off = 0;
  for (size_t i = 0; i < g->keys.size(); ++i) {
    const std::string& k = g->keys[i];
    const bool nbt = k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0;
    if (nbt) {
      t->key_offset.push_back(-1);
      t->key_trainable.push_back(0);
      continue;
    }
    int64_t n = 1;
    for (int64_t d : g->key_shapes[i]) n *= d;
    t->key_offset.push_back(off);
    const bool stat = k.find("running_mean") != std::string::npos || k.find("running_var") != std::string::npos;
    t->key_trainable.push_back(stat ? 0 : 1);
    off += (n + 3) & ~(int64_t)3;          // 16-byte aligned entries
  }
  t->flat_size = off;
  t->in_cp = round_up(c.in_channels, 16);
  // per-conv parameter offsets
  t->convs.resize(g->ops.size());
  for (size_t i = 0; i < g->ops.size(); ++i) {
    const Op& op = g->ops[i];
    if (op.kind == Op::TAIL) {
      const ConvWeights& w = g->weights[op.wi];
      egn_hrnet_train::ConvT& ct = t->convs[i];
      ct.Cin = w.Cin; ct.Cout = w.Cout; ct.ksize = 0;
      ct.taps = (w.k / 1000) * (w.k % 1000);            // kh * kw (encoded by the graph builder)
      ct.Cin_p = g->tensors[op.in].Cp;
      ct.Cout_p = w.Cout;
      ct.w_off = flat_index(t, w.conv_key + ".weight");
      ct.bias_off = flat_index(t, w.conv_key + ".bias");
      t->tail_op = (int)i;
      continue;
    }
    if (op.kind != Op::CONV && op.kind != Op::STEM) continue;
    if (op.write_heatmap) t->head1_op = (int)i;
    const ConvWeights& w = g->weights[op.wi];
    egn_hrnet_train::ConvT& ct = t->convs[i];
    ct.Cin = w.Cin; ct.Cout = w.Cout; ct.ksize = w.k; ct.taps = w.k * w.k;
    ct.Cin_p = op.kind == Op::STEM ? t->in_cp : w.Cin_p;
    ct.Cout_p = g->tensors[op.out].Cp;
    ct.w_off = flat_index(t, w.conv_key + ".weight");
    if (w.has_bias) ct.bias_off = flat_index(t, w.conv_key + ".bias");
    if (!w.bn_key.empty()) {
      ct.gamma_off = flat_index(t, w.bn_key + ".weight");
      ct.beta_off = flat_index(t, w.bn_key + ".bias");
      ct.rm_off = flat_index(t, w.bn_key + ".running_mean");
      ct.rv_off = flat_index(t, w.bn_key + ".running_var");
    }
    if (ct.w_off < 0 || (!w.bn_key.empty() && (ct.gamma_off < 0 || ct.beta_off < 0 || ct.rm_off < 0 || ct.rv_off < 0))) {
      set_error("egn_hrnet_train_create: parameter inventory is missing an entry of '%s'", w.conv_key.c_str());
      free_train(t);
      delete t;
      return EGN_ERR_STATE;
    }
  }
  // workspace plan: every activation, every pre-BN conv output and every activation gradient gets its own buffer
  // (backward needs them all); the network input (padded NHWC) and one dy scratch of the largest conv output
  auto align = [](int64_t n) { return ceil_div64(n, 512) * 512; };
  int64_t top = 0;
  t->x16_off = top;
  top += align((int64_t)c.input_h * c.input_w * t->in_cp);
  t->out_off.assign(g->tensors.size(), -1);
  t->grad_off.assign(g->tensors.size(), -1);
  t->y_off.assign(g->ops.size(), -1);
  int64_t biggest = 0;
  for (size_t id = 0; id < g->tensors.size(); ++id) {
    const int64_t n = align((int64_t)g->tensors[id].H * g->tensors[id].W * g->tensors[id].Cp);
    t->out_off[id] = top;
    top += n;
    t->grad_off[id] = top;
    top += n;
    biggest = std::max(biggest, n);
  }
  for (size_t i = 0; i < g->ops.size(); ++i) {
    const Op& op = g->ops[i];
    if ((op.kind == Op::CONV || op.kind == Op::STEM) && t->convs[i].gamma_off >= 0) {
      t->y_off[i] = top;
      top += align((int64_t)g->tensors[op.out].H * g->tensors[op.out].W * g->tensors[op.out].Cp);
    }
  }
  t->dy_off = top;
  top += biggest;
  t->ws_per_crop = top;
  *out = t;
  return EGN_OK;
}

void egn_hrnet_train_destroy(egn_hrnet_train* t) {
  if (!t) return;
  egn::free_train(t);
  delete t;
}

int64_t egn_hrnet_train_flat_size(const egn_hrnet_train* t) { return t ? t->flat_size : 0; }

int64_t egn_hrnet_train_param_offset(const egn_hrnet_train* t, int i) {
  if (!t || i < 0 || i >= (int)t->key_offset.size()) return -1;
  return t->key_offset[i];
}

int egn_hrnet_train_param_trainable(const egn_hrnet_train* t, int i) {
  if (!t || i < 0 || i >= (int)t->key_trainable.size()) return 0;
  return t->key_trainable[i];
}

size_t egn_hrnet_train_workspace_bytes(const egn_hrnet_train* t, int batch) {
  if (!t || batch <= 0) return 0;
  return (size_t)t->ws_per_crop * (size_t)batch * sizeof(float) + 1024;
}

int64_t egn_hrnet_train_flops_per_sample(const egn_hrnet_train* t) {
  // forward + data gradient + weight gradient of every convolution: 3 x 2 x MACs
  return t ? 6 * t->g->macs : 0;
}

int egn_hrnet_forward_train(egn_hrnet_train* t, float* flat_params, const float* x, int batch, float* heatmap_out,
                            float* coords_out, float momentum, int update_running_stats, void* workspace,
                            size_t workspace_bytes, void* stream) {
  using namespace egn;
  EGN_REQUIRE(t && flat_params && x && heatmap_out, "egn_hrnet_forward_train: null argument");
  EGN_REQUIRE(!coords_out || t->tail_op >= 0, "egn_hrnet_forward_train: coords_out needs the coordinate head");
  EGN_REQUIRE(batch > 0, "egn_hrnet_forward_train: batch must be positive");
  if (int rc = require_device()) return rc;
  if (!workspace || workspace_bytes < egn_hrnet_train_workspace_bytes(t, batch)) {
    set_error("workspace too small: need %zu bytes for a training batch of %d", egn_hrnet_train_workspace_bytes(t, batch), batch);
    return EGN_ERR_WORKSPACE;
  }
  if (int rc = ensure_device_buffers(t)) return rc;
  egn_hrnet* g = t->g;
  cudaStream_t st = as_stream(stream);
  float* base = reinterpret_cast<float*>(train_ws_base(workspace));
  auto act = [&](int id) { return base + (size_t)t->out_off[id] * batch; };
  const egn_hrnet_cfg& c = g->cfg;
  float* x16 = base + (size_t)t->x16_off * batch;
  nchw_to_nhwc_pad_kernel<<<grid_for((int64_t)batch * c.input_h * c.input_w * t->in_cp), 256, 0, st>>>(
      x, x16, batch, c.in_channels, c.input_h, c.input_w, t->in_cp);
  EGN_LAUNCH_CHECK("nchw_to_nhwc_pad_kernel");
  for (size_t i = 0; i < g->ops.size(); ++i) {
    const Op& op = g->ops[i];
    if (op.kind == Op::FUSE) {
      const TensorInfo& to = g->tensors[op.out];
      FuseArgs a{};
      a.out = act(op.out);
      a.nterms = op.nterms;
      for (int j = 0; j < op.nterms; ++j) {
        a.term[j] = act(op.term[j]);
        a.shift[j] = op.shift[j];
      }
      a.B = batch; a.H = to.H; a.W = to.W; a.Cp = to.Cp;
      if (int rc = launch_fuse(Dtype::F32, a, st)) return rc;
      continue;
    }
    if (op.kind == Op::TAIL) {
      egn_hrnet_train::ConvT& ct = t->convs[i];
      if (t->tail_io_batch < batch) {
        cudaFree(t->tail_io);
        t->tail_io = nullptr;
        EGN_CUDA_CHECK(cudaMalloc(&t->tail_io, (size_t)3 * batch * ct.Cout * sizeof(float)));
        t->tail_io_batch = batch;
      }
      pack_tail_kernel<<<grid_for((int64_t)ct.Cout * ct.Cin * ct.taps), 256, 0, st>>>(flat_params + ct.w_off, ct.Cout, ct.Cin,
                                                                                     ct.taps, ct.Cin_p, t->w_tail);
      HeadTailArgs a{};
      a.in = act(op.in);
      a.w = t->w_tail;
      a.bias = flat_params + ct.bias_off;
      a.coords = t->tail_io;                                   // [B, Cout] sigmoid outputs (kept for the backward pass)
      a.logits = t->tail_io + (size_t)batch * ct.Cout;
      a.B = batch;
      a.L = ct.taps * ct.Cin_p;
      a.Cout = ct.Cout;
      a.Cp = ct.Cin_p;
      if (int rc = launch_head_tail(Dtype::F32, a, st)) return rc;
      if (coords_out)
        EGN_CUDA_CHECK(cudaMemcpyAsync(coords_out, t->tail_io, (size_t)batch * ct.Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
      continue;
    }
    if (op.kind != Op::CONV && op.kind != Op::STEM) continue;
    egn_hrnet_train::ConvT& ct = t->convs[i];
    const TensorInfo& to = g->tensors[op.out];
    const bool stem = op.kind == Op::STEM;
    const int H = stem ? c.input_h : g->tensors[op.in].H, W = stem ? c.input_w : g->tensors[op.in].W;
    const int stride = stem ? 2 : op.stride, pad = stem ? 1 : op.pad, relu = stem ? 1 : op.relu;
    // parameters of this step -> kernel layouts
    pack_weights_kernel<<<grid_for((int64_t)ct.Cout * ct.Cin * ct.taps), 256, 0, st>>>(
        flat_params + ct.w_off, ct.Cout, ct.Cin, ct.taps, ct.Cin_p, ct.Cout_p, ct.w_fwd, ct.w_dgrad);
    if (ct.bias_off >= 0)
      copy_bias_kernel<<<ceil_div(ct.Cout_p, 128), 128, 0, st>>>(flat_params + ct.bias_off, ct.Cout, ct.Cout_p, ct.bias_p);
    const bool bn = ct.gamma_off >= 0;
    float* y = bn ? base + (size_t)t->y_off[i] * batch : act(op.out);
    ConvArgs a{};
    a.in = stem ? x16 : act(op.in);
    a.out = y;
    a.res = nullptr;
    a.bias = ct.bias_p;
    a.B = batch; a.H = H; a.W = W; a.Cin_p = ct.Cin_p; a.OH = to.H; a.OW = to.W; a.Cout_p = ct.Cout_p; a.Cout = ct.Cout;
    a.ksize = ct.ksize; a.stride = stride; a.pad = pad; a.relu = 0;
    if (op.coord_maps) {                      // head1: channels Cout, Cout + 1 carry the coordinate maps (hrnet.py:602-606)
      a.coord_maps = 1;
      a.xs = t->d_xs;
      a.ys = t->d_ys;
    }
    if (int rc = launch_conv_simt(Dtype::F32, a, ct.w_fwd, st)) return rc;
    if (!bn) continue;
    const int64_t npix = (int64_t)batch * to.H * to.W;
    EGN_CUDA_CHECK(cudaMemsetAsync(ct.acc, 0, 2 * ct.Cout_p * sizeof(double), st));
    const int lanes = 256 / (ct.Cout_p / 4);
    const int ppb = std::max(lanes * 8, (int)ceil_div64(npix, 148 * 4));
    channel_reduce_kernel<0><<<(unsigned)ceil_div64(npix, ppb), 256, 2 * ct.Cout_p * sizeof(double), st>>>(
        y, nullptr, nullptr, nullptr, nullptr, npix, ct.Cout_p, ppb, ct.acc);
    bn_finalize_kernel<<<ceil_div(ct.Cout_p, 128), 128, 0, st>>>(
        ct.acc, ct.Cout, ct.Cout_p, (double)npix, 1e-5f, momentum, ct.mean, ct.invstd,
        update_running_stats ? flat_params + ct.rm_off : nullptr, update_running_stats ? flat_params + ct.rv_off : nullptr);
    bn_apply_kernel<<<grid_for(npix * (ct.Cout_p / 4)), 256, 0, st>>>(
        y, op.res >= 0 ? act(op.res) : nullptr, ct.mean, ct.invstd, flat_params + ct.gamma_off, flat_params + ct.beta_off,
        npix, ct.Cout, ct.Cout_p, relu, act(op.out));
    EGN_LAUNCH_CHECK("train forward conv + bn");
  }
  // heat-maps: the heat-map conv's NHWC output (its first num_joints channels) -> fp32 NCHW
  const Op& hop = g->ops[t->head1_op];
  const TensorInfo& th = g->tensors[hop.out];
  if (int rc = launch_nhwc_to_nchw(Dtype::F32, act(hop.out), heatmap_out, batch, th.H, th.W, th.Cp, t->convs[t->head1_op].Cout, st))
    return rc;
  t->last_batch = batch;
  return EGN_OK;
}

int egn_hrnet_backward(egn_hrnet_train* t, const float* flat_params, const float* grad_heatmap, const float* grad_coords,
                       int batch, float* flat_grads, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace egn;
  EGN_REQUIRE(t && flat_params && flat_grads && (grad_heatmap || grad_coords), "egn_hrnet_backward: null argument");
  EGN_REQUIRE(!grad_coords || t->tail_op >= 0, "egn_hrnet_backward: grad_coords needs the coordinate head");
  if (t->last_batch != batch || batch <= 0) {
    set_error("egn_hrnet_backward: call egn_hrnet_forward_train with the same batch (%d) and workspace first", batch);
    return EGN_ERR_STATE;
  }
  if (int rc = require_device()) return rc;
  if (!workspace || workspace_bytes < egn_hrnet_train_workspace_bytes(t, batch)) {
    set_error("workspace too small: need %zu bytes for a training batch of %d", egn_hrnet_train_workspace_bytes(t, batch), batch);
    return EGN_ERR_WORKSPACE;
  }
  egn_hrnet* g = t->g;
  cudaStream_t st = as_stream(stream);
  float* base = reinterpret_cast<float*>(train_ws_base(workspace));
  auto act = [&](int id) { return base + (size_t)t->out_off[id] * batch; };
  auto grad = [&](int id) { return base + (size_t)t->grad_off[id] * batch; };
  float* dy_scratch = base + (size_t)t->dy_off * batch;
  float* x16 = base + (size_t)t->x16_off * batch;
  const egn_hrnet_cfg& c = g->cfg;
  EGN_CUDA_CHECK(cudaMemsetAsync(flat_grads, 0, (size_t)t->flat_size * sizeof(float), st));
  std::vector<char> written(g->tensors.size(), 0);     // first writer of a gradient buffer overwrites, later ones add
  // d loss / d heat-maps arrives as fp32 NCHW; it seeds the gradient of the heat-map conv's output (the coordinate
  // head's consumers add to it; its coordinate-map channels are constants and stay zero here)
  {
    const Op& hop = g->ops[t->head1_op];
    const TensorInfo& th = g->tensors[hop.out];
    if (grad_heatmap) {
      nchw_to_nhwc_pad_kernel<<<grid_for((int64_t)batch * th.H * th.W * th.Cp), 256, 0, st>>>(
          grad_heatmap, grad(hop.out), batch, t->convs[t->head1_op].Cout, th.H, th.W, th.Cp);
    } else {
      EGN_CUDA_CHECK(cudaMemsetAsync(grad(hop.out), 0, (size_t)batch * th.H * th.W * th.Cp * sizeof(float), st));
    }
    written[hop.out] = 1;
  }
  for (int i = (int)g->ops.size() - 1; i >= 0; --i) {
    const Op& op = g->ops[i];
    if (op.kind == Op::TAIL) {
      if (!grad_coords) continue;
      egn_hrnet_train::ConvT& ct = t->convs[i];
      const int n = batch * ct.Cout, L = ct.taps * ct.Cin_p;
      float* coords = t->tail_io;
      float* dlogit = t->tail_io + (size_t)2 * batch * ct.Cout;
      tail_dlogit_kernel<<<ceil_div(n, 256), 256, 0, st>>>(grad_coords, coords, n, dlogit);
      tail_wgrad_kernel<<<grid_for((int64_t)ct.Cout * ct.Cin * ct.taps), 256, 0, st>>>(
          dlogit, act(op.in), batch, ct.Cout, ct.Cin, ct.taps, ct.Cin_p, flat_grads + ct.w_off, flat_grads + ct.bias_off);
      tail_dgrad_kernel<<<grid_for((int64_t)batch * L), 256, 0, st>>>(dlogit, t->w_tail, batch, ct.Cout, L, grad(op.in),
                                                                      written[op.in]);
      written[op.in] = 1;
      EGN_LAUNCH_CHECK("coordinate head tail backward");
      continue;
    }
    if (op.out < 0 || !written[op.out]) continue;      // nothing downstream depends on this op
    const TensorInfo& to = g->tensors[op.out];
    if (op.kind == Op::FUSE) {
      for (int j = 0; j < op.nterms; ++j) {
        const int id = op.term[j];
        const TensorInfo& tt = g->tensors[id];
        fuse_bwd_kernel<<<grid_for((int64_t)batch * tt.H * tt.W * (tt.Cp / 4)), 256, 0, st>>>(
            grad(op.out), act(op.out), grad(id), batch, to.H, to.W, to.Cp, op.shift[j], written[id]);
        written[id] = 1;
      }
      EGN_LAUNCH_CHECK("fuse_bwd_kernel");
      continue;
    }
    if (op.kind != Op::CONV && op.kind != Op::STEM) continue;
    egn_hrnet_train::ConvT& ct = t->convs[i];
    const bool stem = op.kind == Op::STEM;
    const int H = stem ? c.input_h : g->tensors[op.in].H, W = stem ? c.input_w : g->tensors[op.in].W;
    const int stride = stem ? 2 : op.stride, pad = stem ? 1 : op.pad, relu = stem ? 1 : op.relu;
    const int64_t npix = (int64_t)batch * to.H * to.W;
    const bool bn = ct.gamma_off >= 0;
    const float* dy = nullptr;
    const int lanes = 256 / (ct.Cout_p / 4);
    const int ppb = std::max(lanes * 8, (int)ceil_div64(npix, 148 * 4));
    EGN_CUDA_CHECK(cudaMemsetAsync(ct.acc, 0, 2 * ct.Cout_p * sizeof(double), st));
    if (bn) {
      const float* y = base + (size_t)t->y_off[i] * batch;
      channel_reduce_kernel<1><<<(unsigned)ceil_div64(npix, ppb), 256, 2 * ct.Cout_p * sizeof(double), st>>>(
          grad(op.out), relu ? act(op.out) : nullptr, y, ct.mean, ct.invstd, npix, ct.Cout_p, ppb, ct.acc);
      bn_bwd_finalize_kernel<<<ceil_div(ct.Cout_p, 128), 128, 0, st>>>(ct.acc, ct.Cout, ct.Cout_p, ct.sums,
                                                                      flat_grads + ct.gamma_off, flat_grads + ct.beta_off);
      float* dres = op.res >= 0 ? grad(op.res) : nullptr;
      bn_bwd_apply_kernel<<<grid_for(npix * (ct.Cout_p / 4)), 256, 0, st>>>(
          grad(op.out), relu ? act(op.out) : nullptr, y, ct.mean, ct.invstd, flat_params + ct.gamma_off, ct.sums, npix,
          ct.Cout, ct.Cout_p, (float)(1.0 / (double)npix), dy_scratch, dres, dres ? (int)written[op.res] : 0);
      if (op.res >= 0) written[op.res] = 1;
      dy = dy_scratch;
    } else {
      // conv + bias only (final layer): dy = dout, dbias = sum dy
      dy = grad(op.out);
      if (ct.bias_off >= 0) {
        channel_reduce_kernel<2><<<(unsigned)ceil_div64(npix, ppb), 256, 2 * ct.Cout_p * sizeof(double), st>>>(
            dy, nullptr, nullptr, nullptr, nullptr, npix, ct.Cout_p, ppb, ct.acc);
        bias_grad_kernel<<<ceil_div(ct.Cout, 128), 128, 0, st>>>(ct.acc, ct.Cout, flat_grads + ct.bias_off);
      }
    }
    // weight gradient
    {
      const size_t wn = (size_t)ct.taps * ct.Cin_p * ct.Cout_p;
      EGN_CUDA_CHECK(cudaMemsetAsync(t->dw_scratch, 0, wn * sizeof(float), st));
      WgradArgs a{};
      a.x = stem ? x16 : act(op.in);
      a.dy = dy;
      a.dw = t->dw_scratch;
      a.B = batch; a.H = H; a.W = W; a.Cin_p = ct.Cin_p; a.OH = to.H; a.OW = to.W; a.Cout_p = ct.Cout_p;
      a.ksize = ct.ksize; a.stride = stride; a.pad = pad;
      const int groups = simt_groups_for(ct.Cout_p);
      const int m_tiles = ceil_div(ct.taps * ct.Cin_p, SG_BM), n_tiles = ceil_div(ct.Cout_p, 64 * groups);
      int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(148 * 4, m_tiles * n_tiles), ceil_div64(npix, 512)));
      a.pix_per_split = (int)(ceil_div64(ceil_div64(npix, splits), SG_BK) * SG_BK);
      splits = ceil_div64(npix, a.pix_per_split);
      dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
      if (groups == 2) conv_wgrad_kernel<2><<<grid, SG_THREADS, 0, st>>>(a);
      else conv_wgrad_kernel<1><<<grid, SG_THREADS, 0, st>>>(a);
      unpack_wgrad_kernel<<<grid_for((int64_t)ct.Cout * ct.Cin * ct.taps), 256, 0, st>>>(
          t->dw_scratch, ct.Cout, ct.Cin, ct.taps, ct.Cin_p, ct.Cout_p, flat_grads + ct.w_off);
    }
    // data gradient (the network input needs none)
    if (!stem) {
      DgradArgs a{};
      a.dy = dy;
      a.w = ct.w_dgrad;
      a.dx = grad(op.in);
      a.B = batch; a.H = H; a.W = W; a.Cin_p = ct.Cin_p; a.OH = to.H; a.OW = to.W; a.Cout_p = ct.Cout_p;
      a.ksize = ct.ksize; a.stride = stride; a.pad = pad; a.accumulate = written[op.in];
      const int groups = simt_groups_for(ct.Cin_p);
      dim3 grid((unsigned)ceil_div64((int64_t)batch * H * W, SG_BM), (unsigned)ceil_div(ct.Cin_p, 64 * groups));
      if (groups == 2) conv_dgrad_kernel<2><<<grid, SG_THREADS, 0, st>>>(a);
      else conv_dgrad_kernel<1><<<grid, SG_THREADS, 0, st>>>(a);
      written[op.in] = 1;
    }
    EGN_LAUNCH_CHECK("train backward conv");
  }
  return EGN_OK;
}

int egn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* trainable_mask,
                  int64_t n, int step, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream) {
  using namespace egn;
  EGN_REQUIRE(n >= 0 && step >= 1, "egn_adam_step: bad n / step (steps count from 1)");
  EGN_REQUIRE(n == 0 || (params && grads && exp_avg && exp_avg_sq), "egn_adam_step: null pointer");
  if (int rc = require_device()) return rc;
  if (n == 0) return EGN_OK;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, trainable_mask, n, lr, beta1,
                                                          beta2, eps, weight_decay, bc1, sqrtf(bc2));
  EGN_LAUNCH_CHECK("adam_kernel");
  return EGN_OK;
}

int egn_sgd_step(float* params, const float* grads, float* momentum_buf, const uint8_t* trainable_mask, int64_t n,
                 int step, float lr, float momentum, float weight_decay, int nesterov, void* stream) {
  using namespace egn;
  EGN_REQUIRE(n >= 0 && step >= 1, "egn_sgd_step: bad n / step (steps count from 1)");
  EGN_REQUIRE(n == 0 || (params && grads && (momentum == 0.f || momentum_buf)), "egn_sgd_step: null pointer");
  if (int rc = require_device()) return rc;
  if (n == 0) return EGN_OK;
  sgd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(params, grads, momentum_buf, trainable_mask, n, lr, momentum,
                                                         weight_decay, nesterov, step == 1);
  EGN_LAUNCH_CHECK("sgd_kernel");
  return EGN_OK;
}

}  // extern "C"
