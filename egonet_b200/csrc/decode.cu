// Heat-map decoders: hard arg-max and the two soft-argmax variants.
//
// Reference semantics (upstream libs/common/img_proc.py):
//   get_max_preds   :608-637  numpy argmax over the flattened H*W map (first
//                             occurrence on ties, NaN counts as the maximum),
//                             x = idx % W, y = floor(idx / W), both zeroed where
//                             max <= 0.
//   soft_arg_max    :678-707  softmax over H*W, expectation of x and y, raw max.
//   soft_arg_max_np :639-676  sum-normalised expectation, zeroed where max <= 0.
//
// One CTA per (crop, joint) map: the 4096-element map is read once with
// coalesced float4 loads, reduced with warp shuffles, and the per-map scalars
// are written by thread 0.  HBM-bound: 16 KB in, 12-16 B out per map.
#include "common.h"

namespace egn {

constexpr int kDecodeThreads = 256;

struct ArgMax {
  float v;
  int i;
};

// numpy ordering: NaN beats everything, otherwise larger value, ties -> lower index
__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  const bool na = va != va, nb = vb != vb;
  if (na || nb) {
    if (na && nb) return ia < ib;
    return na;
  }
  if (va > vb) return true;
  if (va < vb) return false;
  return ia < ib;
}

__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float v = __shfl_xor_sync(0xffffffffu, a.v, o);
    int i = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (better(v, i, a.v, a.i)) {
      a.v = v;
      a.i = i;
    }
  }
  return a;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ ArgMax block_argmax(ArgMax a, ArgMax* smem) {
  a = warp_argmax(a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = a;
  __syncthreads();
  if (warp == 0) {
    ArgMax b = lane < (kDecodeThreads / 32) ? smem[lane] : ArgMax{-INFINITY, 0x7fffffff};
    b = warp_argmax(b);
    if (lane == 0) smem[0] = b;
  }
  __syncthreads();
  ArgMax r = smem[0];
  __syncthreads();
  return r;
}

template <typename T>
__device__ T block_sum(T v, T* smem) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T b = lane < (kDecodeThreads / 32) ? smem[lane] : T(0);
    b = warp_sum(b);
    if (lane == 0) smem[0] = b;
  }
  __syncthreads();
  T r = smem[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ ArgMax scan_argmax(const float* __restrict__ m, int n) {
  ArgMax best{-INFINITY, 0x7fffffff};
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(m) & 15) == 0);
  if (vec) {
    const float4* m4 = reinterpret_cast<const float4*>(m);
    for (int i = threadIdx.x; i < n / 4; i += kDecodeThreads) {
      const float4 v = __ldg(m4 + i);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (better(e[j], 4 * i + j, best.v, best.i)) best = ArgMax{e[j], 4 * i + j};
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kDecodeThreads) {
      const float v = __ldg(m + i);
      if (better(v, i, best.v, best.i)) best = ArgMax{v, i};
    }
  }
  return best;
}

__global__ void __launch_bounds__(kDecodeThreads)
argmax2d_kernel(const float* __restrict__ hm, int H, int W, int32_t* __restrict__ idx,
                float* __restrict__ preds, float* __restrict__ maxvals) {
  __shared__ ArgMax sm[kDecodeThreads / 32];
  const int n = H * W;
  const float* m = hm + (size_t)blockIdx.x * n;
  ArgMax r = block_argmax(scan_argmax(m, n), sm);
  if (threadIdx.x == 0) {
    if (idx) idx[blockIdx.x] = r.i;
    const float keep = r.v > 0.0f ? 1.0f : 0.0f;  // NaN > 0 is false, like numpy
    // reference casts the index to float32 before % and / (img_proc.py:626-629)
    const float fi = (float)r.i;
    preds[2 * blockIdx.x + 0] = fmodf(fi, (float)W) * keep;
    preds[2 * blockIdx.x + 1] = floorf(fi / (float)W) * keep;
    maxvals[blockIdx.x] = r.v;
  }
}

// mode 0: softmax-normalised; mode 1: sum-normalised with max>0 mask.
__global__ void __launch_bounds__(kDecodeThreads)
soft_argmax2d_kernel(const float* __restrict__ hm, int H, int W, int mode,
                     float* __restrict__ preds, float* __restrict__ maxvals) {
  __shared__ ArgMax sm[kDecodeThreads / 32];
  __shared__ float sf[kDecodeThreads / 32];
  const int n = H * W;
  const float* m = hm + (size_t)blockIdx.x * n;
  const ArgMax r = block_argmax(scan_argmax(m, n), sm);
  const float mx = r.v;
  float s = 0.f, sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < n; i += kDecodeThreads) {
    const float v = __ldg(m + i);  // second pass hits L1/L2 (16 KB map)
    const float p = mode == 0 ? expf(v - mx) : v;
    const int y = i / W, x = i - y * W;
    s += p;
    sx += p * (float)x;
    sy += p * (float)y;
  }
  s = block_sum(s, sf);
  sx = block_sum(sx, sf);
  sy = block_sum(sy, sf);
  if (threadIdx.x == 0) {
    float px = sx / s, py = sy / s;
    if (mode == 1 && !(mx > 0.0f)) {
      // reference multiplies by the 0/1 mask, so NaN/inf of a degenerate map stays NaN
      px = px * 0.0f;
      py = py * 0.0f;
    }
    preds[2 * blockIdx.x + 0] = px;
    preds[2 * blockIdx.x + 1] = py;
    maxvals[blockIdx.x] = mx;
  }
}

}  // namespace egn

extern "C" {

int egn_argmax2d(const float* hm, int B, int K, int H, int W, int32_t* idx, float* preds,
                 float* maxvals, void* stream) {
  using namespace egn;
  EGN_REQUIRE(B >= 0 && K > 0 && H > 0 && W > 0, "egn_argmax2d: bad shape B=%d K=%d H=%d W=%d", B, K, H, W);
  EGN_REQUIRE(B == 0 || (hm && preds && maxvals), "egn_argmax2d: null pointer");
  EGN_REQUIRE((int64_t)H * W < (1 << 24), "egn_argmax2d: map too large for exact float32 indices");
  if (int rc = require_device()) return rc;
  if (B * K == 0) return EGN_OK;
  argmax2d_kernel<<<B * K, kDecodeThreads, 0, as_stream(stream)>>>(hm, H, W, idx, preds, maxvals);
  EGN_LAUNCH_CHECK("argmax2d_kernel");
  return EGN_OK;
}

int egn_soft_argmax2d(const float* hm, int B, int K, int H, int W, int mode, float* preds,
                      float* maxvals, void* stream) {
  using namespace egn;
  EGN_REQUIRE(B >= 0 && K > 0 && H > 0 && W > 0, "egn_soft_argmax2d: bad shape");
  EGN_REQUIRE(B == 0 || (hm && preds && maxvals), "egn_soft_argmax2d: null pointer");
  EGN_REQUIRE(mode == EGN_SOFTARGMAX_SOFTMAX || mode == EGN_SOFTARGMAX_SUM,
              "egn_soft_argmax2d: unknown mode %d", mode);
  if (int rc = require_device()) return rc;
  if (B * K == 0) return EGN_OK;
  soft_argmax2d_kernel<<<B * K, kDecodeThreads, 0, as_stream(stream)>>>(hm, H, W, mode, preds,
                                                                       maxvals);
  EGN_LAUNCH_CHECK("soft_argmax2d_kernel");
  return EGN_OK;
}

}  // extern "C"
