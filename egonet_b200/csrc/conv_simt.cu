// CUDA-core kernels of the HC network: generic fused conv (implicit GEMM),
// stem conv, cross-resolution fuse, head tail, layout conversion.
//
// These are (1) the fp32 "exact" precision mode that holds the 1e-4 parity bound
// against the reference's fp32 arithmetic, (2) the on-device comparator for the
// tcgen05 kernels in conv_tc.cu (same layouts, same folded weights, same
// epilogue), and (3) the home of the shapes tensor cores do not fit (stem Cin=3,
// the 4x4 head tail).  upstream: libs/model/heatmapModel/hrnet.py (conv+BN+ReLU
// chains :63-133, fuse :282-300, heads :427-467,601-608).
#include "common.h"
#include "kernels.h"
#include "simt_gemm.cuh"

namespace egn {

// ---------------------------------------------------------------------------
// storage policies: how a logical NHWC element (pixel, channel) of a tensor with Cp padded channels is
// loaded / stored.  Plain<T>: one T per element.  Split16: two fp16 planes per pixel, [hi: Cp][lo: Cp]
// (kernels.h, Dtype::F16X2); loads return hi + lo, stores write hi = rn16(v), lo = rn16(v - hi).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void half4_to_float(const uint2 q, float v[4]) {
  const __half2 a = *reinterpret_cast<const __half2*>(&q.x);
  const __half2 b = *reinterpret_cast<const __half2*>(&q.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ uint2 float4_to_half(const float v[4]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]);
  const __half2 b = __floats2half2_rn(v[2], v[3]);
  uint2 q;
  q.x = *reinterpret_cast<const uint32_t*>(&a);
  q.y = *reinterpret_cast<const uint32_t*>(&b);
  return q;
}

template <typename T> struct Plain;
template <> struct Plain<float> {
  using type = float;
  static __device__ __forceinline__ void load4(const void* base, int64_t pix, int Cp, int c, float v[4]) {
    const float4 q = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + pix * Cp + c);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store4(void* base, int64_t pix, int Cp, int c, const float v[4]) {
    *reinterpret_cast<float4*>(static_cast<float*>(base) + pix * Cp + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t pix, int Cp, int c) {
    return static_cast<const float*>(base)[pix * Cp + c];
  }
};
template <> struct Plain<__half> {
  using type = __half;
  static __device__ __forceinline__ void load4(const void* base, int64_t pix, int Cp, int c, float v[4]) {
    half4_to_float(*reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + pix * Cp + c), v);
  }
  static __device__ __forceinline__ void store4(void* base, int64_t pix, int Cp, int c, const float v[4]) {
    *reinterpret_cast<uint2*>(static_cast<__half*>(base) + pix * Cp + c) = float4_to_half(v);
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t pix, int Cp, int c) {
    return __half2float(static_cast<const __half*>(base)[pix * Cp + c]);
  }
};
struct Split16 {
  static __device__ __forceinline__ void load4(const void* base, int64_t pix, int Cp, int c, float v[4]) {
    const __half* p = static_cast<const __half*>(base) + pix * (2 * Cp) + c;
    float lo[4];
    half4_to_float(*reinterpret_cast<const uint2*>(p), v);
    half4_to_float(*reinterpret_cast<const uint2*>(p + Cp), lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += lo[i];
  }
  static __device__ __forceinline__ void store4(void* base, int64_t pix, int Cp, int c, const float v[4]) {
    __half* p = static_cast<__half*>(base) + pix * (2 * Cp) + c;
    const uint2 hi = float4_to_half(v);
    float h[4], lo[4];
    half4_to_float(hi, h);
#pragma unroll
    for (int i = 0; i < 4; ++i) lo[i] = v[i] - h[i];
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + Cp) = float4_to_half(lo);
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t pix, int Cp, int c) {
    const __half* p = static_cast<const __half*>(base) + pix * (2 * Cp) + c;
    return __half2float(p[0]) + __half2float(p[Cp]);
  }
};

// Dispatch a kernel template over the storage policy of `dt`.
#define EGN_DISPATCH_STORAGE(dt, S, ...)                    \
  do {                                                      \
    if ((dt) == Dtype::F32) { using S = Plain<float>; __VA_ARGS__; }        \
    else if ((dt) == Dtype::F16) { using S = Plain<__half>; __VA_ARGS__; }  \
    else { using S = Split16; __VA_ARGS__; }                \
  } while (0)

// ---------------------------------------------------------------------------
// generic conv (implicit GEMM on CUDA cores): M = output pixels, N = output channels, K = taps x input channels.
// 128 pixels x BN (64 / 128) channels per CTA, 8 x 4 / 8 x 8 outputs per thread (simt_gemm.cuh), the next K slab
// prefetched into registers while the current one is multiplied.
// ---------------------------------------------------------------------------
template <typename S, int GROUPS>
__global__ void __launch_bounds__(SG_THREADS, 2)
conv_simt_kernel(ConvArgs p, const float* __restrict__ wp) {
  constexpr int BN = 64 * GROUPS;
  __shared__ __align__(16) float As[SG_BK][SG_APITCH];
  __shared__ __align__(16) float Bs[SG_BK][BN];
  const int t = threadIdx.x;
  if (t == 0) pdl_trigger();
  pdl_wait();
  const int64_t M = (int64_t)p.B * p.OH * p.OW;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * BN;
  // A-load role: pixels lp and lp + 64, channel quad lq
  const int lp = t >> 2, lq = t & 3;
  int lb[2], loh[2], low[2];
  bool lvalid[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t lm = m0 + lp + 64 * h;
    lvalid[h] = lm < M;
    lb[h] = loh[h] = low[h] = 0;
    if (lvalid[h]) {
      lb[h] = (int)(lm / ((int64_t)p.OH * p.OW));
      const int r = (int)(lm - (int64_t)lb[h] * p.OH * p.OW);
      loh[h] = r / p.OW;
      low[h] = r - loh[h] * p.OW;
    }
  }
  const int tx = t & 15, ty = t >> 4;
  float acc[8][4 * GROUPS];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * GROUPS; ++j) acc[i][j] = 0.f;

  const int taps = p.ksize * p.ksize;
  const int kslabs = p.Cin_p / SG_BK;                  // Cin_p is a multiple of 16
  const int nslab = taps * kslabs;
  float a_reg[2][4];
  float4 b_reg[GROUPS];
  auto fetch = [&](int slab) {
    const int tap = slab / kslabs, c0 = (slab - tap * kslabs) * SG_BK;
    const int r = tap / p.ksize, q = tap - r * p.ksize;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ih = loh[h] * p.stride + r - p.pad, iw = low[h] * p.stride + q - p.pad;
      const bool pv = lvalid[h] && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
      a_reg[h][0] = a_reg[h][1] = a_reg[h][2] = a_reg[h][3] = 0.f;
      if (pv) S::load4(p.in, ((int64_t)lb[h] * p.H + ih) * p.W + iw, p.Cin_p, c0 + lq * 4, a_reg[h]);
    }
    const float* wt = wp + ((size_t)tap * p.Cin_p + c0) * p.Cout_p;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int idx = t + g * SG_THREADS;              // float4 index inside the 16 x BN slab
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      b_reg[g] = n0 + col < p.Cout_p ? __ldg(reinterpret_cast<const float4*>(wt + (size_t)row * p.Cout_p + n0 + col))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c) As[lq * 4 + c][lp + 64 * h] = a_reg[h][c];
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

  // epilogue: bias (+ residual) (+ ReLU), 4 consecutive channels per store
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int n = n0 + sg_col<4>(tx, g, 0);
    if (n >= p.Cout_p) continue;
    const float4 bias = *reinterpret_cast<const float4*>(p.bias + n);
    const float bv[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + sg_row(ty, i);
      if (m >= M) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][g * 4 + j] + bv[j];
      if (p.res) {
        float rv[4];
        S::load4(p.res, m, p.Cout_p, n, rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += rv[j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.heatmap || p.coord_maps) {
        const int b = (int)(m / ((int64_t)p.OH * p.OW));
        const int rr = (int)(m - (int64_t)b * p.OH * p.OW);
        const int oh = rr / p.OW, ow = rr - oh * p.OW;
        if (p.heatmap) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.Cout) p.heatmap[(((int64_t)b * p.Cout + n + j) * p.OH + oh) * p.OW + ow] = v[j];
        }
        if (p.coord_maps) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j == p.Cout) v[j] = p.xs[ow];
            if (n + j == p.Cout + 1) v[j] = p.ys[oh];
          }
        }
      }
      S::store4(p.out, m, p.Cout_p, n, v);
    }
  }
}

// N tile of the FFMA kernels: 128-wide tiles (8 x 8 per thread) unless 64-wide ones waste fewer columns
int simt_groups_for(int Cout_p) {
  const int c128 = ceil_div(Cout_p, 128) * 128, c64 = ceil_div(Cout_p, 64) * 64;
  return (double)c128 <= 1.25 * (double)c64 ? 2 : 1;
}

int launch_conv_simt(Dtype dt, const ConvArgs& a, const float* w_packed, cudaStream_t st) {
  const int64_t M = (int64_t)a.B * a.OH * a.OW;
  const int groups = simt_groups_for(a.Cout_p);
  dim3 grid((unsigned)ceil_div64(M, SG_BM), (unsigned)ceil_div(a.Cout_p, 64 * groups));
  if (groups == 2)
    EGN_DISPATCH_STORAGE(dt, S, EGN_CUDA_CHECK(launch_pdl(conv_simt_kernel<S, 2>, grid, dim3(SG_THREADS), 0, st, a, w_packed)));
  else
    EGN_DISPATCH_STORAGE(dt, S, EGN_CUDA_CHECK(launch_pdl(conv_simt_kernel<S, 1>, grid, dim3(SG_THREADS), 0, st, a, w_packed)));
  EGN_LAUNCH_CHECK("conv_simt_kernel");
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// stem conv1: fp32 NCHW input, 3x3 s2 p1, Cin in {3,5} -> 64 channels NHWC
// ---------------------------------------------------------------------------
// One CTA = an 8 x 32 tile of output pixels; a warp owns one output row of the tile.  Lane (qc, cg):
// four pixels ox = qc + 8i (i = 0..3) x sixteen channels {16q + 4cg + j}: 64 accumulators per thread,
// so each weight float4 read from shared memory feeds 16 FMAs and each input value 16 (LDS : FFMA = 1 : 8;
// the previous one-pixel-per-thread mapping sat at 1 : 3 and was LDS-issue bound).  Patch reads of a warp
// are stride-2 floats (conflict free, broadcast over cg); weight reads are 64 contiguous bytes
// (broadcast over qc).  Every output accumulates its taps in the order c, r, s starting from 0 --
// the same order as the reference-parity fp32 path always had, so results are unchanged bit for bit.
constexpr int STH = 8, STW = 32;                  // output tile (rows x cols)
constexpr int SPH = 2 * STH + 1, SPW = 2 * STW + 1;   // input patch 17 x 65
constexpr int SPP = SPW + 1;                      // patch row pitch (floats)
constexpr int STEM_THREADS = 256;
constexpr int STEM_MAX_CIN = 5;

template <typename S>
__global__ void __launch_bounds__(STEM_THREADS, 2)
stem_kernel(StemArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* w = sm;                                      // [9*Cin][64]
  float* bias = w + 9 * p.Cin * 64;                   // [64]
  float* patch = bias + 64;                           // [Cin][SPH][SPP]
  const int t = threadIdx.x;
  const int tiles_x = (p.OW + STW - 1) / STW;
  const int b = blockIdx.y;
  const int oy0 = (blockIdx.x / tiles_x) * STH, ox0 = (blockIdx.x % tiles_x) * STW;
  const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
  if (t == 0) pdl_trigger();
  for (int e = t; e < 9 * p.Cin * 16; e += STEM_THREADS)
    reinterpret_cast<float4*>(w)[e] = __ldg(reinterpret_cast<const float4*>(p.w) + e);
  if (t < 64) bias[t] = __ldg(p.bias + t);
  pdl_wait();   // weights staged above do not depend on the previous kernel; the input and the output buffer do
  for (int e = t; e < p.Cin * SPH * SPW; e += STEM_THREADS) {
    const int c = e / (SPH * SPW), r = (e / SPW) % SPH, q = e % SPW;
    const int iy = iy0 + r, ix = ix0 + q;
    float v = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
      v = __ldg(p.x + (((int64_t)b * p.Cin + c) * p.H + iy) * p.W + ix);
    patch[(c * SPH + r) * SPP + q] = v;
  }
  __syncthreads();
  const int cg = t & 3, qc = (t >> 2) & 7, ty = t >> 5;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  for (int c = 0; c < p.Cin; ++c) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* prow = patch + (c * SPH + 2 * ty + r) * SPP + 2 * qc;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        float xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = prow[16 * i + s];
        const float* wr = w + ((r * 3 + s) * p.Cin + c) * 64 + cg * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + q * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][q * 4 + 0] = fmaf(xv[i], wv.x, acc[i][q * 4 + 0]);
            acc[i][q * 4 + 1] = fmaf(xv[i], wv.y, acc[i][q * 4 + 1]);
            acc[i][q * 4 + 2] = fmaf(xv[i], wv.z, acc[i][q * 4 + 2]);
            acc[i][q * 4 + 3] = fmaf(xv[i], wv.w, acc[i][q * 4 + 3]);
          }
        }
      }
    }
  }
  const int oy = oy0 + ty;
  if (oy >= p.OH) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + qc + 8 * i;
    if (ox >= p.OW) continue;
    const int64_t opix = ((int64_t)b * p.OH + oy) * p.OW + ox;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(acc[i][q * 4 + j] + bias[q * 16 + cg * 4 + j], 0.f);
      S::store4(p.out, opix, 64, cg * 4 + q * 16, v);
    }
  }
}

int launch_stem(Dtype dt, const StemArgs& a, cudaStream_t st) {
  EGN_REQUIRE(a.Cin >= 1 && a.Cin <= STEM_MAX_CIN, "stem: unsupported input channel count %d", a.Cin);
  const size_t smem = ((size_t)a.Cin * SPH * SPP + 9 * a.Cin * 64 + 64) * sizeof(float);
  dim3 grid(ceil_div(a.OW, STW) * ceil_div(a.OH, STH), a.B);
  EGN_DISPATCH_STORAGE(dt, S, {
    EGN_CUDA_CHECK(cudaFuncSetAttribute(stem_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGN_CUDA_CHECK(launch_pdl(stem_kernel<S>, grid, dim3(STEM_THREADS), smem, st, a));
  });
  EGN_LAUNCH_CHECK("stem_kernel");
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// fuse: out = relu(sum_j up(term_j)), 4 channels per thread
// ---------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs p) {
  // 8 channels (16 bytes of fp16) per thread; all terms' loads are issued before the sum so that
  // several independent requests per thread are in flight
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  const int cq = p.Cp / 8;
  const int64_t total = (int64_t)p.B * p.H * p.W * cq;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % cq) * 8;
    int64_t pix = e / cq;
    const int w = (int)(pix % p.W);
    pix /= p.W;
    const int h = (int)(pix % p.H);
    const int b = (int)(pix / p.H);
    float v[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < p.nterms) {
        const int sh = p.shift[j];
        const int Hs = p.H >> sh, Ws = p.W >> sh;
        const int64_t tp = ((int64_t)b * Hs + (h >> sh)) * Ws + (w >> sh);
        S::load4(p.term[j], tp, p.Cp, c, v[j]);
        S::load4(p.term[j], tp, p.Cp, c + 4, v[j] + 4);
      }
    }
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = v[0][q];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      if (j < p.nterms) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += v[j][q];
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = fmaxf(acc[q], 0.f);
    const int64_t op = ((int64_t)b * p.H + h) * p.W + w;
    S::store4(p.out, op, p.Cp, c, acc);
    S::store4(p.out, op, p.Cp, c + 4, acc + 4);
  }
}

int launch_fuse(Dtype dt, const FuseArgs& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.B * a.H * a.W * (a.Cp / 8);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, threads), 148 * 8);
  EGN_DISPATCH_STORAGE(dt, S, EGN_CUDA_CHECK(launch_pdl(fuse_kernel<S>, dim3(blocks), dim3(threads), 0, st, a)));
  EGN_LAUNCH_CHECK("fuse_kernel");
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// head tail: full-map valid conv + bias + sigmoid; one CTA per crop
// ---------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) head_tail_kernel(HeadTailArgs p) {
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  extern __shared__ float xin[];
  // the crop's [kh*kw pixels][Cp] map flattened to L = kh*kw*Cp logical elements
  for (int e = threadIdx.x; e < p.L; e += blockDim.x)
    xin[e] = S::load1(p.in, (int64_t)blockIdx.x * (p.L / p.Cp) + e / p.Cp, p.Cp, e % p.Cp);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < p.Cout; j += blockDim.x / 32) {
    const float* w = p.w + (size_t)j * p.L;
    float s = 0.f;
    for (int e = lane; e < p.L; e += 32) s = fmaf(xin[e], __ldg(w + e), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s += p.bias[j];
      if (p.logits) p.logits[(int64_t)blockIdx.x * p.Cout + j] = s;
      if (p.coords) p.coords[(int64_t)blockIdx.x * p.Cout + j] = 1.0f / (1.0f + expf(-s));
    }
  }
}

int launch_head_tail(Dtype dt, const HeadTailArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.L * sizeof(float);
  EGN_REQUIRE(smem <= 200 * 1024, "head tail: map too large (%d elements)", a.L);
  EGN_DISPATCH_STORAGE(dt, S, {
    EGN_CUDA_CHECK(cudaFuncSetAttribute(head_tail_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGN_CUDA_CHECK(launch_pdl(head_tail_kernel<S>, dim3(a.B), dim3(256), smem, st, a));
  });
  EGN_LAUNCH_CHECK("head_tail_kernel");
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// NHWC (padded) -> fp32 NCHW, for debug taps
// ---------------------------------------------------------------------------
template <typename S>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ in, float* __restrict__ out, int B, int H,
                                    int W, int Cp, int C) {
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(e % W);
    int64_t r = e / W;
    const int h = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    out[e] = S::load1(in, ((int64_t)b * H + h) * W + w, Cp, c);
  }
}

int launch_nhwc_to_nchw(Dtype dt, const void* in, float* out, int B, int H, int W, int Cp, int C,
                        cudaStream_t st) {
  const int64_t total = (int64_t)B * C * H * W;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 16);
  EGN_DISPATCH_STORAGE(dt, S, (nhwc_to_nchw_kernel<S><<<blocks, 256, 0, st>>>(in, out, B, H, W, Cp, C)));
  EGN_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return EGN_OK;
}

}  // namespace egn
