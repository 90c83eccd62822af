// Launch interfaces of the HC kernels (implemented in conv_simt.cu / conv_tc.cu),
// used by the engine in hrnet_engine.cu.  All activations are NHWC with the
// channel count padded to a multiple of 16 (pad lanes hold exact zeros).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace egn {

// Activation storage.  F16X2 ("split"): every logical element is an unevaluated sum hi + lo of two fp16 values
// (hi = rn16(v), lo = rn16(v - hi): ~22 significant bits), stored per pixel as two planes [hi: Cp][lo: Cp] --
// physically an fp16 NHWC tensor with 2 * Cp channels.  The tensor-core kernels then compute
// x * w = x_hi * w_hi + x_lo * w_hi + x_hi * w_lo (three fp16 MMAs into the same fp32 accumulator), which holds the
// reference's fp32 results to ~1e-6 relative per layer.
enum class Dtype : int { F32 = 0, F16 = 1, F16X2 = 2 };

inline size_t dtype_size(Dtype d) { return d == Dtype::F16 ? 2 : 4; }   // bytes per logical element

// One fused conv launch: out = act(conv(in, w) + bias [+ res]).
struct ConvArgs {
  const void* in;        // [B, H, W, Cin_p]
  void* out;             // [B, OH, OW, Cout_p]
  const void* res;       // [B, OH, OW, Cout_p] or null (added before the ReLU)
  const float* bias;     // [Cout_p] fp32, BN shift folded in, zero in pad lanes
  int B, H, W, Cin_p, OH, OW, Cout_p, Cout;
  int ksize, stride, pad, relu;
  int split;             // 1: in / res / out are F16X2 split tensors (tcgen05 path: error-compensated MMAs)
  // head1 extras
  float* heatmap;        // fp32 NCHW [B, Cout, OH, OW] copy of the un-rounded result, or null
  const float* xs;       // [OW] / [OH] coordinate-map values written to channels
  const float* ys;       //   Cout and Cout+1 when coord_maps != 0 (hrnet.py:461-467,606)
  int coord_maps;
};

// CUDA-core implicit GEMM (exact fp32 accumulation, fp32 weights [tap][Cin_p][Cout_p]).
int launch_conv_simt(Dtype dt, const ConvArgs& a, const float* w_packed, cudaStream_t st);
// 64-channel column groups per CTA (1 or 2) the FFMA kernels use for this output width
int simt_groups_for(int Cout_p);

// Stem conv1: fp32 NCHW network input -> NHWC, 3x3 stride 2 pad 1, Cin in {3,5}, Cout 64.
struct StemArgs {
  const float* x;        // [B, Cin, H, W]
  void* out;             // [B, H/2, W/2, 64]
  const float* w;        // [9][Cin][64] folded fp32
  const float* bias;     // [64]
  int B, Cin, H, W, OH, OW;
};
int launch_stem(Dtype dt, const StemArgs& a, cudaStream_t st);

// Cross-resolution fuse: out = relu(sum_j term_j), term_j sampled with nearest
// up-sampling by 2^shift_j (hrnet.py:241,291-298).  Terms are summed in order.
struct FuseArgs {
  void* out;             // [B, H, W, Cp]
  const void* term[4];
  int shift[4];
  int nterms;
  int B, H, W, Cp;
};
int launch_fuse(Dtype dt, const FuseArgs& a, cudaStream_t st);

// head2 tail: valid kh x kw conv over the whole [kh, kw] map + bias + sigmoid (hrnet.py:457-458).
struct HeadTailArgs {
  const void* in;        // [B, kh, kw, Cp]
  const float* w;        // [Cout][kh*kw*Cp] fp32 (zero in pad lanes)
  const float* bias;     // [Cout]
  float* coords;         // [B, Cout] sigmoid output, or null
  float* logits;         // [B, Cout] or null
  int B, L, Cout, Cp;    // L = kh*kw*Cp logical elements per crop
};
int launch_head_tail(Dtype dt, const HeadTailArgs& a, cudaStream_t st);

// NHWC (padded) activation -> fp32 NCHW [B, C, H, W] (debug taps).
int launch_nhwc_to_nchw(Dtype dt, const void* in, float* out, int B, int H, int W, int Cp, int C,
                        cudaStream_t st);

// ---- tcgen05 path (conv_tc.cu) ---------------------------------------------
struct TcConvPlan;  // opaque: packed fp16 weights + tensor-map recipe for one conv
bool tc_conv_supported(const ConvArgs& shape_only);
// Build the per-layer plan from folded fp32 weights laid out [tap][Cin_p][Cout_p].
int tc_conv_plan_create(const ConvArgs& shape_only, const float* w_folded_host, TcConvPlan** out);
void tc_conv_plan_destroy(TcConvPlan* p);
int launch_conv_tc(TcConvPlan* plan, const ConvArgs& a, cudaStream_t st);
size_t tc_conv_plan_weight_bytes(const TcConvPlan* p);

}  // namespace egn
