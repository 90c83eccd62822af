// Register-tiled fp32 FFMA GEMM core shared by the CUDA-core convolution kernels (conv_simt.cu: forward;
// hrnet_train.cu: data gradient, weight gradient).  A CTA of 256 threads (16 x 16) owns a 128 x BN output tile,
// BN = 16 * GROUPS * GW in {64, 96, 128}; a thread owns 8 rows (ty*4 + i and 64 + ty*4 + i) x TN = GROUPS * GW
// columns (column group g: g*16*GW + tx*GW + j), so each K step issues 8 * TN FFMA for 2 LDS.128 + GROUPS loads.
// Operands are staged K-major in shared memory, As[k][m] / Bs[k][n], one 16-deep K slab at a time; the caller
// prefetches the next slab's global loads into registers while this one is multiplied (two-stage pipeline).
#pragma once

namespace egn {

constexpr int SG_BM = 128, SG_BK = 16, SG_THREADS = 256;
constexpr int SG_APITCH = SG_BM + 4;          // As row pitch (floats): keeps float4 alignment, staggers banks

template <int GROUPS, int GW>
struct SgTile {
  static constexpr int TN = GROUPS * GW;
  static constexpr int BN = 16 * TN;
};

// acc[i][g*GW + j] += As[k][row(i)] * Bs[k][col(g, j)] over the slab
template <int GROUPS, int GW>
__device__ __forceinline__ void sg_slab_fma(const float (*As)[SG_APITCH], const float (*Bs)[16 * GROUPS * GW], int tx, int ty,
                                            float (&acc)[8][GROUPS * GW]) {
#pragma unroll
  for (int k = 0; k < SG_BK; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[GROUPS * GW];
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const float* bp = &Bs[k][g * 16 * GW + tx * GW];
      if (GW == 4) {
        const float4 v = *reinterpret_cast<const float4*>(bp);
        b[g * GW + 0] = v.x; b[g * GW + 1] = v.y; b[g * GW + 2] = v.z; b[g * GW + 3] = v.w;
      } else if (GW == 2) {
        const float2 v = *reinterpret_cast<const float2*>(bp);
        b[g * GW + 0] = v.x; b[g * GW + 1] = v.y;
      } else {
#pragma unroll
        for (int j = 0; j < GW; ++j) b[g * GW + j] = bp[j];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < GROUPS * GW; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// row / column owned by accumulator (i, g, j) of thread (tx, ty)
__device__ __forceinline__ int sg_row(int ty, int i) { return (i < 4 ? 0 : 64) + ty * 4 + (i & 3); }
template <int GW>
__device__ __forceinline__ int sg_col(int tx, int g, int j) { return g * 16 * GW + tx * GW + j; }

}  // namespace egn
