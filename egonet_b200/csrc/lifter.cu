// L: the 2D->3D lifter (residual MLP) with its (de)normalisation fused in.
//
// upstream: EgoNet.lift_2d_to_3d egonet.py:469-486; FCModel.forward
// FCmodel.py:92-105; ResidualBlock.forward FCmodel.py:33-43; normalize_1d /
// unnormalize_1d operations.py:21-52.  Eval mode: Dropout is the identity and
// every BatchNorm1d is folded into the preceding Linear at finalize().
//
// All instances of a batch go through one chain of 2*blocks+2 fp32 GEMM launches
// (upstream calls L once per image with 1-15 rows).  fp32 FFMA accumulation keeps
// the result inside the 1e-4 parity bound; the first launch normalises the fp64
// screen key-points on load and the last one de-normalises to fp64 on store.
// The weights (17.5 MB fp32) are the only real traffic: each launch streams its
// [K,J] matrix once, shared across all rows of the batch through L2.
#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace egn {

constexpr int LBM = 32, LBN = 64, LBK = 32, LTHREADS = 128;
constexpr int LAP = LBM + 4;      // A tile pitch in floats: 16-byte aligned rows, stores spread over 8 banks

struct LinearArgs {
  const float* a_f32;      // [n, K] fp32 input, or null when a_f64 is used
  const double* a_f64;     // [n, K] fp64 input normalised on load: (x - mean_in) / std_in
  const double* mean_in;
  const double* std_in;
  const float* wt;         // [K, J] folded weights (transposed torch layout)
  const float* bias;       // [J] folded bias
  const float* skip;       // [n, J] residual added AFTER the activation, or null
  float* out_f32;          // [n, J] or null
  double* out_f64;         // [n, J] de-normalised: y * std_out + mean_out, or null
  const double* mean_out;
  const double* std_out;
  int n, K, J, relu;
};

// Register-tiled fp32 GEMM: 32 x 64 output tile per CTA, 4 x 4 outputs per thread (two 16-byte shared-memory reads
// per 16 FFMAs; the first version read three words per 8 FFMAs), K slabs of 32 streamed through a 3-stage cp.async
// ring: with one CTA of four warps per SM the chain is bound by the latency of its own loads -- 368 us per 256
// instances with synchronous slabs, 270 us with a one-slab register prefetch (and the same at 64 instances).  Every
// output still accumulates its products in ascending k with fmaf: bit-identical to the simple kernel it replaces.
constexpr int LSTAGES = 3;

__device__ __forceinline__ void cp_async_4(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int n = valid ? 4 : 0;                 // src-size 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(LTHREADS) lifter_linear_kernel(LinearArgs p) {
  __shared__ __align__(16) float As[LSTAGES][LBK][LAP];
  __shared__ __align__(16) float Bs[LSTAGES][LBK][LBN];
  const int t = threadIdx.x;
  const int row0 = blockIdx.y * LBM, col0 = blockIdx.x * LBN;
  const int tx = t % 16, ty = t / 16;  // 4 columns x 4 rows per thread
  const bool vec_b = (p.J % 4 == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int n_slabs = (p.K + LBK - 1) / LBK;
  // slab s -> ring stage s % LSTAGES (asynchronous copies; out-of-range elements are zero-filled)
  auto issue = [&](int s) {
    const int st = s % LSTAGES, k0 = s * LBK;
    // A slab: 32 rows x 32 k (coalesced along k), stored [k][row]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = t + i * LTHREADS;
      const int r = e / LBK, k = e % LBK;
      const int gr = row0 + r, gk = k0 + k;
      const bool ok = gr < p.n && gk < p.K;
      if (p.a_f64) {
        // first layer: fp64 screen key-points normalised on load (three slabs: synchronous)
        As[st][k][r] = ok ? (float)((p.a_f64[(size_t)gr * p.K + gk] - p.mean_in[gk]) / p.std_in[gk]) : 0.f;
      } else {
        cp_async_4(&As[st][k][r], ok ? p.a_f32 + (size_t)gr * p.K + gk : p.a_f32, ok);
      }
    }
    // B slab: 32 k x 64 j (coalesced along j, 16-byte copies when the row pitch allows)
    if (vec_b) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = t + i * LTHREADS;
        const int k = e / 16, j = (e % 16) * 4;
        const int gk = k0 + k, gj = col0 + j;
        const bool ok = gk < p.K && gj + 3 < p.J;
        cp_async_16(&Bs[st][k][j], ok ? p.wt + (size_t)gk * p.J + gj : p.wt, ok);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int e = t + i * LTHREADS;
        const int k = e / LBN, j = e % LBN;
        const int gk = k0 + k, gj = col0 + j;
        const bool ok = gk < p.K && gj < p.J;
        cp_async_4(&Bs[st][k][j], ok ? p.wt + (size_t)gk * p.J + gj : p.wt, ok);
      }
    }
  };
  for (int s = 0; s < LSTAGES - 1; ++s) {
    if (s < n_slabs) issue(s);
    cp_async_commit();
  }
  for (int s = 0; s < n_slabs; ++s) {
    cp_async_wait<LSTAGES - 2>();            // slab s has landed (this thread's copies)
    __syncthreads();                         // ... everybody's; and stage (s - 1) % LSTAGES is no longer being read
    if (s + LSTAGES - 1 < n_slabs) issue(s + LSTAGES - 1);
    cp_async_commit();
    const int st = s % LSTAGES;
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = row0 + ty * 4 + i;
    if (gr >= p.n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = col0 + tx * 4 + j;
      if (gj >= p.J) continue;
      float v = acc[i][j] + p.bias[gj];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.skip) v += p.skip[(size_t)gr * p.J + gj];
      if (p.out_f32) p.out_f32[(size_t)gr * p.J + gj] = v;
      if (p.out_f64) p.out_f64[(size_t)gr * p.J + gj] = (double)v * p.std_out[gj] + p.mean_out[gj];
    }
  }
}

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

}  // namespace egn

struct egn_lifter {
  int nin, nout, neurons, blocks;
  std::map<std::string, egn::HostTensor> raw;
  std::vector<double> stats[4];  // mean_in, std_in, mean_out, std_out
  bool finalized = false;
  // device
  std::vector<float*> d_wt, d_bias;  // 2*blocks+2 layers
  double* d_stats[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace egn {

static void lifter_free_device(egn_lifter* l) {
  for (float* p : l->d_wt) cudaFree(p);
  for (float* p : l->d_bias) cudaFree(p);
  l->d_wt.clear();
  l->d_bias.clear();
  for (auto& p : l->d_stats) {
    cudaFree(p);
    p = nullptr;
  }
  l->finalized = false;
}

static const HostTensor* find(const egn_lifter* l, const std::string& key) {
  auto it = l->raw.find(key);
  return it == l->raw.end() ? nullptr : &it->second;
}

// Fold Linear(+BatchNorm1d) into (Wt [K,J], bias [J]); math in double, stored fp32.
static int fold_linear(const egn_lifter* l, const std::string& lin, const std::string& bn, int K,
                       int J, std::vector<float>* wt, std::vector<float>* bias) {
  const HostTensor* w = find(l, lin + ".weight");
  const HostTensor* b = find(l, lin + ".bias");
  if (!w || !b) {
    set_error("lifter weight '%s.weight/bias' was never set", lin.c_str());
    return EGN_ERR_MISSING;
  }
  if ((int64_t)w->data.size() != (int64_t)K * J || (int64_t)b->data.size() != J) {
    set_error("lifter weight '%s' has the wrong size", lin.c_str());
    return EGN_ERR_INVALID;
  }
  std::vector<double> scale(J, 1.0), shift(J, 0.0);
  if (!bn.empty()) {
    const HostTensor *g = find(l, bn + ".weight"), *be = find(l, bn + ".bias"),
                     *mu = find(l, bn + ".running_mean"), *var = find(l, bn + ".running_var");
    if (!g || !be || !mu || !var) {
      set_error("lifter BatchNorm '%s' is incomplete", bn.c_str());
      return EGN_ERR_MISSING;
    }
    for (int j = 0; j < J; ++j) {
      scale[j] = (double)g->data[j] / std::sqrt((double)var->data[j] + 1e-5);
      shift[j] = (double)be->data[j] - (double)mu->data[j] * scale[j];
    }
  }
  wt->assign((size_t)K * J, 0.f);
  bias->assign(J, 0.f);
  for (int j = 0; j < J; ++j) {
    for (int k = 0; k < K; ++k) (*wt)[(size_t)k * J + j] = (float)((double)w->data[(size_t)j * K + k] * scale[j]);
    (*bias)[j] = (float)((double)b->data[j] * scale[j] + shift[j]);
  }
  return EGN_OK;
}

}  // namespace egn

extern "C" {

int egn_lifter_create(int input_size, int output_size, int num_neurons, int num_blocks,
                      egn_lifter** out) {
  using namespace egn;
  EGN_REQUIRE(out, "egn_lifter_create: null out");
  EGN_REQUIRE(input_size > 0 && output_size > 0 && num_neurons > 0 && num_blocks >= 0,
              "egn_lifter_create: bad sizes");
  egn_lifter* l = new egn_lifter();
  l->nin = input_size;
  l->nout = output_size;
  l->neurons = num_neurons;
  l->blocks = num_blocks;
  *out = l;
  return EGN_OK;
}

void egn_lifter_destroy(egn_lifter* l) {
  if (!l) return;
  egn::lifter_free_device(l);
  delete l;
}

int egn_lifter_set_weight(egn_lifter* l, const char* key, const float* host_data,
                          const int64_t* shape, int ndim) {
  using namespace egn;
  EGN_REQUIRE(l && key, "egn_lifter_set_weight: null argument");
  std::string k(key);
  if (k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return EGN_OK;
  EGN_REQUIRE(host_data && ndim >= 1 && ndim <= 2, "egn_lifter_set_weight: bad tensor for '%s'", key);
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  HostTensor t;
  t.data.assign(host_data, host_data + n);
  t.shape.assign(shape, shape + ndim);
  l->raw[k] = std::move(t);
  l->finalized = false;
  return EGN_OK;
}

int egn_lifter_set_stats(egn_lifter* l, const double* mean_in, const double* std_in,
                         const double* mean_out, const double* std_out) {
  using namespace egn;
  EGN_REQUIRE(l && mean_in && std_in && mean_out && std_out, "egn_lifter_set_stats: null argument");
  l->stats[0].assign(mean_in, mean_in + l->nin);
  l->stats[1].assign(std_in, std_in + l->nin);
  l->stats[2].assign(mean_out, mean_out + l->nout);
  l->stats[3].assign(std_out, std_out + l->nout);
  l->finalized = false;
  return EGN_OK;
}

int egn_lifter_finalize(egn_lifter* l) {
  using namespace egn;
  EGN_REQUIRE(l, "egn_lifter_finalize: null handle");
  if (int rc = require_device()) return rc;
  lifter_free_device(l);
  if (l->stats[0].empty()) {
    set_error("lifter statistics (LS) were never set");
    return EGN_ERR_MISSING;
  }
  struct Layer {
    std::string lin, bn;
    int K, J;
  };
  std::vector<Layer> layers;
  layers.push_back({"w1", "batch_norm1", l->nin, l->neurons});
  for (int i = 0; i < l->blocks; ++i) {
    const std::string p = "res_blocks." + std::to_string(i);
    layers.push_back({p + ".w1", p + ".batch_norm1", l->neurons, l->neurons});
    layers.push_back({p + ".w2", p + ".batch_norm2", l->neurons, l->neurons});
  }
  layers.push_back({"w2", "", l->neurons, l->nout});
  for (const Layer& ly : layers) {
    std::vector<float> wt, bias;
    if (int rc = fold_linear(l, ly.lin, ly.bn, ly.K, ly.J, &wt, &bias)) return rc;
    float *dw = nullptr, *db = nullptr;
    EGN_CUDA_CHECK(cudaMalloc(&dw, wt.size() * sizeof(float)));
    l->d_wt.push_back(dw);
    EGN_CUDA_CHECK(cudaMalloc(&db, bias.size() * sizeof(float)));
    l->d_bias.push_back(db);
    EGN_CUDA_CHECK(cudaMemcpy(dw, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    EGN_CUDA_CHECK(cudaMemcpy(db, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < 4; ++i) {
    EGN_CUDA_CHECK(cudaMalloc(&l->d_stats[i], l->stats[i].size() * sizeof(double)));
    EGN_CUDA_CHECK(cudaMemcpy(l->d_stats[i], l->stats[i].data(), l->stats[i].size() * sizeof(double),
                              cudaMemcpyHostToDevice));
  }
  l->finalized = true;
  return EGN_OK;
}

size_t egn_lifter_workspace_bytes(const egn_lifter* l, int n) {
  if (!l || n <= 0) return 0;
  const size_t per = ((size_t)n * l->neurons * sizeof(float) + 255) / 256 * 256;
  return 3 * per;
}

int egn_lifter_forward(egn_lifter* l, const double* kpts_2d, int n, double* kpts_3d, float* raw_out,
                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace egn;
  EGN_REQUIRE(l, "egn_lifter_forward: null handle");
  EGN_REQUIRE(n >= 0, "egn_lifter_forward: negative n");
  EGN_REQUIRE(n == 0 || (kpts_2d && kpts_3d), "egn_lifter_forward: null argument");
  if (!l->finalized) {
    set_error("egn_lifter_forward called before egn_lifter_finalize");
    return EGN_ERR_STATE;
  }
  if (int rc = require_device()) return rc;
  if (n == 0) return EGN_OK;
  if (!workspace || workspace_bytes < egn_lifter_workspace_bytes(l, n)) {
    set_error("lifter workspace too small: need %zu bytes", egn_lifter_workspace_bytes(l, n));
    return EGN_ERR_WORKSPACE;
  }
  const size_t per = egn_lifter_workspace_bytes(l, n) / 3;
  float* buf[3];
  for (int i = 0; i < 3; ++i) buf[i] = reinterpret_cast<float*>(static_cast<char*>(workspace) + i * per);
  cudaStream_t st = as_stream(stream);
  auto launch = [&](LinearArgs a) -> int {
    dim3 grid(ceil_div(a.J, LBN), ceil_div(a.n, LBM));
    lifter_linear_kernel<<<grid, LTHREADS, 0, st>>>(a);
    EGN_LAUNCH_CHECK("lifter_linear_kernel");
    return EGN_OK;
  };
  int li = 0;
  LinearArgs a{};
  a.n = n;
  // w1 + batch_norm1 + relu, normalising the fp64 input on load
  a.a_f64 = kpts_2d; a.mean_in = l->d_stats[0]; a.std_in = l->d_stats[1];
  a.wt = l->d_wt[li]; a.bias = l->d_bias[li]; a.K = l->nin; a.J = l->neurons; a.relu = 1;
  a.out_f32 = buf[0];
  if (int rc = launch(a)) return rc;
  ++li;
  int cur = 0;
  for (int b = 0; b < l->blocks; ++b) {
    const int t1 = (cur + 1) % 3, t2 = (cur + 2) % 3;
    LinearArgs r{};
    r.n = n; r.K = l->neurons; r.J = l->neurons; r.relu = 1;
    r.a_f32 = buf[cur]; r.wt = l->d_wt[li]; r.bias = l->d_bias[li]; r.out_f32 = buf[t1];
    if (int rc = launch(r)) return rc;
    ++li;
    r.a_f32 = buf[t1]; r.wt = l->d_wt[li]; r.bias = l->d_bias[li]; r.out_f32 = buf[t2];
    r.skip = buf[cur];  // out = x + relu(bn(w2(.)))  (FCmodel.py:42)
    if (int rc = launch(r)) return rc;
    ++li;
    cur = t2;
  }
  LinearArgs f{};
  f.n = n; f.K = l->neurons; f.J = l->nout; f.relu = 0;
  f.a_f32 = buf[cur]; f.wt = l->d_wt[li]; f.bias = l->d_bias[li];
  f.out_f32 = raw_out; f.out_f64 = kpts_3d; f.mean_out = l->d_stats[2]; f.std_out = l->d_stats[3];
  return launch(f);
}

}  // extern "C"
