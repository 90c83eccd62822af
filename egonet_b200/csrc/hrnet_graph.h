// Op graph of the HC network shared by the inference engine (hrnet_engine.cu) and the training engine
// (hrnet_train.cu): tensors, fused-op list and the state_dict inventory, built from an egn_hrnet_cfg.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace egn {

constexpr int kChanAlign = 16;
constexpr int64_t kSizeAlign = 512;  // per-crop buffer sizes are multiples of this many elements

struct TensorInfo {
  int C = 0, Cp = 0, H = 0, W = 0;
  int64_t per_crop = 0;   // elements per crop, rounded up to kSizeAlign
  int64_t offset = -1;    // per-crop element offset inside the workspace
  int def_op = -1, last_use = -1;
  std::string tap;
};

struct ConvWeights {
  std::string conv_key, bn_key;
  bool has_bias = false;
  int Cin = 0, Cout = 0, k = 1;
  int Cin_p = 0, Cout_p = 0;
  bool force_fp32 = false;       // stem / tail keep fp32 weights in every precision mode
  float* d_simt = nullptr;       // [tap][Cin_p][Cout_p] fp32 (SIMT kernels)
  float* d_bias = nullptr;       // [Cout_p]
  TcConvPlan* tc = nullptr;      // tcgen05 plan (fp16 packed weights)
};

struct Op {
  enum Kind { STEM, CONV, FUSE, TAIL } kind = CONV;
  int in = -1, out = -1, res = -1;
  int wi = -1;
  int stride = 1, pad = 0, relu = 0;
  bool write_heatmap = false, coord_maps = false;
  bool use_tc = false;
  int nterms = 0, term[4] = {-1, -1, -1, -1}, shift[4] = {0, 0, 0, 0};
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

}  // namespace egn

struct egn_hrnet {
  egn_hrnet_cfg cfg;
  egn::Dtype dt;
  std::vector<std::string> keys;                 // state_dict order
  std::vector<std::vector<int64_t>> key_shapes;
  std::map<std::string, egn::HostTensor> raw;
  std::vector<egn::TensorInfo> tensors;
  std::vector<egn::ConvWeights> weights;
  std::map<std::string, int> weight_index;       // conv key -> weights[]
  std::vector<egn::Op> ops;
  std::map<std::string, int> taps;               // tap name -> tensor id
  int64_t ws_per_crop = 0;                       // elements
  int64_t macs = 0, act_bytes = 0, weight_bytes = 0;
  int n_launches = 0, n_tc = 0;
  float *d_xs = nullptr, *d_ys = nullptr;
  bool finalized = false;
  std::string build_error;
};

namespace egn {
// state_dict inventory (same traversal as PoseHighResolutionNet.__init__) and op graph; build_graph returns
// EGN_OK or EGN_ERR_INVALID with h->build_error set.
void build_keys(egn_hrnet* h);
int build_graph(egn_hrnet* h);
}  // namespace egn
