// HC engine: builds the HRNet op graph from a config, folds/repacks the
// reference's state_dict, plans the activation workspace and replays the graph
// as a sequence of fused kernels on a caller-provided stream.
//
// upstream structure followed (libs/model/heatmapModel/hrnet.py):
//   module/param inventory   :311-469, 471-561, 174-277
//   forward                  :563-614 (BasicBlock :76-92, Bottleneck :113-133,
//                            HighResolutionModule :282-300)
// Everything the reference runs as conv -> BN -> ReLU (-> add -> ReLU) is ONE
// launch here: BN is folded into the weights at finalize(), bias / residual /
// ReLU live in the conv epilogue; nearest up-sampling + the multi-branch sum +
// ReLU of a fuse row are one launch.
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "hrnet_graph.h"
#include "kernels.h"

namespace egn {

// ---------------------------------------------------------------------------
// state_dict inventory (same traversal as PoseHighResolutionNet.__init__)
// ---------------------------------------------------------------------------
struct KeyBuilder {
  egn_hrnet* h;
  void bn(const std::string& p, int c) {
    for (const char* leaf : {"weight", "bias", "running_mean", "running_var"}) {
      h->keys.push_back(p + "." + leaf);
      h->key_shapes.push_back({c});
    }
    h->keys.push_back(p + ".num_batches_tracked");
    h->key_shapes.push_back({});
  }
  void conv(const std::string& p, int cout, int cin, int kh, int kw, bool bias = false) {
    h->keys.push_back(p + ".weight");
    h->key_shapes.push_back({cout, cin, kh, kw});
    if (bias) {
      h->keys.push_back(p + ".bias");
      h->key_shapes.push_back({cout});
    }
  }
  void basic(const std::string& p, int cin, int cout, bool down) {
    conv(p + ".conv1", cout, cin, 3, 3);
    bn(p + ".bn1", cout);
    conv(p + ".conv2", cout, cout, 3, 3);
    bn(p + ".bn2", cout);
    if (down) {
      conv(p + ".downsample.0", cout, cin, 1, 1);
      bn(p + ".downsample.1", cout);
    }
  }
};

static std::string S(int v) { return std::to_string(v); }

void build_keys(egn_hrnet* h) {
  const egn_hrnet_cfg& c = h->cfg;
  KeyBuilder kb{h};
  kb.conv("conv1", 64, c.in_channels, 3, 3);
  kb.bn("bn1", 64);
  kb.conv("conv2", 64, 64, 3, 3);
  kb.bn("bn2", 64);
  int inpl = 64;
  for (int k = 0; k < 4; ++k) {
    const std::string p = "layer1." + S(k);
    kb.conv(p + ".conv1", 64, inpl, 1, 1);
    kb.bn(p + ".bn1", 64);
    kb.conv(p + ".conv2", 64, 64, 3, 3);
    kb.bn(p + ".bn2", 64);
    kb.conv(p + ".conv3", 256, 64, 1, 1);
    kb.bn(p + ".bn3", 256);
    if (k == 0) {
      kb.conv(p + ".downsample.0", 256, inpl, 1, 1);
      kb.bn(p + ".downsample.1", 256);
    }
    inpl = 256;
  }
  std::vector<int> pre = {256};
  for (int si = 0; si < c.num_stages; ++si) {
    const int nb = c.stage_branches[si];
    std::vector<int> cur(c.stage_channels[si], c.stage_channels[si] + nb);
    const std::string tp = "transition" + S(si + 1);
    for (int i = 0; i < nb; ++i) {
      if (i < (int)pre.size()) {
        if (cur[i] != pre[i]) {
          kb.conv(tp + "." + S(i) + ".0", cur[i], pre[i], 3, 3);
          kb.bn(tp + "." + S(i) + ".1", cur[i]);
        }
      } else {
        for (int j = 0; j < i + 1 - (int)pre.size(); ++j) {
          const int cin = pre.back();
          const int cout = (j == i - (int)pre.size()) ? cur[i] : cin;
          kb.conv(tp + "." + S(i) + "." + S(j) + ".0", cout, cin, 3, 3);
          kb.bn(tp + "." + S(i) + "." + S(j) + ".1", cout);
        }
      }
    }
    const bool last_stage = si == c.num_stages - 1;
    for (int m = 0; m < c.stage_modules[si]; ++m) {
      const std::string mp = "stage" + S(si + 2) + "." + S(m);
      const bool multi = !(last_stage && m == c.stage_modules[si] - 1);
      for (int b = 0; b < nb; ++b)
        for (int k = 0; k < c.stage_blocks[si][b]; ++k)
          kb.basic(mp + ".branches." + S(b) + "." + S(k), cur[b], cur[b], false);
      for (int i = 0; i < (multi ? nb : 1); ++i)
        for (int j = 0; j < nb; ++j) {
          const std::string fp = mp + ".fuse_layers." + S(i) + "." + S(j);
          if (j > i) {
            kb.conv(fp + ".0", cur[i], cur[j], 1, 1);
            kb.bn(fp + ".1", cur[i]);
          } else if (j < i) {
            for (int k = 0; k < i - j; ++k) {
              const int cout = (k == i - j - 1) ? cur[i] : cur[j];
              kb.conv(fp + "." + S(k) + ".0", cout, cur[j], 3, 3);
              kb.bn(fp + "." + S(k) + ".1", cout);
            }
          }
        }
    }
    pre = cur;
  }
  const int nj = c.num_joints;
  if (c.head_type == EGN_HEAD_HEATMAP) {
    kb.conv("final_layer", nj, pre[0], c.final_conv_kernel, c.final_conv_kernel, true);
  } else {
    kb.conv("head1.0", nj, pre[0], 1, 1, true);
    int cin = nj + 2;
    for (int k = 0; k < 4; ++k) {
      kb.basic("head2." + S(k), cin, 2 * nj, true);
      cin = 2 * nj;
    }
    kb.conv("head2.4", 2 * nj, 2 * nj, c.heatmap_h / 16, c.heatmap_w / 16, true);
  }
}

// ---------------------------------------------------------------------------
// op graph
// ---------------------------------------------------------------------------
struct GraphBuilder {
  egn_hrnet* h;

  int tensor(int C, int H, int W, const std::string& tap = "") {
    TensorInfo t;
    t.C = C;
    t.Cp = round_up(C, kChanAlign);
    t.H = H;
    t.W = W;
    t.per_crop = ceil_div64((int64_t)H * W * t.Cp, kSizeAlign) * kSizeAlign;
    t.tap = tap;
    h->tensors.push_back(t);
    const int id = (int)h->tensors.size() - 1;
    if (!tap.empty()) h->taps[tap] = id;
    return id;
  }
  void name_tap(int id, const std::string& tap) {
    h->tensors[id].tap = tap;
    h->taps[tap] = id;
  }
  int weights(const std::string& conv_key, const std::string& bn_key, bool bias, int cin, int cout,
              int k, int cin_p, int cout_p, bool force_fp32 = false) {
    ConvWeights w;
    w.conv_key = conv_key;
    w.bn_key = bn_key;
    w.has_bias = bias;
    w.Cin = cin;
    w.Cout = cout;
    w.k = k;
    w.Cin_p = cin_p;
    w.Cout_p = cout_p;
    w.force_fp32 = force_fp32;
    h->weights.push_back(w);
    h->weight_index[conv_key] = (int)h->weights.size() - 1;
    return (int)h->weights.size() - 1;
  }
  // conv + folded BN (+ residual) (+ ReLU); returns the output tensor id
  int conv(int in, const std::string& ck, const std::string& bk, int cout, int k, int stride, int relu,
           int res = -1, bool bias = false, int out_c = -1) {
    const TensorInfo ti = h->tensors[in];
    const int pad = k == 3 ? 1 : 0;
    const int OH = (ti.H + 2 * pad - k) / stride + 1, OW = (ti.W + 2 * pad - k) / stride + 1;
    const int out = tensor(out_c < 0 ? cout : out_c, OH, OW);
    Op op;
    op.kind = Op::CONV;
    op.in = in;
    op.out = out;
    op.res = res;
    op.stride = stride;
    op.pad = pad;
    op.relu = relu;
    op.wi = weights(ck, bk, bias, ti.C, cout, k, ti.Cp, h->tensors[out].Cp);
    h->ops.push_back(op);
    h->macs += (int64_t)cout * ti.C * k * k * OH * OW;
    return out;
  }
};

int build_graph(egn_hrnet* h) {
  const egn_hrnet_cfg& c = h->cfg;
  GraphBuilder g{h};
  // stem conv1 (reads the fp32 NCHW network input directly)
  const int H1 = (c.input_h + 2 - 3) / 2 + 1, W1 = (c.input_w + 2 - 3) / 2 + 1;
  int t = g.tensor(64, H1, W1, "stem1");
  {
    Op op;
    op.kind = Op::STEM;
    op.out = t;
    op.wi = g.weights("conv1", "bn1", false, c.in_channels, 64, 3, c.in_channels, 64, true);
    h->ops.push_back(op);
    h->macs += (int64_t)64 * c.in_channels * 9 * H1 * W1;
  }
  t = g.conv(t, "conv2", "bn2", 64, 3, 2, 1);
  g.name_tap(t, "stem2");
  for (int k = 0; k < 4; ++k) {
    const std::string p = "layer1." + S(k);
    const int a = g.conv(t, p + ".conv1", p + ".bn1", 64, 1, 1, 1);
    const int b = g.conv(a, p + ".conv2", p + ".bn2", 64, 3, 1, 1);
    const int r = k == 0 ? g.conv(t, p + ".downsample.0", p + ".downsample.1", 256, 1, 1, 0) : t;
    t = g.conv(b, p + ".conv3", p + ".bn3", 256, 1, 1, 1, r);
  }
  g.name_tap(t, "layer1");
  std::vector<int> pre = {256};
  std::vector<int> ys = {t};
  for (int si = 0; si < c.num_stages; ++si) {
    const int nb = c.stage_branches[si];
    std::vector<int> cur(c.stage_channels[si], c.stage_channels[si] + nb);
    const std::string tp = "transition" + S(si + 1);
    std::vector<int> xs;
    for (int i = 0; i < nb; ++i) {
      if (i < (int)pre.size()) {
        if (cur[i] != pre[i]) {
          if (si != 0) {
            h->build_error = "channel-changing transition on an existing branch after stage 2 is not "
                             "supported (upstream hrnet.py:583 feeds it the wrong tensor)";
            return EGN_ERR_INVALID;
          }
          xs.push_back(g.conv(ys[i], tp + "." + S(i) + ".0", tp + "." + S(i) + ".1", cur[i], 3, 1, 1));
        } else {
          xs.push_back(ys[i]);
        }
      } else {
        int u = ys.back();
        for (int j = 0; j < i + 1 - (int)pre.size(); ++j) {
          const int cout = (j == i - (int)pre.size()) ? cur[i] : pre.back();
          u = g.conv(u, tp + "." + S(i) + "." + S(j) + ".0", tp + "." + S(i) + "." + S(j) + ".1", cout, 3, 2, 1);
        }
        xs.push_back(u);
      }
    }
    const bool last_stage = si == c.num_stages - 1;
    for (int m = 0; m < c.stage_modules[si]; ++m) {
      const std::string mp = "stage" + S(si + 2) + "." + S(m);
      const bool multi = !(last_stage && m == c.stage_modules[si] - 1);
      for (int b = 0; b < nb; ++b)
        for (int k = 0; k < c.stage_blocks[si][b]; ++k) {
          const std::string p = mp + ".branches." + S(b) + "." + S(k);
          const int a = g.conv(xs[b], p + ".conv1", p + ".bn1", cur[b], 3, 1, 1);
          xs[b] = g.conv(a, p + ".conv2", p + ".bn2", cur[b], 3, 1, 1, xs[b]);
        }
      std::vector<int> outs;
      for (int i = 0; i < (multi ? nb : 1); ++i) {
        Op f;
        f.kind = Op::FUSE;
        f.nterms = nb;
        for (int j = 0; j < nb; ++j) {
          const std::string fp = mp + ".fuse_layers." + S(i) + "." + S(j);
          if (j == i) {
            f.term[j] = xs[j];
          } else if (j > i) {
            f.term[j] = g.conv(xs[j], fp + ".0", fp + ".1", cur[i], 1, 1, 0);
            f.shift[j] = j - i;
          } else {
            int u = xs[j];
            for (int k = 0; k < i - j; ++k) {
              const bool lastk = k == i - j - 1;
              u = g.conv(u, fp + "." + S(k) + ".0", fp + "." + S(k) + ".1", lastk ? cur[i] : cur[j], 3, 2,
                         lastk ? 0 : 1);
            }
            f.term[j] = u;
          }
        }
        const TensorInfo xi = h->tensors[xs[i]];
        f.out = g.tensor(cur[i], xi.H, xi.W, mp + ".out" + S(i));
        h->ops.push_back(f);
        outs.push_back(f.out);
      }
      xs = outs;
    }
    ys = xs;
    pre = cur;
  }
  const int feat = ys[0];
  const int nj = c.num_joints;
  if (c.head_type == EGN_HEAD_HEATMAP) {
    const int o = g.conv(feat, "final_layer", "", nj, c.final_conv_kernel, 1, 0, -1, true);
    h->ops.back().write_heatmap = true;
    g.name_tap(o, "heatmap");
  } else {
    // head1 output tensor carries the 33 maps plus the two coordinate maps (hrnet.py:602-606)
    int u = g.conv(feat, "head1.0", "", nj, 1, 1, 0, -1, true, nj + 2);
    h->ops.back().write_heatmap = true;
    h->ops.back().coord_maps = true;
    g.name_tap(u, "head1");
    for (int k = 0; k < 4; ++k) {
      const std::string p = "head2." + S(k);
      const int a = g.conv(u, p + ".conv1", p + ".bn1", 2 * nj, 3, 2, 1);
      const int r = g.conv(u, p + ".downsample.0", p + ".downsample.1", 2 * nj, 1, 2, 0);
      u = g.conv(a, p + ".conv2", p + ".bn2", 2 * nj, 3, 1, 1, r);
      g.name_tap(u, p);
    }
    const TensorInfo tu = h->tensors[u];
    const int kh = c.heatmap_h / 16, kw = c.heatmap_w / 16;
    if (tu.H != kh || tu.W != kw) {
      h->build_error = "head2 tail kernel does not cover the final map (heatmap size must be a multiple of 16)";
      return EGN_ERR_INVALID;
    }
    Op op;
    op.kind = Op::TAIL;
    op.in = u;
    op.wi = g.weights("head2.4", "", true, 2 * nj, 2 * nj, kh, tu.Cp, 2 * nj, true);
    h->weights.back().k = kh * 1000 + kw;  // non-square: encoded kh,kw
    h->ops.push_back(op);
    h->macs += (int64_t)(2 * nj) * (2 * nj) * kh * kw;
  }
  // conv weights written by head1 need the true input-channel count of head2.0 (nj + 2)
  return EGN_OK;
}

// liveness + greedy offset assignment (exact-size free lists)
static void plan_workspace(egn_hrnet* h) {
  auto& T = h->tensors;
  for (auto& t : T) {
    t.def_op = -1;
    t.last_use = -1;
    t.offset = -1;
  }
  for (int i = 0; i < (int)h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    auto use = [&](int id) {
      if (id >= 0) T[id].last_use = std::max(T[id].last_use, i);
    };
    if (op.out >= 0 && T[op.out].def_op < 0) T[op.out].def_op = i;
    use(op.in);
    use(op.res);
    for (int j = 0; j < op.nterms; ++j) use(op.term[j]);
  }
  std::multimap<int64_t, int64_t> free_list;  // size -> offset
  int64_t top = 0;
  std::vector<std::vector<int>> dying(h->ops.size());
  for (int id = 0; id < (int)T.size(); ++id)
    if (T[id].last_use >= 0) dying[T[id].last_use].push_back(id);
  for (int i = 0; i < (int)h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    if (op.out >= 0 && T[op.out].offset < 0) {
      auto it = h->cfg.keep_taps ? free_list.end() : free_list.find(T[op.out].per_crop);
      if (it != free_list.end()) {
        T[op.out].offset = it->second;
        free_list.erase(it);
      } else {
        T[op.out].offset = top;
        top += T[op.out].per_crop;
      }
    }
    // an op's output never aliases its own inputs: release inputs after allocating the output
    for (int id : dying[i])
      if (T[id].offset >= 0) free_list.insert({T[id].per_crop, T[id].offset});
  }
  h->ws_per_crop = top;
}

static void count_traffic(egn_hrnet* h) {
  // algorithmic activation bytes per crop: every op reads its inputs once, writes its output once
  const int64_t es = (int64_t)dtype_size(h->dt);
  int64_t bytes = 0;
  for (const Op& op : h->ops) {
    auto sz = [&](int id) { return id < 0 ? 0 : (int64_t)h->tensors[id].H * h->tensors[id].W * h->tensors[id].C * es; };
    if (op.kind == Op::STEM) bytes += (int64_t)h->cfg.in_channels * h->cfg.input_h * h->cfg.input_w * 4;
    bytes += sz(op.in) + sz(op.res) + sz(op.out);
    for (int j = 0; j < op.nterms; ++j) bytes += sz(op.term[j]);
  }
  h->act_bytes = bytes;
  h->n_launches = (int)h->ops.size();
}

// ---------------------------------------------------------------------------
// weight folding / packing
// ---------------------------------------------------------------------------
static const HostTensor* find_raw(const egn_hrnet* h, const std::string& k) {
  auto it = h->raw.find(k);
  return it == h->raw.end() ? nullptr : &it->second;
}

// -> folded [tap][Cin_p][Cout_p] fp32 + bias [Cout_p]; math in double
static int fold_conv(const egn_hrnet* h, const ConvWeights& w, int kh, int kw, bool round_fp16,
                     std::vector<float>* wf, std::vector<float>* bias) {
  const HostTensor* W = find_raw(h, w.conv_key + ".weight");
  if (!W) {
    set_error("weight '%s.weight' was never set", w.conv_key.c_str());
    return EGN_ERR_MISSING;
  }
  if ((int64_t)W->data.size() != (int64_t)w.Cout * w.Cin * kh * kw) {
    set_error("weight '%s.weight' has %zu elements, expected %lld", w.conv_key.c_str(), W->data.size(),
              (long long)w.Cout * w.Cin * kh * kw);
    return EGN_ERR_INVALID;
  }
  std::vector<double> scale(w.Cout, 1.0), shift(w.Cout, 0.0);
  if (!w.bn_key.empty()) {
    const HostTensor *g = find_raw(h, w.bn_key + ".weight"), *b = find_raw(h, w.bn_key + ".bias"),
                     *mu = find_raw(h, w.bn_key + ".running_mean"), *var = find_raw(h, w.bn_key + ".running_var");
    if (!g || !b || !mu || !var) {
      set_error("BatchNorm '%s' is incomplete (weight/bias/running_mean/running_var)", w.bn_key.c_str());
      return EGN_ERR_MISSING;
    }
    for (int o = 0; o < w.Cout; ++o) {
      scale[o] = (double)g->data[o] / std::sqrt((double)var->data[o] + 1e-5);
      shift[o] = (double)b->data[o] - (double)mu->data[o] * scale[o];
    }
  }
  if (w.has_bias) {
    const HostTensor* cb = find_raw(h, w.conv_key + ".bias");
    if (!cb) {
      set_error("weight '%s.bias' was never set", w.conv_key.c_str());
      return EGN_ERR_MISSING;
    }
    for (int o = 0; o < w.Cout; ++o) shift[o] += (double)cb->data[o] * scale[o];
  }
  const int taps = kh * kw;
  wf->assign((size_t)taps * w.Cin_p * w.Cout_p, 0.f);
  bias->assign(w.Cout_p, 0.f);
  for (int o = 0; o < w.Cout; ++o) {
    (*bias)[o] = (float)shift[o];
    for (int ci = 0; ci < w.Cin; ++ci)
      for (int tp = 0; tp < taps; ++tp) {
        float v = (float)((double)W->data[((size_t)o * w.Cin + ci) * taps + tp] * scale[o]);
        if (round_fp16) v = __half2float(__float2half_rn(v));
        (*wf)[((size_t)tp * w.Cin_p + ci) * w.Cout_p + o] = v;
      }
  }
  return EGN_OK;
}

static void free_device(egn_hrnet* h) {
  for (ConvWeights& w : h->weights) {
    cudaFree(w.d_simt);
    cudaFree(w.d_bias);
    w.d_simt = w.d_bias = nullptr;
    if (w.tc) tc_conv_plan_destroy(w.tc);
    w.tc = nullptr;
  }
  cudaFree(h->d_xs);
  cudaFree(h->d_ys);
  h->d_xs = h->d_ys = nullptr;
  h->finalized = false;
}

static ConvArgs conv_shape(const egn_hrnet* h, const Op& op, int B) {
  const TensorInfo& ti = h->tensors[op.in];
  const TensorInfo& to = h->tensors[op.out];
  const ConvWeights& w = h->weights[op.wi];
  ConvArgs a{};
  a.B = B;
  a.H = ti.H;
  a.W = ti.W;
  a.Cin_p = ti.Cp;
  a.OH = to.H;
  a.OW = to.W;
  a.Cout_p = to.Cp;
  a.Cout = w.Cout;
  a.ksize = w.k;
  a.stride = op.stride;
  a.pad = op.pad;
  a.relu = op.relu;
  a.coord_maps = op.coord_maps ? 1 : 0;
  a.split = h->dt == Dtype::F16X2 ? 1 : 0;
  return a;
}

// numpy.linspace(0, 1, n).astype(float32)  (hrnet.py:461-466)
static std::vector<float> linspace01(int n) {
  std::vector<float> v(n);
  const double step = n > 1 ? 1.0 / (double)(n - 1) : 0.0;
  for (int i = 0; i < n; ++i) v[i] = (float)((double)i * step);
  if (n > 1) v[n - 1] = 1.0f;
  return v;
}

}  // namespace egn

namespace egn {
static char* ws_base(void* workspace) {
  // 1 KB aligned base inside the caller's buffer
  uintptr_t p = reinterpret_cast<uintptr_t>(workspace);
  return reinterpret_cast<char*>((p + 1023) & ~uintptr_t(1023));
}

// Replays the op list on `st`.  If `events` is non-null it must hold ops.size()+1 events; one is
// recorded before every op and one after the last (per-op device timing for the roofline report).
static int run_ops(egn_hrnet* h, const float* x, int batch, float* heatmap_out, float* coords_out,
                   float* logits_out, void* workspace, cudaStream_t st, cudaEvent_t* events) {
  char* base = ws_base(workspace);
  const size_t es = dtype_size(h->dt);
  auto ptr = [&](int id) -> void* {
    return id < 0 ? nullptr : base + (size_t)h->tensors[id].offset * (size_t)batch * es;
  };
  int op_index = 0;
  for (const Op& op : h->ops) {
    if (events) cudaEventRecord(events[op_index], st);
    ++op_index;
    switch (op.kind) {
      case Op::STEM: {
        const TensorInfo& to = h->tensors[op.out];
        StemArgs a{};
        a.x = x;
        a.out = ptr(op.out);
        a.w = h->weights[op.wi].d_simt;
        a.bias = h->weights[op.wi].d_bias;
        a.B = batch;
        a.Cin = h->cfg.in_channels;
        a.H = h->cfg.input_h;
        a.W = h->cfg.input_w;
        a.OH = to.H;
        a.OW = to.W;
        if (int rc = launch_stem(h->dt, a, st)) return rc;
        break;
      }
      case Op::CONV: {
        ConvArgs a = conv_shape(h, op, batch);
        a.in = ptr(op.in);
        a.out = ptr(op.out);
        a.res = ptr(op.res);
        a.bias = h->weights[op.wi].d_bias;
        a.heatmap = op.write_heatmap ? heatmap_out : nullptr;
        a.xs = h->d_xs;
        a.ys = h->d_ys;
        const int rc = op.use_tc ? launch_conv_tc(h->weights[op.wi].tc, a, st)
                                 : launch_conv_simt(h->dt, a, h->weights[op.wi].d_simt, st);
        if (rc) return rc;
        break;
      }
      case Op::FUSE: {
        const TensorInfo& to = h->tensors[op.out];
        FuseArgs a{};
        a.out = ptr(op.out);
        a.nterms = op.nterms;
        for (int j = 0; j < op.nterms; ++j) {
          a.term[j] = ptr(op.term[j]);
          a.shift[j] = op.shift[j];
        }
        a.B = batch;
        a.H = to.H;
        a.W = to.W;
        a.Cp = to.Cp;
        if (int rc = launch_fuse(h->dt, a, st)) return rc;
        break;
      }
      case Op::TAIL: {
        if (!coords_out && !logits_out) break;
        const TensorInfo& ti = h->tensors[op.in];
        HeadTailArgs a{};
        a.in = ptr(op.in);
        a.w = h->weights[op.wi].d_simt;
        a.bias = h->weights[op.wi].d_bias;
        a.coords = coords_out;
        a.logits = logits_out;
        a.B = batch;
        a.L = ti.H * ti.W * ti.Cp;
        a.Cp = ti.Cp;
        a.Cout = h->weights[op.wi].Cout;
        if (int rc = launch_head_tail(h->dt, a, st)) return rc;
        break;
      }
    }
  }
  if (events) cudaEventRecord(events[op_index], st);
  return EGN_OK;
}
}  // namespace egn

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int egn_hrnet_create(const egn_hrnet_cfg* cfg, egn_hrnet** out) {
  using namespace egn;
  EGN_REQUIRE(cfg && out, "egn_hrnet_create: null argument");
  const egn_hrnet_cfg& c = *cfg;
  EGN_REQUIRE(c.in_channels >= 1 && c.in_channels <= 5, "in_channels must be 1..5 (got %d)", c.in_channels);
  EGN_REQUIRE(c.input_w >= 32 && c.input_h >= 32 && c.input_w % 32 == 0 && c.input_h % 32 == 0,
              "input size %dx%d must be a multiple of 32", c.input_w, c.input_h);
  EGN_REQUIRE(c.heatmap_w * 4 == c.input_w && c.heatmap_h * 4 == c.input_h,
              "heatmap_size must be input_size / 4 (pixel_shuffle is not supported)");
  EGN_REQUIRE(c.num_joints >= 1 && c.num_joints <= 120, "num_joints out of range");
  EGN_REQUIRE(c.head_type == EGN_HEAD_HEATMAP || c.head_type == EGN_HEAD_COORDINATES,
              "unsupported head_type %d (heatmap / coordinates only, as in the shipped configs)", c.head_type);
  EGN_REQUIRE(c.head_type != EGN_HEAD_HEATMAP || c.final_conv_kernel == 1 || c.final_conv_kernel == 3,
              "final_conv_kernel must be 1 or 3");
  EGN_REQUIRE(c.head_type != EGN_HEAD_COORDINATES || (c.heatmap_w % 16 == 0 && c.heatmap_h % 16 == 0),
              "coordinate head needs heatmap_size divisible by 16");
  EGN_REQUIRE(c.num_stages == 3, "num_stages must be 3 (stage2..stage4)");
  for (int s = 0; s < 3; ++s) {
    EGN_REQUIRE(c.stage_branches[s] == s + 2, "stage%d must have %d branches", s + 2, s + 2);
    EGN_REQUIRE(c.stage_modules[s] >= 1, "stage%d needs at least one module", s + 2);
    for (int b = 0; b < c.stage_branches[s]; ++b) {
      EGN_REQUIRE(c.stage_blocks[s][b] >= 1 && c.stage_channels[s][b] >= 1, "bad stage%d branch %d", s + 2, b);
      EGN_REQUIRE(s == 0 || b >= c.stage_branches[s - 1] || c.stage_channels[s][b] == c.stage_channels[s - 1][b],
                  "branch widths must stay constant across stages");
    }
  }
  EGN_REQUIRE(c.precision == EGN_PREC_FP32 || c.precision == EGN_PREC_FP16 || c.precision == EGN_PREC_FP16X2,
              "unknown precision %d", c.precision);
  EGN_REQUIRE(c.conv_impl == EGN_CONV_AUTO || c.conv_impl == EGN_CONV_SIMT, "unknown conv_impl %d", c.conv_impl);
  egn_hrnet* h = new egn_hrnet();
  h->cfg = c;
  h->dt = c.precision == EGN_PREC_FP32 ? Dtype::F32 : (c.precision == EGN_PREC_FP16 ? Dtype::F16 : Dtype::F16X2);
  build_keys(h);
  if (int rc = build_graph(h)) {
    set_error("egn_hrnet_create: %s", h->build_error.c_str());
    delete h;
    return rc;
  }
  plan_workspace(h);
  count_traffic(h);
  // which convs go to the tensor cores
  for (Op& op : h->ops) {
    if (op.kind != Op::CONV) continue;
    op.use_tc = c.precision != EGN_PREC_FP32 && c.conv_impl == EGN_CONV_AUTO &&
                tc_conv_supported(conv_shape(h, op, 1));
    if (op.use_tc) ++h->n_tc;
  }
  *out = h;
  return EGN_OK;
}

void egn_hrnet_destroy(egn_hrnet* h) {
  if (!h) return;
  egn::free_device(h);
  delete h;
}

int egn_hrnet_num_weights(const egn_hrnet* h) { return h ? (int)h->keys.size() : 0; }

const char* egn_hrnet_weight_key(const egn_hrnet* h, int i) {
  if (!h || i < 0 || i >= (int)h->keys.size()) return nullptr;
  return h->keys[i].c_str();
}

int egn_hrnet_weight_shape(const egn_hrnet* h, int i, int64_t shape[4]) {
  if (!h || i < 0 || i >= (int)h->keys.size()) return -1;
  const auto& s = h->key_shapes[i];
  for (size_t d = 0; d < s.size(); ++d) shape[d] = s[d];
  return (int)s.size();
}

int egn_hrnet_set_weight(egn_hrnet* h, const char* key, const float* host_data, const int64_t* shape, int ndim) {
  using namespace egn;
  EGN_REQUIRE(h && key, "egn_hrnet_set_weight: null argument");
  const std::string k(key);
  if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return EGN_OK;
  auto it = std::find(h->keys.begin(), h->keys.end(), k);
  EGN_REQUIRE(it != h->keys.end(), "egn_hrnet_set_weight: unexpected key '%s'", key);
  const auto& want = h->key_shapes[it - h->keys.begin()];
  EGN_REQUIRE(host_data && ndim == (int)want.size(), "egn_hrnet_set_weight: '%s' expects %zu dims, got %d", key,
              want.size(), ndim);
  int64_t n = 1;
  for (int d = 0; d < ndim; ++d) {
    EGN_REQUIRE(shape[d] == want[d], "egn_hrnet_set_weight: '%s' dim %d is %lld, expected %lld", key, d,
                (long long)shape[d], (long long)want[d]);
    n *= shape[d];
  }
  HostTensor t;
  t.data.assign(host_data, host_data + n);
  t.shape.assign(shape, shape + ndim);
  h->raw[k] = std::move(t);
  h->finalized = false;
  return EGN_OK;
}

int egn_hrnet_finalize(egn_hrnet* h) {
  using namespace egn;
  EGN_REQUIRE(h, "egn_hrnet_finalize: null handle");
  if (int rc = require_device()) return rc;
  free_device(h);
  h->weight_bytes = 0;
  // which weights does each op need, and in which form
  std::vector<int> need_simt(h->weights.size(), 0), need_tc(h->weights.size(), 0);
  std::vector<const Op*> owner(h->weights.size(), nullptr);
  for (const Op& op : h->ops) {
    if (op.wi < 0) continue;
    owner[op.wi] = &op;
    if (op.kind == Op::CONV && op.use_tc) need_tc[op.wi] = 1; else need_simt[op.wi] = 1;
  }
  for (size_t i = 0; i < h->weights.size(); ++i) {
    ConvWeights& w = h->weights[i];
    int kh = w.k, kw = w.k;
    if (w.k >= 1000) {
      kh = w.k / 1000;
      kw = w.k % 1000;
    }
    const bool r16 = h->dt == Dtype::F16 && !w.force_fp32;
    std::vector<float> wf, bias;
    if (int rc = fold_conv(h, w, kh, kw, r16, &wf, &bias)) return rc;
    EGN_CUDA_CHECK(cudaMalloc(&w.d_bias, bias.size() * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemcpy(w.d_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (owner[i] && owner[i]->kind == Op::TAIL) {
      // tail wants [Cout][kh*kw*Cin_p]
      const int L = kh * kw * w.Cin_p;
      std::vector<float> wt((size_t)w.Cout * L, 0.f);
      for (int tp = 0; tp < kh * kw; ++tp)
        for (int ci = 0; ci < w.Cin_p; ++ci)
          for (int o = 0; o < w.Cout; ++o)
            wt[(size_t)o * L + tp * w.Cin_p + ci] = wf[((size_t)tp * w.Cin_p + ci) * w.Cout_p + o];
      EGN_CUDA_CHECK(cudaMalloc(&w.d_simt, wt.size() * sizeof(float)));
      EGN_CUDA_CHECK(cudaMemcpy(w.d_simt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
      h->weight_bytes += (int64_t)wt.size() * 4;
      continue;
    }
    if (need_simt[i]) {
      EGN_CUDA_CHECK(cudaMalloc(&w.d_simt, wf.size() * sizeof(float)));
      EGN_CUDA_CHECK(cudaMemcpy(w.d_simt, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice));
      h->weight_bytes += (int64_t)wf.size() * 4;
    }
    if (need_tc[i]) {
      if (int rc = tc_conv_plan_create(conv_shape(h, *owner[i], 1), wf.data(), &w.tc)) return rc;
      h->weight_bytes += (int64_t)tc_conv_plan_weight_bytes(w.tc);
    }
  }
  if (h->cfg.head_type == EGN_HEAD_COORDINATES) {
    const std::vector<float> xs = linspace01(h->cfg.heatmap_w), ys = linspace01(h->cfg.heatmap_h);
    EGN_CUDA_CHECK(cudaMalloc(&h->d_xs, xs.size() * sizeof(float)));
    EGN_CUDA_CHECK(cudaMalloc(&h->d_ys, ys.size() * sizeof(float)));
    EGN_CUDA_CHECK(cudaMemcpy(h->d_xs, xs.data(), xs.size() * sizeof(float), cudaMemcpyHostToDevice));
    EGN_CUDA_CHECK(cudaMemcpy(h->d_ys, ys.data(), ys.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  EGN_CUDA_CHECK(cudaDeviceSynchronize());
  h->finalized = true;
  return EGN_OK;
}

size_t egn_hrnet_workspace_bytes(const egn_hrnet* h, int batch) {
  if (!h || batch <= 0) return 0;
  return (size_t)h->ws_per_crop * (size_t)batch * egn::dtype_size(h->dt) + 1024;
}


int egn_hrnet_forward(egn_hrnet* h, const float* x, int batch, float* heatmap_out, float* coords_out,
                      float* logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace egn;
  EGN_REQUIRE(h, "egn_hrnet_forward: null handle");
  EGN_REQUIRE(batch >= 0, "egn_hrnet_forward: negative batch");
  EGN_REQUIRE(batch == 0 || x, "egn_hrnet_forward: null input");
  if (!h->finalized) {
    set_error("egn_hrnet_forward called before egn_hrnet_finalize");
    return EGN_ERR_STATE;
  }
  if (int rc = require_device()) return rc;
  if (batch == 0) return EGN_OK;
  if (!workspace || workspace_bytes < egn_hrnet_workspace_bytes(h, batch)) {
    set_error("workspace too small: need %zu bytes for batch %d", egn_hrnet_workspace_bytes(h, batch), batch);
    return EGN_ERR_WORKSPACE;
  }
  EGN_REQUIRE(h->cfg.head_type == EGN_HEAD_COORDINATES || (!coords_out && !logits_out),
              "coords/logits outputs need the coordinate head");
  return egn::run_ops(h, x, batch, heatmap_out, coords_out, logits_out, workspace, as_stream(stream), nullptr);
}

int egn_hrnet_read_tap(egn_hrnet* h, const char* name, int batch, const void* workspace, float* out,
                       int dims[3], void* stream) {
  using namespace egn;
  EGN_REQUIRE(h && name && workspace && out, "egn_hrnet_read_tap: null argument");
  EGN_REQUIRE(h->cfg.keep_taps, "egn_hrnet_read_tap needs a handle created with keep_taps=1");
  auto it = h->taps.find(name);
  EGN_REQUIRE(it != h->taps.end(), "egn_hrnet_read_tap: unknown tap '%s'", name);
  if (int rc = require_device()) return rc;
  const TensorInfo& t = h->tensors[it->second];
  char* base = egn::ws_base(const_cast<void*>(workspace));
  const void* src = base + (size_t)t.offset * (size_t)batch * dtype_size(h->dt);
  if (dims) {
    dims[0] = t.C;
    dims[1] = t.H;
    dims[2] = t.W;
  }
  return launch_nhwc_to_nchw(h->dt, src, out, batch, t.H, t.W, t.Cp, t.C, as_stream(stream));
}

int egn_hrnet_profile(egn_hrnet* h, const float* x, int batch, void* workspace, size_t workspace_bytes,
                      void* stream, float* op_ms) {
  using namespace egn;
  EGN_REQUIRE(h && x && op_ms && batch > 0, "egn_hrnet_profile: bad argument");
  if (!h->finalized) {
    set_error("egn_hrnet_profile called before egn_hrnet_finalize");
    return EGN_ERR_STATE;
  }
  if (int rc = require_device()) return rc;
  if (!workspace || workspace_bytes < egn_hrnet_workspace_bytes(h, batch)) {
    set_error("workspace too small: need %zu bytes for batch %d", egn_hrnet_workspace_bytes(h, batch), batch);
    return EGN_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  std::vector<cudaEvent_t> ev(h->ops.size() + 1);
  for (auto& e : ev) EGN_CUDA_CHECK(cudaEventCreate(&e));
  // coords/logits scratch is not needed: the tail is skipped without outputs, so time it with a dummy
  int rc = run_ops(h, x, batch, nullptr, nullptr, nullptr, workspace, st, ev.data());
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("egn_hrnet_profile: %s", cudaGetErrorString(cudaGetLastError()));
    rc = EGN_ERR_CUDA;
  }
  if (!rc)
    for (size_t i = 0; i < h->ops.size(); ++i) cudaEventElapsedTime(&op_ms[i], ev[i], ev[i + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int egn_hrnet_op_info(const egn_hrnet* h, int i, egn_op_info* o) {
  using namespace egn;
  EGN_REQUIRE(h && o && i >= 0 && i < (int)h->ops.size(), "egn_hrnet_op_info: bad index");
  const Op& op = h->ops[i];
  const int64_t es = (int64_t)dtype_size(h->dt);
  auto sz = [&](int id) { return id < 0 ? (int64_t)0 : (int64_t)h->tensors[id].H * h->tensors[id].W * h->tensors[id].Cp * es; };
  *o = egn_op_info{};
  o->kind = (int)op.kind;
  o->use_tc = op.use_tc ? 1 : 0;
  o->stride = op.stride;
  o->has_res = op.res >= 0;
  if (op.out >= 0) {
    o->OH = h->tensors[op.out].H; o->OW = h->tensors[op.out].W; o->Cout = h->tensors[op.out].C;
  }
  if (op.in >= 0) {
    o->H = h->tensors[op.in].H; o->W = h->tensors[op.in].W; o->Cin = h->tensors[op.in].C;
  }
  o->act_bytes = sz(op.in) + sz(op.res) + sz(op.out);
  for (int j = 0; j < op.nterms; ++j) o->act_bytes += sz(op.term[j]);
  if (op.kind == Op::STEM) {
    o->Cin = h->cfg.in_channels; o->H = h->cfg.input_h; o->W = h->cfg.input_w; o->ksize = 3; o->stride = 2;
    o->act_bytes += (int64_t)h->cfg.in_channels * h->cfg.input_h * h->cfg.input_w * 4;
  }
  if (op.wi >= 0) {
    const ConvWeights& w = h->weights[op.wi];
    const int kh = w.k >= 1000 ? w.k / 1000 : w.k, kw = w.k >= 1000 ? w.k % 1000 : w.k;
    o->ksize = kh;
    const int64_t opix = op.kind == Op::TAIL ? 1 : (int64_t)o->OH * o->OW;
    o->macs = (int64_t)w.Cout * w.Cin * kh * kw * opix;
    o->weight_bytes = (int64_t)w.Cout_p * w.Cin_p * kh * kw * ((op.kind == Op::CONV && !w.force_fp32 && h->dt == Dtype::F16 && op.use_tc) ? 2 : 4);   // fp16x2: hi + lo = 4 bytes
    snprintf(o->name, sizeof(o->name), "%s", w.conv_key.c_str());
  } else if (op.kind == Op::FUSE) {
    snprintf(o->name, sizeof(o->name), "%s", h->tensors[op.out].tap.c_str());
  }
  return EGN_OK;
}

int64_t egn_hrnet_macs_per_crop(const egn_hrnet* h) { return h ? h->macs : 0; }
int egn_hrnet_num_launches(const egn_hrnet* h) { return h ? h->n_launches : 0; }
int egn_hrnet_num_tc_launches(const egn_hrnet* h) { return h ? h->n_tc : 0; }
int64_t egn_hrnet_act_bytes_per_crop(const egn_hrnet* h) { return h ? h->act_bytes : 0; }
int64_t egn_hrnet_weight_bytes(const egn_hrnet* h) { return h ? h->weight_bytes : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------
// single fused conv layer (per-layer parity tests; synchronous convenience call)
// ---------------------------------------------------------------------------
static int conv2d_fused_impl(int impl, int dtype, const void* in, const float* w_oihw_host,
                             const float* bias_host, const void* res, void* out, int B, int H, int W,
                             int Cin, int Cout, int ksize, int stride, int relu, void* stream, int iters,
                             float* avg_ms, float* acc_out = nullptr) {
  using namespace egn;
  EGN_REQUIRE(in && w_oihw_host && out, "egn_conv2d_fused: null pointer");
  EGN_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "egn_conv2d_fused: bad shape");
  EGN_REQUIRE(ksize == 1 || ksize == 3, "egn_conv2d_fused: ksize must be 1 or 3");
  EGN_REQUIRE(stride == 1 || stride == 2, "egn_conv2d_fused: stride must be 1 or 2");
  EGN_REQUIRE(dtype >= 0 && dtype <= 2, "egn_conv2d_fused: dtype 0 (fp32), 1 (fp16) or 2 (fp16x2)");
  EGN_REQUIRE(impl == 0 || (impl == 1 && dtype != 0), "egn_conv2d_fused: the tcgen05 path is fp16 / fp16x2 only");
  const Dtype dt = dtype == 0 ? Dtype::F32 : (dtype == 1 ? Dtype::F16 : Dtype::F16X2);
  if (int rc = require_device()) return rc;
  ConvArgs a{};
  a.B = B; a.H = H; a.W = W;
  a.Cin_p = round_up(Cin, kChanAlign);
  a.Cout_p = round_up(Cout, kChanAlign);
  a.Cout = Cout;
  a.ksize = ksize; a.stride = stride; a.pad = ksize == 3 ? 1 : 0; a.relu = relu;
  a.OH = (H + 2 * a.pad - ksize) / stride + 1;
  a.OW = (W + 2 * a.pad - ksize) / stride + 1;
  a.in = in; a.out = out; a.res = res;
  a.split = dtype == 2 ? 1 : 0;
  a.heatmap = acc_out;
  const int taps = ksize * ksize;
  std::vector<float> wf((size_t)taps * a.Cin_p * a.Cout_p, 0.f), bias(a.Cout_p, 0.f);
  for (int o = 0; o < Cout; ++o) {
    if (bias_host) bias[o] = bias_host[o];
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) {
        float v = w_oihw_host[((size_t)o * Cin + c) * taps + t];
        if (dtype == 1) v = __half2float(__float2half_rn(v));
        wf[((size_t)t * a.Cin_p + c) * a.Cout_p + o] = v;
      }
  }
  float* d_bias = nullptr;
  EGN_CUDA_CHECK(cudaMalloc(&d_bias, bias.size() * sizeof(float)));
  cudaMemcpy(d_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice);
  a.bias = d_bias;
  int rc = EGN_OK;
  cudaStream_t st = as_stream(stream);
  if (impl == 1) {
    if (!tc_conv_supported(a)) {
      set_error("egn_conv2d_fused: shape not supported by the tcgen05 kernel");
      rc = EGN_ERR_INVALID;
    } else {
      TcConvPlan* plan = nullptr;
      rc = tc_conv_plan_create(a, wf.data(), &plan);
      if (!rc) rc = launch_conv_tc(plan, a, st);
      if (!rc && iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < iters && !rc; ++i) rc = launch_conv_tc(plan, a, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (avg_ms) *avg_ms = ms / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
      }
      if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("egn_conv2d_fused(tc): %s", cudaGetErrorString(cudaGetLastError()));
        rc = EGN_ERR_CUDA;
      }
      tc_conv_plan_destroy(plan);
    }
  } else {
    float* d_w = nullptr;
    if (cudaMalloc(&d_w, wf.size() * sizeof(float)) != cudaSuccess) {
      set_error("egn_conv2d_fused: cudaMalloc failed");
      rc = EGN_ERR_CUDA;
    } else {
      cudaMemcpy(d_w, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice);
      rc = launch_conv_simt(dt, a, d_w, st);
      if (!rc && iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < iters && !rc; ++i) rc = launch_conv_simt(dt, a, d_w, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (avg_ms) *avg_ms = ms / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
      }
      if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("egn_conv2d_fused(simt): %s", cudaGetErrorString(cudaGetLastError()));
        rc = EGN_ERR_CUDA;
      }
      cudaFree(d_w);
    }
  }
  cudaFree(d_bias);
  return rc;
}

extern "C" int egn_conv2d_fused(int impl, int dtype, const void* in, const float* w_oihw_host,
                                const float* bias_host, const void* res, void* out, int B, int H, int W,
                                int Cin, int Cout, int ksize, int stride, int relu, void* stream) {
  return conv2d_fused_impl(impl, dtype, in, w_oihw_host, bias_host, res, out, B, H, W, Cin, Cout, ksize, stride,
                           relu, stream, 0, nullptr);
}

extern "C" int egn_conv2d_bench(int impl, int dtype, const void* in, const float* w_oihw_host,
                                const float* bias_host, const void* res, void* out, int B, int H, int W,
                                int Cin, int Cout, int ksize, int stride, int relu, void* stream, int iters,
                                float* avg_ms) {
  using namespace egn;
  EGN_REQUIRE(iters > 0 && avg_ms, "egn_conv2d_bench: iters must be positive");
  return conv2d_fused_impl(impl, dtype, in, w_oihw_host, bias_host, res, out, B, H, W, Cin, Cout, ksize, stride,
                           relu, stream, iters, avg_ms);
}

extern "C" int egn_debug_conv_acc(int impl, const void* in, const float* w_oihw_host, const float* bias_host,
                                  void* out, float* acc_out, int B, int H, int W, int Cin, int Cout, int ksize,
                                  int stride, void* stream) {
  using namespace egn;
  EGN_REQUIRE(acc_out, "egn_debug_conv_acc: null accumulator output");
  return conv2d_fused_impl(impl, 1, in, w_oihw_host, bias_host, nullptr, out, B, H, W, Cin, Cout, ksize, stride, 0,
                           stream, 0, nullptr, acc_out);
}
