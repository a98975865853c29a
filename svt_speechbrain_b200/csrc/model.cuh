// Host-side model objects behind the C ABI: weight registry (reference state_dict names),
// packing into kernel layouts, workspace planning, forward orchestration.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/svt_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "ops.cuh"

namespace svt {

struct RawTensor {
  float* dev = nullptr;  // fp32 copy on the device
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto d : shape) n *= static_cast<size_t>(d);
    return n;
  }
};

// owns device allocations made while packing
class DevicePool {
 public:
  ~DevicePool() { release(); }
  int alloc(size_t bytes, void** out);
  template <typename T>
  int alloc_t(size_t n, T** out) { return alloc(n * sizeof(T), reinterpret_cast<void**>(out)); }
  void release();

 private:
  std::vector<void*> ptrs_;
};

class WeightRegistry {
 public:
  ~WeightRegistry() { clear(); }
  int set(const std::string& name, const float* host, const int64_t* shape, int ndim);
  const RawTensor* find(const std::string& name) const;
  int require(const std::string& name, std::initializer_list<int64_t> shape, const RawTensor** out) const;
  void clear();

 private:
  std::map<std::string, RawTensor> t_;
};

struct LinearW {
  __nv_bfloat16* w = nullptr;  // [N, K]
  float* b = nullptr;          // [N]
  int N = 0, K = 0;
  float* colsum = nullptr;     // [N], LayerNorm-folded weights only (GemmArgs::ln_colsum); b then holds beta.W + b
};
struct NormW {
  float* g = nullptr;
  float* b = nullptr;
};

}  // namespace svt

// ---------------------------------------------------------------------------------------------
namespace svt {
int pack_posconv_weight(const float* w_dev, const float* tap_scale, int D, int groups, int taps, __nv_bfloat16* dst,
                        cudaStream_t stream);
}

namespace svt {
// scratch of the shared transformer body (rows M = clips * Ta): h fp32 [M, D] residual stream (in / out), hb bf16 copy
// [M, D] (in), qkv bf16 [M, 3D], ctx bf16 [M, D], mid bf16 [M, F], pre fp32 [M, D], rowstats fp32
// transformer_rowstats_bytes() (per-row sum / sum of squares links of the folded-LayerNorm chain)
struct TransformerBuffers {
  float* rowstats = nullptr;
  float* gate = nullptr;  // WavLM: [M][heads] fp32
  float* h = nullptr;
  __nv_bfloat16* hb = nullptr;
  __nv_bfloat16* qkv = nullptr;
  __nv_bfloat16* ctx = nullptr;
  __nv_bfloat16* mid = nullptr;
  float* pre = nullptr;
};
}  // namespace svt

struct svt_encoder {
  ~svt_encoder() {
    if (head_w != nullptr) cudaFree(head_w);
    if (head_b != nullptr) cudaFree(head_b);
    drop_rel_tabs();
  }
  void drop_rel_tabs() const {
    for (auto& kv : rel_tabs) cudaFree(kv.second);
    rel_tabs.clear();
  }
  svt_encoder_config cfg{};
  svt::WeightRegistry reg;
  svt::DevicePool pool;
  bool finalized = false;
  bool norm_per_clip = false;     // whole-tensor norms use one (mean, var) per clip instead of one per call
  bool transformer_only = false;  // AV-HuBERT video stream: only the positional conv + layers are packed

  // packed weights
  float* conv0_w = nullptr;  // [k][C] fp32
  float* conv0_b = nullptr;
  void* conv0_tab = nullptr;  // tables of the tensor-core conv0 kernel (layer-norm feature extractors)
  svt::NormW conv0_norm;
  std::vector<svt::LinearW> conv;   // layers 1..n-1: [C, k*C_in]
  std::vector<svt::NormW> conv_norm;
  svt::NormW proj_norm;
  svt::LinearW proj;
  __nv_bfloat16* pos_w = nullptr;  // [G][taps][64][64]
  float* pos_b = nullptr;
  struct PosLayer { __nv_bfloat16* w = nullptr; float* b = nullptr; };
  std::vector<PosLayer> pos_stack;  // data2vec-audio: cfg.pos_conv_layers x (conv weights as pos_w, bias)
  float* pos_bn_scale = nullptr;    // HuBERT conv_pos_batch_norm: gamma / sqrt(running_var + eps), beta - mean * that
  float* pos_bn_shift = nullptr;
  float* ones = nullptr;            // [D] affine of the stack's parameter-free LayerNorms
  float* zeros = nullptr;
  svt::NormW enc_norm;
  struct Layer {
    svt::NormW ln1, ln2;
    svt::LinearW qkv, out, ff1, ff2;
    svt::LinearW qkv_ln, ff1_ln;  // pre-LN models: gamma / beta of ln1 / ln2 folded in (option "ln_fold")
    // WavLM: gru_rel_pos_linear with its 2 x 4 output groups summed ([2][64] + [2]) and gru_rel_pos_const [H]
    float* gate_w2 = nullptr;
    float* gate_b2 = nullptr;
    float* gate_const = nullptr;
    float* gate_w2g = nullptr;  // with ln1 folded in: [H][2][64], [H][2], [H][2]
    float* gate_cg = nullptr;
    float* gate_dg = nullptr;
  };
  std::vector<float> rel_embed;                 // WavLM: layers.0.attention.rel_attn_embed.weight [buckets][H] (host)
  mutable std::map<int, float*> rel_tabs;       // T -> device table [H][stride] of the Toeplitz position bias (own cudaMalloc,
                                                // at most kMaxRelTabs clip lengths cached)
  static constexpr size_t kMaxRelTabs = 32;
  std::vector<Layer> layers;
  float* head_w = nullptr;  // [n_out, D] fp32
  float* head_b = nullptr;
  int head_n = 0;

  // geometry helpers
  int conv_out_len(int L, int upto) const;  // valid frames after conv layers 0..upto
  int t_alloc0(int L) const;                // allocated frames of layer 0 (multiple of prod(strides[1:]) and 4)
};

namespace svt {
int encoder_finalize(svt_encoder* e);
size_t transformer_rowstats_bytes(const svt_encoder* e, size_t M);
int encoder_transformer_forward(const svt_encoder* e, int B, int T, int Ta, const TransformerBuffers& tb, double* stats_out,
                                int stats_stride, const float** final_x_out, cudaStream_t s);
}  // namespace svt

struct svt_fusion {
  svt_fusion_config cfg{};
  svt::WeightRegistry reg;
  svt::DevicePool pool;
  bool finalized = false;
  struct Layer {
    svt::LinearW qkv;   // [3D, D] (q rows pre-scaled by d_h^-0.5)
    svt::LinearW out2;  // [D, 2D] = [alpha*Wo | (1-alpha)*Wo], bias bo
    svt::NormW n1, n2;
    svt::LinearW ff1, ff2;
  };
  Layer layer[2];
};
