// Launch interface of the CUDA kernels (host side).  One kernel per stage of the reference's
// forward pass; every launcher enqueues on `stream` and returns the launch status only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <utility>

#include "device_common.cuh"

namespace fdnn {

// Launch helper shared by the launchers: programmatic dependent launch (PDL) lets the next kernel's
// launch and prologue overlap the tail of the previous one; every kernel executes
// griddepcontrol.wait before it touches memory another kernel of the pass may own.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
bool pdl_enabled();  // FDNN_PDL=0 turns it off

// ---- fp32 input layer (input_layer.cu) ---------------------------------------------------------
struct InputLayerArgs {
  const float *in;     // [M][I] raw frames (not modified)
  const float *shift;  // [I]
  const float *scale;  // [I]
  const float *w0;     // [H][I]
  const float *bias0;  // [H]
  const uint8_t *lut;
  uint8_t *out_u8;  // [M][H]
  int M, I, H;
};
cudaError_t input_layer_configure();
int input_layer_max_dim();  // widest (padded) input the kernel takes
cudaError_t launch_input_layer(const InputLayerArgs &a, cudaStream_t stream);

// ---- int8 layers -------------------------------------------------------------------------------
// tcgen05 path (qlayer_tc.cu): needs K a multiple of 128 and, for hidden layers, N a multiple of 32.
cudaError_t qlayer_tc_configure();
bool qlayer_tc_supported(int N, int K, bool logits);
// Launch shape: N-tile width (64/128/256), cluster size (1/2/4) and which operand the cluster shares.
// The activation map must have a box of 128/cluster rows when share_a (else 128); the weight map a
// box of block_n rows when share_a (else block_n/cluster).
struct TcPlan {
  int block_n;
  int cluster;
  bool share_a;
  bool pair;  // CTA pairs (tcgen05 cta_group::2, qlayer_pair.cu): activation box 128 rows, weight box block_n/2 rows
};
TcPlan qlayer_tc_plan(int M, int N, bool logits, int num_sms, int policy = 0);  // policy: FDNN_POLICY_* (include/fdnn.h)
cudaError_t launch_qlayer_tc(const CUtensorMap &tmap_act, const CUtensorMap &tmap_w, const QLayerArgs &a, bool logits, TcPlan plan, int num_sms,
                             cudaStream_t stream);
// CTA-pair path (qlayer_pair.cu): two CTAs per 256×block_n tile sharing the weight tile.
cudaError_t qlayer_pair_configure();
cudaError_t launch_qlayer_pair(const CUtensorMap &tmap_act, const CUtensorMap &tmap_w, const QLayerArgs &a, bool logits, int block_n, int num_sms,
                               cudaStream_t stream);
// dp4a path (qlayer_simt.cu): any legal network (K a multiple of 16); used for narrow layers.
cudaError_t launch_qlayer_simt(const QLayerArgs &a, bool logits, cudaStream_t stream);

// ---- softmax over output rows (softmax.cu) -----------------------------------------------------
// out[r][:] = softmax(mask ? (mask[r][i] ? logits[r][i] : 0) : logits[r][i]) without max
// subtraction (dnn.cc:534-544, 355-392).  logits and out may alias.
struct SoftmaxArgs {
  const float *logits;  // [rows][ld]
  const int8_t *mask;   // [rows][mask_ld] or nullptr
  float *out;           // [rows][out_ld]
  int rows, O, ld, mask_ld, out_ld;
};
cudaError_t softmax_configure();
cudaError_t launch_softmax(const SoftmaxArgs &a, cudaStream_t stream);

}  // namespace fdnn
