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

// ---- fp32 input layer, certified on the tensor cores (input_tc.cu) ------------------------------
constexpr int kInputTcMaxI = 512;   // widest (padded) input the fixed-point path takes
constexpr int kInputTcPitch = 512;  // bytes per row of a limb plane (K padded with zeros)
// How many fp32 roundings term k of a frame·weight-row dot product goes through in the reference (dnn.cc:219-247,
// 168-172): its product, the adds of its SSE lane (k mod 4) that come after it, the two adds that combine the lanes;
// + 1 for safety.  I = padded input width (multiple of 4).
__host__ __device__ inline double input_round_count(int k, int I) {
  const int n = I / 4, q = k / 4;
  return double((q == 0 ? n - 1 : n - q) + 3 + 1);
}
// Certificate constants (all fp32, every bound rounded UP; derivation in input_tc.cu).
struct InputRowStats {   // per frame, written by the prepare kernel
  float a[5];            // 100 · 2^(e−22) · 256^s: weight of shift class s in z (exact); a[0] < 0 = certify nothing in this row
  float p;               // 100 · u · ‖√c · x'‖₂                     (the reference's own rounding, c = input_round_count)
  float pp;              // 100 · 7u · (‖x'‖₂ + 2^(e−22)·2¹⁷·√I)      (fp32 evaluation of z from the class sums)
  float r1;              // 100 · 2^(e−22) · ½ Σ|X_k|                 (quantisation of the weights)
  float ar;              // 100 · 2^(e−22)                            (… of the frame, with InputNodeStats::e)
  float pad[3];
};
struct InputNodeStats {  // per node of layer 0, built when the model is uploaded
  float c;               // 2^(e_w−22) (exact); < 0 = certify nothing for this node
  float bc;              // fl(bias · 100)
  float q;               // ‖√c · w‖₂
  float qp;              // ‖w‖₂ + 2^(e_w−22)·2¹⁷·√I
  float e;               // 2^(e_w−22) · (½ Σ|W_k| + ¼ I + 255 · Σ (W_k mod 256))
  float f;               // u·|bc| + 2.1u + 1e-9: rounding of bias·100, of the reference's "+ bias" and "· 100" at |z| = 0
  float pad[2];
};
struct InputTcArgs {
  const float *in;      // [M][I] raw frames
  const float *shift, *scale;
  const float *w0;      // [H][I] fp32 (exact path)
  const float *bias0;   // [H]
  const uint8_t *lut;   // doubled sigmoid LUT
  const InputNodeStats *node_stats;  // [H]
  float *xq;            // [M][I] transformed frames (scratch, read by the exact path)
  uint8_t *x_limbs;     // 3 planes of [x_plane_rows][kInputTcPitch] bytes
  size_t x_plane;       // bytes per plane
  int x_plane_rows, w_plane_rows;
  InputRowStats *row_stats;  // [M]
  uint32_t *unc_bits;   // [M][unc_words]: bit n%32 of word n/32 = node n of this frame is left to the exact path
  int unc_words;        // ceil(H / 32)
  uint32_t *unc_t;      // [ceil(M / 32)][H]: the same bits by blocks of 32 frames: bit f%32 of word [f/32][n] (block fix-up kernel)
  uint32_t *unc_count;  // how many bits are set (diagnostics)
  uint8_t *out_u8;      // [M][H]
  int M, I, H;
  int fixup_ctas;       // upper limit for the exact kernel's grid
  int num_sms;
};
cudaError_t input_tc_configure();
bool input_tc_supported(int I, int H);
cudaError_t launch_input_tc(const CUtensorMap &tmap_x, const CUtensorMap &tmap_w, const InputTcArgs &a, cudaStream_t stream);

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
// Fused path (qlayer_fused.cu): every int8 layer of the pass and the softmax in one persistent, cooperatively launched kernel,
// for batches whose layers are a single wave of tiles.
constexpr int kFusedMaxLayers = 16;     // int8 layers per network the fused kernel takes
constexpr int kFusedMaxRowBlocks = 64;  // 128-frame row blocks per launch (one progress counter each)
constexpr int kFusedSyncWords = kFusedMaxRowBlocks + 2;  // + exit counter + tile counter
struct FusedLayer {
  const float *bias;        // [N]
  const uint32_t *fix_ptr;  // risk list grouped by this layer's tile width (fdnn_internal.h)
  const FixEntry *fix_ent;
  float coeff, rcp;
  int fast_div, fast_tail;
  int N, K;
  uint32_t need;            // tiles of a row block that must have been finished (all earlier layers) before this layer reads it
  uint32_t tile_begin;      // number of this layer's first tile in the global (layer, row block, column block) order
  uint32_t n_blocks;        // column blocks of this layer at its tile width
};
struct FusedArgs {
  CUtensorMap act[2];                 // the two u8 activation buffers [M][H] as 3-D maps (128 B of K, row, K block): box 128 rows × 2 K blocks
  CUtensorMap w[kFusedMaxLayers];     // s8 weights [N][K], 3-D likewise: box BNH rows (hidden layers) / 256 rows (output layer) × 2 K blocks
  FusedLayer layer[kFusedMaxLayers];
  int n_layers, M;
  uint8_t *act_buf[2];
  float *out;                         // [M][out_ld]: logits, normalised in place when do_softmax
  int out_ld;
  int do_softmax;
  const uint8_t *lut;
  float one, neg_zero;
  uint32_t *sync;                     // [kFusedSyncWords] zero-initialised; the kernel leaves it zeroed
  uint32_t total_tiles;
  uint32_t tiles_per_row_block;       // Σ over layers of the number of N tiles: a row block has left the output layer
  int debug_flags;                    // FDNN_FUSED_DEBUG: timing experiments (2 no proxy fences, 4 no extra fence before the release, 8 poll without sleeping, 16 spin on the accumulator barrier; RESULTS WRONG: 1 no scan entries, 32 no MMAs, 64 no activation tiles)
  unsigned long long *timeline;       // optional [layers][1024][8] nanosecond stamps (fdnn_ctx_timeline)
};
cudaError_t qlayer_fused_configure();
int qlayer_fused_plan(int M, int H, int O, int num_sms, int policy, int *grid);  // → hidden tile width (64 / 128) or 0 = not applicable
int qlayer_fused_max_softmax_width();
bool qlayer_fused_softmax_pays(int M, int grid);
cudaError_t launch_qlayer_fused(const FusedArgs &a, int bnh, int grid, cudaStream_t stream);

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
