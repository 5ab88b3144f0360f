// Internal declarations shared by the host core (model_host.cc), the C ABI (fdnn_api.cu) and the
// kernels.  Not part of the public boundary (include/fdnn.h).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace fdnn {

// Sigmoid LUT domain.  The reference's table covers k ∈ [−640, 640) with k ≤ −640 → 0 and
// k ≥ 640 → 255 handled by branches (dnn.h:35-42), where k = (int)round(x·100), round half away
// from zero.  The device table folds the rounding and both clamps into one lookup: with
// c = clamp(x·100, −641, 641) and v = trunc(2c) ∈ [−1282, 1282] (2c is exact),
//   |k| = (|v| + 1) >> 1,   lut2[v + 1282] = (k ≤ −640 ? 0 : k ≥ 640 ? 255 : lut[k + 640]).
constexpr int kLutHalf = 640;
constexpr int kLut2Center = 1282;
constexpr int kLut2Size = 2 * kLut2Center + 1;  // 2565
constexpr int kLut2Padded = 2576;               // multiple of 16 bytes

constexpr uint32_t kBlobMagic = 0x424E4446u;  // "FDNB"
constexpr uint32_t kBlobVersion = 6;
constexpr size_t kBlobAlign = 256;

// pmaddubsw (dnn.cc:337-340) clamps every adjacent-pair sum a[2p]·w[2p] + a[2p+1]·w[2p+1] to int16.
// A tensor-core contraction does not.  With a ≤ 255 the clamp can only fire for weight pairs whose
// same-sign magnitudes add up to ≥ 129; those (node, pair) "risk entries" are listed per layer.
// The layer kernel evaluates them against the activation tiles it is streaming anyway and adds
// clamp(v) − v to the raw tensor-core sums, which makes the sums bit-identical to the reference's.
//
// The list is stored three times, once per N-tile width G the tensor-core kernel can use
// (kFixGroups = 64, 128, 256), each ordered by (node / G, 128-byte K block of the pair, node, pair):
// a tile's entries are then one contiguous run, already in the order the K pipeline delivers the
// activation bytes.  ptr[g · k_blocks + kb] .. ptr[g · k_blocks + kb + 1] delimits the entries of
// node group g whose pair is in K block kb; ptr has n_groups · k_blocks + 1 elements.
constexpr int kFixVariants = 3;
constexpr int kFixGroups[kFixVariants] = {64, 128, 256};
constexpr int kFixKBlock = 128;  // bytes of K per block (= 64 pairs)
struct FixEntry {
  uint32_t pair_w;  // pair index p (bits 0-15) | (uint8)w[2p] << 16 | (uint8)w[2p+1] << 24
  uint32_t node;    // node n
};

// One int8 layer inside the blob; all offsets are from the start of the blob, 256-byte aligned.
struct BlobQLayer {
  int32_t nodes;      // N
  int32_t inputs;     // K (multiple of 16)
  float multiplier;   // round(127/max)                      dnn.cc:479
  float coeff;        // multiplier * 255.0f (fp32 product)  dnn.cc:297-298
  float rcp_coeff;    // RN(1 / coeff)
  uint32_t n_fix;     // saturation risk entries
  uint32_t fast_div;  // 1: q=s·rcp; r=fma(−q,coeff,s); q+=r·rcp verified == s/coeff for every reachable s
  uint32_t k_blocks;  // ceil(K / kFixKBlock)
  uint64_t off_w;     // int8  [N][K] row-major (K-major)
  uint64_t off_bias;  // fp32  [N]
  uint64_t off_fix_ptr[kFixVariants];  // uint32 [ceil(N/G)·k_blocks + 1] for G = 64, 128, 256
  uint64_t off_fix_ent[kFixVariants];  // FixEntry [n_fix], sorted by (node/G, K block, node, pair)
};

struct BlobHeader {
  uint32_t magic, version;
  uint64_t total_size;
  int32_t in_dim;       // padded to ×4
  int32_t in_dim_file;  // as stored
  int32_t hidden;       // H
  int32_t out_dim;      // O
  int32_t n_qlayers;
  float cutoff;
  uint64_t off_w0;     // fp32 [H][in_dim]
  uint64_t off_bias0;  // fp32 [H]
  uint64_t off_shift;  // fp32 [in_dim]
  uint64_t off_scale;  // fp32 [in_dim]
  uint64_t off_lut;    // u8 [kLut2Padded], index trunc(2·clamp(x·100)) + 1282
  uint64_t off_qlayers;  // BlobQLayer[n_qlayers]
};

// Thread-local error text behind fdnn_last_error().
void set_error(const std::string &msg);
const char *get_error();

// Parse dnn.bin + quantize + build LUT and fix-up lists into a relocatable blob.
// Returns FDNN_OK or a negative code (include/fdnn.h).
int pack_model(const char *path, float cutoff, std::vector<uint8_t> &blob);
// Structural validation of a blob (bounds, magic, constraints).
int validate_blob(const uint8_t *blob, size_t size);
// Offline tooling: network aligner/writer and feature-file IO.
int align_dnn_bin(const char *in_path, const char *out_path, int input_alignment, int hidden_alignment);
int import_kaldi_nnet1(const char *nnet_path, const char *transform_path, const char *out_path);
int read_feature_bin(const char *path, int *frames, int *dim, std::vector<float> &data);
int write_feature_bin(const char *path, const float *data, int frames, int dim);
int write_output_dump(const char *path, const float *data, int frames, int dim);
// Reference LUT (1280 entries, dnn.cc:100-115).
void build_reference_lut(uint8_t out[1280]);

inline const BlobHeader *blob_header(const uint8_t *blob) { return reinterpret_cast<const BlobHeader *>(blob); }
inline const BlobQLayer *blob_qlayers(const uint8_t *blob) {
  return reinterpret_cast<const BlobQLayer *>(blob + blob_header(blob)->off_qlayers);
}

}  // namespace fdnn
