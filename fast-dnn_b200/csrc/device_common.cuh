// Device-side pieces shared by the kernels: the exact scalar arithmetic of the reference's
// per-element stages and the evaluation of pmaddubsw-saturation risk entries.
#pragma once

#include <cstdint>

#include "fdnn_internal.h"

namespace fdnn {

// Risk list of one int8 layer, in the variant grouped by `group` nodes (BlobQLayer::off_fix_*;
// order described in fdnn_internal.h).
struct FixList {
  const uint32_t *ptr;  // [ceil(N/group) · k_blocks + 1]
  const FixEntry *ent;
  int k_blocks;
  int group;  // 64, 128 or 256
};

struct QLayerArgs {
  const uint8_t *act;  // [M][K] u8, row-major
  const int8_t *w;     // [N][K] s8, row-major
  const float *bias;   // [N]
  const uint8_t *lut;  // doubled sigmoid LUT (kLut2Padded bytes)
  float coeff, rcp;
  int fast_div;
  int M, N, K;
  FixList fix;      // this layer's saturation risk entries
  uint8_t *out_u8;  // hidden mode: u8 activations [M][N]
  float *out_f32;   // logits mode: lin + bias, fp32 [M][out_ld]
  int out_ld;
  // optional per-CTA phase timestamps (SM clock), 8 slots per CTA; nullptr in normal operation
  unsigned long long *timeline;
  int debug_flags;  // profiling experiments only (FDNN_DEBUG): 1 = scan warps skip their entries (results wrong)
};

__device__ __forceinline__ void stamp(unsigned long long *timeline, int slot) {
  if (timeline != nullptr) timeline[size_t(blockIdx.x) * 8 + slot] = clock64();
}

// (float)sum / (multiplier·255)  — dnn.cc:296-311.  IEEE division, or the 3-op form proven equal
// to it for every reachable sum when the model was packed (model_host.cc: verify_fast_div).
__device__ __forceinline__ float dequant(int32_t sum, float coeff, float rcp, int fast_div) {
  float s = __int2float_rn(sum);
  if (fast_div) {
    float q = __fmul_rn(s, rcp);
    float e = __fmaf_rn(-q, coeff, s);
    return __fmaf_rn(e, rcp, q);
  }
  return __fdiv_rn(s, coeff);
}

// Slot in the doubled LUT for QuantizedSigmoid::get(x) (dnn.h:35-42): k = (int)round(x·100), round
// half away from zero, k ≤ −640 → 0, k ≥ 640 → 255.  c + c is exact, and truncating it keeps
// exactly the information round-half-away needs (fdnn_internal.h).  Products that do not fit an
// int32, and NaN, convert to INT_MIN on x86 (cvttss2si), i.e. land in the 0 bucket.
__device__ __forceinline__ int qsig_slot(float x) {
  const float t = __fmul_rn(x, 100.0f);
  float c = fminf(fmaxf(t, -641.0f), 641.0f);  // NaN → −641
  if (!(t < 2147483648.0f)) c = -641.0f;
  return __float2int_rz(__fadd_rn(c, c)) + kLut2Center;
}

// clamp(v) − v for one pair sum v = a0·w0 + a1·w1 (what pmaddubsw's int16 saturation changes)
__device__ __forceinline__ int saturation_delta(uint32_t a01, uint32_t pair_w) {
  const int w0 = int(int8_t(pair_w >> 16)), w1 = int(int8_t(pair_w >> 24));
  const int v = int(a01 & 0xffu) * w0 + int(a01 >> 8) * w1;
  return max(min(v, 32767), -32768) - v;
}

// Slow, self-contained evaluation of every risk entry that touches nodes [col, col+16) of `row`,
// reading the activation bytes from global memory.  Used by the dp4a kernel and as the overflow
// path of the tensor-core kernel.
__device__ __forceinline__ void brute_force_corrections(int32_t (&s)[16], int row, int col, const QLayerArgs &args) {
  const int kb_n = args.fix.k_blocks;
  const int sg = col / args.fix.group;
  const uint32_t e0 = __ldg(args.fix.ptr + size_t(sg) * kb_n), e1 = __ldg(args.fix.ptr + size_t(sg + 1) * kb_n);
  const uint8_t *a_row = args.act + size_t(row) * size_t(args.K);
  for (uint32_t e = e0; e < e1; ++e) {
    const uint2 fe = __ldg(reinterpret_cast<const uint2 *>(args.fix.ent) + e);
    const uint32_t rel = fe.y - uint32_t(col);
    if (rel < 16u) {
      const uint32_t a01 = *reinterpret_cast<const uint16_t *>(a_row + 2 * (fe.x & 0xffffu));
      const int d = saturation_delta(a01, fe.x);
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] += (rel == uint32_t(i)) ? d : 0;
    }
  }
}

// The reference's per-element tail for one row and one aligned chunk of 16 nodes whose exact
// (saturation-corrected) sums are in `s`: dequantise → + bias → {LUT → u8 | fp32 logits} → global.
template <bool kLogits>
__device__ __forceinline__ void finish_chunk(const int32_t (&s)[16], int row, int col, const QLayerArgs &args, const float *bias16,
                                             const uint8_t *lut) {
  const int N = args.N;
  if constexpr (kLogits) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(dequant(s[i], args.coeff, args.rcp, args.fast_div), bias16[i]);
    float *dst = args.out_f32 + size_t(row) * size_t(args.out_ld) + col;
    if (col + 16 <= N && (args.out_ld & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4 *>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (col + i < N) dst[i] = v[i];
    }
  } else {
    uint32_t packed[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float x = __fadd_rn(dequant(s[4 * i + b], args.coeff, args.rcp, args.fast_div), bias16[4 * i + b]);
        w |= uint32_t(lut[qsig_slot(x)]) << (8 * b);
      }
      packed[i] = w;
    }
    // hidden widths are multiples of 16 (dnn.cc:331), so a chunk is always whole and 16-byte aligned
    *reinterpret_cast<uint4 *>(args.out_u8 + size_t(row) * size_t(N) + col) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

}  // namespace fdnn
