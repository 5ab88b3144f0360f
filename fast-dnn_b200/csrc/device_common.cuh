// Device-side pieces shared by all kernels: the exact scalar arithmetic of the reference's
// per-element stages, and the producer/consumer halves of the pmaddubsw-saturation correction.
#pragma once

#include <cstdint>

#include "fdnn_internal.h"

namespace fdnn {

// Correction channel of one int8 layer.  Producers (the kernel writing the layer's INPUT
// activations) post clamp(v) − v for the few weight pairs that can saturate; the layer's own
// kernel adds them to its raw tensor-core sums.  `corr` is kept all-zero between uses: whoever
// consumes a non-zero flag re-zeroes what it read.  Granularity: kFixChunk (16) nodes per flag.
struct CorrChannel {
  int32_t *corr;   // [rows][ld] int32
  uint8_t *flags;  // [ld/16][rows_cap]: non-zero ⇒ corr[row][16c .. 16c+15] holds something
  int ld;          // corr row pitch (elements, multiple of 16)
  int rows_cap;    // flag pitch
};

// Risk list of the consumer layer, seen from the producer (BlobQLayer::off_fix_*).
struct FixList {
  const uint32_t *ptr;  // [n_chunks + 1], chunk = 16 consecutive inputs of the consumer
  const FixEntry *ent;
};

struct QLayerArgs {
  const uint8_t *act;  // [M][K] u8, row-major
  const int8_t *w;     // [N][K] s8, row-major
  const float *bias;   // [N]
  const uint8_t *lut;  // doubled sigmoid LUT (kLut2Padded bytes)
  float coeff, rcp;
  int fast_div;
  int M, N, K;
  CorrChannel self;   // corrections addressed to this layer
  // hidden mode: u8 activations out + corrections for the next layer
  uint8_t *out_u8;  // [M][N]
  FixList next_fix;
  CorrChannel next;
  // logits mode: lin + bias, fp32 [M][out_ld]
  float *out_f32;
  int out_ld;
  // optional per-CTA phase timestamps (SM clock), 8 slots per CTA; nullptr in normal operation
  unsigned long long *timeline;
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

// Producer half: `a` points at this row's 16 freshly written activations of input chunk `chunk`
// (shared memory); evaluates entries [e0, e1) of the consumer's risk list.  `ent` may point to
// shared or global memory; entry e lives at ent[e − ent_base].
__device__ __forceinline__ void post_saturation(const uint8_t *a, int chunk, int row, const FixEntry *ent, uint32_t ent_base, uint32_t e0,
                                                uint32_t e1, const CorrChannel &ch) {
  for (uint32_t e = e0; e < e1; ++e) {
    const uint2 fe = *reinterpret_cast<const uint2 *>(ent + (e - ent_base));
    const int off = 2 * int(fe.x & 0xffffu) - kFixChunk * chunk;
    const int w0 = int(int8_t(fe.x >> 16)), w1 = int(int8_t(fe.x >> 24));
    const int v = int(a[off]) * w0 + int(a[off + 1]) * w1;
    const int d = max(min(v, 32767), -32768) - v;
    if (d != 0) {
      atomicAdd(ch.corr + size_t(row) * size_t(ch.ld) + fe.y, d);
      ch.flags[size_t(fe.y >> 4) * size_t(ch.rows_cap) + size_t(row)] = 1;
    }
  }
}

__device__ __forceinline__ uint8_t load_flag(const CorrChannel &ch, int node_chunk, int row) {
  return ch.flags[size_t(node_chunk) * size_t(ch.rows_cap) + size_t(row)];
}

// Consumer half for one row and one aligned chunk of 16 nodes held in registers; `flag` is the
// (possibly prefetched) flag byte of that (chunk, row).
__device__ __forceinline__ void take_corrections(int32_t (&s)[16], uint8_t flag, int node_chunk, int row, const CorrChannel &ch) {
  if (flag) {
    int4 *c = reinterpret_cast<int4 *>(ch.corr + size_t(row) * size_t(ch.ld) + size_t(node_chunk) * kFixChunk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int4 v = c[j];
      s[4 * j + 0] += v.x;
      s[4 * j + 1] += v.y;
      s[4 * j + 2] += v.z;
      s[4 * j + 3] += v.w;
      c[j] = make_int4(0, 0, 0, 0);
    }
    ch.flags[size_t(node_chunk) * size_t(ch.rows_cap) + size_t(row)] = 0;
  }
}

// The reference's per-element tail for one row and one aligned chunk of 16 nodes whose corrected
// sums are in `s`: dequantise → + bias → {LUT → u8 | fp32 logits} → global memory.  Returns the
// 16 packed bytes in hidden mode (for the saturation scan of the next layer).
template <bool kLogits>
__device__ __forceinline__ uint4 finish_chunk(const int32_t (&s)[16], int row, int col, const QLayerArgs &args, const float *bias16,
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
    return make_uint4(0, 0, 0, 0);
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
    const uint4 out = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    // hidden widths are multiples of 16 (dnn.cc:331), so a chunk is always whole and 16-byte aligned
    *reinterpret_cast<uint4 *>(args.out_u8 + size_t(row) * size_t(N) + col) = out;
    return out;
  }
}

}  // namespace fdnn
