// Device-side pieces shared by all kernels: the exact scalar arithmetic of the reference's
// per-element stages, and the producer/consumer halves of the pmaddubsw-saturation correction.
#pragma once

#include <cstdint>

#include "fdnn_internal.h"

namespace fdnn {

// Correction channel of one int8 layer.  Producers (the kernel writing the layer's INPUT
// activations) post clamp(v) − v for the few weight pairs that can saturate; the layer's own
// kernel adds them to its raw tensor-core sums.  `corr` is kept all-zero between uses: whoever
// consumes a non-zero flag re-zeroes what it read.
struct CorrChannel {
  int32_t *corr;   // [rows][ld] int32
  uint8_t *flags;  // [ceil(N/32)][rows_cap]: non-zero ⇒ corr[row][32c .. 32c+31] holds something
  int ld;          // corr row pitch (elements)
  int rows_cap;    // flag pitch
};

// Risk list of the consumer layer, seen from the producer (BlobQLayer::off_fix_*).
struct FixList {
  const uint32_t *ptr;  // [n_chunks + 1]
  const FixEntry *ent;
};

struct QLayerArgs {
  const uint8_t *act;  // [M][K] u8, row-major
  const int8_t *w;     // [N][K] s8, row-major
  const float *bias;   // [N]
  const uint8_t *lut;  // extended sigmoid LUT (kLutExtPadded bytes)
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
};

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

// Index into the extended LUT for QuantizedSigmoid::get(x) (dnn.h:35-42):
// k = (int)round(x·100) with round-half-away; k ≤ −640 → 0, k ≥ 640 → 255.  Values whose
// rounded product does not fit an int32 (and NaN) convert to INT_MIN on x86, i.e. the 0 bucket.
__device__ __forceinline__ int qsig_index(float x) {
  float t = __fmul_rn(x, 100.0f);
  float c = fminf(fmaxf(t, -641.0f), 641.0f);
  if (!(t < 2147483648.0f)) c = -641.0f;
  float f = truncf(c);
  float d = __fsub_rn(c, f);  // exact
  if (fabsf(d) >= 0.5f) f += copysignf(1.0f, c);
  return __float2int_rz(f) + 641;
}

// Producer half: `a` points at this row's 32 freshly written activations of input chunk `chunk`
// (readable memory, usually shared); evaluates the consumer's risk entries of that chunk.
__device__ __forceinline__ void post_saturation(const uint8_t *a, int chunk, int row, const FixList &fix, const CorrChannel &ch) {
  const uint32_t e0 = __ldg(fix.ptr + chunk), e1 = __ldg(fix.ptr + chunk + 1);
  for (uint32_t e = e0; e < e1; ++e) {
    const uint2 fe = __ldg(reinterpret_cast<const uint2 *>(fix.ent) + e);
    const int off = 2 * int(fe.x & 0xffffu) - kFixChunk * chunk;
    const int w0 = int(int8_t(fe.x >> 16)), w1 = int(int8_t(fe.x >> 24));
    const int v = int(a[off]) * w0 + int(a[off + 1]) * w1;
    const int d = max(min(v, 32767), -32768) - v;
    if (d != 0) {
      atomicAdd(ch.corr + size_t(row) * size_t(ch.ld) + fe.y, d);
      ch.flags[size_t(fe.y >> 5) * size_t(ch.rows_cap) + size_t(row)] = 1;
    }
  }
}

// Consumer half for one row and one aligned chunk of 32 nodes held in registers.
__device__ __forceinline__ void take_corrections(int32_t (&s)[32], int node_chunk, int row, const CorrChannel &ch) {
  uint8_t *flag = ch.flags + size_t(node_chunk) * size_t(ch.rows_cap) + size_t(row);
  if (*flag) {
    int4 *c = reinterpret_cast<int4 *>(ch.corr + size_t(row) * size_t(ch.ld) + size_t(node_chunk) * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int4 v = c[j];
      s[4 * j + 0] += v.x;
      s[4 * j + 1] += v.y;
      s[4 * j + 2] += v.z;
      s[4 * j + 3] += v.w;
      c[j] = make_int4(0, 0, 0, 0);
    }
    *flag = 0;
  }
}

// The reference's per-element tail for one row and one aligned chunk of 32 nodes whose raw
// (unsaturated) sums are in `s`:  + corrections → dequantise → + bias → {LUT → u8 | fp32 logits}.
// `bias32` points at the 32 biases of the chunk, `lut` at the extended LUT (both any address
// space), `scan` at 32 bytes of thread-private scratch (shared memory).
template <bool kLogits>
__device__ __forceinline__ void epilogue_chunk(int32_t (&s)[32], int row, int col, const QLayerArgs &args, const float *bias32,
                                               const uint8_t *lut, uint8_t *scan) {
  const int N = args.N;
  take_corrections(s, col >> 5, row, args.self);
  if constexpr (kLogits) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __fadd_rn(dequant(s[i], args.coeff, args.rcp, args.fast_div), bias32[i]);
    float *dst = args.out_f32 + size_t(row) * size_t(args.out_ld) + col;
    if (col + 32 <= N && (args.out_ld & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<float4 *>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col + i < N) dst[i] = v[i];
    }
  } else {
    uint32_t packed[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float x = __fadd_rn(dequant(s[4 * i + b], args.coeff, args.rcp, args.fast_div), bias32[4 * i + b]);
        w |= uint32_t(lut[qsig_index(x)]) << (8 * b);
      }
      packed[i] = w;
    }
    uint8_t *dst = args.out_u8 + size_t(row) * size_t(N) + col;
    if (col + 32 <= N) {  // hidden widths are multiples of 16, so rows stay 16-byte aligned
      reinterpret_cast<uint4 *>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      reinterpret_cast<uint4 *>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col + i < N) dst[i] = uint8_t(packed[i >> 2] >> (8 * (i & 3)));
    }
    if (args.next_fix.ptr != nullptr) {
      reinterpret_cast<uint4 *>(scan)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      reinterpret_cast<uint4 *>(scan)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      post_saturation(scan, col >> 5, row, args.next_fix, args.next);
    }
  }
}

}  // namespace fdnn
