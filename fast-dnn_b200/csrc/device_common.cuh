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
  // 1: fast_div holds, bias and coeff are finite and |lin + bias|·200 < 2³¹ for every reachable sum (checked
  // when the model is uploaded) — the packed-f32x2 tail below is then bit-identical to the general one
  int fast_tail;
  float one, neg_zero;  // 1.0f and −0.0f as run-time values: keeps ptxas from rewriting fma(a, 1, b) / fma(a, b, −0) into FADD2 / FMUL2
  int M, N, K;
  int row0;         // first row of this launch (CTA-pair kernel only; a multiple of 256): rows [row0, M) are computed
  FixList fix;      // this layer's saturation risk entries
  uint8_t *out_u8;  // hidden mode: u8 activations [M][N]
  float *out_f32;   // logits mode: lin + bias, fp32 [M][out_ld]
  int out_ld;
  // optional per-CTA phase timestamps (SM clock), 8 slots per CTA; nullptr in normal operation
  unsigned long long *timeline;
  int debug_flags;  // profiling experiments only (FDNN_DEBUG; results wrong): 1 = scan warps skip their entries
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

// ---- packed fp32 pairs (Blackwell FFMA2: one issue slot for two IEEE fused multiply-adds) --------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// fp32 logits of one row per lane, 16 consecutive columns each, written so that the four lanes of a quad cover 64 contiguous
// bytes of ONE row per store instruction (4×4 transpose of the 16-byte pieces over the quad, two butterfly steps): a warp-wide
// store is then 8 rows × two full 32-byte sectors instead of 32 rows × half a sector each.  (Measured on B200 with the
// kernel's own clock stamps: the logits epilogue of a 128×256 tile took 15 k of the tile's 40 k cycles at batch 512, at
// ≈ 1.8 cycles per half-sector store request.)  All 32 lanes must call it; the rows of a quad are consecutive.
__device__ __forceinline__ void store_logits_quad(float4 (&p)[4], float *out, size_t ld, int row, int col, int M) {
  const int qi = int(threadIdx.x) & 3;
  {
    const bool hi = (qi & 1) != 0;  // swap bit 0 of (lane in quad, piece)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float4 send = hi ? p[2 * m] : p[2 * m + 1];
      float4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
      recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
      if (hi) p[2 * m] = recv; else p[2 * m + 1] = recv;
    }
  }
  {
    const bool hi = (qi & 2) != 0;  // … and bit 1
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float4 send = hi ? p[m] : p[m + 2];
      float4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 2);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 2);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 2);
      recv.w = __shfl_xor_sync(0xffffffffu, send.w, 2);
      if (hi) p[m] = recv; else p[m + 2] = recv;
    }
  }
  // p[k] is now piece qi (columns col + 4·qi …) of row (row − qi + k)
  const int row0 = row - qi;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (row0 + k < M) *reinterpret_cast<float4 *>(out + size_t(row0 + k) * ld + size_t(col + 4 * qi)) = p[k];
}

// Same results as finish_chunk below when args.fast_tail is set, in ≈ 8 instead of ≈ 15 issue slots per
// element: every rounding step of the reference is kept, two elements per FFMA2 —
//   q = RN(s·rcp) = fma(s, rcp, −0);  e = fma(q, −coeff, s);  lin = fma(e, rcp, q)      (the verified 3-op division)
//   x = RN(lin + bias) = fma(lin, 1, bias)
//   hidden: RN(x·200) = 2·RN(x·100) exactly, so trunc(clamp(RN(x·200), ±1282)) is the doubled-LUT slot of
//   qsig_slot(); fast_tail rules out NaN and |x·200| ≥ 2³¹, the only inputs on which the two differ.
// kQuadStore (logits only): every lane of the warp calls, lane = row, `row_ok` says whether the row exists (row < M)
template <bool kLogits, bool kQuadStore = false>
__device__ __forceinline__ void finish_chunk_fast(const int32_t (&s)[16], int row, int col, const QLayerArgs &args, const float *bias16,
                                                  const uint8_t *lut, bool row_ok = true, int M = 0) {
  const uint64_t rcp2 = pack2(args.rcp, args.rcp), ncoeff2 = pack2(-args.coeff, -args.coeff), one2 = pack2(args.one, args.one),
                 nz2 = pack2(args.neg_zero, args.neg_zero);
  uint64_t x[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const uint64_t sf = pack2(__int2float_rn(s[2 * p]), __int2float_rn(s[2 * p + 1]));
    const uint64_t q = fma2(sf, rcp2, nz2);
    const uint64_t e = fma2(q, ncoeff2, sf);
    const uint64_t lin = fma2(e, rcp2, q);
    x[p] = fma2(lin, one2, *reinterpret_cast<const uint64_t *>(bias16 + 2 * p));
  }
  if constexpr (kLogits) {
    const int N = args.N;
    float *dst = args.out_f32 + size_t(row) * size_t(args.out_ld) + col;
    if (col + 16 <= N && (args.out_ld & 3) == 0) {
      float4 v4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b, c, d;
        unpack2(x[2 * i], a, b);
        unpack2(x[2 * i + 1], c, d);
        v4[i] = make_float4(a, b, c, d);
      }
      if constexpr (kQuadStore) {
        store_logits_quad(v4, args.out_f32, size_t(args.out_ld), row, col, M);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4 *>(dst)[i] = v4[i];
      }
    } else if (row_ok) {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        float a, b;
        unpack2(x[p], a, b);
        if (col + 2 * p < N) dst[2 * p] = a;
        if (col + 2 * p + 1 < N) dst[2 * p + 1] = b;
      }
    }
  } else {
    const uint64_t k200 = pack2(__fmul_rn(args.one, 200.0f), __fmul_rn(args.one, 200.0f));
    uint32_t packed[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint32_t b[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float t0, t1;
        unpack2(fma2(x[2 * w + h], k200, nz2), t0, t1);
        const int v0 = __float2int_rz(fminf(fmaxf(t0, -1282.0f), 1282.0f));
        const int v1 = __float2int_rz(fminf(fmaxf(t1, -1282.0f), 1282.0f));
        b[2 * h] = lut[v0 + kLut2Center];
        b[2 * h + 1] = lut[v1 + kLut2Center];
      }
      packed[w] = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
    }
    *reinterpret_cast<uint4 *>(args.out_u8 + size_t(row) * size_t(args.N) + col) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// The reference's per-element tail for one row and one aligned chunk of 16 nodes whose exact
// (saturation-corrected) sums are in `s`: dequantise → + bias → {LUT → u8 | fp32 logits} → global.
template <bool kLogits, bool kQuadStore = false>
__device__ __forceinline__ void finish_chunk(const int32_t (&s)[16], int row, int col, const QLayerArgs &args, const float *bias16,
                                             const uint8_t *lut, bool row_ok = true, int M = 0) {
  const int N = args.N;
  if constexpr (kLogits) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(dequant(s[i], args.coeff, args.rcp, args.fast_div), bias16[i]);
    float *dst = args.out_f32 + size_t(row) * size_t(args.out_ld) + col;
    if (col + 16 <= N && (args.out_ld & 3) == 0) {
      float4 v4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if constexpr (kQuadStore) {
        store_logits_quad(v4, args.out_f32, size_t(args.out_ld), row, col, M);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4 *>(dst)[i] = v4[i];
      }
    } else if (row_ok) {
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
