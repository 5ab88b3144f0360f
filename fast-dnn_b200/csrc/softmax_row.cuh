// Row softmax of the output layer, as the reference computes it (paths under /root/reference):
//   SoftMax::apply           src/cpp/dnn.cc:534-544   e = exp(x); total = Σ e; e / total — no max subtraction
//   LazyOutputActivations    src/cpp/dnn.cc:355-392   masked-out nodes enter as logit 0 (e = 1) and
//                                                     come back as 1/total, not 0
// The reference adds the exponentials sequentially in fp32, uses glibc's expf and divides; here the
// sum is a fixed-shape tree (deterministic), the exponential is 2^(x·log2e) on the SFU with the
// rounding error of the product folded back in (≈ 3e-7 relative), and the division is a multiply by
// the row's IEEE reciprocal (≤ 1 ulp) — which is where the stated float tolerance of the softmax
// scores comes from (tests/conftest.py: 1e-9 + 2e-5·|ref|; the logits themselves are bit-exact).
// The row routine below is shared by softmax_kernel (softmax.cu) and the softmax phase of the fused layer kernel
// (qlayer_fused.cu), so both produce the same bits.  It is instruction-bound otherwise (ncu: 35 instructions per element with expf and
// __fdiv_rn).  One CTA per row; the exponentials stay in shared memory between the two passes.

#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace fdnn {

constexpr int kSoftmaxThreads = 224;  // threads that work on one row (the shape of the summation tree depends on it): 7 warps,
                                      // so that the 28 warps of the fused layer kernel are exactly four row groups
constexpr int kSoftmaxUnroll = 4;  // (9 — the whole 8000-wide row in flight — measured: 12.4 vs 12.5 us at 512 rows, 240 vs 211 us at 16384)

// e^x.  t = RN(x·log2e_hi) goes to ex2.approx; r = (x·log2e − t) is recovered exactly with one fma
// plus the low part of log2e, and 2^r ≈ 1 + r·ln2 (|r| < 2^-17).  Overflows to +inf above 88.72 like
// the reference's expf does (there is no max subtraction, dnn.cc:534-544), underflows to 0, NaN stays NaN.
__device__ __forceinline__ float exp_fast(float x) {
  const float t = __fmul_rn(x, 1.4426950216293335f);
  const float r = fmaf(x, 1.9259629911e-8f, fmaf(x, 1.4426950216293335f, -t));
  float p;
  asm("ex2.approx.f32 %0, %1;" : "=f"(p) : "f"(t));
  // 0 and +inf are final (x = ±inf makes r = inf − inf = NaN, and inf · r is NaN for r ≤ 0): expf(−inf) = 0 keeps a class
  // that a −inf bias switched off at exactly 0, expf(x > 88.72) = +inf gives the reference's zeros-and-one-NaN row
  if (p == 0.0f || p == __int_as_float(0x7f800000)) return p;
  return fmaf(p, r * 0.6931471805599453f, p);
}

// One row by kSoftmaxThreads threads (tid = 0 … 223; `sync` is a barrier over exactly those threads: __syncthreads in the
// stand-alone kernel, a named barrier in the fused one).  s_e: O floats of shared memory when kCache (the exponentials
// between the two passes; x and y may then alias), s_red: kSoftmaxThreads / 32 floats.
template <bool kCache, class Sync>
__device__ __forceinline__ void softmax_row(const float *x, const int8_t *m, float *y, int O, bool vec, float *s_e, float *s_red, int tid,
                                            Sync sync) {
  constexpr int kThreads = kSoftmaxThreads, kUnroll = kSoftmaxUnroll;
  float part = 0.0f;
  if (vec) {
    // kUnroll loads in flight per thread before anything waits for them (a row is 8 float4 per thread on the 8000-wide
    // output layer: the pass was bound by their latency, one after the other, on short batches)
    for (int i0 = tid; i0 < O / 4; i0 += kThreads * kUnroll) {
      float4 v[kUnroll];
      char4 k[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = i0 + u * kThreads;
        v[u] = i < O / 4 ? reinterpret_cast<const float4 *>(x)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        k[u] = (m && i < O / 4) ? reinterpret_cast<const char4 *>(m)[i] : make_char4(1, 1, 1, 1);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = i0 + u * kThreads;
        if (i < O / 4) {
          const float4 e = make_float4(exp_fast(k[u].x ? v[u].x : 0.0f), exp_fast(k[u].y ? v[u].y : 0.0f), exp_fast(k[u].z ? v[u].z : 0.0f),
                                       exp_fast(k[u].w ? v[u].w : 0.0f));
          if (kCache) reinterpret_cast<float4 *>(s_e)[i] = e;
          part = __fadd_rn(part, __fadd_rn(__fadd_rn(e.x, e.y), __fadd_rn(e.z, e.w)));
        }
      }
    }
  } else {
    for (int i = tid; i < O; i += kThreads) {
      float v = x[i];
      if (m && m[i] == 0) v = 0.0f;
      const float e = exp_fast(v);
      if (kCache) s_e[i] = e;
      part = __fadd_rn(part, e);
    }
  }
  // fixed-shape tree: lanes, then the 8 warp sums (the barrier also orders the s_e writes before the reads below)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
  const int warp = tid / 32, lane = tid % 32;
  if (lane == 0) s_red[warp] = part;
  sync();
  float t = lane < kThreads / 32 ? s_red[lane] : 0.0f;  // 7 warp sums, padded with +0 to 8
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
  const float total = __shfl_sync(0xffffffffu, t, 0);
  const float inv = __fdiv_rn(1.0f, total);
  if (vec) {
    for (int i = tid; i < O / 4; i += kThreads) {
      float4 e;
      if (kCache) {
        e = reinterpret_cast<const float4 *>(s_e)[i];
      } else {
        float4 v = reinterpret_cast<const float4 *>(x)[i];
        if (m) {
          const char4 k = reinterpret_cast<const char4 *>(m)[i];
          v.x = k.x ? v.x : 0.0f;
          v.y = k.y ? v.y : 0.0f;
          v.z = k.z ? v.z : 0.0f;
          v.w = k.w ? v.w : 0.0f;
        }
        e = make_float4(exp_fast(v.x), exp_fast(v.y), exp_fast(v.z), exp_fast(v.w));
      }
      reinterpret_cast<float4 *>(y)[i] = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
    }
  } else {
    for (int i = tid; i < O; i += kThreads) {
      float e;
      if (kCache) {
        e = s_e[i];
      } else {
        float v = x[i];
        if (m && m[i] == 0) v = 0.0f;
        e = exp_fast(v);
      }
      y[i] = e * inv;
    }
  }
}

}  // namespace fdnn
