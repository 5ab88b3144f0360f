// Int8 layer on the CUDA cores (dp4a, u8 × s8 → s32) for layers the tensor-core path does not
// take: inputs narrower than one 128-byte K block or not a multiple of it, and hidden widths that
// are not a multiple of 32 (the reference only asks for multiples of 16, dnn.cc:331).  Same
// arithmetic contract and the same epilogue as qlayer_tc.cu; see that file for the references.
//
// One thread owns one frame and one aligned chunk of 16 nodes, so the per-element tail is shared
// verbatim; saturation corrections are recomputed per thread from the layer's risk list.

#include <cuda_runtime.h>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kRowsPerBlock = 128;

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <bool kLogits>
__global__ void __launch_bounds__(kRowsPerBlock) qlayer_simt_kernel(const QLayerArgs args) {
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  const int row = int(blockIdx.x) * kRowsPerBlock + int(threadIdx.x);
  const int col = int(blockIdx.y) * 16;
  if (row >= args.M) return;
  const int K = args.K, N = args.N;
  const int cols = min(16, N - col);
  int32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = 0;
  const uint4 *a_row = reinterpret_cast<const uint4 *>(args.act + size_t(row) * size_t(K));
  for (int k = 0; k < K / 16; ++k) {
    const uint4 a = a_row[k];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < cols) {  // warp-uniform
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(args.w + size_t(col + i) * size_t(K)) + k);
        int acc = s[i];
        acc = dp4a_u8s8(a.x, w.x, acc);
        acc = dp4a_u8s8(a.y, w.y, acc);
        acc = dp4a_u8s8(a.z, w.z, acc);
        acc = dp4a_u8s8(a.w, w.w, acc);
        s[i] = acc;
      }
    }
  }
  brute_force_corrections(s, row, col, args);  // dp4a does not saturate pair sums either
  float bias16[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) bias16[i] = i < cols ? __ldg(args.bias + col + i) : 0.0f;
  finish_chunk<kLogits>(s, row, col, args, bias16, args.lut);
}

}  // namespace

cudaError_t launch_qlayer_simt(const QLayerArgs &a, bool logits, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  dim3 grid((a.M + kRowsPerBlock - 1) / kRowsPerBlock, (a.N + 15) / 16);
  if (logits) return launch_pdl(qlayer_simt_kernel<true>, grid, dim3(kRowsPerBlock), size_t(0), stream, pdl_enabled(), a);
  return launch_pdl(qlayer_simt_kernel<false>, grid, dim3(kRowsPerBlock), size_t(0), stream, pdl_enabled(), a);
}

}  // namespace fdnn
