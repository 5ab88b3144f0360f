// fp32 input layer, bit-exact with the reference's SSE code (paths under /root/reference):
//   ApplyShiftAndScale   src/cpp/dnn.cc:175-192   x = (x + shift) · scale, two roundings
//   InputActivations     src/cpp/dnn.cc:219-247   four lane sums over k ≡ lane (mod 4): mulps, addps
//   horizontalSum        src/cpp/dnn.cc:168-172   (l0 + l1) + (l2 + l3)
//   AddBias              src/cpp/dnn.cc:250-264
//   QuantizedSigmoid     src/cpp/dnn.cc:267-286, src/cpp/dnn.h:35-42   LUT → u8
//
// The summation order is part of the result, so this is CUDA-core work: every product and every
// add is a separately rounded fp32 operation, never contracted to FMA, and each thread keeps the
// four SSE lanes of every (frame, node) it owns as four accumulators (scalar FMUL + FADD: measured
// on B200, the packed add.rn.f32x2 is slower than two scalar FADDs, and ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into a fused FFMA2 even with --fmad=false, which would change
// results — tools/microbench.cu, DESIGN.md §Measured hardware facts).
//
// Tiling: a CTA (16 warps) computes 64 frames × 128 nodes; a warp 16 frames × 32 nodes; a thread
// 4 frames × 4 nodes, strided (frames fg + 4i, nodes ng + 8j with lane = 8·fg + ng) so that every
// 128-bit shared-memory read is either a broadcast or conflict-free with the 44-float row pitch —
// 8 wavefronts per 128 math instructions.  K is streamed in 40-float chunks, double
// buffered: weights by cp.async, frames through registers so shift/scale is applied on the way in.

#include <cuda_runtime.h>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kTileF = 64;    // frames per CTA
constexpr int kTileN = 128;   // nodes per CTA
constexpr int kChunk = 40;    // floats of K per stage
constexpr int kPitch = 44;    // smem row pitch in floats (≡ 12 mod 32 → conflict-free LDS.128 over consecutive rows)
constexpr int kThreads = 512;
constexpr int kTF = 4, kTN = 4;
constexpr int kStageFloats = (kTileF + kTileN) * kPitch;
constexpr int kSmemBytes = 2 * kStageFloats * 4 + kLut2Padded;
static_assert(kTileF * kTileN <= 2 * kStageFloats * 4, "u8 output tile must fit in the pipeline buffers");

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1) input_layer_kernel(const InputLayerArgs args) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float *stage_buf = reinterpret_cast<float *>(smem_raw);
  uint8_t *s_lut = smem_raw + 2 * kStageFloats * 4;

  const int tid = int(threadIdx.x);
  const int warp = tid / 32, lane = tid % 32;
  const int fg = lane / 8, ng = lane % 8;
  const int wf = (warp / 4) * 16, wn = (warp % 4) * 32;  // warp tile origin inside the CTA tile
  const int f0 = int(blockIdx.y) * kTileF, n0 = int(blockIdx.x) * kTileN;
  const int M = args.M, I = args.I, H = args.H;
  const int n_chunks = (I + kChunk - 1) / kChunk;

  for (int i = tid; i < kLut2Padded / 16; i += kThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(args.lut) + i);
  ptx::griddep_wait();  // the activation buffer we write may still be read by the previous pass
  ptx::griddep_launch_dependents();

  // Frame elements this thread moves per chunk: kTileF rows × (kChunk/4) float4 = 640 vectors.
  constexpr int kXVecs = kTileF * (kChunk / 4);
  constexpr int kXPerThread = (kXVecs + kThreads - 1) / kThreads;
  float4 xr[kXPerThread];

  auto load_x_regs = [&](int c) {
    const int k0 = c * kChunk, kc = min(kChunk, I - k0);
#pragma unroll
    for (int j = 0; j < kXPerThread; ++j) {
      const int v = tid + j * kThreads;
      const int r = v / (kChunk / 4), q = v % (kChunk / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < kXVecs && 4 * q < kc && f0 + r < M) {
        const float4 raw = __ldg(reinterpret_cast<const float4 *>(args.in + size_t(f0 + r) * size_t(I) + k0) + q);
        const float4 sh = __ldg(reinterpret_cast<const float4 *>(args.shift + k0) + q);
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(args.scale + k0) + q);
        val.x = __fmul_rn(__fadd_rn(raw.x, sh.x), sc.x);
        val.y = __fmul_rn(__fadd_rn(raw.y, sh.y), sc.y);
        val.z = __fmul_rn(__fadd_rn(raw.z, sh.z), sc.z);
        val.w = __fmul_rn(__fadd_rn(raw.w, sh.w), sc.w);
      }
      xr[j] = val;
    }
  };
  auto store_x_regs = [&](int buf) {
    float *xs = stage_buf + buf * kStageFloats;
#pragma unroll
    for (int j = 0; j < kXPerThread; ++j) {
      const int v = tid + j * kThreads;
      if (v < kXVecs) {
        const int r = v / (kChunk / 4), q = v % (kChunk / 4);
        *reinterpret_cast<float4 *>(xs + r * kPitch + 4 * q) = xr[j];
      }
    }
  };
  auto issue_w = [&](int c, int buf) {
    const int k0 = c * kChunk, kc = min(kChunk, I - k0);
    float *ws = stage_buf + buf * kStageFloats + kTileF * kPitch;
    for (int v = tid; v < kTileN * (kChunk / 4); v += kThreads) {
      const int r = v / (kChunk / 4), q = v % (kChunk / 4);
      float *dst = ws + r * kPitch + 4 * q;
      if (4 * q < kc && n0 + r < H)
        cp_async16(dst, args.w0 + size_t(n0 + r) * size_t(I) + k0 + 4 * q);
      else
        *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  // acc[i][j] = the four SSE lanes of frame fg+4i, node ng+8j
  float4 acc[kTF][kTN];
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTN; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);

  load_x_regs(0);
  issue_w(0, 0);
  store_x_regs(0);
  cp_async_wait_all();
  __syncthreads();

  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    const bool more = c + 1 < n_chunks;
    if (more) {
      issue_w(c + 1, buf ^ 1);
      load_x_regs(c + 1);
    }
    const float *xs = stage_buf + buf * kStageFloats + (wf + fg) * kPitch;
    const float *ws = stage_buf + buf * kStageFloats + (kTileF + wn + ng) * kPitch;
    const int kc4 = (min(kChunk, I - c * kChunk)) / 4;
#pragma unroll 2
    for (int q = 0; q < kc4; ++q) {
      float4 xv[kTF], wv[kTN];
#pragma unroll
      for (int i = 0; i < kTF; ++i) xv[i] = *reinterpret_cast<const float4 *>(xs + (4 * i) * kPitch + 4 * q);
#pragma unroll
      for (int j = 0; j < kTN; ++j) wv[j] = *reinterpret_cast<const float4 *>(ws + (8 * j) * kPitch + 4 * q);
#pragma unroll
      for (int i = 0; i < kTF; ++i)
#pragma unroll
        for (int j = 0; j < kTN; ++j) {
          acc[i][j].x = __fadd_rn(acc[i][j].x, __fmul_rn(xv[i].x, wv[j].x));
          acc[i][j].y = __fadd_rn(acc[i][j].y, __fmul_rn(xv[i].y, wv[j].y));
          acc[i][j].z = __fadd_rn(acc[i][j].z, __fmul_rn(xv[i].z, wv[j].z));
          acc[i][j].w = __fadd_rn(acc[i][j].w, __fmul_rn(xv[i].w, wv[j].w));
        }
    }
    if (more) store_x_regs(buf ^ 1);
    cp_async_wait_all();
    __syncthreads();
  }

  // ---- epilogue: lanes → h, + bias, LUT; stage the u8 tile in shared memory ------------------------
  uint8_t *s_out = smem_raw;  // [kTileF][kTileN], pipeline buffers are free after the last barrier
  float bias[kTN];
#pragma unroll
  for (int j = 0; j < kTN; ++j) bias[j] = (n0 + wn + ng + 8 * j < H) ? __ldg(args.bias0 + n0 + wn + ng + 8 * j) : 0.0f;
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTN; ++j) {
      const float h = __fadd_rn(__fadd_rn(acc[i][j].x, acc[i][j].y), __fadd_rn(acc[i][j].z, acc[i][j].w));
      s_out[(wf + fg + 4 * i) * kTileN + wn + ng + 8 * j] = s_lut[qsig_slot(__fadd_rn(h, bias[j]))];
    }
  __syncthreads();

  const int cols = min(kTileN, H - n0);  // multiple of 16
  for (int v = tid; v < kTileF * (kTileN / 16); v += kThreads) {
    const int r = v / (kTileN / 16), q = v % (kTileN / 16);
    if (f0 + r < M && 16 * q < cols) {
      *reinterpret_cast<uint4 *>(args.out_u8 + size_t(f0 + r) * size_t(H) + n0 + 16 * q) = *reinterpret_cast<const uint4 *>(s_out + r * kTileN + 16 * q);
    }
  }
}

}  // namespace

cudaError_t input_layer_configure() {
  return cudaFuncSetAttribute(input_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_input_layer(const InputLayerArgs &a, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  dim3 grid((a.H + kTileN - 1) / kTileN, (a.M + kTileF - 1) / kTileF);
  return launch_pdl(input_layer_kernel, grid, dim3(kThreads), size_t(kSmemBytes), stream, pdl_enabled(), a);
}

}  // namespace fdnn
