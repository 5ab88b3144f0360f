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
// Tiling: a CTA (16 warps) computes 64 frames × 128 nodes; a warp 16 frames × 32 nodes.  Shared
// memory bandwidth, not arithmetic, is what a naive register tile runs out of here (a 128-bit load
// costs four wavefronts however much of it is broadcast), so the four SSE lanes are split across
// four adjacent threads: thread (group g, lane r) owns SSE lane r of an 8 × 8 block — 64
// accumulators, and per 4-wide K step 16 one-wavefront 32-bit loads for 128 math instructions.
// The lanes are recombined as (l0 + l1) + (l2 + l3) with two shuffles at the end.  K is streamed
// in 88-float chunks (5 per 440-wide input: fewer barriers than with 40-float ones, 58.7 → 56.9 µs per 512 frames)
// through a 3-stage cp.async ring (frames and weight rows alike, issued two
// chunks ahead); the frame rows of a stage are shifted and scaled in place one iteration before
// they are used, so the loop has one barrier per chunk and no exposed global-memory latency.

#include <cuda_runtime.h>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kTileF = 64;    // frames per CTA
constexpr int kTileN = 128;   // nodes per CTA
constexpr int kChunk = 88;    // floats of K per stage
constexpr int kPitch = 92;    // smem row pitch in floats (≡ 28 mod 32: consecutive rows land in different bank quads)
constexpr int kThreads = 512;
constexpr int kTF = 8, kTN = 8;  // per-thread block of one SSE lane
constexpr int kStages = 3;
constexpr int kStageFloats = (kTileF + kTileN) * kPitch;
constexpr int kMaxI = 1024;   // shift/scale are kept in shared memory
constexpr int kSmemBytes = kStages * kStageFloats * 4 + kLut2Padded + 2 * kMaxI * 4;
static_assert(kTileF * kTileN <= kStages * kStageFloats * 4, "u8 output tile must fit in the pipeline buffers");

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) input_layer_kernel(const InputLayerArgs args) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float *stage_buf = reinterpret_cast<float *>(smem_raw);
  uint8_t *s_lut = smem_raw + kStages * kStageFloats * 4;
  float *s_shift = reinterpret_cast<float *>(s_lut + kLut2Padded);
  float *s_scale = s_shift + kMaxI;

  const int tid = int(threadIdx.x);
  const int warp = tid / 32, lane = tid % 32;
  const int r = lane & 3;                // SSE lane this thread accumulates
  const int fg = (lane >> 2) & 1;        // frames fg + 2i   (i < 8) of the warp tile
  const int ng = lane >> 3;              // nodes  ng + 4j   (j < 8)
  const int wf = (warp / 4) * 16, wn = (warp % 4) * 32;  // warp tile origin inside the CTA tile
  const int f0 = int(blockIdx.y) * kTileF, n0 = int(blockIdx.x) * kTileN;
  const int M = args.M, I = args.I, H = args.H;
  const int n_chunks = (I + kChunk - 1) / kChunk;

  for (int i = tid; i < kLut2Padded / 16; i += kThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(args.lut) + i);
  for (int i = tid; i < I; i += kThreads) {
    s_shift[i] = __ldg(args.shift + i);
    s_scale[i] = __ldg(args.scale + i);
  }
  ptx::griddep_wait();  // the activation buffer we write may still be read by the previous pass
  ptx::griddep_launch_dependents();

  // Stage c of the ring holds, for K chunk c: rows [0, 64) = frames (raw on arrival, shifted and
  // scaled in place one iteration before use), rows [64, 192) = weight rows.  Both arrive by cp.async.
  // The (row, 16-byte column) slots a thread moves are the same for every chunk: resolve them once.
  constexpr int kVecsPerRow = kChunk / 4;
  constexpr int kSlots = ((kTileF + kTileN) * kVecsPerRow + kThreads - 1) / kThreads;  // 4
  constexpr int kXSlots = (kTileF * kVecsPerRow + kThreads - 1) / kThreads;             // 2
  const float *slot_src[kSlots];  // global address of the slot in chunk 0, nullptr if the row does not exist
  int slot_dst[kSlots];           // float offset inside a stage
  int slot_col[kSlots];           // first K index of the slot inside the chunk
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int v = tid + j * kThreads;
    const int row = v / kVecsPerRow, q = v % kVecsPerRow;
    const bool in_range = v < (kTileF + kTileN) * kVecsPerRow;
    const bool is_x = row < kTileF;
    const int src_row = is_x ? f0 + row : n0 + row - kTileF;
    const bool ok = in_range && (is_x ? src_row < M : src_row < H);
    slot_src[j] = ok ? (is_x ? args.in : args.w0) + size_t(src_row) * size_t(I) + 4 * q : nullptr;
    slot_dst[j] = in_range ? row * kPitch + 4 * q : -1;
    slot_col[j] = 4 * q;
  }
  auto issue = [&](int c) {
    if (c < n_chunks) {
      const int k0 = c * kChunk, kc = min(kChunk, I - k0);
      float *st = stage_buf + (c % kStages) * kStageFloats;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        if (slot_dst[j] >= 0) {
          if (slot_src[j] != nullptr && slot_col[j] < kc)
            cp_async16(st + slot_dst[j], slot_src[j] + k0);
          else
            *reinterpret_cast<float4 *>(st + slot_dst[j]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    cp_async_commit();  // (possibly empty) group: keeps the wait_group arithmetic uniform
  };
  // ApplyShiftAndScale (dnn.cc:175-192) on the frame rows of chunk c, in place: add, then multiply.
  // Frame rows are the first kTileF·kVecsPerRow slots, so a thread transforms slots it loaded itself
  // or a neighbour did — either way after the barrier that made them visible.
  auto transform = [&](int c) {
    if (c >= n_chunks) return;
    const int k0 = c * kChunk, kc = min(kChunk, I - k0);
    float *st = stage_buf + (c % kStages) * kStageFloats;
#pragma unroll
    for (int j = 0; j < kXSlots; ++j) {
      if (tid + j * kThreads < kTileF * kVecsPerRow && slot_col[j] < kc) {
        float4 x = *reinterpret_cast<float4 *>(st + slot_dst[j]);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + k0 + slot_col[j]);
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + k0 + slot_col[j]);
        x.x = __fmul_rn(__fadd_rn(x.x, sh.x), sc.x);
        x.y = __fmul_rn(__fadd_rn(x.y, sh.y), sc.y);
        x.z = __fmul_rn(__fadd_rn(x.z, sh.z), sc.z);
        x.w = __fmul_rn(__fadd_rn(x.w, sh.w), sc.w);
        *reinterpret_cast<float4 *>(st + slot_dst[j]) = x;
      }
    }
  };

  // acc[i][j] = SSE lane r of frame fg+2i, node ng+4j
  float acc[kTF][kTN];
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTN; ++j) acc[i][j] = 0.f;

  issue(0);
  issue(1);
  cp_async_wait<1>();  // chunk 0 landed (this thread's part)
  __syncthreads();     // … everyone's part, and s_shift/s_scale
  transform(0);

  for (int c = 0; c < n_chunks; ++c) {
    // chunk c+1 has landed, chunk c has been transformed by the previous iteration, chunk c−1 is done with
    cp_async_wait<0>();
    __syncthreads();
    issue(c + 2);      // into the stage chunk c−1 occupied
    transform(c + 1);  // nobody reads it before the next barrier
    const float *xs = stage_buf + (c % kStages) * kStageFloats + (wf + fg) * kPitch + r;
    const float *ws = stage_buf + (c % kStages) * kStageFloats + (kTileF + wn + ng) * kPitch + r;
    const int kc4 = (min(kChunk, I - c * kChunk)) / 4;
#pragma unroll 1
    for (int q = 0; q < kc4; ++q) {
      float xv[kTF], wv[kTN];
#pragma unroll
      for (int i = 0; i < kTF; ++i) xv[i] = xs[(2 * i) * kPitch + 4 * q];
#pragma unroll
      for (int j = 0; j < kTN; ++j) wv[j] = ws[(4 * j) * kPitch + 4 * q];
#pragma unroll
      for (int i = 0; i < kTF; ++i)
#pragma unroll
        for (int j = 0; j < kTN; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(xv[i], wv[j]));
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: lanes → h, + bias, LUT; stage the u8 tile in shared memory ------------------------
  uint8_t *s_out = smem_raw;  // [kTileF][kTileN], pipeline buffers are free after the last barrier
  // horizontalSum (dnn.cc:168-172): (l0 + l1) + (l2 + l3); fp32 addition is commutative, so every
  // lane of the quad ends up with the same bits.  Lane r then finishes the outputs with j ≡ r (mod 4).
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTN; ++j) {
      const float pair = __fadd_rn(acc[i][j], __shfl_xor_sync(0xffffffffu, acc[i][j], 1));
      acc[i][j] = __fadd_rn(pair, __shfl_xor_sync(0xffffffffu, pair, 2));
    }
#pragma unroll
  for (int j = 0; j < kTN; ++j) {
    if ((j & 3) == r) {
      const int node = wn + ng + 4 * j;
      const float bias = (n0 + node < H) ? __ldg(args.bias0 + n0 + node) : 0.0f;
#pragma unroll
      for (int i = 0; i < kTF; ++i) s_out[(wf + fg + 2 * i) * kTileN + node] = s_lut[qsig_slot(__fadd_rn(acc[i][j], bias))];
    }
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

int input_layer_max_dim() { return kMaxI; }

cudaError_t launch_input_layer(const InputLayerArgs &a, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  if (a.I > kMaxI || a.I % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid((a.H + kTileN - 1) / kTileN, (a.M + kTileF - 1) / kTileF);
  return launch_pdl(input_layer_kernel, grid, dim3(kThreads), size_t(kSmemBytes), stream, pdl_enabled(), a);
}

}  // namespace fdnn
