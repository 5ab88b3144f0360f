// fp32 input layer, certified on the tensor cores and finished exactly where the certificate fails.
//
// The reference's layer 0 (src/cpp/dnn.cc:175-192 ApplyShiftAndScale, :219-247 InputActivations, :168-172
// horizontalSum, :250-286 AddBias + QuantizedSigmoid) produces ONE BYTE per (frame, node): the sigmoid
// bucket k = round-half-away((h + bias)·100) of a 440-term fp32 dot product h whose summation order is
// part of the result.  Reproducing every rounding costs two CUDA-core instructions per MAC
// (input_layer.cu: half of the whole pass).  But the byte only depends on which side of a bucket boundary
// the reference's h falls, and the reference's h is within a RIGOROUS distance of the exact real dot
// product.  So:
//
//   1. input_prep_kernel    x' = fl(fl(x + shift)·scale) exactly as the reference; each row is written as a
//                           block-fixed-point integer vector X (|X_k| ≤ 2²², three 8-bit limbs) together with
//                           its scale, Σ|X_k| and ‖x'‖₂.  W0 gets the same treatment once, at model upload.
//   2. input_tc_kernel      Σ_k X_k·W_k EXACTLY up to the product of the two lowest limbs (bounded by 255·Σ|W₀| and
//                           put into ε_q): the eight limb-pair products on tcgen05.mma kind::i8, s32 accumulators in tensor
//                           memory, one per shift class s = i + j = 1 … 4, two sets so that a tile's epilogue overlaps the
//                           next tile's MMAs.  The frame limbs are the bytes of the two's-complement integer (u8, u8, s8); the
//                           weight limbs are BALANCED base-256 digits (all s8, W = W₀ + 256·W₁ + 65536·W₂ exactly), so that one
//                           instruction multiplies a frame limb with all the weight limbs it needs (N = 192 or 128: three
//                           instructions per K step instead of eight of N = 64 that cost the same ≈ 37 ns each).
//                           z = (h_fix + bias)·100 is then evaluated in fp32 from the five class sums and
//                             D = 100·(u·‖√c·x'‖₂·‖√c·w‖₂ + ε_q) + 100·7u·‖x'‖₂'·‖w‖₂' + 3.1u·|z| + u·|bias·100| + 2.1u
//                           where u = 2⁻²⁴ and c_k counts the roundings term k goes through in the reference (its
//                           product, the adds of its SSE lane that follow it — I/4 − 1 for the first two terms of a
//                           lane down to 1 for the last — and the two combining adds), so that the reference's h is
//                           within Σ γ_{c_k}|x'_k·w_k| ≤ u(1+1e-4)·‖√c·x'‖₂‖√c·w‖₂ of the real dot product (weighted
//                           Cauchy-Schwarz), ε_q = s_x·s_w·(½Σ|X_k| + ½Σ|W_k| + ¼I) bounds the fixed-point quantisation,
//                           the 7u term the fp32 evaluation of z (five int→float conversions and five FMAs over
//                           terms whose absolute values sum to ≤ 100·‖x'‖₂'·‖w‖₂', the primes adding 2¹⁷√I units for
//                           the two's-complement limbs of negative numbers), and the rest the fp32 roundings of
//                           bias·100 here and of "+ bias", "· 100" in the reference.  Every constant is rounded up.
//                           If no half-integer lies within D of z, every value the reference can have produced
//                           rounds to the same k: the byte is certain.  Otherwise (≈ 2.6 % of the elements on the
//                           synthetic network) the element's bit is set in a [frame][node/32] bitmap and in its
//                           transpose by blocks of 32 frames, [frame/32][node].
//   3. input_fixup_block_kernel   the flagged elements, with the reference's exact arithmetic (as input_layer.cu: four
//                           lane sums, FMUL + FADD, never FMA): a CTA per 32 frames × up to 2048 nodes, frames in shared
//                           memory, the listed nodes' weight rows by bulk copy.  (input_fixup_kernel, a warp per frame
//                           reading weights through L1, is the earlier version: FDNN_FIXUP=warp.)
//
// Rows with non-finite or extreme values and nodes with non-finite weights or bias are left to step 3 entirely,
// so the result is bit-identical to input_layer.cu for any input.

#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kTileM = 128, kTileN = 64, kBlockK = 128, kUmmaK = 32;
constexpr int kLimbs = 3;
constexpr int kClasses = 2 * kLimbs - 2;  // shift classes i + j = 1 … 4; class 0 (low limb × low limb) is bounded, not computed
constexpr int kAccSets = 2;               // two sets of accumulators: the next tile's MMAs run under this tile's epilogue
constexpr int kStages = 3;
constexpr int kXBytes = kTileM * kBlockK, kWBytes = kTileN * kBlockK;
constexpr int kStageBytes = kLimbs * (kXBytes + kWBytes);  // 72 KB
constexpr int kEpiWarps = 16;  // four per tensor-memory lane quarter, 16 of the tile's 64 columns each (the certificate is ≈ 40 instructions
                               // per element and latency-bound: eight warps took 5.5 us per tile, more than the MMAs or the operand feed)
constexpr int kTcThreads = (4 + kEpiWarps) * 32;
constexpr int kTmemCols = kAccSets * kClasses * kTileN;  // 2 × 4 × 64 = 512
constexpr int kBarRegion = 128;  // 2·kStages + 1 barriers, the TMEM address
constexpr int kTcSmem = kStages * kStageBytes + kBarRegion + kLut2Padded + kTileN * int(sizeof(InputNodeStats)) + kTileM * 2 * 4;

// ---- 1. prepare: transform, block-fixed-point limbs, row statistics ---------------------------------------
__global__ void __launch_bounds__(256) input_prep_kernel(const InputTcArgs a) {
  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  const int row = int(blockIdx.x) * 8 + warp;
  ptx::griddep_wait();  // launched with programmatic serialization: the previous pass may still be reading what this one writes
  ptx::griddep_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.unc_count = 0u;
  if (row >= a.M) return;
  const int I = a.I, I4 = I / 4;
  constexpr int kVec = kInputTcMaxI / 128;  // float4 per lane
  float4 xv[kVec];
  float mx = 0.0f;
  bool bad = false;
  double n2 = 0.0, n2c = 0.0;
  const float4 *in4 = reinterpret_cast<const float4 *>(a.in + size_t(row) * size_t(I));
  float4 *xq4 = reinterpret_cast<float4 *>(a.xq + size_t(row) * size_t(I));
#pragma unroll
  for (int t = 0; t < kVec; ++t) {
    const int q = lane + 32 * t;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < I4) {
      const float4 x = in4[q], sh = reinterpret_cast<const float4 *>(a.shift)[q], sc = reinterpret_cast<const float4 *>(a.scale)[q];
      v.x = __fmul_rn(__fadd_rn(x.x, sh.x), sc.x);  // dnn.cc:175-192: add, then multiply, each rounded
      v.y = __fmul_rn(__fadd_rn(x.y, sh.y), sc.y);
      v.z = __fmul_rn(__fadd_rn(x.z, sh.z), sc.z);
      v.w = __fmul_rn(__fadd_rn(x.w, sh.w), sc.w);
      xq4[q] = v;
      const double c = input_round_count(4 * q, I);  // the four SSE lanes of a K step go through the same number of adds
      const double sq = double(v.x) * double(v.x) + double(v.y) * double(v.y) + double(v.z) * double(v.z) + double(v.w) * double(v.w);
      n2 += sq;
      n2c += c * sq;
    }
    xv[t] = v;
    const float m4 = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    bad |= !(m4 <= 3.0e38f);  // NaN or inf
    mx = fmaxf(mx, m4);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    n2c += __shfl_xor_sync(0xffffffffu, n2c, o);
    bad |= __shfl_xor_sync(0xffffffffu, int(bad), o) != 0;
  }
  int e = (mx > 0.0f) ? ilogbf(mx) + 1 : 0;  // mx < 2^e
  if (e < -40 || e > 40) bad = true;
  if (bad) e = 0;
  const float qs = ldexpf(1.0f, 22 - e);  // exact power of two
  long long sx = 0;
  uint32_t *p0 = reinterpret_cast<uint32_t *>(a.x_limbs + size_t(row) * kInputTcPitch);
  uint32_t *p1 = reinterpret_cast<uint32_t *>(a.x_limbs + a.x_plane + size_t(row) * kInputTcPitch);
  uint32_t *p2 = reinterpret_cast<uint32_t *>(a.x_limbs + 2 * a.x_plane + size_t(row) * kInputTcPitch);
#pragma unroll
  for (int t = 0; t < kVec; ++t) {
    const int q = lane + 32 * t;
    if (q < I4) {
      const float f[4] = {xv[t].x, xv[t].y, xv[t].z, xv[t].w};
      uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int X = bad ? 0 : __float2int_rn(f[j] * qs);  // |X| ≤ 2²², exact scaling
        sx += X < 0 ? -X : X;
        w0 |= uint32_t(X & 255) << (8 * j);
        w1 |= uint32_t((X >> 8) & 255) << (8 * j);
        w2 |= uint32_t((X >> 16) & 255) << (8 * j);  // signed top limb
      }
      p0[q] = w0;
      p1[q] = w1;
      p2[q] = w2;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
  if (lane == 0) {
    constexpr double u = 5.9604644775390625e-8, up = 1.0 + 1e-6;
    InputRowStats s;
    const double sc = ldexp(1.0, e - 22), nx = sqrt(n2) * (1.0 + 1e-9), nxc = sqrt(n2c) * (1.0 + 1e-9);
    for (int c = 0; c < 5; ++c) s.a[c] = float(100.0 * ldexp(sc, 8 * c));  // exact: 100 · 2^n
    if (bad) s.a[0] = -1.0f;
    s.p = __double2float_ru(100.0 * u * (1.0 + 1e-4) * nxc * up);  // γ_c ≤ c·u·(1 + 1e-4) for c ≤ 1000
    s.pp = __double2float_ru(100.0 * 7.0 * u * (nx + sc * 131072.0 * sqrt(double(I))) * up);
    s.r1 = __double2float_ru(100.0 * sc * 0.5 * double(sx) * up);
    s.ar = __double2float_ru(100.0 * sc * up);
    s.pad[0] = s.pad[1] = s.pad[2] = 0.0f;
    a.row_stats[row] = s;
  }
}

// ---- 2. exact fixed-point dot products on the tensor cores, certificate, byte or list ------------------------
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// instruction descriptor, kind::i8, D = s32, M = 128, N = n (a multiple of 16), A u8 (0) or s8 (1), B s8, both K-major
__device__ __forceinline__ uint32_t idesc_limbs(bool a_signed, int n) {
  return (2u << 4) | (uint32_t(a_signed) << 7) | (1u << 10) | ((uint32_t(n) >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kTcThreads, 1)
input_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const InputTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t *tiles = smem;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *acc_bar = empty_bar + kStages;   // [set] the tile's accumulators are complete
  uint64_t *acc_free = acc_bar + kAccSets;   // [set] the epilogue warps have read them
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + kAccSets);
  uint8_t *s_lut = smem + kStages * kStageBytes + kBarRegion;

  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  // persistent: CTA b takes tiles b, b + grid, …; the TMA ring runs on across tile boundaries, so the next tile's operands
  // arrive while this tile's epilogue drains the (single) set of accumulators
  const int n_blocks = (a.H + kTileN - 1) / kTileN, m_blocks = (a.M + kTileM - 1) / kTileM;
  const int tiles_total = n_blocks * m_blocks;
  const int k_blocks = (a.I + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < kAccSets; ++i) {
      ptx::mbar_init(acc_bar + i, 1);
      ptx::mbar_init(acc_free + i, kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::griddep_wait();  // barriers, tensor memory and descriptors were set up under the prepare kernel's tail
  ptx::griddep_launch_dependents();

  if (warp == 0) {
    // SIX lanes issue, one box each, in one instruction (lanes 0-2 the frame limbs, 3-5 the weight limbs): one thread gets a box
    // accepted only every ≈ 225 ns (tools/feed_bench.cu), and six boxes per 72 KB stage from one lane were what this kernel ran at
    // (52 GB/s per SM, 5.5 us per tile — more than its MMAs or its epilogue take).
    if (lane < 2 * kLimbs) {
      const bool is_x = lane < kLimbs;
      const int l = is_x ? lane : lane - kLimbs;
      uint32_t it = 0;  // K blocks issued so far: stage = it % kStages, use number = it / kStages
      for (int tile = int(blockIdx.x); tile < tiles_total; tile += int(gridDim.x)) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        const int row = is_x ? l * a.x_plane_rows + m_blk * kTileM : l * a.w_plane_rows + n_blk * kTileN;
        // K rotation: tiles start at different K blocks (integer accumulation is order-independent), so the CTAs do not all ask the
        // L2 for the same 128-byte column of the 512-byte-pitch limb planes at the same time
        int kb = (m_blk * 3 + n_blk) % k_blocks;
        for (int turn = 0; turn < k_blocks; ++turn, ++it, kb = (kb + 1 == k_blocks ? 0 : kb + 1)) {
          const int stage = int(it % kStages);
          if (it >= uint32_t(kStages)) ptx::mbar_wait(empty_bar + stage, ((it / kStages) - 1u) & 1u);
          uint8_t *st = tiles + stage * kStageBytes;
          // (a box may land before lane 0's expect_tx is visible: the transaction count dips below zero, the phase cannot complete
          // before lane 0 has arrived)
          if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, kStageBytes);
          ptx::tma_load_2d(is_x ? &tmap_x : &tmap_w, full_bar + stage, is_x ? st + l * kXBytes : st + kLimbs * kXBytes + l * kWBytes, kb * kBlockK, row);
        }
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0, tile_no = 0;
    for (int tile = int(blockIdx.x); tile < tiles_total; tile += int(gridDim.x), ++tile_no) {
    const uint32_t set = tile_no % kAccSets, use = tile_no / kAccSets;
    if (use != 0) {
      ptx::mbar_wait(acc_free + set, (use - 1u) & 1u);  // this set's previous tile has been read
      ptx::tc_fence_after_sync();
    }
    for (int kb = 0; kb < k_blocks; ++kb, ++it) {
      const int stage = int(it % kStages);
      ptx::mbar_wait(full_bar + stage, (it / kStages) & 1u);
      ptx::tc_fence_after_sync();
      if (lane == 0) {
        const uint32_t base = ptx::smem_u32(tiles + stage * kStageBytes);
        // The weight limbs are BALANCED digits (every limb an s8, model upload), and the three limb tiles of a K block lie one after
        // the other in the stage: one instruction multiplies a frame limb with all the weight limbs it needs — N = 192 or 128 instead
        // of three or two instructions of N = 64, which cost the same 37 ns each (tools/int8_peak.cu) — and the products of
        // consecutive shift classes land in consecutive accumulator columns.  Class s = i + j sits at columns (s − 1)·64:
        //   X₁ · [W₀ W₁ W₂] → classes 1 2 3 (columns 0-191)     X₀ · [W₁ W₂] → classes 1 2 (0-127; X₀·W₀ is bounded, not computed)
        //   X₂ · [W₀ W₁ W₂] → classes 2 3 4 (columns 64-255)
        // The very first step of a tile overwrites: X₁ first (classes 1-3), and class 4 by an instruction of its own.
        const uint32_t d0 = tmem_base + uint32_t(int(set) * kClasses * kTileN);
        const uint32_t xa = base, wa = base + uint32_t(kLimbs * kXBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t ko = uint64_t(k * (kUmmaK / 16));
          const uint64_t dx0 = ptx::smem_desc_k_sw128(xa) + ko, dx1 = ptx::smem_desc_k_sw128(xa + uint32_t(kXBytes)) + ko,
                         dx2 = ptx::smem_desc_k_sw128(xa + uint32_t(2 * kXBytes)) + ko;
          const uint64_t dw0 = ptx::smem_desc_k_sw128(wa) + ko, dw1 = ptx::smem_desc_k_sw128(wa + uint32_t(kWBytes)) + ko,
                         dw2 = ptx::smem_desc_k_sw128(wa + uint32_t(2 * kWBytes)) + ko;
          const bool first = kb == 0 && k == 0;
          ptx::mma_i8_ss(d0, dx1, dw0, idesc_limbs(false, 3 * kTileN), uint32_t(!first));
          ptx::mma_i8_ss(d0, dx0, dw1, idesc_limbs(false, 2 * kTileN), 1u);
          if (first) {
            ptx::mma_i8_ss(d0 + uint32_t(3 * kTileN), dx2, dw2, idesc_limbs(true, kTileN), 0u);
            ptx::mma_i8_ss(d0 + uint32_t(kTileN), dx2, dw0, idesc_limbs(true, 2 * kTileN), 1u);
          } else {
            ptx::mma_i8_ss(d0 + uint32_t(kTileN), dx2, dw0, idesc_limbs(true, 3 * kTileN), 1u);
          }
        }
        ptx::mma_commit(empty_bar + stage);
        if (kb == k_blocks - 1) ptx::mma_commit(acc_bar + set);
      }
      __syncwarp();
    }
    }
  } else if (warp >= 4) {
    const int et = int(threadIdx.x) - 128;
    for (int i = et; i < kLut2Padded / 16; i += kEpiWarps * 32) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(a.lut) + i);
    ptx::named_bar_sync(1, kEpiWarps * 32);
    const int quarter = warp & 3, cq = (warp - 4) >> 2, half = cq >> 1;  // column quarter (16 columns), bitmap word (32 columns)
    InputNodeStats *s_node = reinterpret_cast<InputNodeStats *>(s_lut + kLut2Padded);
    uint32_t *s_unc = reinterpret_cast<uint32_t *>(s_node + kTileN);  // [tile row][word]: the odd column quarters' half words
    uint32_t tile_no = 0;
    for (int tile = int(blockIdx.x); tile < tiles_total; tile += int(gridDim.x), ++tile_no) {
    const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
    const int row = m_blk * kTileM + quarter * 32 + lane;
    const bool row_ok = row < a.M;
    InputRowStats rs;
    rs.a[0] = -1.0f;
    if (row_ok) rs = a.row_stats[row];
    // this tile's node constants: shared memory, read as broadcasts (everybody is done with the previous tile's first)
    ptx::named_bar_sync(1, kEpiWarps * 32);
    for (int i = et; i < kTileN; i += kEpiWarps * 32) {
      InputNodeStats ns{};
      ns.c = -1.0f;
      if (n_blk * kTileN + i < a.H) ns = a.node_stats[n_blk * kTileN + i];
      s_node[i] = ns;
    }
    const uint32_t set = tile_no % kAccSets, use = tile_no / kAccSets;
    if (warp == 4) ptx::mbar_wait_parked(acc_bar + set, use & 1u);
    ptx::named_bar_sync(1, kEpiWarps * 32);
    ptx::tc_fence_after_sync();
    uint32_t unc_mask = 0;  // bit c: column (half·32 + c) of this row is uncertain (this warp fills 16 of the word's bits)
    const int col_base = n_blk * kTileN + half * 32;
    const bool row_cert = row_ok && rs.a[0] > 0.0f;
#pragma unroll 1
    for (int g = 2 * (cq & 1); g < 2 * (cq & 1) + 2; ++g) {
      uint32_t acc[kClasses][8];
#pragma unroll
      for (int s = 0; s < kClasses; ++s)
        tmem_ld_32x8(tmem_base + uint32_t(int(set) * kClasses * kTileN + s * kTileN + half * 32 + g * 8) + (uint32_t(quarter * 32) << 16), acc[s]);
      ptx::tmem_ld_wait();
      if (g == 2 * (cq & 1) + 1) {  // accumulators fully read by this warp: the next tile's MMAs may overwrite them
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free + set);
      }
      uint32_t bytes[2] = {0u, 0u};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const InputNodeStats ns = s_node[half * 32 + g * 8 + c];
        // z in fp32 from the class sums; its evaluation error is part of D (InputRowStats::pp · InputNodeStats::qp)
        float v = __int2float_rn(int32_t(acc[0][c])) * rs.a[1];  // acc[s] holds shift class s + 1
#pragma unroll
        for (int s = 1; s < kClasses; ++s) v = fmaf(__int2float_rn(int32_t(acc[s][c])), rs.a[s + 1], v);
        const float z = fmaf(v, ns.c, ns.bc);
        const float az = fabsf(z);
        float D = fmaf(rs.p, ns.q, fmaf(rs.pp, ns.qp, fmaf(rs.r1, ns.c, fmaf(rs.ar, ns.e, fmaf(1.85e-7f, az, ns.f)))));  // 3.1u = 1.85e-7
        D *= 1.00001f;
        const float k = (z + 12582912.0f) - 12582912.0f;  // round to nearest integer, exact for |z| < 2²²
        const bool in_range = az < 700.0f;
        const bool mid = in_range && fabsf(z - k) + D < 0.5f && fabsf(k) < 640.0f;
        const bool hi = z - D > 639.5f && z + D < 2.0e9f;  // (≥ 2³¹ converts to INT_MIN on x86, i.e. bucket 0: left to the exact path)
        const bool lo = z + D < -639.5f;
        const bool certain = row_cert && ns.c > 0.0f && (mid || hi || lo);
        const int slot = mid ? 2 * __float2int_rn(k) + kLut2Center : (hi ? 2 * kLut2Center : 0);
        const uint32_t byte = certain ? uint32_t(s_lut[slot]) : 0u;
        if (!certain && row_ok && col_base + g * 8 + c < a.H) unc_mask |= 1u << (g * 8 + c);
        bytes[c >> 2] |= byte << (8 * (c & 3));
      }
      if (row_ok && col_base + g * 8 < a.H)  // hidden widths are multiples of 16: an 8-column group is whole or absent
        *reinterpret_cast<uint2 *>(a.out_u8 + size_t(row) * size_t(a.H) + col_base + g * 8) = make_uint2(bytes[0], bytes[1]);
    }
    // the two warps of a bitmap word meet: the odd column quarter hands its 16 bits over and is done
    uint32_t n_unc = uint32_t(__popc(unc_mask));
    if (cq & 1) s_unc[(quarter * 32 + lane) * 2 + half] = unc_mask;
    ptx::named_bar_sync(1, kEpiWarps * 32);
    if (!(cq & 1)) {
    unc_mask |= s_unc[(quarter * 32 + lane) * 2 + half];
    // undecided elements → one bitmap word per (frame, 32 nodes); every word of the bitmap has exactly one writer
    if (row_ok && col_base < a.H) a.unc_bits[size_t(row) * size_t(a.unc_words) + size_t(col_base / 32)] = unc_mask;
    // … and the same 32 frames × 32 nodes transposed (five butterfly steps over the warp): lane c ends up with the word of
    // node col_base + c, bit f = frame f of this warp's block of 32 frames, for the block fix-up kernel
    {
      uint32_t t = unc_mask;
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t p = __shfl_xor_sync(0xffffffffu, t, j);
        t = (lane & j) == 0 ? ((t & m) | ((p & m) << j)) : ((t & ~m) | ((p & ~m) >> j));
      }
      const int row0 = m_blk * kTileM + quarter * 32;
      if (row0 < a.M && col_base + lane < a.H) a.unc_t[size_t(row0 / 32) * size_t(a.H) + size_t(col_base + lane)] = t;
    }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_unc += __shfl_xor_sync(0xffffffffu, n_unc, o);
    if (lane == 0 && n_unc != 0u) atomicAdd(a.unc_count, n_unc);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---- 3. exact arithmetic for the undecided elements -------------------------------------------------------------
// One warp per frame: the transformed frame sits in shared memory, the frame's undecided nodes are compacted from
// its bitmap words into a list, and the eight quads of the warp take eight nodes at a time, exactly as the reference
// computes them (dnn.cc:219-247), weights straight from L2.
constexpr int kFixWarps = 8;
// `warps_per_row` (1, 2, 4 or 8, a launch parameter) warps share a frame when there are too few frames to fill the GPU with
// one warp each; they build the same node list and take its 8-node rounds in turn.
__global__ void __launch_bounds__(kFixWarps * 32) input_fixup_kernel(const InputTcArgs a, const int warps_per_row) {
  __shared__ __align__(16) float s_x[kFixWarps][kInputTcMaxI];
  __shared__ uint16_t s_cols[kFixWarps][1024];
  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  const int quad = lane >> 2, r = lane & 3;
  const int I = a.I;
  const int rows_per_cta = kFixWarps / warps_per_row;
  const int sub = warp % warps_per_row;  // which share of the frame's rounds this warp takes
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  for (int row = int(blockIdx.x) * rows_per_cta + warp / warps_per_row; row < a.M; row += int(gridDim.x) * rows_per_cta) {
    __syncwarp();
    for (int k = lane; k < I; k += 32) s_x[warp][k] = a.xq[size_t(row) * size_t(I) + k];
    for (int w0 = 0; w0 < a.unc_words; w0 += 32) {  // 32 bitmap words = up to 1024 nodes per pass
      const uint32_t bits = (w0 + lane < a.unc_words) ? a.unc_bits[size_t(row) * size_t(a.unc_words) + size_t(w0 + lane)] : 0u;
      const uint32_t mine = uint32_t(__popc(bits));
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t n = __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();
      uint32_t pos = incl - mine, b = bits;
      while (b != 0u) {
        s_cols[warp][pos++] = uint16_t(32 * (w0 + lane) + __ffs(b) - 1);
        b &= b - 1u;
      }
      __syncwarp();
      // a round = 8 nodes, one per quad: thread (quad, r) accumulates SSE lane r of its node.  (Measured on B200, 16384
      // frames: two nodes per quad 349 us, a node per thread with 16-byte loads 372 us, this 318 us — the kernel is
      // bound by the latency of scattered weight-row reads from L2, ≈ 4 TB/s, whatever the mapping.)
      for (uint32_t e0 = 8u * uint32_t(sub); e0 < n; e0 += 8u * uint32_t(warps_per_row)) {
        const bool live = e0 + uint32_t(quad) < n;
        const int col = live ? int(s_cols[warp][e0 + quad]) : int(s_cols[warp][e0]);
        const float *__restrict__ w = a.w0 + size_t(col) * size_t(I);
        const float *x = s_x[warp];
        float acc = 0.0f;
        int k = r;
        for (; k + 28 < I; k += 32) {  // eight steps of this lane at a time: loads first, then the dependent adds
          float wv[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) wv[t] = __ldg(w + k + 4 * t);
#pragma unroll
          for (int t = 0; t < 8; ++t) acc = __fadd_rn(acc, __fmul_rn(x[k + 4 * t], wv[t]));  // dnn.cc:219-247, lane r
        }
        for (; k < I; k += 4) acc = __fadd_rn(acc, __fmul_rn(x[k], __ldg(w + k)));
        const float pair = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
        const float h = __fadd_rn(pair, __shfl_xor_sync(0xffffffffu, pair, 2));  // (l0 + l1) + (l2 + l3), dnn.cc:168-172
        if (live && r == 0) a.out_u8[size_t(row) * size_t(a.H) + col] = a.lut[qsig_slot(__fadd_rn(h, a.bias0[col]))];
      }
    }
  }
}


// ---- 3b. the same arithmetic, by blocks of 32 frames (the default) -------------------------------------------------
// The warp-per-frame kernel above lives on scattered 4-byte reads of weight rows through L1 (≈ 4–5 TB/s of L2 reads at
// best).  Here a CTA owns 32 frames (an unc_t word per node) × a range of nodes: the frames sit in shared memory, one SSE
// lane after the other, the nodes that are undecided for at least one of the 32 frames are compacted into a list, and a
// producer warp brings their weight rows in with plain bulk copies (cp.async.bulk, one 4·I-byte row per copy, kFbSlots
// rows per mbarrier, kFbBufs batches in flight — no L1 miss queue in the way, and a row is fetched once for all of the
// block's frames that need it).  Eight consumer warps take a batch's (node, frame) elements eight at a time — a quad per
// element, thread (quad, r) doing SSE lane r exactly as the reference (dnn.cc:219-247, 168-172): FMUL, FADD, never FMA.
constexpr int kFbFrames = 32;       // = the bits of an unc_t word
constexpr int kFbConsumers = 12;    // warps doing the arithmetic (measured on B200, 16384 | 512 frames: 12 warps and 4 buffers
                                    // 224 | 14.5 us, 20 warps and 5 buffers 262 | 15.9 us: warps polling a barrier for a batch
                                    // that has not landed share the shared-memory pipe with the ones doing the arithmetic)
constexpr int kFbBufs = 4;          // at most: batches in flight, and producer warps: each owns a buffer (issuing a bulk copy takes
                                    // a warp ≈ 33 cycles whatever its size, 16 in a row per batch)
constexpr int kFbThreads = (kFbConsumers + kFbBufs) * 32;
constexpr int kFbSmemLimit = 227 * 1024 - 1024;  // dynamic shared memory a CTA of this file can ask for
constexpr int kFbSlots = 16;        // weight rows per batch
constexpr int kFbPassNodes = 2048;  // nodes whose words are compacted at a time
constexpr int kFbPassLoads = (kFbPassNodes + kFbThreads - 1) / kFbThreads;  // unc_t words per thread and pass
constexpr int kFbPassBatches = kFbPassNodes / kFbSlots;                     // 128: each has its own "landed" barrier
constexpr int kFbMaxChunks = kFbSlots * kFbFrames / 8;                      // 64 chunks of 8 elements in a full batch

struct FbLayout {
  int seg;      // floats per SSE lane of a frame: I/4 rounded up to an ODD number of float4, so that the four lanes of a
                // quad read four different 16-byte bank groups
  int fstride;  // floats per frame
  int wstride;  // bytes per weight-row slot, an odd number of 16-byte units: any eight consecutive slots are conflict-free
  int off_lut, off_word, off_node, off_incl, off_cst, off_x, off_w, n_bufs, total;
};
__host__ __device__ inline FbLayout fb_layout(int I) {
  FbLayout L;
  L.seg = (I / 4 + 3) / 4 * 4;
  if ((L.seg / 4) % 2 == 0) L.seg += 4;
  L.fstride = 4 * L.seg;
  L.wstride = 4 * I;
  if ((L.wstride / 16) % 2 == 0) L.wstride += 16;
  L.off_lut = (8 * (kFbPassBatches + kFbBufs) + 32 + 15) / 16 * 16;  // the barriers and the list counter come first
  L.off_word = L.off_lut + kLut2Padded;
  L.off_node = L.off_word + 4 * kFbPassNodes;
  L.off_incl = L.off_node + 2 * kFbPassNodes;
  L.off_cst = L.off_incl + 2 * kFbPassNodes;
  L.off_x = L.off_cst + 2 * (kFbPassBatches + 16);
  L.off_w = L.off_x + 4 * kFbFrames * L.fstride;
  L.n_bufs = (kFbSmemLimit - L.off_w) / (kFbSlots * L.wstride);
  L.n_bufs = L.n_bufs > kFbBufs ? kFbBufs : L.n_bufs;
  L.total = L.off_w + L.n_bufs * kFbSlots * L.wstride;
  return L;
}

__global__ void __launch_bounds__(kFbThreads) input_fixup_block_kernel(const InputTcArgs a, const int words_per_cta, const int debug) {
  extern __shared__ __align__(128) uint8_t fb_smem[];
  const FbLayout L = fb_layout(a.I);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(fb_smem), *empty_bar = full_bar + kFbPassBatches;
  uint32_t *s_count = reinterpret_cast<uint32_t *>(fb_smem + 8 * (kFbPassBatches + kFbBufs));
  uint16_t *l_incl = reinterpret_cast<uint16_t *>(fb_smem + L.off_incl);  // elements up to and including this entry, within its batch
  uint16_t *s_cst = reinterpret_cast<uint16_t *>(fb_smem + L.off_cst);    // chunks before batch b in this pass; [n_batches] = all
  uint8_t *s_lut = fb_smem + L.off_lut;
  uint32_t *l_word = reinterpret_cast<uint32_t *>(fb_smem + L.off_word);
  uint16_t *l_node = reinterpret_cast<uint16_t *>(fb_smem + L.off_node);
  float *s_x = reinterpret_cast<float *>(fb_smem + L.off_x);
  uint8_t *ring = fb_smem + L.off_w;
  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  const int quad = lane >> 2, r = lane & 3;
  const int I = a.I, n4 = I / 4;
  const int fb = int(blockIdx.x), f0 = fb * kFbFrames;
  const int n_lo = int(blockIdx.y) * words_per_cta * 32;
  const int n_hi = min(a.H, n_lo + words_per_cta * 32);
  // a batch's "landed" barrier is used once per pass (any consumer may wait on any batch, in any order, with no phase to
  // lose track of); a buffer's "free again" barrier takes one arrival per chunk of 8 elements, the producer standing in
  // for the chunks a batch does not have, and only the buffer's producer waits on it
  if (threadIdx.x < kFbPassBatches) ptx::mbar_init(full_bar + threadIdx.x, 1);
  if (threadIdx.x < kFbBufs) ptx::mbar_init(empty_bar + threadIdx.x, kFbMaxChunks);
  ptx::fence_barrier_init();
  for (int i = int(threadIdx.x); i < kLut2Padded / 16; i += kFbThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(a.lut) + i);
  ptx::griddep_wait();  // everything above ran under the tensor-core kernel's tail
  ptx::griddep_launch_dependents();
  // the CTA's bitmap words (at most kFbPassNodes of them: the launcher splits wider layers), before anything else that waits for memory
  uint32_t wv[kFbPassLoads];
#pragma unroll
  for (int u = 0; u < kFbPassLoads; ++u) {
    const int i = n_lo + int(threadIdx.x) + u * kFbThreads;
    wv[u] = i < n_hi ? __ldg(a.unc_t + size_t(fb) * size_t(a.H) + size_t(i)) : 0u;
  }
  // the block's transformed frames, eight 16-byte loads per thread, all in flight; they go to shared memory (s_x[f][r][t] =
  // x'[f][4t + r]) only after the node list has been built, so their latency hides under that work
  static_assert(kFbFrames * (kInputTcMaxI / 4) <= 8 * kFbThreads, "a block's frames are one round of eight loads per thread");
  const int rows = min(kFbFrames, a.M - f0);
  const float4 *xsrc = reinterpret_cast<const float4 *>(a.xq + size_t(f0) * size_t(I));  // rows are contiguous: [rows][n4] float4
  float4 xv[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int idx = int(threadIdx.x) + u * kFbThreads;
    xv[u] = idx < rows * n4 ? __ldg(xsrc + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t n_bufs = uint32_t(L.n_bufs);  // batch b uses buffer b % n_bufs for the (b / n_bufs)-th time
  const int p_lo = n_lo;
  {
    if (threadIdx.x == 0) *s_count = 0u;
    __syncthreads();  // barriers initialised, counter zero
#pragma unroll
    for (int u = 0; u < kFbPassLoads; ++u) {  // nodes with a non-zero word → list (order irrelevant)
      const int i = int(threadIdx.x) + u * kFbThreads;
      const uint32_t w = wv[u];
      const uint32_t nz = __ballot_sync(0xffffffffu, w != 0u);
      if (nz != 0u) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(s_count, uint32_t(__popc(nz)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (w != 0u) {
          const uint32_t pos = base + uint32_t(__popc(nz & ((1u << lane) - 1u)));
          l_word[pos] = w;
          l_node[pos] = uint16_t(p_lo + i);
        }
      }
    }
    __syncthreads();
    const uint32_t n_list = *s_count;
    const uint32_t n_batches = (n_list + kFbSlots - 1) / kFbSlots;
    // per batch, once for everybody: the running element count of its entries and how many chunks of 8 that makes
    for (uint32_t b = uint32_t(warp); b < n_batches; b += kFbThreads / 32) {
      const uint32_t entry = b * kFbSlots + uint32_t(lane & (kFbSlots - 1));
      uint32_t incl = entry < n_list ? uint32_t(__popc(l_word[entry])) : 0u;
#pragma unroll
      for (int o = 1; o < kFbSlots; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o, kFbSlots);
        if ((lane & (kFbSlots - 1)) >= o) incl += v;
      }
      if (lane < kFbSlots) l_incl[entry] = uint16_t(incl);
      if (lane == kFbSlots - 1) s_cst[b + 1] = uint16_t((incl + 7u) / 8u);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // the frames have arrived by now
      const int idx = int(threadIdx.x) + u * kFbThreads;
      if (idx < rows * n4) {
        const int f = idx / n4, t = idx - f * n4;
        float *d = s_x + f * L.fstride + t;
        d[0] = xv[u].x;
        d[L.seg] = xv[u].y;
        d[2 * L.seg] = xv[u].z;
        d[3 * L.seg] = xv[u].w;
      }
    }
    __syncthreads();
    if (warp == 0) {  // chunk counts → running totals (s_cst[b] = chunks before batch b)
      constexpr int kPer = kFbPassBatches / 32;
      uint32_t v[kPer], sum = 0;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const uint32_t b = uint32_t(lane * kPer + u);
        v[u] = b < n_batches ? uint32_t(s_cst[b + 1]) : 0u;
        sum += v[u];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - sum;
      __syncwarp();
      if (lane == 0) s_cst[0] = 0;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const uint32_t b = uint32_t(lane * kPer + u);
        run += v[u];
        if (b < n_batches) s_cst[b + 1] = uint16_t(run);
      }
    }
    __syncthreads();
    if (warp >= kFbConsumers) {
      // producers: a batch = up to kFbSlots weight rows, one bulk copy each, all landing on the batch's barrier
      const uint32_t buf = uint32_t(warp - kFbConsumers);
      for (uint32_t b = buf, use = 0; b < n_batches && buf < n_bufs; b += n_bufs, ++use) {
        if (use != 0u) ptx::mbar_wait(empty_bar + buf, (use - 1u) & 1u);
        const uint32_t cnt = min(uint32_t(kFbSlots), n_list - b * kFbSlots);
        const uint32_t row_bytes = (debug & 4) ? 16u : uint32_t(4 * I);
        if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + b, (debug & 2) ? 0u : cnt * row_bytes);
        __syncwarp();
        if (uint32_t(lane) < cnt && !(debug & 2))
          ptx::bulk_load(ring + (buf * kFbSlots + uint32_t(lane)) * uint32_t(L.wstride), a.w0 + size_t(l_node[b * kFbSlots + uint32_t(lane)]) * size_t(I),
                         row_bytes, full_bar + b);
        const uint32_t absent = uint32_t(kFbMaxChunks) - (uint32_t(s_cst[b + 1]) - uint32_t(s_cst[b]));
        if (lane == 0 && absent != 0u) ptx::mbar_arrive_n(empty_bar + buf, absent);
      }
    } else {
      uint32_t c_hi = 0;
      for (uint32_t b = 0; b < n_batches; ++b) {
        const uint32_t c_lo = c_hi;
        c_hi = s_cst[b + 1];
        // chunk c of the pass belongs to consumer warp c % kFbConsumers; most batches have nothing for this warp
        uint32_t j = (uint32_t(warp) + uint32_t(kFbConsumers) - c_lo % uint32_t(kFbConsumers)) % uint32_t(kFbConsumers);
        const uint32_t n_chunks = c_hi - c_lo;
        if (j >= n_chunks) continue;
        const uint32_t buf = b % n_bufs;
        // both half-warps hold the batch's entries: word, node, running element count
        const uint32_t entry = b * kFbSlots + uint32_t(lane & (kFbSlots - 1));
        const bool have = entry < n_list;
        const uint32_t word = have ? l_word[entry] : 0u;
        const uint32_t node = have ? uint32_t(l_node[entry]) : 0u;
        const uint32_t incl = l_incl[entry];
        const uint32_t mine = uint32_t(__popc(word));
        const uint32_t total = __shfl_sync(0xffffffffu, incl, kFbSlots - 1);
        ptx::mbar_wait(full_bar + b, 0u);
        for (; j < n_chunks; j += kFbConsumers) {
          if (!(debug & 1)) {
          const uint32_t e = 8u * j + uint32_t(quad);
          const bool live = e < total;
          const uint32_t ee = live ? e : 0u;
          int idx = 0;  // the entry whose elements include number ee: as many entries end at or before it
#pragma unroll
          for (int i = 0; i < kFbSlots - 1; ++i) idx += int(__shfl_sync(0xffffffffu, incl, i) <= ee);
          const uint32_t ww = __shfl_sync(0xffffffffu, word, idx);
          const int col = int(__shfl_sync(0xffffffffu, node, idx));
          uint32_t rank = ee - (__shfl_sync(0xffffffffu, incl, idx) - __shfl_sync(0xffffffffu, mine, idx));
          uint32_t rest = ww;
          while (rank != 0u) {  // the rank-th set bit = the frame
            rest &= rest - 1u;
            --rank;
          }
          const int f = __ffs(int(rest)) - 1;
          const float bias = __ldg(a.bias0 + col);  // needed at the very end: its latency hides under the dot product
          const float *wp = reinterpret_cast<const float *>(ring + (buf * kFbSlots + uint32_t(idx)) * uint32_t(L.wstride)) + r;
          const float *xp = s_x + f * L.fstride + r * L.seg;
          float acc = 0.0f;
          int t = 0;
#pragma unroll 2
          for (; t + 4 <= n4; t += 4) {  // dnn.cc:219-247, lane r: k = 4t + r
            const float4 xv = *reinterpret_cast<const float4 *>(xp + t);
            const float w0 = wp[4 * t], w1 = wp[4 * t + 4], w2 = wp[4 * t + 8], w3 = wp[4 * t + 12];
            acc = __fadd_rn(acc, __fmul_rn(xv.x, w0));
            acc = __fadd_rn(acc, __fmul_rn(xv.y, w1));
            acc = __fadd_rn(acc, __fmul_rn(xv.z, w2));
            acc = __fadd_rn(acc, __fmul_rn(xv.w, w3));
          }
          for (; t < n4; ++t) acc = __fadd_rn(acc, __fmul_rn(xp[t], wp[4 * t]));
          const float pair = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
          const float h = __fadd_rn(pair, __shfl_xor_sync(0xffffffffu, pair, 2));  // (l0 + l1) + (l2 + l3), dnn.cc:168-172
          if (live && r == 0) a.out_u8[size_t(f0 + f) * size_t(a.H) + size_t(col)] = s_lut[qsig_slot(__fadd_rn(h, bias))];
          }
          __syncwarp();  // every lane has read its operands: the chunk no longer needs the buffer
          if (lane == 0) ptx::mbar_arrive(empty_bar + buf);
        }
      }
    }
  }
}

}  // namespace

cudaError_t input_tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(input_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(input_fixup_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFbSmemLimit);
}

bool input_tc_supported(int I, int H) { return I > 0 && I <= kInputTcMaxI && I % 4 == 0 && H % 16 == 0 && H <= 65536; }

// Enqueues the three kernels.  tmap_x: [3 · x_plane_rows][512] u8, box 128 rows × 128 B, 128B swizzle;
// tmap_w: [3 · w_plane_rows][512], box 64 rows × 128 B.
cudaError_t launch_input_tc(const CUtensorMap &tmap_x, const CUtensorMap &tmap_w, const InputTcArgs &a, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  cudaError_t e = launch_pdl(input_prep_kernel, dim3((a.M + 7) / 8), dim3(256), size_t(0), stream, pdl_enabled(), a);
  if (e != cudaSuccess) return e;
  const int tiles = ((a.H + kTileN - 1) / kTileN) * ((a.M + kTileM - 1) / kTileM);
  e = launch_pdl(input_tc_kernel, dim3(tiles < a.num_sms ? tiles : a.num_sms), dim3(kTcThreads), size_t(kTcSmem), stream, pdl_enabled(), tmap_x, tmap_w, a);
  if (e != cudaSuccess) return e;
  static const bool by_warp = [] {  // FDNN_FIXUP=warp: the warp-per-frame kernel (A/B measurements)
    const char *v = std::getenv("FDNN_FIXUP");
    return v != nullptr && v[0] == 'w';
  }();
  if (by_warp) {
    // enough warps to keep every SM busy: several warps per frame when the batch is short
    int warps_per_row = 1;
    while (warps_per_row < kFixWarps && a.M * warps_per_row < a.fixup_ctas * 2) warps_per_row *= 2;
    const int fix_ctas = (a.M * warps_per_row + kFixWarps - 1) / kFixWarps;
    return launch_pdl(input_fixup_kernel, dim3(fix_ctas < a.fixup_ctas ? fix_ctas : a.fixup_ctas), dim3(kFixWarps * 32), size_t(0), stream, pdl_enabled(), a, warps_per_row);
  }
  // a CTA per block of 32 frames; short batches also split the nodes so that one wave of CTAs covers the GPU
  const int fblocks = (a.M + kFbFrames - 1) / kFbFrames;
  int splits = a.num_sms / fblocks;
  const int min_splits = (a.unc_words * 32 + kFbPassNodes - 1) / kFbPassNodes;  // a CTA lists at most kFbPassNodes nodes
  splits = splits < min_splits ? min_splits : (splits > a.unc_words ? a.unc_words : splits);
  const int words_per_cta = (a.unc_words + splits - 1) / splits;
  splits = (a.unc_words + words_per_cta - 1) / words_per_cta;
  static const int debug = [] {  // FDNN_FB_DEBUG: 1 no arithmetic, 2 no copies, 4 16-byte copies (timing experiments; results wrong)
    const char *v = std::getenv("FDNN_FB_DEBUG");
    return v != nullptr ? std::atoi(v) : 0;
  }();
  return launch_pdl(input_fixup_block_kernel, dim3(fblocks, splits), dim3(kFbThreads), size_t(fb_layout(a.I).total), stream, pdl_enabled(), a, words_per_cta, debug);
}

}  // namespace fdnn
