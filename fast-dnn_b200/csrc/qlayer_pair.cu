// Int8 layer on CTA pairs:  u8 activations [M×K] · s8 weights [N×K]ᵀ → s32 with tcgen05.mma
// cta_group::2, then the reference's per-element tail in the epilogue.  Same job as qlayer_tc.cu
// (reference src/cpp/dnn.cc:289-349 quantizedNodeSum / QuantizedLayerActivations, :250-286 AddBias +
// QuantizedSigmoid, :442-447 output "+= bias"), different machine mapping.
//
// Why pairs.  Measured on B200 (profiles/r1_experiments.md): with one CTA per 128×BN tile the layer is
// bounded by the L2 → shared-memory operand feed (≈ 33-36 B/clk/SM through TMA), not by the tensor
// pipe: a 128×256 tile needs 48 KB of operands per 512 tensor-pipe cycles.  Two CTAs on the two SMs
// of a TPC that issue ONE 256×BN MMA share the weight tile: each stages its own 128 activation rows
// and only HALF of the weight rows, and the tensor cores read the peer's half over the pair link —
// 32 KB per SM per 512 cycles, and one third fewer shared-memory wavefronts for the same math.
//
// Roles per CTA (28 warps, as in qlayer_tc.cu): warp 0 TMA producer (both CTAs; all bytes are counted
// on the LEADER's full barrier via cp.async.bulk.tensor.cta_group::2), warp 1 MMA issuer (leader
// only; tcgen05.commit multicasts "stage consumed" / "accumulator ready" to both CTAs), warp 2 TMEM
// allocator (cta_group::2, both CTAs), warps 4-11 saturation scan of this CTA's 128 rows, warps 12-27
// epilogue of this CTA's 128 rows × BN columns out of its own tensor memory.
//
// The scan reads a stage AFTER the MMA has consumed it (it waits for the commit, not for TMA): the
// peer has no barrier of its own that tells it "my tile landed", the commit is visible in both CTAs,
// and the scan's result is only needed at the end of the tile anyway.  A stage returns to the producer
// when the four scan warps that own it have arrived.

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kBlockM = 128;  // rows per CTA; the pair's MMA is 256 rows
constexpr int kBlockK = 128;  // bytes of K per pipeline stage = one 128B swizzle atom
constexpr int kUmmaK = 32;
constexpr int kScanSets = 2;
constexpr int kScanWarps = 4 * kScanSets;
constexpr int kEpilogueWarps = 16;
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kFirstScanWarp = 4, kFirstEpilogueWarp = kFirstScanWarp + kScanWarps;
constexpr int kThreads = (kFirstEpilogueWarp + kEpilogueWarps) * 32;  // 896
constexpr int kAccStages = 2;
constexpr int kEntCap = 1024;   // risk entries of the tile staged (packed) in shared memory; the rest is read from global
constexpr int kPtrSlots = 132;  // ≥ k_blocks + 1 → K ≤ 16768
constexpr int kRowEvents = 8;

static_assert(kBlockK == kFixKBlock, "risk-list order is tied to the tiling");

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int BN>
struct PairConfig {
  static constexpr int kABytes = kBlockM * kBlockK;     // this CTA's 128 activation rows
  static constexpr int kBBytes = (BN / 2) * kBlockK;    // this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 6 : 8;
  static constexpr int kTmemCols = kAccStages * BN;
  static constexpr int kColsPerWarp = BN / 4;
  static constexpr int kChunks = kColsPerWarp / 16;
  static constexpr int kBiasBytes = kAccStages * BN * 4;
  static constexpr int kEntBytes = kAccStages * kEntCap * 4;
  static constexpr int kPtrBytes = kAccStages * kPtrSlots * 4;
  static constexpr int kEvBytes = kAccStages * kBlockM * kRowEvents * 4;
  static constexpr int kCntBytes = kAccStages * kBlockM * 4;
  static constexpr int kBarBytes = (3 * kStages + 4 * kAccStages) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBiasBytes + kLut2Padded + kEntBytes + kPtrBytes + kEvBytes + kCntBytes + kBarBytes;
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte alignment of swizzled tiles");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BN, bool kLogits>
__global__ void __launch_bounds__(kThreads, 1)
qlayer_pair_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w, const QLayerArgs args) {
  using Cfg = PairConfig<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t *tiles = smem;
  float *s_bias = reinterpret_cast<float *>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint8_t *s_lut = reinterpret_cast<uint8_t *>(s_bias) + Cfg::kBiasBytes;
  uint32_t *s_ent = reinterpret_cast<uint32_t *>(s_lut + kLut2Padded);
  uint32_t *s_ptr = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_ent) + Cfg::kEntBytes);
  uint32_t *s_rowev = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_ptr) + Cfg::kPtrBytes);
  uint32_t *s_rowcnt = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_rowev) + Cfg::kEvBytes);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_rowcnt) + Cfg::kCntBytes);  // leader's copy is the live one
  uint64_t *done_bar = full_bar + Cfg::kStages;    // MMA has consumed the stage (commit, both CTAs)
  uint64_t *empty_bar = done_bar + Cfg::kStages;   // this CTA's scan has released the stage
  uint64_t *tmem_full_bar = empty_bar + Cfg::kStages;
  uint64_t *tmem_empty_bar = tmem_full_bar + kAccStages;  // this CTA's epilogue has drained the stage (for the scan)
  uint64_t *acc_free_bar = tmem_empty_bar + kAccStages;   // both CTAs' epilogues have (leader's copy, for the MMA)
  uint64_t *scan_done_bar = acc_free_bar + kAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(scan_done_bar + kAccStages);

  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  if (threadIdx.x == 0) stamp(args.timeline, 0);
  const int M = args.M, N = args.N, K = args.K;
  const int m_pairs = (M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int n_blocks = (N + BN - 1) / BN;
  const int tiles_total = m_pairs * n_blocks;
  const int k_blocks = (K + kBlockK - 1) / kBlockK;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int first_ct = int(blockIdx.x) / 2, ct_step = int(gridDim.x) / 2;
  // tile ct → (pair of M blocks, N block); both CTAs of a pair walk the same list
  auto decode = [&](int ct, int &m_blk, int &n_blk) {
    m_blk = 2 * (ct / n_blocks) + int(rank);
    n_blk = ct % n_blocks;
  };
  // K rotation: pairs that share an operand start at different K blocks (integer accumulation is
  // order-independent); both CTAs of a pair must agree, so it depends on the pair's tile only
  auto first_k_block = [&](int ct) { return ((ct % n_blocks) + 5 * (ct / n_blocks)) % k_blocks; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_act);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(done_bar + i, 1);
      ptx::mbar_init(empty_bar + i, 4);
    }
    for (int i = 0; i < kAccStages; ++i) {
      ptx::mbar_init(tmem_full_bar + i, 1);
      ptx::mbar_init(tmem_empty_bar + i, kEpilogueWarps);
      ptx::mbar_init(acc_free_bar + i, 2 * kEpilogueWarps);
      ptx::mbar_init(scan_done_bar + i, kScanWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's barriers must exist before TMA, commits or arrivals can target them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  if (threadIdx.x == 0) stamp(args.timeline, 1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): my 128 activation rows + my half of the weight rows =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
        int m_blk, n_blk;
        decode(ct, m_blk, n_blk);
        int kb = first_k_block(ct);
        for (int i = 0; i < k_blocks; ++i, kb = (kb + 1 == k_blocks ? 0 : kb + 1)) {
          ptx::mbar_wait_parked(empty_bar + stage, phase ^ 1);
          uint8_t *sa = tiles + stage * Cfg::kStageBytes;
          uint8_t *sb = sa + Cfg::kABytes;
          if (leader) ptx::mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);  // both CTAs' bytes land on this barrier
          const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(full_bar + stage), 0);
          ptx::tma_load_2d_pair(&tmap_act, bar, sa, kb * kBlockK, m_blk * kBlockM);
          ptx::tma_load_2d_pair(&tmap_w, bar, sb, kb * kBlockK, n_blk * BN + int(rank) * (BN / 2));
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane of the leader CTA issues for the pair =====
    if (leader) {
      constexpr uint32_t idesc = ptx::idesc_i8_u8s8_pair(BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
        ptx::mbar_wait_parked_cluster(acc_free_bar + acc, acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait_parked(full_bar + stage, phase);
          ptx::tc_fence_after_sync();
          if (lane == 0) {
            if (kb == 0) stamp(args.timeline, 2);
            const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::kStageBytes);
            const uint64_t da = ptx::smem_desc_k_sw128(a_addr), db = ptx::smem_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::mma_i8_ss_pair(d_tmem, da + uint64_t(k * (kUmmaK / 16)), db + uint64_t(k * (kUmmaK / 16)), idesc, uint32_t((kb | k) != 0));
            }
            ptx::mma_commit_pair(done_bar + stage);
            if (kb == k_blocks - 1) {
              ptx::mma_commit_pair(tmem_full_bar + acc);
              stamp(args.timeline, 3);
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kFirstScanWarp && warp < kFirstEpilogueWarp) {
    // ===== saturation scan of this CTA's rows (see qlayer_tc.cu for the lane layout) =====
    const int st = int(threadIdx.x) - kFirstScanWarp * 32;
    const int sset = st / kBlockM;
    const int row_sub = lane & 7, ent_sub = lane >> 3;
    const int row_base = ((st % kBlockM) / 32) * 32 + row_sub;  // rows row_base + 8j, j < 4
    const uint32_t swz = uint32_t(row_sub) << 4;
    const int kbn = args.fix.k_blocks;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t it = 0;  // running K-block count across tiles: stage = it % kStages, phase = (it / kStages) & 1
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      decode(ct, m_blk, n_blk);
      const int kb0 = first_k_block(ct);
      auto k_block_of = [&](int turn) { return (kb0 + turn) % k_blocks; };
      const bool scan_on = !(args.debug_flags & 1);
      const uint32_t *gp = args.fix.ptr + size_t(n_blk) * kbn;
      uint32_t *P = s_ptr + acc * kPtrSlots;
      uint32_t *E = s_ent + acc * kEntCap;
      // the event slots of this accumulator stage are free once its previous tile has been drained
      ptx::mbar_wait_parked(tmem_empty_bar + acc, acc_phase ^ 1);
      const uint32_t ent_begin = __ldg(gp);
      for (int i = st; i <= kbn; i += kScanThreads) P[i] = scan_on ? __ldg(gp + i) - ent_begin : 0u;
      const uint32_t n_ent = scan_on ? __ldg(gp + kbn) - ent_begin : 0u;
      const uint32_t staged = min(n_ent, uint32_t(kEntCap));
      // staged form, one word per entry: w0 | w1 << 8 | (node − n0) << 16 | (byte offset of the pair
      // inside its 128-byte K block) << 24 — dp4a of that word with the zero-extended activation
      // pair is exactly a0·w0 + a1·w1
      const uint2 *gent = reinterpret_cast<const uint2 *>(args.fix.ent) + ent_begin;
      for (uint32_t e = uint32_t(st); e < staged; e += kScanThreads) {
        const uint2 fe = __ldg(gent + e);
        E[e] = (fe.x >> 16) | ((fe.y - uint32_t(n_blk * BN)) << 16) | (((2u * (fe.x & 0xffffu)) & 127u) << 24);
      }
      uint32_t *cnt_s = s_rowcnt + acc * kBlockM;
      uint32_t *ev_s = s_rowev + acc * kBlockM * kRowEvents;
      if (st < kBlockM) cnt_s[st] = 0;
      ptx::named_bar_sync(2, kScanThreads);
      auto record = [&](int row, int v, uint32_t node_local) {  // rare: a row's events come from several threads
        const int d = max(min(v, 32767), -32768) - v;
        const uint32_t slot = atomicAdd(cnt_s + row, 1u);
        if (slot < uint32_t(kRowEvents)) ev_s[row * kRowEvents + slot] = (node_local << 24) | (uint32_t(d) & 0xffffffu);
      };
      uint32_t w0 = 0, w1 = 0;
      auto fetch = [&](uint32_t r0, uint32_t r_end) {
        const uint32_t last = max(r_end, 1u) - 1u;
        w0 = E[min(r0 + uint32_t(ent_sub), last)];
        w1 = E[min(r0 + 4u + uint32_t(ent_sub), last)];
      };
      int kb = int((uint32_t(sset) + kScanSets - it % kScanSets) % kScanSets);  // first pipeline turn of this tile owned by this set
      uint32_t r0 = 0, r1 = 0;
      if (kb < k_blocks) {
        r0 = P[k_block_of(kb)];
        r1 = P[k_block_of(kb) + 1];
        fetch(r0, min(r1, staged));
      }
      for (; kb < k_blocks; kb += kScanSets) {
        const uint32_t g = it + uint32_t(kb);
        const int stage = int(g % uint32_t(Cfg::kStages));
        ptx::mbar_wait_parked(done_bar + stage, (g / uint32_t(Cfg::kStages)) & 1u);
        const uint32_t a_swz = (ptx::smem_u32(tiles + stage * Cfg::kStageBytes) + uint32_t(row_base) * 128u) ^ swz;
        const uint32_t fast_end = min(r1, staged);
        for (uint32_t e = r0; e < fast_end; e += 8) {
          if (e != r0) fetch(e, fast_end);
          uint32_t a0[4], a1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a0[j] = ptx::lds_u16((a_swz + uint32_t(j) * 1024u) ^ (w0 >> 24));
            a1[j] = ptx::lds_u16((a_swz + uint32_t(j) * 1024u) ^ (w1 >> 24));
          }
          int v0[4], v1[4];
          uint32_t fired = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v0[j] = dp4a_u8s8(a0[j], w0, 0);
            v1[j] = dp4a_u8s8(a1[j], w1, 0);
            fired |= (uint32_t(v0[j] + 32768) | uint32_t(v1[j] + 32768)) >> 16;  // non-zero ⇔ some v ∉ [−32768, 32767]
          }
          if (fired != 0) {
            const bool ok0 = e + uint32_t(ent_sub) < fast_end, ok1 = e + 4u + uint32_t(ent_sub) < fast_end;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (ok0 && uint32_t(v0[j] + 32768) > 65535u) record(row_base + 8 * j, v0[j], (w0 >> 16) & 0xffu);
              if (ok1 && uint32_t(v1[j] + 32768) > 65535u) record(row_base + 8 * j, v1[j], (w1 >> 16) & 0xffu);
            }
          }
        }
        // beyond the staging capacity (dense risk lists): one entry per pass, 32 rows per warp
        for (uint32_t e = max(r0, staged); e < r1; ++e) {
          const uint2 fe = __ldg(gent + e);
          const uint32_t b = (2u * (fe.x & 0xffffu)) & 127u;
          const int row = (st % kBlockM);
          const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::kStageBytes) + uint32_t(row) * 128u;
          const uint32_t a01s = ptx::lds_u16(a_addr + (((b & 0x70u) ^ (uint32_t(row & 7) << 4)) | (b & 15u)));
          const int v = dp4a_u8s8(a01s, fe.x >> 16, 0);
          if (uint32_t(v + 32768) > 65535u) record(row, v, fe.y - uint32_t(n_blk * BN));
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(empty_bar + stage);
        if (kb + kScanSets < k_blocks) {
          r0 = P[k_block_of(kb + kScanSets)];
          r1 = P[k_block_of(kb + kScanSets) + 1];
          fetch(r0, min(r1, staged));
        }
      }
      it += uint32_t(k_blocks);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(scan_done_bar + acc);
      if (st == 0) stamp(args.timeline, 4);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= kFirstEpilogueWarp) {
    // ===== epilogue: this CTA's TMEM → registers → + saturation events → reference tail → global =====
    const int et = int(threadIdx.x) - kFirstEpilogueWarp * 32;
    const int quarter = warp & 3;
    const int col_group = (warp - kFirstEpilogueWarp) >> 2;
    const int row_local = quarter * 32 + lane;
    if (!kLogits) {
      for (int i = et; i < kLut2Padded / 16; i += kEpilogueThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(args.lut) + i);
    }
    const uint32_t leader_acc_free = ptx::mapa_u32(ptx::smem_u32(acc_free_bar), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      decode(ct, m_blk, n_blk);
      const int n0 = n_blk * BN;
      const int row = m_blk * kBlockM + row_local;
      const bool row_ok = row < M;
      const int col0 = n0 + col_group * Cfg::kColsPerWarp;
      const int n_valid = max(0, min(Cfg::kChunks, (N - col0 + 15) / 16));  // warp-uniform

      float *bias_s = s_bias + acc * BN;
      for (int i = et; i < BN; i += kEpilogueThreads) bias_s[i] = (n0 + i < N) ? __ldg(args.bias + n0 + i) : 0.0f;
      ptx::named_bar_sync(1, kEpilogueThreads);

      ptx::mbar_wait_parked(tmem_full_bar + acc, acc_phase);
      if (et == 0) stamp(args.timeline, 5);
      ptx::mbar_wait_parked(scan_done_bar + acc, acc_phase);
      ptx::tc_fence_after_sync();
      if (et == 0) stamp(args.timeline, 7);
      const uint32_t n_ev = s_rowcnt[acc * kBlockM + row_local];
      const uint32_t *ev = s_rowev + (acc * kBlockM + row_local) * kRowEvents;
      const uint32_t t_addr = tmem_base + uint32_t(acc * BN + col_group * Cfg::kColsPerWarp) + (uint32_t(quarter * 32) << 16);
      auto release_acc = [&]() {  // accumulator stage and its event slots fully read by this warp
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(tmem_empty_bar + acc);
          ptx::mbar_arrive_remote(leader_acc_free + uint32_t(acc) * 8u);
        }
      };
      if (n_valid == 0) release_acc();
#pragma unroll
      for (int j = 0; j < Cfg::kChunks; ++j) {
        if (j < n_valid) {
          uint32_t raw[16];
          ptx::tmem_ld_32x16(t_addr + uint32_t(j * 16), raw);
          ptx::tmem_ld_wait();
          int32_t s[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i] = int32_t(raw[i]);
          const int col = col0 + j * 16;
          if (n_ev != 0 && row_ok) {
            if (n_ev <= uint32_t(kRowEvents)) {
              for (uint32_t k = 0; k < n_ev; ++k) {
                const uint32_t e = ev[k];
                const uint32_t rel = (e >> 24) - uint32_t(col - n0);
                const int d = int(e << 8) >> 8;
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] += (rel == uint32_t(i)) ? d : 0;
              }
            } else {
              brute_force_corrections(s, row, col, args);  // more events than slots in this row: recompute them
            }
          }
          if (j == n_valid - 1) release_acc();  // before the math: the MMA of tile i+2 can start
          if (row_ok) finish_chunk<kLogits>(s, row, col, args, bias_s + (col - n0), s_lut);
        }
      }
      if (et == 0) stamp(args.timeline, 6);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // nobody leaves while the peer's MMA may read its tiles or arrive on its barriers
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool kLogits>
cudaError_t launch_one(const CUtensorMap &ta, const CUtensorMap &tw, const QLayerArgs &a, int num_sms, cudaStream_t stream) {
  using Cfg = PairConfig<BN>;
  const int m_pairs = (a.M + 2 * kBlockM - 1) / (2 * kBlockM), n_blocks = (a.N + BN - 1) / BN;
  const int pair_tiles = m_pairs * n_blocks;
  const int max_pairs = num_sms / 2;
  const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(pairs * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = size_t(Cfg::kSmemBytes);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  attr[n_attr].id = cudaLaunchAttributeClusterDimension;
  attr[n_attr].val.clusterDim.x = 2;
  attr[n_attr].val.clusterDim.y = 1;
  attr[n_attr].val.clusterDim.z = 1;
  ++n_attr;
  cfg.attrs = attr;
  cfg.numAttrs = unsigned(n_attr);
  return cudaLaunchKernelEx(&cfg, qlayer_pair_kernel<BN, kLogits>, ta, tw, a);
}

template <int BN>
cudaError_t configure_one() {
  cudaError_t e = cudaFuncSetAttribute(qlayer_pair_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairConfig<BN>::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(qlayer_pair_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairConfig<BN>::kSmemBytes);
}

}  // namespace

cudaError_t qlayer_pair_configure() {
  cudaError_t e = configure_one<64>();
  if (e == cudaSuccess) e = configure_one<128>();
  if (e == cudaSuccess) e = configure_one<256>();
  return e;
}

// `tmap_act`: box of 128 rows; `tmap_w`: box of block_n / 2 rows.
cudaError_t launch_qlayer_pair(const CUtensorMap &tmap_act, const CUtensorMap &tmap_w, const QLayerArgs &a, bool logits, int block_n, int num_sms,
                               cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  switch (block_n) {
    case 64: return logits ? launch_one<64, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<64, false>(tmap_act, tmap_w, a, num_sms, stream);
    case 128: return logits ? launch_one<128, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<128, false>(tmap_act, tmap_w, a, num_sms, stream);
    case 256: return logits ? launch_one<256, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<256, false>(tmap_act, tmap_w, a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace fdnn
