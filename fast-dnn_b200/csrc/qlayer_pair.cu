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
// Roles per CTA (28 warps): warp 0 TMA producer (both CTAs; all bytes are counted on the LEADER's full
// barrier via cp.async.bulk.tensor.cta_group::2), warp 1 MMA issuer (leader only; tcgen05.commit
// multicasts "stage consumed" / "accumulator ready" to both CTAs), warp 2 TMEM allocator
// (cta_group::2, both CTAs), warps 4-19 saturation scan of this CTA's 128 rows (four sets of four
// warps, a set per pipeline turn), warps 20-27 epilogue of this CTA's 128 rows × BN columns out of its
// own tensor memory.
//
// Instruction issue, not the tensor pipe or shared memory, was what the single-CTA kernel ran out of on
// long streams (ncu: ≈ 80 k warp instructions per 128×256 tile against the 32 k issue slots its 8192
// tensor-pipe cycles offer).  Hence, here: (1) saturation events go into per-(16-column chunk, row)
// cells, so that the epilogue applies at most two 16-way selects per chunk instead of walking the
// row's whole event list for every chunk; (2) the per-element tail runs on packed FFMA2
// (finish_chunk_fast); (3) one epilogue warp polls the tile barriers, the others sleep in a hardware
// barrier; (4) the scan folds the range check into dp4a's accumulator and ORs the results; (5) the scan
// is latency-bound (dependent shared-memory loads), so it gets 16 warps and nothing else to do — the
// epilogue warps, idle most of a tile, stage its entry lists, zero its cells and file its events.
//
// The scan reads a stage AFTER the MMA has consumed it (it waits for the commit, not for TMA): only the
// leader's full barrier sees the TMA bytes, the commit is visible in both CTAs, and the scan's result is
// only needed at the end of the tile anyway.  A stage returns to the producer when the four scan warps
// that own it have arrived.  (Measured alternative, B200: scan concurrent with the MMA, the leader
// relaying "stage landed" to the peer — 102 us instead of 94 us per 16384-frame hidden layer.)

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
constexpr int kScanSets = 4;         // scan warps come in sets of 4 (one lane group per tile row); set s takes pipeline turns ≡ s (mod 4)
constexpr int kScanWarps = 4 * kScanSets;
constexpr int kEpilogueWarps = 8;    // four TMEM lane quarters × two column halves
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kFirstScanWarp = 4, kFirstEpilogueWarp = kFirstScanWarp + kScanWarps;
constexpr int kThreads = (kFirstEpilogueWarp + kEpilogueWarps) * 32;  // 896
constexpr int kAccStages = 2;
constexpr int kEntCap = 512;    // risk entries of the tile staged (packed) in shared memory; the rest is read from global
constexpr int kPtrSlots = 132;  // ≥ k_blocks + 1 → K ≤ 16768
constexpr int kRowEvents = 4;    // overflow events kept per tile row (third and later event of one 16-column chunk)
constexpr int kCellSlots = 1;    // events a (chunk, row) cell holds (1: 16 KB of cells per CTA instead of 32 — the difference is a sixth pipeline stage)
constexpr int kListCap = 32;     // events a scan warp collects per tile; the epilogue warps file them into the cells

static_assert(kBlockK == kFixKBlock, "risk-list order is tied to the tiling");

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// File one saturation event of tile row `row`, column `node_local` (clamp(v) − v = d) for the epilogue:
// d (20 bits) | column within its 16-column chunk << 20 goes into the first free slot of the (chunk, row)
// cell; a third event of one cell goes to the row's overflow list (count << 16 | chunk mask in flag_s).
template <int BN>
__device__ __forceinline__ void file_event(uint32_t *cell, uint32_t *flag_s, uint32_t *ev_s, int row, int d, uint32_t node_local) {
  constexpr int kAllChunks = BN / 16;
  const uint32_t word = (uint32_t(d) & 0xfffffu) | ((node_local & 15u) << 20);
  uint32_t *c0 = cell + (node_local >> 4) * kBlockM + row;
  bool filed = atomicCAS(c0, 0u, word) == 0u;
  if (kCellSlots > 1 && !filed) filed = atomicCAS(c0 + kAllChunks * kBlockM, 0u, word) == 0u;
  if (!filed) {
    const uint32_t slot = atomicAdd(flag_s + row, 0x10000u) >> 16;
    if (slot >= 0x8000u) atomicSub(flag_s + row, 0x10000u);  // dense risk lists: the count must not wrap (such a row recomputes anyway)
    atomicOr(flag_s + row, 1u << (node_local >> 4));
    if (slot < uint32_t(kRowEvents)) ev_s[row * kRowEvents + slot] = (node_local << 24) | (uint32_t(d) & 0xffffffu);
  }
}

template <int BN>
struct PairConfig {
  static constexpr int kABytes = kBlockM * kBlockK;     // this CTA's 128 activation rows
  static constexpr int kBBytes = (BN / 2) * kBlockK;    // this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 6 : (BN == 128 ? 8 : 9);
  static constexpr int kAllChunks = BN / 16;
  static constexpr int kCellWords = kCellSlots * kAllChunks * kBlockM;  // per accumulator stage
  static constexpr int kCellBytes = kAccStages * kCellWords * 4;
  static constexpr int kListBytes = kAccStages * (kScanWarps * kListCap + kScanWarps) * 4;  // lists + their lengths
  static constexpr int kTmemCols = kAccStages * BN;
  static constexpr int kColsPerWarp = BN / (kEpilogueWarps / 4);
  static constexpr int kChunks = kColsPerWarp / 16;
  static constexpr int kBiasBytes = kAccStages * BN * 4;
  static constexpr int kEntBytes = kAccStages * kEntCap * 4;
  static constexpr int kPtrBytes = kAccStages * kPtrSlots * 4;
  static constexpr int kEvBytes = kAccStages * kBlockM * kRowEvents * 4;
  static constexpr int kCntBytes = kAccStages * kBlockM * 4;
  static constexpr int kBarBytes = (3 * kStages + 4 * kAccStages) * 8 + 16;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kBiasBytes + kLut2Padded + kEntBytes + kPtrBytes + kEvBytes + kCntBytes + kCellBytes + kListBytes + kBarBytes;
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
  uint32_t *s_cell = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_rowcnt) + Cfg::kCntBytes);  // [acc][slot][chunk][row]
  uint32_t *s_list = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_cell) + Cfg::kCellBytes);  // [acc][scan warp][kListCap], then [acc][scan warp] lengths
  uint32_t *s_listn = s_list + kAccStages * kScanWarps * kListCap;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_list) + Cfg::kListBytes);  // leader's copy is the live one
  uint64_t *done_bar = full_bar + Cfg::kStages;    // MMA has consumed the stage (commit, both CTAs)
  uint64_t *empty_bar = done_bar + Cfg::kStages;   // this CTA's scan has released the stage
  uint64_t *tmem_full_bar = empty_bar + Cfg::kStages;
  uint64_t *prep_bar = tmem_full_bar + kAccStages;        // this CTA's epilogue has set up the stage's scan state for its next tile
  uint64_t *acc_free_bar = prep_bar + kAccStages;         // both CTAs' epilogues have drained the stage (leader's copy, for the MMA)
  uint64_t *scan_done_bar = acc_free_bar + kAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(scan_done_bar + kAccStages);

  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  auto tstamp = [&](int slot) { stamp(args.timeline, slot); };
  // Stage-level stamps of the first pair (profiling aid, tools/pair_stage_timeline.py): event `ev` of this CTA's g-th pipeline turn
  // goes to timeline[2048 + ((cta · 8 + ev) · 256 + g)], beyond the per-CTA slots the grid uses; SM clocks, so only stamps of one
  // CTA compare.  Events: 0 producer starts waiting for the slot, 1 its TMA boxes are issued, 2 MMA warp sees the stage full
  // (leader), 3 its MMAs and the commit are issued (leader), 4 a scan set sees the commit, 5 the set has released the stage.
  // Compiled in only with -DFDNN_STAGE_STAMPS (make NVFLAGS_EXTRA=-DFDNN_STAGE_STAMPS): the MMA-issuing thread is bound by its own
  // instruction stream, and even stamps that are switched off cost it a compare and a branch per stage.
  auto sstamp = [&](int ev, uint32_t g) {
#ifdef FDNN_STAGE_STAMPS
    if (args.timeline != nullptr && blockIdx.x < 2 && g < 256u) args.timeline[2048 + ((size_t(blockIdx.x) * 8 + size_t(ev)) * 256 + g)] = clock64();
#else
    (void) ev;
    (void) g;
#endif
  };
  if (threadIdx.x == 0) tstamp(0);
  const int M = args.M, N = args.N, K = args.K;  // rows [args.row0, M) of the activation / output buffers
  const int row0 = args.row0;
  const int m_pairs = (M - row0 + 2 * kBlockM - 1) / (2 * kBlockM);
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
      ptx::mbar_init(prep_bar + i, kEpilogueWarps);
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
  if (threadIdx.x == 0) tstamp(1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): my 128 activation rows + my half of the weight rows =====
    // TWO lanes issue, one box each: a thread gets a 128-row box accepted only every ≈ 225 ns (71 GB/s per SM from one lane,
    // 101 GB/s from two, tools/feed_bench.cu) — a single issuing lane, not the L2, was what capped the operand feed.
    if (lane < 2) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;
      for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
        int m_blk, n_blk;
        decode(ct, m_blk, n_blk);
        int kb = first_k_block(ct);
        for (int i = 0; i < k_blocks; ++i, ++g, kb = (kb + 1 == k_blocks ? 0 : kb + 1)) {
          if (lane == 0) sstamp(0, g);
          ptx::mbar_wait_parked(empty_bar + stage, phase ^ 1);
          uint8_t *sa = tiles + stage * Cfg::kStageBytes;
          uint8_t *sb = sa + Cfg::kABytes;
          if (leader && lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);  // both CTAs' bytes land on this barrier
          const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(full_bar + stage), 0);
          // one instruction, two boxes: per-lane tensor map, destination and row coordinate
          ptx::tma_load_2d_pair(lane == 0 ? &tmap_act : &tmap_w, bar, lane == 0 ? sa : sb, kb * kBlockK,
                                lane == 0 ? row0 + m_blk * kBlockM : n_blk * BN + int(rank) * (BN / 2));
          if (lane == 0) sstamp(1, g);
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
      uint32_t g = 0;
      for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
        ptx::mbar_wait_cluster(acc_free_bar + acc, acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb, ++g) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after_sync();
          if (lane == 0) {
            sstamp(2, g);
            if (kb == 0) tstamp(2);
            const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::kStageBytes);
            const uint64_t da = ptx::smem_desc_k_sw128(a_addr), db = ptx::smem_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::mma_i8_ss_pair(d_tmem, da + uint64_t(k * (kUmmaK / 16)), db + uint64_t(k * (kUmmaK / 16)), idesc, uint32_t((kb | k) != 0));
            }
            ptx::mma_commit_pair(done_bar + stage);
            sstamp(3, g);
            if (kb == k_blocks - 1) {
              ptx::mma_commit_pair(tmem_full_bar + acc);
              tstamp(3);
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
    // ===== saturation scan of this CTA's rows =====
    // Thread ↔ tile rows, straight from the 128B-swizzled A tile TMA staged for the tensor core (row r,
    // byte b of the 128-byte K block lives at r·128 + ((b/16 ^ r%8)·16 + b%16)).  One shared-memory
    // instruction covers 8 rows × 4 entries (lane = 8·entry + row%8): the swizzle puts the same byte
    // offset of 8 consecutive rows into 8 different 16-byte chunks, and the packer orders entries so that
    // 4 consecutive ones differ in their word offset within the chunk — 32 lanes, 32 banks.
    // The scan warps do nothing but this loop: the per-tile set-up (entry staging, zeroed event cells) and
    // the filing of the events they find are the epilogue warps' job, so no stage waits for either.
    const int st = int(threadIdx.x) - kFirstScanWarp * 32;
    const int sw = warp - kFirstScanWarp;
    const int sset = st / kBlockM;
    const int row_sub = lane & 7, ent_sub = lane >> 3;
    const int row_base = ((st % kBlockM) / 32) * 32 + row_sub;  // rows row_base + 8j, j < 4
    const uint32_t swz = uint32_t(row_sub) << 4;
    const int kbn = args.fix.k_blocks;
    const uint32_t lanes_below = (1u << lane) - 1u;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t it = 0;  // running K-block count across tiles: stage = it % kStages, phase = (it / kStages) & 1
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      decode(ct, m_blk, n_blk);
      const int kb0 = first_k_block(ct);
      auto k_block_of = [&](int turn) { return (kb0 + turn) % k_blocks; };
      const uint32_t *P = s_ptr + acc * kPtrSlots;  // K-block offsets of this tile's entries, relative to its first one
      const uint32_t *E = s_ent + acc * kEntCap;    // one word per entry: w0 | w1 << 8 | (node − n0) << 16 | byte offset of the pair in its K block << 24
      uint32_t *cell = s_cell + acc * Cfg::kCellWords;
      uint32_t *flag_s = s_rowcnt + acc * kBlockM;
      uint32_t *ev_s = s_rowev + acc * kBlockM * kRowEvents;
      uint32_t *list = s_list + (acc * kScanWarps + sw) * kListCap;
      ptx::mbar_wait_parked(prep_bar + acc, acc_phase);  // the epilogue warps have staged P and E and zeroed the cells
      const uint32_t staged = min(P[kbn], uint32_t(kEntCap));
      uint32_t n_list = 0;  // warp-uniform
      // Events (pair sum left the int16 range; ≈ 1 per 150 evaluations on the synthetic network, i.e. in every
      // other pass of the loop below) are appended to this warp's list with a ballot and a plain store.
      // warp-collective: lanes with `f` append (v32 = pair sum + 32768)
      auto push = [&](bool f, int row, int v32, uint32_t node_local) {
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (m != 0u) {
          if (f) {
            const int v = v32 - 32768;
            const int d = max(min(v, 32767), -32768) - v;
            const uint32_t pos = n_list + uint32_t(__popc(m & lanes_below));
            if (pos < uint32_t(kListCap))
              list[pos] = (uint32_t(d) & 0x1ffffu) | (node_local << 17) | (uint32_t(row) << 25);
            else
              file_event<BN>(cell, flag_s, ev_s, row, d, node_local);  // dense risk lists
          }
          n_list += uint32_t(__popc(m));
        }
      };
      // Entry words are fetched one turn ahead: they do not depend on the data of the stage.
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
        // one warp of the set polls the stage's barrier, the other three wait for it in a hardware barrier: a
        // set idles three turns out of four, and sixteen polling warps would take a third of the issue slots
        if ((sw & 3) == 0) {
          ptx::mbar_wait_parked(done_bar + stage, (g / uint32_t(Cfg::kStages)) & 1u);
          if (lane == 0) sstamp(4, g);
        }
        ptx::named_bar_sync(2 + uint32_t(sset), 128);
        // rows are 128-byte aligned: address = row base | (entry's byte offset with its 16-byte chunk index XORed by row%8)
        const uint32_t a_rows = ptx::smem_u32(tiles + stage * Cfg::kStageBytes) + uint32_t(row_base) * 128u;
        const uint32_t fast_end = min(r1, staged);
        for (uint32_t e = r0; e < fast_end; e += 8) {
          if (e != r0) fetch(e, fast_end);
          const uint32_t x0 = a_rows | ((w0 >> 24) ^ swz), x1 = a_rows | ((w1 >> 24) ^ swz);
          const uint32_t a00 = ptx::lds_u16_off<0>(x0), a01 = ptx::lds_u16_off<1024>(x0), a02 = ptx::lds_u16_off<2048>(x0), a03 = ptx::lds_u16_off<3072>(x0);
          const uint32_t a10 = ptx::lds_u16_off<0>(x1), a11 = ptx::lds_u16_off<1024>(x1), a12 = ptx::lds_u16_off<2048>(x1), a13 = ptx::lds_u16_off<3072>(x1);
          // pair sum + 32768 ∈ [0, 65535] unless pmaddubsw would have saturated: OR them all, look closer only if a high bit is set
          const int v00 = dp4a_u8s8(a00, w0, 32768), v01 = dp4a_u8s8(a01, w0, 32768), v02 = dp4a_u8s8(a02, w0, 32768), v03 = dp4a_u8s8(a03, w0, 32768);
          const int v10 = dp4a_u8s8(a10, w1, 32768), v11 = dp4a_u8s8(a11, w1, 32768), v12 = dp4a_u8s8(a12, w1, 32768), v13 = dp4a_u8s8(a13, w1, 32768);
          const uint32_t fired = uint32_t(v00 | v01 | v02 | v03 | v10 | v11 | v12 | v13) >> 16;
          if (__any_sync(0xffffffffu, fired != 0u)) {
            const bool ok0 = e + uint32_t(ent_sub) < fast_end, ok1 = e + 4u + uint32_t(ent_sub) < fast_end;
            const uint32_t n0l = (w0 >> 16) & 0xffu, n1l = (w1 >> 16) & 0xffu;
            push(ok0 && uint32_t(v00) > 65535u, row_base, v00, n0l);
            push(ok0 && uint32_t(v01) > 65535u, row_base + 8, v01, n0l);
            push(ok0 && uint32_t(v02) > 65535u, row_base + 16, v02, n0l);
            push(ok0 && uint32_t(v03) > 65535u, row_base + 24, v03, n0l);
            push(ok1 && uint32_t(v10) > 65535u, row_base, v10, n1l);
            push(ok1 && uint32_t(v11) > 65535u, row_base + 8, v11, n1l);
            push(ok1 && uint32_t(v12) > 65535u, row_base + 16, v12, n1l);
            push(ok1 && uint32_t(v13) > 65535u, row_base + 24, v13, n1l);
          }
        }
        // beyond the staging capacity (dense risk lists): one entry per pass, 32 rows per warp
        if (r1 > staged) {
          const uint2 *gent = reinterpret_cast<const uint2 *>(args.fix.ent) + __ldg(args.fix.ptr + size_t(n_blk) * kbn);
          for (uint32_t e = max(r0, staged); e < r1; ++e) {
            const uint2 fe = __ldg(gent + e);
            const uint32_t b = (2u * (fe.x & 0xffffu)) & 127u;
            const int row = (st % kBlockM);
            const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::kStageBytes) + uint32_t(row) * 128u;
            const uint32_t a01s = ptx::lds_u16(a_addr + (((b & 0x70u) ^ (uint32_t(row & 7) << 4)) | (b & 15u)));
            const int v32 = dp4a_u8s8(a01s, fe.x >> 16, 32768);
            push(uint32_t(v32) > 65535u, row, v32, fe.y - uint32_t(n_blk * BN));
          }
        }
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(empty_bar + stage);
          if ((sw & 3) == 3) sstamp(5, g);
        }
        if (kb + kScanSets < k_blocks) {
          r0 = P[k_block_of(kb + kScanSets)];
          r1 = P[k_block_of(kb + kScanSets) + 1];
          fetch(r0, min(r1, staged));
        }
      }
      it += uint32_t(k_blocks);
      __syncwarp();
      if (lane == 0) {
        s_listn[acc * kScanWarps + sw] = min(n_list, uint32_t(kListCap));
        ptx::mbar_arrive(scan_done_bar + acc);
      }
      if (st == 0) tstamp(4);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= kFirstEpilogueWarp) {
    // ===== epilogue: this CTA's TMEM → registers → + saturation events → reference tail → global =====
    // 8 warps: warp % 4 selects the TMEM lane quarter (rows), (warp − first) / 4 the column half.  Around the
    // tile's drain they also do the scan warps' housekeeping: file the events the scan found into the
    // (16-column chunk, row) cells, and set up the accumulator stage's scan state for the tile after next.
    const int et = int(threadIdx.x) - kFirstEpilogueWarp * 32;
    const int quarter = warp & 3;
    const int col_group = (warp - kFirstEpilogueWarp) >> 2;
    const int row_local = quarter * 32 + lane;
    const int kbn = args.fix.k_blocks;
    if (!kLogits) {
      for (int i = et; i < kLut2Padded / 16; i += kEpilogueThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(args.lut) + i);
    }
    const uint32_t leader_acc_free = ptx::mapa_u32(ptx::smem_u32(acc_free_bar), 0);
    // scan state of accumulator stage `a` for tile `ct2`: entry offsets per K block, packed entry words, zeroed cells and flags
    auto prepare = [&](int a, int ct2) {
      if (ct2 < tiles_total) {
        int m_blk2, n_blk2;
        decode(ct2, m_blk2, n_blk2);
        const bool scan_on = !(args.debug_flags & 1);
        const uint32_t *gp = args.fix.ptr + size_t(n_blk2) * kbn;
        const uint32_t ent_begin = __ldg(gp);
        uint32_t *P = s_ptr + a * kPtrSlots;
        uint32_t *E = s_ent + a * kEntCap;
        for (int i = et; i <= kbn; i += kEpilogueThreads) P[i] = scan_on ? __ldg(gp + i) - ent_begin : 0u;
        const uint32_t n_ent = scan_on ? __ldg(gp + kbn) - ent_begin : 0u;
        const uint32_t staged = min(n_ent, uint32_t(kEntCap));
        const uint2 *gent = reinterpret_cast<const uint2 *>(args.fix.ent) + ent_begin;
        for (uint32_t e = uint32_t(et); e < staged; e += kEpilogueThreads) {
          const uint2 fe = __ldg(gent + e);
          E[e] = (fe.x >> 16) | ((fe.y - uint32_t(n_blk2 * BN)) << 16) | (((2u * (fe.x & 0xffffu)) & 127u) << 24);
        }
        uint4 *cz = reinterpret_cast<uint4 *>(s_cell + a * Cfg::kCellWords);
        for (int i = et; i < Cfg::kCellWords / 4; i += kEpilogueThreads) cz[i] = make_uint4(0u, 0u, 0u, 0u);
        if (et < kBlockM) s_rowcnt[a * kBlockM + et] = 0;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(prep_bar + a);
    };
    prepare(0, first_ct);
    prepare(1, first_ct + ct_step);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      decode(ct, m_blk, n_blk);
      const int n0 = n_blk * BN;
      const int row = row0 + m_blk * kBlockM + row_local;
      const bool row_ok = row < M;
      const int col0 = n0 + col_group * Cfg::kColsPerWarp;
      const int n_valid = max(0, min(Cfg::kChunks, (N - col0 + 15) / 16));  // warp-uniform

      float *bias_s = s_bias + acc * BN;
      for (int i = et; i < BN; i += kEpilogueThreads) bias_s[i] = (n0 + i < N) ? __ldg(args.bias + n0 + i) : 0.0f;
      // one warp polls the tile's barriers; the others wait for it in the hardware barrier below, which
      // costs no issue slots (and also publishes the bias tile)
      if (warp == kFirstEpilogueWarp) {
        ptx::mbar_wait_parked(tmem_full_bar + acc, acc_phase);
        if (et == 0) tstamp(5);
        ptx::mbar_wait_parked(scan_done_bar + acc, acc_phase);
        if (et == 0) tstamp(7);
        ptx::tc_fence_after_sync();
        ptx::tc_fence_before_sync();
      }
      ptx::named_bar_sync(1, kEpilogueThreads);
      ptx::tc_fence_after_sync();
      // file the scan warps' events: 16 threads per list
      uint32_t *cell_all = s_cell + acc * Cfg::kCellWords;
      uint32_t *flag_s = s_rowcnt + acc * kBlockM;
      uint32_t *ev_s = s_rowev + acc * kBlockM * kRowEvents;
      {
        constexpr int kPerList = kEpilogueThreads / kScanWarps;
        const int lw = et / kPerList;
        const uint32_t n_l = s_listn[acc * kScanWarps + lw];
        const uint32_t *list = s_list + (acc * kScanWarps + lw) * kListCap;
        for (uint32_t i = uint32_t(et % kPerList); i < n_l; i += uint32_t(kPerList)) {
          const uint32_t w = list[i];
          file_event<BN>(cell_all, flag_s, ev_s, int(w >> 25), int(w << 15) >> 15, (w >> 17) & 0xffu);
        }
      }
      ptx::named_bar_sync(1, kEpilogueThreads);
      const uint32_t flags = flag_s[row_local];
      const bool many = (flags >> 16) > uint32_t(kRowEvents);  // more overflow events than slots: this row recomputes from global memory
      const uint32_t *ev = ev_s + row_local * kRowEvents;
      const uint32_t *cell = cell_all + row_local;
      const uint32_t t_addr = tmem_base + uint32_t(acc * BN + col_group * Cfg::kColsPerWarp) + (uint32_t(quarter * 32) << 16);
      auto release_acc = [&]() {  // accumulator stage fully read by this warp
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote_relaxed(leader_acc_free + uint32_t(acc) * 8u);
      };
      if (n_valid == 0) release_acc();
#pragma unroll
      for (int j = 0; j < Cfg::kChunks; ++j) {
        if (j < n_valid) {
          const int chunk = col_group * Cfg::kChunks + j;
          uint32_t e0 = cell[chunk * kBlockM], e1 = kCellSlots > 1 ? cell[(Cfg::kAllChunks + chunk) * kBlockM] : 0u;
          uint32_t raw[16];
          ptx::tmem_ld_32x16(t_addr + uint32_t(j * 16), raw);
          ptx::tmem_ld_wait();
          int32_t s[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i] = int32_t(raw[i]);
          const int col = col0 + j * 16;
          if (many || !row_ok) e0 = e1 = 0u;
          // an empty cell is the zero word: it adds 0 to column 0
          auto apply = [&](uint32_t e) {
            const uint32_t rel = (e >> 20) & 15u;
            const int d = int(e << 12) >> 12;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (rel == uint32_t(i)) s[i] += d;
          };
          if (__any_sync(0xffffffffu, e0 != 0u)) apply(e0);
          if (__any_sync(0xffffffffu, e1 != 0u)) apply(e1);
          if (row_ok && ((flags >> chunk) & 1u) != 0u) {
            if (!many) {
              const uint32_t n_ev = flags >> 16;
              for (uint32_t k = 0; k < n_ev; ++k) {
                const uint32_t e = ev[k];
                const uint32_t rel = (e >> 24) - uint32_t(col - n0);
                const int d = int(e << 8) >> 8;
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] += (rel == uint32_t(i)) ? d : 0;
              }
            }
          }
          if (many && row_ok) brute_force_corrections(s, row, col, args);
          if (j == n_valid - 1) release_acc();  // before the math: the MMA of tile i+2 can start
          if constexpr (kLogits) {  // every lane: the quad-transposed store needs the whole warp
            if (args.fast_tail)
              finish_chunk_fast<true, true>(s, row, col, args, bias_s + (col - n0), s_lut, row_ok, M);
            else
              finish_chunk<true, true>(s, row, col, args, bias_s + (col - n0), s_lut, row_ok, M);
          } else if (row_ok) {
            if (args.fast_tail)
              finish_chunk_fast<false>(s, row, col, args, bias_s + (col - n0), s_lut);
            else
              finish_chunk<false>(s, row, col, args, bias_s + (col - n0), s_lut);
          }
        }
      }
      if (et == 0) tstamp(6);
      // everybody is done with this stage's cells, flags, lists and entry words: set it up for the tile after next
      ptx::named_bar_sync(1, kEpilogueThreads);
      prepare(acc, ct + 2 * ct_step);
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
  const int m_pairs = (a.M - a.row0 + 2 * kBlockM - 1) / (2 * kBlockM), n_blocks = (a.N + BN - 1) / BN;
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
  if (a.M - a.row0 <= 0 || a.row0 % (2 * kBlockM) != 0) return a.M - a.row0 <= 0 ? cudaSuccess : cudaErrorInvalidValue;
  switch (block_n) {
    case 64: return logits ? launch_one<64, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<64, false>(tmap_act, tmap_w, a, num_sms, stream);
    case 128: return logits ? launch_one<128, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<128, false>(tmap_act, tmap_w, a, num_sms, stream);
    case 256: return logits ? launch_one<256, true>(tmap_act, tmap_w, a, num_sms, stream) : launch_one<256, false>(tmap_act, tmap_w, a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace fdnn
