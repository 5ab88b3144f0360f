// Int8 layer on the 5th-generation tensor cores:  u8 activations [M×K] · s8 weights [N×K]ᵀ → s32
// in tensor memory, then the reference's per-element tail in the epilogue.
//
// Replaces, for one whole layer and all frames at once (paths under /root/reference):
//   quantizedNodeSum            src/cpp/dnn.cc:323-349   (the pmaddubsw dot product)
//   QuantizedLayerActivations   src/cpp/dnn.cc:289-318   (sum / (multiplier·255))
//   AddBias + QuantizedSigmoid  src/cpp/dnn.cc:250-286   (hidden mode: + bias, LUT → u8)
//   CalculateOutput's "+= bias" src/cpp/dnn.cc:442-447   (logits mode: fp32 lin + bias)
//
// Both operands are already K-major in the reference's own row-major layouts, so TMA copies
// 128-row × 128-byte boxes straight into 128B-swizzled shared memory and tcgen05.mma kind::i8
// (A unsigned, B signed) consumes them through shared-memory descriptors; no transposes.
//
// Persistent, warp-specialised (28 warps): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 =
// TMEM allocator, warps 4-11 = saturation scan, warps 12-27 = epilogue (four warps per TMEM lane
// quarter, each a quarter of the tile's columns, 16 columns at a time — the tail is ≈20
// instructions per element, so it needs the issue slots of many warps).  Two accumulator stages
// in TMEM let the epilogue of tile i overlap the contraction of tile i+1.
//
// Thread-block clusters: at every batch size the operand tiles come out of L2, and L2→SM
// bandwidth, not the tensor pipe, is what the contraction waits for.  CTAs that need the same
// operand tile therefore form a cluster and fetch it once: each CTA loads 1/C of the shared tile
// and TMA-multicasts it into all C shared memories (kShareA: C neighbouring N tiles share the
// activation tile — small batches; otherwise C neighbouring M tiles share the weight tile —
// streams).  A pipeline stage is then only free when every CTA of the cluster has released it,
// so the MMA commit and the scan warps arrive on the empty barrier of all C CTAs.  Measured on B200
// (profiles/r1_cluster_multicast.md): correct, but slower than independent CTAs at C ≤ 4 — the
// cluster-wide stage hand-off costs more than the multicast saves — so it is opt-in (FDNN_CLUSTER=1).
//
// K rotation: CTAs that share an operand tile start their K loop at different K blocks (integer
// accumulation is order-independent), so they do not all ask L2 for the same lines at the same time.
//
// pmaddubsw's int16 pair saturation is not reproduced by the tensor core.  The scan warps (two
// sets of one thread per tile row, alternating K blocks) walk the layer's risk entries K block by K block, read the two activation
// bytes of each entry from the very A tile that TMA staged for the MMA, and record the rare
// non-zero clamp(v) − v as per-row events in shared memory; the epilogue adds them to the raw sums
// before dequantisation, which makes the sums bit-identical to the reference's.

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace fdnn {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 128;  // bytes of K per pipeline stage = one 128B swizzle atom
constexpr int kUmmaK = 32;    // K per tcgen05.mma for 8-bit operands
constexpr int kScanSets = 2;        // scan warps come in sets of 4 (one thread per tile row); set s takes K blocks ≡ s (mod 2)
constexpr int kScanWarps = 4 * kScanSets;
constexpr int kEpilogueWarps = 16;  // four per TMEM lane quarter, each owning a quarter of the tile's columns
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kFirstScanWarp = 4, kFirstEpilogueWarp = kFirstScanWarp + kScanWarps;
constexpr int kThreads = (kFirstEpilogueWarp + kEpilogueWarps) * 32;  // 896
constexpr int kAccStages = 2;
constexpr int kEntCap = 2048;   // risk entries of the tile staged (packed) in shared memory; the rest is read from global
constexpr int kPtrSlots = 132;  // ≥ k_blocks + 1 → K ≤ 16768
constexpr int kRowEvents = 8;   // saturation events kept per tile row; more → that row recomputes from global memory

static_assert(kBlockK == kFixKBlock && kFixGroups[0] == 64 && kFixGroups[1] == 128 && kFixGroups[2] == 256, "risk-list order is tied to the tiling");

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int BN>
struct TcConfig {
  static constexpr int kABytes = kBlockM * kBlockK;
  static constexpr int kBBytes = BN * kBlockK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = kAccStages * BN;  // 128, 256 or 512: powers of two
  static constexpr int kColsPerWarp = BN / 4;
  static constexpr int kChunks = kColsPerWarp / 16;  // 16-column chunks per epilogue thread
  static constexpr int kBiasBytes = kAccStages * BN * 4;
  static constexpr int kEntBytes = kAccStages * kEntCap * 4;
  static constexpr int kPtrBytes = kAccStages * kPtrSlots * 4;
  static constexpr int kEvBytes = kAccStages * kBlockM * kRowEvents * 4;
  static constexpr int kCntBytes = kAccStages * kBlockM * 4;
  static constexpr int kBarBytes = (2 * kStages + 3 * kAccStages) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBiasBytes + kLut2Padded + kEntBytes + kPtrBytes +
                                    kEvBytes + kCntBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// Which tile of the layer a CTA works on in round `ct` of its cluster.  Tiles past the edge of the
// matrix ("dummy" tiles of a partially filled cluster) run the whole protocol on zero-filled
// operands and store nothing.
template <int BN, int C, bool kShareA>
struct TileMap {
  int m_blocks, n_blocks, groups, total;
  __device__ TileMap(int M, int N) {
    m_blocks = (M + kBlockM - 1) / kBlockM;
    n_blocks = (N + BN - 1) / BN;
    groups = kShareA ? (n_blocks + C - 1) / C : (m_blocks + C - 1) / C;
    total = C == 1 ? m_blocks * n_blocks : (kShareA ? m_blocks * groups : groups * n_blocks);
  }
  __device__ void decode(int ct, int rank, int &m_blk, int &n_blk) const {
    if (C == 1) {
      m_blk = ct / n_blocks;
      n_blk = ct % n_blocks;
    } else if (kShareA) {
      m_blk = ct / groups;
      n_blk = (ct % groups) * C + rank;
    } else {
      m_blk = (ct / n_blocks) * C + rank;
      n_blk = ct % n_blocks;
    }
  }
};

// First K block of a tile's (rotated) K loop: spreads the CTAs that share an activation tile (same
// m_blk) and those that share a weight tile (same n_blk) over different K blocks.  In a cluster the
// shared operand is multicast K block by K block, so its members must agree: no rotation there.
template <int C>
__device__ __forceinline__ int first_k_block(int m_blk, int n_blk, int k_blocks) {
  return C == 1 ? (n_blk + 5 * m_blk) % k_blocks : 0;
}

template <int BN, bool kLogits, int C, bool kShareA>
__global__ void __launch_bounds__(kThreads, 1)
qlayer_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w, const QLayerArgs args) {
  using Cfg = TcConfig<BN>;
  static_assert(C == 1 || C == 2 || C == 4, "cluster size");
  static_assert(kBlockM % C == 0 && BN % C == 0, "operand slices");
  constexpr uint16_t kClusterMask = uint16_t((1u << C) - 1u);
  // 128B-swizzled tiles need 1024-byte alignment.  Dynamic shared memory starts at the base of the
  // CTA's window (this kernel has no static shared memory); keeping the pointers derived from the
  // array itself (no integer round-trip) lets the compiler emit LDS/STS instead of generic accesses.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t *tiles = smem;
  float *s_bias = reinterpret_cast<float *>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint8_t *s_lut = reinterpret_cast<uint8_t *>(s_bias) + Cfg::kBiasBytes;
  uint32_t *s_ent = reinterpret_cast<uint32_t *>(s_lut + kLut2Padded);
  uint32_t *s_ptr = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_ent) + Cfg::kEntBytes);
  uint32_t *s_rowev = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_ptr) + Cfg::kPtrBytes);
  uint32_t *s_rowcnt = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_rowev) + Cfg::kEvBytes);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_rowcnt) + Cfg::kCntBytes);
  uint64_t *empty_bar = full_bar + Cfg::kStages;
  uint64_t *tmem_full_bar = empty_bar + Cfg::kStages;
  uint64_t *tmem_empty_bar = tmem_full_bar + kAccStages;
  uint64_t *scan_done_bar = tmem_empty_bar + kAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(scan_done_bar + kAccStages);

  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  if (threadIdx.x == 0) stamp(args.timeline, 0);
  const int M = args.M, N = args.N, K = args.K;
  const TileMap<BN, C, kShareA> tmap(M, N);
  const int n_blocks = tmap.n_blocks;
  const int tiles_total = tmap.total;
  const int k_blocks = (K + kBlockK - 1) / kBlockK;
  const int rank = C == 1 ? 0 : int(ptx::cluster_ctarank());
  const int first_ct = int(blockIdx.x) / C, ct_step = int(gridDim.x) / C;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_act);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, C * (1 + 4));  // per CTA of the cluster: MMA commit + the 4 warps of the scan set that owns the stage
    }
    for (int i = 0; i < kAccStages; ++i) {
      ptx::mbar_init(tmem_full_bar + i, 1);
      ptx::mbar_init(tmem_empty_bar + i, kEpilogueThreads);
      ptx::mbar_init(scan_done_bar + i, kScanThreads);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (C > 1) ptx::cluster_sync_all();  // peers' barriers must exist before anything is multicast into them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Weights are model constants: the weight boxes of the first tile's first stages go out BEFORE the wait for the previous kernel, so
  // they arrive while this CTA (launched early, programmatic dependent launch) would only sit there; after the wait the same stages
  // need their activation boxes only.  (Not in a cluster: a stage's multicast halves come from different CTAs.)
  const int pre_w = (C == 1 && first_ct < tiles_total) ? min(Cfg::kStages, k_blocks) : 0;  // stages whose weight box is in flight
  if (warp == 0 && lane == 0 && pre_w > 0) {
    int m_blk, n_blk;
    tmap.decode(first_ct, rank, m_blk, n_blk);
    int kb = first_k_block<C>(m_blk, n_blk, k_blocks);
    for (int s = 0; s < pre_w; ++s, kb = (kb + 1 == k_blocks ? 0 : kb + 1)) {
      ptx::mbar_arrive_expect_tx(full_bar + s, Cfg::kStageBytes);
      ptx::tma_load_2d(&tmap_w, full_bar + s, tiles + s * Cfg::kStageBytes + Cfg::kABytes, kb * kBlockK, n_blk * BN);
    }
  }
  // everything above touched only constants and this CTA's own shared/tensor memory; from here on
  // we read the previous kernel's activations and write buffers it may still be reading
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  if (threadIdx.x == 0) stamp(args.timeline, 1);

  if (warp == 0) {
    // ===== TMA producer =====
    // Without clusters TWO lanes issue, one box each (lane 0 the activation box and the barrier, lane 1 the weight box): one thread
    // gets a 128-row box accepted only every ≈ 225 ns — 71 GB/s per SM from one lane, 101 GB/s from two (tools/feed_bench.cu).
    if (lane < (C == 1 ? 2 : 1)) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
        int m_blk, n_blk;
        tmap.decode(ct, rank, m_blk, n_blk);
        int kb = first_k_block<C>(m_blk, n_blk, k_blocks);
        for (int i = 0; i < k_blocks; ++i, kb = (kb + 1 == k_blocks ? 0 : kb + 1)) {
          uint8_t *sa = tiles + stage * Cfg::kStageBytes;
          uint8_t *sb = sa + Cfg::kABytes;
          if (C == 1 && ct == first_ct && i < pre_w) {
            // armed and its weight box issued before the wait for the previous kernel (above): the activation box completes it
            if (lane == 0) ptx::tma_load_2d(&tmap_act, full_bar + stage, sa, kb * kBlockK, m_blk * kBlockM);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (C == 1)
            ptx::mbar_wait(empty_bar + stage, phase ^ 1);
          else
            ptx::mbar_wait_cluster(empty_bar + stage, phase ^ 1);
          if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
          if (C == 1) {
            // (the weight box may land before lane 0's expect_tx: the transaction count dips below zero, the phase cannot complete
            // before lane 0 has arrived)
            // one instruction, two boxes: per-lane tensor map, destination and row coordinate
            ptx::tma_load_2d(lane == 0 ? &tmap_act : &tmap_w, full_bar + stage, lane == 0 ? sa : sb, kb * kBlockK, lane == 0 ? m_blk * kBlockM : n_blk * BN);
          } else if (kShareA) {
            // my quarter (half) of the activation tile, to everybody; my own weight tile
            constexpr int kRows = kBlockM / C;
            ptx::tma_load_2d_multicast(&tmap_act, full_bar + stage, sa + rank * kRows * kBlockK, kb * kBlockK, m_blk * kBlockM + rank * kRows,
                                       kClusterMask);
            ptx::tma_load_2d(&tmap_w, full_bar + stage, sb, kb * kBlockK, n_blk * BN);
          } else {
            constexpr int kRows = BN / C;
            ptx::tma_load_2d(&tmap_act, full_bar + stage, sa, kb * kBlockK, m_blk * kBlockM);
            ptx::tma_load_2d_multicast(&tmap_w, full_bar + stage, sb + rank * kRows * kBlockK, kb * kBlockK, n_blk * BN + rank * kRows,
                                       kClusterMask);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one fixed lane issues and commits) =====
    constexpr uint32_t idesc = ptx::idesc_i8_u8s8(BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      ptx::mbar_wait(tmem_empty_bar + acc, acc_phase ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {  // kb counts pipeline turns here; which K block a turn carries does not matter to the MMA
        ptx::mbar_wait(full_bar + stage, phase);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          if (kb == 0) stamp(args.timeline, 2);
          const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint64_t da = ptx::smem_desc_k_sw128(a_addr), db = ptx::smem_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advancing K inside the swizzle atom = advancing the start address (16-byte units)
            ptx::mma_i8_ss(d_tmem, da + uint64_t(k * (kUmmaK / 16)), db + uint64_t(k * (kUmmaK / 16)), idesc, uint32_t((kb | k) != 0));
          }
          if (C == 1)
            ptx::mma_commit(empty_bar + stage);
          else
            ptx::mma_commit_multicast(empty_bar + stage, kClusterMask);
          if (kb == k_blocks - 1) {
            ptx::mma_commit(tmem_full_bar + acc);
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
  } else if (warp >= kFirstScanWarp && warp < kFirstEpilogueWarp) {
    // ===== saturation scan: thread = tile row, reads its two activation bytes per risk entry =====
    // straight from the 128B-swizzled A tile TMA staged for the tensor core (row r, byte b of the
    // 128-byte K block lives at r·128 + ((b/16 ^ r%8)·16 + b%16)).
    // Work split inside a scan warp: 8 rows × 4 entries per shared-memory instruction (lane = 8·entry
    // + row%8).  The 128B swizzle puts the same byte offset of 8 consecutive rows into 8 different
    // 16-byte chunks, and the packer orders entries so that 4 consecutive ones differ in their word
    // offset within the chunk: 32 lanes, 32 banks.
    const int st = int(threadIdx.x) - kFirstScanWarp * 32;  // 0 .. kScanThreads − 1
    const int sset = st / kBlockM;                           // which K blocks this thread's set takes
    const int row_sub = lane & 7, ent_sub = lane >> 3;
    const int row_base = ((st % kBlockM) / 32) * 32 + row_sub;  // rows row_base + 8j, j < 4
    static_assert(Cfg::kStages % kScanSets == 0, "a pipeline stage must always belong to the same scan set");
    const uint32_t swz = uint32_t(row_sub) << 4;
    const int kbn = args.fix.k_blocks;  // == k_blocks; the list variant is the one grouped by BN nodes
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t it = 0;  // running K-block count across tiles: stage = it % kStages, phase = (it / kStages) & 1
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      tmap.decode(ct, rank, m_blk, n_blk);
      const int kb0 = first_k_block<C>(m_blk, n_blk, k_blocks);
      auto k_block_of = [&](int turn) { return (kb0 + turn) % k_blocks; };  // which K block pipeline turn `turn` carries
      const bool real = n_blk < n_blocks;  // a dummy tile has no risk entries
      const uint32_t *gp = args.fix.ptr + size_t(real ? n_blk : 0) * kbn;
      uint32_t *P = s_ptr + acc * kPtrSlots;  // K-block offsets of this tile's entries, relative to its first one
      uint32_t *E = s_ent + acc * kEntCap;
      // the event slots of this accumulator stage are free once its previous tile has been drained
      ptx::mbar_wait(tmem_empty_bar + acc, acc_phase ^ 1);
      const uint32_t ent_begin = __ldg(gp);
      for (int i = st; i <= kbn; i += kScanThreads) P[i] = (real && !(args.debug_flags & 1)) ? __ldg(gp + i) - ent_begin : 0u;
      const uint32_t n_ent = (real && !(args.debug_flags & 1)) ? __ldg(gp + kbn) - ent_begin : 0u;
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
      // Entry words are fetched one turn ahead: they do not depend on the data TMA is bringing, so
      // after the barrier only activation load → dp4a → range check remains on the critical path.
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
        ptx::mbar_wait(full_bar + stage, (g / uint32_t(Cfg::kStages)) & 1u);
        // rows are 128-byte aligned, so "row base + (offset with its 16-byte chunk index XORed by row%8)"
        // is a single XOR of the entry's byte offset into a pre-swizzled base
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
          // branch-free common case: collect "pair sum left the int16 range" bits, look closer only if any is set
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
        if (lane == 0) {
          if (C == 1) {
            ptx::mbar_arrive(empty_bar + stage);
          } else {
#pragma unroll
            for (int p = 0; p < C; ++p) ptx::mbar_arrive_cluster(empty_bar + stage, uint32_t(p));
          }
        }
        if (kb + kScanSets < k_blocks) {
          r0 = P[k_block_of(kb + kScanSets)];
          r1 = P[k_block_of(kb + kScanSets) + 1];
          fetch(r0, min(r1, staged));
        }
      }
      it += uint32_t(k_blocks);
      ptx::mbar_arrive(scan_done_bar + acc);
      if (st == 0) stamp(args.timeline, 4);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= kFirstEpilogueWarp) {
    // ===== epilogue: TMEM → registers → + saturation events → reference tail → global =====
    // 16 warps: warp % 4 selects the TMEM lane quarter (rows), (warp − 12) / 4 the column quarter.
    const int et = int(threadIdx.x) - kFirstEpilogueWarp * 32;
    const int quarter = warp & 3;
    const int col_group = (warp - kFirstEpilogueWarp) >> 2;
    const int row_local = quarter * 32 + lane;
    // the sigmoid table is only needed here, so fetching it overlaps the first tile's contraction
    // (the per-tile named barrier below orders it before its first use)
    if (!kLogits) {
      for (int i = et; i < kLut2Padded / 16; i += kEpilogueThreads) reinterpret_cast<uint4 *>(s_lut)[i] = __ldg(reinterpret_cast<const uint4 *>(args.lut) + i);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int ct = first_ct; ct < tiles_total; ct += ct_step) {
      int m_blk, n_blk;
      tmap.decode(ct, rank, m_blk, n_blk);
      const int n0 = n_blk * BN;
      const int row = m_blk * kBlockM + row_local;
      const bool row_ok = row < M;
      const int col0 = n0 + col_group * Cfg::kColsPerWarp;  // first column of this thread
      const int n_valid = max(0, min(Cfg::kChunks, (N - col0 + 15) / 16));  // warp-uniform

      float *bias_s = s_bias + acc * BN;
      for (int i = et; i < BN; i += kEpilogueThreads) bias_s[i] = (n0 + i < N) ? __ldg(args.bias + n0 + i) : 0.0f;
      ptx::named_bar_sync(1, kEpilogueThreads);

      ptx::mbar_wait_relaxed(tmem_full_bar + acc, acc_phase);
      if (et == 0) stamp(args.timeline, 5);
      ptx::mbar_wait(scan_done_bar + acc, acc_phase);
      ptx::tc_fence_after_sync();
      if (et == 0) stamp(args.timeline, 7);
      const uint32_t n_ev = s_rowcnt[acc * kBlockM + row_local];
      const uint32_t *ev = s_rowev + (acc * kBlockM + row_local) * kRowEvents;
      const uint32_t t_addr = tmem_base + uint32_t(acc * BN + col_group * Cfg::kColsPerWarp) + (uint32_t(quarter * 32) << 16);
      if (n_valid == 0) {
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(tmem_empty_bar + acc);
      }
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
          if (j == n_valid - 1) {
            // accumulator stage and its event slots fully read: hand both back before doing the math
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(tmem_empty_bar + acc);
          }
          if constexpr (kLogits) {  // every lane: the quad-transposed store needs the whole warp
            finish_chunk<true, true>(s, row, col, args, bias_s + (col - n0), s_lut, row_ok, M);
          } else {
            if (row_ok) finish_chunk<false>(s, row, col, args, bias_s + (col - n0), s_lut);
          }
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
  if (C > 1) ptx::cluster_sync_all();  // nobody leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool kLogits, int C, bool kShareA>
cudaError_t launch_one(const CUtensorMap &ta, const CUtensorMap &tw, const QLayerArgs &a, int num_sms, cudaStream_t stream) {
  using Cfg = TcConfig<BN>;
  const int m_blocks = (a.M + kBlockM - 1) / kBlockM, n_blocks = (a.N + BN - 1) / BN;
  const int cluster_tiles = C == 1 ? m_blocks * n_blocks : (kShareA ? m_blocks * ((n_blocks + C - 1) / C) : ((m_blocks + C - 1) / C) * n_blocks);
  // clusters of 4 cannot use every SM of a GPC; stay below what can be co-resident so the grid is one wave
  const int max_clusters = C == 1 ? num_sms : (C == 2 ? num_sms / 2 : (num_sms * 7 / 8) / 4);
  const int clusters = cluster_tiles < max_clusters ? cluster_tiles : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(clusters * C));
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
  if (C > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = C;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = unsigned(n_attr);
  return cudaLaunchKernelEx(&cfg, qlayer_tc_kernel<BN, kLogits, C, kShareA>, ta, tw, a);
}

template <int BN, int C, bool kShareA>
cudaError_t configure_one() {
  cudaError_t e = cudaFuncSetAttribute(qlayer_tc_kernel<BN, false, C, kShareA>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<BN>::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(qlayer_tc_kernel<BN, true, C, kShareA>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<BN>::kSmemBytes);
}

template <int BN, int C, bool kShareA>
cudaError_t launch_mode(const CUtensorMap &ta, const CUtensorMap &tw, const QLayerArgs &a, bool logits, int num_sms, cudaStream_t stream) {
  return logits ? launch_one<BN, true, C, kShareA>(ta, tw, a, num_sms, stream) : launch_one<BN, false, C, kShareA>(ta, tw, a, num_sms, stream);
}

}  // namespace

// Opt in to the large dynamic shared memory carve-out on the current device (once per device).
cudaError_t qlayer_tc_configure() {
  cudaError_t e = configure_one<64, 1, true>();
  if (e == cudaSuccess) e = configure_one<128, 1, true>();
  if (e == cudaSuccess) e = configure_one<256, 1, true>();
  if (e == cudaSuccess) e = configure_one<64, 4, true>();
  if (e == cudaSuccess) e = configure_one<128, 4, true>();
  if (e == cudaSuccess) e = configure_one<256, 2, false>();
  return e;
}

bool qlayer_tc_supported(int N, int K, bool logits) {
  if (K < kBlockK || K % kBlockK != 0) return false;
  if (K / kBlockK + 1 > kPtrSlots || K / 2 > 65536) return false;
  if (!logits && N % 16 != 0) return false;
  return true;
}

// Tile width and cluster shape for a launch.  Small batches: narrow tiles so that every SM has one,
// optionally (FDNN_CLUSTER=1) four neighbouring N tiles share the activation tile.  Streams: 128×256
// tiles, optionally two neighbouring M tiles share the weight tile.
TcPlan qlayer_tc_plan(int M, int N, bool logits, int num_sms, int policy) {
  static const bool clusters = [] {
    const char *e = std::getenv("FDNN_CLUSTER");
    return e && e[0] == '1';
  }();
  const int m_blocks = (M + kBlockM - 1) / kBlockM;
  TcPlan p{64, 1, true, false};
  // CTA pairs (qlayer_pair.cu) once there is at least a full wave of 256×256 pair tiles: measured on B200
  // (profiles/r1_experiments.md) they beat single-CTA tiles from there on (16384 frames: hidden layer 87 vs
  // 116 us, output layer 389 vs 509 us) and are level or behind below.  FDNN_PAIR=<bn> forces pairs with
  // that tile width for every layer, FDNN_PAIR=0 turns them off (tuning experiments, parity tests).
  {
    const char *e = std::getenv("FDNN_PAIR");
    const int forced = (e && e[0]) ? std::atoi(e) : -1;
    const bool can_pair = num_sms >= 2;
    if (can_pair && (forced == 64 || forced == 128 || forced == 256)) return TcPlan{forced, 2, false, true};
    const int pair_tiles = ((M + 255) / 256) * ((N + 255) / 256);
    static const int logits_min = [] {  // FDNN_PAIR_LOGITS_TILES: from how many pair tiles on the output layer runs on pairs (tuning)
      const char *v = std::getenv("FDNN_PAIR_LOGITS_TILES");
      return (v && v[0]) ? std::atoi(v) : 0;
    }();
    const int min_tiles = (logits && logits_min > 0) ? logits_min : num_sms / 2;
    if (can_pair && forced != 0 && pair_tiles >= min_tiles) return TcPlan{256, 2, false, true};
  }
  if (const char *e = std::getenv("FDNN_FORCE_BN")) {  // tuning experiments
    const int bn = std::atoi(e);
    if ((bn == 64 || bn == 128 || bn == 256) && N >= 4096) {
      p.block_n = bn;
      return p;
    }
  }
  if (const char *e = std::getenv("FDNN_BN_ALL")) {  // tuning experiments: single-CTA tiles of this width for every layer
    const int bn = std::atoi(e);
    if (bn == 64 || bn == 128 || bn == 256) {
      p.block_n = bn;
      return p;
    }
  }
  if (m_blocks * ((N + 255) / 256) >= 2 * num_sms) {
    p.block_n = 256;
    if (clusters && m_blocks % 2 == 0) {
      p.cluster = 2;
      p.share_a = false;
    }
    return p;
  }
  // Below that, measured on B200 (tools/stage_times.py): per-tile time grows faster than the tile
  // (fewer pipeline stages fit), so take the narrowest tile that still finishes in one round; if even
  // 128-wide tiles need a second round, the 8000-wide output layer is better off with one round of
  // 128×256 tiles (batch 512: 30.9 us vs 34.7 us), the 2048-wide hidden layers with two rounds of
  // 128×128 (batch 2048: 27 us vs 37 us).
  auto tiles = [&](int bn) { return m_blocks * ((N + bn - 1) / bn); };
  // Throughput policy (several contexts in flight): 128-wide tiles where one caller alone would take 64-wide ones.
  // Measured on B200, headline network, 4 × 512 frames in flight: 93 us per step instead of 112 us (one caller
  // alone: 157 instead of 144 us) — 64 CTAs per layer instead of 128, 32 MB instead of 48 MB through L2 → SM.
  if (tiles(64) <= num_sms && policy != 1)
    p.block_n = 64;
  else if (tiles(128) <= num_sms)
    p.block_n = 128;
  else
    p.block_n = (logits && tiles(256) <= num_sms) ? 256 : 128;
  if (clusters && p.block_n <= 128 && (N + p.block_n - 1) / p.block_n >= 4) {
    p.cluster = 4;
    p.share_a = true;
  }
  return p;
}

cudaError_t launch_qlayer_tc(const CUtensorMap &tmap_act, const CUtensorMap &tmap_w, const QLayerArgs &a, bool logits, TcPlan plan, int num_sms,
                             cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  if (plan.cluster == 1) {
    switch (plan.block_n) {
      case 64: return launch_mode<64, 1, true>(tmap_act, tmap_w, a, logits, num_sms, stream);
      case 128: return launch_mode<128, 1, true>(tmap_act, tmap_w, a, logits, num_sms, stream);
      case 256: return launch_mode<256, 1, true>(tmap_act, tmap_w, a, logits, num_sms, stream);
    }
  } else if (plan.cluster == 4 && plan.share_a) {
    if (plan.block_n == 64) return launch_mode<64, 4, true>(tmap_act, tmap_w, a, logits, num_sms, stream);
    if (plan.block_n == 128) return launch_mode<128, 4, true>(tmap_act, tmap_w, a, logits, num_sms, stream);
  } else if (plan.cluster == 2 && !plan.share_a && plan.block_n == 256) {
    return launch_mode<256, 2, false>(tmap_act, tmap_w, a, logits, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace fdnn
