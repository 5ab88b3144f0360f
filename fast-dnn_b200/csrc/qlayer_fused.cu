// Every int8 layer of one pass — and the softmax — in ONE persistent kernel, for batches that are a single wave of tiles.
//
// Replaces, for a whole batch (paths under /root/reference/src/cpp): the layer loop of CalculateUntilLastHiddenLayer
// (dnn.cc:413-423: QuantizedLayerActivations + AddBias + QuantizedSigmoid per hidden layer), CalculateOutput (dnn.cc:428-454)
// and SoftMax::apply (dnn.cc:534-544).  The arithmetic per element is that of qlayer_tc.cu (same tcgen05 kind::i8
// contraction, same saturation scan, same verified tail in device_common.cuh) and of softmax_row.cuh: results are
// bit-identical to the layer-by-layer path.
//
// Why.  At batch 512 a 2048×2048 layer is 128 tiles of 128×64 and ≈ 4 us of operand streaming (48 MB from the L2 at 12 TB/s:
// 192 KB in flight per SM over a ≈ 2 us round trip TMA → MMA issue → scan → release; tools/feed_bench.cu reaches 22 TB/s with
// the same bytes in flight when nothing holds a stage), but as a kernel of its own it took 14-15 us: launch ramp, barrier and tensor-memory set-up,
// table loads, the first TMA round trip and the drain were paid seven times per pass (profiles/r1f_summary.md).  Here they are
// paid once: a CTA keeps its tensor memory, barriers, tables and TMA ring across layers, and between layers waits only for the
// DATA it needs — the 128 frames of its tile's row block must have left the previous layer, which is a counter per row block
// (one release-add per finished tile, one acquire-poll by the TMA producer) instead of a grid-wide barrier.  Weight tiles do
// not depend on the previous layer at all, so the producer has the next layer's first weight tiles in flight while it waits.
// Roles run free of each other across layers; the only CTA-wide barriers are at kernel start and before the softmax phase.
//
// Tiles are handed out dynamically, in (layer, row block, column block) order, from one global counter: a tile only ever waits
// for tiles with smaller numbers, every claimed tile belongs to a CTA that is running and works its claims off in order, so the
// kernel makes progress with ANY number of resident CTAs — no co-residency requirement, no cooperative launch, and the fused
// kernels of several contexts can share the GPU (they simply split the SMs).  The producer claims; the other roles learn the
// tile through a small ring in shared memory.
//
// Pipeline geometry changes once, between the hidden layers (tiles 128×BNH) and the output layer (tiles 128×256): the ring of
// shared-memory stages is re-cut there, so the producer first waits until every stage has been released and the MMA issuer
// until both accumulators have been drained.  Barrier phases are tracked per slot (one bit each), so the barrier objects
// themselves live through the change.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "softmax_row.cuh"

namespace fdnn {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 128;  // bytes of K per pipeline stage = one 128B swizzle atom
constexpr int kUmmaK = 32;
constexpr int kScanSets = 2;  // scan warps come in sets of 4 (one thread per tile row); set s takes pipeline turns ≡ s (mod 2)
constexpr int kScanWarps = 4 * kScanSets;
constexpr int kEpilogueWarps = 16;
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kFirstScanWarp = 4, kFirstEpilogueWarp = kFirstScanWarp + kScanWarps;
constexpr int kThreads = (kFirstEpilogueWarp + kEpilogueWarps) * 32;  // 896
constexpr int kAccStages = 2;
constexpr int kEntCap = 2048;
constexpr int kPtrSlots = 132;
constexpr int kRowEvents = 8;
constexpr int kMaxStages = 8;
constexpr int kRingBytes = 192 * 1024;
constexpr int kLogitsBN = 256;
constexpr int kTmemCols = 512;
constexpr int kPrepThreads = 64;  // warps 2 and 3
constexpr int kSched = 4;  // depth of the claimed-tile ring between the producer and the other roles
constexpr int kSoftmaxGroups = 4;                                   // groups of kSoftmaxThreads threads, a row each
constexpr int kSoftmaxRowFloats = kRingBytes / kSoftmaxGroups / 4;  // 12288: widest row the fused softmax takes

static_assert(kBlockK == kFixKBlock, "risk-list order is tied to the tiling");
static_assert(kSoftmaxGroups * kSoftmaxThreads == kThreads, "softmax groups");

// A pipeline stage is TWO 128-byte K blocks of both operands, each operand brought by ONE 3-D TMA box: what a stage costs the
// producer is a fixed ≈ 300-350 ns of barrier and TMA instruction issue whatever it carries (tools/feed_bench.cu: 24 / 32 / 48 KB
// stages → 80 / 107 / 160 GB/s per SM), so the bytes per stage, not the stage count, decide how fast a tile's operands arrive.
constexpr int kKbPerStage = 2;
template <int BN>
struct Geo {
  static constexpr int kATile = kBlockM * kBlockK;   // one K block of the activation rows
  static constexpr int kWTile = BN * kBlockK;        // … of the weight rows
  static constexpr int kABytes = kKbPerStage * kATile;
  static constexpr int kStageBytes = kABytes + kKbPerStage * kWTile;
  static constexpr int kStages = BN == 256 ? 2 : (BN == 128 ? 2 : 4);
  static constexpr int kColsPerWarp = BN / 4;
  static constexpr int kChunks = kColsPerWarp / 16;
  static_assert(kStages * kStageBytes <= kRingBytes && kStages <= kMaxStages && kStages % kScanSets == 0, "ring");
  static_assert(kAccStages * BN <= kTmemCols, "tensor memory");
};

constexpr int kSmemBias = kRingBytes;
constexpr int kSmemLut = kSmemBias + kAccStages * kLogitsBN * 4;
constexpr int kSmemEnt = kSmemLut + kLut2Padded;
constexpr int kSmemPtr = kSmemEnt + kAccStages * kEntCap * 4;
constexpr int kSmemRowEv = kSmemPtr + kAccStages * kPtrSlots * 4;
constexpr int kSmemRowCnt = kSmemRowEv + kAccStages * kBlockM * kRowEvents * 4;
constexpr int kSmemBars = kSmemRowCnt + kAccStages * kBlockM * 4;
constexpr int kNumBars = 2 * kMaxStages + 4 * kAccStages + 2 * kSched;
constexpr int kSmemRed = kSmemBars + kNumBars * 8 + 16 + kSched * 4;
constexpr int kSmemTotal = kSmemRed + kSoftmaxGroups * (kSoftmaxThreads / 32) * 4;
static_assert(kSmemBars % 8 == 0 && kSmemTotal <= 232448, "shared memory budget");

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t *p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy accesses to global memory ↔ async-proxy (TMA) accesses of the same bytes
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// wait until `*counter` ≥ need (bounded: a protocol bug traps instead of hanging the GPU)
__device__ __forceinline__ void wait_counter(const uint32_t *counter, uint32_t need, bool sleep = true) {
  uint32_t polls = 0;
  while (ld_acquire_gpu(counter) < need) {
    if (sleep) __nanosleep(40);
    if (++polls > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// profiling aid (fdnn_ctx_timeline): [layer][CTA][8] nanosecond stamps — 0 row block ready (producer), 1 first stage landed (MMA),
// 2 accumulator ready (epilogue), 3 scan done (epilogue), 4 stores issued, 5 tile released, 6 weights requested, 7 producer enters layer
__device__ __forceinline__ void fstamp(const FusedArgs &p, int layer, int slot) {
  if (p.timeline != nullptr) p.timeline[(size_t(layer) * 1024 + blockIdx.x) * 8 + slot] = global_ns();
}

// first stage (pair of K blocks) of a tile's rotated K loop, in stages: CTAs that share an operand start at different K blocks
__device__ __forceinline__ int first_k_stage(int m_blk, int n_blk, int k_stages) { return (n_blk + 5 * m_blk) % k_stages; }

__device__ __forceinline__ uint32_t bit(uint32_t word, uint32_t i) { return (word >> i) & 1u; }

struct Shared {
  uint8_t *ring;
  float *bias;
  uint8_t *lut;
  uint32_t *ent, *ptr, *rowev, *rowcnt;
  uint64_t *full_bar, *empty_bar, *tmem_full_bar, *tmem_empty_bar, *scan_done_bar, *prep_bar;
  uint64_t *sched_full_bar, *sched_empty_bar;
  uint32_t *tmem_slot;
  int *tile_ring;
  float *red;
};

// progress counters in global memory (FusedArgs::sync): [0, kFusedMaxRowBlocks) finished tiles per row block, then …
constexpr int kSyncExit = kFusedMaxRowBlocks;      // CTAs that have left the kernel (the last one zeroes everything)
constexpr int kSyncTile = kFusedMaxRowBlocks + 1;  // next tile to hand out

struct Tile {
  int layer, m_blk, n_blk;
};
__device__ __forceinline__ Tile decode_tile(const FusedArgs &p, int id) {
  int j = 0;
  while (j + 1 < p.n_layers && id >= int(p.layer[j + 1].tile_begin)) ++j;
  const int rel = id - int(p.layer[j].tile_begin), nb = int(p.layer[j].n_blocks);
  return Tile{j, rel / nb, rel % nb};
}

// consumer side of the claimed-tile ring: every lane of the warp gets the id, `arrive` lanes (one per warp) release the slot
struct Sched {
  uint32_t q = 0, ph = 0;
};
__device__ __forceinline__ int next_tile(const Shared &sh, Sched &sc, bool arrive) {
  ptx::mbar_wait(sh.sched_full_bar + sc.q, bit(sc.ph, sc.q));
  const int id = sh.tile_ring[sc.q];
  __syncwarp();
  if (arrive) ptx::mbar_arrive(sh.sched_empty_bar + sc.q);
  sc.ph ^= 1u << sc.q;
  sc.q = sc.q + 1 == uint32_t(kSched) ? 0u : sc.q + 1;
  return id;
}

// Running state of a role across tiles and layers.
struct Flow {
  uint32_t slot = 0;      // next ring slot (all roles count pipeline turns identically)
  uint32_t turn = 0;      // pipeline turns since the ring was (re)cut: slot = turn % stages, scan set = turn % 2
  uint32_t ring_ph = 0;   // bit s: parity of this role's next wait on slot s
  uint32_t acc = 0;       // next accumulator stage
  uint32_t acc_ph = 0;    // bit a: parity of this role's next wait on accumulator a
};

template <int BN>
__device__ __forceinline__ void advance(Flow &f) {
  f.ring_ph ^= 1u << f.slot;
  f.slot = f.slot + 1 == uint32_t(Geo<BN>::kStages) ? 0u : f.slot + 1;
  ++f.turn;
}
__device__ __forceinline__ void advance_acc(Flow &f) {
  f.acc_ph ^= 1u << f.acc;
  f.acc ^= 1u;
}

// ---- one layer, one role each ---------------------------------------------------------------------------------------------

// The whole producer warp, converged.  The first `pre` stages of a tile are issued by `pre` lanes at once (lane i owns stage i: its
// weight box first — weights do not depend on the previous layer — its activation box once the row block is ready), the remaining
// stages by two lanes (lane 0 the activation box and the barrier, lane 1 the weight box).
template <int BN>
__device__ __forceinline__ void produce_tile(const FusedArgs &p, const Tile &t, const Shared &sh, Flow &f, int lane) {
  using G = Geo<BN>;
  const int j = t.layer, m_blk = t.m_blk, n_blk = t.n_blk;
  const FusedLayer &L = p.layer[j];
  const int k_stages = L.K / (kBlockK * kKbPerStage);
  const CUtensorMap *amap = &p.act[j & 1];
  const CUtensorMap *wmap = &p.w[j];
  const int ks0 = first_k_stage(m_blk, n_blk, k_stages);
  const int pre = min(G::kStages, k_stages);
  if (lane == 0) fstamp(p, j, 7);
  const uint32_t my_slot = (f.slot + uint32_t(lane)) % uint32_t(G::kStages);
  const int my_kb = ((ks0 + lane) % k_stages) * kKbPerStage;
  uint8_t *my_stage = sh.ring + my_slot * G::kStageBytes;
  if (lane < pre) {
    ptx::mbar_wait(sh.empty_bar + my_slot, bit(f.ring_ph, my_slot) ^ 1u);
    // (a weight box may land before … no: the same lane posts the expect_tx first)
    ptx::mbar_arrive_expect_tx(sh.full_bar + my_slot, (p.debug_flags & 64) ? G::kStageBytes - G::kABytes : G::kStageBytes);
    ptx::tma_load_3d(wmap, sh.full_bar + my_slot, my_stage + G::kABytes, 0, n_blk * BN, my_kb);
  }
  if (lane == 0) {
    fstamp(p, j, 6);
    if (j > 0) {
      wait_counter(p.sync + m_blk, L.need, (p.debug_flags & 8) == 0);
      if (!(p.debug_flags & 2)) fence_proxy_async_global();
    }
    fstamp(p, j, 0);
  }
  __syncwarp();
  if (lane < pre) {
    if (j > 0 && !(p.debug_flags & 2)) fence_proxy_async_global();
    if (!(p.debug_flags & 64)) ptx::tma_load_3d(amap, sh.full_bar + my_slot, my_stage, 0, m_blk * kBlockM, my_kb);
  }
  for (int i = 0; i < pre; ++i) advance<BN>(f);
  int ks = (ks0 + pre) % k_stages;
  for (int i = pre; i < k_stages; ++i, ks = (ks + 1 == k_stages ? 0 : ks + 1)) {
    if (lane < 2) {
      ptx::mbar_wait(sh.empty_bar + f.slot, bit(f.ring_ph, f.slot) ^ 1u);
      uint8_t *sa = sh.ring + f.slot * G::kStageBytes;
      if (lane == 0) ptx::mbar_arrive_expect_tx(sh.full_bar + f.slot, (p.debug_flags & 64) ? G::kStageBytes - G::kABytes : G::kStageBytes);
      // one instruction, two boxes: per-lane tensor map, destination and row coordinate (the weight box may land before lane 0's
      // expect_tx is visible: the transaction count dips below zero, the phase cannot complete before lane 0 has arrived)
      if (!(p.debug_flags & 64) || lane == 1)
        ptx::tma_load_3d(lane == 0 ? amap : wmap, sh.full_bar + f.slot, lane == 0 ? sa : sa + G::kABytes, 0, lane == 0 ? m_blk * kBlockM : n_blk * BN,
                         ks * kKbPerStage);
    }
    advance<BN>(f);
  }
}

template <int BN>
__device__ __forceinline__ void mma_tile(const FusedArgs &p, const Tile &t, const Shared &sh, Flow &f, uint32_t tmem_base, int lane) {
  using G = Geo<BN>;
  const int j = t.layer;
  const int k_stages = p.layer[j].K / (kBlockK * kKbPerStage);
  constexpr uint32_t idesc = ptx::idesc_i8_u8s8(BN);
  ptx::mbar_wait(sh.tmem_empty_bar + f.acc, bit(f.acc_ph, f.acc) ^ 1u);
  ptx::tc_fence_after_sync();
  const uint32_t d_tmem = tmem_base + f.acc * uint32_t(BN);
  for (int ks = 0; ks < k_stages; ++ks) {
    ptx::mbar_wait(sh.full_bar + f.slot, bit(f.ring_ph, f.slot));
    ptx::tc_fence_after_sync();
    if (lane == 0) {
      if (ks == 0) fstamp(p, j, 1);
      const uint32_t s_addr = ptx::smem_u32(sh.ring + f.slot * G::kStageBytes);
      if (!(p.debug_flags & 32)) {
#pragma unroll
        for (int kt = 0; kt < kKbPerStage; ++kt) {
          const uint64_t da = ptx::smem_desc_k_sw128(s_addr + uint32_t(kt * G::kATile));
          const uint64_t db = ptx::smem_desc_k_sw128(s_addr + uint32_t(G::kABytes + kt * G::kWTile));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::mma_i8_ss(d_tmem, da + uint64_t(k * (kUmmaK / 16)), db + uint64_t(k * (kUmmaK / 16)), idesc, uint32_t((ks | kt | k) != 0));
        }
      }
      ptx::mma_commit(sh.empty_bar + f.slot);
      if (ks == k_stages - 1) ptx::mma_commit(sh.tmem_full_bar + f.acc);
    }
    __syncwarp();
    advance<BN>(f);
  }
  advance_acc(f);
}

// The per-tile state of the saturation scan, staged by warps 2-3 ahead of the scan warps: entry offsets per K block, packed
// entry words (w0 | w1 << 8 | (node − n0) << 16 | byte offset of the pair in its stage's 256 bytes of K << 24), zeroed event counters.
template <int BN>
__device__ __forceinline__ void prep_tile(const FusedArgs &p, const Tile &t, const Shared &sh, Flow &f, int pt, int lane) {
  const FusedLayer &L = p.layer[t.layer];
  const int kbn = L.K / kBlockK;
  const uint32_t acc = f.acc;
  const uint32_t *gp = L.fix_ptr + size_t(t.n_blk) * kbn;
  uint32_t *P = sh.ptr + acc * kPtrSlots;
  uint32_t *E = sh.ent + acc * kEntCap;
  // the event slots and lists of this accumulator stage are free once its previous tile has been drained
  ptx::mbar_wait(sh.tmem_empty_bar + acc, bit(f.acc_ph, acc) ^ 1u);
  const bool scan_on = !(p.debug_flags & 1);
  const uint32_t ent_begin = __ldg(gp);
  for (int i = pt; i <= kbn; i += kPrepThreads) P[i] = scan_on ? __ldg(gp + i) - ent_begin : 0u;
  const uint32_t n_ent = scan_on ? __ldg(gp + kbn) - ent_begin : 0u;
  const uint32_t staged = min(n_ent, uint32_t(kEntCap));
  const uint2 *gent = reinterpret_cast<const uint2 *>(L.fix_ent) + ent_begin;
  for (uint32_t e = uint32_t(pt); e < staged; e += kPrepThreads) {
    const uint2 fe = __ldg(gent + e);
    // bits 24-30: byte offset of the pair in its K block; bit 31: which of the stage's two K blocks (stages start at even K blocks)
    E[e] = (fe.x >> 16) | ((fe.y - uint32_t(t.n_blk * BN)) << 16) | (((2u * (fe.x & 0xffffu)) & 255u) << 24);
  }
  for (int i = pt; i < kBlockM; i += kPrepThreads) sh.rowcnt[acc * kBlockM + i] = 0;
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(sh.prep_bar + acc);
  advance_acc(f);
}

template <int BN>
__device__ __forceinline__ void scan_tile(const FusedArgs &p, const Tile &t, const Shared &sh, Flow &f, int st, int lane) {
  using G = Geo<BN>;
  const FusedLayer &L = p.layer[t.layer];
  const int k_blocks = L.K / kBlockK, k_stages = k_blocks / kKbPerStage;
  const int sset = st / kBlockM;
  const int row_sub = lane & 7, ent_sub = lane >> 3;
  const int row_base = ((st % kBlockM) / 32) * 32 + row_sub;
  const uint32_t swz = uint32_t(row_sub) << 4;
  const int kbn = k_blocks;
  const uint32_t *fix_ptr = L.fix_ptr;
  const FixEntry *fix_ent = L.fix_ent;
  {
    const int m_blk = t.m_blk, n_blk = t.n_blk;
    const int ks0 = first_k_stage(m_blk, n_blk, k_stages);
    auto k_block_of = [&](int turn) { return ((ks0 + turn) % k_stages) * kKbPerStage; };  // first K block of pipeline turn `turn`
    const uint32_t acc = f.acc;
    const uint32_t *gp = fix_ptr + size_t(n_blk) * kbn;
    uint32_t *P = sh.ptr + acc * kPtrSlots;
    uint32_t *E = sh.ent + acc * kEntCap;
    // the two prep warps have staged this tile's entry offsets (P) and packed entries (E) and zeroed the event counters of
    // this accumulator stage — one or more tiles ahead, so no global-memory latency sits between two tiles of a CTA
    ptx::mbar_wait(sh.prep_bar + acc, bit(f.acc_ph, acc));
    const uint32_t staged = min(P[kbn], uint32_t(kEntCap));
    uint32_t *cnt_s = sh.rowcnt + acc * kBlockM;
    uint32_t *ev_s = sh.rowev + acc * kBlockM * kRowEvents;
    auto record = [&](int row, int v, uint32_t node_local) {
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
    // this set's first pipeline turn (stage) of the tile; every role counts turns identically, so ring slot and scan set follow from it
    int turn = int((uint32_t(sset) + kScanSets - f.turn % kScanSets) % kScanSets);
    uint32_t r0 = 0, r1 = 0;
    if (turn < k_stages) {
      r0 = P[k_block_of(turn)];
      r1 = P[k_block_of(turn) + kKbPerStage];
      fetch(r0, min(r1, staged));
    }
    for (; turn < k_stages; turn += kScanSets) {
      const uint32_t slot = (f.slot + uint32_t(turn)) % uint32_t(G::kStages);
      ptx::mbar_wait(sh.full_bar + slot, bit(f.ring_ph, slot));
      f.ring_ph ^= 1u << slot;
      // activation bytes of the stage: two 128B-swizzled tiles of 128 rows, one per K block (entry word bit 31 says which)
      const uint32_t a_swz = (ptx::smem_u32(sh.ring + slot * G::kStageBytes) + uint32_t(row_base) * 128u) ^ swz;
      const uint32_t fast_end = min(r1, staged);
      for (uint32_t e = r0; e < fast_end; e += 8) {
        if (e != r0) fetch(e, fast_end);
        const uint32_t o0 = (w0 >> 24) & 127u, t0 = (w0 >> 31) * uint32_t(G::kATile), o1 = (w1 >> 24) & 127u, t1 = (w1 >> 31) * uint32_t(G::kATile);
        uint32_t a0[4], a1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a0[q] = ptx::lds_u16(((a_swz + uint32_t(q) * 1024u) ^ o0) + t0);
          a1[q] = ptx::lds_u16(((a_swz + uint32_t(q) * 1024u) ^ o1) + t1);
        }
        int v0[4], v1[4];
        uint32_t fired = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v0[q] = dp4a_u8s8(a0[q], w0, 0);
          v1[q] = dp4a_u8s8(a1[q], w1, 0);
          fired |= (uint32_t(v0[q] + 32768) | uint32_t(v1[q] + 32768)) >> 16;
        }
        if (fired != 0) {
          const bool ok0 = e + uint32_t(ent_sub) < fast_end, ok1 = e + 4u + uint32_t(ent_sub) < fast_end;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ok0 && uint32_t(v0[q] + 32768) > 65535u) record(row_base + 8 * q, v0[q], (w0 >> 16) & 0xffu);
            if (ok1 && uint32_t(v1[q] + 32768) > 65535u) record(row_base + 8 * q, v1[q], (w1 >> 16) & 0xffu);
          }
        }
      }
      for (uint32_t e = max(r0, staged); e < r1; ++e) {  // beyond the staging capacity (dense risk lists)
        const uint2 *gent = reinterpret_cast<const uint2 *>(fix_ent) + __ldg(gp);
        const uint2 fe = __ldg(gent + e);
        const uint32_t b2 = (2u * (fe.x & 0xffffu)) & 255u, b = b2 & 127u;
        const int row = (st % kBlockM);
        const uint32_t a_addr = ptx::smem_u32(sh.ring + slot * G::kStageBytes) + (b2 >> 7) * uint32_t(G::kATile) + uint32_t(row) * 128u;
        const uint32_t a01s = ptx::lds_u16(a_addr + (((b & 0x70u) ^ (uint32_t(row & 7) << 4)) | (b & 15u)));
        const int v = dp4a_u8s8(a01s, fe.x >> 16, 0);
        if (uint32_t(v + 32768) > 65535u) record(row, v, fe.y - uint32_t(n_blk * BN));
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(sh.empty_bar + slot);
      if (turn + kScanSets < k_stages) {
        r0 = P[k_block_of(turn + kScanSets)];
        r1 = P[k_block_of(turn + kScanSets) + kKbPerStage];
        fetch(r0, min(r1, staged));
      }
    }
    // account for the whole tile's turns (the phase bits of this set's own slots were flipped above, one by one)
    f.slot = (f.slot + uint32_t(k_stages)) % uint32_t(G::kStages);
    f.turn += uint32_t(k_stages);
    ptx::mbar_arrive(sh.scan_done_bar + acc);
    advance_acc(f);
  }
}

template <int BN, bool kLogits>
__device__ __forceinline__ void epilogue_tile(const FusedArgs &p, const Tile &t, const Shared &sh, Flow &f, uint32_t tmem_base, int warp, int lane) {
  using G = Geo<BN>;
  const int j = t.layer;
  const FusedLayer &L = p.layer[j];
  const int M = p.M, N = L.N;
  const int et = int(threadIdx.x) - kFirstEpilogueWarp * 32;
  const int quarter = warp & 3;
  const int col_group = (warp - kFirstEpilogueWarp) >> 2;
  const int row_local = quarter * 32 + lane;
  QLayerArgs a{};
  a.act = p.act_buf[j & 1];
  a.bias = L.bias;
  a.lut = p.lut;
  a.coeff = L.coeff;
  a.rcp = L.rcp;
  a.fast_div = L.fast_div;
  a.fast_tail = L.fast_tail;
  a.one = p.one;
  a.neg_zero = p.neg_zero;
  a.M = M;
  a.N = N;
  a.K = L.K;
  a.fix.ptr = L.fix_ptr;
  a.fix.ent = L.fix_ent;
  a.fix.k_blocks = L.K / kBlockK;
  a.fix.group = BN;
  a.out_u8 = p.act_buf[(j + 1) & 1];
  a.out_f32 = p.out;
  a.out_ld = p.out_ld;
  {
    const int m_blk = t.m_blk, n_blk = t.n_blk;
    const int n0 = n_blk * BN;
    const int row = m_blk * kBlockM + row_local;
    const bool row_ok = row < M;
    const int col0 = n0 + col_group * G::kColsPerWarp;
    const int n_valid = max(0, min(G::kChunks, (N - col0 + 15) / 16));  // warp-uniform
    const uint32_t acc = f.acc;

    float *bias_s = sh.bias + acc * kLogitsBN;
    for (int i = et; i < BN; i += kEpilogueThreads) bias_s[i] = (n0 + i < N) ? __ldg(L.bias + n0 + i) : 0.0f;
    ptx::named_bar_sync(1, kEpilogueThreads);

    if (p.debug_flags & 16)
      ptx::mbar_wait(sh.tmem_full_bar + acc, bit(f.acc_ph, acc));
    else
      ptx::mbar_wait_relaxed(sh.tmem_full_bar + acc, bit(f.acc_ph, acc));
    if (et == 0) fstamp(p, j, 2);
    ptx::mbar_wait(sh.scan_done_bar + acc, bit(f.acc_ph, acc));
    ptx::tc_fence_after_sync();
    if (et == 0) fstamp(p, j, 3);
    const uint32_t n_ev = sh.rowcnt[acc * kBlockM + row_local];
    const uint32_t *ev = sh.rowev + (acc * kBlockM + row_local) * kRowEvents;
    const uint32_t t_addr = tmem_base + acc * uint32_t(BN) + uint32_t(col_group * G::kColsPerWarp) + (uint32_t(quarter * 32) << 16);
    if (n_valid == 0) {
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(sh.tmem_empty_bar + acc);
    }
#pragma unroll
    for (int c = 0; c < G::kChunks; ++c) {
      if (c < n_valid) {
        uint32_t raw[16];
        ptx::tmem_ld_32x16(t_addr + uint32_t(c * 16), raw);
        ptx::tmem_ld_wait();
        int32_t s[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = int32_t(raw[i]);
        const int col = col0 + c * 16;
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
            brute_force_corrections(s, row, col, a);
          }
        }
        if (c == n_valid - 1) {
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(sh.tmem_empty_bar + acc);
        }
        if constexpr (kLogits) {
          finish_chunk<true, true>(s, row, col, a, bias_s + (col - n0), sh.lut, row_ok, M);
        } else {
          if (row_ok) {
            if (a.fast_tail)
              finish_chunk_fast<false>(s, row, col, a, bias_s + (col - n0), sh.lut);
            else
              finish_chunk<false>(s, row, col, a, bias_s + (col - n0), sh.lut);
          }
        }
      }
    }
    // tile done: every store of the 16 epilogue warps ordered before ONE release-add on the row block's counter
    if (!(p.debug_flags & 2)) fence_proxy_async_global();
    ptx::named_bar_sync(3, kEpilogueThreads);
    if (et == 0) {
      fstamp(p, j, 4);
      if (p.debug_flags & 4) __threadfence();  // not needed: the release below is cumulative over what the barrier ordered before it
      red_release_gpu_add(p.sync + m_blk, 1u);
      fstamp(p, j, 5);
    }
    advance_acc(f);
  }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------

template <int BNH>
__global__ void __launch_bounds__(kThreads, 1) qlayer_fused_kernel(const __grid_constant__ FusedArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  Shared sh;
  sh.ring = smem;
  sh.bias = reinterpret_cast<float *>(smem + kSmemBias);
  sh.lut = smem + kSmemLut;
  sh.ent = reinterpret_cast<uint32_t *>(smem + kSmemEnt);
  sh.ptr = reinterpret_cast<uint32_t *>(smem + kSmemPtr);
  sh.rowev = reinterpret_cast<uint32_t *>(smem + kSmemRowEv);
  sh.rowcnt = reinterpret_cast<uint32_t *>(smem + kSmemRowCnt);
  sh.full_bar = reinterpret_cast<uint64_t *>(smem + kSmemBars);
  sh.empty_bar = sh.full_bar + kMaxStages;
  sh.tmem_full_bar = sh.empty_bar + kMaxStages;
  sh.tmem_empty_bar = sh.tmem_full_bar + kAccStages;
  sh.scan_done_bar = sh.tmem_empty_bar + kAccStages;
  sh.prep_bar = sh.scan_done_bar + kAccStages;
  sh.sched_full_bar = sh.prep_bar + kAccStages;
  sh.sched_empty_bar = sh.sched_full_bar + kSched;
  sh.tmem_slot = reinterpret_cast<uint32_t *>(sh.sched_empty_bar + kSched);
  sh.tile_ring = reinterpret_cast<int *>(sh.tmem_slot + 4);
  sh.red = reinterpret_cast<float *>(smem + kSmemRed);

  const int warp = int(threadIdx.x) / 32, lane = int(threadIdx.x) % 32;
  const int nl = p.n_layers;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.act[0]);
    ptx::prefetch_tensormap(&p.act[1]);
    for (int j = 0; j < nl; ++j) ptx::prefetch_tensormap(&p.w[j]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      ptx::mbar_init(sh.full_bar + i, 1);
      ptx::mbar_init(sh.empty_bar + i, 1 + 4);  // MMA commit + the 4 warps of the scan set that owns the stage
    }
    for (int i = 0; i < kAccStages; ++i) {
      ptx::mbar_init(sh.tmem_full_bar + i, 1);
      ptx::mbar_init(sh.tmem_empty_bar + i, kEpilogueThreads);
      ptx::mbar_init(sh.scan_done_bar + i, kScanThreads);
      ptx::mbar_init(sh.prep_bar + i, kPrepThreads / 32);
    }
    for (int i = 0; i < kSched; ++i) {
      ptx::mbar_init(sh.sched_full_bar + i, 1);
      ptx::mbar_init(sh.sched_empty_bar + i, 1 + 2 + kScanWarps + kEpilogueWarps);  // one lane of the MMA warp and of every prep / scan / epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(sh.tmem_slot);
  if (warp >= kFirstEpilogueWarp) {
    const int et = int(threadIdx.x) - kFirstEpilogueWarp * 32;
    for (int i = et; i < kLut2Padded / 16; i += kEpilogueThreads) reinterpret_cast<uint4 *>(sh.lut)[i] = __ldg(reinterpret_cast<const uint4 *>(p.lut) + i);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *sh.tmem_slot;
  // everything above touched only constants and this CTA's own shared / tensor memory; from here on we read what the previous
  // kernel wrote (and the counters the previous fused kernel of this context zeroed on its way out)
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();

  Flow f;
  Sched sc;
  bool recut = false;  // the ring has been re-cut for the output layer
  const int last = nl - 1;
  if (warp == 0) {
    {  // the whole warp, converged (produce_tile)
      const int total = int(p.total_tiles);
      // (Claiming one tile ahead, to take the atomic's round trip off the path between two tiles, was measured and dropped: no gain
      // at 512 frames, a loss at 1024 — it changes who gets which tile.  profiles/r2_experiments.md)
      for (;;) {
        int id = 0;
        if (lane == 0) {
          ptx::mbar_wait(sh.sched_empty_bar + sc.q, bit(sc.ph, sc.q) ^ 1u);
          id = int(atomicAdd(p.sync + kSyncTile, 1u));
          if (id >= total) id = -1;
          sh.tile_ring[sc.q] = id;
          ptx::mbar_arrive(sh.sched_full_bar + sc.q);
        }
        id = __shfl_sync(0xffffffffu, id, 0);
        sc.ph ^= 1u << sc.q;
        sc.q = sc.q + 1 == uint32_t(kSched) ? 0u : sc.q + 1;
        if (id < 0) break;
        const Tile t = decode_tile(p, id);
        if (t.layer == last) {
          if (!recut) {  // every stage of the old cut must have been released
            if (lane == 0)
              for (uint32_t s2 = 0; s2 < uint32_t(kMaxStages); ++s2) ptx::mbar_wait(sh.empty_bar + s2, bit(f.ring_ph, s2) ^ 1u);
            __syncwarp();
            f.slot = 0;
            f.turn = 0;
            recut = true;
          }
          produce_tile<kLogitsBN>(p, t, sh, f, lane);
        } else {
          produce_tile<BNH>(p, t, sh, f, lane);
        }
      }
    }
  } else if (warp == 1) {
    for (;;) {
      const int id = next_tile(sh, sc, lane == 0);
      if (id < 0) break;
      const Tile t = decode_tile(p, id);
      if (t.layer == last) {
        if (!recut) {  // both accumulators drained: the output layer's accumulators overlap both of the hidden layers'
          for (uint32_t a = 0; a < uint32_t(kAccStages); ++a) ptx::mbar_wait(sh.tmem_empty_bar + a, bit(f.acc_ph, a) ^ 1u);
          f.slot = 0;
          f.turn = 0;
          recut = true;
        }
        mma_tile<kLogitsBN>(p, t, sh, f, tmem_base, lane);
      } else {
        mma_tile<BNH>(p, t, sh, f, tmem_base, lane);
      }
    }
  } else if (warp == 2 || warp == 3) {
    const int pt = int(threadIdx.x) - 64;
    for (;;) {
      const int id = next_tile(sh, sc, lane == 0);
      if (id < 0) break;
      const Tile t = decode_tile(p, id);
      if (t.layer == last)
        prep_tile<kLogitsBN>(p, t, sh, f, pt, lane);
      else
        prep_tile<BNH>(p, t, sh, f, pt, lane);
    }
  } else if (warp >= kFirstScanWarp && warp < kFirstEpilogueWarp) {
    const int st = int(threadIdx.x) - kFirstScanWarp * 32;
    for (;;) {
      const int id = next_tile(sh, sc, lane == 0);
      if (id < 0) break;
      const Tile t = decode_tile(p, id);
      if (t.layer == last) {
        if (!recut) {
          f.slot = 0;
          f.turn = 0;
          recut = true;
        }
        scan_tile<kLogitsBN>(p, t, sh, f, st, lane);
      } else {
        scan_tile<BNH>(p, t, sh, f, st, lane);
      }
    }
  } else if (warp >= kFirstEpilogueWarp) {
    for (;;) {
      const int id = next_tile(sh, sc, lane == 0);
      if (id < 0) break;
      const Tile t = decode_tile(p, id);
      if (t.layer == last)
        epilogue_tile<kLogitsBN, true>(p, t, sh, f, tmem_base, warp, lane);
      else
        epilogue_tile<BNH, false>(p, t, sh, f, tmem_base, warp, lane);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();  // this CTA's tiles are done: ring, tables and tensor memory are free
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }

  // ---- softmax phase: rows round-robin over (CTA, group); a row is ready when its row block has left the output layer ----
  if (p.do_softmax) {
    const int g = warp / (kSoftmaxThreads / 32), tid = int(threadIdx.x) % kSoftmaxThreads;
    float *s_e = reinterpret_cast<float *>(sh.ring) + size_t(g) * kSoftmaxRowFloats;
    float *s_red = sh.red + g * (kSoftmaxThreads / 32);
    const int O = p.layer[nl - 1].N;
    const bool vec = (O % 4 == 0) && (p.out_ld % 4 == 0);
    const uint32_t done = p.tiles_per_row_block;
    int ready_block = -1;
    for (int r = int(blockIdx.x) * kSoftmaxGroups + g; r < p.M; r += int(gridDim.x) * kSoftmaxGroups) {
      const int m_blk = r / kBlockM;
      if (m_blk != ready_block) {
        if (tid == 0) wait_counter(p.sync + m_blk, done);
        ptx::named_bar_sync(4 + uint32_t(g), kSoftmaxThreads);
        ready_block = m_blk;
      }
      float *x = p.out + size_t(r) * size_t(p.out_ld);
      softmax_row<true>(x, nullptr, x, O, vec, s_e, s_red, tid, [g] { ptx::named_bar_sync(4 + uint32_t(g), kSoftmaxThreads); });
      ptx::named_bar_sync(4 + uint32_t(g), kSoftmaxThreads);  // s_e and s_red are reused by the next row
    }
  }

  // ---- the last CTA out resets the counters for the next launch ----
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m_blocks = (p.M + kBlockM - 1) / kBlockM;
    __threadfence();
    const uint32_t old = atomicAdd(p.sync + kSyncExit, 1u);
    if (old == gridDim.x - 1) {
      for (int i = 0; i < m_blocks; ++i) p.sync[i] = 0u;
      p.sync[kSyncTile] = 0u;
      __threadfence();
      p.sync[kSyncExit] = 0u;
    }
  }
}

template <int BNH>
cudaError_t launch_one(const FusedArgs &a, int grid, cudaStream_t stream) {
  static const bool coop = [] {  // not needed (the tile scheduler is deadlock-free at any residency); timing experiments only
    const char *e = std::getenv("FDNN_FUSED_COOP");
    return e && e[0] == '1';
  }();
  const bool pdl = !coop;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = size_t(kSmemTotal);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (coop) {
    attr[n_attr].id = cudaLaunchAttributeCooperative;
    attr[n_attr].val.cooperative = 1;
    ++n_attr;
  }
  if (pdl && pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = unsigned(n_attr);
  return cudaLaunchKernelEx(&cfg, qlayer_fused_kernel<BNH>, a);
}

}  // namespace

cudaError_t qlayer_fused_configure() {
  cudaError_t e = cudaFuncSetAttribute(qlayer_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qlayer_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
  return e;
}

int qlayer_fused_max_softmax_width() { return kSoftmaxRowFloats; }

// Tile width of the hidden layers and grid for a batch of M frames, or 0 when the batch is not a single wave of tiles (the
// layer-by-layer kernels take it then).  H = hidden width, O = outputs.
int qlayer_fused_plan(int M, int H, int O, int num_sms, int policy, int *grid) {
  const char *env = std::getenv("FDNN_FUSED");  // read per call: tests run the same context both ways
  if ((env && env[0] == '0') || M <= 0) return 0;
  const int m_blocks = (M + kBlockM - 1) / kBlockM;
  if (m_blocks > kFusedMaxRowBlocks) return 0;
  auto tiles = [&](int n, int bn) { return m_blocks * ((n + bn - 1) / bn); };
  // Where it pays (measured on B200, tools/fused_times.py): one caller at a time ("latency" policy; several contexts in flight
  // interleave better kernel by kernel), and layers wide enough that a tile's operand stream, not the hand-over between
  // layers, is what a layer costs — small networks run faster layer by layer with programmatic dependent launch.
  const char *force = std::getenv("FDNN_FUSED");
  const bool forced = force && force[0] == '2';
  int bnh = 0;
  if (tiles(H, 64) <= num_sms)
    bnh = 64;
  else if (tiles(H, 128) <= num_sms)
    bnh = 128;
  if (bnh == 0) return 0;
  if (!forced && (policy == 1 || tiles(H, bnh) < num_sms / 3)) return 0;
  // the output layer may take several rounds of 128×256 tiles on the same CTAs
  *grid = std::min(num_sms, std::max(tiles(H, bnh), tiles(O, kLogitsBN)));
  if (policy == 1) {
    // (FDNN_FUSED=2 only.)  Several contexts in flight: a kernel that takes every SM serialises the passes; with fewer CTAs
    // than tiles the fused kernels of several contexts share the GPU.  Measured on B200 (profiles/r2_experiments.md): 74.5 us per
    // step at best (grid 64, 4 contexts) against 68.5 us layer by layer, so the "throughput" policy stays layer by layer.
    static const int tgrid = [] {
      const char *e = std::getenv("FDNN_FUSED_TGRID");
      const int v = e ? std::atoi(e) : 64;
      return v > 0 ? v : 64;
    }();
    *grid = std::min(*grid, tgrid);
  }
  return bnh;
}

// rows per (CTA, group) slot up to which the fused kernel also normalises the rows; beyond, the stand-alone softmax kernel
// (a CTA per row, many rows per SM) is faster
bool qlayer_fused_softmax_pays(int M, int grid) { return M <= 2 * grid * kSoftmaxGroups; }

cudaError_t launch_qlayer_fused(const FusedArgs &a, int bnh, int grid, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  if (bnh == 64) return launch_one<64>(a, grid, stream);
  if (bnh == 128) return launch_one<128>(a, grid, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fdnn
