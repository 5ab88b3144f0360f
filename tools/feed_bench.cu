// How fast can an SM pull operand tiles out of the L2 into shared memory — by TMA, by cp.async (LDGSTS), or by both at once?
// Every int8 layer kernel of this repository is bounded by that feed, so this probe decides how tiles are staged.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../fast-dnn_b200/csrc -o feed_bench feed_bench.cu -lcuda
//   ./feed_bench
//
// One CTA per SM (optionally fewer) streams 128-row × 128-byte tiles of an L2-resident u8 matrix [rows][2048] into a ring of
// shared-memory stages; a consumer warp releases each stage as soon as it has landed (no math).  Modes per stage of 32 KB:
//   tma      two 16 KB TMA boxes (128B swizzle), completion by complete_tx
//   ldgsts   32 KB by cp.async 16-byte copies from `kCopyWarps` warps, completion by cp.async.mbarrier.arrive
//   mixed    16 KB by TMA + 16 KB by cp.async
// Prints GB/s per SM and in total for each mode and CTA count.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace fdnn;

constexpr int kStages = 6;
constexpr int kStageBytes = 32768;
constexpr int kTileBytes = 16384;  // 128 rows × 128 bytes
constexpr int kCopyWarps = 4;
constexpr int kThreads = 32 * (2 + kCopyWarps);
constexpr int kK = 2048;

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}

// mode: 0 tma, 1 ldgsts, 2 mixed
__global__ void __launch_bounds__(kThreads, 1) feed_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *base, int rows, int turns, int mode,
                                                            unsigned long long *sink, int box_rows = 128, int issuers = 1, int k_tiles_arg = kK / 128) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int tma_tiles = mode == 0 ? 2 : (mode == 2 ? 1 : 0);  // 16 KB tiles per stage brought by TMA
  const int cp_tiles = 2 - tma_tiles;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      // arrivals: the TMA producer's arrive.expect_tx (if any TMA tile) + one per copying thread (noinc arrivals count against the expected number)
      ptx::mbar_init(full_bar + i, (tma_tiles ? 1 : 0) + (cp_tiles ? kCopyWarps * 32 : 0));
      ptx::mbar_init(empty_bar + i, 1);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int row_tiles = rows / 128, k_tiles = k_tiles_arg;
  const int total_tiles = row_tiles * k_tiles;
  if (warp == 0) {
    if (lane < issuers && tma_tiles) {
      // `issuers` lanes share the boxes of a stage (lane 0 arms the barrier): does one thread's issue rate matter?
      int stage = 0;
      uint32_t phase = 0;
      const int box_bytes = box_rows * 128, boxes = tma_tiles * kTileBytes / box_bytes;
      for (int t = 0; t < turns; ++t) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, tma_tiles * kTileBytes);
        __syncwarp((1u << issuers) - 1u);
        for (int j = lane; j < boxes; j += issuers) {
          const int tile = int((uint64_t(blockIdx.x) * 7919u + uint64_t(t) * 8 + j) % uint64_t(total_tiles - 2));
          ptx::tma_load_2d(&tmap, full_bar + stage, smem + stage * kStageBytes + j * box_bytes, (tile % k_tiles) * 128, (tile / k_tiles) * 128);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int t = 0; t < turns; ++t) {
      ptx::mbar_wait(full_bar + stage, phase);
      if (lane == 0) {
        acc += smem[stage * kStageBytes + (t & 1023)];
        ptx::mbar_arrive(empty_bar + stage);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && sink) sink[blockIdx.x] = acc;
  } else if (cp_tiles) {
    const int cw = warp - 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < turns; ++t) {
      ptx::mbar_wait(empty_bar + stage, phase ^ 1);
      for (int j = tma_tiles; j < 2; ++j) {
        const int tile = int((uint64_t(blockIdx.x) * 7919u + uint64_t(t) * 2 + j) % uint64_t(total_tiles));
        const uint8_t *g = base + size_t(tile / k_tiles) * 128 * kK + size_t(tile % k_tiles) * 128;
        const uint32_t s = ptx::smem_u32(smem + stage * kStageBytes + j * kTileBytes);
        // 1024 16-byte pieces per tile: piece p = row·8 + chunk; a warp instruction covers 4 rows × 128 contiguous bytes
        for (int p = cw * 32 + lane; p < 1024; p += kCopyWarps * 32) {
          const int r = p >> 3, c = p & 7;
          cp_async_16(s + uint32_t(r * 128 + ((c ^ (r & 7)) << 4)), g + size_t(r) * kK + c * 16);
        }
      }
      cp_async_arrive_noinc(full_bar + stage);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
}

// TMA only, everything a parameter: rows per box, how many lanes of the producer warp issue boxes, stage size, stages in flight.
__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int turns, int box_rows, int issuers, int stage_bytes,
                                                         int stages, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, 1);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int k_tiles = kK / 128, row_boxes = rows / box_rows, total = row_boxes * k_tiles;
  const int box_bytes = box_rows * 128, boxes = stage_bytes / box_bytes;
  if (warp == 0) {
    if (lane < issuers) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < turns; ++t) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, uint32_t(stage_bytes));
        __syncwarp((1u << issuers) - 1u);
        for (int j = lane; j < boxes; j += issuers) {
          const int tile = int((uint64_t(blockIdx.x) * 7919u + uint64_t(t) * 16 + j) % uint64_t(total));
          ptx::tma_load_2d(&tmap, full_bar + stage, smem + stage * stage_bytes + j * box_bytes, (tile % k_tiles) * 128, (tile / k_tiles) * box_rows);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int t = 0; t < turns; ++t) {
      ptx::mbar_wait(full_bar + stage, phase);
      if (lane == 0) {
        acc += smem[stage * stage_bytes + (t & 1023)];
        ptx::mbar_arrive(empty_bar + stage);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && sink) sink[blockIdx.x] = acc;
  }
}

// The layer kernels' stage: one 128-row box from one tensor map (activations) + one w_rows-row box from ANOTHER map (weights).
// lanes = 1: one lane issues both boxes; lanes = 2: lane 0 the activation box, lane 1 the weight box, in one instruction.
__global__ void __launch_bounds__(96, 1) two_map_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int rows,
                                                        int turns, int w_rows, int lanes, int stages, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int stage_bytes = (128 + w_rows) * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, 1);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int k_tiles = kK / 128;
  if (lanes == 3 && (warp == 0 || warp == 2)) {
    if (lane == 0) {  // two warps, one box each
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < turns; ++t) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        uint8_t *sa = smem + stage * stage_bytes;
        const int ta = int((uint64_t(blockIdx.x % 4) * 16 + uint64_t(t)) % uint64_t(k_tiles));
        const int tw = int((uint64_t(blockIdx.x) * 7919u + uint64_t(t)) % uint64_t((rows / w_rows) * k_tiles));
        if (warp == 0) {
          ptx::mbar_arrive_expect_tx(full_bar + stage, uint32_t(stage_bytes));
          ptx::tma_load_2d(&map_a, full_bar + stage, sa, ta * 128, int(blockIdx.x % 4) * 128);
        } else {
          ptx::tma_load_2d(&map_w, full_bar + stage, sa + 128 * 128, (tw % k_tiles) * 128, (tw / k_tiles) * w_rows);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    if (lane < lanes) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < turns; ++t) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        uint8_t *sa = smem + stage * stage_bytes;
        if (lane == 0) ptx::mbar_arrive_expect_tx(full_bar + stage, uint32_t(stage_bytes));
        const int ta = int((uint64_t(blockIdx.x % 4) * 16 + uint64_t(t)) % uint64_t(k_tiles));         // 4 row blocks shared by many CTAs, K walks
        const int tw = int((uint64_t(blockIdx.x) * 7919u + uint64_t(t)) % uint64_t((rows / w_rows) * k_tiles));
        if (lanes == 1) {
          ptx::tma_load_2d(&map_a, full_bar + stage, sa, ta * 128, int(blockIdx.x % 4) * 128);
          ptx::tma_load_2d(&map_w, full_bar + stage, sa + 128 * 128, (tw % k_tiles) * 128, (tw / k_tiles) * w_rows);
        } else {
          ptx::tma_load_2d(lane == 0 ? &map_a : &map_w, full_bar + stage, lane == 0 ? sa : sa + 128 * 128, lane == 0 ? ta * 128 : (tw % k_tiles) * 128,
                           lane == 0 ? int(blockIdx.x % 4) * 128 : (tw / k_tiles) * w_rows);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int t = 0; t < turns; ++t) {
      ptx::mbar_wait(full_bar + stage, phase);
      if (lane == 0) {
        acc += smem[stage * stage_bytes + (t & 1023)];
        ptx::mbar_arrive(empty_bar + stage);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && sink) sink[blockIdx.x] = acc;
  }
}

// mode "multicast": clusters of C CTAs; every CTA issues 1/C of each stage's bytes and multicasts them to all C CTAs, so each CTA
// RECEIVES a whole 32 KB stage while its TMA unit only ISSUES 32/C KB.  A stage is free when all C consumers have released it.
template <int C>
__global__ void __launch_bounds__(64, 1) feed_multicast_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int turns, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, C);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::cluster_sync_all();
  const int row_tiles = rows / 128, k_tiles = kK / 128, total_tiles = row_tiles * k_tiles;
  constexpr int kSliceRows = 256 / C;  // the stage is 256 rows × 128 bytes (two tiles); each CTA brings 256/C rows
  const int cluster_id = int(blockIdx.x) / C;
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < turns; ++t) {
        ptx::mbar_wait_cluster(empty_bar + stage, phase ^ 1);
        ptx::mbar_arrive_expect_tx(full_bar + stage, kStageBytes);
        const int tile = int((uint64_t(cluster_id) * 7919u + uint64_t(t) * 2) % uint64_t(total_tiles - 1));
        // rows [tile_row·128 + rank·kSliceRows, +kSliceRows) of a 256-row stage → same offset in every CTA of the cluster
        // (the tensor map's box is kSliceRows rows)
        ptx::tma_load_2d_multicast(&tmap, full_bar + stage, smem + stage * kStageBytes + rank * kSliceRows * 128, (tile % k_tiles) * 128,
                                   ((tile / k_tiles) * 128) % (rows - 256) + int(rank) * kSliceRows, uint16_t((1u << C) - 1u));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int t = 0; t < turns; ++t) {
      ptx::mbar_wait(full_bar + stage, phase);
      if (lane == 0) {
        acc += smem[stage * kStageBytes + (t & 1023)];
#pragma unroll
        for (int p = 0; p < C; ++p) ptx::mbar_arrive_cluster(empty_bar + stage, uint32_t(p));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && sink) sink[blockIdx.x] = acc;
  }
  ptx::cluster_sync_all();
}

template <int C>
float run_multicast(const CUtensorMap &map, int rows, int ctas, int turns, unsigned long long *sink) {
  cudaFuncSetAttribute(feed_multicast_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(ctas / C * C));
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = kStages * kStageBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, feed_multicast_kernel<C>, map, rows, 200, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, feed_multicast_kernel<C>, map, rows, turns, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) std::fprintf(stderr, "multicast<%d>: %s\n", C, cudaGetErrorString(e));
  return ms;
}

#define CK(x)                                                                \
  do {                                                                       \
    cudaError_t e_ = (x);                                                    \
    if (e_ != cudaSuccess) {                                                 \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));          \
      std::exit(1);                                                          \
    }                                                                        \
  } while (0)

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int rows = 16384;  // 32 MB matrix: L2-resident
  uint8_t *d = nullptr;
  CK(cudaMalloc(&d, size_t(rows) * kK));
  CK(cudaMemset(d, 1, size_t(rows) * kK));
  unsigned long long *sink = nullptr;
  CK(cudaMalloc(&sink, 1024 * 8));
  using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t dims[2] = {cuuint64_t(kK), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(kK)};
  cuuint32_t box[2] = {128u, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  if (reinterpret_cast<EncodeFn>(fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    std::fprintf(stderr, "cuTensorMapEncodeTiled failed\n");
    return 1;
  }
  CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char *names[3] = {"tma", "ldgsts", "mixed"};
  const int turns = 4000;  // 128 MB per CTA
  std::printf("{\"stage_bytes\": %d, \"stages\": %d, \"copy_warps\": %d, \"results\": [", kStageBytes, kStages, kCopyWarps);
  bool first = true;
  for (int ctas : {sms, sms / 2, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      feed_kernel<<<ctas, kThreads, kStages * kStageBytes>>>(map, d, rows, 200, mode, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      feed_kernel<<<ctas, kThreads, kStages * kStageBytes>>>(map, d, rows, turns, mode, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double bytes = double(ctas) * turns * kStageBytes;
      std::printf("%s{\"ctas\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"total_gbs\": %.1f, \"per_sm_gbs\": %.1f}", first ? "" : ", ", ctas, names[mode], ms,
                  bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / ctas);
      first = false;
    }
  }
  // TMA variants: box height, number of issuing lanes, stage size
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608));
  for (int box_rows : {32, 64, 128, 256}) {
    CUtensorMap mb;
    cuuint32_t boxb[2] = {128u, cuuint32_t(box_rows)};
    if (reinterpret_cast<EncodeFn>(fnp)(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, boxb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      continue;
    for (int stage_bytes : {32768, 49152, 65536}) {
      const int stages = 196608 / stage_bytes > 8 ? 8 : 196608 / stage_bytes;
      for (int issuers : {1, 2, 4, 8}) {
        if (stage_bytes % (box_rows * 128) != 0 || stage_bytes / (box_rows * 128) < issuers) continue;
        const int tn = turns * 32768 / stage_bytes;
        tma_rate_kernel<<<sms, 64, stages * stage_bytes>>>(mb, rows, 200, box_rows, issuers, stage_bytes, stages, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        tma_rate_kernel<<<sms, 64, stages * stage_bytes>>>(mb, rows, tn, box_rows, issuers, stage_bytes, stages, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf(", {\"ctas\": %d, \"mode\": \"tma\", \"box_rows\": %d, \"issuing_lanes\": %d, \"stage_bytes\": %d, \"stages\": %d, \"per_sm_gbs\": %.1f}", sms, box_rows,
                    issuers, stage_bytes, stages, double(sms) * tn * stage_bytes / (ms * 1e-3) / 1e9 / sms);
      }
    }
  }
  // the layer kernels' stage shape: activation box + weight box from two tensor maps
  CK(cudaFuncSetAttribute(two_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608));
  for (int w_rows : {64, 128, 256}) {
    CUtensorMap mw2;
    cuuint32_t boxw[2] = {128u, cuuint32_t(w_rows)};
    if (reinterpret_cast<EncodeFn>(fnp)(&mw2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, boxw, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      continue;
    const int stage_bytes = (128 + w_rows) * 128, stages = 196608 / stage_bytes > 8 ? 8 : 196608 / stage_bytes;
    for (int ctas : {128, 64}) {
      for (int lanes : {1, 2, 3}) {
        two_map_kernel<<<ctas, 96, stages * stage_bytes>>>(map, mw2, rows, 200, w_rows, lanes, stages, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        two_map_kernel<<<ctas, 96, stages * stage_bytes>>>(map, mw2, rows, turns, w_rows, lanes, stages, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf(", {\"ctas\": %d, \"mode\": \"two maps: 128-row box + %d-row box, issue mode %d (1 one lane, 2 two lanes, 3 two warps), %d stages\", \"per_sm_gbs\": %.1f}", ctas, w_rows, lanes,
                    stages, double(turns) * stage_bytes / (ms * 1e-3) / 1e9);
      }
    }
  }
  {  // no swizzle, 128-row boxes
    CUtensorMap mn;
    if (reinterpret_cast<EncodeFn>(fnp)(&mn, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mn, d, rows, 200, 0, sink, 128, 1);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mn, d, rows, turns, 0, sink, 128, 1);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::printf(", {\"ctas\": %d, \"mode\": \"tma no swizzle\", \"ms\": %.3f, \"per_sm_gbs\": %.1f}", sms, ms, double(sms) * turns * kStageBytes / (ms * 1e-3) / 1e9 / sms);
    }
  }
  {  // rows only 512 bytes apart (the limb planes of the input layer's certificate kernel): the same buffer viewed as [rows·4][512]
    CUtensorMap mp;
    cuuint64_t dims5[2] = {512u, cuuint64_t(rows) * 4};
    cuuint64_t strides5[1] = {512u};
    if (reinterpret_cast<EncodeFn>(fnp)(&mp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims5, strides5, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      for (int issuers : {1, 2}) {
        feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mp, d, rows * 4, 200, 0, sink, 128, issuers, 4);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mp, d, rows * 4, turns, 0, sink, 128, issuers, 4);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf(", {\"ctas\": %d, \"mode\": \"tma rows 512 bytes apart\", \"issuing_lanes\": %d, \"ms\": %.3f, \"per_sm_gbs\": %.1f}", sms, issuers, ms,
                    double(sms) * turns * kStageBytes / (ms * 1e-3) / 1e9 / sms);
      }
    }
  }
  {  // a wide 2-D box: 256 bytes of K per row is not expressible with 128B swizzle; instead a matrix viewed as [rows/2][4096]: rows twice as long
    CUtensorMap mw;
    cuuint64_t dims2[2] = {cuuint64_t(2 * kK), cuuint64_t(rows / 2)};
    cuuint64_t strides2[1] = {cuuint64_t(2 * kK)};
    if (reinterpret_cast<EncodeFn>(fnp)(&mw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims2, strides2, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mw, d, rows / 2, 200, 0, sink, 128, 1);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      feed_kernel<<<sms, kThreads, kStages * kStageBytes>>>(mw, d, rows / 2, turns, 0, sink, 128, 1);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::printf(", {\"ctas\": %d, \"mode\": \"tma rows 4096 bytes apart\", \"ms\": %.3f, \"per_sm_gbs\": %.1f}", sms, ms, double(sms) * turns * kStageBytes / (ms * 1e-3) / 1e9 / sms);
    }
  }
  // multicast: tensor maps with boxes of 256/C rows
  for (int c : {2, 4}) {
    CUtensorMap mc;
    cuuint32_t boxc[2] = {128u, cuuint32_t(256 / c)};
    if (reinterpret_cast<EncodeFn>(fnp)(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, boxc, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      continue;
    for (int ctas : {sms, sms / 2}) {
      const int used = c == 2 ? ctas / 2 * 2 : (ctas * 7 / 8) / 4 * 4;
      const float ms = c == 2 ? run_multicast<2>(mc, rows, used, turns, sink) : run_multicast<4>(mc, rows, used, turns, sink);
      const double bytes = double(used) * turns * kStageBytes;  // bytes RECEIVED by the CTAs
      std::printf(", {\"ctas\": %d, \"mode\": \"multicast%d\", \"ms\": %.3f, \"total_gbs_received\": %.1f, \"per_sm_gbs_received\": %.1f}", used, c, ms,
                  bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / used);
    }
  }
  std::printf("]}\n");
  return 0;
}
