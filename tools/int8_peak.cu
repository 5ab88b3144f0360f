// Calibration of the dense int8 tensor-core rate of THIS GPU (SURVEY.md §7 hard part 10): a tcgen05.mma kind::i8 loop with
// both operands resident in shared memory (no TMA, no epilogue, nothing but UTCIMMA and the commits that bound the number of
// instructions in flight), one CTA — or one CTA pair — per SM.  The figure it prints is the denominator of every
// "fraction of int8 peak" this repository reports (bench.py reads tools/int8_peak.json or runs this binary).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../fast-dnn_b200/csrc -o int8_peak int8_peak.cu
//   ./int8_peak [seconds of sustained run, default 2]
//
// Prints one JSON object: burst (≈ 2 ms launches, best of 10) and sustained (back-to-back launches for the given time)
// TOP/s for cta_group::1 (128×256×32 per instruction) and cta_group::2 (256×256×32 per instruction).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace fdnn;

constexpr int kN = 256;         // accumulator columns per MMA
constexpr int kStageK = 128;    // bytes of K per resident operand tile (one swizzle atom) = 4 MMAs of K = 32
constexpr int kGroup = 16;      // MMAs per commit
constexpr int kThreads = 128;

template <int kCtaGroup>
__global__ void __launch_bounds__(kThreads, 1) mma_loop_kernel(int groups, int random_fill, unsigned long long *sink, int n_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // A: 128 rows × 128 bytes; B: 256 (cta_group::1) or 128 (this CTA's half, cta_group::2) rows × 128 bytes; contents are
  // irrelevant to the rate (zeros), the layout is the K-major 128B-swizzled one the layer kernels use
  uint8_t *a_tile = smem;
  uint8_t *b_tile = smem + 128 * kStageK;
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (128 + kN) * kStageK / 16; i += kThreads) {
    // operand bytes: zeros, or pseudo-random bytes (what real activations and weights look like to the multipliers: the
    // power drawn, and with it the clock the chip sustains, depends on how much the operands toggle)
    uint32_t h = uint32_t(i) * 2654435761u + blockIdx.x * 40503u;
    uint4 v;
    h ^= h >> 15; h *= 2246822519u; v.x = h;
    h ^= h >> 13; h *= 3266489917u; v.y = h;
    h ^= h >> 16; h *= 668265263u; v.z = h;
    h ^= h >> 15; h *= 374761393u; v.w = h;
    reinterpret_cast<uint4 *>(smem)[i] = random_fill ? v : make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (kCtaGroup == 1)
      ptx::tmem_alloc<512>(&tmem_slot);
    else
      ptx::tmem_alloc_pair<512>(&tmem_slot);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill → async-proxy (tensor core) reads
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kCtaGroup == 2) ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const bool issuer = warp == 0 && (kCtaGroup == 1 || ptx::cluster_ctarank() == 0);
  if (issuer) {
    const uint32_t idesc = kCtaGroup == 1 ? ptx::idesc_i8_u8s8(uint32_t(n_cols)) : ptx::idesc_i8_u8s8_pair(uint32_t(n_cols));
    const uint64_t da = ptx::smem_desc_k_sw128(ptx::smem_u32(a_tile)), db = ptx::smem_desc_k_sw128(ptx::smem_u32(b_tile));
    uint32_t phase[2] = {0, 0};
    for (int g = 0; g < groups; ++g) {
      const int b = g & 1;
      if (g >= 2) {  // at most two groups of kGroup instructions in flight
        ptx::mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1;
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kGroup; ++i) {
          const uint32_t d = tmem_base + uint32_t((i & 1) * kN);  // alternate the two accumulator halves
          const uint64_t koff = uint64_t((i & 3) * 2);
          if (kCtaGroup == 1)
            ptx::mma_i8_ss(d, da + koff, db + koff, idesc, 1u);
          else
            ptx::mma_i8_ss_pair(d, da + koff, db + koff, idesc, 1u);
        }
        if (kCtaGroup == 1)
          ptx::mma_commit(&bars[b]);
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           ptx::smem_u32(&bars[b])),
                       "h"(uint16_t(1))
                       : "memory");
      }
      __syncwarp();
    }
    for (int b = 0; b < 2 && b < groups; ++b) ptx::mbar_wait(&bars[(groups - 1 - b) & 1], phase[(groups - 1 - b) & 1]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kCtaGroup == 2) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    if (kCtaGroup == 1)
      ptx::tmem_dealloc<512>(tmem_base);
    else
      ptx::tmem_dealloc_pair<512>(tmem_base);
  }
  if (threadIdx.x == 0 && sink != nullptr && groups < 0) *sink = tmem_base;
}

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                        \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

template <int kCtaGroup>
cudaError_t launch(int ctas, int groups, int random_fill, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(ctas));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (128 + kN) * kStageK;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kCtaGroup > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, mma_loop_kernel<kCtaGroup>, groups, random_fill, static_cast<unsigned long long *>(nullptr), int(kN));
}

template <int kCtaGroup>
void measure(int sms, double seconds, int random_fill, double *burst, double *sustained, double *sustained_first, double *sustained_last) {
  CK(cudaFuncSetAttribute(mma_loop_kernel<kCtaGroup>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + kN) * kStageK));
  const int ctas = sms / kCtaGroup * kCtaGroup;
  // ops per instruction and issuing CTA: 2 · M · N · K with M = 128 per CTA of the group
  const double ops_per_group = 2.0 * (128.0 * kCtaGroup) * kN * 32.0 * kGroup;
  const double issuers = double(ctas) / kCtaGroup;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int groups = 2000;  // 32 000 MMAs per issuer ≈ 2 ms
  for (int i = 0; i < 3; ++i) CK(launch<kCtaGroup>(ctas, groups, random_fill, nullptr));
  CK(cudaDeviceSynchronize());
  double best = 0;
  for (int i = 0; i < 10; ++i) {
    CK(cudaEventRecord(e0));
    CK(launch<kCtaGroup>(ctas, groups, random_fill, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::max(best, ops_per_group * groups * issuers / (ms * 1e-3) / 1e12);
  }
  *burst = best;
  // sustained: back-to-back launches for `seconds`, rate of the whole interval and of its first / last tenth
  std::vector<cudaEvent_t> ev;
  const auto t0 = std::chrono::steady_clock::now();
  cudaEvent_t first;
  CK(cudaEventCreate(&first));
  CK(cudaEventRecord(first));
  ev.push_back(first);
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    for (int i = 0; i < 8; ++i) CK(launch<kCtaGroup>(ctas, groups, random_fill, nullptr));
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e));
    ev.push_back(e);
    CK(cudaEventSynchronize(e));
  }
  auto rate = [&](size_t a, size_t b) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev[a], ev[b]));
    return ops_per_group * groups * issuers * 8.0 * double(b - a) / (ms * 1e-3) / 1e12;
  };
  const size_t n = ev.size() - 1, tenth = std::max<size_t>(1, n / 10);
  *sustained = rate(0, n);
  *sustained_first = rate(0, tenth);
  *sustained_last = rate(n - tenth, n);
  for (auto e : ev) cudaEventDestroy(e);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

int main(int argc, char **argv) {
  const double seconds = argc > 1 ? std::atof(argv[1]) : 2.0;
  int dev = 0, sms = 0, clock_khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  double r[4][4];
  measure<1>(sms, seconds, 1, &r[0][0], &r[0][1], &r[0][2], &r[0][3]);
  measure<2>(sms, seconds, 1, &r[1][0], &r[1][1], &r[1][2], &r[1][3]);
  measure<1>(sms, seconds / 2, 0, &r[2][0], &r[2][1], &r[2][2], &r[2][3]);
  measure<2>(sms, seconds / 2, 0, &r[3][0], &r[3][1], &r[3][2], &r[3][3]);
  // what the instruction rate would give at the maximum SM clock: 128·256·32 MACs per 128 cycles and SM
  const double nominal = 2.0 * 8192.0 * sms * (clock_khz * 1e3) / 1e12;
  const char *names[4] = {"cta_group_1", "cta_group_2", "cta_group_1_zero_operands", "cta_group_2_zero_operands"};
  std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f, \"int8_tops_at_max_clock_8192_mac_per_clk_sm\": %.1f", prop.name, sms,
              clock_khz / 1e3, nominal);
  for (int i = 0; i < 4; ++i)
    std::printf(", \"%s\": {\"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"sustained_first_tenth\": %.1f, \"sustained_last_tenth\": %.1f}", names[i],
                r[i][0], r[i][1], r[i][2], r[i][3]);
  // instruction rate at narrower tiles (cta_group::1, M = 128): is a layer with 64- or 128-wide tiles bound by how fast one thread
  // can issue tcgen05.mma rather than by the tensor pipe?
  {
    CK(cudaFuncSetAttribute(mma_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + kN) * kStageK));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::printf(", \"narrow_tiles\": [");
    bool first = true;
    for (int n : {32, 64, 128, 256}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(unsigned(sms));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = (128 + kN) * kStageK;
      const int groups = 4000;
      CK(cudaLaunchKernelEx(&cfg, mma_loop_kernel<1>, 200, 1, static_cast<unsigned long long *>(nullptr), n));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      CK(cudaLaunchKernelEx(&cfg, mma_loop_kernel<1>, groups, 1, static_cast<unsigned long long *>(nullptr), n));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double mmas = double(groups) * kGroup;
      std::printf("%s{\"n\": %d, \"ns_per_mma\": %.1f, \"tops\": %.1f}", first ? "" : ", ", n, ms * 1e6 / mmas, 2.0 * 128 * n * 32 * mmas * sms / (ms * 1e-3) / 1e12);
      first = false;
    }
    std::printf("]");
  }
  std::printf(", \"sustained_seconds\": %.1f, \"how\": \"tcgen05.mma kind::i8 (u8 x s8 -> s32), operands resident in shared memory "
              "(pseudo-random bytes; zeros for the *_zero_operands entries), %d MMAs per commit, 2 commits in flight, one CTA (pair) per SM; "
              "ops = 2*M*N*K\"}\n",
              seconds, kGroup);
  return 0;
}
